cd $GRAFT_REPO_ROOT
for v in 0 1 2 3 8 9 10 11; do echo "== pow7 variant $v"; timeout 60 tools/microbench/bin/phase_pv$v.bin 4 | grep -E "^S |^P |^perm"; done 2>&1 | tee gpurun_out/pow7_variants_r02k.txt
