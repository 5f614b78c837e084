cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for f in tools/microbench/bin/leaf_*.bin; do printf "%-28s " $(basename $f .bin); timeout 120 $f 21 135 4 2>&1 | tail -1; done | tee gpurun_out/leaf_variants_r02a.txt
