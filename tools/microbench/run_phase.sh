cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in base expl comp; do echo "== phase_$v"; timeout 120 tools/microbench/bin/phase_$v.bin 4; done 2>&1 | tee gpurun_out/phase_overlap_r02i.txt
for v in base expl; do printf "%-12s " leaf_$v; timeout 120 tools/microbench/bin/leaf_$v.bin 21 135 4 2>&1 | tail -1; done | tee -a gpurun_out/phase_overlap_r02i.txt
