cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
( QP_BENCH_PINNED=1 timeout 300 python tools/bench_multi.py 20 ; timeout 300 python tools/bench_multi.py 20 ) 2>&1 | grep -v "^\[qp_" | tee gpurun_out/r02m_multi_device_commit_n8.txt
timeout 600 python tools/bench_mprove.py 18 20 2>&1 | tail -4 | tee gpurun_out/r02m_mprove_n8.txt
