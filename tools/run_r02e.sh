cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02e_n1.json 2> gpurun_out/bench_r02e_n1.err; tail -c 300 gpurun_out/bench_r02e_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02e_n1.json')); print({k:d[k] for k in ('value','e2e','e2e_pageable','kernel_ms','cap_equal_cpu')}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print(d['roofline']); print([ (r['degree_bits'], r.get('gpu_ms')) for r in d['prove']['runs']] if d.get('prove') and 'runs' in d['prove'] else d.get('prove'))"
# launch list of the same command (after it exited 0 without ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02e.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-prove > gpurun_out/ncu_launch.log 2>&1
# full capture of the commit kernels: leaf hash + the four NTT pass kernels, one launch each
ncu --set full --clock-control none --import-source on -k regex:'leaf_hash_kernel|strided_pass_kernel|final_pass_kernel|tree_level_kernel' -c 12 -o gpurun_out/prof_r02e_commit python bench.py --steps 1 --warmup 0 --no-cpu --no-prove > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_r02e_commit.ncu-rep
