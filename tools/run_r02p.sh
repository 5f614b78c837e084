# Round-2 final measurement on one B200 (gpurun): bench line, launch list of the same command restricted to the
# library's kernels, one `ncu --set full` capture of the commit kernels and one of the lookup / quotient kernels.
cd $GRAFT_REPO_ROOT
python bench.py --steps 10 --warmup 3 > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; tail -c 300 gpurun_out/r02p_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench_n1.json')); print({k:d[k] for k in ('value','e2e','e2e_pageable','kernel_ms','cap_equal_cpu','gpu_launches')}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores']); print(d['roofline']); p=d.get('prove') or {}; print([(r['degree_bits'], round(r['ms'],2), r.get('bytes_equal_cpu')) for r in p.get('runs',[])], [(r['degree_bits'], round(r['ms'],2), r.get('bytes_equal_cpu'), r.get('scopes_ms')) for r in p.get('runs_with_lookups',[])], p.get('error'))"
MINE='regex:ntt::|merkle::|quotient::|openings::|fri::|_kernel'
# launch list of the same command (after it exited 0 without ncu): the library's kernels only, so that the step's
# launches are not crowded out by torch's witness-generation kernels
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "$MINE" -c 4000 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-prove > gpurun_out/r02p_ncu_launch.log 2>&1
tail -2 gpurun_out/r02p_ncu_launch.log; wc -l gpurun_out/r02p_launches.csv
# full capture of the commit kernels, one launch of each kind
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:leaf_hash_kernel|strided_pass_kernel|final_pass_kernel|tree_level_kernel' -c 12 -o gpurun_out/r02p_prof_commit python bench.py --steps 1 --warmup 0 --no-cpu --no-prove > gpurun_out/r02p_ncu_full.log 2>&1
# the plonk-layer kernels of a 2^14-row proof with lookup tables
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:lookup_|quotient_kernel|poseidon_gate_kernel|combine_kernel|perm_chunks' -c 8 -o gpurun_out/r02p_prof_plonk python tools/bench_prove.py --degrees 14 --cpu --reps 1 --recursion --lookups > gpurun_out/r02p_ncu_plonk.log 2>&1
# the reports stay on the box (gpurun_out/ travels back only below 64 MiB): their summaries and raw pages come home
for r in commit plonk; do
  python tools/ncu_summary.py gpurun_out/r02p_prof_$r.ncu-rep gpurun_out/r02p_ncu_${r}_kernels.md
  ncu -i gpurun_out/r02p_prof_$r.ncu-rep --page raw --csv > gpurun_out/r02p_ncu_${r}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02p_prof_$r.ncu-rep
done
python tools/ncu_summary.py --launches gpurun_out/r02p_launches.csv gpurun_out/r02p_launches.md
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
