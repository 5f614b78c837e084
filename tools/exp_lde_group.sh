cd $GRAFT_REPO_ROOT
for g in 0 1 2 4 8; do
  QP_LDE_GROUP=$g python - <<PY
import json, torch, sys
sys.path.insert(0, ".")
import bench, qp_plonky2_b200 as qp
ctx = qp.Context(0, max_lde_log=23)
d = bench.synth_columns_torch(0, 135, 1 << 20, "cuda")
best = None
for it in range(4):
    b = qp.PolynomialBatch.from_values(ctx, d, 3, False, 4)
    k = dict(b.kernel_ms); cap = b.merkle_tree.cap[0].tolist(); b.free()
    if it and (best is None or k["lde"] < best["lde"]): best = k
print("QP_LDE_GROUP=$g", json.dumps({"lde_ms": best["lde"], "intt": best["intt"], "cap0": cap[:2]}))
PY
done
python tools/bench_stage.py 2>&1 | tail -3
# DRAM traffic of the LDE passes at group 1 (4 columns only, enough launches to see the per-launch bytes)
QP_LDE_GROUP=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'strided_pass_kernel<8>|final_pass_kernel<12>' -s 20 -c 8 --csv --log-file gpurun_out/lde_group1_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-prove > /dev/null 2>&1
tail -9 gpurun_out/lde_group1_traffic.csv | cut -d, -f5,12-
