#!/usr/bin/env python3
"""End-to-end commit from 135 separate PAGEABLE host columns (qp_batch_from_values_cols) under
different staging settings: QP_STAGE_MODE (0 pinned ring, 1 driver-staged) x QP_STAGE_THREADS."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, time, json, numpy as np, torch
sys.path.insert(0, %r)
import bench, qp_plonky2_b200 as qp
ctx = qp.Context(0, max_lde_log=23)
cols = list(bench.synth_columns_numpy(0, 135, 1 << 20))
cols = [np.array(c, copy=True) for c in cols]
best = 1e9
for it in range(5):
    t0 = time.perf_counter()
    b = qp.PolynomialBatch.from_values_cols(ctx, cols, 3, False, 4)
    cap = b.merkle_tree.cap
    dt = (time.perf_counter() - t0) * 1e3
    b.free()
    if it: best = min(best, dt)
print(json.dumps({"e2e_pageable_ms": best, "cap0": [int(x) for x in cap[0]]}))
"""

if __name__ == "__main__":
    for mode, thr in [(0, 1), (0, 2), (0, 4), (0, 8), (0, 12), (1, 1)]:
        env = dict(os.environ, QP_STAGE_MODE=str(mode), QP_STAGE_THREADS=str(thr))
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
        print("mode=%d threads=%-2d" % (mode, thr), r.stdout.strip() or r.stderr[-300:], flush=True)
