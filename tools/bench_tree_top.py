#!/usr/bin/env python3
"""Where does the 16-lanes-per-permutation tree kernel (merkle::tree_top_kernel) beat the per-thread
one?  Times MerkleTree::new on short leaves (so the levels dominate) for several thresholds; each
threshold runs in its own process because the library reads QP_TREE_TOP_NODES once.
    python tools/bench_tree_top.py > gpurun_out/tree_top.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, time, json
sys.path.insert(0, %r)
import numpy as np, torch
import qp_plonky2_b200 as qp
ctx = qp.Context(0, max_lde_log=12)
out = {}
for lg in (7, 11, 15, 17, 19):
    leaves = torch.randint(0, 2**62, (1 << lg, 8), dtype=torch.int64, device="cuda")
    best = None
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t = qp.MerkleTree(ctx, leaves, 4)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e6
        if rep and (best is None or dt < best):
            best = dt
        cap = t.cap.copy()
        t.free()
    out[lg] = (round(best, 1), int(cap[0][0]))
print(json.dumps(out))
""" % ROOT


def main():
    for thr in (0, 16, 256, 2048, 8192, 32768):
        env = dict(os.environ, QP_TREE_TOP_NODES=str(thr))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print("threshold %6d:" % thr, r.stdout.strip() or r.stderr[-400:])


if __name__ == "__main__":
    main()
