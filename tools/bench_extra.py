#!/usr/bin/env python3
"""Secondary measurements on one B200 (BASELINE.json configs[1], [3]; FRI commit phase):
  * MerkleTree::new Poseidon sweep, 2^16..2^24 leaves x 135 elements, cap_height 4 and 0
    (plonky2/benches/merkle.rs shape)
  * recursion-shaped commits (bench_recursion: n = 2^12..2^14, columns 85/143/20/16, rate 3, cap 4)
  * fri_committed_trees on a 2^(d+3) domain, arities from ConstantArityBits(4,5), plus PoW grinding
Each number is the best of `reps` runs, wall clock around the C-ABI call with device-resident
inputs (cudaDeviceSynchronize on both sides), i.e. latency as a caller sees it.
    python tools/bench_extra.py > gpurun_out/extra.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qp_plonky2_b200 as qp  # noqa: E402


def rnd(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 2**62, shape, dtype=torch.int64, device="cuda", generator=g)


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        if hasattr(r, "free"):
            r.free()
    return min(ts)


def main():
    ctx = qp.Context(0, max_lde_log=24)
    out = {"merkle_sweep": [], "recursion_commits": [], "fri_commit": []}
    for lg in (16, 18, 20, 22, 24):
        if lg == 24:
            leaves = None
        leaves = rnd((1 << lg, 135), lg)
        for cap_h in (4, 0):
            ms = best(lambda: qp.MerkleTree(ctx, leaves, cap_h), reps=3 if lg >= 22 else 5)
            perms = (1 << lg) * 17 + (1 << lg) - (1 << cap_h)
            out["merkle_sweep"].append({"leaves_log": lg, "leaf_len": 135, "cap_height": cap_h, "ms": ms,
                                        "perms_per_s": perms / ms * 1e3,
                                        "note": "includes the device copy of the caller's leaves"})
        del leaves
        torch.cuda.empty_cache()
    for lg in (12, 13, 14):
        for cols in (85, 143, 20, 16):
            v = rnd((cols, 1 << lg), 100 + lg)
            ms = best(lambda: qp.PolynomialBatch.from_values(ctx, v, 3, False, 4), reps=10)
            out["recursion_commits"].append({"rows_log": lg, "cols": cols, "rate_bits": 3, "cap_height": 4, "ms": ms})
    for d in (12, 13, 14, 20):
        n = 1 << (d + 3)
        co = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
        co[: 1 << d] = rnd((1 << d, 2), 7)
        # values = coset FFT of the coefficients (natural order), computed with the library itself
        planes = co.t().contiguous().cpu().numpy().view(np.uint64)
        vals = ctx.coset_fft(planes, shift=14293326489335486720)
        va = torch.from_numpy(np.ascontiguousarray(vals.T).view(np.int64)).cuda()
        ar = qp.fri_reduction_arity_bits(d, 3, 4)

        def run():
            ch = qp.Challenger()
            r = qp.fri_committed_trees(ctx, co, va, ch, 3, 4, ar)
            run.ch = ch
            return r

        ms = best(run, reps=5)
        ch = run.ch
        t0 = time.perf_counter()
        w = qp.fri_proof_of_work(ctx, ch, 16)
        pow_ms = (time.perf_counter() - t0) * 1e3
        out["fri_commit"].append({"degree_bits": d, "lde_bits": d + 3, "arity_bits": ar, "commit_phase_ms": ms,
                                  "pow16_ms": pow_ms, "pow_witness": w})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
