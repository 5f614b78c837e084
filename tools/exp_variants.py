#!/usr/bin/env python3
"""Kernel experiments: build the library with different -D flags and time the commit kernels.

    python tools/exp_variants.py build  NAME "-DFOO=1 -DBAR=2" ...   (here, no GPU needed)
    python tools/exp_variants.py run [rows_log]                        (on the GPU box)
Variants live in gpurun_out/variants/NAME.so (scratch, not committed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "variants")  # *.so is git-ignored but travels with gpurun


def build(name, flags):
    os.makedirs(VDIR, exist_ok=True)
    out = os.path.join(VDIR, name + ".so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
           "-fPIC", "-shared", "-o", out] + flags.split() + [
        os.path.join(ROOT, "qp-plonky2_b200/csrc/qp_plonky2.cu"), os.path.join(ROOT, "qp-plonky2_b200/host/transcript.cpp"),
        os.path.join(ROOT, "qp-plonky2_b200/host/plonk_host.cpp"), os.path.join(ROOT, "qp-plonky2_b200/host/prover.cpp")]
    subprocess.check_call(cmd)
    print("built", out)


CHILD = r"""
import sys, json, torch
sys.path.insert(0, %r)
import qp_plonky2_b200 as qp
rows_log = %d
ctx = qp.Context(0, max_lde_log=rows_log + 3)
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.randint(0, 2**62, (135, 1 << rows_log), dtype=torch.int64, device="cuda", generator=g)
best = None
for it in range(4):
    b = qp.PolynomialBatch.from_values(ctx, d, 3, False, 4)
    k = dict(b.kernel_ms); cap = b.merkle_tree.cap[0].tolist(); b.free()
    if it and (best is None or k["leaf_hash"] < best["leaf_hash"]): best = k
print(json.dumps({"kernel_ms": best, "cap0": cap}))
"""


def run(rows_log):
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        env = dict(os.environ, QP_PLONKY2_LIB=os.path.join(VDIR, f))
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, rows_log)], env=env, capture_output=True, text=True)
        print(f, r.stdout.strip() or r.stderr[-400:])


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for i in range(2, len(sys.argv), 2):
            build(sys.argv[i], sys.argv[i + 1])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 18)
