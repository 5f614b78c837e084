#!/usr/bin/env python3
"""prove() over 1 / 2 / 4 / 8 GPUs of one box from ONE process (qp_mprove) against the single-context qp_prove:
synthetic 143-wire circuits (NoopGate, ConstantGate, PublicInputGate, ArithmeticGate rows with copy constraints)
under standard_recursion_config, witness in host memory.  Prints one JSON line per size with the proof latency
per device count, the TimingTree scopes of the best run and whether every proof had the same bytes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qp_plonky2_b200 as qp  # noqa: E402
from qp_plonky2_b200 import plonk, prover  # noqa: E402
from synth_circuit import SynthCircuit  # noqa: E402

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [16, 18]
    n_dev = torch.cuda.device_count()
    for db in sizes:
        t0 = time.time()
        sc = SynthCircuit(db, seed=900 + db)
        c = sc.common
        gen_s = time.time() - t0
        # the witness in PINNED host memory (what a prover that generates it for the GPU would use; pageable memory
        # makes the host staging copy, ~10 GB/s on these boxes, the bottleneck of every upload): QP_BENCH_PAGEABLE=1
        wires = sc.wires
        if not os.environ.get("QP_BENCH_PAGEABLE"):
            pinned = torch.from_numpy(np.ascontiguousarray(sc.wires).view(np.int64)).pin_memory()
            wires = pinned.numpy().view(np.uint64)
        REC_PLACEHOLDER = None
        rec = {"witness_memory": "pageable" if os.environ.get("QP_BENCH_PAGEABLE") else "pinned", "degree_bits": db, "num_wires": c.num_wires, "witness_gen_s": round(gen_s, 1), "ms": {}, "scopes_ms": {}}
        ctx = qp.Context(0, max_lde_log=db + c.rate_bits)
        circ = plonk.Circuit(ctx, c, sc.sigmas)
        pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
        ref = None
        best = 1e30
        for it in range(3):
            t0 = time.perf_counter()
            ref = prover.prove(pd, wires, sc.public_inputs)
            best = min(best, (time.perf_counter() - t0) * 1e3)
        rec["ms"]["single_ctx"] = round(best, 2)
        cs_cap = pd.constants_sigmas_commitment.merkle_tree.cap
        digest = pd.circuit_digest
        pd.constants_sigmas_commitment.free()
        circ.free()
        ctx.close()
        same = True
        only = [int(x) for x in os.environ.get("QP_MPROVE_DEVICES", "").split(",") if x]
        for D in [x for x in (1, 2, 4, 8) if x <= n_dev and (not only or x in only)]:
            m = qp.MultiContext(list(range(D)), max_lde_log=db + c.rate_bits)
            mpd = prover.MultiProverData(m, c, sc.sigmas, sc.constants_sigmas())
            best, timing = 1e30, {}
            for it in range(3):
                t = {}
                t0 = time.perf_counter()
                got = prover.mprove(mpd, wires, sc.public_inputs, t)
                dt = (time.perf_counter() - t0) * 1e3
                if dt < best:
                    best, timing = dt, t
                same = same and got == ref
            rec["ms"]["devices_%d" % D] = round(best, 2)
            rec["scopes_ms"]["devices_%d" % D] = {k: round(v, 2) for k, v in timing.items()}
            mpd.constants_sigmas_commitment.free()
            for x in mpd.circuits:
                x.free()
            m.close()
        rec["all_proofs_identical"] = same
        rec["proof_bytes"] = len(ref)
        if db <= 18:
            import verifier

            rec["verifier_accepts"] = verifier.verify(ref, c, prover.FriConfig(c.rate_bits, c.cap_height), cs_cap,
                                                      digest) is None
        print(json.dumps(rec), flush=True)
