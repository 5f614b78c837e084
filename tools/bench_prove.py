#!/usr/bin/env python3
"""prove() latency on one B200 for synthetic circuits of bench_recursion's shapes (BASELINE.json
configs[1]: degree 2^12..2^14, 143 wires / 80 routed, standard_recursion_config) and larger, with
the reference's TimingTree scope names; optionally the oracle's CPU prove() next to it.
    python tools/bench_prove.py [--degrees 12 13 14] [--cpu 12] > gpurun_out/prove.json
The inputs come from tests/synth_circuit.py, which needs nothing from oracle/; the oracle is imported only
for the CPU leg (--cpu), the counterpart of bench.py's cpu_baseline."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qp_plonky2_b200 as qp  # noqa: E402
from qp_plonky2_b200 import plonk, prover  # noqa: E402


def measure(degrees, cpu_degrees=(), reps=5, device=0, verbose=True, poseidon=True, recursion=False, lookups=False):
    """-> list of per-degree records (ms = best of reps - 1 timed runs after one warm-up)."""
    import torch
    from synth_circuit import SynthCircuit

    class A:
        pass
    a = A()
    a.degrees, a.cpu, a.reps = list(degrees), list(cpu_degrees), reps
    ctx = qp.Context(device, max_lde_log=max(a.degrees) + 3)
    out = {"prove": []}
    for lg in a.degrees:
        sc = SynthCircuit(lg, seed=lg, poseidon=poseidon, extra_gates=recursion, recursion_gates=recursion, lookups=lookups)
        c = sc.common
        circ = plonk.Circuit(ctx, c, sc.sigmas)
        pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
        wires_dev = torch.from_numpy(sc.wires.view(np.int64)).cuda()
        best, best_t, nbytes = None, None, 0
        for rep in range(a.reps):
            t = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proof = prover.prove(pd, wires_dev, sc.public_inputs, t)
            ms = (time.perf_counter() - t0) * 1e3
            nbytes = len(proof)
            if rep and (best is None or ms < best):
                best, best_t = ms, t
        # end to end from the witness as the reference holds it: one pageable host vector per wire (qp_prove_cols)
        cols = [np.array(sc.wires[w], copy=True) for w in range(c.num_wires)]
        best_host = None
        for rep in range(a.reps):
            t0 = time.perf_counter()
            proof_h = prover.prove(pd, cols, sc.public_inputs)
            ms = (time.perf_counter() - t0) * 1e3
            if rep and (best_host is None or ms < best_host):
                best_host = ms
        rec = {"degree_bits": lg, "num_wires": c.num_wires, "gates": [g.id().split(" ")[0].split("(")[0] for g in c.gates],
               "ms": best, "proof_bytes": nbytes, "scopes_ms": best_t,
               "witness": "device-resident",
               "ms_from_host_witness": best_host, "host_witness": "MatrixWitness.wire_values: %d pageable vectors" % c.num_wires,
               "host_witness_bytes_equal": proof_h == proof}
        if lookups:
            rec["lookup_tables"] = [len(t) for t in c.luts]
        if lg in a.cpu:
            import oracle
            from oracle import prover as oprover
            o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
            t0 = time.perf_counter()
            want, _ = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                                    degree_bits=lg, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                                    num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                                    num_partial_products=c.num_partial_products)
            rec["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
            rec["cpu_threads"] = oracle.lib().orc_num_threads()
            rec["bytes_equal_cpu"] = want == proof
        out["prove"].append(rec)
        if verbose:
            print(json.dumps(rec), file=sys.stderr)
        circ.free()
    ctx.close()
    return out["prove"]


def measure_factorial(cpu=True, reps=5, device=0):
    """BASELINE.json configs[0]: the `factorial` example (plonky2/examples/factorial.rs) as a real circuit
    (tests/factorial_circuit.py), standard_recursion_config -- proof latency on the device from the host witness, the
    oracle's CPU prove() beside it, equal bytes."""
    from factorial_circuit import factorial_circuit

    sc = factorial_circuit()
    c = sc.common
    ctx = qp.Context(device, max_lde_log=c.degree_bits + 3)
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
    best, proof = None, None
    for rep in range(reps):
        t0 = time.perf_counter()
        proof = prover.prove(pd, sc.wires, sc.public_inputs)
        ms = (time.perf_counter() - t0) * 1e3
        if rep and (best is None or ms < best):
            best = ms
    rec = {"what": "factorial example: 99 chained multiplications, in-circuit public-input hash; host witness",
           "degree_bits": c.degree_bits, "public_inputs": [int(x) for x in sc.public_inputs], "ms": best,
           "proof_bytes": len(proof)}
    if cpu:
        import oracle
        from oracle import prover as oprover
        o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
        t0 = time.perf_counter()
        want, _ = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                                degree_bits=c.degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                                num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                                num_partial_products=c.num_partial_products)
        rec["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        rec["cpu_threads"] = oracle.lib().orc_num_threads()
        rec["bytes_equal_cpu"] = want == proof
    circ.free()
    ctx.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", type=int, nargs="*", default=[12, 13, 14, 16])
    ap.add_argument("--cpu", type=int, nargs="*", default=[12])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-poseidon", action="store_true")
    ap.add_argument("--recursion", action="store_true",
                    help="all 14 gate types of a recursive verifier circuit (four selector groups)")
    ap.add_argument("--lookups", action="store_true", help="two lookup tables with their LookupGate / LookupTableGate rows")
    a = ap.parse_args()
    print(json.dumps({"prove": measure(a.degrees, a.cpu, a.reps, poseidon=not a.no_poseidon, recursion=a.recursion,
                                       lookups=a.lookups)}))


if __name__ == "__main__":
    main()
