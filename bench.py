#!/usr/bin/env python3
"""Headline benchmark: PolynomialBatch::from_values commit (iNTT + coset LDE + Poseidon Merkle
tree + cap) of 2^20 rows x 135 columns at rate_bits = 3, cap_height = 4 (BASELINE.json
configs[2]) on N B200s, beside the CPU restatement of the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows-log L]

One "step" = one whole commit of one synthetic witness matrix.  `value` = ms per commit with
the witness already resident in HBM (CUDA events, max over ranks); `e2e` = ms per commit through
the C ABI with HOST (pinned) buffers, the host->device copy of the witness and the device->host
read of the cap inside the timed region.  N > 1: coset-sharded strong scaling (dist.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_LOG, COLS, RATE_BITS, CAP_HEIGHT = 20, 135, 3, 4
METRIC = "commit_ms_2^20x135_rate3"


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.kill()
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 3 + k and s[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference path (oracle, kind "port")
# ---------------------------------------------------------------------------------------------
def cpu_commit_ms(rows_log, sample_rows_log, repeats=1):
    """Time oracle PolynomialBatch::from_values on a bounded sample (2^sample_rows_log rows x 135
    columns, all host threads) and scale by the row ratio to the full workload."""
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm uses every host core
    os.environ["OMP_NUM_THREADS"] = os.environ.get("QP_ORACLE_THREADS", str(os.cpu_count()))
    import oracle

    vals = oracle.rand_felts((COLS, 1 << sample_rows_log), 42)
    best, scopes = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        b = oracle.PolynomialBatch.from_values(vals, RATE_BITS, CAP_HEIGHT)
        dt = (time.perf_counter() - t0) * 1e3
        if best is None or dt < best:
            best, scopes = dt, b.scope_ms
        del b
    scale = 1 << (rows_log - sample_rows_log)
    return best * scale, best, scopes, oracle.lib().orc_num_threads()


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    sample = min(args.rows_log, 16)
    times = []
    for _ in range(args.warmup):
        cpu_commit_ms(args.rows_log, sample)
    for _ in range(args.steps):
        full_ms, _, scopes, cores = cpu_commit_ms(args.rows_log, sample)
        times.append(full_ms)
    ms = sum(times) / len(times)
    desc = "oracle (C/OpenMP restatement of the reference's rayon path) on 2^%d rows x %d cols, x%d" % (
        sample, COLS, 1 << (args.rows_log - sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": desc,
                         "scopes_ms_sample": scopes},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "PolynomialBatch::from_values 2^%d rows x %d cols, rate_bits=%d, cap_height=%d, "
                        "blinding=false (LDE + Poseidon Merkle commit)" % (args.rows_log, COLS, RATE_BITS, CAP_HEIGHT),
            "parallelism": "single GPU" if world == 1 else "coset-sharded x%d (column-sharded iNTT, "
                           "NCCL all-gather of coefficients and cap)" % world,
            "l2": "inputs (%.2f GB) and LDE (%.2f GB) exceed the 126 MB L2" % (
                COLS * 8 * 2 ** args.rows_log / 1e9, COLS * 8 * 2 ** (args.rows_log + RATE_BITS) / 1e9)}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries exactly ONE JSON line: everything native libraries print there (NCCL's version
    # banner, ...) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch

    import qp_plonky2_b200 as qp
    import qp_plonky2_b200.dist as qd

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one stream for torch (NCCL, tensor ops) and the library, so their work is ordered
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = qp.Context(local, max_lde_log=args.rows_log + RATE_BITS, stream=stream.cuda_stream)
    n = 1 << args.rows_log
    c_lo, c_hi = (0, COLS) if world == 1 else qd.column_shard(COLS, world, rank)

    # synthetic witness: uniform canonical Goldilocks elements, seed 42 (same on every run)
    # (per column, so that a rank generates only its own column shard and every N sees the same matrix)
    gen = torch.Generator(device=dev)
    d_vals = torch.empty((c_hi - c_lo, n), dtype=torch.int64, device=dev)
    for c in range(c_lo, c_hi):
        gen.manual_seed(42 + c)
        col = torch.randint(0, 2**63 - 1, (n,), dtype=torch.int64, device=dev, generator=gen)
        d_vals[c - c_lo] = col * 2 + torch.randint(0, 2, (n,), dtype=torch.int64, device=dev, generator=gen)  # 64 random bits
    del col
    h_vals = torch.empty(d_vals.shape, dtype=torch.int64).pin_memory()
    h_vals.copy_(d_vals)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        def step(src):
            b = qp.PolynomialBatch.from_values(ctx, src, RATE_BITS, False, CAP_HEIGHT)
            cap = b.merkle_tree.cap
            return b, cap
    else:
        pc = qd.padded_cols(COLS, world)

        def ifft_fn(v, rows):
            out = torch.zeros((rows, n), dtype=torch.int64, device=dev)
            ctx.ifft_columns(v, out_device=out[: v.shape[0]])
            return out

        def begin_fn(first, count):
            return qp.PolynomialBatch.begin(ctx, COLS, args.rows_log, RATE_BITS, False, CAP_HEIGHT,
                                            block_first=first, block_count=count)

        def put_fn(batch, rows, c0):
            batch.put_coeffs(rows, c0)

        def end_fn(batch):
            batch.end()
            return batch, torch.from_numpy(batch.merkle_tree.cap.view(np.int64)).to(dev)

        def step(src):
            b, cap = qd.sharded_commit(src, COLS, args.rows_log, RATE_BITS, CAP_HEIGHT, rank=rank, world=world,
                                       ifft_fn=ifft_fn, begin_fn=begin_fn, put_fn=put_fn, end_fn=end_fn,
                                       all_gather_fn=qd.torch_all_gather,
                                       all_gather_async_fn=qd.torch_all_gather_async)
            return b, cap.cpu().numpy().view(np.uint64)

    # ---- device-resident arm ----
    caps = []
    for _ in range(args.warmup):
        b, cap = step(d_vals)
        b.free()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    kernel_ms = {"intt": 0.0, "lde": 0.0, "leaf_hash": 0.0, "tree_levels": 0.0}
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        b, cap = step(d_vals)
        torch.cuda.synchronize()
        dev_ms += (time.perf_counter() - t0) * 1e3
        for k in kernel_ms:
            kernel_ms[k] += b.kernel_ms.get(k, 0.0)
        caps.append(cap)
        b.free()
    barrier()
    launches = ctx.launch_count - launches0
    # ---- end-to-end arm: host (pinned) buffers through the C ABI ----
    for _ in range(min(args.warmup, 2)):
        b, cap = step(h_vals)
        b.free()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        b, cap = step(h_vals)
        torch.cuda.synchronize()
        e2e_ms += (time.perf_counter() - t0) * 1e3
        assert (cap == caps[0]).all(), "e2e cap differs from device-resident cap"
        b.free()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    dev_ms /= args.steps
    e2e_ms /= args.steps
    for k in kernel_ms:
        kernel_ms[k] /= args.steps
    t = torch.tensor([dev_ms, e2e_ms, kernel_ms["intt"], kernel_ms["lde"], kernel_ms["leaf_hash"],
                      kernel_ms["tree_levels"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_intt, k_lde, k_leaf, k_tree = [float(x) for x in t.cpu()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (leaf_hash_kernel) and of the LDE ----
    peak, peak_src = peaks()
    N = n << RATE_BITS
    n_loc = N // world
    leaf_bytes = n_loc * (COLS * 8 + 32)                 # 8 B per element read + 32 B digest per leaf
    perms = n_loc * ((COLS + 7) // 8)
    # DRAM bytes of one launch from the committed ncu --set full capture of this kernel at this size
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        rec = json.load(open(tp)).get("leaf_hash_kernel", {})
        if rec.get("rows_log") == args.rows_log and rec.get("cols", 135) == COLS:
            traffic = rec.get("dram_bytes")
    roof = {"kernel": "merkle::leaf_hash_kernel", "bound": "hbm", "achieved": leaf_bytes / (k_leaf * 1e-3) / 1e9,
            "peak": peak, "unit": "GB/s", "frac": leaf_bytes / (k_leaf * 1e-3) / 1e9 / peak, "traffic": traffic,
            "peak_source": peak_src, "ms": k_leaf,
            "note": "integer-issue bound, not HBM bound: %.3g Poseidon permutations/s per GPU" % (perms / (k_leaf * 1e-3)),
            "permutations_per_s": perms / (k_leaf * 1e-3)}
    lde_bytes = COLS * n * 8 + COLS * n_loc * 8          # 8 B per coeff in + 8 B per value out
    roof_lde = {"kernel": "ntt::strided_pass_kernel + ntt::final_pass_kernel (LDE)", "bound": "hbm",
                "achieved": lde_bytes / (k_lde * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": lde_bytes / (k_lde * 1e-3) / 1e9 / peak, "ms": k_lde}
    intt_bytes = (c_hi - c_lo) * n * 16
    roof_intt = None  # the sharded path runs the iNTT outside the batch handle (qp_ifft_columns)
    if k_intt > 0:
        roof_intt = {"kernel": "ntt (iNTT)", "bound": "hbm", "achieved": intt_bytes / k_intt / 1e6, "peak": peak,
                     "unit": "GB/s", "frac": intt_bytes / k_intt / 1e6 / peak, "ms": k_intt}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = min(args.rows_log, 16)
        full_ms, sample_ms, scopes, cores = cpu_commit_ms(args.rows_log, sample)
        cpu = {"value": full_ms, "unit": "ms", "cores": cores, "kind": "port",
               "sample": "oracle from_values on 2^%d rows x %d cols (%.0f ms), scaled x%d by rows; faithful "
                         "C/OpenMP restatement, not the rustc-compiled reference" % (sample, COLS, sample_ms,
                                                                                     1 << (args.rows_log - sample)),
               "scopes_ms_sample": scopes}

    # issue-slot roofline of the same kernel: ncu (profiles/r01h_ncu_full.md) counts 16.08k warp
    # instructions per warp-permutation; one warp instruction per cycle per SM sub-partition is the ceiling
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    instr_per_warp_perm = 16080.0
    issue_peak = 148 * 4 * sm_mhz * 1e6 * 32 / instr_per_warp_perm
    roof_issue = {"kernel": "merkle::leaf_hash_kernel", "bound": "issue slots (fma-heavy + ALU + FP64 pipes share one "
                  "issue port per SM sub-partition)", "achieved": perms / (k_leaf * 1e-3), "peak": issue_peak,
                  "unit": "Poseidon permutations/s", "frac": perms / (k_leaf * 1e-3) / issue_peak,
                  "instr_per_warp_permutation": instr_per_warp_perm}

    # ---- secondary metric of BASELINE.json: proof latency at bench_recursion's degrees (N=1 only) ----
    prove = None
    if world == 1 and not args.no_prove:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_prove
            prove = {"what": "prove() from the witness on (wires/Z/quotient commitments, openings, FRI proof, "
                             "serialisation) for a synthetic 143-wire circuit with all 14 gate types of a "
                             "recursive verifier (22% PoseidonGate rows, four selector groups, copy constraints), "
                             "standard_recursion_config; witness device-resident; best of 4",
                     "runs": bench_prove.measure([12, 13, 14], [] if args.no_cpu else [12], reps=5, device=local,
                                                 verbose=False, recursion=True)}
        except Exception as e:  # the headline metric does not depend on it
            prove = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h_vals.numel() * 8),
                "d2h_bytes_per_step": int((1 << CAP_HEIGHT) * 32 // world)},
        "gpu_launches": int(launches),
        "roofline": roof, "roofline_issue": roof_issue, "roofline_lde": roof_lde, "roofline_intt": roof_intt,
        "prove": prove,
        "kernel_ms": {"intt": k_intt, "lde": k_lde, "leaf_hash": k_leaf, "tree_levels": k_tree},
        "cpu_baseline": cpu, "clocks": clocks,
        "cap0": [int(x) for x in caps[0][0]],
    }
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def main():
    global COLS, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows-log", type=int, default=ROWS_LOG)
    ap.add_argument("--cols", type=int, default=COLS,
                    help="columns of the witness matrix (BASELINE.json configs[4]: --rows-log 23 --cols 400 --gpus 8)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    args = ap.parse_args()
    COLS = args.cols
    METRIC = "commit_ms_2^%dx%d_rate%d" % (args.rows_log, COLS, RATE_BITS)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
