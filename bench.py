#!/usr/bin/env python3
"""Headline benchmark: PolynomialBatch::from_values commit (iNTT + coset LDE + Poseidon Merkle
tree + cap) of 2^20 rows x 135 columns at rate_bits = 3, cap_height = 4 (BASELINE.json
configs[2]) on N B200s, beside the CPU restatement of the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows-log L]

One "step" = one whole commit of one synthetic witness matrix.  `value` = ms per commit with
the witness already resident in HBM (CUDA events on the library's stream, max over ranks; the wall
clock around the same region is reported as `wall_ms_per_step`); `e2e` = ms per commit through the
C ABI with HOST (pinned) buffers, the host->device copy of the witness and the device->host read of
the cap inside the timed region; `e2e_pageable` = the same through `qp_batch_from_values_cols` from
135 separate pageable column vectors (what the reference's `Vec<PolynomialValues>` is).
N > 1: coset-sharded strong scaling (dist.py).

Both arms commit to the SAME witness (`synth_columns_*`: SplitMix64 of (seed, column, row), reduced
to canonical form): at N = 1 the CPU restatement runs once on the very matrix the GPU committed to, at
full size, its wall time is `cpu_baseline.value` and its Merkle cap must equal the GPU's
(`cap_equal_cpu`).  `--impl reference` times the same full-size commit per step (no extrapolation).

    python bench.py --workload merkle [--leaves-log L] [--gpus N]     MerkleTree::new sweep (configs[3])
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
# synthetic witness: column c, row i = SplitMix64((seed << 40) + (c << 32) + i) reduced to [0, p).
# The numpy form (CPU arm) and the torch form (generated on the device) are the same function, so
# every arm, every rank and every N commits to the same matrix.
# ---------------------------------------------------------------------------------------------
P_GL = 0xFFFFFFFF00000001
SEED = 42


def synth_columns_numpy(c_lo, c_hi, n, seed=SEED):
    out = np.empty((c_hi - c_lo, n), dtype=np.uint64)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for c in range(c_lo, c_hi):
            z = (i + np.uint64((seed << 40) + (c << 32))) * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[c - c_lo] = np.where(z >= np.uint64(P_GL), z - np.uint64(P_GL), z)
    return out


def synth_columns_torch(c_lo, c_hi, n, dev, seed=SEED):
    import torch

    def s64(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >> 63 else x

    def lsr(z, k):  # logical shift right on int64 storage
        return (z >> k) & ((1 << (64 - k)) - 1)

    out = torch.empty((c_hi - c_lo, n), dtype=torch.int64, device=dev)
    i = torch.arange(n, dtype=torch.int64, device=dev)
    for c in range(c_lo, c_hi):
        z = (i + ((seed << 40) + (c << 32))) * s64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        ge = (z < 0) & (z >= -0xFFFFFFFF)  # unsigned z >= p
        out[c - c_lo] = torch.where(ge, z + 0xFFFFFFFF, z)
    return out


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference path (oracle, kind "port"),
# at FULL size on the same witness -- nothing is extrapolated
# ---------------------------------------------------------------------------------------------
def oracle_threads():
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm uses every host core
    os.environ["OMP_NUM_THREADS"] = os.environ.get("QP_ORACLE_THREADS", str(os.cpu_count()))
    import oracle

    return oracle, oracle.lib().orc_num_threads()


def cpu_commit(vals):
    """One oracle PolynomialBatch::from_values on `vals` (all host threads): (ms, scopes, batch)."""
    oracle, _ = oracle_threads()
    t0 = time.perf_counter()
    b = oracle.PolynomialBatch.from_values(vals, RATE_BITS, CAP_HEIGHT)
    return (time.perf_counter() - t0) * 1e3, b.scope_ms, b


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    if args.workload == "merkle":
        return run_reference_merkle(args)
    _, cores = oracle_threads()
    vals = synth_columns_numpy(0, COLS, 1 << args.rows_log)
    # every step is one FULL-size commit; QP_REF_BUDGET_S bounds the whole run (steps_run says how many ran)
    budget = float(os.environ.get("QP_REF_BUDGET_S", 420))   # "ends within a few minutes"
    t_start = time.perf_counter()
    times, scopes, cap0 = [], None, None
    warm = 0
    for _ in range(args.warmup):
        if warm and time.perf_counter() - t_start > budget / 4:
            break
        _, _, b = cpu_commit(vals)
        cap0 = [int(x) for x in b.cap[0]]
        del b
        warm += 1
    for _ in range(args.steps):
        if times and time.perf_counter() - t_start + times[-1] / 1e3 > budget:
            break
        ms, scopes, b = cpu_commit(vals)
        cap0 = [int(x) for x in b.cap[0]]
        del b
        times.append(ms)
    ms = sum(times) / len(times)
    desc = ("oracle (C/OpenMP restatement of the reference's rayon path) on the full 2^%d rows x %d cols, "
            "%d timed runs, nothing scaled" % (args.rows_log, COLS, len(times)))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": desc,
                         "scopes_ms": scopes, "min_ms": min(times), "max_ms": max(times)},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cap0": cap0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# configs[3]: MerkleTree::new Poseidon sweep, 2^16 .. 2^24 leaves x 135 elements, cap height 4,
# subtree-sharded at N > 1 (one shard of whole cap subtrees per GPU, qp_merkle_tree_new_shard)
# ---------------------------------------------------------------------------------------------
MERKLE_LEAF_LEN = 135


def synth_leaves_numpy(first, count, leaf_len=MERKLE_LEAF_LEN, seed=SEED + 1):
    """Leaf-major rows [count][leaf_len]: element (i, c) = SplitMix64((seed << 40) + i * leaf_len + c), canonical."""
    out = np.empty((count, leaf_len), dtype=np.uint64)
    step = 1 << 16
    with np.errstate(over="ignore"):
        for r0 in range(0, count, step):
            r1 = min(r0 + step, count)
            idx = (np.arange(first + r0, first + r1, dtype=np.uint64)[:, None] * np.uint64(leaf_len)
                   + np.arange(leaf_len, dtype=np.uint64)[None, :])
            z = (idx + np.uint64(seed << 40)) * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[r0:r1] = np.where(z >= np.uint64(P_GL), z - np.uint64(P_GL), z)
    return out


def synth_leaves_torch(first, count, dev, leaf_len=MERKLE_LEAF_LEN, seed=SEED + 1):
    import torch

    def s64(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >> 63 else x

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    out = torch.empty((count, leaf_len), dtype=torch.int64, device=dev)
    step = 1 << 20
    col = torch.arange(leaf_len, dtype=torch.int64, device=dev)[None, :]
    for r0 in range(0, count, step):
        r1 = min(r0 + step, count)
        z = torch.arange(first + r0, first + r1, dtype=torch.int64, device=dev)[:, None] * leaf_len + col
        z = (z + (seed << 40)) * s64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        ge = (z < 0) & (z >= -0xFFFFFFFF)
        out[r0:r1] = torch.where(ge, z + 0xFFFFFFFF, z)
    return out


def merkle_config(sizes, world):
    return {"workload": "MerkleTree::new (Poseidon, hash_leaf + two_to_one) sweep: 2^%s leaves x %d Goldilocks elements, "
                        "cap_height=%d" % ("/".join(str(x) for x in sizes), MERKLE_LEAF_LEN, CAP_HEIGHT),
            "parallelism": "single GPU" if world == 1 else "subtree-sharded x%d (whole cap subtrees per GPU, no data-path "
                           "collective; all-gather of the cap)" % world,
            "l2": "leaves of every size from 2^18 up (%.0f MB at 2^18) exceed the 126 MB L2; the smaller sizes are "
                  "flushed by the larger ones between timed iterations" % (MERKLE_LEAF_LEN * 8 * 2 ** 18 / 1e6)}


def run_reference_merkle(args):
    oracle, cores = oracle_threads()
    sizes = [args.leaves_log] if args.leaves_log else [16, 18, 20]
    sweep = []
    for L in sizes:
        leaves = synth_leaves_numpy(0, 1 << L)
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            t = oracle.MerkleTree(leaves, CAP_HEIGHT)
            dt = (time.perf_counter() - t0) * 1e3
            if it >= args.warmup:
                times.append(dt)
        sweep.append({"leaves_log": L, "ms": sum(times) / len(times), "cap0": [int(x) for x in t.cap[0]]})
        del t, leaves
    ms = sweep[-1]["ms"]
    metric = "merkle_tree_new_ms_2^%dx%d" % (sizes[-1], MERKLE_LEAF_LEN)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": merkle_config(sizes, 1), "sweep": sweep,
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
                         "sample": "oracle MerkleTree::new at full size, every step"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_merkle(args):
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch

    import qp_plonky2_b200 as qp
    import qp_plonky2_b200.dist as qd

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = qp.Context(local, max_lde_log=4, stream=stream.cuda_stream)
    sizes = [args.leaves_log] if args.leaves_log else [16, 18, 20, 22, 24]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    sweep = []
    for L in sizes:
        n_loc = (1 << L) // world
        d_leaves = synth_leaves_torch(rank * n_loc, n_loc, dev)
        torch.cuda.synchronize()

        def step(src):
            t = qp.MerkleTree(ctx, src, CAP_HEIGHT, shard=rank, n_shards=world)
            cap = t.cap
            if dist is not None:
                cap = qd.torch_all_gather(torch.from_numpy(cap.view(np.int64)).to(dev)).cpu().numpy().view(np.uint64)
            return t, cap

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(args.warmup):
            step(d_leaves)[0].free()
        dev_ms, cap = 0.0, None
        for _ in range(args.steps):
            barrier()
            ev0.record(stream)
            t, cap = step(d_leaves)
            ev1.record(stream)
            torch.cuda.synchronize()
            dev_ms += ev0.elapsed_time(ev1)
            t.free()
        dev_ms /= args.steps
        # end to end from pinned host rows (H2D inside), sizes whose leaves fit comfortably in host memory
        e2e_ms = None
        if L <= 22:
            h_leaves = torch.empty(d_leaves.shape, dtype=torch.int64).pin_memory()
            h_leaves.copy_(d_leaves)
            torch.cuda.synchronize()
            for _ in range(min(args.warmup, 2)):
                step(h_leaves)[0].free()
            e2e_ms = 0.0
            for _ in range(args.steps):
                barrier()
                t0 = time.perf_counter()
                t, cap2 = step(h_leaves)
                torch.cuda.synchronize()
                e2e_ms += (time.perf_counter() - t0) * 1e3
                assert (cap2 == cap).all()
                t.free()
            e2e_ms /= args.steps
        tt = torch.tensor([dev_ms, e2e_ms if e2e_ms is not None else 0.0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_max = [float(x) for x in tt.cpu()]
        perms = (1 << L) * ((MERKLE_LEAF_LEN + 7) // 8) + (1 << L) - (1 << CAP_HEIGHT)
        rec = {"leaves_log": L, "ms": dev_ms, "e2e_ms": e2e_max if e2e_ms is not None else None,
               "permutations_per_s": perms / (dev_ms * 1e-3), "cap0": [int(x) for x in cap[0]]}
        # the CPU restatement beside it, same rows (rank 0, N = 1, sizes it finishes in seconds)
        if world == 1 and not args.no_cpu and L <= 20:
            oracle, cores = oracle_threads()
            host = d_leaves.cpu().numpy().view(np.uint64)
            t0 = time.perf_counter()
            ot = oracle.MerkleTree(host, CAP_HEIGHT)
            rec["cpu_ms"] = (time.perf_counter() - t0) * 1e3
            rec["cpu_cores"] = cores
            rec["cap_equal_cpu"] = bool((ot.cap == cap).all())
            assert rec["cap_equal_cpu"], "GPU MerkleTree::new cap differs from the CPU oracle's"
            del ot, host
        sweep.append(rec)
        del d_leaves
        torch.cuda.empty_cache()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        last = sweep[-1]
        with_cpu = [r for r in sweep if "cpu_ms" in r]
        cpu = None
        if with_cpu:
            r = with_cpu[-1]
            cpu = {"value": r["cpu_ms"], "unit": "ms", "cores": r["cpu_cores"], "kind": "port",
                   "sample": "oracle MerkleTree::new on the same 2^%d x %d rows, full size, 1 run (the largest size the "
                             "CPU arm runs; GPU value there: %.3f ms)" % (r["leaves_log"], MERKLE_LEAF_LEN, r["ms"])}
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        ipp = 16080.0
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            ipp = float(json.load(open(tp)).get("leaf_hash_kernel", {}).get("instr_per_warp_permutation", ipp))
        peak_perm = 148 * 4 * sm_mhz * 1e6 * 32 / ipp * world
        line = {
            "metric": "merkle_tree_new_ms_2^%dx%d" % (last["leaves_log"], MERKLE_LEAF_LEN), "value": last["ms"], "unit": "ms",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": last["ms"],
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": merkle_config(sizes, world), "sweep": sweep,
            "e2e": {"value": next((r["e2e_ms"] for r in reversed(sweep) if r["e2e_ms"] is not None), None), "unit": "ms",
                    "what": "largest size with a host-resident arm (2^%d)" % max([r["leaves_log"] for r in sweep if r["e2e_ms"] is not None], default=0),
                    "h2d_bytes_per_step": int(max([(1 << r["leaves_log"]) for r in sweep if r["e2e_ms"] is not None], default=0)
                                              * MERKLE_LEAF_LEN * 8 // world),
                    "d2h_bytes_per_step": int((1 << CAP_HEIGHT) * 32 // world)},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "merkle::leaf_hash_kernel + tree levels", "bound": "issue",
                         "achieved": last["permutations_per_s"], "peak": peak_perm, "unit": "Poseidon permutations/s",
                         "frac": last["permutations_per_s"] / peak_perm, "traffic": None,
                         "instr_per_warp_permutation": ipp},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def workload_config(args, world):
    return {"workload": "PolynomialBatch::from_values 2^%d rows x %d cols, rate_bits=%d, cap_height=%d, "
                        "blinding=false (LDE + Poseidon Merkle commit)" % (args.rows_log, COLS, RATE_BITS, CAP_HEIGHT),
            "parallelism": "single GPU" if world == 1 else "coset-sharded x%d (column-sharded upload + iNTT, per-owner "
                           "NCCL broadcasts of 8-column pieces in column order overlapped with LDE + leaf hashing, "
                           "all-gather of the cap)" % world,
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
        # high-priority NCCL stream: a broadcast kernel launched while a leaf-hash kernel (thousands of blocks)
        # is running must not queue behind all of that kernel's blocks, or every piece of the pipeline pays it
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=not os.environ.get("QP_BENCH_LOW_PRIO"))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one stream for torch (NCCL, tensor ops) and the library, so their work is ordered
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = qp.Context(local, max_lde_log=args.rows_log + RATE_BITS, stream=stream.cuda_stream)
    n = 1 << args.rows_log
    c_lo, c_hi = (0, COLS) if world == 1 else qd.column_shard(COLS, world, rank)

    # synthetic witness: uniform canonical Goldilocks elements (synth_columns_*), the same matrix in every
    # arm; a rank generates only its own column shard
    d_vals = synth_columns_torch(c_lo, c_hi, n, dev)
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
        # coset-sharded commit, pipelined (dist.sharded_commit_pipelined): pieces of 8 columns travel in
        # global column order, each uploaded + inverse-transformed by its owner on a producer stream (its own
        # context) and broadcast (NCCL) straight into every rank's coefficient matrix; the main stream extends
        # and hashes a piece as soon as it is there
        s2 = torch.cuda.Stream(device=dev, priority=0 if os.environ.get("QP_BENCH_LOW_PRIO") else -1)  # producer: ahead of the hashing
        copy_stream = torch.cuda.Stream(device=dev)
        ctx2 = qp.Context(local, max_lde_log=args.rows_log, stream=s2.cuda_stream)
        d_in = torch.empty((c_hi - c_lo, n), dtype=torch.int64, device=dev)   # staging of this rank's host columns
        cur = {}

        def begin_fn(first, count):
            b = qp.PolynomialBatch.begin(ctx, COLS, args.rows_log, RATE_BITS, False, CAP_HEIGHT,
                                         block_first=first, block_count=count)
            ev = torch.cuda.Event()
            ev.record(stream)          # the coefficient matrix is a stream-ordered allocation of the main stream
            s2.wait_event(ev)
            cur["coeffs"] = b.coeffs_slot(0, COLS)      # one zero-copy view of the whole matrix, sliced per piece
            return b

        def produce_fn(b, c0, c1):
            slot = cur["coeffs"][c0:c1]
            src = cur["src"]
            with torch.cuda.stream(s2):
                if src.is_cuda:
                    v = src[c0 - c_lo: c1 - c_lo]
                else:
                    v = d_in[c0 - c_lo: c1 - c_lo]
                    with torch.cuda.stream(copy_stream):
                        v.copy_(src[c0 - c_lo: c1 - c_lo], non_blocking=True)
                        e = torch.cuda.Event()
                        e.record(copy_stream)
                    s2.wait_event(e)
                ctx2.ifft_columns(v, out_device=slot, sync=False)
            return slot

        def slot_fn(b, c0, c1):
            return cur["coeffs"][c0:c1]

        def broadcast_fn(buf, src_rank):
            with torch.cuda.stream(s2):    # ordered after the producer's transform (and the allocation)
                w = dist.broadcast(buf, src_rank, async_op=True)
            return w.wait                  # called on the main stream: extend / hash wait for the piece

        def extend_fn(b, c0, c1):
            b.extend_columns(c0, c1 - c0, absorb=True)

        def end_fn(batch):
            batch.end()
            return batch, torch.from_numpy(batch.merkle_tree.cap.view(np.int64)).to(dev)

        def step(src):
            cur["src"] = src
            b, cap = qd.sharded_commit_pipelined(COLS, args.rows_log, RATE_BITS, CAP_HEIGHT, rank=rank, world=world,
                                                 begin_fn=begin_fn, produce_fn=produce_fn, slot_fn=slot_fn,
                                                 broadcast_fn=broadcast_fn, extend_fn=extend_fn, end_fn=end_fn,
                                                 all_gather_fn=qd.torch_all_gather, piece_cols=args.piece_cols,
                                                 lookahead=args.lookahead)
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
    # CUDA events on the stream the library launches on (= torch's current stream, see above)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    kernel_ms = {"intt": 0.0, "lde": 0.0, "leaf_hash": 0.0, "tree_levels": 0.0}
    dev_ms = wall_ms = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        b, cap = step(d_vals)
        ev1.record(stream)
        torch.cuda.synchronize()
        wall_ms += (time.perf_counter() - t0) * 1e3
        dev_ms += ev0.elapsed_time(ev1)
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
        t0 = time.perf_counter()   # host buffers in, host cap out: the wall clock IS the end-to-end time
        b, cap = step(h_vals)
        torch.cuda.synchronize()
        e2e_ms += (time.perf_counter() - t0) * 1e3
        assert (cap == caps[0]).all(), "e2e cap differs from device-resident cap"
        b.free()
    barrier()
    # ---- end to end from what the reference passes: one pageable heap vector per column ----
    e2e_pg_ms, e2e_pg_min = None, 1e30
    if world == 1 and hasattr(qp.PolynomialBatch, "from_values_cols"):
        cols_pg = [np.array(h_vals[c].numpy(), copy=True).view(np.uint64) for c in range(COLS)]
        for _ in range(min(args.warmup, 2)):
            qp.PolynomialBatch.from_values_cols(ctx, cols_pg, RATE_BITS, False, CAP_HEIGHT).free()
        e2e_pg_ms = 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            b = qp.PolynomialBatch.from_values_cols(ctx, cols_pg, RATE_BITS, False, CAP_HEIGHT)
            cap = b.merkle_tree.cap
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            e2e_pg_ms += dt
            e2e_pg_min = min(e2e_pg_min, dt)
            assert (cap == caps[0]).all(), "pageable-columns cap differs from device-resident cap"
            b.free()
        e2e_pg_ms /= args.steps
        del cols_pg
    clocks = sampler.stop() if rank == 0 else None

    dev_ms /= args.steps
    wall_ms /= args.steps
    e2e_ms /= args.steps
    for k in kernel_ms:
        kernel_ms[k] /= args.steps
    t = torch.tensor([dev_ms, e2e_ms, kernel_ms["intt"], kernel_ms["lde"], kernel_ms["leaf_hash"],
                      kernel_ms["tree_levels"], wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_intt, k_lde, k_leaf, k_tree, wall_ms = [float(x) for x in t.cpu()]
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
    roof_hbm = {"kernel": "merkle::leaf_hash_kernel", "bound": "hbm", "achieved": leaf_bytes / (k_leaf * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": leaf_bytes / (k_leaf * 1e-3) / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "ms": k_leaf,
                "note": "reported because the contract asks for it; this kernel is bound by instruction issue, "
                        "not by HBM (see `roofline`)"}
    lde_bytes = COLS * n * 8 + COLS * n_loc * 8          # 8 B per coeff in + 8 B per value out
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    roof_lde = {"kernel": "ntt::strided_pass_kernel + ntt::final_pass_kernel (LDE)", "bound": "hbm",
                "achieved": lde_bytes / (k_lde * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": lde_bytes / (k_lde * 1e-3) / 1e9 / peak, "ms": k_lde,
                "traffic": tj.get("lde_passes", {}).get("dram_bytes") if world == 1 and args.rows_log == 20 and COLS == 135 else None,
                "note": "ALU-pipe bound (74-83 % busy in ncu, profiles/): 64-bit modular butterflies cost ~15 integer "
                        "instructions per element per stage; the HBM fraction is reported because the contract asks"}
    intt_bytes = (c_hi - c_lo) * n * 16
    roof_intt = None  # the sharded path runs the iNTT outside the batch handle (qp_ifft_columns)
    if k_intt > 0:
        roof_intt = {"kernel": "ntt (iNTT)", "bound": "hbm", "achieved": intt_bytes / k_intt / 1e6, "peak": peak,
                     "unit": "GB/s", "frac": intt_bytes / k_intt / 1e6 / peak, "ms": k_intt,
                     "traffic": tj.get("intt_passes", {}).get("dram_bytes") if args.rows_log == 20 and COLS == 135 else None}

    # ---- CPU baseline: the oracle, once, at FULL size on the very witness the GPU committed to ----
    cpu, cap_equal_cpu = None, None
    if world == 1 and not args.no_cpu:
        _, cores = oracle_threads()
        host = h_vals.numpy().view(np.uint64)
        cpu_ms, scopes, ob = cpu_commit(host)
        cap_equal_cpu = bool((ob.cap == caps[0]).all())
        # and the digests of the first and the last cap subtree, against the device's
        b, _ = step(d_vals)
        dg = b.merkle_tree.digests
        per = dg.shape[0] >> CAP_HEIGHT
        digests_equal = bool((dg[:per] == ob.digests[:per]).all() and (dg[-per:] == ob.digests[-per:]).all())
        b.free()
        del ob, dg
        cpu = {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port",
               "sample": "full size, 1 run: oracle from_values on the same 2^%d rows x %d cols the GPU arm committed to; "
                         "faithful C/OpenMP restatement, not the rustc-compiled reference" % (args.rows_log, COLS),
               "scopes_ms": scopes, "cap_equal": cap_equal_cpu, "digest_blocks_equal": digests_equal}
        assert cap_equal_cpu and digests_equal, "GPU commitment differs from the CPU oracle's at full size"

    # issue-slot roofline of the same kernel: ncu counts warp instructions per warp-permutation
    # (profiles/traffic.json "instr_per_warp_permutation"); one warp instruction per cycle per SM
    # sub-partition is the ceiling
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    instr_per_warp_perm = 16080.0
    if os.path.exists(tp):
        instr_per_warp_perm = float(json.load(open(tp)).get("leaf_hash_kernel", {}).get("instr_per_warp_permutation", 16080.0))
    issue_peak = 148 * 4 * sm_mhz * 1e6 * 32 / instr_per_warp_perm
    roof = {"kernel": "merkle::leaf_hash_kernel", "bound": "issue", "achieved": perms / (k_leaf * 1e-3),
            "peak": issue_peak, "unit": "Poseidon permutations/s", "frac": perms / (k_leaf * 1e-3) / issue_peak,
            "traffic": traffic, "ms": k_leaf, "instr_per_warp_permutation": instr_per_warp_perm,
            "note": "the bound that binds: one warp instruction per cycle per SM sub-partition at the ncu-measured "
                    "instruction count; the HBM view of the same launch is `roofline_hbm` (SURVEY 8d: HBM is not "
                    "the yardstick for the hash kernels)"}

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
                                                 verbose=False, recursion=True),
                     # the same circuit with two lookup tables (lookup argument: prover.rs:489-636, vanishing_poly.rs:521-680)
                     "runs_with_lookups": bench_prove.measure([12], [] if args.no_cpu else [12], reps=5, device=local,
                                                              verbose=False, recursion=True, lookups=True)}
        except Exception as e:  # the headline metric does not depend on it
            prove = {"error": repr(e)}
        try:   # BASELINE.json configs[0]: the factorial example as a real circuit
            prove["factorial_example"] = bench_prove.measure_factorial(cpu=not args.no_cpu, device=local)
        except Exception as e:
            prove["factorial_example"] = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "wall_ms_per_step": wall_ms,
        "config": workload_config(args, world),
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h_vals.numel() * 8),
                "d2h_bytes_per_step": int((1 << CAP_HEIGHT) * 32 // world)},
        "e2e_pageable": None if e2e_pg_ms is None else {
            "value": e2e_pg_ms, "min": e2e_pg_min, "unit": "ms", "what": "qp_batch_from_values_cols: %d separate pageable column "
            "vectors (the reference's Vec<PolynomialValues>), staged through the library's pinned ring" % COLS},
        "gpu_launches": int(launches), "cap_equal_cpu": cap_equal_cpu,
        "roofline": roof, "roofline_hbm": roof_hbm, "roofline_lde": roof_lde, "roofline_intt": roof_intt,
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
    ap.add_argument("--workload", default="commit", choices=["commit", "merkle"],
                    help="commit: PolynomialBatch::from_values (headline); merkle: MerkleTree::new sweep (configs[3])")
    ap.add_argument("--leaves-log", type=int, default=None,
                    help="merkle workload: one size (default: the sweep 2^16..2^24 on the GPU, 2^16..2^20 on the CPU arm)")
    ap.add_argument("--piece-cols", type=int, default=8, help="N > 1: columns per broadcast piece")
    ap.add_argument("--lookahead", type=int, default=3, help="N > 1: pieces the host issues ahead of the consumer")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    args = ap.parse_args()
    COLS = args.cols
    METRIC = "commit_ms_2^%dx%d_rate%d" % (args.rows_log, COLS, RATE_BITS)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "merkle":
        run_merkle(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
