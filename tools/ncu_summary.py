#!/usr/bin/env python3
"""Summarise an Nsight Compute report (.ncu-rep) into the small text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_full.md
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r01_launches.md

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe inst %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe cycles % (IMAD)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe cycles %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe inst % (DFMA)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe inst % (I2F)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (max unit)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not_selected"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg_throttle"),
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % rep)
        f.write("Per-launch values under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
        for r in rows[2:]:
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]][:90])
            for k, label in KEYS:
                if k in idx and r[idx[k]] not in ("", "-nan"):
                    f.write("| %s (`%s`) | %s | %s |\n" % (label, k, r[idx[k]], U[idx[k]]))
    print("wrote", out)


def launches(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        name = r[ki].split("(")[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none) of %s\n\n" % path)
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, t) in agg.items():
            f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, n, t, 100 * t / tot))
        f.write("| **total** | | %.3f | |\n" % tot)
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
