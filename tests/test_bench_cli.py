"""bench.py's command line on the CPU: the reference arm (the oracle on a bounded sample) prints one
JSON line with the contract's keys; guards the GPU-box runs against argument / syntax slips."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows-log", "8",
                          "--cols", "9", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "commit_ms_2^8x9_rate3" and rec["unit"] == "ms"
    assert rec["higher_is_better"] is False and rec["cpu_baseline"]["kind"] == "port"
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]


def test_reference_arm_merkle_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "merkle",
                          "--leaves-log", "8", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "merkle_tree_new_ms_2^8x135" and rec["sweep"][0]["leaves_log"] == 8


def test_both_arms_generate_the_same_witness():
    """bench.py's numpy (CPU arm) and torch (device arm) witness generators are one function."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench

    a = bench.synth_columns_numpy(3, 9, 1 << 10)
    b = bench.synth_columns_torch(3, 9, 1 << 10, "cpu").numpy().view(np.uint64)
    assert (a == b).all() and int(a.max()) < bench.P_GL
    c = bench.synth_leaves_numpy(17, 300)
    d = bench.synth_leaves_torch(17, 300, "cpu").numpy().view(np.uint64)
    assert (c == d).all() and int(c.max()) < bench.P_GL
