#!/usr/bin/env python3
"""Single-process multi-device commit (qp_mbatch_from_values_cols) at the headline shape, from 135 pageable
host columns, on every GPU of the box: ms per commit (wall, best of 4) and the cap's first digest."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import qp_plonky2_b200 as qp  # noqa: E402

if __name__ == "__main__":
    rows_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    D = torch.cuda.device_count()
    cols = [np.array(c, copy=True) for c in bench.synth_columns_numpy(0, 135, 1 << rows_log)]
    memory = "pageable"
    if os.environ.get("QP_BENCH_PINNED"):     # the same columns in pinned host memory (no staging copy)
        keep = [torch.from_numpy(c.view(np.int64)).pin_memory() for c in cols]
        cols = [k.numpy().view(np.uint64) for k in keep]
        memory = "pinned"
    for d in [x for x in (1, 2, 4, 8) if x <= D]:
        m = qp.MultiContext(list(range(d)), max_lde_log=rows_log + 3)
        best, cap0 = 1e9, None
        for it in range(5):
            t0 = time.perf_counter()
            mb = qp.MultiBatch.from_values_cols(m, cols, 3, False, 4)
            cap = mb.cap
            dt = (time.perf_counter() - t0) * 1e3
            mb.free()
            if it:
                best = min(best, dt)
            cap0 = [int(x) for x in cap[0]]
        m.close()
        print(json.dumps({"devices": d, "host_memory": memory, "e2e_ms": best, "cap0": cap0}), flush=True)
