"""World-size-2 (and 4) `gloo` test of the multi-GPU plumbing (qp-plonky2_b200/dist.py): column
sharding of the iNTT, the padded all-gather of coefficients, coset-block assignment and the
all-gather of cap entries.  The oracle stands in for the device compute, so this runs on CPU."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cols, lg_n, rate, cap_h, ret):
    import torch.distributed as dist

    import oracle
    import qp_plonky2_b200.dist as qd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vals = oracle.rand_felts((n_cols, 1 << lg_n), 42)
        want = oracle.PolynomialBatch.from_values(vals, rate, cap_h)
        lo, hi = qd.column_shard(n_cols, world, rank)

        def ifft_fn(v, rows):
            out = np.zeros((rows, 1 << lg_n), dtype=np.uint64)
            for k in range(v.shape[0]):
                out[k] = oracle.ifft(v[k])
            return out

        state = {}

        def begin_fn(first, count):
            state["first"], state["count"] = first, count
            state["seen"] = np.zeros(n_cols, dtype=int)
            return np.zeros((n_cols, 1 << lg_n), dtype=np.uint64)

        def put_fn(batch, rows, c0):
            batch[c0:c0 + rows.shape[0]] = rows
            state["seen"][c0:c0 + rows.shape[0]] += 1

        def end_fn(coeffs_all):
            assert (state["seen"] == 1).all(), "every column must be put exactly once"
            assert (coeffs_all == want.polynomials).all(), "gathered coefficients differ"
            b = oracle.PolynomialBatch.from_coeffs(coeffs_all, rate, cap_h)
            per_block = (1 << cap_h) >> rate
            first, count = state["first"], state["count"]
            return b, b.cap[first * per_block : (first + count) * per_block].copy()

        _, cap = qd.sharded_commit(vals[lo:hi], n_cols, lg_n, rate, cap_h, rank=rank, world=world,
                                   ifft_fn=ifft_fn, begin_fn=begin_fn, put_fn=put_fn, end_fn=end_fn,
                                   all_gather_fn=qd.torch_all_gather,
                                   all_gather_async_fn=qd.torch_all_gather_async)
        ret[rank] = bool((cap == want.cap).all())
    finally:
        dist.destroy_process_group()


def _worker_pipelined(rank, world, port, n_cols, lg_n, rate, cap_h, piece_cols, ret):
    import torch.distributed as dist

    import oracle
    import qp_plonky2_b200.dist as qd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << lg_n
        vals = oracle.rand_felts((n_cols, n), 43)
        want = oracle.PolynomialBatch.from_values(vals, rate, cap_h)
        lo, hi = qd.column_shard(n_cols, world, rank)
        state = {"order": []}

        def begin_fn(first, count):
            state["first"], state["count"] = first, count
            return np.zeros((n_cols, n), dtype=np.uint64)           # the coefficient matrix under construction

        def produce_fn(coeffs, c0, c1):
            assert lo <= c0 < c1 <= hi, "asked to produce a column this rank does not own"
            for c in range(c0, c1):
                coeffs[c] = oracle.ifft(vals[c])
            return coeffs[c0:c1]

        def slot_fn(coeffs, c0, c1):
            assert c1 <= lo or c0 >= hi, "a foreign slot inside this rank's own shard"
            return coeffs[c0:c1]

        def extend_fn(coeffs, c0, c1):
            assert (coeffs[c0:c1] == want.polynomials[c0:c1]).all(), "piece arrived wrong or late"
            state["order"].append((c0, c1))

        def end_fn(coeffs):
            assert state["order"] == [(a, b) for _, a, b in qd.pipeline_pieces(n_cols, world, piece_cols)]
            assert state["order"][0][0] == 0 and state["order"][-1][1] == n_cols   # global column order, complete
            b = oracle.PolynomialBatch.from_coeffs(coeffs, rate, cap_h)
            per_block = (1 << cap_h) >> rate
            first, count = state["first"], state["count"]
            return b, b.cap[first * per_block: (first + count) * per_block].copy()

        _, cap = qd.sharded_commit_pipelined(n_cols, lg_n, rate, cap_h, rank=rank, world=world, begin_fn=begin_fn,
                                             produce_fn=produce_fn, slot_fn=slot_fn,
                                             broadcast_fn=qd.torch_broadcast_async, extend_fn=extend_fn, end_fn=end_fn,
                                             all_gather_fn=qd.torch_all_gather, piece_cols=piece_cols)
        ret[rank] = bool((cap == want.cap).all())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cols,piece_cols", [(2, 5, 8), (2, 19, 4), (4, 35, 8), (4, 7, 1)])
def test_sharded_commit_pipelined_plumbing(world, n_cols, piece_cols):
    """The per-owner broadcast pipeline (dist.sharded_commit_pipelined): every rank sees every column exactly
    once, in global column order, with the owner's coefficients, and the gathered cap is the oracle's."""
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_pipelined, args=(world, port, n_cols, 6, 3, 4, piece_cols, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_pipeline_pieces():
    import qp_plonky2_b200.dist as qd

    for n_cols, world, pc in [(135, 8, 8), (135, 2, 8), (143, 4, 16), (5, 8, 8), (400, 8, 8)]:
        pieces = qd.pipeline_pieces(n_cols, world, pc)
        assert pieces[0][1] == 0 and pieces[-1][2] == n_cols
        assert all(a[2] == b[1] for a, b in zip(pieces, pieces[1:]))             # contiguous, in order
        for owner, c0, c1 in pieces:
            lo, hi = qd.column_shard(n_cols, world, owner)
            assert lo <= c0 < c1 <= hi and c1 - c0 <= pc                        # never straddles an owner


@pytest.mark.parametrize("world,n_cols", [(2, 5), (2, 8), (4, 7), (2, 19), (4, 35)])
def test_sharded_commit_plumbing(world, n_cols):
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_cols, 6, 3, 4, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_arithmetic():
    import qp_plonky2_b200.dist as qd

    for n_cols in (1, 5, 135, 143):
        for world in (1, 2, 4, 8):
            spans = [qd.column_shard(n_cols, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_cols
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
            assert qd.padded_cols(n_cols, world) == max(b - a for a, b in spans)
    assert [qd.block_shard(3, 4, 8, r) for r in range(8)] == [(r, 1) for r in range(8)]
    assert [qd.block_shard(3, 4, 2, r) for r in range(2)] == [(0, 4), (4, 4)]
    with pytest.raises(ValueError):
        qd.block_shard(3, 4, 16, 0)      # more ranks than cosets
    with pytest.raises(ValueError):
        qd.block_shard(3, 1, 4, 0)       # a shard would be smaller than a cap subtree
    with pytest.raises(ValueError):
        qd.block_shard(3, 4, 3, 0)       # not a power of two
