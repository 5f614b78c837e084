"""Multi-GPU commit: one process per GPU, torch.distributed for the plumbing.

Decomposition: leaf i of the commitment is the point g*w_N^bitrev(i), so the
top rate_bits bits of the leaf index select one of the 2^rate_bits cosets of the size-n
subgroup -- which is simultaneously a contiguous range of leaves, i.e. whole cap subtrees.
  1. "IFFT": columns are sharded across ranks (each rank uploads and inverts only its columns);
  2. all-gather of the coefficients (the ONE data-path collective: n_cols * n * 8 bytes total),
     issued in a few pieces so that
  3. "FFT + blinding" of the columns that have arrived overlaps the rest of the transfer; then
     "build Merkle tree": rank q computes, for ALL columns, the coset blocks
     [q * 2^r / W, (q+1) * 2^r / W) and hashes exactly those leaves -- complete rows, local;
  4. all-gather of the cap entries (2^cap_height * 32 bytes) for the host transcript.
The reference has no counterpart (it is single-process rayon); the result is bit-identical to
the single-GPU commitment because each leaf and each subtree is computed by the same kernels.

The compute steps are injected as callables so that the CPU `gloo` tests can drive the same
plumbing with the oracle standing in for the device.
"""
import numpy as np


def column_shard(n_cols: int, world: int, rank: int):
    """Contiguous column range of `rank`, sizes differing by at most one."""
    base, extra = divmod(n_cols, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def block_shard(rate_bits: int, cap_height: int, world: int, rank: int):
    """Coset-block range (first, count) of `rank`.  Needs world | 2^rate_bits and a shard no
    smaller than a cap subtree; otherwise the path does not shard this way."""
    blocks = 1 << rate_bits
    if world & (world - 1) or world > blocks or (1 << cap_height) < world:
        raise ValueError(
            "coset sharding needs world (=%d) to be a power of two <= 2^rate_bits (=%d) and <= 2^cap_height"
            % (world, blocks)
        )
    count = blocks // world
    return rank * count, count


def padded_cols(n_cols: int, world: int) -> int:
    return -(-n_cols // world)


GATHER_CHUNKS = 4  # the coefficient all-gather is issued in this many pieces


def sharded_commit(values_local, n_cols, degree_log, rate_bits, cap_height, *, rank, world, ifft_fn,
                   begin_fn, put_fn, end_fn, all_gather_fn, all_gather_async_fn=None):
    """Run steps 1-4.  values_local: this rank's columns [c_local][n] (column_shard).
    ifft_fn(values_local, out_rows) -> coefficients [out_rows][n] (first c_local rows meaningful)
    all_gather_fn(x) -> concatenation over ranks along axis 0
    all_gather_async_fn(x) -> (result, wait): the same, started now and complete after wait()
    begin_fn(block_first, block_count) -> batch under construction (PolynomialBatch.begin)
    put_fn(batch, coeff_rows, c0)      -> columns [c0, c0 + len) are in; their LDE may start
    end_fn(batch) -> (batch, cap_local [k][4])
    The all-gather of the coefficients runs in GATHER_CHUNKS pieces, all started up front: the LDE
    of piece k overlaps the transfer of the later pieces.  Returns (batch_local, cap_full)."""
    if all_gather_async_fn is None:
        def all_gather_async_fn(x):
            return all_gather_fn(x), (lambda: None)
    pc = padded_cols(n_cols, world)
    coeffs_local = ifft_fn(values_local, pc)                      # [pc][n], zero rows as padding
    step = -(-pc // GATHER_CHUNKS)
    pieces = []
    for r0 in range(0, pc, step):
        r1 = min(r0 + step, pc)
        pieces.append((r0, r1) + tuple(all_gather_async_fn(coeffs_local[r0:r1])))   # [world * (r1 - r0)][n]
    first, count = block_shard(rate_bits, cap_height, world, rank)
    batch = begin_fn(first, count)
    for r0, r1, gathered, wait in pieces:
        wait()
        for r in range(world):
            lo, hi = column_shard(n_cols, world, r)
            v1 = min(r1, hi - lo)                                 # rank r's rows past hi - lo are padding
            if v1 > r0:
                base = r * (r1 - r0)
                put_fn(batch, gathered[base: base + (v1 - r0)], lo + r0)
    batch, cap_local = end_fn(batch)
    cap_full = all_gather_fn(cap_local)
    return batch, cap_full


# ---- the pipelined form: per-owner broadcasts in global column order ------------------------------
# hash_leaf's sponge absorbs the columns strictly in order (core/src/hashing.rs:150-168), so what hides
# communication behind hashing is the arrival ORDER: the columns travel as small pieces in global column
# order, each broadcast from the rank that owns (uploads + inverse-transforms) it straight into every
# rank's coefficient matrix; a rank extends a piece to its cosets and advances its leaf sponges as soon
# as the piece is there, while the later pieces -- the other ranks' uploads, transforms and broadcasts --
# are still in flight.  Only the very first piece has nothing to hide behind.

def pipeline_pieces(n_cols: int, world: int, piece_cols: int = 8):
    """The global column order cut into pieces of at most piece_cols columns that never straddle an
    owner: [(owner, c0, c1)]."""
    out = []
    for r in range(world):
        lo, hi = column_shard(n_cols, world, r)
        for c0 in range(lo, hi, piece_cols):
            out.append((r, c0, min(c0 + piece_cols, hi)))
    return out


def sharded_commit_pipelined(n_cols, degree_log, rate_bits, cap_height, *, rank, world, begin_fn, produce_fn,
                             slot_fn, broadcast_fn, extend_fn, end_fn, all_gather_fn, piece_cols=8, lookahead=3):
    """begin_fn(block_first, block_count) -> batch under construction
    produce_fn(batch, c0, c1) -> buffer: this rank owns columns [c0, c1): put their coefficients in place
                                (upload + inverse transform into the batch's coefficient matrix)
    slot_fn(batch, c0, c1)    -> buffer: where columns [c0, c1) of another owner are to be received
    broadcast_fn(buffer, src) -> wait: start the broadcast from rank src; wait() orders the consumer after it
    extend_fn(batch, c0, c1)  -> LDE of the columns to this rank's cosets + leaf sponges over the prefix
    end_fn(batch) -> (batch, cap_local [k][4]);  all_gather_fn(cap_local) -> cap [2^cap_height][4]
    Every rank issues the same broadcasts in the same order (a collective); nothing else is exchanged."""
    first, count = block_shard(rate_bits, cap_height, world, rank)
    batch = begin_fn(first, count)
    pieces = pipeline_pieces(n_cols, world, piece_cols)
    waits = []

    def issue():
        owner, c0, c1 = pieces[len(waits)]
        buf = produce_fn(batch, c0, c1) if owner == rank else slot_fn(batch, c0, c1)
        waits.append(broadcast_fn(buf, owner))

    # the host stays `lookahead` pieces ahead of the consumer: far enough that transfers overlap compute,
    # near enough that the first extend is queued before the host has walked the whole list
    for k, (_, c0, c1) in enumerate(pieces):
        while len(waits) < min(len(pieces), k + 1 + lookahead):
            issue()
        waits[k]()
        extend_fn(batch, c0, c1)
    batch, cap_local = end_fn(batch)
    return batch, all_gather_fn(cap_local)


# ---- torch.distributed glue (NCCL on GPUs, gloo on CPU) --------------------------------------

def torch_all_gather(x, group=None):
    """all_gather along axis 0 for torch tensors (int64 storage) or numpy uint64 arrays."""
    import torch
    import torch.distributed as dist

    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.int64)) if is_np else x.contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.numpy().view(np.uint64) if is_np else out


def torch_broadcast_async(x, src, group=None):
    """Broadcast x (torch tensor, or numpy uint64 array updated in place) from rank src, started
    asynchronously: -> wait.  NCCL: the collective is ordered after the work already queued on the CURRENT
    stream, and wait() makes the stream that is current THEN wait for it; gloo: wait() blocks the host."""
    import torch
    import torch.distributed as dist

    if isinstance(x, np.ndarray):
        t = torch.from_numpy(x.view(np.int64))      # shares memory: the broadcast lands in x
        work = dist.broadcast(t, src, group=group, async_op=True)
        return work.wait
    work = dist.broadcast(x, src, group=group, async_op=True)
    return work.wait


def torch_all_gather_async(x, group=None):
    """all_gather along axis 0 started asynchronously: -> (result, wait).  wait() makes the current
    stream (NCCL) or the host (gloo) wait for the collective."""
    import torch
    import torch.distributed as dist

    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.int64)) if is_np else x.contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    work = dist.all_gather_into_tensor(out, t, group=group, async_op=True)
    res = out.numpy().view(np.uint64) if is_np else out
    return res, work.wait
