"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle
on the same seeded inputs -- bit-exact.  Mirrors the reference's own tests:
Poseidon KATs, FFT identities, every-leaf Merkle round trips, from_values -> openings."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from oracle import pyref

P = oracle.P


@pytest.fixture(scope="module")
def qp():
    import qp_plonky2_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(qp):
    c = qp.Context(0, max_lde_log=24)
    yield c
    c.close()


def u64(x):
    return np.array(x, dtype=np.uint64)


# ---- Poseidon ------------------------------------------------------------------------------

def test_poseidon_kat(ctx, golden):
    """core/src/poseidon_goldilocks.rs:455-490"""
    kat = golden("poseidon_kat.json")
    inp = u64([[int(x) for x in v["input"]] for v in kat])
    out = u64([[int(x) for x in v["output"]] for v in kat])
    assert (ctx.poseidon(inp) == out).all()


def test_poseidon_random_and_noncanonical(ctx):
    st = oracle.rand_felts((4096, 12), 1)
    st[0] = u64([2**64 - 1] * 12)           # raw u64 >= p are legal internal values
    st[1] = u64([P] * 12)
    st[2] = u64([P - 1, P, P + 1, 2**64 - 1, 0, 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63, P - 2**32, 5])
    want = np.stack([oracle.poseidon(s) for s in st])
    assert (ctx.poseidon(st) == want).all()


# ---- transforms ------------------------------------------------------------------------------

@pytest.mark.parametrize("lg", list(range(0, 15)) + [16, 17])
def test_fft_matches_oracle(ctx, lg):
    """fft / coset_fft: natural order, every pass structure (single tile, 2 and 3 passes)"""
    n = 1 << lg
    c = oracle.rand_felts((3, n), 100 + lg)
    c[0, 0] = np.uint64(2**64 - 1)          # non-canonical inputs
    if n > 1:
        c[0, 1] = np.uint64(P)
    shift = oracle.lib().orc_gl_coset_shift()
    got = ctx.coset_fft(c, shift=1)
    got_c = ctx.coset_fft(c, shift=shift)
    for k in range(3):
        assert (got[k] == oracle.fft(c[k])).all()
        assert (got_c[k] == oracle.coset_fft(c[k], shift)).all()
    br = ctx.coset_fft(c, shift=shift, bit_reversed=True)
    assert (br[1] == oracle.reverse_index_bits(got_c[1])).all()


def test_fft_equals_naive(ctx):
    """field/src/fft.rs:215-249: fft == direct evaluation, degree 200 -> 256"""
    c = np.zeros(256, dtype=np.uint64)
    c[:200] = oracle.rand_felts(200, 5)
    assert (ctx.coset_fft(c) == oracle.fft_naive(c)).all()
    assert ctx.coset_fft(c).tolist() == pyref.fft([int(x) for x in c])


@pytest.mark.parametrize("lg", list(range(0, 15)) + [16, 18])
def test_ifft_matches_oracle(ctx, lg):
    n = 1 << lg
    v = oracle.rand_felts((2, n), 200 + lg)
    v[1, 0] = np.uint64(2**64 - 1)
    got = ctx.ifft_columns(v)
    for k in range(2):
        assert (got[k] == oracle.ifft(v[k])).all()


def test_fft_ifft_round_trip_large(ctx):
    """size-independent property at the full NTT size: ifft(fft(x)) == x for n = 2^20"""
    x = oracle.rand_felts((2, 1 << 20), 9)
    y = ctx.coset_fft(x)
    assert (ctx.ifft_columns(y) == x).all()
    # linearity: fft(a + b) = fft(a) + fft(b)
    s = ((x[0].astype(object) + x[1].astype(object)) % P).astype(np.uint64)
    ys = ctx.coset_fft(s)
    assert (ys == ((y[0].astype(object) + y[1].astype(object)) % P).astype(np.uint64)).all()


# ---- Merkle tree -----------------------------------------------------------------------------

@pytest.mark.parametrize("lg,leaf_len,cap_h", [
    (0, 4, 0), (1, 2, 0), (1, 2, 1), (3, 0, 1), (4, 1, 0), (5, 7, 5), (8, 7, 0), (8, 7, 1), (8, 7, 8),
    (6, 8, 2), (6, 9, 2), (7, 16, 3), (9, 135, 4), (10, 143, 4), (12, 32, 4),
])
def test_merkle_tree_new_parity(qp, ctx, lg, leaf_len, cap_h):
    """MerkleTree::new: cap, digests (reference layout), get, prove -- vs oracle"""
    leaves = oracle.rand_felts((1 << lg, leaf_len), 300 + lg)
    t = qp.MerkleTree(ctx, leaves, cap_h)
    if leaf_len == 0:
        # oracle binding needs a non-null pointer; compare with pyref instead
        dig, cap = pyref.merkle_tree([[] for _ in range(1 << lg)], cap_h)
        assert t.cap.tolist() == cap and t.digests.tolist() == dig
        return
    o = oracle.MerkleTree(leaves, cap_h)
    assert (t.cap == o.cap).all()
    assert (t.digests == o.digests).all()
    for i in {0, (1 << lg) - 1, (1 << lg) // 3}:
        assert (t.get(i) == leaves[i]).all()
        pr = t.prove(i)
        assert (pr == o.prove(i)).all()
        assert oracle.merkle_verify(leaves[i], i, t.cap, pr)


@pytest.mark.parametrize("lg,leaf_len,cap_h,shards", [(10, 135, 4, 2), (12, 7, 4, 8), (9, 20, 3, 8), (8, 135, 0, 1), (11, 9, 4, 16)])
def test_merkle_tree_new_shards_concatenate_to_the_tree(qp, ctx, lg, leaf_len, cap_h, shards):
    """qp_merkle_tree_new_shard: the shards' digests blocks and cap entries, concatenated, are the reference's
    arrays (merkle_tree.rs:85-119 parallelises over the same cap subtrees), and a shard-local opening verifies
    against the whole cap at the global leaf index."""
    leaves = oracle.rand_felts((1 << lg, leaf_len), 300 + lg)
    want = oracle.MerkleTree(leaves, cap_h)
    per = (1 << lg) // shards
    caps, digs = [], []
    for s in range(shards):
        t = qp.MerkleTree(ctx, leaves[s * per:(s + 1) * per], cap_h, shard=s, n_shards=shards)
        caps.append(t.cap)
        digs.append(t.digests)
        assert oracle.merkle_verify(leaves[s * per + 3 % per], s * per + 3 % per, want.cap, t.prove(3 % per))
        t.free()
    assert (np.concatenate(caps) == want.cap).all()
    assert (np.concatenate(digs) == want.digests).all()
    with pytest.raises(qp.QpError):
        qp.MerkleTree(ctx, leaves[:per], cap_h, shard=0, n_shards=2 << cap_h)     # a shard smaller than a cap subtree


@pytest.mark.parametrize("lg,leaf_len,cap_h", [(4, 3, 0), (6, 135, 4), (10, 20, 2), (12, 9, 4)])
def test_merkle_tree_new_pipelined_host_rows(qp, ctx, lg, leaf_len, cap_h, monkeypatch):
    """MerkleTree::new on large host rows: the upload runs in 16 row slices and the leaves of a slice are hashed
    while the next slice crosses PCIe -- same digests, cap and openings as the oracle."""
    monkeypatch.setenv("QP_PIPELINE_MIN_BYTES", "1")
    leaves = oracle.rand_felts((1 << lg, leaf_len), 330 + lg)
    want = oracle.MerkleTree(leaves, cap_h)
    t = qp.MerkleTree(ctx, leaves, cap_h)
    assert (t.cap == want.cap).all() and (t.digests == want.digests).all()
    assert (t.get((1 << lg) - 1) == leaves[-1]).all()
    assert oracle.merkle_verify(leaves[7], 7, want.cap, t.prove(7))


def test_merkle_every_leaf_verifies(qp, ctx):
    """plonky2/src/hash/merkle_tree.rs:224-282: n = 2^8 x 7 elements, cap heights 0/1/8"""
    leaves = oracle.rand_felts((1 << 8, 7), 42)
    for cap_h in (0, 1, 8):
        t = qp.MerkleTree(ctx, leaves, cap_h)
        cap = t.cap
        for i in range(1 << 8):
            assert oracle.merkle_verify(leaves[i], i, cap, t.prove(i))


def test_merkle_errors(qp, ctx):
    """should_panic cases: cap too tall (merkle_tree.rs:164-170), non power of two (log2_strict)"""
    with pytest.raises(qp.QpError) as e:
        qp.MerkleTree(ctx, oracle.rand_felts((1 << 8, 7), 1), 9)
    assert e.value.code == 2
    with pytest.raises(qp.QpError) as e:
        qp.MerkleTree(ctx, oracle.rand_felts((12, 7), 1), 0)
    assert e.value.code == 3
    t = qp.MerkleTree(ctx, oracle.rand_felts((8, 3), 1), 0)
    with pytest.raises(qp.QpError):
        t.prove(8)


def test_leaf_hash_domain_separation(qp, ctx):
    """core/src/merkle_tree.rs:386-475 on the device: hash_leaf != two_to_one / length binding"""
    def hash_leaf_dev(x):
        return qp.MerkleTree(ctx, u64(x).reshape(1, -1), 0).cap[0]

    left, right = oracle.hash_no_pad(u64([1, 2])), oracle.hash_no_pad(u64([3, 4]))
    internal = qp.MerkleTree(ctx, np.stack([u64([1, 2]), u64([3, 4])]), 0)  # not the same thing; see below
    cat = np.concatenate([left, right])
    assert (hash_leaf_dev(cat) == oracle.hash_leaf(cat)).all()
    assert not (hash_leaf_dev(cat) == oracle.two_to_one(left, right)).all()
    assert not (hash_leaf_dev([1, 2, 3, 4, 5]) == hash_leaf_dev([1, 2, 3, 4, 5, 0])).all()
    # the device two_to_one (internal node) equals the oracle's compress of the leaf digests
    want = oracle.two_to_one(oracle.hash_leaf(u64([1, 2])), oracle.hash_leaf(u64([3, 4])))
    assert (internal.cap[0] == want).all()


# ---- PolynomialBatch -------------------------------------------------------------------------

CASES = [
    # lg_n, cols, rate, cap_h, salt
    (0, 1, 0, 0, False), (0, 2, 3, 2, False), (1, 3, 1, 0, False), (3, 3, 1, 0, True), (4, 5, 3, 2, False),
    (4, 2, 2, 4, True), (5, 9, 3, 4, False), (7, 135, 3, 4, False), (9, 20, 3, 4, False), (10, 16, 3, 0, False),
    (12, 143, 3, 4, False), (13, 7, 3, 4, True), (14, 3, 3, 4, False), (14, 2, 0, 3, False), (10, 4, 2, 12, False),
]


@pytest.mark.parametrize("lg_n,cols,rate,cap_h,salt", CASES)
def test_batch_from_values_parity(qp, ctx, lg_n, cols, rate, cap_h, salt):
    """from_values -> coefficients, LDE leaves (leaf order), digests, cap -- all bit-exact"""
    vals = oracle.rand_felts((cols, 1 << lg_n), 400 + lg_n)
    vals[0, 0] = np.uint64(2**64 - 1)  # non-canonical input is canonicalised on the way
    s = oracle.rand_felts((4, (1 << lg_n) << rate), 12) if salt else None
    want = oracle.PolynomialBatch.from_values(vals, rate, cap_h, salt=s)
    got = qp.PolynomialBatch.from_values(ctx, vals, rate, bool(salt), cap_h, salt=s)
    assert (got.polynomials == want.polynomials).all()
    assert (got.merkle_tree.cap == want.cap).all()
    assert (got.merkle_tree.digests == want.digests).all()
    assert (got.merkle_tree.leaves() == want.leaves).all()
    N = (1 << lg_n) << rate
    for idx in {0, N - 1, N // 3}:
        assert (got.get_lde_values(idx) == want.get_lde_values(idx)).all()
        pr = got.merkle_tree.prove(idx)
        assert oracle.merkle_verify(want.leaves[idx], idx, got.merkle_tree.cap, pr)
    assert (got.merkle_tree.get_many([N - 1, 0]) == want.leaves[[N - 1, 0]]).all()
    # from_coeffs on the coefficients gives the same commitment (oracle.rs:180-189)
    got2 = qp.PolynomialBatch.from_coeffs(ctx, want.polynomials, rate, bool(salt), cap_h, salt=s)
    assert (got2.merkle_tree.cap == want.cap).all()
    assert set(got.timing) == {"IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"}


def test_batch_pipelined_host_upload(qp, ctx):
    """host inputs >= 64 MiB take the column-group pipeline (upload overlapped with iNTT + LDE);
    2^16 x 135 is 70.8 MB.  Same commitment as the oracle, bit for bit."""
    vals = oracle.rand_felts((135, 1 << 16), 4242)
    want = oracle.PolynomialBatch.from_values(vals, 3, 4)
    got = qp.PolynomialBatch.from_values(ctx, vals, 3, False, 4)
    assert (got.merkle_tree.cap == want.cap).all()
    assert (got.polynomials == want.polynomials).all()
    assert (got.merkle_tree.digests == want.digests).all()
    idx = [0, 1, 77777, (1 << 19) - 1]
    assert (got.merkle_tree.get_many(idx) == want.leaves[idx]).all()


def test_batch_device_input(qp, ctx):
    """inputs already resident in HBM (torch CUDA tensor) give the same commitment"""
    import torch

    vals = oracle.rand_felts((6, 1 << 11), 77)
    want = oracle.PolynomialBatch.from_values(vals, 3, 4)
    d = torch.from_numpy(vals.view(np.int64)).cuda()
    got = qp.PolynomialBatch.from_values(ctx, d, 3, False, 4)
    assert (got.merkle_tree.cap == want.cap).all()


@pytest.mark.parametrize("shards", [2, 4, 8])
def test_batch_coset_sharding(qp, ctx, shards):
    """multi-GPU decomposition on one device: each shard holds block_count coset
    blocks = a contiguous leaf range; caps and digests concatenate to the whole tree."""
    lg_n, cols, rate, cap_h = 9, 5, 3, 4
    vals = oracle.rand_felts((cols, 1 << lg_n), 55)
    want = oracle.PolynomialBatch.from_values(vals, rate, cap_h)
    bc = (1 << rate) // shards
    caps, digs, leaves = [], [], []
    for s in range(shards):
        b = qp.PolynomialBatch.from_values(ctx, vals, rate, False, cap_h, block_first=s * bc, block_count=bc)
        caps.append(b.merkle_tree.cap)
        digs.append(b.merkle_tree.digests)
        leaves.append(b.merkle_tree.leaves())
    assert (np.concatenate(caps) == want.cap).all()
    assert (np.concatenate(digs) == want.digests).all()
    assert (np.concatenate(leaves) == want.leaves).all()


def test_batch_errors(qp, ctx):
    v = oracle.rand_felts((2, 16), 1)
    with pytest.raises(qp.QpError) as e:          # cap too tall
        qp.PolynomialBatch.from_values(ctx, v, 1, False, 6)
    assert e.value.code == 2
    with pytest.raises(qp.QpError) as e:          # not a power of two
        qp.PolynomialBatch.from_values(ctx, oracle.rand_felts((2, 12), 1), 1, False, 0)
    assert e.value.code == 3
    with pytest.raises(qp.QpError) as e:          # ragged columns (oracle.rs:277)
        qp.PolynomialBatch.from_values(ctx, [[1, 2, 3, 4], [1, 2]], 1, False, 0)
    assert e.value.code == 4
    with pytest.raises(qp.QpError) as e:          # blinding without salt
        qp.PolynomialBatch.from_values(ctx, v, 1, True, 0)
    assert e.value.code == 7
    with pytest.raises(qp.QpError):               # empty batch
        qp.PolynomialBatch.from_values(ctx, np.zeros((0, 16), dtype=np.uint64), 1, False, 0)


def test_batch_lde_is_low_degree_extension_large(qp, ctx):
    """2^16 x 4 at rate 3: leaf i = every column evaluated at g w_N^bitrev(i), checked by direct
    Horner evaluation of the returned coefficients (no oracle FFT involved)."""
    lg_n, rate = 16, 3
    vals = oracle.rand_felts((4, 1 << lg_n), 3)
    b = qp.PolynomialBatch.from_values(ctx, vals, rate, False, 4)
    co = [[int(x) for x in col] for col in b.polynomials]
    g, w = pyref.GENERATOR, pyref.primitive_root_of_unity(lg_n + rate)
    for idx in (1, 12345, (1 << (lg_n + rate)) - 1):
        x = g * pow(w, idx, P) % P
        row = b.get_lde_values(idx)
        for c in range(4):
            acc = 0
            for ci in reversed(co[c]):
                acc = (acc * x + ci) % P
            assert int(row[c]) == acc


# ---- FRI commit phase --------------------------------------------------------------------------

def _lowdeg_ext(deg_bits, rate, seed):
    n = 1 << (deg_bits + rate)
    co = np.zeros((n, 2), dtype=np.uint64)
    co[: 1 << deg_bits] = oracle.rand_felts((1 << deg_bits, 2), seed)
    g = oracle.lib().orc_gl_coset_shift()
    va = np.stack([oracle.coset_fft(co[:, 0], g), oracle.coset_fft(co[:, 1], g)], axis=1)
    return co, va


@pytest.mark.parametrize("deg_bits,rate,cap_h,arities", [
    (4, 1, 0, [1, 1]), (5, 2, 1, [2, 1]), (6, 1, 0, [3, 2]), (9, 3, 4, [4, 4]), (12, 3, 4, [4, 4]),
    (14, 3, 4, [4, 4, 4]), (8, 3, 4, []), (10, 3, 2, [4]),
])
def test_fri_committed_trees_parity(qp, ctx, deg_bits, rate, cap_h, arities):
    """fri_committed_trees: caps, betas (through the transcript), final poly, trees -- vs oracle"""
    co, va = _lowdeg_ext(deg_bits, rate, 500 + deg_bits)
    ca, cb = qp.Challenger(), oracle.Challenger()
    ca.observe_elements([1, 2, 3])
    cb.observe(u64([1, 2, 3]))
    r = qp.fri_committed_trees(ctx, co, va, ca, rate, cap_h, arities)
    o = oracle.fri_committed_trees(co, va, rate, cap_h, arities, cb, keep_trees=True)
    assert (r.caps == o["caps"]).all()
    assert (r.final_poly == o["final_poly"]).all()
    assert ca.get_challenge() == cb.get_challenge()      # transcripts stayed in sync
    for k in range(len(arities)):
        assert (r.tree_digests(k) == o["digests"][k]).all()
        nl = o["leaves"][k].shape[0]
        for i in {0, nl - 1, nl // 3}:
            assert (r.tree_get(k, i) == o["leaves"][k][i]).all()
            pr = r.tree_prove(k, i)
            assert oracle.merkle_verify(o["leaves"][k][i], i, r.caps[k], pr)


def test_fri_fold_consistency(qp, ctx):
    """FRI soundness identity, oracle-free: the final polynomial evaluated at x^(arity^rounds)
    equals the iterated fold of the committed values.  Here: degree after folding drops so the
    final poly has exactly 2^deg_bits >> sum(arities) coefficients and the rest are zero."""
    co, va = _lowdeg_ext(10, 3, 9)
    ch = qp.Challenger()
    r = qp.fri_committed_trees(ctx, co, va, ch, 3, 4, [4, 4])
    assert r.final_poly.shape == (4, 2)


def test_fri_proof_of_work(qp, ctx):
    """fri_proof_of_work: smallest witness; transcript equal to the oracle's afterwards"""
    for bits in (4, 10, 16):
        ca, cb = qp.Challenger(), oracle.Challenger()
        ca.observe_elements([5, 6, 7, bits])
        cb.observe(u64([5, 6, 7, bits]))
        wa = qp.fri_proof_of_work(ctx, ca, bits)
        wb = oracle.fri_proof_of_work(cb, bits)
        assert wa == wb
        assert ca.get_challenge() == cb.get_challenge()


def test_challenger_matches_oracle(qp):
    a, b = qp.Challenger(), oracle.Challenger()
    rng = np.random.default_rng(3)
    for i in range(1, 12):
        xs = oracle.rand_felts(int(rng.integers(0, 20)), 50 + i)
        a.observe_elements(xs)
        b.observe(xs)
        for _ in range(i):
            assert a.get_challenge() == b.get_challenge()
    assert qp.fri_reduction_arity_bits(14, 3, 4) == [4, 4, 4]
    assert qp.fri_reduction_arity_bits(20, 3, 4) == oracle.fri_reduction_arity_bits(20, 3, 4)


# ---- BASELINE.json full-size configurations: size-independent properties -------------------------

def test_full_size_commit_properties(qp, ctx):
    """2^20 rows x 135 columns, rate 3, cap 4 (BASELINE.json configs[2]): size-independent
    properties -- (1) random leaves verify against the cap with the ORACLE's verifier, (2) rows are
    the Horner evaluation of the returned coefficients, (3) the forward transform of the coefficients
    returns the input.  The bit-for-bit comparison with the oracle at this size is the next test."""
    import torch

    lg_n, cols, rate, cap_h = 20, 135, 3, 4
    gen = torch.Generator(device="cuda").manual_seed(42)
    d = torch.randint(0, 2**62, (cols, 1 << lg_n), dtype=torch.int64, device="cuda", generator=gen)
    b = qp.PolynomialBatch.from_values(ctx, d, rate, False, cap_h)
    cap = b.merkle_tree.cap
    assert cap.shape == (16, 4)
    N = 1 << (lg_n + rate)
    rng = np.random.default_rng(1)
    idxs = [0, N - 1] + [int(x) for x in rng.integers(0, N, 6)]
    rows = b.merkle_tree.get_many(idxs)
    for i, row in zip(idxs, rows):
        assert oracle.merkle_verify(row, i, cap, b.merkle_tree.prove(i))
    # Horner check of two columns at two points
    coeffs = b.polynomials
    g, w = pyref.GENERATOR, pyref.primitive_root_of_unity(lg_n + rate)
    for i, row in list(zip(idxs, rows))[:2]:
        nat = int(format(i, "0%db" % (lg_n + rate))[::-1], 2)
        x = g * pow(w, nat, P) % P
        for c in (0, 134):
            acc = 0
            for ci in coeffs[c][::-1].tolist():
                acc = (acc * x + ci) % P
            assert int(row[c]) == acc
    # iNTT really inverts: forward transform of the coefficients returns the input values
    back = ctx.coset_fft(coeffs[:2], shift=1)
    assert (back == d[:2].cpu().numpy().view(np.uint64)).all()
    print("full-size timing (ms):", b.timing)


def test_full_size_commit_equals_oracle(qp, ctx):
    """2^20 rows x 135 columns, rate 3, cap 4 (BASELINE.json configs[2]) against the oracle AT FULL SIZE
    (about 21 GB of host memory and half a minute of host time on the GPU box): coefficients, the whole
    digest array, the cap and sampled LDE rows, bit for bit, on bench.py's own witness."""
    import torch

    import bench

    lg_n, cols, rate, cap_h = 20, 135, 3, 4
    d = bench.synth_columns_torch(0, cols, 1 << lg_n, "cuda")
    host = d.cpu().numpy().view(np.uint64)
    assert (host == bench.synth_columns_numpy(0, cols, 1 << lg_n)).all()      # both arms of bench.py see one matrix
    assert int(host.max()) < P
    b = qp.PolynomialBatch.from_values(ctx, d, rate, False, cap_h)
    want = oracle.PolynomialBatch.from_values(host, rate, cap_h)
    assert (b.merkle_tree.cap == want.cap).all(), "cap differs at full size"
    assert (b.polynomials == want.polynomials).all(), "coefficients differ at full size"
    assert (b.merkle_tree.digests == want.digests).all(), "digests differ at full size"
    N = 1 << (lg_n + rate)
    rng = np.random.default_rng(7)
    idxs = [0, 1, N // 2, N - 1] + [int(x) for x in rng.integers(0, N, 60)]
    rows = b.merkle_tree.get_many(idxs)
    for i, row in zip(idxs, rows):
        assert (row == want.leaves[i]).all(), "LDE row %d differs at full size" % i
    b.free()


def test_config_c_row_count_commit_properties(qp):
    """2^23 rows (BASELINE.json configs[4]'s row count; three-pass transforms of size 2^23, LDE of
    2^26 points) x 3 columns: size-independent properties, as above -- leaves verify against the cap
    with the oracle's verifier, a row is the Horner evaluation of the returned coefficients, the
    forward transform of the coefficients returns the input, and a coset shard of the same
    commitment (what one of 8 GPUs computes) reproduces its two cap entries."""
    import torch

    lg_n, cols, rate, cap_h = 23, 3, 3, 4
    ctx = qp.Context(0, max_lde_log=lg_n + rate)
    try:
        gen = torch.Generator(device="cuda").manual_seed(23)
        d = torch.randint(0, 2**62, (cols, 1 << lg_n), dtype=torch.int64, device="cuda", generator=gen)
        b = qp.PolynomialBatch.from_values(ctx, d, rate, False, cap_h)
        cap = b.merkle_tree.cap
        N = 1 << (lg_n + rate)
        idxs = [0, N - 1, 123456789 % N, (N // 2) + 77]
        rows = b.merkle_tree.get_many(idxs)
        for i, row in zip(idxs, rows):
            assert oracle.merkle_verify(row, i, cap, b.merkle_tree.prove(i))
        coeffs = b.polynomials
        g, w = pyref.GENERATOR, pyref.primitive_root_of_unity(lg_n + rate)
        i, row = idxs[2], rows[2]
        nat = int(format(i, "0%db" % (lg_n + rate))[::-1], 2)
        x = g * pow(w, nat, P) % P
        acc = 0
        for ci in coeffs[1][::-1].tolist():
            acc = (acc * x + ci) % P
        assert int(row[1]) == acc
        back = ctx.coset_fft(coeffs[:1], shift=1)
        assert (back == d[:1].cpu().numpy().view(np.uint64)).all()
        shard = qp.PolynomialBatch.from_coeffs(ctx, coeffs, rate, False, cap_h, block_first=5, block_count=1)
        assert (shard.merkle_tree.cap == cap[10:12]).all()
        shard.free()
        b.free()
    finally:
        ctx.close()


# ---- opening side of prove_openings -----------------------------------------

def test_eval_polys_at_ext_point(qp, ctx):
    """OpeningSet::new: every committed polynomial evaluated at zeta (plonky2/src/plonk/proof.rs:289-327)"""
    for lg_n, cols in ((0, 2), (5, 7), (12, 20), (16, 3)):
        vals = oracle.rand_felts((cols, 1 << lg_n), 700 + lg_n)
        b = qp.PolynomialBatch.from_values(ctx, vals, 1, False, 0)
        co = b.polynomials
        for pt in ((5, 0), (1234567891011, 987654321), (0, 0), (P - 1, P - 2)):
            got = b.eval_polys(pt)
            for c in range(cols):
                assert (got[c] == oracle.eval_poly_ext(co[c], pt)).all()


@pytest.mark.parametrize("lg_n,rate,zero_point", [(1, 1, False), (5, 2, False), (6, 3, True), (11, 3, False), (13, 3, False)])
def test_fri_from_openings_parity(qp, ctx, lg_n, rate, zero_point):
    """reduce_openings_to_unmasked_final_poly + final coset FFT + commit phase, all on the device,
    against the oracle's reduce_openings -> coset_fft -> fri_committed_trees."""
    n = 1 << lg_n
    cap_h = min(2, lg_n + rate)
    rng = np.random.default_rng(lg_n)
    o1 = qp.PolynomialBatch.from_values(ctx, oracle.rand_felts((5, n), 1), rate, False, cap_h)
    o2 = qp.PolynomialBatch.from_values(ctx, oracle.rand_felts((3, n), 2), rate, False, cap_h)
    oracles = [o1, o2]
    coeffs = [o1.polynomials, o2.polynomials]
    dev_batches, ref_batches = [], []
    for b in range(3):
        k = int(rng.integers(1, 9))
        pt = tuple(int(x) for x in oracle.rand_felts(2, 40 + b))
        if zero_point and b == 1:
            pt = (0, 0)
        sh = tuple(int(x) for x in oracle.rand_felts(2, 50 + b))
        dt, rt = [], []
        for _ in range(k):
            oi = int(rng.integers(0, 2))
            pi = int(rng.integers(0, coeffs[oi].shape[0]))
            w = tuple(int(x) for x in oracle.rand_felts(2, int(rng.integers(0, 1 << 30))))
            dt.append((oracles[oi], pi, w))
            rt.append((coeffs[oi][pi], w))
        dev_batches.append(dict(point=pt, shift=sh, terms=dt))
        ref_batches.append(dict(point=pt, shift=sh, terms=rt))
    f = qp.fri_from_openings(ctx, dev_batches, lg_n, rate, cap_h)
    want_final = oracle.reduce_openings(ref_batches, lg_n)
    assert (qp.fri_initial_coeffs(f, lg_n) == want_final).all()
    # commit phase on top of it
    N = n << rate
    co = np.zeros((N, 2), dtype=np.uint64)
    co[:n] = want_final
    g = oracle.lib().orc_gl_coset_shift()
    va = np.stack([oracle.coset_fft(co[:, 0], g), oracle.coset_fft(co[:, 1], g)], axis=1)
    arities = [a for a in ([2, 1] if lg_n >= 5 else [1]) if a <= lg_n]
    while sum(arities) > lg_n or (lg_n + rate - sum(arities[:1])) < cap_h:
        arities.pop()
    ca, cb = qp.Challenger(), oracle.Challenger()
    qp.fri_commit_phase(f, ca, rate, arities)
    o = oracle.fri_committed_trees(co, va, rate, cap_h, arities, cb)
    assert (f.caps == o["caps"]).all()
    assert (f.final_poly == o["final_poly"]).all()
    assert ca.get_challenge() == cb.get_challenge()


@pytest.mark.parametrize("lg_n,rate,cap_h,pow_bits,queries", [(5, 2, 1, 4, 3), (9, 3, 4, 8, 5), (12, 3, 4, 16, 28)])
def test_fri_proof_bytes_parity(qp, ctx, lg_n, rate, cap_h, pow_bits, queries):
    """Whole FRI proof (commit phase, final poly, PoW, query openings of the initial trees and of
    every commit-phase tree) serialised like write_fri_proof -- byte for byte against the oracle.
    This is prove_openings from the alpha-reduction on (plonky2/src/fri/oracle.rs:320-358)."""
    n = 1 << lg_n
    v1, v2 = oracle.rand_felts((6, n), 11), oracle.rand_felts((3, n), 12)
    salt = oracle.rand_felts((4, n << rate), 13)
    d1 = qp.PolynomialBatch.from_values(ctx, v1, rate, False, cap_h)
    d2 = qp.PolynomialBatch.from_values(ctx, v2, rate, True, cap_h, salt=salt)
    o1 = oracle.PolynomialBatch.from_values(v1, rate, cap_h)
    o2 = oracle.PolynomialBatch.from_values(v2, rate, cap_h, salt=salt)
    rng = np.random.default_rng(7)
    dev_b, ref_b = [], []
    for b in range(2):
        pt = tuple(int(x) for x in oracle.rand_felts(2, 60 + b))
        sh = tuple(int(x) for x in oracle.rand_felts(2, 70 + b))
        dt, rt = [], []
        for (dd, oo, cols) in ((d1, o1, 6), (d2, o2, 3)):
            for pi in range(cols):
                w = tuple(int(x) for x in oracle.rand_felts(2, int(rng.integers(0, 1 << 30))))
                dt.append((dd, pi, w))
                rt.append((oo.polynomials[pi], w))
        dev_b.append(dict(point=pt, shift=sh, terms=dt))
        ref_b.append(dict(point=pt, shift=sh, terms=rt))
    arities = oracle.fri_reduction_arity_bits(lg_n, rate, cap_h, arity_bits=min(4, max(1, lg_n // 3)), final_poly_bits=2)
    ca, cb = qp.Challenger(), oracle.Challenger()
    ca.observe_elements([9, 8, 7])
    cb.observe(u64([9, 8, 7]))
    f = qp.fri_from_openings(ctx, dev_b, lg_n, rate, cap_h)
    got = qp.fri_proof(ctx, [d1, d2], f, ca, rate, arities, pow_bits, queries)
    fin = oracle.reduce_openings(ref_b, lg_n)
    N = n << rate
    co = np.zeros((N, 2), dtype=np.uint64)
    co[:n] = fin
    g = oracle.lib().orc_gl_coset_shift()
    va = np.stack([oracle.coset_fft(co[:, 0], g), oracle.coset_fft(co[:, 1], g)], axis=1)
    want = oracle.fri_proof_bytes([o1, o2], co, va, cb, rate, cap_h, arities, pow_bits, queries)
    assert len(got) == len(want)
    assert got == want
    assert ca.get_challenge() == cb.get_challenge()


@pytest.mark.parametrize("cols,lg_n,blinding", [(5, 8, False), (16, 9, False), (17, 9, True), (40, 10, False),
                                                (135, 10, False), (143, 11, True), (8, 7, True)])
def test_from_values_pipelined_upload_with_partial_leaf_hashing(qp, ctx, cols, lg_n, blinding, monkeypatch):
    """Host input: the upload runs in 16-column groups and the leaf sponge absorbs each group as soon
    as its LDE exists (merkle::leaf_hash_kernel in pieces).  Same commitment as the oracle's, for
    leaf lengths on and off the 8-element chunk boundary, with and without salt."""
    monkeypatch.setenv("QP_PIPELINE_MIN_BYTES", "1")
    n = 1 << lg_n
    vals = oracle.rand_felts((cols, n), 900 + cols)
    salt = oracle.rand_felts((4, n << 3), 901 + cols) if blinding else None
    got = qp.PolynomialBatch.from_values(ctx, vals, 3, blinding, 4, salt=salt)
    want = oracle.PolynomialBatch.from_values(vals, 3, 4, salt=salt)
    assert (got.merkle_tree.cap == want.cap).all()
    assert (got.merkle_tree.digests == want.digests).all()
    assert (got.polynomials == want.polynomials).all()
    assert (got.merkle_tree.leaves() == want.leaves).all()
    assert got.kernel_ms["leaf_hash"] > 0


@pytest.mark.parametrize("cols,lg_n,blinding,pipelined", [(1, 6, False, False), (9, 8, True, False), (5, 8, False, True),
                                                          (33, 9, True, True), (135, 10, False, True), (50, 11, False, True)])
def test_from_values_cols_pageable_columns(qp, ctx, cols, lg_n, blinding, pipelined, monkeypatch):
    """qp_batch_from_values_cols: one separately allocated pageable vector per column (the reference's
    Vec<PolynomialValues>, oracle.rs:168-175), small (gathered) and large (staged through the pinned ring
    in 16-column groups, more groups than ring slots) -- the oracle's commitment either way."""
    if pipelined:
        monkeypatch.setenv("QP_PIPELINE_MIN_BYTES", "1")
    n = 1 << lg_n
    vals = oracle.rand_felts((cols, n), 950 + cols)
    columns = [np.array(vals[c], copy=True) for c in range(cols)]      # separate heap allocations
    salt = oracle.rand_felts((4, n << 3), 951 + cols) if blinding else None
    got = qp.PolynomialBatch.from_values_cols(ctx, columns, 3, blinding, 4, salt=salt)
    want = oracle.PolynomialBatch.from_values(vals, 3, 4, salt=salt)
    assert (got.merkle_tree.cap == want.cap).all()
    assert (got.merkle_tree.digests == want.digests).all()
    assert (got.polynomials == want.polynomials).all()
    assert (got.merkle_tree.leaves() == want.leaves).all()
    with pytest.raises(qp.QpError) as e:
        qp.PolynomialBatch.from_values_cols(ctx, columns[:1] + [columns[0][: n // 2]], 3, False, 4)
    assert e.value.code == 4   # "Polynomial degrees inconsistent" (oracle.rs:277)


@pytest.mark.parametrize("cols,lg_n,blinding,first,count", [(7, 8, False, 0, 8), (19, 10, True, 0, 8), (35, 9, False, 4, 4),
                                                            (5, 6, False, 6, 2)])
def test_batch_in_pieces_equals_from_coeffs(qp, ctx, cols, lg_n, blinding, first, count):
    """qp_batch_begin / put_coeffs / end (columns arriving in arbitrary pieces, as from the chunked
    all-gather of the multi-GPU path) gives the batch from_coeffs gives."""
    n = 1 << lg_n
    co = oracle.rand_felts((cols, n), 800 + cols)
    salt = oracle.rand_felts((4, n << 3), 801 + cols) if blinding else None
    want = qp.PolynomialBatch.from_coeffs(ctx, co, 3, blinding, 4, salt=salt, block_first=first, block_count=count)
    b = qp.PolynomialBatch.begin(ctx, cols, lg_n, 3, blinding, 4, block_first=first, block_count=count)
    order = list(range(0, cols, 3))
    for c0 in order[1::2] + order[0::2]:          # out of order, ragged last piece
        b.put_coeffs(co[c0:c0 + 3], c0)
    b.end(salt)
    assert (b.merkle_tree.cap == want.merkle_tree.cap).all()
    assert (b.merkle_tree.digests == want.merkle_tree.digests).all()
    assert (b.polynomials == want.polynomials).all()
    assert (b.merkle_tree.leaves() == want.merkle_tree.leaves()).all()
    with pytest.raises(qp.QpError):
        b.put_coeffs(co[:2], cols - 1)            # past the last column


# ---- BatchMerkleTree (the oracle of batch FRI) ---------------------------------

@pytest.mark.parametrize("cols,lg_n,blinding,first,count,piece,in_order", [
    (7, 8, False, 0, 8, 8, True), (19, 10, True, 0, 8, 8, True), (35, 9, False, 4, 4, 5, True),
    (135, 8, False, 0, 8, 8, True), (33, 7, False, 6, 2, 8, False), (16, 6, True, 0, 8, 3, False)])
def test_batch_extend_columns_in_place_with_sponge_absorption(qp, ctx, cols, lg_n, blinding, first, count, piece, in_order):
    """qp_batch_coeffs_slot / qp_batch_extend_columns: coefficients written in place (as a collective's
    receive buffer would be), extended piece by piece with the leaf sponges advancing over the completed
    column prefix -- the batch from_coeffs gives, for in-order and out-of-order arrival, salted or not,
    whole or a coset shard."""
    import torch

    n = 1 << lg_n
    co = oracle.rand_felts((cols, n), 820 + cols)
    salt = oracle.rand_felts((4, n << 3), 821 + cols) if blinding else None
    want = qp.PolynomialBatch.from_coeffs(ctx, co, 3, blinding, 4, salt=salt, block_first=first, block_count=count)
    b = qp.PolynomialBatch.begin(ctx, cols, lg_n, 3, blinding, 4, block_first=first, block_count=count)
    starts = list(range(0, cols, piece))
    if not in_order:
        starts = starts[1::2] + starts[0::2]
    d_co = torch.from_numpy(co.view(np.int64)).cuda()
    for c0 in starts:
        k = min(piece, cols - c0)
        b.coeffs_slot(c0, k).copy_(d_co[c0:c0 + k])
        torch.cuda.synchronize()
        b.extend_columns(c0, k, absorb=True)
    b.end(salt=salt)
    assert (b.merkle_tree.cap == want.merkle_tree.cap).all()
    assert (b.merkle_tree.digests == want.merkle_tree.digests).all()
    assert (b.polynomials == want.polynomials).all()
    assert (b.merkle_tree.leaves() == want.merkle_tree.leaves()).all()
    with pytest.raises(qp.QpError):
        b.extend_columns(cols, 1)
    b2 = qp.PolynomialBatch.begin(ctx, cols, lg_n, 3, False, 4)
    b2.extend_columns(0, 1)
    if cols > 1:
        with pytest.raises(qp.QpError):
            b2.end()          # columns missing
    b2.free()


@pytest.mark.parametrize("shapes,cap_h", [([(4, 2)], 0), ([(4, 2), (2, 3)], 0), ([(64, 7), (16, 3), (8, 20)], 2),
                                           ([(4096, 135), (512, 9), (32, 1)], 4), ([(256, 0), (16, 4)], 4),
                                           ([(1 << 14, 20), (1 << 13, 16)], 4)])
def test_batch_merkle_tree_parity(qp, ctx, shapes, cap_h):
    """BatchMerkleTree::new / open_batch / values against the big-integer restatement
    (plonky2/src/hash/batch_merkle_tree.rs:40-164), every digest; openings through the restated
    verify_batch_merkle_proof_to_cap."""
    rng = np.random.default_rng(len(shapes) * 100 + cap_h)
    mats = [rng.integers(0, P, size=s, dtype=np.uint64) for s in shapes]
    t = qp.BatchMerkleTree(ctx, mats, cap_h)
    assert t.leaf_heights == [s[0].bit_length() - 1 for s in shapes]
    lists = [m.tolist() for m in mats]
    small = shapes[0][0] <= 4096
    if small:
        digests, cap, heights = pyref.batch_merkle_tree(lists, cap_h)
        assert t.digests.tolist() == [list(d) for d in digests]
        assert t.cap.tolist() == [list(c) for c in cap]
    cap = t.cap.tolist()
    n = shapes[0][0]
    for i in sorted({0, 1, n // 2, n - 1, int(rng.integers(0, n))}):
        rows = [[int(x) for x in r] for r in t.values(i)]
        assert rows == [m[i >> (t.leaf_heights[0] - h)] for m, h in zip(lists, t.leaf_heights)]
        proof = [list(map(int, s)) for s in t.open_batch(i)]
        assert pyref.batch_merkle_verify(rows, t.leaf_heights, i, cap, proof)
        if rows[-1]:
            rows[-1][0] ^= 1
            assert not pyref.batch_merkle_verify(rows, t.leaf_heights, i, cap, proof)
    t.free()


def test_batch_merkle_tree_errors(qp, ctx):
    """the reference's asserts, batch_merkle_tree.rs:41-55"""
    z = lambda h, w: np.zeros((h, w), dtype=np.uint64)
    with pytest.raises(qp.QpError):
        qp.BatchMerkleTree(ctx, [z(8, 2), z(8, 2)], 0)    # duplicate heights
    with pytest.raises(qp.QpError):
        qp.BatchMerkleTree(ctx, [z(4, 2), z(8, 2)], 0)    # not sorted tallest first
    with pytest.raises(qp.QpError):
        qp.BatchMerkleTree(ctx, [z(6, 2)], 0)             # not a power of two
    with pytest.raises(qp.QpError):
        qp.BatchMerkleTree(ctx, [z(16, 2), z(4, 2)], 3)   # cap_height > log2(last height)


# ---- byte form of a PolynomialBatch (CircuitData (de)serialization) -----------------

@pytest.mark.parametrize("lg_n,cols,rate,cap_h,salted", [(5, 3, 3, 4, False), (9, 7, 3, 4, False), (6, 4, 1, 0, True),
                                                        (12, 84, 3, 4, False)])
def test_polynomial_batch_bytes_round_trip(qp, ctx, lg_n, cols, rate, cap_h, salted):
    """write_polynomial_batch / read_polynomial_batch (serialization/mod.rs:1803-1822, 758-784): the
    device batch serialises to the oracle's bytes; the bytes deserialise to a batch with the same cap,
    digests, leaves and polynomials; truncated or tampered bytes are refused."""
    vals = oracle.rand_felts((cols, 1 << lg_n), 1200 + lg_n)
    salt = oracle.rand_felts((4, 1 << (lg_n + rate)), 77) if salted else None
    b = qp.PolynomialBatch.from_values(ctx, vals, rate, salted, cap_h, salt=salt)
    want = oracle.PolynomialBatch.from_values(vals, rate, cap_h, salt=salt)
    data = b.to_bytes()
    assert data == oracle.serialize_polynomial_batch(want, rate, salted)
    back, used = qp.PolynomialBatch.from_bytes(ctx, data + b"trailing")
    assert used == len(data) and back.blinding == salted and back.n_cols == cols and back.degree_log == lg_n
    assert (back.merkle_tree.cap == want.cap).all() and (back.merkle_tree.digests == want.digests).all()
    assert (back.merkle_tree.leaves() == want.leaves).all() and (back.polynomials == want.polynomials).all()
    assert back.to_bytes() == data
    with pytest.raises(qp.QpError):
        qp.PolynomialBatch.from_bytes(ctx, data[:-9])
    bad = bytearray(data)
    bad[8 + 8 + 3] ^= 1          # a coefficient of the first polynomial: the stored cap is no longer theirs
    with pytest.raises(qp.QpError):
        qp.PolynomialBatch.from_bytes(ctx, bytes(bad))
    bad = bytearray(data)
    bad[0] ^= 1                  # the polynomial count
    with pytest.raises(qp.QpError):
        qp.PolynomialBatch.from_bytes(ctx, bytes(bad))


# ---- BatchFriOracle -----------------------------------------------------------

@pytest.mark.parametrize("lens,rate,cap_h,from_values", [([64, 64, 64, 16, 16, 8], 3, 2, False), ([32, 32, 4], 1, 0, True),
                                                        ([128], 3, 4, True), ([256, 256, 64, 64, 64, 64, 64], 2, 4, True),
                                                        ([1 << 12] * 20 + [1 << 10] * 9 + [1 << 7] * 3, 3, 4, False)])
def test_batch_fri_oracle_parity(qp, ctx, lens, rate, cap_h, from_values):
    """BatchFriOracle::from_values / from_coeffs (plonky2/src/batch_fri/oracle.rs:78-160): per-degree
    LDEs in leaf order under one BatchMerkleTree -- cap, every digest, coefficients, rows and
    openings against the big-integer restatement (small cases) / its verifier (all cases)."""
    rng = np.random.default_rng(sum(lens) + rate)
    polys = [rng.integers(0, P, size=n, dtype=np.uint64) for n in lens]
    o = (qp.BatchFriOracle.from_values if from_values else qp.BatchFriOracle.from_coeffs)(ctx, polys, rate, False, cap_h)
    assert o.degree_bits == sorted({n.bit_length() - 1 for n in lens}, reverse=True)
    coeffs = [pyref.ifft(p.tolist()) for p in polys] if from_values and max(lens) <= 256 else None
    if not from_values:
        coeffs = [p.tolist() for p in polys]
    got_polys = o.polynomials
    if coeffs is not None:
        assert [c.tolist() for c in got_polys] == coeffs
    heights = [d + rate for d in o.degree_bits]
    cap = o.cap.tolist()
    if max(lens) <= 256:
        mats, digests, want_cap, want_bits = pyref.batch_fri_from_coeffs(coeffs, rate, cap_h)
        assert o.digests.tolist() == [list(d) for d in digests] and cap == [list(c) for c in want_cap]
        assert want_bits == o.degree_bits
    N = lens[0] << rate
    g, w = pyref.GENERATOR, pyref.primitive_root_of_unity(heights[0])
    for i in sorted({0, N - 1, int(rng.integers(0, N))}):
        rows = [[int(x) for x in r] for r in o.values(i)]
        proof = [list(map(int, s)) for s in o.open_batch(i)]
        assert pyref.batch_merkle_verify(rows, heights, i, cap, proof)
        # a row of the tallest group is that group's polynomials at g w^bitrev(i) (get_lde_values)
        x = g * pow(w, pyref.bitrev(i, heights[0]), P) % P
        acc = 0
        for ci in reversed(got_polys[0].tolist()):
            acc = (acc * x + ci) % P
        assert rows[0][0] == acc
    o.free()


def test_batch_fri_oracle_errors(qp, ctx):
    z = lambda n: np.zeros(n, dtype=np.uint64)
    with pytest.raises(qp.QpError):
        qp.BatchFriOracle.from_coeffs(ctx, [z(8), z(16)], 1, False, 0)     # degrees must not increase (oracle.rs:118)
    with pytest.raises(qp.QpError):
        qp.BatchFriOracle.from_coeffs(ctx, [z(16), z(4)], 1, False, 4)     # cap above the shortest matrix
    with pytest.raises(qp.QpError):
        qp.BatchFriOracle.from_coeffs(ctx, [z(16)], 1, True, 0)            # blinding: not implemented
    with pytest.raises(qp.QpError):
        qp.BatchFriOracle.from_coeffs(ctx, [z(12)], 1, False, 0)           # not a power of two


def test_batch_open_many(qp, ctx):
    """qp_batch_open_many = get_leaves + prove_many in one round trip: same rows and paths, and each pair
    verifies against the cap (fri_prover_query_rounds, plonky2/src/fri/prover.rs:246-253)."""
    vals = oracle.rand_felts((7, 1 << 8), 31)
    b = qp.PolynomialBatch.from_values(ctx, vals, 3, False, 2)
    idx = [0, 5, 2047, 1024, 5, 77]
    rows, paths = b.merkle_tree.open_many(idx)
    assert (rows == b.merkle_tree.get_many(idx)).all()
    cap = b.merkle_tree.cap
    for k, i in enumerate(idx):
        assert (paths[k] == b.merkle_tree.prove(i)).all()
        assert oracle.merkle_verify(rows[k], i, cap, paths[k])
    with pytest.raises(qp.QpError):
        b.merkle_tree.open_many([1 << 11])
    b.free()


# ---- multi-device context (one process, a list of GPUs) ------------------------------------------

def _check_multi(qp, devices, cols, lg_n, blinding, monkeypatch=None):
    n = 1 << lg_n
    vals = oracle.rand_felts((cols, n), 970 + cols)
    columns = [np.array(vals[c], copy=True) for c in range(cols)]
    salt = oracle.rand_felts((4, n << 3), 971 + cols) if blinding else None
    want = oracle.PolynomialBatch.from_values(vals, 3, 4, salt=salt)
    m = qp.MultiContext(devices, max_lde_log=lg_n + 3)
    try:
        mb = qp.MultiBatch.from_values_cols(m, columns, 3, blinding, 4, salt=salt)
        assert (mb.cap == want.cap).all(), "multi-device cap differs from the oracle's"
        D = len(devices)
        per_dig = want.digests.shape[0] // D
        per_leaf = want.leaves.shape[0] // D
        for i, sh in enumerate(mb.shards):
            assert (sh.merkle_tree.digests == want.digests[i * per_dig:(i + 1) * per_dig]).all()
            assert (sh.merkle_tree.leaves() == want.leaves[i * per_leaf:(i + 1) * per_leaf]).all()
            assert (sh.polynomials == want.polynomials).all()
            # a local opening verifies against the GLOBAL cap at the global index
            li = 5 % per_leaf
            row = sh.merkle_tree.get_many([li])[0]
            assert oracle.merkle_verify(row, i * per_leaf + li, mb.cap, sh.merkle_tree.prove(li))
        mb.free()
    finally:
        m.close()


@pytest.mark.parametrize("cols,lg_n,blinding", [(5, 6, False), (19, 9, True), (135, 8, False)])
def test_multi_context_single_device(qp, cols, lg_n, blinding):
    """qp_mctx_* with one device: the producer / consumer threads, piece events and in-place slots of the
    multi-device commit on a single GPU -- the oracle's commitment."""
    _check_multi(qp, [0], cols, lg_n, blinding)


@pytest.mark.parametrize("cols,lg_n,blinding", [(5, 6, False), (19, 9, True), (135, 10, False), (40, 12, False)])
def test_multi_context_all_devices(qp, cols, lg_n, blinding):
    """One process driving every GPU of the box (2, 4 or 8): coset = cap-subtree shards, coefficient pieces as
    peer copies in column order, the cap and every shard's digests / rows / openings against the oracle."""
    import torch

    D = torch.cuda.device_count()
    D = 8 if D >= 8 else 4 if D >= 4 else 2 if D >= 2 else 1
    if D < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    _check_multi(qp, list(range(D)), cols, lg_n, blinding)
