"""CPU tests of the oracle (oracle/plonky2_oracle.c) against the reference's own pins and the
independent big-integer restatement (oracle/pyref.py).  Mirrors the reference's test strategy
(SURVEY.md section 4): known-answer vectors, algebraic self-consistency, Merkle round trips."""
import numpy as np
import pytest

import oracle
from oracle import pyref

P = oracle.P


def u64(x):
    return np.array(x, dtype=np.uint64)


# ---- known-answer vectors -----------------------------------------------------------------

def test_poseidon_kat(golden):
    """core/src/poseidon_goldilocks.rs:455-490 (test_vectors) + poseidon.rs:743-756 (consistency)"""
    for v in golden("poseidon_kat.json"):
        i = [int(x) for x in v["input"]]
        o = u64([int(x) for x in v["output"]])
        assert (oracle.poseidon(i) == o).all()
        assert (oracle.poseidon_naive(i) == o).all()
        assert pyref.poseidon(i) == [int(x) for x in o]


def test_poseidon_fast_equals_naive_random():
    st = oracle.rand_felts((64, 12), 7)
    for s in st:
        assert (oracle.poseidon(s) == oracle.poseidon_naive(s)).all()
    # non-canonical inputs (raw u64 >= p) are legal internal values
    s = u64([2**64 - 1] * 12)
    assert (oracle.poseidon(s) == u64(pyref.poseidon([int(x) for x in s]))).all()


def test_reverse_index_bits_table(golden):
    """plonky2/src/util/mod.rs:56-123"""
    g = golden("reverse_index_bits.json")
    assert oracle.reverse_index_bits(u64(g["small_in"])).tolist() == g["small_out"]
    assert oracle.reverse_index_bits(np.arange(256, dtype=np.uint64)).tolist() == g["output256"]
    a = np.arange(1 << 16, dtype=np.uint64)
    b = oracle.reverse_index_bits(a)
    assert (oracle.reverse_index_bits(b) == a).all()
    assert pyref.reverse_index_bits(list(range(256))) == g["output256"]


def test_field_ops_boundary(golden):
    """field/src/prime_field_testing.rs:8-70 -- add/sub/mul vs big-integer arithmetic"""
    xs = [int(x) for x in golden("field_inputs.json")]
    L = oracle.lib()
    for a in xs:
        for b in xs:
            assert L.orc_gl_canon(L.orc_gl_add(a, b)) == (a + b) % P
            assert L.orc_gl_canon(L.orc_gl_sub(a, b)) == (a - b) % P
            assert L.orc_gl_canon(L.orc_gl_mul(a, b)) == (a * b) % P
    for a in xs[1:]:
        assert L.orc_gl_canon(L.orc_gl_mul(a, L.orc_gl_inv(a))) == 1
    # non-canonical operands
    for a in (P, P + 5, 2**64 - 1):
        for b in (P, 2**64 - 1, 3):
            assert L.orc_gl_canon(L.orc_gl_add(a, b)) == (a + b) % P
            assert L.orc_gl_canon(L.orc_gl_sub(a, b)) == (a - b) % P
            assert L.orc_gl_canon(L.orc_gl_mul(a, b)) == (a * b) % P
    for k in range(0, 33):
        assert L.orc_gl_canon(L.orc_gl_mul(L.orc_gl_inverse_2exp(k), pow(2, k, P))) == 1
        w = L.orc_gl_canon(L.orc_gl_primitive_root(k))
        assert pow(w, 1 << k, P) == 1 and (k == 0 or pow(w, 1 << (k - 1), P) == P - 1)
        assert w == pyref.primitive_root_of_unity(k)


# ---- FFT self-consistency (field/src/fft.rs:215-249, polynomial/mod.rs:477-516) -----------

def test_fft_equals_naive_and_inverse():
    deg, lg = 200, 8
    c = np.zeros(1 << lg, dtype=np.uint64)
    c[:deg] = oracle.rand_felts(deg, 1)
    naive = oracle.fft_naive(c)
    assert (oracle.fft(c) == naive).all()
    assert (oracle.ifft(naive) == c).all()
    assert oracle.fft(c).tolist() == pyref.fft([int(x) for x in c])
    assert oracle.ifft(naive).tolist() == pyref.ifft([int(x) for x in naive])


def test_fft_zero_factor():
    """fft_with_options(zero_factor=r) == fft when the top 1-2^-r of the input is zero"""
    lg = 8
    for r in range(0, 4):
        c = np.zeros(1 << lg, dtype=np.uint64)
        m = (1 << lg) >> r
        c[:m] = oracle.rand_felts(m, 10 + r)
        assert (oracle.fft(c, zero_factor=r) == oracle.fft(c)).all()


def test_coset_fft_equals_naive():
    lg = 8
    c = oracle.rand_felts(1 << lg, 3)
    shift = oracle.lib().orc_gl_coset_shift()
    assert shift == pyref.GENERATOR
    assert (oracle.coset_fft(c, shift) == oracle.fft_naive(c, shift)).all()
    assert oracle.coset_fft(c, shift).tolist() == pyref.coset_fft([int(x) for x in c], shift)


# ---- hashing rules -------------------------------------------------------------------------

@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 7, 8, 9, 15, 16, 17, 64, 135, 143])
def test_hash_leaf_vs_pyref(n):
    x = oracle.rand_felts(n, 100 + n)
    assert oracle.hash_leaf(x).tolist() == pyref.hash_leaf([int(v) for v in x])
    assert oracle.hash_no_pad(x).tolist() == pyref.hash_no_pad([int(v) for v in x])


def test_hash_leaf_domain_separation():
    """core/src/merkle_tree.rs:386-475"""
    left = oracle.hash_no_pad(u64([1, 2]))
    right = oracle.hash_no_pad(u64([3, 4]))
    internal = oracle.two_to_one(left, right)
    cat = np.concatenate([left, right])
    assert (oracle.hash_no_pad(cat) == internal).all()
    assert not (oracle.hash_leaf(cat) == internal).all()
    leaf = u64([1, 2, 3, 4, 5])
    z = u64([1, 2, 3, 4, 5, 0])
    assert not (oracle.hash_leaf(leaf) == oracle.hash_leaf(z)).all()
    t = oracle.MerkleTree(np.stack([leaf, u64([9] * 5)]), 0)
    pr = t.prove(0)
    assert oracle.merkle_verify(leaf, 0, t.cap, pr)
    assert not oracle.merkle_verify(z, 0, t.cap, pr)


def test_internal_node_cannot_masquerade_as_leaf():
    """core/src/merkle_tree.rs:328-384"""
    leaves = oracle.rand_felts((4, 7), 5)
    t = oracle.MerkleTree(leaves, 0)
    assert oracle.merkle_verify(leaves[0], 0, t.cap, t.prove(0))
    h0, h1 = oracle.hash_no_pad(leaves[0]), oracle.hash_no_pad(leaves[1])
    fake = np.concatenate([h0, h1])
    right_internal = oracle.two_to_one(oracle.hash_leaf(leaves[2]), oracle.hash_leaf(leaves[3]))
    assert not oracle.merkle_verify(fake, 0, t.cap, right_internal.reshape(1, 4))


# ---- Merkle tree (plonky2/src/hash/merkle_tree.rs:224-282) ---------------------------------

@pytest.mark.parametrize("cap_height", [0, 1, 8])
def test_merkle_every_leaf_round_trip(cap_height):
    n, k = 1 << 8, 7
    leaves = oracle.rand_felts((n, k), 42)
    t = oracle.MerkleTree(leaves, cap_height)
    for i in range(n):
        assert oracle.merkle_verify(leaves[i], i, t.cap, t.prove(i))
    bad = leaves[3].copy()
    bad[0] ^= np.uint64(1)
    assert not oracle.merkle_verify(bad, 3, t.cap, t.prove(3))


def test_merkle_cap_too_tall():
    with pytest.raises(ValueError):
        oracle.MerkleTree(oracle.rand_felts((1 << 8, 7), 1), 9)


@pytest.mark.parametrize("lg,cap_height,k", [(4, 0, 5), (4, 2, 9), (5, 5, 3), (3, 1, 16), (1, 0, 2), (0, 0, 4)])
def test_merkle_layout_vs_pyref(lg, cap_height, k):
    leaves = oracle.rand_felts((1 << lg, k), 9)
    t = oracle.MerkleTree(leaves, cap_height)
    dig, cap = pyref.merkle_tree([[int(x) for x in l] for l in leaves], cap_height)
    assert t.cap.tolist() == cap
    assert t.digests.tolist() == dig
    for i in range(1 << lg):
        assert t.prove(i).tolist() == pyref.merkle_prove(i, 1 << lg, cap_height, dig)


# ---- PolynomialBatch ------------------------------------------------------------------------

@pytest.mark.parametrize("lg_n,cols,rate,cap_h,salt", [(3, 3, 1, 0, False), (4, 5, 3, 2, False), (4, 2, 2, 4, True), (5, 9, 3, 4, False)])
def test_batch_vs_pyref(lg_n, cols, rate, cap_h, salt):
    vals = oracle.rand_felts((cols, 1 << lg_n), 11)
    s = oracle.rand_felts((4, (1 << lg_n) << rate), 12) if salt else None
    b = oracle.PolynomialBatch.from_values(vals, rate, cap_h, salt=s)
    coeffs, leaves, dig, cap = pyref.batch_from_values(
        [[int(x) for x in c] for c in vals], rate, cap_h, None if s is None else [[int(x) for x in c] for c in s])
    assert b.polynomials.tolist() == coeffs
    assert b.leaves.tolist() == leaves
    assert b.digests.tolist() == dig
    assert b.cap.tolist() == cap
    b2 = oracle.PolynomialBatch.from_coeffs(b.polynomials, rate, cap_h, salt=s)
    assert (b2.cap == b.cap).all() and (b2.leaves == b.leaves).all()


def test_batch_lde_is_low_degree_extension():
    """leaf i holds every column at g*w_N^bitrev(i) (oracle.rs:208-209, 286-291): check against
    direct evaluation of the interpolant, and that the coset restricted to step 2^r hits ... the
    original values only through the polynomial (values live on H, LDE on gH)."""
    lg_n, rate = 4, 2
    n, N = 1 << lg_n, 1 << (lg_n + rate)
    vals = oracle.rand_felts((3, n), 21)
    b = oracle.PolynomialBatch.from_values(vals, rate, 1)
    g, w = pyref.GENERATOR, pyref.primitive_root_of_unity(lg_n + rate)
    for c in range(3):
        co = [int(x) for x in b.polynomials[c]]
        # coefficients interpolate the values on H
        wn = pyref.primitive_root_of_unity(lg_n)
        for k in range(n):
            x = pow(wn, k, P)
            assert sum(ci * pow(x, i, P) for i, ci in enumerate(co)) % P == int(vals[c][k])
        for idx in range(0, N, 5):
            x = g * pow(w, idx, P) % P
            want = sum(ci * pow(x, i, P) for i, ci in enumerate(co)) % P
            assert int(b.get_lde_values(idx)[c]) == want


# ---- Challenger + FRI commit phase ----------------------------------------------------------

def test_challenger_vs_pyref():
    a, b = oracle.Challenger(), pyref.Challenger()
    rng = np.random.default_rng(3)
    for i in range(1, 12):
        xs = oracle.rand_felts(int(rng.integers(0, 20)), 50 + i)
        a.observe(xs)
        b.observe([int(x) for x in xs])
        for _ in range(i):
            assert a.get_challenge() == b.get_challenge()


def test_fri_arity_bits():
    """core/src/fri.rs:50-61 with standard_recursion_config (rate 3, cap 4, ConstantArityBits(4,5))"""
    assert oracle.fri_reduction_arity_bits(12, 3, 4) == [4, 4]
    assert oracle.fri_reduction_arity_bits(13, 3, 4) == [4, 4]
    assert oracle.fri_reduction_arity_bits(14, 3, 4) == [4, 4, 4]
    assert oracle.fri_reduction_arity_bits(20, 3, 4) == [4, 4, 4, 4]
    assert oracle.fri_reduction_arity_bits(23, 3, 4) == [4, 4, 4, 4, 4]


@pytest.mark.parametrize("deg_bits,rate,cap_h,arities", [(4, 1, 0, [1, 1]), (5, 2, 1, [2, 1]), (6, 1, 0, [3, 2])])
def test_fri_commit_vs_pyref(deg_bits, rate, cap_h, arities):
    n = 1 << (deg_bits + rate)
    co = np.zeros((n, 2), dtype=np.uint64)
    co[: 1 << deg_bits] = oracle.rand_felts((1 << deg_bits, 2), 77)
    g = pyref.GENERATOR
    l0 = pyref.coset_fft([int(x) for x in co[:, 0]], g)
    l1 = pyref.coset_fft([int(x) for x in co[:, 1]], g)
    va = u64(list(zip(l0, l1)))
    ca, cb = oracle.Challenger(), pyref.Challenger()
    ca.observe(u64([1, 2, 3]))
    cb.observe([1, 2, 3])
    r = oracle.fri_committed_trees(co, va, rate, cap_h, arities, ca, keep_trees=True)
    caps, betas, final, trees = pyref.fri_committed_trees(
        [tuple(int(x) for x in c) for c in co], [tuple(int(x) for x in v) for v in va], rate, cap_h, arities, cb)
    assert r["caps"].tolist() == caps
    assert r["betas"].tolist() == [list(b) for b in betas]
    assert r["final_poly"].tolist() == [list(c) for c in final]
    for k, (lv, dg) in enumerate(trees):
        assert r["leaves"][k].tolist() == lv
        assert r["digests"][k].tolist() == dg
    assert ca.get_challenge() == cb.get_challenge()
    # the folded polynomial really has degree < 2^deg_bits / prod(arity): trailing coeffs zero
    assert len(final) == (1 << deg_bits) >> sum(arities)


def test_fri_pow_smallest_witness():
    ch = oracle.Challenger()
    ch.observe(u64([5, 6, 7]))
    ref = ch.clone()
    w = oracle.fri_proof_of_work(ch, 8)
    # recompute with pyref: w is the first candidate whose response has >= 8 leading zeros
    pc = pyref.Challenger()
    pc.observe([5, 6, 7])
    for cand in range(w + 1):
        c2 = pyref.Challenger()
        c2.state, c2.inp, c2.out = list(pc.state), list(pc.inp), list(pc.out)
        c2.observe([cand])
        resp = c2.get_challenge()
        ok = resp < (1 << 56)
        assert ok == (cand == w)
    del ref


# ---- opening side ---------------------------------------------------------------------------

def _random_opening_batches(n, n_polys, seed, zero_point=False):
    rng = np.random.default_rng(seed)
    polys = oracle.rand_felts((n_polys, n), seed)
    batches = []
    for b in range(3):
        k = int(rng.integers(1, n_polys + 1))
        idx = rng.integers(0, n_polys, k)        # repeated polynomials are allowed
        w = oracle.rand_felts((k, 2), seed + 10 + b)
        pt = oracle.rand_felts(2, seed + 20 + b)
        if zero_point and b == 1:
            pt = np.zeros(2, dtype=np.uint64)
        batches.append(dict(point=tuple(int(x) for x in pt), shift=tuple(int(x) for x in oracle.rand_felts(2, seed + 30 + b)),
                            terms=[(polys[i], (int(a), int(c))) for i, (a, c) in zip(idx, w)], idx=[int(i) for i in idx]))
    return polys, batches


@pytest.mark.parametrize("lg,zero_point", [(0, False), (1, False), (5, False), (6, True)])
def test_reduce_openings_vs_pyref(lg, zero_point):
    n = 1 << lg
    polys, batches = _random_opening_batches(n, 6, 900 + lg, zero_point)
    got = oracle.reduce_openings(batches, lg)
    pb = [dict(point=b["point"], shift=b["shift"], terms=[([int(x) for x in p], w) for p, w in b["terms"]]) for b in batches]
    want = pyref.reduce_openings(pb, n)
    assert got.tolist() == [list(x) for x in want]
    # the quotient identity itself: (X - z) * q(X) + comp(z) == comp(X) for a single batch with shift anything
    one = [batches[0]]
    q = oracle.reduce_openings(one, lg)
    z = batches[0]["point"]
    comp = [(0, 0)] * n
    for p, w in pb[0]["terms"]:
        comp = [pyref.ext_add(comp[j], ((w[0] * p[j]) % P, (w[1] * p[j]) % P)) for j in range(n)]
    cz = (0, 0)
    for c in reversed(comp):
        cz = pyref.ext_add(pyref.ext_mul(cz, z), c)
    qq = [tuple(int(x) for x in r) for r in q]
    for j in range(n):
        lhs = pyref.ext_add(qq[j - 1] if j > 0 else (0, 0), tuple((-x) % P for x in pyref.ext_mul(z, qq[j])))
        if j == 0:
            lhs = pyref.ext_add(lhs, cz)
        assert lhs == comp[j]


def test_eval_poly_ext_vs_pyref():
    c = oracle.rand_felts(37, 5)
    pt = (123456789, 987654321)
    assert tuple(int(x) for x in oracle.eval_poly_ext(c, pt)) == pyref.eval_poly_ext([int(x) for x in c], pt)


def test_batch_merkle_tree_restatement():
    """BatchMerkleTree (plonky2/src/hash/batch_merkle_tree.rs): the reference's own structural tests
    (commit_single / commit_mixed, :179-258) and its every-leaf open -> verify_batch_merkle_proof_to_cap
    round trip (:286-336), on the big-integer restatement."""
    from oracle import pyref

    # commit_single: one matrix, cap height 0 -- an ordinary Merkle tree
    mat_1 = [[0, 1], [2, 1], [2, 2], [0, 0]]
    digests, cap, heights = pyref.batch_merkle_tree([mat_1], 0)
    h = [pyref.hash_leaf(r) for r in mat_1]
    assert digests[0:2] == h[0:2] and digests[4:6] == h[2:4] and heights == [2]
    layer_1 = [pyref.two_to_one(h[0], h[1]), pyref.two_to_one(h[2], h[3])]
    assert digests[2:4] == layer_1 and cap == [pyref.two_to_one(*layer_1)]
    # commit_mixed: a second matrix of half the height joins one level up, hashed WITH the digests
    mat_2 = [[1, 2, 1], [0, 2, 2]]
    digests, cap, heights = pyref.batch_merkle_tree([mat_1, mat_2], 0)
    assert digests[0:2] == h[0:2] and digests[2:4] == h[2:4] and heights == [2, 1]
    layer_2 = [pyref.hash_leaf(list(layer_1[0]) + mat_2[0]), pyref.hash_leaf(list(layer_1[1]) + mat_2[1])]
    assert digests[4:6] == layer_2 and cap == [pyref.two_to_one(*layer_2)]
    assert pyref.batch_merkle_open(1, [mat_1, mat_2], 0, digests) == [h[0], layer_2[1]]   # :279-280
    # random matrices, every leaf
    rng = np.random.default_rng(3)
    mats = [rng.integers(0, P, size=(32, 5), dtype=np.uint64).tolist(), rng.integers(0, P, size=(8, 3), dtype=np.uint64).tolist(),
            rng.integers(0, P, size=(4, 9), dtype=np.uint64).tolist()]
    for cap_h in (0, 1, 2):
        digests, cap, heights = pyref.batch_merkle_tree(mats, cap_h)
        assert len(digests) == 2 * (32 - (1 << cap_h)) and len(cap) == 1 << cap_h
        for i in range(32):
            proof = pyref.batch_merkle_open(i, mats, cap_h, digests)
            rows = [m[i >> (5 - hh)] for m, hh in zip(mats, heights)]
            assert pyref.batch_merkle_verify(rows, heights, i, cap, proof)
            bad = [list(r) for r in rows]
            bad[1][0] ^= 1
            assert not pyref.batch_merkle_verify(bad, heights, i, cap, proof)


def test_batch_fri_oracle_restatement():
    """BatchFriOracle::from_coeffs (plonky2/src/batch_fri/oracle.rs:105-160): with a single degree it is
    PolynomialBatch::from_coeffs (same leaves, digests, cap); with several, every row opens against
    the cap through verify_batch_merkle_proof_to_cap."""
    from oracle import pyref

    rng = np.random.default_rng(9)
    polys = [rng.integers(0, P, size=8, dtype=np.uint64).tolist() for _ in range(3)]
    mats, digests, cap, bits = pyref.batch_fri_from_coeffs(polys, 2, 1)
    leaves, d1, c1 = pyref.batch_from_coeffs(polys, 2, 1)
    assert mats == [leaves] and digests == d1 and cap == c1 and bits == [3]
    polys += [rng.integers(0, P, size=2, dtype=np.uint64).tolist()]
    mats, digests, cap, bits = pyref.batch_fri_from_coeffs(polys, 2, 1)
    assert bits == [3, 1] and [len(m) for m in mats] == [32, 8] and [len(m[0]) for m in mats] == [3, 1]
    for i in range(32):
        rows = [mats[0][i], mats[1][i >> 2]]
        assert pyref.batch_merkle_verify(rows, [5, 3], i, cap, pyref.batch_merkle_open(i, mats, 1, digests))
