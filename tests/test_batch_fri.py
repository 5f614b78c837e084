"""Batch FRI prover (plonky2/src/batch_fri/prover.rs, batch_fri/oracle.rs:163-229): the big-integer
restatement against the restated batch verifier (CPU), and the device path against both (GPU).

The configurations are the reference's own tests (batch_fri/prover.rs:232-483): `single_polynomial`
(k = 9, values 1..n) and `multiple_polynomials` (k = 9, 8, 6), reduction_arity_bits [1, 2, 1], rate_bits 1,
cap_height 5, no proof of work, 10 query rounds; the acceptance criterion is theirs too
(verify_batch_fri_proof, batch_fri/verifier.rs)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import verifier  # noqa: E402
from oracle import P, pyref  # noqa: E402

RATE, CAP, ARITIES, QUERIES = 1, 5, [1, 2, 1], 10


def _transcript_prefix(ch, cap, polys):
    """The reference tests' transcript before the opening proof: observe the cap, draw two alphas and zeta,
    observe every polynomial's value at zeta.  `ch` has observe / get_challenge / get_extension_challenge."""
    ch.observe([int(x) for d in cap for x in d])
    ch.get_challenge(), ch.get_challenge()
    zeta = ch.get_extension_challenge()
    evals = [pyref.eval_poly_ext(p, zeta) for p in polys]
    for e in evals:
        ch.observe(list(e))
    return zeta, evals


def _instances(zeta, degree_groups):
    """One instance per degree: every polynomial of that degree opened at zeta (raw expressions)."""
    inst, pi = [], 0
    for k in degree_groups:
        inst.append(dict(oracles=[k], batches=[dict(point=zeta, openings=[[(0, pi + j, "one")] for j in range(k)])]))
        pi += k
    return inst


def _clone(ch):
    c = pyref.Challenger()
    c.state, c.inp, c.out = list(ch.state), list(ch.inp), list(ch.out)
    return c


def _case(lens, seed, counting=False):
    rng = np.random.default_rng(seed)
    values = [list(range(1, n + 1)) if counting else [int(x) for x in rng.integers(0, P, size=n, dtype=np.uint64)]
              for n in lens]
    polys = [pyref.ifft(v) for v in values]
    mats, digests, cap, bits = pyref.batch_fri_from_coeffs(polys, RATE, CAP)
    groups = [sum(1 for n in lens if n == 1 << d) for d in bits]
    return values, polys, (mats, digests), cap, bits, groups


@pytest.mark.parametrize("lens,counting", [([512], True), ([512, 256, 64], False), ([512, 512, 256, 64, 64], False)])
def test_batch_fri_restatement_is_accepted(lens, counting):
    """oracle/pyref.py's BatchFriOracle::prove_openings -> verify_batch_fri_proof accepts; a flipped bit in any
    region of the proof is rejected."""
    values, polys, tree, cap, bits, groups = _case(lens, 11, counting)
    ch = pyref.Challenger()
    zeta, evals = _transcript_prefix(ch, cap, polys)
    inst = _instances(zeta, groups)
    vch = _clone(ch)
    proof = pyref.batch_prove_openings_bytes(bits, inst, [polys], [tree], ch, RATE, CAP, ARITIES, 0, QUERIES)
    openings, pi = [], 0
    for k in groups:
        openings.append([[evals[pi + j] for j in range(k)]])
        pi += k
    assert verifier.verify_batch_fri_proof(bits, inst, openings, _clone(vch), [cap], proof, RATE, CAP, ARITIES, 0,
                                           QUERIES) is None
    # corrupt: a commit-phase cap, an initial evaluation, a query step, the final polynomial, a claimed opening
    cap_bytes = 32 << CAP
    for pos in (3, cap_bytes * len(ARITIES) + 2, len(proof) // 2, len(proof) - 20):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert verifier.verify_batch_fri_proof(bits, inst, openings, _clone(vch), [cap], bytes(bad), RATE, CAP, ARITIES,
                                               0, QUERIES) is not None, pos
    wrong = [[list(b) for b in o] for o in openings]
    wrong[-1][0][0] = ((wrong[-1][0][0][0] + 1) % P, wrong[-1][0][0][1])
    assert verifier.verify_batch_fri_proof(bits, inst, wrong, _clone(vch), [cap], proof, RATE, CAP, ARITIES, 0,
                                           QUERIES) is not None


def test_batch_fri_single_degree_equals_fri():
    """With one degree the batch commit phase is fri_committed_trees."""
    rng = np.random.default_rng(5)
    n = 64
    co = [(int(a), int(b)) for a, b in rng.integers(0, P, size=(n, 2), dtype=np.uint64)]
    co[n >> RATE:] = [(0, 0)] * (n - (n >> RATE))
    va = list(zip(pyref.coset_fft([c[0] for c in co], pyref.GENERATOR), pyref.coset_fft([c[1] for c in co], pyref.GENERATOR)))
    a = pyref.fri_committed_trees(co, va, RATE, 2, [1, 2], pyref.Challenger())
    b = pyref.batch_fri_committed_trees(co, [va], RATE, 2, [1, 2], pyref.Challenger())
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]


# ---- device ------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def qp():
    import qp_plonky2_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(qp):
    c = qp.Context(0, max_lde_log=20)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lens,counting,pow_bits", [([512], True, 0), ([512, 256, 64], False, 0),
                                                    ([512, 512, 256, 64, 64], False, 3)])
def test_batch_fri_prove_openings_matches_restatement(qp, ctx, lens, counting, pow_bits):
    """BatchFriOracle::prove_openings on the device: byte-identical to the restatement, accepted by the verifier."""
    values, polys, tree, cap, bits, groups = _case(lens, 11, counting)
    o = qp.BatchFriOracle.from_values(ctx, [np.array(v, dtype=np.uint64) for v in values], RATE, False, CAP)
    assert o.cap.tolist() == [list(c) for c in cap]
    ch, och = qp.Challenger(), pyref.Challenger()
    zeta, evals = _transcript_prefix(och, cap, polys)
    ch.observe_cap(o.cap)
    ch.get_n_challenges(2)
    assert ch.get_extension_challenge() == zeta
    for e in evals:
        ch.observe_elements(list(e))
    inst = _instances(zeta, groups)
    vch = _clone(och)
    got = qp.BatchFriOracle.prove_openings(ctx, bits, inst, [o], ch, RATE, CAP, ARITIES, pow_bits, QUERIES)
    want = pyref.batch_prove_openings_bytes(bits, inst, [polys], [tree], och, RATE, CAP, ARITIES, pow_bits, QUERIES)
    assert got == want
    assert ch.get_challenge() == och.get_challenge()          # transcripts left in the same state
    openings, pi = [], 0
    for k in groups:
        openings.append([[evals[pi + j] for j in range(k)]])
        pi += k
    assert verifier.verify_batch_fri_proof(bits, inst, openings, vch, [cap], got, RATE, CAP, ARITIES, pow_bits,
                                           QUERIES) is None
    o.free()


@pytest.mark.gpu
def test_batch_fri_proof_two_points_and_larger(qp, ctx):
    """A starky-shaped opening: 2^12 / 2^10 / 2^7 rows, several columns per degree, every polynomial opened at
    zeta and the tallest ones also at g * zeta, arity-4 schedule; accepted by the restated verifier."""
    rate, cap_h, arities, queries, pow_bits = 1, 3, [2, 3, 2, 2], 12, 4
    lens = [1 << 12] * 5 + [1 << 10] * 3 + [1 << 7] * 2
    rng = np.random.default_rng(3)
    polys = [rng.integers(0, P, size=n, dtype=np.uint64) for n in lens]
    o = qp.BatchFriOracle.from_coeffs(ctx, polys, rate, False, cap_h)
    bits, groups = o.degree_bits, o.group_sizes
    cap = o.cap
    ch, och = qp.Challenger(), pyref.Challenger()
    ch.observe_cap(cap)
    och.observe([int(x) for x in cap.reshape(-1)])
    zeta = ch.get_extension_challenge()
    assert och.get_extension_challenge() == zeta
    g = pyref.primitive_root_of_unity(bits[0])
    zeta_next = (zeta[0] * g % P, zeta[1] * g % P)
    inst, openings, pi = [], [], 0
    for d, k in zip(bits, groups):
        at = lambda pt, idx: [pyref.eval_poly_ext([int(c) for c in polys[i]], pt) for i in idx]
        idx = list(range(pi, pi + k))
        batches = [dict(point=zeta, openings=[[(0, i, "one")] for i in idx])]
        vals = [at(zeta, idx)]
        if d == bits[0]:
            batches.append(dict(point=zeta_next, openings=[[(0, i, "one")] for i in idx[:2]]))
            vals.append(at(zeta_next, idx[:2]))
        inst.append(dict(oracles=[k], batches=batches))
        openings.append(vals)
        pi += k
    for inst_vals in openings:
        for vals in inst_vals:
            for e in vals:
                ch.observe_elements(list(e))
                och.observe(list(e))
    proof = qp.BatchFriOracle.prove_openings(ctx, bits, inst, [o], ch, rate, cap_h, arities, pow_bits, queries)
    assert verifier.verify_batch_fri_proof(bits, inst, openings, och, [cap.tolist()], proof, rate, cap_h, arities,
                                           pow_bits, queries) is None
    bad = bytearray(proof)
    bad[len(bad) // 3] ^= 4
    och2 = pyref.Challenger()
    och2.observe([int(x) for x in cap.reshape(-1)])
    och2.get_extension_challenge()
    for inst_vals in openings:
        for vals in inst_vals:
            for e in vals:
                och2.observe(list(e))
    assert verifier.verify_batch_fri_proof(bits, inst, openings, och2, [cap.tolist()], bytes(bad), rate, cap_h, arities,
                                           pow_bits, queries) is not None
    o.free()


@pytest.mark.gpu
def test_batch_fri_errors(qp, ctx):
    """The reference's asserts (batch_fri/prover.rs:36-52,142): degrees must strictly decrease and every
    polynomial must be reached by the reduction schedule before the last fold."""
    rng = np.random.default_rng(9)

    def fri(n):
        co = np.zeros((n, 2), dtype=np.uint64)
        co[: n // 2] = rng.integers(0, P, size=(n // 2, 2), dtype=np.uint64)
        return qp.fri_begin(ctx, co, co, 1, 0)

    o = qp.BatchFriOracle.from_coeffs(ctx, [rng.integers(0, P, size=32, dtype=np.uint64)], 1, False, 0)
    with pytest.raises(qp.QpError):      # 2^4 is never the current length under [1, 1]
        qp.batch_fri_proof(ctx, [o], [fri(64), fri(16)], qp.Challenger(), 1, 0, [1, 1], 0, 2)
    with pytest.raises(qp.QpError):      # not strictly decreasing
        qp.batch_fri_proof(ctx, [o], [fri(64), fri(32), fri(32)], qp.Challenger(), 1, 0, [1, 1, 1], 0, 2)
    o.free()
