"""GPU parity of the permutation argument and the quotient polynomials
(plonky2/src/plonk/prover.rs:402-480,640-866) through the C ABI, against the CPU oracle --
bit-exact at oracle-sized circuits, and through the verifier identity at large ones."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from qp_plonky2_b200 import plonk

from synth_circuit import SynthCircuit, verifier_identity_holds

P = oracle.P


@pytest.fixture(scope="module")
def qp():
    import qp_plonky2_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(qp):
    c = qp.Context(0, max_lde_log=22)
    yield c
    c.close()


def challenges(seed):
    return tuple(oracle.rand_felts((2,), seed + k) for k in range(3))


@pytest.mark.parametrize("degree_bits,qdf,nc", [(5, 8, 2), (8, 8, 2), (11, 8, 2), (12, 8, 2), (9, 4, 2), (10, 8, 1),
                                                 (13, 4, 2)])
def test_partial_products_and_zs_match_oracle(qp, ctx, degree_bits, qdf, nc):
    sc = SynthCircuit(degree_bits, seed=degree_bits, quotient_degree_factor=qdf, num_challenges=nc)
    betas, gammas, _ = challenges(50 + degree_bits)
    betas, gammas = betas[:nc], gammas[:nc]
    want = sc.oracle_circuit.partial_products_and_zs(sc.wires, sc.sigmas, betas, gammas)
    circ = plonk.Circuit(ctx, sc.common, sc.sigmas)
    got = circ.partial_products_and_zs(sc.wires, betas, gammas)
    assert got.shape == want.shape
    assert (got == want).all()
    circ.free()


@pytest.mark.parametrize("degree_bits,qdf,nc", [(5, 8, 2), (9, 8, 2), (10, 4, 2), (11, 8, 1), (12, 8, 2)])
def test_quotient_polys_match_oracle(qp, ctx, degree_bits, qdf, nc):
    sc = SynthCircuit(degree_bits, seed=20 + degree_bits, quotient_degree_factor=qdf, num_challenges=nc)
    c = sc.common
    betas, gammas, alphas = (a[:nc] for a in challenges(70 + degree_bits))
    zs = sc.oracle_circuit.partial_products_and_zs(sc.wires, sc.sigmas, betas, gammas)
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    o_w = oracle.PolynomialBatch.from_values(sc.wires, c.rate_bits, c.cap_height)
    o_z = oracle.PolynomialBatch.from_values(zs, c.rate_bits, c.cap_height)
    want = sc.oracle_circuit.compute_quotient_polys(c.rate_bits, o_cs.leaves, o_w.leaves, o_z.leaves, betas, gammas,
                                                    alphas, sc.public_inputs_hash)
    g_cs = qp.PolynomialBatch.from_values(ctx, sc.constants_sigmas(), c.rate_bits, False, c.cap_height)
    g_w = qp.PolynomialBatch.from_values(ctx, sc.wires, c.rate_bits, False, c.cap_height)
    g_z = qp.PolynomialBatch.from_values(ctx, zs, c.rate_bits, False, c.cap_height)
    circ = plonk.Circuit(ctx, sc.common, sc.sigmas)
    got = circ.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash)
    assert (got == want).all()
    # the chunks commit like the reference's from_coeffs of quotient_poly.chunks(degree), prover.rs:309-333
    chunks = got.reshape(nc << c.quotient_degree_bits, sc.n)
    g_q = qp.PolynomialBatch.from_coeffs(ctx, chunks, c.rate_bits, False, c.cap_height)
    o_q = oracle.PolynomialBatch.from_coeffs(want.reshape(chunks.shape), c.rate_bits, c.cap_height)
    assert (g_q.merkle_tree.cap == o_q.cap).all()
    for b in (g_cs, g_w, g_z, g_q):
        b.free()
    circ.free()


def test_quotient_errors(qp, ctx):
    sc = SynthCircuit(6, seed=3)
    c = sc.common
    circ = plonk.Circuit(ctx, sc.common, None)
    with pytest.raises(qp.QpError):     # created without sigmas
        circ.partial_products_and_zs(sc.wires, [1, 2], [3, 4])
    small = qp.PolynomialBatch.from_values(ctx, sc.wires[:, :32], c.rate_bits, False, 2)
    g_w = qp.PolynomialBatch.from_values(ctx, sc.wires, c.rate_bits, False, c.cap_height)
    with pytest.raises(qp.QpError) as e:  # "Polynomial degrees inconsistent"
        circ.compute_quotient_polys(small, g_w, g_w, [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])
    assert e.value.code == 4
    with pytest.raises(qp.QpError):     # too few polynomials in the Z batch
        few = qp.PolynomialBatch.from_values(ctx, sc.wires[:3], c.rate_bits, False, c.cap_height)
        circ.compute_quotient_polys(g_w, g_w, few, [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])


def test_large_circuit_verifier_identity(qp, ctx):
    """2^16 rows x 143 wires (the oracle would take minutes): the whole device pipeline -- Z and
    partial products, three commitments, quotient, quotient commitment -- checked through the
    verifier identity at random points, which holds only for the right quotient."""
    sc = SynthCircuit(16, seed=99)
    c = sc.common
    betas, gammas, alphas = challenges(500)
    circ = plonk.Circuit(ctx, sc.common, sc.sigmas)
    zs = circ.partial_products_and_zs(sc.wires, betas, gammas)
    assert (zs[:2, 0] == 1).all()
    g_cs = qp.PolynomialBatch.from_values(ctx, sc.constants_sigmas(), c.rate_bits, False, c.cap_height)
    g_w = qp.PolynomialBatch.from_values(ctx, sc.wires, c.rate_bits, False, c.cap_height)
    g_z = qp.PolynomialBatch.from_values(ctx, zs, c.rate_bits, False, c.cap_height)
    q = circ.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash)
    for x0 in (5, 0xDEADBEEFCAFEF00D % P):
        assert verifier_identity_holds(sc, g_cs.polynomials, g_w.polynomials, g_z.polynomials, q, betas, gammas,
                                       alphas, x0)
