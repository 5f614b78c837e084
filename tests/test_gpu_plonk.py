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


@pytest.mark.parametrize("degree_bits,qdf,nc,poseidon", [(5, 8, 2, False), (9, 8, 2, False), (10, 4, 2, False),
                                                          (11, 8, 1, False), (12, 8, 2, False), (6, 8, 2, True),
                                                          (10, 8, 2, True)])
def test_quotient_polys_match_oracle(qp, ctx, degree_bits, qdf, nc, poseidon):
    sc = SynthCircuit(degree_bits, seed=20 + degree_bits, quotient_degree_factor=qdf, num_challenges=nc,
                      poseidon=poseidon)
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


def test_inconsistent_lookup_declarations_are_refused(qp, ctx):
    """qp_circuit_create checks the lookup share of the circuit data against the reference's own formulas
    (circuit_builder.rs:1183-1194,1284-1290; gadgets/lookup.rs:80-160): lookup polynomials / selectors declared
    without tables, a wrong polynomial count, table rows that do not hold the table, or quotient evaluation before
    the lookup challenges are set -- all refused, never proved as if the lookups were not there."""
    sc = SynthCircuit(5, seed=3)
    c = sc.common
    c.num_lookup_polys, c.num_lookup_selectors = 2, 3
    try:
        with pytest.raises(qp.QpError) as e:
            plonk.Circuit(ctx, c, sc.sigmas)
        assert e.value.code == 5
    finally:
        c.num_lookup_polys = c.num_lookup_selectors = 0
    plonk.Circuit(ctx, c, sc.sigmas)
    lk = SynthCircuit(6, seed=4, lookups=True)
    c = lk.common
    good_polys, good_rows = c.num_lookup_polys, list(c.lookup_rows)
    try:
        c.num_lookup_polys = good_polys + 1
        with pytest.raises(qp.QpError):
            plonk.Circuit(ctx, c, lk.sigmas)
        c.num_lookup_polys = good_polys
        c.lookup_rows = [(good_rows[0][0], good_rows[0][1], good_rows[0][2] + 1)] + good_rows[1:]   # one row too many
        with pytest.raises(qp.QpError):
            plonk.Circuit(ctx, c, lk.sigmas, program_source="twin")
    finally:
        c.num_lookup_polys, c.lookup_rows = good_polys, good_rows
    circ = plonk.Circuit(ctx, c, lk.sigmas)
    g_w = qp.PolynomialBatch.from_values(ctx, lk.wires, c.rate_bits, False, c.cap_height)
    g_cs = qp.PolynomialBatch.from_values(ctx, lk.constants_sigmas(), c.rate_bits, False, c.cap_height)
    with pytest.raises(qp.QpError):     # lookup challenges not set
        circ.compute_quotient_polys(g_cs, g_w, g_w, [1, 2], [3, 4], [5, 6], [0, 0, 0, 0])


@pytest.mark.parametrize("degree_bits,nc,poseidon,device_witness", [(5, 2, False, False), (8, 2, True, False),
                                                                    (10, 1, False, True), (12, 2, True, True)])
def test_lookup_polys_match_oracle(qp, ctx, degree_bits, nc, poseidon, device_witness):
    """compute_all_lookup_polys (plonky2/src/plonk/prover.rs:489-636): RE and the partial Sum / LDC polynomials of
    two tables, bit for bit against the oracle, from a host and from a device-resident witness."""
    import torch

    sc = SynthCircuit(degree_bits, seed=30 + degree_bits, num_challenges=nc, poseidon=poseidon, lookups=True)
    c = sc.common
    deltas = oracle.rand_felts((4 * nc,), 90 + degree_bits)
    want = sc.oracle_circuit.lookup_polys(sc.wires, deltas)
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    w = torch.from_numpy(np.ascontiguousarray(sc.wires).view(np.int64)).cuda() if device_witness else sc.wires
    got = circ.lookup_polys(w, deltas)
    assert got.shape == want.shape == (nc * c.num_lookup_polys, 1 << degree_bits)
    assert (got == want).all()
    circ.free()


@pytest.mark.parametrize("degree_bits,nc,poseidon,rec", [(5, 2, False, False), (7, 2, True, True), (9, 1, False, False),
                                                         (10, 2, True, False)])
def test_quotient_polys_with_lookups_match_oracle(qp, ctx, degree_bits, nc, poseidon, rec):
    """compute_quotient_polys with the lookup terms of the vanishing polynomial (check_lookup_constraints_batch,
    vanishing_poly.rs:521-680; the gate constraints' powers of alpha start after them): bit-exact against the oracle."""
    sc = SynthCircuit(degree_bits, seed=40 + degree_bits, num_challenges=nc, poseidon=poseidon, extra_gates=rec,
                      recursion_gates=rec, lookups=True)
    c = sc.common
    betas, gammas, alphas = (a[:nc] for a in challenges(170 + degree_bits))
    deltas = np.concatenate([betas, gammas, oracle.rand_felts((2 * nc,), 180 + degree_bits)])
    oc = sc.oracle_circuit
    zs = np.concatenate([oc.partial_products_and_zs(sc.wires, sc.sigmas, betas, gammas), oc.lookup_polys(sc.wires, deltas)])
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    o_w = oracle.PolynomialBatch.from_values(sc.wires, c.rate_bits, c.cap_height)
    o_z = oracle.PolynomialBatch.from_values(zs, c.rate_bits, c.cap_height)
    want = oc.compute_quotient_polys(c.rate_bits, o_cs.leaves, o_w.leaves, o_z.leaves, betas, gammas, alphas,
                                     sc.public_inputs_hash, deltas=deltas)
    g_cs = qp.PolynomialBatch.from_values(ctx, sc.constants_sigmas(), c.rate_bits, False, c.cap_height)
    g_w = qp.PolynomialBatch.from_values(ctx, sc.wires, c.rate_bits, False, c.cap_height)
    g_z = qp.PolynomialBatch.from_values(ctx, zs, c.rate_bits, False, c.cap_height)
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    got = circ.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash, deltas=deltas)
    assert (got == want).all()
    # other challenges give another quotient: the terms are really in
    other = circ.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash,
                                        deltas=np.concatenate([betas, gammas, oracle.rand_felts((2 * nc,), 5)]))
    assert not (other == want).all()
    for b in (g_cs, g_w, g_z):
        b.free()
    circ.free()


@pytest.mark.parametrize("degree_bits,pow_bits,queries,poseidon,rec,zk", [
    (6, 6, 4, False, False, False), (9, 16, 28, True, False, False), (8, 10, 7, True, True, False),
    (7, 8, 6, True, False, True)])
def test_full_proof_with_lookups_bytes_match_oracle(qp, ctx, degree_bits, pow_bits, queries, poseidon, rec, zk):
    """prove() of a circuit with two lookup tables (deltas drawn after the gammas, lookup polynomials committed with
    the Z's, lookup_zs / lookup_zs_next in the opening set and in both FRI batches): byte for byte the oracle's proof,
    accepted by the restated verifier; also in zero-knowledge mode."""
    import verifier
    from oracle import prover as oprover
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(degree_bits, seed=80 + degree_bits, poseidon=poseidon, extra_gates=rec, recursion_gates=rec,
                      lookups=True)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    cfg = prover.FriConfig(c.rate_bits, c.cap_height, pow_bits, 4, 5, queries)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas(), cfg)
    N = (1 << degree_bits) << c.rate_bits
    salts = [oracle.rand_felts((4, N), 720 + k) for k in range(3)] if zk else None
    got = prover.prove(pd, sc.wires, sc.public_inputs, salts=salts)
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    want, _ = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                            degree_bits=degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                            num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                            num_partial_products=c.num_partial_products, rate_bits=c.rate_bits,
                            cap_height=c.cap_height, proof_of_work_bits=pow_bits, num_query_rounds=queries, salts=salts)
    assert len(got) == len(want)
    assert got == want
    cap = pd.constants_sigmas_commitment.merkle_tree.cap
    assert verifier.verify(got, c, pd.fri, cap, pd.circuit_digest, hiding=zk) is None


def test_large_lookup_proof_is_accepted_by_the_restated_verifier(qp, ctx):
    """2^13 rows, standard_recursion_config, all gate types + two lookup tables: beyond the size the oracle prover
    is run at, the restated verifier accepts the device's proof and rejects it after a bit flip in a lookup opening."""
    import verifier
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(13, seed=313, poseidon=True, extra_gates=True, recursion_gates=True, lookups=True)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
    proof = prover.prove(pd, sc.wires, sc.public_inputs)
    cap = pd.constants_sigmas_commitment.merkle_tree.cap
    assert verifier.verify(proof, c, pd.fri, cap, pd.circuit_digest) is None
    at = 3 * (4 << c.cap_height) * 8 + 16 * (c.num_constants + c.num_routed_wires + c.num_wires + 2 * c.num_challenges) + 5
    bad = bytearray(proof)
    bad[at] ^= 1
    assert verifier.verify(bytes(bad), c, pd.fri, cap, pd.circuit_digest) is not None


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


@pytest.mark.parametrize("degree_bits,qdf,pow_bits,queries,poseidon,extra,rec", [
    (6, 8, 6, 4, False, False, False), (9, 8, 16, 28, False, False, False), (8, 4, 10, 7, False, False, False),
    (11, 8, 16, 28, False, False, False), (10, 8, 16, 28, True, False, False), (9, 8, 12, 10, True, True, False),
    (9, 8, 12, 10, True, True, True), (7, 8, 8, 6, False, False, True)])
def test_full_proof_bytes_match_oracle(qp, ctx, degree_bits, qdf, pow_bits, queries, poseidon, extra, rec):
    """prove_with_partition_witness (plonky2/src/plonk/prover.rs:176-398) end to end on the device,
    serialised like write_proof_with_public_inputs -- byte for byte against the oracle's prove()."""
    from oracle import prover as oprover
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(degree_bits, seed=60 + degree_bits, quotient_degree_factor=qdf, poseidon=poseidon,
                      extra_gates=extra, recursion_gates=rec)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    cfg = prover.FriConfig(c.rate_bits, c.cap_height, pow_bits, 4, 5, queries)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas(), cfg)
    timing = {}
    got = prover.prove(pd, sc.wires, sc.public_inputs, timing)
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    assert (pd.circuit_digest == oprover.circuit_digest(o_cs.cap, degree_bits)).all()
    want, info = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                               degree_bits=degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                               num_challenges=c.num_challenges, quotient_degree_factor=qdf,
                               num_partial_products=c.num_partial_products, rate_bits=c.rate_bits,
                               cap_height=c.cap_height, proof_of_work_bits=pow_bits, num_query_rounds=queries)
    assert len(got) == len(want)
    assert got == want
    assert set(timing) >= {"compute wires commitment", "compute quotient polys", "compute opening proofs"}


@pytest.mark.parametrize("degree_bits,poseidon,rec,device_witness", [(6, False, False, False), (9, True, False, False),
                                                                      (8, True, True, True)])
def test_zero_knowledge_proof_bytes_match_oracle(qp, ctx, degree_bits, poseidon, rec, device_witness):
    """qp_prove_zk: config.zero_knowledge (plonky2/src/plonk/prover.rs:210,280,328) -- salted wires / Z /
    quotient commitments, leaf_hiding observed as 1, salted leaves in the query openings.  With the salt
    injected the proof is the oracle's byte for byte (host and device-resident witness), the restated verifier
    accepts it in hiding mode only, and without salt the call is the plain prove()."""
    import torch
    import verifier
    from oracle import prover as oprover
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(degree_bits, seed=70 + degree_bits, poseidon=poseidon, extra_gates=rec, recursion_gates=rec)
    c = sc.common
    N = (1 << degree_bits) << c.rate_bits
    salts = [oracle.rand_felts((4, N), 710 + k) for k in range(3)]
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    cfg = prover.FriConfig(c.rate_bits, c.cap_height, 8, 4, 5, 6)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas(), cfg)
    if device_witness:
        w = torch.from_numpy(np.ascontiguousarray(sc.wires).view(np.int64)).cuda()
        sl = [torch.from_numpy(x.view(np.int64)).cuda() for x in salts]
        got = prover.prove(pd, w, sc.public_inputs, salts=sl)
    else:
        got = prover.prove(pd, sc.wires, sc.public_inputs, salts=salts)
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    want, _ = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                            degree_bits=degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                            num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                            num_partial_products=c.num_partial_products, rate_bits=c.rate_bits,
                            cap_height=c.cap_height, proof_of_work_bits=8, num_query_rounds=6, salts=salts)
    assert len(got) == len(want)
    assert got == want
    cap = pd.constants_sigmas_commitment.merkle_tree.cap
    assert verifier.verify(got, c, pd.fri, cap, pd.circuit_digest, hiding=True) is None
    assert verifier.verify(got, c, pd.fri, cap, pd.circuit_digest, hiding=False) is not None
    plain = prover.prove(pd, sc.wires, sc.public_inputs)
    assert len(plain) == len(got) - 6 * 3 * 4 * 8 and verifier.verify(plain, c, pd.fri, cap, pd.circuit_digest) is None


@pytest.mark.parametrize("degree_bits,lookups,zk", [(7, False, False), (9, True, False), (8, False, True)])
def test_proof_from_witness_column_vectors(qp, ctx, degree_bits, lookups, zk):
    """qp_prove_cols / qp_mprove_cols: the witness as the reference holds it -- MatrixWitness.wire_values, one
    separately allocated (pageable) host vector per wire -- gives byte for byte the proof of the matrix form (host and
    device-resident), with lookup tables and in zero-knowledge mode, and over the multi-device driver."""
    import torch
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(degree_bits, seed=610 + degree_bits, poseidon=True, lookups=lookups)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    cfg = prover.FriConfig(c.rate_bits, c.cap_height, 8, 4, 5, 6)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas(), cfg)
    N = (1 << degree_bits) << c.rate_bits
    salts = [oracle.rand_felts((4, N), 730 + k) for k in range(3)] if zk else None
    want = prover.prove(pd, sc.wires, sc.public_inputs, salts=salts)
    cols = [np.array(sc.wires[w], copy=True) for w in range(c.num_wires)]     # one heap vector per wire
    assert prover.prove(pd, cols, sc.public_inputs, salts=salts) == want
    w_dev = torch.from_numpy(np.ascontiguousarray(sc.wires).view(np.int64)).cuda()
    s_dev = [torch.from_numpy(x.view(np.int64)).cuda() for x in salts] if zk else None
    assert prover.prove(pd, w_dev, sc.public_inputs, salts=s_dev) == want
    with pytest.raises(ValueError):
        prover.prove(pd, cols[:-1], sc.public_inputs)
    if not zk:
        m = qp.MultiContext([0], max_lde_log=degree_bits + c.rate_bits)
        try:
            mpd = prover.MultiProverData(m, c, sc.sigmas, sc.constants_sigmas(), cfg)
            assert prover.mprove(mpd, cols, sc.public_inputs) == want
            assert prover.mprove(mpd, sc.wires, sc.public_inputs) == want
            mpd.constants_sigmas_commitment.free()
            for x in mpd.circuits:
                x.free()
        finally:
            m.close()
    circ.free()


def test_factorial_example_circuit_on_the_device(qp, ctx):
    """BASELINE.json configs[0] (plonky2/examples/factorial.rs) as a real circuit (tests/factorial_circuit.py: chained
    multiplications, constants, the in-circuit public-input hash) under standard_recursion_config: the device's proof
    is the oracle's byte for byte, the restated verifier accepts it with public inputs (1, 100! mod p), and a proof for
    a wrong claimed result is refused."""
    import math
    import verifier
    from factorial_circuit import factorial_circuit
    from oracle import prover as oprover
    from qp_plonky2_b200 import prover

    sc = factorial_circuit()
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas())          # standard_recursion_config
    got = prover.prove(pd, sc.wires, sc.public_inputs)
    o_cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    want, _ = oprover.prove(sc.oracle_circuit, o_cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                            degree_bits=c.degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                            num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                            num_partial_products=c.num_partial_products)
    assert got == want
    assert prover.prove(pd, [np.array(sc.wires[w], copy=True) for w in range(c.num_wires)], sc.public_inputs) == want
    cap = pd.constants_sigmas_commitment.merkle_tree.cap
    assert verifier.verify(got, c, pd.fri, cap, pd.circuit_digest) is None
    assert [int(x) for x in np.frombuffer(got[-16:], dtype="<u8")] == [1, math.factorial(100) % P]
    forged = prover.prove(pd, sc.wires, [1, (sc.public_inputs[1] + 1) % P])
    assert verifier.verify(forged, c, pd.fri, cap, pd.circuit_digest) is not None
    circ.free()


def test_large_proof_openings_pass_the_verifier(qp, ctx):
    """2^15-row proof on the device (the oracle's quotient would take a minute): parse the opening
    set back out of the proof bytes, re-derive the challenges with the host transcript and run the
    restated verifier's algebraic check at zeta."""
    from qp_plonky2_b200 import prover
    from synth_circuit import verifier_plonk_identity

    sc = SynthCircuit(15, seed=77, poseidon=True)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
    proof = prover.prove(pd, sc.wires, sc.public_inputs)
    cap_words = (1 << c.cap_height) * 4
    n_open = (c.num_constants + c.num_routed_wires + c.num_wires + c.num_challenges * (2 + c.num_partial_products) +
              c.num_challenges * c.quotient_degree_factor)
    words = np.frombuffer(proof[: 8 * (3 * cap_words + 2 * n_open)], dtype="<u8")
    caps = [words[i * cap_words:(i + 1) * cap_words] for i in range(3)]
    pos = 3 * cap_words
    nc = c.num_challenges
    sizes = [("constants", c.num_constants), ("plonk_sigmas", c.num_routed_wires), ("wires", c.num_wires),
             ("plonk_zs", nc), ("plonk_zs_next", nc), ("partial_products", nc * c.num_partial_products),
             ("quotient_polys", nc * c.quotient_degree_factor)]
    openings = {}
    for name, k in sizes:
        openings[name] = words[pos:pos + 2 * k].reshape(k, 2)
        pos += 2 * k
    # replay the transcript (prover.rs:216-345 / verifier get_challenges.rs:39-96)
    ch = qp.Challenger()
    pd.fri.observe(ch, c.degree_bits, pd.reduction_arity_bits)
    ch.observe_elements(pd.circuit_digest)
    pih = prover.hash_no_pad(sc.public_inputs)
    ch.observe_elements(pih)
    ch.observe_cap(caps[0])
    betas, gammas = ch.get_n_challenges(nc), ch.get_n_challenges(nc)
    ch.observe_cap(caps[1])
    alphas = ch.get_n_challenges(nc)
    ch.observe_cap(caps[2])
    zeta = ch.get_extension_challenge()
    assert verifier_plonk_identity(c, openings, zeta, betas, gammas, alphas, pih)


@pytest.mark.parametrize("degree_bits,poseidon,extra,rec", [(10, False, False, False), (14, True, False, False),
                                                            (12, True, True, False), (12, True, True, True)])
def test_device_proof_is_accepted_by_the_restated_verifier(qp, ctx, degree_bits, poseidon, extra, rec):
    """The reference's own acceptance criterion: the full verifier (tests/verifier.py: transcript,
    plonk identity at zeta, PoW, 28 FRI query rounds with every Merkle path, folding consistency,
    final polynomial) accepts the device's proof under standard_recursion_config -- at a size where
    the oracle prover is no longer run -- and rejects it after a bit flip."""
    import verifier
    from qp_plonky2_b200 import prover

    sc = SynthCircuit(degree_bits, seed=300 + degree_bits, poseidon=poseidon, extra_gates=extra, recursion_gates=rec)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas())
    proof = prover.prove(pd, sc.wires, sc.public_inputs)
    cap = pd.constants_sigmas_commitment.merkle_tree.cap
    assert verifier.verify(proof, c, pd.fri, cap, pd.circuit_digest) is None
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 4
    assert verifier.verify(bytes(bad), c, pd.fri, cap, pd.circuit_digest) is not None


@pytest.mark.parametrize("source", ["dag", "twin"])
def test_quotient_from_an_external_recording(qp, ctx, source):
    """The device runs any valid program: the gates recorded by the Python mirror and compiled by the
    host library (qp_program_from_dag -- what a shim with its own gates does), or recorded AND compiled
    by the mirror (a WAIT after every load, no segments), give the oracle's quotient -- with
    PoseidonGate interpreted, not native."""
    sc = SynthCircuit(6, seed=21, poseidon=True, extra_gates=True, recursion_gates=True)
    c = sc.common
    circ = plonk.Circuit(ctx, c, sc.sigmas, program_source=source)
    betas, gammas, alphas = [11, 12], [13, 14], [15, 16]
    zs = circ.partial_products_and_zs(sc.wires, betas, gammas)
    g_cs = qp.PolynomialBatch.from_values(ctx, sc.constants_sigmas(), c.rate_bits, False, c.cap_height)
    g_w = qp.PolynomialBatch.from_values(ctx, sc.wires, c.rate_bits, False, c.cap_height)
    g_z = qp.PolynomialBatch.from_values(ctx, zs, c.rate_bits, False, c.cap_height)
    got = circ.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash)
    ref = plonk.Circuit(ctx, c, sc.sigmas)
    want = ref.compute_quotient_polys(g_cs, g_w, g_z, betas, gammas, alphas, sc.public_inputs_hash)
    assert (got == want).all()


@pytest.mark.parametrize("degree_bits,poseidon,rec,lookups", [(7, False, False, False), (10, True, False, False),
                                                              (12, True, True, False), (9, True, False, True)])
def test_multi_device_prove_equals_single_device_prove(qp, ctx, degree_bits, poseidon, rec, lookups):
    """qp_mprove (BASELINE.json configs[4]: prove() with coset-sharded commitments and quotient evaluation over
    every GPU of the box, one process): the proof is byte for byte the single-device proof, and the restated
    verifier accepts it.  On a one-GPU box the multi-device driver runs with a single device (same code path:
    shard gathering, peer copies onto itself, sharded query openings)."""
    import torch
    import verifier
    from qp_plonky2_b200 import prover

    D = torch.cuda.device_count()
    D = 8 if D >= 8 else 4 if D >= 4 else 2 if D >= 2 else 1
    sc = SynthCircuit(degree_bits, seed=500 + degree_bits, poseidon=poseidon, extra_gates=rec, recursion_gates=rec,
                      lookups=lookups)
    c = sc.common
    cfg = prover.FriConfig(c.rate_bits, c.cap_height, 10, 4, 5, 12)
    circ = plonk.Circuit(ctx, c, sc.sigmas)
    pd = prover.ProverData(ctx, circ, sc.constants_sigmas(), cfg)
    want = prover.prove(pd, sc.wires, sc.public_inputs)
    m = qp.MultiContext(list(range(D)), max_lde_log=degree_bits + c.rate_bits)
    try:
        mpd = prover.MultiProverData(m, c, sc.sigmas, sc.constants_sigmas(), cfg)
        assert (mpd.circuit_digest == pd.circuit_digest).all()
        timing = {}
        got = prover.mprove(mpd, sc.wires, sc.public_inputs, timing)
        assert len(got) == len(want)
        assert got == want
        cap = pd.constants_sigmas_commitment.merkle_tree.cap
        assert verifier.verify(got, c, pd.fri, cap, pd.circuit_digest) is None
        assert set(timing) >= {"compute wires commitment", "compute quotient polys"}
        mpd.constants_sigmas_commitment.free()
        for x in mpd.circuits:
            x.free()
    finally:
        m.close()
