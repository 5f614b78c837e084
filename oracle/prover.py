"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement of prove_with_partition_witness from the full witness on
(plonky2/src/plonk/prover.rs:176-398) out of the oracle's pieces, and of
write_proof_with_public_inputs (plonky2/src/util/serialization/mod.rs:2040-2079).  Same scope as
the device prover: lookup argument included, smallest PoW witness, zero-knowledge salt injected.  Parity status: "unpinned" by
golden data (the reference cannot be built here); pinned structurally by the verifier identity
(tests/test_plonk_oracle.py) and, for the FRI part, by tests/test_oracle.py.
"""
import numpy as np

import oracle
from oracle import P

W = 7


def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def circuit_digest(cap, degree_bits):
    ds = oracle.hash_no_pad(np.array([1, 0, 0, 0, 0, 0, 0, 1], dtype=np.uint64))
    return oracle.hash_no_pad(np.concatenate([np.asarray(cap, dtype=np.uint64).reshape(-1), ds,
                                              np.array([degree_bits], dtype=np.uint64)]))


def prove(oc, cs_batch, num_constants, wires, sigmas, public_inputs, pih_wires_check=None, *, degree_bits,
          num_wires, num_routed_wires, num_challenges, quotient_degree_factor, num_partial_products,
          rate_bits=3, cap_height=4, proof_of_work_bits=16, arity_bits=4, final_poly_bits=5, num_query_rounds=28,
          salts=None):
    """oc: oracle.Circuit; cs_batch: oracle.PolynomialBatch of constants + sigmas.  salts: None, or the injected
    salt columns (wires, zs, quotient), each [4][N]: config.zero_knowledge (prover.rs:210,280,328) -- the three
    oracles are salted (fri/oracle.rs:259-263) and leaf_hiding is observed as 1 (core/src/fri.rs:311)."""
    n = 1 << degree_bits
    nc = num_challenges
    public_inputs = [int(x) % P for x in public_inputs]
    pih = oracle.hash_no_pad(np.array(public_inputs, dtype=np.uint64))
    zk = salts is not None
    sw, sz, sq = salts if zk else (None, None, None)
    wb = oracle.PolynomialBatch.from_values(wires, rate_bits, cap_height, salt=sw)
    ch = oracle.Challenger()
    arities = oracle.fri_reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits, final_poly_bits)
    # FriParams::observe, core/src/fri.rs:289-321
    ch.observe([rate_bits, cap_height, proof_of_work_bits, 1, arity_bits, final_poly_bits, num_query_rounds, int(zk),
                degree_bits] + list(arities))
    ch.observe(circuit_digest(cs_batch.cap, degree_bits))
    ch.observe(pih)
    ch.observe(wb.cap.reshape(-1))
    betas = [ch.get_challenge() for _ in range(nc)]
    gammas = [ch.get_challenge() for _ in range(nc)]
    # prover.rs:227-243: with lookup tables, 2 nc more challenges; deltas = betas ++ gammas ++ those
    has_lookup = oc.num_lookup_polys != 0
    deltas = betas + gammas + [ch.get_challenge() for _ in range(2 * nc)] if has_lookup else None
    zs = oc.partial_products_and_zs(wires, sigmas, betas, gammas)
    n_zs = zs.shape[0]
    if has_lookup:   # compute_all_lookup_polys, committed after the Z's and partial products (prover.rs:262-271)
        zs = np.concatenate([zs, oc.lookup_polys(wires, deltas)])
    zb = oracle.PolynomialBatch.from_values(zs, rate_bits, cap_height, salt=sz)
    ch.observe(zb.cap.reshape(-1))
    alphas = [ch.get_challenge() for _ in range(nc)]
    def unsalted(b):   # get_lde_values strips the salt (fri/oracle.rs:290)
        return np.ascontiguousarray(b.leaves[:, : b.leaves.shape[1] - 4]) if b.blinding else b.leaves

    q = oc.compute_quotient_polys(rate_bits, cs_batch.leaves, unsalted(wb), unsalted(zb), betas, gammas, alphas, pih,
                                  deltas=deltas)
    qd = quotient_degree_factor * n
    assert not q[:, qd:].any(), "Quotient has failed, the vanishing polynomial is not divisible by Z_H"
    chunks = np.ascontiguousarray(q[:, :qd]).reshape(nc * quotient_degree_factor, n)
    qb = oracle.PolynomialBatch.from_coeffs(chunks, rate_bits, cap_height, salt=sq)
    ch.observe(qb.cap.reshape(-1))
    zeta = ch.get_extension_challenge()
    g = oracle.lib().orc_gl_primitive_root(degree_bits)
    zeta_next = (g * zeta[0] % P, g * zeta[1] % P)

    def ev(batch, point):
        return np.stack([oracle.eval_poly_ext(p, point) for p in batch.polynomials])

    cs_eval, wires_eval, zs_eval = ev(cs_batch, zeta), ev(wb, zeta), ev(zb, zeta)
    zs_next_eval, quotient_eval = ev(zb, zeta_next), ev(qb, zeta)
    n_pre = num_constants + num_routed_wires
    constants, sig = cs_eval[:num_constants], cs_eval[num_constants:n_pre]
    plonk_zs, plonk_zs_next, pps = zs_eval[:nc], zs_next_eval[:nc], zs_eval[nc:n_zs]
    lookup_zs, lookup_zs_next = zs_eval[n_zs:], zs_next_eval[n_zs:]      # lookup_range, circuit_data.rs:582
    # observe_openings(to_fri_openings()), proof.rs:328-368
    for v in (constants, sig, wires_eval, plonk_zs, pps, quotient_eval, lookup_zs, plonk_zs_next, lookup_zs_next):
        ch.observe(np.asarray(v).reshape(-1))
    alpha = ch.get_extension_challenge()
    batches = []
    # fri_all_openings / fri_next_batch_openings, circuit_data.rs:711-747
    zeta_polys = (list(cs_batch.polynomials[:n_pre]) + list(wb.polynomials) + list(zb.polynomials[:n_zs]) +
                  list(qb.polynomials) + list(zb.polynomials[n_zs:]))
    for point, polys in ((zeta, zeta_polys), (zeta_next, list(zb.polynomials[:nc]) + list(zb.polynomials[n_zs:]))):
        terms, w = [], (1, 0)
        for p in polys:
            terms.append((p, w))
            w = ext_mul(w, alpha)
        batches.append(dict(point=point, shift=w, terms=terms))
    final = oracle.reduce_openings(batches, degree_bits)
    N = n << rate_bits
    co = np.zeros((N, 2), dtype=np.uint64)
    co[:n] = final
    gs = oracle.lib().orc_gl_coset_shift()
    va = np.stack([oracle.coset_fft(co[:, 0], gs), oracle.coset_fft(co[:, 1], gs)], axis=1)
    fri_bytes = oracle.fri_proof_bytes([cs_batch, wb, zb, qb], co, va, ch, rate_bits, cap_height, arities,
                                       proof_of_work_bits, num_query_rounds)
    out = bytearray()
    for cap in (wb.cap, zb.cap, qb.cap):
        out += np.ascontiguousarray(cap).astype("<u8").tobytes()
    # write_opening_set, serialization/mod.rs:1495-1508
    for v in (constants, sig, wires_eval, plonk_zs, plonk_zs_next, lookup_zs, lookup_zs_next, pps, quotient_eval):
        out += np.ascontiguousarray(v).astype("<u8").tobytes()
    out += fri_bytes
    out += np.array([len(public_inputs)] + public_inputs, dtype="<u8").tobytes()
    return bytes(out), dict(zeta=zeta, alphas=alphas, betas=betas, gammas=gammas, deltas=deltas, pih=pih, openings=dict(
        constants=constants, plonk_sigmas=sig, wires=wires_eval, plonk_zs=plonk_zs, plonk_zs_next=plonk_zs_next,
        partial_products=pps, quotient_polys=quotient_eval, lookup_zs=lookup_zs, lookup_zs_next=lookup_zs_next))
