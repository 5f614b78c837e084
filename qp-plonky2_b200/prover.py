"""prove_with_partition_witness from the full witness on (plonky2/src/plonk/prover.rs:176-398), every
polynomial-sized step on the device:

    wires commitment            PolynomialBatch.from_values            prover.rs:201-214
    transcript                  Challenger (host, core/src/challenger.rs)   prover.rs:216-234
    Z / partial products        Circuit.partial_products_and_zs        prover.rs:250-261
    their commitment            PolynomialBatch.from_values            prover.rs:275-287
    quotient polynomials        Circuit.compute_quotient_polys         prover.rs:293-307
    quotient commitment         PolynomialBatch.from_coeffs            prover.rs:322-334
    openings at zeta, g zeta    PolynomialBatch.eval_polys             prover.rs:349-361 (OpeningSet::new)
    opening proof               fri_from_openings + fri_proof          prover.rs:364-380 (prove_openings)
    serialization               write_proof_with_public_inputs         util/serialization/mod.rs:2040-2079

Out of scope: circuit building and witness generation -- the caller brings
the constants/sigmas commitment and the witness matrix, like `prove_with_partition_witness` gets
them from ProverOnlyCircuitData and the generators.  No lookups, no zero-knowledge blinding.
Deterministic where the reference is not: the PoW witness is the smallest one (serial `find`).
"""
import numpy as np

P = 0xFFFFFFFF00000001
W = 7  # F_p^2 = F_p[X] / (X^2 - 7), field/src/goldilocks_extensions.rs:13-26


def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


class FriConfig:
    """core/src/fri.rs FriConfig with FriReductionStrategy::ConstantArityBits; defaults are
    standard_recursion_config (core/src/circuit_config.rs:51-70)."""

    def __init__(self, rate_bits=3, cap_height=4, proof_of_work_bits=16, arity_bits=4, final_poly_bits=5,
                 num_query_rounds=28):
        self.rate_bits, self.cap_height = rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.arity_bits, self.final_poly_bits = arity_bits, final_poly_bits

    def observe(self, challenger, degree_bits, reduction_arity_bits):
        """FriConfig::observe + FriParams::observe, core/src/fri.rs:289-321."""
        challenger.observe_elements([self.rate_bits, self.cap_height, self.proof_of_work_bits])
        challenger.observe_elements([1, self.arity_bits, self.final_poly_bits])  # ConstantArityBits.serialize()
        challenger.observe_element(self.num_query_rounds)
        challenger.observe_element(0)  # leaf_hiding
        challenger.observe_element(degree_bits)
        challenger.observe_elements(list(reduction_arity_bits))


def hash_no_pad(ctx, elems):
    """hash_no_pad (core/src/hashing.rs:68-95) through the device permutation: a handful of
    permutations per proof (public inputs, circuit digest)."""
    state = np.zeros(12, dtype=np.uint64)
    elems = [int(e) % P for e in elems]
    for i in range(0, len(elems), 8):
        chunk = elems[i:i + 8]
        state[: len(chunk)] = chunk
        state = ctx.poseidon(state.reshape(1, 12))[0]
    if not elems:
        pass
    return state[:4].copy()


def circuit_digest(ctx, constants_sigmas_cap, degree_bits):
    """circuit_builder.rs:1289-1303: hash_no_pad(cap || hash_pad(domain_separator = []) || degree_bits)."""
    # hash_pad (core/src/config.rs:46-54) of the empty vector: [1, 0, 0, 0, 0, 0, 0, 1]
    ds = hash_no_pad(ctx, [1, 0, 0, 0, 0, 0, 0, 1])
    return hash_no_pad(ctx, list(np.asarray(constants_sigmas_cap).reshape(-1)) + list(ds) + [degree_bits])


class ProverData:
    """What prove() needs from ProverOnlyCircuitData (circuit_data.rs:497-540): the constants /
    sigmas commitment (circuit_builder.rs:1207-1227) and the circuit digest."""

    def __init__(self, ctx, circuit, constants_sigmas_values, fri_config=None):
        from . import PolynomialBatch, fri_reduction_arity_bits
        c = circuit.common
        self.ctx, self.circuit = ctx, circuit
        self.fri = fri_config or FriConfig(c.rate_bits, c.cap_height)
        self.constants_sigmas_commitment = PolynomialBatch.from_values(
            ctx, constants_sigmas_values, self.fri.rate_bits, False, self.fri.cap_height)
        self.reduction_arity_bits = fri_reduction_arity_bits(
            c.degree_bits, self.fri.rate_bits, self.fri.cap_height, self.fri.arity_bits, self.fri.final_poly_bits)
        self.circuit_digest = circuit_digest(ctx, self.constants_sigmas_commitment.merkle_tree.cap, c.degree_bits)


def _ext_bytes(arr):
    return np.ascontiguousarray(arr, dtype=np.uint64).astype("<u8").tobytes()


def prove(prover_data, wires, public_inputs, timing=None):
    """-> bytes of ProofWithPublicInputs.  wires: witness matrix [num_wires][n] (numpy or torch
    CUDA); public_inputs: list of field elements (already part of the witness)."""
    import time

    from . import Challenger, PolynomialBatch, _is_torch, fri_from_openings, fri_proof
    pd = prover_data
    ctx, circuit = pd.ctx, pd.circuit
    c = circuit.common
    fri = pd.fri
    nc = c.num_challenges
    t = {}
    t0 = time.perf_counter()

    def lap(name):
        nonlocal t0
        ctx.synchronize()
        t1 = time.perf_counter()
        t[name] = (t1 - t0) * 1e3
        t0 = t1

    public_inputs = [int(x) % P for x in public_inputs]
    public_inputs_hash = hash_no_pad(ctx, public_inputs)
    wires_commitment = PolynomialBatch.from_values(ctx, wires, fri.rate_bits, False, fri.cap_height)
    lap("compute wires commitment")
    challenger = Challenger()
    fri.observe(challenger, c.degree_bits, pd.reduction_arity_bits)
    challenger.observe_elements(pd.circuit_digest)
    challenger.observe_elements(public_inputs_hash)
    wires_cap = wires_commitment.merkle_tree.cap
    challenger.observe_cap(wires_cap)
    betas = challenger.get_n_challenges(nc)
    gammas = challenger.get_n_challenges(nc)
    # a device-resident witness keeps every intermediate polynomial on the device
    on_device = _is_torch(wires) and wires.is_cuda
    n = 1 << c.degree_bits
    zs_pp = None
    if on_device:
        import torch
        zs_pp = torch.empty((nc * (1 + c.num_partial_products), n), dtype=torch.int64, device=wires.device)
    zs_pp = circuit.partial_products_and_zs(wires, betas, gammas, out_device=zs_pp)
    lap("compute partial products")
    zs_commitment = PolynomialBatch.from_values(ctx, zs_pp, fri.rate_bits, False, fri.cap_height)
    lap("commit to partial products, Z's")
    zs_cap = zs_commitment.merkle_tree.cap
    challenger.observe_cap(zs_cap)
    alphas = challenger.get_n_challenges(nc)
    quotient = None
    if on_device:
        quotient = torch.empty((nc, n << c.quotient_degree_bits), dtype=torch.int64, device=wires.device)
    quotient = circuit.compute_quotient_polys(pd.constants_sigmas_commitment, wires_commitment, zs_commitment,
                                              betas, gammas, alphas, public_inputs_hash, out_device=quotient)
    lap("compute quotient polys")
    # prover.rs:309-320: trim to quotient_degree = factor * n and split into degree-n chunks
    qd = c.quotient_degree_factor * n
    if quotient.shape[1] > qd and bool(quotient[:, qd:].any()):
        raise ValueError("Quotient has failed, the vanishing polynomial is not divisible by Z_H")
    chunks = quotient[:, :qd].contiguous() if on_device else np.ascontiguousarray(quotient[:, :qd])
    chunks = chunks.reshape(nc * c.quotient_degree_factor, n)
    if on_device:
        torch.cuda.current_stream().synchronize()  # a trimming copy runs on torch's stream, the library on its own
    quotient_commitment = PolynomialBatch.from_coeffs(ctx, chunks, fri.rate_bits, False, fri.cap_height)
    lap("commit to quotient polys")
    quotient_cap = quotient_commitment.merkle_tree.cap
    challenger.observe_cap(quotient_cap)
    zeta = challenger.get_extension_challenge()
    if ext_pow(zeta, n) == (1, 0):
        raise ValueError("Opening point is in the subgroup.")
    g = pow(7277203076849721926, 1 << (32 - c.degree_bits), P)  # primitive_root_of_unity(degree_bits)
    zeta_next = ((g * zeta[0]) % P, (g * zeta[1]) % P)
    # OpeningSet::new, proof.rs:289-327
    cs_eval = pd.constants_sigmas_commitment.eval_polys(zeta)
    wires_eval = wires_commitment.eval_polys(zeta)
    zs_eval = zs_commitment.eval_polys(zeta)
    zs_next_eval = zs_commitment.eval_polys(zeta_next)
    quotient_eval = quotient_commitment.eval_polys(zeta)
    lap("construct the opening set")
    n_pre = c.num_constants + c.num_routed_wires
    constants, plonk_sigmas = cs_eval[: c.num_constants], cs_eval[c.num_constants: n_pre]
    plonk_zs, plonk_zs_next, partial_products = zs_eval[:nc], zs_next_eval[:nc], zs_eval[nc:]
    # observe_openings(to_fri_openings()), proof.rs:328-368, core/src/fri.rs:349-356
    for v in (constants, plonk_sigmas, wires_eval, plonk_zs, partial_products, quotient_eval, plonk_zs_next):
        challenger.observe_elements(np.asarray(v).reshape(-1))
    # prove_openings (fri/oracle.rs:320-358) on get_fri_instance(zeta) (circuit_data.rs:592-612)
    alpha = challenger.get_extension_challenge()
    oracles = [pd.constants_sigmas_commitment, wires_commitment, zs_commitment, quotient_commitment]
    zeta_polys = ([(oracles[0], i) for i in range(n_pre)] + [(oracles[1], i) for i in range(c.num_wires)] +
                  [(oracles[2], i) for i in range(nc * (1 + c.num_partial_products))] +
                  [(oracles[3], i) for i in range(nc * c.quotient_degree_factor)])
    next_polys = [(oracles[2], i) for i in range(nc)]
    batches = []
    for point, polys in ((zeta, zeta_polys), (zeta_next, next_polys)):
        terms, w = [], (1, 0)
        for o, i in polys:          # ReducingFactor::reduce_polys: sum_i alpha^i p_i, reducing.rs:63-72
            terms.append((o, i, w))
            w = ext_mul(w, alpha)
        batches.append(dict(point=point, shift=w, terms=terms))  # shift_poly: alpha^count, reducing.rs:94-97
    f = fri_from_openings(ctx, batches, c.degree_bits, fri.rate_bits, fri.cap_height)
    fri_bytes = fri_proof(ctx, oracles, f, challenger, fri.rate_bits, pd.reduction_arity_bits,
                          fri.proof_of_work_bits, fri.num_query_rounds)
    lap("compute opening proofs")
    f.free()
    # write_proof_with_public_inputs, serialization/mod.rs:2040-2079 (+ write_opening_set :1495-1508)
    out = bytearray()
    for cap in (wires_cap, zs_cap, quotient_cap):
        out += _ext_bytes(cap)
    for v in (constants, plonk_sigmas, wires_eval, plonk_zs, plonk_zs_next, partial_products, quotient_eval):
        out += _ext_bytes(v)
    out += fri_bytes
    out += np.array([len(public_inputs)] + public_inputs, dtype="<u8").tobytes()
    for b in (wires_commitment, zs_commitment, quotient_commitment):
        b.free()
    if timing is not None:
        timing.update(t)
    return bytes(out)
