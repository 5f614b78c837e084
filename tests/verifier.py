"""Test infrastructure: the reference's VERIFIER restated, run on serialised proofs.

    verify()                    verifier/src/plonk/verifier.rs:17-117 (transcript, plonk identity, FRI)
    get_challenges              verifier/src/plonk/get_challenges.rs:39-96, core/src/fri.rs:358-412
    verify_fri_proof            core/src/fri_verifier.rs:69-118 (PoW, query rounds)
    fri_combine_initial         core/src/fri_verifier.rs:132-174
    compute_evaluation          core/src/fri_verifier.rs:26-54 (interpolate the coset, evaluate at beta)
    verify_merkle_proof_to_cap  core/src/merkle_proofs.rs:42-97 (oracle.merkle_verify)
    proof layout                plonky2/src/util/serialization/mod.rs:1495-1508,1618-1667,2040-2079

Nothing above the Poseidon permutation has golden vectors in the reference, and the reference cannot
be compiled here; "the verifier accepts" is the reference's own acceptance criterion for those layers
(SURVEY.md section 4), so this restatement is what pins the prover -- the oracle's and the device's --
structurally: a proof is accepted only if commitments, openings, quotient, FRI folding, Merkle paths,
PoW and the Fiat-Shamir order are all consistent with each other.
"""
import numpy as np

import oracle
from oracle import P
from synth_circuit import Ext, verifier_plonk_identity

GEN = 14293326489335486720           # MULTIPLICATIVE_GROUP_GENERATOR = coset shift
TWO_ADIC = 7277203076849721926       # POWER_OF_TWO_GENERATOR (order 2^32)


def root_of_unity(bits):
    return pow(TWO_ADIC, 1 << (32 - bits), P)


def reverse_bits(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


class Reader:
    def __init__(self, data):
        self.d, self.p = data, 0

    def u64s(self, n):
        out = np.frombuffer(self.d, dtype="<u8", count=n, offset=self.p).copy()
        self.p += 8 * n
        return out

    def u8(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def ext(self, n):
        return self.u64s(2 * n).reshape(n, 2)

    def merkle_proof(self):   # write_merkle_proof: u8 length + digests
        k = self.u8()
        return self.u64s(4 * k).reshape(k, 4)


def parse_proof(proof, common, fri, arities, hiding=False):
    """-> dict with caps, openings, FRI proof parts, public inputs.  hiding: the prover oracles' leaves carry
    four salt elements each (validate_fri_proof_shape, core/src/fri_verifier... / fri/validate_shape.rs)."""
    c = common
    nc = c.num_challenges
    cap_words = 4 << fri.cap_height
    r = Reader(proof)
    out = {"caps": [r.u64s(cap_words).reshape(-1, 4) for _ in range(3)]}
    n_lk = nc * getattr(c, "num_lookup_polys", 0)
    sizes = [("constants", c.num_constants), ("plonk_sigmas", c.num_routed_wires), ("wires", c.num_wires),
             ("plonk_zs", nc), ("plonk_zs_next", nc), ("lookup_zs", n_lk), ("lookup_zs_next", n_lk),
             ("partial_products", nc * c.num_partial_products), ("quotient_polys", nc * c.quotient_degree_factor)]
    out["openings"] = {name: r.ext(k) for name, k in sizes}
    salt = 4 if hiding else 0
    leaf_lens = [c.num_constants + c.num_routed_wires, c.num_wires + salt,
                 nc * (1 + c.num_partial_products) + n_lk + salt, nc * c.quotient_degree_factor + salt]
    out["commit_caps"] = [r.u64s(cap_words).reshape(-1, 4) for _ in arities]
    rounds = []
    for _ in range(fri.num_query_rounds):
        initial = [(r.u64s(n), r.merkle_proof()) for n in leaf_lens]
        steps = [(r.ext(1 << a), r.merkle_proof()) for a in arities]
        rounds.append((initial, steps))
    out["query_rounds"] = rounds
    lde_bits = c.degree_bits + fri.rate_bits
    final_len = (1 << (lde_bits - sum(arities))) >> fri.rate_bits
    out["final_poly"] = r.ext(final_len)
    out["pow_witness"] = int(r.u64s(1)[0])
    npi = int(r.u64s(1)[0])
    out["public_inputs"] = [int(x) for x in r.u64s(npi)]
    assert r.p == len(proof), "trailing bytes"
    return out


def _e(v):
    return Ext(int(v[0]), int(v[1]))


def _reduce(values, alpha):
    """ReducingFactor::reduce: sum_i alpha^i v_i (core/src/reducing.rs:46-49)."""
    acc = Ext(0)
    for v in reversed(values):
        acc = acc * alpha + v
    return acc


def _interpolate(points, values, x):
    """Lagrange interpolation through (points[i], values[i]) evaluated at x (field/src/interpolation.rs)."""
    total = Ext(0)
    for i, (pi, vi) in enumerate(zip(points, values)):
        num, den = Ext(1), Ext(1)
        for j, pj in enumerate(points):
            if j != i:
                num = num * (x - pj)
                den = den * (Ext(pi) - pj)
        total = total + vi * num * den.inv()
    return total


def verify(proof, common, fri, constants_sigmas_cap, circuit_digest, hiding=False):
    """-> None if the proof is accepted, else a string naming the failed check.  hiding = config.zero_knowledge:
    leaf_hiding is observed as 1 and the salt is stripped from the blinded oracles' leaves
    (unsalted_eval, core/src/fri_verifier.rs:222-228)."""
    c = common
    nc = c.num_challenges
    n = 1 << c.degree_bits
    arities = oracle.fri_reduction_arity_bits(c.degree_bits, fri.rate_bits, fri.cap_height, fri.arity_bits,
                                              fri.final_poly_bits)
    try:
        pr = parse_proof(proof, c, fri, arities, hiding)
    except Exception as e:  # validate_fri_proof_shape
        return "malformed proof: %r" % (e,)
    op = pr["openings"]
    # ---- challenges (get_challenges.rs:39-96) ----
    ch = oracle.Challenger()
    ch.observe([fri.rate_bits, fri.cap_height, fri.proof_of_work_bits, 1, fri.arity_bits, fri.final_poly_bits,
                fri.num_query_rounds, int(hiding), c.degree_bits] + list(arities))
    ch.observe(np.asarray(circuit_digest, dtype=np.uint64))
    pih = oracle.hash_no_pad(np.array(pr["public_inputs"], dtype=np.uint64))
    ch.observe(pih)
    ch.observe(pr["caps"][0].reshape(-1))
    betas = [ch.get_challenge() for _ in range(nc)]
    gammas = [ch.get_challenge() for _ in range(nc)]
    has_lookup = getattr(c, "num_lookup_polys", 0) != 0
    # get_challenges.rs:53-68: four lookup challenges per challenge, the first 2 nc of them are the betas and gammas
    deltas = betas + gammas + [ch.get_challenge() for _ in range(2 * nc)] if has_lookup else None
    ch.observe(pr["caps"][1].reshape(-1))
    alphas = [ch.get_challenge() for _ in range(nc)]
    ch.observe(pr["caps"][2].reshape(-1))
    zeta = ch.get_extension_challenge()
    # observe_openings(to_fri_openings()), plonky2/src/plonk/proof.rs:328-368
    for name in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys", "lookup_zs",
                 "plonk_zs_next", "lookup_zs_next"):
        ch.observe(op[name].reshape(-1))
    fri_alpha = _e(ch.get_extension_challenge())
    fri_betas = []
    for cap in pr["commit_caps"]:
        ch.observe(cap.reshape(-1))
        fri_betas.append(_e(ch.get_extension_challenge()))
    ch.observe(pr["final_poly"].reshape(-1))
    ch.observe([pr["pow_witness"]])
    pow_response = ch.get_challenge()
    lde_bits = c.degree_bits + fri.rate_bits
    N = 1 << lde_bits
    x_indices = [ch.get_challenge() % N for _ in range(fri.num_query_rounds)]
    # ---- plonk identity at zeta (verifier.rs:60-100) ----
    if not verifier_plonk_identity(c, op, zeta, betas, gammas, alphas, pih, deltas):
        return "vanishing(zeta) != Z_H(zeta) * quotient(zeta)"
    # ---- FRI (fri_verifier.rs:69-118) ----
    if pow_response >> (64 - fri.proof_of_work_bits) if fri.proof_of_work_bits else 0:
        return "Invalid proof of work witness."
    z = _e(zeta)
    g = root_of_unity(c.degree_bits)
    points = [z, z * g]
    # openings per batch, in the order of get_fri_instance (circuit_data.rs:592-612, 741-749)
    batch_open = [[_e(v) for name in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products",
                                      "quotient_polys", "lookup_zs") for v in op[name]],
                  [_e(v) for name in ("plonk_zs_next", "lookup_zs_next") for v in op[name]]]
    n_zs = nc * (1 + c.num_partial_products)
    reduced_openings = [_reduce(vals, fri_alpha) for vals in batch_open]
    caps = [np.asarray(constants_sigmas_cap, dtype=np.uint64).reshape(-1, 4)] + pr["caps"]
    w_lde = root_of_unity(lde_bits)
    for x_index, (initial, steps) in zip(x_indices, pr["query_rounds"]):
        for (evals, path), cap in zip(initial, caps):
            if not oracle.merkle_verify(evals, x_index, cap, path):
                return "initial tree Merkle proof"
        subgroup_x = GEN * pow(w_lde, reverse_bits(x_index, lde_bits), P) % P
        # fri_combine_initial: evaluations at x of every opened polynomial, from the leaves
        # the blinded oracles (wires, zs, quotient: core/src/plonk_common.rs:20-35) lose their salt here
        leaf = [[Ext(int(v)) for v in (evals[: len(evals) - 4] if hiding and k else evals)]
                for k, (evals, _) in enumerate(initial)]
        # fri_all_openings / fri_next_batch_openings (circuit_data.rs:711-747): the lookup polynomials of oracle 2
        # come after the quotient polynomials in the zeta batch, and after the Z's in the zeta_next batch
        batch_evals = [leaf[0] + leaf[1] + leaf[2][:n_zs] + leaf[3] + leaf[2][n_zs:], leaf[2][:nc] + leaf[2][n_zs:]]
        total = Ext(0)
        for evals, ro, pt in zip(batch_evals, reduced_openings, points):
            num = _reduce(evals, fri_alpha) - ro
            den = Ext(subgroup_x) - pt
            total = total * fri_alpha.pow(len(evals)) + num * den.inv()
        old_eval = total
        xi = x_index
        for k, a in enumerate(arities):
            evals, path = steps[k]
            arity = 1 << a
            coset_index, within = xi >> a, xi & (arity - 1)
            if not (_e(evals[within]) == old_eval):
                return "FRI consistency at reduction %d" % k
            # compute_evaluation, fri_verifier.rs:26-54
            ga = root_of_unity(a)
            ev = [None] * arity
            for i in range(arity):
                ev[reverse_bits(i, a)] = _e(evals[i])       # undo the bit-reversed leaf order
            start = subgroup_x * pow(ga, arity - reverse_bits(within, a), P) % P
            pts = [start * pow(ga, i, P) % P for i in range(arity)]
            old_eval = _interpolate(pts, ev, fri_betas[k])
            if not oracle.merkle_verify(evals.reshape(-1), coset_index, pr["commit_caps"][k], path):
                return "commit-phase Merkle proof %d" % k
            subgroup_x = pow(subgroup_x, arity, P)
            xi = coset_index
        fp = Ext(0)
        for coef in reversed(pr["final_poly"]):               # final_poly.eval(subgroup_x)
            fp = fp * subgroup_x + _e(coef)
        if not (fp == old_eval):
            return "Final polynomial evaluation is invalid."
    return None


# ---------------------------------------------------------------------------------------------
# batch FRI: verify_batch_fri_proof (plonky2/src/batch_fri/verifier.rs:23-247)
# ---------------------------------------------------------------------------------------------
def parse_fri_proof(data, oracle_leaf_lens, lde_bits, rate_bits, cap_height, arities, num_query_rounds):
    """write_fri_proof's layout (serialization/mod.rs:1654-1667) -> dict."""
    r = Reader(data)
    cap_words = 4 << cap_height
    out = {"commit_caps": [r.u64s(cap_words).reshape(-1, 4) for _ in arities]}
    rounds = []
    for _ in range(num_query_rounds):
        initial = [(r.u64s(n), r.merkle_proof()) for n in oracle_leaf_lens]
        steps = [(r.ext(1 << a), r.merkle_proof()) for a in arities]
        rounds.append((initial, steps))
    out["query_rounds"] = rounds
    out["final_poly"] = r.ext((1 << (lde_bits - sum(arities))) >> rate_bits)
    out["pow_witness"] = int(r.u64s(1)[0])
    assert r.p == len(data), "trailing bytes"
    return out


def _coefficient(coeff, point):
    if coeff == "one" or coeff == ("one",):
        return Ext(1)
    if coeff[0] == "point_power":
        return point.pow(coeff[1])
    return Ext(int(coeff[1][0]), int(coeff[1][1]))


def verify_batch_fri_proof(degree_bits, instances, openings, challenger, initial_caps, proof, rate_bits, cap_height,
                           arities, pow_bits, num_query_rounds):
    """-> None if accepted, else the failed check.  degree_bits: polynomial degrees, tallest first;
    instances[i] = dict(oracles=[num_polys per oracle], batches=[dict(point=(a, b), openings=[expression, ...])]);
    openings[i][b] = list of claimed (c0, c1) values of instance i's batch b; challenger: an oracle.pyref
    Challenger in the state the prover's had right before prove_openings; initial_caps[o]: cap of oracle o."""
    from oracle import pyref

    lde = [d + rate_bits for d in degree_bits]
    leaf_lens = [sum(inst["oracles"][o] for inst in instances) for o in range(len(initial_caps))]
    try:
        pr = parse_fri_proof(proof, leaf_lens, lde[0], rate_bits, cap_height, arities, num_query_rounds)
    except Exception as e:
        return "malformed proof: %r" % (e,)
    # fri_challenges (core/src/fri.rs:358-420)
    ch = challenger
    fri_alpha = _e(ch.get_extension_challenge())
    fri_betas = []
    for cap in pr["commit_caps"]:
        ch.observe([int(v) for v in cap.reshape(-1)])
        fri_betas.append(_e(ch.get_extension_challenge()))
    ch.observe([int(v) for v in pr["final_poly"].reshape(-1)])
    ch.observe([pr["pow_witness"]])
    pow_response = ch.get_challenge()
    N = 1 << lde[0]
    x_indices = [ch.get_challenge() % N for _ in range(num_query_rounds)]
    if pow_bits and pow_response >> (64 - pow_bits):
        return "Invalid proof of work witness."
    # PrecomputedReducedOpenings::from_os_and_alpha
    reduced = [[_reduce([_e(v) for v in vals], fri_alpha) for vals in inst_open] for inst_open in openings]

    def combine_initial(index, initial, subgroup_x):        # batch_fri_combine_initial, verifier.rs:112-152
        total = Ext(0)
        for b, ro in zip(instances[index]["batches"], reduced[index]):
            pt = _e(b["point"])
            evals = []
            for expr in b["openings"]:
                acc = Ext(0)
                for oi, pi, coeff in expr:
                    acc = acc + _coefficient(coeff, pt) * Ext(int(initial[oi][0][pi]))
                evals.append(acc)
            num = _reduce(evals, fri_alpha) - ro
            total = total * fri_alpha.pow(len(evals)) + num * (Ext(subgroup_x) - pt).inv()
        return total

    for x_index, (initial, steps) in zip(x_indices, pr["query_rounds"]):
        # batch_fri_verify_initial_proof: per oracle, the rows of every degree under one batch Merkle proof
        for o, ((evals, path), cap) in enumerate(zip(initial, initial_caps)):
            rows, pos = [], 0
            for inst in instances:
                k = inst["oracles"][o]
                rows.append([int(v) for v in evals[pos:pos + k]])
                pos += k
            if not pyref.batch_merkle_verify(rows, lde, x_index, [[int(v) for v in d] for d in cap],
                                             [[int(v) for v in d] for d in path]):
                return "initial batch Merkle proof"
        n = lde[0]
        subgroup_x = GEN * pow(root_of_unity(n), reverse_bits(x_index, n), P) % P
        batch_index = 0
        old_eval = combine_initial(batch_index, initial, subgroup_x)
        batch_index += 1
        xi = x_index
        for k, a in enumerate(arities):
            evals, path = steps[k]
            arity = 1 << a
            coset_index, within = xi >> a, xi & (arity - 1)
            if not (_e(evals[within]) == old_eval):
                return "FRI consistency at reduction %d" % k
            ga = root_of_unity(a)
            ev = [None] * arity
            for i in range(arity):
                ev[reverse_bits(i, a)] = _e(evals[i])
            start = subgroup_x * pow(ga, arity - reverse_bits(within, a), P) % P
            pts = [start * pow(ga, i, P) % P for i in range(arity)]
            old_eval = _interpolate(pts, ev, fri_betas[k])
            if not pyref.merkle_verify([int(v) for v in evals.reshape(-1)], coset_index,
                                       [[int(v) for v in d] for d in pr["commit_caps"][k]],
                                       [[int(v) for v in d] for d in path]):
                return "commit-phase Merkle proof %d" % k
            subgroup_x = pow(subgroup_x, arity, P)
            xi = coset_index
            n -= a
            if batch_index < len(lde) and n == lde[batch_index]:
                x_init = GEN * pow(root_of_unity(n), reverse_bits(xi, n), P) % P
                old_eval = old_eval * fri_betas[k] + combine_initial(batch_index, initial, x_init)
                batch_index += 1
        if batch_index != len(instances):
            return "Wrong number of folded instances."
        fp = Ext(0)
        for coef in reversed(pr["final_poly"]):
            fp = fp * subgroup_x + _e(coef)
        if not (fp == old_eval):
            return "Final polynomial evaluation is invalid."
    return None
