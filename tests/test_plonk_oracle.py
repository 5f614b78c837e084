"""CPU tests of the plonk oracle (permutation argument, quotient) and of the host-side constraint
program compiler.  No GPU.  The reference has no golden vectors for these layers and cannot be
built here, so the oracle is pinned STRUCTURALLY: the restated verifier identity
vanishing(x) == Z_H(x) * quotient(x) (verifier/src/plonk/verifier.rs:86-100) must hold at random
points for a satisfying witness and must fail for a corrupted one."""
import numpy as np
import pytest

import oracle
from qp_plonky2_b200 import plonk

import gate_witness
from synth_circuit import SynthCircuit, verifier_identity_holds

P = oracle.P


def run_program(code, pool, n_regs, wires, consts, pih, alphas):
    """Python big-integer interpreter of the constraint program (same semantics as
    quotient::quotient_kernel): returns G[a] = sum_gates filter * sum_k alpha^k c_k."""
    r = [0] * n_regs
    nc = len(alphas)
    G, h = [0] * nc, [0] * nc
    in_flight = []   # asynchronous column loads: (register, value); at most 3 stay in flight after an issue

    def load(dst, v):
        r[dst] = None              # unreadable until it lands
        in_flight.append((dst, v))
        while len(in_flight) > 3:
            d, x = in_flight.pop(0)
            assert r[d] is None, "register overwritten while its load was in flight"
            r[d] = x

    for ins in code:
        op, dst, a, b, c = plonk.decode_word(ins)
        if op == plonk.OP_END:     # segments are self-contained: no register survives an END
            assert not in_flight and h == [0] * nc
            r = [None] * n_regs
        elif op == plonk.OP_NATIVE_POSEIDON:
            # a PoseidonGate handed to the device's native evaluator: dst = group end, a = gate index,
            # b = group start, c = selector column | many_selectors << 16 (quotient.cuh)
            sel = int(consts[c & 0xFFFF])
            f = 1
            for j in range(b, dst):
                if j != a:
                    f = f * (j - sel) % P
            if c >> 16:
                f = f * (plonk.UNUSED_SELECTOR - sel) % P
            cons = plonk.PoseidonGate().eval_unfiltered(None, lambda k: gate_witness.ModP(wires[k]), None)
            for q in range(nc):
                acc = sum(int(v) * pow(int(alphas[q]), k, P) for k, v in enumerate(cons)) % P
                G[q] = (G[q] + f * acc) % P
        elif op == plonk.OP_WAIT:
            for d, x in in_flight:
                assert r[d] is None
                r[d] = x
            in_flight.clear()
        elif op == plonk.OP_LDW:
            load(dst, int(wires[c]))
        elif op == plonk.OP_LDK:
            load(dst, int(consts[c]))
        elif op == plonk.OP_LDP:
            r[dst] = int(pih[c])
        elif op == plonk.OP_LDI:
            r[dst] = int(pool[c])
        elif op == plonk.OP_ADD:
            r[dst] = (r[a] + r[b]) % P
        elif op == plonk.OP_SUB:
            r[dst] = (r[a] - r[b]) % P
        elif op == plonk.OP_MUL:
            r[dst] = r[a] * r[b] % P
        elif op == plonk.OP_MULI:
            r[dst] = r[a] * int(pool[c]) % P
        elif op == plonk.OP_ADDI:
            r[dst] = (r[a] + int(pool[c])) % P
        elif op == plonk.OP_FMAI:
            r[dst] = (r[a] * int(pool[c]) + r[b]) % P
        elif op == plonk.OP_EMIT:   # constraint index in c
            h = [(h[k] + r[a] * pow(int(alphas[k]), c, P)) % P for k in range(nc)]
        elif op == plonk.OP_GATE:
            G = [(G[k] + r[a] * h[k]) % P for k in range(nc)]
            h = [0] * nc
    return G


@pytest.mark.parametrize("qdf", [8, 4])
def test_selectors_info_matches_reference_rule(qdf):
    sc = SynthCircuit(6, seed=3, quotient_degree_factor=qdf)
    c = sc.common
    # gates sorted by (degree, id): circuit_builder.rs:1177-1179
    assert [g.id().split(" ")[0] for g in c.gates] == ["NoopGate", "ConstantGate", "PublicInputGate", "ArithmeticGate"]
    if qdf == 8:   # 3 + 4 - 1 <= 9: one selector (selectors.rs:112-131)
        assert c.groups == [(0, 4)] and c.num_selectors == 1
    else:          # greedy groups with size + degree < 5 (selectors.rs:140-150)
        assert c.groups == [(0, 3), (3, 4)] and c.selector_indices == [0, 0, 0, 1]
    assert c.num_partial_products == -(-80 // qdf) - 1


@pytest.mark.parametrize("native", [False, True, "dag"])
@pytest.mark.parametrize("qdf,poseidon,extra,rec", [(8, False, False, False), (4, False, False, False),
                                                    (8, True, False, False), (8, True, True, False),
                                                    (4, False, True, False), (8, True, True, True),
                                                    (8, False, False, True)])
def test_constraint_program_matches_oracle_gate_evaluation(qdf, poseidon, native, extra, rec):
    """The compiled program and the oracle's hand-written gate evaluators agree on random
    (non-satisfying) inputs: compare the full vanishing value with the permutation terms zeroed
    out by Z = partial products = 0... simpler: both sides computed in full."""
    sc = SynthCircuit(5, seed=5, quotient_degree_factor=qdf, poseidon=poseidon, extra_gates=extra,
                      recursion_gates=rec)
    c = sc.common
    if native == "dag":   # this module's recording of the gates through the native compiler (qp_program_from_dag)
        code, pool, n_regs = c.constraint_program(native_compile=True)
        assert any(int(w) & 0xFF == plonk.OP_FMAI for w in code) or not (poseidon or extra or rec)
    elif native:   # qp-plonky2_b200/host/plonk_host.cpp, the compiler whose output the device runs
        prog = plonk.native_constraint_program(c.gates, qdf + 1)
        code, pool, n_regs = prog["code"], prog["pool"], prog["n_regs"]
        assert prog["selector_indices"] == c.selector_indices and prog["groups"] == c.groups
        assert prog["order"] == list(range(len(c.gates)))   # already sorted
        assert prog["num_gate_constraints"] == c.num_gate_constraints
    else:
        code, pool, n_regs = c.constraint_program()
    if poseidon and extra and rec:   # the 14 gates of a recursive verifier circuit: four selector groups
        assert c.groups == [(0, 7), (7, 11), (11, 13), (13, 14)] and c.num_constants == 4 + 2
        assert [g.id().split(" ")[0].split("(")[0] for g in c.gates] == [
            "NoopGate", "ConstantGate", "PoseidonMdsGate", "PublicInputGate", "BaseSumGate", "ReducingExtensionGate",
            "ReducingGate", "ArithmeticExtensionGate", "ArithmeticGate", "MulExtensionGate", "ExponentiationGate",
            "RandomAccessGate", "CosetInterpolationGate", "PoseidonGate"]
    if poseidon and not extra:   # degree 7 forces a second selector group (selectors.rs:140-150)
        assert c.groups == [(0, 4), (4, 5)] and c.num_gate_constraints == 123
    if poseidon and extra and not rec:       # Noop, Constant, PublicInput, BaseSum, ArithmeticExtension, Arithmetic | MulExtension, Poseidon
        assert c.groups == [(0, 6), (6, 8)]
    rng = np.random.default_rng(7)
    for trial in range(4 * len(c.gates) if rec else 8):
        wires = oracle.rand_felts((c.num_wires,), 100 + trial)
        consts = oracle.rand_felts((c.num_constants,), 200 + trial)
        # a selector value that is a real gate index some of the time
        consts[0] = rng.integers(0, c.groups[0][1])
        if len(c.groups) > 1 and trial % 2:
            consts[0], consts[1] = plonk.UNUSED_SELECTOR, rng.integers(c.groups[1][0], c.groups[1][1])
        if rec:   # every gate in turn is the selected one (its group's selector = its index, the others unused)
            gi = trial % len(c.gates)
            consts[:c.num_selectors] = plonk.UNUSED_SELECTOR
            consts[c.selector_indices[gi]] = gi
        pih = oracle.rand_felts((4,), 300 + trial)
        alphas = oracle.rand_felts((2,), 400 + trial)
        nc, np_ = c.num_challenges, c.num_partial_products
        zeros = np.zeros(nc * np_, dtype=np.uint64)
        ones = np.ones(nc, dtype=np.uint64)
        # with Z(x) = Z(gx) = 1... the permutation terms do not vanish in general, so evaluate
        # them separately through the oracle with an EMPTY gate list and subtract.
        full = sc.oracle_circuit.eval_vanishing_poly_base(
            12345, consts, wires, ones, ones, zeros, oracle.rand_felts((80,), 9), [3, 4], [5, 6], alphas, pih)
        empty = oracle.Circuit(c.degree_bits, c.quotient_degree_bits, nc, c.num_routed_wires, c.num_wires,
                               c.num_constants, np_, qdf, c.num_selectors, [], c.k_is)
        perm = empty.eval_vanishing_poly_base(
            12345, consts, wires, ones, ones, zeros, oracle.rand_felts((80,), 9), [3, 4], [5, 6], alphas, pih)
        G = run_program(code, pool, n_regs, wires, consts, pih, alphas)
        base = nc + nc * (np_ + 1)
        for a in range(nc):
            want = (int(full[a]) - int(perm[a])) % P
            assert G[a] * pow(int(alphas[a]), base, P) % P == want


@pytest.mark.parametrize("degree_bits,qdf,poseidon", [(7, 8, False), (6, 4, False), (6, 8, True)])
def test_oracle_quotient_satisfies_verifier_identity(degree_bits, qdf, poseidon):
    sc = SynthCircuit(degree_bits, seed=11, quotient_degree_factor=qdf, poseidon=poseidon)
    c = sc.common
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    wb = oracle.PolynomialBatch.from_values(sc.wires, c.rate_bits, c.cap_height)
    betas, gammas, alphas = (oracle.rand_felts((2,), s) for s in (21, 22, 23))
    zs = sc.oracle_circuit.partial_products_and_zs(sc.wires, sc.sigmas, betas, gammas)
    # Z starts at 1 and the grand product closes (the witness satisfies the copy constraints)
    assert (zs[:2, 0] == 1).all()
    zb = oracle.PolynomialBatch.from_values(zs, c.rate_bits, c.cap_height)
    q = sc.oracle_circuit.compute_quotient_polys(c.rate_bits, cs.leaves, wb.leaves, zb.leaves, betas, gammas,
                                                 alphas, sc.public_inputs_hash)
    assert q.shape == (2, sc.n << c.quotient_degree_bits)
    for x0 in (3, 0x123456789ABCDEF, P - 2):
        assert verifier_identity_holds(sc, cs.polynomials, wb.polynomials, zb.polynomials, q, betas, gammas, alphas, x0)
    # a corrupted quotient fails
    q2 = q.copy()
    q2[0, 5] ^= 1
    assert not verifier_identity_holds(sc, cs.polynomials, wb.polynomials, zb.polynomials, q2, betas, gammas, alphas, 3)


def test_oracle_detects_unsatisfied_witness():
    """With a broken gate constraint the 'quotient' is not a polynomial of the right degree any
    more, so the identity fails at a random point."""
    sc = SynthCircuit(6, seed=13)
    c = sc.common
    arith = np.nonzero(sc.row_gate == 3)[0]
    sc.wires[3, arith[0]] ^= 1  # output of op 0 in one arithmetic row
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    wb = oracle.PolynomialBatch.from_values(sc.wires, c.rate_bits, c.cap_height)
    betas, gammas, alphas = (oracle.rand_felts((2,), s) for s in (31, 32, 33))
    zs = sc.oracle_circuit.partial_products_and_zs(sc.wires, sc.sigmas, betas, gammas)
    zb = oracle.PolynomialBatch.from_values(zs, c.rate_bits, c.cap_height)
    q = sc.oracle_circuit.compute_quotient_polys(c.rate_bits, cs.leaves, wb.leaves, zb.leaves, betas, gammas,
                                                 alphas, sc.public_inputs_hash)
    assert not verifier_identity_holds(sc, cs.polynomials, wb.polynomials, zb.polynomials, q, betas, gammas, alphas, 77)


def _oracle_prove(sc, **kw):
    from oracle import prover as oprover

    c = sc.common
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    return oprover.prove(sc.oracle_circuit, cs, c.num_constants, sc.wires, sc.sigmas, sc.public_inputs,
                         degree_bits=c.degree_bits, num_wires=c.num_wires, num_routed_wires=c.num_routed_wires,
                         num_challenges=c.num_challenges, quotient_degree_factor=c.quotient_degree_factor,
                         num_partial_products=c.num_partial_products, rate_bits=c.rate_bits,
                         cap_height=c.cap_height, **kw)


@pytest.mark.parametrize("degree_bits,qdf,poseidon", [(6, 8, False), (7, 4, False), (6, 8, True)])
def test_oracle_proof_openings_satisfy_the_verifier(degree_bits, qdf, poseidon):
    """Full oracle prove(): the opening set inside the proof passes the verifier's algebraic check
    at the Fiat-Shamir point zeta in F_p^2 (verifier/src/plonk/verifier.rs:60-100)."""
    from synth_circuit import verifier_plonk_identity

    sc = SynthCircuit(degree_bits, seed=41, quotient_degree_factor=qdf, poseidon=poseidon)
    proof, info = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=4)
    assert verifier_plonk_identity(sc.common, info["openings"], info["zeta"], info["betas"], info["gammas"],
                                   info["alphas"], info["pih"])
    # tampering with one opening breaks it
    bad = dict(info["openings"])
    bad["wires"] = info["openings"]["wires"].copy()
    bad["wires"][3, 0] ^= 1
    assert not verifier_plonk_identity(sc.common, bad, info["zeta"], info["betas"], info["gammas"], info["alphas"],
                                       info["pih"])
    # layout: 3 caps, opening set, FRI proof, public inputs (serialization/mod.rs:2040-2079)
    c = sc.common
    n_open = (c.num_constants + c.num_routed_wires + c.num_wires + 2 * c.num_challenges +
              c.num_challenges * c.num_partial_products + c.num_challenges * c.quotient_degree_factor)
    assert proof[: 3 * 16 * 32] != b"" and len(proof) > 3 * 16 * 32 + n_open * 16
    tail = np.frombuffer(proof[-8 * (1 + len(sc.public_inputs)):], dtype="<u8")
    assert int(tail[0]) == len(sc.public_inputs) and [int(x) for x in tail[1:]] == sc.public_inputs
    # deterministic
    assert _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=4)[0] == proof


def test_native_program_sorts_gates_like_the_builder():
    """circuit_builder.rs:1177-1179: by (degree, id) whatever order the caller lists them in."""
    gates = [plonk.PoseidonGate(), plonk.ArithmeticGate(20), plonk.PublicInputGate(), plonk.NoopGate(), plonk.ConstantGate(2)]
    prog = plonk.native_constraint_program(gates, 9)
    assert [gates[i].id().split(" ")[0].split("(")[0] for i in prog["order"]] == [
        "NoopGate", "ConstantGate", "PublicInputGate", "ArithmeticGate", "PoseidonGate"]
    assert prog["groups"] == [(0, 4), (4, 5)] and prog["selector_indices"] == [0, 0, 0, 0, 1]
    with pytest.raises(ValueError):   # "... has too high degree. Consider increasing `quotient_degree_factor`."
        plonk.native_constraint_program(gates, 7)
    # the whole gate set of a recursive verifier circuit, listed in a scrambled order
    gates = [plonk.CosetInterpolationGate.with_max_degree(4, 8), plonk.PoseidonGate(), plonk.ReducingGate(46),
             plonk.ExponentiationGate(70), plonk.ArithmeticGate(20), plonk.PublicInputGate(), plonk.PoseidonMdsGate(),
             plonk.RandomAccessGate.new_from_config(143, 80, 4), plonk.NoopGate(), plonk.MulExtensionGate(13),
             plonk.ReducingExtensionGate(34), plonk.ConstantGate(2), plonk.BaseSumGate2(63),
             plonk.ArithmeticExtensionGate(10)]
    prog = plonk.native_constraint_program(gates, 9)
    want = plonk.sort_gates(gates)
    assert [gates[i].id() for i in prog["order"]] == [g.id() for g in want]
    assert prog["groups"] == [(0, 7), (7, 11), (11, 13), (13, 14)]
    assert prog["num_gate_constants"] == 2 and prog["num_gate_constraints"] == 123
    # the program is cut into self-contained segments of similar cost (PoseidonGate, half of the work,
    # between constraints) and loads with distant uses are repeated instead of pinning registers:
    # BaseSumGate alone would need 65
    ends = [i for i, w in enumerate(prog["code"]) if int(w) & 0xFF == plonk.OP_END]
    assert list(prog["segments"]) == [0] + [e + 1 for e in ends] and len(ends) >= 3
    # PoseidonGate is not compiled: one word, a segment of its own, for the device's native evaluator
    op, dst, a, b, c = plonk.decode_word(prog["code"][-1])
    assert (op, dst, a, b, c) == (plonk.OP_NATIVE_POSEIDON, 14, 13, 13, 3 | 1 << 16)
    sizes = np.diff([0] + ends)
    assert sizes.max() < 3 * sizes.min() and prog["n_regs"] <= 40
    g = plonk.RandomAccessGate.new_from_config(143, 80, 4)      # random_access.rs:58-76 under the standard config
    assert (g.num_copies, g.num_extra_constants, g.degree, g.num_constraints) == (4, 2, 5, 26)
    g = plonk.CosetInterpolationGate.with_max_degree(4, 8)      # coset_interpolation.rs:49-75
    assert (g.degree, g.num_intermediates, g.num_wires(), g.num_constraints) == (6, 2, 48, 13)
    assert plonk.ExponentiationGate.new_from_config(143, 80).num_power_bits == 70


def test_host_hash_no_pad_and_circuit_digest_match_oracle():
    import ctypes as C
    import qp_plonky2_b200 as qp
    from oracle import prover as oprover

    L = qp.lib()
    for n in (0, 1, 7, 8, 9, 16, 71):
        x = oracle.rand_felts((n,), 900 + n)
        out = np.zeros(4, dtype=np.uint64)
        L.qp_hash_no_pad(x.ctypes.data if n else None, n, out.ctypes.data)
        assert (out == oracle.hash_no_pad(x)).all(), n
    cap = oracle.rand_felts((16, 4), 3)
    out = np.zeros(4, dtype=np.uint64)
    L.qp_circuit_digest(cap.ctypes.data, 16, 13, out.ctypes.data)
    assert (out == oprover.circuit_digest(cap, 13)).all()


class _Fri:
    def __init__(self, rate_bits=3, cap_height=4, proof_of_work_bits=16, arity_bits=4, final_poly_bits=5,
                 num_query_rounds=28):
        self.rate_bits, self.cap_height, self.proof_of_work_bits = rate_bits, cap_height, proof_of_work_bits
        self.arity_bits, self.final_poly_bits, self.num_query_rounds = arity_bits, final_poly_bits, num_query_rounds


@pytest.mark.parametrize("degree_bits,qdf,poseidon,pow_bits,queries,extra,rec", [
    (6, 8, False, 6, 5, False, False), (8, 8, True, 10, 9, False, False), (7, 4, False, 5, 4, False, False),
    (7, 8, True, 6, 5, True, False), (7, 8, True, 6, 5, True, True)])
def test_restated_verifier_accepts_oracle_proofs_and_rejects_tampering(degree_bits, qdf, poseidon, pow_bits, queries,
                                                                      extra, rec):
    """The reference's acceptance criterion for everything above the permutation:
    the full verifier -- transcript, plonk identity, PoW, FRI query rounds with every Merkle path,
    folding consistency, final polynomial -- accepts the oracle's proof and rejects a flipped bit in
    any part of it."""
    from oracle import prover as oprover
    import verifier

    sc = SynthCircuit(degree_bits, seed=91, quotient_degree_factor=qdf, poseidon=poseidon, extra_gates=extra,
                      recursion_gates=rec)
    c = sc.common
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    proof, info = _oracle_prove(sc, proof_of_work_bits=pow_bits, num_query_rounds=queries)
    fri = _Fri(c.rate_bits, c.cap_height, pow_bits, 4, 5, queries)
    digest = oprover.circuit_digest(cs.cap, degree_bits)
    assert verifier.verify(proof, c, fri, cs.cap, digest) is None
    # tampering: one bit in each region of the proof
    cap_bytes = 3 * (4 << c.cap_height) * 8
    n_open = (c.num_constants + c.num_routed_wires + c.num_wires + c.num_challenges * (2 + c.num_partial_products) +
              c.num_challenges * c.quotient_degree_factor)
    spots = {"wires cap": 5, "quotient cap": cap_bytes - 9, "an opening": cap_bytes + 16 * 7,
             "commit-phase cap": cap_bytes + 16 * n_open + 3, "a query leaf": cap_bytes + 16 * n_open + 4000,
             "final poly / pow": len(proof) - 8 * (1 + len(sc.public_inputs)) - 5, "a public input": len(proof) - 3}
    for what, at in spots.items():
        bad = bytearray(proof)
        bad[at] ^= 1
        assert verifier.verify(bytes(bad), c, fri, cs.cap, digest) is not None, what
    # and a wrong circuit digest changes every challenge
    assert verifier.verify(proof, c, fri, cs.cap, digest ^ np.uint64(1)) is not None


@pytest.mark.parametrize("degree_bits,qdf,poseidon", [(6, 8, False), (7, 8, True)])
def test_restated_verifier_accepts_zero_knowledge_oracle_proofs(degree_bits, qdf, poseidon):
    """config.zero_knowledge (plonky2/src/plonk/prover.rs:210,280,328): salted wires / Z / quotient oracles and
    leaf_hiding = 1.  The restated verifier accepts the oracle's zk proof in hiding mode (stripping the salt,
    core/src/fri_verifier.rs:222-228), refuses it in non-hiding mode, refuses a plain proof in hiding mode, and
    a different salt gives different commitments but the same openings-independent acceptance."""
    from oracle import prover as oprover
    import verifier

    sc = SynthCircuit(degree_bits, seed=92, quotient_degree_factor=qdf, poseidon=poseidon)
    c = sc.common
    N = (1 << degree_bits) << c.rate_bits
    salts = [oracle.rand_felts((4, N), 700 + k) for k in range(3)]
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    zk_proof, _ = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=5, salts=salts)
    plain, _ = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=5)
    fri = _Fri(c.rate_bits, c.cap_height, 6, 4, 5, 5)
    digest = oprover.circuit_digest(cs.cap, degree_bits)
    assert len(zk_proof) == len(plain) + 5 * 3 * 4 * 8          # 3 salted leaves per query round
    assert verifier.verify(zk_proof, c, fri, cs.cap, digest, hiding=True) is None
    assert verifier.verify(zk_proof, c, fri, cs.cap, digest, hiding=False) is not None
    assert verifier.verify(plain, c, fri, cs.cap, digest, hiding=True) is not None
    other, _ = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=5,
                             salts=[oracle.rand_felts((4, N), 800 + k) for k in range(3)])
    assert other[:64] != zk_proof[:64]                            # the wires cap depends on the salt
    assert verifier.verify(other, c, fri, cs.cap, digest, hiding=True) is None
    bad = bytearray(zk_proof)
    bad[len(bad) // 2] ^= 2
    assert verifier.verify(bytes(bad), c, fri, cs.cap, digest, hiding=True) is not None


# ---------------------------------------------------------------------------------------------
# lookup argument (gates/lookup.rs, gates/lookup_table.rs, prover.rs:489-636, vanishing_poly.rs:212-381)
# ---------------------------------------------------------------------------------------------
def test_keccak256_known_answers():
    """keccak_hash::keccak = Keccak-256 with the original padding (the lut_hash inside the lookup gates' ids decides
    the gate ORDER, hence selectors and constraint indices): the published vectors of Keccak-256, for the host
    library's implementation and for the Python mirror's, plus a multi-block message."""
    import ctypes as C
    import qp_plonky2_b200 as qp
    from qp_plonky2_b200 import plonk

    def native(data):
        out = (C.c_uint8 * 32)()
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        qp.lib().qp_keccak256(buf, len(data), out)
        return bytes(out)

    kats = {b"": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
            b"abc": "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"}
    for msg, want in kats.items():
        assert plonk.keccak256(msg).hex() == want
        assert native(msg).hex() == want
    long = bytes(range(256)) * 3   # 768 bytes: six rate blocks
    assert native(long) == plonk.keccak256(long)
    assert native(bytes(135)) == plonk.keccak256(bytes(135)) and native(bytes(136)) == plonk.keccak256(bytes(136))


def _lookup_polys_bigint(c, wires, deltas):
    """compute_lookup_polys (prover.rs:489-636) with Python integers, written from the reference loop by loop."""
    n, nr = 1 << c.degree_bits, c.num_routed_wires
    num_lu_slots, num_lut_slots = nr // 2, nr // 3
    max_lu = c.quotient_degree_factor - 1
    npl = -(-num_lu_slots // max_lu)
    max_lut = -(-num_lut_slots // npl)
    W = wires.astype(object)
    out = []
    for ch in range(c.num_challenges):
        a, b, alpha, delta = (int(v) for v in deltas[4 * ch:4 * ch + 4])
        polys = [[0] * (n + 1) for _ in range(npl + 1)]
        for last_lu, last_lut, first_lut in c.lookup_rows:
            for row in range(first_lut, last_lut - 1, -1):
                looked = [(W[3 * s, row] + a * W[3 * s + 1, row]) % P for s in range(num_lut_slots)]
                inv = [pow((alpha - v) % P, P - 2, P) for v in looked]
                lookup = [(W[3 * s, row] + b * W[3 * s + 1, row]) % P for s in range(num_lut_slots)]
                re = polys[0][row + 1]
                for e in lookup:
                    re = (re * delta + e) % P
                polys[0][row] = re
                for slot in range(npl):
                    prev = polys[slot][row] if slot else polys[npl][row + 1]
                    for s in range(slot * max_lut, min((slot + 1) * max_lut, num_lut_slots)):
                        prev = (prev + W[3 * s + 2, row] * inv[s]) % P
                    polys[slot + 1][row] = prev
            for row in range(last_lut - 1, last_lu - 1, -1):
                looking = [(W[2 * s, row] + a * W[2 * s + 1, row]) % P for s in range(num_lu_slots)]
                inv = [pow((alpha - v) % P, P - 2, P) for v in looking]
                for slot in range(npl):
                    prev = polys[npl][row + 1] if slot == 0 else polys[slot][row]
                    tot = sum(inv[slot * max_lu:min((slot + 1) * max_lu, num_lu_slots)]) % P
                    polys[slot + 1][row] = (prev - tot) % P
        out += [p[:n] for p in polys]
    return np.array(out, dtype=object)


@pytest.mark.parametrize("degree_bits,poseidon", [(5, False), (7, True)])
def test_oracle_lookup_polys_match_bigint_restatement(degree_bits, poseidon):
    """RE and the partial Sum / LDC polynomials of the C oracle == a big-integer restatement; the argument closes
    (the last partial polynomial is 0 at last_lu_row: Sum(end) == LDC(end)), RE at last_lut_row is the table's
    polynomial at delta, and a wrong multiplicity breaks the closing."""
    sc = SynthCircuit(degree_bits, seed=61, poseidon=poseidon, lookups=True)
    c, oc = sc.common, sc.oracle_circuit
    assert c.num_lookup_selectors == 6 and c.num_lookup_polys == 1 + -(-(c.num_routed_wires // 2) // 7)
    deltas = oracle.rand_felts((4 * c.num_challenges,), 62)
    got = oc.lookup_polys(sc.wires, deltas)
    want = _lookup_polys_bigint(c, sc.wires, deltas)
    assert got.shape == want.shape and (got.astype(object) == want).all()
    nlp = c.num_lookup_polys
    for ch in range(c.num_challenges):
        d4 = np.ascontiguousarray(deltas[4 * ch:4 * ch + 4])
        for t, (last_lu, last_lut, first_lut) in enumerate(c.lookup_rows):
            assert got[ch * nlp + nlp - 1][last_lu] == 0
            assert int(got[ch * nlp][last_lut]) == int(oracle.lib().orc_lut_poly_eval(oracle.C.byref(oc.c), t, oracle._ptr(d4)))
            assert got[ch * nlp][first_lut + 1] == 0 and got[ch * nlp + 1][first_lut + 1] == 0
    bad = sc.wires.copy()
    bad[2, c.lookup_rows[0][2]] += 1     # multiplicity of the first entry of table 0
    assert oc.lookup_polys(bad, deltas)[nlp - 1][c.lookup_rows[0][0]] != 0


def test_native_program_orders_lookup_gates_like_the_builder():
    """The lookup gates have no constraints but their ids (with the Keccak of the table) take part in the sort, and
    the gates' constants start after the 4 + n_luts lookup selectors (gate.rs:179): the host library's compiler
    reproduces the Python mirror's selectors and the oracle's gate constraint values."""
    from qp_plonky2_b200 import plonk

    sc = SynthCircuit(6, seed=63, poseidon=True, lookups=True)
    c = sc.common
    nat = plonk.native_constraint_program(c.gates, c.quotient_degree_factor + 1, c.num_routed_wires, c.luts, c.lookup_rows)
    assert nat["selector_indices"] == c.selector_indices and nat["groups"] == c.groups
    assert nat["order"] == list(range(len(c.gates)))
    shuffled = list(reversed(c.gates))
    nat2 = plonk.native_constraint_program(shuffled, c.quotient_degree_factor + 1, c.num_routed_wires, c.luts, c.lookup_rows)
    assert [shuffled[i].id() for i in nat2["order"]] == [g.id() for g in c.gates]


@pytest.mark.parametrize("degree_bits,poseidon,rec", [(6, False, False), (7, True, True)])
def test_restated_verifier_accepts_oracle_proofs_with_lookups(degree_bits, poseidon, rec):
    """prove() with two lookup tables: the restated verifier -- whose lookup constraints are written from the
    reference's extension-field check_lookup_constraints, independently of the oracle's -- accepts the oracle's
    proof; it rejects a flipped lookup opening, and it rejects a proof whose witness looks up a pair that is not in
    the table."""
    from oracle import prover as oprover
    import verifier

    sc = SynthCircuit(degree_bits, seed=93, poseidon=poseidon, extra_gates=rec, recursion_gates=rec, lookups=True)
    c = sc.common
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    proof, info = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=5)
    fri = _Fri(c.rate_bits, c.cap_height, 6, 4, 5, 5)
    digest = oprover.circuit_digest(cs.cap, degree_bits)
    assert verifier.verify(proof, c, fri, cs.cap, digest) is None
    nc, nlp = c.num_challenges, c.num_lookup_polys
    assert info["openings"]["lookup_zs"].shape == (nc * nlp, 2) and len(info["deltas"]) == 4 * nc
    assert info["deltas"][:2 * nc] == info["betas"] + info["gammas"]
    cap_bytes = 3 * (4 << c.cap_height) * 8
    at_lookup = cap_bytes + 16 * (c.num_constants + c.num_routed_wires + c.num_wires + 2 * nc) + 16 * 3
    for at in (at_lookup, at_lookup + 16 * nc * nlp):       # a lookup_zs opening, a lookup_zs_next opening
        bad = bytearray(proof)
        bad[at] ^= 1
        assert verifier.verify(bytes(bad), c, fri, cs.cap, digest) is not None
    # a looking pair that is not in the table: the oracle's prover still produces a proof (like the reference's
    # would from such a witness), and the verifier refuses it
    wires = sc.wires.copy()
    wires[1, c.lookup_rows[0][0]] = (int(wires[1, c.lookup_rows[0][0]]) + 1) % P
    good_wires, sc.wires = sc.wires, wires
    try:
        forged, _ = _oracle_prove(sc, proof_of_work_bits=6, num_query_rounds=5)
    finally:
        sc.wires = good_wires
    assert verifier.verify(forged, c, fri, cs.cap, digest) == "vanishing(zeta) != Z_H(zeta) * quotient(zeta)"


def test_factorial_example_circuit_proves_and_verifies():
    """BASELINE.json configs[0], `cargo run --example factorial` (plonky2/examples/factorial.rs): a REAL circuit --
    99 chained multiplications, constants, the in-circuit public-input hash tied to the PublicInputGate -- under
    standard_recursion_config.  The oracle's proof of "1 * 2 * ... * 100 = 100! mod p" passes the restated verifier;
    a wrong claimed result, a broken multiplication and a broken copy constraint do not."""
    import math

    from factorial_circuit import factorial_circuit
    from oracle import prover as oprover
    import verifier

    sc = factorial_circuit()
    c = sc.common
    assert sc.public_inputs == [1, math.factorial(100) % P]
    cs = oracle.PolynomialBatch.from_values(sc.constants_sigmas(), c.rate_bits, c.cap_height)
    fri = _Fri(c.rate_bits, c.cap_height, 16, 4, 5, 28)
    digest = oprover.circuit_digest(cs.cap, c.degree_bits)
    proof, _ = _oracle_prove(sc)
    assert verifier.verify(proof, c, fri, cs.cap, digest) is None
    tail = np.frombuffer(proof[-24:], dtype="<u8")
    assert [int(x) for x in tail] == [2, 1, math.factorial(100) % P]
    good_pi, good_wires = list(sc.public_inputs), sc.wires
    try:
        sc.public_inputs = [1, (good_pi[1] + 1) % P]                 # claim another result: the PI hash no longer matches
        bad, _ = _oracle_prove(sc)
        assert verifier.verify(bad, c, fri, cs.cap, digest) is not None
        sc.public_inputs = good_pi
        w = good_wires.copy()
        arith_row = int(np.nonzero(sc.row_gate == [g.id().startswith("ArithmeticGate") for g in c.gates].index(True))[0][0])
        w[3, arith_row] = (int(w[3, arith_row]) + 1) % P             # first product wrong: gate constraint AND its copies fail
        sc.wires = w
        bad, _ = _oracle_prove(sc)
        assert verifier.verify(bad, c, fri, cs.cap, digest) is not None
        w = good_wires.copy()
        w[4, arith_row] = (int(w[4, arith_row]) + 1) % P             # second multiplicand != first product: only the copy
        w[7, arith_row] = int(w[4, arith_row]) * int(w[5, arith_row]) % P   # constraint breaks (the gate itself holds)
        sc.wires = w
        bad, _ = _oracle_prove(sc)
        assert verifier.verify(bad, c, fri, cs.cap, digest) is not None
    finally:
        sc.public_inputs, sc.wires = good_pi, good_wires
