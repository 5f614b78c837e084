"""Test infrastructure: a synthetic Plonk circuit with a satisfying witness.

The reference builds circuits with CircuitBuilder (out of scope); the
permutation argument and the quotient only consume its OUTPUT -- gate list, selector and constant
polynomials, sigma polynomials, and a witness.  This module fabricates those directly:

  * rows are assigned one of {NoopGate, ConstantGate(2), PublicInputGate, ArithmeticGate(routed/4)}
    and, with `poseidon=True`, PoseidonGate (the gate recursion circuits spend their rows on;
    degree 7, 123 constraints, its own selector group) -- gate semantics: plonky2/src/gates/*.rs --
    with random gate constants;
  * with `lookups=True`, two lookup tables with their LookupGate / LookupTableGate rows laid out the way
    CircuitBuilder::add_all_lookups does (gadgets/lookup.rs:80-160: per table the LookupGate rows, then the
    LookupTableGate rows holding the table upside down, then a NoopGate row), the wires set_lookup_wires fills
    (prover.rs:39-141: multiplicities, padding with the table's first entry) and the lookup selectors
    (gates/selectors.rs:27-75);
  * the witness satisfies every gate; inputs of arithmetic operations are, with probability 1/2,
    COPIES of unconstrained cells elsewhere in the trace, and every copy class is wired as one cycle
    of the permutation sigma (what CircuitBuilder::sigma_vecs derives from `connect` calls,
    circuit_builder.rs:1199-1205), so Z and the partial products are non-trivial.
"""
import numpy as np

from qp_plonky2_b200 import plonk

import gate_witness

P = plonk.P


def rand_felts(shape, seed):
    """Uniform canonical Goldilocks elements from a seeded PCG64 (the same stream as the oracle's
    helper of that name).  Building a circuit needs nothing from oracle/: the benchmarks construct their
    inputs with this module, and only the tests' checkers (`oracle_circuit`, `eval_poly`) import it."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a).astype(np.uint64)


_M32 = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def _mulmod(a, b):
    """a * b mod p on uint64 arrays without Python integers: 32-bit halves, then reduce128 (goldilocks_field.rs:
    390-403) -- the witness of a 2^20-row circuit in seconds."""
    a, b = np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64)
    a0, a1, b0, b1 = a & _M32, a >> _S32, b & _M32, b >> _S32
    ll, lh, hl, hh = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = (ll >> _S32) + (lh & _M32) + (hl & _M32)
    lo = (ll & _M32) | ((mid & _M32) << _S32)
    hi = hh + (lh >> _S32) + (hl >> _S32) + (mid >> _S32)
    hi_hi, hi_lo = hi >> _S32, hi & _M32
    t = lo - hi_hi
    t = np.where(lo < hi_hi, t - _M32, t)      # a borrow is -2^64 = -(2^32 - 1)
    u = hi_lo * _M32
    r = t + u
    r = np.where(r < u, r + _M32, r)           # a carry is +2^64 = 2^32 - 1
    return np.where(r >= np.uint64(P), r - np.uint64(P), r)


def _addmod(a, b):
    a, b = np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64)   # canonical inputs
    r = a + b
    r = np.where(r < a, r + _M32, r)
    return np.where(r >= np.uint64(P), r - np.uint64(P), r)


class SynthCircuit:
    def __init__(self, degree_bits, seed=1, num_wires=143, num_routed_wires=80, num_challenges=2,
                 quotient_degree_factor=8, rate_bits=3, cap_height=4, poseidon=False, extra_gates=False,
                 recursion_gates=False, lookups=False):
        rng = np.random.Generator(np.random.PCG64(seed))
        n = 1 << degree_bits
        nr = num_routed_wires
        self.n = n
        gates = [plonk.NoopGate(), plonk.ConstantGate(2), plonk.PublicInputGate(),
                 plonk.ArithmeticGate.new_from_config(nr)]
        luts, lookup_rows, lookup_layout = [], [], []
        if lookups:
            # two tables; 2.4 and exactly 1 LookupGate rows of lookups (the last row of the first is padded)
            lu_slots, lut_slots = nr // 2, nr // 3
            luts = [[(i, (3 * i * i + 5) & 0xffff) for i in range(lut_slots + 11)],
                    [(2 * i + 1, i ^ 0x5a5a) for i in range(2 * lut_slots + 8)]]
            row = 1   # row 0 is the PublicInputGate
            for t, (lut, n_lookups) in enumerate(zip(luts, (2 * lu_slots + lu_slots // 2 - 1, lu_slots))):
                last_lu = row
                last_lut = last_lu + -(-n_lookups // lu_slots)
                first_lut = last_lut + -(-len(lut) // lut_slots) - 1
                lookup_rows.append((last_lu, last_lut, first_lut))
                lu_gate, lut_gate = plonk.LookupGate(nr, lut, t), plonk.LookupTableGate(nr, lut, t, last_lut)
                gates += [lu_gate, lut_gate]
                lookup_layout.append((lu_gate, lut_gate, n_lookups))
                row = first_lut + 2   # the NoopGate row after the table
            assert row <= n, "circuit too small for the lookup rows"
        if poseidon:
            gates.append(plonk.PoseidonGate())
        if extra_gates:   # the extension-field arithmetic and bit-decomposition gates of recursion circuits
            gates += [plonk.ArithmeticExtensionGate.new_from_config(nr), plonk.MulExtensionGate.new_from_config(nr),
                      plonk.BaseSumGate2(63)]
        if recursion_gates:   # the rest of a recursive verifier's gate set (standard_recursion_config shapes)
            gates += [plonk.RandomAccessGate.new_from_config(num_wires, nr, 4),
                      plonk.ReducingGate(plonk.ReducingGate.max_coeffs_len(num_wires, nr)),
                      plonk.ReducingExtensionGate(plonk.ReducingExtensionGate.max_coeffs_len(num_wires, nr)),
                      plonk.PoseidonMdsGate(), plonk.ExponentiationGate.new_from_config(num_wires, nr),
                      plonk.CosetInterpolationGate.with_max_degree(4, quotient_degree_factor)]
        self.common = c = plonk.CommonCircuitData(degree_bits, gates, num_wires, nr, num_challenges,
                                                  quotient_degree_factor, rate_bits, cap_height, luts=luts,
                                                  lookup_rows=lookup_rows)
        self._oracle_circuit = None
        idx = {g.id().split(" ")[0].split("(")[0]: i for i, g in enumerate(c.gates)
               if not isinstance(g, (plonk.LookupGate, plonk.LookupTableGate))}
        gc0 = c.num_selectors + c.num_lookup_selectors   # first gate constant (gate.rs:179)
        # gate per row: mostly arithmetic (and Poseidon), a few of the others; row 0 is the public-input gate
        kinds_p = {"NoopGate": 0.2, "ConstantGate": 0.1, "ArithmeticGate": 0.7}
        if poseidon:
            kinds_p.update({"ArithmeticGate": 0.3, "PoseidonGate": 0.4})
        if extra_gates:
            kinds_p["ArithmeticGate"] -= 0.2
            kinds_p.update({"ArithmeticExtensionGate": 0.08, "MulExtensionGate": 0.06, "BaseSumGate": 0.06})
        rec_names = ["RandomAccessGate", "ReducingGate", "ReducingExtensionGate", "PoseidonMdsGate",
                     "ExponentiationGate", "CosetInterpolationGate"]
        if recursion_gates:
            kinds_p[max(kinds_p, key=kinds_p.get)] -= 0.18   # from the most frequent gate (Poseidon / Arithmetic)
            kinds_p.update({k: 0.03 for k in rec_names})
        row_gate = rng.choice([idx[k] for k in kinds_p], size=n, p=list(kinds_p.values()))
        row_gate[0] = idx["PublicInputGate"]
        for (lu_gate, lut_gate, _), (last_lu, last_lut, first_lut) in zip(lookup_layout, lookup_rows):
            row_gate[last_lu:last_lut] = c.gates.index(lu_gate)
            row_gate[last_lut:first_lut + 1] = c.gates.index(lut_gate)
            row_gate[first_lut + 1] = idx["NoopGate"]
        self.row_gate = row_gate
        # constants: selector polynomials (selectors.rs:141-159), then the gate constants
        consts = np.zeros((c.num_constants, n), dtype=np.uint64)
        for s, (a, b) in enumerate(c.groups):
            in_group = (row_gate >= a) & (row_gate < b)
            consts[s] = np.where(in_group, row_gate, plonk.UNUSED_SELECTOR if c.num_selectors > 1 else row_gate)
        gate_consts = rand_felts((c.num_gate_constants, n), seed + 1)
        uses = (row_gate == idx["ConstantGate"]) | (row_gate == idx["ArithmeticGate"])
        if extra_gates:
            uses |= (row_gate == idx["ArithmeticExtensionGate"]) | (row_gate == idx["MulExtensionGate"])
        if recursion_gates:
            uses |= row_gate == idx["RandomAccessGate"]
        consts[gc0:] = np.where(uses[None, :], gate_consts, 0)
        # lookup selectors: TransSre, TransLdc, InitSre, LastLdc, then one "ends" selector per table
        for t, (last_lu, last_lut, first_lut) in enumerate(lookup_rows):
            ls = consts[c.num_selectors:]
            ls[0, last_lut:first_lut + 1] = 1
            ls[1, last_lu:last_lut] = 1
            ls[2, first_lut + 1] = 1
            ls[3, last_lu] = 1
            ls[4 + t, last_lut] = 1
        self.constants = consts
        # witness
        wires = rand_felts((num_wires, n), seed + 2)
        for (_, _, n_lookups), (last_lu, last_lut, first_lut), lut in zip(lookup_layout, lookup_rows, luts):
            lu_slots, lut_slots = nr // 2, nr // 3
            wires[:, first_lut + 1] = 0   # "Will ensure the next row's wires will be all zeros", gadgets/lookup.rs:150
            looked_up = [int(k) for k in rng.integers(0, len(lut), size=n_lookups)]
            looked_up += [0] * (-n_lookups % lu_slots)            # padding of the last LookupGate row: the first entry
            mult = np.bincount(looked_up, minlength=len(lut))
            for k, e in enumerate(looked_up):
                r, s_ = last_lu + k // lu_slots, k % lu_slots
                wires[2 * s_, r], wires[2 * s_ + 1, r] = lut[e]
            for k in range(-(-len(lut) // lut_slots) * lut_slots):
                r, s_ = first_lut - k // lut_slots, k % lut_slots
                e = k if k < len(lut) else 0                     # unused slots: the first entry, multiplicity 0
                wires[3 * s_, r], wires[3 * s_ + 1, r] = lut[e]
                wires[3 * s_ + 2, r] = mult[k] if k < len(lut) else 0
        # public inputs and their hash (prover.rs:185-186); the PublicInputGate row carries the hash
        self.public_inputs = [int(x) for x in rand_felts((3,), seed + 3)]
        from qp_plonky2_b200 import prover as _prover   # host-side hash_no_pad (qp_hash_no_pad), checked against the oracle's in tests
        self.public_inputs_hash = _prover.hash_no_pad(self.public_inputs)
        pi_rows = row_gate == idx["PublicInputGate"]
        for k in range(4):
            wires[k, pi_rows] = self.public_inputs_hash[k]
        const_rows = row_gate == idx["ConstantGate"]
        for k in range(2):
            wires[k, const_rows] = consts[gc0 + k, const_rows]
        # copy constraints: free cells = routed wires of Noop rows (unconstrained)
        noop_rows = np.nonzero(row_gate == idx["NoopGate"])[0]
        noop_rows = noop_rows[~np.isin(noop_rows, [r[2] + 1 for r in lookup_rows])]   # the tables' zero rows stay zero
        arith_rows = np.nonzero(row_gate == idx["ArithmeticGate"])[0]
        sigma_row = np.tile(np.arange(n, dtype=np.int64), (nr, 1))
        sigma_col = np.tile(np.arange(nr, dtype=np.int64)[:, None], (1, n))
        num_ops = nr // 4
        if len(noop_rows) and len(arith_rows):
            in_cols = np.array([4 * o + k for o in range(num_ops) for k in range(3)])
            tc, tr = np.meshgrid(in_cols, arith_rows, indexing="ij")
            tc, tr = tc.ravel(), tr.ravel()
            pick = rng.random(tc.size) < 0.5
            tc, tr = tc[pick], tr[pick]
            sr = rng.choice(noop_rows, size=tc.size)
            sc = rng.integers(0, nr, size=tc.size)
            wires[tc, tr] = wires[sc, sr]
            # classes keyed by source cell; one cycle per class: source -> t1 -> t2 -> ... -> source
            key = sc.astype(np.int64) * n + sr
            order = np.argsort(key, kind="stable")
            key, tc, tr, sc, sr = key[order], tc[order], tr[order], sc[order], sr[order]
            first = np.r_[True, key[1:] != key[:-1]]
            last = np.r_[key[1:] != key[:-1], True]
            # source -> first target of its class
            sigma_col[sc[first], sr[first]] = tc[first]
            sigma_row[sc[first], sr[first]] = tr[first]
            # target -> next target, last target -> source
            nxt_c = np.r_[tc[1:], 0]
            nxt_r = np.r_[tr[1:], 0]
            nxt_c[last], nxt_r[last] = sc[last], sr[last]
            sigma_col[tc, tr] = nxt_c
            sigma_row[tc, tr] = nxt_r
        # arithmetic outputs (arithmetic_base.rs:181)
        if len(arith_rows):
            c0 = consts[gc0][arith_rows]
            c1 = consts[gc0 + 1][arith_rows]
            for o in range(num_ops):
                m0, m1, ad = wires[4 * o][arith_rows], wires[4 * o + 1][arith_rows], wires[4 * o + 2][arith_rows]
                wires[4 * o + 3][arith_rows] = _addmod(_mulmod(_mulmod(m0, m1), c0), _mulmod(ad, c1))
        if extra_gates:
            c0 = consts[gc0].astype(object)
            c1 = consts[gc0 + 1].astype(object)
            W = wires.astype(object)
            rows = np.nonzero(row_gate == idx["ArithmeticExtensionGate"])[0]
            for o in range(nr // 8):   # arithmetic_extension.rs:92-110
                b = 8 * o
                p0 = (W[b, rows] * W[b + 2, rows] + 7 * W[b + 1, rows] * W[b + 3, rows]) % P
                p1 = (W[b, rows] * W[b + 3, rows] + W[b + 1, rows] * W[b + 2, rows]) % P
                wires[b + 6, rows] = ((p0 * c0[rows] + W[b + 4, rows] * c1[rows]) % P).astype(np.uint64)
                wires[b + 7, rows] = ((p1 * c0[rows] + W[b + 5, rows] * c1[rows]) % P).astype(np.uint64)
            rows = np.nonzero(row_gate == idx["MulExtensionGate"])[0]
            for o in range(nr // 6):   # multiplication_extension.rs:86-101
                b = 6 * o
                p0 = (W[b, rows] * W[b + 2, rows] + 7 * W[b + 1, rows] * W[b + 3, rows]) % P
                p1 = (W[b, rows] * W[b + 3, rows] + W[b + 1, rows] * W[b + 2, rows]) % P
                wires[b + 4, rows] = ((p0 * c0[rows]) % P).astype(np.uint64)
                wires[b + 5, rows] = ((p1 * c0[rows]) % P).astype(np.uint64)
            rows = np.nonzero(row_gate == idx["BaseSumGate"])[0]
            bits = rng.integers(0, 2, size=(63, len(rows)), dtype=np.uint64)   # base_sum.rs: limbs in {0, 1}
            wires[1:64, rows] = bits
            total = np.zeros(len(rows), dtype=object)
            for i in range(62, -1, -1):
                total = (total * 2 + bits[i].astype(object)) % P
            wires[0, rows] = total.astype(np.uint64)
        if recursion_gates:   # one generator call per row, like the reference's SimpleGenerators
            G = {g.id().split(" ")[0].split("(")[0]: g for g in c.gates}
            felt = lambda: int(rng.integers(0, P, dtype=np.uint64))
            ext = lambda: (felt(), felt())
            for r in range(n):
                name = next((k for k in rec_names if row_gate[r] == idx[k]), None)
                if name is None:
                    continue
                g = G[name]
                if name == "RandomAccessGate":
                    row = gate_witness.generate(g, [(int(rng.integers(0, g.vec_size)), [felt() for _ in range(g.vec_size)])
                                      for _ in range(g.num_copies)],
                                     [int(consts[gc0 + i, r]) for i in range(g.num_extra_constants)])
                elif name == "ReducingGate":
                    row = gate_witness.generate(g, ext(), ext(), [felt() for _ in range(g.num_coeffs)])
                elif name == "ReducingExtensionGate":
                    row = gate_witness.generate(g, ext(), ext(), [ext() for _ in range(g.num_coeffs)])
                elif name == "PoseidonMdsGate":
                    row = gate_witness.generate(g, [ext() for _ in range(12)])
                elif name == "ExponentiationGate":
                    row = gate_witness.generate(g, felt(), [int(b) for b in rng.integers(0, 2, size=g.num_power_bits)])
                else:
                    row = gate_witness.generate(g, felt() or 1, [ext() for _ in range(g.num_points)], ext())
                for w, v in row.items():
                    wires[w, r] = v
        # Poseidon rows: inputs and swap are free, everything else follows (PoseidonGenerator)
        if poseidon:
            pg = plonk.PoseidonGate()
            for r in np.nonzero(row_gate == idx["PoseidonGate"])[0]:
                row = gate_witness.generate(pg, [int(wires[i, r]) for i in range(12)], int(rng.integers(0, 2)))
                for w, v in row.items():
                    wires[w, r] = v
        self.wires = wires
        # sigma polynomials' values: k_is[col] * w^row  (circuit_builder.rs sigma_vecs)
        w = plonk.primitive_root_of_unity(degree_bits)
        sub = np.ones(n, dtype=np.uint64)           # w^i by doubling: sub[m .. 2m) = sub[0 .. m) * w^m
        m, wm = 1, w
        while m < n:
            sub[m:2 * m] = _mulmod(sub[:m], np.uint64(wm))
            m, wm = 2 * m, wm * wm % P
        k_arr = np.array([int(k) for k in c.k_is], dtype=np.uint64)
        self.sigmas = _mulmod(k_arr[sigma_col], sub[sigma_row])
        self.subgroup = sub

    @property
    def oracle_circuit(self):
        """The same circuit as the ORACLE describes it (checker side, tests only)."""
        if self._oracle_circuit is None:
            import oracle
            c, nr = self.common, self.common.num_routed_wires
            degree_bits, num_challenges, num_wires = c.degree_bits, c.num_challenges, c.num_wires
            quotient_degree_factor = c.quotient_degree_factor
            kinds = {"NoopGate": oracle.GATE_NOOP, "ConstantGate": oracle.GATE_CONSTANT,
                     "PublicInputGate": oracle.GATE_PUBLIC_INPUT, "ArithmeticGate": oracle.GATE_ARITHMETIC,
                     "PoseidonGate": oracle.GATE_POSEIDON, "ArithmeticExtensionGate": oracle.GATE_ARITHMETIC_EXT,
                     "MulExtensionGate": oracle.GATE_MUL_EXT, "BaseSumGate": oracle.GATE_BASE_SUM_2,
                     "RandomAccessGate": oracle.GATE_RANDOM_ACCESS, "ReducingGate": oracle.GATE_REDUCING,
                     "ReducingExtensionGate": oracle.GATE_REDUCING_EXT, "PoseidonMdsGate": oracle.GATE_POSEIDON_MDS,
                     "ExponentiationGate": oracle.GATE_EXPONENTIATION,
                     "CosetInterpolationGate": oracle.GATE_COSET_INTERPOLATION}
            kinds.update({"LookupGate": oracle.GATE_LOOKUP, "LookupTableGate": oracle.GATE_LOOKUP_TABLE})
            og = []
            for i, g in enumerate(c.gates):
                name = g.id().split(" ")[0].split("(")[0]
                param = g.param
                og.append((kinds[name], param, c.selector_indices[i], c.groups[c.selector_indices[i]]))
            self._oracle_circuit = oracle.Circuit(degree_bits, c.quotient_degree_bits, num_challenges, nr, num_wires,
                                                 c.num_constants, c.num_partial_products, quotient_degree_factor,
                                                 c.num_selectors, og, c.k_is, luts=c.luts, lookup_rows=c.lookup_rows)
        return self._oracle_circuit

    def constants_sigmas(self):
        return np.ascontiguousarray(np.concatenate([self.constants, self.sigmas], axis=0))


def eval_poly(coeffs, x):
    """Horner over F_p (PolynomialCoeffs::eval)."""
    import oracle
    out = oracle.lib().orc_eval_poly_ext  # F_p^2 evaluator with point (x, 0)
    res = np.zeros(2, dtype=np.uint64)
    pt = np.array([x, 0], dtype=np.uint64)
    c = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out(oracle._ptr(c), c.size, oracle._ptr(pt), oracle._ptr(res))
    return int(res[0])


def verifier_identity_holds(sc, cs_polys, wires_polys, zs_polys, quotient_coeffs, betas, gammas, alphas, x0):
    """The verifier's check (verifier/src/plonk/verifier.rs:86-100) at a base-field point x0:
    vanishing(x0) == Z_H(x0) * quotient(x0), with every opening computed from the coefficient
    vectors.  Size-independent property: it holds iff the committed quotient is the right one."""
    c = sc.common
    n = 1 << c.degree_bits
    g = plonk.primitive_root_of_unity(c.degree_bits)
    x0 = int(x0)
    cs = [eval_poly(p, x0) for p in cs_polys]
    wv = [eval_poly(p, x0) for p in wires_polys]
    zv = [eval_poly(p, x0) for p in zs_polys]
    zn = [eval_poly(p, x0 * g % P) for p in zs_polys[: c.num_challenges]]
    nc = c.num_challenges
    van = sc.oracle_circuit.eval_vanishing_poly_base(
        x0, cs[: c.num_constants], wv, zv[:nc], zn, zv[nc:], cs[c.num_constants:], betas, gammas, alphas,
        sc.public_inputs_hash)
    z_h = (pow(x0, n, P) - 1) % P
    ok = True
    for a in range(nc):
        q = eval_poly(quotient_coeffs[a], x0)
        ok &= int(van[a]) == z_h * q % P
    return ok


class Ext:
    """F_p^2 = F_p[X]/(X^2 - 7) element for the restated verifier."""

    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a, self.b = int(a) % P, int(b) % P

    @staticmethod
    def of(x):
        return x if isinstance(x, Ext) else Ext(x)

    def __add__(self, o):
        o = Ext.of(o)
        return Ext(self.a + o.a, self.b + o.b)

    def __sub__(self, o):
        o = Ext.of(o)
        return Ext(self.a - o.a, self.b - o.b)

    def __rsub__(self, o):
        return Ext.of(o) - self

    def __mul__(self, o):
        o = Ext.of(o)
        return Ext(self.a * o.a + 7 * self.b * o.b, self.a * o.b + self.b * o.a)

    __radd__ = __add__
    __rmul__ = __mul__

    def __eq__(self, o):
        o = Ext.of(o)
        return self.a == o.a and self.b == o.b

    def pow(self, e):
        r, x = Ext(1), self
        while e:
            if e & 1:
                r = r * x
            x = x * x
            e >>= 1
        return r

    def inv(self):
        # (a + bX)^-1 = (a - bX) / (a^2 - 7 b^2)
        d = pow((self.a * self.a - 7 * self.b * self.b) % P, P - 2, P)
        return Ext(self.a * d, -self.b * d)


def check_lookup_constraints(common, wires, local_lookup_zs, next_lookup_zs, lookup_selectors, deltas):
    """check_lookup_constraints (verifier/src/plonk/vanishing_poly.rs:212-381) for one challenge, over Ext values;
    deltas = this challenge's (ChallengeA, ChallengeB, ChallengeAlpha, ChallengeDelta).  Written from the reference's
    extension-field form (sum over i of prod over j != i), independently of the oracle's base-field one."""
    c = common
    num_lu_slots, num_lut_slots = c.num_routed_wires // 2, c.num_routed_wires // 3
    lu_degree = c.quotient_degree_factor - 1            # lookup_accumulator_degree
    num_sldc = len(local_lookup_zs) - 1
    lut_degree = -(-num_lut_slots // num_sldc)
    a, b, alpha, delta = (int(v) for v in deltas)
    TRANS_SRE, TRANS_LDC, INIT_SRE, LAST_LDC, START_END = 0, 1, 2, 3, 4
    z_re, next_z_re = local_lookup_zs[0], next_lookup_zs[0]
    z_x, z_gx = local_lookup_zs[1:], next_lookup_zs[1:]
    looked = [wires[3 * s] + wires[3 * s + 1] * a for s in range(num_lut_slots)]
    looking = [wires[2 * s] + wires[2 * s + 1] * a for s in range(num_lu_slots)]
    lookup = [wires[3 * s] + wires[3 * s + 1] * b for s in range(num_lut_slots)]
    out = [lookup_selectors[LAST_LDC] * z_x[num_sldc - 1], lookup_selectors[INIT_SRE] * z_x[0],
           lookup_selectors[INIT_SRE] * z_re]
    for r in range(START_END, c.num_lookup_selectors):
        lut = c.luts[r - START_END]
        # get_lut_poly(..).eval(delta): coefficients = the combos in table order, padded, REVERSED
        coeffs = [(i + b * o) % P for i, o in lut]
        coeffs += [coeffs[0]] * ((num_lut_slots - len(lut) % num_lut_slots) % num_lut_slots)
        coeffs.reverse()
        ev = sum(cf * pow(delta, k, P) for k, cf in enumerate(coeffs)) % P
        out.append(lookup_selectors[r] * (z_re - ev))
    cur = next_z_re
    for e in lookup:
        cur = cur * delta + e
    out.append(lookup_selectors[TRANS_SRE] * (z_re - cur))

    def prod(vals):
        r = Ext(1)
        for v in vals:
            r = r * v
        return r

    for poly in range(num_sldc):
        t_rng = range(poly * lut_degree, min((poly + 1) * lut_degree, num_lut_slots))
        u_rng = range(poly * lu_degree, min((poly + 1) * lu_degree, num_lu_slots))
        lut_prod = prod(Ext(alpha) - looked[i] for i in t_rng)
        lu_prod = prod(Ext(alpha) - looking[i] for i in u_rng)
        lu_sum_prods = Ext(0)
        for i in u_rng:
            lu_sum_prods = lu_sum_prods + prod(Ext(alpha) - looking[j] for j in u_rng if j != i)
        lut_sum_prods_with_mul = Ext(0)
        for i in t_rng:
            lut_sum_prods_with_mul = lut_sum_prods_with_mul + wires[3 * i + 2] * prod(
                Ext(alpha) - looked[j] for j in t_rng if j != i)
        prev = z_gx[num_sldc - 1] if poly == 0 else z_x[poly - 1]
        out.append(lookup_selectors[TRANS_SRE] * (lut_prod * (z_x[poly] - prev) - lut_sum_prods_with_mul))
        out.append(lookup_selectors[TRANS_LDC] * (lu_prod * (z_x[poly] - prev) + lu_sum_prods))
    return out


def verifier_plonk_identity(common, openings, zeta, betas, gammas, alphas, pih, deltas=None):
    """verify_with_challenges' algebraic check (verifier/src/plonk/verifier.rs:60-100): evaluate the
    vanishing polynomial at zeta from the OPENINGS (eval_vanishing_poly, vanishing_poly.rs:38-150)
    and compare with Z_H(zeta) * sum_i zeta^(n i) quotient_chunk_i(zeta), per challenge.  deltas: the flat
    lookup challenges (4 per challenge) of a circuit with lookup tables."""
    c = common
    n = 1 << c.degree_bits
    z = Ext(*zeta)
    E = lambda arr: [Ext(int(v[0]), int(v[1])) for v in arr]
    consts, sig, wires = E(openings["constants"]), E(openings["plonk_sigmas"]), E(openings["wires"])
    zs, zs_next, pps = E(openings["plonk_zs"]), E(openings["plonk_zs_next"]), E(openings["partial_products"])
    qs = E(openings["quotient_polys"])
    nc, nr, npp, md = c.num_challenges, c.num_routed_wires, c.num_partial_products, c.quotient_degree_factor
    z_h = z.pow(n) - 1
    l_0 = z_h * ((z - 1) * n).inv()   # eval_l_0, plonk_common.rs
    terms = []
    for i in range(nc):
        terms.append(l_0 * (zs[i] - 1))
    for i in range(nc):
        num = [wires[j] + z * (int(c.k_is[j]) * int(betas[i]) % P) + int(gammas[i]) for j in range(nr)]
        den = [wires[j] + sig[j] * int(betas[i]) + int(gammas[i]) for j in range(nr)]
        accs = [zs[i]] + pps[i * npp:(i + 1) * npp] + [zs_next[i]]
        for w in range(npp + 1):
            pn = pd = Ext(1)
            for j in range(w * md, min((w + 1) * md, nr)):
                pn, pd = pn * num[j], pd * den[j]
            terms.append(accs[w] * pn - accs[w + 1] * pd)
    nlp = getattr(c, "num_lookup_polys", 0)
    if nlp:
        lz, lzn = E(openings["lookup_zs"]), E(openings["lookup_zs_next"])
        sel = consts[c.num_selectors:c.num_selectors + c.num_lookup_selectors]
        for i in range(nc):
            terms += check_lookup_constraints(c, wires, lz[i * nlp:(i + 1) * nlp], lzn[i * nlp:(i + 1) * nlp], sel,
                                              deltas[4 * i:4 * i + 4])
    gate_terms = [Ext(0)] * c.num_gate_constraints
    prefix = c.num_selectors + getattr(c, "num_lookup_selectors", 0)
    for gi, g in enumerate(c.gates):
        sel = c.selector_indices[gi]
        a, b = c.groups[sel]
        s = consts[sel]
        f = Ext(1)
        for j in range(a, b):
            if j != gi:
                f = f * (Ext(j) - s)
        if c.num_selectors > 1:
            f = f * (Ext(plonk.UNUSED_SELECTOR) - s)
        cons = g.eval_unfiltered(lambda k: consts[prefix + k], lambda k: wires[k], lambda k: Ext(int(pih[k])))
        for k, cv in enumerate(cons):
            gate_terms[k] = gate_terms[k] + f * cv
    terms += gate_terms
    zeta_pow_deg = z.pow(n)
    for a in range(nc):
        acc = Ext(0)
        for t in reversed(terms):
            acc = acc * int(alphas[a]) + t
        chunk = qs[a * c.quotient_degree_factor:(a + 1) * c.quotient_degree_factor]
        q = Ext(0)
        for v in reversed(chunk):   # reduce_with_powers(chunk, zeta^n)
            q = q * zeta_pow_deg + v
        if not (acc == z_h * q):
            return False
    return True
