"""Test infrastructure: witness generation for the gates of qp_plonky2_b200.plonk -- the reference's SimpleGenerators
(plonky2/src/gates/*.rs), which the product does not need (the prover starts from the full witness, SURVEY.md
section 8f).  generate(gate, ...) -> {wire index: value} for one row."""
from qp_plonky2_b200 import plonk
from qp_plonky2_b200.plonk import P, _ext_mul, _ext_add, _ext_sub, _ext_scalar  # noqa: F401


class ModP:
    """Integer mod p with the operators the gate bodies use (witness generation, self-checks)."""

    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v) % P

    @staticmethod
    def _o(o):
        return o.v if isinstance(o, ModP) else int(o)

    def __add__(self, o):
        return ModP(self.v + self._o(o))

    def __sub__(self, o):
        return ModP(self.v - self._o(o))

    def __rsub__(self, o):
        return ModP(self._o(o) - self.v)

    def __mul__(self, o):
        return ModP(self.v * self._o(o))

    __radd__ = __add__
    __rmul__ = __mul__

    def __int__(self):
        return self.v


def _generate_RandomAccessGate(self, copy_inputs, extra_constants):
    """RandomAccessGenerator (random_access.rs:370-406): copy_inputs[copy] = (index, [items])."""
    row = {}
    for copy, (index, items) in enumerate(copy_inputs):
        row[self.wire_access_index(copy)] = index
        for i, v in enumerate(items):
            row[self.wire_list_item(i, copy)] = v
        row[self.wire_claimed_element(copy)] = items[index]
        for i in range(self.bits):
            row[self.wire_bit(i, copy)] = (index >> i) & 1
    for i, v in enumerate(extra_constants):
        row[self.wire_extra_constant(i)] = v
    return row


def _generate_ReducingGate(self, alpha, old_acc, coeffs):
    """ReducingGenerator (reducing.rs:210-242); coeffs: base elements (or pairs for the extension gate)."""
    row = {2: alpha[0], 3: alpha[1], 4: old_acc[0], 5: old_acc[1]}
    acc = (ModP(old_acc[0]), ModP(old_acc[1]))
    al = (ModP(alpha[0]), ModP(alpha[1]))
    for i, c in enumerate(coeffs):
        c = c if self.EXT_COEFFS else (c, 0)
        w = 6 + (2 * i if self.EXT_COEFFS else i)
        row[w] = c[0]
        if self.EXT_COEFFS:
            row[w + 1] = c[1]
        t = _ext_mul(acc, al)
        acc = (t[0] + c[0], t[1] + c[1])
        a = self.wires_accs(i)
        row[a], row[a + 1] = acc[0].v, acc[1].v
    return row


def _generate_PoseidonMdsGate(self, inputs):
    """PoseidonMdsGenerator: inputs = 12 pairs."""
    row = {}
    for i, (a, b) in enumerate(inputs):
        row[2 * i], row[2 * i + 1] = a, b
    comp = self._mds_ext([(ModP(a), ModP(b)) for a, b in inputs])
    for i in range(12):
        row[24 + 2 * i], row[25 + 2 * i] = comp[i][0].v, comp[i][1].v
    return row


def _generate_ExponentiationGate(self, base, power_bits):
    """ExponentiationGenerator (exponentiation.rs:270-306); power_bits little-endian."""
    n = self.num_power_bits
    row = {0: base}
    cur = 1
    for i in range(n):
        row[1 + i] = power_bits[i]
    for i in range(n):
        if power_bits[n - i - 1] == 1:
            cur = cur * base % P
        row[2 + n + i] = cur
        last = cur
        cur = cur * cur % P
    row[1 + n] = last
    return row


def _generate_CosetInterpolationGate(self, shift, values, point):
    """InterpolationGenerator (coset_interpolation.rs:461-530)."""
    row = {0: shift}
    for i, (a, b) in enumerate(values):
        row[self.wires_value(i)], row[self.wires_value(i) + 1] = a, b
    w = self.wires_evaluation_point()
    row[w], row[w + 1] = point
    sinv = pow(shift, P - 2, P)
    row[self.wire_shift_inverse()] = sinv
    shifted = (ModP(point[0] * sinv), ModP(point[1] * sinv))
    w = self.wires_shifted_evaluation_point()
    row[w], row[w + 1] = shifted[0].v, shifted[1].v
    vals = [(ModP(a), ModP(b)) for a, b in values]
    ev, prod = self._first(vals, shifted)
    for i in range(self.num_intermediates):
        for w, v in ((self.wires_intermediate_eval(i), ev), (self.wires_intermediate_prod(i), prod)):
            row[w], row[w + 1] = v[0].v, v[1].v
        start = 1 + (self.degree - 1) * (i + 1)
        end = min(start + self.degree - 1, self.num_points)
        ev, prod = self._partial(start, end, vals, shifted, ev, prod)
    w = self.wires_evaluation_value()
    row[w], row[w + 1] = ev[0].v, ev[1].v
    return row


def _generate_PoseidonGate(self, inputs, swap):
    """PoseidonGenerator (poseidon.rs:424-520): -> {wire index: value} for one row."""

    class M:  # integer mod p with the operators the body uses
        __slots__ = ("v",)

        def __init__(self, v):
            self.v = int(v) % P

        def _o(self, o):
            return o.v if isinstance(o, M) else int(o)

        def __add__(self, o):
            return M(self.v + self._o(o))

        def __sub__(self, o):
            return M(self.v - self._o(o))

        def __mul__(self, o):
            return M(self.v * self._o(o))

    row = {i: M(inputs[i]) for i in range(12)}
    row[self.WIRE_SWAP] = M(swap)

    def on_sbox_in(w, computed, negate=False):
        row[w] = computed
        return computed

    out = self._run(lambda i: row[i], on_sbox_in, lambda c: None)
    for i in range(12):
        row[12 + i] = out[i]
    return {w: v.v for w, v in row.items()}


_GENERATORS = {
    plonk.RandomAccessGate: _generate_RandomAccessGate,
    plonk.ReducingGate: _generate_ReducingGate,
    plonk.PoseidonMdsGate: _generate_PoseidonMdsGate,
    plonk.ExponentiationGate: _generate_ExponentiationGate,
    plonk.CosetInterpolationGate: _generate_CosetInterpolationGate,
    plonk.PoseidonGate: _generate_PoseidonGate,
}


def generate(gate, *args):
    """Run the generator of `gate`'s type (subclasses resolve to their nearest generator)."""
    for cls in type(gate).__mro__:
        if cls in _GENERATORS:
            return _GENERATORS[cls](gate, *args)
    raise TypeError("no witness generator for %r" % (type(gate).__name__,))
