"""Host mirror of the circuit data that the permutation argument and the quotient evaluation need.

Reference interface (citations reference-relative):
    Gate trait (id / degree / num_constants / num_constraints / eval_unfiltered)  plonky2/src/gates/gate.rs
    NoopGate, ConstantGate, PublicInputGate, ArithmeticGate                       plonky2/src/gates/{noop,constant,public_input,arithmetic_base}.rs
    selector_polynomials -> SelectorsInfo                                         plonky2/src/gates/selectors.rs:99-166
    get_unique_coset_shifts (k_is)                                                field/src/cosets.rs:9-24
    CommonCircuitData fields                                                      plonky2/src/plonk/circuit_data.rs:412-470
    wires_permutation_partial_products_and_zs / compute_quotient_polys            plonky2/src/plonk/prover.rs:402-480,640-866

The device never sees gate objects: `CommonCircuitData.constraint_program()` runs every gate's
`eval_unfiltered` over a recording value type and compiles the result, with the gate filters
(gate.rs:326-333), into the straight-line program of include/qp_plonky2_b200.h.  A Rust shim does
the same with a recording `Field` type over `Gate::eval_unfiltered_base_one`.
"""
import ctypes as C

import numpy as np

P = 0xFFFFFFFF00000001
MULTIPLICATIVE_GROUP_GENERATOR = 14293326489335486720  # field/src/goldilocks_field.rs:84
UNUSED_SELECTOR = 0xFFFFFFFF  # core/src/selectors.rs

(OP_END, OP_LDW, OP_LDK, OP_LDP, OP_LDI, OP_ADD, OP_SUB, OP_MUL, OP_EMIT, OP_GATE, OP_MULI, OP_ADDI, OP_WAIT, OP_FMAI,
 OP_NATIVE_POSEIDON) = range(15)


def encode_word(op, dst=0, a=0, b=0, c=0):
    """op | dst << 8 | a << 16 | b << 24 | c << 32 (include/qp_plonky2_b200.h)."""
    assert dst < 256 and a < 256 and b < 256
    return op | (dst << 8) | (a << 16) | (b << 24) | (c << 32)


def decode_word(ins):
    ins = int(ins)
    return ins & 0xFF, (ins >> 8) & 0xFF, (ins >> 16) & 0xFF, (ins >> 24) & 0xFF, ins >> 32


# ---- recording value type ------------------------------------------------------------------------
class Val:
    """A node of the constraint expression DAG."""

    __slots__ = ("prog", "idx")

    def __init__(self, prog, idx):
        self.prog, self.idx = prog, idx

    def _bin(self, op, other):
        if not isinstance(other, Val):
            # constants go into the instruction (pool index), not into a register
            if op == OP_MUL:
                return self.prog._node(OP_MULI, self.idx, self.prog.pool_slot(other))
            if op == OP_ADD:
                return self.prog._node(OP_ADDI, self.idx, self.prog.pool_slot(other))
            return self.prog._node(OP_ADDI, self.idx, self.prog.pool_slot(-int(other)))
        return self.prog._node(op, self.idx, other.idx)

    def __add__(self, o):
        return self._bin(OP_ADD, o)

    def __sub__(self, o):
        return self._bin(OP_SUB, o)

    def __mul__(self, o):
        return self._bin(OP_MUL, o)

    def __rsub__(self, o):
        return self.prog.imm(o) - self

    __radd__ = __add__
    __rmul__ = __mul__


class ConstraintProgram:
    """Builds the program: leaves are loads, inner nodes field operations; `emit_gate` appends a
    gate's constraints (each with its index k: the kernel adds alpha^k * value) and its filter."""

    def __init__(self):
        self.nodes = []    # (op, a, b)
        self.memo = {}
        self.pool = []
        self.pool_index = {}
        self.actions = []  # (OP_EMIT | OP_GATE, node)

    def _node(self, op, a=0, b=0):
        key = (op, a, b)
        if op in (OP_ADD, OP_MUL) and a > b:
            key = (op, b, a)
        i = self.memo.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self.memo[key] = i
        return Val(self, i)

    def wire(self, i):
        return self._node(OP_LDW, i)

    def constant(self, i):
        """Polynomial i of the constants_sigmas oracle."""
        return self._node(OP_LDK, i)

    def public_input_hash(self, i):
        return self._node(OP_LDP, i)

    def pool_slot(self, v):
        v = int(v) % P
        k = self.pool_index.get(v)
        if k is None:
            k = len(self.pool)
            self.pool.append(v)
            self.pool_index[v] = k
        return k

    def imm(self, v):
        return self._node(OP_LDI, self.pool_slot(v))

    def emit_gate(self, constraints, filt):
        for k, c in enumerate(constraints):
            self.actions.append((OP_EMIT, c.idx, k))
        self.actions.append((OP_GATE, filt.idx, 0))
        # values are not shared across gates: a wire loaded for one gate would otherwise stay in a
        # register until the last gate that reads it
        self.memo = {}

    def compile_native(self):
        """The recording handed to the native compiler (qp_program_from_dag, host/plonk_host.cpp): the
        scheduling, segmentation and register allocation the device wants.  -> (code, pool, n_regs)"""
        from . import lib
        nodes = np.array([(op, a, b) for op, a, b in self.nodes], dtype=np.uint32).reshape(-1, 3)
        acts = np.array([(op, node, k) for op, node, k in self.actions], dtype=np.uint32).reshape(-1, 3)
        pool = np.array(self.pool or [0], dtype=np.uint64)
        h = C.c_void_p()
        rc = lib().qp_program_from_dag(nodes.ctypes.data, len(nodes), pool.ctypes.data, len(pool), acts.ctypes.data,
                                       len(acts), C.byref(h))
        if rc:
            raise ValueError("qp_program_from_dag failed (%d)" % rc)
        try:
            ptr = C.c_void_p()
            n = lib().qp_program_code(h, C.byref(ptr))
            code = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
            n = lib().qp_program_pool(h, C.byref(ptr))
            out_pool = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,)).copy()
            return code, out_pool, int(lib().qp_program_regs(h))
        finally:
            lib().qp_program_free(h)

    def compile(self):
        """-> (code uint64[], pool uint64[], n_regs).  Nodes are scheduled lazily in action order
        and registers are reused after a node's last use."""
        nodes = self.nodes
        # schedule: post-order from each action's root
        order, seen = [], set()
        sched = []  # ("node", i) | ("act", op, i)
        for op, root, k in self.actions:
            stack = [(root, False)]
            while stack:
                i, done = stack.pop()
                if i in seen:
                    continue
                nop, a, b = nodes[i]
                if done or nop in (OP_LDW, OP_LDK, OP_LDP, OP_LDI):
                    seen.add(i)
                    order.append(i)
                    sched.append(("node", i))
                    continue
                stack.append((i, True))
                for ch in ((a,) if nop in (OP_MULI, OP_ADDI) else (b, a)):
                    if ch not in seen:
                        stack.append((ch, False))
            sched.append(("act", op, root, k))
        last_use = {}
        for t, s in enumerate(sched):
            if s[0] == "node":
                nop, a, b = nodes[s[1]]
                if nop in (OP_ADD, OP_SUB, OP_MUL):
                    last_use[a] = t
                    last_use[b] = t
                elif nop in (OP_MULI, OP_ADDI):
                    last_use[a] = t
            else:
                last_use[s[2]] = t
        reg_of, free, n_regs, code = {}, [], 0, []
        for t, s in enumerate(sched):
            if s[0] == "node":
                i = s[1]
                nop, a, b = nodes[i]
                ra = rb = rc = 0
                if nop in (OP_ADD, OP_SUB, OP_MUL):
                    ra, rb = reg_of[a], reg_of[b]
                    for ch in {a, b}:
                        if last_use.get(ch) == t:
                            free.append(reg_of[ch])
                elif nop in (OP_MULI, OP_ADDI):
                    ra, rc = reg_of[a], b
                    if last_use.get(a) == t:
                        free.append(reg_of[a])
                else:
                    rc = a
                if free:
                    r = free.pop()
                else:
                    r = n_regs
                    n_regs += 1
                reg_of[i] = r
                code.append(encode_word(nop, r, ra, rb, rc))
                if nop in (OP_LDW, OP_LDK):
                    # column loads are asynchronous on the device; this mirror does not overlap
                    # them (the native compiler hoists them three loads ahead): wait at once
                    code.append(OP_WAIT)
                if i not in last_use:  # dead value
                    free.append(r)
            else:
                _, op, root, k = s
                code.append(encode_word(op, 0, reg_of[root], 0, k))
                if last_use.get(root) == t:
                    free.append(reg_of[root])
        assert n_regs <= 256
        return (np.array(code, dtype=np.uint64), np.array(self.pool or [0], dtype=np.uint64), max(n_regs, 1))


# ---- gates ------------------------------------------------------------------------------------------
(GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC, GATE_POSEIDON, GATE_ARITHMETIC_EXT, GATE_MUL_EXT,
 GATE_BASE_SUM_2, GATE_RANDOM_ACCESS, GATE_REDUCING, GATE_REDUCING_EXT, GATE_POSEIDON_MDS, GATE_EXPONENTIATION,
 GATE_COSET_INTERPOLATION, GATE_LOOKUP, GATE_LOOKUP_TABLE) = range(16)  # qp_plonky2_host.h

# Debug rendering of PhantomData<F> inside gate ids: core::any::type_name of the field type.  The
# field crate's package is `qp-plonky2-field` with no [lib] rename (field/Cargo.toml:2), so the path
# starts with `qp_plonky2_field`.
_PHANTOM = "PhantomData<qp_plonky2_field::goldilocks_field::GoldilocksField>"


class Gate:
    kind = None   # QP_GATE_* of the native host library
    param = 0

    def id(self):
        raise NotImplementedError

    degree = 0
    num_constants = 0
    num_constraints = 0

    def eval_unfiltered(self, consts, wires, pih):
        """consts(i) / wires(i) / pih(i) -> values; returns the list of constraints."""
        return []


class NoopGate(Gate):  # plonky2/src/gates/noop.rs
    kind = GATE_NOOP

    def id(self):
        return "NoopGate"


class ConstantGate(Gate):  # plonky2/src/gates/constant.rs
    degree = 1

    kind = GATE_CONSTANT

    def __init__(self, num_consts):
        self.num_consts = self.num_constants = self.num_constraints = self.param = num_consts

    def id(self):
        return "ConstantGate { num_consts: %d }" % self.num_consts

    def eval_unfiltered(self, consts, wires, pih):
        return [consts(i) - wires(i) for i in range(self.num_consts)]  # constant.rs:121-129


class PublicInputGate(Gate):  # plonky2/src/gates/public_input.rs
    degree = 1
    num_constraints = 4
    kind = GATE_PUBLIC_INPUT

    def id(self):
        return "PublicInputGate"

    def eval_unfiltered(self, consts, wires, pih):
        return [wires(i) - pih(i) for i in range(4)]  # public_input.rs:103-113


class ArithmeticGate(Gate):  # plonky2/src/gates/arithmetic_base.rs
    degree = 3
    num_constants = 2

    kind = GATE_ARITHMETIC

    def __init__(self, num_ops):
        self.num_ops = self.num_constraints = self.param = num_ops

    @staticmethod
    def new_from_config(num_routed_wires):
        return ArithmeticGate(num_routed_wires // 4)  # arithmetic_base.rs:44-47

    def id(self):
        return "ArithmeticGate { num_ops: %d }" % self.num_ops

    def eval_unfiltered(self, consts, wires, pih):
        c0, c1 = consts(0), consts(1)
        out = []
        for i in range(self.num_ops):  # arithmetic_base.rs:168-185
            m0, m1, addend, output = wires(4 * i), wires(4 * i + 1), wires(4 * i + 2), wires(4 * i + 3)
            out.append(output - (m0 * m1 * c0 + addend * c1))
        return out


def _ext_mul(x, y):
    """(a0 + a1 X)(b0 + b1 X) mod X^2 - 7, field/src/extension/quadratic.rs:186-199."""
    return (x[0] * y[0] + (x[1] * y[1]) * 7, x[0] * y[1] + x[1] * y[0])


class ArithmeticExtensionGate(Gate):  # plonky2/src/gates/arithmetic_extension.rs (D = 2)
    degree = 3
    num_constants = 2
    kind = GATE_ARITHMETIC_EXT

    def __init__(self, num_ops):
        self.num_ops = self.param = num_ops
        self.num_constraints = 2 * num_ops

    @staticmethod
    def new_from_config(num_routed_wires):
        return ArithmeticExtensionGate(num_routed_wires // 8)

    def id(self):
        return "ArithmeticExtensionGate { num_ops: %d }" % self.num_ops

    def eval_unfiltered(self, consts, wires, pih):
        c0, c1 = consts(0), consts(1)
        out = []
        for i in range(self.num_ops):  # arithmetic_extension.rs:92-110
            w = [wires(8 * i + k) for k in range(8)]
            pr = _ext_mul((w[0], w[1]), (w[2], w[3]))
            out.append(w[6] - (pr[0] * c0 + w[4] * c1))
            out.append(w[7] - (pr[1] * c0 + w[5] * c1))
        return out


class MulExtensionGate(Gate):  # plonky2/src/gates/multiplication_extension.rs (D = 2)
    degree = 3
    num_constants = 1
    kind = GATE_MUL_EXT

    def __init__(self, num_ops):
        self.num_ops = self.param = num_ops
        self.num_constraints = 2 * num_ops

    @staticmethod
    def new_from_config(num_routed_wires):
        return MulExtensionGate(num_routed_wires // 6)

    def id(self):
        return "MulExtensionGate { num_ops: %d }" % self.num_ops

    def eval_unfiltered(self, consts, wires, pih):
        c0 = consts(0)
        out = []
        for i in range(self.num_ops):  # multiplication_extension.rs:86-101
            w = [wires(6 * i + k) for k in range(6)]
            pr = _ext_mul((w[0], w[1]), (w[2], w[3]))
            out.append(w[4] - pr[0] * c0)
            out.append(w[5] - pr[1] * c0)
        return out


class BaseSumGate2(Gate):  # plonky2/src/gates/base_sum.rs with B = 2
    degree = 2
    kind = GATE_BASE_SUM_2

    def __init__(self, num_limbs):
        self.num_limbs = self.param = num_limbs
        self.num_constraints = 1 + num_limbs

    def id(self):
        return "BaseSumGate { num_limbs: %d } + Base: 2" % self.num_limbs

    def eval_unfiltered(self, consts, wires, pih):
        acc = wires(self.num_limbs)           # reduce_with_powers(limbs, 2), base_sum.rs:153-170
        for i in range(self.num_limbs - 1, 0, -1):
            acc = acc * 2 + wires(i)
        out = [acc - wires(0)]
        for i in range(1, self.num_limbs + 1):
            out.append(wires(i) * (wires(i) - 1))
        return out


def _ext_add(x, y):
    return (x[0] + y[0], x[1] + y[1])


def _ext_sub(x, y):
    return (x[0] - y[0], x[1] - y[1])


def _ext_scalar(x, k):
    return (x[0] * k, x[1] * k)


def primitive_root_of_unity(n_log):
    """field/src/types.rs:280-284 with POWER_OF_TWO_GENERATOR of goldilocks_field.rs:91."""
    g = 7277203076849721926
    for _ in range(32 - n_log):
        g = g * g % P
    return g


def two_adic_subgroup(n_log):
    """field/src/types.rs:292-295."""
    g, out, v = primitive_root_of_unity(n_log), [], 1
    for _ in range(1 << n_log):
        out.append(v)
        v = v * g % P
    return out


def barycentric_weights(xs):
    """field/src/interpolation.rs:53-65."""
    out = []
    for i, xi in enumerate(xs):
        d = 1
        for j, xj in enumerate(xs):
            if j != i:
                d = d * (xi - xj) % P
        out.append(pow(d, P - 2, P))
    return out


class RandomAccessGate(Gate):  # plonky2/src/gates/random_access.rs
    kind = GATE_RANDOM_ACCESS

    def __init__(self, num_copies, bits, num_extra_constants):
        self.num_copies, self.bits, self.num_extra_constants = num_copies, bits, num_extra_constants
        assert bits < 256 and num_copies < 256 and num_extra_constants < 65536
        self.param = bits | (num_copies << 8) | (num_extra_constants << 16)
        self.degree = bits + 1                                                # random_access.rs:275-277
        self.num_constants = num_extra_constants
        self.num_constraints = num_copies * (bits + 2) + num_extra_constants  # :279-282

    @staticmethod
    def new_from_config(num_wires, num_routed_wires, bits, config_num_constants=2):
        vec_size = 1 << bits                                                  # random_access.rs:58-76
        max_copies = min(num_routed_wires // (2 + vec_size), num_wires // (2 + vec_size + bits))
        max_extra = num_routed_wires - (2 + vec_size) * max_copies
        return RandomAccessGate(max_copies, bits, min(max_extra, config_num_constants))

    def id(self):
        return "RandomAccessGate { bits: %d, num_copies: %d, num_extra_constants: %d, _phantom: %s }<D=2>" % (
            self.bits, self.num_copies, self.num_extra_constants, _PHANTOM)

    vec_size = property(lambda self: 1 << self.bits)

    def wire_access_index(self, copy):
        return (2 + self.vec_size) * copy

    def wire_claimed_element(self, copy):
        return (2 + self.vec_size) * copy + 1

    def wire_list_item(self, i, copy):
        return (2 + self.vec_size) * copy + 2 + i

    def wire_extra_constant(self, i):
        return (2 + self.vec_size) * self.num_copies + i

    def num_routed_wires(self):
        return (2 + self.vec_size) * self.num_copies + self.num_extra_constants

    def wire_bit(self, i, copy):
        return self.num_routed_wires() + copy * self.bits + i

    def eval_unfiltered(self, consts, wires, pih):
        out = []
        for copy in range(self.num_copies):  # random_access.rs:144-189 (packed form :307-352)
            access_index = wires(self.wire_access_index(copy))
            items = [wires(self.wire_list_item(i, copy)) for i in range(self.vec_size)]
            claimed = wires(self.wire_claimed_element(copy))
            bits = [wires(self.wire_bit(i, copy)) for i in range(self.bits)]
            for b in bits:
                out.append(b * (b - 1))
            acc = None
            for b in reversed(bits):
                acc = b if acc is None else acc * 2 + b   # fold from ZERO: 0.double() + b = b
            out.append(acc - access_index)
            for b in bits:
                items = [x + b * (y - x) for x, y in zip(items[0::2], items[1::2])]
            out.append(items[0] - claimed)
        for i in range(self.num_extra_constants):
            out.append(consts(i) - wires(self.wire_extra_constant(i)))
        return out


class ReducingGate(Gate):  # plonky2/src/gates/reducing.rs (D = 2)
    degree = 2
    kind = GATE_REDUCING
    EXT_COEFFS = False

    def __init__(self, num_coeffs):
        assert num_coeffs > 0
        self.num_coeffs = self.param = num_coeffs
        self.num_constraints = 2 * num_coeffs

    @staticmethod
    def max_coeffs_len(num_wires, num_routed_wires):
        return min(num_routed_wires - 6, (num_wires - 4) // 3)  # reducing.rs:36-38

    def id(self):
        return "ReducingGate { num_coeffs: %d }" % self.num_coeffs

    def coeff(self, wires, i):
        return (wires(6 + i), 0)

    def start_accs(self):
        return 6 + self.num_coeffs

    def wires_accs(self, i):
        return 0 if i == self.num_coeffs - 1 else self.start_accs() + 2 * i

    def eval_unfiltered(self, consts, wires, pih):
        ext = lambda w: (wires(w), wires(w + 1))
        alpha, acc = ext(2), ext(4)
        out = []
        for i in range(self.num_coeffs):  # reducing.rs:109-133
            nxt = ext(self.wires_accs(i))
            c = self.coeff(wires, i)
            t = _ext_mul(acc, alpha)
            out.append(t[0] + c[0] - nxt[0])
            out.append((t[1] + c[1] - nxt[1]) if self.EXT_COEFFS else (t[1] - nxt[1]))
            acc = nxt
        return out


class ReducingExtensionGate(ReducingGate):  # plonky2/src/gates/reducing_extension.rs (D = 2)
    kind = GATE_REDUCING_EXT
    EXT_COEFFS = True

    @staticmethod
    def max_coeffs_len(num_wires, num_routed_wires):
        return min((num_routed_wires - 6) // 2, (num_wires - 4) // 4)  # reducing_extension.rs:37-41

    def id(self):
        return "ReducingExtensionGate { num_coeffs: %d }" % self.num_coeffs

    def coeff(self, wires, i):
        return (wires(6 + 2 * i), wires(7 + 2 * i))

    def start_accs(self):
        return 6 + 2 * self.num_coeffs


class PoseidonMdsGate(Gate):  # plonky2/src/gates/poseidon_mds.rs (D = 2)
    degree = 1
    kind = GATE_POSEIDON_MDS
    num_constraints = 24

    def id(self):
        return "PoseidonMdsGate(%s)<WIDTH=12>" % _PHANTOM

    @staticmethod
    def _mds_ext(inp):
        k = _poseidon_constants()
        circ, diag = k["POSEIDON_MDS_CIRC"], k["POSEIDON_MDS_DIAG"]
        out = []
        for r in range(12):  # mds_row_shf_field, core/src/poseidon.rs:200-215
            acc = _ext_scalar(inp[r], circ[0] + diag[r])
            for i in range(1, 12):
                acc = _ext_add(acc, _ext_scalar(inp[(i + r) % 12], circ[i]))
            out.append(acc)
        return out

    def eval_unfiltered(self, consts, wires, pih):
        inp = [(wires(2 * i), wires(2 * i + 1)) for i in range(12)]
        comp = self._mds_ext(inp)
        out = []
        for i in range(12):  # poseidon_mds.rs:150-169
            out.append(wires(24 + 2 * i) - comp[i][0])
            out.append(wires(25 + 2 * i) - comp[i][1])
        return out


class ExponentiationGate(Gate):  # plonky2/src/gates/exponentiation.rs
    degree = 4
    kind = GATE_EXPONENTIATION

    def __init__(self, num_power_bits):
        self.num_power_bits = self.param = num_power_bits
        self.num_constraints = num_power_bits + 1

    @staticmethod
    def new_from_config(num_wires, num_routed_wires):
        return ExponentiationGate(min(num_routed_wires - 2, (num_wires - 2) // 2))  # exponentiation.rs:51-60

    def id(self):
        return "ExponentiationGate { num_power_bits: %d, _phantom: %s }<D=2>" % (self.num_power_bits, _PHANTOM)

    def eval_unfiltered(self, consts, wires, pih):
        n = self.num_power_bits
        base, output = wires(0), wires(1 + n)
        out = []
        for i in range(n):  # exponentiation.rs:210-245
            cur_bit = wires(1 + (n - i - 1))
            mul_by = cur_bit * base + (1 - cur_bit)
            if i == 0:
                computed = mul_by
            else:
                prev = wires(2 + n + i - 1)
                computed = prev * prev * mul_by
            out.append(computed - wires(2 + n + i))
        out.append(output - wires(2 + n + n - 1))
        return out


class CosetInterpolationGate(Gate):  # plonky2/src/gates/coset_interpolation.rs (D = 2)
    kind = GATE_COSET_INTERPOLATION

    def __init__(self, subgroup_bits, degree):
        self.subgroup_bits, self.degree = subgroup_bits, degree
        self.param = subgroup_bits | (degree << 8)
        self.num_points = 1 << subgroup_bits
        self.num_intermediates = (self.num_points - 2) // (degree - 1)
        self.num_constraints = 1 + 2 + 2 + 4 * self.num_intermediates   # coset_interpolation.rs:406-410
        self.domain = two_adic_subgroup(subgroup_bits)
        self.weights = barycentric_weights(self.domain)
        self.start_intermediates = 1 + 2 * self.num_points + 4

    @staticmethod
    def with_max_degree(subgroup_bits, max_degree):
        n_points = 1 << subgroup_bits        # coset_interpolation.rs:49-75
        n_intermediates = (n_points - 2) // (max_degree - 1)
        return CosetInterpolationGate(subgroup_bits, (n_points - 2) // (n_intermediates + 1) + 2)

    def id(self):
        return "CosetInterpolationGate { subgroup_bits: %d, degree: %d, barycentric_weights: [%s], _phantom: %s }<D=2>" % (
            self.subgroup_bits, self.degree, ", ".join(str(w) for w in self.weights), _PHANTOM)

    def wires_value(self, i):
        return 1 + 2 * i

    def wires_evaluation_point(self):
        return 1 + 2 * self.num_points

    def wires_evaluation_value(self):
        return 3 + 2 * self.num_points

    def wires_intermediate_eval(self, i):
        return self.start_intermediates + 2 * i

    def wires_intermediate_prod(self, i):
        return self.start_intermediates + 2 * (self.num_intermediates + i)

    def wires_shifted_evaluation_point(self):
        return self.start_intermediates + 4 * self.num_intermediates

    def wire_shift_inverse(self):
        return self.start_intermediates + 2 * (2 * self.num_intermediates + 1)

    def num_wires(self):
        return self.wire_shift_inverse() + 1

    def _partial(self, lo, hi, values, x, ev, prod):
        """partial_interpolate, coset_interpolation.rs:572-599."""
        for j in range(lo, hi):
            val = _ext_scalar(values[j], self.weights[j])
            term = (x[0] - self.domain[j], x[1])
            ev = _ext_add(_ext_mul(ev, term), _ext_mul(val, prod))
            prod = _ext_mul(prod, term)
        return ev, prod

    def _first(self, values, x):
        # initial_eval = 0, initial_partial_prod = 1: the first step is (w_0 v_0, x - x_0)
        ev = _ext_scalar(values[0], self.weights[0])
        prod = (x[0] - self.domain[0], x[1])
        return self._partial(1, self.degree, values, x, ev, prod)

    def eval_unfiltered(self, consts, wires, pih):
        ext = lambda w: (wires(w), wires(w + 1))
        shift, shift_inv = wires(0), wires(self.wire_shift_inverse())
        point, shifted = ext(self.wires_evaluation_point()), ext(self.wires_shifted_evaluation_point())
        out = [shift * shift_inv - 1]          # coset_interpolation.rs:260-307
        out += list(_ext_sub(point, _ext_scalar(shifted, shift)))
        values = [ext(self.wires_value(i)) for i in range(self.num_points)]
        ev, prod = self._first(values, shifted)
        for i in range(self.num_intermediates):
            iev, iprod = ext(self.wires_intermediate_eval(i)), ext(self.wires_intermediate_prod(i))
            out += list(_ext_sub(iev, ev)) + list(_ext_sub(iprod, prod))
            start = 1 + (self.degree - 1) * (i + 1)
            end = min(start + self.degree - 1, self.num_points)
            ev, prod = self._partial(start, end, values, shifted, iev, iprod)
        out += list(_ext_sub(ext(self.wires_evaluation_value()), ev))
        return out


def _poseidon_constants():
    """The tables of core/src/poseidon_goldilocks.rs as generated into csrc/poseidon_constants.h
    (tools/gen_poseidon_constants.py)."""
    import os
    import re
    global _PC
    if _PC is None:
        text = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "poseidon_constants.h")).read()
        _PC = {}
        for m in re.finditer(r"static const uint64_t (\w+)\[(\d+)\] = \{(.*?)\};", text, re.S):
            _PC[m.group(1)] = [int(t, 16) for t in re.findall(r"0x([0-9a-fA-F]+)ULL", m.group(3))]
    return _PC


_PC = None


class PoseidonGate(Gate):
    """plonky2/src/gates/poseidon.rs: one width-12 permutation per row, S-box inputs as wires.
    `arith` supplies the value type: the recording `Val`s for the constraint program, plain
    integers mod p for witness generation (the gate's generator, poseidon.rs:424-520)."""
    degree = 7
    kind = GATE_POSEIDON
    num_constraints = 12 * 7 + 22 + 12 + 1 + 4   # poseidon.rs:416-422
    WIRE_SWAP = 24
    START_DELTA = 25
    START_FULL_0 = 29
    START_PARTIAL = 29 + 12 * 3
    START_FULL_1 = 29 + 12 * 3 + 22
    END = 29 + 12 * 3 + 22 + 12 * 4            # 135 wires

    def id(self):
        return "PoseidonGate(%s)<WIDTH=12>" % _PHANTOM

    @staticmethod
    def _mds(state):
        k = _poseidon_constants()
        circ, diag = k["POSEIDON_MDS_CIRC"], k["POSEIDON_MDS_DIAG"]
        out = []
        for r in range(12):   # core/src/poseidon.rs:178-198
            acc = state[r] * (circ[0] + diag[r])
            for i in range(1, 12):
                acc = acc + state[(i + r) % 12] * circ[i]
            out.append(acc)
        return out

    @staticmethod
    def _sbox(x):
        x2 = x * x
        x4 = x2 * x2
        return x * x2 * x4

    def _run(self, wires, on_sbox_in, on_constraint):
        """The body shared by eval_unfiltered and the witness generator (poseidon.rs:204-283).
        on_sbox_in(wire_index, computed_state) -> value to continue with."""
        k = _poseidon_constants()
        rc = k["POSEIDON_ALL_ROUND_CONSTANTS"]
        swap = wires(self.WIRE_SWAP)
        on_constraint(swap * (swap - 1))
        state = [None] * 12
        for i in range(4):
            lhs, rhs = wires(i), wires(i + 4)
            delta = on_sbox_in(self.START_DELTA + i, swap * (rhs - lhs), negate=True)
            state[i], state[i + 4] = lhs + delta, rhs - delta
        for i in range(8, 12):
            state[i] = wires(i)
        rnd = 0
        for r in range(4):
            state = [state[i] + rc[12 * rnd + i] for i in range(12)]
            if r != 0:
                state = [on_sbox_in(self.START_FULL_0 + 12 * (r - 1) + i, state[i]) for i in range(12)]
            state = self._mds([self._sbox(x) for x in state])
            rnd += 1
        # partial rounds, fast form (core/src/poseidon.rs:302-342,378-408,584-596)
        state = [state[i] + k["POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT"][i] for i in range(12)]
        init = k["POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX"]
        t = [state[0]] + [None] * 11
        for c in range(1, 12):
            acc = state[1] * init[c - 1]
            for r in range(2, 12):
                acc = acc + state[r] * init[(r - 1) * 11 + (c - 1)]
            t[c] = acc
        state = t
        m00 = k["POSEIDON_MDS_CIRC"][0] + k["POSEIDON_MDS_DIAG"][0]
        for r in range(22):
            s0 = self._sbox(on_sbox_in(self.START_PARTIAL + r, state[0]))
            if r != 21:
                s0 = s0 + k["POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS"][r]
            w_hat = k["POSEIDON_FAST_PARTIAL_ROUND_W_HATS"][11 * r: 11 * r + 11]
            vs = k["POSEIDON_FAST_PARTIAL_ROUND_VS"][11 * r: 11 * r + 11]
            d = s0 * m00
            for j in range(1, 12):
                d = d + state[j] * w_hat[j - 1]
            state = [d] + [state[j] + s0 * vs[j - 1] for j in range(1, 12)]
        rnd += 22
        for r in range(4):
            state = [state[i] + rc[12 * rnd + i] for i in range(12)]
            state = [on_sbox_in(self.START_FULL_1 + 12 * r + i, state[i]) for i in range(12)]
            state = self._mds([self._sbox(x) for x in state])
            rnd += 1
        return state

    def eval_unfiltered(self, consts, wires, pih):
        cons = []

        def on_sbox_in(w, computed, negate=False):
            cons.append(computed - wires(w))   # delta: swap*(rhs-lhs) - delta_i; S-box: state - sbox_in
            return wires(w)

        out = self._run(wires, on_sbox_in, cons.append)
        for i in range(12):
            cons.append(out[i] - wires(12 + i))
        return cons


def keccak256(data):
    """keccak_hash::keccak (Keccak-256 with the original 0x01 padding; hashlib's sha3_256 pads with 0x06)."""
    rc = [0x0000000000000001, 0x0000000000008082, 0x800000000000808a, 0x8000000080008000, 0x000000000000808b,
          0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008a, 0x0000000000000088,
          0x0000000080008009, 0x000000008000000a, 0x000000008000808b, 0x800000000000008b, 0x8000000000008089,
          0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800a, 0x800000008000000a,
          0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
    m64 = (1 << 64) - 1
    rol = lambda x, k: ((x << k) | (x >> (64 - k))) & m64 if k else x
    msg = bytearray(data) + b"\x01"
    msg += b"\x00" * (-len(msg) % 136)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]   # a[x][y]
    for off in range(0, len(msg), 136):
        for i in range(17):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        for rnd in range(24):
            c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
            d = [c[(x - 1) % 5] ^ rol(c[(x + 1) % 5], 1) for x in range(5)]
            a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
            b = [[0] * 5 for _ in range(5)]
            x, y, cur = 1, 0, a[1][0]
            b[0][0] = a[0][0]
            for t in range(24):           # rho + pi along the (x, y) -> (y, 2x + 3y) orbit
                x, y = y, (2 * x + 3 * y) % 5
                nxt = a[x][y]
                b[x][y] = rol(cur, ((t + 1) * (t + 2) // 2) % 64)
                cur = nxt
            a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & m64 & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
            a[0][0] ^= rc[rnd]
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def _lut_hash(lut):
    """keccak over input.to_le_bytes() ++ output.to_le_bytes() of every entry, rendered like `{:?}` of [u8; 32]
    (gates/lookup.rs:44-55, gates/lookup_table.rs:50-62)."""
    data = b"".join(int(i).to_bytes(2, "little") + int(o).to_bytes(2, "little") for i, o in lut)
    return "[" + ", ".join(str(b) for b in keccak256(data)) + "]"


class LookupGate(Gate):  # plonky2/src/gates/lookup.rs: no gate constraints (the lookup argument's terms are global)
    kind = GATE_LOOKUP

    def __init__(self, num_routed_wires, lut, lut_index):
        self.num_slots, self.lut, self.param = num_routed_wires // 2, lut, lut_index

    def id(self):
        return "LookupGate {num_slots: %d, lut_hash: %s}" % (self.num_slots, _lut_hash(self.lut))


class LookupTableGate(Gate):  # plonky2/src/gates/lookup_table.rs
    kind = GATE_LOOKUP_TABLE

    def __init__(self, num_routed_wires, lut, lut_index, last_lut_row):
        self.num_slots, self.lut, self.param, self.last_lut_row = num_routed_wires // 3, lut, lut_index, last_lut_row

    def id(self):
        return "LookupTableGate {num_slots: %d, lut_hash: %s, last_lut_row: %d}" % (
            self.num_slots, _lut_hash(self.lut), self.last_lut_row)


def sort_gates(gates):
    """circuit_builder.rs:1177-1179: by (degree, id)."""
    return sorted(gates, key=lambda g: (g.degree, g.id()))


def selectors_info(gates, max_degree):
    """selector_polynomials' grouping (selectors.rs:99-166) for SORTED gates:
    -> (selector_indices[gate], groups[(start, end)])."""
    num_gates = len(gates)
    max_gate_degree = gates[-1].degree
    if max_gate_degree + num_gates - 1 <= max_degree:
        return [0] * num_gates, [(0, num_gates)]
    if max_gate_degree >= max_degree:
        raise ValueError("%s has too high degree. Consider increasing `quotient_degree_factor`." % gates[-1].id())
    groups, start = [], 0
    while start < num_gates:
        size = 0
        while start + size < num_gates and size + gates[start + size].degree < max_degree:
            size += 1
        groups.append((start, start + size))
        start += size
    idx = []
    for i in range(num_gates):
        idx.append(next(k for k, (a, b) in enumerate(groups) if a <= i < b))
    return idx, groups


def get_unique_coset_shifts(num_shifts):
    """field/src/cosets.rs:9-24: g^0 .. g^(num_shifts-1)."""
    out, v = [], 1
    for _ in range(num_shifts):
        out.append(v)
        v = v * MULTIPLICATIVE_GROUP_GENERATOR % P
    return np.array(out, dtype=np.uint64)


def log2_ceil(x):
    return max(0, (int(x) - 1).bit_length())


class CommonCircuitData:
    """The fields of CommonCircuitData (circuit_data.rs:412-470) that the permutation argument, the lookup
    argument and the quotient use.  `gates` in any order (sorted here like the builder does).  luts: the lookup
    tables ([(input, output), ...] each); lookup_rows: prover_data.lookup_rows, one (last_lu_row, last_lut_row,
    first_lut_row) per table (the LookupGate / LookupTableGate of every table must be in `gates`)."""

    def __init__(self, degree_bits, gates, num_wires=143, num_routed_wires=80, num_challenges=2,
                 quotient_degree_factor=8, rate_bits=3, cap_height=4, luts=None, lookup_rows=None):
        self.luts = [[(int(i), int(o)) for i, o in t] for t in (luts or [])]
        self.lookup_rows = [tuple(int(v) for v in r) for r in (lookup_rows or [])]
        assert len(self.luts) == len(self.lookup_rows)
        self.degree_bits = degree_bits
        self.num_wires, self.num_routed_wires = num_wires, num_routed_wires
        self.num_challenges = num_challenges
        self.quotient_degree_factor = quotient_degree_factor
        self.rate_bits, self.cap_height = rate_bits, cap_height
        self.gates = sort_gates(gates)
        # circuit_builder.rs:1180-1181: selector_polynomials(gates, instances, quotient_degree_factor + 1)
        self.selector_indices, self.groups = selectors_info(self.gates, quotient_degree_factor + 1)
        self.num_selectors = len(self.groups)
        # selectors_lookup + selector_ends_lookups, gates/selectors.rs:27-75; circuit_builder.rs:1183-1194,1284-1290
        self.num_lookup_selectors = 4 + len(self.luts) if self.luts else 0
        self.num_lookup_polys = -(-(num_routed_wires // 2) // (quotient_degree_factor - 1)) + 1 if self.luts else 0
        self.num_gate_constants = max(g.num_constants for g in self.gates)
        self.num_constants = self.num_selectors + self.num_lookup_selectors + self.num_gate_constants  # circuit_builder.rs:1180-1197
        self.num_gate_constraints = max(g.num_constraints for g in self.gates)
        self.k_is = get_unique_coset_shifts(num_routed_wires)
        # util/partial_products.rs:41-48
        self.num_partial_products = -(-num_routed_wires // quotient_degree_factor) - 1
        self.quotient_degree_bits = log2_ceil(quotient_degree_factor)

    def gate_index(self, gate_id):
        return next(i for i, g in enumerate(self.gates) if g.id() == gate_id)

    def constraint_program(self, native_compile=False):
        """evaluate_gate_constraints_base_batch (vanishing_poly.rs:700-726) as a program.  With
        `native_compile` the recording made here goes through the native compiler."""
        prog = ConstraintProgram()
        prefix = self.num_selectors + self.num_lookup_selectors  # gate.rs:179 remove_prefix
        for i, g in enumerate(self.gates):
            sel = self.selector_indices[i]
            start, end = self.groups[sel]
            s = prog.constant(sel)
            # compute_filter, gate.rs:326-333
            factors = [j - s for j in range(start, end) if j != i]
            if self.num_selectors > 1:
                factors.append(UNUSED_SELECTOR - s)
            filt = prog.imm(1)
            for f in factors:
                filt = filt * f
            cons = g.eval_unfiltered(lambda k: prog.constant(prefix + k), prog.wire, prog.public_input_hash)
            assert len(cons) == g.num_constraints
            if cons:
                prog.emit_gate(cons, filt)
        return prog.compile_native() if native_compile else prog.compile()


class _GateDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("param", C.c_uint32)]


class _LookupTable(C.Structure):   # qp_lookup_table, include/qp_plonky2_b200.h
    _fields_ = [("table", C.c_void_p), ("len", C.c_size_t), ("last_lu_row", C.c_uint32), ("last_lut_row", C.c_uint32),
                ("first_lut_row", C.c_uint32)]


def _lookup_tables(luts, lookup_rows):
    """-> (ctypes array of qp_lookup_table, the numpy tables it points into)."""
    keep = [np.ascontiguousarray(t, dtype=np.uint16).reshape(-1, 2) for t in luts]
    arr = (_LookupTable * max(1, len(keep)))()
    for k, (t, r) in enumerate(zip(keep, lookup_rows)):
        arr[k] = _LookupTable(t.ctypes.data, len(t), r[0], r[1], r[2])
    return arr, keep


def native_constraint_program(gates, max_degree, num_routed_wires=0, luts=(), lookup_rows=()):
    """The host library's compiler (qp-plonky2_b200/host/plonk_host.cpp, qp_program_create[_lookups]): the
    program the device runs.  -> dict(code, pool, n_regs, num_selectors, selector_indices, groups,
    order) with gates in the reference's sorted order; `order[i]` = index in `gates` of sorted gate i."""
    from . import lib
    arr = (_GateDesc * len(gates))(*[_GateDesc(g.kind, g.param) for g in gates])
    h = C.c_void_p()
    if luts:
        lt, keep = _lookup_tables(luts, lookup_rows)
        rc = lib().qp_program_create_lookups(arr, len(gates), max_degree, num_routed_wires, lt, len(luts), C.byref(h))
        del keep
    else:
        rc = lib().qp_program_create(arr, len(gates), max_degree, C.byref(h))
    if rc:
        raise ValueError("qp_program_create failed (%d): unsupported gate, or a gate of too high degree" % rc)
    try:
        ptr = C.c_void_p()
        n = lib().qp_program_code(h, C.byref(ptr))
        code = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        n = lib().qp_program_pool(h, C.byref(ptr))
        pool = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        sel, groups, order = [], {}, []
        for i in range(len(gates)):
            o, s_, a, b = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
            lib().qp_program_gate(h, i, C.byref(o), C.byref(s_), C.byref(a), C.byref(b))
            order.append(o.value)
            sel.append(s_.value)
            groups[s_.value] = (a.value, b.value)
        n = lib().qp_program_segments(h, C.byref(ptr))
        segments = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        return dict(code=code, pool=pool, n_regs=int(lib().qp_program_regs(h)), segments=segments,
                    num_selectors=int(lib().qp_program_num_selectors(h)), selector_indices=sel,
                    groups=[groups[k] for k in sorted(groups)], order=order,
                    num_gate_constants=int(lib().qp_program_num_gate_constants(h)),
                    num_gate_constraints=int(lib().qp_program_num_gate_constraints(h)))
    finally:
        lib().qp_program_free(h)


class _Desc(C.Structure):
    _fields_ = [
        ("degree_bits", C.c_uint32), ("quotient_degree_bits", C.c_uint32), ("num_challenges", C.c_uint32),
        ("num_routed_wires", C.c_uint32), ("num_wires", C.c_uint32), ("num_constants", C.c_uint32),
        ("num_partial_products", C.c_uint32), ("max_degree", C.c_uint32),
        ("k_is", C.c_void_p), ("sigmas", C.c_void_p), ("sigmas_space", C.c_int),
        ("program", C.c_void_p), ("program_len", C.c_size_t), ("pool", C.c_void_p), ("pool_len", C.c_size_t),
        ("program_regs", C.c_uint32), ("num_lookup_polys", C.c_uint32), ("num_lookup_selectors", C.c_uint32),
        ("num_selectors", C.c_uint32), ("luts", C.c_void_p), ("n_luts", C.c_size_t),
    ]


class Circuit:
    """Device-resident circuit data: k_is, sigmas (prover_data.sigmas as columns), the compiled
    gate program.  One per circuit, like ProverOnlyCircuitData."""

    def __init__(self, ctx, common, sigmas=None, program_source="native"):
        """program_source: "native" = the host library records and compiles the gates (qp_program_create);
        "dag" = this module's recording, compiled by the host library (qp_program_from_dag -- the route
        of a shim with gates the library does not know); "twin" = recorded and compiled here."""
        from . import _buf, lib  # late: this module is imported by the package
        self.ctx, self.common = ctx, common
        if program_source == "native":
            native = native_constraint_program(common.gates, common.quotient_degree_factor + 1, common.num_routed_wires,
                                               getattr(common, "luts", ()), getattr(common, "lookup_rows", ()))
            assert native["selector_indices"] == common.selector_indices and native["groups"] == common.groups
            code, pool, n_regs = native["code"], native["pool"], native["n_regs"]
        else:
            code, pool, n_regs = common.constraint_program(native_compile=program_source == "dag")
        self.program = (code, pool, n_regs)
        k_is = np.ascontiguousarray(common.k_is, dtype=np.uint64)
        d = _Desc()
        d.degree_bits, d.quotient_degree_bits = common.degree_bits, common.quotient_degree_bits
        d.num_challenges = common.num_challenges
        d.num_routed_wires, d.num_wires = common.num_routed_wires, common.num_wires
        d.num_constants = common.num_constants
        d.num_partial_products, d.max_degree = common.num_partial_products, common.quotient_degree_factor
        d.k_is = k_is.ctypes.data
        keep = None
        if sigmas is not None:
            ptr, space, keep, shape = _buf(sigmas)
            assert tuple(shape) == (common.num_routed_wires, 1 << common.degree_bits)
            d.sigmas, d.sigmas_space = ptr.value, space
        d.program, d.program_len = code.ctypes.data, code.size
        d.pool, d.pool_len = pool.ctypes.data, pool.size
        d.program_regs = n_regs
        # the lookup argument's share of the circuit data (common_data.luts, prover_data.lookup_rows)
        d.num_lookup_polys = getattr(common, "num_lookup_polys", 0)
        d.num_lookup_selectors = getattr(common, "num_lookup_selectors", 0)
        d.num_selectors = common.num_selectors
        luts = getattr(common, "luts", [])
        lt, keep_luts = _lookup_tables(luts, getattr(common, "lookup_rows", []))
        if luts:
            d.luts, d.n_luts = C.addressof(lt), len(luts)
        self._h = C.c_void_p()
        ctx.check(lib().qp_circuit_create(ctx._h, C.byref(d), C.byref(self._h)))
        del keep, keep_luts, lt

    def partial_products_and_zs(self, wires, betas, gammas, out_device=None):
        """all_wires_permutation_partial_products + Z-first ordering (prover.rs:255-261,402-480):
        -> value columns [(nc + nc*np)][n] for the second from_values of prove()."""
        from . import QP_DEVICE, QP_HOST, _buf, _np_ptr, lib
        c = self.common
        ptr, space, keep, shape = _buf(wires)
        assert shape[0] >= c.num_routed_wires and shape[1] == 1 << c.degree_bits
        b = np.ascontiguousarray(betas, dtype=np.uint64)
        g = np.ascontiguousarray(gammas, dtype=np.uint64)
        rows = c.num_challenges * (1 + c.num_partial_products)
        if out_device is not None:
            self.ctx.check(lib().qp_circuit_partial_products_and_zs(
                self._h, ptr, space, _np_ptr(b), _np_ptr(g), C.c_void_p(out_device.data_ptr()), QP_DEVICE))
            return out_device
        out = np.zeros((rows, 1 << c.degree_bits), dtype=np.uint64)
        self.ctx.check(lib().qp_circuit_partial_products_and_zs(
            self._h, ptr, space, _np_ptr(b), _np_ptr(g), _np_ptr(out), QP_HOST))
        return out

    def lookup_polys(self, wires, deltas):
        """compute_all_lookup_polys (prover.rs:489-636) -> value columns [nc * num_lookup_polys][n].
        deltas: [nc][4] (ChallengeA, ChallengeB, ChallengeAlpha, ChallengeDelta per challenge)."""
        from . import QP_HOST, _buf, _np_ptr, lib
        c = self.common
        ptr, space, keep, shape = _buf(wires)
        assert shape[0] >= c.num_routed_wires and shape[1] == 1 << c.degree_bits
        d = np.ascontiguousarray(deltas, dtype=np.uint64)
        assert d.size == 4 * c.num_challenges
        out = np.zeros((c.num_challenges * c.num_lookup_polys, 1 << c.degree_bits), dtype=np.uint64)
        self.ctx.check(lib().qp_circuit_lookup_polys(self._h, ptr, space, _np_ptr(d), _np_ptr(out), QP_HOST))
        return out

    def set_lookup_challenges(self, deltas):
        """The lookup challenges the next quotient evaluation uses (qp_circuit_set_lookup_challenges)."""
        from . import _np_ptr, lib
        d = np.ascontiguousarray(deltas, dtype=np.uint64)
        assert d.size == 4 * self.common.num_challenges
        self.ctx.check(lib().qp_circuit_set_lookup_challenges(self._h, _np_ptr(d)))

    def compute_quotient_polys(self, constants_sigmas, wires, zs_partial_products, betas, gammas, alphas,
                               public_inputs_hash, out_device=None, deltas=None):
        """compute_quotient_polys (prover.rs:640-866) -> coefficients [nc][n << quotient_degree_bits].
        deltas: the lookup challenges of a circuit with lookup tables."""
        from . import QP_DEVICE, QP_HOST, _np_ptr, lib
        c = self.common
        if deltas is not None:
            self.set_lookup_challenges(deltas)
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (betas, gammas, alphas, public_inputs_hash)]
        n_lde = 1 << (c.degree_bits + c.quotient_degree_bits)
        if out_device is not None:
            self.ctx.check(lib().qp_circuit_compute_quotient_polys(
                self._h, constants_sigmas._h, wires._h, zs_partial_products._h, *[_np_ptr(a) for a in arrs],
                C.c_void_p(out_device.data_ptr()), QP_DEVICE))
            return out_device
        out = np.zeros((c.num_challenges, n_lde), dtype=np.uint64)
        self.ctx.check(lib().qp_circuit_compute_quotient_polys(
            self._h, constants_sigmas._h, wires._h, zs_partial_products._h, *[_np_ptr(a) for a in arrs],
            _np_ptr(out), QP_HOST))
        return out

    def free(self):
        from . import lib
        if self._h:
            lib().qp_circuit_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
