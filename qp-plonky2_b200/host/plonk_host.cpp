// Host side of the quotient path (include/qp_plonky2_host.h): the gate set of a circuit compiled
// into the straight-line constraint program that quotient::quotient_kernel interprets.
//
// Reference semantics (reference-relative citations):
//   gate order                      plonky2/src/plonk/circuit_builder.rs:1177-1179  (degree, then id)
//   selector groups                 plonky2/src/gates/selectors.rs:99-166
//   compute_filter                  plonky2/src/gates/gate.rs:326-333
//   evaluate_gate_constraints       plonky2/src/plonk/vanishing_poly.rs:700-726
//   NoopGate / ConstantGate / PublicInputGate / ArithmeticGate / PoseidonGate / ArithmeticExtensionGate /
//   MulExtensionGate / BaseSumGate<2> / RandomAccessGate / ReducingGate / ReducingExtensionGate /
//   PoseidonMdsGate / ExponentiationGate / CosetInterpolationGate
//                                   plonky2/src/gates/{noop,constant,public_input,arithmetic_base,poseidon,
//                                   arithmetic_extension,multiplication_extension,base_sum,random_access,
//                                   reducing,reducing_extension,poseidon_mds,exponentiation,
//                                   coset_interpolation}.rs
// In a Rust build this role is played by the shim's recording field type run over
// Gate::eval_unfiltered_base_one; here the same recording evaluation is written in
// C++ for the gates above.  Pure host logic: no field arithmetic on data happens here.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/qp_plonky2_host.h"
#include "../csrc/poseidon_constants.h"

namespace {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t UNUSED_SELECTOR = 0xFFFFFFFFULL;  // core/src/selectors.rs

enum Op : unsigned { END = 0, LDW, LDK, LDP, LDI, ADD, SUB, MUL, EMIT, GATE, MULI, ADDI, WAIT, FMAI, NATIVE_POSEIDON };

struct Recorder;

// A node of the constraint expression DAG (the recording "field element").
struct Val {
    Recorder* rec = nullptr;
    int idx = -1;
};

struct Recorder {
    struct Node {
        unsigned op;
        int a, b;
    };
    struct Action {
        unsigned op;
        int node;
        unsigned k;
    };
    std::vector<Node> nodes;
    std::map<std::tuple<unsigned, int, int>, int> memo;
    std::vector<uint64_t> pool;
    std::map<uint64_t, int> pool_index;
    std::vector<Action> actions;

    Val node(unsigned op, int a = 0, int b = 0) {
        if ((op == ADD || op == MUL) && a > b) std::swap(a, b);
        auto key = std::make_tuple(op, a, b);
        auto it = memo.find(key);
        if (it == memo.end()) {
            nodes.push_back({op, a, b});
            it = memo.emplace(key, (int)nodes.size() - 1).first;
        }
        return Val{this, it->second};
    }
    int pool_slot(uint64_t v) {
        v %= P;
        auto it = pool_index.find(v);
        if (it == pool_index.end()) {
            pool.push_back(v);
            it = pool_index.emplace(v, (int)pool.size() - 1).first;
        }
        return it->second;
    }
    Val wire(int i) { return node(LDW, i); }
    Val constant(int i) { return node(LDK, i); }
    Val pih(int i) { return node(LDP, i); }
    Val imm(uint64_t v) { return node(LDI, pool_slot(v)); }
    void emit_gate(const std::vector<Val>& constraints, Val filter) {
        for (size_t k = 0; k < constraints.size(); k++) actions.push_back({EMIT, constraints[k].idx, (unsigned)k});
        actions.push_back({GATE, filter.idx, 0});
        memo.clear();  // values are not shared across gates (register lifetime)
    }
};

inline Val operator+(Val a, Val b) { return a.rec->node(ADD, a.idx, b.idx); }
inline Val operator-(Val a, Val b) { return a.rec->node(SUB, a.idx, b.idx); }
inline Val operator*(Val a, Val b) { return a.rec->node(MUL, a.idx, b.idx); }
inline Val operator+(Val a, uint64_t c) { return a.rec->node(ADDI, a.idx, a.rec->pool_slot(c)); }
inline Val operator-(Val a, uint64_t c) { return a.rec->node(ADDI, a.idx, a.rec->pool_slot(P - c % P)); }
inline Val operator*(Val a, uint64_t c) { return a.rec->node(MULI, a.idx, a.rec->pool_slot(c)); }

// ---- gates ---------------------------------------------------------------------------------------
struct GateInfo {
    uint32_t kind, param;
    unsigned degree, num_constants, num_constraints;
    std::string id;
};

// Debug rendering of PhantomData<F>: core::any::type_name of the field type; the field crate is the
// package `qp-plonky2-field` with no [lib] rename (field/Cargo.toml:2).
const char* const PHANTOM = "PhantomData<qp_plonky2_field::goldilocks_field::GoldilocksField>";

// host-side constants of the gates (domain, barycentric weights): plain modular arithmetic
uint64_t mulmod(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % P); }
uint64_t powmod(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    for (; e; e >>= 1, a = mulmod(a, a))
        if (e & 1) r = mulmod(r, a);
    return r;
}
std::vector<uint64_t> two_adic_subgroup(unsigned n_log) {  // field/src/types.rs:280-295
    uint64_t g = 7277203076849721926ULL;                   // POWER_OF_TWO_GENERATOR, goldilocks_field.rs:91
    for (unsigned i = n_log; i < 32; i++) g = mulmod(g, g);
    std::vector<uint64_t> out(1u << n_log);
    uint64_t v = 1;
    for (auto& o : out) {
        o = v;
        v = mulmod(v, g);
    }
    return out;
}
std::vector<uint64_t> barycentric_weights(const std::vector<uint64_t>& xs) {  // field/src/interpolation.rs:53-65
    std::vector<uint64_t> out;
    for (size_t i = 0; i < xs.size(); i++) {
        uint64_t d = 1;
        for (size_t j = 0; j < xs.size(); j++)
            if (j != i) d = mulmod(d, xs[i] >= xs[j] ? xs[i] - xs[j] : xs[i] + (P - xs[j]));
        out.push_back(powmod(d, P - 2));
    }
    return out;
}

struct RandomAccessShape {  // random_access.rs:78-127
    unsigned bits, copies, extra, vec;
    explicit RandomAccessShape(uint32_t param)
        : bits(param & 0xFF), copies((param >> 8) & 0xFF), extra(param >> 16), vec(1u << (param & 0xFF)) {}
    int access_index(unsigned c) const { return (int)((2 + vec) * c); }
    int claimed(unsigned c) const { return (int)((2 + vec) * c + 1); }
    int item(unsigned i, unsigned c) const { return (int)((2 + vec) * c + 2 + i); }
    int extra_constant(unsigned i) const { return (int)((2 + vec) * copies + i); }
    int bit(unsigned i, unsigned c) const { return (int)((2 + vec) * copies + extra + c * bits + i); }
};

struct CosetInterpolationShape {  // coset_interpolation.rs:77-155
    unsigned bits, degree, points, inter, start_inter;
    explicit CosetInterpolationShape(uint32_t param)
        : bits(param & 0xFF), degree(param >> 8), points(1u << (param & 0xFF)) {
        inter = degree > 1 ? (points - 2) / (degree - 1) : 0;
        start_inter = 1 + 2 * points + 4;
    }
    int value(unsigned i) const { return (int)(1 + 2 * i); }
    int point() const { return (int)(1 + 2 * points); }
    int eval_value() const { return (int)(3 + 2 * points); }
    int inter_eval(unsigned i) const { return (int)(start_inter + 2 * i); }
    int inter_prod(unsigned i) const { return (int)(start_inter + 2 * (inter + i)); }
    int shifted_point() const { return (int)(start_inter + 4 * inter); }
    int shift_inverse() const { return (int)(start_inter + 2 * (2 * inter + 1)); }
};

GateInfo gate_info(uint32_t kind, uint32_t param) {
    switch (kind) {
        case QP_GATE_RANDOM_ACCESS: {
            RandomAccessShape g(param);
            if (!g.bits || g.bits > 6 || !g.copies) break;
            return {kind, param, g.bits + 1, g.extra, g.copies * (g.bits + 2) + g.extra,
                    "RandomAccessGate { bits: " + std::to_string(g.bits) + ", num_copies: " + std::to_string(g.copies) +
                        ", num_extra_constants: " + std::to_string(g.extra) + ", _phantom: " + PHANTOM + " }<D=2>"};
        }
        case QP_GATE_REDUCING:
            if (!param) break;
            return {kind, param, 2, 0, 2 * param, "ReducingGate { num_coeffs: " + std::to_string(param) + " }"};
        case QP_GATE_REDUCING_EXT:
            if (!param) break;
            return {kind, param, 2, 0, 2 * param, "ReducingExtensionGate { num_coeffs: " + std::to_string(param) + " }"};
        case QP_GATE_POSEIDON_MDS:
            return {kind, param, 1, 0, 24, std::string("PoseidonMdsGate(") + PHANTOM + ")<WIDTH=12>"};
        case QP_GATE_EXPONENTIATION:
            if (!param) break;
            return {kind, param, 4, 0, param + 1,
                    "ExponentiationGate { num_power_bits: " + std::to_string(param) + ", _phantom: " + PHANTOM + " }<D=2>"};
        case QP_GATE_COSET_INTERPOLATION: {
            CosetInterpolationShape g(param);
            if (!g.bits || g.bits > 6 || g.degree < 2) break;
            std::string w;
            for (uint64_t x : barycentric_weights(two_adic_subgroup(g.bits))) w += (w.empty() ? "" : ", ") + std::to_string(x);
            return {kind, param, g.degree, 0, 5 + 4 * g.inter,
                    "CosetInterpolationGate { subgroup_bits: " + std::to_string(g.bits) + ", degree: " +
                        std::to_string(g.degree) + ", barycentric_weights: [" + w + "], _phantom: " + PHANTOM + " }<D=2>"};
        }
        case QP_GATE_NOOP: return {kind, param, 0, 0, 0, "NoopGate"};
        case QP_GATE_CONSTANT:
            return {kind, param, 1, param, param, "ConstantGate { num_consts: " + std::to_string(param) + " }"};
        case QP_GATE_PUBLIC_INPUT: return {kind, param, 1, 0, 4, "PublicInputGate"};
        case QP_GATE_ARITHMETIC:
            return {kind, param, 3, 2, param, "ArithmeticGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_POSEIDON:
            return {kind, param, 7, 0, 123,
                    std::string("PoseidonGate(") + PHANTOM + ")<WIDTH=12>"};
        case QP_GATE_ARITHMETIC_EXT:
            return {kind, param, 3, 2, 2 * param, "ArithmeticExtensionGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_MUL_EXT:
            return {kind, param, 3, 1, 2 * param, "MulExtensionGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_BASE_SUM_2:
            return {kind, param, 2, 0, 1 + param, "BaseSumGate { num_limbs: " + std::to_string(param) + " } + Base: 2"};
    }
    return {kind, param, 0, 0, 0, ""};
}

// Keccak-f[1600] / Keccak-256 with the original 0x01 padding (keccak_hash::keccak = tiny-keccak's Keccak::v256):
// the lut_hash of LookupGate / LookupTableGate (gates/lookup.rs:44-55, gates/lookup_table.rs:50-62).
void keccak_f(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            const uint64_t t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            const int j = PIL[i];
            const uint64_t b = st[j];
            st[j] = (t << ROT[i]) | (t >> (64 - ROT[i]));
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];
    }
}

std::string lut_hash_debug(const qp_lookup_table& t) {  // "{:?}" of [u8; 32]
    std::vector<uint8_t> bytes;  // input.to_le_bytes() ++ output.to_le_bytes() per entry
    for (size_t i = 0; i < 2 * t.len; i++) {
        bytes.push_back((uint8_t)(t.table[i] & 0xff));
        bytes.push_back((uint8_t)(t.table[i] >> 8));
    }
    uint8_t h[32];
    qp_keccak256(bytes.data(), bytes.size(), h);
    std::string out = "[";
    for (int i = 0; i < 32; i++) out += (i ? ", " : "") + std::to_string((unsigned)h[i]);
    return out + "]";
}

// F_p^2 = F_p[X]/(X^2 - 7) over recording values (field/src/extension/quadratic.rs:186-199)
struct ExtVal {
    Val a, b;
};
ExtVal ext_mul(ExtVal x, ExtVal y) { return ExtVal{x.a * y.a + (x.b * y.b) * 7, x.a * y.b + x.b * y.a}; }
ExtVal ext_add(ExtVal x, ExtVal y) { return ExtVal{x.a + y.a, x.b + y.b}; }
ExtVal ext_sub(ExtVal x, ExtVal y) { return ExtVal{x.a - y.a, x.b - y.b}; }
ExtVal ext_scalar(ExtVal x, uint64_t k) { return ExtVal{x.a * k, x.b * k}; }
ExtVal ext_wires(Recorder& R, int w) { return ExtVal{R.wire(w), R.wire(w + 1)}; }

Val sbox(Val x) {  // core/src/poseidon.rs:546-552
    Val x2 = x * x, x4 = x2 * x2;
    return x * x2 * x4;
}

void mds(Val (&s)[12]) {  // core/src/poseidon.rs:178-198
    Val out[12];
    for (int r = 0; r < 12; r++) {
        Val acc = s[r] * (POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[r]);
        for (int i = 1; i < 12; i++) acc = acc + s[(i + r) % 12] * POSEIDON_MDS_CIRC[i];
        out[r] = acc;
    }
    for (int r = 0; r < 12; r++) s[r] = out[r];
}

// PoseidonGate::eval_unfiltered_base_one, plonky2/src/gates/poseidon.rs:204-283 (wire layout :43-100)
void eval_poseidon(Recorder& R, std::vector<Val>& c) {
    const int SWAP = 24, DELTA = 25, FULL0 = 29, PARTIAL = 65, FULL1 = 87;
    Val swap = R.wire(SWAP);
    c.push_back(swap * (swap - 1));
    Val st[12];
    for (int i = 0; i < 4; i++) {
        Val lhs = R.wire(i), rhs = R.wire(i + 4), delta = R.wire(DELTA + i);
        c.push_back(swap * (rhs - lhs) - delta);
        st[i] = lhs + delta;
        st[i + 4] = rhs - delta;
    }
    for (int i = 8; i < 12; i++) st[i] = R.wire(i);
    unsigned round = 0;
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = st[i] + POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i];
        if (r != 0)
            for (int i = 0; i < 12; i++) {
                Val in = R.wire(FULL0 + 12 * (r - 1) + i);
                c.push_back(st[i] - in);
                st[i] = in;
            }
        for (int i = 0; i < 12; i++) st[i] = sbox(st[i]);
        mds(st);
        round++;
    }
    // partial rounds in the fast form, core/src/poseidon.rs:302-342,378-408
    for (int i = 0; i < 12; i++) st[i] = st[i] + POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i];
    {
        Val t[12];
        t[0] = st[0];
        for (int cc = 1; cc < 12; cc++) {
            Val acc = st[1] * POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[cc - 1];
            for (int r = 2; r < 12; r++) acc = acc + st[r] * POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (cc - 1)];
            t[cc] = acc;
        }
        for (int i = 0; i < 12; i++) st[i] = t[i];
    }
    const uint64_t m00 = POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0];
    for (int r = 0; r < 22; r++) {
        Val in = R.wire(PARTIAL + r);
        c.push_back(st[0] - in);
        Val s0 = sbox(in);
        if (r != 21) s0 = s0 + POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[r];
        Val d = s0 * m00;
        for (int j = 1; j < 12; j++) d = d + st[j] * POSEIDON_FAST_PARTIAL_ROUND_W_HATS[11 * r + j - 1];
        for (int j = 1; j < 12; j++) st[j] = st[j] + s0 * POSEIDON_FAST_PARTIAL_ROUND_VS[11 * r + j - 1];
        st[0] = d;
    }
    round += 22;
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = st[i] + POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i];
        for (int i = 0; i < 12; i++) {
            Val in = R.wire(FULL1 + 12 * r + i);
            c.push_back(st[i] - in);
            st[i] = in;
        }
        for (int i = 0; i < 12; i++) st[i] = sbox(st[i]);
        mds(st);
        round++;
    }
    for (int i = 0; i < 12; i++) c.push_back(st[i] - R.wire(12 + i));
}

// Gate::eval_unfiltered for the supported gates; `prefix` = selector columns removed (gate.rs:179)
void eval_gate(Recorder& R, const GateInfo& gi, unsigned prefix, std::vector<Val>& c) {
    const GateInfo& g = gi;
    switch (g.kind) {
        case QP_GATE_CONSTANT:  // constant.rs:121-129
            for (unsigned i = 0; i < g.param; i++) c.push_back(R.constant(prefix + i) - R.wire(i));
            break;
        case QP_GATE_PUBLIC_INPUT:  // public_input.rs:103-113
            for (int i = 0; i < 4; i++) c.push_back(R.wire(i) - R.pih(i));
            break;
        case QP_GATE_ARITHMETIC: {  // arithmetic_base.rs:168-185
            Val c0 = R.constant(prefix), c1 = R.constant(prefix + 1);
            for (unsigned i = 0; i < g.param; i++) {
                Val m0 = R.wire(4 * i), m1 = R.wire(4 * i + 1), ad = R.wire(4 * i + 2), out = R.wire(4 * i + 3);
                c.push_back(out - (m0 * m1 * c0 + ad * c1));
            }
            break;
        }
        case QP_GATE_POSEIDON: eval_poseidon(R, c); break;
        case QP_GATE_ARITHMETIC_EXT: {  // arithmetic_extension.rs:92-110, D = 2
            Val c0 = R.constant(prefix), c1 = R.constant(prefix + 1);
            for (unsigned i = 0; i < g.param; i++) {
                ExtVal m0{R.wire(8 * i), R.wire(8 * i + 1)}, m1{R.wire(8 * i + 2), R.wire(8 * i + 3)};
                ExtVal ad{R.wire(8 * i + 4), R.wire(8 * i + 5)}, out{R.wire(8 * i + 6), R.wire(8 * i + 7)};
                ExtVal pr = ext_mul(m0, m1);
                c.push_back(out.a - (pr.a * c0 + ad.a * c1));
                c.push_back(out.b - (pr.b * c0 + ad.b * c1));
            }
            break;
        }
        case QP_GATE_MUL_EXT: {  // multiplication_extension.rs:86-101
            Val c0 = R.constant(prefix);
            for (unsigned i = 0; i < g.param; i++) {
                ExtVal m0{R.wire(6 * i), R.wire(6 * i + 1)}, m1{R.wire(6 * i + 2), R.wire(6 * i + 3)};
                ExtVal out{R.wire(6 * i + 4), R.wire(6 * i + 5)};
                ExtVal pr = ext_mul(m0, m1);
                c.push_back(out.a - pr.a * c0);
                c.push_back(out.b - pr.b * c0);
            }
            break;
        }
        case QP_GATE_BASE_SUM_2: {  // base_sum.rs:153-170 with B = 2
            // reduce_with_powers(limbs, 2): Horner from the last limb
            Val acc = R.wire((int)g.param);
            for (int i = (int)g.param - 1; i >= 1; i--) acc = acc * 2 + R.wire(i);
            c.push_back(acc - R.wire(0));
            for (unsigned i = 1; i <= g.param; i++) {
                Val limb = R.wire((int)i);
                c.push_back(limb * (limb - 1));  // (limb - 0)(limb - 1)
            }
            break;
        }
        case QP_GATE_RANDOM_ACCESS: {  // random_access.rs:144-189 (packed form :307-352)
            RandomAccessShape g(gi.param);
            for (unsigned copy = 0; copy < g.copies; copy++) {
                Val access_index = R.wire(g.access_index(copy));
                std::vector<Val> items, bits;
                for (unsigned i = 0; i < g.vec; i++) items.push_back(R.wire(g.item(i, copy)));
                Val claimed = R.wire(g.claimed(copy));
                for (unsigned i = 0; i < g.bits; i++) bits.push_back(R.wire(g.bit(i, copy)));
                for (Val b : bits) c.push_back(b * (b - 1));
                Val acc = bits.back();  // fold from ZERO: 0.double() + b = b
                for (int i = (int)g.bits - 2; i >= 0; i--) acc = acc * 2 + bits[i];
                c.push_back(acc - access_index);
                for (Val b : bits) {
                    std::vector<Val> next;
                    for (size_t k = 0; k + 1 < items.size(); k += 2) next.push_back(items[k] + b * (items[k + 1] - items[k]));
                    items = next;
                }
                c.push_back(items[0] - claimed);
            }
            for (unsigned i = 0; i < g.extra; i++) c.push_back(R.constant(prefix + i) - R.wire(g.extra_constant(i)));
            break;
        }
        case QP_GATE_REDUCING:        // reducing.rs:109-133
        case QP_GATE_REDUCING_EXT: {  // reducing_extension.rs:113-132
            const bool ext = gi.kind == QP_GATE_REDUCING_EXT;
            const unsigned n = gi.param, start_accs = 6 + (ext ? 2 * n : n);
            ExtVal alpha = ext_wires(R, 2), acc = ext_wires(R, 4);
            for (unsigned i = 0; i < n; i++) {
                ExtVal nxt = ext_wires(R, i == n - 1 ? 0 : (int)(start_accs + 2 * i));
                ExtVal t = ext_mul(acc, alpha);
                if (ext) {
                    ExtVal co = ext_wires(R, (int)(6 + 2 * i));
                    c.push_back(t.a + co.a - nxt.a);
                    c.push_back(t.b + co.b - nxt.b);
                } else {
                    c.push_back(t.a + R.wire((int)(6 + i)) - nxt.a);
                    c.push_back(t.b - nxt.b);
                }
                acc = nxt;
            }
            break;
        }
        case QP_GATE_POSEIDON_MDS: {  // poseidon_mds.rs:150-169, mds_row_shf_field core/src/poseidon.rs:200-215
            ExtVal in[12];
            for (int i = 0; i < 12; i++) in[i] = ext_wires(R, 2 * i);
            for (int r = 0; r < 12; r++) {
                ExtVal acc = ext_scalar(in[r], POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[r]);
                for (int i = 1; i < 12; i++) acc = ext_add(acc, ext_scalar(in[(i + r) % 12], POSEIDON_MDS_CIRC[i]));
                c.push_back(R.wire(24 + 2 * r) - acc.a);
                c.push_back(R.wire(25 + 2 * r) - acc.b);
            }
            break;
        }
        case QP_GATE_EXPONENTIATION: {  // exponentiation.rs:210-245
            const int n = (int)gi.param;
            Val base = R.wire(0), output = R.wire(1 + n);
            for (int i = 0; i < n; i++) {
                Val cur_bit = R.wire(1 + (n - i - 1));
                Val mul_by = cur_bit * base + (R.imm(1) - cur_bit);
                Val computed = mul_by;
                if (i != 0) {
                    Val prev = R.wire(2 + n + i - 1);
                    computed = prev * prev * mul_by;
                }
                c.push_back(computed - R.wire(2 + n + i));
            }
            c.push_back(output - R.wire(2 + n + n - 1));
            break;
        }
        case QP_GATE_COSET_INTERPOLATION: {  // coset_interpolation.rs:260-307, partial_interpolate :572-599
            CosetInterpolationShape g(gi.param);
            const std::vector<uint64_t> domain = two_adic_subgroup(g.bits), weights = barycentric_weights(domain);
            Val shift = R.wire(0), shift_inv = R.wire(g.shift_inverse());
            ExtVal point = ext_wires(R, g.point()), x = ext_wires(R, g.shifted_point());
            c.push_back(shift * shift_inv - 1);
            c.push_back(point.a - x.a * shift);
            c.push_back(point.b - x.b * shift);
            std::vector<ExtVal> values;
            for (unsigned i = 0; i < g.points; i++) values.push_back(ext_wires(R, g.value(i)));
            auto partial = [&](unsigned lo, unsigned hi, ExtVal& ev, ExtVal& prod) {
                for (unsigned j = lo; j < hi; j++) {
                    ExtVal val = ext_scalar(values[j], weights[j]);
                    ExtVal term{x.a - domain[j], x.b};
                    ev = ext_add(ext_mul(ev, term), ext_mul(val, prod));
                    prod = ext_mul(prod, term);
                }
            };
            // initial_eval = 0, initial_partial_prod = 1: the first step is (w_0 v_0, x - x_0)
            ExtVal ev = ext_scalar(values[0], weights[0]), prod{x.a - domain[0], x.b};
            partial(1, g.degree, ev, prod);
            for (unsigned i = 0; i < g.inter; i++) {
                ExtVal iev = ext_wires(R, g.inter_eval(i)), iprod = ext_wires(R, g.inter_prod(i));
                c.push_back(iev.a - ev.a);
                c.push_back(iev.b - ev.b);
                c.push_back(iprod.a - prod.a);
                c.push_back(iprod.b - prod.b);
                const unsigned start = 1 + (g.degree - 1) * (i + 1), end = std::min(start + g.degree - 1, g.points);
                ev = iev;
                prod = iprod;
                partial(start, end, ev, prod);
            }
            ExtVal out = ext_wires(R, g.eval_value());
            c.push_back(out.a - ev.a);
            c.push_back(out.b - ev.b);
            break;
        }
        default: break;
    }
}

}  // namespace

struct qp_program {
    std::vector<GateInfo> gates;  // sorted
    std::vector<uint32_t> order;  // sorted position -> index in the caller's list
    std::vector<uint32_t> selector_indices;
    std::vector<uint32_t> groups;  // [start0, end0, start1, end1, ...]
    std::vector<uint64_t> code, pool;
    std::vector<uint32_t> segments;  // word offset of every self-contained segment of `code`
    uint32_t n_regs = 1;
    uint32_t num_gate_constants = 0, num_gate_constraints = 0;
};

// ---- scheduling, segmentation and register allocation -----------------------------------------
// The device evaluates the program once per point of the quotient domain with its registers in
// shared memory, so what matters is (1) few registers -- they bound the points resident per SM --
// and (2) independent pieces that different thread blocks can take for small circuits.
//   * values are scheduled lazily in post-order from each constraint (they die young);
//   * a loaded column whose next use is far away is dropped and loaded again (`REMAT_GAP`): a
//     BaseSumGate reads its 63 limbs twice, 250 steps apart, and would otherwise pin 65 registers;
//   * the action list is cut at gate boundaries -- and inside a gate that is much larger than the
//     rest (PoseidonGate) between constraints -- into segments of similar cost.  Every segment is
//     self-contained (it recomputes what it shares with another), ends with OP_END, and the
//     result is the sum over segments because sum_gates filter * sum_k alpha^k c_k is additive.
namespace {

constexpr int REMAT_GAP = 64;          // steps between two uses of a load beyond which it is reloaded
constexpr unsigned TARGET_SEGMENTS = 6;
constexpr size_t MIN_SEGMENT_STEPS = 384;

struct Step {
    unsigned op;   // node op, FMAI, or EMIT / GATE / WAIT
    int dst;       // value id defined (node steps), -1 for actions
    int a, b;      // value ids of the register operands (-1: none)
    int c;         // column (LDW / LDK), pih index (LDP), pool slot (LDI / MULI / ADDI / FMAI), constraint (EMIT)
};

// instruction word: op | dst << 8 | a << 16 | b << 24 | c << 32   (include/qp_plonky2_b200.h)
uint64_t word(unsigned op, unsigned dst, unsigned a, unsigned b, unsigned c) {
    return op | ((uint64_t)dst << 8) | ((uint64_t)a << 16) | ((uint64_t)b << 24) | ((uint64_t)c << 32);
}

bool is_leaf(unsigned op) { return op >= LDW && op <= LDI; }
bool is_unary(unsigned op) { return op == MULI || op == ADDI; }

// Schedules `acts` as one self-contained piece; appends the steps; returns how many were added
// after each action (cumulative), for the cost model.
struct Scheduler {
    const Recorder& R;
    std::vector<Step> steps;
    std::vector<int> val_of;      // node -> current value id (-1: not computed in this piece)
    std::vector<int> touched_at;  // node -> step index of its last use (leaves)
    std::vector<int> uses;        // node -> number of consumers inside this piece
    std::vector<uint32_t> weight; // node -> size of its expression tree (saturating): which operand goes first
    int n_vals = 0;
    Scheduler(const Recorder& r, const std::vector<Recorder::Action>& acts)
        : R(r), val_of(r.nodes.size(), -1), touched_at(r.nodes.size(), -1), uses(r.nodes.size(), 0),
          weight(r.nodes.size(), 1) {
        // operands were recorded before their consumers, so one pass in index order sizes every tree
        for (size_t i = 0; i < r.nodes.size(); i++) {
            const auto& nd = r.nodes[i];
            if (is_leaf(nd.op)) continue;
            uint64_t w = 1 + (uint64_t)weight[nd.a] + (is_unary(nd.op) ? 0 : weight[nd.b]);
            weight[i] = (uint32_t)std::min<uint64_t>(w, 1u << 30);
        }
        std::vector<char> seen(r.nodes.size(), 0);
        std::vector<int> stack;
        for (const auto& a : acts) {
            uses[a.node]++;
            stack.push_back(a.node);
        }
        while (!stack.empty()) {
            const int i = stack.back();
            stack.pop_back();
            if (seen[i]) continue;
            seen[i] = 1;
            const auto& nd = R.nodes[i];
            if (is_leaf(nd.op)) continue;
            uses[nd.a]++;
            stack.push_back(nd.a);
            if (!is_unary(nd.op)) {
                uses[nd.b]++;
                stack.push_back(nd.b);
            }
        }
    }
    // x * constant + y as ONE instruction when the product has no other consumer (the MDS layers
    // and the reducing gates are chains of these): which operand of ADD node nd is that product
    int fusable(const Recorder::Node& nd) const {
        if (nd.op != ADD) return -1;
        for (int m : {nd.a, nd.b})
            if (R.nodes[m].op == MULI && uses[m] == 1 && val_of[m] < 0 && nd.a != nd.b) return m;
        return -1;
    }

    int use_leaf(int i) {  // value id of a loaded column / constant, (re)loading it if it went stale
        const auto& nd = R.nodes[i];
        if (val_of[i] < 0 || (int)steps.size() - touched_at[i] > REMAT_GAP) {
            val_of[i] = n_vals++;
            steps.push_back({nd.op, val_of[i], -1, -1, is_leaf(nd.op) ? nd.a : 0});
        }
        touched_at[i] = (int)steps.size();
        return val_of[i];
    }
    int operand(int i) { return is_leaf(R.nodes[i].op) ? use_leaf(i) : val_of[i]; }
    void compute(int root) {
        if (is_leaf(R.nodes[root].op) || val_of[root] >= 0) return;
        std::vector<std::pair<int, bool>> stack{{root, false}};
        while (!stack.empty()) {
            auto [i, done] = stack.back();
            stack.pop_back();
            const auto& nd = R.nodes[i];
            if (val_of[i] >= 0) continue;
            const int prod = fusable(nd);
            auto want = [&](int k) {
                if (!is_leaf(R.nodes[k].op) && val_of[k] < 0) stack.push_back({k, false});
            };
            if (!done) {
                // the operand with the larger expression goes first (it is popped last-pushed-first):
                // its intermediate values are gone by the time the smaller one is computed, and a long
                // accumulation chain pulls in its small side terms just in time instead of all up front
                stack.push_back({i, true});
                int first = nd.a, second = is_unary(nd.op) ? -1 : nd.b;
                if (prod >= 0) {
                    first = R.nodes[prod].a;
                    second = prod == nd.a ? nd.b : nd.a;
                }
                if (second >= 0 && weight[second] > weight[first]) std::swap(first, second);
                if (second >= 0) want(second);
                want(first);
                continue;
            }
            if (prod >= 0) {
                const int x = operand(R.nodes[prod].a), y = operand(prod == nd.a ? nd.b : nd.a);
                val_of[i] = n_vals++;
                steps.push_back({FMAI, val_of[i], x, y, R.nodes[prod].b});
                continue;
            }
            const int va = operand(nd.a);
            if (is_unary(nd.op)) {
                val_of[i] = n_vals++;
                steps.push_back({nd.op, val_of[i], va, -1, nd.b});
                continue;
            }
            const int vb = operand(nd.b);
            val_of[i] = n_vals++;
            steps.push_back({nd.op, val_of[i], va, vb, 0});
        }
    }
    void action(const Recorder::Action& act) {
        compute(act.node);
        steps.push_back({act.op, -1, operand(act.node), -1, (int)act.k});
    }
};

// Column loads (LDW / LDK) are asynchronous on the device (cp.async into the register file): the
// load that sat at the position of its first use is issued QP_PROGRAM_LOAD_LEAD loads earlier, and
// the device lets at most that many loads stay in flight after issuing one -- so a value has
// arrived when its consumer runs, and HBM latency overlaps the arithmetic in between.  The last
// loads of a piece have no successor to wait for them: an explicit WAIT does.
void hoist_loads(std::vector<Step>& steps) {
    constexpr size_t D = QP_PROGRAM_LOAD_LEAD;
    std::vector<size_t> pos;
    for (size_t t = 0; t < steps.size(); t++)
        if (steps[t].op == LDW || steps[t].op == LDK) pos.push_back(t);
    const size_t m = pos.size();
    if (!m) return;
    std::vector<Step> out;
    out.reserve(steps.size() + 1);
    for (size_t j = 0; j < std::min(D, m); j++) out.push_back(steps[pos[j]]);
    size_t j = 0;  // next original load position
    for (size_t t = 0; t < steps.size(); t++) {
        if (j < m && t == pos[j]) {
            if (j + D < m) out.push_back(steps[pos[j + D]]);
            else if (j + D == m || (m < D && j == 0)) out.push_back({WAIT, -1, -1, -1, 0});
            j++;
            continue;
        }
        out.push_back(steps[t]);
    }
    steps.swap(out);
}

// registers by linear scan over one piece; appends the encoded words (+ OP_END)
void encode(const Scheduler& S, std::vector<uint64_t>& code, uint32_t& n_regs) {
    std::vector<int> last_use(S.n_vals, -1);
    for (size_t t = 0; t < S.steps.size(); t++) {
        const Step& s = S.steps[t];
        if (s.a >= 0) last_use[s.a] = (int)t;
        if (s.b >= 0) last_use[s.b] = (int)t;
    }
    std::vector<int> reg_of(S.n_vals, -1), free_regs;
    uint32_t regs = 0;
    for (size_t t = 0; t < S.steps.size(); t++) {
        const Step& s = S.steps[t];
        const unsigned ra = s.a >= 0 ? (unsigned)reg_of[s.a] : 0, rb = s.b >= 0 ? (unsigned)reg_of[s.b] : 0;
        if (s.a >= 0 && last_use[s.a] == (int)t) free_regs.push_back(reg_of[s.a]);
        if (s.b >= 0 && s.b != s.a && last_use[s.b] == (int)t) free_regs.push_back(reg_of[s.b]);
        unsigned rd = 0;
        if (s.dst >= 0) {
            if (!free_regs.empty()) {
                rd = (unsigned)free_regs.back();
                free_regs.pop_back();
            } else {
                rd = regs++;
            }
            reg_of[s.dst] = (int)rd;
            if (last_use[s.dst] < 0) free_regs.push_back((int)rd);  // dead value
        }
        code.push_back(word(s.op, rd, ra, rb, (unsigned)s.c));
    }
    code.push_back(END);
    if (regs > n_regs) n_regs = regs;
}

}  // namespace

static void compile(Recorder& R, qp_program* out) {
    // the gates: runs of EMITs closed by a GATE
    struct GateActs {
        size_t first, last;  // [first, last] in R.actions, last = the GATE action
        std::vector<size_t> cum;  // steps after each action when the gate is scheduled alone
    };
    std::vector<GateActs> gates;
    for (size_t i = 0, start = 0; i < R.actions.size(); i++)
        if (R.actions[i].op == GATE) {
            gates.push_back({start, i, {}});
            start = i + 1;
        }
    size_t total = 0;
    for (auto& g : gates) {
        Scheduler S(R, std::vector<Recorder::Action>(R.actions.begin() + g.first, R.actions.begin() + g.last + 1));
        for (size_t i = g.first; i <= g.last; i++) {
            S.action(R.actions[i]);
            g.cum.push_back(S.steps.size());
        }
        total += g.cum.back();
    }
    const size_t target = std::max(MIN_SEGMENT_STEPS, total / TARGET_SEGMENTS);
    // pieces = lists of actions; a gate much larger than the target is cut between constraints
    std::vector<std::vector<Recorder::Action>> pieces;
    std::vector<Recorder::Action> cur;
    size_t cur_cost = 0;
    auto flush = [&]() {
        if (!cur.empty()) pieces.push_back(cur);
        cur.clear();
        cur_cost = 0;
    };
    for (const auto& g : gates) {
        const size_t cost = g.cum.back(), n_emit = g.last - g.first;
        const size_t parts = std::min<size_t>(std::max<size_t>(1, (cost + target / 2) / target), std::max<size_t>(1, n_emit));
        if (parts > 1) {
            flush();
            size_t i = g.first;
            for (size_t part = 0; part < parts; part++) {
                const size_t until = part + 1 == parts ? cost : cost * (part + 1) / parts;
                std::vector<Recorder::Action> acts;
                while (i < g.last && (g.cum[i - g.first] <= until || acts.empty())) acts.push_back(R.actions[i++]);
                if (part + 1 == parts)
                    while (i < g.last) acts.push_back(R.actions[i++]);
                if (acts.empty()) continue;
                acts.push_back(R.actions[g.last]);  // every part closes with the gate's filter
                pieces.push_back(acts);
            }
            continue;
        }
        if (cur_cost && cur_cost + cost > target + target / 2) flush();
        for (size_t i = g.first; i <= g.last; i++) cur.push_back(R.actions[i]);
        cur_cost += cost;
        if (cur_cost >= target) flush();
    }
    flush();
    uint32_t n_regs = 0;
    for (const auto& acts : pieces) {
        Scheduler S(R, acts);
        for (const auto& a : acts) S.action(a);
        hoist_loads(S.steps);
        out->segments.push_back((uint32_t)out->code.size());
        encode(S, out->code, n_regs);
    }
    if (!out->code.empty()) out->code.pop_back();  // the library appends the final OP_END itself
    out->pool = R.pool;
    if (out->pool.empty()) out->pool.push_back(0);
    out->n_regs = n_regs ? n_regs : 1;
}

extern "C" void qp_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t st[25] = {0};
    const size_t rate = 136;
    std::vector<uint8_t> msg(data, data + len);
    msg.push_back(0x01);
    while (msg.size() % rate) msg.push_back(0);
    msg.back() |= 0x80;
    for (size_t off = 0; off < msg.size(); off += rate) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w = 0;
            for (int k = 0; k < 8; k++) w |= (uint64_t)msg[off + 8 * i + k] << (8 * k);
            st[i] ^= w;
        }
        keccak_f(st);
    }
    for (int i = 0; i < 32; i++) out[i] = (uint8_t)(st[i / 8] >> (8 * (i % 8)));
}

extern "C" int qp_program_create(const qp_gate_desc* gates, size_t n_gates, unsigned max_degree, qp_program** out) {
    return qp_program_create_lookups(gates, n_gates, max_degree, 0, nullptr, 0, out);
}

extern "C" int qp_program_create_lookups(const qp_gate_desc* gates, size_t n_gates, unsigned max_degree,
                                         unsigned num_routed_wires, const qp_lookup_table* luts, size_t n_luts,
                                         qp_program** out) {
    if (!gates || !n_gates || !out || (n_luts && (!luts || !num_routed_wires))) return QP_ERR_BAD_ARG;
    *out = nullptr;
    for (size_t t = 0; t < n_luts; t++)
        if (!luts[t].table || !luts[t].len) return QP_ERR_BAD_ARG;  // "Empty LUTs are not supported."
    // gates/selectors.rs:27-75: TransSre, TransLdc, InitSre, LastLdc, then one "ends" selector per table
    const unsigned num_lookup_selectors = n_luts ? 4 + (unsigned)n_luts : 0;
    auto* p = new qp_program();
    std::vector<std::pair<GateInfo, uint32_t>> gs;
    for (size_t i = 0; i < n_gates; i++) {
        GateInfo g = gate_info(gates[i].kind, gates[i].param);
        if ((gates[i].kind == QP_GATE_LOOKUP || gates[i].kind == QP_GATE_LOOKUP_TABLE) && gates[i].param < n_luts) {
            const qp_lookup_table& t = luts[gates[i].param];
            if (gates[i].kind == QP_GATE_LOOKUP)   // lookup.rs:72-77; degree 0, no constants, no constraints
                g = {gates[i].kind, gates[i].param, 0, 0, 0,
                     "LookupGate {num_slots: " + std::to_string(num_routed_wires / 2) + ", lut_hash: " + lut_hash_debug(t) + "}"};
            else                                    // lookup_table.rs:86-92
                g = {gates[i].kind, gates[i].param, 0, 0, 0,
                     "LookupTableGate {num_slots: " + std::to_string(num_routed_wires / 3) + ", lut_hash: " + lut_hash_debug(t) +
                         ", last_lut_row: " + std::to_string(t.last_lut_row) + "}"};
        }
        if (g.id.empty()) {
            delete p;
            return QP_ERR_BAD_ARG;
        }
        gs.push_back({g, (uint32_t)i});
    }
    // circuit_builder.rs:1177-1179
    std::stable_sort(gs.begin(), gs.end(), [](const auto& a, const auto& b) {
        return std::make_pair(a.first.degree, a.first.id) < std::make_pair(b.first.degree, b.first.id);
    });
    for (auto& g : gs) {
        p->gates.push_back(g.first);
        p->order.push_back(g.second);
        p->num_gate_constants = std::max(p->num_gate_constants, g.first.num_constants);
        p->num_gate_constraints = std::max(p->num_gate_constraints, g.first.num_constraints);
    }
    // selector_polynomials' grouping, selectors.rs:99-166
    const unsigned num_gates = (unsigned)n_gates, max_gate_degree = p->gates.back().degree;
    if (max_gate_degree + num_gates - 1 <= max_degree) {
        p->selector_indices.assign(num_gates, 0);
        p->groups = {0, num_gates};
    } else {
        if (max_gate_degree >= max_degree) {  // "... has too high degree. Consider increasing `quotient_degree_factor`."
            delete p;
            return QP_ERR_BAD_ARG;
        }
        unsigned start = 0;
        while (start < num_gates) {
            unsigned size = 0;
            while (start + size < num_gates && size + p->gates[start + size].degree < max_degree) size++;
            p->groups.push_back(start);
            p->groups.push_back(start + size);
            for (unsigned i = 0; i < size; i++) p->selector_indices.push_back((uint32_t)(p->groups.size() / 2 - 1));
            start += size;
        }
    }
    const unsigned num_selectors = (unsigned)p->groups.size() / 2;
    // evaluate_gate_constraints_base_batch as a program, vanishing_poly.rs:700-726
    // PoseidonGate -- half of a recursion circuit's constraint work -- is not compiled: the device has
    // a native evaluator for it (quotient::poseidon_gate_kernel); the program carries one word with
    // the gate's filter parameters in a segment of its own.  QP_NATIVE_POSEIDON=0 compiles it like
    // any other gate (the two must agree: tests/test_gpu_plonk.py).
    const char* env_native = getenv("QP_NATIVE_POSEIDON");
    const bool native_poseidon = !(env_native && env_native[0] == '0');
    std::vector<uint64_t> native_words;
    Recorder R;
    for (unsigned i = 0; i < num_gates; i++) {
        const unsigned sel = p->selector_indices[i];
        if (native_poseidon && p->gates[i].kind == QP_GATE_POSEIDON && native_words.empty() && num_gates < 256) {
            native_words.push_back(word(NATIVE_POSEIDON, p->groups[2 * sel + 1], i, p->groups[2 * sel],
                                        sel | ((num_selectors > 1 ? 1u : 0u) << 16)));
            continue;
        }
        Val s = R.constant((int)sel);
        Val filt = R.imm(1);
        for (unsigned j = p->groups[2 * sel]; j < p->groups[2 * sel + 1]; j++)  // compute_filter, gate.rs:326-333
            if (j != i) filt = filt * (R.imm(j) - s);
        if (num_selectors > 1) filt = filt * (R.imm(UNUSED_SELECTOR) - s);
        std::vector<Val> cons;
        eval_gate(R, p->gates[i], num_selectors + num_lookup_selectors /* gate.rs:179 */, cons);
        if (!cons.empty()) R.emit_gate(cons, filt);
        else R.memo.clear();
    }
    compile(R, p);
    for (uint64_t w : native_words) {  // a segment of its own (the library closes the last segment)
        if (!p->code.empty()) p->code.push_back(END);
        p->segments.push_back((uint32_t)p->code.size());
        p->code.push_back(w);
    }
    *out = p;
    return QP_OK;
}

// The compiler on somebody else's recording: a shim (or the Python twin) that ran the gates'
// eval_unfiltered over its own recording type hands over the node list and gets the same
// scheduling, segments and register allocation as the built-in gates.
extern "C" int qp_program_from_dag(const qp_dag_node* nodes, size_t n_nodes, const uint64_t* pool, size_t pool_len,
                                   const qp_dag_action* actions, size_t n_actions, qp_program** out) {
    if (!out || (n_nodes && !nodes) || (n_actions && !actions) || (pool_len && !pool)) return QP_ERR_BAD_ARG;
    *out = nullptr;
    Recorder R;
    R.pool.assign(pool, pool + pool_len);
    for (size_t i = 0; i < n_nodes; i++) {
        const qp_dag_node& nd = nodes[i];
        bool ok = false;
        switch (nd.op) {
            case LDW: case LDK: ok = true; break;
            case LDP: ok = nd.a < 4; break;
            case LDI: ok = nd.a < pool_len; break;
            case ADD: case SUB: case MUL: ok = nd.a < i && nd.b < i; break;   // operands before consumers
            case MULI: case ADDI: ok = nd.a < i && nd.b < pool_len; break;
            default: break;
        }
        if (!ok) return QP_ERR_BAD_ARG;
        R.nodes.push_back({nd.op, (int)nd.a, (int)nd.b});
    }
    bool open_gate = false;
    for (size_t i = 0; i < n_actions; i++) {
        const qp_dag_action& a = actions[i];
        if ((a.op != EMIT && a.op != GATE) || a.node >= n_nodes || a.k >= 65536) return QP_ERR_BAD_ARG;
        R.actions.push_back({a.op, (int)a.node, a.k});
        open_gate = a.op == EMIT;
    }
    if (open_gate) return QP_ERR_BAD_ARG;  // constraints without a closing GATE (filter)
    auto* p = new qp_program();
    compile(R, p);
    *out = p;
    return QP_OK;
}

extern "C" void qp_program_free(qp_program* p) { delete p; }
extern "C" size_t qp_program_code(const qp_program* p, const uint64_t** code) {
    if (code) *code = p->code.data();
    return p->code.size();
}
extern "C" size_t qp_program_pool(const qp_program* p, const uint64_t** pool) {
    if (pool) *pool = p->pool.data();
    return p->pool.size();
}
extern "C" unsigned qp_program_regs(const qp_program* p) { return p->n_regs; }
extern "C" size_t qp_program_segments(const qp_program* p, const uint32_t** offsets) {
    if (offsets) *offsets = p->segments.data();
    return p->segments.size();
}
extern "C" unsigned qp_program_num_selectors(const qp_program* p) { return (unsigned)p->groups.size() / 2; }
extern "C" unsigned qp_program_num_gate_constants(const qp_program* p) { return p->num_gate_constants; }
extern "C" unsigned qp_program_num_gate_constraints(const qp_program* p) { return p->num_gate_constraints; }
extern "C" int qp_program_gate(const qp_program* p, unsigned sorted_index, unsigned* original_index,
                               unsigned* selector_index, unsigned* group_start, unsigned* group_end) {
    if (!p || sorted_index >= p->gates.size()) return QP_ERR_BAD_ARG;
    const unsigned sel = p->selector_indices[sorted_index];
    if (original_index) *original_index = p->order[sorted_index];
    if (selector_index) *selector_index = sel;
    if (group_start) *group_start = p->groups[2 * sel];
    if (group_end) *group_end = p->groups[2 * sel + 1];
    return QP_OK;
}
