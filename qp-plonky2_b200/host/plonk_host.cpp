// Host side of the quotient path (include/qp_plonky2_host.h): the gate set of a circuit compiled
// into the straight-line constraint program that quotient::quotient_kernel interprets.
//
// Reference semantics (reference-relative citations):
//   gate order                      plonky2/src/plonk/circuit_builder.rs:1177-1179  (degree, then id)
//   selector groups                 plonky2/src/gates/selectors.rs:99-166
//   compute_filter                  plonky2/src/gates/gate.rs:326-333
//   evaluate_gate_constraints       plonky2/src/plonk/vanishing_poly.rs:700-726
//   NoopGate / ConstantGate / PublicInputGate / ArithmeticGate / PoseidonGate / ArithmeticExtensionGate /
//   MulExtensionGate / BaseSumGate<2>
//                                   plonky2/src/gates/{noop,constant,public_input,arithmetic_base,poseidon,
//                                   arithmetic_extension,multiplication_extension,base_sum}.rs
// In a Rust build this role is played by the shim's recording field type run over
// Gate::eval_unfiltered_base_one; here the same recording evaluation is written in
// C++ for the gates above.  Pure host logic: no field arithmetic on data happens here.
#include <algorithm>
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

enum Op : unsigned { END = 0, LDW, LDK, LDP, LDI, ADD, SUB, MUL, EMIT, GATE, MULI, ADDI };

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

GateInfo gate_info(uint32_t kind, uint32_t param) {
    switch (kind) {
        case QP_GATE_NOOP: return {kind, param, 0, 0, 0, "NoopGate"};
        case QP_GATE_CONSTANT:
            return {kind, param, 1, param, param, "ConstantGate { num_consts: " + std::to_string(param) + " }"};
        case QP_GATE_PUBLIC_INPUT: return {kind, param, 1, 0, 4, "PublicInputGate"};
        case QP_GATE_ARITHMETIC:
            return {kind, param, 3, 2, param, "ArithmeticGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_POSEIDON:
            return {kind, param, 7, 0, 123,
                    "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>"};
        case QP_GATE_ARITHMETIC_EXT:
            return {kind, param, 3, 2, 2 * param, "ArithmeticExtensionGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_MUL_EXT:
            return {kind, param, 3, 1, 2 * param, "MulExtensionGate { num_ops: " + std::to_string(param) + " }"};
        case QP_GATE_BASE_SUM_2:
            return {kind, param, 2, 0, 1 + param, "BaseSumGate { num_limbs: " + std::to_string(param) + " } + Base: 2"};
    }
    return {kind, param, 0, 0, 0, ""};
}

// F_p^2 = F_p[X]/(X^2 - 7) over recording values (field/src/extension/quadratic.rs:186-199)
struct ExtVal {
    Val a, b;
};
ExtVal ext_mul(ExtVal x, ExtVal y) { return ExtVal{x.a * y.a + (x.b * y.b) * 7, x.a * y.b + x.b * y.a}; }

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
void eval_gate(Recorder& R, const GateInfo& g, unsigned prefix, std::vector<Val>& c) {
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
    uint32_t n_regs = 1;
    uint32_t num_gate_constants = 0, num_gate_constraints = 0;
};

static void compile(Recorder& R, qp_program* out) {
    // lazy post-order schedule from each action's root, registers reused after the last use
    struct Step {
        bool is_node;
        int i;
        unsigned op, k;
    };
    std::vector<Step> sched;
    std::vector<char> seen(R.nodes.size(), 0);
    auto is_leaf = [](unsigned op) { return op >= LDW && op <= LDI; };
    auto is_unary = [](unsigned op) { return op == MULI || op == ADDI; };
    for (const auto& act : R.actions) {
        std::vector<std::pair<int, bool>> stack{{act.node, false}};
        while (!stack.empty()) {
            auto [i, done] = stack.back();
            stack.pop_back();
            if (seen[i]) continue;
            const auto& nd = R.nodes[i];
            if (done || is_leaf(nd.op)) {
                seen[i] = 1;
                sched.push_back({true, i, 0, 0});
                continue;
            }
            stack.push_back({i, true});
            if (!is_unary(nd.op) && !seen[nd.b]) stack.push_back({nd.b, false});
            if (!seen[nd.a]) stack.push_back({nd.a, false});
        }
        sched.push_back({false, act.node, act.op, act.k});
    }
    std::vector<int> last_use(R.nodes.size(), -1);
    for (size_t t = 0; t < sched.size(); t++) {
        const Step& s = sched[t];
        if (s.is_node) {
            const auto& nd = R.nodes[s.i];
            if (nd.op == ADD || nd.op == SUB || nd.op == MUL) last_use[nd.a] = last_use[nd.b] = (int)t;
            else if (is_unary(nd.op)) last_use[nd.a] = (int)t;
        } else {
            last_use[s.i] = (int)t;
        }
    }
    std::vector<int> reg_of(R.nodes.size(), -1), free_regs;
    uint32_t n_regs = 0;
    for (size_t t = 0; t < sched.size(); t++) {
        const Step& s = sched[t];
        if (s.is_node) {
            const auto& nd = R.nodes[s.i];
            uint64_t ra = 0, rb = 0;
            if (nd.op == ADD || nd.op == SUB || nd.op == MUL) {
                ra = reg_of[nd.a];
                rb = reg_of[nd.b];
                if (last_use[nd.a] == (int)t) free_regs.push_back(reg_of[nd.a]);
                if (nd.b != nd.a && last_use[nd.b] == (int)t) free_regs.push_back(reg_of[nd.b]);
            } else if (is_unary(nd.op)) {
                ra = reg_of[nd.a];
                rb = nd.b;
                if (last_use[nd.a] == (int)t) free_regs.push_back(reg_of[nd.a]);
            } else {
                ra = nd.a;
            }
            int r;
            if (!free_regs.empty()) {
                r = free_regs.back();
                free_regs.pop_back();
            } else {
                r = (int)n_regs++;
            }
            reg_of[s.i] = r;
            out->code.push_back(nd.op | ((uint64_t)r << 8) | (ra << 24) | (rb << 40));
            if (last_use[s.i] < 0) free_regs.push_back(r);
        } else {
            out->code.push_back(s.op | ((uint64_t)reg_of[s.i] << 24) | ((uint64_t)s.k << 40));
            if (last_use[s.i] == (int)t) free_regs.push_back(reg_of[s.i]);
        }
    }
    out->pool = R.pool;
    if (out->pool.empty()) out->pool.push_back(0);
    out->n_regs = n_regs ? n_regs : 1;
}

extern "C" int qp_program_create(const qp_gate_desc* gates, size_t n_gates, unsigned max_degree, qp_program** out) {
    if (!gates || !n_gates || !out) return QP_ERR_BAD_ARG;
    *out = nullptr;
    auto* p = new qp_program();
    std::vector<std::pair<GateInfo, uint32_t>> gs;
    for (size_t i = 0; i < n_gates; i++) {
        GateInfo g = gate_info(gates[i].kind, gates[i].param);
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
    Recorder R;
    for (unsigned i = 0; i < num_gates; i++) {
        const unsigned sel = p->selector_indices[i];
        Val s = R.constant((int)sel);
        Val filt = R.imm(1);
        for (unsigned j = p->groups[2 * sel]; j < p->groups[2 * sel + 1]; j++)  // compute_filter, gate.rs:326-333
            if (j != i) filt = filt * (R.imm(j) - s);
        if (num_selectors > 1) filt = filt * (R.imm(UNUSED_SELECTOR) - s);
        std::vector<Val> cons;
        eval_gate(R, p->gates[i], num_selectors /* + num_lookup_selectors = 0 */, cons);
        if (!cons.empty()) R.emit_gate(cons, filt);
        else R.memo.clear();
    }
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
