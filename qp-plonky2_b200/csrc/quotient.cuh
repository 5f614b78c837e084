// Plonk permutation argument and quotient evaluation on the device.
//
// Reference semantics:
//   wires_permutation_partial_products_and_zs   plonky2/src/plonk/prover.rs:402-480
//   compute_quotient_polys                      plonky2/src/plonk/prover.rs:640-866
//   eval_vanishing_poly_base_batch              plonky2/src/plonk/vanishing_poly.rs:166-330
//   check_partial_products                      plonky2/src/util/partial_products.rs:52-95
//   ZeroPolyOnCoset                             field/src/zero_poly_coset.rs
//   check_lookup_constraints_batch              plonky2/src/plonk/vanishing_poly.rs:521-680
//   compute_lookup_polys                        plonky2/src/plonk/prover.rs:489-636
//   reduce_with_powers_multi                    core/src/plonk_common.rs:68-85
//
// B200 formulation.  The three oracles keep their LDE column-major in leaf order (ntt.cuh), so
// the points of the quotient domain (every `step`-th point of the LDE) are the FIRST n << qdb
// positions of every column: one thread per position reads all columns coalesced, and the
// reference's per-batch gather/transposes (prover.rs:745-800) do not exist.  The gate set is
// data (common_data.gates): the host compiles every gate's filtered constraints into one
// straight-line program over F_p (plonk.py / the Rust shim's recording field type) and the
// kernel interprets it with its register file in shared memory -- exact arithmetic, so any
// evaluation order gives the reference's values.  Constraints carry their index (the power of
// alpha they get), so the program evaluates them in the gate's natural order and registers die
// young (a Poseidon gate needs ~40 live values, not the ~330 a last-to-first Horner order forces).
#pragma once
#include "goldilocks.cuh"
#include "poseidon.cuh"

namespace quotient {

// Constraint program: 64-bit words  op | dst << 8 | a << 16 | b << 24 | c << 32  (dst, a, b:
// registers; c: a column, a pool slot or a constraint index).
enum : unsigned {
    OP_END = 0,   // end of a segment
    OP_LDW = 1,   // dst = wire column c              (asynchronous, see LOAD_LEAD)
    OP_LDK = 2,   // dst = constants_sigmas column c  (asynchronous)
    OP_LDP = 3,   // dst = public_inputs_hash[c]
    OP_LDI = 4,   // dst = pool[c]
    OP_ADD = 5,
    OP_SUB = 6,
    OP_MUL = 7,
    OP_EMIT = 8,  // constraint c of the current gate: h += alpha^c r[a]
    OP_GATE = 9,  // end of a gate: G += r[a] (filter) * h, h = 0
    OP_MULI = 10, // dst = r[a] * pool[c]
    OP_ADDI = 11, // dst = r[a] + pool[c]
    OP_WAIT = 12, // every column load issued so far has arrived
    OP_FMAI = 13, // dst = r[a] * pool[c] + r[b]
    // A segment that is this single word hands a whole PoseidonGate to poseidon_gate_kernel instead
    // of the interpreter: dst = group end, a = the gate's index (= its selector value), b = group
    // start, c = selector column | (num_selectors > 1) << 16.
    OP_NATIVE_POSEIDON = 14,
};

// LDW / LDK are asynchronous copies global -> register file (cp.async, 8 bytes per thread): issuing
// one leaves at most LOAD_LEAD loads in flight, so the compiler (plonk_host.cpp, hoist_loads) issues
// each load LOAD_LEAD loads ahead of its first use and HBM latency overlaps the arithmetic between.
constexpr int LOAD_LEAD = 3;

// the register file and the tables live in shared memory and are addressed with 32-bit shared
// addresses computed once per thread (one LEA per operand in the interpreter loop)
__device__ __forceinline__ uint64_t lds64(unsigned addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void load_async(unsigned dst_shared, const uint64_t* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n"
                 "cp.async.commit_group;\n"
                 "cp.async.wait_group %2;\n" ::"r"(dst_shared), "l"(src), "n"(LOAD_LEAD)
                 : "memory");
}
__device__ __forceinline__ void load_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

constexpr int MAX_CHALLENGES = 4;
constexpr int BLOCK = 128;

__device__ __forceinline__ uint64_t inverse(uint64_t a) { return gl::pow(a, gl::P - 2); }

__device__ __forceinline__ size_t brev(size_t x, unsigned bits) {
    return bits ? (size_t)(__brevll((unsigned long long)x) >> (64 - bits)) : 0;
}

struct Params {
    // domains
    unsigned degree_bits, qdb, lg_lde;     // lg_lde = degree_bits + qdb
    unsigned nc, nr, np, max_degree, num_constants;
    // oracles: column-major LDE in leaf order, column stride = their local LDE length
    const uint64_t* cs;
    size_t cs_stride;
    const uint64_t* wires;
    size_t wires_stride;
    const uint64_t* zs;
    size_t zs_stride;
    // tables
    const uint64_t* tw_row;      // tw_row[i] = w_{2^lg_lde}^i
    const uint64_t* k_is;        // [nr]
    const uint64_t* zh_eval;     // [2^qdb]  g^n v^k - 1
    const uint64_t* zh_inv;      // [2^qdb]
    const uint64_t* alpha_pows;  // [nc][apow_stride]: alpha^t, t <= max(base, constraints per gate); base = nc + nc (np + 1)
    unsigned apow_stride;
    uint64_t betas[MAX_CHALLENGES], gammas[MAX_CHALLENGES], alphas[MAX_CHALLENGES];
    uint64_t pih[4];
    // gate program: n_seg self-contained segments (each ends with OP_END) at seg_off[k]
    const uint64_t* program;
    const uint32_t* seg_off;
    unsigned n_seg;
    const uint64_t* pool;
    unsigned pool_len;
    unsigned n_regs;
    // partial_out == 0: out = the quotient values [nc][2^lg_lde], natural order.
    // partial_out != 0: out = partial sums [gridDim.y (+ 1)][nc][2^lg_lde] for combine_kernel.
    uint64_t* out;
    unsigned partial_out;
    // A coset shard (multi-GPU): the kernel covers positions [pos_first, pos_first + pos_count) of the quotient
    // domain in leaf order; the columns it reads are the shard's (position p lives at p - pos_first), and with
    // out_leaf_order the outputs stay in leaf order, local: out[..][a][p - pos_first], rows of 2^out_lg words.
    // Whole domain: pos_first = 0, pos_count = 2^lg_lde, out_lg = lg_lde, out_leaf_order = 0 (natural order).
    size_t pos_first, pos_count;
    unsigned out_lg, out_leaf_order;
    // Lookup argument (n_luts == 0: the circuit has none).  The terms of challenge ch are vanishing terms
    // base + ch * lookup_terms .. (base = nc + nc (np + 1)); the gate constraints follow them, so gate
    // constraint k carries alpha^(gate_base + k), gate_base = base + nc * lookup_terms (vanishing_poly.rs:313-318).
    unsigned n_luts, nlp;              // tables; num_lookup_polys = 1 (RE) + the partial SLDC polynomials
    unsigned num_selectors;            // the lookup selectors are constants columns num_selectors .. + 4 + n_luts
    unsigned num_lu_slots, num_lut_slots, lu_degree, lut_degree;
    unsigned lookup_terms;             // 4 + n_luts + 2 (nlp - 1) per challenge
    unsigned gate_base;
    const uint64_t* lookup_consts;     // [nc][4 + n_luts]: ChallengeA, ChallengeB, ChallengeAlpha, ChallengeDelta, then
                                       // get_lut_poly(t).eval(delta) of every table (prover.rs:687-716)
};

// shared memory: pool | register file [n_regs][BLOCK].  The register file bounds the points resident
// per SM, so the powers of alpha (2 KB, read by EMIT only) stay in global memory behind L1: with the
// recursion gate set that is the difference between four and five resident blocks.
__host__ __device__ inline size_t smem_words(unsigned pool_len, unsigned n_regs) {
    return (size_t)pool_len + (size_t)n_regs * BLOCK;
}

// One thread per point of the quotient domain, enumerated in leaf order.  The work of a point is
// 1 + n_seg independent units -- the permutation terms and the segments of the gate program --
// dealt round-robin over blockIdx.y, so that a 2^12-row circuit (256 tiles) still fills 148 SMs.
__global__ void __launch_bounds__(BLOCK, 6) quotient_kernel(Params p) {
    extern __shared__ uint64_t smem[];
    uint64_t* const sh_pool = smem;
    uint64_t* const regs = sh_pool + p.pool_len;  // [n_regs][BLOCK]
    const uint64_t* const __restrict__ sh_apow = p.alpha_pows;
    for (unsigned k = threadIdx.x; k < p.pool_len; k += BLOCK) sh_pool[k] = p.pool[k];
    __syncthreads();
    const size_t n_lde = (size_t)1 << p.lg_lde;
    const size_t pos_raw = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool live = pos_raw < p.pos_count;
    const size_t pos = live ? pos_raw : p.pos_count - 1;   // position inside the shard (what the columns are indexed by)
    const size_t i = brev(p.pos_first + pos, p.lg_lde);    // natural index: the point is g w^i
    const unsigned zi = (unsigned)(i & (((size_t)1 << p.qdb) - 1));
    uint64_t res[MAX_CHALLENGES], G[MAX_CHALLENGES], h[MAX_CHALLENGES];
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++) res[a] = G[a] = h[a] = 0;
    auto add_term = [&](unsigned t, uint64_t term) {
#pragma unroll
        for (int a = 0; a < MAX_CHALLENGES; a++)
            if (a < (int)p.nc) res[a] = gl::add(res[a], gl::mul(term, __ldg(&sh_apow[a * p.apow_stride + t])));
    };
    constexpr unsigned REG_SHIFT = 10;  // one register = BLOCK * 8 bytes
    static_assert(BLOCK * 8 == 1 << REG_SHIFT, "register stride");
    const unsigned r_base = (unsigned)__cvta_generic_to_shared(regs + threadIdx.x);
    const unsigned pool_base = (unsigned)__cvta_generic_to_shared(sh_pool);
    for (unsigned unit = blockIdx.y; unit < 1 + p.n_seg; unit += gridDim.y) {
        if (unit == 0) {
            // (the next row of the same coset: same high bits of the leaf index, so inside the shard)
            const size_t pos_next = brev((i + ((size_t)1 << p.qdb)) & (n_lde - 1), p.lg_lde) - p.pos_first;
            const uint64_t x = gl::mul(gl::GENERATOR, p.tw_row[i]);
            // L_0(x) = Z_H(x) / (n (x - 1)), zero_poly_coset.rs:93-96
            const uint64_t l_0 = gl::mul(p.zh_eval[zi], inverse(gl::mul((uint64_t)1 << p.degree_bits, gl::sub(x, 1))));
            // the L_0(x) (Z(x) - 1) terms, vanishing_poly.rs:268-272
            for (unsigned ch = 0; ch < p.nc; ch++) add_term(ch, gl::mul(l_0, gl::sub(p.zs[ch * p.zs_stride + pos], 1)));
            // partial product checks, vanishing_poly.rs:297-320
            uint64_t pn[MAX_CHALLENGES], pd[MAX_CHALLENGES], bx[MAX_CHALLENGES];
#pragma unroll
            for (int c = 0; c < MAX_CHALLENGES; c++) {
                pn[c] = pd[c] = 1;
                bx[c] = c < (int)p.nc ? gl::mul(p.betas[c], x) : 0;
            }
            unsigned w = 0, in_chunk = 0;
            for (unsigned j = 0; j < p.nr; j++) {
                const uint64_t wire = p.wires[j * p.wires_stride + pos];
                const uint64_t sigma = p.cs[(p.num_constants + j) * p.cs_stride + pos];
                const uint64_t k = p.k_is[j];
#pragma unroll
                for (int c = 0; c < MAX_CHALLENGES; c++) {
                    if (c < (int)p.nc) {
                        const uint64_t wg = gl::add(wire, p.gammas[c]);
                        pn[c] = gl::mul(pn[c], gl::add(wg, gl::mul(bx[c], k)));
                        pd[c] = gl::mul(pd[c], gl::add(wg, gl::mul(p.betas[c], sigma)));
                    }
                }
                if (++in_chunk == p.max_degree || j + 1 == p.nr) {
#pragma unroll
                    for (int c = 0; c < MAX_CHALLENGES; c++) {
                        if (c < (int)p.nc) {
                            // accumulators Z(x), p_0 .. p_{np-1}, Z(g x): util/partial_products.rs:60-63
                            const uint64_t prev = w == 0 ? p.zs[c * p.zs_stride + pos]
                                                         : p.zs[(p.nc + c * p.np + w - 1) * p.zs_stride + pos];
                            const uint64_t next = w == p.np ? p.zs[c * p.zs_stride + pos_next]
                                                            : p.zs[(p.nc + c * p.np + w) * p.zs_stride + pos];
                            add_term(p.nc + c * (p.np + 1) + w, gl::sub(gl::mul(prev, pn[c]), gl::mul(next, pd[c])));
                            pn[c] = pd[c] = 1;
                        }
                    }
                    w++;
                    in_chunk = 0;
                }
            }
            continue;
        }
        // gate constraints (vanishing_poly.rs:700-726): G = sum over gates of filter * sum_k alpha^k c_k
        const uint64_t* pc = p.program + p.seg_off[unit - 1];
        uint64_t next = __ldg(pc);  // the instruction stream is fetched one word ahead of its use
        for (;;) {
            const unsigned lo = (unsigned)next, c = (unsigned)(next >> 32);
            const unsigned op = lo & 0xff;
            if (op == OP_END) break;  // (every segment ends with its loads consumed: nothing is in flight)
            next = __ldg(++pc);
            const unsigned rd = r_base + (((lo >> 8) & 0xff) << REG_SHIFT);
            const unsigned ra = r_base + (((lo >> 16) & 0xff) << REG_SHIFT);
            const unsigned rb = r_base + ((lo >> 24) << REG_SHIFT);
            // most frequent first (MDS layers and reducing gates are chains of FMAI)
            if (op == OP_FMAI) {
                sts64(rd, gl::add(gl::mul(lds64(ra), lds64(pool_base + c * 8)), lds64(rb)));
            } else if (op == OP_ADD) {
                sts64(rd, gl::add(lds64(ra), lds64(rb)));
            } else if (op == OP_MUL) {
                sts64(rd, gl::mul(lds64(ra), lds64(rb)));
            } else if (op == OP_LDW) {
                load_async(rd, &p.wires[c * p.wires_stride + pos]);
            } else if (op == OP_SUB) {
                sts64(rd, gl::sub(lds64(ra), lds64(rb)));
            } else if (op == OP_MULI) {
                sts64(rd, gl::mul(lds64(ra), lds64(pool_base + c * 8)));
            } else if (op == OP_EMIT) {
                const uint64_t v = lds64(ra);
#pragma unroll
                for (int k = 0; k < MAX_CHALLENGES; k++)
                    if (k < (int)p.nc) h[k] = gl::add(h[k], gl::mul(v, __ldg(&sh_apow[k * p.apow_stride + c])));
            } else if (op == OP_ADDI) {
                sts64(rd, gl::add(lds64(ra), lds64(pool_base + c * 8)));
            } else if (op == OP_LDI) {
                sts64(rd, lds64(pool_base + c * 8));
            } else if (op == OP_LDK) {
                load_async(rd, &p.cs[c * p.cs_stride + pos]);
            } else if (op == OP_WAIT) {
                load_wait_all();
            } else if (op == OP_GATE) {
                const uint64_t f = lds64(ra);
#pragma unroll
                for (int k = 0; k < MAX_CHALLENGES; k++)
                    if (k < (int)p.nc) {
                        G[k] = gl::add(G[k], gl::mul(f, h[k]));
                        h[k] = 0;
                    }
            } else if (op == OP_LDP) {
                sts64(rd, p.pih[c & 3]);
            }
        }
    }
    if (!live) return;
    const uint64_t zinv = p.zh_inv[zi];
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++)
        if (a < (int)p.nc) {
            const uint64_t total = gl::add(res[a], gl::mul(G[a], __ldg(&sh_apow[a * p.apow_stride + p.gate_base])));
            const size_t oi = p.out_leaf_order ? pos : i;
            if (!p.partial_out)
                p.out[((size_t)a << p.out_lg) + oi] = gl::canon(gl::mul(total, zinv));  // prover.rs:848-853
            else
                p.out[(((size_t)blockIdx.y * p.nc + a) << p.out_lg) + oi] = total;
        }
}

// ---- PoseidonGate, natively -----------------------------------------------------------------------
// A recursive verifier spends its rows on PoseidonGate, and its 123 degree-7 constraints are half of
// the interpreted program (3.6 k of 7.8 k operations at ~50 instructions each).  The gate is the
// permutation with every S-box input replaced by a wire (plonky2/src/gates/poseidon.rs:204-283), so
// it runs here on the permutation's own building blocks -- x^7 on the integer pipes, the MDS layer
// with the next round's constants on the FP64 pipe (poseidon::mds_layer_split) -- with the state in
// registers: one thread per point, no shared memory, ~27 k instructions instead of ~200 k.  The
// partial rounds are evaluated in the naive form (S-box on lane 0, full MDS; core/src/poseidon.rs:
// 613-633) where the reference's gate uses the sparse fast form: the two are the same function with
// the same lane-0 values (the S-box inputs -- the only partial-round values the gate constrains).
struct NativePoseidon {
    unsigned present, index, group_start, group_end, sel_column, many_selectors;
};

__global__ void __launch_bounds__(128) poseidon_gate_kernel(Params p, NativePoseidon g, uint64_t* __restrict__ out) {
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= p.pos_count) return;
    const size_t i = brev(p.pos_first + pos, p.lg_lde);
    auto wire = [&](unsigned c) { return p.wires[c * p.wires_stride + pos]; };
    uint64_t h[MAX_CHALLENGES];
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++) h[a] = 0;
    unsigned k = 0;  // constraint index, in the gate's order (poseidon.rs:204-283)
    auto emit = [&](uint64_t v) {
#pragma unroll
        for (int a = 0; a < MAX_CHALLENGES; a++)
            if (a < (int)p.nc) h[a] = gl::add(h[a], gl::mul(v, __ldg(&p.alpha_pows[a * p.apow_stride + k])));
        k++;
    };
    constexpr unsigned SWAP = 24, DELTA = 25, FULL0 = 29, PARTIAL = 65, FULL1 = 87;
    const uint64_t swap = wire(SWAP);
    emit(gl::mul(swap, gl::sub(swap, 1)));
    uint64_t s[12];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint64_t lhs = wire(j), rhs = wire(j + 4), delta = wire(DELTA + j);
        emit(gl::sub(gl::mul(swap, gl::sub(rhs, lhs)), delta));
        s[j] = gl::add(lhs, delta);
        s[j + 4] = gl::sub(rhs, delta);
    }
#pragma unroll
    for (int j = 8; j < 12; j++) s[j] = wire(j);
#pragma unroll
    for (int j = 0; j < 12; j++) s[j] = gl::add(s[j], poseidon::c_rc[j]);  // constant layer of round 0
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        if (r != 0) {
#pragma unroll
            for (int j = 0; j < 12; j++) {
                const uint64_t in = wire(FULL0 + 12 * (r - 1) + j);
                emit(gl::sub(s[j], in));
                s[j] = in;
            }
        }
        poseidon::sbox_all(s);
        poseidon::mds_layer_split(s, r + 1);  // + the constants of the next round
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        const uint64_t in = wire(PARTIAL + r);
        emit(gl::sub(s[0], in));
        s[0] = gl::pow7(in);
        poseidon::mds_layer_split(s, 4 + r + 1);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const uint64_t in = wire(FULL1 + 12 * r + j);
            emit(gl::sub(s[j], in));
            s[j] = in;
        }
        poseidon::sbox_all(s);
        poseidon::mds_layer_split(s, 26 + r + 1);  // row 30 of the constants is zero
    }
#pragma unroll
    for (int j = 0; j < 12; j++) emit(gl::sub(s[j], wire(12 + j)));
    // compute_filter, gates/gate.rs:326-333
    const uint64_t sel = p.cs[g.sel_column * p.cs_stride + pos];
    uint64_t f = 1;
    for (unsigned j = g.group_start; j < g.group_end; j++)
        if (j != g.index) f = gl::mul(f, gl::sub((uint64_t)j, sel));
    if (g.many_selectors) f = gl::mul(f, gl::sub(0xFFFFFFFFull, sel));
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++)
        if (a < (int)p.nc)
            out[((size_t)a << p.out_lg) + (p.out_leaf_order ? pos : i)] =
                gl::mul(gl::mul(f, h[a]), __ldg(&p.alpha_pows[a * p.apow_stride + p.gate_base]));
}

// ---- lookup argument ---------------------------------------------------------------------------------
// The lookup terms of the vanishing polynomial (check_lookup_constraints_batch, vanishing_poly.rs:521-680),
// one thread per point like poseidon_gate_kernel: a partial sum of its own for combine_kernel, so the
// interpreter kernel's registers stay what the gate program needs.
__global__ void __launch_bounds__(128) lookup_terms_kernel(Params p, uint64_t* __restrict__ out) {
    const size_t pos = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= p.pos_count) return;
    const size_t n_lde = (size_t)1 << p.lg_lde;
    const size_t i = brev(p.pos_first + pos, p.lg_lde);
    const unsigned base = p.nc + p.nc * (p.np + 1);
    uint64_t res[MAX_CHALLENGES];
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++) res[a] = 0;
    auto add_term = [&](unsigned t, uint64_t term) {
#pragma unroll
        for (int a = 0; a < MAX_CHALLENGES; a++)
            if (a < (int)p.nc) res[a] = gl::add(res[a], gl::mul(term, __ldg(&p.alpha_pows[a * p.apow_stride + t])));
    };
    // challenge by challenge (vanishing_poly.rs:263-283)
    const size_t pos_next = brev((i + ((size_t)1 << p.qdb)) & (n_lde - 1), p.lg_lde) - p.pos_first;
    const unsigned num_sldc = p.nlp - 1;
    const uint64_t* const sel = p.cs + (size_t)p.num_selectors * p.cs_stride + pos;  // column k: sel[k * cs_stride]
    const uint64_t s_trans_sre = sel[0], s_trans_ldc = sel[p.cs_stride], s_init_sre = sel[2 * p.cs_stride],
                   s_last_ldc = sel[3 * p.cs_stride];
    for (unsigned ch = 0; ch < p.nc; ch++) {
        const uint64_t* const d = p.lookup_consts + (size_t)ch * (4 + p.n_luts);
        const uint64_t da = d[0], db = d[1], dalpha = d[2], ddelta = d[3];
        const uint64_t* const z = p.zs + ((size_t)p.nc * (1 + p.np) + (size_t)ch * p.nlp) * p.zs_stride;
        const uint64_t z_re = z[pos], next_z_re = z[pos_next];
        unsigned t = base + ch * p.lookup_terms;
        add_term(t++, gl::mul(s_last_ldc, z[(size_t)num_sldc * p.zs_stride + pos]));   // last LDC
        add_term(t++, gl::mul(s_init_sre, z[p.zs_stride + pos]));                       // initial Sum
        add_term(t++, gl::mul(s_init_sre, z_re));                                       // initial RE
        for (unsigned r = 0; r < p.n_luts; r++)                                         // final RE of every table
            add_term(t++, gl::mul(sel[(size_t)(4 + r) * p.cs_stride], gl::sub(z_re, d[4 + r])));
        uint64_t cur = next_z_re;                                                       // RE row transition
        for (unsigned s = 0; s < p.num_lut_slots; s++)
            cur = gl::add(gl::mul(cur, ddelta), gl::add(p.wires[(size_t)(3 * s) * p.wires_stride + pos],
                                                        gl::mul(db, p.wires[(size_t)(3 * s + 1) * p.wires_stride + pos])));
        add_term(t++, gl::mul(s_trans_sre, gl::sub(z_re, cur)));
        for (unsigned poly = 0; poly < num_sldc; poly++) {
            // prod = prod_i f_i and sum = sum_i m_i prod_{j != i} f_j, f_i = alpha - combo_i, built slot by
            // slot: (prod, sum) <- (prod f, sum f + m prod)
            uint64_t lut_prod = 1, lut_sum = 0, lu_prod = 1, lu_sum = 0;
            const unsigned t1 = min((poly + 1) * p.lut_degree, p.num_lut_slots);
            for (unsigned s = poly * p.lut_degree; s < t1; s++) {
                const uint64_t* const w = p.wires + (size_t)(3 * s) * p.wires_stride + pos;
                const uint64_t f = gl::sub(dalpha, gl::add(w[0], gl::mul(da, w[p.wires_stride])));
                lut_sum = gl::add(gl::mul(lut_sum, f), gl::mul(w[2 * p.wires_stride], lut_prod));
                lut_prod = gl::mul(lut_prod, f);
            }
            const unsigned u1 = min((poly + 1) * p.lu_degree, p.num_lu_slots);
            for (unsigned s = poly * p.lu_degree; s < u1; s++) {
                const uint64_t* const w = p.wires + (size_t)(2 * s) * p.wires_stride + pos;
                const uint64_t f = gl::sub(dalpha, gl::add(w[0], gl::mul(da, w[p.wires_stride])));
                lu_sum = gl::add(gl::mul(lu_sum, f), lu_prod);
                lu_prod = gl::mul(lu_prod, f);
            }
            // the previous accumulator: the previous polynomial of this row, or the last one of the next row
            const uint64_t prev = poly == 0 ? z[(size_t)num_sldc * p.zs_stride + pos_next] : z[(size_t)poly * p.zs_stride + pos];
            const uint64_t diff = gl::sub(z[(size_t)(poly + 1) * p.zs_stride + pos], prev);
            add_term(t++, gl::mul(s_trans_sre, gl::sub(gl::mul(lut_prod, diff), lut_sum)));
            add_term(t++, gl::mul(s_trans_ldc, gl::add(gl::mul(lu_prod, diff), lu_sum)));
        }
    }
#pragma unroll
    for (int a = 0; a < MAX_CHALLENGES; a++)
        if (a < (int)p.nc) out[((size_t)a << p.out_lg) + (p.out_leaf_order ? pos : i)] = res[a];
}

// out[a][i] = Z_H(x_i)^-1 * sum_y partial[y][a][i]   (the units of quotient_kernel, summed)
// (rows of 2^out_lg words; leaf_order: the word at row offset o is position pos_first + o of the domain,
//  whose natural index is its bit reversal over lg_lde bits)
__global__ void combine_kernel(const uint64_t* __restrict__ partial, unsigned n_parts, unsigned nc, unsigned out_lg,
                               unsigned qdb, const uint64_t* __restrict__ zh_inv, uint64_t* __restrict__ out,
                               unsigned lg_lde, size_t pos_first, unsigned leaf_order) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t words = (size_t)nc << out_lg;
    if (id >= words) return;
    uint64_t acc = partial[id];
    for (unsigned y = 1; y < n_parts; y++) acc = gl::add(acc, partial[(size_t)y * words + id]);
    const size_t o = id & (((size_t)1 << out_lg) - 1);
    const size_t i = leaf_order ? brev(pos_first + o, lg_lde) : o;
    out[id] = gl::canon(gl::mul(acc, zh_inv[i & (((size_t)1 << qdb) - 1)]));
}

// coeffs[v][i] *= s^i with s^i = lo[i & mask] * hi[i >> split]   (coset_ifft's g^-i, polynomial/mod.rs:82-87)
__global__ void scale_powers_kernel(uint64_t* __restrict__ data, size_t n_vec, unsigned lg, const uint64_t* lo,
                                    const uint64_t* hi, int split) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (n_vec << lg)) return;
    const size_t i = id & (((size_t)1 << lg) - 1);
    const uint64_t f = gl::mul(lo[i & (((size_t)1 << split) - 1)], hi[i >> split]);
    data[id] = gl::canon(gl::mul(data[id], f));
}

// ---- Z and partial products ------------------------------------------------------------------
struct PermParams {
    unsigned degree_bits, nc, nr, np, max_degree;
    const uint64_t* wires;   // [num_wires][n] witness columns (values on H, natural order)
    const uint64_t* sigmas;  // [nr][n]
    const uint64_t* k_is;
    const uint64_t* tw_row;  // w_n^i
    uint64_t betas[MAX_CHALLENGES], gammas[MAX_CHALLENGES];
    uint64_t* chunk;         // [nc][np + 1][n]  quotient chunk products
    uint64_t* rowprod;       // [nc][n]          product of a row's chunks
};

constexpr int MAX_CHUNKS = 32;

// Thread per (row, challenge): chunk products of numerators / denominators, one batch inversion
// per row (Montgomery's trick; the reference inverts every denominator, prover.rs:449-455 -- the
// quotient of the products is the same field element).
__global__ void __launch_bounds__(128) perm_chunks_kernel(PermParams p) {
    const size_t n = (size_t)1 << p.degree_bits;
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n * p.nc) return;
    const unsigned ch = (unsigned)(id >> p.degree_bits);
    const size_t i = id & (n - 1);
    const uint64_t beta = p.betas[ch], gamma = p.gammas[ch];
    const uint64_t bx = gl::mul(beta, p.tw_row[i]);
    uint64_t nprod[MAX_CHUNKS], dprod[MAX_CHUNKS];
    const unsigned n_chunks = p.np + 1;
    unsigned j = 0;
    for (unsigned k = 0; k < n_chunks; k++) {
        uint64_t pn = 1, pd = 1;
        for (unsigned e = 0; e < p.max_degree && j < p.nr; e++, j++) {
            const uint64_t wg = gl::add(p.wires[j * n + i], gamma);
            pn = gl::mul(pn, gl::add(wg, gl::mul(bx, p.k_is[j])));
            pd = gl::mul(pd, gl::add(wg, gl::mul(beta, p.sigmas[j * n + i])));
        }
        nprod[k] = pn;
        dprod[k] = pd;
    }
    // batch inversion of dprod[0..n_chunks)
    uint64_t pref[MAX_CHUNKS];
    uint64_t acc = 1;
    for (unsigned k = 0; k < n_chunks; k++) {
        pref[k] = acc;
        acc = gl::mul(acc, dprod[k]);
    }
    uint64_t inv = inverse(acc);
    uint64_t row = 1;
    for (unsigned k = n_chunks; k-- > 0;) {
        const uint64_t dinv = gl::mul(inv, pref[k]);
        inv = gl::mul(inv, dprod[k]);
        nprod[k] = gl::mul(nprod[k], dinv);  // quotient chunk product
    }
    for (unsigned k = 0; k < n_chunks; k++) {
        p.chunk[((size_t)ch * n_chunks + k) * n + i] = nprod[k];
        row = gl::mul(row, nprod[k]);
    }
    p.rowprod[(size_t)ch * n + i] = row;
}

// Exclusive prefix PRODUCT over each of `n_vec` vectors of 2^lg elements, three phases.
constexpr int SCAN_BLOCK = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

// Tiles never straddle vectors: block b works on tile (b % tiles_per_vec) of vector (b / tiles_per_vec),
// tiles_per_vec = max(1, n / SCAN_TILE); elements past the vector's end count as 1.
// phase 1: per-tile products
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_products_kernel(const uint64_t* __restrict__ in, size_t n,
                                                                        size_t tiles_per_vec,
                                                                        uint64_t* __restrict__ tile_prod) {
    __shared__ uint64_t sh[SCAN_BLOCK];
    const size_t vec = blockIdx.x / tiles_per_vec, tile = blockIdx.x % tiles_per_vec;
    const size_t off = tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    const size_t base = vec * n + off;
    uint64_t acc = 1;
#pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++)
        if (off + e < n) acc = gl::mul(acc, in[base + e]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SCAN_BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = gl::mul(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_prod[blockIdx.x] = sh[0];
}

// phase 2: exclusive scan of the tile products of every vector (one block per vector, serial over
// tiles in chunks of SCAN_BLOCK)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles_kernel(uint64_t* tile_prod, size_t tiles_per_vec) {
    __shared__ uint64_t sh[SCAN_BLOCK];
    __shared__ uint64_t carry;
    uint64_t* v = tile_prod + (size_t)blockIdx.x * tiles_per_vec;
    if (threadIdx.x == 0) carry = 1;
    __syncthreads();
    for (size_t t0 = 0; t0 < tiles_per_vec; t0 += SCAN_BLOCK) {
        const size_t t = t0 + threadIdx.x;
        const uint64_t mine = t < tiles_per_vec ? v[t] : 1;
        sh[threadIdx.x] = mine;
        __syncthreads();
        // Hillis-Steele inclusive scan
        for (int s = 1; s < SCAN_BLOCK; s <<= 1) {
            uint64_t o = 1;
            if ((int)threadIdx.x >= s) o = sh[threadIdx.x - s];
            __syncthreads();
            if ((int)threadIdx.x >= s) sh[threadIdx.x] = gl::mul(sh[threadIdx.x], o);
            __syncthreads();
        }
        const uint64_t incl = sh[threadIdx.x];
        const uint64_t c = carry;
        const uint64_t excl = threadIdx.x == 0 ? c : gl::mul(c, sh[threadIdx.x - 1]);
        __syncthreads();
        if (t < tiles_per_vec) v[t] = excl;
        if (threadIdx.x == SCAN_BLOCK - 1) carry = gl::mul(c, incl);
        __syncthreads();
    }
}

// phase 3: Z(w^i) = exclusive prefix product; partial products p_k(w^i) = Z * chunk_0 ... chunk_k.
// out[(nc + nc np)][n]: Z_0 .. Z_{nc-1}, then the np partial products of challenge 0, 1, ...
// (the order they are committed in, prover.rs:255-261).
__global__ void __launch_bounds__(SCAN_BLOCK) perm_finish_kernel(PermParams p, size_t tiles_per_vec,
                                                                 const uint64_t* __restrict__ tile_excl,
                                                                 uint64_t* __restrict__ out) {
    __shared__ uint64_t sh[SCAN_BLOCK];
    const size_t n = (size_t)1 << p.degree_bits;
    const size_t vec = blockIdx.x / tiles_per_vec, tile = blockIdx.x % tiles_per_vec;
    const size_t off = tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    const size_t base = vec * n + off;
    uint64_t v[SCAN_ITEMS];
    uint64_t acc = 1;
#pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++) {
        v[e] = off + e < n ? p.rowprod[base + e] : 1;
        acc = gl::mul(acc, v[e]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 1; s < SCAN_BLOCK; s <<= 1) {
        uint64_t o = 1;
        if ((int)threadIdx.x >= s) o = sh[threadIdx.x - s];
        __syncthreads();
        if ((int)threadIdx.x >= s) sh[threadIdx.x] = gl::mul(sh[threadIdx.x], o);
        __syncthreads();
    }
    uint64_t z = tile_excl[blockIdx.x];
    if (threadIdx.x > 0) z = gl::mul(z, sh[threadIdx.x - 1]);
    const unsigned n_chunks = p.np + 1;
#pragma unroll
    for (int e = 0; e < SCAN_ITEMS; e++) {
        if (off + e < n) {
            const unsigned ch = (unsigned)vec;
            const size_t i = off + e;
            out[(size_t)ch * n + i] = gl::canon(z);
            uint64_t a = z;
            for (unsigned k = 0; k < p.np; k++) {
                a = gl::mul(a, p.chunk[((size_t)ch * n_chunks + k) * n + i]);
                out[((size_t)p.nc + (size_t)ch * p.np + k) * n + i] = gl::canon(a);
            }
            z = gl::mul(z, v[e]);
        }
    }
}

// ---- lookup polynomials (compute_lookup_polys, plonky2/src/plonk/prover.rs:489-636) ---------------------
// RE and the partial Sum / LDC polynomials are recurrences down the rows of a table's gates (LookupTableGate
// rows first_lut_row .. last_lut_row, then LookupGate rows last_lut_row - 1 .. last_lu_row): every row adds
// its own contribution to the value handed over by the row below it.  The contributions need the field
// inversions (one per slot) and are independent: phase 1 computes them with one thread per (challenge, row);
// phase 2 walks the chain, which is then additions and one multiplication per row.
struct LookupPolyParams {
    unsigned degree_bits, nc, np1;     // np1 = num_lookup_polys = 1 + num_partial_lookups
    unsigned num_lu_slots, num_lut_slots, lu_degree, lut_degree;
    const uint64_t* wires;             // [>= num_routed_wires][n] witness columns
    const uint32_t* rows;              // [n_rows]: row | (LookupTableGate row ? 1u << 31 : 0), all tables
    unsigned n_rows;
    const uint32_t* tables;            // [n_luts][3]: last_lu_row, last_lut_row, first_lut_row
    unsigned n_luts;
    const uint64_t* consts;            // [nc][4 + n_luts] (Params::lookup_consts)
    uint64_t* out;                     // [nc][np1][n], zero outside the tables' rows
};

// phase 1: out[ch][0][row] = sum_s combo_b(s) delta^(slots - 1 - s) (LookupTableGate rows: what the row adds to
// delta^slots RE(row + 1)); out[ch][k + 1][row] = the row's share of partial polynomial k: + sum m_s / (alpha -
// combo_a(s)) over the LookupTableGate slots of k, - sum 1 / (alpha - combo_a(s)) over the LookupGate slots of k.
// One block per (challenge, row), one thread per slot: the Fermat inversions -- ~100 dependent multiplications
// each -- are the whole cost and run side by side (one thread per row took 0.3 ms for the 11 lookup rows of a
// 2^14-row circuit, a tenth of a small proof; profiles/r02p_ncu_plonk_kernels.md); the block's last thread
// evaluates the RE contribution meanwhile.  blockDim.x > max(slots per row, partial polynomials).
__global__ void lookup_rows_kernel(LookupPolyParams p) {
    extern __shared__ uint64_t term[];   // [blockDim.x]
    const unsigned ch = blockIdx.x / p.n_rows, code = p.rows[blockIdx.x % p.n_rows];
    const size_t n = (size_t)1 << p.degree_bits, row = code & 0x7fffffffu;
    const uint64_t* const d = p.consts + (size_t)ch * (4 + p.n_luts);
    const uint64_t da = d[0], db = d[1], dalpha = d[2], ddelta = d[3];
    uint64_t* const out = p.out + (size_t)ch * p.np1 * n + row;
    auto wire = [&](unsigned c) { return p.wires[(size_t)c * n + row]; };
    const bool lut = (code >> 31) != 0;
    const unsigned n_slots = lut ? p.num_lut_slots : p.num_lu_slots, deg = lut ? p.lut_degree : p.lu_degree;
    const unsigned s = threadIdx.x;
    uint64_t t = 0;
    if (s < n_slots) {
        if (lut)
            t = gl::mul(wire(3 * s + 2), inverse(gl::sub(dalpha, gl::add(wire(3 * s), gl::mul(da, wire(3 * s + 1))))));
        else
            t = inverse(gl::sub(dalpha, gl::add(wire(2 * s), gl::mul(da, wire(2 * s + 1)))));
    } else if (lut && s == blockDim.x - 1) {
        uint64_t re = 0;
        for (unsigned q = 0; q < p.num_lut_slots; q++)
            re = gl::add(gl::mul(re, ddelta), gl::add(wire(3 * q), gl::mul(db, wire(3 * q + 1))));
        out[0] = re;
    }
    term[s] = t;
    __syncthreads();
    if (s < p.np1 - 1) {
        uint64_t sum = 0;
        const unsigned hi = min((s + 1) * deg, n_slots);
        for (unsigned q = s * deg; q < hi; q++) sum = gl::add(sum, term[q]);
        out[(size_t)(s + 1) * n] = lut ? sum : gl::sub(0, sum);
    }
}

// phase 2: one thread per challenge walks the tables in the reference's order, rows downwards:
// RE(row) = delta^slots RE(row + 1) + share; SLDC_0(row) = SLDC_last(row + 1) + share_0, SLDC_k(row) = SLDC_(k-1)(row)
// + share_k.  The chain is serial by definition; the values handed from row to row stay in registers, so a row costs
// one multiplication and a few additions on shares whose loads do not depend on the chain.
__global__ void lookup_chain_kernel(LookupPolyParams p) {
    const unsigned ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.nc) return;
    const size_t n = (size_t)1 << p.degree_bits;
    const unsigned npl = p.np1 - 1;
    uint64_t* const out = p.out + (size_t)ch * p.np1 * n;
    const uint64_t delta_pow = gl::pow(p.consts[(size_t)ch * (4 + p.n_luts) + 3], p.num_lut_slots);
    for (unsigned t = 0; t < p.n_luts; t++) {
        const size_t last_lu = p.tables[3 * t], last_lut = p.tables[3 * t + 1], first_lut = p.tables[3 * t + 2];
        // the row after the table (no table's row, qp_circuit_create): whatever the polynomials hold there
        uint64_t re = out[first_lut + 1], last = out[(size_t)npl * n + first_lut + 1];
        for (size_t row = first_lut + 1; row-- > last_lu;) {
            if (row >= last_lut) {
                re = gl::canon(gl::add(gl::mul(re, delta_pow), out[row]));
                out[row] = re;
            }
            uint64_t acc = last;
            for (unsigned k = 0; k < npl; k++) {
                acc = gl::canon(gl::add(acc, out[(size_t)(k + 1) * n + row]));
                out[(size_t)(k + 1) * n + row] = acc;
            }
            last = acc;
        }
    }
}

}  // namespace quotient
