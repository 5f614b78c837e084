// Goldilocks NTT for sm_100a: multi-pass, radix-2^R register butterflies, shared-memory
// exchange, decimation in frequency (natural order in -> bit-reversed order out).
//
// What is computed (reference semantics):
//   forward:  y_k = sum_i x_i w^(ik), w = primitive_root_of_unity(lg n)      field/src/fft.rs:159-202
//   inverse:  c_i = y_{(n-i) mod n} / n                                       field/src/fft.rs:68-91
//   coset LDE: zero-pad to N = n 2^r, scale by g^i, forward size-N transform  field/src/polynomial/mod.rs:199-201,280-293
//             with output kept in leaf order (index bit-reversed)             plonky2/src/fri/oracle.rs:208-209
// How (B200 formulation, not the reference's radix-2 DIT over a bit-reversed copy):
//   * The size-N transform of a zero-padded input is 2^r independent size-n transforms of
//     c_i (g w_N^e)^i, e = bitrev_r(block): the reference's zero-factor rounds (fft.rs:179-192)
//     are exactly that statement.  Block q of the leaf-ordered output is coset e = bitrev_r(q).
//   * In-place DIF leaves X[bitrev(p)] at position p, which IS leaf order -- the transpose and
//     reverse_index_bits of the reference never run.
//   * Each CTA owns a 4096-element tile (256 threads x 16 elements); a pass performs up to 12
//     stages as rounds of radix 16/8/4/2 butterflies held in registers.  Inside a butterfly all
//     twiddles are powers of two (w_64 = 8 in this field, so w_16 = 2^12, w_8 = 2^24, w_4 = 2^48),
//     and the inter-round twiddles w_M^(j k) come from one table of full rows.
//   * The kernel is integer-issue bound on B200 (about 13 integer instructions per element per
//     stage against ~44 available per 8 bytes of HBM traffic), see DESIGN.md.
#pragma once
#include "goldilocks.cuh"

namespace ntt {

constexpr int TILE_LOG = 12;
constexpr int TILE = 1 << TILE_LOG;  // elements per CTA
constexpr int THREADS = 256;
constexpr int EPT = 16;              // elements per thread
constexpr int SMEM_ELEMS = TILE + (TILE >> 4);
#ifndef QP_NTT_MIN_BLOCKS   // resident CTAs per SM the pass kernels are compiled for (register cap)
#define QP_NTT_MIN_BLOCKS 4   // 64 registers: 32 resident warps per SM; measured 1-2 % faster than 3 CTAs at 78 registers
#endif

__host__ __device__ constexpr int pick_radix(int k) {
    return (k >= 4 && k != 5 && k != 6 && k != 9) ? 4 : (k >= 3 ? 3 : k);
}

__host__ __device__ constexpr uint64_t pow2_mod(int s) {
    // 2^s mod p for 0 <= s < 96 (2^64 = 2^32 - 1 mod p)
    return s < 64 ? (1ULL << s) : (((1ULL << 32) - 1) << (s - 64));
}

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, int bits) {
    return bits ? (__brev(x) >> (32 - bits)) : 0;
}

enum : int { MODE_STRIDED = 0, MODE_FINAL = 1 };
enum : int { OUT_NATURAL = 0, OUT_INVERSE = 1 };

struct PassParams {
    const uint64_t* src;
    uint64_t* dst;
    const uint64_t* tw;  // tw[(1 << lg) + e] = w_{2^lg}^e, 0 <= e < 2^lg  (full rows)
    int L;               // log2 of the vector length
    int s_lo;            // STRIDED: lowest stage of this pass (row stride 2^s_lo)
    int inner_bits;      // vector id v -> outer = v >> inner_bits, inner = v & mask
    unsigned n_vec;
    size_t src_outer_stride;  // src vector = src + outer * src_outer_stride (+ inner * src_inner_stride)
    size_t src_inner_stride;
    size_t dst_outer_stride;
    size_t dst_inner_stride;
    // optional scaling on load: x_i *= scale_lo[inner][i & (2^split - 1)] * scale_hi[inner][i >> split]
    const uint64_t* scale_lo;
    const uint64_t* scale_hi;
    int scale_split;
    // FINAL: output mode
    int out_mode;
    uint64_t n_inv;
};


__device__ __forceinline__ int padi(int idx) { return idx + (idx >> 4); }

// In-register 2^R-point DFT, DIF, constant twiddles w_{2^R}^e = 2^(e*192/2^R); register b ends
// up holding y_{bitrev_R(b)}.
#ifndef QP_NTT_LAZY   // 1: butterflies in the lazy three-word form (goldilocks.cuh); 0: two-fix modular add / sub
#define QP_NTT_LAZY 1
#endif
template <int R>
__device__ __forceinline__ void dft_regs(uint64_t (&x)[1 << R]) {
    constexpr int N = 1 << R;
#if QP_NTT_LAZY
    gl::lz y[N];
#pragma unroll
    for (int q = 0; q < N; q++) y[q] = gl::lz_from(x[q]);
#pragma unroll
    for (int u = 0; u < R; u++) {
        const int half = N >> (u + 1);
#pragma unroll
        for (int q = 0; q < N; q++) {
            if ((q & half) == 0) {
                const gl::lz a = y[q], b = y[q + half];
                y[q] = gl::lz_add(a, b);
                const int e = (q & (half - 1)) << u;  // exponent of w_{2^R}
                const int sh = e * (192 / N);         // < 96
                const gl::lz d = gl::lz_sub(a, b);
                y[q + half] = (e == 0) ? d : gl::lz_mul_pow2(d, sh);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < N; q++) x[q] = gl::lz_reduce(y[q]);
#else
#pragma unroll
    for (int u = 0; u < R; u++) {
        const int half = N >> (u + 1);
#pragma unroll
        for (int q = 0; q < N; q++) {
            if ((q & half) == 0) {
                uint64_t a = x[q], b = x[q + half];
                x[q] = gl::add(a, b);
                const int e = (q & (half - 1)) << u;  // exponent of w_{2^R}
                const int sh = e * (192 / N);         // < 96
                uint64_t d = gl::sub(a, b);
                x[q + half] = (e == 0) ? d : gl::mul(d, pow2_mod(sh));
            }
        }
    }
#endif
}

template <int R>
__device__ __forceinline__ constexpr int brev_c(int b) {
    int r = 0;
    for (int i = 0; i < R; i++) r |= ((b >> i) & 1) << (R - 1 - i);
    return r;
}

// One radix-2^R round over the whole tile.
//   MODE_STRIDED: tile = 2^K rows (stride 2^s_lo) x T = TILE/2^K contiguous columns; butterfly id
//                 decodes as (c fastest, jt, blk); smem index = row * T + c.
//   MODE_FINAL:   tile = T chunks x 2^K contiguous rows; butterfly id decodes as (jt fastest, blk,
//                 chunk); smem index = chunk * 2^K + row.
// FIRST rounds read global memory (through `ld`), the others shared memory.  LAST rounds of a
// strided pass write global memory; every other round writes back to the slots it read.
template <int MODE, int K, int R, int MT_LOG, bool FIRST, bool LAST, class Loader, class Storer>
__device__ __forceinline__ void radix_round(uint64_t* smem, const PassParams& p, int lgM_global,
                                            unsigned j_base, Loader ld, Storer st) {
    constexpr int N = 1 << R;
    constexpr int T_LOG = TILE_LOG - K;
    constexpr int ST_LOG = MT_LOG - R;  // log2 rows between a butterfly's elements
    const uint64_t* row_tw = p.tw + ((size_t)1 << lgM_global);
#pragma unroll
    for (int u = 0; u < (EPT >> R); u++) {
        const int bf = u * THREADS + threadIdx.x;
        int c, jt, blk;
        if (MODE == MODE_STRIDED) {
            c = bf & ((1 << T_LOG) - 1);
            jt = (bf >> T_LOG) & ((1 << ST_LOG) - 1);
            blk = bf >> (T_LOG + ST_LOG);
        } else {
            jt = bf & ((1 << ST_LOG) - 1);
            blk = (bf >> ST_LOG) & ((1 << (K - MT_LOG)) - 1);
            c = bf >> (K - R);
        }
        const int row0 = (blk << MT_LOG) + jt;
        uint64_t x[N];
#pragma unroll
        for (int q = 0; q < N; q++) {
            const int row = row0 + (q << ST_LOG);
            if (FIRST) {
                x[q] = ld(row, c);
            } else {
                const int idx = (MODE == MODE_STRIDED) ? (row << T_LOG) + c : (c << K) + row;
                x[q] = smem[padi(idx)];
            }
        }
        dft_regs<R>(x);
        // inter-round twiddles: register b holds y_k, k = bitrev_R(b); multiply by w_M^(j k)
        unsigned j;
        bool need_tw;
        if (MODE == MODE_STRIDED) {
            j = ((unsigned)jt << p.s_lo) + j_base + c;
            need_tw = true;
        } else {
            j = jt;
            need_tw = ST_LOG > 0;
        }
        if (need_tw) {
#pragma unroll
            for (int b = 1; b < N; b++) {
                const unsigned k = brev_c<R>(b);
                uint64_t w = __ldg(row_tw + (size_t)j * k);
                x[b] = gl::mul(x[b], w);
            }
        }
#pragma unroll
        for (int q = 0; q < N; q++) {
            const int row = row0 + (q << ST_LOG);
            if (LAST && MODE == MODE_STRIDED) {
                st(row, c, x[q]);
            } else {
                const int idx = (MODE == MODE_STRIDED) ? (row << T_LOG) + c : (c << K) + row;
                smem[padi(idx)] = x[q];
            }
        }
    }
}

template <int MODE, int K, int K_REM, bool FIRST, class Loader, class Storer>
__device__ __forceinline__ void run_rounds(uint64_t* smem, const PassParams& p, unsigned j_base,
                                           Loader ld, Storer st) {
    if constexpr (K_REM > 0) {
        constexpr int R = pick_radix(K_REM);
        constexpr bool LAST = (K_REM == R);
        const int lgM = K_REM + (MODE == MODE_STRIDED ? p.s_lo : 0);
        if (!FIRST) __syncthreads();
        radix_round<MODE, K, R, K_REM, FIRST, LAST>(smem, p, lgM, j_base, ld, st);
        run_rounds<MODE, K, K_REM - R, false>(smem, p, j_base, ld, st);
    }
}

__device__ __forceinline__ uint64_t load_scaled(const PassParams& p, const uint64_t* src_vec,
                                                unsigned inner, unsigned i) {
    uint64_t v = src_vec[i];
    if (p.scale_lo) {
        const unsigned lo_mask = (1u << p.scale_split) - 1;
        uint64_t f0 = __ldg(p.scale_lo + ((size_t)inner << p.scale_split) + (i & lo_mask));
        uint64_t f1 = __ldg(p.scale_hi + ((size_t)inner << (p.L - p.scale_split)) + (i >> p.scale_split));
        v = gl::mul(v, gl::mul(f0, f1));
    }
    return v;
}

// Strided pass: stages s_lo .. s_lo+K-1 of every vector.
template <int K>
__global__ void __launch_bounds__(THREADS, QP_NTT_MIN_BLOCKS) strided_pass_kernel(PassParams p) {
    constexpr int T_LOG = TILE_LOG - K;
    extern __shared__ uint64_t smem[];
    const unsigned tiles_per_vec = 1u << (p.L - TILE_LOG);
    const unsigned v = blockIdx.x / tiles_per_vec;
    const unsigned tile = blockIdx.x % tiles_per_vec;
    const unsigned outer = v >> p.inner_bits, inner = v & ((1u << p.inner_bits) - 1);
    const uint64_t* src_vec = p.src + outer * p.src_outer_stride + inner * p.src_inner_stride;
    uint64_t* dst_vec = p.dst + outer * p.dst_outer_stride + inner * p.dst_inner_stride;
    const unsigned col_groups_log = p.s_lo - T_LOG;  // 2^s_lo / T column groups per block
    const unsigned a = tile >> col_groups_log;
    const unsigned c0 = (tile & ((1u << col_groups_log) - 1)) << T_LOG;
    const unsigned base = (a << (p.s_lo + K)) + c0;
    auto ld = [&](int row, int c) -> uint64_t {
        return load_scaled(p, src_vec, inner, base + ((unsigned)row << p.s_lo) + c);
    };
    auto st = [&](int row, int c, uint64_t val) {
        dst_vec[base + ((unsigned)row << p.s_lo) + c] = val;
    };
    run_rounds<MODE_STRIDED, K, K, true>(smem, p, c0, ld, st);
}

// Final pass: the last K stages (contiguous 2^K-element chunks), T = TILE/2^K chunks per CTA.
// Chunks are enumerated over all vectors: g = blockIdx.x * T + u, vector = g >> (L-K).
template <int K>
__global__ void __launch_bounds__(THREADS, QP_NTT_MIN_BLOCKS) final_pass_kernel(PassParams p) {
    constexpr int T_LOG = TILE_LOG - K;
    constexpr int T = 1 << T_LOG;
    extern __shared__ uint64_t smem[];
    const int cb = p.L - K;  // log2 chunks per vector
    const size_t total_chunks = (size_t)p.n_vec << cb;
    const size_t g0 = (size_t)blockIdx.x * T;
    const unsigned inner_mask = (1u << p.inner_bits) - 1;
    auto chunk_info = [&](int c, unsigned& v, unsigned& pc_lin, bool& ok) {
        size_t g = g0 + c;
        ok = g < total_chunks;
        v = (unsigned)(g >> cb);
        pc_lin = (unsigned)(g & (((size_t)1 << cb) - 1));
    };
    auto ld = [&](int row, int c) -> uint64_t {
        unsigned v, pc_lin;
        bool ok;
        chunk_info(c, v, pc_lin, ok);
        if (!ok) return 0;
        const unsigned outer = v >> p.inner_bits, inner = v & inner_mask;
        const uint64_t* src_vec = p.src + outer * p.src_outer_stride + inner * p.src_inner_stride;
        // OUT_INVERSE visits chunks in bit-reversed order so that the permuted stores coalesce
        const unsigned pc = (p.out_mode == OUT_INVERSE) ? bitrev32(pc_lin, cb) : pc_lin;
        return load_scaled(p, src_vec, inner, (pc << K) + row);
    };
    auto st = [&](int, int, uint64_t) {};
    if constexpr (K == 0) {
        // length-1 vectors: no butterflies, every "chunk" is one element
#pragma unroll
        for (int u = 0; u < EPT; u++) {
            const int e = u * THREADS + threadIdx.x;
            smem[padi(e)] = ld(0, e);
        }
    }
    run_rounds<MODE_FINAL, K, K, true>(smem, p, 0, ld, st);
    __syncthreads();
    // copy-out: smem[(c << K) + row] holds position `row` of chunk c (bit-reversed order)
    if (p.out_mode == OUT_NATURAL) {
#pragma unroll
        for (int u = 0; u < EPT; u++) {
            const int e = u * THREADS + threadIdx.x;
            const int row = e & ((1 << K) - 1), c = e >> K;
            unsigned v, pc_lin;
            bool ok;
            chunk_info(c, v, pc_lin, ok);
            if (ok) {
                const unsigned outer = v >> p.inner_bits, inner = v & inner_mask;
                uint64_t* dst_vec = p.dst + outer * p.dst_outer_stride + inner * p.dst_inner_stride;
                dst_vec[((size_t)pc_lin << K) + row] = gl::canon(smem[padi(e)]);
            }
        }
    } else {
        // position p = (pc << K) + q holds Y[bitrev_L(p)] = Y[(bitrev_K(q) << cb) + pc_lin];
        // coefficient index i = (n - that) mod n, value scaled by 1/n.  Threads take c fastest
        // so that consecutive threads write consecutive (descending) addresses.
        const unsigned nmask = (1u << p.L) - 1;
#pragma unroll
        for (int u = 0; u < EPT; u++) {
            const int e = u * THREADS + threadIdx.x;
            const int c = e & (T - 1), q = e >> T_LOG;
            unsigned v, pc_lin;
            bool ok;
            chunk_info(c, v, pc_lin, ok);
            if (ok) {
                const unsigned outer = v >> p.inner_bits, inner = v & inner_mask;
                uint64_t* dst_vec = p.dst + outer * p.dst_outer_stride + inner * p.dst_inner_stride;
                const unsigned yk = (bitrev32((unsigned)q, K) << cb) + pc_lin;
                const unsigned i = (0u - yk) & nmask;
                uint64_t val = gl::mul(smem[padi((c << K) + q)], p.n_inv);
                dst_vec[i] = gl::canon(val);
            }
        }
    }
}


}  // namespace ntt
