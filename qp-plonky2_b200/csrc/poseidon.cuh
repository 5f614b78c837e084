// Poseidon-Goldilocks permutation (width 12, 4 + 22 + 4 rounds, x^7) for sm_100a, one
// permutation per thread with the whole state in registers.
//
// Reference semantics: core/src/poseidon.rs:599-633 (poseidon / poseidon_naive -- the two are
// the same function; the KATs at core/src/poseidon_goldilocks.rs:455-490 pin it).
//
// B200 formulation (not the reference's), details further down and in DESIGN.md section 4.3:
//   * the linear layers run on the FP64 pipe as EXACT integer arithmetic (DFMA on 32-bit halves,
//     every row sum < 2^50), the S-box on the integer pipes: the work is spread over four issue
//     pipes instead of saturating the fma-heavy one with IMAD.WIDE;
//   * the reference replaces the 22 partial rounds' MDS by sparse matrices with full 64-bit
//     entries (FAST_PARTIAL_*, 23 modular multiplies per round); here partial rounds are fused in
//     PAIRS with the small-entry matrix M^2 instead;
//   * one loop body per round type keeps the code inside the instruction cache.
// Exact arithmetic => identical outputs to the reference.
#pragma once
#include <cuda_runtime.h>

#include "goldilocks.cuh"
#include "poseidon_constants.h"

namespace poseidon {

static constexpr int WIDTH = 12;
static constexpr int RATE = 8;
static constexpr int N_ROUNDS = 30;
static constexpr int N_PARTIAL_PAIRS = 11;  // rounds 4..25 in fused pairs

struct Mat {
    uint32_t a[12][12];
};

// M = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...)  (poseidon_goldilocks.rs:24-25):
// M[r][c] = CIRC[(c - r) mod 12] + (r == c ? DIAG[r] : 0), core/src/poseidon.rs:178-198.
__host__ __device__ constexpr Mat mds_matrix() {
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    Mat m{};
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) m.a[r][c] = C[(c - r + 12) % 12] + ((r == 0 && c == 0) ? 8u : 0u);
    return m;
}
__host__ __device__ constexpr Mat mat_mul(const Mat& x, const Mat& y) {
    Mat z{};
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) {
            uint64_t acc = 0;
            for (int k = 0; k < 12; k++) acc += (uint64_t)x.a[r][k] * y.a[k][c];
            z.a[r][c] = (uint32_t)acc;  // < 2^17 for M^2 (row sums of M are <= 264)
        }
    return z;
}
static constexpr Mat M1 = mds_matrix();
// device code cannot name namespace-scope constexpr objects: each device function re-declares
// the matrices as local compile-time constants (all entries fold into immediates)
#define QP_POSEIDON_MATS                                 \
    constexpr Mat m1 = mds_matrix();                     \
    constexpr Mat m2 = mat_mul(m1, m1);                  \
    (void)m2;

// Device constants, uploaded once per context by upload_constants():
//   c_rc[12 * r + i]      ALL_ROUND_CONSTANTS (core/src/poseidon.rs:57-155), plus a zero row 30
__constant__ uint64_t c_rc[WIDTH * (N_ROUNDS + 1)];
// FP64-pipe formulation (see f64 below): the same constants as (2^52 + low half, 2^52 + high half)
// doubles, and the constants of the 11 fused PAIRS of partial rounds:
//   c_pair_k_d[g]        rc'_0                      (rc', rc'' = constants of the two rounds
//   c_pair_K_d[g][0..11] M rc' + rc''                FOLLOWING the pair's first round)
__constant__ double c_rc_d[WIDTH * (N_ROUNDS + 1)][2];
// c_rc_dd[6 * round + r] = c_rc_d[12 * round + r + 6] - c_rc_d[12 * round + r]   (split layer, below)
__constant__ double c_rc_dd[6 * (N_ROUNDS + 1)][2];
__constant__ double c_pair_k_d[N_PARTIAL_PAIRS][2];
__constant__ double c_pair_K_d[N_PARTIAL_PAIRS][WIDTH][2];
__constant__ double c_pair_K_dd[N_PARTIAL_PAIRS][6][2];   // c_pair_K_d[g][r + 6] - c_pair_K_d[g][r]

// Host side: derive the group constants (mod p, exact) and upload everything.
static inline cudaError_t upload_constants(cudaStream_t stream) {
    typedef unsigned __int128 u128;
    const uint64_t P = gl::P;
    static uint64_t rc[WIDTH * (N_ROUNDS + 1)];
    for (int i = 0; i < WIDTH * N_ROUNDS; i++) rc[i] = POSEIDON_ALL_ROUND_CONSTANTS[i];
    for (int i = 0; i < WIDTH; i++) rc[WIDTH * N_ROUNDS + i] = 0;
    auto matvec = [&](const Mat& m, const uint64_t* v, uint64_t* out) {
        for (int r = 0; r < 12; r++) {
            u128 acc = 0;
            for (int c = 0; c < 12; c++) acc += (u128)m.a[r][c] * (v[c] % P);
            out[r] = (uint64_t)(acc % P);
        }
    };
    // FP64 tables
    static double rcd[WIDTH * (N_ROUNDS + 1)][2];
    static uint64_t pk[N_PARTIAL_PAIRS], pK[N_PARTIAL_PAIRS][WIDTH];
    static double pkd[N_PARTIAL_PAIRS][2], pKd[N_PARTIAL_PAIRS][WIDTH][2];
    // (2^52 + low half, 2^52 + high half) of c' = c - (1 + K) 2^64 mod p with 2^32 - K added to the
    // high part, K = 0x43300000: the exponent words of the two finished accumulators then cancel
    // inside fold_row_f64 (derivation there)
    auto split = [&](uint64_t v, double* out) {
        const uint64_t K = 0x43300000ULL;
        const uint64_t two64 = (uint64_t)((((u128)1) << 64) % P);
        const uint64_t excess = (uint64_t)((u128)(1 + K) * two64 % P);
        const uint64_t c = (uint64_t)(((u128)(v % P) + P - excess) % P);
        out[0] = 4503599627370496.0 + (double)(uint32_t)c;
        out[1] = 4503599627370496.0 + (double)(uint32_t)(c >> 32) + (double)((1ULL << 32) - K);
    };
    for (int i = 0; i < WIDTH * (N_ROUNDS + 1); i++) split(rc[i], rcd[i]);
    static double rcdd[6 * (N_ROUNDS + 1)][2];
    for (int r = 0; r <= N_ROUNDS; r++)
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < 2; k++) rcdd[6 * r + i][k] = rcd[12 * r + i + 6][k] - rcd[12 * r + i][k];  // exact
    for (int g = 0; g < N_PARTIAL_PAIRS; g++) {
        const int r = 4 + 2 * g;
        const uint64_t* r1 = rc + 12 * (r + 1);
        const uint64_t* r2 = rc + 12 * (r + 2);
        uint64_t m1r1[12];
        matvec(M1, r1, m1r1);
        pk[g] = r1[0];
        split(pk[g], pkd[g]);
        for (int i = 0; i < 12; i++) {
            pK[g][i] = (uint64_t)(((u128)m1r1[i] + r2[i]) % P);
            split(pK[g][i], pKd[g][i]);
        }
    }
    static double pKdd[N_PARTIAL_PAIRS][6][2];
    for (int g = 0; g < N_PARTIAL_PAIRS; g++)
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < 2; k++) pKdd[g][i][k] = pKd[g][i + 6][k] - pKd[g][i][k];  // exact
    cudaError_t e = cudaMemcpyToSymbolAsync(c_rc, rc, sizeof rc, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_pair_K_dd, pKdd, sizeof pKdd, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_rc_d, rcd, sizeof rcd, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_rc_dd, rcdd, sizeof rcdd, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_pair_k_d, pkd, sizeof pkd, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_pair_K_d, pKd, sizeof pKd, 0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // the sources are static host arrays
    return e;
}

// (x0^7 - x0) split into halves
__device__ __forceinline__ void sbox_delta(uint64_t x0, uint32_t& dlo, uint32_t& dhi) {
    gl::unpack(gl::sub(gl::pow7(x0), x0), dlo, dhi);
}

// ---- FP64-pipe formulation of the linear layers ---------------------------------------------
// Measured on B200 (tools/microbench/int_pipes.cu, profiles/r01_int_pipe_microbench.txt):
// IMAD.WIDE.U32 costs 4.3 issue cycles per warp instruction and overlaps with almost nothing,
// while DFMA costs 2.3 on its own pipe and overlaps with the integer pipes.  A product
// (32-bit half) x (matrix entry < 2^17) summed over a row stays below 2^50, so it is EXACT in a
// double: the linear layers run on the FP64 pipe, bit-exactly, at half the cost.
//   in : u32 half h -> double: bits (0x43300000 : h) are 2^52 + h; one DADD removes the 2^52.
//   out: accumulators start at 2^52 + (constant half), so the finished sum is 2^52 + v with
//        v < 2^52 sitting in the mantissa: no conversion instruction, just the two words.
//        (fold_row_f64 also cancels the exponent words, so the read-back is free.)
// The 22 partial rounds are fused in PAIRS (entries of M^2 < 2^17 keep two 32-bit limbs exact):
//     x'' = M^2 x + d1 M^2 e0 + d2 M e0 + K,   d = x0^7 - x0,
// 362 DFMA per two rounds before the split below.  (Moving rows back to IMAD.WIDE was measured:
// every row on the FP64 pipe is best, profiles/r01d_variants_fp64_rows.txt.)
#ifndef QP_POSEIDON_I2F   // 1: cvt.rn.f64.u32 (I2F, XU pipe, otherwise idle) instead of the 2^52 trick for the inputs
#define QP_POSEIDON_I2F 1
#endif

namespace f64 {
__device__ __forceinline__ double from_u32(uint32_t h) {
#if QP_POSEIDON_I2F
    return (double)h;
#else
    return __hiloint2double(0x43300000, (int)h) - 4503599627370496.0;
#endif
}
}  // namespace f64

// Finished FP64 accumulators -> field element.  al = 2^52 + a, ah = 2^52 + h + 2^32 - K with
// K = 0x43300000 (the exponent word of 2^52) and a, h < 2^51.  Read as raw words, al is the
// integer a + K 2^32 and ah is h + 2^32 - K + K 2^32, so
//     raw(al) + raw(ah) 2^32 = a + h 2^32 + (1 + K) 2^64 :
// the exponent words cancel except for a constant multiple of 2^64, which the host subtracted
// from the round constant (upload_constants).  No masking, no conversion instruction.
__device__ __forceinline__ uint64_t fold_row_f64(double al, double ah) {
    // high words are K + small: their sum stays below 2^32 as gl::fold3 requires
    return gl::fold3((uint32_t)__double2loint(al), (uint32_t)__double2hiint(al),
                     (uint32_t)__double2loint(ah), (uint32_t)__double2hiint(ah));
}

// Full-round linear layer with the circulant split by the CRT  z^12 - 1 = (z^6 - 1)(z^6 + 1):
// with u = x[0..6) + x[6..12), v = x[0..6) - x[6..12),
//     P_r = sum_j c+_j u[(j + r) mod 6]              (cyclic,     c+_j = (c_j + c_{j+6}) / 2)
//     Q_r = sum_j c-_j s(j + r) v[(j + r) mod 6]     (negacyclic, c-_j = (c_j - c_{j+6}) / 2, s = -1 on wrap)
//     y_r = P_r + Q_r,   y_{r+6} = P_r - Q_r          (r < 6)
// Every c_j + c_{j+6} and c_j - c_{j+6} of this MDS matrix is even, so the halved coefficients
// (15,14,40,17,18,24) and (2,1,1,-1,-16,4) are integers and everything stays exact: 103 FP64
// operations per plane instead of 144.  P_r starts at the constant of row r; row r + 6 adds the
// difference of the two constants (exact: both are 2^52 + a 33-bit integer).
#ifndef QP_MDS_COLMAJOR   // 1: accumulate column by column (all twelve chains consume u[0], v[0] first, ...)
#define QP_MDS_COLMAJOR 0
#endif
#ifndef QP_ROUND_FUSED    // 1: full rounds run S-box -> conversion -> accumulation per lane pair (j, j + 6)
#define QP_ROUND_FUSED 0
#endif

// sign-adjusted negacyclic coefficient: row r, column j of the 6 x 6 negacyclic block
__host__ __device__ constexpr double negacyc(const double (&cm)[6], int r, int j) {
    return (j >= r) ? cm[j - r] : -cm[j - r + 6];
}

__device__ __forceinline__ void mds_layer_split(uint64_t (&s)[12], int round) {
    constexpr double CP[6] = {15, 14, 40, 17, 18, 24};
    constexpr double CM[6] = {2, 1, 1, -1, -16, 4};
    uint32_t w[2][12];
#pragma unroll
    for (int i = 0; i < 12; i++) gl::unpack(s[i], w[0][i], w[1][i]);
    double y[2][12];
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        double d[12], u[6], v[6];
#pragma unroll
        for (int i = 0; i < 12; i++) d[i] = f64::from_u32(w[pl][i]);
#pragma unroll
        for (int j = 0; j < 6; j++) {
            u[j] = d[j] + d[j + 6];
            v[j] = d[j] - d[j + 6];
        }
#if QP_MDS_COLMAJOR
        double P[6], Q[6];
#pragma unroll
        for (int r = 0; r < 6; r++) P[r] = c_rc_d[12 * round + r][pl];
#pragma unroll
        for (int j = 0; j < 6; j++) {
#pragma unroll
            for (int r = 0; r < 6; r++) {
                P[r] = fma(u[j], CP[(j - r + 6) % 6], P[r]);
                Q[r] = (j == 0) ? v[j] * negacyc(CM, r, j) : fma(v[j], negacyc(CM, r, j), Q[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            y[pl][r] = P[r] + Q[r];
            y[pl][r + 6] = (P[r] - Q[r]) + c_rc_dd[6 * round + r][pl];
        }
#else
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double P = c_rc_d[12 * round + r][pl];
            double Q = v[r] * CM[0];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int idx = (j + r) % 6;
                P = fma(u[idx], CP[j], P);
                if (j > 0) Q = fma(v[idx], (j + r < 6) ? CM[j] : -CM[j], Q);
            }
            y[pl][r] = P + Q;
            y[pl][r + 6] = (P - Q) + c_rc_dd[6 * round + r][pl];
        }
#endif
        y[pl][0] = fma(d[0], 8.0, y[pl][0]);  // the diagonal entry of row 0
    }
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = fold_row_f64(y[0][r], y[1][r]);
}

// Scheduling fence experiments (ptxas groups all integer work before all FP64 work otherwise)
#ifndef QP_SCHED_FENCE
#define QP_SCHED_FENCE 0
#endif
__device__ __forceinline__ void sched_fence() {
#if QP_SCHED_FENCE == 1
    asm volatile("bar.warp.sync 0xffffffff;");
#elif QP_SCHED_FENCE == 2
    asm volatile("pmevent 1;");
#elif QP_SCHED_FENCE == 3
    { uint32_t t_; asm volatile("mov.u32 %0, %%clock;" : "=r"(t_)); }
#elif QP_SCHED_FENCE == 4
    { uint32_t t_; asm volatile("{ .reg .pred p; mov.u32 %0, %%laneid; setp.eq.u32 p, %0, 77; @p trap; }" : "=r"(t_)); }
#elif QP_SCHED_FENCE == 5
    asm volatile("nanosleep.u32 0;");
#endif
}

// S-box layer + linear layer of a full round, lane pair by lane pair: the FP64 accumulation of pair j
// depends only on the S-boxes of lanes j and j + 6, so the integer work of the next pair can overlap it.
__device__ __forceinline__ void full_round_fused(uint64_t (&s)[12], int round) {
    constexpr double CP[6] = {15, 14, 40, 17, 18, 24};
    constexpr double CM[6] = {2, 1, 1, -1, -16, 4};
    double P[2][6], Q[2][6], d0[2];
#pragma unroll
    for (int pl = 0; pl < 2; pl++)
#pragma unroll
        for (int r = 0; r < 6; r++) P[pl][r] = c_rc_d[12 * round + r][pl];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const uint64_t a = gl::pow7(s[j]), b = gl::pow7(s[j + 6]);
        uint32_t wa[2], wb[2];
        gl::unpack(a, wa[0], wa[1]);
        gl::unpack(b, wb[0], wb[1]);
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const double da = f64::from_u32(wa[pl]), db = f64::from_u32(wb[pl]);
            if (j == 0) d0[pl] = da;
            const double u = da + db, v = da - db;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                P[pl][r] = fma(u, CP[(j - r + 6) % 6], P[pl][r]);
                Q[pl][r] = (j == 0) ? v * negacyc(CM, r, j) : fma(v, negacyc(CM, r, j), Q[pl][r]);
            }
        }
        sched_fence();
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double y0[2], y1[2];
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            y0[pl] = P[pl][r] + Q[pl][r];
            y1[pl] = (P[pl][r] - Q[pl][r]) + c_rc_dd[6 * round + r][pl];
            if (r == 0) y0[pl] = fma(d0[pl], 8.0, y0[pl]);
        }
        s[r] = fold_row_f64(y0[0], y0[1]);
        s[r + 6] = fold_row_f64(y1[0], y1[1]);
    }
}

// Software-pipelined form: segment j holds the S-boxes of lane pair j AND the FP64 accumulation of pair
// j - 1 (independent work for two different pipes), segments separated by scheduling fences.
__device__ __forceinline__ void full_round_pipelined(uint64_t (&s)[12], int round) {
    constexpr double CP[6] = {15, 14, 40, 17, 18, 24};
    constexpr double CM[6] = {2, 1, 1, -1, -16, 4};
    double P[2][6], Q[2][6], d0[2], da[2], db[2];
#pragma unroll
    for (int pl = 0; pl < 2; pl++)
#pragma unroll
        for (int r = 0; r < 6; r++) P[pl][r] = c_rc_d[12 * round + r][pl];
#pragma unroll
    for (int j = 0; j <= 6; j++) {
        if (j > 0) {
            // accumulation of pair j - 1
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
                const double u = da[pl] + db[pl], v = da[pl] - db[pl];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    P[pl][r] = fma(u, CP[(j - 1 - r + 6) % 6], P[pl][r]);
                    Q[pl][r] = (j == 1) ? v * negacyc(CM, r, j - 1) : fma(v, negacyc(CM, r, j - 1), Q[pl][r]);
                }
            }
        }
        if (j < 6) {
            const uint64_t a = gl::pow7(s[j]), b = gl::pow7(s[j + 6]);
            uint32_t wa[2], wb[2];
            gl::unpack(a, wa[0], wa[1]);
            gl::unpack(b, wb[0], wb[1]);
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
                da[pl] = f64::from_u32(wa[pl]);
                db[pl] = f64::from_u32(wb[pl]);
                if (j == 0) d0[pl] = da[pl];
            }
            sched_fence();
        }
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double y0[2], y1[2];
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            y0[pl] = P[pl][r] + Q[pl][r];
            y1[pl] = (P[pl][r] - Q[pl][r]) + c_rc_dd[6 * round + r][pl];
            if (r == 0) y0[pl] = fma(d0[pl], 8.0, y0[pl]);
        }
        s[r] = fold_row_f64(y0[0], y0[1]);
        s[r + 6] = fold_row_f64(y1[0], y1[1]);
    }
}

// The same CRT split for the pair layer.  M = C + 8 e0 e0^T (C circulant), so
//     M^2 x = C^2 x + 8 (C e0) x0 + e0 * 8 (M x)_0 ,
// C^2 is circulant again (first row c * c, cyclic) with even half-sums / half-differences, and
// (M x)_0 is needed anyway for lane 0 after the first round: 154 FP64 operations per plane
// instead of 181.
struct Split6 {
    double p[6], m[6];
};
__host__ __device__ constexpr Split6 split_c2() {
    constexpr long long C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    long long c2[12] = {};
    for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) c2[(a + b) % 12] += C[a] * C[b];
    Split6 o{};
    for (int j = 0; j < 6; j++) {
        o.p[j] = (double)((c2[j] + c2[j + 6]) / 2);
        o.m[j] = (double)((c2[j] - c2[j + 6]) / 2);
    }
    return o;
}
__host__ __device__ constexpr bool split_c2_exact() {
    constexpr long long C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    long long c2[12] = {};
    for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) c2[(a + b) % 12] += C[a] * C[b];
    for (int j = 0; j < 6; j++)
        if ((c2[j] + c2[j + 6]) % 2 || (c2[j] - c2[j + 6]) % 2) return false;
    return true;
}
static_assert(split_c2_exact(), "C^2 does not split into integer half-sums");

__device__ __forceinline__ void partial_pair_split(uint64_t (&s)[12], int g) {
    QP_POSEIDON_MATS
    constexpr Split6 S = split_c2();
    constexpr double C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t w[2][12];
#pragma unroll
    for (int i = 0; i < 12; i++) gl::unpack(s[i], w[0][i], w[1][i]);
    uint32_t dw[2];
    sbox_delta(s[0], dw[0], dw[1]);
    double d[2][12], sx[2], acc0[2], e1[2];
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
#pragma unroll
        for (int i = 0; i < 12; i++) d[pl][i] = f64::from_u32(w[pl][i]);
        e1[pl] = f64::from_u32(dw[pl]);
        // (M x)_0 = c . x + 8 x0
        double t = d[pl][0] * (C[0] + 8.0);
#pragma unroll
        for (int k = 1; k < 12; k++) t = fma(d[pl][k], C[k], t);
        sx[pl] = t;
        // lane 0 after the first round: (M x)_0 + d1 M00 + rc'_0
        acc0[pl] = fma(e1[pl], (double)m1.a[0][0], t + c_pair_k_d[g][pl]);
    }
    sbox_delta(fold_row_f64(acc0[0], acc0[1]), dw[0], dw[1]);
    double y[2][12];
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        const double e2 = f64::from_u32(dw[pl]);
        double u[6], v[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
            u[j] = d[pl][j] + d[pl][j + 6];
            v[j] = d[pl][j] - d[pl][j + 6];
        }
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double P = c_pair_K_d[g][r][pl];
            double Q = v[r] * S.m[0];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int idx = (j + r) % 6;
                P = fma(u[idx], S.p[j], P);
                if (j > 0) Q = fma(v[idx], (j + r < 6) ? S.m[j] : -S.m[j], Q);
            }
            y[pl][r] = P + Q;
            y[pl][r + 6] = (P - Q) + c_pair_K_dd[g][r][pl];
        }
#pragma unroll
        for (int r = 0; r < 12; r++) {
            // + 8 (C e0)_r x0 + d1 (M^2)_r0 + d2 M_r0
            double t = fma(d[pl][0], 8.0 * C[(12 - r) % 12], y[pl][r]);
            t = fma(e1[pl], (double)m2.a[r][0], t);
            y[pl][r] = fma(e2, (double)m1.a[r][0], t);
        }
        y[pl][0] = fma(sx[pl], 8.0, y[pl][0]);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = fold_row_f64(y[0][r], y[1][r]);
}

// S-box layer as 12/L iterations x L lanes with a register rotation (smaller code) -- with the
// linear layers on the FP64 pipe the code fits the instruction cache fully unrolled, and L = 12 (no
// rotation moves) is the measured best on B200 (leaf hash at 2^21 x 135: L = 6 25.46 ms, L = 12 25.06 ms).
#ifndef QP_POSEIDON_SBOX_ROT
#define QP_POSEIDON_SBOX_ROT 12
#endif

// S-box layer on all 12 lanes (poseidon.rs:554-562)
__device__ __forceinline__ void sbox_all(uint64_t (&s)[12]) {
#if QP_POSEIDON_SBOX_ROT
    constexpr int L = QP_POSEIDON_SBOX_ROT;
#pragma unroll 1
    for (int it = 0; it < 12 / L; it++) {
        uint64_t t[L];
#pragma unroll
        for (int i = 0; i < L; i++) t[i] = gl::pow7(s[i]);
#pragma unroll
        for (int i = 0; i < 12 - L; i++) s[i] = s[i + L];
#pragma unroll
        for (int i = 0; i < L; i++) s[12 - L + i] = t[i];
    }
#else
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::pow7(s[i]);
#endif
}

// The permutation.  State lanes may be any u64 representatives; outputs likewise.
// SYNC: all threads of the block run it together and meet at a barrier per round, which keeps
// the warps of an SM inside the same window of code (instruction-cache working set).
template <bool SYNC = false>
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
    // round 0 constant layer (poseidon.rs:504-513); every later round gets its constants from
    // the linear layer that precedes it
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add1(s[i], c_rc[i]);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        // four full rounds (poseidon.rs:574-581)
        const int base = half * 26;
#pragma unroll 1
        for (int r = base; r < base + 4; r++) {
            if (SYNC) __syncthreads();
#if QP_ROUND_FUSED == 2
            full_round_pipelined(s, r + 1);
#elif QP_ROUND_FUSED
            full_round_fused(s, r + 1);
#else
            sbox_all(s);
            mds_layer_split(s, r + 1);  // row 30 is zero
#endif
        }
        if (half == 0) {
            // 22 partial rounds (poseidon.rs:623-628) as 11 fused pairs
#pragma unroll 1
            for (int g = 0; g < N_PARTIAL_PAIRS; g++) {
                if (SYNC) __syncthreads();
                partial_pair_split(s, g);
            }
        }
    }
}

}  // namespace poseidon
