// Opening-side kernels: everything `prove_openings` does with the committed polynomials before the
// FRI commit phase, kept on the device so that the fold inputs never cross PCIe.
//
// Reference semantics:
//   PolynomialCoeffs::eval / OpeningSet::new          field/src/polynomial/mod.rs:155-160, plonky2/src/plonk/proof.rs:289-327
//   reduce_openings_to_unmasked_final_poly            plonky2/src/fri/oracle.rs:129-165
//     composition = sum_i alpha^i expr_i(X)           core/src/reducing.rs:63-72 (reduce_polys)
//     quotient    = (composition - composition(z)) / (X - z)   field/src/polynomial/division.rs:77-90 (divide_by_linear)
//     final       = final * alpha^count + quotient    core/src/reducing.rs:94-97 (shift_poly)
//
// divide_by_linear is a serial Horner scan in the reference.  Here it is the same function written
// as an additive suffix scan:  q_k = sum_{i>k} c_i z^(i-k-1) = z^-(k+1) * sum_{i>=k+1} c_i z^i,
// exact in the field, so the coefficients are identical (z != 0; z = 0 degenerates to a shift).
#pragma once
#include "fri.cuh"

namespace openings {

using fri::Ext;
using fri::ext_mul;

__device__ __forceinline__ Ext ext_add(Ext a, Ext b) { return Ext{gl::add(a.c0, b.c0), gl::add(a.c1, b.c1)}; }

// ext * base
__device__ __forceinline__ Ext ext_scale(Ext a, uint64_t b) { return Ext{gl::mul(a.c0, b), gl::mul(a.c1, b)}; }

// pw[k][i] = z^i for i < n (planes [2][n]); sq[b] = z^(2^b) (host-computed, [32][2])
__global__ void power_table_kernel(const uint64_t* __restrict__ sq, unsigned lg_n, uint64_t* __restrict__ pw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)1 << lg_n;
    if (i >= n) return;
    Ext acc{1, 0};
    for (unsigned b = 0; b < lg_n; b++)
        if ((i >> b) & 1) acc = ext_mul(acc, Ext{sq[2 * b], sq[2 * b + 1]});
    pw[i] = gl::canon(acc.c0);
    pw[n + i] = gl::canon(acc.c1);
}

// out[p] = sum_j c_p[j] * z^j.  A polynomial is cut into gridDim.y segments (one block each: a batch has 20 to
// ~150 polynomials, fewer than the SMs, and one block per polynomial leaves the machine at eight warps per SM --
// 2.3 ms per batch at 2^20 coefficients); the segments' sums meet in eval_polys_finish_kernel.
__global__ void __launch_bounds__(256)
eval_polys_kernel(const uint64_t* __restrict__ coeffs, size_t n, const uint64_t* __restrict__ pw,
                  uint64_t* __restrict__ partial /* [n_polys][gridDim.y][2] */) {
    const uint64_t* c = coeffs + (size_t)blockIdx.x * n;
    const size_t seg = (n + gridDim.y - 1) / gridDim.y;
    const size_t j0 = (size_t)blockIdx.y * seg, j1 = j0 + seg < n ? j0 + seg : n;
    Ext acc{0, 0};
    for (size_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const uint64_t v = c[j];
        acc = ext_add(acc, ext_scale(Ext{pw[j], pw[n + j]}, v));
    }
    __shared__ uint64_t s0[256], s1[256];
    s0[threadIdx.x] = acc.c0;
    s1[threadIdx.x] = acc.c1;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) {
            s0[threadIdx.x] = gl::add(s0[threadIdx.x], s0[threadIdx.x + h]);
            s1[threadIdx.x] = gl::add(s1[threadIdx.x], s1[threadIdx.x + h]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint64_t* o = partial + 2 * ((size_t)blockIdx.x * gridDim.y + blockIdx.y);
        o[0] = s0[0];
        o[1] = s1[0];
    }
}

// out[p] = sum over the segments of polynomial p (one thread per polynomial)
__global__ void eval_polys_finish_kernel(const uint64_t* __restrict__ partial, unsigned n_polys, unsigned n_seg,
                                         uint64_t* __restrict__ out) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_polys) return;
    Ext acc{0, 0};
    for (unsigned s = 0; s < n_seg; s++)
        acc = ext_add(acc, Ext{partial[2 * ((size_t)p * n_seg + s)], partial[2 * ((size_t)p * n_seg + s) + 1]});
    out[2 * p] = gl::canon(acc.c0);
    out[2 * p + 1] = gl::canon(acc.c1);
}

// d[j] = (sum_t w_t * poly_t[j]) * z^j   (composition polynomial times the power table)
// polys: device array of pointers, weights [n_terms][2].
__global__ void __launch_bounds__(256)
weighted_sum_kernel(const uint64_t* const* __restrict__ polys, const uint64_t* __restrict__ weights,
                    unsigned n_terms, size_t n, const uint64_t* __restrict__ pw, uint64_t* __restrict__ d) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Ext acc{0, 0};
    for (unsigned t = 0; t < n_terms; t++) {
        const uint64_t v = polys[t][j];
        acc = ext_add(acc, ext_scale(Ext{weights[2 * t], weights[2 * t + 1]}, v));
    }
    if (pw) acc = ext_mul(acc, Ext{pw[j], pw[n + j]});
    d[j] = acc.c0;
    d[n + j] = acc.c1;
}

// Additive suffix scan of planes [2][n], in place, in three launches.
//   phase 1: each block scans its 1024-element segment (suffix sums), writes the segment total
//   phase 2: one block turns the totals into exclusive suffix sums of the totals
//   phase 3: each block adds its carry
constexpr int SCAN_SEG = 1024;

__global__ void __launch_bounds__(256)
suffix_scan_segments_kernel(uint64_t* __restrict__ d, size_t n, uint64_t* __restrict__ totals /* [2][n_seg] */) {
    __shared__ uint64_t sh[2][SCAN_SEG];
    const size_t seg = blockIdx.x, base = seg * SCAN_SEG;
    const size_t n_seg = gridDim.x;
    for (int k = 0; k < 2; k++)
        for (int e = threadIdx.x; e < SCAN_SEG; e += 256) sh[k][e] = base + e < n ? d[k * n + base + e] : 0;
    __syncthreads();
    // Hillis-Steele suffix scan: sh[e] += sh[e + off]
    for (int off = 1; off < SCAN_SEG; off <<= 1) {
        uint64_t t[2][4];
        for (int k = 0; k < 2; k++)
            for (int q = 0; q < 4; q++) {
                const int e = threadIdx.x + 256 * q;
                t[k][q] = e + off < SCAN_SEG ? gl::add(sh[k][e], sh[k][e + off]) : sh[k][e];
            }
        __syncthreads();
        for (int k = 0; k < 2; k++)
            for (int q = 0; q < 4; q++) sh[k][threadIdx.x + 256 * q] = t[k][q];
        __syncthreads();
    }
    for (int k = 0; k < 2; k++)
        for (int e = threadIdx.x; e < SCAN_SEG; e += 256)
            if (base + e < n) d[k * n + base + e] = sh[k][e];
    if (threadIdx.x == 0) {
        totals[seg] = sh[0][0];
        totals[n_seg + seg] = sh[1][0];
    }
}

// totals[k][s] <- sum_{s' > s} totals[k][s']   (single block; n_seg is at most a few thousand)
__global__ void suffix_scan_totals_kernel(uint64_t* __restrict__ totals, size_t n_seg) {
    if (threadIdx.x < 2) {
        uint64_t* t = totals + threadIdx.x * n_seg;
        uint64_t acc = 0;
        for (size_t s = n_seg; s-- > 0;) {
            const uint64_t v = t[s];
            t[s] = acc;
            acc = gl::add(acc, v);
        }
    }
}

// quotient/accumulate: with T = suffix sums of d (after adding the segment carry),
//   q_k = T_{k+1} * zinv^(k+1)  (k < n-1),  q_{n-1} = 0          (divide_by_linear + the zero pad)
//   final_k = final_k * shift + q_k                                 (shift_poly, then +=)
// ipw = power table of z^-1.  `first` = the running sum is still empty.
__global__ void __launch_bounds__(256)
quotient_accumulate_kernel(const uint64_t* __restrict__ T, const uint64_t* __restrict__ totals, size_t n,
                           const uint64_t* __restrict__ ipw, uint64_t s0, uint64_t s1, int first,
                           uint64_t* __restrict__ fin) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t n_seg = (n + SCAN_SEG - 1) / SCAN_SEG;
    Ext q{0, 0};
    if (k + 1 < n) {
        const size_t i = k + 1, seg = i / SCAN_SEG;
        Ext t{gl::add(T[i], totals[seg]), gl::add(T[n + i], totals[n_seg + seg])};
        q = ext_mul(t, Ext{ipw[i], ipw[n + i]});
    }
    if (!first) q = ext_add(ext_mul(Ext{fin[k], fin[n + k]}, Ext{s0, s1}), q);
    fin[k] = gl::canon(q.c0);
    fin[n + k] = gl::canon(q.c1);
}

// z == 0: (p(X) - p(0)) / X is a shift of the coefficients.
__global__ void shift_accumulate_kernel(const uint64_t* __restrict__ comp, size_t n, uint64_t s0, uint64_t s1,
                                        int first, uint64_t* __restrict__ fin) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Ext q{0, 0};
    if (k + 1 < n) q = Ext{comp[k + 1], comp[n + k + 1]};
    if (!first) q = ext_add(ext_mul(Ext{fin[k], fin[n + k]}, Ext{s0, s1}), q);
    fin[k] = gl::canon(q.c0);
    fin[n + k] = gl::canon(q.c1);
}

}  // namespace openings
