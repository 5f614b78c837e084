// FRI commit-phase kernels over F_p^2 = F_p[X]/(X^2 - 7) (field/src/goldilocks_extensions.rs:13-26).
//
// Device layout: an array of n extension elements is kept as two base-field planes
// plane[k][i] (k = coordinate), because an F_p^2 FFT with base-field twiddles is two independent
// base-field FFTs (field/src/extension/mod.rs:73-76) and the NTT kernels want unit stride.
// Values are kept in bit-reversed order -- the order fri_committed_trees commits them in
// (plonky2/src/fri/prover.rs:98-104) and the order the DIF NTT produces.
#pragma once
#include "poseidon.cuh"
#include "merkle.cuh"

namespace fri {


struct Ext {
    uint64_t c0, c1;
};

// field/src/extension/quadratic.rs:186-199 with W = 7
__device__ __forceinline__ Ext ext_mul(Ext a, Ext b) {
    uint64_t a1b1 = gl::mul(a.c1, b.c1);
    // 7 * x: 67-bit product reduced with reduce96
    unsigned __int128 w = (unsigned __int128)a1b1 * 7u;
    uint64_t w7 = gl::reduce96((uint64_t)w, (uint32_t)(w >> 64));
    Ext r;
    r.c0 = gl::add(gl::mul(a.c0, b.c0), w7);
    r.c1 = gl::add(gl::mul(a.c0, b.c1), gl::mul(a.c1, b.c0));
    return r;
}

// [n][2] interleaved -> planes [2][n]; if BITREV, plane[k][i] = in[bitrev(i)][k]
// (reverse_index_bits_in_place(values), prover.rs:98).
__global__ void ext_to_planes_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ planes,
                                     unsigned lg_n, int bitrev) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)1 << lg_n;
    if (i >= n) return;
    const size_t src = bitrev ? (lg_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - lg_n)) : 0) : i;
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(in)[src];
    planes[i] = gl::canon(v.x);
    planes[n + i] = gl::canon(v.y);
}

// planes [2][n] -> [count][2] interleaved rows starting at element `first`
__global__ void planes_to_ext_kernel(const uint64_t* __restrict__ planes, size_t n, size_t first,
                                     size_t count, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    ulonglong2 v;
    v.x = gl::canon(planes[first + i]);
    v.y = gl::canon(planes[n + first + i]);
    reinterpret_cast<ulonglong2*>(out)[i] = v;
}

// Fold: out[i] = sum_{j < arity} c[i*arity + j] beta^j  (reduce_with_powers,
// core/src/plonk_common.rs:87-98; prover.rs:111-117).  Planes in, planes out (natural order).
__global__ void fold_kernel(const uint64_t* __restrict__ in, size_t n_in, unsigned arity_bits,
                            uint64_t beta0, uint64_t beta1, uint64_t* __restrict__ out) {
    const size_t n_out = n_in >> arity_bits;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const unsigned arity = 1u << arity_bits;
    const uint64_t* p0 = in + (i << arity_bits);
    const uint64_t* p1 = in + n_in + (i << arity_bits);
    Ext beta{beta0, beta1};
    Ext acc{0, 0};
    for (unsigned j = arity; j-- > 0;) {
        acc = ext_mul(acc, beta);
        acc.c0 = gl::add(acc.c0, p0[j]);
        acc.c1 = gl::add(acc.c1, p1[j]);
    }
    out[i] = gl::canon(acc.c0);
    out[n_out + i] = gl::canon(acc.c1);
}

// Batch FRI: when the folded codeword has reached the length of the next (smaller) polynomial, it absorbs
// that polynomial's values, v <- v * beta + w (plonky2/src/batch_fri/prover.rs:126-137).  Both operands are
// planes in bit-reversed order over domains of the same size, so the update is elementwise.
__global__ void mix_kernel(uint64_t* __restrict__ v, const uint64_t* __restrict__ w, size_t n, uint64_t beta0,
                           uint64_t beta1) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ext r = ext_mul(Ext{v[i], v[n + i]}, Ext{beta0, beta1});
    v[i] = gl::canon(gl::add(r.c0, w[i]));
    v[n + i] = gl::canon(gl::add(r.c1, w[n + i]));
}

// Proof-of-work search (plonky2/src/fri/prover.rs:185-200).  Candidate w = base + global thread
// id; `found` holds the smallest successful candidate so far (init UINT64_MAX).  A launch covers
// far more candidates than the expected 2^min_lz, and a block whose candidates are all larger than
// a witness already found returns at once: the search costs the waves it needs, not a host round
// trip per batch.  (A block is only skipped when a SMALLER candidate succeeded, so the minimum --
// the serial `find` rule -- is kept whatever order the blocks run in.)
__global__ void __launch_bounds__(128)
pow_kernel(const uint64_t* __restrict__ state12, unsigned witness_pos, unsigned min_lz,
           uint64_t base, unsigned long long* found) {
    const uint64_t block_base = base + (uint64_t)blockIdx.x * blockDim.x;
    if (*(volatile unsigned long long*)found < block_base) return;
    const uint64_t w = block_base + threadIdx.x;
    if (w >= gl::P) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = (k == (int)witness_pos) ? w : state12[k];
    poseidon::permute(s);
    const uint64_t resp = gl::canon(s[7]);
    const unsigned lz = resp ? (unsigned)__clzll((long long)resp) : 64u;
    if (lz >= min_lz) atomicMin(found, (unsigned long long)w);
}


}  // namespace fri
