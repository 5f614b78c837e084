// Poseidon Merkle tree kernels: leaf hashing straight off the (column-major, leaf-ordered) LDE,
// internal levels, cap -- digests written in the reference's interleaved layout.
//
// Reference semantics:
//   hash_leaf   core/src/hashing.rs:150-168   state[8] = len+1, overwrite-mode absorb of 8-chunks
//   two_to_one  core/src/hashing.rs:47-64     [l0..l3, r0..r3, 0,0,0,0] -> permute -> [0..4)
//   layout      plonky2/src/hash/merkle_tree.rs:56-83,121-160: per cap-subtree block, node k of
//               layer l (0 = leaf digests) lives at 2*((k>>1) << (l+1)) + 2*(2^l - 1) + (k&1);
//               subtree roots go to `cap`.
#pragma once
#include "poseidon.cuh"

namespace merkle {

struct TreeShape {
    unsigned lg_leaves;   // log2(#leaves)
    unsigned cap_height;
    __host__ __device__ unsigned num_layers() const { return lg_leaves - cap_height; }
    __host__ __device__ size_t sub_digests() const {  // digests per cap subtree
        return 2 * (((size_t)1 << num_layers()) - 1);
    }
};


// Slot (in digests, units of 4 u64) of node k of `layer` inside subtree t; layer < num_layers.
__device__ __forceinline__ size_t digest_slot(const TreeShape& sh, unsigned layer, size_t t, size_t k) {
    return t * sh.sub_digests() + 2 * ((k >> 1) << (layer + 1)) + 2 * (((size_t)1 << layer) - 1) + (k & 1);
}

__device__ __forceinline__ void store_digest(uint64_t* dst, const uint64_t (&s)[12]) {
    ulonglong2 a, b;
    a.x = gl::canon(s[0]);
    a.y = gl::canon(s[1]);
    b.x = gl::canon(s[2]);
    b.y = gl::canon(s[3]);
    reinterpret_cast<ulonglong2*>(dst)[0] = a;
    reinterpret_cast<ulonglong2*>(dst)[1] = b;
}

// Where leaf i's digest goes: layer 0 of its subtree, or straight into the cap if the tree is
// all cap (merkle_tree.rs:94-103).
__device__ __forceinline__ uint64_t* leaf_digest_ptr(const TreeShape& sh, uint64_t* digests,
                                                     uint64_t* cap, size_t i) {
    const unsigned nl = sh.num_layers();
    if (nl == 0) return cap + 4 * i;
    const size_t t = i >> nl, k = i & (((size_t)1 << nl) - 1);
    return digests + 4 * digest_slot(sh, 0, t, k);
}

// Where the elements of a leaf live.
//   AffineLayout: element c of leaf i = data[c * col_stride + i * row_stride]
//       column-major LDE (the commit path):    col_stride = N, row_stride = 1  (coalesced across the warp)
//       leaf-major rows (MerkleTree::new API): col_stride = 1, row_stride = leaf_len
//   ExtPlanesLayout: FRI commit leaves = 2^arity_bits consecutive F_p^2 values flattened
//       (flatten, field/src/extension/mod.rs:129-138) out of two coordinate planes.
struct AffineLayout {
    const uint64_t* data;
    size_t col_stride, row_stride;
    __device__ __forceinline__ uint64_t get(size_t i, unsigned c) const {
        return data[(size_t)c * col_stride + i * row_stride];
    }
};
struct ExtPlanesLayout {
    const uint64_t* planes;
    size_t n;
    unsigned arity_bits;
    __device__ __forceinline__ uint64_t get(size_t i, unsigned c) const {
        return planes[(size_t)(c & 1) * n + (i << arity_bits) + (c >> 1)];
    }
};

// BatchMerkleTree stage leaves straight off a column-major LDE: leaf i = cap[i] (the digest the
// previous stage left at this height, 4 words) || the columns at row i
// (plonky2/src/hash/batch_merkle_tree.rs:91-100, batch_fri/oracle.rs:133-147).
struct CapPrefixLayout {
    const uint64_t* cap;
    const uint64_t* data;
    size_t col_stride;
    __device__ __forceinline__ uint64_t get(size_t i, unsigned c) const {
        return c < 4 ? cap[4 * i + c] : data[(size_t)(c - 4) * col_stride + i];
    }
};

// One thread per leaf.
#ifndef QP_LEAF_MIN_BLOCKS
#define QP_LEAF_MIN_BLOCKS 1
#endif
#ifndef QP_LEAF_BLOCK       // threads per block of the leaf kernel
#define QP_LEAF_BLOCK 128
#endif
#ifndef QP_LEAF_PREFETCH   // 1: the next 8-element chunk is fetched into registers during the current permutation
#define QP_LEAF_PREFETCH 1
#endif
#ifndef QP_LEAF_SYNC        // 1: barrier per Poseidon round (all warps of the block in lockstep)
#define QP_LEAF_SYNC 0
#endif
// A leaf's sponge can be advanced in pieces: chunks [chunk_first, chunk_first + chunk_count) of
// 8 elements are absorbed in this launch; the 12-word sponge state of every leaf travels between
// launches in `state` ([12][n_leaves], column-major so that loads coalesce).  The commit from host
// memory uses this to hash the columns that have already arrived while later ones are still
// crossing PCIe.  chunk_first = 0 starts from the initial state (hashing.rs:156-158); the launch
// that absorbs the last chunk writes the digest.  (0, all, nullptr) is the whole hash in one go.
template <class Layout>
__global__ void __launch_bounds__(QP_LEAF_BLOCK, QP_LEAF_MIN_BLOCKS)
leaf_hash_kernel(Layout lay, unsigned leaf_len, TreeShape sh, uint64_t* __restrict__ digests,
                 uint64_t* __restrict__ cap, unsigned chunk_first, unsigned chunk_count,
                 uint64_t* __restrict__ state, size_t leaf_first = 0, size_t leaf_count = ~(size_t)0) {
    const size_t n_leaves = (size_t)1 << sh.lg_leaves;
    // this launch covers leaves [leaf_first, leaf_first + leaf_count) (rows that arrive over time)
    const size_t leaf_end = leaf_count > n_leaves - leaf_first ? n_leaves : leaf_first + leaf_count;
    const size_t i_raw = leaf_first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // threads past the end hash the last leaf again and drop the result, so that every thread
    // of the block reaches the same barriers
    const bool live = i_raw < leaf_end;
    const size_t i = live ? i_raw : leaf_end - 1;
    // hashing.rs:160-163: one permutation per 8-chunk, the last chunk may be short (its missing
    // lanes keep the previous state).  Software pipeline: fetch chunk ch+1 while permuting ch.
    const unsigned n_chunks = (leaf_len + 7) / 8;
    const unsigned ch_end = (chunk_count > n_chunks - chunk_first) ? n_chunks : chunk_first + chunk_count;
    uint64_t s[12];
    if (chunk_first == 0) {
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = 0;
        s[8] = (uint64_t)leaf_len + 1;
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = state[(size_t)k * n_leaves + i];
    }
#if QP_LEAF_PREFETCH
    uint64_t nxt[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (chunk_first * 8 + k < leaf_len) nxt[k] = lay.get(i, chunk_first * 8 + k);
#pragma unroll 1
    for (unsigned ch = chunk_first; ch < ch_end; ch++) {
        const unsigned c = ch * 8;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c + k < leaf_len) s[k] = nxt[k];
        if (ch + 1 < ch_end) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (c + 8 + k < leaf_len) nxt[k] = lay.get(i, c + 8 + k);
        }
        poseidon::permute<QP_LEAF_SYNC != 0>(s);
    }
#else
    // no register prefetch: 16 registers fewer, the load latency is left to the other resident warps
#pragma unroll 1
    for (unsigned ch = chunk_first; ch < ch_end; ch++) {
        const unsigned c = ch * 8;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c + k < leaf_len) s[k] = lay.get(i, c + k);
        poseidon::permute<QP_LEAF_SYNC != 0>(s);
    }
#endif
    if (!live) return;
    if (ch_end == n_chunks) {
        store_digest(leaf_digest_ptr(sh, digests, cap, i), s);
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) state[(size_t)k * n_leaves + i] = s[k];
    }
}

// One thread per node of `layer` (1 <= layer <= num_layers): two_to_one of its children.
__global__ void __launch_bounds__(128)
tree_level_kernel(TreeShape sh, unsigned layer, uint64_t* __restrict__ digests,
                  uint64_t* __restrict__ cap) {
    const unsigned nl = sh.num_layers();
    const size_t per_sub = (size_t)1 << (nl - layer);
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (per_sub << sh.cap_height)) return;
    const size_t t = idx >> (nl - layer), k = idx & (per_sub - 1);
    // children = the sibling pair k of layer-1: adjacent, left first
    const uint64_t* ch = digests + 4 * digest_slot(sh, layer - 1, t, 2 * k);
    uint64_t s[12];
    const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(ch);
    ulonglong2 v0 = c2[0], v1 = c2[1], v2 = c2[2], v3 = c2[3];
    s[0] = v0.x; s[1] = v0.y; s[2] = v1.x; s[3] = v1.y;
    s[4] = v2.x; s[5] = v2.y; s[6] = v3.x; s[7] = v3.y;
    s[8] = s[9] = s[10] = s[11] = 0;
    poseidon::permute(s);
    uint64_t* out = (layer == nl) ? cap + 4 * t : digests + 4 * digest_slot(sh, layer, t, k);
    store_digest(out, s);
}

// ---- the upper levels: one permutation per 16 lanes ----------------------------------------------
// A level with fewer nodes than the GPU has threads costs the LATENCY of one per-thread permutation
// (16 k dependent-ish instructions, ~28 us on B200) whatever its size, and a 2^15-leaf tree -- the
// size recursion proofs commit to -- has eleven such levels.  Here lane l of a 16-lane group holds
// state element l (lanes 12..15 idle): the S-box is one x^7 per lane, the MDS layer eleven shuffles
// and 24 small multiply-adds per lane (the matrix is circulant: out_l = sum_i C_i s_{(l+i) mod 12},
// + 8 s_0 in lane 0; core/src/poseidon.rs:178-198), about 3.8 k instructions per permutation on the
// critical path instead of 16 k -- the reference's poseidon_naive round structure
// (core/src/poseidon.rs:613-633), which the KATs pin to the same function.  It does several times the
// total work of the per-thread kernel, so it is used only where that one is latency-bound: measured
// on B200 (tools/bench_tree_top.py, profiles/r01r_tree_top_threshold.txt) the crossover is at about
// 2^11 nodes per level (MerkleTree::new on 2^15 short leaves: 333 -> 224 us; 2^11 leaves: 223 -> 133 us).
// One block = 16 groups = 16 adjacent nodes of `layer`; it then walks up to `up` further levels
// (8, 4, 2, 1 nodes) inside the block, a barrier between levels.
constexpr int TOP_BLOCK = 256, TOP_GROUPS = TOP_BLOCK / 16, TOP_MAX_UP = 4;

__device__ __forceinline__ uint64_t shfl16(uint64_t v, int src) {
    uint32_t lo, hi;
    gl::unpack(v, lo, hi);
    lo = __shfl_sync(0xffffffffu, lo, src, 16);
    hi = __shfl_sync(0xffffffffu, hi, src, 16);
    return gl::pack(lo, hi);
}

// lane l's element of the permuted state; `rc` = ALL_ROUND_CONSTANTS in shared memory
__device__ __forceinline__ uint64_t permute_lanes(uint64_t s, int l, const uint64_t* __restrict__ rc) {
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};  // poseidon_goldilocks.rs:24
    const int lm = l < 12 ? l : 0;
#pragma unroll 1
    for (int r = 0; r < poseidon::N_ROUNDS; r++) {
        const uint64_t t = gl::add(s, rc[12 * r + lm]);              // constant_layer
        const bool full = r < 4 || r >= 26;
        s = (full || l == 0) ? gl::pow7(t) : t;                      // sbox_layer / sbox on lane 0
        uint32_t x0, x1;
        gl::unpack(s, x0, x1);
        uint64_t lo = l == 0 ? (uint64_t)x0 * 8 : 0, hi = l == 0 ? (uint64_t)x1 * 8 : 0;  // MDS_MATRIX_DIAG
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const int src = lm + i < 12 ? lm + i : lm + i - 12;
            const uint32_t y0 = __shfl_sync(0xffffffffu, x0, src, 16), y1 = __shfl_sync(0xffffffffu, x1, src, 16);
            lo += (uint64_t)y0 * C[i];
            hi += (uint64_t)y1 * C[i];
        }
        uint32_t w0, w1, v0, v1;
        gl::unpack(lo, w0, w1);
        gl::unpack(hi, v0, v1);
        s = gl::fold3(w0, w1, v0, v1);                               // lo + hi 2^32 mod p
    }
    return s;
}

__global__ void __launch_bounds__(TOP_BLOCK)
tree_top_kernel(TreeShape sh, unsigned layer, unsigned up, uint64_t* __restrict__ digests,
                uint64_t* __restrict__ cap) {
    __shared__ uint64_t rc[12 * poseidon::N_ROUNDS];
    for (int i = threadIdx.x; i < 12 * poseidon::N_ROUNDS; i += TOP_BLOCK) rc[i] = poseidon::c_rc[i];
    __syncthreads();
    const unsigned nl = sh.num_layers();
    const int l = threadIdx.x & 15, g = threadIdx.x >> 4;
    for (unsigned j = 0; j <= up; j++) {
        const unsigned ly = layer + j;
        const size_t per_sub = (size_t)1 << (nl - ly);
        const size_t idx = (((size_t)blockIdx.x * TOP_GROUPS) >> j) + g;
        const bool active = g < (TOP_GROUPS >> j) && idx < (per_sub << sh.cap_height);
        const size_t t = idx >> (nl - ly), k = idx & (per_sub - 1);
        // children = the sibling pair k of ly-1: eight adjacent words, left digest first
        uint64_t s = 0;
        if (active && l < 8) s = digests[4 * digest_slot(sh, ly - 1, t, 2 * k) + l];
        s = permute_lanes(s, l, rc);
        if (active && l < 4) {
            uint64_t* out = (ly == nl) ? cap + 4 * t : digests + 4 * digest_slot(sh, ly, t, k);
            out[l] = gl::canon(s);
        }
        __syncthreads();  // the next level of this block reads what this one wrote
    }
}

// Merkle opening: siblings bottom-up (merkle_tree.rs:121-160).  One thread per (query, layer).
__global__ void merkle_paths_kernel(TreeShape sh, const uint64_t* __restrict__ digests,
                                    const uint64_t* __restrict__ leaf_indices, unsigned n_queries,
                                    uint64_t* __restrict__ out /* [n_queries][num_layers][4] */) {
    const unsigned nl = sh.num_layers();
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_queries * nl) return;
    const unsigned qy = id / nl, layer = id % nl;
    const size_t leaf = leaf_indices[qy];
    const size_t t = leaf >> nl, k = (leaf & (((size_t)1 << nl) - 1)) >> layer;
    const uint64_t* srcp = digests + 4 * digest_slot(sh, layer, t, k ^ 1);
    uint64_t* dst = out + 4 * ((size_t)qy * nl + layer);
#pragma unroll
    for (int e = 0; e < 4; e++) dst[e] = srcp[e];
}

// BatchMerkleTree stage leaves: out[i] = cap[i] (4 words) || rows[i] (width words)
// (batch_merkle_tree.rs:91-100).
__global__ void concat_cap_rows_kernel(const uint64_t* __restrict__ cap, const uint64_t* __restrict__ rows,
                                       size_t n_rows, size_t width, uint64_t* __restrict__ out) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t w4 = width + 4;
    if (id >= n_rows * w4) return;
    const size_t i = id / w4, c = id % w4;
    out[id] = c < 4 ? cap[4 * i + c] : rows[i * width + (c - 4)];
}

// Gather whole leaves (rows): out[q][c] = element c of leaf idx[q] (canonical).
template <class Layout>
__global__ void gather_rows_kernel(Layout lay, unsigned leaf_len,
                                   const uint64_t* __restrict__ leaf_indices, size_t first,
                                   size_t n_rows, uint64_t* __restrict__ out) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_rows * leaf_len) return;
    const size_t qy = id / leaf_len;
    const unsigned c = (unsigned)(id % leaf_len);
    const size_t leaf = leaf_indices ? (size_t)leaf_indices[qy] : first + qy;
    out[id] = gl::canon(lay.get(leaf, c));
}


}  // namespace merkle
