// Poseidon-Goldilocks permutation (width 12, 4 + 22 + 4 rounds, x^7) for sm_100a, one
// permutation per thread with the whole state in registers.
//
// Reference semantics: core/src/poseidon.rs:599-633 (poseidon / poseidon_naive -- the two are
// the same function; the KATs at core/src/poseidon_goldilocks.rs:455-490 pin it).
//
// B200 formulation (not the reference's): the reference speeds up the 22 partial rounds with
// the FAST_PARTIAL_* sparse-matrix tables, whose entries are full 64-bit constants (23 full
// modular multiplies per round).  On the GPU integer pipe a dense multiply by the MDS matrix is
// CHEAPER than that, because every MDS coefficient is < 2^6: the state is split into 32-bit
// halves and each row is 2 x 12 IMAD.WIDE.U32 accumulations (no carries: 12 * 41 * 2^32 < 2^42)
// followed by ONE 96-bit reduction.  So all 30 rounds use the same naive round function
//      state <- MDS * sbox(state)            (sbox on lane 0 only in partial rounds)
// and the next round's constants ride in as the initial value of the row accumulators, which
// makes the constant layer free.  Exact arithmetic => identical outputs to the reference.
#pragma once
#include "goldilocks.cuh"
#include "poseidon_constants.h"

namespace poseidon {

static constexpr int WIDTH = 12;
static constexpr int RATE = 8;
static constexpr int N_ROUNDS = 30;

// ALL_ROUND_CONSTANTS (core/src/poseidon.rs:57-155), uploaded once per context.
__constant__ uint64_t c_round_constants[WIDTH * N_ROUNDS];


// One MDS row (core/src/poseidon.rs:178-198: sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r]) plus an
// additive 64-bit constant `rc` (the NEXT round's constant for this lane, canonical).
template <int R>
__device__ __forceinline__ uint64_t mds_row(const uint32_t (&lo)[12], const uint32_t (&hi)[12],
                                            uint64_t rc) {
    constexpr uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t rc0, rc1;
    gl::unpack(rc, rc0, rc1);
    uint64_t al = rc0, ah = rc1;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const int k = (i + R) % 12;
        uint32_t c = C[i] + ((R == 0 && i == 0) ? 8u : 0u);  // DIAG = [8, 0, ..., 0]
        al += (uint64_t)lo[k] * c;
        ah += (uint64_t)hi[k] * c;
    }
    // value = al + ah * 2^32, al, ah < 2^43:  (s2 : s1 : s0) then reduce96.
    uint32_t al0, al1, ah0, ah1, s1, s2;
    gl::unpack(al, al0, al1);
    gl::unpack(ah, ah0, ah1);
    asm("{\n\t"
        "add.cc.u32  %0, %2, %3;\n\t"
        "addc.u32    %1, %4, 0;\n\t"
        "}"
        : "=&r"(s1), "=&r"(s2)
        : "r"(al1), "r"(ah0), "r"(ah1));
    return gl::reduce96(gl::pack(al0, s1), s2);
}

// state <- MDS(state) + rc[0..12]   (rc may be nullptr => no constants)
__device__ __forceinline__ void mds_layer(uint64_t (&s)[12], const uint64_t* rc) {
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) gl::unpack(s[i], lo[i], hi[i]);
#define QP_ROW(R) s[R] = mds_row<R>(lo, hi, rc ? rc[R] : 0ULL);
    QP_ROW(0) QP_ROW(1) QP_ROW(2) QP_ROW(3) QP_ROW(4) QP_ROW(5)
    QP_ROW(6) QP_ROW(7) QP_ROW(8) QP_ROW(9) QP_ROW(10) QP_ROW(11)
#undef QP_ROW
}

// The permutation.  State lanes may be any u64 representatives; outputs likewise.
__device__ __forceinline__ void permute(uint64_t (&s)[12]) {
    // round 0 constant layer (poseidon.rs:504-513); later rounds get theirs from mds_layer
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::add1(s[i], c_round_constants[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl::pow7(s[i]);
        mds_layer(s, c_round_constants + 12 * (r + 1));
    }
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s[0] = gl::pow7(s[0]);
        mds_layer(s, c_round_constants + 12 * (r + 1));
    }
#pragma unroll 1
    for (int r = 26; r < 29; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl::pow7(s[i]);
        mds_layer(s, c_round_constants + 12 * (r + 1));
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl::pow7(s[i]);
    mds_layer(s, nullptr);
}


}  // namespace poseidon
