// Goldilocks field arithmetic for sm_100a, p = 2^64 - 2^32 + 1.
//
// Values are raw u64 and may be NON-canonical (any representative in [0, 2^64)), exactly as
// the reference carries them (field/src/goldilocks_field.rs:33-37); canonicalise only where
// bytes leave the device path (digests, LDE rows, caps).  All routines are total over u64
// inputs.  The multiply is 4x IMAD.WIDE.U32 (fma pipe) and the reduction is
//      x = x0 + x1 b + x2 b^2 + x3 b^3,  b = 2^32,  b^2 = b - 1,  b^3 = -1  (mod p)
//        = (x1:x0) - x3 + x2*(2^32-1)
// (reference reduce128, goldilocks_field.rs:390-403), done with carry-chained 32-bit PTX so
// no 64-bit compare/select sequences are generated.
#pragma once
#include <cstdint>

namespace gl {

static constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
static constexpr uint64_t EPS = 0xFFFFFFFFULL;
// field/src/goldilocks_field.rs:84,91
static constexpr uint64_t GENERATOR = 14293326489335486720ULL;
static constexpr uint64_t POWER_OF_TWO_GENERATOR = 7277203076849721926ULL;

__host__ __device__ __forceinline__ uint64_t canon(uint64_t a) { return a >= P ? a - P : a; }


__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack(uint64_t v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}

// ---- add / sub ---------------------------------------------------------------------------
// 32-bit carry chains: ptxas turns each of these into IADD3/IADD3.X pairs with the carry held
// in a predicate (4-5 SASS instructions for the single-correction forms), where the obvious
// 64-bit C (`s = a + b; if (s < a) s += EPS`) costs 8 (compare + select).

// NOTE on carry flags: ptxas keeps CC.CF in "ARM style" -- after add.cc it is the carry, after
// sub.cc it is NOT-borrow, and subc computes a + ~b + CF.  So `subc m, 0, 0` yields the borrow
// mask (0xffffffff on borrow) after a SUBTRACT chain, but the inverse of the carry mask after an
// ADD chain.  After additions the carry is therefore turned into a predicate (addc / setp) and
// the mask selected from it; ptxas folds that into one SEL on the carry predicate.

// a + b (mod p) when at most one wrap can occur, i.e. a + b < 2^64 + p: true whenever one
// operand is <= p.  5 SASS instructions.
__device__ __forceinline__ uint64_t add1(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1, s0, s1;
    unpack(a, a0, a1);
    unpack(b, b0, b1);
    asm("{\n\t"
        ".reg .u32 c, m;\n\t"
        ".reg .pred p;\n\t"
        "add.cc.u32   %0, %2, %4;\n\t"
        "addc.cc.u32  %1, %3, %5;\n\t"
        "addc.u32     c, 0, 0;\n\t"
        "setp.ne.u32  p, c, 0;\n\t"
        "selp.u32     m, 0xffffffff, 0, p;\n\t"   // carry * EPS
        "add.cc.u32   %0, %0, m;\n\t"
        "addc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(s0), "=&r"(s1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(s0, s1);
}

// a + b (mod p), total over u64 x u64 (goldilocks_field.rs:249-265).  A second wrap needs both
// operands non-canonical; after it the sum is < 2^32, so a third cannot occur.
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1, s0, s1;
    unpack(a, a0, a1);
    unpack(b, b0, b1);
    asm("{\n\t"
        ".reg .u32 c, m;\n\t"
        ".reg .pred p;\n\t"
        "add.cc.u32   %0, %2, %4;\n\t"
        "addc.cc.u32  %1, %3, %5;\n\t"
        "addc.u32     c, 0, 0;\n\t"
        "setp.ne.u32  p, c, 0;\n\t"
        "selp.u32     m, 0xffffffff, 0, p;\n\t"
        "add.cc.u32   %0, %0, m;\n\t"
        "addc.cc.u32  %1, %1, 0;\n\t"
        "addc.u32     c, 0, 0;\n\t"
        "setp.ne.u32  p, c, 0;\n\t"
        "selp.u32     m, 0xffffffff, 0, p;\n\t"
        "add.cc.u32   %0, %0, m;\n\t"
        "addc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(s0), "=&r"(s1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(s0, s1);
}

// a - b (mod p) when at most one wrap can occur: b canonical, or a - b > -(2^64 - 2^32).
__device__ __forceinline__ uint64_t sub1(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1, s0, s1;
    unpack(a, a0, a1);
    unpack(b, b0, b1);
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32   %0, %2, %4;\n\t"
        "subc.cc.u32  %1, %3, %5;\n\t"
        "subc.u32     m, 0, 0;\n\t"   // m = borrow ? 0xffffffff : 0
        "sub.cc.u32   %0, %0, m;\n\t"
        "subc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(s0), "=&r"(s1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(s0, s1);
}

// a - b (mod p), total over u64 x u64 (goldilocks_field.rs:280-294).
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) {
    uint32_t a0, a1, b0, b1, s0, s1;
    unpack(a, a0, a1);
    unpack(b, b0, b1);
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32   %0, %2, %4;\n\t"
        "subc.cc.u32  %1, %3, %5;\n\t"
        "subc.u32     m, 0, 0;\n\t"
        "sub.cc.u32   %0, %0, m;\n\t"
        "subc.cc.u32  %1, %1, 0;\n\t"
        "subc.u32     m, 0, 0;\n\t"
        "sub.cc.u32   %0, %0, m;\n\t"
        "subc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(s0), "=&r"(s1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return pack(s0, s1);
}

__device__ __forceinline__ uint64_t neg(uint64_t a) {
    uint64_t c = canon(a);
    return c ? P - c : 0;
}

// ---- reduction / multiplication ----------------------------------------------------------

// (hi:lo), hi 32 bits  ->  lo + hi * EPS with the single possible wrap folded back
// (reduce96, goldilocks_field.rs:381-385).  hi * EPS = (hi << 32) - hi is formed on the ALU pipe
// (2 instructions) instead of one IMAD.WIDE: on B200 the fma-heavy pipe is the scarce one
// (IMAD.WIDE = 4 pipe cycles).  hi * EPS <= (2^32-1)^2 < p, so one correction is enough.
__device__ __forceinline__ uint64_t reduce96(uint64_t lo, uint32_t hi) {
    uint32_t p0, p1;
    asm("{\n\t"
        "sub.cc.u32   %0, 0, %2;\n\t"    // low word:  -hi
        "subc.u32     %1, %2, 0;\n\t"    // high word: hi - (hi != 0)
        "}"
        : "=&r"(p0), "=&r"(p1)
        : "r"(hi));
    return add1(lo, pack(p0, p1));
}

// Reduce the 128-bit value (hi:lo) mod p; output is some u64 representative
// (reduce128, goldilocks_field.rs:390-403):  lo - hi_hi + hi_lo * EPS.
__device__ __forceinline__ uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint32_t x0, x1, x2, x3, t0, t1;
    unpack(lo, x0, x1);
    unpack(hi, x2, x3);
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32   %0, %2, %4;\n\t"
        "subc.cc.u32  %1, %3, 0;\n\t"
        "subc.u32     m, 0, 0;\n\t"
        "sub.cc.u32   %0, %0, m;\n\t"   // borrow only if lo < 2^32, so this cannot borrow again
        "subc.u32     %1, %1, 0;\n\t"
        "}"
        : "=&r"(t0), "=&r"(t1)
        : "r"(x0), "r"(x1), "r"(x3));
    return reduce96(pack(t0, t1), x2);
}

// a * b (mod p) (goldilocks_field.rs:303-310).  The 128-bit product is written as a C
// multiply so ptxas emits its fused 7-instruction IMAD.WIDE.U32 sequence (carry in a predicate).
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return reduce128((uint64_t)p, (uint64_t)(p >> 64));
}

__device__ __forceinline__ uint64_t sqr(uint64_t a) { return mul(a, a); }

// x^7 (core/src/poseidon.rs:546-552)
__device__ __forceinline__ uint64_t pow7(uint64_t x) {
    uint64_t x2 = sqr(x);
    uint64_t x4 = sqr(x2);
    uint64_t x3 = mul(x, x2);
    return mul(x3, x4);
}

__device__ __forceinline__ uint64_t pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, a);
        a = sqr(a);
        e >>= 1;
    }
    return r;
}


// ---- host-side helpers (table setup only; never on a data path) ----
static inline uint64_t host_mul(uint64_t a, uint64_t b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    return (uint64_t)(x % P);
}
static inline uint64_t host_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    a %= P;
    while (e) {
        if (e & 1) r = host_mul(r, a);
        a = host_mul(a, a);
        e >>= 1;
    }
    return r;
}
// field/src/types.rs:280-284
static inline uint64_t host_primitive_root(unsigned k) {
    return host_pow(POWER_OF_TWO_GENERATOR, 1ULL << (32 - k));
}

}  // namespace gl
