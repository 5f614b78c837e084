// Goldilocks field arithmetic for sm_100a, p = 2^64 - 2^32 + 1.
//
// Values are raw u64 and may be NON-canonical (any representative in [0, 2^64)), exactly as
// the reference carries them (field/src/goldilocks_field.rs:33-37); canonicalise only where
// bytes leave the device path (digests, LDE rows, caps).  All routines are total over u64
// inputs.  The multiply is 4x IMAD.WIDE.U32 (fma-heavy pipe) and the reduction is
//      x = x0 + x1 b + x2 b^2 + x3 b^3,  b = 2^32,  b^2 = b - 1,  b^3 = -1  (mod p)
//        = (x0 - x2 - x3) + (x1 + x2) b
// (reference reduce128, goldilocks_field.rs:390-403), one signed multi-word sum and one
// fold of its overflow word: 16 SASS instructions per modular multiplication.
#pragma once
#include <cstdint>

// 1: x^7 uses the explicit 4 x IMAD.WIDE product (mul_hv); measured slower than the compiler's
// fused sequence on B200 (39.8 vs 38.8 ms), so off by default.
#ifndef QP_POSEIDON_EXPLICIT_MUL
#define QP_POSEIDON_EXPLICIT_MUL 0
#endif
// 1: the two squarings of x^7 use three IMAD.WIDE (cross term doubled) instead of the compiler's four
#ifndef QP_POSEIDON_SQR_HV   // measured on B200: leaf hash 25.69 -> 25.44 ms at 2^21 leaves x 135
#define QP_POSEIDON_SQR_HV 1
#endif

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
// Reductions.  With B = 2^32:  B^2 = B - 1 and B^3 = -1 (mod p), so a value given by 32-bit words
//      x0 + x1 B + x2 B^2 + x3 B^3  =  (x0 - x2 - x3) + (x1 + x2) B            (reduce128,
// goldilocks_field.rs:390-403 computes the same thing as lo - hi_hi + hi_lo * EPS).  The right-hand
// side is evaluated as ONE signed multi-word sum u + v B whose overflow word w is in {-1, 0, 1};
// w B^2 = w EPS is folded back once, and the bounds below show that this last addition cannot
// wrap.  Written with 64-bit C arithmetic on purpose: ptxas turns it into 3-input IADD3 with dual
// carry predicates (9 SASS instructions for reduce128, against 13 for the borrow/carry-chain form
// that PTX add.cc/sub.cc can express).

// r + w * EPS for a small signed w:  (r1 + w : r0) - sign_extend(w), as a two-word subtraction
__device__ __forceinline__ uint64_t add_w_eps(uint64_t r, int32_t w) {
    uint32_t r0, r1, lo, hi;
    unpack(r, r0, r1);
    asm("{\n\t"
        ".reg .u32 t, sgn;\n\t"
        "add.u32      t, %3, %4;\n\t"
        "shr.s32      sgn, %4, 31;\n\t"
        "sub.cc.u32   %0, %2, %4;\n\t"
        "subc.u32     %1, t, sgn;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi)
        : "r"(r0), "r"(r1), "r"(w));
    return pack(lo, hi);
}

// (hi:lo), hi 32 bits  ->  lo + hi * EPS  (reduce96, goldilocks_field.rs:381-385).
// u = x0 - hi, v = x1 + hi + (u >> 32) in [0, 2^33 - 2]; w = 1 implies r <= B^2 - B - 1, so r + EPS < B^2.
__device__ __forceinline__ uint64_t reduce96(uint64_t lo, uint32_t hi) {
    uint32_t x0, x1;
    unpack(lo, x0, x1);
    const int64_t u = (int64_t)(uint64_t)x0 - (int64_t)(uint64_t)hi;
    const int64_t v = (int64_t)((uint64_t)x1 + hi) + (u >> 32);
    return add_w_eps(pack((uint32_t)u, (uint32_t)v), (int32_t)(v >> 32));
}

// Reduce the 128-bit value (hi:lo) mod p; output is some u64 representative.
// u in [-(2B - 2), B), v = x1 + x2 + (u >> 32) in [-2, 2B - 2].  w = 1: the value is at most
// 2B^2 - 2B, so r <= B^2 - 2B and r + EPS < B^2.  w = -1: the value is at least -(B - 1), so
// r >= B^2 - B + 1 > EPS.
__device__ __forceinline__ uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint32_t x0, x1, x2, x3;
    unpack(lo, x0, x1);
    unpack(hi, x2, x3);
    const int64_t u = (int64_t)(uint64_t)x0 - (int64_t)(uint64_t)x2 - (int64_t)(uint64_t)x3;
    const int64_t v = (int64_t)((uint64_t)x1 + x2) + (u >> 32);
    return add_w_eps(pack((uint32_t)u, (uint32_t)v), (int32_t)(v >> 32));
}

// w0 + (w1 + v0) B + v1 B^2  ->  field element, for w1 + v1 < 2^32: the sum of two 64-bit
// accumulators (w1:w0) + (v1:v0) B with small high words, as the Poseidon linear layers produce.
// = (w0 - v1) + (w1 + v0 + v1) B; v <= 2B - 2, so w <= 1 and r + EPS < B^2 as in reduce96.
__device__ __forceinline__ uint64_t fold3(uint32_t w0, uint32_t w1, uint32_t v0, uint32_t v1) {
    const int64_t u = (int64_t)(uint64_t)w0 - (int64_t)(uint64_t)v1;
    const int64_t v = (int64_t)((uint64_t)w1 + v0 + v1) + (u >> 32);
    return add_w_eps(pack((uint32_t)u, (uint32_t)v), (int32_t)(v >> 32));
}

// 64 x 64 -> 128 bit product as (lo, hi).  Four IMAD.WIDE.U32 and a 32-bit carry chain that
// ptxas folds into five IADD3/IADD3.X with dual carry predicates.  (Letting the compiler expand
// `(unsigned __int128)a * b` gives 4 IMAD.WIDE + IMAD.WIDE.X + IMAD.X + MOV: 28 cycles of the
// fma-heavy pipe instead of 16 -- that pipe is the bottleneck of the Poseidon kernels.)
__device__ __forceinline__ void mul_wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
    uint32_t a0, a1, b0, b1;
    unpack(a, a0, a1);
    unpack(b, b0, b1);
    uint64_t L, M, N, H;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(L) : "r"(a0), "r"(b0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(M) : "r"(a0), "r"(b1));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(N) : "r"(a1), "r"(b0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(H) : "r"(a1), "r"(b1));
    uint32_t l0, l1, m0, m1, n0, n1, h0, h1, x1, x2, x3;
    unpack(L, l0, l1);
    unpack(M, m0, m1);
    unpack(N, n0, n1);
    unpack(H, h0, h1);
    asm("{\n\t"
        "add.cc.u32  %0, %3, %4;\n\t"
        "addc.cc.u32 %1, %5, %6;\n\t"
        "addc.u32    %2, %7, 0;\n\t"
        "add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.u32    %2, %2, 0;\n\t"
        "}"
        : "=&r"(x1), "=&r"(x2), "=&r"(x3)
        : "r"(l1), "r"(m0), "r"(h0), "r"(m1), "r"(h1), "r"(n0), "r"(n1));
    lo = pack(l0, x1);
    hi = pack(x2, x3);
}

// a^2 as (lo, hi): three IMAD.WIDE.U32 (the cross term is added twice).
__device__ __forceinline__ void sqr_wide(uint64_t a, uint64_t& lo, uint64_t& hi) {
    uint32_t a0, a1;
    unpack(a, a0, a1);
    uint64_t L, M, H;
    asm("mul.wide.u32 %0, %1, %1;" : "=l"(L) : "r"(a0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(M) : "r"(a0), "r"(a1));
    asm("mul.wide.u32 %0, %1, %1;" : "=l"(H) : "r"(a1));
    uint32_t l0, l1, m0, m1, h0, h1, x1, x2, x3;
    unpack(L, l0, l1);
    unpack(M, m0, m1);
    unpack(H, h0, h1);
    asm("{\n\t"
        "add.cc.u32  %0, %3, %4;\n\t"
        "addc.cc.u32 %1, %5, %6;\n\t"
        "addc.u32    %2, %7, 0;\n\t"
        "add.cc.u32  %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.u32    %2, %2, 0;\n\t"
        "}"
        : "=&r"(x1), "=&r"(x2), "=&r"(x3)
        : "r"(l1), "r"(m0), "r"(h0), "r"(m1), "r"(h1));
    lo = pack(l0, x1);
    hi = pack(x2, x3);
}

// a * b (mod p) (goldilocks_field.rs:303-310).  Two forms with the same result:
//   mul    -- compiler-expanded 128-bit product: fewest ALU instructions (the NTT kernels are
//             ALU-pipe bound);
//   mul_hv -- explicit 4 x IMAD.WIDE form: fewest fma-heavy pipe cycles (the Poseidon kernels are
//             bound by that pipe).
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return reduce128((uint64_t)p, (uint64_t)(p >> 64));
}
__device__ __forceinline__ uint64_t mul_hv(uint64_t a, uint64_t b) {
    uint64_t lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
}
__device__ __forceinline__ uint64_t sqr(uint64_t a) { return mul(a, a); }
__device__ __forceinline__ uint64_t sqr_hv(uint64_t a) {
    uint64_t lo, hi;
    sqr_wide(a, lo, hi);
    return reduce128(lo, hi);
}

// x^7 (core/src/poseidon.rs:546-552)
__device__ __forceinline__ uint64_t pow7(uint64_t x) {
#if defined(QP_POW7_VARIANT)
    // measurement variants (tools/microbench): bit 0 / bit 1 = first / second squaring explicit (sqr_hv) or the
    // compiler's; bit 2 = multiplications explicit (mul_hv); bit 3 = the chain x^2, x^3, x^6, x^7
    constexpr int V = QP_POW7_VARIANT;
    auto S1 = [](uint64_t a) { return (V & 1) ? sqr_hv(a) : sqr(a); };
    auto S2 = [](uint64_t a) { return (V & 2) ? sqr_hv(a) : sqr(a); };
    auto M = [](uint64_t a, uint64_t b) { return (V & 4) ? mul_hv(a, b) : mul(a, b); };
    if (V & 8) {
        const uint64_t x2 = S1(x), x3 = M(x2, x), x6 = S2(x3);
        return M(x6, x);
    }
    const uint64_t x2 = S1(x), x4 = S2(x2), x3 = M(x, x2);
    return M(x3, x4);
#elif QP_POSEIDON_SQR_HV
    uint64_t x2 = sqr_hv(x);
    uint64_t x4 = sqr_hv(x2);
    uint64_t x3 = mul(x, x2);
    return mul(x3, x4);
#elif QP_POSEIDON_EXPLICIT_MUL
    uint64_t x2 = sqr_hv(x);
    uint64_t x4 = sqr_hv(x2);
    uint64_t x3 = mul_hv(x, x2);
    return mul_hv(x3, x4);
#else
    uint64_t x2 = sqr(x);
    uint64_t x4 = sqr(x2);
    uint64_t x3 = mul(x, x2);
    return mul(x3, x4);
#endif
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


// ---- lazy (redundant) form for butterfly networks -----------------------------------------------
// Inside a radix-2^R butterfly a value is carried as three 32-bit words  w0 + w1 B + w2 B^2  (B = 2^32)
// whose top word is a SMALL SIGNED integer: addition and subtraction are then plain 3-word carry chains
// (3 IADD3 each, no modular fix-up -- the two-fix gl::add / gl::sub cost 8), and a multiplication by 2^s,
// the only twiddle a radix-16 butterfly needs in this field (w_64 = 8), is a funnel shift followed by
// the fold  B^2 = B - 1, B^3 = -1, B^4 = -B, B^5 = 1 - B  back to two words plus a small carry.
// Bounds: lz_from gives |x| < 2^64 and lz_mul_pow2 gives |x| < 2^66; each lz_add / lz_sub at most doubles
// the bound, so after the <= 4 levels of a radix-16 butterfly |x| < 2^70: w2 stays in [-64, 63], and a
// shift by r <= 31 bits fits the fourth word that lz_mul_pow2 computes.
struct lz {
    uint32_t w0, w1, w2;
};
__device__ __forceinline__ lz lz_from(uint64_t x) {
    lz r;
    unpack(x, r.w0, r.w1);
    r.w2 = 0;
    return r;
}
__device__ __forceinline__ lz lz_add(const lz& a, const lz& b) {
    lz r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=&r"(r.w0), "=&r"(r.w1), "=&r"(r.w2)
        : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
}
__device__ __forceinline__ lz lz_sub(const lz& a, const lz& b) {
    lz r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=&r"(r.w0), "=&r"(r.w1), "=&r"(r.w2)
        : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(b.w0), "r"(b.w1), "r"(b.w2));
    return r;
}
// a * 2^s mod p for a compile-time-foldable 0 < s < 96.  With s = 32 q + r:  Y = a 2^r as four words
// (y3 signed), then  Y B^q = u + v B  with
//   q = 0: u =  y0 - y2 - y3,  v = y1 + y2          q = 1: u = -y1 - y2,  v = y0 + y1 - y3
//   q = 2: u = -y0 - y1 + y3,  v = y0 - y2 - y3
__device__ __forceinline__ lz lz_mul_pow2(const lz& a, int s) {
    const int q = s >> 5, r = s & 31;
    uint32_t y0, y1, y2;
    int32_t y3;
    if (r == 0) {
        y0 = a.w0;
        y1 = a.w1;
        y2 = a.w2;
        y3 = (int32_t)a.w2 >> 31;
    } else {
        y0 = a.w0 << r;
        y1 = __funnelshift_l(a.w0, a.w1, r);
        y2 = __funnelshift_l(a.w1, a.w2, r);
        y3 = (int32_t)a.w2 >> (32 - r);
    }
    int64_t u, v;
    if (q == 0) {
        u = (int64_t)(uint64_t)y0 - (int64_t)(uint64_t)y2 - (int64_t)y3;
        v = (int64_t)((uint64_t)y1 + y2);
    } else if (q == 1) {
        u = -(int64_t)((uint64_t)y1 + y2);
        v = (int64_t)((uint64_t)y0 + y1) - (int64_t)y3;
    } else {
        u = (int64_t)y3 - (int64_t)((uint64_t)y0 + y1);
        v = (int64_t)(uint64_t)y0 - (int64_t)(uint64_t)y2 - (int64_t)y3;
    }
    const int64_t t = v + (u >> 32);
    lz o;
    o.w0 = (uint32_t)u;
    o.w1 = (uint32_t)t;
    o.w2 = (uint32_t)(t >> 32);
    return o;
}
// -> some u64 representative:  w0 + w1 B + w2 (B - 1)  =  (w0 - w2) + (w1 + w2) B, overflow word in {-1, 0, 1}
__device__ __forceinline__ uint64_t lz_reduce(const lz& a) {
    const int64_t w2 = (int64_t)(int32_t)a.w2;
    const int64_t u = (int64_t)(uint64_t)a.w0 - w2;
    const int64_t v = (int64_t)(uint64_t)a.w1 + w2 + (u >> 32);
    return add_w_eps(pack((uint32_t)u, (uint32_t)v), (int32_t)(v >> 32));
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
