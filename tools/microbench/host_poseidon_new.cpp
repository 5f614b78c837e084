#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../../qp-plonky2_b200/csrc/poseidon_constants.h"
typedef unsigned __int128 u128;
constexpr uint64_t P = 0xFFFFFFFF00000001ULL, EPS=0xFFFFFFFFULL;
inline uint64_t reduce128(u128 x) {
    const uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t0 = lo - hi_hi; if (lo < hi_hi) t0 -= EPS;
    const uint64_t t1 = hi_lo * EPS; uint64_t r = t0 + t1; if (r < t1) r += EPS;
    return r >= P ? r - P : r;
}
inline uint64_t fmul(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }
inline uint64_t fadd(uint64_t a, uint64_t b) { const uint64_t s = a + b; return (s < a || s >= P) ? s - P : s; }
// lo + hi 2^32 with lo, hi < 2^42 -> mod p  (2^64 = EPS)
inline uint64_t fold(uint64_t lo, uint64_t hi) {
    // value = lo + (hi & EPS) 2^32 + (hi >> 32) 2^64
    const u128 v = (u128)lo + ((u128)(hi & EPS) << 32) + (u128)(hi >> 32) * EPS;
    const uint64_t l = (uint64_t)v, h = (uint64_t)(v >> 64);  // h <= 1
    uint64_t r = l + h * EPS; if (r < l) r += EPS;
    return r >= P ? r - P : r;
}
__attribute__((target_clones("avx512f","avx2","default")))
void mds(uint64_t s[12]) {
    uint32_t lo[24], hi[24];
    uint64_t al[12], ah[12];
    for (int i = 0; i < 12; i++) { lo[i] = lo[i + 12] = (uint32_t)s[i]; hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32); }
    for (int r = 0; r < 12; r++) { al[r] = 0; ah[r] = 0; }
    for (int i = 0; i < 12; i++) {
        const uint32_t c = (uint32_t)POSEIDON_MDS_CIRC[i];
        for (int r = 0; r < 12; r++) { al[r] += (uint64_t)lo[r + i] * c; ah[r] += (uint64_t)hi[r + i] * c; }
    }
    al[0] += lo[0] * POSEIDON_MDS_DIAG[0]; ah[0] += hi[0] * POSEIDON_MDS_DIAG[0];
    for (int r = 0; r < 12; r++) s[r] = fold(al[r], ah[r]);
}
void host_permute(uint64_t s[12]) {
    for (int i = 0; i < 12; i++) s[i] %= P;
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = fadd(s[i], POSEIDON_ALL_ROUND_CONSTANTS[12 * r + i]);
        const int lanes = (r < 4 || r >= 26) ? 12 : 1;
        for (int i = 0; i < lanes; i++) { const uint64_t x = s[i], x2 = fmul(x, x), x4 = fmul(x2, x2); s[i] = fmul(fmul(x, x2), x4); }
        mds(s);
    }
}
int main(){ uint64_t s[12]={1,2,3,4,5,6,7,8,9,10,11,12};
 auto t0=std::chrono::steady_clock::now(); for(int k=0;k<20000;k++) host_permute(s); auto t1=std::chrono::steady_clock::now();
 printf("%f us/perm %llu\n", std::chrono::duration<double,std::micro>(t1-t0).count()/20000, (unsigned long long)s[0]); }
