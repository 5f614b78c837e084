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
void host_permute(uint64_t s[12]) {
    for (int i = 0; i < 12; i++) s[i] %= P;
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = fadd(s[i], POSEIDON_ALL_ROUND_CONSTANTS[12 * r + i]);
        const int lanes = (r < 4 || r >= 26) ? 12 : 1;
        for (int i = 0; i < lanes; i++) { const uint64_t x = s[i], x2 = fmul(x, x), x4 = fmul(x2, x2); s[i] = fmul(fmul(x, x2), x4); }
        uint64_t t[12];
        for (int row = 0; row < 12; row++) {
            u128 acc = (u128)s[row] * POSEIDON_MDS_DIAG[row];
            for (int i = 0; i < 12; i++) acc += (u128)s[(i + row) % 12] * POSEIDON_MDS_CIRC[i];
            t[row] = reduce128(acc);
        }
        std::memcpy(s, t, sizeof t);
    }
}
int main(){ uint64_t s[12]={1,2,3,4,5,6,7,8,9,10,11,12};
 auto t0=std::chrono::steady_clock::now(); for(int k=0;k<20000;k++) host_permute(s); auto t1=std::chrono::steady_clock::now();
 printf("%f us/perm %llu\n", std::chrono::duration<double,std::micro>(t1-t0).count()/20000, (unsigned long long)s[0]); }
