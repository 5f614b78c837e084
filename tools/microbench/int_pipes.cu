// Integer-pipe throughput microbenchmark for sm_100a (B200): cycles per warp instruction per SMSP.
// Each kernel runs NCHAIN independent dependency chains per thread so latency is hidden, with
// enough warps per SM (32) to saturate the pipe.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_pipes tools/microbench/int_pipes.cu && /tmp/int_pipes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, NCHAIN = 8;

#define KERNEL(NAME, BODY)                                                         \
    __global__ void NAME(uint32_t* out, uint32_t seed) {                           \
        uint32_t a[NCHAIN];                                                        \
        uint64_t w[NCHAIN];                                                        \
        uint32_t b = seed | 3u, c = seed * 7u + 1u;                                \
        for (int i = 0; i < NCHAIN; i++) { a[i] = threadIdx.x + i + seed; w[i] = a[i]; } \
        for (int it = 0; it < ITER; it++) {                                        \
            _Pragma("unroll") for (int i = 0; i < NCHAIN; i++) { BODY }            \
        }                                                                          \
        uint32_t s = 0;                                                            \
        for (int i = 0; i < NCHAIN; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32); \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                            \
    }

KERNEL(k_imad_wide, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));)
KERNEL(k_imad_wide_imm, asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(w[i]) : "r"(a[i]));)
KERNEL(k_imad_lo, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_imad_hi, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_dp4a, asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x3215;" : "+r"(a[i]) : "r"(b));)
KERNEL(k_mix_wide_add, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
KERNEL(k_mix_wide_2add, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_mix_lo_add, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
// independent ALU work (no register shared with the wide chain)
KERNEL(k_mix_wide_add_indep, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
KERNEL(k_mix_2wide_add, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c)); asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(w[i]) : "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
KERNEL(k_mix_wide_imad, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_mulwide_add, { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b)); a[i] = (uint32_t)t ^ (uint32_t)(t >> 32); } asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
KERNEL(k_mix_imad_2add, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_mix_dp4a_add, asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));)
KERNEL(k_mix_dp4a_lop, asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));)
KERNEL(k_add64, asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 3));)
KERNEL(k_ffma, { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[i] = __float_as_uint(f); })

template <class K>
void run(const char* name, K k, int instr_per_body) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t* out;
    const int blocks = sms * 4, threads = 256;  // 32 warps per SM
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<blocks, threads>>>(out, 12345);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(out, 12345 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // warp instructions per SMSP = warps/SMSP * ITER * NCHAIN * instr_per_body
    const double warp_instr = 8.0 * ITER * NCHAIN * instr_per_body;
    const double cycles = best * 1e-3 * (double)khz * 1e3;
    printf("%-18s %8.3f ms  %6.2f cycles per warp-instruction per SMSP  (%.1f lanes/clk/SM)\n", name, best,
           cycles / warp_instr, 128.0 / (cycles / warp_instr));
    cudaFree(out);
}

int main() {
    run("IMAD.WIDE reg", k_imad_wide, 1);
    run("IMAD.WIDE imm", k_imad_wide_imm, 1);
    run("IMAD (lo)", k_imad_lo, 1);
    run("IMAD.HI", k_imad_hi, 1);
    run("IDP.4A (dp4a)", k_dp4a, 1);
    run("IADD3", k_iadd3, 1);
    run("LOP3", k_lop3, 1);
    run("SHF", k_shf, 1);
    run("PRMT", k_prmt, 1);
    run("FFMA", k_ffma, 1);
    run("WIDE+IADD", k_mix_wide_add, 2);
    run("WIDE+IADD+LOP", k_mix_wide_2add, 3);
    run("IMAD+IADD", k_mix_lo_add, 2);
    run("WIDE+IADD indep", k_mix_wide_add_indep, 2);
    run("2WIDE+IADD", k_mix_2wide_add, 3);
    run("WIDE+IMAD", k_mix_wide_imad, 2);
    run("MULWIDE+xor+IADD", k_mulwide_add, 3);
    run("IMAD+IADD+LOP", k_mix_imad_2add, 3);
    run("DP4A+IADD", k_mix_dp4a_add, 2);
    run("DP4A+LOP", k_mix_dp4a_lop, 2);
    run("ADD64 (2 instr)", k_add64, 2);
    return 0;
}
