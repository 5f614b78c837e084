// Integer / FP64 pipe throughput microbenchmark for sm_100a (B200): cycles per warp instruction per
// SMSP.  Each kernel runs NCHAIN independent dependency chains per thread (latency hidden) with 32
// warps per SM.  Every multiply takes an operand from its own chain, so ptxas cannot hoist the
// product out of the loop (the first version of this file measured IADD3 chains by accident:
// check the loop body with `cuobjdump -sass` before believing a number).  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/int_pipes tools/microbench/int_pipes.cu && /tmp/int_pipes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048, NCHAIN = 8;

#define KERNEL(NAME, BODY)                                                                         \
    __global__ void NAME(uint32_t* out, uint32_t seed) {                                           \
        uint32_t a[NCHAIN];                                                                        \
        uint64_t w[NCHAIN];                                                                        \
        double d[NCHAIN];                                                                          \
        uint32_t b = seed | 3u, c = seed * 7u + 1u;                                                \
        double db = 1.0 + seed * 1e-9, dc = 0.5 + seed * 1e-7;                                     \
        for (int i = 0; i < NCHAIN; i++) {                                                         \
            a[i] = threadIdx.x + i + seed;                                                         \
            w[i] = a[i] * 0x100000001ull;                                                          \
            d[i] = (double)a[i];                                                                   \
        }                                                                                          \
        _Pragma("unroll 4") for (int it = 0; it < ITER; it++) {                                    \
            _Pragma("unroll") for (int i = 0; i < NCHAIN; i++) { BODY }                            \
        }                                                                                          \
        uint32_t s = 0;                                                                            \
        for (int i = 0; i < NCHAIN; i++)                                                           \
            s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)__double2loint(d[i]) + \
                 (uint32_t)__double2hiint(d[i]);                                                   \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                            \
    }

#define LO(x) ((uint32_t)(x))
#define HI(x) ((uint32_t)((x) >> 32))
// w = lo(w) * b + w : dependent wide multiply-accumulate
#define WIDE asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[i])), "r"(b));
#define WIDE_IMM asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(w[i]) : "r"(HI(w[i])));
#define MULWIDE asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(LO(w[i])), "r"(HI(w[i])));
#define IMADLO asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define IMADHI asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define IADD asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
#define LOP asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
#define SHF asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
#define DFMA asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(db), "d"(dc));
#define DADD asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dc));
#define DMUL asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(db));
#define FFMA                                                                                     \
    {                                                                                            \
        float f = __uint_as_float(a[i]);                                                         \
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));          \
        a[i] = __float_as_uint(f);                                                               \
    }
// 64-bit add with carry (2 instructions)
#define ADD64 asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b << 3));

#define I2F                                                                                      \
    {                                                                                            \
        double t_;                                                                               \
        asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t_) : "r"(a[i]));                           \
        a[i] ^= (uint32_t)__double2hiint(t_);                                                    \
    }
KERNEL(k_i2f, I2F)
KERNEL(k_mulwide_lop, MULWIDE LOP)
KERNEL(k_mulwide_lop_shf, MULWIDE LOP SHF)
KERNEL(k_mulwide_iadd, MULWIDE IADD)
KERNEL(k_mulwide_dfma, MULWIDE DFMA)
KERNEL(k_mulwide_2dfma, MULWIDE DFMA DFMA)
KERNEL(k_mulwide_imadlo, MULWIDE IMADLO)
KERNEL(k_mulwide_ffma, MULWIDE FFMA)
KERNEL(k_i2f_dfma, I2F DFMA)
KERNEL(k_wide, WIDE)
KERNEL(k_wide_imm, WIDE_IMM)
KERNEL(k_mulwide, MULWIDE)
KERNEL(k_imad_lo, IMADLO)
KERNEL(k_imad_hi, IMADHI)
KERNEL(k_iadd, IADD)
KERNEL(k_lop, LOP)
KERNEL(k_shf, SHF)
KERNEL(k_ffma, FFMA)
KERNEL(k_dfma, DFMA)
KERNEL(k_dadd, DADD)
KERNEL(k_dmul, DMUL)
KERNEL(k_add64, ADD64)
KERNEL(k_wide_lop, WIDE LOP)
KERNEL(k_wide_2lop, WIDE LOP SHF)
KERNEL(k_wide_iadd, WIDE IADD)
KERNEL(k_wide_imadlo, WIDE IMADLO)
KERNEL(k_wide_dfma, WIDE DFMA)
KERNEL(k_wide_2dfma, WIDE DFMA DFMA)
KERNEL(k_wide_dfma_lop, WIDE DFMA LOP)
KERNEL(k_wide_2dfma_2lop, WIDE DFMA DFMA LOP SHF)
KERNEL(k_dfma_lop, DFMA LOP)
KERNEL(k_dfma_ffma, DFMA FFMA)
KERNEL(k_imadlo_lop, IMADLO LOP)
KERNEL(k_imadlo_dfma, IMADLO DFMA)
KERNEL(k_ffma_lop, FFMA LOP)
KERNEL(k_ffma_wide, FFMA WIDE)

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
    printf("%-22s %8.3f ms  %6.2f cycles per body per warp per SMSP (%d instr: %5.2f each)\n", name, best,
           cycles / warp_instr * instr_per_body, instr_per_body, cycles / warp_instr);
    cudaFree(out);
}

int main() {
    run("I2F.F64.U32 (+LOP)", k_i2f, 2);
    run("I2F+LOP+DFMA", k_i2f_dfma, 3);
    run("MULWIDE+LOP", k_mulwide_lop, 2);
    run("MULWIDE+LOP+SHF", k_mulwide_lop_shf, 3);
    run("MULWIDE+IADD", k_mulwide_iadd, 2);
    run("MULWIDE+DFMA", k_mulwide_dfma, 2);
    run("MULWIDE+2DFMA", k_mulwide_2dfma, 3);
    run("MULWIDE+IMADLO", k_mulwide_imadlo, 2);
    run("MULWIDE+FFMA", k_mulwide_ffma, 2);
    run("IMAD.WIDE reg", k_wide, 1);
    run("IMAD.WIDE imm", k_wide_imm, 1);
    run("MUL.WIDE", k_mulwide, 1);
    run("IMAD (lo)", k_imad_lo, 1);
    run("IMAD.HI", k_imad_hi, 1);
    run("IADD3", k_iadd, 1);
    run("LOP3", k_lop, 1);
    run("SHF", k_shf, 1);
    run("FFMA", k_ffma, 1);
    run("DFMA", k_dfma, 1);
    run("DADD", k_dadd, 1);
    run("DMUL", k_dmul, 1);
    run("ADD64 (2 instr)", k_add64, 2);
    run("WIDE+LOP", k_wide_lop, 2);
    run("WIDE+LOP+SHF", k_wide_2lop, 3);
    run("WIDE+IADD", k_wide_iadd, 2);
    run("WIDE+IMADLO", k_wide_imadlo, 2);
    run("WIDE+DFMA", k_wide_dfma, 2);
    run("WIDE+2DFMA", k_wide_2dfma, 3);
    run("WIDE+DFMA+LOP", k_wide_dfma_lop, 3);
    run("WIDE+2DFMA+LOP+SHF", k_wide_2dfma_2lop, 5);
    run("DFMA+LOP", k_dfma_lop, 2);
    run("DFMA+FFMA", k_dfma_ffma, 2);
    run("IMADLO+LOP", k_imadlo_lop, 2);
    run("IMADLO+DFMA", k_imadlo_dfma, 2);
    run("FFMA+LOP", k_ffma_lop, 2);
    run("FFMA+WIDE", k_ffma_wide, 2);
    return 0;
}
