// Cycles per modular multiplication on B200 for the forms used by the S-box, in clean loops of 12 independent
// chains per thread (the ILP of an S-box layer) at 4 and 5 warps per SM sub-partition:
//   mul    gl::mul      compiler 128-bit product + reduce128
//   mulhv  gl::mul_hv   four MUL.WIDE + carry chain + reduce128
//   sqrhv  gl::sqr_hv   three MUL.WIDE
//   pow7   gl::pow7     (two squarings, two multiplications: a dependent chain of three)
//   red    reduce128 alone on a synthetic 128-bit value
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o modmul_rate.bin modmul_rate.cu
#include <cstdio>
#include <cstdlib>

#include "../../qp-plonky2_b200/csrc/goldilocks.cuh"

constexpr int ITER = 512, NCH = 12;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(uint64_t* out, uint64_t seed) {
    uint64_t x[NCH], y[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        x[i] = (seed + threadIdx.x + 131 * i + blockIdx.x) * 0x9E3779B97F4A7C15ULL;
        y[i] = x[i] ^ 0x1234567890abcdefULL;
    }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) x[i] = gl::mul(x[i], y[i]);
            if (MODE == 1) x[i] = gl::mul_hv(x[i], y[i]);
            if (MODE == 2) x[i] = gl::sqr_hv(x[i]);
            if (MODE == 3) x[i] = gl::pow7(x[i]);
            if (MODE == 4) x[i] = gl::reduce128(x[i], y[i]) ^ y[i];
            if (MODE == 5) {   // product only (no reduction): keeps the low half, folds the high half in with a xor
                uint64_t lo, hi;
                gl::mul_wide(x[i], y[i], lo, hi);
                x[i] = lo ^ hi;
            }
            if (MODE == 6) x[i] = gl::sqr(x[i]);
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) acc ^= x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
static void run(const char* name, K kern, int warps_per_smsp, double mults_per_iter) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint64_t* out;
    const int blocks = sms * warps_per_smsp;   // 128-thread blocks: one warp per sub-partition each
    cudaMalloc(&out, (size_t)blocks * 128 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, 128>>>(out, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<blocks, 128>>>(out, 2 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    const double cycles = best * 1e-3 * khz * 1e3;
    printf("%-6s %d warps/SMSP  %8.3f ms  %6.2f cycles per operation per warp (SMSP time)\n", name, warps_per_smsp, best,
           cycles / (warps_per_smsp * (double)ITER * NCH * mults_per_iter));
}

int main() {
    for (int w : {4, 5, 8}) {
        run("mul", k<0>, w, 1);
        run("mulhv", k<1>, w, 1);
        run("sqrhv", k<2>, w, 1);
        run("sqr", k<6>, w, 1);
        run("pow7", k<3>, w, 4);
        run("red", k<4>, w, 1);
        run("prod", k<5>, w, 1);
    }
    return 0;
}
