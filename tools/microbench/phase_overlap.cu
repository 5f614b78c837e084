// Do the integer (S-box) and FP64 (linear layer) halves of the Poseidon permutation overlap on B200?
// Times, with the occupancy of the leaf-hash kernel (4 and 8 warps per SM sub-partition):
//   S    : all warps run S-box layers only            (12 x gl::pow7 per layer)
//   M    : all warps run full-round linear layers only (poseidon::mds_layer_split)
//   P    : all warps run partial-round pairs only      (poseidon::partial_pair_split)
//   S|M  : even warps run S, odd warps run M  -- perfect overlap would cost max(S, M) / 2 per unit, none (S + M) / 2
//   perm : the whole permutation
// and prints cycles per layer per warp per SM sub-partition, so that
//   8 * (S + M) + 11 * P   (no overlap at all; P contains its two S-boxes)   can be compared with the measured permutation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DQP_...] -o phase_overlap.bin phase_overlap.cu
#include <cstdio>
#include <cstdlib>

#include "../../qp-plonky2_b200/csrc/poseidon.cuh"

constexpr int ITER = 256;

template <int MODE>   // 0 S, 1 M, 2 P, 3 S|M by warp parity, 4 perm, 5 S|P by warp parity
__global__ void __launch_bounds__(256, 1) k(uint64_t* out, uint64_t seed) {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = (seed + threadIdx.x + 131 * i + blockIdx.x) * 0x9E3779B97F4A7C15ULL;
    const bool odd = (threadIdx.x >> 7) & 1;   // warps 0-3 / 4-7 of the block: one of each kind per sub-partition
    if (MODE == 0 || ((MODE == 3 || MODE == 5) && !odd)) {
#pragma unroll 1
        for (int it = 0; it < ITER; it++) poseidon::sbox_all(s);
    } else if (MODE == 1 || (MODE == 3 && odd)) {
#pragma unroll 1
        for (int it = 0; it < ITER; it++) poseidon::mds_layer_split(s, 1 + (it & 3));
    } else if (MODE == 2 || (MODE == 5 && odd)) {
#pragma unroll 1
        for (int it = 0; it < ITER; it++) poseidon::partial_pair_split(s, it % 11);
    } else if (MODE == 6) {
        // two independent states per thread: the S-box layer of one and the linear layer of the other sit in the
        // same basic block, so ptxas is free to interleave MUL.WIDE / IADD3 with DFMA inside one warp
        uint64_t t[12];
#pragma unroll
        for (int i = 0; i < 12; i++) t[i] = s[i] ^ 0x5555555555555555ULL;
#pragma unroll 1
        for (int it = 0; it < ITER / 2; it++) {
            poseidon::sbox_all(s);
            poseidon::mds_layer_split(t, 1 + (it & 3));
            poseidon::sbox_all(t);
            poseidon::mds_layer_split(s, 1 + (it & 3));
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] ^= t[i];
    } else {
#pragma unroll 1
        for (int it = 0; it < ITER / 16; it++) poseidon::permute<false>(s);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= s[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
static double run(const char* name, K kern, int warps_per_smsp, double units_per_warp) {
    const int blocks_per_sm = warps_per_smsp / 2;   // 256-thread blocks: two warps per sub-partition each
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint64_t* out;
    const int blocks = sms * blocks_per_sm;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, 256>>>(out, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<blocks, 256>>>(out, 2 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    const double cycles = best * 1e-3 * khz * 1e3;
    const double per_unit = cycles / (warps_per_smsp * units_per_warp);
    printf("%-8s %d warps/SMSP  %8.3f ms  %9.1f cycles per unit per warp-slot\n", name, warps_per_smsp, best, per_unit);
    return per_unit;
}

int main(int argc, char** argv) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    if (poseidon::upload_constants(st) != cudaSuccess) return 1;
    for (int w : {4, 8}) {
        if (argc > 1 && atoi(argv[1]) != w) continue;
        const double S = run("S", k<0>, w, ITER);
        const double M = run("M", k<1>, w, ITER);
        const double P = run("P", k<2>, w, ITER);
        // mixed kernels: every warp does ITER units of its own kind; report per (S unit + M unit) pair of warps
        const double SM = run("S|M", k<3>, w, ITER) * 2;
        const double SP = run("S|P", k<5>, w, ITER) * 2;
        const double perm = run("perm", k<4>, w, ITER / 16);
        // per (S + M) of ONE state: the kernel does ITER S-layers and ITER linear layers per thread
        const double SM2 = run("S+M x2", k<6>, w, ITER);
        printf("  two states per thread, S-box layer of one next to the linear layer of the other: %.0f per (S + M) (sum %.0f)\n", SM2, S + M);
        printf("  %d warps/SMSP: S %.0f  M %.0f  P %.0f | one S-warp + one M-warp side by side: %.0f per (S+M) "
               "(sum %.0f, max %.0f) | S+P side by side %.0f (sum %.0f)\n",
               w, S, M, P, SM, S + M, S > M ? S : M, SP, S + P);
        printf("  permutation %.0f cycles per warp-slot; 8 (S + M) + 11 P = %.0f\n", perm, 8 * (S + M) + 11 * P);
    }
    return 0;
}
