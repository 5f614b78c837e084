// Leaf-hash kernel experiments: times merkle::leaf_hash_kernel (and tree_level_kernel) on a synthetic
// column-major LDE and prints a checksum of the digests, so that variants built with different -D flags
// can be compared for speed AND equality of results in one gpurun call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 [-DQP_...] \
//        -o leaf_NAME.bin tools/microbench/leaf_hash_bench.cu
//   ./leaf_NAME.bin [lg_leaves=21] [leaf_len=135] [reps=5]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../qp-plonky2_b200/csrc/merkle.cuh"

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
            return 1;                                                              \
        }                                                                          \
    } while (0)

__global__ void fill_kernel(uint64_t* p, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ULL;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    p[i] = z ^ (z >> 31);
}

__global__ void checksum_kernel(const uint64_t* p, size_t n, unsigned long long* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) v += p[i] * (2 * i + 1);
    atomicAdd(out, v);
}

int main(int argc, char** argv) {
    const unsigned lg = argc > 1 ? atoi(argv[1]) : 21;
    const unsigned leaf_len = argc > 2 ? atoi(argv[2]) : 135;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    const size_t n = (size_t)1 << lg;
    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    CK(poseidon::upload_constants(st));
    uint64_t *lde, *digests, *cap;
    unsigned long long* sum;
    CK(cudaMalloc(&lde, n * leaf_len * 8));
    merkle::TreeShape sh{lg, 4};
    const size_t n_dig = 2 * (n - 16);
    CK(cudaMalloc(&digests, n_dig * 32));
    CK(cudaMalloc(&cap, 16 * 32));
    CK(cudaMalloc(&sum, 8));
    fill_kernel<<<(unsigned)((n * leaf_len + 255) / 256), 256, 0, st>>>(lde, n * leaf_len, 42);
    CK(cudaMemsetAsync(digests, 0, n_dig * 32, st));
    merkle::AffineLayout lay{lde, n, 1};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f, best_lvl = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0, st));
        merkle::leaf_hash_kernel<merkle::AffineLayout><<<(unsigned)((n + QP_LEAF_BLOCK - 1) / QP_LEAF_BLOCK), QP_LEAF_BLOCK, 0, st>>>(
            lay, leaf_len, sh, digests, cap, 0u, ~0u, nullptr);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        // first internal level (n/2 two_to_one permutations)
        CK(cudaEventRecord(e0, st));
        merkle::tree_level_kernel<<<(unsigned)((n / 2 + 127) / 128), 128, 0, st>>>(sh, 1, digests, cap);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_lvl) best_lvl = ms;
    }
    CK(cudaMemsetAsync(sum, 0, 8, st));
    checksum_kernel<<<1024, 256, 0, st>>>(digests, n_dig * 4, sum);
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, sum, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double perms = (double)n * ((leaf_len + 7) / 8);
    printf("leaf_hash %8.3f ms  %.4g perm/s   level1 %7.3f ms %.4g perm/s   checksum %016llx\n", best,
           perms / (best * 1e-3), best_lvl, (double)(n / 2) / (best_lvl * 1e-3), h);
    return 0;
}
