// qp_plonky2_b200: C-ABI library (include/qp_plonky2_b200.h) over the sm_100a kernels.
// Single translation unit: the __constant__ round-constant table is shared by every kernel.
//
// There is no CPU data path in this file: host code only plans launches, owns handles and
// moves bytes.  If no CUDA device is usable every entry point returns QP_ERR_CUDA.
#include "../../include/qp_plonky2_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fri.cuh"
#include "goldilocks.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "openings.cuh"
#include "poseidon.cuh"
#include "quotient.cuh"

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct ScaleTables {
    uint64_t* lo = nullptr;  // [n_inner][2^split]
    uint64_t* hi = nullptr;  // [n_inner][2^(L-split)]
    int split = 0;
};

struct qp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t* tw = nullptr;  // tw[(1<<lg)+e] = w_{2^lg}^e, e < 2^lg
    unsigned tw_lg = 0;
    int sm_count = 148;
    // small device -> host results (caps, openings, Merkle paths) go through pinned memory: a copy
    // into a pageable buffer is staged by the driver and costs several microseconds more, and a
    // proof makes dozens of them.  Guarded by `mu`: finished handles may be read from any thread.
    static constexpr size_t STAGE_WORDS = 32768;
    uint64_t* stage = nullptr;
    std::atomic<uint64_t> launches{0};
    std::string err;
    std::mutex mu;
    // coset scale tables keyed by (L, rate_bits, block_first, block_count) for the LDE
    std::map<std::vector<uint64_t>, ScaleTables> scale_cache;
    cudaEvent_t ev[8] = {};
    // host -> device uploads of large inputs run on their own stream, in column groups, so that
    // the transfer of group g+1 overlaps the transforms of group g (qp_batch_from_values)
    cudaStream_t copy_stream = nullptr;
    static constexpr int MAX_GROUPS = 16;
    cudaEvent_t copy_ev[MAX_GROUPS] = {};
    cudaEvent_t ready_ev = nullptr;
    cudaEvent_t grp_ev[3 * MAX_GROUPS] = {};  // per column group: start, after LDE, after partial leaf hash
    // pinned staging ring for pageable host columns (qp_batch_from_values_cols): allocated on first
    // use, kept for the life of the context
    static constexpr int RING_SLOTS = 3;
    uint64_t* ring[RING_SLOTS] = {};
    size_t ring_words = 0;
};

#define CUDA_TRY(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return QP_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                          \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                     \
        (ctx)->launches++;                                                                   \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#kernel) + ": " + cudaGetErrorString(_e);               \
            return QP_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static int fail(qp_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// stream-ordered allocations from the device's default pool (kept cached between commits)
static int dev_alloc(qp_ctx* ctx, uint64_t** p, size_t n_words) {
    *p = nullptr;
    if (n_words == 0) return QP_OK;
    cudaSetDevice(ctx->device);  // a handle may be read while another device is current (multi-device callers)
    cudaError_t e = cudaMallocAsync((void**)p, n_words * 8, ctx->stream);
    static const bool trace_alloc = getenv("QP_TRACE_ALLOC") != nullptr;
    if (trace_alloc)
        fprintf(stderr, "[alloc] dev %d ctx %p stream %p: %zu bytes -> %p %s\n", ctx->device, (void*)ctx, (void*)ctx->stream,
                n_words * 8, (void*)*p, cudaGetErrorString(e));
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (trace_alloc) {   // can the device itself still give memory?
            void* q = nullptr;
            cudaError_t e2 = cudaMalloc(&q, n_words * 8);
            fprintf(stderr, "[alloc]   plain cudaMalloc of the same size: %s\n", cudaGetErrorString(e2));
            if (e2 == cudaSuccess) cudaFree(q);
            cudaGetLastError();
        }
        // what the device and the pool held when the request failed (an out-of-memory with free memory left points
        // at the pool's mappings, not at the request)
        size_t free_b = 0, total_b = 0;
        uint64_t reserved = 0, used = 0;
        cudaMemGetInfo(&free_b, &total_b);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        }
        cudaGetLastError();
        char buf[256];
        snprintf(buf, sizeof buf, "cudaMallocAsync of %zu bytes on device %d: %s (device free %zu of %zu MiB; pool reserved %llu MiB, in use %llu MiB)",
                 n_words * 8, ctx->device, cudaGetErrorString(e), free_b >> 20, total_b >> 20,
                 (unsigned long long)(reserved >> 20), (unsigned long long)(used >> 20));
        ctx->err = buf;
        return e == cudaErrorMemoryAllocation ? QP_ERR_TOO_LARGE : QP_ERR_CUDA;
    }
    return QP_OK;
}
static void dev_free(qp_ctx* ctx, uint64_t* p) {
    static const bool trace_alloc = getenv("QP_TRACE_ALLOC") != nullptr;
    if (p) {
        cudaError_t e = cudaFreeAsync(p, ctx->stream);
        if (trace_alloc) fprintf(stderr, "[free ] dev %d ctx %p stream %p: %p %s\n", ctx->device, (void*)ctx, (void*)ctx->stream, (void*)p, cudaGetErrorString(e));
    }
}

// Stream-ordered temporaries of one call: everything allocated through the scope is freed when the
// scope ends, on every path (the early returns of CUDA_TRY / LAUNCH included); keep() hands a buffer
// over to a longer-lived owner.
struct TempScope {
    qp_ctx* ctx;
    std::vector<uint64_t*> bufs;
    explicit TempScope(qp_ctx* c) : ctx(c) {}
    TempScope(const TempScope&) = delete;
    TempScope& operator=(const TempScope&) = delete;
    ~TempScope() {
        for (uint64_t* p : bufs) dev_free(ctx, p);
    }
    int alloc(uint64_t** p, size_t n_words) {
        int rc = dev_alloc(ctx, p, n_words);
        if (!rc && *p) bufs.push_back(*p);
        return rc;
    }
    void adopt(uint64_t* p) {
        if (p) bufs.push_back(p);
    }
    uint64_t* keep(uint64_t* p) {
        bufs.erase(std::remove(bufs.begin(), bufs.end(), p), bufs.end());
        return p;
    }
};

static int copy_out(qp_ctx* ctx, uint64_t* dst, int space, const uint64_t* src_dev, size_t n_words) {
    if (!dst) return fail(ctx, QP_ERR_BAD_ARG, "null output buffer");
    if (n_words == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (space != QP_DEVICE && ctx->stage && n_words <= qp_ctx::STAGE_WORDS) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage, src_dev, n_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        std::memcpy(dst, ctx->stage, n_words * 8);
        return QP_OK;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src_dev, n_words * 8,
                                  space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                  ctx->stream));
    if (space != QP_DEVICE) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}

// caller buffer -> device pointer (copying if it is host memory); *owned tells who frees
static int to_device(qp_ctx* ctx, const uint64_t* p, int space, size_t n_words, const uint64_t** dev,
                     uint64_t** owned) {
    *owned = nullptr;
    if (!p && n_words) return fail(ctx, QP_ERR_BAD_ARG, "null input buffer");
    if (space == QP_DEVICE) {
        *dev = p;
        return QP_OK;
    }
    int rc = dev_alloc(ctx, owned, n_words);
    if (rc) return rc;
    if (n_words)
        CUDA_TRY(ctx, cudaMemcpyAsync(*owned, p, n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
    *dev = *owned;
    return QP_OK;
}

// ---- table construction kernels -------------------------------------------------------------
// tw[(1<<lg) + e] = w_{2^lg}^e.  One thread per entry: e-th power by square-and-multiply of the
// level's generator (host-computed, passed in gens[lg]).
__global__ void build_twiddles_kernel(uint64_t* tw, unsigned max_lg, const uint64_t* gens) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= ((size_t)2 << max_lg)) return;
    if (id == 0) {
        tw[0] = 0;
        return;
    }
    const unsigned lg = 63 - __clzll((long long)id);
    const uint64_t e = id - ((uint64_t)1 << lg);
    tw[id] = gl::canon(gl::pow(gens[lg], e));
}

// lo[q][c] = s_q^c, hi[q][t] = s_q^(t << split)
__global__ void build_scale_kernel(uint64_t* lo, uint64_t* hi, const uint64_t* shifts, unsigned n_inner,
                                   int L, int split) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_lo = (size_t)1 << split, n_hi = (size_t)1 << (L - split);
    if (id < n_inner * n_lo) {
        const unsigned q = id >> split;
        lo[id] = gl::canon(gl::pow(shifts[q], id & (n_lo - 1)));
    } else if (id < n_inner * (n_lo + n_hi)) {
        const size_t k = id - n_inner * n_lo;
        const unsigned q = k >> (L - split);
        hi[k] = gl::canon(gl::pow(shifts[q], (k & (n_hi - 1)) << split));
    }
}

// out[i] = in[bitrev(i)] for `n_vec` vectors of 2^lg elements (canonicalising)
__global__ void bitrev_permute_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                      unsigned lg, size_t n_vec) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (n_vec << lg)) return;
    const size_t v = id >> lg, i = id & (((size_t)1 << lg) - 1);
    const size_t src = lg ? (size_t)(__brevll((unsigned long long)i) >> (64 - lg)) : 0;
    out[id] = gl::canon(in[(v << lg) + src]);
}

// salt columns: dst[k][i_loc] = salt[k][bitrev_lgN(first_leaf + i_loc)]
__global__ void salt_to_leaf_order_kernel(const uint64_t* __restrict__ salt, uint64_t* __restrict__ dst,
                                          unsigned lg_N, size_t first_leaf, size_t n_loc) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= QP_SALT_SIZE * n_loc) return;
    const size_t k = id / n_loc, i = id % n_loc;
    const size_t leaf = first_leaf + i;
    const size_t src = lg_N ? (size_t)(__brevll((unsigned long long)leaf) >> (64 - lg_N)) : 0;
    dst[id] = gl::canon(salt[(k << lg_N) + src]);
}

__global__ void __launch_bounds__(128) permute_states_kernel(uint64_t* states, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = states[i * 12 + k];
    poseidon::permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) states[i * 12 + k] = gl::canon(s[k]);
}

// ---------------------------------------------------------------------------------------------
extern "C" int qp_ctx_create(int device, void* stream, unsigned max_lde_log, qp_ctx** out) {
    if (!out) return QP_ERR_BAD_ARG;
    *out = nullptr;
    // 32-bit indices inside the transform kernels (1u << L, grid sizes): 2^30 points is the ceiling
    if (max_lde_log > 30) return QP_ERR_TOO_LARGE;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return QP_ERR_CUDA;
    }
    qp_ctx* ctx = new qp_ctx();
    ctx->device = device;
    CUDA_TRY(ctx, cudaSetDevice(device));
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return QP_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (cudaMallocHost((void**)&ctx->stage, qp_ctx::STAGE_WORDS * 8) != cudaSuccess) {
        cudaGetLastError();
        ctx->stage = nullptr;  // not fatal: results are then copied straight into the caller's buffer
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sm_count < 1) ctx->sm_count = 148;
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (auto& e : ctx->copy_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ready_ev, cudaEventDisableTiming);
    for (auto& e : ctx->grp_ev) cudaEventCreate(&e);
    // keep freed blocks cached in the pool: commits allocate and free multi-GB buffers
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    // Poseidon round constants (+ derived constants of the fused partial rounds) -> constant memory
    if (poseidon::upload_constants(ctx->stream) != cudaSuccess) {
        delete ctx;
        return QP_ERR_CUDA;
    }
    // twiddle table of full rows up to max_lde_log (fft_root_table, field/src/fft.rs:14-33)
    unsigned lg = max_lde_log < 1 ? 1 : max_lde_log;
    ctx->tw_lg = lg;
    int rc = dev_alloc(ctx, &ctx->tw, (size_t)2 << lg);
    if (rc) {
        delete ctx;
        return rc;
    }
    uint64_t gens[33];
    for (unsigned k = 0; k <= 32; k++) gens[k] = gl::host_primitive_root(k);
    uint64_t* d_gens = nullptr;
    rc = dev_alloc(ctx, &d_gens, 33);
    if (rc) {
        delete ctx;
        return rc;
    }
    cudaMemcpyAsync(d_gens, gens, sizeof gens, cudaMemcpyHostToDevice, ctx->stream);
    build_twiddles_kernel<<<cdiv((size_t)2 << lg, 256), 256, 0, ctx->stream>>>(ctx->tw, lg, d_gens);
    ctx->launches++;
    dev_free(ctx, d_gens);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        delete ctx;
        return QP_ERR_CUDA;
    }
    *out = ctx;
    return QP_OK;
}

extern "C" void qp_ctx_destroy(qp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->scale_cache) {
        cudaFreeAsync(kv.second.lo, ctx->stream);
        cudaFreeAsync(kv.second.hi, ctx->stream);
    }
    cudaFreeAsync(ctx->tw, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    {
        // the release threshold is raised for the life of a context (a 9 GB LDE is recycled between commits);
        // when a context goes away, what the pool holds and nobody uses goes back to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    for (auto& e : ctx->ev) cudaEventDestroy(e);
    for (auto& e : ctx->copy_ev) cudaEventDestroy(e);
    cudaEventDestroy(ctx->ready_ev);
    for (auto& e : ctx->grp_ev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (auto& r : ctx->ring)
        if (r) cudaFreeHost(r);
    delete ctx;
}

extern "C" const char* qp_last_error(const qp_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" int qp_ctx_synchronize(qp_ctx* ctx) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}
extern "C" uint64_t qp_ctx_launch_count(const qp_ctx* ctx) { return ctx ? ctx->launches.load() : (uint64_t)0; }

// ---------------------------------------------------------------------------------------------
// NTT driver
// ---------------------------------------------------------------------------------------------
struct NttJob {
    const uint64_t* src = nullptr;
    uint64_t* dst = nullptr;
    int L = 0;
    unsigned n_vec = 1;
    int inner_bits = 0;
    size_t src_outer = 0, src_inner = 0, dst_outer = 0, dst_inner = 0;
    const ScaleTables* scale = nullptr;
    int out_mode = ntt::OUT_NATURAL;
    // OUT_INVERSE permutes across CTAs in its final pass, so that pass cannot run in place: the
    // strided passes work in `scratch` ([n_vec][2^L], may alias src if the caller owns src) and
    // the final pass goes scratch -> dst.  nullptr = allocate a temporary.
    uint64_t* scratch = nullptr;
};

template <int K>
static int launch_strided(qp_ctx* ctx, const ntt::PassParams& p) {
    const size_t grid64 = (size_t)p.n_vec << (p.L - ntt::TILE_LOG);
    if (grid64 > 0x7fffffffULL) return fail(ctx, QP_ERR_TOO_LARGE, "transform batch exceeds the launch grid limit");
    const unsigned grid = (unsigned)grid64;
    LAUNCH(ctx, ntt::strided_pass_kernel<K>, grid, ntt::THREADS, ntt::SMEM_ELEMS * 8, p);
    return QP_OK;
}
template <int K>
static int launch_final(qp_ctx* ctx, const ntt::PassParams& p) {
    const size_t chunks = (size_t)p.n_vec << (p.L - K);
    if (((chunks - 1) >> (ntt::TILE_LOG - K)) + 1 > 0x7fffffffULL)
        return fail(ctx, QP_ERR_TOO_LARGE, "transform batch exceeds the launch grid limit");
    const unsigned grid = cdiv(chunks, (size_t)1 << (ntt::TILE_LOG - K));
    LAUNCH(ctx, ntt::final_pass_kernel<K>, grid, ntt::THREADS, ntt::SMEM_ELEMS * 8, p);
    return QP_OK;
}

static int dispatch_strided(qp_ctx* ctx, int K, const ntt::PassParams& p) {
    switch (K) {
        case 1: return launch_strided<1>(ctx, p);
        case 2: return launch_strided<2>(ctx, p);
        case 3: return launch_strided<3>(ctx, p);
        case 4: return launch_strided<4>(ctx, p);
        case 5: return launch_strided<5>(ctx, p);
        case 6: return launch_strided<6>(ctx, p);
        case 7: return launch_strided<7>(ctx, p);
        case 8: return launch_strided<8>(ctx, p);
        case 9: return launch_strided<9>(ctx, p);
        case 10: return launch_strided<10>(ctx, p);
    }
    return fail(ctx, QP_ERR_BAD_ARG, "bad strided K");
}
static int dispatch_final(qp_ctx* ctx, int K, const ntt::PassParams& p) {
    switch (K) {
        case 0: return launch_final<0>(ctx, p);
        case 1: return launch_final<1>(ctx, p);
        case 2: return launch_final<2>(ctx, p);
        case 3: return launch_final<3>(ctx, p);
        case 4: return launch_final<4>(ctx, p);
        case 5: return launch_final<5>(ctx, p);
        case 6: return launch_final<6>(ctx, p);
        case 7: return launch_final<7>(ctx, p);
        case 8: return launch_final<8>(ctx, p);
        case 9: return launch_final<9>(ctx, p);
        case 10: return launch_final<10>(ctx, p);
        case 11: return launch_final<11>(ctx, p);
        case 12: return launch_final<12>(ctx, p);
    }
    return fail(ctx, QP_ERR_BAD_ARG, "bad final K");
}

// Forward DIF transform of every vector; result at dst in bit-reversed order (OUT_NATURAL means
// "leave each element where DIF puts it") or as inverse-transform coefficients (OUT_INVERSE).
static int run_ntt(qp_ctx* ctx, const NttJob& job) {
    if ((unsigned)job.L > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "transform larger than the context's twiddle table");
    if (job.n_vec == 0) return QP_OK;
    const int L = job.L;
    // The final pass handles the last Kf stages.  Natural output: Kf = 12 (one contiguous tile
    // per CTA).  Inverse output: Kf = 10 so that 4 chunks per CTA give 32-byte store runs.
    int Kf = (job.out_mode == ntt::OUT_INVERSE) ? 10 : 12;
    if (L <= ntt::TILE_LOG) Kf = L;
    // strided passes above it: K <= kmax so that the tile keeps >= 2^(12-kmax) contiguous columns
    const int kmax = (job.out_mode == ntt::OUT_INVERSE) ? 10 : 8;
    const int rem = L - Kf;
    const int n_strided = rem > 0 ? (rem + kmax - 1) / kmax : 0;
    ntt::PassParams p{};
    p.tw = ctx->tw;
    p.L = L;
    p.n_vec = job.n_vec;
    p.inner_bits = job.inner_bits;
    p.out_mode = job.out_mode;
    p.n_inv = gl::P - ((gl::P - 1) >> L);  // inverse_2exp, field/src/types.rs:239-278
    // where the strided passes live
    uint64_t* work = job.dst;
    size_t work_outer = job.dst_outer, work_inner = job.dst_inner;
    uint64_t* tmp = nullptr;
    if (job.out_mode == ntt::OUT_INVERSE && n_strided > 0) {
        work = job.scratch;
        if (!work) {
            int rc = dev_alloc(ctx, &tmp, (size_t)job.n_vec << L);
            if (rc) return rc;
            work = tmp;
        }
        work_outer = (size_t)1 << (L + job.inner_bits);
        work_inner = (size_t)1 << L;
    }
    bool first = true;
    int s_hi = L;  // stages [s_lo, s_hi) remain above
    for (int i = 0; i < n_strided; i++) {
        const int K = rem / n_strided + (i < rem % n_strided ? 1 : 0);
        const int s_lo = s_hi - K;
        if (s_lo < ntt::TILE_LOG - K) return fail(ctx, QP_ERR_BAD_ARG, "internal: strided tile wider than block");
        p.s_lo = s_lo;
        p.src = first ? job.src : work;
        p.dst = work;
        p.src_outer_stride = first ? job.src_outer : work_outer;
        p.src_inner_stride = first ? job.src_inner : work_inner;
        p.dst_outer_stride = work_outer;
        p.dst_inner_stride = work_inner;
        p.scale_lo = (first && job.scale) ? job.scale->lo : nullptr;
        p.scale_hi = (first && job.scale) ? job.scale->hi : nullptr;
        p.scale_split = job.scale ? job.scale->split : 0;
        int rc = dispatch_strided(ctx, K, p);
        if (rc) {
            dev_free(ctx, tmp);
            return rc;
        }
        first = false;
        s_hi = s_lo;
    }
    p.s_lo = 0;
    p.src = first ? job.src : work;
    p.dst = job.dst;
    p.src_outer_stride = first ? job.src_outer : work_outer;
    p.src_inner_stride = first ? job.src_inner : work_inner;
    p.dst_outer_stride = job.dst_outer;
    p.dst_inner_stride = job.dst_inner;
    p.scale_lo = (first && job.scale) ? job.scale->lo : nullptr;
    p.scale_hi = (first && job.scale) ? job.scale->hi : nullptr;
    p.scale_split = job.scale ? job.scale->split : 0;
    int rc = dispatch_final(ctx, Kf, p);
    dev_free(ctx, tmp);
    return rc;
}

// Scale tables for x_i *= shift_q^i, q < n_inner (shifts on host).
static int build_scale(qp_ctx* ctx, int L, const std::vector<uint64_t>& shifts, ScaleTables* out) {
    const unsigned n_inner = (unsigned)shifts.size();
    out->split = L / 2;
    const size_t n_lo = (size_t)1 << out->split, n_hi = (size_t)1 << (L - out->split);
    TempScope tmp(ctx);
    uint64_t *lo = nullptr, *hi = nullptr, *d_shifts = nullptr;
    int rc = tmp.alloc(&lo, n_inner * n_lo);
    if (!rc) rc = tmp.alloc(&hi, n_inner * n_hi);
    if (!rc) rc = tmp.alloc(&d_shifts, n_inner);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_shifts, shifts.data(), n_inner * 8, cudaMemcpyHostToDevice, ctx->stream));
    // `shifts` may die before the copy runs
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    LAUNCH(ctx, build_scale_kernel, cdiv(n_inner * (n_lo + n_hi), 256), 256, 0, lo, hi, d_shifts, n_inner, L, out->split);
    out->lo = tmp.keep(lo);
    out->hi = tmp.keep(hi);
    return QP_OK;
}

static unsigned host_bitrev(unsigned x, unsigned bits) {
    unsigned r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// Coset tables of the LDE: leaf-order block q is the coset g * w_N^bitrev_r(q) (see ntt.cuh).
static int lde_scale(qp_ctx* ctx, int L, unsigned rate_bits, unsigned block_first, unsigned block_count,
                     const ScaleTables** out) {
    std::vector<uint64_t> key = {(uint64_t)L, rate_bits, block_first, block_count};
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->scale_cache.find(key);
    if (it == ctx->scale_cache.end()) {
        const uint64_t wN = gl::host_primitive_root(L + rate_bits);
        std::vector<uint64_t> shifts(block_count);
        for (unsigned b = 0; b < block_count; b++)
            shifts[b] = gl::host_mul(gl::GENERATOR, gl::host_pow(wN, host_bitrev(block_first + b, rate_bits)));
        ScaleTables st;
        int rc = build_scale(ctx, L, shifts, &st);
        if (rc) return rc;
        it = ctx->scale_cache.emplace(key, st).first;
    }
    *out = &it->second;
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// Merkle driver
// ---------------------------------------------------------------------------------------------
struct TreeBuf {
    merkle::TreeShape shape{0, 0};
    uint64_t* digests = nullptr;  // [n_digests][4]
    uint64_t* cap = nullptr;      // [2^cap_height][4]
    size_t n_digests() const { return 2 * (((size_t)1 << shape.lg_leaves) - ((size_t)1 << shape.cap_height)); }
    size_t n_cap() const { return (size_t)1 << shape.cap_height; }
};

// Leaf hashing, possibly in pieces (merkle::leaf_hash_kernel), and the levels above it.
template <class Layout>
static int hash_leaves(qp_ctx* ctx, Layout lay, unsigned leaf_len, TreeBuf* t, unsigned chunk_first = 0,
                       unsigned chunk_count = ~0u, uint64_t* state = nullptr, size_t leaf_first = 0,
                       size_t leaf_count = ~(size_t)0) {
    if (!t->digests) {
        int rc = dev_alloc(ctx, &t->digests, t->n_digests() * 4);
        if (rc) return rc;
        rc = dev_alloc(ctx, &t->cap, t->n_cap() * 4);
        if (rc) return rc;
    }
    const size_t n_leaves = (size_t)1 << t->shape.lg_leaves;
    if (leaf_first >= n_leaves) return QP_OK;
    const size_t n_here = std::min(leaf_count, n_leaves - leaf_first);
    LAUNCH(ctx, merkle::leaf_hash_kernel<Layout>, cdiv(n_here, QP_LEAF_BLOCK), QP_LEAF_BLOCK, 0, lay, leaf_len,
           t->shape, t->digests, t->cap, chunk_first, chunk_count, state, leaf_first, n_here);
    return QP_OK;
}

// levels with at most this many nodes go to the 16-lanes-per-permutation kernel (merkle.cuh);
// the environment variable is for measurements (tools/bench_tree_top.py)
#ifndef QP_TREE_TOP_NODES
#define QP_TREE_TOP_NODES 2048
#endif
static size_t tree_top_nodes() {
    static const size_t v = [] {
        const char* e = getenv("QP_TREE_TOP_NODES");
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)QP_TREE_TOP_NODES;
    }();
    return v;
}
static int build_tree_levels(qp_ctx* ctx, TreeBuf* t) {
    const unsigned nl = t->shape.num_layers();
    for (unsigned layer = 1; layer <= nl; layer++) {
        const size_t nodes = (size_t)1 << (t->shape.lg_leaves - layer);
        if (nodes <= tree_top_nodes()) {
            // latency-bound levels: 16 lanes per permutation, up to five levels per launch
            const unsigned up = std::min<unsigned>(merkle::TOP_MAX_UP, nl - layer);
            LAUNCH(ctx, merkle::tree_top_kernel, cdiv(nodes, (size_t)merkle::TOP_GROUPS), merkle::TOP_BLOCK, 0, t->shape,
                   layer, up, t->digests, t->cap);
            layer += up;
            continue;
        }
        LAUNCH(ctx, merkle::tree_level_kernel, cdiv(nodes, 128), 128, 0, t->shape, layer, t->digests, t->cap);
    }
    return QP_OK;
}

template <class Layout>
static int build_tree(qp_ctx* ctx, Layout lay, unsigned leaf_len, TreeBuf* t, cudaEvent_t after_leaves = nullptr) {
    int rc = hash_leaves(ctx, lay, leaf_len, t);
    if (rc) return rc;
    if (after_leaves) cudaEventRecord(after_leaves, ctx->stream);
    return build_tree_levels(ctx, t);
}

// Siblings for many leaves at once: out[q][layer][4]
static int tree_prove_many(qp_ctx* ctx, const TreeBuf& t, const uint64_t* leaf_indices, unsigned n_q,
                           uint64_t* siblings_out) {
    const unsigned nl = t.shape.num_layers();
    if (n_q == 0 || nl == 0) return QP_OK;
    if (!leaf_indices || !siblings_out) return fail(ctx, QP_ERR_BAD_ARG, "null buffer");
    for (unsigned i = 0; i < n_q; i++)
        if (leaf_indices[i] >> t.shape.lg_leaves) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TempScope tmp(ctx);
    uint64_t* d_idx = nullptr;
    uint64_t* d_out = nullptr;
    int rc = tmp.alloc(&d_idx, n_q);
    if (!rc) rc = tmp.alloc(&d_out, (size_t)n_q * nl * 4);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_idx, leaf_indices, (size_t)n_q * 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, merkle::merkle_paths_kernel, cdiv((size_t)n_q * nl, 128), 128, 0, t.shape, t.digests, d_idx, n_q, d_out);
    return copy_out(ctx, siblings_out, QP_HOST, d_out, (size_t)n_q * nl * 4);
}

static int tree_prove(qp_ctx* ctx, const TreeBuf& t, size_t leaf_index, uint64_t* siblings_out) {
    if (!siblings_out && t.shape.num_layers()) return fail(ctx, QP_ERR_BAD_ARG, "null output buffer");
    if (leaf_index >> t.shape.lg_leaves) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned nl = t.shape.num_layers();
    if (nl == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TempScope tmp(ctx);
    uint64_t* d_idx = nullptr;
    uint64_t* d_out = nullptr;
    int rc = tmp.alloc(&d_idx, 1);
    if (!rc) rc = tmp.alloc(&d_out, (size_t)nl * 4);
    if (rc) return rc;
    uint64_t idx = leaf_index;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_idx, &idx, 8, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, merkle::merkle_paths_kernel, cdiv(nl, 64), 64, 0, t.shape, t.digests, d_idx, 1u, d_out);
    return copy_out(ctx, siblings_out, QP_HOST, d_out, (size_t)nl * 4);  // synchronises: `idx` outlives the copy
}

// ---------------------------------------------------------------------------------------------
// PolynomialBatch
// ---------------------------------------------------------------------------------------------
struct qp_batch {
    qp_ctx* ctx = nullptr;
    size_t n_cols = 0;
    unsigned degree_log = 0, rate_bits = 0, cap_height = 0;  // global parameters
    bool blinding = false;
    unsigned block_first = 0, block_count = 0;
    size_t leaf_len = 0;     // n_cols (+4 when blinding)
    size_t n_local = 0;      // local leaves = block_count << degree_log
    uint64_t* coeffs = nullptr;  // [n_cols][n]
    bool coeffs_peer_buf = false;  // coeffs is a peer-readable buffer (peer_buf_acquire), not a pool allocation
    uint64_t* lde = nullptr;     // [leaf_len][n_local], leaf order
    TreeBuf tree;                // local tree: lg_leaves = log2(n_local), local cap height
    float ms[4] = {0, 0, 0, 0};
    float ms_leaf_hash = 0, ms_tree_levels = 0;
    // a batch under construction (qp_batch_begin .. qp_batch_end): which columns have their LDE, how
    // far the leaf sponges have absorbed the column prefix, and their state ([12][n_local])
    std::vector<uint8_t> col_ready;
    size_t cols_prefix = 0;
    unsigned chunks_done = 0;
    uint64_t* sponge_state = nullptr;
};

// Buffers that OTHER devices read (the coefficient matrices of a multi-device commit, multi_device.inl): plain
// cudaMalloc memory, which cudaDeviceEnablePeerAccess makes visible to the peers.  (Granting the peers access to the
// stream-ordered pool instead -- cudaMemPoolSetAccess -- was measured to stop the pool from growing: allocations of
// a few hundred MB failed with "out of memory" at 1-2 GB of pool and 180 GB of device memory free, while cudaMalloc
// of the same size succeeded; profiles/r02q_pool_peer_access.txt.)  cudaMalloc / cudaFree synchronise the device, so
// released buffers are kept per device and handed out again: a prover that commits the same shapes proof after
// proof allocates them once.
struct PeerBufCache {
    std::mutex mu;
    std::vector<std::pair<size_t, uint64_t*>> free_list[64];   // per device: (words, pointer)
    static constexpr size_t MAX_CACHED = 6;
};
static PeerBufCache g_peer_bufs;

static int peer_buf_acquire(qp_ctx* ctx, size_t words, uint64_t** out) {
    *out = nullptr;
    if (words == 0) return QP_OK;
    if (ctx->device < 0 || ctx->device >= 64) return fail(ctx, QP_ERR_BAD_ARG, "device index out of range");
    {
        std::lock_guard<std::mutex> lock(g_peer_bufs.mu);
        auto& fl = g_peer_bufs.free_list[ctx->device];
        size_t best = fl.size();
        for (size_t i = 0; i < fl.size(); i++)   // smallest cached buffer that fits without wasting more than half
            if (fl[i].first >= words && fl[i].first <= 2 * words && (best == fl.size() || fl[i].first < fl[best].first)) best = i;
        if (best != fl.size()) {
            *out = fl[best].second;
            fl.erase(fl.begin() + best);
            return QP_OK;
        }
    }
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc((void**)out, words * 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        ctx->err = std::string("cudaMalloc (peer-readable buffer): ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? QP_ERR_TOO_LARGE : QP_ERR_CUDA;
    }
    return QP_OK;
}
// every piece of work that reads the buffer must be complete (the callers synchronise their streams first)
static void peer_buf_release(int device, uint64_t* p, size_t words) {
    if (!p) return;
    uint64_t* victim = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_peer_bufs.mu);
        auto& fl = g_peer_bufs.free_list[device];
        fl.emplace_back(words, p);
        if (fl.size() > PeerBufCache::MAX_CACHED) {   // the oldest one goes back to the driver
            victim = fl.front().second;
            fl.erase(fl.begin());
        }
    }
    if (victim) {
        cudaSetDevice(device);
        cudaFree(victim);
    }
}
static void peer_buf_trim(int device) {
    std::vector<std::pair<size_t, uint64_t*>> all;
    {
        std::lock_guard<std::mutex> lock(g_peer_bufs.mu);
        all.swap(g_peer_bufs.free_list[device]);
    }
    cudaSetDevice(device);
    for (auto& e : all) cudaFree(e.second);
}

static bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
static unsigned ilog2(size_t x) {
    unsigned k = 0;
    while (((size_t)1 << k) < x) k++;
    return k;
}

// Phase 1: the batch object and its LDE buffer (takes ownership of d_coeffs).
static int batch_create(qp_ctx* ctx, uint64_t* d_coeffs, size_t n_cols, unsigned degree_log, unsigned rate_bits,
                        int blinding, unsigned cap_height, unsigned block_first, unsigned block_count,
                        qp_batch** out) {
    qp_batch* b = new qp_batch();
    b->ctx = ctx;
    b->n_cols = n_cols;
    b->degree_log = degree_log;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->blinding = blinding != 0;
    b->block_first = block_first;
    b->block_count = block_count;
    b->leaf_len = n_cols + (blinding ? QP_SALT_SIZE : 0);
    b->n_local = (size_t)block_count << degree_log;
    b->coeffs = d_coeffs;
    const unsigned shard_bits = rate_bits - ilog2(block_count);  // log2(#shards)
    b->tree.shape.lg_leaves = ilog2(b->n_local);
    b->tree.shape.cap_height = cap_height - shard_bits;
    *out = b;
    return dev_alloc(ctx, &b->lde, b->leaf_len * b->n_local);
}

// Phase 2: "FFT + blinding" (oracle.rs:202-206, 267-283) for columns [c0, c1).
// Columns per transform launch pair (QP_LDE_GROUP; 0 = all columns of the call in one pair).  A strided pass
// writes the whole intermediate before the final pass reads it back: with few columns per pair the
// intermediate (group x 2^rate_bits x n x 8 bytes) can stay in the 126 MB L2 between the two.
static size_t lde_group() {
    static const size_t v = [] {
        const char* e = getenv("QP_LDE_GROUP");
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)0;
    }();
    return v;
}

static int batch_lde_columns(qp_batch* b, size_t c0, size_t c1) {
    qp_ctx* ctx = b->ctx;
    const size_t n = (size_t)1 << b->degree_log;
    const ScaleTables* st = nullptr;
    int rc = lde_scale(ctx, (int)b->degree_log, b->rate_bits, b->block_first, b->block_count, &st);
    if (rc) return rc;
    const size_t grp = lde_group() ? lde_group() : (c1 - c0);
    for (size_t a = c0; a < c1 && !rc; a += grp) {
        const size_t e = std::min(a + grp, c1);
        NttJob job;
        job.src = b->coeffs + a * n;
        job.dst = b->lde + a * b->n_local;
        job.L = (int)b->degree_log;
        job.n_vec = (unsigned)((e - a) * b->block_count);
        job.inner_bits = (int)ilog2(b->block_count);
        job.src_outer = n;
        job.src_inner = 0;
        job.dst_outer = b->n_local;
        job.dst_inner = n;
        job.scale = st;
        job.out_mode = ntt::OUT_NATURAL;
        rc = run_ntt(ctx, job);
    }
    return rc;
}

// Phase 3: salt columns, "build Merkle tree" (oracle.rs:210-214), timings.  ev[1] must have been
// recorded where the LDE phase started.
static int batch_finish(qp_batch* b, const uint64_t* salt_dev, unsigned chunks_done = 0, uint64_t* state = nullptr) {
    qp_ctx* ctx = b->ctx;
    if (b->blinding) {
        LAUNCH(ctx, salt_to_leaf_order_kernel, cdiv(QP_SALT_SIZE * b->n_local, 256), 256, 0, salt_dev,
               b->lde + b->n_cols * b->n_local, b->degree_log + b->rate_bits,
               (size_t)b->block_first << b->degree_log, b->n_local);
    }
    cudaEventRecord(ctx->ev[2], ctx->stream);
    // "transpose LDEs" is fused away: the LDE is already in leaf order, column-major.
    merkle::AffineLayout lay{b->lde, b->n_local, 1};
    // chunks [0, chunks_done) of every leaf were absorbed while the input was still arriving
    int rc = hash_leaves(ctx, lay, (unsigned)b->leaf_len, &b->tree, chunks_done, ~0u, state);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[5], ctx->stream);
    rc = build_tree_levels(ctx, &b->tree);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[3], ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&b->ms[1], ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&b->ms[3], ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&b->ms_leaf_hash, ctx->ev[2], ctx->ev[5]);
    cudaEventElapsedTime(&b->ms_tree_levels, ctx->ev[5], ctx->ev[3]);
    return QP_OK;
}

// from_coeffs on device-resident coefficients (takes ownership of d_coeffs)
static int batch_from_device_coeffs(qp_ctx* ctx, uint64_t* d_coeffs, size_t n_cols, unsigned degree_log,
                                    unsigned rate_bits, int blinding, unsigned cap_height,
                                    const uint64_t* salt_dev, unsigned block_first, unsigned block_count,
                                    qp_batch** out) {
    int rc = batch_create(ctx, d_coeffs, n_cols, degree_log, rate_bits, blinding, cap_height, block_first,
                          block_count, out);
    if (rc) return rc;
    cudaEventRecord(ctx->ev[1], ctx->stream);
    rc = batch_lde_columns(*out, 0, n_cols);
    if (rc) return rc;
    return batch_finish(*out, salt_dev);
}

static int check_batch_args(qp_ctx* ctx, size_t n_cols, unsigned degree_log, unsigned rate_bits, int blinding,
                            unsigned cap_height, const uint64_t* salt, unsigned block_first,
                            unsigned block_count, qp_batch** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (n_cols == 0) return fail(ctx, QP_ERR_BAD_ARG, "polynomials[0]: empty batch (index out of bounds)");
    if (degree_log + rate_bits > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "LDE larger than the context's max_lde_log");
    if (cap_height > degree_log + rate_bits)
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most log2(leaves.len())");
    if (blinding && !salt) return fail(ctx, QP_ERR_BLINDING_NO_SALT, "blinding needs injected salt");
    if (!is_pow2(block_count) || block_count > (1u << rate_bits) || block_first % block_count ||
        block_first + block_count > (1u << rate_bits))
        return fail(ctx, QP_ERR_BAD_ARG, "bad coset block range");
    const unsigned shard_bits = rate_bits - ilog2(block_count);
    if (cap_height < shard_bits) return fail(ctx, QP_ERR_BAD_ARG, "shard smaller than a cap subtree");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return QP_OK;
}

extern "C" int qp_batch_from_coeffs(qp_ctx* ctx, const uint64_t* coeffs, int space, size_t n_cols,
                                    unsigned degree_log, unsigned rate_bits, int blinding,
                                    unsigned cap_height, const uint64_t* salt, unsigned block_first,
                                    unsigned block_count, qp_batch** out) {
    int rc = check_batch_args(ctx, n_cols, degree_log, rate_bits, blinding, cap_height, salt, block_first,
                              block_count, out);
    if (rc) return rc;
    if (!coeffs) return fail(ctx, QP_ERR_BAD_ARG, "null coeffs");
    const size_t n = (size_t)1 << degree_log;
    TempScope tmp(ctx);
    uint64_t* d_coeffs = nullptr;
    rc = tmp.alloc(&d_coeffs, n_cols * n);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_coeffs, coeffs, n_cols * n * 8,
                                  space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                  ctx->stream));
    const uint64_t* d_salt = nullptr;
    if (blinding) {
        uint64_t* salt_owned = nullptr;
        rc = to_device(ctx, salt, space, (size_t)QP_SALT_SIZE << (degree_log + rate_bits), &d_salt, &salt_owned);
        tmp.adopt(salt_owned);
        if (rc) return rc;
    }
    // the batch owns the coefficients from here on
    rc = batch_from_device_coeffs(ctx, tmp.keep(d_coeffs), n_cols, degree_log, rate_bits, blinding, cap_height, d_salt,
                                  block_first, block_count, out);
    if (rc) {
        qp_batch_free(*out);
        *out = nullptr;
    }
    return rc;
}

static int ifft_device(qp_ctx* ctx, const uint64_t* d_values, size_t n_cols, unsigned degree_log,
                       uint64_t* d_coeffs, uint64_t* scratch = nullptr) {
    NttJob job;
    job.src = d_values;
    job.dst = d_coeffs;
    job.L = (int)degree_log;
    job.n_vec = (unsigned)n_cols;
    job.inner_bits = 0;
    job.src_outer = (size_t)1 << degree_log;
    job.dst_outer = (size_t)1 << degree_log;
    job.out_mode = ntt::OUT_INVERSE;
    job.scratch = scratch;
    return run_ntt(ctx, job);
}

// QP_STAGE_MODE (measurements): 0 = stage pageable columns through the pinned ring (default), 1 = hand them
// to cudaMemcpyAsync directly
static int stage_mode() {
    static const int v = [] {
        const char* e = getenv("QP_STAGE_MODE");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// Host -> pinned staging of pageable columns, a few threads wide (one memcpy thread tops out well
// below what PCIe 5 moves).
static void stage_columns(uint64_t* dst, const uint64_t* const* cols, size_t c0, size_t c1, size_t n) {
    const size_t total = (c1 - c0) * n;
    unsigned n_thr = total * 8 >= ((size_t)8 << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1;
    const char* e = getenv("QP_STAGE_THREADS");
    if (e) n_thr = (unsigned)std::max(1, atoi(e));
    auto work = [&](unsigned t) {
        // thread t copies the t-th slice of every column (keeps all threads on distinct pages)
        const size_t lo = n * t / n_thr, hi = n * (t + 1) / n_thr;
        for (size_t c = c0; c < c1; c++) std::memcpy(dst + (c - c0) * n + lo, cols[c] + lo, (hi - lo) * 8);
    };
    if (n_thr == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n_thr; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

static bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Make sure the context's pinned ring holds `need` words per slot.
static int ensure_ring(qp_ctx* ctx, size_t need) {
    if (ctx->ring_words >= need) return QP_OK;
    for (auto& r : ctx->ring) {
        if (r) cudaFreeHost(r);
        r = nullptr;
    }
    ctx->ring_words = 0;
    for (auto& r : ctx->ring)
        if (cudaMallocHost((void**)&r, need * 8) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, QP_ERR_TOO_LARGE, "cannot allocate the pinned staging ring");
        }
    ctx->ring_words = need;
    return QP_OK;
}

// Upload host columns [c0, c1) to d_dst ([c1 - c0][n]) on the context's copy stream and record
// copy_ev[g % MAX_GROUPS]: as they are when the source is pinned, through the pinned ring (slot g % RING_SLOTS,
// reused once the copy that last read it has completed) when it is pageable.
static int upload_columns(qp_ctx* ctx, const uint64_t* const* cols, bool pinned_src, size_t c0, size_t c1, size_t n,
                          uint64_t* d_dst, int g) {
    if (pinned_src || stage_mode() == 1) {
        // (mode 1: pageable source handed to the driver as it is; the call returns once the source has been read)
        for (size_t c = c0; c < c1; c++)
            CUDA_TRY(ctx, cudaMemcpyAsync(d_dst + (c - c0) * n, cols[c], n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
        uint64_t* slot = ctx->ring[g % qp_ctx::RING_SLOTS];
        if (g >= qp_ctx::RING_SLOTS)
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->copy_ev[(g - qp_ctx::RING_SLOTS) % qp_ctx::MAX_GROUPS]));
        stage_columns(slot, cols, c0, c1, n);
        CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, slot, (c1 - c0) * n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->copy_ev[g % qp_ctx::MAX_GROUPS], ctx->copy_stream));
    return QP_OK;
}

// The commit from HOST columns: `cols[c]` points at column c (2^degree_log words).  Columns travel in
// groups of 16 (= two 8-element sponge chunks): the upload of group g + 1 overlaps the iNTT + LDE +
// partial leaf hash of group g.  pinned_src: the columns can be handed to the copy engine as they
// are; otherwise (pageable memory -- what a Rust Vec is) they are staged through the context's pinned
// ring first.
static int batch_from_host_columns(qp_ctx* ctx, const uint64_t* const* cols, bool pinned_src, size_t n_cols,
                                   unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                                   const uint64_t* salt, unsigned block_first, unsigned block_count, qp_batch** out) {
    const size_t n = (size_t)1 << degree_log;
    TempScope tmp(ctx);
    uint64_t *d_coeffs = nullptr, *d_values = nullptr, *d_state = nullptr;
    int rc = tmp.alloc(&d_coeffs, n_cols * n);
    if (!rc) rc = tmp.alloc(&d_values, n_cols * n);
    if (rc) return rc;
    rc = batch_create(ctx, tmp.keep(d_coeffs), n_cols, degree_log, rate_bits, blinding, cap_height, block_first,
                      block_count, out);
    qp_batch* b = *out;
    auto bail = [&](int code) {
        // later copies may still be in flight into d_values / out of the caller's memory
        cudaStreamSynchronize(ctx->copy_stream);
        qp_batch_free(b);
        *out = nullptr;
        return code;
    };
    if (rc) return bail(rc);
    size_t per = 16;
    while ((n_cols + per - 1) / per > (size_t)qp_ctx::MAX_GROUPS) per += 16;
    const int n_groups = (int)((n_cols + per - 1) / per);
    rc = tmp.alloc(&d_state, 12 * b->n_local);
    if (rc) return bail(rc);
    if (!pinned_src) {
        rc = ensure_ring(ctx, std::min(per, n_cols) * n);
        if (rc) return bail(rc);
    }
    // the copy stream may only touch d_values once the (stream-ordered) allocation has happened
    cudaEventRecord(ctx->ready_ev, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_stream, ctx->ready_ev, 0);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    merkle::AffineLayout lay{b->lde, b->n_local, 1};
    unsigned chunks_done = 0;
    int hashed_groups = 0;
    // pinned source: queue every upload up front; pageable source: stage group g (host work) right
    // before queueing its compute, so the host copies group g + 1 while the device works on group g
    auto upload = [&](int g) -> int {
        const size_t c0 = g * per, c1 = std::min(c0 + per, n_cols);
        return upload_columns(ctx, cols, pinned_src, c0, c1, n, d_values + c0 * n, g);
    };
    // pinned source: every upload is queued up front.  Pageable source: a stager thread walks the groups
    // (stage into the ring -- host-blocking -- then queue the copy) and publishes issued[g]; this thread
    // queues the compute of a group once its copy has been queued, so staging never waits for launches.
    std::vector<std::atomic<int>> issued(n_groups);
    for (auto& f : issued) f.store(0);
    std::atomic<int> stage_rc{QP_OK};
    std::thread stager;
    if (pinned_src) {
        for (int g = 0; g < n_groups && !rc; g++) rc = upload(g);
        for (auto& f : issued) f.store(1);
    } else {
        stager = std::thread([&] {
            cudaSetDevice(ctx->device);
            for (int g = 0; g < n_groups; g++) {
                int r = upload(g);
                if (r) stage_rc.store(r);
                issued[g].store(1, std::memory_order_release);
                if (r) {
                    for (int k = g + 1; k < n_groups; k++) issued[k].store(1, std::memory_order_release);
                    return;
                }
            }
        });
    }
    for (int g = 0; g < n_groups && !rc; g++) {
        const size_t c0 = g * per, c1 = std::min(c0 + per, n_cols);
        while (!issued[g].load(std::memory_order_acquire)) std::this_thread::yield();
        if (stage_rc.load()) {
            rc = stage_rc.load();
            break;
        }
        cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[g], 0);
        cudaEventRecord(ctx->grp_ev[3 * g], ctx->stream);
        rc = ifft_device(ctx, d_values + c0 * n, c1 - c0, degree_log, b->coeffs + c0 * n, d_values + c0 * n);
        if (!rc) rc = batch_lde_columns(b, c0, c1);
        cudaEventRecord(ctx->grp_ev[3 * g + 1], ctx->stream);
        // absorb the complete chunks of this group, leaving at least the last chunk of the leaf
        // (and the salt, if any) to batch_finish
        const unsigned chunk_end = (unsigned)(c1 / 8);
        const unsigned n_chunks = (unsigned)((b->leaf_len + 7) / 8);
        if (!rc && chunk_end > chunks_done && chunk_end < n_chunks) {
            rc = hash_leaves(ctx, lay, (unsigned)b->leaf_len, &b->tree, chunks_done, chunk_end - chunks_done, d_state);
            chunks_done = chunk_end;
            hashed_groups = g + 1;
        }
        cudaEventRecord(ctx->grp_ev[3 * g + 2], ctx->stream);
    }
    if (stager.joinable()) stager.join();
    if (!rc) rc = stage_rc.load();
    if (rc) return bail(rc);
    const uint64_t* d_salt = nullptr;
    if (blinding) {
        uint64_t* salt_owned = nullptr;
        rc = to_device(ctx, salt, QP_HOST, (size_t)QP_SALT_SIZE << (degree_log + rate_bits), &d_salt, &salt_owned);
        tmp.adopt(salt_owned);
    }
    if (!rc) rc = batch_finish(b, d_salt, chunks_done, d_state);  // synchronises the stream
    if (rc) return bail(rc);
    // scopes: the transforms and the partial leaf hashes interleave, so sum them per group
    float t_ntt = 0, t_hash = 0, ms = 0;
    for (int g = 0; g < n_groups; g++) {
        cudaEventElapsedTime(&ms, ctx->grp_ev[3 * g], ctx->grp_ev[3 * g + 1]);
        t_ntt += ms;
        if (g < hashed_groups) {
            cudaEventElapsedTime(&ms, ctx->grp_ev[3 * g + 1], ctx->grp_ev[3 * g + 2]);
            t_hash += ms;
        }
    }
    b->ms[1] = t_ntt;             // "IFFT" + "FFT + blinding" (the upload wait is not in it)
    b->ms[3] += t_hash;           // "build Merkle tree"
    b->ms_leaf_hash += t_hash;
    return QP_OK;
}

static size_t pipeline_threshold() {
    // (QP_PIPELINE_MIN_BYTES lowers the threshold so that the tests can drive the pipelined path with small inputs)
    const char* thr_env = getenv("QP_PIPELINE_MIN_BYTES");
    return thr_env ? (size_t)strtoull(thr_env, nullptr, 10) : ((size_t)64 << 20);
}

extern "C" int qp_batch_from_values(qp_ctx* ctx, const uint64_t* values, int space, size_t n_cols,
                                    unsigned degree_log, unsigned rate_bits, int blinding,
                                    unsigned cap_height, const uint64_t* salt, unsigned block_first,
                                    unsigned block_count, qp_batch** out) {
    int rc = check_batch_args(ctx, n_cols, degree_log, rate_bits, blinding, cap_height, salt, block_first,
                              block_count, out);
    if (rc) return rc;
    if (!values) return fail(ctx, QP_ERR_BAD_ARG, "null values");
    const size_t n = (size_t)1 << degree_log;
    // Host input of at least 64 MiB: upload in column groups on the copy stream and run the
    // iNTT + LDE (+ partial leaf hash) of group g while group g+1 is still in flight.
    if (space != QP_DEVICE && n_cols * n * 8 >= pipeline_threshold() && n_cols >= 2) {
        std::vector<const uint64_t*> cols(n_cols);
        for (size_t c = 0; c < n_cols; c++) cols[c] = values + c * n;
        return batch_from_host_columns(ctx, cols.data(), host_pointer_is_pinned(values), n_cols, degree_log, rate_bits,
                                       blinding, cap_height, salt, block_first, block_count, out);
    }
    TempScope tmp(ctx);
    uint64_t* d_coeffs = nullptr;
    rc = tmp.alloc(&d_coeffs, n_cols * n);
    if (rc) return rc;
    const uint64_t* d_values = nullptr;
    uint64_t* values_owned = nullptr;
    rc = to_device(ctx, values, space, n_cols * n, &d_values, &values_owned);
    tmp.adopt(values_owned);
    if (rc) return rc;
    // "IFFT" (oracle.rs:176-180)
    cudaEventRecord(ctx->ev[0], ctx->stream);
    rc = ifft_device(ctx, d_values, n_cols, degree_log, d_coeffs, values_owned);
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (rc) return rc;
    const uint64_t* d_salt = nullptr;
    if (blinding) {
        uint64_t* salt_owned = nullptr;
        rc = to_device(ctx, salt, space, (size_t)QP_SALT_SIZE << (degree_log + rate_bits), &d_salt, &salt_owned);
        tmp.adopt(salt_owned);
        if (rc) return rc;
    }
    rc = batch_from_device_coeffs(ctx, tmp.keep(d_coeffs), n_cols, degree_log, rate_bits, blinding, cap_height, d_salt,
                                  block_first, block_count, out);
    if (rc) {
        qp_batch_free(*out);
        *out = nullptr;
        return rc;
    }
    cudaEventElapsedTime(&(*out)->ms[0], ctx->ev[0], ctx->ev[4]);
    return QP_OK;
}

// PolynomialBatch::from_values on what the reference actually passes (oracle.rs:168-175): one heap
// vector per column, pageable.  Small inputs are gathered into one upload; large ones go through
// the staged pipeline above.
extern "C" int qp_batch_from_values_cols(qp_ctx* ctx, const uint64_t* const* cols, size_t n_cols,
                                         unsigned degree_log, unsigned rate_bits, int blinding, unsigned cap_height,
                                         const uint64_t* salt, unsigned block_first, unsigned block_count,
                                         qp_batch** out) {
    int rc = check_batch_args(ctx, n_cols, degree_log, rate_bits, blinding, cap_height, salt, block_first,
                              block_count, out);
    if (rc) return rc;
    if (!cols) return fail(ctx, QP_ERR_BAD_ARG, "null column table");
    for (size_t c = 0; c < n_cols; c++)
        if (!cols[c]) return fail(ctx, QP_ERR_BAD_ARG, "null column");
    const size_t n = (size_t)1 << degree_log;
    if (n_cols * n * 8 >= pipeline_threshold() && n_cols >= 2)
        return batch_from_host_columns(ctx, cols, host_pointer_is_pinned(cols[0]) && host_pointer_is_pinned(cols[n_cols - 1]),
                                       n_cols, degree_log, rate_bits, blinding, cap_height, salt, block_first,
                                       block_count, out);
    std::vector<uint64_t> flat(n_cols * n);
    for (size_t c = 0; c < n_cols; c++) std::memcpy(flat.data() + c * n, cols[c], n * 8);
    // (the small-input path synchronises before it returns, so `flat` outlives its upload)
    return qp_batch_from_values(ctx, flat.data(), QP_HOST, n_cols, degree_log, rate_bits, blinding, cap_height, salt,
                                block_first, block_count, out);
}

// from_coeffs in pieces (multi-GPU: the LDE of the columns that have arrived overlaps the
// all-gather of the rest): begin allocates, put copies coefficient columns [c0, c0 + count) in and
// extends them, end hashes.  Same kernels and the same result as qp_batch_from_coeffs.
extern "C" int qp_batch_begin(qp_ctx* ctx, size_t n_cols, unsigned degree_log, unsigned rate_bits, int blinding,
                              unsigned cap_height, unsigned block_first, unsigned block_count, qp_batch** out) {
    static const uint64_t dummy_salt = 0;  // presence is checked at qp_batch_end
    int rc = check_batch_args(ctx, n_cols, degree_log, rate_bits, blinding, cap_height,
                              blinding ? &dummy_salt : nullptr, block_first, block_count, out);
    if (rc) return rc;
    uint64_t* d_coeffs = nullptr;
    rc = dev_alloc(ctx, &d_coeffs, n_cols << degree_log);
    if (rc) return rc;
    rc = batch_create(ctx, d_coeffs, n_cols, degree_log, rate_bits, blinding, cap_height, block_first, block_count, out);
    if (rc) {
        if (*out) {
            qp_batch_free(*out);
            *out = nullptr;
        } else {
            dev_free(ctx, d_coeffs);
        }
        return rc;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    return QP_OK;
}

// The LDE of columns [c0, c0 + count) of a batch under construction (their coefficients are in place),
// and -- when `absorb` -- the leaf sponges advanced over every complete 8-column chunk of the column
// PREFIX that is now extended (merkle::leaf_hash_kernel in pieces): with columns arriving in order, the
// hashing of the early columns overlaps the arrival of the late ones, and qp_batch_end only has the last
// chunk left.
static int batch_extend(qp_batch* b, size_t c0, size_t count, bool absorb) {
    qp_ctx* ctx = b->ctx;
    int rc = batch_lde_columns(b, c0, c0 + count);
    if (rc) return rc;
    if (b->col_ready.size() != b->n_cols) b->col_ready.assign(b->n_cols, 0);
    for (size_t c = c0; c < c0 + count; c++) b->col_ready[c] = 1;
    while (b->cols_prefix < b->n_cols && b->col_ready[b->cols_prefix]) b->cols_prefix++;
    if (!absorb) return QP_OK;
    const unsigned chunk_end = (unsigned)(b->cols_prefix / 8);
    const unsigned n_chunks = (unsigned)((b->leaf_len + 7) / 8);
    if (chunk_end > b->chunks_done && chunk_end < n_chunks) {
        if (!b->sponge_state) {
            rc = dev_alloc(ctx, &b->sponge_state, 12 * b->n_local);
            if (rc) return rc;
        }
        merkle::AffineLayout lay{b->lde, b->n_local, 1};
        rc = hash_leaves(ctx, lay, (unsigned)b->leaf_len, &b->tree, b->chunks_done, chunk_end - b->chunks_done,
                         b->sponge_state);
        b->chunks_done = chunk_end;
    }
    return rc;
}

extern "C" int qp_batch_put_coeffs(qp_batch* b, const uint64_t* coeffs, int space, size_t c0, size_t count) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (!coeffs || c0 + count > b->n_cols) return fail(ctx, QP_ERR_BAD_ARG, "column range out of bounds");
    if (count == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << b->degree_log;
    CUDA_TRY(ctx, cudaMemcpyAsync(b->coeffs + c0 * n, coeffs, count * n * 8,
                                  space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    if (space != QP_DEVICE) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffer may go away
    return batch_extend(b, c0, count, false);
}

extern "C" uint64_t* qp_batch_coeffs_slot(qp_batch* b, size_t c0) {
    if (!b || c0 >= b->n_cols) return nullptr;
    return b->coeffs + (c0 << b->degree_log);
}

extern "C" int qp_batch_extend_columns(qp_batch* b, size_t c0, size_t count, int absorb) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (c0 + count > b->n_cols) return fail(ctx, QP_ERR_BAD_ARG, "column range out of bounds");
    if (count == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return batch_extend(b, c0, count, absorb != 0);
}

extern "C" int qp_batch_end(qp_batch* b, const uint64_t* salt, int space) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (b->blinding && !salt) return fail(ctx, QP_ERR_BLINDING_NO_SALT, "Cannot set blinding without salt");
    if (b->cols_prefix != b->n_cols && !b->col_ready.empty())
        return fail(ctx, QP_ERR_BAD_ARG, "qp_batch_end before every column was put");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TempScope tmp(ctx);
    const uint64_t* d_salt = nullptr;
    int rc = QP_OK;
    if (b->blinding) {
        uint64_t* salt_owned = nullptr;
        rc = to_device(ctx, salt, space, (size_t)QP_SALT_SIZE << (b->degree_log + b->rate_bits), &d_salt, &salt_owned);
        tmp.adopt(salt_owned);
    }
    if (!rc) rc = batch_finish(b, d_salt, b->chunks_done, b->sponge_state);
    dev_free(ctx, b->sponge_state);
    b->sponge_state = nullptr;
    return rc;
}

extern "C" int qp_ifft_columns(qp_ctx* ctx, const uint64_t* values, int space, size_t n_cols,
                               unsigned degree_log, uint64_t* coeffs_out, int out_space) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!values || !coeffs_out) return fail(ctx, QP_ERR_BAD_ARG, "null buffer");
    if (degree_log > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "transform larger than max_lde_log");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << degree_log;
    const uint64_t* d_values = nullptr;
    uint64_t* values_owned = nullptr;
    int rc = to_device(ctx, values, space, n_cols * n, &d_values, &values_owned);
    if (rc) return rc;
    uint64_t* d_coeffs = coeffs_out;
    uint64_t* coeffs_owned = nullptr;
    if (out_space != QP_DEVICE) {
        rc = dev_alloc(ctx, &coeffs_owned, n_cols * n);
        if (rc) return rc;
        d_coeffs = coeffs_owned;
    }
    rc = ifft_device(ctx, d_values, n_cols, degree_log, d_coeffs, values_owned);
    if (!rc && out_space != QP_DEVICE) rc = copy_out(ctx, coeffs_out, QP_HOST, d_coeffs, n_cols * n);
    dev_free(ctx, values_owned);
    dev_free(ctx, coeffs_owned);
    return rc;
}

extern "C" void qp_batch_free(qp_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->coeffs_peer_buf) {
        cudaStreamSynchronize(b->ctx->stream);   // readers of the matrix on this device; the peers' copies ended with the commit
        peer_buf_release(b->ctx->device, b->coeffs, b->n_cols << b->degree_log);
    } else {
        dev_free(b->ctx, b->coeffs);
    }
    dev_free(b->ctx, b->sponge_state);
    dev_free(b->ctx, b->lde);
    dev_free(b->ctx, b->tree.digests);
    dev_free(b->ctx, b->tree.cap);
    delete b;
}

extern "C" size_t qp_batch_cap_len(const qp_batch* b) { return b ? b->tree.n_cap() : 0; }
extern "C" size_t qp_batch_digests_len(const qp_batch* b) { return b ? b->tree.n_digests() : 0; }
extern "C" size_t qp_batch_leaf_len(const qp_batch* b) { return b ? b->leaf_len : 0; }
extern "C" const uint64_t* qp_batch_device_lde(const qp_batch* b) { return b ? b->lde : nullptr; }
extern "C" const uint64_t* qp_batch_device_coeffs(const qp_batch* b) { return b ? b->coeffs : nullptr; }

extern "C" int qp_batch_cap(const qp_batch* b, uint64_t* out, int space) {
    if (!b) return QP_ERR_BAD_ARG;
    return copy_out(b->ctx, out, space, b->tree.cap, b->tree.n_cap() * 4);
}
extern "C" int qp_batch_coeffs(const qp_batch* b, uint64_t* out, int space) {
    if (!b) return QP_ERR_BAD_ARG;
    return copy_out(b->ctx, out, space, b->coeffs, b->n_cols << b->degree_log);
}
extern "C" int qp_batch_digests(const qp_batch* b, uint64_t* out, int space) {
    if (!b) return QP_ERR_BAD_ARG;
    return copy_out(b->ctx, out, space, b->tree.digests, b->tree.n_digests() * 4);
}

static int batch_gather(const qp_batch* b, const uint64_t* idx_host, size_t first, size_t count,
                        unsigned row_len, uint64_t* out, int space) {
    qp_ctx* ctx = b->ctx;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null output buffer");
    if (count == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TempScope tmp(ctx);
    uint64_t* d_idx = nullptr;
    int rc;
    if (idx_host) {
        for (size_t i = 0; i < count; i++)
            if (idx_host[i] >= b->n_local) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
        rc = tmp.alloc(&d_idx, count);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(d_idx, idx_host, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    } else if (first + count > b->n_local) {
        return fail(ctx, QP_ERR_BAD_ARG, "leaf range out of bounds");
    }
    uint64_t* d_out = out;
    if (space != QP_DEVICE) {
        rc = tmp.alloc(&d_out, count * row_len);
        if (rc) return rc;
    }
    merkle::AffineLayout lay{b->lde, b->n_local, 1};
    LAUNCH(ctx, merkle::gather_rows_kernel<merkle::AffineLayout>, cdiv(count * row_len, 256), 256, 0, lay,
           row_len, d_idx, first, count, d_out);
    if (space != QP_DEVICE) return copy_out(ctx, out, QP_HOST, d_out, count * row_len);
    // device output: the index upload must not outlive the caller's buffer
    if (idx_host) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}

extern "C" int qp_batch_leaves(const qp_batch* b, size_t first, size_t count, uint64_t* out, int space) {
    if (!b) return QP_ERR_BAD_ARG;
    return batch_gather(b, nullptr, first, count, (unsigned)b->leaf_len, out, space);
}

extern "C" int qp_batch_get_lde_values(const qp_batch* b, size_t index, size_t step, uint64_t* out) {
    if (!b) return QP_ERR_BAD_ARG;
    const unsigned bits = b->degree_log + b->rate_bits;
    const size_t nat = index * step;
    if (nat >> bits) return fail(b->ctx, QP_ERR_BAD_ARG, "LDE index out of range");
    size_t leaf = 0;
    for (unsigned i = 0; i < bits; i++) leaf |= ((nat >> i) & 1) << (bits - 1 - i);
    const size_t first_leaf = (size_t)b->block_first << b->degree_log;
    if (leaf < first_leaf || leaf >= first_leaf + b->n_local)
        return fail(b->ctx, QP_ERR_BAD_ARG, "LDE row lives on another shard");
    // salt stripped (oracle.rs:290)
    return batch_gather(b, nullptr, leaf - first_leaf, 1, (unsigned)b->n_cols, out, QP_HOST);
}

extern "C" int qp_batch_get_leaves(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* out) {
    if (!b) return QP_ERR_BAD_ARG;
    if (!leaf_indices && n) return fail(b->ctx, QP_ERR_BAD_ARG, "null indices");
    return batch_gather(b, leaf_indices, 0, n, (unsigned)b->leaf_len, out, QP_HOST);
}

extern "C" int qp_batch_prove(const qp_batch* b, size_t leaf_index, uint64_t* siblings_out) {
    if (!b) return QP_ERR_BAD_ARG;
    return tree_prove(b->ctx, b->tree, leaf_index, siblings_out);
}

extern "C" int qp_batch_prove_many(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* siblings_out) {
    if (!b) return QP_ERR_BAD_ARG;
    return tree_prove_many(b->ctx, b->tree, leaf_indices, n, siblings_out);
}

// Rows and Merkle paths of n leaves in one round trip (the query openings of fri_proof: one call per
// oracle instead of two): one index upload, two kernels, one device buffer back.
extern "C" int qp_batch_open_many(const qp_batch* b, const uint64_t* leaf_indices, unsigned n, uint64_t* rows_out,
                                  uint64_t* siblings_out) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (n == 0) return QP_OK;
    if (!leaf_indices || !rows_out || (!siblings_out && b->tree.shape.num_layers()))
        return fail(ctx, QP_ERR_BAD_ARG, "null buffer");
    for (unsigned i = 0; i < n; i++)
        if (leaf_indices[i] >= b->n_local) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned nl = b->tree.shape.num_layers();
    const size_t row_words = (size_t)n * b->leaf_len, path_words = (size_t)n * nl * 4;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TempScope tmp(ctx);
    uint64_t* d_idx = nullptr;
    uint64_t* d_out = nullptr;
    int rc = tmp.alloc(&d_idx, n);
    if (!rc) rc = tmp.alloc(&d_out, row_words + path_words ? row_words + path_words : 1);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_idx, leaf_indices, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (row_words) {
        merkle::AffineLayout lay{b->lde, b->n_local, 1};
        LAUNCH(ctx, merkle::gather_rows_kernel<merkle::AffineLayout>, cdiv(row_words, 256), 256, 0, lay,
               (unsigned)b->leaf_len, d_idx, (size_t)0, (size_t)n, d_out);
    }
    if (path_words)
        LAUNCH(ctx, merkle::merkle_paths_kernel, cdiv((size_t)n * nl, 128), 128, 0, b->tree.shape, b->tree.digests, d_idx, n,
               d_out + row_words);
    std::vector<uint64_t> host(row_words + path_words);
    rc = copy_out(ctx, host.data(), QP_HOST, d_out, row_words + path_words);
    if (!rc) {
        std::memcpy(rows_out, host.data(), row_words * 8);
        if (path_words) std::memcpy(siblings_out, host.data() + row_words, path_words * 8);
    }
    return rc;
}

extern "C" int qp_batch_timing(const qp_batch* b, double ms[4]) {
    if (!b || !ms) return QP_ERR_BAD_ARG;
    for (int i = 0; i < 4; i++) ms[i] = b->ms[i];
    return QP_OK;
}
extern "C" int qp_batch_kernel_timing(const qp_batch* b, double ms[4]) {
    if (!b || !ms) return QP_ERR_BAD_ARG;
    ms[0] = b->ms[0];          // iNTT passes
    ms[1] = b->ms[1];          // LDE passes (+ salt)
    ms[2] = b->ms_leaf_hash;   // leaf_hash_kernel alone
    ms[3] = b->ms_tree_levels; // tree_level_kernel x num_layers
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// write_polynomial_batch / read_polynomial_batch (plonky2/src/util/serialization/mod.rs:1803-1822,
// 758-784; write_merkle_tree :1476-1491): the byte form CircuitData keeps its constants/sigmas
// commitment in.  All integers are u64 little-endian, field elements canonical, a bool one byte.
// ---------------------------------------------------------------------------------------------
extern "C" int qp_batch_describe(const qp_batch* b, uint64_t out[5]) {
    if (!b || !out) return QP_ERR_BAD_ARG;
    out[0] = b->n_cols;
    out[1] = b->degree_log;
    out[2] = b->rate_bits;
    out[3] = b->cap_height;
    out[4] = b->blinding ? 1 : 0;
    return QP_OK;
}

extern "C" size_t qp_batch_serialized_len(const qp_batch* b) {
    if (!b) return 0;
    const size_t n = (size_t)1 << b->degree_log, N = n << b->rate_bits;
    size_t words = 1 + b->n_cols * (1 + n);                      // polynomials
    words += 1 + N * (1 + b->leaf_len);                          // leaves
    words += 1 + b->tree.n_digests() * 4;                        // digests
    words += 1 + b->tree.n_cap() * 4;                            // cap height, cap
    words += 2;                                                  // degree_log, rate_bits
    return words * 8 + 1;                                        // + blinding
}

extern "C" int qp_batch_serialize(const qp_batch* b, uint8_t* out, size_t capacity) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (b->block_first != 0 || b->block_count != (1u << b->rate_bits))
        return fail(ctx, QP_ERR_BAD_ARG, "only a whole (unsharded) batch has the reference's byte form");
    if (!out || capacity < qp_batch_serialized_len(b)) return fail(ctx, QP_ERR_BAD_ARG, "output buffer too small");
    const size_t n = (size_t)1 << b->degree_log, N = n << b->rate_bits;
    uint8_t* w = out;
    auto put = [&](uint64_t x) {
        std::memcpy(w, &x, 8);
        w += 8;
    };
    std::vector<uint64_t> host(std::max(b->n_cols * n, std::max(b->tree.n_digests(), b->tree.n_cap()) * 4));
    int rc = qp_batch_coeffs(b, host.data(), QP_HOST);
    if (rc) return rc;
    put(b->n_cols);
    for (size_t c = 0; c < b->n_cols; c++) {
        put(n);
        std::memcpy(w, host.data() + c * n, n * 8);
        w += n * 8;
    }
    put(N);
    const size_t chunk = std::max<size_t>(1, ((size_t)1 << 22) / (b->leaf_len ? b->leaf_len : 1));  // ~32 MB of rows at a time
    std::vector<uint64_t> rows(std::min(chunk, N) * b->leaf_len);
    for (size_t first = 0; first < N; first += chunk) {
        const size_t cnt = std::min(chunk, N - first);
        if (b->leaf_len) {
            rc = qp_batch_leaves(b, first, cnt, rows.data(), QP_HOST);
            if (rc) return rc;
        }
        for (size_t i = 0; i < cnt; i++) {
            put(b->leaf_len);
            std::memcpy(w, rows.data() + i * b->leaf_len, b->leaf_len * 8);
            w += b->leaf_len * 8;
        }
    }
    put(b->tree.n_digests());
    if (b->tree.n_digests()) {
        rc = qp_batch_digests(b, host.data(), QP_HOST);
        if (rc) return rc;
        std::memcpy(w, host.data(), b->tree.n_digests() * 32);
        w += b->tree.n_digests() * 32;
    }
    put(b->cap_height);
    rc = qp_batch_cap(b, host.data(), QP_HOST);
    if (rc) return rc;
    std::memcpy(w, host.data(), b->tree.n_cap() * 32);
    w += b->tree.n_cap() * 32;
    put(b->degree_log);
    put(b->rate_bits);
    *w++ = b->blinding ? 1 : 0;
    return (size_t)(w - out) == qp_batch_serialized_len(b) ? QP_OK : fail(ctx, QP_ERR_BAD_ARG, "internal: length");
}

// The device batch is rebuilt from the polynomials (and the salt columns found in the leaves); the
// bytes' own cap must come out again -- the reference trusts the stored tree, this rejects a
// tree that does not belong to its polynomials.
extern "C" int qp_batch_deserialize(qp_ctx* ctx, const uint8_t* data, size_t len, qp_batch** out, size_t* consumed) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!data || !out) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    size_t pos = 0;
    bool ok = true;
    auto get = [&]() -> uint64_t {
        uint64_t x = 0;
        if (pos + 8 > len) ok = false;
        else std::memcpy(&x, data + pos, 8);
        pos += 8;
        return x;
    };
    auto truncated = [&]() { return fail(ctx, QP_ERR_BAD_ARG, "serialized PolynomialBatch is truncated or malformed"); };
    const uint64_t n_cols = get();
    if (!ok || n_cols > ((uint64_t)1 << 24)) return truncated();
    std::vector<uint64_t> coeffs;
    uint64_t n = 0;
    for (uint64_t c = 0; c < n_cols; c++) {
        const uint64_t plen = get();
        if (!ok || (c && plen != n) || plen > len / 8 || pos + plen * 8 > len) return truncated();
        n = plen;
        coeffs.resize((c + 1) * n);
        std::memcpy(coeffs.data() + c * n, data + pos, n * 8);
        pos += n * 8;
    }
    const uint64_t N = get();
    if (!ok || N > len / 8) return truncated();
    uint64_t leaf_len = 0;
    std::vector<uint64_t> salt_rows;  // [N][4] when the leaves are salted
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t l = get();
        if (!ok || (i && l != leaf_len) || l > len / 8 || pos + l * 8 > len) return truncated();
        leaf_len = l;
        if (l == n_cols + QP_SALT_SIZE) {
            if (salt_rows.empty()) salt_rows.resize(N * QP_SALT_SIZE);
            std::memcpy(salt_rows.data() + i * QP_SALT_SIZE, data + pos + n_cols * 8, QP_SALT_SIZE * 8);
        }
        pos += l * 8;
    }
    const uint64_t n_digests = get();
    if (!ok || n_digests > len / 32 || pos + n_digests * 32 > len) return truncated();
    pos += n_digests * 32;
    const uint64_t cap_height = get();
    if (!ok || cap_height > 32 || pos + (((size_t)1 << cap_height) * 32) > len) return truncated();
    const uint8_t* cap_bytes = data + pos;
    pos += ((size_t)1 << cap_height) * 32;
    const uint64_t degree_log = get(), rate_bits = get();
    if (!ok || pos + 1 > len) return truncated();
    const int blinding = data[pos++] ? 1 : 0;
    if (degree_log > 32 || rate_bits > 8 || n_cols == 0 || n != (uint64_t)1 << degree_log || N != n << rate_bits ||
        leaf_len != n_cols + (blinding ? QP_SALT_SIZE : 0) || cap_height > degree_log + rate_bits ||
        n_digests != 2 * (N - ((uint64_t)1 << cap_height)))
        return fail(ctx, QP_ERR_BAD_ARG, "serialized PolynomialBatch is inconsistent");
    std::vector<uint64_t> salt;
    if (blinding) {  // leaf i is the point bitrev(i): back to [QP_SALT_SIZE][N] in natural point order
        salt.resize((size_t)QP_SALT_SIZE * N);
        const unsigned bits = (unsigned)(degree_log + rate_bits);
        for (uint64_t i = 0; i < N; i++) {
            uint64_t nat = 0;
            for (unsigned k = 0; k < bits; k++) nat |= ((i >> k) & 1) << (bits - 1 - k);
            for (unsigned k = 0; k < QP_SALT_SIZE; k++) salt[(size_t)k * N + nat] = salt_rows[i * QP_SALT_SIZE + k];
        }
    }
    qp_batch* b = nullptr;
    int rc = qp_batch_from_coeffs(ctx, coeffs.data(), QP_HOST, n_cols, (unsigned)degree_log, (unsigned)rate_bits, blinding,
                                  (unsigned)cap_height, blinding ? salt.data() : nullptr, 0, 1u << rate_bits, &b);
    if (rc) return rc;
    std::vector<uint64_t> cap(((size_t)1 << cap_height) * 4);
    rc = qp_batch_cap(b, cap.data(), QP_HOST);
    if (!rc && std::memcmp(cap.data(), cap_bytes, cap.size() * 8) != 0)
        rc = fail(ctx, QP_ERR_BAD_ARG, "serialized Merkle cap does not belong to the serialized polynomials");
    if (rc) {
        qp_batch_free(b);
        return rc;
    }
    *out = b;
    if (consumed) *consumed = pos;
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// MerkleTree::new on caller rows
// ---------------------------------------------------------------------------------------------
struct qp_tree {
    qp_ctx* ctx = nullptr;
    uint64_t* leaves = nullptr;  // leaf-major [n][leaf_len] (device copy)
    size_t n_leaves = 0, leaf_len = 0;
    TreeBuf tree;
};

extern "C" int qp_merkle_tree_new(qp_ctx* ctx, const uint64_t* leaves, int space, size_t n_leaves,
                                  size_t leaf_len, unsigned cap_height, qp_tree** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!is_pow2(n_leaves)) return fail(ctx, QP_ERR_NOT_POW2, "Not a power of two");
    if (cap_height > ilog2(n_leaves))
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most log2(leaves.len())");
    if (!leaves && leaf_len) return fail(ctx, QP_ERR_BAD_ARG, "null leaves");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    qp_tree* t = new qp_tree();
    t->ctx = ctx;
    t->n_leaves = n_leaves;
    t->leaf_len = leaf_len;
    t->tree.shape.lg_leaves = ilog2(n_leaves);
    t->tree.shape.cap_height = cap_height;
    *out = t;
    int rc = dev_alloc(ctx, &t->leaves, n_leaves * leaf_len);
    merkle::AffineLayout lay{t->leaves, 1, leaf_len};
    const size_t bytes = n_leaves * leaf_len * 8;
    if (!rc && space != QP_DEVICE && bytes >= pipeline_threshold() && n_leaves >= 16) {
        // host rows of at least 64 MiB: upload in row slices on the copy stream and hash the leaves of a
        // slice while the next one is crossing PCIe (leaf rows are independent: hash_leaf per row)
        const int n_slices = 16;
        const size_t per = n_leaves / n_slices;
        cudaEventRecord(ctx->ready_ev, ctx->stream);
        cudaStreamWaitEvent(ctx->copy_stream, ctx->ready_ev, 0);
        for (int g = 0; g < n_slices && !rc; g++) {
            cudaError_t e = cudaMemcpyAsync(t->leaves + g * per * leaf_len, leaves + g * per * leaf_len, per * leaf_len * 8,
                                            cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_ev[g], ctx->copy_stream);
            if (e != cudaSuccess) rc = fail(ctx, QP_ERR_CUDA, cudaGetErrorString(e));
        }
        for (int g = 0; g < n_slices && !rc; g++) {
            cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[g], 0);
            rc = hash_leaves(ctx, lay, (unsigned)leaf_len, &t->tree, 0, ~0u, nullptr, g * per, per);
        }
        if (rc) cudaStreamSynchronize(ctx->copy_stream);   // nothing may still be writing into t->leaves when it is freed
        if (!rc) rc = build_tree_levels(ctx, &t->tree);
    } else {
        if (!rc && leaf_len)
            CUDA_TRY(ctx, cudaMemcpyAsync(t->leaves, leaves, bytes,
                                          space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                          ctx->stream));
        if (!rc) rc = build_tree(ctx, lay, (unsigned)leaf_len, &t->tree);
    }
    if (!rc) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (rc) {
        qp_tree_free(t);
        *out = nullptr;
    }
    return rc;
}

// MerkleTree::new, one shard of it: the reference parallelises over the 2^cap_height cap subtrees
// (fill_digests_buf, plonky2/src/hash/merkle_tree.rs:85-119) and so does the multi-GPU form -- shard s of S
// (a power of two <= 2^cap_height) is given the leaves of subtrees [s 2^h / S, (s + 1) 2^h / S) and returns
// exactly their block of `digests` and their cap entries; concatenated over the shards these are the
// reference's arrays.  (Whole subtrees form a Merkle tree of cap height h - log2 S: same kernels.)
extern "C" int qp_merkle_tree_new_shard(qp_ctx* ctx, const uint64_t* shard_leaves, int space, size_t n_leaves_total,
                                        size_t leaf_len, unsigned cap_height, unsigned shard, unsigned n_shards,
                                        qp_tree** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!is_pow2(n_leaves_total)) return fail(ctx, QP_ERR_NOT_POW2, "Not a power of two");
    if (cap_height > ilog2(n_leaves_total))
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most log2(leaves.len())");
    if (!is_pow2(n_shards) || shard >= n_shards || n_shards > ((size_t)1 << cap_height))
        return fail(ctx, QP_ERR_BAD_ARG, "shards must be a power of two of whole cap subtrees");
    return qp_merkle_tree_new(ctx, shard_leaves, space, n_leaves_total / n_shards, leaf_len, cap_height - ilog2(n_shards), out);
}

extern "C" void qp_tree_free(qp_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    dev_free(t->ctx, t->leaves);
    dev_free(t->ctx, t->tree.digests);
    dev_free(t->ctx, t->tree.cap);
    delete t;
}
extern "C" int qp_tree_cap(const qp_tree* t, uint64_t* out, int space) {
    if (!t) return QP_ERR_BAD_ARG;
    return copy_out(t->ctx, out, space, t->tree.cap, t->tree.n_cap() * 4);
}
extern "C" int qp_tree_digests(const qp_tree* t, uint64_t* out, int space) {
    if (!t) return QP_ERR_BAD_ARG;
    return copy_out(t->ctx, out, space, t->tree.digests, t->tree.n_digests() * 4);
}
extern "C" size_t qp_tree_digests_len(const qp_tree* t) { return t ? t->tree.n_digests() : 0; }
extern "C" int qp_tree_prove(const qp_tree* t, size_t leaf_index, uint64_t* siblings_out) {
    if (!t) return QP_ERR_BAD_ARG;
    return tree_prove(t->ctx, t->tree, leaf_index, siblings_out);
}
extern "C" int qp_tree_get(const qp_tree* t, size_t leaf_index, uint64_t* out) {
    if (!t) return QP_ERR_BAD_ARG;
    if (leaf_index >= t->n_leaves) return fail(t->ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    if (t->leaf_len == 0) return QP_OK;
    return copy_out(t->ctx, out, QP_HOST, t->leaves + leaf_index * t->leaf_len, t->leaf_len);
}

// ---------------------------------------------------------------------------------------------
// Poseidon / transforms
// ---------------------------------------------------------------------------------------------
extern "C" int qp_poseidon_permute(qp_ctx* ctx, uint64_t* states, int space, size_t count) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!states && count) return fail(ctx, QP_ERR_BAD_ARG, "null states");
    if (count == 0) return QP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t* d_in = nullptr;
    uint64_t* owned = nullptr;
    int rc = to_device(ctx, states, space, count * 12, &d_in, &owned);
    if (rc) return rc;
    uint64_t* d = const_cast<uint64_t*>(d_in);
    LAUNCH(ctx, permute_states_kernel, cdiv(count, 128), 128, 0, d, count);
    if (space != QP_DEVICE) rc = copy_out(ctx, states, QP_HOST, d, count * 12);
    dev_free(ctx, owned);
    return rc;
}

// coset FFT of device-resident vectors; output bit-reversed in `dst` (device).
// Tables of shift^i, i < 2^lg_n, cached per (size, shift): a FRI fold round then costs no allocation and no
// host round trip (the shifts of a proof are g^(arity^round): a handful of values per size).  shift = 1: none.
static int shift_tables(qp_ctx* ctx, unsigned lg_n, uint64_t shift, const ScaleTables** out) {
    *out = nullptr;
    if (gl::canon(shift) == 1) return QP_OK;
    std::vector<uint64_t> key = {(uint64_t)lg_n, gl::canon(shift), ~0ULL, ~0ULL};
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->scale_cache.find(key);
    if (it == ctx->scale_cache.end()) {
        ScaleTables t;
        int rc = build_scale(ctx, (int)lg_n, std::vector<uint64_t>{shift}, &t);
        if (rc) return rc;
        it = ctx->scale_cache.emplace(key, t).first;
    }
    *out = &it->second;
    return QP_OK;
}

static int coset_fft_device(qp_ctx* ctx, const uint64_t* d_src, uint64_t* d_dst, size_t n_vec, unsigned lg_n,
                            uint64_t shift) {
    const ScaleTables* st = nullptr;
    int rc_st = shift_tables(ctx, lg_n, shift, &st);
    if (rc_st) return rc_st;
    NttJob job;
    job.src = d_src;
    job.dst = d_dst;
    job.L = (int)lg_n;
    job.n_vec = (unsigned)n_vec;
    job.src_outer = job.dst_outer = (size_t)1 << lg_n;
    job.scale = st;
    return run_ntt(ctx, job);
}

extern "C" int qp_coset_fft(qp_ctx* ctx, const uint64_t* coeffs, int space, size_t n_vec, unsigned lg_n,
                            uint64_t shift, int bit_reversed, uint64_t* out, int out_space) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!coeffs || !out) return fail(ctx, QP_ERR_BAD_ARG, "null buffer");
    if (lg_n > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "transform larger than max_lde_log");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t words = n_vec << lg_n;
    const uint64_t* d_in = nullptr;
    uint64_t* in_owned = nullptr;
    int rc = to_device(ctx, coeffs, space, words, &d_in, &in_owned);
    if (rc) return rc;
    uint64_t* d_tmp = nullptr;
    rc = dev_alloc(ctx, &d_tmp, words);
    if (rc) return rc;
    rc = coset_fft_device(ctx, d_in, d_tmp, n_vec, lg_n, shift);
    uint64_t* d_res = d_tmp;
    uint64_t* d_nat = nullptr;
    if (!rc && !bit_reversed) {
        rc = dev_alloc(ctx, &d_nat, words);
        if (!rc) {
            LAUNCH(ctx, bitrev_permute_kernel, cdiv(words, 256), 256, 0, d_tmp, d_nat, lg_n, n_vec);
            d_res = d_nat;
        }
    }
    if (!rc) rc = copy_out(ctx, out, out_space, d_res, words);
    if (!rc && out_space == QP_DEVICE) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, in_owned);
    dev_free(ctx, d_tmp);
    dev_free(ctx, d_nat);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// FRI commit phase
// ---------------------------------------------------------------------------------------------
struct FriRound {
    unsigned arity_bits = 0;
    unsigned lg_n = 0;           // values in this round
    uint64_t* values = nullptr;  // planes [2][n], bit-reversed order (= leaf order)
    TreeBuf tree;
};

struct qp_fri {
    qp_ctx* ctx = nullptr;
    unsigned lg_n = 0, rate_bits = 0, cap_height = 0;
    unsigned cur_lg = 0;          // current number of coefficients (log2)
    uint64_t* coeffs = nullptr;   // planes [2][2^cur_lg], natural order
    uint64_t* values = nullptr;   // planes of the NEXT round to commit (bit-reversed)
    uint64_t shift = gl::GENERATOR;
    uint64_t* initial = nullptr;  // planes [2][2^initial_lg]: the unmasked final poly (from openings)
    unsigned initial_lg = 0;
    std::vector<FriRound> rounds;
    bool committed = false;       // a commit_round awaits its fold_round
};

extern "C" int qp_fri_begin(qp_ctx* ctx, const uint64_t* coeffs_ext, const uint64_t* values_ext, int space,
                            unsigned lg_n, unsigned rate_bits, unsigned cap_height, qp_fri** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!coeffs_ext || !values_ext) return fail(ctx, QP_ERR_BAD_ARG, "null input");
    if (lg_n > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "FRI domain larger than max_lde_log");
    if (rate_bits > lg_n) return fail(ctx, QP_ERR_BAD_ARG, "rate_bits > lde_bits");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << lg_n;
    qp_fri* f = new qp_fri();
    f->ctx = ctx;
    f->lg_n = f->cur_lg = lg_n;
    f->rate_bits = rate_bits;
    f->cap_height = cap_height;
    *out = f;
    const uint64_t *d_c = nullptr, *d_v = nullptr;
    uint64_t *c_owned = nullptr, *v_owned = nullptr;
    int rc = to_device(ctx, coeffs_ext, space, 2 * n, &d_c, &c_owned);
    if (!rc) rc = to_device(ctx, values_ext, space, 2 * n, &d_v, &v_owned);
    if (!rc) rc = dev_alloc(ctx, &f->coeffs, 2 * n);
    if (!rc) rc = dev_alloc(ctx, &f->values, 2 * n);
    if (!rc) {
        LAUNCH(ctx, fri::ext_to_planes_kernel, cdiv(n, 256), 256, 0, d_c, f->coeffs, lg_n, 0);
        // reverse_index_bits_in_place(values) (prover.rs:98)
        LAUNCH(ctx, fri::ext_to_planes_kernel, cdiv(n, 256), 256, 0, d_v, f->values, lg_n, 1);
    }
    dev_free(ctx, c_owned);
    dev_free(ctx, v_owned);
    if (rc) {
        qp_fri_free(f);
        *out = nullptr;
    }
    return rc;
}

// ---- host-side F_p^2 helpers (a handful of scalars per call; table seeds only) ----------------
namespace hx {
typedef unsigned __int128 u128;
struct E {
    uint64_t a, b;
};
static inline uint64_t fmul(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x * y) % gl::P); }
static inline uint64_t fadd(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x + y) % gl::P); }
static inline uint64_t fsub(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x + gl::P - y % gl::P) % gl::P); }
static inline E mul(E x, E y) {  // X^2 = 7
    return E{fadd(fmul(x.a, y.a), fmul(7, fmul(x.b, y.b))), fadd(fmul(x.a, y.b), fmul(x.b, y.a))};
}
static inline E inv(E x) {  // (a - bX) / (a^2 - 7 b^2)
    uint64_t nrm = fsub(fmul(x.a, x.a), fmul(7, fmul(x.b, x.b)));
    uint64_t ni = gl::host_pow(nrm, gl::P - 2);
    return E{fmul(x.a, ni), fmul(fsub(0, x.b), ni)};
}
}  // namespace hx

// planes [2][2^lg_n] of z^i on the device
static int build_power_table(qp_ctx* ctx, hx::E z, unsigned lg_n, uint64_t** out) {
    uint64_t sq[64];
    hx::E cur = z;
    for (int b = 0; b < 32; b++) {
        sq[2 * b] = cur.a;
        sq[2 * b + 1] = cur.b;
        cur = hx::mul(cur, cur);
    }
    uint64_t* d_sq = nullptr;
    int rc = dev_alloc(ctx, &d_sq, 64);
    if (!rc) rc = dev_alloc(ctx, out, (size_t)2 << lg_n);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_sq, sq, sizeof sq, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // sq is a stack array
    LAUNCH(ctx, openings::power_table_kernel, cdiv((size_t)1 << lg_n, 256), 256, 0, d_sq, lg_n, *out);
    dev_free(ctx, d_sq);
    return QP_OK;
}

extern "C" int qp_batch_eval_polys(const qp_batch* b, const uint64_t point[2], uint64_t* out) {
    if (!b) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = b->ctx;
    if (!point || !out) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    uint64_t* pw = nullptr;
    uint64_t* d_out = nullptr;
    int rc = build_power_table(ctx, hx::E{point[0] % gl::P, point[1] % gl::P}, b->degree_log, &pw);
    // segments per polynomial: about eight resident blocks per SM in total, at least 1024 coefficients each
    const size_t n = (size_t)1 << b->degree_log;
    unsigned n_seg = (unsigned)cdiv((size_t)ctx->sm_count * 8, b->n_cols ? b->n_cols : 1);
    if (n_seg > n / 1024) n_seg = (unsigned)(n / 1024);
    if (n_seg < 1) n_seg = 1;
    uint64_t* d_partial = nullptr;
    if (!rc) rc = dev_alloc(ctx, &d_out, 2 * b->n_cols);
    if (!rc) rc = dev_alloc(ctx, &d_partial, 2 * b->n_cols * n_seg);
    if (!rc) {
        LAUNCH(ctx, openings::eval_polys_kernel, dim3((unsigned)b->n_cols, n_seg), 256, 0, b->coeffs, n, pw, d_partial);
        LAUNCH(ctx, openings::eval_polys_finish_kernel, cdiv(b->n_cols, 128), 128, 0, d_partial, (unsigned)b->n_cols, n_seg, d_out);
        rc = copy_out(ctx, out, QP_HOST, d_out, 2 * b->n_cols);
    }
    dev_free(ctx, pw);
    dev_free(ctx, d_out);
    dev_free(ctx, d_partial);
    return rc;
}

// values <- coset LDE (shift g) of the coefficient planes, in bit-reversed order; also keeps the
// zero-padded coefficient planes the fold kernel reads.
static int fri_from_device_planes(qp_ctx* ctx, uint64_t* fin /* [2][n], consumed */, unsigned degree_log,
                                  unsigned rate_bits, unsigned cap_height, qp_fri** out) {
    const size_t n = (size_t)1 << degree_log, N = n << rate_bits;
    qp_fri* f = new qp_fri();
    f->ctx = ctx;
    f->lg_n = f->cur_lg = degree_log + rate_bits;
    f->rate_bits = rate_bits;
    f->cap_height = cap_height;
    f->initial = fin;
    f->initial_lg = degree_log;
    *out = f;
    int rc = dev_alloc(ctx, &f->coeffs, 2 * N);
    if (!rc) rc = dev_alloc(ctx, &f->values, 2 * N);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(f->coeffs, 0, 2 * N * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(f->coeffs, fin, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(f->coeffs + N, fin + n, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    const ScaleTables* st = nullptr;
    rc = lde_scale(ctx, (int)degree_log, rate_bits, 0, 1u << rate_bits, &st);
    if (rc) return rc;
    NttJob job;
    job.src = fin;
    job.dst = f->values;
    job.L = (int)degree_log;
    job.n_vec = 2u << rate_bits;
    job.inner_bits = (int)rate_bits;
    job.src_outer = n;
    job.src_inner = 0;
    job.dst_outer = N;
    job.dst_inner = n;
    job.scale = st;
    return run_ntt(ctx, job);
}

extern "C" int qp_fri_begin_from_openings(qp_ctx* ctx, const qp_opening_batch* batches, size_t n_batches,
                                          unsigned degree_log, unsigned rate_bits, unsigned cap_height,
                                          qp_fri** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!batches || n_batches == 0) return fail(ctx, QP_ERR_BAD_ARG, "no opening batches");
    if (degree_log + rate_bits > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "FRI domain larger than max_lde_log");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << degree_log;
    const size_t n_seg = (n + openings::SCAN_SEG - 1) / openings::SCAN_SEG;
    uint64_t *fin = nullptr, *d = nullptr, *totals = nullptr;
    int rc = dev_alloc(ctx, &fin, 2 * n);
    if (!rc) rc = dev_alloc(ctx, &d, 2 * n);
    if (!rc) rc = dev_alloc(ctx, &totals, 2 * n_seg);
    for (size_t bi = 0; bi < n_batches && !rc; bi++) {
        const qp_opening_batch& ob = batches[bi];
        if (!ob.terms && ob.n_terms) rc = fail(ctx, QP_ERR_BAD_ARG, "null terms");
        std::vector<const uint64_t*> ptrs(ob.n_terms);
        std::vector<uint64_t> w(2 * ob.n_terms);
        for (size_t t = 0; t < ob.n_terms && !rc; t++) {
            const qp_batch* pb = ob.terms[t].batch;
            if (!pb || pb->ctx != ctx || pb->degree_log != degree_log || ob.terms[t].poly_index >= pb->n_cols)
                rc = fail(ctx, QP_ERR_DEGREE_MISMATCH, "opening term: bad oracle / polynomial index / degree");
            else {
                ptrs[t] = pb->coeffs + ob.terms[t].poly_index * n;
                w[2 * t] = ob.terms[t].weight[0] % gl::P;
                w[2 * t + 1] = ob.terms[t].weight[1] % gl::P;
            }
        }
        if (rc) break;
        uint64_t *d_ptrs = nullptr, *d_w = nullptr, *pw = nullptr, *ipw = nullptr;
        rc = dev_alloc(ctx, &d_ptrs, ob.n_terms ? ob.n_terms : 1);
        if (!rc) rc = dev_alloc(ctx, &d_w, ob.n_terms ? 2 * ob.n_terms : 1);
        if (!rc && ob.n_terms) {
            cudaMemcpyAsync(d_ptrs, ptrs.data(), ob.n_terms * 8, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(d_w, w.data(), 2 * ob.n_terms * 8, cudaMemcpyHostToDevice, ctx->stream);
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die at the end of the iteration
        }
        const hx::E z{ob.point[0] % gl::P, ob.point[1] % gl::P};
        const bool z_zero = z.a == 0 && z.b == 0;
        if (!rc && !z_zero) rc = build_power_table(ctx, z, degree_log, &pw);
        if (!rc && !z_zero) rc = build_power_table(ctx, hx::inv(z), degree_log, &ipw);
        if (!rc) {
            LAUNCH(ctx, openings::weighted_sum_kernel, cdiv(n, 256), 256, 0, (const uint64_t* const*)d_ptrs, d_w,
                   (unsigned)ob.n_terms, n, pw, d);
            const uint64_t s0 = ob.shift[0] % gl::P, s1 = ob.shift[1] % gl::P;
            if (z_zero) {
                LAUNCH(ctx, openings::shift_accumulate_kernel, cdiv(n, 256), 256, 0, d, n, s0, s1, bi == 0, fin);
            } else {
                LAUNCH(ctx, openings::suffix_scan_segments_kernel, (unsigned)n_seg, 256, 0, d, n, totals);
                LAUNCH(ctx, openings::suffix_scan_totals_kernel, 1, 32, 0, totals, n_seg);
                LAUNCH(ctx, openings::quotient_accumulate_kernel, cdiv(n, 256), 256, 0, d, totals, n, ipw, s0, s1,
                       bi == 0, fin);
            }
        }
        dev_free(ctx, d_ptrs);
        dev_free(ctx, d_w);
        dev_free(ctx, pw);
        dev_free(ctx, ipw);
    }
    dev_free(ctx, d);
    dev_free(ctx, totals);
    if (rc) {
        dev_free(ctx, fin);
        return rc;
    }
    rc = fri_from_device_planes(ctx, fin, degree_log, rate_bits, cap_height, out);
    if (rc) {
        qp_fri_free(*out);
        *out = nullptr;
    }
    return rc;
}

extern "C" int qp_fri_initial_coeffs(const qp_fri* f, uint64_t* out) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (!f->initial) return fail(ctx, QP_ERR_BAD_ARG, "FRI state was not built from openings");
    const size_t n = (size_t)1 << f->initial_lg;
    uint64_t* dtmp = nullptr;
    int rc = dev_alloc(ctx, &dtmp, 2 * n);
    if (rc) return rc;
    LAUNCH(ctx, fri::planes_to_ext_kernel, cdiv(n, 256), 256, 0, f->initial, n, (size_t)0, n, dtmp);
    rc = copy_out(ctx, out, QP_HOST, dtmp, 2 * n);
    dev_free(ctx, dtmp);
    return rc;
}

extern "C" int qp_fri_commit_round(qp_fri* f, unsigned arity_bits, uint64_t* cap_out) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (f->committed || !f->values) return fail(ctx, QP_ERR_BAD_ARG, "commit_round called out of order");
    if (arity_bits > f->cur_lg) return fail(ctx, QP_ERR_BAD_ARG, "arity exceeds domain");
    const unsigned lg_leaves = f->cur_lg - arity_bits;
    if (f->cap_height > lg_leaves)
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most log2(leaves.len())");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FriRound r;
    r.arity_bits = arity_bits;
    r.lg_n = f->cur_lg;
    r.values = f->values;
    f->values = nullptr;
    r.tree.shape.lg_leaves = lg_leaves;
    r.tree.shape.cap_height = f->cap_height;
    // leaves = chunks of `arity` consecutive (bit-reversed) values, flattened (prover.rs:99-104)
    merkle::ExtPlanesLayout lay{r.values, (size_t)1 << r.lg_n, arity_bits};
    int rc = build_tree(ctx, lay, 2u << arity_bits, &r.tree);
    f->rounds.push_back(r);
    if (rc) return rc;
    f->committed = true;
    return copy_out(ctx, cap_out, QP_HOST, r.tree.cap, r.tree.n_cap() * 4);
}

extern "C" int qp_fri_fold_round(qp_fri* f, const uint64_t beta[2], int is_last) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (!f->committed || !beta) return fail(ctx, QP_ERR_BAD_ARG, "fold_round called out of order");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const unsigned ab = f->rounds.back().arity_bits;
    const size_t n_in = (size_t)1 << f->cur_lg, n_out = n_in >> ab;
    uint64_t* folded = nullptr;
    int rc = dev_alloc(ctx, &folded, 2 * n_out);
    if (rc) return rc;
    LAUNCH(ctx, fri::fold_kernel, cdiv(n_out, 128), 128, 0, f->coeffs, n_in, ab, beta[0], beta[1], folded);
    dev_free(ctx, f->coeffs);
    f->coeffs = folded;
    f->cur_lg -= ab;
    f->committed = false;
    if (is_last) return QP_OK;
    // shift <- shift^arity ; values <- coset_fft(coeffs, shift) (prover.rs:121-122), kept bit-reversed
    f->shift = gl::host_pow(f->shift, (uint64_t)1 << ab);
    rc = dev_alloc(ctx, &f->values, 2 * n_out);
    if (rc) return rc;
    return coset_fft_device(ctx, f->coeffs, f->values, 2, f->cur_lg, f->shift);
}

// Batch FRI (plonky2/src/batch_fri/prover.rs:122-141): after a fold round that is not the last, if the next
// polynomial of the batch has exactly the current length,
//     final_values = final_values * beta + values[polynomial_index];  final_coeffs = final_values.coset_ifft(shift)
// `lower` is the FRI state of that polynomial (qp_fri_begin / qp_fri_begin_from_openings), untouched so far: its
// `values` are the LDE values on a domain of the same size, in the same bit-reversed order as ours.
extern "C" int qp_fri_mix_values(qp_fri* f, const qp_fri* lower, const uint64_t beta[2]) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (!lower || !beta) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    if (lower->ctx != ctx) return fail(ctx, QP_ERR_BAD_ARG, "FRI state belongs to another context");
    if (f->committed || !f->values) return fail(ctx, QP_ERR_BAD_ARG, "mix_values called out of order");
    if (!lower->values || !lower->rounds.empty() || lower->cur_lg != f->cur_lg)
        return fail(ctx, QP_ERR_DEGREE_MISMATCH, "the polynomial to absorb must have the current length and no rounds");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const unsigned lg = f->cur_lg;
    const size_t n = (size_t)1 << lg;
    LAUNCH(ctx, fri::mix_kernel, cdiv(n, 256), 256, 0, f->values, lower->values, n, beta[0] % gl::P, beta[1] % gl::P);
    // coset_ifft (polynomial/mod.rs:58-88): natural-order values -> inverse transform -> coefficient i times shift^-i
    TempScope tmp(ctx);
    uint64_t* nat = nullptr;
    int rc = tmp.alloc(&nat, 2 * n);
    if (rc) return rc;
    LAUNCH(ctx, bitrev_permute_kernel, cdiv(2 * n, 256), 256, 0, f->values, nat, lg, (size_t)2);
    NttJob job;
    job.src = nat;
    job.dst = f->coeffs;
    job.L = (int)lg;
    job.n_vec = 2;
    job.inner_bits = 0;
    job.src_outer = n;
    job.dst_outer = n;
    job.out_mode = ntt::OUT_INVERSE;
    job.scratch = nat;
    rc = run_ntt(ctx, job);
    if (rc) return rc;
    const ScaleTables* st = nullptr;
    rc = shift_tables(ctx, lg, gl::host_pow(f->shift, gl::P - 2), &st);
    if (rc) return rc;
    if (st)
        LAUNCH(ctx, quotient::scale_powers_kernel, cdiv(2 * n, 256), 256, 0, f->coeffs, (size_t)2, lg, st->lo, st->hi,
               st->split);
    return QP_OK;
}

/* log2 of the current codeword length (after the folds done so far) */
extern "C" unsigned qp_fri_domain_bits(const qp_fri* f) { return f ? f->cur_lg : 0; }

extern "C" int qp_fri_final_poly(qp_fri* f, uint64_t* out, size_t* len_out) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (f->cur_lg < f->rate_bits) return fail(ctx, QP_ERR_BAD_ARG, "final polynomial shorter than the rate");
    const size_t len = ((size_t)1 << f->cur_lg) >> f->rate_bits;
    if (len_out) *len_out = len;
    if (!out) return QP_OK;
    uint64_t* d = nullptr;
    int rc = dev_alloc(ctx, &d, 2 * len);
    if (rc) return rc;
    LAUNCH(ctx, fri::planes_to_ext_kernel, cdiv(len, 256), 256, 0, f->coeffs, (size_t)1 << f->cur_lg, (size_t)0,
           len, d);
    rc = copy_out(ctx, out, QP_HOST, d, 2 * len);
    dev_free(ctx, d);
    return rc;
}

extern "C" unsigned qp_fri_num_rounds(const qp_fri* f) { return f ? (unsigned)f->rounds.size() : 0; }

extern "C" int qp_fri_tree_get(const qp_fri* f, unsigned round, size_t leaf_index, uint64_t* out) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (round >= f->rounds.size()) return fail(ctx, QP_ERR_BAD_ARG, "round out of range");
    const FriRound& r = f->rounds[round];
    if (leaf_index >> r.tree.shape.lg_leaves) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned row_len = 2u << r.arity_bits;
    uint64_t* d = nullptr;
    int rc = dev_alloc(ctx, &d, row_len);
    if (rc) return rc;
    merkle::ExtPlanesLayout lay{r.values, (size_t)1 << r.lg_n, r.arity_bits};
    LAUNCH(ctx, merkle::gather_rows_kernel<merkle::ExtPlanesLayout>, 1, 64, 0, lay, row_len,
           (const uint64_t*)nullptr, leaf_index, (size_t)1, d);
    rc = copy_out(ctx, out, QP_HOST, d, row_len);
    dev_free(ctx, d);
    return rc;
}

extern "C" int qp_fri_tree_prove(const qp_fri* f, unsigned round, size_t leaf_index, uint64_t* siblings_out) {
    if (!f) return QP_ERR_BAD_ARG;
    if (round >= f->rounds.size()) return fail(f->ctx, QP_ERR_BAD_ARG, "round out of range");
    return tree_prove(f->ctx, f->rounds[round].tree, leaf_index, siblings_out);
}
extern "C" int qp_fri_tree_open_many(const qp_fri* f, unsigned round, const uint64_t* leaf_indices, unsigned n,
                                     uint64_t* leaves_out, uint64_t* siblings_out) {
    if (!f) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = f->ctx;
    if (round >= f->rounds.size()) return fail(ctx, QP_ERR_BAD_ARG, "round out of range");
    if (n == 0) return QP_OK;
    if (!leaf_indices || !leaves_out) return fail(ctx, QP_ERR_BAD_ARG, "null buffer");
    const FriRound& r = f->rounds[round];
    for (unsigned i = 0; i < n; i++)
        if (leaf_indices[i] >> r.tree.shape.lg_leaves) return fail(ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned row_len = 2u << r.arity_bits;
    uint64_t *d_idx = nullptr, *d_rows = nullptr;
    int rc = dev_alloc(ctx, &d_idx, n);
    if (!rc) rc = dev_alloc(ctx, &d_rows, (size_t)n * row_len);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_idx, leaf_indices, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    merkle::ExtPlanesLayout lay{r.values, (size_t)1 << r.lg_n, r.arity_bits};
    LAUNCH(ctx, merkle::gather_rows_kernel<merkle::ExtPlanesLayout>, cdiv((size_t)n * row_len, 128), 128, 0, lay,
           row_len, (const uint64_t*)d_idx, (size_t)0, (size_t)n, d_rows);
    rc = copy_out(ctx, leaves_out, QP_HOST, d_rows, (size_t)n * row_len);
    dev_free(ctx, d_idx);
    dev_free(ctx, d_rows);
    if (!rc) rc = tree_prove_many(ctx, r.tree, leaf_indices, n, siblings_out);
    return rc;
}

extern "C" int qp_fri_tree_digests(const qp_fri* f, unsigned round, uint64_t* out, int space) {
    if (!f) return QP_ERR_BAD_ARG;
    if (round >= f->rounds.size()) return fail(f->ctx, QP_ERR_BAD_ARG, "round out of range");
    const TreeBuf& t = f->rounds[round].tree;
    return copy_out(f->ctx, out, space, t.digests, t.n_digests() * 4);
}
extern "C" size_t qp_fri_tree_digests_len(const qp_fri* f, unsigned round) {
    return (f && round < f->rounds.size()) ? f->rounds[round].tree.n_digests() : 0;
}

extern "C" void qp_fri_free(qp_fri* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    dev_free(f->ctx, f->coeffs);
    dev_free(f->ctx, f->values);
    dev_free(f->ctx, f->initial);
    for (auto& r : f->rounds) {
        dev_free(f->ctx, r.values);
        dev_free(f->ctx, r.tree.digests);
        dev_free(f->ctx, r.tree.cap);
    }
    delete f;
}

extern "C" int qp_fri_proof_of_work(qp_ctx* ctx, const uint64_t state12[12], unsigned witness_pos,
                                    unsigned min_leading_zeros, uint64_t* witness_out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!state12 || !witness_out || witness_pos >= 12) return fail(ctx, QP_ERR_BAD_ARG, "bad PoW arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    uint64_t* d_state = nullptr;
    uint64_t* d_found = nullptr;
    int rc = dev_alloc(ctx, &d_state, 12);
    if (!rc) rc = dev_alloc(ctx, &d_found, 1);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_state, state12, 96, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_found, 0xff, 8, ctx->stream));
    // Batches in increasing candidate order; the first batch with a hit contains the global
    // minimum.  A batch is 8 times the expected work 2^min_leading_zeros (a miss has probability
    // e^-8; at least 2^17 = one wave): the kernel's blocks stop once a smaller witness exists, but
    // every launched block still costs a few nanoseconds to retire (2^22 candidates: 0.14 ms).
    const unsigned lg_batch = min_leading_zeros + 3 < 17 ? 17 : (min_leading_zeros + 3 > 26 ? 26 : min_leading_zeros + 3);
    uint64_t batch = (uint64_t)1 << lg_batch;
    uint64_t base = 0;
    uint64_t found = UINT64_MAX;
    while (true) {
        const uint64_t remaining = gl::P - base;
        const uint64_t cnt = batch < remaining ? batch : remaining;
        LAUNCH(ctx, fri::pow_kernel, cdiv(cnt, 128), 128, 0, d_state, witness_pos, min_leading_zeros, base,
               (unsigned long long*)d_found);
        CUDA_TRY(ctx, cudaMemcpyAsync(&found, d_found, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (found != UINT64_MAX) break;
        base += cnt;
        if (base >= gl::P) break;
        if (batch < ((uint64_t)1 << 26)) batch <<= 2;
    }
    dev_free(ctx, d_state);
    dev_free(ctx, d_found);
    if (found == UINT64_MAX) return fail(ctx, QP_ERR_BAD_ARG, "Proof of work failed. This is highly unlikely!");
    *witness_out = found;
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// BatchMerkleTree::new (plonky2/src/hash/batch_merkle_tree.rs): a chain of trees, each capped at
// the height of the next matrix
// ---------------------------------------------------------------------------------------------
struct qp_batch_tree {
    qp_ctx* ctx = nullptr;
    std::vector<uint64_t*> rows;     // the caller's matrices on the device, row-major
    std::vector<size_t> heights, widths;
    std::vector<TreeBuf> stages;     // stage j: leaves = matrix j (|| cap of stage j-1), cap at height j+1
    uint64_t* scratch = nullptr;     // a stage's concatenated leaves while it is being built
    unsigned cap_height = 0;
};

extern "C" void qp_batch_tree_free(qp_batch_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    for (uint64_t* p : t->rows) dev_free(t->ctx, p);
    dev_free(t->ctx, t->scratch);
    for (TreeBuf& s : t->stages) {
        dev_free(t->ctx, s.digests);
        dev_free(t->ctx, s.cap);
    }
    delete t;
}

static int batch_tree_build(qp_ctx* ctx, qp_batch_tree* t, const uint64_t* const* matrices, int space) {
    const size_t n_matrices = t->heights.size();
    for (size_t j = 0; j < n_matrices; j++) {
        const size_t h = t->heights[j], w = t->widths[j];
        int rc = dev_alloc(ctx, &t->rows[j], h * w ? h * w : 1);
        if (rc) return rc;
        if (h * w)
            CUDA_TRY(ctx, cudaMemcpyAsync(t->rows[j], matrices[j], h * w * 8,
                                          space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                          ctx->stream));
        TreeBuf& st = t->stages[j];
        st.shape.lg_leaves = ilog2(h);
        st.shape.cap_height = j + 1 < n_matrices ? ilog2(t->heights[j + 1]) : t->cap_height;
        if (j == 0) {
            merkle::AffineLayout lay{t->rows[0], 1, w};
            rc = build_tree(ctx, lay, (unsigned)w, &st);
            if (rc) return rc;
            continue;
        }
        // new_leaves[i] = cap_hash[i] || cur[i], batch_merkle_tree.rs:91-100
        rc = dev_alloc(ctx, &t->scratch, h * (w + 4));
        if (rc) return rc;
        LAUNCH(ctx, merkle::concat_cap_rows_kernel, cdiv(h * (w + 4), 256), 256, 0, t->stages[j - 1].cap, t->rows[j], h, w,
               t->scratch);
        merkle::AffineLayout lay{t->scratch, 1, w + 4};
        rc = build_tree(ctx, lay, (unsigned)(w + 4), &st);
        if (rc) return rc;
        dev_free(ctx, t->scratch);  // stream-ordered: released after the kernels above
        t->scratch = nullptr;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}

extern "C" int qp_batch_merkle_tree_new(qp_ctx* ctx, const uint64_t* const* matrices, int space, const size_t* heights,
                                        const size_t* widths, size_t n_matrices, unsigned cap_height,
                                        qp_batch_tree** out) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!n_matrices || !matrices || !heights || !widths) return fail(ctx, QP_ERR_BAD_ARG, "no leaves");
    for (size_t j = 0; j < n_matrices; j++) {
        if (!is_pow2(heights[j])) return fail(ctx, QP_ERR_NOT_POW2, "Not a power of two");
        if (j && heights[j - 1] <= heights[j])
            return fail(ctx, QP_ERR_BAD_ARG, "leaves must be sorted by height, tallest first, no duplicates");
        if (!matrices[j] && widths[j]) return fail(ctx, QP_ERR_BAD_ARG, "null matrix");
    }
    if (cap_height > ilog2(heights[n_matrices - 1]))
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most last_leaves_cap_height");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    qp_batch_tree* t = new qp_batch_tree();
    t->ctx = ctx;
    t->cap_height = cap_height;
    t->heights.assign(heights, heights + n_matrices);
    t->widths.assign(widths, widths + n_matrices);
    t->rows.assign(n_matrices, nullptr);
    t->stages.assign(n_matrices, TreeBuf{});
    const int rc = batch_tree_build(ctx, t, matrices, space);
    if (rc) {
        qp_batch_tree_free(t);
        return rc;
    }
    *out = t;
    return QP_OK;
}

extern "C" int qp_batch_tree_cap(const qp_batch_tree* t, uint64_t* out, int space) {
    if (!t) return QP_ERR_BAD_ARG;
    return copy_out(t->ctx, out, space, t->stages.back().cap, t->stages.back().n_cap() * 4);
}
extern "C" size_t qp_batch_tree_digests_len(const qp_batch_tree* t) {
    size_t n = 0;
    if (t)
        for (const TreeBuf& s : t->stages) n += s.n_digests();
    return n;
}
extern "C" int qp_batch_tree_digests(const qp_batch_tree* t, uint64_t* out, int space) {
    if (!t) return QP_ERR_BAD_ARG;
    size_t pos = 0;
    for (const TreeBuf& s : t->stages) {
        int rc = copy_out(t->ctx, out + pos, space, s.digests, s.n_digests() * 4);
        if (rc) return rc;
        pos += s.n_digests() * 4;
    }
    return QP_OK;
}
extern "C" int qp_batch_tree_open(const qp_batch_tree* t, size_t leaf_index, uint64_t* siblings_out) {
    if (!t) return QP_ERR_BAD_ARG;
    if (leaf_index >= t->heights[0]) return fail(t->ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned lg0 = ilog2(t->heights[0]);
    size_t pos = 0;
    for (const TreeBuf& s : t->stages) {  // batch_merkle_tree.rs:139-150
        int rc = tree_prove(t->ctx, s, leaf_index >> (lg0 - s.shape.lg_leaves), siblings_out + pos);
        if (rc) return rc;
        pos += 4 * (size_t)s.shape.num_layers();
    }
    return QP_OK;
}
extern "C" int qp_batch_tree_values(const qp_batch_tree* t, size_t leaf_index, uint64_t* out) {
    if (!t) return QP_ERR_BAD_ARG;
    if (leaf_index >= t->heights[0]) return fail(t->ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    const unsigned lg0 = ilog2(t->heights[0]);
    size_t pos = 0;
    for (size_t j = 0; j < t->rows.size(); j++) {
        const size_t row = leaf_index >> (lg0 - ilog2(t->heights[j])), w = t->widths[j];
        if (w) {
            int rc = copy_out(t->ctx, out + pos, QP_HOST, t->rows[j] + row * w, w);
            if (rc) return rc;
        }
        pos += w;
    }
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// BatchFriOracle::from_values / from_coeffs (plonky2/src/batch_fri/oracle.rs:78-160): one LDE per
// run of equal-length polynomials, one BatchMerkleTree over the runs.  Group 0 is an ordinary
// batch whose tree is capped at the next group's height; every later group hashes its LDE rows
// together with the cap below it (merkle::CapPrefixLayout) -- no leaf-major copy is ever made.
// ---------------------------------------------------------------------------------------------
struct qp_batch_fri {
    qp_ctx* ctx = nullptr;
    std::vector<qp_batch*> groups;  // coefficients, LDE and the stage tree of every group
    unsigned rate_bits = 0, cap_height = 0;
};

/* group g as a PolynomialBatch view (coefficients, LDE rows); owned by the oracle */
extern "C" const qp_batch* qp_batch_fri_group_batch(const qp_batch_fri* o, size_t g) {
    return (o && g < o->groups.size()) ? o->groups[g] : nullptr;
}

extern "C" void qp_batch_fri_free(qp_batch_fri* o) {
    if (!o) return;
    for (qp_batch* b : o->groups) qp_batch_free(b);
    delete o;
}

static int batch_fri_build(qp_ctx* ctx, qp_batch_fri* o, const uint64_t* const* polys, const unsigned* degree_bits,
                           size_t n_polys, int space, bool is_values) {
    size_t start = 0;
    for (size_t i = 0; i < n_polys; i++) {
        if (i + 1 < n_polys && degree_bits[i] == degree_bits[i + 1]) continue;
        const size_t cols = i + 1 - start;
        const unsigned d = degree_bits[start];
        const size_t n = (size_t)1 << d;
        // next group's height decides where this stage's tree stops (batch_merkle_tree.rs:63-70)
        const unsigned stage_cap = i + 1 < n_polys ? degree_bits[i + 1] + o->rate_bits : o->cap_height;
        uint64_t* d_coeffs = nullptr;
        int rc = dev_alloc(ctx, &d_coeffs, cols * n);
        if (rc) return rc;
        for (size_t c = 0; c < cols; c++)
            CUDA_TRY(ctx, cudaMemcpyAsync(d_coeffs + c * n, polys[start + c], n * 8,
                                          space == QP_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                          ctx->stream));
        if (is_values) {  // "IFFT", oracle.rs:86-90 (out of place: the inverse transform permutes across CTAs)
            uint64_t* d_vals = d_coeffs;
            rc = dev_alloc(ctx, &d_coeffs, cols * n);
            if (!rc) rc = qp_ifft_columns(ctx, d_vals, QP_DEVICE, cols, d, d_coeffs, QP_DEVICE);
            dev_free(ctx, d_vals);
            if (rc) {
                dev_free(ctx, d_coeffs);
                return rc;
            }
        }
        qp_batch* b = nullptr;
        if (o->groups.empty()) {
            rc = batch_from_device_coeffs(ctx, d_coeffs, cols, d, o->rate_bits, 0, stage_cap, nullptr, 0, 1u << o->rate_bits, &b);
            if (b) o->groups.push_back(b);
            if (rc) return rc;
        } else {
            rc = batch_create(ctx, d_coeffs, cols, d, o->rate_bits, 0, stage_cap, 0, 1u << o->rate_bits, &b);
            if (b) o->groups.push_back(b);
            if (!rc) rc = batch_lde_columns(b, 0, cols);
            if (rc) return rc;
            const qp_batch* below = o->groups[o->groups.size() - 2];
            merkle::CapPrefixLayout lay{below->tree.cap, b->lde, b->n_local};
            rc = hash_leaves(ctx, lay, (unsigned)(cols + 4), &b->tree);
            if (!rc) rc = build_tree_levels(ctx, &b->tree);
            if (rc) return rc;
        }
        start = i + 1;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}

static int batch_fri_new(qp_ctx* ctx, const uint64_t* const* polys, const unsigned* degree_bits, size_t n_polys, int space,
                         unsigned rate_bits, int blinding, unsigned cap_height, qp_batch_fri** out, bool is_values) {
    if (!ctx) return QP_ERR_BAD_ARG;
    if (!out) return fail(ctx, QP_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!n_polys || !polys || !degree_bits) return fail(ctx, QP_ERR_BAD_ARG, "no polynomials");
    if (blinding) return fail(ctx, QP_ERR_BAD_ARG, "salt injection is not implemented for batch oracles");
    for (size_t i = 0; i < n_polys; i++) {
        if (!polys[i]) return fail(ctx, QP_ERR_BAD_ARG, "null polynomial");
        if (i && degree_bits[i - 1] < degree_bits[i])  // oracle.rs:118
            return fail(ctx, QP_ERR_BAD_ARG, "polynomials must be sorted by degree, largest first");
        if (degree_bits[i] + rate_bits > ctx->tw_lg) return fail(ctx, QP_ERR_TOO_LARGE, "LDE larger than the context's max_lde_log");
    }
    if (cap_height > degree_bits[n_polys - 1] + rate_bits)
        return fail(ctx, QP_ERR_CAP_HEIGHT, "cap_height should be at most last_leaves_cap_height");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    qp_batch_fri* o = new qp_batch_fri();
    o->ctx = ctx;
    o->rate_bits = rate_bits;
    o->cap_height = cap_height;
    const int rc = batch_fri_build(ctx, o, polys, degree_bits, n_polys, space, is_values);
    if (rc) {
        qp_batch_fri_free(o);
        return rc;
    }
    *out = o;
    return QP_OK;
}

extern "C" int qp_batch_fri_from_values(qp_ctx* ctx, const uint64_t* const* polys, const unsigned* degree_bits,
                                        size_t n_polys, int space, unsigned rate_bits, int blinding, unsigned cap_height,
                                        qp_batch_fri** out) {
    return batch_fri_new(ctx, polys, degree_bits, n_polys, space, rate_bits, blinding, cap_height, out, true);
}
extern "C" int qp_batch_fri_from_coeffs(qp_ctx* ctx, const uint64_t* const* polys, const unsigned* degree_bits,
                                        size_t n_polys, int space, unsigned rate_bits, int blinding, unsigned cap_height,
                                        qp_batch_fri** out) {
    return batch_fri_new(ctx, polys, degree_bits, n_polys, space, rate_bits, blinding, cap_height, out, false);
}
extern "C" size_t qp_batch_fri_num_groups(const qp_batch_fri* o) { return o ? o->groups.size() : 0; }
extern "C" int qp_batch_fri_group(const qp_batch_fri* o, size_t g, unsigned* degree_bits, size_t* n_polys) {
    if (!o || g >= o->groups.size()) return QP_ERR_BAD_ARG;
    if (degree_bits) *degree_bits = o->groups[g]->degree_log;
    if (n_polys) *n_polys = o->groups[g]->n_cols;
    return QP_OK;
}
extern "C" int qp_batch_fri_coeffs(const qp_batch_fri* o, size_t g, uint64_t* out, int space) {
    if (!o || g >= o->groups.size()) return QP_ERR_BAD_ARG;
    return qp_batch_coeffs(o->groups[g], out, space);
}
extern "C" int qp_batch_fri_cap(const qp_batch_fri* o, uint64_t* out, int space) {
    if (!o) return QP_ERR_BAD_ARG;
    const TreeBuf& t = o->groups.back()->tree;
    return copy_out(o->ctx, out, space, t.cap, t.n_cap() * 4);
}
extern "C" size_t qp_batch_fri_digests_len(const qp_batch_fri* o) {
    size_t n = 0;
    if (o)
        for (const qp_batch* b : o->groups) n += b->tree.n_digests();
    return n;
}
extern "C" int qp_batch_fri_digests(const qp_batch_fri* o, uint64_t* out, int space) {
    if (!o) return QP_ERR_BAD_ARG;
    size_t pos = 0;
    for (const qp_batch* b : o->groups) {
        int rc = copy_out(o->ctx, out + pos, space, b->tree.digests, b->tree.n_digests() * 4);
        if (rc) return rc;
        pos += b->tree.n_digests() * 4;
    }
    return QP_OK;
}
extern "C" int qp_batch_fri_open(const qp_batch_fri* o, size_t leaf_index, uint64_t* siblings_out) {
    if (!o) return QP_ERR_BAD_ARG;
    const unsigned lg0 = o->groups[0]->tree.shape.lg_leaves;
    if (leaf_index >> lg0) return fail(o->ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    size_t pos = 0;
    for (const qp_batch* b : o->groups) {  // batch_merkle_tree.rs:139-150
        int rc = tree_prove(o->ctx, b->tree, leaf_index >> (lg0 - b->tree.shape.lg_leaves), siblings_out + pos);
        if (rc) return rc;
        pos += 4 * (size_t)b->tree.shape.num_layers();
    }
    return QP_OK;
}
extern "C" int qp_batch_fri_values(const qp_batch_fri* o, size_t leaf_index, uint64_t* out) {
    if (!o) return QP_ERR_BAD_ARG;
    const unsigned lg0 = o->groups[0]->tree.shape.lg_leaves;
    if (leaf_index >> lg0) return fail(o->ctx, QP_ERR_BAD_ARG, "leaf index out of range");
    size_t pos = 0;
    for (const qp_batch* b : o->groups) {
        int rc = batch_gather(b, nullptr, leaf_index >> (lg0 - b->tree.shape.lg_leaves), 1, (unsigned)b->n_cols, out + pos,
                              QP_HOST);
        if (rc) return rc;
        pos += b->n_cols;
    }
    return QP_OK;
}

// ---------------------------------------------------------------------------------------------
// Plonk permutation argument and quotient polynomials (quotient.cuh)
// ---------------------------------------------------------------------------------------------
struct qp_circuit {
    qp_ctx* ctx = nullptr;
    qp_circuit_desc d{};          // scalar fields only; the pointers below are the device copies
    uint64_t* k_is = nullptr;     // [num_routed_wires]
    uint64_t* sigmas = nullptr;   // [num_routed_wires][n], may be null (quotient-only circuits)
    uint64_t* program = nullptr;  // program_len + 1 words (OP_END appended)
    uint32_t* seg_off = nullptr;  // first word of every OP_END-terminated segment the interpreter runs
    unsigned n_seg = 0;
    quotient::NativePoseidon native{};  // a PoseidonGate handed to poseidon_gate_kernel (OP_NATIVE_POSEIDON)
    uint64_t* pool = nullptr;
    uint64_t* zh = nullptr;       // [2][2^qdb]: Z_H on the coset, and its inverses
    unsigned max_emit = 0;        // largest constraint index in the program
    ScaleTables g_inv;            // powers of 1/g up to 2^(degree_bits + qdb)
    // lookup argument (common_data.luts, prover_data.lookup_rows); none: luts empty
    std::vector<std::vector<uint16_t>> luts;  // [t][2 len]: (input, output) pairs
    std::vector<uint32_t> lookup_tables;      // [n_luts][3]: last_lu_row, last_lut_row, first_lut_row
    uint64_t* d_lookup_tables = nullptr;      // the same on the device (uint32)
    uint64_t* d_lookup_rows = nullptr;        // uint32 [n_lookup_rows]: row | LookupTableGate row << 31
    unsigned n_lookup_rows = 0;
    uint64_t* lookup_consts = nullptr;        // [nc][4 + n_luts]: the challenges' deltas and table evaluations
    bool lookup_challenges_set = false;
    unsigned lookup_terms() const { return d.n_luts ? 4 + (unsigned)d.n_luts + 2 * (d.num_lookup_polys - 1) : 0; }
};

extern "C" int qp_circuit_create(qp_ctx* ctx, const qp_circuit_desc* d, qp_circuit** out) {
    if (!ctx || !d || !out) return QP_ERR_BAD_ARG;
    *out = nullptr;
    if (d->num_challenges == 0 || d->num_challenges > (unsigned)quotient::MAX_CHALLENGES)
        return fail(ctx, QP_ERR_BAD_ARG, "num_challenges must be 1..4");
    if (d->max_degree < 2) return fail(ctx, QP_ERR_BAD_ARG, "max_degree > 1 (util/partial_products.rs:17)");
    if (d->n_luts || d->num_lookup_polys || d->num_lookup_selectors) {
        // circuit_builder.rs:1183-1194,1284-1290: 4 + n_luts lookup selectors after the gate selectors; per challenge
        // RE + ceil(num_lu_slots / lookup_accumulator_degree) partial polynomials, lookup_accumulator_degree =
        // quotient_degree_factor - 1 (circuit_data.rs:557-559) = max_degree - 1 here
        const unsigned lu_slots = d->num_routed_wires / 2, lut_slots = d->num_routed_wires / 3;
        if (!d->n_luts || !d->luts || d->max_degree < 3 || lut_slots == 0 ||
            d->num_lookup_selectors != 4 + d->n_luts ||
            d->num_lookup_polys != 1 + (lu_slots + d->max_degree - 2) / (d->max_degree - 1) ||
            (size_t)d->num_selectors + d->num_lookup_selectors > d->num_constants)
            return fail(ctx, QP_ERR_BAD_ARG, "inconsistent lookup declaration (num_lookup_polys / num_lookup_selectors / luts)");
        const size_t n_rows = (size_t)1 << d->degree_bits;
        for (size_t t = 0; t < d->n_luts; t++) {
            const qp_lookup_table& lt = d->luts[t];
            if (!lt.table || !lt.len) return fail(ctx, QP_ERR_BAD_ARG, "Empty LUTs are not supported.");
            // the rows are upside down: LookupGate rows [last_lu_row, last_lut_row), LookupTableGate rows
            // [last_lut_row, first_lut_row] holding the table, and the row after them is read (prover.rs:553,580)
            if (!(lt.last_lu_row <= lt.last_lut_row && lt.last_lut_row <= lt.first_lut_row && (size_t)lt.first_lut_row + 1 < n_rows) ||
                lt.first_lut_row - lt.last_lut_row + 1 != (lt.len + lut_slots - 1) / lut_slots)
                return fail(ctx, QP_ERR_BAD_ARG, "lookup rows inconsistent with the table");
            for (size_t u = 0; u < d->n_luts; u++)   // the row after a table belongs to no table (the builder's NoopGate)
                if (d->luts[u].last_lu_row <= lt.first_lut_row + 1 && lt.first_lut_row + 1 <= d->luts[u].first_lut_row && u != t)
                    return fail(ctx, QP_ERR_BAD_ARG, "lookup tables' rows overlap");
        }
    }
    if (d->num_routed_wires == 0 || d->num_routed_wires > d->num_wires || !d->k_is)
        return fail(ctx, QP_ERR_BAD_ARG, "bad wire counts");
    // num_partial_products(n, max_degree) = ceil(n / max_degree) - 1, util/partial_products.rs:41-48
    if (d->num_partial_products + 1 != (d->num_routed_wires + d->max_degree - 1) / d->max_degree ||
        d->num_partial_products + 1 > (unsigned)quotient::MAX_CHUNKS)
        return fail(ctx, QP_ERR_BAD_ARG, "num_partial_products inconsistent with num_routed_wires / max_degree");
    if (d->degree_bits + d->quotient_degree_bits > ctx->tw_lg)
        return fail(ctx, QP_ERR_TOO_LARGE, "quotient domain larger than the context's twiddle table");
    if (d->program_len && !d->program) return fail(ctx, QP_ERR_BAD_ARG, "null program");
    if (d->program_regs > 256 || (size_t)d->program_regs * quotient::BLOCK * 8 + d->pool_len * 8 > 190 * 1024)
        return fail(ctx, QP_ERR_TOO_LARGE, "constraint program needs too many registers");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    qp_circuit* c = new qp_circuit();
    c->ctx = ctx;
    c->d = *d;
    c->d.k_is = c->d.sigmas = c->d.program = c->d.pool = nullptr;  // the caller's host pointers die with the call
    c->d.luts = nullptr;
    struct Guard {  // every early return below releases what has been allocated so far
        qp_circuit* c;
        ~Guard() {
            if (c) qp_circuit_free(c);
        }
    } guard{c};
    const size_t n = (size_t)1 << d->degree_bits;
    int rc = dev_alloc(ctx, &c->k_is, d->num_routed_wires);
    if (!rc) rc = dev_alloc(ctx, &c->program, d->program_len + 1);
    if (!rc) rc = dev_alloc(ctx, &c->pool, d->pool_len ? d->pool_len : 1);
    if (!rc) rc = dev_alloc(ctx, &c->zh, (size_t)2 << d->quotient_degree_bits);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(c->k_is, d->k_is, d->num_routed_wires * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint64_t> prog(d->program, d->program + d->program_len);
    prog.push_back(quotient::OP_END);
    std::vector<uint32_t> seg_off;
    for (size_t k = 0, start = 0; k < prog.size(); k++)
        if ((prog[k] & 0xff) == quotient::OP_END) {
            if (k == start + 1 && (prog[start] & 0xff) == quotient::OP_NATIVE_POSEIDON) {
                const uint64_t w = prog[start];
                if (c->native.present) return fail(ctx, QP_ERR_BAD_ARG, "more than one native PoseidonGate segment");
                c->native = {1u, (unsigned)((w >> 16) & 0xff), (unsigned)((w >> 24) & 0xff), (unsigned)((w >> 8) & 0xff),
                             (unsigned)((w >> 32) & 0xffff), (unsigned)((w >> 48) & 1)};
                if (c->native.sel_column >= d->num_constants || c->native.group_start > c->native.index ||
                    c->native.index >= c->native.group_end || d->num_wires < 135)
                    return fail(ctx, QP_ERR_BAD_ARG, "malformed native PoseidonGate segment");
            } else if (k > start) {
                for (size_t q = start; q < k; q++)
                    if ((prog[q] & 0xff) == quotient::OP_NATIVE_POSEIDON)
                        return fail(ctx, QP_ERR_BAD_ARG, "a native PoseidonGate word must be a segment of its own");
                seg_off.push_back((uint32_t)start);
            }
            start = k + 1;
        }
    c->n_seg = (unsigned)seg_off.size();
    if (!seg_off.empty()) {
        rc = dev_alloc(ctx, (uint64_t**)&c->seg_off, (seg_off.size() + 1) / 2);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(c->seg_off, seg_off.data(), seg_off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (uint64_t ins : prog) {
        const unsigned op = ins & 0xff, dst = (ins >> 8) & 0xff, a = (ins >> 16) & 0xff, b = (ins >> 24) & 0xff;
        const uint64_t cc = ins >> 32;
        const unsigned nr = d->program_regs;
        bool ok = op <= quotient::OP_NATIVE_POSEIDON;
        switch (op) {
            case quotient::OP_ADD: case quotient::OP_SUB: case quotient::OP_MUL: ok = dst < nr && a < nr && b < nr; break;
            case quotient::OP_FMAI: ok = dst < nr && a < nr && b < nr && cc < d->pool_len; break;
            case quotient::OP_MULI: case quotient::OP_ADDI: ok = dst < nr && a < nr && cc < d->pool_len; break;
            case quotient::OP_LDW: ok = dst < nr && cc < d->num_wires; break;
            case quotient::OP_LDK: ok = dst < nr && cc < (uint64_t)d->num_constants + d->num_routed_wires; break;
            case quotient::OP_LDP: ok = dst < nr && cc < 4; break;
            case quotient::OP_LDI: ok = dst < nr && cc < d->pool_len; break;
            case quotient::OP_EMIT: ok = a < nr && cc < 65536; break;
            case quotient::OP_GATE: ok = a < nr; break;
            default: break;
        }
        if (!ok) return fail(ctx, QP_ERR_BAD_ARG, "malformed constraint program");
        if (op == quotient::OP_EMIT && cc > c->max_emit) c->max_emit = (unsigned)cc;
    }
    if (c->native.present && c->max_emit < 122) c->max_emit = 122;  // PoseidonGate has 123 constraints
    CUDA_TRY(ctx, cudaMemcpyAsync(c->program, prog.data(), prog.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (d->pool_len)
        CUDA_TRY(ctx, cudaMemcpyAsync(c->pool, d->pool, d->pool_len * 8, cudaMemcpyHostToDevice, ctx->stream));
    // ZeroPolyOnCoset (field/src/zero_poly_coset.rs:24-66): g^n v^k - 1 and inverses, k < 2^qdb
    const size_t rate = (size_t)1 << d->quotient_degree_bits;
    std::vector<uint64_t> zh(2 * rate);
    const uint64_t g_pow_n = gl::host_pow(gl::GENERATOR, (uint64_t)1 << d->degree_bits);
    const uint64_t v = gl::host_primitive_root(d->quotient_degree_bits);
    for (size_t k = 0; k < rate; k++) {
        const uint64_t t = gl::host_mul(g_pow_n, gl::host_pow(v, k));
        zh[k] = t ? t - 1 : gl::P - 1;
        zh[rate + k] = gl::host_pow(zh[k], gl::P - 2);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(c->zh, zh.data(), zh.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint32_t> lookup_rows;
    if (d->n_luts) {
        for (size_t t = 0; t < d->n_luts; t++) {
            const qp_lookup_table& lt = d->luts[t];
            c->luts.emplace_back(lt.table, lt.table + 2 * lt.len);
            c->lookup_tables.insert(c->lookup_tables.end(), {lt.last_lu_row, lt.last_lut_row, lt.first_lut_row});
            for (uint32_t r = lt.last_lu_row; r <= lt.first_lut_row; r++)
                lookup_rows.push_back(r | (r >= lt.last_lut_row ? 1u << 31 : 0u));
        }
        c->n_lookup_rows = (unsigned)lookup_rows.size();
        rc = dev_alloc(ctx, &c->d_lookup_tables, (c->lookup_tables.size() + 1) / 2);
        if (!rc) rc = dev_alloc(ctx, &c->d_lookup_rows, (lookup_rows.size() + 1) / 2);
        if (!rc) rc = dev_alloc(ctx, &c->lookup_consts, (size_t)d->num_challenges * (4 + d->n_luts));
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(c->d_lookup_tables, c->lookup_tables.data(), c->lookup_tables.size() * 4,
                                      cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(c->d_lookup_rows, lookup_rows.data(), lookup_rows.size() * 4, cudaMemcpyHostToDevice,
                                      ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors die here
    if (d->sigmas) {
        const uint64_t* dev = nullptr;
        uint64_t* owned = nullptr;
        rc = to_device(ctx, d->sigmas, d->sigmas_space, (size_t)d->num_routed_wires * n, &dev, &owned);
        if (rc) return rc;
        if (!owned) {
            rc = dev_alloc(ctx, &owned, (size_t)d->num_routed_wires * n);
            if (rc) return rc;
            CUDA_TRY(ctx, cudaMemcpyAsync(owned, dev, (size_t)d->num_routed_wires * n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        c->sigmas = owned;
    }
    rc = build_scale(ctx, (int)(d->degree_bits + d->quotient_degree_bits), {gl::host_pow(gl::GENERATOR, gl::P - 2)}, &c->g_inv);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    c->d.k_is = c->d.sigmas = c->d.program = c->d.pool = nullptr;
    c->d.luts = nullptr;
    guard.c = nullptr;
    *out = c;
    return QP_OK;
}

extern "C" int qp_circuit_describe(const qp_circuit* c, qp_circuit_desc* out) {
    if (!c || !out) return QP_ERR_BAD_ARG;
    *out = c->d;  // scalar fields; the pointers were cleared at creation
    return QP_OK;
}
extern "C" int qp_circuit_has_sigmas(const qp_circuit* c) { return c && c->sigmas ? 1 : 0; }

// Device-to-device copy between two contexts' devices (peer copy over NVLink where the topology has it),
// ordered after the work queued on src_ctx's stream and complete when the call returns.
extern "C" int qp_memcpy_peer(qp_ctx* dst_ctx, uint64_t* dst, qp_ctx* src_ctx, const uint64_t* src, size_t n_words) {
    if (!dst_ctx || !src_ctx) return QP_ERR_BAD_ARG;
    if ((!dst || !src) && n_words) return fail(src_ctx, QP_ERR_BAD_ARG, "null buffer");
    if (n_words == 0) return QP_OK;
    CUDA_TRY(src_ctx, cudaSetDevice(src_ctx->device));
    CUDA_TRY(src_ctx, cudaMemcpyPeerAsync(dst, dst_ctx->device, src, src_ctx->device, n_words * 8, src_ctx->stream));
    CUDA_TRY(src_ctx, cudaStreamSynchronize(src_ctx->stream));
    return QP_OK;
}

// Stream-ordered device scratch for host-side drivers that keep intermediates on the device.
extern "C" int qp_dev_alloc(qp_ctx* ctx, size_t n_words, uint64_t** out) {
    if (!ctx || !out) return QP_ERR_BAD_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return dev_alloc(ctx, out, n_words);
}
extern "C" void qp_dev_free(qp_ctx* ctx, uint64_t* p) {
    if (ctx && p) dev_free(ctx, p);
}
extern "C" int qp_memcpy(qp_ctx* ctx, uint64_t* dst, int dst_space, const uint64_t* src, int src_space, size_t n_words) {
    if (!ctx || (n_words && (!dst || !src))) return QP_ERR_BAD_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, n_words * 8, cudaMemcpyDefault, ctx->stream));
    (void)dst_space;
    (void)src_space;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return QP_OK;
}

extern "C" void qp_circuit_free(qp_circuit* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    dev_free(c->ctx, c->k_is);
    dev_free(c->ctx, c->sigmas);
    dev_free(c->ctx, c->program);
    dev_free(c->ctx, (uint64_t*)c->seg_off);
    dev_free(c->ctx, c->pool);
    dev_free(c->ctx, c->zh);
    dev_free(c->ctx, c->d_lookup_tables);
    dev_free(c->ctx, c->d_lookup_rows);
    dev_free(c->ctx, c->lookup_consts);
    cudaStreamSynchronize(c->ctx->stream);
    dev_free(c->ctx, c->g_inv.lo);
    dev_free(c->ctx, c->g_inv.hi);
    delete c;
}

// all_wires_permutation_partial_products (prover.rs:402-480) + the Z-first reordering of
// prover.rs:255-261: out[(nc + nc * np)][n].
extern "C" int qp_circuit_partial_products_and_zs(qp_circuit* c, const uint64_t* wires, int space,
                                                  const uint64_t* betas, const uint64_t* gammas, uint64_t* out,
                                                  int out_space) {
    if (!c) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = c->ctx;
    if (!wires || !betas || !gammas || !out) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    if (!c->sigmas) return fail(ctx, QP_ERR_BAD_ARG, "circuit was created without sigmas");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const qp_circuit_desc& d = c->d;
    const size_t n = (size_t)1 << d.degree_bits;
    const unsigned nc = d.num_challenges, np = d.num_partial_products;
    const uint64_t* d_wires = nullptr;
    uint64_t* owned = nullptr;
    // only the routed wires are read
    int rc = to_device(ctx, wires, space, (size_t)d.num_routed_wires * n, &d_wires, &owned);
    if (rc) return rc;
    quotient::PermParams p{};
    p.degree_bits = d.degree_bits;
    p.nc = nc;
    p.nr = d.num_routed_wires;
    p.np = np;
    p.max_degree = d.max_degree;
    p.wires = d_wires;
    p.sigmas = c->sigmas;
    p.k_is = c->k_is;
    p.tw_row = ctx->tw + ((size_t)1 << d.degree_bits);
    for (unsigned i = 0; i < nc; i++) {
        p.betas[i] = betas[i];
        p.gammas[i] = gammas[i];
    }
    const size_t tiles_per_vec = n / quotient::SCAN_TILE ? n / quotient::SCAN_TILE : 1;
    const size_t n_tiles = tiles_per_vec * nc;
    uint64_t *tile = nullptr, *d_out = nullptr;
    rc = dev_alloc(ctx, &p.chunk, (size_t)nc * (np + 1) * n);
    if (!rc) rc = dev_alloc(ctx, &p.rowprod, (size_t)nc * n);
    if (!rc) rc = dev_alloc(ctx, &tile, n_tiles);
    const size_t out_words = (size_t)(nc + nc * np) * n;
    if (!rc) {
        if (out_space == QP_DEVICE) d_out = out;
        else rc = dev_alloc(ctx, &d_out, out_words);
    }
    if (rc) return rc;
    LAUNCH(ctx, quotient::perm_chunks_kernel, cdiv(n * nc, 128), 128, 0, p);
    LAUNCH(ctx, quotient::scan_tile_products_kernel, (unsigned)n_tiles, quotient::SCAN_BLOCK, 0, p.rowprod, n,
           tiles_per_vec, tile);
    LAUNCH(ctx, quotient::scan_tiles_kernel, nc, quotient::SCAN_BLOCK, 0, tile, tiles_per_vec);
    LAUNCH(ctx, quotient::perm_finish_kernel, (unsigned)n_tiles, quotient::SCAN_BLOCK, 0, p, tiles_per_vec, tile, d_out);
    if (out_space != QP_DEVICE) {
        rc = copy_out(ctx, out, QP_HOST, d_out, out_words);
        dev_free(ctx, d_out);
    }
    dev_free(ctx, p.chunk);
    dev_free(ctx, p.rowprod);
    dev_free(ctx, tile);
    dev_free(ctx, owned);
    if (!rc) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// The lookup challenges of the proof in progress and get_lut_poly(t).eval(delta) of every table
// (vanishing_poly.rs:29-52, prover.rs:687-716): the table's combos input + b output, padded with its first entry
// to whole LookupTableGate rows, are the coefficients in reverse order -- a Horner pass in table order.
static int lookup_set_challenges(qp_circuit* c, const uint64_t* deltas) {
    qp_ctx* ctx = c->ctx;
    const qp_circuit_desc& d = c->d;
    if (!d.n_luts) return fail(ctx, QP_ERR_BAD_ARG, "the circuit has no lookup tables");
    if (!deltas) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    const unsigned slots = d.num_routed_wires / 3, w = 4 + (unsigned)d.n_luts;
    std::vector<uint64_t> consts((size_t)d.num_challenges * w);
    for (unsigned ch = 0; ch < d.num_challenges; ch++) {
        uint64_t* k = &consts[(size_t)ch * w];
        for (int j = 0; j < 4; j++) k[j] = deltas[4 * ch + j] % gl::P;
        for (size_t t = 0; t < d.n_luts; t++) {
            const std::vector<uint16_t>& tab = c->luts[t];
            const size_t len = tab.size() / 2, padded = (slots - len % slots) % slots;
            uint64_t acc = 0;
            for (size_t e = 0; e < len + padded; e++) {
                const size_t q = e < len ? e : 0;
                const uint64_t combo = (uint64_t)(((unsigned __int128)gl::host_mul(k[1], tab[2 * q + 1]) + tab[2 * q]) % gl::P);
                acc = (uint64_t)(((unsigned __int128)gl::host_mul(acc, k[3]) + combo) % gl::P);
            }
            k[4 + t] = acc;
        }
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(c->lookup_consts, consts.data(), consts.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // consts dies here
    c->lookup_challenges_set = true;
    return QP_OK;
}

extern "C" int qp_circuit_set_lookup_challenges(qp_circuit* c, const uint64_t* deltas) {
    if (!c) return QP_ERR_BAD_ARG;
    return lookup_set_challenges(c, deltas);
}

// compute_all_lookup_polys (prover.rs:489-636): out[nc * num_lookup_polys][n] value columns.
extern "C" int qp_circuit_lookup_polys(qp_circuit* c, const uint64_t* wires, int space, const uint64_t* deltas,
                                       uint64_t* out, int out_space) {
    if (!c) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = c->ctx;
    if (!wires || !deltas || !out) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    int rc = lookup_set_challenges(c, deltas);
    if (rc) return rc;
    const qp_circuit_desc& d = c->d;
    const size_t n = (size_t)1 << d.degree_bits;
    const size_t out_words = (size_t)d.num_challenges * d.num_lookup_polys * n;
    TempScope tmp(ctx);
    const uint64_t* d_wires = nullptr;
    uint64_t *owned = nullptr, *d_out = out;
    rc = to_device(ctx, wires, space, (size_t)d.num_routed_wires * n, &d_wires, &owned);  // only routed wires are read
    if (rc) return rc;
    tmp.adopt(owned);
    if (out_space != QP_DEVICE) {
        rc = tmp.alloc(&d_out, out_words);
        if (rc) return rc;
    }
    quotient::LookupPolyParams p{};
    p.degree_bits = d.degree_bits;
    p.nc = d.num_challenges;
    p.np1 = d.num_lookup_polys;
    p.num_lu_slots = d.num_routed_wires / 2;
    p.num_lut_slots = d.num_routed_wires / 3;
    p.lu_degree = d.max_degree - 1;
    p.lut_degree = (p.num_lut_slots + p.np1 - 2) / (p.np1 - 1);
    p.wires = d_wires;
    p.rows = (const uint32_t*)c->d_lookup_rows;
    p.n_rows = c->n_lookup_rows;
    p.tables = (const uint32_t*)c->d_lookup_tables;
    p.n_luts = (unsigned)d.n_luts;
    p.consts = c->lookup_consts;
    p.out = d_out;
    CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, out_words * 8, ctx->stream));
    // one block per (challenge, row); one thread per slot and a spare one for the RE contribution
    const unsigned slots = std::max(std::max(p.num_lu_slots, p.num_lut_slots), p.np1);
    const unsigned block = std::max(64u, (slots + 1 + 31) / 32 * 32);
    if (block > 1024) return fail(ctx, QP_ERR_UNSUPPORTED, "more than 1023 lookup slots per row");
    LAUNCH(ctx, quotient::lookup_rows_kernel, p.n_rows * p.nc, block, block * 8, p);
    LAUNCH(ctx, quotient::lookup_chain_kernel, 1, 32, 0, p);
    if (out_space != QP_DEVICE) rc = copy_out(ctx, out, QP_HOST, d_out, out_words);
    if (!rc) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// compute_quotient_polys (prover.rs:640-866): out[num_challenges][n << qdb] coefficients.
// The quotient VALUES over the part of the quotient domain the three batches cover (all of it for whole
// batches; a coset shard's positions for the shards of a multi-GPU commitment -- every term of the vanishing
// polynomial at a point needs only that point and the next row of the same coset, prover.rs:679,750).
// whole domain: d_vals = [nc][n_lde] in natural order.  shard: d_vals = [nc][pos_count] in leaf order, local.
static int quotient_values(qp_circuit* c, const qp_batch* constants_sigmas, const qp_batch* wires,
                           const qp_batch* zs_partial_products, const uint64_t* betas, const uint64_t* gammas,
                           const uint64_t* alphas, const uint64_t public_inputs_hash[4], bool shard_layout,
                           uint64_t* d_vals, size_t pos_first, size_t pos_count) {
    qp_ctx* ctx = c->ctx;
    const qp_circuit_desc& d = c->d;
    const unsigned nc = d.num_challenges, np = d.num_partial_products;
    const unsigned lg_lde = d.degree_bits + d.quotient_degree_bits;
    quotient::Params p{};
    p.degree_bits = d.degree_bits;
    p.qdb = d.quotient_degree_bits;
    p.lg_lde = lg_lde;
    p.nc = nc;
    p.nr = d.num_routed_wires;
    p.np = np;
    p.max_degree = d.max_degree;
    p.num_constants = d.num_constants;
    p.cs = constants_sigmas->lde;
    p.cs_stride = constants_sigmas->n_local;
    p.wires = wires->lde;
    p.wires_stride = wires->n_local;
    p.zs = zs_partial_products->lde;
    p.zs_stride = zs_partial_products->n_local;
    p.tw_row = ctx->tw + ((size_t)1 << lg_lde);
    p.k_is = c->k_is;
    p.zh_eval = c->zh;
    p.zh_inv = c->zh + ((size_t)1 << d.quotient_degree_bits);
    p.program = c->program;
    p.seg_off = c->seg_off;
    p.n_seg = c->n_seg;
    p.pool = c->pool;
    p.pool_len = (unsigned)d.pool_len;
    p.n_regs = d.program_regs ? d.program_regs : 1;
    p.pos_first = pos_first;
    p.pos_count = pos_count;
    p.out_leaf_order = shard_layout ? 1 : 0;
    p.out_lg = shard_layout ? ilog2(pos_count) : lg_lde;
    const unsigned base = nc + nc * (np + 1);
    if (d.n_luts) {
        if (!c->lookup_challenges_set)
            return fail(ctx, QP_ERR_BAD_ARG, "lookup challenges not set (qp_circuit_set_lookup_challenges / qp_circuit_lookup_polys)");
        p.n_luts = (unsigned)d.n_luts;
        p.nlp = d.num_lookup_polys;
        p.num_selectors = d.num_selectors;
        p.num_lu_slots = d.num_routed_wires / 2;
        p.num_lut_slots = d.num_routed_wires / 3;
        p.lu_degree = d.max_degree - 1;
        p.lut_degree = (p.num_lut_slots + p.nlp - 2) / (p.nlp - 1);
        p.lookup_terms = c->lookup_terms();
        p.lookup_consts = c->lookup_consts;
    }
    p.gate_base = base + nc * c->lookup_terms();
    const unsigned stride = (p.gate_base > c->max_emit ? p.gate_base : c->max_emit) + 1;
    p.apow_stride = stride;
    std::vector<uint64_t> apow((size_t)nc * stride);
    for (unsigned a = 0; a < nc; a++) {
        p.betas[a] = betas[a];
        p.gammas[a] = gammas[a];
        p.alphas[a] = alphas[a];
        uint64_t pw = 1;
        for (unsigned t = 0; t < stride; t++) {
            apow[(size_t)a * stride + t] = pw;
            pw = gl::host_mul(pw, alphas[a] % gl::P);
        }
    }
    for (int i = 0; i < 4; i++) p.pih[i] = public_inputs_hash[i];
    TempScope tmp(ctx);
    uint64_t* d_apow = nullptr;
    int rc = tmp.alloc(&d_apow, apow.size());
    if (rc) return rc;
    const size_t out_words = (size_t)nc << p.out_lg;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_apow, apow.data(), apow.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    p.alpha_pows = d_apow;
    const size_t smem = quotient::smem_words(p.pool_len, p.n_regs) * 8;
    if (smem > 200 * 1024) return fail(ctx, QP_ERR_TOO_LARGE, "constraint program needs too much shared memory");
    if (smem > 48 * 1024)
        CUDA_TRY(ctx, cudaFuncSetAttribute(quotient::quotient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // The units of a point (permutation terms + program segments) are spread over blockIdx.y while
    // the tiles alone leave SMs idle: aim at ~8 resident blocks per SM.
    const unsigned tiles = (unsigned)cdiv(pos_count, quotient::BLOCK), units = 1 + c->n_seg;
    unsigned ny = (unsigned)cdiv((size_t)ctx->sm_count * 8, (size_t)tiles);
    if (ny > units) ny = units;
    if (ny < 1) ny = 1;
    // a native PoseidonGate is one more partial sum, from its own kernel; so are the lookup terms
    const unsigned n_parts = ny + (c->native.present ? 1 : 0) + (d.n_luts ? 1 : 0);
    uint64_t* d_partial = nullptr;
    if (n_parts > 1) {
        rc = tmp.alloc(&d_partial, (size_t)n_parts * out_words);
        if (rc) return rc;
    }
    p.out = n_parts > 1 ? d_partial : d_vals;
    p.partial_out = n_parts > 1;
    LAUNCH(ctx, quotient::quotient_kernel, dim3(tiles, ny), quotient::BLOCK, smem, p);
    if (c->native.present)
        LAUNCH(ctx, quotient::poseidon_gate_kernel, cdiv(pos_count, 128), 128, 0, p, c->native, d_partial + (size_t)ny * out_words);
    if (d.n_luts)
        LAUNCH(ctx, quotient::lookup_terms_kernel, cdiv(pos_count, 128), 128, 0, p, d_partial + (size_t)(n_parts - 1) * out_words);
    if (n_parts > 1)
        LAUNCH(ctx, quotient::combine_kernel, cdiv(out_words, 256), 256, 0, d_partial, n_parts, nc, p.out_lg,
               d.quotient_degree_bits, p.zh_inv, d_vals, lg_lde, pos_first, p.out_leaf_order);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // apow dies at return
    return QP_OK;
}

// values (natural order, [nc][n_lde], consumed as scratch) -> quotient coefficients:
// values.coset_ifft(F::coset_shift()), polynomial/mod.rs:58-88
static int quotient_finish(qp_circuit* c, uint64_t* d_vals, uint64_t* d_out) {
    qp_ctx* ctx = c->ctx;
    const qp_circuit_desc& d = c->d;
    const unsigned lg_lde = d.degree_bits + d.quotient_degree_bits;
    const size_t n_lde = (size_t)1 << lg_lde;
    NttJob job;
    job.src = d_vals;
    job.dst = d_out;
    job.L = (int)lg_lde;
    job.n_vec = d.num_challenges;
    job.inner_bits = 0;
    job.src_outer = n_lde;
    job.dst_outer = n_lde;
    job.out_mode = ntt::OUT_INVERSE;
    job.scratch = d_vals;
    int rc = run_ntt(ctx, job);
    if (rc) return rc;
    LAUNCH(ctx, quotient::scale_powers_kernel, cdiv((size_t)d.num_challenges * n_lde, 256), 256, 0, d_out,
           (size_t)d.num_challenges, lg_lde, c->g_inv.lo, c->g_inv.hi, c->g_inv.split);
    return QP_OK;
}

static int quotient_check_batches(qp_circuit* c, const qp_batch* const bs[3], bool whole) {
    qp_ctx* ctx = c->ctx;
    const qp_circuit_desc& d = c->d;
    const unsigned nc = d.num_challenges, np = d.num_partial_products;
    for (int k = 0; k < 3; k++) {
        const qp_batch* b = bs[k];
        if (b->ctx != ctx) return fail(ctx, QP_ERR_BAD_ARG, "batch belongs to another context");
        if (b->degree_log != d.degree_bits) return fail(ctx, QP_ERR_DEGREE_MISMATCH, "Polynomial degrees inconsistent");
        // "Having constraints of degree higher than the rate is not supported yet", prover.rs:662-666
        if (d.quotient_degree_bits > b->rate_bits) return fail(ctx, QP_ERR_BAD_ARG, "quotient degree exceeds the rate");
        if (whole && (b->block_first != 0 || b->block_count != (1u << b->rate_bits)))
            return fail(ctx, QP_ERR_BAD_ARG, "quotient evaluation needs unsharded batches (shards: qp_circuit_quotient_values_shard)");
        if (b->block_first != bs[0]->block_first || b->block_count != bs[0]->block_count || b->rate_bits != bs[0]->rate_bits)
            return fail(ctx, QP_ERR_BAD_ARG, "the three batches must be the same coset shard");
    }
    if (bs[0]->n_cols < (size_t)d.num_constants + d.num_routed_wires || bs[1]->n_cols < d.num_wires ||
        bs[2]->n_cols < (size_t)nc + (size_t)nc * np + (size_t)nc * d.num_lookup_polys)
        return fail(ctx, QP_ERR_BAD_ARG, "batch has too few polynomials for this circuit");
    return QP_OK;
}

extern "C" int qp_circuit_compute_quotient_polys(qp_circuit* c, const qp_batch* constants_sigmas,
                                                 const qp_batch* wires, const qp_batch* zs_partial_products,
                                                 const uint64_t* betas, const uint64_t* gammas,
                                                 const uint64_t* alphas, const uint64_t public_inputs_hash[4],
                                                 uint64_t* out, int out_space) {
    if (!c) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = c->ctx;
    if (!constants_sigmas || !wires || !zs_partial_products || !betas || !gammas || !alphas ||
        !public_inputs_hash || !out)
        return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    const qp_circuit_desc& d = c->d;
    const qp_batch* bs[3] = {constants_sigmas, wires, zs_partial_products};
    int rc = quotient_check_batches(c, bs, true);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n_lde = (size_t)1 << (d.degree_bits + d.quotient_degree_bits);
    const size_t out_words = (size_t)d.num_challenges * n_lde;
    TempScope tmp(ctx);
    uint64_t *d_out = out, *d_vals = nullptr;
    if (out_space != QP_DEVICE) rc = tmp.alloc(&d_out, out_words);
    if (!rc) rc = tmp.alloc(&d_vals, out_words);
    if (!rc) rc = quotient_values(c, constants_sigmas, wires, zs_partial_products, betas, gammas, alphas,
                                  public_inputs_hash, false, d_vals, 0, n_lde);
    if (!rc) rc = quotient_finish(c, d_vals, d_out);
    if (!rc && out_space != QP_DEVICE) rc = copy_out(ctx, out, QP_HOST, d_out, out_words);
    if (!rc) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// Multi-GPU form, step 1: the quotient values of ONE coset shard (the three batches are the shards of the
// same device).  out_dev: [nc][*pos_count] on the shard's device, leaf order, local; *pos_first = the shard's
// first position in the quotient domain (leaf order).  *pos_count = 0 (nothing written) if the shard lies
// outside the quotient domain (quotient_degree_bits < rate_bits and a late coset).
extern "C" int qp_circuit_quotient_values_shard(qp_circuit* c, const qp_batch* constants_sigmas, const qp_batch* wires,
                                                const qp_batch* zs_partial_products, const uint64_t* betas,
                                                const uint64_t* gammas, const uint64_t* alphas,
                                                const uint64_t public_inputs_hash[4], uint64_t* out_dev,
                                                size_t* pos_first, size_t* pos_count) {
    if (!c) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = c->ctx;
    if (!constants_sigmas || !wires || !zs_partial_products || !betas || !gammas || !alphas || !public_inputs_hash ||
        !out_dev || !pos_first || !pos_count)
        return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    const qp_circuit_desc& d = c->d;
    const qp_batch* bs[3] = {constants_sigmas, wires, zs_partial_products};
    int rc = quotient_check_batches(c, bs, false);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n_lde = (size_t)1 << (d.degree_bits + d.quotient_degree_bits);
    const size_t first = (size_t)wires->block_first << d.degree_bits;
    *pos_first = first;
    *pos_count = first >= n_lde ? 0 : std::min(wires->n_local, n_lde - first);
    if (*pos_count == 0) return QP_OK;
    return quotient_values(c, constants_sigmas, wires, zs_partial_products, betas, gammas, alphas, public_inputs_hash,
                           true, out_dev, *pos_first, *pos_count);
}

// Multi-GPU form, step 2 (on the device that gathered the shards' values): vals_leaf_order = [nc][n_lde] on
// the circuit's device, the quotient values of the whole domain in LEAF order (the shards' blocks side by
// side; consumed) -> quotient coefficients [nc][n_lde] (out: device or host).
extern "C" int qp_circuit_quotient_finish(qp_circuit* c, uint64_t* vals_leaf_order, uint64_t* out, int out_space) {
    if (!c) return QP_ERR_BAD_ARG;
    qp_ctx* ctx = c->ctx;
    if (!vals_leaf_order || !out) return fail(ctx, QP_ERR_BAD_ARG, "null argument");
    const qp_circuit_desc& d = c->d;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const unsigned lg_lde = d.degree_bits + d.quotient_degree_bits;
    const size_t out_words = (size_t)d.num_challenges << lg_lde;
    TempScope tmp(ctx);
    uint64_t *d_nat = nullptr, *d_out = out;
    int rc = tmp.alloc(&d_nat, out_words);
    if (!rc && out_space != QP_DEVICE) rc = tmp.alloc(&d_out, out_words);
    if (rc) return rc;
    LAUNCH(ctx, bitrev_permute_kernel, cdiv(out_words, 256), 256, 0, vals_leaf_order, d_nat, lg_lde, (size_t)d.num_challenges);
    rc = quotient_finish(c, d_nat, d_out);
    if (!rc && out_space != QP_DEVICE) rc = copy_out(ctx, out, QP_HOST, d_out, out_words);
    if (!rc) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

#include "multi_device.inl"
