// How far does a stream-ordered memory pool grow once a peer device has been granted access to it?
// (qp_mprove on two devices failed with "out of memory" at ~1 GiB of pool with 180 GB of device memory free.)
//   nvcc -O2 -o pool_peer_growth.bin pool_peer_growth.cu && ./pool_peer_growth.bin
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

static size_t grow(const char* what, cudaMemPool_t pool, cudaStream_t s, size_t chunk, size_t limit) {
    std::vector<void*> ptrs;
    size_t total = 0;
    cudaError_t e = cudaSuccess;
    while (total < limit) {
        void* p = nullptr;
        e = pool ? cudaMallocFromPoolAsync(&p, chunk, pool, s) : cudaMallocAsync(&p, chunk, s);
        if (e != cudaSuccess) break;
        ptrs.push_back(p);
        total += chunk;
    }
    cudaGetLastError();
    printf("%-70s grew to %6zu MiB in chunks of %4zu MiB: %s\n", what, total >> 20, chunk >> 20, cudaGetErrorString(e));
    for (void* p : ptrs) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    return total;
}

static cudaMemPool_t make_pool(int dev) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) printf("cudaMemPoolCreate: %s\n", cudaGetErrorString(e));
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    return pool;
}

static void grant(cudaMemPool_t pool, int to_dev) {
    cudaMemAccessDesc desc = {};
    desc.location.type = cudaMemLocationTypeDevice;
    desc.location.id = to_dev;
    desc.flags = cudaMemAccessFlagsProtReadWrite;
    cudaError_t e = cudaMemPoolSetAccess(pool, &desc, 1);
    if (e != cudaSuccess) printf("cudaMemPoolSetAccess: %s\n", cudaGetErrorString(e));
}

int main() {
    int n = 0;
    cudaGetDeviceCount(&n);
    printf("%d devices\n", n);
    if (n < 2) return 0;
    int can01 = 0, can10 = 0;
    cudaDeviceCanAccessPeer(&can01, 0, 1);
    cudaDeviceCanAccessPeer(&can10, 1, 0);
    printf("can access peer 0->1 %d, 1->0 %d\n", can01, can10);
    const size_t LIM = (size_t)12 << 30;
    cudaStream_t s0, s1;
    cudaSetDevice(0); cudaStreamCreate(&s0);
    cudaSetDevice(1); cudaStreamCreate(&s1);
    // 1. no peers anywhere
    cudaSetDevice(1);
    cudaMemPool_t a = make_pool(1);
    grow("explicit pool on device 1, no peer access", a, s1, (size_t)96 << 20, LIM);
    // 2. grant device 0 access WITHOUT cudaDeviceEnablePeerAccess
    grant(a, 0);
    grow("same pool after cudaMemPoolSetAccess(device 0) [cached memory]", a, s1, (size_t)96 << 20, LIM);
    cudaMemPoolTrimTo(a, 0);
    grow("same pool after a trim to 0 [fresh memory, peer-mapped]", a, s1, (size_t)96 << 20, LIM);
    cudaMemPoolTrimTo(a, 0);
    grow("same, 2 MiB chunks", a, s1, (size_t)2 << 20, (size_t)4 << 30);
    cudaMemPoolDestroy(a);
    // 3. with cudaDeviceEnablePeerAccess both ways
    cudaSetDevice(0); printf("enable 0->1: %s\n", cudaGetErrorString(cudaDeviceEnablePeerAccess(1, 0)));
    cudaSetDevice(1); printf("enable 1->0: %s\n", cudaGetErrorString(cudaDeviceEnablePeerAccess(0, 0)));
    cudaGetLastError();
    cudaMemPool_t b = make_pool(1);
    grant(b, 0);
    grow("new pool on device 1 with access for device 0, after EnablePeerAccess", b, s1, (size_t)96 << 20, LIM);
    cudaMemPoolTrimTo(b, 0);
    // 4. a big default-pool user on device 0 first (like a single-device prove), then peers
    cudaSetDevice(0);
    grow("default pool of device 0 (no peer access)", nullptr, s0, (size_t)256 << 20, LIM);
    cudaMemPool_t def0;
    cudaDeviceGetDefaultMemPool(&def0, 0);
    cudaMemPoolTrimTo(def0, 0);
    cudaMemPool_t c = make_pool(0);
    grant(c, 1);
    grow("new pool on device 0 with access for device 1", c, s0, (size_t)96 << 20, LIM);
    cudaSetDevice(1);
    grow("pool b on device 1 again", b, s1, (size_t)96 << 20, LIM);
    // 5. both pools hold memory at the same time
    {
        std::vector<void*> keep;
        cudaSetDevice(0);
        for (int i = 0; i < 40; i++) { void* p = nullptr; if (cudaMallocFromPoolAsync(&p, (size_t)96 << 20, c, s0) == cudaSuccess) keep.push_back(p); }
        cudaGetLastError();
        printf("device 0 holds %zu x 96 MiB\n", keep.size());
        cudaSetDevice(1);
        grow("pool b on device 1 while pool c on device 0 holds memory", b, s1, (size_t)96 << 20, LIM);
        cudaSetDevice(0);
        for (void* p : keep) cudaFreeAsync(p, s0);
        cudaStreamSynchronize(s0);
    }
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    printf("device 0 free %zu of %zu MiB\n", fr >> 20, tot >> 20);
    return 0;
}
