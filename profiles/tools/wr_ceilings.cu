// wr_ceilings.cu -- measurement tool (not product): which STORE pattern reaches the write bandwidth torch's fill_
// reaches on this B200 (7.48 TB/s, profiles/tools/fill_probe.py), for the output of to_husl u8 -> f64 (1.59 GB)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/tools/wr_ceilings.cu -o profiles/tools/wr_ceilings
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// persistent grid-stride, 16 bytes per thread per iteration
__global__ void k_fill16(uint4 *__restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_uint4(i, 1, 2, 3);
}
// persistent grid-stride, 32 bytes per thread per iteration (STG.256)
__global__ void k_fill32(double *__restrict__ b, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(b + 4 * i), "d"(1.0), "d"(2.0), "d"(3.0), "d"((double)i) : "memory");
}
// one-shot grid (not persistent): a CTA of 256 threads writes U x 4 KB contiguous
template <int U>
__global__ void k_fill_oneshot(uint4 *__restrict__ b, size_t n) {
    size_t base = (size_t)blockIdx.x * (256 * U) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t i = base + 256 * u; if (i < n) b[i] = make_uint4(i, 1, 2, 3); }
}
// persistent, but every CTA owns one CONTIGUOUS range of the buffer
__global__ void k_fill_blocked(uint4 *__restrict__ b, size_t n) {
    const size_t per = (n + gridDim.x - 1) / gridDim.x, lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) b[i] = make_uint4(i, 1, 2, 3);
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk stores: every warp fills a TILE-byte staging buffer once and sends it tile after tile (grid-stride over tiles)
template <int TILE, int DEPTH>
__global__ void k_fill_bulk(unsigned char *__restrict__ out, size_t tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + warp * TILE;
    for (int i = lane * 16; i < TILE; i += 512) *reinterpret_cast<uint4 *>(buf + i) = make_uint4(i, 1, 2, 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t nw = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + warp; t < tiles; t += nw) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * TILE), "r"(smem_addr(buf)), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH) : "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// the same with ONE bulk copy per CTA per iteration (CTA-sized contiguous chunk)
template <int TILE>
__global__ void k_fill_bulk_cta(unsigned char *__restrict__ out, size_t chunks, int chunk_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x * 16; i < chunk_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(i, 1, 2, 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t t = blockIdx.x; t < chunks; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * (size_t)chunk_bytes), "r"(smem_addr(smem)), "r"(chunk_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// persistent CTAs that take contiguous chunks from a global counter (dynamic schedule) instead of a fixed share
__global__ void k_fill_dynamic(uint4 *__restrict__ b, size_t n, int chunk16, unsigned *counter) {
    __shared__ unsigned s_chunk;
    const size_t chunks = (n + chunk16 - 1) / chunk16;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(counter, 1u);
        __syncthreads();
        const size_t c = s_chunk;
        if (c >= chunks) break;
        const size_t lo = c * chunk16, hi = lo + chunk16 < n ? lo + chunk16 : n;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) b[i] = make_uint4(i, 1, 2, 3);
    }
}
// per-WARP dynamic schedule: each warp takes the next 3 KB tile (or a run of R tiles) from the counter
template <int R>
__global__ void k_fill_dynamic_warp(uint4 *__restrict__ b, size_t tiles, unsigned *counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned t0 = 0;
        if (lane == 0) t0 = atomicAdd(counter, (unsigned)R);
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        if (t0 >= tiles) break;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t t = (size_t)t0 + r;
            if (t < tiles) {
                uint4 *o = b + t * 192 + lane;
#pragma unroll
                for (int k = 0; k < 6; ++k) o[32 * k] = make_uint4(k, 1, 2, 3);
            }
        }
    }
}

// fixed shares again, every CTA recording when it started and finished (globaltimer, ns) and on which SM it ran
__global__ void k_fill16_timed(uint4 *__restrict__ b, size_t n, unsigned long long *t) {
    unsigned long long t0, t1; unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_uint4(i, 1, 2, 3);
    __syncthreads();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    if (threadIdx.x == 0) { t[3 * blockIdx.x] = t0; t[3 * blockIdx.x + 1] = t1; t[3 * blockIdx.x + 2] = sm; }
}

template <typename F>
float best_ms(F launch, int reps = 15) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    const size_t px = 8ull * 3840 * 2160, bytes = px * 24, n16 = bytes / 16;
    void *hsl;
    CK(cudaMalloc(&hsl, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char *name, int a, int b, float ms) { printf("%-44s %5d %5d %8.4f ms  %8.1f GB/s\n", name, a, b, ms, bytes / ms / 1e6); fflush(stdout); };
    rep("cudaMemsetAsync", 0, 0, best_ms([&] { cudaMemsetAsync(hsl, 1, bytes); }));
    for (int block : {128, 256, 512, 1024})
        for (int per_sm : {1, 2, 4, 8, 16}) {
            if (block * per_sm > 2048) continue;
            rep("persistent STG.128  (block, CTAs/SM)", block, per_sm, best_ms([&] { k_fill16<<<sms * per_sm, block>>>((uint4 *)hsl, n16); }));
        }
    for (int block : {256, 512, 1024})
        for (int per_sm : {1, 2, 4, 8}) {
            if (block * per_sm > 2048) continue;
            rep("persistent STG.256  (block, CTAs/SM)", block, per_sm, best_ms([&] { k_fill32<<<sms * per_sm, block>>>((double *)hsl, bytes / 32); }));
        }
    rep("one-shot grid, 4 KB per CTA", 256, 1, best_ms([&] { k_fill_oneshot<1><<<(unsigned)((n16 + 255) / 256), 256>>>((uint4 *)hsl, n16); }));
    rep("one-shot grid, 16 KB per CTA", 256, 4, best_ms([&] { k_fill_oneshot<4><<<(unsigned)((n16 + 1023) / 1024), 256>>>((uint4 *)hsl, n16); }));
    rep("one-shot grid, 64 KB per CTA", 256, 16, best_ms([&] { k_fill_oneshot<16><<<(unsigned)((n16 + 4095) / 4096), 256>>>((uint4 *)hsl, n16); }));
    for (int per_sm : {1, 2, 4, 8})
        rep("persistent, contiguous range per CTA", 256, per_sm, best_ms([&] { k_fill_blocked<<<sms * per_sm, 256>>>((uint4 *)hsl, n16); }));
    for (int warps : {4, 8, 16, 32}) {
        cudaFuncSetAttribute(k_fill_bulk<3072, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_fill_bulk<3072, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_fill_bulk<6144, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_fill_bulk<12288, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        rep("bulk 3 KB per warp, 1 in flight (warps, 1)", warps, 1, best_ms([&] { k_fill_bulk<3072, 1><<<sms, warps * 32, warps * 3072>>>((unsigned char *)hsl, bytes / 3072); }));
        rep("bulk 3 KB per warp, 4 in flight (warps, 1)", warps, 1, best_ms([&] { k_fill_bulk<3072, 4><<<sms, warps * 32, warps * 3072>>>((unsigned char *)hsl, bytes / 3072); }));
        rep("bulk 6 KB per warp, 2 in flight (warps, 1)", warps, 1, best_ms([&] { k_fill_bulk<6144, 2><<<sms, warps * 32, warps * 6144>>>((unsigned char *)hsl, bytes / 6144); }));
        if (warps <= 16) rep("bulk 12 KB per warp, 2 in flight (warps, 1)", warps, 1, best_ms([&] { k_fill_bulk<12288, 2><<<sms, warps * 32, warps * 12288>>>((unsigned char *)hsl, bytes / 12288); }));
    }
    {
        unsigned long long *t, ht[3 * 148];
        CK(cudaMalloc(&t, sizeof ht));
        for (int rep_i = 0; rep_i < 2; ++rep_i) {
            k_fill16_timed<<<sms, 1024>>>((uint4 *)hsl, n16, t);
            CK(cudaMemcpy(ht, t, sizeof ht, cudaMemcpyDeviceToHost));
        }
        unsigned long long first = ~0ull;
        for (int i = 0; i < sms; ++i) first = ht[3 * i] < first ? ht[3 * i] : first;
        printf("fixed shares, 1024 x 1: per-CTA finish time (us after the first start), by SM id:\n");
        std::vector<std::pair<unsigned, double>> v;
        for (int i = 0; i < sms; ++i) v.push_back({(unsigned)ht[3 * i + 2], (ht[3 * i + 1] - first) / 1000.0});
        std::sort(v.begin(), v.end());
        for (int i = 0; i < sms; ++i) printf("%3u:%6.1f%s", v[i].first, v[i].second, (i % 12 == 11) ? "\n" : " ");
        printf("\n");
    }
    unsigned *counter;
    CK(cudaMalloc(&counter, 4));
    for (int block : {256, 512, 1024})
        for (int per_sm : {1, 2, 4})
            for (int kb : {4, 16, 48, 192}) {
                if (block * per_sm > 2048) continue;
                char name[96]; snprintf(name, sizeof name, "dynamic chunks of %d KB (block, CTAs/SM)", kb);
                rep(name, block, per_sm, best_ms([&] { cudaMemsetAsync(counter, 0, 4); k_fill_dynamic<<<sms * per_sm, block>>>((uint4 *)hsl, n16, kb * 64, counter); }));
            }
    for (int block : {512, 1024}) {
        rep("dynamic per warp, 1 tile (block, 1)", block, 1, best_ms([&] { cudaMemsetAsync(counter, 0, 4); k_fill_dynamic_warp<1><<<sms, block>>>((uint4 *)hsl, bytes / 3072, counter); }));
        rep("dynamic per warp, 4 tiles (block, 1)", block, 1, best_ms([&] { cudaMemsetAsync(counter, 0, 4); k_fill_dynamic_warp<4><<<sms, block>>>((uint4 *)hsl, bytes / 3072, counter); }));
        rep("dynamic per warp, 16 tiles (block, 1)", block, 1, best_ms([&] { cudaMemsetAsync(counter, 0, 4); k_fill_dynamic_warp<16><<<sms, block>>>((uint4 *)hsl, bytes / 3072, counter); }));
    }
    cudaFuncSetAttribute(k_fill_bulk_cta<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int kb : {12, 24, 48, 96})
        for (int per_sm : {1, 2}) {
            if (kb * per_sm > 200) continue;
            rep("bulk per CTA (KB per copy, CTAs/SM)", kb, per_sm, best_ms([&] { k_fill_bulk_cta<0><<<sms * per_sm, 128, kb * 1024>>>((unsigned char *)hsl, bytes / (kb * 1024), kb * 1024); }));
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
