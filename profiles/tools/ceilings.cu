// ceilings.cu -- measurement tool (not product): what HBM bandwidth does the B200 give
// to streams with the SAME read/write mix as the HUSL kernels, with no math at all?
//   copy      : 50 % read / 50 % write (the MEASURED_PEAKS.json pattern)
//   fwd-like  : read 3 B/px, write 24 B/px (to_husl u8 -> f64) or 12 B/px (f32)
//   inv-like  : read 24 / 12 B/px, write 3 B/px
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/tools/ceilings.cu -o gpurun_out/ceilings
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_copy(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// one thread = 4 px: reads 12 B, writes W x 16 B  (W = 6: f64 out, W = 3: f32 out), warp-coalesced
template <int W>
__global__ void k_fwd_like(const uint32_t *__restrict__ in, uint4 *__restrict__ out, size_t tiles) {
    const int lane = threadIdx.x & 31;
    size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t t = warp; t < tiles; t += nw) {
        const uint32_t *p = in + t * 96 + lane;
        uint32_t w0 = __ldg(p), w1 = __ldg(p + 32), w2 = __ldg(p + 64);
        uint4 v = make_uint4(w0, w1, w2, w0 ^ w1);
        uint4 *o = out + t * (32 * W) + lane;
#pragma unroll
        for (int k = 0; k < W; ++k) { o[32 * k] = v; v.x += k; }
    }
}
template <int W>
__global__ void k_inv_like(const uint4 *__restrict__ in, uint32_t *__restrict__ out, size_t tiles) {
    const int lane = threadIdx.x & 31;
    size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t t = warp; t < tiles; t += nw) {
        const uint4 *p = in + t * (32 * W) + lane;
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < W; ++k) { uint4 v = __ldcs(p + 32 * k); acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
        uint32_t *o = out + t * 96 + lane;
        o[0] = acc.x ^ acc.w; o[32] = acc.y; o[64] = acc.z;
    }
}
__global__ void k_fill(uint4 *__restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_uint4(i, 1, 2, 3);
}
__global__ void k_read(const uint4 *__restrict__ a, uint32_t *sink, size_t n) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = __ldcs(a + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
float best_ms(F launch, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main() {
    const size_t px = 8ull * 3840 * 2160, tiles = px / 128;
    void *rgb, *hsl; uint32_t *sink;
    CK(cudaMalloc(&rgb, px * 3)); CK(cudaMalloc(&hsl, px * 24)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(rgb, 1, px * 3)); CK(cudaMemset(hsl, 1, px * 24));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, block = 256;
    auto rep = [&](const char *name, double bytes, float ms) { printf("%-28s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    size_t n16 = px * 24 / 32;   // copy half the big buffer into the other half
    rep("copy 0.8GB->0.8GB", 2.0 * n16 * 16, best_ms([&] { k_copy<<<grid, block>>>((const uint4 *)hsl, (uint4 *)hsl + n16, n16); }));
    rep("fill 1.6GB (write only)", px * 24.0, best_ms([&] { k_fill<<<grid, block>>>((uint4 *)hsl, px * 24 / 16); }));
    rep("read 1.6GB (read only)", px * 24.0, best_ms([&] { k_read<<<grid, block>>>((const uint4 *)hsl, sink, px * 24 / 16); }));
    rep("fwd-like u8->f64 (27 B/px)", px * 27.0, best_ms([&] { k_fwd_like<6><<<grid, block>>>((const uint32_t *)rgb, (uint4 *)hsl, tiles); }));
    rep("fwd-like u8->f32 (15 B/px)", px * 15.0, best_ms([&] { k_fwd_like<3><<<grid, block>>>((const uint32_t *)rgb, (uint4 *)hsl, tiles); }));
    rep("inv-like f64->u8 (27 B/px)", px * 27.0, best_ms([&] { k_inv_like<6><<<grid, block>>>((const uint4 *)hsl, (uint32_t *)rgb, tiles); }));
    rep("inv-like f32->u8 (15 B/px)", px * 15.0, best_ms([&] { k_inv_like<3><<<grid, block>>>((const uint4 *)hsl, (uint32_t *)rgb, tiles); }));
    CK(cudaDeviceSynchronize());
    return 0;
}
