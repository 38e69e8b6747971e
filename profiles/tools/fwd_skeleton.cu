// fwd_skeleton.cu -- measurement tool (not product): the memory operations of to_husl u8 -> f64 (k_fwd_tma<double>) without
// its arithmetic, element by element, to see which of them holds the kernel at 0.30 ms when a bare persistent bulk-store fill
// of the same 1.59 GB takes 0.256 ms (wr_ceilings.cu).  The full kernel with its arithmetic removed (HUSL_FWD_SKELETON builds,
// session 3) runs at 0.300 - 0.303 ms against 0.2988 ms: the arithmetic is free, the structure is the limit.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/tools/fwd_skeleton.cu -o profiles/tools/fwd_skeleton
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t ld_hint(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// FLAGS: 32 = the input reads carry an L2 evict-first policy and do not allocate in L1;
// FLAGS: 1 = read the 12 input bytes per lane (prefetched one tile ahead), 2 = rewrite the staging image every tile
// (6 x STS.128 per lane + fence.proxy.async), 4 = keep 64 KB of table in shared memory and do 12 dependent lookups per lane
template <int FLAGS, int OST, int PFD = 0>
__global__ void __launch_bounds__(512, 1)
k_skel(const uint32_t *__restrict__ in, unsigned char *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TB = 3072, TAB = (FLAGS & 4) ? 65536 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (FLAGS & 4) { for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<float2 *>(smem)[i] = make_float2(i, 1); __syncthreads(); }
    unsigned char *obuf = smem + TAB + warp * (OST * TB);
    for (int i = lane * 16; i < OST * TB; i += 512) *reinterpret_cast<uint4 *>(obuf + i) = make_uint4(i, 1, 2, 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long nw = (long long)gridDim.x * 16;
    long long tile = (long long)blockIdx.x * 16 + warp;
    uint64_t pol = 0;
    if (FLAGS & 32) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t w0 = 1, w1 = 2, w2 = 3;
    if ((FLAGS & 1) && tile < n_tiles) { const uint32_t *p = in + tile * 96 + lane * 3; w0 = __ldg(p); w1 = __ldg(p + 1); w2 = __ldg(p + 2); }
    int st = 0;
    uint32_t sink = 0;
    for (; tile < n_tiles; tile += nw) {
        sink += w0 ^ w1 ^ w2;                      // the loaded words are used even when the staging image is not rewritten
        uint32_t n0 = w0, n1 = w1, n2 = w2;
        if ((FLAGS & 1) && tile + nw < n_tiles) {
            const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
            if (FLAGS & 32) { n0 = ld_hint(p, pol); n1 = ld_hint(p + 1, pol); n2 = ld_hint(p + 2, pol); }
            else { n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); }
        }
        float acc = 0;
        if (FLAGS & 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc += reinterpret_cast<const float2 *>(smem)[((w0 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w1 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w2 >> (8 * k)) & 255) * 32 + lane].x;
            }
        }
        if ((FLAGS & 16) && lane == 0 && tile + PFD * nw < n_tiles)      // L2 prefetch of the input PFD tiles ahead (fire and forget)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + (tile + PFD * nw) * 96), "r"(384) : "memory");
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OST - 1) : "memory");
        __syncwarp();
        if (FLAGS & 2) {
            unsigned char *wr = obuf + st * TB + lane * 96;
#pragma unroll
            for (int k = 0; k < 6; ++k) *reinterpret_cast<double2 *>(wr + 16 * k) = make_double2((double)__uint_as_float(w0 + k) + acc, (double)__uint_as_float(w1 ^ w2));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + tile * TB), "r"(smem_addr(obuf + st * TB)), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        st = (st + 1 == OST) ? 0 : st + 1;
        w0 = n0; w1 = n1; w2 = n2;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (sink == 0x12345678u) out[0] = 1;
}

template <typename F> float best_ms(F launch, int reps = 15);
// ONE-SHOT grid: the kernel without arithmetic and without its table, every CTA (16 warps) converts K tiles per warp and exits;
// the hardware hands CTAs out as SMs become free (what a ticket schedule imitates)
template <int K>
__global__ void __launch_bounds__(512, 1)
k_skel_oneshot(const uint32_t *__restrict__ in, unsigned char *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TB = 3072, OST = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *obuf = smem + warp * (OST * TB);
    long long tile = (long long)blockIdx.x * (16 * K) + warp;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) { const uint32_t *p = in + tile * 96 + lane * 3; w0 = __ldg(p); w1 = __ldg(p + 1); w2 = __ldg(p + 2); }
    int st = 0;
#pragma unroll 1
    for (int k = 0; k < K && tile < n_tiles; ++k, tile += 16) {
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (k + 1 < K && tile + 16 < n_tiles) { const uint32_t *p = in + (tile + 16) * 96 + lane * 3; n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OST - 1) : "memory");
        __syncwarp();
        unsigned char *wr = obuf + st * TB + lane * 96;
#pragma unroll
        for (int j = 0; j < 6; ++j) *reinterpret_cast<double2 *>(wr + 16 * j) = make_double2((double)__uint_as_float(w0 + j), (double)__uint_as_float(w1 ^ w2));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + tile * TB), "r"(smem_addr(obuf + st * TB)), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        st ^= 1;
        w0 = n0; w1 = n1; w2 = n2;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// PERSISTENT CTAs with WARP-granular tickets: warp slot w of every CTA draws from counter[w]; ticket t -> tile t * 16 + w, so the
// tiles are handed out in address order, no CTA barrier.  Tickets are drawn two tiles ahead (one for the input prefetch).
template <bool TABLE, int RUN>
__global__ void __launch_bounds__(512, 1)
k_skel_tickets(const uint32_t *__restrict__ in, unsigned char *__restrict__ out, long long n_tiles, unsigned *ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TB = 3072, OST = 2, TAB = TABLE ? 65536 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (TABLE) { for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<float2 *>(smem)[i] = make_float2(i, 1); __syncthreads(); }
    unsigned char *obuf = smem + TAB + warp * (OST * TB);
    auto draw = [&]() -> long long {       // next run of RUN tiles of this warp slot: first tile index
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ctr + warp * 64, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        return ((long long)t * RUN) * 16 + warp;
    };
    long long cur = draw(), nxt = draw();
    unsigned pend = 0;                     // lane 0: the ticket drawn during this run, collected (shuffled) when the run ends
    if (lane == 0) pend = atomicAdd(ctr + warp * 64, 1u);
    int r = 0;
    long long tile = cur;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) { const uint32_t *p = in + tile * 96 + lane * 3; w0 = __ldg(p); w1 = __ldg(p + 1); w2 = __ldg(p + 2); }
    int st = 0;
    float acc = 0;
    while (tile < n_tiles) {
        // the tile after this one: the next of the run, or the first of the next run (whose ticket is already here)
        long long ntile;
        if (r + 1 < RUN) ntile = tile + 16; else ntile = nxt;
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (ntile < n_tiles) { const uint32_t *p = in + ntile * 96 + lane * 3; n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); }
        if (r + 1 == RUN) {                  // ntile took `nxt`; the ticket after it was requested a whole run ago
            nxt = ((long long)__shfl_sync(0xffffffffu, pend, 0) * RUN) * 16 + warp;
            if (lane == 0) pend = atomicAdd(ctr + warp * 64, 1u);
        }
        if (TABLE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc += reinterpret_cast<const float2 *>(smem)[((w0 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w1 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w2 >> (8 * k)) & 255) * 32 + lane].x;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OST - 1) : "memory");
        __syncwarp();
        unsigned char *wr = obuf + st * TB + lane * 96;
#pragma unroll
        for (int j = 0; j < 6; ++j) *reinterpret_cast<double2 *>(wr + 16 * j) = make_double2((double)__uint_as_float(w0 + j) + acc, (double)__uint_as_float(w1 ^ w2));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + tile * TB), "r"(smem_addr(obuf + st * TB)), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        st ^= 1;
        w0 = n0; w1 = n1; w2 = n2;
        tile = ntile;
        r = (r + 1 == RUN) ? 0 : r + 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <bool TABLE, int RUN>
float run_tickets(const uint32_t *in, unsigned char *out, long long tiles, int sms, unsigned *ctr) {
    const int sm = (TABLE ? 65536 : 0) + 16 * 2 * 3072;
    cudaFuncSetAttribute(k_skel_tickets<TABLE, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    return best_ms([&] { cudaMemsetAsync(ctr, 0, 4096); k_skel_tickets<TABLE, RUN><<<sms, 512, sm>>>(in, out, tiles, ctr); });
}
template <int K>
float run_oneshot(const uint32_t *in, unsigned char *out, long long tiles) {
    cudaFuncSetAttribute(k_skel_oneshot<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 3072);
    return best_ms([&] { k_skel_oneshot<K><<<(unsigned)((tiles + 16 * K - 1) / (16 * K)), 512, 16 * 2 * 3072>>>(in, out, tiles); });
}
// The same with the input arriving through a ring of TMA bulk loads (cp.async.bulk + mbarrier, IST tiles ahead): loads of the
// async proxy are not something the generic-proxy MEMBAR of fence.proxy.async waits for.
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, int bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int IST, int OST, bool TABLE>
__global__ void __launch_bounds__(512, 1)
k_skel_tma_in(const unsigned char *__restrict__ in, unsigned char *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TB = 3072, TAB = TABLE ? 65536 : 0, PER = OST * TB + IST * 384 + 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (TABLE) { for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<float2 *>(smem)[i] = make_float2(i, 1); __syncthreads(); }
    unsigned char *obuf = smem + TAB + warp * PER, *ibuf = obuf + OST * TB;
    const uint32_t ibuf_s = smem_addr(ibuf), bar_s = smem_addr(ibuf + IST * 384);
    const long long nw = (long long)gridDim.x * 16;
    const long long first = (long long)blockIdx.x * 16 + warp;
    if (lane == 0) {
        for (int s = 0; s < IST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < IST; ++s) {
            const long long t = first + s * nw;
            if (t < n_tiles) { mbar_expect_tx(bar_s + 8 * s, 384); bulk_load(ibuf_s + s * 384, in + t * 384, 384, bar_s + 8 * s); }
        }
    }
    __syncwarp();
    int st = 0, ist = 0; uint32_t phase = 0;
    for (long long tile = first; tile < n_tiles; tile += nw) {
        mbar_wait(bar_s + 8 * ist, phase);
        const uint32_t *rd = reinterpret_cast<const uint32_t *>(ibuf + ist * 384) + lane * 3;
        const uint32_t w0 = rd[0], w1 = rd[1], w2 = rd[2];
        __syncwarp();
        if (lane == 0) {
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) { mbar_expect_tx(bar_s + 8 * ist, 384); bulk_load(ibuf_s + ist * 384, in + t * 384, 384, bar_s + 8 * ist); }
        }
        float acc = 0;
        if (TABLE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc += reinterpret_cast<const float2 *>(smem)[((w0 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w1 >> (8 * k)) & 255) * 32 + lane].x;
                acc += reinterpret_cast<const float2 *>(smem)[((w2 >> (8 * k)) & 255) * 32 + lane].x;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(OST - 1) : "memory");
        __syncwarp();
        unsigned char *wr = obuf + st * TB + lane * 96;
#pragma unroll
        for (int k = 0; k < 6; ++k) *reinterpret_cast<double2 *>(wr + 16 * k) = make_double2((double)__uint_as_float(w0 + k) + acc, (double)__uint_as_float(w1 ^ w2));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + tile * TB), "r"(smem_addr(obuf + st * TB)), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        st = (st + 1 == OST) ? 0 : st + 1;
        if (++ist == IST) { ist = 0; phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int IST, int OST, bool TABLE>
float run_tma(const unsigned char *in, unsigned char *out, long long tiles, int sms) {
    const int sm = (TABLE ? 65536 : 0) + 16 * (OST * 3072 + IST * 384 + 64);
    cudaFuncSetAttribute(k_skel_tma_in<IST, OST, TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    return best_ms([&] { k_skel_tma_in<IST, OST, TABLE><<<sms, 512, sm>>>(in, out, tiles); });
}

template <typename F>
float best_ms(F launch, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); t.push_back(ms); }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

template <int FLAGS, int OST, int PFD = 0>
float run(const uint32_t *in, unsigned char *out, long long tiles, int sms) {
    const int sm = ((FLAGS & 4) ? 65536 : 0) + 16 * OST * 3072;
    cudaFuncSetAttribute(k_skel<FLAGS, OST, PFD>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    return best_ms([&] { k_skel<FLAGS, OST, PFD><<<sms, 512, sm>>>(in, out, tiles); });
}

int main() {
    const size_t px = 8ull * 3840 * 2160;
    const long long tiles = px / 128;
    void *rgb, *hsl;
    CK(cudaMalloc(&rgb, px * 3)); CK(cudaMalloc(&hsl, px * 24));
    CK(cudaMemset(rgb, 7, px * 3));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char *name, float ms) { printf("%-78s %8.4f ms  %7.1f GB/s written\n", name, ms, px * 24.0 / ms / 1e6); fflush(stdout); };
    const uint32_t *in = (const uint32_t *)rgb; unsigned char *out = (unsigned char *)hsl;
    rep("bulk stores of a constant staging image, 2 buffers per warp", run<0, 2>(in, out, tiles, sms));
    rep("+ input reads (12 B per lane, one tile ahead)", run<1, 2>(in, out, tiles, sms));
    rep("+ staging image rewritten every tile (6 STS.128 + fence.proxy.async)", run<2, 2>(in, out, tiles, sms));
    rep("+ both", run<3, 2>(in, out, tiles, sms));
    rep("+ both + 64 KB table and 12 lookups per lane (= the kernel without arithmetic)", run<7, 2>(in, out, tiles, sms));
    rep("the same with 3 staging buffers per warp", run<7, 3>(in, out, tiles, sms));
    rep("staging rewritten, 1 buffer per warp", run<2, 1>(in, out, tiles, sms));
    rep("staging rewritten, 3 buffers per warp", run<2, 3>(in, out, tiles, sms));
    rep("staging rewritten, 4 buffers per warp", run<2, 4>(in, out, tiles, sms));
    unsigned *ctr; CK(cudaMalloc(&ctr, 4096));
    rep("PERSISTENT, warp-granular tickets (16 counters), runs of 1 tile, no table", run_tickets<false, 1>(in, out, tiles, sms, ctr));
    rep("... runs of 1 tile, with the table and its lookups", run_tickets<true, 1>(in, out, tiles, sms, ctr));
    rep("... runs of 2 tiles, with the table", run_tickets<true, 2>(in, out, tiles, sms, ctr));
    rep("... runs of 4 tiles, with the table", run_tickets<true, 4>(in, out, tiles, sms, ctr));
    rep("ONE-SHOT grid: reads + staging + bulk stores, 1 tile per warp per CTA", run_oneshot<1>(in, out, tiles));
    rep("... 2 tiles per warp per CTA", run_oneshot<2>(in, out, tiles));
    rep("ONE-SHOT grid: reads + staging + bulk stores, 4 tiles per warp per CTA", run_oneshot<4>(in, out, tiles));
    rep("... 16 tiles per warp per CTA", run_oneshot<16>(in, out, tiles));
    rep("... 64 tiles per warp per CTA", run_oneshot<64>(in, out, tiles));
    rep("kernel without arithmetic, input reads with L2 evict-first + L1 no-allocate", run<39, 2>(in, out, tiles, sms));
    rep("kernel without arithmetic + L2 prefetch of the input 4 tiles ahead", run<23, 2, 4>(in, out, tiles, sms));
    rep("... 16 tiles ahead", run<23, 2, 16>(in, out, tiles, sms));
    rep("... 64 tiles ahead", run<23, 2, 64>(in, out, tiles, sms));
    rep("input through TMA bulk loads (2 tiles ahead) + staging rewritten + bulk stores", run_tma<2, 2, false>((const unsigned char *)rgb, out, tiles, sms));
    rep("... 4 tiles ahead", run_tma<4, 2, false>((const unsigned char *)rgb, out, tiles, sms));
    rep("... 4 tiles ahead + table lookups (= the kernel without arithmetic, TMA input)", run_tma<4, 2, true>((const unsigned char *)rgb, out, tiles, sms));
    rep("... 8 tiles ahead + table lookups", run_tma<8, 2, true>((const unsigned char *)rgb, out, tiles, sms));
    CK(cudaDeviceSynchronize());
    return 0;
}
