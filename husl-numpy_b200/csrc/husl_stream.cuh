// husl_stream.cuh -- asynchronously staged streaming kernels of the HUSL hot path (sm_100a).
//
//   k_fwd_tma<TOut>   (a) husl_from_rgb : uint8 RGB -> HUSL, results leave through TMA bulk stores
//                         (dispatched for float64 output, the HBM-bound case)
//   k_inv_tma<TIn>    (b) rgb_from_husl : HUSL arrives through a ring of TMA bulk loads + mbarriers
//                         (dispatched for float64 input)
//                         (float64 input with a 16- but not 32-byte aligned base, and host -> host calls that
//                         write their uint8 result in place over the link)
//   k_inv_direct      (b) float64 input, 32-byte aligned: NO ring -- a thread reads its own 4 pixels with three
//                         256-bit loads; the kernel that ships for device arrays since session 3 of round 2
//   k_inv_cp          (b) rgb_from_husl, float32 input with an edit on load: the ring filled by per-lane cp.async
//   k_inv_direct32    (b) float32 input without an edit: no ring either (three 128-bit loads per thread)
//
// Same per-pixel math as husl_kernels.cuh (husl_math.cuh); what changes is how
// the wide side of each kernel (the 12 / 24 B-per-pixel HUSL stream) moves:
//
//   * a warp owns a tile of 128 consecutive pixels; its HUSL tile is ONE
//     contiguous 1536 B (float) / 3072 B (double) block of global memory;
//   * the inverse kernel fetches tiles with the bulk async-copy engine
//     (`cp.async.bulk.shared::cluster.global`, SASS UBLKCP) into a per-warp
//     ring of STAGES shared-memory buffers, completion signalled on an
//     mbarrier per stage: loads run STAGES tiles ahead of the math without
//     holding registers, and no LDG/STS instruction is issued for them;
//   * the forward kernel writes its results straight into a shared-memory image
//     of the output tile (lane l owns bytes [l*48, l*48+48) resp. [l*96, ..+96),
//     i.e. its 4 pixels) and one elected lane hands the tile to the bulk engine
//     (`cp.async.bulk.global.shared::cta`): no LDS/STG instruction is issued for
//     the transposition to coalesced stores;
//   * the uint8 side (3 B/px) of the inverse goes out the same way (384 B per
//     tile); the uint8 side of the forward stays a register-prefetched LDG of 3
//     words per lane (12 B = 4 whole pixels);
//   * the bulk engine's bookkeeping (mbarrier arm/wait, elect, proxy fence) costs
//     about 60 warp-instructions per tile: worth it where HBM is the limit
//     (float64), not where the math pipes are (float32) -- measured in
//     profiles/r01_ab_variants.txt.
//
// Requirements (checked by the host before launching; otherwise the generic
// kernels run): HUSL pointer 16-byte aligned, uint8 pointer 16-byte aligned
// (inverse) / 4-byte aligned (forward).
#pragma once
#include "husl_kernels.cuh"

namespace husl {

// ---- PTX wrappers: mbarrier + bulk async copy ------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// L2 eviction-priority policies for the bulk copies (A/B knobs, see nphusl_cuda_api.cu)
#ifndef HUSL_INVCP_UNROLL
#define HUSL_INVCP_UNROLL 1    // tile loop of k_inv_cp unrolled by this factor (A/B knob)
#endif
#ifndef HUSL_FWD_ST_HINT
#define HUSL_FWD_ST_HINT 0      // 0 none | 1 evict_first | 2 evict_last | 3 no_allocate-like (evict_first, fraction 1)
#endif
#ifndef HUSL_INV_LD_HINT
#define HUSL_INV_LD_HINT 0
#endif
#ifndef HUSL_INV_ST_HINT
#define HUSL_INV_ST_HINT 0
#endif
template <int KIND>
__device__ __forceinline__ uint64_t l2_policy() {
    uint64_t p = 0;
    if (KIND == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    if (KIND == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    if (KIND == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void *dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(pol)
                 : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async (bulk copy) proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T> struct Tile {
    static constexpr int BYTES = TILE_PX * 3 * (int)sizeof(T);   // 1536 / 3072
    static constexpr int LANE = BYTES / 32;                       // 48 / 96
};
constexpr int U8_TILE = TILE_PX * 3;                               // 384

// ---------------------------------------------------------------------------------
// (a) husl_from_rgb, uint8 interleaved RGB -> interleaved HUSL through TMA stores
// ---------------------------------------------------------------------------------
// shared memory: [ 64 KB table | warps x OST x Tile<TOut>::BYTES ]
template <typename TOut, int BLOCK, int MINB, int OST>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_tma(const uint32_t *__restrict__ in, TOut *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const unsigned char *tab = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    constexpr int TB = Tile<TOut>::BYTES;
    unsigned char *obuf = smem + TAB_BYTES + warp * (OST * TB);
    const uint32_t obuf_s = smem_addr(obuf);
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;

    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) {
        const uint32_t *p = in + tile * 96 + lane * 3;
        w0 = ldg_u32(p); w1 = ldg_u32(p + 1); w2 = ldg_u32(p + 2);
    }
    int st = 0;
    for (; tile < n_tiles; tile += nw) {
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (tile + nw < n_tiles) {
            const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
            n0 = ldg_u32(p); n1 = ldg_u32(p + 1); n2 = ldg_u32(p + 2);
        }
        LinPx px[4];
        {
            float2 v;
            v = lut_lin<0>(tab, w0, lane8); px[0].hr = v.x; px[0].lr = v.y;
            v = lut_lin<1>(tab, w0, lane8); px[0].hg = v.x; px[0].lg = v.y;
            v = lut_lin<2>(tab, w0, lane8); px[0].hb = v.x; px[0].lb = v.y;
            v = lut_lin<3>(tab, w0, lane8); px[1].hr = v.x; px[1].lr = v.y;
            v = lut_lin<0>(tab, w1, lane8); px[1].hg = v.x; px[1].lg = v.y;
            v = lut_lin<1>(tab, w1, lane8); px[1].hb = v.x; px[1].lb = v.y;
            v = lut_lin<2>(tab, w1, lane8); px[2].hr = v.x; px[2].lr = v.y;
            v = lut_lin<3>(tab, w1, lane8); px[2].hg = v.x; px[2].lg = v.y;
            v = lut_lin<0>(tab, w2, lane8); px[2].hb = v.x; px[2].lb = v.y;
            v = lut_lin<1>(tab, w2, lane8); px[3].hr = v.x; px[3].lr = v.y;
            v = lut_lin<2>(tab, w2, lane8); px[3].hg = v.x; px[3].lg = v.y;
            v = lut_lin<3>(tab, w2, lane8); px[3].hb = v.x; px[3].lb = v.y;
        }
        float r[12];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            husl_from_diff<true>(make_diff(px[j]), r[3 * j], r[3 * j + 1], r[3 * j + 2]);

        // the bulk store that last used this buffer must have finished reading it
        if (lane == 0) bulk_wait_read<OST - 1>();
        __syncwarp();
        unsigned char *wr = obuf + st * TB + lane * Tile<TOut>::LANE;
        if (sizeof(TOut) == 4) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                *reinterpret_cast<float4 *>(wr + 16 * k) =
                    make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        } else {
            // Lanes are 96 bytes (24 banks) apart, so lanes l and l + 4 of a quarter-warp would hit the same four
            // banks with their k-th 16-byte chunk (2-way conflict on every STS.128: 17.4 M conflicts per launch in
            // the round-1 ncu capture).  Lanes whose bit 2 is set write their chunks one position ahead
            // (k + 1 mod 6): the eight lanes of a quarter-warp then cover all 32 banks.
            const bool rot = (lane & 4) != 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                constexpr int K6 = 6;
                const int k1 = (k + 1) % K6;
                const float a = rot ? r[2 * k1] : r[2 * k], b = rot ? r[2 * k1 + 1] : r[2 * k + 1];
                *reinterpret_cast<double2 *>(wr + (rot ? 16 * k1 : 16 * k)) = make_double2((double)a, (double)b);
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (HUSL_FWD_ST_HINT) bulk_store_hint(reinterpret_cast<unsigned char *>(out) + tile * TB, obuf_s + st * TB, TB, l2_policy<HUSL_FWD_ST_HINT>());
            else bulk_store(reinterpret_cast<unsigned char *>(out) + tile * TB, obuf_s + st * TB, TB);
            bulk_commit();
        }
        st = (st + 1 == OST) ? 0 : st + 1;
        w0 = n0; w1 = n1; w2 = n2;
    }
    if (lane == 0) bulk_wait_all<0>();     // shared memory must outlive the copies
}

// ---------------------------------------------------------------------------------
// (b) rgb_from_husl: interleaved HUSL through TMA loads -> uint8 RGB through TMA stores
// ---------------------------------------------------------------------------------
// shared memory per warp: [ IST x Tile<TIn>::BYTES | 2 x 384 B uint8 tiles | IST mbarriers ]
template <typename TIn, int IST>
struct InvSmem {
    static constexpr int PER_WARP = IST * Tile<TIn>::BYTES + 2 * U8_TILE + 64;
};

// EDIT: the caller's HUSL-space select + edit (husl_edit.cuh) is applied to the loaded values before
// the inverse -- to_rgb(hsl, edit=...) -- so a to_husl -> edit -> to_rgb pipeline needs no pass of its own
// over the 24 B/px HUSL array for the edit.
template <typename TIn, int BLOCK, int MINB, int IST, int ED = -1>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_tma(const TIn *__restrict__ in, uint32_t *__restrict__ out, long long n_tiles, const EditArg<ED> ea = EditArg<ED>()) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<TIn>::BYTES;
    unsigned char *base = smem + warp * InvSmem<TIn, IST>::PER_WARP;
    unsigned char *ibuf = base;
    unsigned char *obuf = base + IST * TB;
    const uint32_t ibuf_s = smem_addr(ibuf), obuf_s = smem_addr(obuf);
    const uint32_t bar_s = smem_addr(base + IST * TB + 2 * U8_TILE);
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    const long long first = (long long)blockIdx.x * (BLOCK / 32) + warp;
    // Tiles are walked from the END of the image: in a to_husl -> (edit) -> to_rgb
    // pipeline the producer wrote the HUSL array front to back, so its last
    // ~100 MB are still in the 126 MB L2 when this kernel starts.
    const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + (n_tiles - 1) * (long long)Tile<TIn>::BYTES;
    unsigned char *dstb = reinterpret_cast<unsigned char *>(out) + (n_tiles - 1) * (long long)U8_TILE;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < IST; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
#pragma unroll
        for (int s = 0; s < IST; ++s) {
            const long long t = first + s * nw;
            if (t < n_tiles) {
                mbar_expect_tx(bar_s + 8 * s, TB);
                if (HUSL_INV_LD_HINT) bulk_load_hint(ibuf_s + s * TB, src - t * TB, TB, bar_s + 8 * s, l2_policy<HUSL_INV_LD_HINT>());
                else bulk_load(ibuf_s + s * TB, src - t * TB, TB, bar_s + 8 * s);
            }
        }
    }
    __syncwarp();

    int st = 0, ost = 0;
    uint32_t phase = 0;
    for (long long tile = first; tile < n_tiles; tile += nw) {
        mbar_wait(bar_s + 8 * st, phase);
        float hsl[12];
        const unsigned char *rd = ibuf + st * TB + lane * Tile<TIn>::LANE;
        if (sizeof(TIn) == 4) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 f = *reinterpret_cast<const float4 *>(rd + 16 * k);
                hsl[4 * k] = f.x; hsl[4 * k + 1] = f.y; hsl[4 * k + 2] = f.z; hsl[4 * k + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double2 d = *reinterpret_cast<const double2 *>(rd + 16 * k);
                hsl[2 * k] = (float)d.x; hsl[2 * k + 1] = (float)d.y;
            }
        }
        __syncwarp();                       // every lane has its 4 pixels in registers
        if (lane == 0) {                    // refill this stage IST tiles ahead
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) {
                mbar_expect_tx(bar_s + 8 * st, TB);
                if (HUSL_INV_LD_HINT) bulk_load_hint(ibuf_s + st * TB, src - t * TB, TB, bar_s + 8 * st, l2_policy<HUSL_INV_LD_HINT>());
                else bulk_load(ibuf_s + st * TB, src - t * TB, TB, bar_s + 8 * st);
            }
        }
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            edit_on_load<ED>(ea, hsl[3 * j], hsl[3 * j + 1], hsl[3 * j + 2]);
            // the edit-on-load instantiation uses the fused kernels' formulation (k_edit), so that
            // to_rgb(to_husl(x), edit=e) == nphusl.edit(x, e) bit for bit for float64 HUSL as well
            rgb255_from_husl<false, wide_in<TIn>() && ED < 0>(hsl[3 * j], hsl[3 * j + 1], hsl[3 * j + 2], R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        const uint32_t o0 = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        const uint32_t o1 = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
        const uint32_t o2 = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
        if (lane == 0) bulk_wait_read<1>();          // the store issued two tiles ago has released its buffer
        __syncwarp();
        uint32_t *sw = reinterpret_cast<uint32_t *>(obuf + ost * U8_TILE);
        sw[lane * 3] = o0; sw[lane * 3 + 1] = o1; sw[lane * 3 + 2] = o2;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (HUSL_INV_ST_HINT) bulk_store_hint(dstb - tile * U8_TILE, obuf_s + ost * U8_TILE, U8_TILE, l2_policy<HUSL_INV_ST_HINT>());
            else bulk_store(dstb - tile * U8_TILE, obuf_s + ost * U8_TILE, U8_TILE);
            bulk_commit();
        }
        ost ^= 1;
        if (++st == IST) { st = 0; phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------
// (b) again, float64 HUSL with 32-byte aligned input: no ring at all
// ---------------------------------------------------------------------------------
// A thread reads its own 4 pixels (96 bytes) with three 256-bit loads (LDG.E.ENL2.256: every request is one whole DRAM
// sector) and writes its 12 result bytes with three 32-bit stores; no shared memory, no mbarriers, no bulk copies.  Groups are
// walked from the END of the image like k_inv_tma (the producer's last ~100 MB are still in L2).  Measured against the
// TMA ring (profiles/r02_ab_inv.txt, session 3): to_rgb 0.294 -> 0.280 ms stand-alone, the to_husl -> to_rgb(edit) step
// 0.598 -> 0.580 ms.  What matters is how many bytes an SM keeps in flight: one CTA of 1024 threads (32 warps x 3 KB) is the
// optimum, 768 x 1 and 512 x 2 are within 1 %, 16 warps or 36 - 64 warps per SM lose 10 - 25 %.
template <int BLOCK, int MINB, int ED = -1>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_direct(const double *__restrict__ in, uint32_t *__restrict__ out, long long n_groups, const EditArg<ED> ea = EditArg<ED>()) {
    const long long stride = (long long)gridDim.x * BLOCK;
    for (long long g = (long long)blockIdx.x * BLOCK + threadIdx.x; g < n_groups; g += stride) {
        const long long r = n_groups - 1 - g;
        double v[12];
        ldg_f64x4(in + 12 * r, v[0], v[1], v[2], v[3]);
        ldg_f64x4(in + 12 * r + 4, v[4], v[5], v[6], v[7]);
        ldg_f64x4(in + 12 * r + 8, v[8], v[9], v[10], v[11]);
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float h = (float)v[3 * j], s = (float)v[3 * j + 1], l = (float)v[3 * j + 2], R, G, B;
            edit_on_load<ED>(ea, h, s, l);
            rgb255_from_husl<false, wide_in<double>() && ED < 0>(h, s, l, R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        uint32_t *dst = out + 3 * r;
        dst[0] = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        dst[1] = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
        dst[2] = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
    }
}

// The float32 inverse without a ring: 4 pixels (48 bytes = three 128-bit loads, lanes 48 bytes apart: neighbouring requests
// share their sectors through L1) per thread.  Only the edit-free instantiation fits the 32 registers that 2 x 1024 threads per SM
// need, and only that shape beats the cp.async ring (0.1646 against 0.1694 ms; 1024 x 1, 512 x 2 ... 256 x 8: 0.177 - 0.21 ms), so
// to_rgb(edit=...) of float32 HUSL stays on k_inv_cp.
template <int BLOCK, int MINB, int ED = -1>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_direct32(const float4 *__restrict__ in, uint32_t *__restrict__ out, long long n_groups, const EditArg<ED> ea = EditArg<ED>()) {
    const long long stride = (long long)gridDim.x * BLOCK;
    for (long long g = (long long)blockIdx.x * BLOCK + threadIdx.x; g < n_groups; g += stride) {
        const long long r = n_groups - 1 - g;
        const float4 a = __ldg(in + 3 * r), b = __ldg(in + 3 * r + 1), c = __ldg(in + 3 * r + 2);
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float h = v[3 * j], s = v[3 * j + 1], l = v[3 * j + 2], R, G, B;
            edit_on_load<ED>(ea, h, s, l);
            rgb255_from_husl<false, false>(h, s, l, R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        uint32_t *dst = out + 3 * r;
        dst[0] = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        dst[1] = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
        dst[2] = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
    }
}

// ---------------------------------------------------------------------------------
// (b') rgb_from_husl, float32 HUSL: per-lane 16-byte async copies (LDGSTS) ring
// ---------------------------------------------------------------------------------
// The float32 inverse is bound by instruction issue (about 80 instructions per
// pixel against 15 bytes), so the ~60 warp-instructions of mbarrier / elect /
// proxy-fence bookkeeping the bulk engine needs per 128-pixel tile cost more
// than they save.  This variant keeps the asynchronous, register-free prefetch
// ring but fills it with three coalesced `cp.async.cg` 16-byte copies per lane
// (the shared-memory image equals the global one, so lanes read their 48-byte
// records conflict-free exactly as in k_inv_tma) and commits / waits with plain
// async-groups.  uint8 results go out through a 384-byte staging transposition.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int IST>
struct InvCpSmem {
    static constexpr int PER_WARP = IST * Tile<float>::BYTES + U8_TILE;
};

template <int BLOCK, int MINB, int IST, int ED = -1>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_cp(const float *__restrict__ in, uint32_t *__restrict__ out, long long n_tiles, const EditArg<ED> ea = EditArg<ED>()) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<float>::BYTES;
    unsigned char *ibuf = smem + warp * InvCpSmem<IST>::PER_WARP;
    uint32_t *sw = reinterpret_cast<uint32_t *>(ibuf + IST * TB);
    const uint32_t ld_s = smem_addr(ibuf) + lane * 16;       // chunk `lane` of a stage
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    const long long first = (long long)blockIdx.x * (BLOCK / 32) + warp;
    // walked from the end of the image (see k_inv_tma)
    const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + (n_tiles - 1) * (long long)TB + lane * 16;
    uint32_t *dst = out + (n_tiles - 1) * 96LL + lane;

#pragma unroll
    for (int s = 0; s < IST; ++s) {
        const long long t = first + s * nw;
        if (t < n_tiles) {
            const unsigned char *g = src - t * TB;
#pragma unroll
            for (int k = 0; k < 3; ++k) cp_async16(ld_s + s * TB + 512 * k, g + 512 * k);
        }
        cp_async_commit();
    }
    int st = 0;
    constexpr int UNROLL = HUSL_INVCP_UNROLL;
#pragma unroll UNROLL
    for (long long tile = first; tile < n_tiles; tile += nw) {
        cp_async_wait<IST - 1>();
        __syncwarp();
        float hsl[12];
        const unsigned char *rd = ibuf + st * TB + lane * 48;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 f = *reinterpret_cast<const float4 *>(rd + 16 * k);
            hsl[4 * k] = f.x; hsl[4 * k + 1] = f.y; hsl[4 * k + 2] = f.z; hsl[4 * k + 3] = f.w;
        }
        __syncwarp();
        {
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) {
                const unsigned char *g = src - t * TB;
#pragma unroll
                for (int k = 0; k < 3; ++k) cp_async16(ld_s + st * TB + 512 * k, g + 512 * k);
            }
            cp_async_commit();
        }
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            edit_on_load<ED>(ea, hsl[3 * j], hsl[3 * j + 1], hsl[3 * j + 2]);
            rgb255_from_husl(hsl[3 * j], hsl[3 * j + 1], hsl[3 * j + 2], R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        const uint32_t o0 = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        const uint32_t o1 = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
        const uint32_t o2 = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
        sw[lane * 3] = o0; sw[lane * 3 + 1] = o1; sw[lane * 3 + 2] = o2;
        __syncwarp();
        uint32_t *d = dst - tile * 96;
        d[0] = sw[lane]; d[32] = sw[lane + 32]; d[64] = sw[lane + 64];
        __syncwarp();
        st = (st + 1 == IST) ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// (b'') rgb_from_husl, float32 HUSL PLANES H[], S[], L[] (planar / channels-first callers)
// ---------------------------------------------------------------------------------
// The same LDGSTS ring as k_inv_cp; a stage holds the three 512-byte plane segments of
// one 128-pixel tile.  Every lane copies and later reads the SAME 16 bytes of each plane
// (its 4 pixels), so the ring needs no warp synchronisation at all.  CHW = true writes
// uint8 planes R[], G[], B[] (one word = the lane's 4 pixels), CHW = false interleaved
// RGB through the 384-byte staging transposition.
template <int BLOCK, int MINB, int IST, bool CHW>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_planes_cp(const float *__restrict__ H, const float *__restrict__ S, const float *__restrict__ L,
                uint32_t *__restrict__ o0p, uint32_t *__restrict__ o1p, uint32_t *__restrict__ o2p, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<float>::BYTES;                       // 3 x 512
    unsigned char *ibuf = smem + warp * InvCpSmem<IST>::PER_WARP;
    uint32_t *sw = reinterpret_cast<uint32_t *>(ibuf + IST * TB);
    const uint32_t ld_s = smem_addr(ibuf) + lane * 16;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    const long long first = (long long)blockIdx.x * (BLOCK / 32) + warp;
    // walked from the end of the image (see k_inv_tma)
    const long long last_px = (n_tiles - 1) * (long long)TILE_PX + lane * 4;
    const float *hp = H + last_px, *sp = S + last_px, *lp = L + last_px;

#pragma unroll
    for (int s = 0; s < IST; ++s) {
        const long long t = first + s * nw;
        if (t < n_tiles) {
            cp_async16(ld_s + s * TB, hp - t * TILE_PX);
            cp_async16(ld_s + s * TB + 512, sp - t * TILE_PX);
            cp_async16(ld_s + s * TB + 1024, lp - t * TILE_PX);
        }
        cp_async_commit();
    }
    int st = 0;
    for (long long tile = first; tile < n_tiles; tile += nw) {
        cp_async_wait<IST - 1>();
        const unsigned char *rd = ibuf + st * TB + lane * 16;
        const float4 h = *reinterpret_cast<const float4 *>(rd);
        const float4 s = *reinterpret_cast<const float4 *>(rd + 512);
        const float4 l = *reinterpret_cast<const float4 *>(rd + 1024);
        {
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) {
                cp_async16(ld_s + st * TB, hp - t * TILE_PX);
                cp_async16(ld_s + st * TB + 512, sp - t * TILE_PX);
                cp_async16(ld_s + st * TB + 1024, lp - t * TILE_PX);
            }
            cp_async_commit();
        }
        const float hh[4] = {h.x, h.y, h.z, h.w}, ss[4] = {s.x, s.y, s.z, s.w}, ll[4] = {l.x, l.y, l.z, l.w};
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            rgb255_from_husl(hh[j], ss[j], ll[j], R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        const long long rt = n_tiles - 1 - tile;                 // the tile this iteration really holds
        if (CHW) {
            const long long w = rt * 32 + lane;
            o0p[w] = __byte_perm(__byte_perm(q[0], q[3], 0x0040), __byte_perm(q[6], q[9], 0x0040), 0x5410);
            o1p[w] = __byte_perm(__byte_perm(q[1], q[4], 0x0040), __byte_perm(q[7], q[10], 0x0040), 0x5410);
            o2p[w] = __byte_perm(__byte_perm(q[2], q[5], 0x0040), __byte_perm(q[8], q[11], 0x0040), 0x5410);
        } else {
            const uint32_t a = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
            const uint32_t b = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
            const uint32_t c = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
            sw[lane * 3] = a; sw[lane * 3 + 1] = b; sw[lane * 3 + 2] = c;
            __syncwarp();
            uint32_t *d = o0p + rt * 96 + lane;
            d[0] = sw[lane]; d[32] = sw[lane + 32]; d[64] = sw[lane + 64];
            __syncwarp();
        }
        st = (st + 1 == IST) ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

}  // namespace husl
