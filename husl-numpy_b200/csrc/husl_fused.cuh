// husl_fused.cuh -- the callers' side of the hot path, moved onto the GPU
// (SURVEY.md section 8f "next" rows 1 and 2):
//
//   k_fwd_planar / k_inv_planar   HUSL kept as three planes (H[], S[], L[])
//       instead of interleaved triplets: what callers slice out anyway
//       (reference README.md:77-98 works on hsl[..., 0] and hsl[..., 2]); every
//       plane access is a natural 16 / 32-byte vector per lane, no transposition.
//   k_edit                        to_husl -> select -> affine H/S/L edit -> to_rgb
//       in ONE pass over uint8 pixels (6 bytes of HBM traffic per pixel instead
//       of 15 + 24 + 15 for three calls with a float32 intermediate).
#pragma once
#include "husl_kernels.cuh"
#include "husl_edit.cuh"

#ifndef HUSL_PLANAR_UNROLL
#define HUSL_PLANAR_UNROLL 1   // tile loops of k_fwd_planar / k_fwd_chw (float32 planes) unrolled by this factor (A/B knob)
#endif
#ifndef HUSL_EDITCHW_UNROLL
#define HUSL_EDITCHW_UNROLL 2  // tile loop of k_edit_chw: 0.2998 -> 0.2975 ms
#endif
#ifndef HUSL_EDIT_UNROLL
#define HUSL_EDIT_UNROLL 2     // tile loop of k_edit unrolled by two: 0.2920 -> 0.2894 ms (interleaved A/B)
#endif

namespace husl {


// lanes hold 4 consecutive pixels: unpack 3 words -> 4 x LinPx through the shared table
__device__ __forceinline__ void lookup4(const unsigned char *tab, uint32_t w0, uint32_t w1, uint32_t w2,
                                        uint32_t lane8, LinPx (&px)[4]) {
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

__device__ __forceinline__ void pack12(const uint32_t (&q)[12], uint32_t &o0, uint32_t &o1, uint32_t &o2) {
    o0 = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
    o1 = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
    o2 = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
}

// ---------------------------------------------------------------------------------
// planar forward: uint8 interleaved RGB -> H[], S[], L[]
// ---------------------------------------------------------------------------------
template <typename TOut, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_planar(const uint32_t *__restrict__ in, TOut *__restrict__ H, TOut *__restrict__ S, TOut *__restrict__ L,
             long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) {
        const uint32_t *p = in + tile * 96 + lane * 3;
        w0 = ldg_u32(p); w1 = ldg_u32(p + 1); w2 = ldg_u32(p + 2);
    }
    constexpr int UNROLL = sizeof(TOut) == 4 ? HUSL_PLANAR_UNROLL : 1;
#pragma unroll UNROLL
    for (; tile < n_tiles; tile += nw) {
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (tile + nw < n_tiles) {
            const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
            n0 = ldg_u32(p); n1 = ldg_u32(p + 1); n2 = ldg_u32(p + 2);
        }
        LinPx px[4];
        lookup4(smem, w0, w1, w2, lane8, px);
        float h[4], s[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) husl_from_diff<true>(make_diff(px[j]), h[j], s[j], l[j]);
        const long long o = tile * TILE_PX + lane * 4;
        if (sizeof(TOut) == 4) {
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(H) + o) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(S) + o) = make_float4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(L) + o) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            stg_f64x4(reinterpret_cast<double *>(H) + o, h[0], h[1], h[2], h[3]);
            stg_f64x4(reinterpret_cast<double *>(S) + o, s[0], s[1], s[2], s[3]);
            stg_f64x4(reinterpret_cast<double *>(L) + o, l[0], l[1], l[2], l[3]);
        }
        w0 = n0; w1 = n1; w2 = n2;
    }
}

// ---------------------------------------------------------------------------------
// planar inverse: H[], S[], L[] -> uint8 interleaved RGB
// ---------------------------------------------------------------------------------
template <typename TIn, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_planar(const TIn *__restrict__ H, const TIn *__restrict__ S, const TIn *__restrict__ L,
             uint32_t *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sw = reinterpret_cast<uint32_t *>(smem) + warp * 96;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    for (long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp; tile < n_tiles; tile += nw) {
        const long long o = tile * TILE_PX + lane * 4;
        float h[4], s[4], l[4];
        if (sizeof(TIn) == 4) {
            const float4 a = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(H) + o));
            const float4 b = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(S) + o));
            const float4 c = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(L) + o));
            h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w;
            s[0] = b.x; s[1] = b.y; s[2] = b.z; s[3] = b.w;
            l[0] = c.x; l[1] = c.y; l[2] = c.z; l[3] = c.w;
        } else {
            double d[12];
            ldg_f64x4(reinterpret_cast<const double *>(H) + o, d[0], d[1], d[2], d[3]);
            ldg_f64x4(reinterpret_cast<const double *>(S) + o, d[4], d[5], d[6], d[7]);
            ldg_f64x4(reinterpret_cast<const double *>(L) + o, d[8], d[9], d[10], d[11]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { h[j] = (float)d[j]; s[j] = (float)d[4 + j]; l[j] = (float)d[8 + j]; }
        }
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            rgb255_from_husl<false, wide_in<TIn>()>(h[j], s[j], l[j], R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        uint32_t o0, o1, o2;
        pack12(q, o0, o1, o2);
        sw[lane * 3] = o0; sw[lane * 3 + 1] = o1; sw[lane * 3 + 2] = o2;
        __syncwarp();
        uint32_t *dst = out + tile * 96 + lane;
        dst[0] = sw[lane]; dst[32] = sw[lane + 32]; dst[64] = sw[lane + 64];
        __syncwarp();
    }
}

// one pixel per thread: tails, unaligned planes
template <typename T>
__global__ void __launch_bounds__(256)
k_fwd_planar_any(const uint8_t *__restrict__ in, T *__restrict__ H, T *__restrict__ S, T *__restrict__ L, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float h, s, l;
        husl_from_diff<true>(make_diff(load_px<uint8_t>(in, i, 3)), h, s, l);
        H[i] = (T)h; S[i] = (T)s; L[i] = (T)l;
    }
}
template <typename T>
__global__ void __launch_bounds__(256)
k_inv_planar_any(const T *__restrict__ H, const T *__restrict__ S, const T *__restrict__ L, uint8_t *__restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float R, G, B;
        rgb255_from_husl<false, wide_in<T>()>((float)H[i], (float)S[i], (float)L[i], R, G, B);
        out[3 * i] = (uint8_t)(round_u8_bits(R) & 0xFF);
        out[3 * i + 1] = (uint8_t)(round_u8_bits(G) & 0xFF);
        out[3 * i + 2] = (uint8_t)(round_u8_bits(B) & 0xFF);
    }
}

// ---------------------------------------------------------------------------------
// fused edit: uint8 RGB -> HUSL (registers) -> select + affine edit -> uint8 RGB
// ---------------------------------------------------------------------------------
// (selection / edit rule, apply_edit and EditShape: husl_edit.cuh)
template <int BLOCK, int MINB, bool EH, bool ES, bool RSL>
__global__ void __launch_bounds__(BLOCK, MINB)
k_edit(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long n_tiles, const nphusl_edit e) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) {
        const uint32_t *p = in + tile * 96 + lane * 3;
        w0 = ldg_u32(p); w1 = ldg_u32(p + 1); w2 = ldg_u32(p + 2);
    }
    constexpr int UNROLL = HUSL_EDIT_UNROLL;
#pragma unroll UNROLL
    for (; tile < n_tiles; tile += nw) {
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (tile + nw < n_tiles) {
            const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
            n0 = ldg_u32(p); n1 = ldg_u32(p + 1); n2 = ldg_u32(p + 2);
        }
        LinPx px[4];
        lookup4(smem, w0, w1, w2, lane8, px);
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float h, s, l, R, G, B;
            husl_from_diff<true>(make_diff(px[j]), h, s, l);
            apply_edit<EH, ES, RSL>(e, h, s, l);
            rgb255_from_husl(h, s, l, R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        uint32_t o0, o1, o2;
        pack12(q, o0, o1, o2);
        // the lane's own 12 bytes straight to global: for this issue-bound kernel three strided 32-bit stores
        // beat the staging transposition (6 shared-memory instructions + 2 warp syncs per tile) by 1 %
        uint32_t *dst = out + tile * 96 + lane * 3;
        dst[0] = o0; dst[1] = o1; dst[2] = o2;
        w0 = n0; w1 = n1; w2 = n2;
    }
}

__global__ void __launch_bounds__(256)
k_edit_any(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n, const nphusl_edit e) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float h, s, l, R, G, B;
        husl_from_diff<true>(make_diff(load_px<uint8_t>(in, i, 3)), h, s, l);
        apply_edit(e, h, s, l);
        rgb255_from_husl(h, s, l, R, G, B);
        out[3 * i] = (uint8_t)(round_u8_bits(R) & 0xFF);
        out[3 * i + 1] = (uint8_t)(round_u8_bits(G) & 0xFF);
        out[3 * i + 2] = (uint8_t)(round_u8_bits(B) & 0xFF);
    }
}

// ---------------------------------------------------------------------------------
// fused banded map: uint8 RGB -> HUSL (registers) -> bands (husl_edit.cuh: apply_bands) -> uint8 RGB,
// the callers' hue_watermelon (examples.py:56-71) in one pass; k_edit's data movement
// ---------------------------------------------------------------------------------
template <int BLOCK, int MINB, bool SAT>
__global__ void __launch_bounds__(BLOCK, MINB)
k_bands(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long n_tiles, const BandsDev a) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
    if (tile < n_tiles) {
        const uint32_t *p = in + tile * 96 + lane * 3;
        w0 = ldg_u32(p); w1 = ldg_u32(p + 1); w2 = ldg_u32(p + 2);
    }
    for (; tile < n_tiles; tile += nw) {
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        if (tile + nw < n_tiles) {
            const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
            n0 = ldg_u32(p); n1 = ldg_u32(p + 1); n2 = ldg_u32(p + 2);
        }
        LinPx px[4];
        lookup4(smem, w0, w1, w2, lane8, px);
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float h, s, l, R, G, B;
            husl_from_diff<true>(make_diff(px[j]), h, s, l);
            apply_bands<SAT>(a, h, s, l);
            rgb255_from_husl(h, s, l, R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        uint32_t o0, o1, o2;
        pack12(q, o0, o1, o2);
        uint32_t *dst = out + tile * 96 + lane * 3;
        dst[0] = o0; dst[1] = o1; dst[2] = o2;
        w0 = n0; w1 = n1; w2 = n2;
    }
}

__global__ void __launch_bounds__(256)
k_bands_any(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n, const BandsDev a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float h, s, l, R, G, B;
        husl_from_diff<true>(make_diff(load_px<uint8_t>(in, i, 3)), h, s, l);
        apply_bands<true>(a, h, s, l);
        rgb255_from_husl(h, s, l, R, G, B);
        out[3 * i] = (uint8_t)(round_u8_bits(R) & 0xFF);
        out[3 * i + 1] = (uint8_t)(round_u8_bits(G) & 0xFF);
        out[3 * i + 2] = (uint8_t)(round_u8_bits(B) & 0xFF);
    }
}

// ---------------------------------------------------------------------------------
// channels-first images (torch / vision layout): R[], G[], B[] uint8 planes <-> H[], S[], L[] planes
// ---------------------------------------------------------------------------------
// Every access is one natural vector per lane per plane (a 32-bit word = 4 pixels
// of a uint8 plane, 16 / 32 bytes of a float / double plane): no staging at all.
template <typename TOut, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_chw(const uint32_t *__restrict__ R, const uint32_t *__restrict__ G, const uint32_t *__restrict__ B,
          TOut *__restrict__ H, TOut *__restrict__ S, TOut *__restrict__ L, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;
    uint32_t wr = 0, wg = 0, wb = 0;
    if (tile < n_tiles) { const long long o = tile * 32 + lane; wr = ldg_u32(R + o); wg = ldg_u32(G + o); wb = ldg_u32(B + o); }
    constexpr int UNROLL = sizeof(TOut) == 4 ? HUSL_PLANAR_UNROLL : 1;
#pragma unroll UNROLL
    for (; tile < n_tiles; tile += nw) {
        uint32_t nr = 0, ng = 0, nb = 0;
        if (tile + nw < n_tiles) { const long long o = (tile + nw) * 32 + lane; nr = ldg_u32(R + o); ng = ldg_u32(G + o); nb = ldg_u32(B + o); }
        LinPx px[4];
        float2 v;
#define CHW_PX(J)                                                                  \
        v = lut_lin<J>(smem, wr, lane8); px[J].hr = v.x; px[J].lr = v.y;           \
        v = lut_lin<J>(smem, wg, lane8); px[J].hg = v.x; px[J].lg = v.y;           \
        v = lut_lin<J>(smem, wb, lane8); px[J].hb = v.x; px[J].lb = v.y;
        CHW_PX(0) CHW_PX(1) CHW_PX(2) CHW_PX(3)
#undef CHW_PX
        float h[4], s[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) husl_from_diff<true>(make_diff(px[j]), h[j], s[j], l[j]);
        const long long o = tile * TILE_PX + lane * 4;
        if (sizeof(TOut) == 4) {
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(H) + o) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(S) + o) = make_float4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(L) + o) = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            stg_f64x4(reinterpret_cast<double *>(H) + o, h[0], h[1], h[2], h[3]);
            stg_f64x4(reinterpret_cast<double *>(S) + o, s[0], s[1], s[2], s[3]);
            stg_f64x4(reinterpret_cast<double *>(L) + o, l[0], l[1], l[2], l[3]);
        }
        wr = nr; wg = ng; wb = nb;
    }
}

template <typename TIn, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_chw(const TIn *__restrict__ H, const TIn *__restrict__ S, const TIn *__restrict__ L,
          uint32_t *__restrict__ R, uint32_t *__restrict__ G, uint32_t *__restrict__ B, long long n_tiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    for (long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp; tile < n_tiles; tile += nw) {
        const long long o = tile * TILE_PX + lane * 4;
        float h[4], s[4], l[4];
        if (sizeof(TIn) == 4) {
            const float4 a = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(H) + o));
            const float4 b = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(S) + o));
            const float4 c = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(L) + o));
            h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w;
            s[0] = b.x; s[1] = b.y; s[2] = b.z; s[3] = b.w;
            l[0] = c.x; l[1] = c.y; l[2] = c.z; l[3] = c.w;
        } else {
            double d[12];
            ldg_f64x4(reinterpret_cast<const double *>(H) + o, d[0], d[1], d[2], d[3]);
            ldg_f64x4(reinterpret_cast<const double *>(S) + o, d[4], d[5], d[6], d[7]);
            ldg_f64x4(reinterpret_cast<const double *>(L) + o, d[8], d[9], d[10], d[11]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { h[j] = (float)d[j]; s[j] = (float)d[4 + j]; l[j] = (float)d[8 + j]; }
        }
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r, g, b;
            rgb255_from_husl<false, wide_in<TIn>()>(h[j], s[j], l[j], r, g, b);
            q[3 * j] = round_u8_bits(r); q[3 * j + 1] = round_u8_bits(g); q[3 * j + 2] = round_u8_bits(b);
        }
        // plane words: byte j of a word = pixel j of this lane
        const long long w = tile * 32 + lane;
        R[w] = __byte_perm(__byte_perm(q[0], q[3], 0x0040), __byte_perm(q[6], q[9], 0x0040), 0x5410);
        G[w] = __byte_perm(__byte_perm(q[1], q[4], 0x0040), __byte_perm(q[7], q[10], 0x0040), 0x5410);
        B[w] = __byte_perm(__byte_perm(q[2], q[5], 0x0040), __byte_perm(q[8], q[11], 0x0040), 0x5410);
    }
}

// tails / unaligned planes, one pixel per thread
template <typename T>
__global__ void __launch_bounds__(256)
k_fwd_chw_any(const uint8_t *__restrict__ R, const uint8_t *__restrict__ G, const uint8_t *__restrict__ B,
              T *__restrict__ H, T *__restrict__ S, T *__restrict__ L, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 r = g_lin_tab[R[i]], g = g_lin_tab[G[i]], b = g_lin_tab[B[i]];
        LinPx p;
        p.hr = r.x; p.lr = r.y; p.hg = g.x; p.lg = g.y; p.hb = b.x; p.lb = b.y;
        float h, s, l;
        husl_from_diff<true>(make_diff(p), h, s, l);
        H[i] = (T)h; S[i] = (T)s; L[i] = (T)l;
    }
}
template <typename T>
__global__ void __launch_bounds__(256)
k_inv_chw_any(const T *__restrict__ H, const T *__restrict__ S, const T *__restrict__ L,
              uint8_t *__restrict__ R, uint8_t *__restrict__ G, uint8_t *__restrict__ B, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float r, g, b;
        rgb255_from_husl<false, wide_in<T>()>((float)H[i], (float)S[i], (float)L[i], r, g, b);
        R[i] = (uint8_t)(round_u8_bits(r) & 0xFF);
        G[i] = (uint8_t)(round_u8_bits(g) & 0xFF);
        B[i] = (uint8_t)(round_u8_bits(b) & 0xFF);
    }
}

// fused edit on channels-first uint8 planes (what nvJPEG / video decoders hand out)
template <int BLOCK, int MINB, bool EH, bool ES, bool RSL>
__global__ void __launch_bounds__(BLOCK, MINB)
k_edit_chw(const uint32_t *__restrict__ R, const uint32_t *__restrict__ G, const uint32_t *__restrict__ B,
           uint32_t *__restrict__ Ro, uint32_t *__restrict__ Go, uint32_t *__restrict__ Bo,
           long long n_tiles, const nphusl_edit e) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    constexpr int UNROLL = HUSL_EDITCHW_UNROLL;
#pragma unroll UNROLL
    for (long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp; tile < n_tiles; tile += nw) {
        const long long w = tile * 32 + lane;
        const uint32_t wr = ldg_u32(R + w), wg = ldg_u32(G + w), wb = ldg_u32(B + w);
        LinPx px[4];
        float2 v;
#define CHW_PX(J)                                                                  \
        v = lut_lin<J>(smem, wr, lane8); px[J].hr = v.x; px[J].lr = v.y;           \
        v = lut_lin<J>(smem, wg, lane8); px[J].hg = v.x; px[J].lg = v.y;           \
        v = lut_lin<J>(smem, wb, lane8); px[J].hb = v.x; px[J].lb = v.y;
        CHW_PX(0) CHW_PX(1) CHW_PX(2) CHW_PX(3)
#undef CHW_PX
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float h, s, l, r, g, b;
            husl_from_diff<true>(make_diff(px[j]), h, s, l);
            apply_edit<EH, ES, RSL>(e, h, s, l);
            rgb255_from_husl(h, s, l, r, g, b);
            q[3 * j] = round_u8_bits(r); q[3 * j + 1] = round_u8_bits(g); q[3 * j + 2] = round_u8_bits(b);
        }
        Ro[w] = __byte_perm(__byte_perm(q[0], q[3], 0x0040), __byte_perm(q[6], q[9], 0x0040), 0x5410);
        Go[w] = __byte_perm(__byte_perm(q[1], q[4], 0x0040), __byte_perm(q[7], q[10], 0x0040), 0x5410);
        Bo[w] = __byte_perm(__byte_perm(q[2], q[5], 0x0040), __byte_perm(q[8], q[11], 0x0040), 0x5410);
    }
}

__global__ void __launch_bounds__(256)
k_edit_chw_any(const uint8_t *__restrict__ R, const uint8_t *__restrict__ G, const uint8_t *__restrict__ B,
               uint8_t *__restrict__ Ro, uint8_t *__restrict__ Go, uint8_t *__restrict__ Bo, long long n, const nphusl_edit e) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 r = g_lin_tab[R[i]], g = g_lin_tab[G[i]], b = g_lin_tab[B[i]];
        LinPx p;
        p.hr = r.x; p.lr = r.y; p.hg = g.x; p.lg = g.y; p.hb = b.x; p.lb = b.y;
        float h, s, l, fr, fg, fb;
        husl_from_diff<true>(make_diff(p), h, s, l);
        apply_edit(e, h, s, l);
        rgb255_from_husl(h, s, l, fr, fg, fb);
        Ro[i] = (uint8_t)(round_u8_bits(fr) & 0xFF);
        Go[i] = (uint8_t)(round_u8_bits(fg) & 0xFF);
        Bo[i] = (uint8_t)(round_u8_bits(fb) & 0xFF);
    }
}

}  // namespace husl
