// husl_float.cuh -- tiled streaming kernels for FLOAT RGB (SURVEY.md section 8a rows a2 / a3:
// the Cython backend's own signatures are float64 RGB in [0,1] -> HUSL,
// _cython_opt.pyx:22-93, and HUSL -> float64 RGB in [0,1], _cython_opt.pyx:177-246).
//
//   k_fwd_flt<TIn, TOut, HUE>   float32 / float64 interleaved RGB -> HUSL (or hue)
//   k_inv_flt<TIn, TOut>        HUSL -> float32 / float64 interleaved RGB in [0,1]
//
// The generic one-pixel-per-thread kernels (k_fwd_any / k_inv_any) touch 12 to 24
// cache lines per warp instruction on these 12 / 24-byte-per-pixel records.  Here both
// sides move like the wide side of k_inv_cp / k_fwd_u8: a warp owns 128 consecutive
// pixels; the input tile (1536 or 3072 contiguous bytes) arrives through a ring of
// per-lane 16-byte cp.async copies whose shared-memory image equals the global one
// (lane l reads its 4 pixels at byte l * LANE), results leave through the per-warp
// staging transposition as coalesced 16-byte stores.  Float inputs are linearised in
// float64 (husl_math.cuh: lin_f64) and keep the explicit lightness clamps.
#pragma once
#include "husl_stream.cuh"

namespace husl {

template <typename TIn, typename TOut, int IST>
struct FltSmem {
    static constexpr int PER_WARP = IST * Tile<TIn>::BYTES + Pack<TOut>::STAGE;
};
// hue output needs no staging (one vector per lane)
template <typename TIn, int IST>
struct FltHueSmem {
    static constexpr int PER_WARP = IST * Tile<TIn>::BYTES;
};

template <typename T>
__device__ __forceinline__ void ring_fill(uint32_t dst_s, const unsigned char *g) {
#pragma unroll
    for (int k = 0; k < Tile<T>::BYTES / 512; ++k) cp_async16(dst_s + 512 * k, g + 512 * k);
}

// lane's 12 values (4 pixels x 3) out of a tile image in shared memory
template <typename T>
__device__ __forceinline__ void read12(const unsigned char *rd, double (&v)[12]) {
    if (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 f = *reinterpret_cast<const float4 *>(rd + 16 * k);
            v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double2 d = *reinterpret_cast<const double2 *>(rd + 16 * k);
            v[2 * k] = d.x; v[2 * k + 1] = d.y;
        }
    }
}

// the lane's 12 values in their own type (float channels reach lin_f64's float overload unconverted)
template <typename T>
__device__ __forceinline__ void read12n(const unsigned char *rd, T (&v)[12]) {
    constexpr int PER = 16 / (int)sizeof(T);
#pragma unroll
    for (int k = 0; k < 12 / PER; ++k) {
        const float4 raw = *reinterpret_cast<const float4 *>(rd + 16 * k);
        const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int q = 0; q < PER; ++q) v[PER * k + q] = e[q];
    }
}

template <typename T>
__device__ __forceinline__ void read12f(const unsigned char *rd, float (&v)[12]) {
    if (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 f = *reinterpret_cast<const float4 *>(rd + 16 * k);
            v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double2 d = *reinterpret_cast<const double2 *>(rd + 16 * k);
            v[2 * k] = (float)d.x; v[2 * k + 1] = (float)d.y;
        }
    }
}

// lane's 12 results -> staging (padded record layout) -> coalesced 16-byte stores of the tile
template <typename TOut>
__device__ __forceinline__ void store12(unsigned char *stage, int lane, const float (&r)[12], TOut *tile_out) {
    constexpr int CH = Pack<TOut>::CHUNKS;
    unsigned char *wr = stage + lane * Pack<TOut>::LANE_STRIDE;
    if (sizeof(TOut) == 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            *reinterpret_cast<float4 *>(wr + 16 * k) = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            *reinterpret_cast<double2 *>(wr + 16 * k) = make_double2((double)r[2 * k], (double)r[2 * k + 1]);
    }
    __syncwarp();
    float4 *o = reinterpret_cast<float4 *>(tile_out) + lane;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int p = 32 * k + lane;
        o[32 * k] = *reinterpret_cast<const float4 *>(stage + (p / CH) * Pack<TOut>::LANE_STRIDE + (p % CH) * 16);
    }
    __syncwarp();
}

// one pixel's four quantities from its three channels: float64 channels are linearised in float64
// (lin_f64), float32 channels go the float64-free way (diff_from_f32)
__device__ __forceinline__ Diff diff_from_rgb(double r, double g, double b) {
    LinPx p;
    lin_f64(r, p.hr, p.lr); lin_f64(g, p.hg, p.lg); lin_f64(b, p.hb, p.lb);
    return make_diff(p);
}
__device__ __forceinline__ Diff diff_from_rgb(float r, float g, float b) { return diff_from_f32(r, g, b); }

template <typename TIn, typename TOut, bool HUE, int BLOCK, int MINB, int IST>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_flt(const TIn *__restrict__ in, TOut *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<TIn>::BYTES;
    constexpr int PER_WARP = HUE ? FltHueSmem<TIn, IST>::PER_WARP : FltSmem<TIn, TOut, IST>::PER_WARP;
    unsigned char *ibuf = smem + warp * PER_WARP;
    unsigned char *stage = ibuf + IST * TB;
    const uint32_t ld_s = smem_addr(ibuf) + lane * 16;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    const long long first = (long long)blockIdx.x * (BLOCK / 32) + warp;
    const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + lane * 16;

#pragma unroll
    for (int s = 0; s < IST; ++s) {
        const long long t = first + s * nw;
        if (t < n_tiles) ring_fill<TIn>(ld_s + s * TB, src + t * TB);
        cp_async_commit();
    }
    int st = 0;
    for (long long tile = first; tile < n_tiles; tile += nw) {
        cp_async_wait<IST - 1>();
        __syncwarp();
        TIn v[12];
        read12n<TIn>(ibuf + st * TB + lane * Tile<TIn>::LANE, v);
        __syncwarp();
        {
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) ring_fill<TIn>(ld_s + st * TB, src + t * TB);
            cp_async_commit();
        }
        if (HUE) {
            float h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = hue_from_diff(diff_from_rgb(v[3 * j], v[3 * j + 1], v[3 * j + 2]));
            if (sizeof(TOut) == 4)
                *(reinterpret_cast<float4 *>(out) + tile * 32 + lane) = make_float4(h[0], h[1], h[2], h[3]);
            else
                stg_f64x4(reinterpret_cast<double *>(out) + tile * 128 + lane * 4, (double)h[0], (double)h[1], (double)h[2], (double)h[3]);
        } else {
            float r[12];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                husl_from_diff<false>(diff_from_rgb(v[3 * j], v[3 * j + 1], v[3 * j + 2]), r[3 * j], r[3 * j + 1], r[3 * j + 2]);
            store12<TOut>(stage, lane, r, out + tile * (TILE_PX * 3));
        }
        st = (st + 1 == IST) ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

template <typename TIn, typename TOut, int BLOCK, int MINB, int IST>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_flt(const TIn *__restrict__ in, TOut *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<TIn>::BYTES;
    unsigned char *ibuf = smem + warp * FltSmem<TIn, TOut, IST>::PER_WARP;
    unsigned char *stage = ibuf + IST * TB;
    const uint32_t ld_s = smem_addr(ibuf) + lane * 16;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    const long long first = (long long)blockIdx.x * (BLOCK / 32) + warp;
    const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + lane * 16;

#pragma unroll
    for (int s = 0; s < IST; ++s) {
        const long long t = first + s * nw;
        if (t < n_tiles) ring_fill<TIn>(ld_s + s * TB, src + t * TB);
        cp_async_commit();
    }
    int st = 0;
    for (long long tile = first; tile < n_tiles; tile += nw) {
        cp_async_wait<IST - 1>();
        __syncwarp();
        float v[12];
        read12f<TIn>(ibuf + st * TB + lane * Tile<TIn>::LANE, v);
        __syncwarp();
        {
            const long long t = tile + (long long)IST * nw;
            if (t < n_tiles) ring_fill<TIn>(ld_s + st * TB, src + t * TB);
            cp_async_commit();
        }
        float r[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            rgb255_from_husl<true, wide_in<TIn>()>(v[3 * j], v[3 * j + 1], v[3 * j + 2], R, G, B);
            r[3 * j] = R * (1.0f / 255.0f); r[3 * j + 1] = G * (1.0f / 255.0f); r[3 * j + 2] = B * (1.0f / 255.0f);
        }
        store12<TOut>(stage, lane, r, out + tile * (TILE_PX * 3));
        st = (st + 1 == IST) ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

// float64 HUSL -> float64 RGB in [0,1] (what the reference's _husl_to_rgb returns, _cython_opt.pyx:177-246) with
// 256-bit accesses: a thread owns 4 pixels = 96 bytes on either side, three LDG.256 in, three STG.256 out, no shared
// memory at all -- the access pattern that took the explicit edit pass from 0.76 to 0.93 of the copy peak
// (husl_edit.cuh: k_husl_edit_wide).  32-byte aligned bases; n_groups = n / 4, the rest goes to k_inv_any.
// `out` may be `in` (a thread reads its 96 bytes before it writes them): plain loads, no __restrict__.
__global__ void __launch_bounds__(256)
k_inv_flt_wide(const double *in, double *out, long long n_groups) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        double v[12];
        ld_f64x4(in + 12 * g, v[0], v[1], v[2], v[3]);
        ld_f64x4(in + 12 * g + 4, v[4], v[5], v[6], v[7]);
        ld_f64x4(in + 12 * g + 8, v[8], v[9], v[10], v[11]);
        double r[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            rgb255_from_husl<true, wide_in<double>()>((float)v[3 * j], (float)v[3 * j + 1], (float)v[3 * j + 2], R, G, B);
            r[3 * j] = (double)(R * (1.0f / 255.0f)); r[3 * j + 1] = (double)(G * (1.0f / 255.0f)); r[3 * j + 2] = (double)(B * (1.0f / 255.0f));
        }
        stg_f64x4(out + 12 * g, r[0], r[1], r[2], r[3]);
        stg_f64x4(out + 12 * g + 4, r[4], r[5], r[6], r[7]);
        stg_f64x4(out + 12 * g + 8, r[8], r[9], r[10], r[11]);
    }
}

}  // namespace husl
