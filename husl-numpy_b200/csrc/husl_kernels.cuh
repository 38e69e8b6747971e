// husl_kernels.cuh -- the sm_100a kernels of the HUSL hot path.
//
//   k_fwd_u8<TOut, HUE>   (a) husl_from_rgb / (c) hue_from_rgb, uint8 RGB in
//   k_inv<TIn>            (b) rgb_from_husl, uint8 RGB out (to_rgb_int fused)
//   k_fwd_any / k_inv_any generic one-pixel-per-thread versions: float / gray
//                         inputs (float64 linearisation), float RGB outputs,
//                         tails and unaligned pointers
//
// Data movement of the two streaming kernels (no tensor cores: a 3x3 matrix per
// pixel is not a dense contraction; the limits are HBM bytes and FP32/MUFU
// issue):
//
//   * a warp owns a tile of 128 consecutive pixels, 4 per lane, and walks the
//     image with a grid stride; grids are sized as SMs x resident CTAs;
//   * interleaved uint8 RGB is read as three coalesced 32-bit words per lane
//     (12 bytes = 4 pixels, so no pixel straddles a lane) and unpacked with one
//     PRMT per channel that directly forms the shared-memory table address;
//   * the sRGB linearisation is a 256-entry table of hi+lo float pairs held in
//     shared memory, replicated once per lane (32 x 2 KB) so that every lookup
//     is a single conflict-free 64-bit access;
//   * results are transposed through a per-warp shared-memory staging buffer so
//     that every global access is a fully coalesced 16-byte vector per lane
//     (lane-contiguous 12 / 24 / 48 / 96-byte records would otherwise touch 12
//     to 24 cache lines per instruction);
//   * float64 I/O (the reference's dtype) is converted in registers; the
//     arithmetic is float32 (+ hi/lo pairs where it matters).
#pragma once
#include "husl_math.cuh"
#include "husl_edit.cuh"

#ifndef HUSL_FWD_UNROLL
#define HUSL_FWD_UNROLL 2   // tile loop of k_fwd_u8<.., HUE = false> unrolled by this factor
#endif
#ifndef HUSL_FWD_PF2
#define HUSL_FWD_PF2 0      // A/B knob: register prefetch of k_fwd_u8 two tiles ahead instead of one
#endif

namespace husl {

// 256 x {hi, lo}: hi = float(lin), lo = float(lin - hi); filled per device by
// the host from float64 (nphusl.py:278-286 formula at c/255).
__device__ float2 g_lin_tab[256];

constexpr int TILE_PX = 128;            // pixels per warp tile (4 per lane)
constexpr int TAB_BYTES = 256 * 32 * 8; // replicated table: [value][lane] float2

__device__ __forceinline__ uint32_t ldg_u32(const uint32_t *p) { return __ldg(p); }

// 256-bit global store (PTX 8.8+, sm_100+); p must be 32-byte aligned
__device__ __forceinline__ void ldg_f64x4(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// the same through the coherent path (buffers the kernel may also write)
__device__ __forceinline__ void ld_f64x4(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg_f64x4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// table lookup for byte K of word w: ONE PRMT builds (byte << 8) | lane*8, the
// byte offset of tab[byte][lane]
template <int K>
__device__ __forceinline__ float2 lut_lin(const unsigned char *tab, uint32_t w, uint32_t lane8) {
    const uint32_t off = __byte_perm(w, lane8, 0x5504 | (K << 4));
    return *reinterpret_cast<const float2 *>(tab + off);
}

template <typename T> struct Pack;
template <> struct Pack<float> {
    static constexpr int CHUNKS = 3;        // 16-byte chunks per lane per tile (4 px x 3 x 4 B)
    static constexpr int LANE_STRIDE = 48;  // 12 l mod 32 over a quarter warp: conflict-free STS.128
    static constexpr int STAGE = 32 * 48;   // staging bytes per warp
};
template <> struct Pack<double> {
    static constexpr int CHUNKS = 6;
    static constexpr int LANE_STRIDE = 112; // 96 B padded to 112: 28 l mod 32 is conflict-free
    static constexpr int STAGE = 32 * 112;
};

__device__ __forceinline__ void fill_table(unsigned char *smem) {
    float2 *tab = reinterpret_cast<float2 *>(smem);
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = g_lin_tab[i >> 5];
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// (a) husl_from_rgb and (c) hue_from_rgb, uint8 interleaved RGB
// ---------------------------------------------------------------------------------
template <typename TOut, bool HUE, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_u8(const uint32_t *__restrict__ in, TOut *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const unsigned char *tab = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    unsigned char *stage = smem + TAB_BYTES + warp * Pack<TOut>::STAGE;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp;

    // loop-invariant staging offsets: write = lane-contiguous record, read =
    // chunk 32k+lane of the tile (record p/CHUNKS, padded stride for double)
    constexpr int CH = Pack<TOut>::CHUNKS;
    uint32_t rd_off[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int p = 32 * k + lane;
        rd_off[k] = (p / CH) * Pack<TOut>::LANE_STRIDE + (p % CH) * 16;
    }

    // register prefetch of the next tile's words (PF2: two tiles ahead)
    constexpr bool PF2 = HUSL_FWD_PF2;   // measured: a second tile in flight does not pay (profiles/r01_ab_variants.txt)
    uint32_t w0 = 0, w1 = 0, w2 = 0, x0 = 0, x1 = 0, x2 = 0;
    if (tile < n_tiles) {
        const uint32_t *p = in + tile * 96 + lane * 3;
        w0 = ldg_u32(p); w1 = ldg_u32(p + 1); w2 = ldg_u32(p + 2);
    }
    if (PF2 && tile + nw < n_tiles) {
        const uint32_t *p = in + (tile + nw) * 96 + lane * 3;
        x0 = ldg_u32(p); x1 = ldg_u32(p + 1); x2 = ldg_u32(p + 2);
    }
    // unrolled by two for the HUSL output (the w = n register moves become renaming: 0.1737 -> 0.1714 ms);
    // the hue-only loop is faster rolled (0.1005 vs 0.1050)
    constexpr int UNROLL = HUE ? 1 : HUSL_FWD_UNROLL;
#pragma unroll UNROLL
    for (; tile < n_tiles; tile += nw) {
        // prefetch the words of the tile 1 (2) iterations ahead while this one is computed
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        const long long ahead = tile + (PF2 ? 2 : 1) * nw;
        if (ahead < n_tiles) {
            const uint32_t *p = in + ahead * 96 + lane * 3;
            n0 = ldg_u32(p); n1 = ldg_u32(p + 1); n2 = ldg_u32(p + 2);
        }
        // bytes: w0 = r0 g0 b0 r1 | w1 = g1 b1 r2 g2 | w2 = b2 r3 g3 b3
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
        if (HUE) {
            float h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = hue_from_diff(make_diff(px[j]));
            if (sizeof(TOut) == 4) {
                float4 *o = reinterpret_cast<float4 *>(out) + tile * 32 + lane;
                *o = make_float4(h[0], h[1], h[2], h[3]);
            } else {
                // one 256-bit store per lane (sm_100: STG.E.256): 32 lanes x 32 B contiguous
                stg_f64x4(reinterpret_cast<double *>(out) + tile * 128 + lane * 4,
                          (double)h[0], (double)h[1], (double)h[2], (double)h[3]);
            }
        } else {
            float r[12];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                husl_from_diff<true>(make_diff(px[j]), r[3 * j], r[3 * j + 1], r[3 * j + 2]);
            unsigned char *wr = stage + lane * Pack<TOut>::LANE_STRIDE;
            if (sizeof(TOut) == 4) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    *reinterpret_cast<float4 *>(wr + 16 * k) =
                        make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    *reinterpret_cast<double2 *>(wr + 16 * k) =
                        make_double2((double)r[2 * k], (double)r[2 * k + 1]);
            }
            __syncwarp();
            float4 *o = reinterpret_cast<float4 *>(out) + tile * (32 * CH) + lane;
#pragma unroll
            for (int k = 0; k < CH; ++k)
                o[32 * k] = *reinterpret_cast<const float4 *>(stage + rd_off[k]);
            __syncwarp();
        }
        if (PF2) { w0 = x0; w1 = x1; w2 = x2; x0 = n0; x1 = n1; x2 = n2; }
        else { w0 = n0; w1 = n1; w2 = n2; }
    }
}

// ---------------------------------------------------------------------------------
// (b) rgb_from_husl: float32 / float64 interleaved HUSL in, uint8 RGB out
// ---------------------------------------------------------------------------------
template <typename TIn, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_inv_u8(const TIn *__restrict__ in, uint32_t *__restrict__ out, long long n_tiles) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *stage = smem + warp * Pack<TIn>::STAGE;
    const long long nw = (long long)gridDim.x * (BLOCK / 32);
    constexpr int CH = Pack<TIn>::CHUNKS;
    uint32_t st_off[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int p = 32 * k + lane;
        st_off[k] = (p / CH) * Pack<TIn>::LANE_STRIDE + (p % CH) * 16;
    }
    for (long long tile = (long long)blockIdx.x * (BLOCK / 32) + warp; tile < n_tiles; tile += nw) {
        // coalesced 16-byte loads -> staging (record layout) -> lane-contiguous reads
        const float4 *src = reinterpret_cast<const float4 *>(in) + tile * (32 * CH) + lane;
        float4 v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = __ldcs(src + 32 * k);
#pragma unroll
        for (int k = 0; k < CH; ++k) *reinterpret_cast<float4 *>(stage + st_off[k]) = v[k];
        __syncwarp();
        float hsl[12];
        const unsigned char *rd = stage + lane * Pack<TIn>::LANE_STRIDE;
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
        __syncwarp();
        uint32_t q[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float R, G, B;
            rgb255_from_husl<false, wide_in<TIn>()>(hsl[3 * j], hsl[3 * j + 1], hsl[3 * j + 2], R, G, B);
            q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
        }
        // pack 12 low bytes into 3 words (PRMT), lane-contiguous 12 B -> staging -> coalesced
        uint32_t o0 = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        uint32_t o1 = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
        uint32_t o2 = __byte_perm(__byte_perm(q[8], q[9], 0x0040), __byte_perm(q[10], q[11], 0x0040), 0x5410);
        uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
        sw[lane * 3] = o0; sw[lane * 3 + 1] = o1; sw[lane * 3 + 2] = o2;
        __syncwarp();
        uint32_t *dst = out + tile * 96 + lane;
        dst[0] = sw[lane]; dst[32] = sw[lane + 32]; dst[64] = sw[lane + 64];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// generic kernels: one pixel per thread, any alignment, every dtype combination
// ---------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ LinPx load_px(const TIn *in, long long i, int channels);

template <>
__device__ __forceinline__ LinPx load_px<uint8_t>(const uint8_t *in, long long i, int channels) {
    LinPx p;
    float2 r, g, b;
    if (channels == 3) {
        r = g_lin_tab[in[3 * i]]; g = g_lin_tab[in[3 * i + 1]]; b = g_lin_tab[in[3 * i + 2]];
    } else if (channels == 4) {
        // RGBA: colour premultiplied by alpha (composited over black, the arithmetic of
        // transform.reshape_rgba_input, transform.py:221-230), kept in float: c/255 * a/255
        const double a = (double)in[4 * i + 3] * (1.0 / 65025.0);
        lin_f64((double)in[4 * i] * a, p.hr, p.lr);
        lin_f64((double)in[4 * i + 1] * a, p.hg, p.lg);
        lin_f64((double)in[4 * i + 2] * a, p.hb, p.lb);
        return p;
    } else {
        r = g = b = g_lin_tab[in[i]];
    }
    p.hr = r.x; p.lr = r.y; p.hg = g.x; p.lg = g.y; p.hb = b.x; p.lb = b.y;
    return p;
}

template <typename TIn>
__device__ __forceinline__ LinPx load_px(const TIn *in, long long i, int channels) {
    LinPx p;
    if (channels == 3) {
        lin_f64(in[3 * i], p.hr, p.lr);
        lin_f64(in[3 * i + 1], p.hg, p.lg);
        lin_f64(in[3 * i + 2], p.hb, p.lb);
    } else if (channels == 4) {      // float RGBA in [0,1]: rgb * alpha (transform.py:221-230)
        const double a = (double)in[4 * i + 3];
        lin_f64((double)in[4 * i] * a, p.hr, p.lr);
        lin_f64((double)in[4 * i + 1] * a, p.hg, p.lg);
        lin_f64((double)in[4 * i + 2] * a, p.hb, p.lb);
    } else {
        lin_f64(in[i], p.hg, p.lg);
        p.hr = p.hb = p.hg; p.lr = p.lb = p.lg;
    }
    return p;
}

// the pixel's four quantities (husl_math.cuh: Diff); three float32 channels take the float64-free route
template <typename TIn>
__device__ __forceinline__ Diff load_diff(const TIn *in, long long i, int channels) {
    return make_diff(load_px<TIn>(in, i, channels));
}
template <>
__device__ __forceinline__ Diff load_diff<float>(const float *in, long long i, int channels) {
    if (channels == 3) return diff_from_f32(in[3 * i], in[3 * i + 1], in[3 * i + 2]);
    return make_diff(load_px<float>(in, i, channels));
}

template <typename TIn, typename TOut, bool HUE>
__global__ void __launch_bounds__(256)
k_fwd_any(const TIn *__restrict__ in, TOut *__restrict__ out, long long n, int channels) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Diff d = load_diff<TIn>(in, i, channels);
        if (HUE) {
            out[i] = (TOut)hue_from_diff(d);
        } else {
            float H, S, L;
            if (sizeof(TIn) == 1 && channels != 4) husl_from_diff<true>(d, H, S, L);
            else husl_from_diff<false>(d, H, S, L);
            out[3 * i] = (TOut)H; out[3 * i + 1] = (TOut)S; out[3 * i + 2] = (TOut)L;
        }
    }
}

// Grayscale -> hue (reshape_image_input's grayscale route into to_hue, transform.py:168-173, 210-215).
// R = G = B makes both hue forms equal their grey offsets (husl_math.cuh: hue_forms), so every non-black grey
// has ONE hue -- hue_deg(GREY_V, GREY_U), the hue of white -- and black has 0: no linearisation, no arctangent
// per pixel, the kernel is a pure 16 B/px (float64 -> float64) stream.  Bit-identical to k_fwd_any<.., HUE>
// on a grey image, including what that kernel answers for negative, NaN and overflowing inputs.
template <typename TIn>
__device__ __forceinline__ float gray_hue_px(TIn v, float grey_hue) {
    if (sizeof(TIn) == 1) return v != 0 ? grey_hue : 0.0f;
    if (sizeof(TIn) == 4) {
        // float32 channels: the general kernels form the pixel with diff_from_f32
        const float c = (float)v;
        if (c <= 1e15f) {
            // lit <=> the linear value of G there is a positive float
            const bool lit = (c > SRGB_T) ? true : (c * (1.0f / 12.92f) > 0.0f);
            return lit ? grey_hue : 0.0f;
        }
        return hue_from_diff(diff_from_f32(c, c, c));      // NaN, +inf, or a power that overflows
    }
    const double c = (double)v;
    if (c <= 1e15) {
        // lit <=> the linear value the general kernel forms is a positive float (lin_f64)
        const bool lit = (c > 0.04045) ? true : ((float)(c * (1.0 / 12.92)) > 0.0f);
        return lit ? grey_hue : 0.0f;
    }
    // NaN, or so large that the float linear value overflows: whatever the general code answers
    LinPx p;
    lin_f64(c, p.hg, p.lg);
    p.hr = p.hb = p.hg; p.lr = p.lb = p.lg;
    return hue_from_diff(make_diff(p));
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
k_gray_hue(const TIn *__restrict__ in, TOut *__restrict__ out, long long n) {
    const float grey_hue = hue_deg(GREY_V, GREY_U);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (TOut)gray_hue_px<TIn>(in[i], grey_hue);
}

template <typename TIn, typename TOut, int ED = -1>
__global__ void __launch_bounds__(256)
k_inv_any(const TIn *__restrict__ in, TOut *__restrict__ out, long long n, const EditArg<ED> ea = EditArg<ED>()) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float R, G, B;
        float h = (float)in[3 * i], s = (float)in[3 * i + 1], l = (float)in[3 * i + 2];
        edit_on_load<ED>(ea, h, s, l);
        rgb255_from_husl<sizeof(TOut) != 1, wide_in<TIn>() && ED < 0>(h, s, l, R, G, B);
        if (sizeof(TOut) == 1) {
            out[3 * i] = (TOut)(round_u8_bits(R) & 0xFF);
            out[3 * i + 1] = (TOut)(round_u8_bits(G) & 0xFF);
            out[3 * i + 2] = (TOut)(round_u8_bits(B) & 0xFF);
        } else {   // float RGB in [0,1], what the reference's _husl_to_rgb returns
            out[3 * i] = (TOut)(R * (1.0f / 255.0f));
            out[3 * i + 1] = (TOut)(G * (1.0f / 255.0f));
            out[3 * i + 2] = (TOut)(B * (1.0f / 255.0f));
        }
    }
}

}  // namespace husl
