// husl_strided.cuh -- the hot path on STRIDED device images (SURVEY.md section 8f
// row 2, "strided / planar inputs"): cropped views `img[y0:y1, x0:x1]`, channel
// slices `rgba[..., :3]`, channel-reversed views `bgr[..., ::-1]`, row-padded
// (pitched) frames from a decoder, `(3, H, W)` planes seen as `(H, W, 3)` by a
// permute.  The reference makes such inputs contiguous with a host copy before
// any backend sees them (np.ascontiguousarray via the memoryview cast of
// _simd_opt.pyx:36 / _cython_opt.pyx:22); here the kernel walks the view itself.
//
// One pixel per thread over a (frames, rows, cols) view described by byte
// strides; the x dimension of the grid walks the columns of a row so a warp still
// touches neighbouring pixels, y walks rows, z walks frames.  Same per-pixel math as every other kernel
// (husl_math.cuh); this is the generality path, the contiguous layouts keep
// their dedicated streaming kernels.
#pragma once
#include "husl_fused.cuh"

namespace husl {

struct View2 {            // device-side copy of nphusl_view (byte strides, may be negative)
    char *base;
    long long frames, rows, cols, fs, rs, cs, es;
};

template <typename T>
__device__ __forceinline__ LinPx load_px_strided(const char *p, long long es);

template <>
__device__ __forceinline__ LinPx load_px_strided<uint8_t>(const char *p, long long es) {
    const float2 r = g_lin_tab[*reinterpret_cast<const uint8_t *>(p)];
    const float2 g = g_lin_tab[*reinterpret_cast<const uint8_t *>(p + es)];
    const float2 b = g_lin_tab[*reinterpret_cast<const uint8_t *>(p + 2 * es)];
    LinPx px;
    px.hr = r.x; px.lr = r.y; px.hg = g.x; px.lg = g.y; px.hb = b.x; px.lb = b.y;
    return px;
}

template <typename T>
__device__ __forceinline__ LinPx load_px_strided(const char *p, long long es) {
    LinPx px;
    lin_f64(*reinterpret_cast<const T *>(p), px.hr, px.lr);
    lin_f64(*reinterpret_cast<const T *>(p + es), px.hg, px.lg);
    lin_f64(*reinterpret_cast<const T *>(p + 2 * es), px.hb, px.lb);
    return px;
}

template <typename T>
__device__ __forceinline__ Diff load_diff_strided(const char *p, long long es) {
    return make_diff(load_px_strided<T>(p, es));
}
template <>
__device__ __forceinline__ Diff load_diff_strided<float>(const char *p, long long es) {
    return diff_from_f32(*reinterpret_cast<const float *>(p), *reinterpret_cast<const float *>(p + es),
                         *reinterpret_cast<const float *>(p + 2 * es));
}

// RGB view -> HUSL view (HUE: one value per pixel, out.es unused)
template <typename TIn, typename TOut, bool HUE>
__global__ void __launch_bounds__(256)
k_fwd_strided(const View2 in, const View2 out) {
    const long long xstep = (long long)gridDim.x * blockDim.x;
    for (long long f = blockIdx.z; f < in.frames; f += gridDim.z)
    for (long long y = blockIdx.y; y < in.rows; y += gridDim.y) {
        const char *irow = in.base + f * in.fs + y * in.rs;
        char *orow = out.base + f * out.fs + y * out.rs;
        for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < in.cols; x += xstep) {
            const Diff d = load_diff_strided<TIn>(irow + x * in.cs, in.es);
            char *o = orow + x * out.cs;
            if (HUE) {
                *reinterpret_cast<TOut *>(o) = (TOut)hue_from_diff(d);
            } else {
                float H, S, L;
                if (sizeof(TIn) == 1) husl_from_diff<true>(d, H, S, L);
                else husl_from_diff<false>(d, H, S, L);
                *reinterpret_cast<TOut *>(o) = (TOut)H;
                *reinterpret_cast<TOut *>(o + out.es) = (TOut)S;
                *reinterpret_cast<TOut *>(o + 2 * out.es) = (TOut)L;
            }
        }
    }
}

// HUSL view -> RGB view (uint8 with the to_rgb_int tail, or float RGB in [0,1])
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
k_inv_strided(const View2 in, const View2 out) {
    const long long xstep = (long long)gridDim.x * blockDim.x;
    for (long long f = blockIdx.z; f < in.frames; f += gridDim.z)
    for (long long y = blockIdx.y; y < in.rows; y += gridDim.y) {
        const char *irow = in.base + f * in.fs + y * in.rs;
        char *orow = out.base + f * out.fs + y * out.rs;
        for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < in.cols; x += xstep) {
            const char *p = irow + x * in.cs;
            float R, G, B;
            rgb255_from_husl<sizeof(TOut) != 1, wide_in<TIn>()>((float)*reinterpret_cast<const TIn *>(p), (float)*reinterpret_cast<const TIn *>(p + in.es),
                                                (float)*reinterpret_cast<const TIn *>(p + 2 * in.es), R, G, B);
            char *o = orow + x * out.cs;
            if (sizeof(TOut) == 1) {
                *reinterpret_cast<uint8_t *>(o) = (uint8_t)(round_u8_bits(R) & 0xFF);
                *reinterpret_cast<uint8_t *>(o + out.es) = (uint8_t)(round_u8_bits(G) & 0xFF);
                *reinterpret_cast<uint8_t *>(o + 2 * out.es) = (uint8_t)(round_u8_bits(B) & 0xFF);
            } else {
                *reinterpret_cast<TOut *>(o) = (TOut)(R * (1.0f / 255.0f));
                *reinterpret_cast<TOut *>(o + out.es) = (TOut)(G * (1.0f / 255.0f));
                *reinterpret_cast<TOut *>(o + 2 * out.es) = (TOut)(B * (1.0f / 255.0f));
            }
        }
    }
}

// fused to_husl -> select -> edit -> to_rgb on a uint8 view (in place when both views alias)
__global__ void __launch_bounds__(256)
k_edit_strided(const View2 in, const View2 out, const nphusl_edit e) {
    const long long xstep = (long long)gridDim.x * blockDim.x;
    for (long long f = blockIdx.z; f < in.frames; f += gridDim.z)
    for (long long y = blockIdx.y; y < in.rows; y += gridDim.y) {
        const char *irow = in.base + f * in.fs + y * in.rs;
        char *orow = out.base + f * out.fs + y * out.rs;
        for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < in.cols; x += xstep) {
            float h, s, l, R, G, B;
            husl_from_diff<true>(make_diff(load_px_strided<uint8_t>(irow + x * in.cs, in.es)), h, s, l);
            apply_edit(e, h, s, l);
            rgb255_from_husl(h, s, l, R, G, B);
            char *o = orow + x * out.cs;
            *reinterpret_cast<uint8_t *>(o) = (uint8_t)(round_u8_bits(R) & 0xFF);
            *reinterpret_cast<uint8_t *>(o + out.es) = (uint8_t)(round_u8_bits(G) & 0xFF);
            *reinterpret_cast<uint8_t *>(o + 2 * out.es) = (uint8_t)(round_u8_bits(B) & 0xFF);
        }
    }
}

}  // namespace husl
