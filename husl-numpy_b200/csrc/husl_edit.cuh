// husl_edit.cuh -- the HUSL-space select + affine edit shared by the fused uint8 kernels
// (husl_fused.cuh, husl_nv12*.cuh, husl_rows.cuh), by the edit-on-load variants of the inverse
// kernels (husl_stream.cuh: to_rgb(hsl, edit=...)) and by the in-place HUSL edit kernel below.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/nphusl_cuda.h"

namespace husl {

// Selection: a pixel is selected when H lies on the arc [h_lo, h_hi] (h_lo > h_hi
// means the arc through 360/0) AND S in [s_lo, s_hi] AND L in [l_lo, l_hi];
// `invert` selects the complement.  Selected pixels get H' = H*h_mul + h_add
// (wrapped to [0, 360)), S' = clamp(S*s_mul + s_add, 0, 100), L' likewise; a
// channel whose (mul, add) is (1, 0) is left exactly as it is, like the channels a
// caller's `hsl[..., 2][mask] *= 0.5` never touches.  Every pixel -- selected or
// not -- goes through the float32 HUSL round trip, exactly as
// to_rgb(edit(to_husl(rgb, dtype=float32))) does.
// EH / ES / RSL are compile-time promises of the dispatcher ("the hue / saturation
// channel may be edited", "a saturation or lightness range may be bounded"): the
// streaming kernels are instantiated per combination so that an edit which only
// scales L inside a hue arc (README.md:77-81) carries no hue-wrap, saturation or
// range arithmetic at all.  <true, true, true> is valid for every descriptor.
template <bool EH = true, bool ES = true, bool RSL = true>
__device__ __forceinline__ void apply_edit(const nphusl_edit &e, float &h, float &s, float &l) {
    bool sel = (e.h_lo <= e.h_hi) ? (h >= e.h_lo && h <= e.h_hi) : (h >= e.h_lo || h <= e.h_hi);
    if (RSL) sel = sel && s >= e.s_lo && s <= e.s_hi && l >= e.l_lo && l <= e.l_hi;
    sel = e.invert ? !sel : sel;
    if (EH) {
        float hn = __fadd_rn(__fmul_rn(h, e.h_mul), e.h_add);
        hn = hn - 360.0f * floorf(hn * (1.0f / 360.0f));
        hn = (hn >= 360.0f || hn < 0.0f) ? 0.0f : hn;
        h = (sel && (e.h_mul != 1.0f || e.h_add != 0.0f)) ? hn : h;
    }
    if (ES) {
        const float sn = fminf(fmaxf(__fadd_rn(__fmul_rn(s, e.s_mul), e.s_add), 0.0f), 100.0f);
        s = (sel && (e.s_mul != 1.0f || e.s_add != 0.0f)) ? sn : s;
    }
    const float ln = fminf(fmaxf(__fadd_rn(__fmul_rn(l, e.l_mul), e.l_add), 0.0f), 100.0f);
    l = (sel && (e.l_mul != 1.0f || e.l_add != 0.0f)) ? ln : l;
}

// what the dispatcher may promise for a descriptor (host side)
struct EditShape {
    bool eh, es, rsl;
    __host__ explicit EditShape(const nphusl_edit &e)
        : eh(e.h_mul != 1.0f || e.h_add != 0.0f), es(e.s_mul != 1.0f || e.s_add != 0.0f),
          rsl(!(e.s_lo == -INFINITY && e.s_hi == INFINITY && e.l_lo == -INFINITY && e.l_hi == INFINITY)) {}
    __host__ int code() const { return (eh ? 1 : 0) | (es ? 2 : 0) | (rsl ? 4 : 0); }
};

// Edit stage of kernels that can run with or without one.  ED = -1: no edit (the argument is an empty
// struct and the stage compiles to nothing); ED = 0..7: EditShape bits (1 = hue may be edited, 2 = saturation
// may be edited, 4 = a saturation / lightness range may be bounded), so an edit that only scales L inside a
// hue arc (README.md:77-81) adds a handful of instructions; 7 is valid for every descriptor.
template <int ED> struct EditArg { nphusl_edit e; };
template <> struct EditArg<-1> {};
template <int ED>
__device__ __forceinline__ void edit_on_load(const EditArg<ED> &a, float &h, float &s, float &l) {
    apply_edit<(ED & 1) != 0, (ED & 2) != 0, (ED & 4) != 0>(a.e, h, s, l);
}
template <>
__device__ __forceinline__ void edit_on_load<-1>(const EditArg<-1> &, float &, float &, float &) {}

// ---------------------------------------------------------------------------------
// banded map (nphusl_bands, include/nphusl_cuda.h): the callers' hue_watermelon / melonize
// (examples.py:56-103) -- bands of one channel painted alternately with two hues, optionally a
// saturated stripe around every band centre
// ---------------------------------------------------------------------------------
struct BandsDev {
    nphusl_bands b;
    float last;          // index of the last band (host side, bands_dev())
    float rwidth;        // 1 / width: only proposes a band, the edges decide
};

// SAT: the descriptor has s_half > 0.  x * (1 / width) proposes a band; the edges RN(k * width) themselves
// decide, one band down or up if the product was rounded across an edge.  A value on an edge, outside
// [0, total) or NaN fails the open-interval test of every band.  The stripe test looks at the centres of
// the band and of its two neighbours: the centre nearest to x is always one of those three.
template <bool SAT>
__device__ __forceinline__ void apply_bands(const BandsDev &a, float &h, float &s, float l) {
    const float x = a.b.src == NPHUSL_SRC_H ? h : (a.b.src == NPHUSL_SRC_S ? s : l);
    const float w = a.b.width, total = a.b.total;
    const float k0 = fminf(fmaxf(floorf(__fmul_rn(x, a.rwidth)), 0.0f), a.last);
    const float e0 = __fmul_rn(k0, w), e1 = __fmul_rn(k0 + 1.0f, w);        // edges of band k0
    const bool down = x <= e0 && k0 > 0.0f, up = x >= e1 && k0 < a.last;
    const float k = down ? k0 - 1.0f : (up ? k0 + 1.0f : k0);
    const float lo = down ? __fmul_rn(k, w) : (up ? e1 : e0);
    const float hi = fminf(down ? e0 : (up ? __fmul_rn(k + 1.0f, w) : e1), total);
    const bool in_band = x > lo && x < hi;
    const bool odd = ((int)k & 1) != 0;
    if (SAT) {
        const float hw = a.b.s_half;
        // centres of bands k - 1, k, k + 1 (edges lo_m = RN((k - 1) w), lo, hi, hi_p = min(RN((k + 2) w), total); hi = RN((k + 1) w) whenever band k + 1 exists)
        const float lo_m = __fmul_rn(k - 1.0f, w), hi_p = fminf(__fmul_rn(k + 2.0f, w), total);
        const float c0 = __fmul_rn(0.5f, __fadd_rn(lo, hi));
        const float cm = __fmul_rn(0.5f, __fadd_rn(lo_m, lo));
        const float cp = __fmul_rn(0.5f, __fadd_rn(hi, hi_p));
        bool stripe = x > __fsub_rn(c0, hw) && x < __fadd_rn(c0, hw);
        stripe = stripe || (k > 0.0f && x > __fsub_rn(cm, hw) && x < __fadd_rn(cm, hw));
        stripe = stripe || (k < a.last && x > __fsub_rn(cp, hw) && x < __fadd_rn(cp, hw));
        s = stripe ? a.b.s_value : s;
    }
    h = in_band ? (odd ? a.b.h_odd : a.b.h_even) : h;
}

__host__ inline BandsDev bands_dev(const nphusl_bands &b) {
    BandsDev a;
    a.b = b;
    a.rwidth = 1.0f / b.width;
    // the last band is the last k whose lower edge RN(k * width) lies below `total` (chunk(): it ends at `total`)
    volatile float lo;
    a.last = ceilf(b.total / b.width) - 1.0f;
    for (lo = (a.last + 1.0f) * b.width; lo < b.total; lo = (a.last + 1.0f) * b.width) a.last += 1.0f;
    for (lo = a.last * b.width; a.last > 0.0f && lo >= b.total; lo = a.last * b.width) a.last -= 1.0f;
    return a;
}

// ED = 8 / 9: a banded map (without / with the saturation stripe) in the edit stage of the inverse kernels
template <> struct EditArg<8> { BandsDev a; };
template <> struct EditArg<9> { BandsDev a; };
template <>
__device__ __forceinline__ void edit_on_load<8>(const EditArg<8> &e, float &h, float &s, float &l) { apply_bands<false>(e.a, h, s, l); }
template <>
__device__ __forceinline__ void edit_on_load<9>(const EditArg<9> &e, float &h, float &s, float &l) { apply_bands<true>(e.a, h, s, l); }

#ifndef HUSL_HOST_EMUL      // the kernels below are device code only (tests/emul builds apply_edit / apply_bands for the host)
// ---------------------------------------------------------------------------------
// in-place (or out-of-place) edit of an interleaved HUSL array: what a caller's
// `hsl[..., 2][mask] *= 0.5` (README.md:77-81) does, as one pass over the array
// ---------------------------------------------------------------------------------
// One thread per pixel, three scalar accesses each: neighbouring threads cover neighbouring
// triplets, so every 32-byte sector is fully used; the array is read once and written once
// (24 / 48 B per pixel).  The arithmetic is float32 on the float32 value of each channel,
// exactly what the fused kernels apply in registers; unselected pixels and untouched
// channels are written back bit for bit.
template <typename T>
__global__ void __launch_bounds__(256)
k_husl_edit(const T *in, T *out, long long n, const nphusl_edit e) {   // out may be in
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T h0 = in[3 * i], s0 = in[3 * i + 1], l0 = in[3 * i + 2];
        float h = (float)h0, s = (float)s0, l = (float)l0;
        const float hf = h, sf = s, lf = l;
        apply_edit<true, true, true>(e, h, s, l);
        out[3 * i] = (h == hf) ? h0 : (T)h;
        out[3 * i + 1] = (s == sf) ? s0 : (T)s;
        out[3 * i + 2] = (l == lf) ? l0 : (T)l;
    }
}

// The same pass with 32-byte accesses: a thread owns 96 consecutive bytes (4 float64 or 8 float32 pixels) and
// moves them as three 256-bit loads and three 256-bit stores (LDG / STG.E.ENL2.256, sm_100) -- every access is one
// whole DRAM sector.  The one-pixel-per-thread kernel above spends 8-byte requests: ncu shows it waiting on the
// load / store queue (lg_throttle 14.7, long_scoreboard 79 cycles per issue) at 4.95 TB/s against the 5.78 TB/s a
// 50/50 read-write stream reaches on this part.  Plain (not .nc) loads: `out` may be `in`, and a thread reads its own
// 96 bytes before it writes them.  Needs 32-byte aligned bases; the pixels after the last whole group go to k_husl_edit.
__device__ __forceinline__ void ld_256(const void *p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_256(void *p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ float f32_lo(double d) { return __int_as_float(__double2loint(d)); }
__device__ __forceinline__ float f32_hi(double d) { return __int_as_float(__double2hiint(d)); }

template <typename T>
__global__ void __launch_bounds__(256)
k_husl_edit_wide(const unsigned char *in, unsigned char *out, long long n_groups, const nphusl_edit e) {   // out may be in
    constexpr int PX = 96 / (3 * (int)sizeof(T));      // pixels per group: 4 (float64) or 8 (float32)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        double q[3][4];
        ld_256(in + g * 96, q[0]); ld_256(in + g * 96 + 32, q[1]); ld_256(in + g * 96 + 64, q[2]);
        T v[3 * PX];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            if (sizeof(T) == 8) v[i] = (T)q[i / 4][i % 4];
            else { v[2 * i] = (T)f32_lo(q[i / 4][i % 4]); v[2 * i + 1] = (T)f32_hi(q[i / 4][i % 4]); }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            float h = (float)v[3 * p], s = (float)v[3 * p + 1], l = (float)v[3 * p + 2];
            const float hf = h, sf = s, lf = l;
            apply_edit<true, true, true>(e, h, s, l);
            v[3 * p] = (h == hf) ? v[3 * p] : (T)h;
            v[3 * p + 1] = (s == sf) ? v[3 * p + 1] : (T)s;
            v[3 * p + 2] = (l == lf) ? v[3 * p + 2] : (T)l;
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            if (sizeof(T) == 8) q[i / 4][i % 4] = (double)v[i];
            else q[i / 4][i % 4] = __hiloint2double(__float_as_int((float)v[2 * i + 1]), __float_as_int((float)v[2 * i]));
        }
        st_256(out + g * 96, q[0]); st_256(out + g * 96 + 32, q[1]); st_256(out + g * 96 + 64, q[2]);
    }
}

#endif  // HUSL_HOST_EMUL

}  // namespace husl
