// husl_nv12.cuh -- video-decoder surfaces as the load stage of the HUSL hot path
// (SURVEY.md section 8f "next" row 3: decode adjacency).
//
// NVDEC hands out NV12: a pitched luma plane Y[h][pitch] and, below it, one
// pitched plane of interleaved half-resolution chroma (U, V byte pairs, one pair
// per 2 x 2 luma block).  The reference's callers get RGB arrays from imageio /
// moviepy (examples.py:117-134) after a colour conversion on the CPU; here that
// conversion is the first ~14 instructions of the kernel that goes on to HUSL:
//
//   k_nv12<MODE, TOut>    NV12 -> { interleaved HUSL | hue | uint8 RGB | fused edit -> uint8 RGB }
//   k_nv12_any<MODE,TOut> one pixel per thread: widths that are not a multiple of 4,
//                         unaligned planes
//
// Colour conversion (integer, so a numpy restatement is bit-exact -- tests use
// oracle.nv12_to_rgb): with 16.16 fixed-point coefficients of the BT.601 / BT.709
// matrix (nphusl_cuda_api.cu: nv12_coef),
//     c = Y - y_off, d = U - 128, e = V - 128
//     R = clamp8((ky*c          + rv*e + 32768) >> 16)
//     G = clamp8((ky*c - gu*d   - gv*e + 32768) >> 16)
//     B = clamp8((ky*c + bu*d          + 32768) >> 16)
// chroma is replicated over its 2 x 2 block (nearest sample, what NPP's
// nppiNV12ToRGB does).  The 8-bit R, G, B then index the same hi+lo
// linearisation table as every other uint8 kernel, so the HUSL values equal
// to_husl(nv12_to_rgb(frame)) bit for bit.
//
// A warp owns 128 consecutive pixels of ONE row (4 per lane: one 32-bit word of
// luma, one 32-bit word of chroma = two U,V pairs); the output tile is the
// contiguous run out[row][x0 : x0+128], written exactly like k_fwd_u8 / k_edit
// write theirs.  Rows whose width is not a multiple of 128 end in a partial tile
// (whole groups of 4 pixels, predicated chunk by chunk).
#pragma once
#include "husl_fused.cuh"

#ifndef HUSL_NV12_EDIT_UNROLL
#define HUSL_NV12_EDIT_UNROLL 2   // row loop of k_nv12<NV12_EDIT>: 0.3386 -> 0.3355 ms
#endif
#ifndef HUSL_NV12_UNROLL
#define HUSL_NV12_UNROLL 2     // row loop of k_nv12<NV12_HUSL> unrolled by two: f32 0.1940 -> 0.1866 ms (interleaved A/B)
#endif

namespace husl {

template <> struct Pack<uint8_t> {          // uint8 RGB tile: 384 B, no padding (k_edit's staging)
    static constexpr int CHUNKS = 1;
    static constexpr int LANE_STRIDE = 12;
    static constexpr int STAGE = 384;
};

// ky, rv, gu, gv, bu: 16.16 coefficients; cr, cg, cb: everything constant folded into one
// term per channel (rounding 32768, -ky*y_off, -+128*chroma coefficients), so that
//     R = clamp8((ky*Y + rv*V + cr) >> 16), G = clamp8((ky*Y - gu*U - gv*V + cg) >> 16), ...
// is the documented formula with no subtraction left per pixel (all integers: bit-identical).
struct Nv12Coef {
    int ky, rv, gu, gv, bu, cr, cg, cb;
};

struct Nv12Src {
    const uint8_t *y, *uv;
    long long y_pitch, uv_pitch;
    int width, height;
    Nv12Coef k;
};

enum { NV12_HUSL = 0, NV12_HUE = 1, NV12_RGB = 2, NV12_EDIT = 3 };

// Channel values stay in 16.16 form after the clamp -- max(min(v, 0xFFFFFF), 0) is one VIMNMX.RELU
// on sm_100 -- so that byte 2 IS the 8-bit channel (floor(v / 65536) in 0..255): the table lookups
// and the uint8 packing pick that byte with the PRMT they need anyway, and no shift is issued.
__device__ __forceinline__ uint32_t clamp16_16(int v) { return (uint32_t)__vimin_s32_relu(v, 0xFFFFFF); }

// one pixel: luma byte + the three chroma terms of its 2 x 2 block -> R, G, B as 16.16 values (byte 2 = 0..255)
__device__ __forceinline__ void yuv_px(const Nv12Coef &k, int yb, int cr, int cg, int cb,
                                       uint32_t &R, uint32_t &G, uint32_t &B) {
    const int yv = k.ky * yb;
    R = clamp16_16(yv + cr);
    G = clamp16_16(yv + cg);
    B = clamp16_16(yv + cb);
}

__device__ __forceinline__ void chroma_terms(const Nv12Coef &k, int u, int v, int &cr, int &cg, int &cb) {
    cr = k.rv * v + k.cr;
    cg = k.cg - k.gu * u - k.gv * v;
    cb = k.bu * u + k.cb;
}

// 4 pixels of a lane: y4 = Y0 Y1 Y2 Y3, uv4 = U0 V0 U1 V1 (little endian bytes) -> 12 channel values
__device__ __forceinline__ void yuv4(const Nv12Coef &k, uint32_t y4, uint32_t uv4, uint32_t (&c)[12]) {
    int cr0, cg0, cb0, cr1, cg1, cb1;
    // byte K zero-extended: one PRMT each
    chroma_terms(k, __byte_perm(uv4, 0, 0x4440), __byte_perm(uv4, 0, 0x4441), cr0, cg0, cb0);
    chroma_terms(k, __byte_perm(uv4, 0, 0x4442), __byte_perm(uv4, 0, 0x4443), cr1, cg1, cb1);
    yuv_px(k, __byte_perm(y4, 0, 0x4440), cr0, cg0, cb0, c[0], c[1], c[2]);
    yuv_px(k, __byte_perm(y4, 0, 0x4441), cr0, cg0, cb0, c[3], c[4], c[5]);
    yuv_px(k, __byte_perm(y4, 0, 0x4442), cr1, cg1, cb1, c[6], c[7], c[8]);
    yuv_px(k, __byte_perm(y4, 0, 0x4443), cr1, cg1, cb1, c[9], c[10], c[11]);
}

__device__ __forceinline__ LinPx lin_px(const unsigned char *tab, uint32_t r, uint32_t g, uint32_t b, uint32_t lane8) {
    LinPx p;
    float2 v;
    v = lut_lin<2>(tab, r, lane8); p.hr = v.x; p.lr = v.y;      // byte 2 of a 16.16 channel value
    v = lut_lin<2>(tab, g, lane8); p.hg = v.x; p.lg = v.y;
    v = lut_lin<2>(tab, b, lane8); p.hb = v.x; p.lb = v.y;
    return p;
}

// shared memory: [ 64 KB table (not for NV12_RGB) | warps x staging ]
template <int MODE, typename TOut>
struct Nv12Smem {
    static constexpr int TAB = MODE == NV12_RGB ? 0 : TAB_BYTES;
    static constexpr int STAGE = MODE == NV12_HUSL ? Pack<TOut>::STAGE : (MODE == NV12_HUE ? 0 : 384);
    static constexpr int bytes(int block) { return TAB + (block / 32) * STAGE; }
};

// EH / ES / RSL: the edit-shape promises of husl_fused.cuh (NV12_EDIT only)
template <int MODE, typename TOut, int BLOCK, int MINB, bool EH = true, bool ES = true, bool RSL = true>
__global__ void __launch_bounds__(BLOCK, MINB)
k_nv12(const Nv12Src src, TOut *__restrict__ out, const nphusl_edit e) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int TAB = Nv12Smem<MODE, TOut>::TAB;
    if (TAB) fill_table(smem);
    const unsigned char *tab = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    unsigned char *stage = smem + TAB + warp * Nv12Smem<MODE, TOut>::STAGE;
    const int W = src.width, H = src.height;
    // Tiles are (row, 128-pixel segment) pairs.  A warp keeps ONE segment column and walks down the
    // rows with an even row step, so every address advances by a constant (luma pointer by
    // step*y_pitch, chroma pointer by step/2*uv_pitch, output by step*W pixels): no division and no
    // 64-bit multiply per tile.  slot -> (segment, first row); slots beyond the grid loop again.
    const int segs = (W + TILE_PX - 1) / TILE_PX;
    const long long n_warps = (long long)gridDim.x * (BLOCK / 32);
    const int step = 2 * (int)max(1LL, min(n_warps / (2LL * segs), (long long)((H + 1) / 2)));
    const long long n_slots = (long long)segs * step;

    constexpr int CH = MODE == NV12_HUSL ? Pack<TOut>::CHUNKS : 1;
    uint32_t rd_off[CH];
    if constexpr (MODE == NV12_HUSL) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int p = 32 * k + lane;
            rd_off[k] = (p / CH) * Pack<TOut>::LANE_STRIDE + (p % CH) * 16;
        }
    }
    constexpr int OUT_PX_BYTES = (MODE == NV12_HUE ? 1 : 3) * (int)sizeof(TOut);

    for (long long slot = (long long)blockIdx.x * (BLOCK / 32) + warp; slot < n_slots; slot += n_warps) {
    const int seg = (int)(slot % segs);
    int row = (int)(slot / segs);
    const int x0 = seg * TILE_PX;
    const int valid = min(TILE_PX, W - x0);                       // pixels of this column's tiles, a multiple of 4
    const bool lane_on = lane * 4 < valid;
    const uint8_t *yp = src.y + (long long)row * src.y_pitch + x0 + lane * 4;
    const uint8_t *uvp = src.uv + (long long)(row >> 1) * src.uv_pitch + x0 + lane * 4;
    unsigned char *op = reinterpret_cast<unsigned char *>(out) + ((long long)row * W + x0) * OUT_PX_BYTES;
    const long long y_step = (long long)step * src.y_pitch, uv_step = (long long)(step / 2) * src.uv_pitch;
    const long long o_step = (long long)step * W * OUT_PX_BYTES;

    uint32_t y4 = 0, uv4 = 0x80808080u;
    if (row < H && lane_on) {
        y4 = ldg_u32(reinterpret_cast<const uint32_t *>(yp));
        uv4 = ldg_u32(reinterpret_cast<const uint32_t *>(uvp));
    }
    // float32 HUSL output only: the float64 variant (HBM-bound, 512 x 1) loses 11 % unrolled (0.2905 -> 0.3252 ms)
    constexpr int UNROLL = (MODE == NV12_HUSL && sizeof(TOut) == 4) ? HUSL_NV12_UNROLL : (MODE == NV12_EDIT ? HUSL_NV12_EDIT_UNROLL : 1);
#pragma unroll UNROLL
    for (; row < H; row += step, op += o_step) {
        yp += y_step; uvp += uv_step;
        uint32_t ny = 0, nuv = 0x80808080u;
        if (row + step < H && lane_on) {
            ny = ldg_u32(reinterpret_cast<const uint32_t *>(yp));
            nuv = ldg_u32(reinterpret_cast<const uint32_t *>(uvp));
        }

        uint32_t c[12];
        yuv4(src.k, y4, uv4, c);
        if constexpr (MODE == NV12_RGB || MODE == NV12_EDIT) {
            uint32_t q[12];
            if constexpr (MODE == NV12_EDIT) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float h, s, l, R, G, B;
                    husl_from_diff<true>(make_diff(lin_px(tab, c[3 * j], c[3 * j + 1], c[3 * j + 2], lane8)), h, s, l);
                    apply_edit<EH, ES, RSL>(e, h, s, l);
                    rgb255_from_husl(h, s, l, R, G, B);
                    q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) q[j] = c[j];
            }
            uint32_t o0, o1, o2;
            if constexpr (MODE == NV12_EDIT) {
                pack12(q, o0, o1, o2);
            } else {                                              // byte 2 of the 16.16 values
                o0 = __byte_perm(__byte_perm(q[0], q[1], 0x0062), __byte_perm(q[2], q[3], 0x0062), 0x5410);
                o1 = __byte_perm(__byte_perm(q[4], q[5], 0x0062), __byte_perm(q[6], q[7], 0x0062), 0x5410);
                o2 = __byte_perm(__byte_perm(q[8], q[9], 0x0062), __byte_perm(q[10], q[11], 0x0062), 0x5410);
            }
            uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
            sw[lane * 3] = o0; sw[lane * 3 + 1] = o1; sw[lane * 3 + 2] = o2;
            __syncwarp();
            uint32_t *dst = reinterpret_cast<uint32_t *>(op) + lane;
            const int words = valid * 3 / 4;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (32 * k + lane < words) dst[32 * k] = sw[32 * k + lane];
            __syncwarp();
        } else if constexpr (MODE == NV12_HUE) {
            float h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                h[j] = hue_from_diff(make_diff(lin_px(tab, c[3 * j], c[3 * j + 1], c[3 * j + 2], lane8)));
            if (lane_on) {
                if (sizeof(TOut) == 4)
                    *reinterpret_cast<float4 *>(reinterpret_cast<float *>(op) + lane * 4) = make_float4(h[0], h[1], h[2], h[3]);
                else
                    stg_f64x4(reinterpret_cast<double *>(op) + lane * 4, (double)h[0], (double)h[1], (double)h[2], (double)h[3]);
            }
        } else {
            float r[12];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                husl_from_diff<true>(make_diff(lin_px(tab, c[3 * j], c[3 * j + 1], c[3 * j + 2], lane8)),
                                     r[3 * j], r[3 * j + 1], r[3 * j + 2]);
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
            float4 *o = reinterpret_cast<float4 *>(op) + lane;
            const int chunks = valid / 4 * CH;                    // 16-byte chunks of this tile
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (32 * k + lane < chunks) o[32 * k] = *reinterpret_cast<const float4 *>(stage + rd_off[k]);
            __syncwarp();
        }
        y4 = ny; uv4 = nuv;
    }
    }
}

// one pixel per thread: any width, any alignment
template <int MODE, typename TOut>
__global__ void __launch_bounds__(256)
k_nv12_any(const Nv12Src src, TOut *__restrict__ out, const nphusl_edit e) {
    const long long n = (long long)src.width * src.height;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long row = i / src.width;
        const int x = (int)(i - row * src.width);
        const uint8_t *uvp = src.uv + (row >> 1) * src.uv_pitch + (x & ~1);
        int cr, cg, cb;
        chroma_terms(src.k, uvp[0], uvp[1], cr, cg, cb);
        uint32_t R, G, B;
        yuv_px(src.k, src.y[row * src.y_pitch + x], cr, cg, cb, R, G, B);
        R >>= 16; G >>= 16; B >>= 16;
        if constexpr (MODE == NV12_RGB) {
            uint8_t *o = reinterpret_cast<uint8_t *>(out) + 3 * i;
            o[0] = (uint8_t)R; o[1] = (uint8_t)G; o[2] = (uint8_t)B;
        } else {
        const float2 r = g_lin_tab[R], g = g_lin_tab[G], b = g_lin_tab[B];
        LinPx p;
        p.hr = r.x; p.lr = r.y; p.hg = g.x; p.lg = g.y; p.hb = b.x; p.lb = b.y;
        const Diff d = make_diff(p);
        if constexpr (MODE == NV12_HUE) {
            out[i] = (TOut)hue_from_diff(d);
        } else {
            float h, s, l;
            husl_from_diff<true>(d, h, s, l);
            if constexpr (MODE == NV12_EDIT) {
                float fr, fg, fb;
                apply_edit(e, h, s, l);
                rgb255_from_husl(h, s, l, fr, fg, fb);
                uint8_t *o = reinterpret_cast<uint8_t *>(out) + 3 * i;
                o[0] = (uint8_t)(round_u8_bits(fr) & 0xFF);
                o[1] = (uint8_t)(round_u8_bits(fg) & 0xFF);
                o[2] = (uint8_t)(round_u8_bits(fb) & 0xFF);
            } else {
                out[3 * i] = (TOut)h; out[3 * i + 1] = (TOut)s; out[3 * i + 2] = (TOut)l;
            }
        }
        }
    }
}

}  // namespace husl
