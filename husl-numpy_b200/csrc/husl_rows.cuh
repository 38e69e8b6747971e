// husl_rows.cuh -- strided device images whose ROWS are dense (cropped views img[:, y0:y1, x0:x1],
// pitched decoder frames, batches of ROIs): the tile kernels' data movement on a (frames, rows,
// cols) view.  husl_strided.cuh walks ANY byte strides one pixel per thread; when the pixels of
// a row are contiguous (col_stride = 3 elements, ch_stride = 1 element), row starts are aligned
// (uint8: 4 bytes, float HUSL: 16 bytes, hue: 4 elements) and cols % 4 == 0, a warp can own 128
// consecutive pixels of one row exactly as in the contiguous kernels:
//
//   k_rows_fwd<TOut, HUE>   uint8 RGB view -> HUSL (or hue) view
//   k_rows_inv<TIn>         HUSL view -> uint8 RGB view
//   k_rows_edit<EH,ES,RSL>  fused edit, uint8 view -> uint8 view (in place when they alias)
//
// A warp keeps one 128-pixel segment column and walks down the (frame, row) sequence with a
// constant row step (husl_nv12.cuh uses the same walk); rows whose width is not a multiple of
// 128 end in a partial tile, predicated per 16-byte chunk / per word.
#pragma once
#include "husl_strided.cuh"
#include "husl_nv12.cuh"

#ifndef HUSL_ROWS_UNROLL
#define HUSL_ROWS_UNROLL 1     // row loop of k_rows_fwd (HUSL output) unrolled by this factor (A/B knob)
#endif

namespace husl {

// The walk: slot -> (segment column, first row); rows are numbered through all frames.  Only the
// first (frame, row) of a slot costs a division; afterwards the pair advances by the constant
// (dq frames, dr rows) with one carry.
struct RowWalk {
    int segs;
    long long n_slots, n_warps, rows, frames, dq, dr;
    __device__ RowWalk(long long cols, long long frames_, long long rows_, int warps_per_block) {
        segs = (int)((cols + TILE_PX - 1) / TILE_PX);
        n_warps = (long long)gridDim.x * warps_per_block;
        rows = rows_; frames = frames_;
        const long long step = max(1LL, min(n_warps / segs, frames * rows));
        n_slots = (long long)segs * step;
        dq = step / rows; dr = step % rows;
    }
};
struct RowPos {
    long long f, y;
    // slot -> segment column `seg` and the first (frame, row); 32-bit divisions whenever the numbers allow,
    // none at all for a view that collapsed to one long row (a fully dense image)
    __device__ RowPos(const RowWalk &w, long long slot, int &seg) {
        if (w.frames * w.rows == 1) { seg = (int)slot; f = 0; y = 0; return; }
        long long r0;
        if (slot < 0x7fffffffLL) { const unsigned s32 = (unsigned)slot, g = (unsigned)w.segs; r0 = s32 / g; seg = (int)(s32 - (unsigned)r0 * g); }
        else { r0 = slot / w.segs; seg = (int)(slot - r0 * w.segs); }
        if (r0 < 0x7fffffffLL && w.rows < 0x7fffffffLL) { const unsigned r32 = (unsigned)r0, n = (unsigned)w.rows; f = r32 / n; y = r32 - (unsigned)f * n; }
        else { f = r0 / w.rows; y = r0 - f * w.rows; }
    }
    __device__ RowPos next(const RowWalk &w) const {
        RowPos n = *this;
        n.y += w.dr; n.f += w.dq;
        if (n.y >= w.rows) { n.y -= w.rows; ++n.f; }
        return n;
    }
    __device__ bool in(const RowWalk &w) const { return f < w.frames; }
    __device__ long long off(const View2 &v) const { return f * v.fs + y * v.rs; }
};

// PXB = bytes per stored pixel of the uint8 view: 3 (RGB), or 4 (RGBA / RGBX memory seen through
// rgba[..., :3]: one 128-bit load per lane, the fourth byte is ignored); BGR: the view's channel
// stride is -1 (BGRA memory presented as RGB), in.base then points at the R byte = byte 2 of a pixel.
template <typename TOut, bool HUE, int BLOCK, int MINB, int PXB = 3, bool BGR = false>
__global__ void __launch_bounds__(BLOCK, MINB)
k_rows_fwd(const View2 in, const View2 out) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    unsigned char *stage = smem + TAB_BYTES + warp * (HUE ? 0 : Pack<TOut>::STAGE);
    const RowWalk w(in.cols, in.frames, in.rows, BLOCK / 32);
    constexpr int CH = Pack<TOut>::CHUNKS;
    uint32_t rd_off[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int p = 32 * k + lane;
        rd_off[k] = (p / CH) * Pack<TOut>::LANE_STRIDE + (p % CH) * 16;
    }
    for (long long slot = (long long)blockIdx.x * (BLOCK / 32) + warp; slot < w.n_slots; slot += w.n_warps) {
        int seg;
        RowPos pos(w, slot, seg);
        const long long x0 = (long long)seg * TILE_PX;
        const int valid = (int)min((long long)TILE_PX, in.cols - x0);
        const bool lane_on = lane * 4 < valid;
        const long long in_off = x0 * PXB + lane * (4 * PXB) - (BGR ? 2 : 0);
        const long long out_off = x0 * (HUE ? 1 : 3) * (long long)sizeof(TOut);
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        auto fetch = [&](const RowPos &p_, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
            a = b = c = d = 0;
            if (p_.in(w) && lane_on) {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(in.base + p_.off(in) + in_off);
                if (PXB == 4) { const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p)); a = q.x; b = q.y; c = q.z; d = q.w; }
                else { a = ldg_u32(p); b = ldg_u32(p + 1); c = ldg_u32(p + 2); }
            }
        };
        fetch(pos, w0, w1, w2, w3);
        constexpr int UNROLL = HUE ? 1 : HUSL_ROWS_UNROLL;
#pragma unroll UNROLL
        for (; pos.in(w);) {
            const RowPos nxt = pos.next(w);
            uint32_t n0, n1, n2, n3;
            fetch(nxt, n0, n1, n2, n3);
            unsigned char *op = reinterpret_cast<unsigned char *>(out.base) + pos.off(out) + out_off;
            LinPx px[4];
            if (PXB == 4) {
                const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 v;
                    v = lut_lin<BGR ? 2 : 0>(smem, ww[j], lane8); px[j].hr = v.x; px[j].lr = v.y;
                    v = lut_lin<1>(smem, ww[j], lane8); px[j].hg = v.x; px[j].lg = v.y;
                    v = lut_lin<BGR ? 0 : 2>(smem, ww[j], lane8); px[j].hb = v.x; px[j].lb = v.y;
                }
            } else {
                lookup4(smem, w0, w1, w2, lane8, px);
            }
            if (HUE) {
                float h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = hue_from_diff(make_diff(px[j]));
                if (lane_on) {
                    if (sizeof(TOut) == 4)
                        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(op) + lane * 4) = make_float4(h[0], h[1], h[2], h[3]);
                    else
                        stg_f64x4(reinterpret_cast<double *>(op) + lane * 4, (double)h[0], (double)h[1], (double)h[2], (double)h[3]);
                }
            } else {
                float r[12];
#pragma unroll
                for (int j = 0; j < 4; ++j) husl_from_diff<true>(make_diff(px[j]), r[3 * j], r[3 * j + 1], r[3 * j + 2]);
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
                const int chunks = valid / 4 * CH;
#pragma unroll
                for (int k = 0; k < CH; ++k)
                    if (32 * k + lane < chunks) o[32 * k] = *reinterpret_cast<const float4 *>(stage + rd_off[k]);
                __syncwarp();
            }
            w0 = n0; w1 = n1; w2 = n2; w3 = n3;
            pos = nxt;
        }
    }
}

template <typename TIn, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_rows_inv(const View2 in, const View2 out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const RowWalk w(in.cols, in.frames, in.rows, BLOCK / 32);
    for (long long slot = (long long)blockIdx.x * (BLOCK / 32) + warp; slot < w.n_slots; slot += w.n_warps) {
        int seg;
        RowPos pos(w, slot, seg);
        const long long x0 = (long long)seg * TILE_PX;
        const int valid = (int)min((long long)TILE_PX, in.cols - x0);
        if (lane * 4 >= valid) continue;                              // no warp-level synchronisation below
        const long long in_off = (x0 + lane * 4) * 3 * (long long)sizeof(TIn);
        const long long out_off = (x0 + lane * 4) * 3;
        for (; pos.in(w); pos = pos.next(w)) {
            const unsigned char *ip = reinterpret_cast<const unsigned char *>(in.base) + pos.off(in) + in_off;
            float v[12];
            if (sizeof(TIn) == 4) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 q = __ldcs(reinterpret_cast<const float4 *>(ip) + k);
                    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double a, b, c, d;
                    ldg_f64x4(reinterpret_cast<const double *>(ip) + 4 * k, a, b, c, d);
                    v[4 * k] = (float)a; v[4 * k + 1] = (float)b; v[4 * k + 2] = (float)c; v[4 * k + 3] = (float)d;
                }
            }
            uint32_t q[12];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float Rr, G, B;
                rgb255_from_husl<false, wide_in<TIn>()>(v[3 * j], v[3 * j + 1], v[3 * j + 2], Rr, G, B);
                q[3 * j] = round_u8_bits(Rr); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
            }
            uint32_t o0, o1, o2;
            pack12(q, o0, o1, o2);
            uint32_t *dst = reinterpret_cast<uint32_t *>(out.base + pos.off(out) + out_off);
            dst[0] = o0; dst[1] = o1; dst[2] = o2;
        }
    }
}

// PXB = 4: RGBA / BGRA memory edited IN PLACE through rgba[..., :3] (out aliases in): one 128-bit load
// and one 128-bit store per lane, the fourth byte of every pixel is written back unchanged.
template <int BLOCK, int MINB, bool EH, bool ES, bool RSL, int PXB = 3, bool BGR = false>
__global__ void __launch_bounds__(BLOCK, MINB)
k_rows_edit(const View2 in, const View2 out, const nphusl_edit e) {
    extern __shared__ __align__(16) unsigned char smem[];
    fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const RowWalk w(in.cols, in.frames, in.rows, BLOCK / 32);
    for (long long slot = (long long)blockIdx.x * (BLOCK / 32) + warp; slot < w.n_slots; slot += w.n_warps) {
        int seg;
        RowPos pos(w, slot, seg);
        const long long x0 = (long long)seg * TILE_PX;
        const int valid = (int)min((long long)TILE_PX, in.cols - x0);
        if (lane * 4 >= valid) continue;                              // no warp-level synchronisation below
        const long long off = (x0 + lane * 4) * PXB - (BGR ? 2 : 0);
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        auto fetch = [&](const RowPos &p_, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
            a = b = c = d = 0;
            if (p_.in(w)) {
                const uint32_t *p = reinterpret_cast<const uint32_t *>(in.base + p_.off(in) + off);
                if (PXB == 4) { const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p)); a = q.x; b = q.y; c = q.z; d = q.w; }
                else { a = ldg_u32(p); b = ldg_u32(p + 1); c = ldg_u32(p + 2); }
            }
        };
        fetch(pos, w0, w1, w2, w3);
        for (; pos.in(w);) {
            const RowPos nxt = pos.next(w);
            uint32_t n0, n1, n2, n3;
            fetch(nxt, n0, n1, n2, n3);
            LinPx px[4];
            const uint32_t ww[4] = {w0, w1, w2, w3};
            if (PXB == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 v;
                    v = lut_lin<BGR ? 2 : 0>(smem, ww[j], lane8); px[j].hr = v.x; px[j].lr = v.y;
                    v = lut_lin<1>(smem, ww[j], lane8); px[j].hg = v.x; px[j].lg = v.y;
                    v = lut_lin<BGR ? 0 : 2>(smem, ww[j], lane8); px[j].hb = v.x; px[j].lb = v.y;
                }
            } else {
                lookup4(smem, w0, w1, w2, lane8, px);
            }
            uint32_t q[12];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float h, s, l, Rr, G, B;
                husl_from_diff<true>(make_diff(px[j]), h, s, l);
                apply_edit<EH, ES, RSL>(e, h, s, l);
                rgb255_from_husl(h, s, l, Rr, G, B);
                q[3 * j] = round_u8_bits(Rr); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
            }
            if (PXB == 4) {
                uint4 o;
                uint32_t *ow = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // bytes: (first, G, last, untouched 4th byte of the input word)
                    const uint32_t lo = __byte_perm(q[3 * j + (BGR ? 2 : 0)], q[3 * j + 1], 0x0040);
                    const uint32_t hi = __byte_perm(q[3 * j + (BGR ? 0 : 2)], ww[j], 0x0070);
                    ow[j] = __byte_perm(lo, hi, 0x5410);
                }
                *reinterpret_cast<uint4 *>(out.base + pos.off(out) + off) = o;
            } else {
                uint32_t o0, o1, o2;
                pack12(q, o0, o1, o2);
                uint32_t *dst = reinterpret_cast<uint32_t *>(out.base + pos.off(out) + off);
                dst[0] = o0; dst[1] = o1; dst[2] = o2;
            }
            w0 = n0; w1 = n1; w2 = n2; w3 = n3;
            pos = nxt;
        }
    }
}

}  // namespace husl
