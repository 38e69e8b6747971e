// husl_nv12_out.cuh -- the encode side of decode adjacency: uint8 RGB -> NV12, and the whole
// video loop NV12 -> HUSL -> select + edit -> NV12 in one pass (3 bytes of HBM traffic per pixel;
// what a hardware encoder, NVENC, takes as input).
//
//   k_rgb_to_nv12          interleaved uint8 RGB (H, W, 3) -> pitched Y plane + interleaved U,V plane
//   k_nv12_edit_nv12       NV12 -> RGB -> HUSL -> edit -> RGB -> NV12, all in registers
//   k_*_any                one thread per 2 x 2 block: widths that are not a multiple of 4, unaligned planes
//
// RGB -> YUV (integer, so oracle.rgb_to_nv12 in numpy is bit-exact): 16.16 coefficients of the
// BT.601 / BT.709 matrix (nphusl_cuda_api.cu: nv12_out_coef),
//     Y = clamp8((yr*R + yg*G + yb*B + 32768 + (y_off << 16)) >> 16)
//     U = clamp8((ur*Ra + ug*Ga + ub*Ba + 32768 + (128 << 16)) >> 16),  V likewise
// where (Ra, Ga, Ba) is the rounded mean of the 2 x 2 block, (sum + 2) >> 2 per channel.
// A lane owns a 4 x 2 block of pixels (two 2 x 2 chroma blocks): one luma word per row and one
// chroma word; a warp owns 128 pixels of a ROW PAIR and walks down the pairs like k_nv12.
#pragma once
#include "husl_nv12.cuh"

namespace husl {

struct Nv12OutCoef {
    int yr, yg, yb, yc;          // yc = 32768 + (y_off << 16)
    int ur, ug, ub, vr, vg, vb;  // chroma rows
    int cc;                      // 32768 + (128 << 16)
};

struct Nv12Dst {
    uint8_t *y, *uv;
    long long y_pitch, uv_pitch;
    Nv12OutCoef k;
};

// one row of 4 pixels: 12 channel values (low byte = value) -> luma word; adds the row's channel sums per chroma block
__device__ __forceinline__ uint32_t luma4(const Nv12OutCoef &k, const uint32_t (&q)[12], int (&sum)[6]) {
    uint32_t y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int R = (int)__byte_perm(q[3 * j], 0, 0x4440), G = (int)__byte_perm(q[3 * j + 1], 0, 0x4440),
                  B = (int)__byte_perm(q[3 * j + 2], 0, 0x4440);
        y[j] = clamp16_16(k.yr * R + k.yg * G + k.yb * B + k.yc);
        sum[3 * (j >> 1)] += R; sum[3 * (j >> 1) + 1] += G; sum[3 * (j >> 1) + 2] += B;
    }
    return __byte_perm(__byte_perm(y[0], y[1], 0x0062), __byte_perm(y[2], y[3], 0x0062), 0x5410);
}

// two chroma blocks (channel sums of 4 pixels each) -> U0 V0 U1 V1
__device__ __forceinline__ uint32_t chroma2(const Nv12OutCoef &k, const int (&sum)[6]) {
    uint32_t c[4];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int R = (sum[3 * b] + 2) >> 2, G = (sum[3 * b + 1] + 2) >> 2, B = (sum[3 * b + 2] + 2) >> 2;
        c[2 * b] = clamp16_16(k.ur * R + k.ug * G + k.ub * B + k.cc);
        c[2 * b + 1] = clamp16_16(k.vr * R + k.vg * G + k.vb * B + k.cc);
    }
    return __byte_perm(__byte_perm(c[0], c[1], 0x0062), __byte_perm(c[2], c[3], 0x0062), 0x5410);
}

// 4 pixels: 3 RGB words -> 12 values whose low byte is the channel
__device__ __forceinline__ void unpack12(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&q)[12]) {
    q[0] = w0; q[1] = w0 >> 8; q[2] = w0 >> 16; q[3] = w0 >> 24;
    q[4] = w1; q[5] = w1 >> 8; q[6] = w1 >> 16; q[7] = w1 >> 24;
    q[8] = w2; q[9] = w2 >> 8; q[10] = w2 >> 16; q[11] = w2 >> 24;
}

// SRC_NV12 = true: fused NV12 -> edit -> NV12 (src); false: uint8 RGB (rgb, C-contiguous H x W x 3) -> NV12
template <bool SRC_NV12, int BLOCK, int MINB, bool EH, bool ES, bool RSL>
__global__ void __launch_bounds__(BLOCK, MINB)
k_to_nv12(const Nv12Src src, const uint8_t *__restrict__ rgb, int W, int H, const Nv12Dst dst, const nphusl_edit e) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (SRC_NV12) fill_table(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane8 = lane * 8;
    const int segs = (W + TILE_PX - 1) / TILE_PX, pairs = H / 2;
    const long long n_warps = (long long)gridDim.x * (BLOCK / 32);
    const int step = (int)max(1LL, min(n_warps / segs, (long long)pairs));       // in row pairs
    const long long n_slots = (long long)segs * step;
    for (long long slot = (long long)blockIdx.x * (BLOCK / 32) + warp; slot < n_slots; slot += n_warps) {
        const int x = (int)(slot % segs) * TILE_PX + lane * 4;
        if (x >= W) continue;                                                     // no warp-level synchronisation below
        for (int pr = (int)(slot / segs); pr < pairs; pr += step) {
            uint32_t qa[12], qb[12];
            if (SRC_NV12) {
                const uint32_t ya = ldg_u32(reinterpret_cast<const uint32_t *>(src.y + (2LL * pr) * src.y_pitch + x));
                const uint32_t yb = ldg_u32(reinterpret_cast<const uint32_t *>(src.y + (2LL * pr + 1) * src.y_pitch + x));
                const uint32_t uv = ldg_u32(reinterpret_cast<const uint32_t *>(src.uv + (long long)pr * src.uv_pitch + x));
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t c[12];
                    yuv4(src.k, r ? yb : ya, uv, c);
                    uint32_t (&q)[12] = r ? qb : qa;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float h, s, l, R, G, B;
                        husl_from_diff<true>(make_diff(lin_px(smem, c[3 * j], c[3 * j + 1], c[3 * j + 2], lane8)), h, s, l);
                        apply_edit<EH, ES, RSL>(e, h, s, l);
                        rgb255_from_husl(h, s, l, R, G, B);
                        q[3 * j] = round_u8_bits(R); q[3 * j + 1] = round_u8_bits(G); q[3 * j + 2] = round_u8_bits(B);
                    }
                }
            } else {
                const uint32_t *pa = reinterpret_cast<const uint32_t *>(rgb + ((2LL * pr) * W + x) * 3);
                const uint32_t *pb = reinterpret_cast<const uint32_t *>(rgb + ((2LL * pr + 1) * W + x) * 3);
                unpack12(ldg_u32(pa), ldg_u32(pa + 1), ldg_u32(pa + 2), qa);
                unpack12(ldg_u32(pb), ldg_u32(pb + 1), ldg_u32(pb + 2), qb);
            }
            int sum[6] = {0, 0, 0, 0, 0, 0};
            const uint32_t oa = luma4(dst.k, qa, sum), ob = luma4(dst.k, qb, sum);
            *reinterpret_cast<uint32_t *>(dst.y + (2LL * pr) * dst.y_pitch + x) = oa;
            *reinterpret_cast<uint32_t *>(dst.y + (2LL * pr + 1) * dst.y_pitch + x) = ob;
            *reinterpret_cast<uint32_t *>(dst.uv + (long long)pr * dst.uv_pitch + x) = chroma2(dst.k, sum);
        }
    }
}

// one thread per 2 x 2 block: even width / height only, any alignment
template <bool SRC_NV12>
__global__ void __launch_bounds__(256)
k_to_nv12_any(const Nv12Src src, const uint8_t *__restrict__ rgb, int W, int H, const Nv12Dst dst, const nphusl_edit e) {
    const long long bw = W / 2, n = bw * (H / 2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long pr = i / bw;
        const int x = (int)(i - pr * bw) * 2;
        int sR = 0, sG = 0, sB = 0;
        int cr = 0, cg = 0, cb = 0;
        if (SRC_NV12) {
            const uint8_t *uvp = src.uv + pr * src.uv_pitch + x;
            chroma_terms(src.k, uvp[0], uvp[1], cr, cg, cb);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const long long row = 2 * pr + r;
                uint32_t R, G, B;
                if (SRC_NV12) {
                    yuv_px(src.k, src.y[row * src.y_pitch + x + dx], cr, cg, cb, R, G, B);
                    const float2 tr = g_lin_tab[R >> 16], tg = g_lin_tab[G >> 16], tb = g_lin_tab[B >> 16];
                    LinPx p;
                    p.hr = tr.x; p.lr = tr.y; p.hg = tg.x; p.lg = tg.y; p.hb = tb.x; p.lb = tb.y;
                    float h, s, l, fr, fg, fb;
                    husl_from_diff<true>(make_diff(p), h, s, l);
                    apply_edit(e, h, s, l);
                    rgb255_from_husl(h, s, l, fr, fg, fb);
                    R = round_u8_bits(fr) & 0xFF; G = round_u8_bits(fg) & 0xFF; B = round_u8_bits(fb) & 0xFF;
                } else {
                    const uint8_t *p = rgb + (row * W + x + dx) * 3;
                    R = p[0]; G = p[1]; B = p[2];
                }
                dst.y[row * dst.y_pitch + x + dx] =
                    (uint8_t)(clamp16_16(dst.k.yr * (int)R + dst.k.yg * (int)G + dst.k.yb * (int)B + dst.k.yc) >> 16);
                sR += (int)R; sG += (int)G; sB += (int)B;
            }
        const int R = (sR + 2) >> 2, G = (sG + 2) >> 2, B = (sB + 2) >> 2;
        uint8_t *o = dst.uv + pr * dst.uv_pitch + x;
        o[0] = (uint8_t)(clamp16_16(dst.k.ur * R + dst.k.ug * G + dst.k.ub * B + dst.k.cc) >> 16);
        o[1] = (uint8_t)(clamp16_16(dst.k.vr * R + dst.k.vg * G + dst.k.vb * B + dst.k.cc) >> 16);
    }
}

}  // namespace husl
