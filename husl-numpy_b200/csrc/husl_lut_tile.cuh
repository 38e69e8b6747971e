// husl_lut_tile.cuh -- the bit-exact LUT mode (husl_lut.cuh: the reference's default _simd.c build) as a
// tile-streaming kernel.
//
// Same values as k_fwd_lut, bit for bit; what changes is everything around the arithmetic:
//   * a warp owns a tile of 128 consecutive pixels, lanes read their 4 pixels as three coalesced 32-bit
//     words (register-prefetched one tile ahead) instead of twelve scalar byte loads per warp-row;
//   * the nine products coef * linear[c] of the XYZ matrix come from an 18 KB shared-memory table
//     (LutParams::prod: 9 DMUL per pixel become 9 LDS.64) and the light table sits beside it as
//     pre-divided doubles (24 KB at the default 3 x 1024 entries) -- an L2 round trip less per pixel;
//     the 2 MB chroma table stays in L2 behind ld.global.nc;
//   * the per-pixel code is straight-line (husl_lut.cuh), and two pixels of a lane are in flight at a
//     time: the kernel is latency-bound (a pixel is a dependent chain of ~100 FP64 operations plus the
//     chroma gather), so what it needs is independent chains per scheduler -- 2 CTAs x 16 warps per SM;
//   * results are transposed through a 3 KB per-warp staging buffer and leave as coalesced 16-byte
//     stores (no 8-byte stores at a 24-byte lane stride).
// Bound by FP64 latency / issue (about 100 DP instructions per pixel: three divisions by per-pixel
// divisors, one square root, one reciprocal -- bit-exactness against the reference's IEEE double chain
// demands them), not by HBM (27 B/px).
#pragma once
#include "husl_stream.cuh"
#include "husl_lut.cuh"

namespace husl {

constexpr int LUT_PROD_BYTES = 9 * 256 * 8;
constexpr int LUT_LIGHT_SMEM_MAX = 3 * 1024 * 8;      // light tables up to the default size live in shared memory

template <bool INTERP, typename TT, int BLOCK, int MINB, bool LIGHT_SMEM, int NP>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_lut_tile(const uint32_t *__restrict__ in, double *__restrict__ out, long long n_tiles, const LutParams t) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *sprod = reinterpret_cast<double *>(smem);
    double *slight = reinterpret_cast<double *>(smem + LUT_PROD_BYTES);
    const int n_light = LIGHT_SMEM ? 3 * t.seg : 0;
    for (int i = threadIdx.x; i < 9 * 256; i += BLOCK) sprod[i] = t.prod[i];
    for (int i = threadIdx.x; i < n_light; i += BLOCK) slight[i] = t.light_q[i];
    __syncthreads();
    const double *light_q = LIGHT_SMEM ? slight : t.light_q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<double>::BYTES;
    unsigned char *stage = smem + LUT_PROD_BYTES + (LIGHT_SMEM ? LUT_LIGHT_SMEM_MAX : 0) + warp * TB;
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
        double *wr = reinterpret_cast<double *>(stage + lane * Tile<double>::LANE);
        // the lane's 12-byte record: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3; LUT_NP pixels in flight at a time
        const uint32_t c[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll
        for (int q = 0; q < 4 / NP; ++q) {
            uint32_t cc[3 * NP];
            double r[3 * NP];
#pragma unroll
            for (int k = 0; k < 3 * NP; ++k) cc[k] = c[3 * NP * q + k];
            lut_pixels<INTERP, TT, NP>(t, sprod, light_q, cc, r);
#pragma unroll
            for (int k = 0; k < 3 * NP; ++k) wr[3 * NP * q + k] = r[k];
        }
        __syncwarp();
        // the tile image in shared memory equals the global one: six coalesced 512-byte rows
        const float4 *rd = reinterpret_cast<const float4 *>(stage) + lane;
        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(out) + tile * TB) + lane;
#pragma unroll
        for (int k = 0; k < 6; ++k) dst[32 * k] = rd[32 * k];
        __syncwarp();
        w0 = n0; w1 = n1; w2 = n2;
    }
}

}  // namespace husl
