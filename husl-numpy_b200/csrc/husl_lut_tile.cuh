// husl_lut_tile.cuh -- the bit-exact LUT mode (husl_lut.cuh: the reference's default _simd.c build) on the
// streaming skeleton of husl_stream.cuh.
//
// Same values as k_fwd_lut, bit for bit; what changes is everything around the arithmetic:
//   * a warp owns a tile of 128 consecutive pixels, lanes read their 4 pixels as three coalesced 32-bit
//     words (register-prefetched one tile ahead) instead of twelve scalar byte loads per warp-row;
//   * the nine products coef * linear[c] of the XYZ matrix come from a 18 KB shared-memory table
//     (LutParams::prod): 9 DMUL per pixel become 9 LDS.64;
//   * results are written into a shared-memory image of the output tile and leave through ONE 3 KB bulk
//     async copy per tile (cp.async.bulk.global.shared::cta, UBLKCP), double-buffered -- no 8-byte
//     stores at a 24-byte lane stride;
//   * the light table is read as pre-divided doubles, the chroma table stays in L2 behind ld.global.nc.
// The kernel is bound by the FP64 pipe (about 115 DP instructions per pixel: two correctly rounded
// divisions by per-pixel divisors, one square root, one reciprocal -- bit-exactness against the
// reference's IEEE double chain demands them), not by HBM (27 B/px).
#pragma once
#include "husl_stream.cuh"
#include "husl_lut.cuh"

namespace husl {

constexpr int LUT_PROD_BYTES = 9 * 256 * 8;

template <bool INTERP, typename TT, int BLOCK, int MINB, int OST>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_lut_tile(const uint32_t *__restrict__ in, double *__restrict__ out, long long n_tiles, const LutParams t) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *sprod = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < 9 * 256; i += BLOCK) sprod[i] = t.prod[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<double>::BYTES;
    unsigned char *obuf = smem + LUT_PROD_BYTES + warp * (OST * TB);
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
        // the bulk store that last used this buffer must have finished reading it
        if (lane == 0) bulk_wait_read<OST - 1>();
        __syncwarp();
        double *wr = reinterpret_cast<double *>(obuf + st * TB + lane * Tile<double>::LANE);
        // bytes of the lane's 12-byte record: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const uint32_t c[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll 2
        for (int j = 0; j < 4; ++j) {
            double H, S, L;
            lut_pixel<INTERP, TT>(t, sprod, c[3 * j], c[3 * j + 1], c[3 * j + 2], H, S, L);
            wr[3 * j] = H; wr[3 * j + 1] = S; wr[3 * j + 2] = L;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store(reinterpret_cast<unsigned char *>(out) + tile * TB, obuf_s + st * TB, TB);
            bulk_commit();
        }
        st = (st + 1 == OST) ? 0 : st + 1;
        w0 = n0; w1 = n1; w2 = n2;
    }
    if (lane == 0) bulk_wait_all<0>();     // shared memory must outlive the copies
}

}  // namespace husl
