// husl_lut_tile.cuh -- the bit-exact LUT mode (husl_lut.cuh: the reference's default _simd.c build) as a
// tile-streaming kernel.
//
// Same values as k_fwd_lut, bit for bit; what changes is everything around the arithmetic:
//   * a warp owns a tile of 128 consecutive pixels, lanes read their 4 pixels as three coalesced 32-bit
//     words (register-prefetched one tile ahead) instead of twelve scalar byte loads per warp-row;
//   * the 256-entry table of linear light sits in shared memory REPLICATED per half-warp lane
//     ([256][16] doubles, 32 KB): lane l reads column l & 15, so the three lookups of a pixel are
//     conflict-free 64-bit loads (2 wavefronts each).  Round 2's first tile kernel held the nine products
//     coef * linear[c] as a 9 x 256 table instead; ncu showed what that costs: nine lookups per pixel at
//     random addresses are ~6.5 shared-memory wavefronts each, and the kernel ran with the L1 data pipe
//     96.6 % busy (l1tex__data_pipe_lsu_wavefronts) -- bound by bank conflicts, not by FP64.  Nine DMUL
//     per pixel are cheaper than those wavefronts;
//   * the light table sits beside it as pre-divided doubles (24 KB at the default 3 x 1024 entries); the
//     2 MB chroma table stays in L2 behind ld.global.nc;
//   * the per-pixel code is straight-line (husl_lut.cuh), NP pixels of a lane in flight at a time;
//   * results are written into a per-warp staging image of the output tile and leave through bulk copies
//     (cp.async.bulk, no read-back through registers).  A lane's 96-byte record would start every fourth lane
//     in the same 16-byte bank group; one 16-byte pad per four records makes the 128-bit stores of a
//     quarter-warp hit eight different groups, and the tile leaves as eight 384-byte copies issued by
//     lanes 0..7 instead of one.
#pragma once
#include "husl_stream.cuh"
#include "husl_lut.cuh"

namespace husl {

#ifndef NPHUSL_LUT_COPIES
#define NPHUSL_LUT_COPIES 16
#endif
constexpr int LUT_COPIES = NPHUSL_LUT_COPIES;            // replicas of the linear table (16: one per half-warp lane)
constexpr int LUT_LIN_BYTES = 256 * LUT_COPIES * 8;
constexpr int LUT_LIGHT_SMEM_MAX = 3 * 1024 * 8;          // light tables up to the default size live in shared memory
constexpr int LUT_STAGE_BYTES = Tile<double>::BYTES + 8 * 16;   // 32 records of 96 bytes + a 16-byte pad per 4 records

template <bool LIGHT_SMEM, int BLOCK> constexpr int lut_tile_smem() {
    return LUT_LIN_BYTES + (LIGHT_SMEM ? LUT_LIGHT_SMEM_MAX : 0) + (BLOCK / 32) * LUT_STAGE_BYTES;
}

// The complete function (husl_lut.cuh, FAST = false) for the four pixels of a lane whose straight-line pass
// met a pixel it does not answer (black, white, u == 0): rarely run, kept out of line and rolled so that it
// costs the streaming loop no registers.  w0..w2 = the lane's 12-byte record, wr = its staging record.
template <bool INTERP, typename TT>
__device__ __noinline__ void lut_redo_lane(const LutParams t, uint32_t w0, uint32_t w1, uint32_t w2, double *wr) {
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        uint32_t c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = 3 * p + k;
            const uint32_t w = j < 4 ? w0 : (j < 8 ? w1 : w2);
            c[k] = (w >> (8 * (j & 3))) & 255u;
        }
        double r[3];
        const auto xyz = [&](int, double &x, double &y, double &z) {
            lut_xyz(__ldg(t.linear + c[0]), __ldg(t.linear + c[1]), __ldg(t.linear + c[2]), x, y, z);
        };
        lut_pixels<INTERP, TT, 1, false>(t, xyz, t.light_q, c, r);
        wr[3 * p] = r[0]; wr[3 * p + 1] = r[1]; wr[3 * p + 2] = r[2];
    }
}

template <bool INTERP, typename TT, int BLOCK, int MINB, bool LIGHT_SMEM, int NP>
__global__ void __launch_bounds__(BLOCK, MINB)
k_fwd_lut_tile(const uint32_t *__restrict__ in, double *__restrict__ out, long long n_tiles, const LutParams t) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *slin = reinterpret_cast<double *>(smem);
    double *slight = reinterpret_cast<double *>(smem + LUT_LIN_BYTES);
    const int n_light = LIGHT_SMEM ? 3 * t.seg : 0;
    for (int i = threadIdx.x; i < 256 * LUT_COPIES; i += BLOCK) slin[i] = t.linear[i / LUT_COPIES];
    for (int i = threadIdx.x; i < n_light; i += BLOCK) slight[i] = t.light_q[i];
    __syncthreads();
    const double *light_q = LIGHT_SMEM ? slight : t.light_q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TB = Tile<double>::BYTES;
    unsigned char *stage = smem + LUT_LIN_BYTES + (LIGHT_SMEM ? LUT_LIGHT_SMEM_MAX : 0) + warp * LUT_STAGE_BYTES;
    const double *mylin = slin + (lane & (LUT_COPIES - 1));
    double *wr = reinterpret_cast<double *>(stage + lane * Tile<double>::LANE + (lane >> 2) * 16);
    // the tile leaves as eight bulk copies of 384 bytes (four records between two pads), one per lane 0..7
    const uint32_t seg_s = smem_addr(stage) + lane * (4 * Tile<double>::LANE + 16);
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
        // the lane's 12-byte record: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3; NP pixels in flight at a time
        const uint32_t c[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
        // the bulk copies of the previous tile must have finished reading the staging buffer
        if (lane < 8) bulk_wait_read<0>();
        __syncwarp();
        bool special = false;
#pragma unroll
        for (int q = 0; q < 4 / NP; ++q) {
            uint32_t cc[3 * NP];
            double lin[3 * NP], r[3 * NP];
#pragma unroll
            for (int k = 0; k < 3 * NP; ++k) {
                cc[k] = c[3 * NP * q + k];
                lin[k] = mylin[cc[k] * LUT_COPIES];
            }
            const auto xyz = [&](int p, double &x, double &y, double &z) {
                lut_xyz(lin[3 * p], lin[3 * p + 1], lin[3 * p + 2], x, y, z);
            };
            special |= lut_pixels<INTERP, TT, NP, true>(t, xyz, light_q, cc, r);
#pragma unroll
            for (int k = 0; k < 3 * NP; ++k) wr[3 * NP * q + k] = r[k];
        }
        if (__any_sync(0xffffffffu, special)) {       // black, white or u == 0 somewhere in the tile (rare)
            if (special) lut_redo_lane<INTERP, TT>(t, w0, w1, w2, wr);
        }
        fence_async_smem();
        __syncwarp();
        if (lane < 8) {
            bulk_store(reinterpret_cast<unsigned char *>(out) + tile * TB + lane * (4 * Tile<double>::LANE), seg_s, 4 * Tile<double>::LANE);
            bulk_commit();
        }
        w0 = n0; w1 = n1; w2 = n2;
    }
    if (lane < 8) bulk_wait_all<0>();     // shared memory must outlive the copies
}

}  // namespace husl
