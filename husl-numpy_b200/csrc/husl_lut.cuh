// husl_lut.cuh -- optional LUT mode: the reference's DEFAULT _simd.c build
// (-DUSE_LIGHT_LUT -DUSE_CHROMA_LUT -DUSE_HUE_ATAN2_APPROX, setup.py:99-128,
// optionally -DINTERPOLATE_CHROMA) reproduced on the device, plus device-side
// generation of the tables make_lookup_tables / _linear_lookup.c produce.
//
// The goal here is BIT-EXACT agreement with nphusl/_simd.c compiled without
// -ffast-math, so this file is IEEE double arithmetic in the reference's
// operation order, written with __dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn so
// nvcc can neither contract a*b+c into an FMA nor reassociate.  It is a
// gather-bound float64 kernel (2 MB chroma table lives in L2); the fast path of
// the library is the float32 EXACT mode in husl_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace husl {

// Tables keep the element type the reference generator was asked for
// (make_lookup_tables:4 `table_type`, LUT/light.py:21-32): uint16 holding the value
// times the int scale (the default, LIGHT_SCALE = CHROMA_SCALE = 100), or float /
// double holding the 4- / 3-decimal literal itself (scales 1).  The kernels are
// templated on the element type; the arithmetic on the promoted double is the same.
//
// A derived table (k_lut_prepare, built once when tables are installed) takes an IEEE operation
// whose operands come from a table out of the per-pixel path without changing a bit:
//   light_q[i]  = light_table_big[i] / LIGHT_SCALE                                  (_simd.c:495)
struct LutParams {
    const double *linear;     // 256, 6-decimal values (_linear_lookup.c:9-42)
    const void *light;        // 3*seg entries (light_table_big)
    const void *chroma;       // nh*nl entries (chroma_table)
    const double *light_q;    // 3*seg (derived)
    int seg, nh, nl;
    double ystep[3], ythresh[2], hstep, lstep;
    double r_ystep[3], r_hstep, r_lstep;   // RN(1 / step), for the exactly rounded divisions below
    double yorg[3], ybase[3];              // per light segment: origin (0, ythresh[0], ythresh[1]) and index base (0, seg, 2 seg)
    double light_scale;       // LIGHT_SCALE
    double sat_scale;         // 100 * CHROMA_SCALE (_simd.c:334)
};

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dd(double a, double b) { return __ddiv_rn(a, b); }

// ---- straight-line IEEE division / reciprocal / square root -----------------------------------------
// __ddiv_rn, __drcp_rn and __dsqrt_rn each carry a range check and a CALL to a slow path (denormals,
// overflow); five of them per pixel cut the pixel's instruction stream into branch regions the compiler
// cannot interleave, and this kernel lives on instruction-level parallelism: one pixel is a chain of
// ~100 dependent FP64 operations plus a table gather that is answered by L2.  Every operand below is a
// normal number well inside the double range (linear light, Luv coordinates, table steps), so the fast
// paths alone are written out: a hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20), Newton / Goldschmidt
// refinement in FMAs, and a final residual correction (Markstein) that makes the last rounding the
// correctly rounded one.  What the kernel computes is pinned bit for bit by the sha256 of all 2^24 colours
// against the reference's own C output (tests/test_cuda_parity.py::test_lut_mode_bit_exact_on_cube).
__device__ __forceinline__ double rcp_seed(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
}
__device__ __forceinline__ double rsqrt_seed(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}
// 1/b to within an ulp: r0 (1 + e + e^2), e = 1 - b r0
__device__ __forceinline__ double rcp_near(double b) {
    const double r0 = rcp_seed(b);
    const double e = __fma_rn(-b, r0, 1.0);
    return __fma_rn(r0, __fma_rn(e, e, e), r0);
}
// a / b given rb ~ 1/b (within an ulp): q0 = RN(a*rb) is within an ulp of the quotient, the residual
// r = a - q0*b is exact in one FMA, and RN(q0 + r*rb) is the correctly rounded quotient (Markstein's
// division step).  3 DP instructions; used directly where the divisor is a run-time constant of the
// tables (index steps) or shared by two quotients (u', v').
__device__ __forceinline__ double cdiv(double a, double b, double rb) {
    const double q0 = __dmul_rn(a, rb);
    const double r = __fma_rn(-q0, b, a);
    return __fma_rn(r, rb, q0);
}
__device__ __forceinline__ double div_rn(double a, double b) { return cdiv(a, b, rcp_near(b)); }
// sqrt(a), a > 0: Goldschmidt on g ~ sqrt(a), h ~ 1/(2 sqrt(a)), then g + (a - g*g) * h
template <bool GUARD = true> __device__ __forceinline__ double sqrt_rn(double a) {
    const double y = rsqrt_seed(a);
    double g = __dmul_rn(a, y), h = __dmul_rn(0.5, y);
    double r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
    r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g);
    if (GUARD) h = __fma_rn(h, r, h);   // the correction below is ~2^-53 g: h good to 2^-40 after one refinement serves it
    const double d = __fma_rn(-g, g, a);
    const double res = __fma_rn(d, h, g);
    return (GUARD && a == 0.0) ? 0.0 : res;
}

// (unsigned short)fmax(0, fmin(hi, (double)roundf((float)x))) of _simd.c:403-404,494 in integer registers:
// roundf is trunc(x + copysign(0.5, x)) with the sum rounded toward zero (so 0.49999997f stays below 1),
// exact for |x| < 2^23; the clamp to [0, hi] then runs on the integer.  x is an index of a few thousand at
// most, never NaN for a pixel whose result is used (black and white are answered with constants).
__device__ __forceinline__ int round_clamp(double x, int hi) {
    const float f = __double2float_rn(x);
    const int i = __float2int_rz(__fadd_rz(f, copysignf(0.5f, f)));
    return min(max(i, 0), hi);
}
// ... and with floorf (:364-365)
__device__ __forceinline__ int floor_clamp(double x, int hi) {
    const int i = __float2int_rd(__double2float_rn(x));
    return min(max(i, 0), hi);
}
// The same two for the tile kernel's straight-line path (FAST): every index it rounds is >= 0 for a pixel whose
// result is used (y - origin >= 0 by the choice of the segment, H in [0, 360], L >= 0), so copysign and the
// lower clamp go; an unsigned minimum keeps whatever a NaN / infinity of a pixel that is answered elsewhere
// (black, white, u == 0: see lut_pixels) converts to inside the table.
__device__ __forceinline__ int round_clamp_pos(double x, int hi) {
    const float f = __double2float_rn(x);
    return (int)min((unsigned)__float2int_rz(__fadd_rz(f, 0.5f)), (unsigned)hi);
}
__device__ __forceinline__ int floor_clamp_pos(double x, int hi) {
    return (int)min((unsigned)__float2int_rd(__double2float_rn(x)), (unsigned)hi);
}
template <bool FAST> __device__ __forceinline__ int round_idx(double x, int hi) {
    return FAST ? round_clamp_pos(x, hi) : round_clamp(x, hi);
}
template <bool FAST> __device__ __forceinline__ int floor_idx(double x, int hi) {
    return FAST ? floor_clamp_pos(x, hi) : floor_clamp(x, hi);
}

// constants whose low word is not zero live in constant memory: FP64 instructions take them as operands
// straight from the bank instead of re-materialising two 32-bit halves per use
__constant__ double LUT_K[21] = {0.78539816339744830962, 0.2447, 0.0663, 180.0 / 3.14159265358979323846,
                                 3.141592653589793, 1.5707963267948966, 0.28, 0.19783000664283, 0.46831999493879,
                                 19.916405993809086, 1e-10, (double)1e-10f,
                                 0.412391, 0.357584, 0.180481, 0.212639, 0.715169, 0.072192, 0.019331, 0.119195, 0.950532};
enum { K_PI4 = 0, K_A, K_B, K_DEG, K_PI, K_PIBY2, K_028, K_REFU, K_REFV, K_WHITE_HUE, K_TINY, K_TINYF,
       K_XR, K_XG, K_XB, K_YR, K_YG, K_YB, K_ZR, K_ZG, K_ZB };

// _simd.c:485-496.  The three segments differ only in (origin, step, index base): selected first, divided
// once.  Segment 0 is y / step exactly as written there (y - 0.0 and + 0.0 change nothing).
__device__ __forceinline__ double lut_light(const LutParams &t, const double *light_q, double y) {
    const bool s0 = y < t.ythresh[0], s1 = y < t.ythresh[1];
    const double org = s0 ? 0.0 : (s1 ? t.ythresh[0] : t.ythresh[1]);
    const double step = s0 ? t.ystep[0] : (s1 ? t.ystep[1] : t.ystep[2]);
    const double rstep = s0 ? t.r_ystep[0] : (s1 ? t.r_ystep[1] : t.r_ystep[2]);
    const double base = s0 ? 0.0 : (s1 ? (double)t.seg : (double)(t.seg * 2));
    const double idx = da(cdiv(ds(y, org), step, rstep), base);
    return light_q[round_clamp(idx, 3 * t.seg - 1)];
}

// The same with the segment's four parameters read from the kernel parameters by segment number (the count of
// thresholds at or below y): indexed constant loads (LDC, the otherwise idle ADU pipe) instead of eight
// 64-bit selects; a shared-memory table costs four wavefronts per 16-byte load even when all lanes read one of
// three addresses, and the L1 data pipe is one of this kernel's three near-limits.
__device__ __forceinline__ double lut_light_seg(const LutParams &t, const double *light_q, double y) {
    const int row = (y >= t.ythresh[0] ? 1 : 0) + (y >= t.ythresh[1] ? 1 : 0);
    const double idx = da(cdiv(ds(y, t.yorg[row]), t.ystep[row], t.r_ystep[row]), t.ybase[row]);
    return light_q[round_clamp_pos(idx, 3 * t.seg - 1)];
}

// _simd.c:279-281
__device__ __forceinline__ double lut_atan(double z) {
    const double az = fabs(z);
    return ds(dm(LUT_K[K_PI4], z), dm(dm(z, ds(az, 1.0)), da(LUT_K[K_A], dm(LUT_K[K_B], az))));
}

// _simd.c:246-266 with :288-315 inlined, branch-free.  to_hue calls the atan2 approximation only when
// |v/u| >= 1, and that function divides the same two numbers again: its z is this z, its "|z| < 1" branch
// cannot be taken, and what is left is the u == 0 guard and one division.  Both sides are evaluated and
// one is selected (a warp of a natural image needs both anyway).
__device__ __forceinline__ double lut_hue(double u, double v) {
    const double DEG_PER_RAD = LUT_K[K_DEG], PI_ = LUT_K[K_PI], PIBY2 = LUT_K[K_PIBY2];
    const double z = div_rn(v, u);
    double ha = dm(lut_atan(z), DEG_PER_RAD);
    ha = (u < 0) ? da(ha, 180.0) : ((v < 0) ? da(ha, 360.0) : ha);
    double at = ds(PIBY2, div_rn(z, da(dm(z, z), LUT_K[K_028])));
    at = (v < 0.0) ? ds(at, PI_) : at;
    at = (u == 0.0) ? ((v > 0.0) ? PIBY2 : ((v == 0.0) ? 0.0 : -PIBY2)) : at;
    double hb = dm(at, DEG_PER_RAD);
    hb = (hb < 0.0) ? da(hb, 360.0) : hb;
    return (fabs(z) < 1.0) ? ha : hb;
}

// lut_hue for u != 0 (the tile kernel answers u == 0 elsewhere, see lut_pixels), with the sign tests and the
// selections on the integer side: FP64 compares and selects of double pairs share the FP64 pipe / cost two
// FSEL each, and nvcc turns the nested ?: above into divergent branches.  With u != 0 neither u nor v can be
// -0.0 (v = 13 L (v' - REF_V) with L > 0; x - x is +0), so "< 0" is the sign bit; the additive constants
// 180, 360 (low words zero) and 0 are selected as high words and added unconditionally: ha + 0, at - 0 and
// hb + 0 return their operand bit for bit because none of them can be -0.0 here (ha = -0 needs z = -0, i.e.
// v = +0 with u < 0, which takes + 180; at and hb are differences that round to +0 when they vanish).
__device__ __forceinline__ double lut_hue_nz(double u, double v, bool &z_not_finite) {
    const double DEG_PER_RAD = LUT_K[K_DEG], PI_ = LUT_K[K_PI], PIBY2 = LUT_K[K_PIBY2];
    const int uh = __double2hiint(u), vh = __double2hiint(v);
    const double z = div_rn(v, u);
    double ha = dm(lut_atan(z), DEG_PER_RAD);
    ha = da(ha, __hiloint2double((uh < 0) ? 0x40668000 : ((vh < 0) ? 0x40768000 : 0), 0));
    double at = ds(PIBY2, div_rn(z, da(dm(z, z), LUT_K[K_028])));
    at = ds(at, __hiloint2double((vh < 0) ? __double2hiint(PI_) : 0, (vh < 0) ? __double2loint(PI_) : 0));
    double hb = dm(at, DEG_PER_RAD);
    hb = da(hb, __hiloint2double((__double2hiint(hb) < 0) ? 0x40768000 : 0, 0));
    const unsigned az = (unsigned)(__double2hiint(z) & 0x7fffffff);
    const bool below_one = az < 0x3ff00000u;     // |z| < 1, false for NaN
    z_not_finite = az >= 0x7ff00000u;            // u == 0 (v / 0 is an infinity or NaN), or a NaN u or v
    return __hiloint2double(below_one ? __double2hiint(ha) : __double2hiint(hb),
                            below_one ? __double2loint(ha) : __double2loint(hb));
}

// _simd.c:399-406 / :360-387
template <bool INTERP, typename TT, bool FAST>
__device__ __forceinline__ double lut_chroma(const LutParams &t, double l, double h) {
    const TT *chroma = static_cast<const TT *>(t.chroma);
    const double hs = cdiv(h, t.hstep, t.r_hstep), ls = cdiv(l, t.lstep, t.r_lstep);
    if (!INTERP) {
        const int hi = round_idx<FAST>(hs, t.nh - 1), li = round_idx<FAST>(ls, t.nl - 1);
        const TT raw = __ldg(chroma + (unsigned)(hi * t.nl + li));
        if (FAST && sizeof(TT) == 2) {      // fmax(1e-10, entry) of an unsigned integer entry: 1e-10 where it is 0
            const double cv = (double)raw, tiny = LUT_K[K_TINY];
            return __hiloint2double(raw == 0 ? __double2hiint(tiny) : __double2hiint(cv), raw == 0 ? __double2loint(tiny) : __double2loint(cv));
        }
        return fmax(LUT_K[K_TINY], (double)raw);
    }
    const int hf = floor_idx<FAST>(hs, t.nh - 2), lf = floor_idx<FAST>(ls, t.nl - 2);
    const TT *c0 = chroma + (unsigned)(hf * t.nl + lf);
    const double c00 = __ldg(c0), c10 = __ldg(c0 + t.nl), c01 = __ldg(c0 + 1), c11 = __ldg(c0 + t.nl + 1);
    const double hn = ds(hs, (double)hf), ln = ds(ls, (double)lf);
    const double hv = ds(1.0, hn), lv = ds(1.0, ln);
    const double c = da(da(da(dm(dm(c00, hv), lv), dm(dm(c10, hn), lv)), dm(dm(c01, hv), ln)), dm(dm(c11, hn), ln));
    return fmax(LUT_K[K_TINYF], c);
}

// NP pixels of rgb_to_husl_nd, default build (_simd.c:128-225), stage by stage.  `xyz(p, x, y, z)` hands over
// pixel p's X, Y, Z (_simd.c:171-173; the two callers below differ in where the table operands come from),
// `light_q` is the pre-divided light table (shared memory in the tile kernel, global / L1 in the
// one-pixel-per-thread kernel).  c[3p .. 3p+2] = R, G, B bytes of pixel p; o[3p .. 3p+2] = H, S, L.
// Walking the NP pixels through each stage together gives the scheduler NP independent dependency chains
// and loads each stage's constants once.
//
// FAST = false is the complete function.  FAST = true (tile kernel) leaves out everything that only three
// kinds of pixel need -- the constants for pure white and pure black (_simd.c:205-212), the u == 0 guard of
// the arctangent (:288-315) and sqrt(0) (u = v = 0 is a special case of u == 0) -- and returns whether any
// of the NP pixels was of those kinds (or had a v / u that is not finite for another reason); the caller then
// redoes that lane's pixels with FAST = false.  On the 2^24 colours with the default tables those are black,
// white and (0, 0, 1), whose L index rounds to 0.  FAST also trims three FP64 operations whose result
// provably (4x: a power of two) or on every one of the 2^24 colours (sha256 of the cube) stays the same.
template <bool INTERP, typename TT, int NP, bool FAST, typename XYZ>
__device__ __forceinline__ bool lut_pixels(const LutParams &t, const XYZ &xyz, const double *light_q,
                                           const uint32_t (&c)[3 * NP], double (&o)[3 * NP]) {
    double y[NP], vu[NP], vv[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        double x, z;
        xyz(p, x, y[p], z);
        const double vs = da(da(x, dm(15.0, y[p])), dm(3.0, z));
        const double rvs = rcp_near(vs);
        if (FAST) {     // RN(4x / vs) = 4 RN(x / vs): the factor is an exponent increment (x = 0 is black, answered elsewhere)
            const double q = cdiv(x, vs, rvs);
            vu[p] = __hiloint2double(__double2hiint(q) + 0x00200000, __double2loint(q));
        } else {
            vu[p] = cdiv(dm(4.0, x), vs, rvs);
        }
        vv[p] = cdiv(dm(9.0, y[p]), vs, rvs);
    }
    double l[NP], u[NP], v[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) l[p] = FAST ? lut_light_seg(t, light_q, y[p]) : lut_light(t, light_q, y[p]);
    {
        const double REF_U = LUT_K[K_REFU], REF_V = LUT_K[K_REFV];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const double l13 = dm(l[p], 13.0);
            u[p] = dm(l13, ds(vu[p], REF_U));
            v[p] = dm(l13, ds(vv[p], REF_V));
        }
    }
    double h[NP], mc[NP];
    bool special = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (FAST) {
            bool znf;
            h[p] = lut_hue_nz(u[p], v[p], znf);
            special = special || znf;
        } else {
            h[p] = lut_hue(u[p], v[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) mc[p] = lut_chroma<INTERP, TT, FAST>(t, l[p], h[p]);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double sat = div_rn(dm(t.sat_scale, sqrt_rn<!FAST>(da(dm(u[p], u[p]), dm(v[p], v[p])))), mc[p]);
        const uint32_t r = c[3 * p], g = c[3 * p + 1], b = c[3 * p + 2];
        if (FAST) {
            // black: r + g + b == 0, white: == 765 (u == 0 was seen by lut_hue_nz: v / u is not finite)
            const uint32_t sum = r + g + b;
            special = special || sum == 0u || sum == 765u;
            // fmin(sat, 100): sat >= 0, so "sat >= 100" is a compare of the high words (100.0 = 0x40590000'00000000)
            const bool cap = (unsigned)__double2hiint(sat) >= 0x40590000u;
            o[3 * p] = h[p];
            o[3 * p + 1] = __hiloint2double(cap ? 0x40590000 : __double2hiint(sat), cap ? 0 : __double2loint(sat));
            o[3 * p + 2] = l[p];
        } else {
            // _simd.c:205-212: pure white and pure black are answered with constants
            const bool white = (r & g & b) == 255u, black = (r | g | b) == 0u;
            o[3 * p] = white ? LUT_K[K_WHITE_HUE] : (black ? 0.0 : h[p]);
            o[3 * p + 1] = (white || black) ? 0.0 : fmin(sat, 100.0);
            o[3 * p + 2] = white ? 100.0 : (black ? 0.0 : l[p]);
        }
    }
    return special;
}

// X, Y, Z of one pixel from its three table values of linear light: the nine products and six sums of
// _simd.c:171-173 in the order written there, (c0 r + c1 g) + c2 b per row
__device__ __forceinline__ void lut_xyz(double lr, double lg, double lb, double &x, double &y, double &z) {
    x = da(da(dm(LUT_K[K_XR], lr), dm(LUT_K[K_XG], lg)), dm(LUT_K[K_XB], lb));
    y = da(da(dm(LUT_K[K_YR], lr), dm(LUT_K[K_YG], lg)), dm(LUT_K[K_YB], lb));
    z = da(da(dm(LUT_K[K_ZR], lr), dm(LUT_K[K_ZG], lg)), dm(LUT_K[K_ZB], lb));
}

// rgb_to_husl_nd, one pixel per thread: tails (< 128 pixels) and unaligned buffers
template <bool INTERP, typename TT>
__global__ void __launch_bounds__(256)
k_fwd_lut(const uint8_t *__restrict__ rgb, double *__restrict__ hsl, long long n, LutParams t) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t c[3] = {rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]};
        double r[3];
        const auto xyz = [&](int, double &x, double &y, double &z) {
            lut_xyz(__ldg(t.linear + c[0]), __ldg(t.linear + c[1]), __ldg(t.linear + c[2]), x, y, z);
        };
        lut_pixels<INTERP, TT, 1, false>(t, xyz, t.light_q, c, r);
        double *o = hsl + 3 * i;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    }
}

// derived table, see LutParams
template <typename TT>
__global__ void k_lut_prepare(const TT *light, int n_light, double light_scale, double *light_q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_light) light_q[i] = __ddiv_rn((double)light[i], light_scale);
}

// ---- device-side table generation (make_lookup_tables:8-9) -------------------------------
// float64 restatement of nphusl.py:268-275 / :165-220 evaluated per table entry.
__device__ __forceinline__ double gen_to_light(double y) {
    if (y > 0.0088564516) return pow(y, 1.0 / 3.0) * 116 - 16;
    return y * 903.2962962;
}

__device__ double gen_max_chroma(double L, double H) {
    const double M[3][3] = {{3.240969941904521, -1.537383177570093, -0.498610760293},
                            {-0.96924363628087, 1.87596750150772, 0.041555057407175},
                            {0.055630079696993, -0.20397695888897, 1.056971514242878}};
    const double hrad = (H / 360.0) * (3.14159265358979323846 * 2);
    const double sub1 = pow(L + 16.0, 3.0) / 1560896.0;
    const double sub2 = (sub1 < 0.0088564516) ? (L / 903.2962962) : sub1;
    const double sn = sin(hrad), cs = cos(hrad);
    double best = INFINITY;
    for (int i = 0; i < 3; i++) {
        const double t1 = 284517.0 * M[i][0] - 94839.0 * M[i][2];
        const double t2 = 838422.0 * M[i][2] + 769860.0 * M[i][1] + 731718.0 * M[i][0];
        const double bt = 632260.0 * M[i][2] - 126452.0 * M[i][1];
        for (int t = 0; t < 2; t++) {
            const double top1 = sub2 * t1;
            double top2 = L * sub2 * t2;
            double bottom = sub2 * bt;
            if (t) { top2 -= L * 769860.0; bottom += 126452.0; }
            double len = (top2 / bottom) / (sn - (top1 / bottom) * cs);
            if (isnan(len) || len < 0) len = INFINITY;
            if (len < best) best = len;
        }
    }
    return best;
}

// compute_linear((float)j / 255.0), _linear_lookup.c:45-61 (before "%f" rounding)
__global__ void k_gen_linear(double *out) {
    const int j = threadIdx.x;
    // (float)j / 255.0 in C promotes to double; j is exact in float, so == j/255.0
    const double x = (double)j / 255.0;
    out[j] = x > 0.04045 ? pow((x + 0.055) / (1.0 + 0.055), 2.4) : x / 12.92;
}

// One table entry as the generated C source holds it (LUT/light.py:86-97, LUT/chroma.py:60-70):
// integer types: int(round(v * int_scale)); float / double: the "{:7.Nf}" literal, i.e. v rounded
// to N decimals (rint(v * 10^N) / 10^N is the double nearest to that decimal), narrowed for float.
template <typename TT>
__device__ __forceinline__ TT lut_entry(double v, double int_scale, double pow10) {
    if (sizeof(TT) == 2) return (TT)(long long)rint(v * int_scale);
    return (TT)(rint(v * pow10) / pow10);
}

// LUT/light.py:62-80, 100-107: entry k of segment i = to_light(start_i + k*step_i)
template <typename TT>
__global__ void k_gen_light(TT *table, int seg, double s0, double s1, double s2,
                            double st0, double st1, double st2, double int_scale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * seg) return;
    const int i = idx / seg, k = idx % seg;
    const double start = i == 0 ? s0 : (i == 1 ? s1 : s2);
    const double step = i == 0 ? st0 : (i == 1 ? st1 : st2);
    table[idx] = lut_entry<TT>(gen_to_light(start + k * step), int_scale, 1e4);
}

// LUT/chroma.py:73-84: entry [h][l] = max_chroma(l*lstep, h*hstep), column l = 0 left 0
template <typename TT>
__global__ void k_gen_chroma(TT *table, int nh, int nl, double hstep, double lstep, double int_scale) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nh * nl) return;
    const int i = (int)(idx / nl), j = (int)(idx % nl);
    TT v = 0;
    if (j > 0) v = lut_entry<TT>(gen_max_chroma(0.0 + j * lstep, 0.0 + i * hstep), int_scale, 1e3);
    table[idx] = v;
}

}  // namespace husl
