// husl_math.cuh -- per-pixel HUSL <-> RGB arithmetic for sm_100a.
//
// What is computed is the reference's float64 chain (nphusl/nphusl.py:110-392,
// nphusl/_cython_opt.pyx:53-309); HOW is different:
//
//  * forward works on linear-light DIFFERENCES dr = r-g, db = b-g and on
//    1-g, formed from hi+lo float pairs, so the near-grey hue and near-white
//    saturation cancellations of U = 13L(u'-REF_U) never happen in float32;
//  * hue is atan2 of two linear forms (no division by D, no L);
//  * saturation is the six-line ray intersection in closed form:
//        S/100 = max( (D - sD*min(r,g,b)) / D,
//                     1 - sD*Y*(1-max(r,g,b)) / (D*(1-Y)) )
//    -- no trigonometry, one reciprocal;
//  * inverse intersects the hue ray with the same six lines in (u',v') and
//    evaluates lin_i = Y*A_i(u',v')/(4v') with ONE reciprocal, then the sRGB
//    companding with MUFU lg2/ex2 and a magic-number round-half-even to uint8.
//
// All constants come from husl_coeffs.h (float64-derived, gen_coeffs.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "husl_coeffs.h"

namespace husl {

// ---- MUFU wrappers -----------------------------------------------------------
#ifdef HUSL_HOST_EMUL
// host build of this header for tests/emul (development tool, never shipped):
// float32 libm stands in for the MUFU unit, emul_noise() perturbs by a few ulp
__device__ __forceinline__ float mufu_rcp(float x) { return emul_noise(1.0f / x); }
__device__ __forceinline__ float mufu_lg2(float x) { return emul_noise(log2f(x)); }
__device__ __forceinline__ float mufu_ex2(float x) { return emul_noise(exp2f(x)); }
__device__ __forceinline__ float mufu_sin(float x) { return emul_noise(sinf(x)); }
__device__ __forceinline__ float mufu_cos(float x) { return emul_noise(cosf(x)); }
__device__ __forceinline__ uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
__device__ __forceinline__ float __saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }   // NaN -> 0 like the .SAT modifier
// round-to-nearest single operations the compiler may not contract (the host build runs with -ffp-contract=off)
__device__ __forceinline__ float __fmul_rn(float a, float b) { return a * b; }
__device__ __forceinline__ float __fadd_rn(float a, float b) { return a + b; }
__device__ __forceinline__ float __fsub_rn(float a, float b) { return a - b; }
#else
__device__ __forceinline__ float mufu_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_sin(float x) {
    float r;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_cos(float x) {
    float r;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif

// ---- hue: atan2 in degrees, [0, 360] -------------------------------------------
// reference: H = degrees(arctan2(V, U)), +360 if negative (nphusl.py:231-239,
// _cython_opt.pyx:75-79).  U, V share the positive factor 13L/D, so the hue is
// atan2(Nv, Nu).  First-quadrant angle of (|Nu|, |Nv|) without the usual
// min/max swap: atan(y/x) = 45 + atan((y-x)/(y+x)), (y-x)/(y+x) in [-1, 1] where
// the odd minimax polynomial is valid on both signs; then two reflections.
// Callers never pass (0, 0): Nu, Nv carry the GREY_U / GREY_V offsets.
__device__ __forceinline__ float hue_deg(float nv, float nu) {
    const float ax = fabsf(nu), ay = fabsf(nv);
    const float q = (ay - ax) * mufu_rcp(ay + ax);
    const float s = q * q;
    static_assert(ATAN_TERMS == 7, "Horner chain below is written for 7 terms");
    float p = fmaf(ATAN_C6, s, ATAN_C5);
    p = fmaf(p, s, ATAN_C4);
    p = fmaf(p, s, ATAN_C3);
    p = fmaf(p, s, ATAN_C2);
    p = fmaf(p, s, ATAN_C1);
    p = fmaf(p, s, ATAN_C0);
    float r = fmaf(p, q, 45.0f);           // [0, 90]
    r = (nu < 0.0f) ? 180.0f - r : r;      // [0, 180]
    r = (nv < 0.0f) ? 360.0f - r : r;      // [0, 360]
    return r;
}

// ---- sRGB linearisation of a FLOAT channel ------------------------------------------
// reference: c/12.92 below 0.04045, ((c+0.055)/1.055)^2.4 above (nphusl.py:278-286).
// Needed to ~1e-11 relative, not float32's 6e-8: hue and saturation are built
// from DIFFERENCES of linear channels, so the result is returned as a hi+lo
// float pair like the uint8 table entries.  x^2.4 = (x * r^2)^4 with r = x^(-1/5):
// r is seeded by MUFU lg2/ex2 (1e-6 relative) and polished by ONE division-free
// Newton step in float64, r <- r*(6 - x*r^5)/5 (error ~3e-12); 11 DP operations
// instead of the ~120 instructions of a general pow().
// The XU pipe (MUFU and every float <-> double conversion, 16 lanes per SM) is what bounds the float-input
// kernels -- ncu: 81 % busy in k_fwd_flt<float, float> with 8 XU operations per channel -- so conversions are
// spent sparingly: the MUFU seed takes a float argument the caller already has (`xf`), and the result is split
// into hi + lo by Veltkamp's trick in FP64 (t = v (2^29 + 1); h = t - (t - v) is v rounded to 24 bits) so that
// only the two final narrowings convert: 4 conversions + 2 MUFU per channel.
__device__ __forceinline__ void split_f64(double v, float &hi, float &lo) {
    const double t = v * 536870913.0;
    const double h = t - (t - v);
    hi = (float)h;
    lo = (float)(v - h);
}

// xf: (c + 0.055) / 1.055 to float accuracy (only seeds the Newton step)
__device__ __forceinline__ void lin_f64_seeded(double c, float xf, float &hi, float &lo) {
    const double x = fma(c, 1.0 / 1.055, 0.055 / 1.055);
    double r = (double)mufu_ex2(mufu_lg2(xf) * -0.2f);
    const double r2 = r * r;
    const double r5 = r2 * r2 * r;
    r = r * fma(-x, r5, 6.0) * 0.2;
    double t = x * (r * r);
    t = t * t;
    t = t * t;
    split_f64((c > 0.04045) ? t : c * (1.0 / 12.92), hi, lo);
}

__device__ __forceinline__ void lin_f64(double c, float &hi, float &lo) {
    lin_f64_seeded(c, (float)fma(c, 1.0 / 1.055, 0.055 / 1.055), hi, lo);
}
// a float32 channel: the seed argument is formed in float, no narrowing conversion at all
__device__ __forceinline__ void lin_f64(float c, float &hi, float &lo) {
    lin_f64_seeded((double)c, fmaf(c, 1.0f / 1.055f, 0.055f / 1.055f), hi, lo);
}

// ---- forward core --------------------------------------------------------------
// Linear-light channel values arrive as hi+lo float pairs (exact to ~2^-48):
// from the 256-entry table for uint8 input, from a float64 pow for float input.
struct LinPx {
    float hr, lr, hg, lg, hb, lb;
};

struct Diff {       // dr = r-g, db = b-g, cg = 1-g, g
    float dr, db, cg, g;
};

__device__ __forceinline__ Diff make_diff(const LinPx &p) {
    Diff d;
    d.dr = (p.hr - p.hg) + (p.lr - p.lg);
    d.db = (p.hb - p.hg) + (p.lb - p.lg);
    d.cg = (1.0f - p.hg) - p.lg;
    d.g = p.hg;
    return d;
}

// ---- float32 channels: the same four quantities without float64 ------------------------------------
// The float64 linearisation above exists because hue and saturation are built from DIFFERENCES of linear
// channels: two float32-accurate powers of nearly equal channels cancel.  For FLOAT32 input the differences
// can be formed without cancellation instead.  With T the segment switch, P(c) = ((c + 0.055)/1.055)^2.4 and
// J = P(T) - T/12.92 (2.3e-9: the two sRGB segments do not quite meet),
//     lin(c) = min(c, T)/12.92 + [P(max(c, T)) - P(T)] + J [c > T],
// so lin(r) - lin(g) = (min(r,T) - min(g,T))/12.92 + [P(rb) - P(gb)] + J ([r > T] - [g > T]),  cb = max(c, T):
// three terms of one sign.  The first is an exact float difference times a constant.  The middle one is
// P(gb) ((1 + t)^2.4 - 1) with t = (rb - gb)/(gb + 0.055), again from an exact difference: a degree-4
// polynomial in t for |t| < 1/4 (relative error 2e-7, gen_coeffs.powg_coeffs), and for larger |t| the plain
// difference of the two MUFU powers, whose relative error 7e-7 (P_r + P_g)/|P_r - P_g| stays below 3e-6
// there.  1 - lin(g) is the same difference with r = 1.  A relative error e of (dr, db) moves the hue by at
// most cond(hue forms) e = 1.5 e rad = 86 e degrees: 3e-6 -> 2.6e-4 degrees against the bar of 1e-3
// (tests/test_emul_math.py sweeps near-grey, near-switch, near-0 and near-1 pixels against the oracle).
// 7 XU operations (6 MUFU lg2/ex2, one reciprocal) and no FP64 per pixel, against 18 XU + 45 FP64 for three
// lin_f64 calls: the float32-input kernels leave the XU pipe (ncu: 81 % busy) for plain FP32 issue.
// NaN channels propagate (the selects below, unlike fmaxf, keep a NaN); the reference returns NaN there too.
__device__ __forceinline__ float srgb_pow(float cb) {           // P(cb), cb >= T
    return mufu_ex2(mufu_lg2(fmaf(cb, 1.0f / 1.055f, 0.055f / 1.055f)) * 2.4f);
}
__device__ __forceinline__ float pow_gain(float t) {            // ((1 + t)^2.4 - 1) / t, |t| < POWG_TAU
    static_assert(POWG_DEG == 4, "Horner chain below is written for degree 4");
    float p = fmaf(POWG_C4, t, POWG_C3);
    p = fmaf(p, t, POWG_C2);
    p = fmaf(p, t, POWG_C1);
    return fmaf(p, t, POWG_C0);
}
// lin(c) - lin(g) given c's and g's clamped parts; pc = P(cb), pg = P(gb), inv = 1/(gb + 0.055)
__device__ __forceinline__ float lin_diff(float cb, float cl, float pc, float jc, float gb, float gl, float pg, float jg, float inv) {
    const float t = (cb - gb) * inv;
    const float pw = (fabsf(t) < POWG_TAU) ? (pg * t) * pow_gain(t) : pc - pg;
    return fmaf(cl - gl, 1.0f / 12.92f, pw) + (jc - jg);
}
__device__ __forceinline__ Diff diff_from_f32(float r, float g, float b) {
    const bool lr = r <= SRGB_T, lg = g <= SRGB_T, lb = b <= SRGB_T;      // false for NaN: the NaN stays in cb
    const float rb = lr ? SRGB_T : r, gb = lg ? SRGB_T : g, bb = lb ? SRGB_T : b;
    const float rl = fminf(r, SRGB_T), gl = fminf(g, SRGB_T), bl = fminf(b, SRGB_T);
    const float jr = lr ? 0.0f : SRGB_J, jg = lg ? 0.0f : SRGB_J, jb = lb ? 0.0f : SRGB_J;
    const float pr = srgb_pow(rb), pg = srgb_pow(gb), pb = srgb_pow(bb);
    const float inv = mufu_rcp(gb + 0.055f);
    Diff d;
    d.dr = lin_diff(rb, rl, pr, jr, gb, gl, pg, jg, inv);
    d.db = lin_diff(bb, bl, pb, jb, gb, gl, pg, jg, inv);
    d.cg = lin_diff(1.0f, SRGB_T, 1.0f, SRGB_J, gb, gl, pg, jg, inv);
    d.g = lg ? g * (1.0f / 12.92f) : pg;
    return d;
}

// 4X - REF_U*D and 9Y - REF_V*D as linear forms of the differences.  Their
// g-coefficients vanish (white point), so exact greys would give atan2(0, 0);
// the float64 reference evaluates atan2 of its own rounding residue there and
// lands at 17.6 .. 21.1 degrees (19.916 for white, 0 for black).  The constant
// offsets 1e-13*(cos, sin)(WHITE_HUE) -- nine orders of magnitude below the
// smallest non-grey |Nu|, |Nv| -- give every grey the hue of white with no
// select; black is handled by the callers (H = 0, like the reference).
__device__ __forceinline__ void hue_forms(const Diff &d, float &nu, float &nv) {
    nu = fmaf(U_DR, d.dr, fmaf(U_DB, d.db, GREY_U));
    nv = fmaf(V_DR, d.dr, fmaf(V_DB, d.db, GREY_V));
}

// IS_U8: with uint8 input the lightness clamps (L > 99.99 -> S = 0, L = 100;
// L < 0.01 -> S = 0, L = 0; nphusl.py:146-152) can only fire for pure white and
// pure black (nearest neighbours (255,255,254) -> L = 99.975, (0,0,1) -> L =
// 0.0198).  For those two colours D*(1-Y) is exactly 0, both saturation
// candidates evaluate to 0*inf = NaN and the three-input max with 0 (FMNMX3
// returns the non-NaN operand) yields S = 0; L = 116*ex2(lg2(1)/3) - 16 = 100 and
// KAPPA*0 = 0 exactly.  So the uint8 kernels need no clamp instructions at all.
// Float inputs keep the explicit clamps, expressed on Y.
template <bool IS_U8>
__device__ __forceinline__ void husl_from_diff(const Diff &d, float &H, float &S, float &L) {
    static_assert(Y_S == 1.0f, "Y = g + t and 1 - Y = (1-g) - t below rely on the Y row summing to 1 in float32");
    const float t = fmaf(Y_DR, d.dr, Y_DB * d.db);
    const float Y = d.g + t;                               // luminance
    const float omY = d.cg - t;                            // 1 - Y, cancellation-free
    const float Dp = fmaf(D_DR, d.dr, D_DB * d.db);
    const float D = fmaf(D_S, d.g, Dp);                    // X + 15Y + 3Z
    float nu, nv;
    hue_forms(d, nu, nv);
    float h = hue_deg(nv, nu);
    h = (D > 0.0f) ? h : 0.0f;                             // black

    // saturation: channel-hits-0 and channel-hits-1 gamut faces
    const float mnd = fminf(fminf(d.dr, d.db), 0.0f);      // min(r,g,b) - g
    const float mxd = fmaxf(fmaxf(d.dr, d.db), 0.0f);      // max(r,g,b) - g
    const float T0 = fmaf(-D_S, mnd, Dp);                  // D - sD*min(r,g,b) >= 0
    const float omx = d.cg - mxd;                          // 1 - max(r,g,b)  >= 0
    const float R100 = 100.0f * mufu_rcp(D * omY);
    const float s0 = (T0 * omY) * R100;
    const float s1 = fmaf((-D_S * Y) * omx, R100, 100.0f);
    float s = fmaxf(fmaxf(s0, s1), 0.0f);                  // one FMNMX3; NaN (black, white) -> 0

    // lightness: 116*cbrt(Y)-16 above EPSILON, KAPPA*Y below (nphusl.py:268-275)
    const float cb = mufu_ex2(mufu_lg2(Y) * (1.0f / 3.0f));
    float l = (Y > EPSILON) ? fmaf(116.0f, cb, -16.0f) : KAPPA * Y;

    if (!IS_U8) {
        // L > 99.99 -> (S,L) = (0,100); L < 0.01 -> (0,0), expressed on Y so the
        // clamp does not depend on the cbrt approximation
        const bool white = Y > Y_HI;
        const bool black = Y < Y_LO;
        s = (white || black) ? 0.0f : s;
        l = white ? 100.0f : l;
        l = black ? 0.0f : l;
    }
    H = h; S = s; L = l;
}

__device__ __forceinline__ float hue_from_diff(const Diff &d) {
    float nu, nv;
    hue_forms(d, nu, nv);
    const float h = hue_deg(nv, nu);
    return (fmaxf(fmaxf(d.dr, d.db), d.g) > 0.0f) ? h : 0.0f;   // black <=> r = g = b = 0
}

// ---- inverse core --------------------------------------------------------------
// sRGB companding of one linear channel, scaled to 0..255 (nphusl.py:330-335)
__device__ __forceinline__ float from_linear_255(float v) {
    const float pw = mufu_ex2(mufu_lg2(v) * (1.0f / 2.4f));
    return (v <= 0.0031308f) ? (12.92f * 255.0f) * v : fmaf(1.055f * 255.0f, pw, -0.055f * 255.0f);
}

// (H,S,L) -> companded RGB scaled to 0..255 (not rounded).
// EXACT_EXT = false (uint8 output): the lightness clamps (L > 99.99 -> white,
// L < 0.01 -> black, nphusl.py:351-357) cost one select: S is zeroed there, Y is
// saturated to [0, 1] by the .SAT modifier of the multiply that produces it, and
// what is left of Y between the clamp threshold and the end point (>= 0.99974 /
// <= 1.1e-5) rounds to 255 / 0 in every channel.  EXACT_EXT = true (float RGB
// output) returns exactly 1 / 0 there like the reference.
// WIDE selects the long-hand formulation below instead (explicit selects for the
// end points, factored 1-Y, unfolded constants: 100 instead of 92 instructions per
// pixel in k_inv_tma, same values to float32 rounding).  It is a scheduling choice
// for float64 HUSL input: that kernel is HBM-bound at 16 warps/SM and, inside the
// to_husl -> to_rgb pipeline where it starts on the producer's dirty L2 lines, the
// step takes 0.594 ms with the long form against 0.600 - 0.608 ms with every
// shorter form tried, although the short forms are faster stand-alone (0.280 vs
// 0.284 ms).  The float32 kernels (48 warps/SM, issue-bound) gain 7 % from the
// short form (profiles/r01_ab_variants.txt, runs 22-27).
// Round 2: ONE formulation for every HUSL input type (HUSL_WIDE_F64 = 0): to_rgb of a float64 array, of the
// same values as float32, with or without an edit applied on load, and the fused uint8 kernels all produce
// the same bits.  The long-hand form stays selectable as an A/B knob (it was worth 1 % of the float64
// to_husl -> to_rgb pipeline and cost the bit-for-bit equality of those paths).
#ifndef HUSL_WIDE_F64
#define HUSL_WIDE_F64 0
#endif
template <typename T> __host__ __device__ constexpr bool wide_in() { return HUSL_WIDE_F64 && sizeof(T) == 8; }

template <bool EXACT_EXT = false, bool WIDE = false>
__device__ __forceinline__ void rgb255_from_husl(float H, float S, float L,
                                                 float &R, float &G, float &B) {
    const float th = H * 0.017453292519943295f;
    const float sn = mufu_sin(th), cs = mufu_cos(th);
    const float w = fmaf(L, 1.0f / 116.0f, 16.0f / 116.0f);
    if (WIDE) {
        // Y and 1-Y from L (nphusl.py:385-392); 1-Y = (1-w)(1+w+w^2), w = (L+16)/116
        const float omw = fmaf(L, -1.0f / 116.0f, 100.0f / 116.0f);
        const bool big = L > 8.0f;
        float Y = big ? w * w * w : L * INV_KAPPA;
        const float omY = big ? omw * fmaf(w, w, 1.0f + w) : 1.0f - Y;
        const float g0 = fmaf(AU0, cs, AV0 * sn);
        const float g1 = fmaf(AU1, cs, AV1 * sn);
        const float g2 = fmaf(AU2, cs, AV2 * sn);
        const float gmin = fminf(fminf(g0, g1), g2);
        const float gmax = fmaxf(fmaxf(g0, g1), g2);
        // 1/tau_max = max(-gmin/AN, (Y*gmax - 4 sin)/(4 REF_V (1-Y)))  (six lines)
        const float c = REF_V4 * omY;
        float den = fmaxf((gmin * (-1.0f / AN)) * c, fmaf(Y, gmax, -4.0f * sn));
        float N = (S * 0.01f) * c;                              // tau = N / den
        // L > 99.99 -> white, L < 0.01 -> black (nphusl.py:351-357)
        const bool white = L > 99.99f;
        const bool black = L < 0.01f;
        const bool ext = white || black;
        N = ext ? 0.0f : N;
        den = ext ? 1.0f : den;
        Y = white ? 1.0f : Y;
        Y = black ? 0.0f : Y;
        // lin_i = Y*(AN*den + N*g_i) / (4*(REF_V*den + N*sin))
        const float yr = Y * mufu_rcp(fmaf(4.0f * N, sn, REF_V4 * den));
        const float base = AN * den;
        R = from_linear_255(fmaf(N, g0, base) * yr);
        G = from_linear_255(fmaf(N, g1, base) * yr);
        B = from_linear_255(fmaf(N, g2, base) * yr);
        return;
    }
    // Y from L (nphusl.py:385-392); near white 1-Y only scales a chroma that vanishes with it
    float Y = (L > 8.0f) ? __saturatef(w * w * w) : __saturatef(L * INV_KAPPA);
    // L outside [0.01, 99.99] (or NaN): one compare on |L - 50| for the uint8 path (the +-2e-6
    // it blurs the lower threshold by is far inside the band that rounds to black anyway)
    const bool ext = EXACT_EXT ? !(L >= 0.01f && L <= 99.99f) : !(fabsf(L - 50.0f) <= (99.99f - 50.0f));
    if (EXACT_EXT) Y = (L > 99.99f) ? 1.0f : ((L < 0.01f) ? 0.0f : Y);
    const float omY = 1.0f - Y;
    // directional derivatives (a quarter of them) of the three channel forms along the hue ray
    const float g0 = fmaf(0.25f * AU0, cs, (0.25f * AV0) * sn);
    const float g1 = fmaf(0.25f * AU1, cs, (0.25f * AV1) * sn);
    const float g2 = fmaf(0.25f * AU2, cs, (0.25f * AV2) * sn);
    const float gmin = fminf(fminf(g0, g1), g2);
    const float gmax = fmaxf(fmaxf(g0, g1), g2);
    // 1/(4 tau_max) = max(-gmin/AN, (Y*gmax - sin)/(4 REF_V (1-Y)))  (six lines), both sides times
    // 4 REF_V (1-Y) with AN == 4 REF_V: den = max(-gmin*(1-Y), Y*gmax - sin); > 0 inside the gamut,
    // the tiny floor keeps 0/0 out when L is at or beyond an end point
    static_assert(AN == REF_V4, "the inverse folds AN and 4*REF_V into one constant");
    const float den = fmaxf(fmaxf(-gmin * omY, fmaf(Y, gmax, -sn)), 1e-30f);
    const float N = ((ext ? 0.0f : S) * (0.01f * REF_V4)) * omY;   // tau = N / (4 den)
    // lin_i = Y*(AN*den + N*g_i) / (4 REF_V*den + N*sin)   (AN == 4 REF_V)
    const float base = AN * den;
    const float yr = Y * mufu_rcp(fmaf(N, sn, base));
    R = from_linear_255(fmaf(N, g0, base) * yr);
    G = from_linear_255(fmaf(N, g1, base) * yr);
    B = from_linear_255(fmaf(N, g2, base) * yr);
}

// transform.to_rgb_int (transform.py:16-20): round-half-even, abs, uint8 wrap.
// |v| + 1.5*2^23 leaves round-half-even(|v|) in the low mantissa bits.
__device__ __forceinline__ uint32_t round_u8_bits(float v255) {
    return __float_as_uint(fabsf(v255) + 12582912.0f);
}

}  // namespace husl
