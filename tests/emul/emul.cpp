// emul.cpp -- DEVELOPMENT/TEST TOOL, NOT PRODUCT.
// Compiles the per-pixel device math of csrc/husl_math.cuh for the HOST (g++),
// with the MUFU approximations replaced by float32 libm calls plus an optional
// pseudo-random perturbation of a few ulp, so the float32 FORMULATION can be
// swept over all 2^24 colours against the float64 oracle on a CPU-only box
// before GPU time is spent (tests/test_emul_math.py).  The real kernels are
// only ever checked on the GPU (tests/test_cuda_parity.py); nothing in the
// product imports or links this file.
#include <math.h>
#include <stdint.h>
#include <string.h>
#define HUSL_HOST_EMUL 1
static int g_noise_ulps = 0;
static inline float emul_noise(float v) {
    if (!g_noise_ulps || v == 0.0f || !isfinite(v)) return v;
    static thread_local uint32_t s = 0x9E3779B9u;
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    int k = (int)(s % (2 * g_noise_ulps + 1)) - g_noise_ulps;
    uint32_t b; memcpy(&b, &v, 4); b += k; memcpy(&v, &b, 4);
    return v;
}
#include "../../husl-numpy_b200/csrc/husl_math.cuh"
#include "../../husl-numpy_b200/csrc/husl_edit.cuh"

using namespace husl;

static void split(double v, float &hi, float &lo) { hi = (float)v; lo = (float)(v - (double)hi); }
static double lin_d(double c) { return c > 0.04045 ? pow((c + 0.055) / 1.055, 2.4) : c / 12.92; }

extern "C" {
void emul_set_noise(int ulps) { g_noise_ulps = ulps; }

// the HUSL-space stages of the fused / edit-on-load kernels on float32 HUSL triplets (csrc/husl_edit.cuh)
void emul_apply_bands(float *hsl, long long n, const nphusl_bands *b, int stripe) {
    const BandsDev a = bands_dev(*b);
    for (long long i = 0; i < n; ++i) {
        if (stripe) apply_bands<true>(a, hsl[3 * i], hsl[3 * i + 1], hsl[3 * i + 2]);
        else apply_bands<false>(a, hsl[3 * i], hsl[3 * i + 1], hsl[3 * i + 2]);
    }
}
void emul_apply_edit(float *hsl, long long n, const nphusl_edit *e) {
    for (long long i = 0; i < n; ++i) apply_edit<true, true, true>(*e, hsl[3 * i], hsl[3 * i + 1], hsl[3 * i + 2]);
}
int emul_bands_last(const nphusl_bands *b) { return (int)bands_dev(*b).last; }

// device linearisation of float channels vs the float64 formula: returns max relative error
double emul_lin_max_rel_err(const double *c, long long n) {
    double worst = 0;
    for (long long i = 0; i < n; ++i) {
        float hi, lo;
        lin_f64(c[i], hi, lo);
        const double want = lin_d(c[i]), got = (double)hi + (double)lo;
        const double e = want != 0 ? fabs(got - want) / fabs(want) : fabs(got);
        if (e > worst) worst = e;
    }
    return worst;
}

void emul_fwd_u8(const uint8_t *rgb, float *out, long long n) {
    float hi[256], lo[256];
    for (int i = 0; i < 256; ++i) split(lin_d(i / 255.0), hi[i], lo[i]);
#pragma omp parallel for
    for (long long i = 0; i < n; ++i) {
        LinPx p;
        p.hr = hi[rgb[3 * i]]; p.lr = lo[rgb[3 * i]];
        p.hg = hi[rgb[3 * i + 1]]; p.lg = lo[rgb[3 * i + 1]];
        p.hb = hi[rgb[3 * i + 2]]; p.lb = lo[rgb[3 * i + 2]];
        husl_from_diff<true>(make_diff(p), out[3 * i], out[3 * i + 1], out[3 * i + 2]);
    }
}

void emul_fwd_f64(const double *rgb, float *out, long long n) {
#pragma omp parallel for
    for (long long i = 0; i < n; ++i) {
        LinPx p;
        lin_f64(rgb[3 * i], p.hr, p.lr);          // the device routine (Newton-polished x^2.4)
        lin_f64(rgb[3 * i + 1], p.hg, p.lg);
        lin_f64(rgb[3 * i + 2], p.hb, p.lb);
        husl_from_diff<false>(make_diff(p), out[3 * i], out[3 * i + 1], out[3 * i + 2]);
    }
}

// float32 channels through the float64-free formulation (diff_from_f32)
void emul_fwd_f32(const float *rgb, float *out, long long n) {
#pragma omp parallel for
    for (long long i = 0; i < n; ++i)
        husl_from_diff<false>(diff_from_f32(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}
// ... and its four intermediate quantities against float64: max relative error of dr, db, 1 - g, g
void emul_diff_f32_err(const float *rgb, long long n, double *worst) {
    double w[4] = {0, 0, 0, 0};
    for (long long i = 0; i < n; ++i) {
        const Diff d = diff_from_f32(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        const double lr = lin_d(rgb[3 * i]), lg = lin_d(rgb[3 * i + 1]), lb = lin_d(rgb[3 * i + 2]);
        const double want[4] = {lr - lg, lb - lg, 1.0 - lg, lg};
        const double got[4] = {d.dr, d.db, d.cg, d.g};
        for (int k = 0; k < 4; ++k) {
            const double e = want[k] != 0 ? fabs(got[k] - want[k]) / fabs(want[k]) : fabs(got[k]);
            if (e > w[k]) w[k] = e;
        }
    }
    for (int k = 0; k < 4; ++k) worst[k] = w[k];
}

// HUSL float32 -> companded RGB scaled to 0..255 (unrounded) and the uint8 bytes
void emul_inv(const float *hsl, float *rgb255, uint8_t *u8, long long n) {
#pragma omp parallel for
    for (long long i = 0; i < n; ++i) {
        float R, G, B;
        if (rgb255) {      // float RGB output: exact end points (EXACT_EXT), as k_inv_any<T, float/double>
            rgb255_from_husl<true>(hsl[3 * i], hsl[3 * i + 1], hsl[3 * i + 2], R, G, B);
            rgb255[3 * i] = R; rgb255[3 * i + 1] = G; rgb255[3 * i + 2] = B;
        }
        if (u8) {          // uint8 output: the streaming kernels' variant
            rgb255_from_husl<false>(hsl[3 * i], hsl[3 * i + 1], hsl[3 * i + 2], R, G, B);
            u8[3 * i] = (uint8_t)(round_u8_bits(R) & 0xFF);
            u8[3 * i + 1] = (uint8_t)(round_u8_bits(G) & 0xFF);
            u8[3 * i + 2] = (uint8_t)(round_u8_bits(B) & 0xFF);
        }
    }
}
}
