/*
 * husl_oracle.c -- CPU restatement of the reference's HUSL hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs may load this library; the product package
 * (husl-numpy_b200/) never does and has no CPU fallback.
 *
 * What is restated (citations are file:line into the reference tree,
 * TadLeonard/husl-numpy 1.5.0):
 *
 *   float64 "reference-of-record" path (NumPy, == Cython to 6e-11):
 *     nphusl/nphusl.py:110-392, nphusl/_cython_opt.pyx:53-309,
 *     nphusl/constants.py:3-21, nphusl/transform.py:16-37
 *   C/SIMD path, exact variant and default LUT variant:
 *     nphusl/_simd.c:87-509, nphusl/_scale_const.c:40-42,
 *     nphusl/_linear_lookup.c:9-75
 *   lookup-table generators:
 *     LUT/light.py:62-151, LUT/chroma.py:73-131
 *
 * Parity pin: oracle/pin_oracle.py runs this file against the reference itself
 * (its NumPy path imported from /root/reference, and oracle/_ref/libsimd_*.so
 * compiled from the reference's own _simd.c) over the full 2^24 colour cube and
 * the reference tests' known-answer literals; tests/golden/ holds vectors made
 * by the reference (tests/golden/make_golden.py) which tests/test_oracle.py
 * re-checks on every run.  Results are recorded in DESIGN.md.
 *
 * Every function is scalar per-pixel code in IEEE double, operations in the
 * reference's order; loops are OpenMP-parallel when compiled with -fopenmp.
 * Build WITHOUT -ffast-math (oracle/Makefile) so the operation order is kept.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- constants: nphusl/constants.py:3-21 ------------------------------- */
static const double M[3][3] = {
    {3.240969941904521, -1.537383177570093, -0.498610760293},
    {-0.96924363628087, 1.87596750150772, 0.041555057407175},
    {0.055630079696993, -0.20397695888897, 1.056971514242878}};
static const double M_INV[3][3] = {
    {0.41239079926595, 0.35758433938387, 0.18048078840183},
    {0.21263900587151, 0.71516867876775, 0.072192315360733},
    {0.019330818715591, 0.11919477979462, 0.95053215224966}};
static const double REF_Y = 1.0;
static const double REF_U = 0.19783000664283;
static const double REF_V = 0.46831999493879;
static const double KAPPA = 903.2962962;
static const double EPSILON = 0.0088564516;
static const double L_MAX = 99.99; /* nphusl/nphusl.py:106 */
static const double L_MIN = 0.01;  /* nphusl/nphusl.py:107 */

/* ======================================================================== */
/* float64 path (NumPy semantics)                                            */
/* ======================================================================== */

/* nphusl.py:278-286 (_to_linear), _cython_opt.pyx:170-174 */
static double to_linear(double c) {
    if (c > 0.04045) return pow((c + 0.055) / (1 + 0.055), 2.4);
    return c / 12.92;
}

/* nphusl.py:330-335 (_from_linear), _cython_opt.pyx:252-256 */
static double from_linear(double v) {
    if (v <= 0.0031308) return 12.92 * v;
    return 1.055 * pow(v, 1 / 2.4) - 0.055;
}

/* nphusl.py:268-275 (_to_light), _cython_opt.pyx:161-165 */
static double to_light(double y) {
    if (y > EPSILON) return pow(y / REF_Y, 1.0 / 3.0) * 116 - 16;
    return (y / REF_Y) * KAPPA;
}

/* nphusl.py:385-392 (_from_light) */
static double from_light(double l) {
    if (l > 8) return REF_Y * pow((l + 16) / 116, 3.0);
    return REF_Y * l / KAPPA;
}

/* nphusl.py:165-220 (_max_lh_chroma, _bounds, _ray_length): smallest positive
 * ray length from the origin, at angle H, to the six gamut boundary lines at
 * lightness L; +inf when none is positive.  _cython_opt.pyx:271-309 is the
 * scalar version of the same thing (init 1e5 instead of inf). */
double oracle_max_chroma(double L, double H) {
    double hrad = (H / 360.0) * (M_PI * 2);
    double sub1 = pow(L + 16.0, 3) / 1560896.0;
    double sub2 = (sub1 < EPSILON) ? (L / KAPPA) : sub1;
    double sn = sin(hrad), cs = cos(hrad);
    double best = INFINITY;
    for (int i = 0; i < 3; i++) {
        double t1 = 284517.0 * M[i][0] - 94839.0 * M[i][2];
        double t2 = 838422.0 * M[i][2] + 769860.0 * M[i][1] + 731718.0 * M[i][0];
        double bt = 632260.0 * M[i][2] - 126452.0 * M[i][1];
        for (int t = 0; t < 2; t++) {
            double top1 = sub2 * t1;
            double top2 = L * sub2 * t2;
            double bottom = sub2 * bt;
            if (t) {
                top2 -= L * 769860.0;
                bottom += 126452.0;
            }
            double m1 = top1 / bottom, b1 = top2 / bottom;
            double len = b1 / (sn - m1 * cs);
            if (isnan(len) || len < 0) len = INFINITY;
            if (len < best) best = len;
        }
    }
    return best;
}

/* One pixel, float RGB in [0,1] -> (L, U, V): nphusl.py:289-296 (_dot_product
 * with M_INV), :246-265 (_xyz_to_luv incl. the inf/NaN scrubbing) */
static void rgb_to_luv(const double rgb[3], double *L, double *U, double *V) {
    double r = to_linear(rgb[0]), g = to_linear(rgb[1]), b = to_linear(rgb[2]);
    double x = M_INV[0][0] * r + M_INV[0][1] * g + M_INV[0][2] * b;
    double y = M_INV[1][0] * r + M_INV[1][1] * g + M_INV[1][2] * b;
    double z = M_INV[2][0] * r + M_INV[2][1] * g + M_INV[2][2] * b;
    double d = x + (15 * y) + (3 * z);
    double uvar = (4 * x) / d, vvar = (9 * y) / d;
    if (isinf(uvar)) uvar = 0;
    if (isinf(vvar)) vvar = 0;
    double l = to_light(y);
    double u = l * 13 * (uvar - REF_U);
    double v = l * 13 * (vvar - REF_V);
    if (l == 0) u = v = 0;              /* luv_flat[L == 0] = 0 */
    if (isnan(u)) u = 0;                /* np.nan_to_num */
    if (isnan(v)) v = 0;
    if (isnan(l)) l = 0;
    *L = l; *U = u; *V = v;
}

/* nphusl.py:223-243 (_luv_to_lch): C and H (degrees, [0,360)) */
static void luv_to_ch(double u, double v, double *C, double *H) {
    if (u == 0.0) u = 0.0;              /* -0.0 scrub of the U channel, :225-226 */
    *C = sqrt(u * u + v * v);
    double h = atan2(v, u) * (180.0 / M_PI);
    if (h < 0.0) h += 360.0;
    *H = h;
}

/* nphusl.py:110-115 + :136-159: float RGB [0,1] -> HUSL */
void oracle_rgb_to_husl_f64(const double *rgb, double *hsl, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double L, U, V, C, H, S;
        rgb_to_luv(rgb + 3 * i, &L, &U, &V);
        luv_to_ch(U, V, &C, &H);
        if (L > L_MAX) { S = 0.0; L = 100.0; }
        else if (L < L_MIN) { S = 0.0; L = 0.0; }
        else S = (C / oracle_max_chroma(L, H)) * 100.0;
        hsl[3 * i] = H; hsl[3 * i + 1] = S; hsl[3 * i + 2] = L;
    }
}

/* transform.py:35-37 (scale_down_rgb) then the float path */
void oracle_rgb_u8_to_husl(const uint8_t *rgb, double *hsl, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double f[3] = {rgb[3 * i] / 255.0, rgb[3 * i + 1] / 255.0, rgb[3 * i + 2] / 255.0};
        oracle_rgb_to_husl_f64(f, hsl + 3 * i, 1);
    }
}

/* nphusl.py:122-128 (_rgb_to_hue), _cython_opt.pyx:114-156 */
void oracle_rgb_to_hue_f64(const double *rgb, double *hue, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double L, U, V, C, H;
        rgb_to_luv(rgb + 3 * i, &L, &U, &V);
        luv_to_ch(U, V, &C, &H);
        hue[i] = H;
    }
}

void oracle_rgb_u8_to_hue(const uint8_t *rgb, double *hue, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double f[3] = {rgb[3 * i] / 255.0, rgb[3 * i + 1] / 255.0, rgb[3 * i + 2] / 255.0};
        oracle_rgb_to_hue_f64(f, hue + i, 1);
    }
}

/* nphusl.py:305-392: HUSL -> float RGB in [0,1] (no clamp) */
void oracle_husl_to_rgb_f64(const double *hsl, double *rgb, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double H = hsl[3 * i], S = hsl[3 * i + 1], L = hsl[3 * i + 2];
        /* _husl_to_lch :338-358 */
        double C = oracle_max_chroma(L, H) / 100.0 * S;
        if (L > L_MAX) { L = 100; C = 0; }
        else if (L < L_MIN) { L = 0; C = 0; }
        /* _lch_to_luv :319-327 */
        double hrad = H * (M_PI / 180.0);
        double U = cos(hrad) * C, V = sin(hrad) * C;
        /* _luv_to_xyz :361-382 */
        double Yv = from_light(L);
        double L13 = 13.0 * L;
        double uvar = U / L13 + REF_U, vvar = V / L13 + REF_V;
        if (isinf(uvar)) uvar = 0;
        if (isinf(vvar)) vvar = 0;
        double Y = Yv * REF_Y;
        double X = -(9 * Y * uvar) / ((uvar - 4.0) * vvar - uvar * vvar);
        double Z = (9.0 * Y - (15.0 * vvar * Y) - (vvar * X)) / (3.0 * vvar);
        if (L == 0) X = Y = Z = 0;
        double xyz[3] = {X, Y, Z};
        for (int k = 0; k < 3; k++) {   /* np.nan_to_num */
            if (isnan(xyz[k])) xyz[k] = 0;
            else if (isinf(xyz[k])) xyz[k] = xyz[k] > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
        }
        /* _xyz_to_rgb: _dot_product(M, xyz) :289-296 then _from_linear */
        for (int k = 0; k < 3; k++)
            rgb[3 * i + k] = from_linear(M[k][0] * xyz[0] + M[k][1] * xyz[1] + M[k][2] * xyz[2]);
    }
}

/* transform.py:16-20, 30-32 (to_rgb_int): round half-to-even of c*255, abs,
 * astype(uint8) -- no clamp, C-cast wrap for values past 255 */
uint8_t oracle_to_rgb_int(double c) {
    double v = fabs(nearbyint(c * 255.0));   /* default rounding mode = half-even */
    if (isnan(v)) return 0;
    if (v >= 9.2e18) return 0;
    return (uint8_t)(long long)v;
}

void oracle_husl_to_rgb_u8(const double *hsl, uint8_t *rgb, size_t n) {
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n; i++) {
        double f[3];
        oracle_husl_to_rgb_f64(hsl + 3 * i, f, 1);
        for (int k = 0; k < 3; k++) rgb[3 * i + k] = oracle_to_rgb_int(f[k]);
    }
}

/* ======================================================================== */
/* lookup tables                                                             */
/* ======================================================================== */

/* nphusl/_linear_lookup.c:45-75: table printed with "%f" (6 decimals) from
 * compute_linear((float)j / 255.0); the shipped array :9-42 is that print-out */
void oracle_linear_table6(double *t256) {
    for (int j = 0; j < 256; j++) {
        double v = to_linear((float)j / 255.0);
        char buf[64];
        snprintf(buf, sizeof buf, "%f", v);
        t256[j] = strtod(buf, NULL);
    }
}

/* LUT/light.py:62-80, 100-151 with `-y 0.070 0.350 -s N` (make_lookup_tables:8):
 * three Y segments [0,.07) [.07,.35) [.35,1+step]; entries round(L*100) as
 * uint16; steps3/thresh2 receive Y_IDX_STEP_{0,1,2} and Y_THRESH_{0,1}.
 * np.arange(start, stop, step)[k] = start + k*step. */
/* One table entry as the generated C source holds it (LUT/light.py:86-97, LUT/chroma.py:60-70):
 * kind 0 = integer types: int(round(v * int_scale)); kind 1 / 2 = `float` / `double` tables: the
 * "{:7.Nf}" literal (N = 4 for light, 3 for chroma), which the C compiler reads as a double and,
 * for a float table, narrows to float. */
static double table_entry(double v, int kind, double int_scale, int decimals) {
    if (kind == 0) return (double)(long)nearbyint(v * int_scale);
    char buf[64];
    snprintf(buf, sizeof buf, "%.*f", decimals, v);
    const double d = strtod(buf, NULL);
    return kind == 1 ? (double)(float)d : d;
}

void oracle_light_table_typed(double *table, int seg, double *steps3, double *thresh2, int kind, double int_scale) {
    double ysteps[3] = {0.070, 0.350, 1.0};
    double start = 0.0;
    for (int i = 0; i < 3; i++) {
        double stop = ysteps[i], step;
        if (i == 2) {
            step = (stop - start) / (seg - 1);
            stop += step;
            step = (stop - start) / seg;
        } else {
            step = (stop - start) / seg;
        }
        steps3[i] = step;
        if (i < 2) {
            char buf[64];               /* "{:0.04f}".format(step) :77 */
            snprintf(buf, sizeof buf, "%0.4f", stop);
            thresh2[i] = strtod(buf, NULL);
        }
        for (int k = 0; k < seg; k++)
            table[i * seg + k] = table_entry(to_light(start + k * step), kind, int_scale, 4);
        start = stop;
    }
}

void oracle_light_table(uint16_t *table, int seg, double *steps3, double *thresh2) {
    double *t = (double *)malloc(sizeof(double) * 3 * (size_t)seg);
    oracle_light_table_typed(t, seg, steps3, thresh2, 0, 100);
    for (int i = 0; i < 3 * seg; i++) table[i] = (uint16_t)t[i];
    free(t);
}

/* LUT/chroma.py:73-84, 63-65: hues = arange(0, 360+hstep, hstep), lights =
 * arange(0, 100+lstep, lstep), entry [h][l] = round(max_chroma(L,H)*100), L=0
 * column left 0 */
void oracle_chroma_table_typed(double *table, int nh, int nl, double *h_step, double *l_step, int kind, double int_scale) {
    double hs = 360.0 / (nh - 1), ls = 100.0 / (nl - 1);
    *h_step = hs; *l_step = ls;
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < nh; i++) {
        table[i * nl] = 0;
        for (int j = 1; j < nl; j++)
            table[i * nl + j] = table_entry(oracle_max_chroma(0.0 + j * ls, 0.0 + i * hs), kind, int_scale, 3);
    }
}

void oracle_chroma_table(uint16_t *table, int nh, int nl, double *h_step, double *l_step) {
    double *t = (double *)malloc(sizeof(double) * (size_t)nh * nl);
    oracle_chroma_table_typed(t, nh, nl, h_step, l_step, 0, 100);
    for (size_t i = 0; i < (size_t)nh * nl; i++) table[i] = (uint16_t)t[i];
    free(t);
}

/* ======================================================================== */
/* C/SIMD path: nphusl/_simd.c                                               */
/* ======================================================================== */

typedef struct {
    const double *linear;       /* 256, 6-decimal (_linear_lookup.c:9-42) */
    const double *light;        /* 3*seg  (light_table_big), the entries of any l_table_t as doubles */
    const double *chroma;       /* nh*nl  (chroma_table), the entries of any c_table_t as doubles */
    double light_scale, chroma_scale;   /* LIGHT_SCALE, CHROMA_SCALE: the int scale, 1 for float tables */
    int seg, nh, nl;
    double ystep[3], ythresh[2], hstep, lstep;
} simd_tables;

static const double WHITE_HUE = 19.916405993809086;   /* _simd.c:79 */
static const double DEG_PER_RAD = 180.0 / M_PI;        /* _simd.c:228 */
static const double SCALE_SUB1[3] = {969398.790856, -279707.331753, -84414.418054};
static const double SCALE_BOTTOM[3] = {-120846.461733, -210946.241904, 694074.104001};
static const double SCALE_SUB2 = 769860.000000;        /* _scale_const.c:40-42 */

/* _simd.c:485-496 (USE_LIGHT_LUT) */
static double simd_light_lut(const simd_tables *t, double y) {
    double idx;
    if (y < t->ythresh[0]) idx = y / t->ystep[0];
    else if (y < t->ythresh[1]) idx = ((y - t->ythresh[0]) / t->ystep[1]) + t->seg;
    else idx = ((y - t->ythresh[1]) / t->ystep[2]) + t->seg * 2;
    /* roundf takes a float: idx is narrowed first */
    const unsigned short r = fmax(0, fmin(3 * t->seg - 1, roundf(idx)));
    return (double)t->light[r] / t->light_scale;
}

/* _simd.c:503-509 (no LUT) */
static double simd_light_exact(double y) {
    if (y > EPSILON) return 116 * cbrt(y / REF_Y) - 16;
    return (y / REF_Y) * KAPPA;
}

/* _simd.c:279-281 */
static double simd_atan_approx(double z) {
    return (M_PI / 4) * z - z * (fabs(z) - 1) * (0.2447 + 0.0663 * fabs(z));
}

/* _simd.c:288-315 */
static double simd_atan2_approx(double y, double x) {
    const double PI_ = 3.141592653589793, PIBY2 = 1.5707963267948966;
    if (x == 0.0) {
        if (y > 0.0) return PIBY2;
        if (y == 0.0) return 0.0;
        return -PIBY2;
    }
    const double z = y / x;
    double at;
    if (fabs(z) < 1.0) {
        at = z / (1.0 + 0.28 * z * z);
        if (x < 0.0) return (y < 0.0) ? at - PI_ : at + PI_;
    } else {
        at = PIBY2 - z / (z * z + 0.28);
        if (y < 0.0) return at - PI_;
    }
    return at;
}

/* _simd.c:246-266 (USE_HUE_ATAN2_APPROX) */
static double simd_hue_approx(double u, double v) {
    const double z = v / u;
    double hue;
    if (fabs(z) < 1.0) {
        hue = simd_atan_approx(z) * DEG_PER_RAD;
        if (u < 0) hue += 180.0;
        else if (v < 0) hue += 360.0;
    } else {
        hue = simd_atan2_approx(v, u) * DEG_PER_RAD;
        if (hue < 0.0) hue += 360.0;
    }
    return hue;
}

/* _simd.c:318-324 (exact variant: float atan2f) */
static double simd_hue_exact(double u, double v) {
    double hue = atan2f(v, u) * DEG_PER_RAD;
    if (hue < 0) hue += 360;
    return hue;
}

/* _simd.c:399-406 (rounded LUT) */
static double simd_chroma_round(const simd_tables *t, double l, double h) {
    const double hsc = h / t->hstep, lsc = l / t->lstep;
    const unsigned short hi = fmax(0.0, fmin(t->nh - 1, roundf(hsc)));
    const unsigned short li = fmax(0.0, fmin(t->nl - 1, roundf(lsc)));
    return (double)fmax(1e-10, t->chroma[(size_t)hi * t->nl + li]);
}

/* _simd.c:360-387 (INTERPOLATE_CHROMA) */
static double simd_chroma_interp(const simd_tables *t, double l, double h) {
    const double hidx = h / t->hstep, lidx = l / t->lstep;
    const unsigned short hf = fmax(0.0, fmin(t->nh - 2, floorf(hidx)));
    const unsigned short lf = fmax(0.0, fmin(t->nl - 2, floorf(lidx)));
    const double c00 = t->chroma[(size_t)hf * t->nl + lf];        /* c_table_t promoted to double by the products below */
    const double c10 = t->chroma[(size_t)(hf + 1) * t->nl + lf];
    const double c01 = t->chroma[(size_t)hf * t->nl + lf + 1];
    const double c11 = t->chroma[(size_t)(hf + 1) * t->nl + lf + 1];
    const double hn = hidx - hf, ln = lidx - lf, hv = 1 - hn, lv = 1 - ln;
    const double c = c00 * hv * lv + c10 * hn * lv + c01 * hv * ln + c11 * hn * ln;
    return fmax(1e-10f, c);
}

/* _simd.c:424-466 (closed form with float sinf/cosf and 6-decimal scalars) */
static double simd_chroma_exact(double l, double h) {
    double sub1 = pow(l + 16.0, 3) / 1560896.0;
    double sub2 = sub1 > EPSILON ? sub1 : l / KAPPA;
    double top2 = SCALE_SUB2 * l * sub2;
    double top2_b = top2 - 769860.0 * l;
    double theta = h / 360.0 * M_PI * 2.0;
    double sn = sinf(theta), cs = cosf(theta);
    double best[3];
    for (int i = 0; i < 3; i++) {
        double top1 = SCALE_SUB1[i] * sub2;
        double bottom = SCALE_BOTTOM[i] * sub2;
        double bottom_b = bottom + 126452.0;
        double mn = 10000.0, len;
        len = (top2 / bottom) / (sn - (top1 / bottom) * cs);
        mn = len > 0 ? len : mn;
        len = (top2_b / bottom_b) / (sn - (top1 / bottom_b) * cs);
        mn = len > 0 ? fmin(len, mn) : mn;
        best[i] = mn;
    }
    return fmin(fmin(best[0], best[1]), best[2]);
}

/* variant: 0 = exact (no -D flags), 1 = default LUT (light+chroma LUT, atan
 * approx), 2 = LUT + INTERPOLATE_CHROMA.  _simd.c:87-225 */
void oracle_simd_rgb_to_husl(const uint8_t *rgb, double *hsl, size_t n_pixels,
                             int variant, const double *linear256,
                             const double *light, int seg, double light_scale,
                             const double *ystep3, const double *ythresh2,
                             const double *chroma, int nh, int nl, double chroma_scale_in,
                             double hstep, double lstep) {
    simd_tables t;
    memset(&t, 0, sizeof t);
    t.linear = linear256; t.light = light; t.chroma = chroma;
    t.seg = seg; t.nh = nh; t.nl = nl; t.hstep = hstep; t.lstep = lstep;
    t.light_scale = light_scale; t.chroma_scale = chroma_scale_in;
    if (variant) {
        memcpy(t.ystep, ystep3, sizeof t.ystep);
        memcpy(t.ythresh, ythresh2, sizeof t.ythresh);
    }
    const double chroma_scale = variant ? chroma_scale_in : 1;     /* CHROMA_SCALE (_simd.c:64 without the LUT) */
    long i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (long)n_pixels; i++) {
        const uint8_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        double *o = hsl + 3 * i;
        if (r == 255 && g == 255 && b == 255) { o[0] = WHITE_HUE; o[1] = 0.0; o[2] = 100.0; continue; }
        if (!r && !g && !b) { o[0] = o[1] = o[2] = 0; continue; }
        /* to_linear_rgb :155-161, to_xyz :168-174 */
        const double rl = t.linear[r], gl = t.linear[g], bl = t.linear[b];
        const double x = 0.412391 * rl + 0.357584 * gl + 0.180481 * bl;
        const double y = 0.212639 * rl + 0.715169 * gl + 0.072192 * bl;
        const double z = 0.019331 * rl + 0.119195 * gl + 0.950532 * bl;
        /* to_luv :178-188 */
        const double var_scale = x + 15 * y + 3 * z;
        const double var_u = 4 * x / var_scale, var_v = 9 * y / var_scale;
        const double l = variant ? simd_light_lut(&t, y) : simd_light_exact(y);
        const double l13 = l * 13;
        const double u = l13 * (var_u - REF_U), v = l13 * (var_v - REF_V);
        /* rgbluv_to_husl_nd :213-221, to_saturation :333-336 */
        const double h = variant ? simd_hue_approx(u, v) : simd_hue_exact(u, v);
        const double mc = variant == 0 ? simd_chroma_exact(l, h)
                        : variant == 1 ? simd_chroma_round(&t, l, h)
                                       : simd_chroma_interp(&t, l, h);
        const double sat = (100 * chroma_scale) * sqrt(u * u + v * v) / mc;
        o[0] = h; o[1] = fmin(sat, 100.0f); o[2] = l;
    }
}

/* timing helper for bench.py's cpu_baseline("port") leg */
int oracle_omp_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
