/* A plain C client of include/nphusl_cuda.h -- what a non-Python host (the reference's
 * _simd_opt.pyx wrapper is compiled C, nphusl/_simd_opt.pyx:16-33) links against.
 * No CUDA headers, no Python: pointers and sizes only.
 *
 *   cabi_client            -> link check: prints the ABI version and the device count, exit 0
 *   cabi_client roundtrip  -> needs a GPU: host uint8 RGB -> HUSL (float64, the reference's
 *                             rgb_to_husl_nd result type) -> uint8 RGB, exact round trip over a
 *                             ramp of colours; black / white / red known answers; exit 0 on success
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nphusl_cuda.h"

static int fail(const char *what) {
    fprintf(stderr, "cabi_client: %s: %s\n", what, nphusl_cuda_last_error());
    return 1;
}

int main(int argc, char **argv) {
    printf("abi %d devices %d\n", nphusl_cuda_abi_version(), nphusl_cuda_device_count());
    if (argc < 2 || strcmp(argv[1], "roundtrip") != 0) return 0;
    if (nphusl_cuda_device_count() < 1) return fail("no device");

    const size_t n = 1 << 16;
    unsigned char *rgb = malloc(3 * n), *back = malloc(3 * n);
    double *hsl = malloc(3 * n * sizeof(double));      /* caller-owned, unlike rgb_to_husl_nd's malloc'd result */
    for (size_t i = 0; i < n; ++i) {
        rgb[3 * i] = (unsigned char)(i & 0xFF);
        rgb[3 * i + 1] = (unsigned char)((i >> 8) & 0xFF);
        rgb[3 * i + 2] = (unsigned char)((i * 7) & 0xFF);
    }
    rgb[0] = rgb[1] = rgb[2] = 0;                      /* black */
    rgb[3] = rgb[4] = rgb[5] = 255;                    /* white */
    rgb[6] = 255; rgb[7] = 0; rgb[8] = 0;              /* red   */
    if (nphusl_cuda_rgb_to_husl(rgb, NPHUSL_U8, NPHUSL_HOST, 3, hsl, NPHUSL_F64, NPHUSL_HOST, n, NPHUSL_EXACT, 0, NULL))
        return fail("rgb_to_husl");
    if (nphusl_cuda_husl_to_rgb(hsl, NPHUSL_F64, NPHUSL_HOST, back, NPHUSL_U8, NPHUSL_HOST, n, 0, NULL))
        return fail("husl_to_rgb");
    if (memcmp(rgb, back, 3 * n) != 0) { fprintf(stderr, "cabi_client: round trip differs\n"); return 1; }
    /* known answers of the reference (tests/test.py: black -> (0,0,0), white -> L 100 S 0; red 12.177 / 100 / 53.237) */
    if (hsl[0] != 0.0 || hsl[1] != 0.0 || hsl[2] != 0.0) { fprintf(stderr, "black -> %g %g %g\n", hsl[0], hsl[1], hsl[2]); return 1; }
    if (hsl[4] != 0.0 || fabs(hsl[5] - 100.0) > 1e-3) { fprintf(stderr, "white -> %g %g %g\n", hsl[3], hsl[4], hsl[5]); return 1; }
    if (fabs(hsl[6] - 12.17705) > 1e-3 || fabs(hsl[7] - 100.0) > 1e-3 || fabs(hsl[8] - 53.23712) > 1e-3) {
        fprintf(stderr, "red -> %g %g %g\n", hsl[6], hsl[7], hsl[8]);
        return 1;
    }
    /* device buffers through the same calls */
    void *d_rgb = NULL, *d_hue = NULL;
    float *hue = malloc(n * sizeof(float));
    if (nphusl_cuda_malloc(&d_rgb, 3 * n, 0) || nphusl_cuda_malloc(&d_hue, n * sizeof(float), 0)) return fail("malloc");
    if (nphusl_cuda_memcpy(d_rgb, rgb, 3 * n, 1, 0, NULL)) return fail("memcpy h2d");
    if (nphusl_cuda_rgb_to_hue(d_rgb, NPHUSL_U8, NPHUSL_DEVICE, 3, d_hue, NPHUSL_F32, NPHUSL_DEVICE, n, 0, NULL)) return fail("rgb_to_hue");
    if (nphusl_cuda_memcpy(hue, d_hue, n * sizeof(float), 2, 0, NULL)) return fail("memcpy d2h");
    if (nphusl_cuda_device_synchronize(0)) return fail("synchronize");
    for (size_t i = 0; i < n; ++i)
        if (fabs((double)hue[i] - hsl[3 * i]) > 1e-3 && fabs(fabs((double)hue[i] - hsl[3 * i]) - 360.0) > 1e-3 && hsl[3 * i + 1] > 0.1) {
            fprintf(stderr, "hue[%zu] = %g vs %g\n", i, hue[i], hsl[3 * i]);
            return 1;
        }
    nphusl_cuda_free(d_rgb, 0); nphusl_cuda_free(d_hue, 0);
    printf("roundtrip ok: %zu px, %llu kernel launches\n", n, nphusl_cuda_launch_count());
    free(rgb); free(back); free(hsl); free(hue);
    return 0;
}
