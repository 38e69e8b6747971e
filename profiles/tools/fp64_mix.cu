// Does an FP64 instruction cost the scheduler one issue slot or two?  Independent DFMA and FFMA/IMAD chains per thread,
// enough warps to hide latency: instructions per cycle per scheduler for pure DFMA, pure FFMA and 1:1 / 1:2 mixes.
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NF>
__global__ void mix(double *out, long long *cyc, double a, int n) {
    double x[4]; float f[8];
    for (int k = 0; k < 4; ++k) x[k] = a + k + threadIdx.x * 1e-9;
    for (int k = 0; k < 8; ++k) f[k] = (float)a + k;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < ND; ++k) x[k] = __fma_rn(x[k], 1.0000001, 1e-9);
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k] = __fmaf_rn(f[k], 1.0000001f, 1e-9f);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < 4; ++k) s += x[k];
    for (int k = 0; k < 8; ++k) s += f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ND, int NF>
void run(double *out, long long *cyc, const char *name) {
    const int n = 2048, warps = 32;
    long long h;
    for (int rep = 0; rep < 2; ++rep) { mix<ND, NF><<<148, 32 * warps>>>(out, cyc, 1.0000001, n); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst = (double)(ND + NF) * 8 * n * (warps / 4);     // warp instructions per scheduler
    printf("%-28s %d DFMA + %d FFMA per group: %.3f instr/cycle/scheduler  (DFMA %.3f, FFMA %.3f)\n", name, ND, NF, inst / h,
           inst / h * ND / (ND + NF), inst / h * NF / (ND + NF));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    run<4, 0>(out, cyc, "pure DFMA");
    run<0, 8>(out, cyc, "pure FFMA");
    run<4, 4>(out, cyc, "1:1");
    run<2, 4>(out, cyc, "1:2");
    run<2, 6>(out, cyc, "1:3");
    run<1, 4>(out, cyc, "1:4");
    return 0;
}
