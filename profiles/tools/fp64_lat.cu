// Dependent-issue latency and throughput of the FP64 pipe and of the conversions the LUT / float kernels lean on
// (B200, sm_100a).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void lat(double *out, long long *cyc, double a, double b, int n) {
    double x = a + threadIdx.x * 1e-9, y = b;
    float f = (float)a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (OP == 0) x = __fma_rn(x, y, y);                              // DFMA chain
            if (OP == 1) x = __dadd_rn(x, y);                                // DADD chain
            if (OP == 2) { f = (float)x; x = (double)f + y; }               // F2F.F32.F64 + F2F.F64.F32 + DADD
            if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y; }   // MUFU.RCP64H + DADD
            if (OP == 4) f = __fmaf_rn(f, 1.0000001f, 1e-9f);                // FFMA chain (reference point)
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + f;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const int n = 4096;
    const char *names[5] = {"DFMA", "DADD", "F2F.F32.F64 + F2F.F64.F32 + DADD", "MUFU.RCP64H + DADD", "FFMA"};
    for (int op = 0; op < 5; ++op) {
        for (int warps = 1; warps <= 16; warps *= 2) {
            // one CTA per SM, `warps` warps on each: per-SM issue behaviour (4 schedulers -> warps/4 per scheduler)
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) lat<0><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
                if (op == 1) lat<1><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
                if (op == 2) lat<2><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
                if (op == 3) lat<3><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
                if (op == 4) lat<4><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-36s warps/SM %2d : %.2f cycles per chained op per warp  (%.2f warp-ops per cycle per SM)\n", names[op], warps,
                   (double)h / (n * 16.0), warps * n * 16.0 / (double)h);
        }
    }
    return 0;
}
