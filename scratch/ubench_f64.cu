// FP64 issue ceilings on B200: mma.sync.m8n8k4.f64 (DMMA) against DFMA, per SM, for 1/2/4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(int iters, double* out, double a0) {
    double acc[NACC][2];
    for (int i = 0; i < NACC; ++i) { acc[i][0] = threadIdx.x; acc[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = a0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(int iters, double* out, double a0) {
    double acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    double a = a0 + threadIdx.x * 1e-9, b = a0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(a, b, acc[i]);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    double* out;
    cudaMalloc(&out, 148 * 1024 * 8 * 8);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = warps * 32;
        float ms = timeit([&] { k_dmma<16><<<148, threads>>>(iters, out, 1.0); });
        double flops = 148.0 * warps * iters * 16 * 512.0;
        printf("DMMA  16 acc, %2d warps/SM: %7.2f TF/s  (%.1f cycles per DMMA per SMSP at 1.965 GHz)\n", warps, flops / ms / 1e9,
               ms * 1e-3 * 1.965e9 / (iters * 16.0 * warps / 4));
        ms = timeit([&] { k_dmma<4><<<148, threads>>>(iters, out, 1.0); });
        flops = 148.0 * warps * iters * 4 * 512.0;
        printf("DMMA   4 acc, %2d warps/SM: %7.2f TF/s\n", warps, flops / ms / 1e9);
        ms = timeit([&] { k_dfma<32><<<148, threads>>>(iters, out, 1.0); });
        flops = 148.0 * warps * 32.0 * iters * 32 * 2.0;
        printf("DFMA  32 acc, %2d warps/SM: %7.2f TF/s  (%.2f cycles per warp-DFMA per SMSP)\n", warps, flops / ms / 1e9,
               ms * 1e-3 * 1.965e9 / (iters * 32.0 * warps / 4));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
