// Sustained rate of mma.sync.m8n8k4.f64 (the FP64 tensor-core path of sm_100a) next to the DFMA rate: the roofline
// denominator of k_tree_m2l_dmma.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/probe_dmma scripts/probe_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a, double b)
{
    double c[CH][2];
#pragma unroll
    for (int k = 0; k < CH; k++) c[k][0] = c[k][1] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < CH; k++) dmma(c[k][0], c[k][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += c[k][0] + c[k][1];
    if (s == 123456789.0) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = threadIdx.x + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = c[k] * a + b;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k];
    if (s == 123456789.0) out[0] = s;
}
template <class F> static double timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3;
}
int main()
{
    double *d; cudaMalloc(&d, 64);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000;
    for (int wpsm : {4, 8, 16, 32, 64}) {
        const int blocks = sm * wpsm / 8;
        double t8 = timeit([&] { k_dmma<8><<<blocks, 256>>>(d, iters, 0.999999, 1e-7); });
        double t4 = timeit([&] { k_dmma<4><<<blocks, 256>>>(d, iters, 0.999999, 1e-7); });
        double f8 = (double)blocks * 8 * iters * 4 * 8 * 256 / t8, f4 = (double)blocks * 8 * iters * 4 * 4 * 256 / t4;
        printf("warps/SM %2d: DMMA 8 chains %.3e FMA/s (%.1f TFLOP/s), 4 chains %.3e FMA/s\n", wpsm, f8, 2 * f8 / 1e12, f4);
    }
    double t = timeit([&] { k_dfma<<<sm * 8, 256>>>(d, iters, 0.999999, 1e-7); });
    printf("DFMA (64 warps/SM, 8 chains): %.3e FMA/s (%.1f TFLOP/s)\n", (double)sm * 8 * 256 * iters * 64 / t, 2.0 * sm * 8 * 256 * iters * 64 / t / 1e12);
    return 0;
}
