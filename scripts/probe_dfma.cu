// Probe: is the DFMA-rate microbenchmark of capi.cu (8 chains, 2048 threads/SM, multiplier and addend in constant
// registers) the best the FP64 pipe does?  Variants: 4/8/16 chains, operands in registers, 1024/2048 threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/probe_dfma scripts/probe_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C, bool REGOPS>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double c[C], ar = a, br = b;
    if (REGOPS) { ar += 1e-300 * threadIdx.x; br += 1e-300 * threadIdx.x; }
#pragma unroll
    for (int i = 0; i < C; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 64 / C; u++)
#pragma unroll
            for (int i = 0; i < C; i++) c[i] = fma(c[i], ar, br);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; i++) s += c[i];
    if (s == 123456789.0) out[0] = s;
}
template <int C, bool REGOPS>
void run(const char *name, int blocks_per_sm, double *d, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 200000; float ms = 0, best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<C, REGOPS><<<sms * blocks_per_sm, 256>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double rate = (double)sms * blocks_per_sm * 256.0 * iters * 64.0 / (best * 1e-3);
    printf("%-44s %8.1f ms  %.4e DFMA/s  = %.1f %% of %d SM x 64 x 1.965 GHz\n", name, best, rate, 100 * rate / (sms * 64 * 1.965e9), sms);
}
int main()
{
    double *d; cudaMalloc(&d, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    run<8, false>("8 chains, const operands, 2048 thr/SM", 8, d, sms);
    run<8, false>("8 chains, const operands, 1024 thr/SM", 4, d, sms);
    run<4, false>("4 chains, const operands, 2048 thr/SM", 8, d, sms);
    run<16, false>("16 chains, const operands, 1024 thr/SM", 4, d, sms);
    run<8, true>("8 chains, register operands, 2048 thr/SM", 8, d, sms);
    run<16, true>("16 chains, register operands, 1024 thr/SM", 4, d, sms);
    run<8, true>("8 chains, register operands, 512 thr/SM", 2, d, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
