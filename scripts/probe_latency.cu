// Probe: single-warp dependent-issue latencies that bound the solve kernel (one CTA, latency-bound).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/probe_latency scripts/probe_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double *out, long long *cyc, double a, double b, int *chase_g)
{
    __shared__ double sm[256];
    __shared__ int chase[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sm[i] = 1.0 + i; chase[i] = (i + 33) & 255; }
    __syncthreads();
    double x = threadIdx.x, y = a;
    long long t0, t1;
    // 0: dependent DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dadd_rn(x, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 1: dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dmul_rn(x, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 2: dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, a, b);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 3: shfl_xor(double) + DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __dadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 1));
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 4: LDS pointer chase
    int p = threadIdx.x & 255;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = chase[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 5: generic-pointer chase into shared memory
    const int *gp = chase; asm volatile("" : "+l"(gp));
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = gp[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 6: global pointer chase (L1/L2 hit)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = __ldcg(chase_g + p);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 7: __syncthreads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 8: IEEE division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) y = __ddiv_rn(y, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    // 9: IEEE sqrt chain
    y = fabs(y) + 2.0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) y = __dsqrt_rn(y + a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    // 10: global L1-cached pointer chase (ld.ca default)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = chase_g[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
    out[threadIdx.x] = x + y + p;
}
int main()
{
    double *out; long long *cyc, h[16]; int *cg, hc[256];
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&cg, 1024);
    for (int i = 0; i < 256; i++) hc[i] = (i + 33) & 255;
    cudaMemcpy(cg, hc, 1024, cudaMemcpyHostToDevice);
    const char *names[] = {"DADD dep", "DMUL dep", "DFMA dep", "SHFL64+DADD", "LDS chase", "generic->smem chase", "ld.cg global chase (L2)",
                           "__syncthreads", "ddiv_rn dep", "dsqrt_rn(+add) dep", "ld.ca global chase (L1)"};
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; rep++) k<<<1, threads>>>(out, cyc, 0.999999, 1e-9, cg);
        cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
        printf("threads=%d  %s\n", threads, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 11; i++) printf("  %-26s %7.1f cycles/op\n", names[i], (double)h[i] / N);
    }
    return 0;
}
