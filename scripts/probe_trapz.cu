// Probe: cost of one 79-term group trapz (the solve kernel's building block): cold (first call) vs warm, group-mask
// shuffles vs full-mask shuffles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/_build/probe_trapz scripts/probe_trapz.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../ludvm_b200/csrc/common.cuh"
using namespace ludvm;
namespace ludvm { int set_error(int c, const char *, ...) { return c; } }

__device__ __noinline__ double trapz_gm(const double *a, const double *b, const double *dx, int P)
{
    auto term = [=](int j) { return dx[j] * (a[j + 1] * b[j + 1] + a[j] * b[j]) / 2.0; };
    return 0.0 + pw_group(term, 0, P - 1, (int)(threadIdx.x & 7));
}

// full-mask variant of the leaf (all four groups of the warp must have the same length)
template <class F>
__device__ __forceinline__ double leaf_full(F f, int off, int m, int lane8)
{
    const unsigned full = 0xffffffffu;
    const int body = (m >= 8) ? (m & ~7) : 0, cnt = body >> 3, rem = m - body;
    double t[16], tail = 0.0;
#pragma unroll
    for (int b = 0; b < 16; b++) if (b < cnt) t[b] = f(off + 8 * b + lane8);
    if (lane8 < rem) tail = f(off + body + lane8);
    double a = -0.0;
    if (cnt) {
        a = t[0];
#pragma unroll
        for (int b = 1; b < 16; b++) if (b < cnt) a = __dadd_rn(a, t[b]);
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) a = __dadd_rn(a, __shfl_xor_sync(full, a, s));
    }
    for (int k = 0; k < rem; k++) a = __dadd_rn(a, __shfl_sync(full, tail, k, 8));
    return a;
}
__device__ __noinline__ double trapz_full(const double *a, const double *b, const double *dx, int P)
{
    auto term = [=](int j) { return dx[j] * (a[j + 1] * b[j + 1] + a[j] * b[j]) / 2.0; };
    return 0.0 + leaf_full(term, 0, P - 1, (int)(threadIdx.x & 7));
}

__global__ void k(double *out, long long *cyc, int P)
{
    __shared__ double A[4][128], B[128], DX[128], B2[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) { for (int q = 0; q < 4; q++) A[q][i] = 1.0 / (1 + i + q); B[i] = 0.5 + i; B2[i] = 1.0; DX[i] = 1e-2 * (i + 1); }
    __syncthreads();
    int grp = threadIdx.x >> 3;
    double acc = 0;
    long long t0, t1;
    for (int rep = 0; rep < 4; rep++) {
        __syncthreads();
        t0 = clock64();
        if (grp < 4) acc += trapz_gm(A[grp], B, DX, P);
        __syncthreads();
        t1 = clock64(); if (threadIdx.x == 0) cyc[rep] = t1 - t0;
    }
    for (int rep = 0; rep < 4; rep++) {
        __syncthreads();
        t0 = clock64();
        if (grp < 4) acc += trapz_full(A[grp], B, DX, P);
        __syncthreads();
        t1 = clock64(); if (threadIdx.x == 0) cyc[4 + rep] = t1 - t0;
    }
    for (int rep = 0; rep < 4; rep++) {   // a and b both vary across the groups of the warp
        __syncthreads();
        t0 = clock64();
        if (grp < 4) acc += trapz_gm(A[0] + (grp & 1) * 128, B + (grp >> 1) * 128, DX, P);
        __syncthreads();
        t1 = clock64(); if (threadIdx.x == 0) cyc[12 + rep] = t1 - t0;
    }
    for (int rep = 0; rep < 4; rep++) {   // all 32 groups busy
        __syncthreads();
        t0 = clock64();
        acc += trapz_gm(A[grp & 3], B, DX, P);
        __syncthreads();
        t1 = clock64(); if (threadIdx.x == 0) cyc[8 + rep] = t1 - t0;
    }
    out[threadIdx.x] = acc;
}
int main()
{
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 16 * 8);
    k<<<1, 256>>>(out, cyc, 80);
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    printf("group-mask, warp 0 only : %lld %lld %lld %lld cycles (first = cold)\n", h[0], h[1], h[2], h[3]);
    printf("full-mask,  warp 0 only : %lld %lld %lld %lld\n", h[4], h[5], h[6], h[7]);
    printf("group-mask, 8 warps     : %lld %lld %lld %lld\n", h[8], h[9], h[10], h[11]);
    printf("group-mask, a,b vary    : %lld %lld %lld %lld\n", h[12], h[13], h[14], h[15]);
    return 0;
}
