// Probe: cycle cost of the block-wide reductions of block_reduce.cuh (one CTA of 256 threads), piece by piece.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o scripts/_build/probe_block scripts/probe_block.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../ludvm_b200/csrc/block_reduce.cuh"
using namespace ludvm;
namespace ludvm { int set_error(int c, const char *, ...) { return c; } }

#define T(k) do { __syncthreads(); if (threadIdx.x == 0) cyc[k] = clock64(); } while (0)

__global__ void __launch_bounds__(256) k(const double *cosn, const double *parts, double *out, long long *cyc, int P, int Nc, int nn)
{
    extern __shared__ double sm[];
    double *Wu = sm, *dth = Wu + P, *ones = dth + P, *A = ones + P, *u1 = A + 64, *w1 = u1 + P, *stage = w1 + P;
    for (int j = threadIdx.x; j < P; j += blockDim.x) { Wu[j] = 1.0 / (1 + j); dth[j] = 1e-2 * (j + 1); ones[j] = 1.0; }
    const int cap = 12000;
    for (int rep = 0; rep < 3; rep++) {
        T(0);
        block_trapz(Wu, 0, 0, cosn, 0, P, dth, P, Nc, stage, cap, A);          // 30 Fourier integrals
        T(1);
        block_trapz(Wu, 1, 0, ones, 1, 0, dth, P, 4, stage, cap, A + 32);      // 4 integrals
        T(2);
        block_fold(parts, parts + 64 * P, P, P, 3, true, stage, cap, u1, w1);  // exact, 8 nodes
        T(3);
        block_fold(parts, parts + 64 * P, P, P, 6, true, stage, cap, u1, w1);  // exact, 64 nodes
        T(4);
        block_fold(parts, parts + 64 * P, P, P, nn, false, stage, cap, u1, w1);  // fast, nn chunks
        T(5);
        // pieces of block_trapz(30): staging only
        {
            const int n = P - 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
            for (int q = warp; q < Nc; q += nwarps) {
                const double *b = cosn + (size_t)q * P;
                double *srow = stage + q * P;
#pragma unroll 4
                for (int j = lane; j < n; j += 32) srow[j] = dth[j] * (Wu[j + 1] * b[j + 1] + Wu[j] * b[j]) / 2.0;
            }
        }
        T(6);
        {   // group sums only
            const int grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
            for (int q = grp; q < Nc; q += ngrp) { double v = 0.0 + sum_group(stage + q * P, 0, P - 1); if ((threadIdx.x & 7) == 0) A[q] = v; }
        }
        T(7);
        {   // one thread per sum (sequential numpy tree)
            if (threadIdx.x < Nc) { const double *v = stage + threadIdx.x * P; auto f = [v](int j) { return v[j]; }; A[threadIdx.x] = 0.0 + pw_seq(f, 0, P - 1); }
        }
        T(8);
        for (int i = 0; i < 16; i++) __syncthreads();
        T(9);
        if (threadIdx.x == 0 && rep == 2) for (int i = 0; i < 9; i++) cyc[16 + i] = cyc[i + 1] - cyc[i];
    }
    out[threadIdx.x] = A[threadIdx.x & 31] + u1[threadIdx.x % P];
}
int main()
{
    const int P = 80, Nc = 30;
    double *cosn, *parts, *out; long long *cyc, h[32];
    cudaMalloc(&cosn, 8 * P * Nc); cudaMalloc(&parts, 8 * P * 128); cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8 * 32);
    cudaMemset(cosn, 0, 8 * P * Nc); cudaMemset(parts, 0, 8 * P * 128);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000);
    k<<<1, 256, 120000>>>(cosn, parts, out, cyc, P, Nc, 64);
    cudaMemcpy(h, cyc, 8 * 32, cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    const char *nm[] = {"block_trapz x30 (global b)", "block_trapz x4", "block_fold exact 8 nodes", "block_fold exact 64 nodes", "block_fold fast 64 chunks",
                        "  staging of 30 integrands only", "  30 group sums only", "  30 thread sums (pw_seq)", "16 barriers"};
    for (int i = 0; i < 9; i++) printf("%-34s %7lld cycles\n", nm[i], h[16 + i]);
    return 0;
}
