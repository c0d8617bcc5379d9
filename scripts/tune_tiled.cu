// Tuning harness for the fast fp64 tiled all-pairs kernel: rows/thread R, CTA size T, inner unroll U, tile size.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/tune_tiled scripts/tune_tiled.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ludvm_b200/csrc/common.cuh"
using namespace ludvm;
namespace ludvm { int set_error(int c, const char *, ...) { return c; } }

template <int R, int T, int U, int TILE, int MINB>
__global__ void __launch_bounds__(T, MINB) k(const double *__restrict__ xs, const double *__restrict__ zs,
                                              const double *__restrict__ gs, double vc4, int n, int chunk_len,
                                              double *pu, double *pw)
{
    __shared__ double sx[TILE], sz[TILE], sg[TILE];
    double tx[R], tz[R], au[R], aw[R];
    int base = blockIdx.x * (T * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; r++) { int row = min(base + r * T, n - 1); tx[r] = xs[row]; tz[r] = zs[row]; au[r] = 0; aw[r] = 0; }
    int c0 = blockIdx.y * chunk_len, c1 = min(n, c0 + chunk_len);
    for (int t0 = c0; t0 < c1; t0 += TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < TILE; j += T) { int s = t0 + j; bool ok = s < c1; sx[j] = ok ? xs[s] : 0; sz[j] = ok ? zs[s] : 0; sg[j] = ok ? gs[s] * LUDVM_INV_TWO_PI : 0; }
        __syncthreads();
#pragma unroll U
        for (int j = 0; j < TILE; j++) {
            double x = sx[j], z = sz[j], g = sg[j];
#pragma unroll
            for (int r = 0; r < R; r++) pair_fast(tx[r], tz[r], x, z, g, vc4, au[r], aw[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) { int row = base + r * T; if (row < n) { pu[(size_t)blockIdx.y * n + row] = au[r]; pw[(size_t)blockIdx.y * n + row] = aw[r]; } }
}

template <int R, int T, int U, int TILE, int MINB>
void run(const char *name, const double *x, const double *z, const double *g, int n, double *pu, double *pw, double dfma)
{
    int chunks = 8, chunk_len = ((n + chunks - 1) / chunks + TILE - 1) / TILE * TILE;
    dim3 grid((n + T * R - 1) / (T * R), chunks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<R, T, U, TILE, MINB><<<grid, T>>>(x, z, g, 0.065 * 0.065 * 0.065 * 0.065, n, chunk_len, pu, pw);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<R, T, U, TILE, MINB>);
    double pps = (double)n * n / (best * 1e-3);
    printf("%-28s regs=%3d  ms=%8.3f  pairs/s=%.4e  frac=%.4f  err=%s\n", name, fa.numRegs, best, pps, pps * 13 / dfma, cudaGetErrorString(cudaGetLastError()));
}

__global__ void fma_rate(double *out, int iters, double a, double b)
{
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) { c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b; c4 = c4 * a + b; c5 = c5 * a + b; c6 = c6 * a + b; c7 = c7 * a + b; }
    }
    double s = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
    if (s == 123456789.0) out[0] = s;
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 262144;
    std::vector<double> hx(n), hz(n), hg(n);
    srand(1);
    for (int i = 0; i < n; i++) { hx[i] = -20.0 * rand() / RAND_MAX; hz[i] = 8.0 * rand() / RAND_MAX - 4; hg[i] = 0.01 * (rand() / (double)RAND_MAX - 0.5); }
    double *x, *z, *g, *pu, *pw;
    cudaMalloc(&x, n * 8); cudaMalloc(&z, n * 8); cudaMalloc(&g, n * 8); cudaMalloc(&pu, (size_t)n * 8 * 8); cudaMalloc(&pw, (size_t)n * 8 * 8);
    cudaMemcpy(x, hx.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(z, hz.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(g, hg.data(), n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double dfma = 0;
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); fma_rate<<<148 * 8, 256>>>(pu, 20000, 0.999999, 1e-7); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); double r = 148.0 * 8 * 256 * 20000 * 64 / (ms * 1e-3); if (r > dfma) dfma = r; }
    printf("n=%d  dfma/s=%.4e\n", n, dfma);
#define RUN(R, T, U, TILE, MINB) run<R, T, U, TILE, MINB>("R" #R " T" #T " U" #U " tile" #TILE " minb" #MINB, x, z, g, n, pu, pw, dfma)
    RUN(4, 256, 4, 512, 2);
    RUN(4, 256, 2, 512, 2);
    RUN(4, 256, 8, 512, 2);
    RUN(4, 256, 1, 512, 2);
    RUN(2, 256, 4, 512, 2);
    RUN(2, 256, 4, 512, 3);
    RUN(2, 256, 8, 512, 4);
    RUN(2, 512, 4, 512, 2);
    RUN(3, 256, 4, 512, 2);
    RUN(6, 256, 2, 512, 1);
    RUN(8, 256, 1, 512, 1);
    RUN(8, 128, 2, 512, 2);
    RUN(4, 128, 4, 512, 4);
    RUN(4, 128, 4, 512, 3);
    RUN(4, 512, 4, 512, 1);
    RUN(4, 256, 4, 1024, 2);
    RUN(4, 256, 4, 256, 2);
    RUN(1, 256, 8, 512, 4);
    RUN(1, 512, 8, 512, 4);
    RUN(1, 1024, 8, 1024, 2);
    return 0;
}
