// Probe for a 12-slot fast pair: candidate 1/sqrt(q) seeds that are accurate enough for ONE second-order step
// (5 FP64 slots for rsqrt + circulation scaling instead of 6), their worst-case error, and the throughput of the
// tiled all-pairs loop built on each.
//   V0  MUFU.RSQ64H(hi word) + third-order step                     (production until r01e: 13 slots per pair)
//   V1  MUFU.RSQ64H(hi word rounded with the low word's top bit) + second-order step
//   V2  fp32 MUFU.RSQ on q repacked as fp32 by integer operations  + second-order step
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/probe_rsq2 scripts/probe_rsq2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed64(double q) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q)); return y; }

template <int V> __device__ __forceinline__ double scaled_rsqrt(double q, double gs);
template <> __device__ __forceinline__ double scaled_rsqrt<0>(double q, double gs)
{
    double y0 = seed64(q);
    double t = q * y0, e = fma(-t, y0, 1.0), p = fma(0.375, e, 0.5), ye = y0 * e;
    return gs * fma(ye, p, y0);
}
template <> __device__ __forceinline__ double scaled_rsqrt<1>(double q, double gs)
{
    int hi = __double2hiint(q);
    unsigned lo = (unsigned)__double2loint(q);
    double y0 = seed64(__hiloint2double(hi + (int)(lo >> 31), 0));
    double t = q * y0, e = fma(-t, y0, 1.0), f = fma(e, 0.5, 1.0), a = gs * y0;
    return a * f;
}
template <> __device__ __forceinline__ double scaled_rsqrt<2>(double q, double gs)
{
    int hi = __double2hiint(q);
    unsigned lo = (unsigned)__double2loint(q);
    unsigned fb = __funnelshift_l(lo, (unsigned)(hi - 0x38000000), 3);
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(__uint_as_float(fb)));
    unsigned yb = __float_as_uint(yf);
    double y0 = __hiloint2double((int)(yb >> 3) + 0x38000000, (int)(yb << 29));
    double t = q * y0, e = fma(-t, y0, 1.0), f = fma(e, 0.5, 1.0), a = gs * y0;
    return a * f;
}

template <int V>
__global__ void k_err(long n, double lo_exp, double hi_exp, double *out)
{
    double m = 0;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double f = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        double q = exp2(lo_exp + (hi_exp - lo_exp) * f);
        if (i & 1) q = __hiloint2double(__double2hiint(q), (i & 2) ? 0x7fffffff : (int)0x80000000u);   // worst low words for V1
        double ref = 1.0 / sqrt(q);
        m = fmax(m, fabs(scaled_rsqrt<V>(q, 1.0) - ref) / ref);
    }
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(~0u, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(m));
}

template <int V>
__device__ __forceinline__ void pair(double xp, double zp, double xw, double zw, double gs, double vc4, double &au, double &aw)
{
    double dx = xp - xw, dz = zp - zw;
    double r2 = fma(dz, dz, dx * dx);
    double q = fma(r2, r2, vc4);
    double gg = scaled_rsqrt<V>(q, gs);
    au = fma(gg, dz, au);
    aw = fma(-gg, dx, aw);
}

template <int V, int U>
__global__ void __launch_bounds__(256, 2) k_tiled(const double *__restrict__ xs, const double *__restrict__ zs,
                                                   const double *__restrict__ gs, double vc4, int n, int chunk_len, double *pu, double *pw)
{
    constexpr int R = 4, T = 256, TILE = 512;
    __shared__ double sx[TILE], sz[TILE], sg[TILE];
    double tx[R], tz[R], au[R], aw[R];
    int base = blockIdx.x * (T * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; r++) { int row = min(base + r * T, n - 1); tx[r] = xs[row]; tz[r] = zs[row]; au[r] = 0; aw[r] = 0; }
    int c0 = blockIdx.y * chunk_len, c1 = min(n, c0 + chunk_len);
    for (int t0 = c0; t0 < c1; t0 += TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < TILE; j += T) { int s = t0 + j; bool ok = s < c1; sx[j] = ok ? xs[s] : 0; sz[j] = ok ? zs[s] : 0; sg[j] = ok ? gs[s] * 0.15915494309189535 : 0; }
        __syncthreads();
#pragma unroll U
        for (int j = 0; j < TILE; j++) {
            double x = sx[j], z = sz[j], g = sg[j];
#pragma unroll
            for (int r = 0; r < R; r++) pair<V>(tx[r], tz[r], x, z, g, vc4, au[r], aw[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) { int row = base + r * T; if (row < n) { pu[(size_t)blockIdx.y * n + row] = au[r]; pw[(size_t)blockIdx.y * n + row] = aw[r]; } }
}

template <int V, int U>
void run(const double *x, const double *z, const double *g, int n, double *pu, double *pw)
{
    int chunks = 8, chunk_len = ((n + chunks - 1) / chunks + 511) / 512 * 512;
    dim3 grid((n + 1023) / 1024, chunks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_tiled<V, U><<<grid, 256>>>(x, z, g, 1.78e-5, n, chunk_len, pu, pw);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    printf("V%d unroll %d: %.2f ms, %.4g pairs/s  (%s)\n", V, U, best, (double)n * n / best * 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *d, h;
    cudaMalloc(&d, 8);
    const double ranges[][2] = {{-20, 20}, {-60, 60}, {0, 2}, {-36, 22}};
    for (int v = 0; v < 3; v++)
        for (auto &r : ranges) {
            cudaMemset(d, 0, 8);
            if (v == 0) k_err<0><<<148 * 8, 256>>>(1L << 30, r[0], r[1], d);
            if (v == 1) k_err<1><<<148 * 8, 256>>>(1L << 30, r[0], r[1], d);
            if (v == 2) k_err<2><<<148 * 8, 256>>>(1L << 30, r[0], r[1], d);
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("V%d q in 2^[%g,%g]: max rel err of the refined 1/sqrt %.3e  (%s)\n", v, r[0], r[1], h, cudaGetErrorString(cudaGetLastError()));
        }
    int n = 1 << 19;
    std::vector<double> hx(n), hz(n), hg(n);
    srand(1);
    for (int i = 0; i < n; i++) { hx[i] = -20.0 * rand() / RAND_MAX; hz[i] = 8.0 * rand() / RAND_MAX - 4; hg[i] = 1e-2 * (rand() / (double)RAND_MAX - 0.5); }
    double *x, *z, *g, *pu, *pw;
    cudaMalloc(&x, n * 8); cudaMalloc(&z, n * 8); cudaMalloc(&g, n * 8); cudaMalloc(&pu, (size_t)n * 64); cudaMalloc(&pw, (size_t)n * 64);
    cudaMemcpy(x, hx.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(z, hz.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(g, hg.data(), n * 8, cudaMemcpyHostToDevice);
    run<0, 2>(x, z, g, n, pu, pw); run<1, 2>(x, z, g, n, pu, pw); run<2, 2>(x, z, g, n, pu, pw);
    run<0, 4>(x, z, g, n, pu, pw); run<1, 4>(x, z, g, n, pu, pw); run<2, 4>(x, z, g, n, pu, pw);
    run<1, 1>(x, z, g, n, pu, pw); run<2, 1>(x, z, g, n, pu, pw); run<1, 8>(x, z, g, n, pu, pw); run<2, 8>(x, z, g, n, pu, pw);
    return 0;
}
