// Probe: accuracy of the MUFU.RSQ64H seed (rsqrt.approx.ftz.f64) and of the refinement steps built on it, over the
// range of q = r^4 + vc^4 the Biot-Savart kernels see.  Evidence for the choice of a third-order step in
// rsqrt_fast (common.cuh): a plain Newton step leaves ~(3/2) e^2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/probe_rsq scripts/probe_rsq.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double q)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(q));
    return y0;
}

__global__ void k(long n, double lo_exp, double hi_exp, double *out)
{
    // out[0] max rel err seed, [1] newton, [2] third order, [3] newton with halved seed via exponent trick
    double m0 = 0, m1 = 0, m2 = 0;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double f = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        double q = exp2(lo_exp + (hi_exp - lo_exp) * f);
        double ref = 1.0 / sqrt(q);
        double y0 = seed(q);
        double t = q * y0, e = fma(-t, y0, 1.0);
        double yn = fma(0.5 * y0, e, y0);
        double y3 = fma(y0 * e, fma(0.375, e, 0.5), y0);
        m0 = fmax(m0, fabs(y0 - ref) / ref);
        m1 = fmax(m1, fabs(yn - ref) / ref);
        m2 = fmax(m2, fabs(y3 - ref) / ref);
    }
    for (int o = 16; o; o >>= 1) {
        m0 = fmax(m0, __shfl_xor_sync(~0u, m0, o));
        m1 = fmax(m1, __shfl_xor_sync(~0u, m1, o));
        m2 = fmax(m2, __shfl_xor_sync(~0u, m2, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(m0));
        atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(m1));
        atomicMax((unsigned long long *)&out[2], (unsigned long long)__double_as_longlong(m2));
    }
}

int main()
{
    double *d, h[4];
    cudaMalloc(&d, 32);
    const double ranges[][2] = {{-20, 20}, {-60, 60}, {0, 1}, {-16, 18}};
    for (auto &r : ranges) {
        cudaMemset(d, 0, 32);
        k<<<148 * 8, 256>>>(1L << 30, r[0], r[1], d);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("q in 2^[%g,%g]: max rel err seed %.3e (2^%.2f)  newton %.3e  third-order %.3e   err=%s\n", r[0], r[1],
               h[0], log2(h[0]), h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
