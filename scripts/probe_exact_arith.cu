// Probe: the branch-free restatements of the fast paths of __dsqrt_rn and __ddiv_rn in common.cuh (dsqrt_rn_try,
// ddiv2_rn_try, and the stage-by-stage pair_exact_try_batch<4> built from them) against the library routines, bit
// for bit, over 2^32 random operand sets each, drawn (a) from the ranges the Biot-Savart kernels see and (b) from the
// whole exponent range including zeros, subnormals, huge and tiny values.  Wherever the restatement does NOT raise
// its range flag the result must equal the library's; the flagged share (library routines used instead) is printed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/probe_exact_arith scripts/probe_exact_arith.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../ludvm_b200/csrc/common.cuh"
using namespace ludvm;
namespace ludvm { int set_error(int c, const char *, ...) { return c; } }

__device__ __forceinline__ unsigned long long rnd(unsigned long long &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
__device__ __forceinline__ double uni(unsigned long long &s) { return (double)(rnd(s) >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ bool differ(double a, double b) { return __double_as_longlong(a) != __double_as_longlong(b) && !(a != a && b != b); }

// what: 0 = sqrt, 1 = div2, 2 = pair batch
__global__ void k(long n, int what, int wide, unsigned long long *out)
{
    unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1), mism = 0, flagged = 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (what == 0) {
            double q;
            if (wide) q = __longlong_as_double((long long)rnd(s));                 // any bit pattern (negatives, NaN, inf too)
            else { double r2 = uni(s) * 3200.0 * uni(s); q = r2 * r2 + 1.78e-5 * (double)(rnd(s) & 3) * uni(s); }
            unsigned worst = 0;
            double v = dsqrt_rn_try(q, worst);
            if (ex_bad(worst)) flagged++; else if (differ(v, __dsqrt_rn(q))) mism++;
        } else if (what == 1) {
            double a1, a2, b;
            if (wide) {
                a1 = __longlong_as_double((long long)rnd(s)); a2 = __longlong_as_double((long long)rnd(s)); b = __longlong_as_double((long long)rnd(s));
                if ((rnd(s) & 63) == 0) a1 = 0.0;
            } else {
                a1 = (uni(s) - 0.5) * 80.0; a2 = (uni(s) - 0.5) * 80.0;
                if ((rnd(s) & 255) == 0) a1 = 0.0;
                double r2 = a1 * a1 + a2 * a2;
                b = 6.283185307179586 * sqrt(r2 * r2 + 1.78e-5 * (double)(rnd(s) & 3));
            }
            unsigned worst = 0;
            double q1, q2;
            ddiv2_rn_try(a1, a2, b, q1, q2, worst);
            if (ex_bad(worst)) flagged++; else if (differ(q1, __ddiv_rn(a1, b)) || differ(q2, __ddiv_rn(a2, b))) mism++;
        } else {
            double xp[4], zp[4], xw[4], zw[4], g[4], vc4[4], tu[4], tw[4];
            for (int c = 0; c < 4; c++) {
                double sc = wide ? exp2((double)((int)(rnd(s) % 600) - 300)) : 1.0;
                xp[c] = (uni(s) - 1.0) * 20.0 * sc; zp[c] = (uni(s) - 0.5) * 8.0 * sc;
                xw[c] = (uni(s) - 1.0) * 20.0 * sc; zw[c] = (uni(s) - 0.5) * 8.0 * sc;
                if ((rnd(s) & 255) == 0) { xw[c] = xp[c]; if (rnd(s) & 1) zw[c] = zp[c]; }
                g[c] = (uni(s) - 0.5) * 0.1;
                vc4[c] = wide ? uni(s) * sc : 1.78e-5 * uni(s);
            }
            if (!pair_exact_try_batch<4>(xp, zp, xw, zw, g, vc4, tu, tw)) flagged++;
            else for (int c = 0; c < 4; c++) {
                double ru, rw;
                pair_exact_ref(xp[c], zp[c], xw[c], zw[c], g[c], vc4[c], ru, rw);
                if (differ(ru, tu[c]) || differ(rw, tw[c])) mism++;
            }
        }
    }
    if (mism) atomicAdd(out, mism);
    if (flagged) atomicAdd(out + 1, flagged);
}
int main()
{
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    const char *names[3] = {"dsqrt_rn_try vs __dsqrt_rn", "ddiv2_rn_try vs __ddiv_rn (2 quotients)", "pair_exact_try_batch<4> vs library pair"};
    for (int what = 0; what < 3; what++)
        for (int wide = 0; wide < 2; wide++) {
            long n = what == 2 ? 1L << 30 : 1L << 32;
            cudaMemset(d, 0, 16);
            k<<<148 * 8, 256>>>(n, what, wide, d);
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-42s %-18s 2^%d operand sets: %llu mismatches, %llu flagged for the library path (%.3g %%)   (%s)\n", names[what],
                   wide ? "whole range" : "Biot-Savart range", what == 2 ? 30 : 32, h[0], h[1], 100.0 * (double)h[1] / (double)n,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
