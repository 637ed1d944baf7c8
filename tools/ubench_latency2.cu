// Finer-grained dependent-chain latencies of PLL sub-sequences.  sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include "fmrx_pll_core.h"
using namespace pllcore;
#define N 8192

template <int OP> __global__ void chain(double seed, float fseed, long long *out, double *sink)
{
    double a = seed; float fa = fseed; int ia = 3;
    double b = 0.75, c2 = 0.1;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        if (OP == 0) { a = (double)__double2float_rn(a); }                                 // d->f->d
        if (OP == 1) { fa = __double2float_rn(__dadd_rn((double)fa, 1e-9)); }              // f->d, dadd, d->f
        if (OP == 2) { double q = __fma_rn(a, FMRX_2_OVER_PI, FMRX_RINT_MAGIC); double nd = __dadd_rn(q, -FMRX_RINT_MAGIC);
                       double r = __fma_rn(-nd, FMRX_PIO2_1, a); r = __fma_rn(-nd, FMRX_PIO2_2, r); r = __fma_rn(-nd, FMRX_PIO2_3, r);
                       a = __dadd_rn(r, 1000.25); }                                          // reduction: 6 ops + 1
        if (OP == 3) { double sn, cs, r; int n; sincos_reduced(a, sn, cs, r, n); a = __dadd_rn(sn, 1000.25); }   // full reduced sincos + 1
        if (OP == 4) { double sn, cs, r; int n; sincos_reduced(a, sn, cs, r, n); double s2, c3; rotate_quadrant(n, sn, cs, s2, c3); a = __dadd_rn(c3, 1000.25); }
        if (OP == 5) { a = (a > b) ? c2 : __dadd_rn(a, 0.3); }                             // dsetp + sel
        if (OP == 6) { a = -a; a = __dadd_rn(a, 0.3); }                                    // negate + dadd
        if (OP == 7) { fa = (ia & 1) ? -fa : fa; fa = __fadd_rn(fa, 0.3f); ia += (fa > 0.f); }   // float select chain
        if (OP == 8) { double z = __dmul_rn(a, a); double z2 = __dmul_rn(z, z); double z4 = __dmul_rn(z2, z2);
                       double s12 = __fma_rn(8.3e-3, z, -1.6e-1), s34 = __fma_rn(2.7e-6, z, -1.9e-4), s56 = __fma_rn(1.5e-10, z, -2.5e-8);
                       double ps = __fma_rn(s34, z2, s12); ps = __fma_rn(s56, z4, ps); a = __fma_rn(__dmul_rn(z, a), ps, a); }   // sin poly: 5 levels
        if (OP == 9) { float f = __double2float_rn(a); float g = __fmul_rn(f, 1.0001f); a = (double)g; }                        // d->f, fmul, f->d
        if (OP == 10) { a = __fma_rn(a, 1.0000001, 1e-9); a = __fma_rn(a, 0.9999999, 1e-9); }  // 2 dfma
        if (OP == 11) { float f = __double2float_rn(a); a = __hiloint2double(__float_as_int(f) >> 3, 0); a = __dadd_rn(a, 1.0); } // d->f + int + dadd
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[OP] = (t1 - t0); sink[OP] = a + fa + ia; }
}

int main()
{
    long long *d_out, h[16]; double *d_sink;
    cudaMalloc(&d_out, 16 * sizeof(long long)); cudaMalloc(&d_sink, 16 * sizeof(double));
    const char *names[] = { "d->f->d", "f->d, DADD, d->f", "reduction (fma,add,3fma)+dadd", "sincos_reduced + dadd",
                            "sincos_reduced + rotate(double) + dadd", "DSETP+select / DADD", "negate + DADD", "float select + FADD",
                            "sin poly (z, estrin, final)", "d->f, FMUL, f->d", "2 DFMA", "d->f, int shift, DADD" };
    for (int rep = 0; rep < 2; rep++) {
        chain<0><<<1, 32>>>(0.6, 0.5f, d_out, d_sink); chain<1><<<1, 32>>>(0.6, 0.5f, d_out, d_sink);
        chain<2><<<1, 32>>>(1000.6, 0.5f, d_out, d_sink); chain<3><<<1, 32>>>(1000.6, 0.5f, d_out, d_sink);
        chain<4><<<1, 32>>>(1000.6, 0.5f, d_out, d_sink); chain<5><<<1, 32>>>(0.6, 0.5f, d_out, d_sink);
        chain<6><<<1, 32>>>(0.6, 0.5f, d_out, d_sink); chain<7><<<1, 32>>>(0.6, 0.5f, d_out, d_sink);
        chain<8><<<1, 32>>>(0.6, 0.5f, d_out, d_sink); chain<9><<<1, 32>>>(0.6, 0.5f, d_out, d_sink);
        chain<10><<<1, 32>>>(0.6, 0.5f, d_out, d_sink); chain<11><<<1, 32>>>(0.6, 0.5f, d_out, d_sink);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 12; i++) printf("%-42s %7.2f cycles/iter\n", names[i], (double)h[i] / N);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
