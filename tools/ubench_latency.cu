// Dependent-chain latency micro-benchmark for the ops on the PLL critical path.
// One warp, one block; reports SM cycles per dependent op.  sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include "fmrx_pll_core.h"

#define N 4096
template <int OP> __global__ void chain(double seed, float fseed, long long *out, double *sink)
{
    double a = seed, b = 1.0000001, c = 1e-9;
    float fa = fseed, fb = 1.0000001f, fc = 1e-9f;
    int ia = (int)seed;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) a = __fma_rn(a, b, c);
        if (OP == 1) a = __dadd_rn(a, c);
        if (OP == 2) a = __dmul_rn(a, b);
        if (OP == 3) fa = __fmaf_rn(fa, fb, fc);
        if (OP == 4) fa = __fadd_rn(fa, fc);
        if (OP == 5) fa = __fmul_rn(fa, fb);
        if (OP == 6) { a = (double)fa; fa = __double2float_rn(a); }       // 2 conversions
        if (OP == 7) { a = __fma_rn(a, b, c); fa = __double2float_rn(a); a = (double)fa; }  // dfma + 2 cvt
        if (OP == 8) fa = __shfl_sync(0xffffffffu, fa, (i + 1) & 31);
        if (OP == 9) ia = ia * 3 + i;
        if (OP == 10) a = (a > 0.5) ? -a : a + 1.0;   // select-ish + dadd
        if (OP == 11) { int hi = __double2hiint(a); a = __hiloint2double(hi ^ 0x100000, __double2loint(a)); a = __dadd_rn(a, c); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[OP] = (t1 - t0); sink[OP] = a + fa + ia; }
}

__global__ void pll_chain(float x, int n, long long *out, float *sink, int variant)
{
    using namespace pllcore;
    Consts k; k.kp = 0.01f * 2.666f; k.ki = 0.01f * 0.01f * 3.555f;
    k.w = (2.0 * 3.14159265358979323846) * (double)(19000.0f / 240000.0f);
    const TrigK K = trig_constants();
    Chain c; c.integ = 0; c.ph = 0; c.fi = 0.5403023f; c.fq = 0.84147096f; c.toff = 100000.0f;
    chain_load(c, k);
    float acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        // a pilot the loop can lock to: 19 kHz at 240 kS/s
        float xx = x * __sinf(0.4974188f * (float)(i + 100000));
        acc += chain_step(c, k, K, xx, nullptr);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; sink[0] = acc + c.ph; }
}

__global__ void sincos_chain(float x, int n, long long *out, float *sink)
{
    using namespace pllcore;
    float v = x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        double sr, cr, r, nd;
        sincos_reduced(trig_constants(), (double)v, sr, cr, r, nd);
        v = __fadd_rn(__double2float_rn(cr), x);   // dependent
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; sink[0] = v; }
}

int main()
{
    long long *d_out, h[16];
    double *d_sink; float *f_sink;
    cudaMalloc(&d_out, 16 * sizeof(long long));
    cudaMalloc(&d_sink, 16 * sizeof(double));
    cudaMalloc(&f_sink, 16 * sizeof(float));
    const char *names[] = { "DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "F2F.f32->f64 + F2F.f64->f32",
                            "DFMA + 2 cvt", "SHFL", "IMAD", "DSETP+sel+DADD", "hi/lo int xor + DADD" };
    for (int rep = 0; rep < 2; rep++) {
        chain<0><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<1><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        chain<2><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<3><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        chain<4><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<5><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        chain<6><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<7><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        chain<8><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<9><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        chain<10><<<1, 32>>>(0.5, 0.5f, d_out, d_sink); chain<11><<<1, 32>>>(0.5, 0.5f, d_out, d_sink);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 12; i++)
        printf("%-32s %7.2f cycles/iter\n", names[i], (double)h[i] / N);
    for (int rep = 0; rep < 2; rep++) {
        pll_chain<<<1, 32>>>(0.05f, 20000, d_out, f_sink, 0);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("%-32s %7.2f cycles/iter\n", "pll chain_step (no memory)", (double)h[0] / 20000);
    for (int rep = 0; rep < 2; rep++) {
        sincos_chain<<<1, 32>>>(1000.5f, 20000, d_out, f_sink);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("%-32s %7.2f cycles/iter\n", "sincos_f32arg + cvt + fadd", (double)h[0] / 20000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); pll_chain<<<1, 32>>>(0.05f, 200000, d_out, f_sink, 0); cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("pll chain: %.1f ns/iter, %.1f cycles/iter => SM clock %.0f MHz\n", ms * 1e6 / 200000, (double)h[0] / 200000,
           (double)h[0] / (ms * 1e3));
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
