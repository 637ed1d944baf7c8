// SHFL / LDS dependent-lookup latency (for the candidate-select idea).  sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define N 8192
template <int OP> __global__ void chain(int seed, long long *out, int *sink)
{
    __shared__ int tab[64];
    __shared__ double dtab[64];
    tab[threadIdx.x] = (threadIdx.x * 7 + 3) & 31; tab[threadIdx.x + 32] = threadIdx.x;
    dtab[threadIdx.x] = threadIdx.x * 0.5; dtab[threadIdx.x + 32] = 1.0;
    __syncwarp();
    int ia = seed; double a = seed * 0.25;
    int lane_val = (threadIdx.x * 5 + 1) & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        if (OP == 0) ia = __shfl_sync(0xffffffffu, lane_val, ia & 31);                 // SHFL with dependent source lane
        if (OP == 1) ia = tab[ia & 31];                                                // LDS dependent
        if (OP == 2) { a = __fma_rn(a, 1.0, 6755399441055744.0); ia = __double2loint(a); ia = __shfl_sync(0xffffffffu, lane_val, ia & 31); a = (double)0 + __hiloint2double(0x3ff00000 + (ia << 10), 0); }  // dfma + lo32 + shfl + int->hi
        if (OP == 3) { a = dtab[ia & 31]; ia = __double2loint(__dadd_rn(a, 6755399441055744.0)); }   // LDS.64 + dadd magic + lo32
        if (OP == 4) { ia = (int)a; a = (double)ia + 0.5; }                            // F2I.F64 + I2F.F64 (+dadd)
        if (OP == 5) { ia = ia * 3 + 1; ia &= 0xffff; }                                // 2 int ops
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[OP] = (t1 - t0); sink[OP] = ia + (int)a; }
}
int main()
{
    long long *d_out, h[8]; int *d_sink;
    cudaMalloc(&d_out, 8 * sizeof(long long)); cudaMalloc(&d_sink, 8 * sizeof(int));
    const char *names[] = { "SHFL.IDX (dependent lane)", "LDS (dependent address)", "DFMA+lo32+SHFL+int->double", "LDS.64 + DADD + lo32", "F2I.F64 + I2F.F64 + DADD", "IMAD + LOP3" };
    for (int rep = 0; rep < 2; rep++) {
        chain<0><<<1, 32>>>(3, d_out, d_sink); chain<1><<<1, 32>>>(3, d_out, d_sink); chain<2><<<1, 32>>>(3, d_out, d_sink);
        chain<3><<<1, 32>>>(3, d_out, d_sink); chain<4><<<1, 32>>>(3, d_out, d_sink); chain<5><<<1, 32>>>(3, d_out, d_sink);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 6; i++) printf("%-32s %7.2f cycles/iter\n", names[i], (double)h[i] / N);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
