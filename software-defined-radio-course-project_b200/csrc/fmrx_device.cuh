// fmrx_device.cuh -- device-side arithmetic primitives for the FM receive chain.
//
// Everything upstream of (and including) the PLL must reproduce the
// reference's IEEE operation sequence bit for bit (g++ -O3 on baseline x86-64:
// no FMA contraction, round-to-nearest, denormals kept).  So every float
// operation on a parity-critical path goes through the explicit _rn intrinsics
// below, which nvcc never contracts into FFMA, and the library is compiled
// with -fmad=false -ftz=false -prec-div=true -prec-sqrt=true as a second line
// of defence.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "fmrx_pll_core.h"

namespace fmrx {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float d2f(double a) { return __double2float_rn(a); }

// src/iofunc.cpp:67 -- (float(u8) - 128.0)/128.0 evaluated in double and stored
// to float.  Every step is exact, so the float sequence below is identical.
__device__ __forceinline__ float unpack_u8(uint32_t v)
{
    return fmul(fsub((float)v, 128.0f), 0.0078125f);
}

// The same value from byte K of a packed word without an integer->float conversion:
// 0x47800000 | b is the float 65536 + b/128 (one ulp there is 1/128), and subtracting
// 65537 leaves (b - 128)/128 exactly -- every step is exact, so the bits are those of
// unpack_u8 (b = 128 gives +0 either way).
template <int K> __device__ __forceinline__ float unpack_u8_byte(uint32_t word)
{
    return fsub(__uint_as_float(__byte_perm(word, 0x47800000u, 0x7650 + K)), 65537.0f);
}

// src/filter.cpp:110-132, one sample of the FM discriminator.
__device__ __forceinline__ float fm_discriminate(float ci, float cq, float pi_, float pq_)
{
    const float di = fsub(ci, pi_);
    const float dq = fsub(cq, pq_);
    // std::pow(float, int) promotes to double: i*i and q*q are exact there,
    // their sum is rounded once in double and once more to float (:118).
    const double dd = dadd(dmul((double)ci, (double)ci), dmul((double)cq, (double)cq));
    const float den = d2f(dd);
    const float num = fsub(fmul(ci, dq), fmul(cq, di));
    return (den != 0.0f) ? fdiv(num, den) : 0.0f;
}

// src/filter.cpp:170 with the PLL's float trigArg: cos evaluated in double on
// the float expression (trigArg*scale)+adjust, rounded to float.
__device__ __forceinline__ float nco_from_trig(float trig_arg, float scale, float adjust)
{
    const float a = fadd(fmul(trig_arg, scale), adjust);
    return pllcore::cos_of_float(a);   // exact reduction for |a| <= 2^24, library cos beyond
}

// src/filter.cpp:180-183
__device__ __forceinline__ float mix2(float a, float b) { return fmul(2.0f, fmul(a, b)); }

// src/project.cpp:185-191: static_cast<short>(x*16384) as x86-64 executes it
// (cvttss2si to 32 bits, "integer indefinite" 0x80000000 when out of range,
// low 16 bits kept), NaN -> 0.
__device__ __forceinline__ uint32_t pcm_s16(float v)
{
    if (v != v)
        return 0u;
    const float s = fmul(v, 16384.0f);
    int w;
    if (!(s >= -2147483648.0f && s < 2147483648.0f))
        w = (int)0x80000000;
    else
        w = __float2int_rz(s);
    return (uint32_t)w & 0xffffu;
}

// PLL loop constants, src/filter.cpp:139-143,167.
struct PllParams {
    float kp;       // normBandwidth * 2.666f
    float ki;       // (normBandwidth*normBandwidth) * 3.555f
    double w;       // (2*PI) * (double)(freq/Fs)
    float scale;    // nocoScale
    float adjust;   // phaseAdjust
};

}  // namespace fmrx
