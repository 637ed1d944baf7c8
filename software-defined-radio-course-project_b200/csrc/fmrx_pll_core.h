// fmrx_pll_core.h -- the arithmetic core of one PLL step (src/filter.cpp:157-171),
// written once for the device (k_pll) and for a host build used only by the CPU
// tests (tests/pll_model.cpp), so the low-latency formulation can be checked bit
// for bit against the oracle on hours of signal without a GPU.
//
// Why a special formulation.  Per IF sample the reference evaluates, in double,
// atan2(eQ,eI), cos(trigArg) and sin(trigArg) and rounds each to float.  The
// recurrence is one dependent chain per capture, so its LATENCY is the throughput
// bound of the whole receive chain.  Stock libdevice sin/cos fall into the
// Payne-Hanek slow path once |trigArg| > 105615 (0.9 s into a capture), and stock
// atan2 is a division plus a degree-19 polynomial.  Here:
//
//  * sincos: trigArg is a FLOAT (24-bit significand, |x| < 2^24) promoted to
//    double, so a 3-term Cody-Waite reduction with 29/29/53-bit pieces of pi/2 is
//    exact in its first two steps (n < 2^24: n*P1 and n*P2 are exact products) and
//    rounds once; the kernels are the fdlibm minimax polynomials, Estrin-ordered.
//  * atan2: its arguments are eI = fl(x*fI), eQ = fl(x*(-fQ)) with (fI,fQ) =
//    fl(cos), fl(sin) of the PREVIOUS trigArg, whose reduced argument r and
//    quadrant n are already known.  So atan2(eQ,eI) = -(phi + delta) where phi is
//    the wrapped previous trigArg (r + quadrant constant, shifted by pi when x<0)
//    and delta is the tiny rotation caused by the four float roundings, obtained
//    exactly from the FMA residuals t_i = eI - x*cos, t_q = eQ + x*sin:
//        cross = -(cos*t_q + sin*t_i)/x,  dot = (cos*t_i - sin*t_q)/x,
//        delta = cross*(1 - dot)            (|cross|,|dot| <~ 2^-23; O(2^-69) dropped)
//    Anything unusual (x = 0, subnormal products, |phi| near pi, NaN) fails the
//    guard and takes the reference formulation (true atan2) for that step.
//
// Both evaluate the same real function the reference does, to ~1 ulp of double,
// and round to float where the reference rounds; they differ from glibc's result
// only when the exact value lies within ~1e-16 relative of a float rounding
// boundary (~1e-8 per sample), the same class of event as using any other libm.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FMRX_HD __host__ __device__ __forceinline__
#else
#define FMRX_HD static inline
#endif

namespace pllcore {

// ---- exact-rounding primitives ---------------------------------------------
#if defined(__CUDA_ARCH__)
FMRX_HD double p_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
FMRX_HD double p_mul(double a, double b) { return __dmul_rn(a, b); }
FMRX_HD double p_add(double a, double b) { return __dadd_rn(a, b); }
FMRX_HD float p_fmulf(float a, float b) { return __fmul_rn(a, b); }
FMRX_HD float p_faddf(float a, float b) { return __fadd_rn(a, b); }
FMRX_HD float p_d2f(double a) { return __double2float_rn(a); }
FMRX_HD int p_lo32(double a) { return __double2loint(a); }
#else
// host build: compiled with -ffp-contract=off, so * and + are single IEEE ops
FMRX_HD double p_fma(double a, double b, double c) { return fma(a, b, c); }
FMRX_HD double p_mul(double a, double b) { return a * b; }
FMRX_HD double p_add(double a, double b) { return a + b; }
FMRX_HD float p_fmulf(float a, float b) { return a * b; }
FMRX_HD float p_faddf(float a, float b) { return a + b; }
FMRX_HD float p_d2f(double a) { return (float)a; }
FMRX_HD int p_lo32(double a)
{
    uint64_t u;
    memcpy(&u, &a, sizeof(u));
    return (int)(uint32_t)u;
}
#endif

// ---- constants ---------------------------------------------------------------
// pi/2 = P1 + P2 + P3: 29 + 29 + 53 bits (n*P1, n*P2 exact for |n| < 2^24, which
// covers every float argument |x| <= 2^24)
#define FMRX_PIO2_1 1.570796325802803       /* 0x1.921fb54000000p+0  */
#define FMRX_PIO2_2 9.920935774287987e-10   /* 0x1.10b4611000000p-30 */
#define FMRX_PIO2_3 2.2517417741562176e-18  /* 0x1.4c4c6628b80dcp-59 */
#define FMRX_FAST_TRIG_MAX 16777216.0f      /* |x| <= 2^24 */
#define FMRX_2_OVER_PI 0.6366197723675814  /* 0x1.45f306dc9c883p-1  */
#define FMRX_RINT_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */
#define FMRX_PIO2_HI 1.5707963267948966
#define FMRX_PIO2_LO 6.123233995736766e-17
#define FMRX_PI_HI 3.141592653589793
#define FMRX_PI_LO 1.2246467991473532e-16

// What one sincos leaves behind for the next step's atan2 shortcut.
struct Trig {
    double cs, sn;   // cos, sin of the float argument, in double (~1 ulp)
    double r;        // reduced argument in [-pi/4, pi/4]
    int n;           // quadrant count: argument = r + n*pi/2
};

// sin and cos of a float-valued argument |x| <= 2^24 (x = (double)float).
FMRX_HD Trig sincos_f32arg(double x)
{
    // n = rint(x * 2/pi) by the add-magic trick; low word of the sum is n
    const double qm = p_add(p_mul(x, FMRX_2_OVER_PI), FMRX_RINT_MAGIC);
    const double nd = p_add(qm, -FMRX_RINT_MAGIC);
    double r = p_fma(-nd, FMRX_PIO2_1, x);      // exact
    r = p_fma(-nd, FMRX_PIO2_2, r);             // exact product, one rounding
    r = p_fma(-nd, FMRX_PIO2_3, r);
    const double z = p_mul(r, r);
    const double z2 = p_mul(z, z);
    const double z4 = p_mul(z2, z2);
    // fdlibm __kernel_sin: sin r = r + r*z*(S1 + z*S2 + ... + z^5*S6)
    const double s12 = p_fma(8.33333333332248946124e-03, z, -1.66666666666666324348e-01);
    const double s34 = p_fma(2.75573137070700676789e-06, z, -1.98412698298579493134e-04);
    const double s56 = p_fma(1.58969099521155010221e-10, z, -2.50507602534068634195e-08);
    double ps = p_fma(s34, z2, s12);
    ps = p_fma(s56, z4, ps);
    const double sn = p_fma(p_mul(z, r), ps, r);
    // fdlibm __kernel_cos: cos r = 1 - z/2 + z^2*(C1 + z*C2 + ... + z^5*C6)
    const double c12 = p_fma(-1.38888888888741095749e-03, z, 4.16666666666666019037e-02);
    const double c34 = p_fma(-2.75573143513906633035e-07, z, 2.48015872894767294178e-05);
    const double c56 = p_fma(-1.13596475577881948265e-11, z, 2.08757232129817482790e-09);
    double pc = p_fma(c34, z2, c12);
    pc = p_fma(c56, z4, pc);
    const double cs = p_fma(z2, pc, p_fma(-0.5, z, 1.0));

    Trig t;
    t.r = r;
    t.n = p_lo32(qm);
    // rotate by n quadrants
    const int q = t.n & 3;
    const double a = (q & 1) ? cs : sn;          // |sin|
    const double b = (q & 1) ? sn : cs;          // |cos|
    t.sn = (q & 2) ? -a : a;
    t.cs = ((q + 1) & 2) ? -b : b;
    return t;
}

// cos of a float, rounded to float, for the NCO output (src/filter.cpp:170): same
// reduction; arguments beyond its range take the library cos.
FMRX_HD float cos_of_float(float a)
{
    if (fabsf(a) <= FMRX_FAST_TRIG_MAX)
        return p_d2f(sincos_f32arg((double)a).cs);
    return p_d2f(cos((double)a));
}

// PLL state carried between steps (superset of the reference's five floats: the
// double-precision leftovers of the last sincos feed the atan2 shortcut; they are
// a pure function of (trigOffset, phaseEst), so they are recomputed, not stored,
// when a state is loaded).
struct Chain {
    float integ, ph, fi, fq, toff;
    Trig trig;       // of the last trigArg (what fi, fq were rounded from)
    float ta;        // last trigArg
};

struct Consts {
    float kp, ki;
    double w;        // (2*PI)*(double)(freq/Fs)
};

// Rebuild the derived fields after loading (integ, ph, fi, fq, toff).  The
// reference's initial state is fi=1, fq=0 with no trigArg behind it; that pair is
// cos/sin of 0, and any state saved by this code satisfies fi,fq = fl(cos,sin)(ta)
// with ta = fl(w*toff + ph).  `consistent` tells the step whether the shortcut
// may trust trig for the next sample.
FMRX_HD bool chain_load(Chain &c, const Consts &k)
{
    const float ta = p_d2f(p_add(p_mul(k.w, (double)c.toff), (double)c.ph));
    c.ta = ta;
    if (!(fabsf(ta) <= FMRX_FAST_TRIG_MAX))
        return false;
    c.trig = sincos_f32arg((double)ta);
    return p_d2f(c.trig.cs) == c.fi && p_d2f(c.trig.sn) == c.fq;
}

// atan2(eq, ei) for ei = fl(x*fi), eq = fl(x*(-fq)); returns false if the guard
// rejects the shortcut (caller then evaluates the true atan2).
FMRX_HD bool atan2_shortcut(const Trig &t, float x, double inv_x, float ei, float eq, double *out)
{
    const double xd = (double)x;
    const double ti = p_fma(-xd, t.cs, (double)ei);   // = x*du (exact residual, rounded once)
    const double tq = p_fma(xd, t.sn, (double)eq);    // = -x*dv
    const double cross = -p_mul(p_fma(t.sn, ti, p_mul(t.cs, tq)), inv_x);
    const double dotc = p_mul(p_fma(-t.sn, tq, p_mul(t.cs, ti)), inv_x);
    const double delta = p_fma(-cross, dotc, cross);
    // phi: the wrapped angle of (cos, sin), turned by pi when x < 0
    const int kk = (t.n + (x < 0.0f ? 2 : 0)) & 3;
    double chi, clo;
    if (kk == 0) {
        chi = 0.0; clo = 0.0;
    } else if (kk == 1) {
        chi = FMRX_PIO2_HI; clo = FMRX_PIO2_LO;
    } else if (kk == 3) {
        chi = -FMRX_PIO2_HI; clo = -FMRX_PIO2_LO;
    } else if (t.r > 0.0) {
        chi = -FMRX_PI_HI; clo = -FMRX_PI_LO;
    } else {
        chi = FMRX_PI_HI; clo = FMRX_PI_LO;
    }
    const double phi = p_add(p_add(t.r, clo), chi);
    const double alpha = p_add(phi, delta);
    *out = -alpha;
    // guard: roundings must be tiny rotations, and stay clear of the +-pi seam
    return fabs(cross) < 0x1p-20 && fabs(dotc) < 0x1p-20 && fabs(phi) < 3.125;
}

#if defined(__CUDA_ARCH__)
FMRX_HD double ref_atan2(double y, double x) { return atan2(y, x); }
#else
FMRX_HD double ref_atan2(double y, double x) { return atan2(y, x); }
#endif

// One step.  x = pilot sample, inv_x = 1.0/(double)x (computed off the chain).
// trig_valid (in/out): c.trig describes the trigArg that fi, fq were rounded from.
// Returns the float trigArg of this step; *slow counts guard rejections.
FMRX_HD float chain_step(Chain &c, const Consts &k, float x, double inv_x, bool &trig_valid,
                         unsigned *slow)
{
    const float ei = p_fmulf(x, c.fi);                               // :159
    const float eq = p_fmulf(x, -c.fq);                              // :160
    double ang;
    if (!(trig_valid && atan2_shortcut(c.trig, x, inv_x, ei, eq, &ang))) {
        ang = ref_atan2((double)eq, (double)ei);                     // :161
        if (slow)
            ++*slow;
    }
    const float ed = p_d2f(ang);
    c.integ = p_faddf(c.integ, p_fmulf(k.ki, ed));                   // :163
    c.ph = p_faddf(c.ph, p_faddf(p_fmulf(k.kp, ed), c.integ));       // :164
    c.toff = p_faddf(c.toff, 1.0f);                                  // :166
    c.ta = p_d2f(p_add(p_mul(k.w, (double)c.toff), (double)c.ph));   // :167
    if (fabsf(c.ta) <= FMRX_FAST_TRIG_MAX) {
        c.trig = sincos_f32arg((double)c.ta);
        c.fi = p_d2f(c.trig.cs);                                     // :168
        c.fq = p_d2f(c.trig.sn);                                     // :169
        trig_valid = true;
    } else {   // beyond the exact-reduction range (or NaN): library sin/cos
        c.fi = p_d2f(cos((double)c.ta));
        c.fq = p_d2f(sin((double)c.ta));
        trig_valid = false;
    }
    return c.ta;
}

}  // namespace pllcore
