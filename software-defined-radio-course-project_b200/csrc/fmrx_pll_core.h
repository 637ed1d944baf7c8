// fmrx_pll_core.h -- the arithmetic core of one PLL step (src/filter.cpp:157-171),
// written once for the device (k_pll) and for a host build used only by the CPU
// tests (tests/pll_model.cpp), so the low-latency formulation can be checked bit
// for bit against the oracle on hours of signal without a GPU.
//
// Why a special formulation.  Per IF sample the reference evaluates, in double,
// atan2(eQ,eI), cos(trigArg) and sin(trigArg) and rounds each to float.  The
// recurrence is one dependent chain per capture, so its LATENCY (and, with one
// warp per capture, its instruction count) is the throughput bound of the whole
// receive chain.  Measured on B200 (tools/ubench_latency*.cu): DFMA/DADD/DMUL
// 8.5 cycles, FP32 4.4, each f32<->f64 conversion 18, DSETP+select 12.  Stock
// libdevice sin/cos fall into the Payne-Hanek slow path once |trigArg| > 105615
// (0.9 s into a capture) and stock atan2 is a division plus a degree-19
// polynomial: 760 cycles per sample.  The formulation here takes about 300 as a
// sequential step (chain_step_spec: what k_pll falls back to), and -- the point of
// it -- splits into a part that hangs off the previous trigArg alone (make_feedback,
// error_from_feedback: evaluated by k_pll's candidate warps for a few hypothetical
// trigArgs ahead of time) and the loop filter (four float additions) that k_pll's
// warp 0 is left with:
//
//  * sincos: trigArg is a FLOAT (24-bit significand, |x| <= 2^24) promoted to
//    double, so a 3-term Cody-Waite reduction with 29/29/53-bit pieces of pi/2 is
//    exact in its first two steps (n < 2^24: n*P1 and n*P2 are exact products) and
//    rounds once; the kernels are the fdlibm minimax polynomials, Estrin-ordered.
//  * no quadrant rotation on the chain: the feedback pair is kept UNROTATED,
//    (cf, sf) = fl32(cos r, sin r) of the reduced argument r.  The reference's
//    fI, fQ are a signed permutation of that pair and float multiplication commutes
//    with signed permutations, so eI, eQ are the same signed permutation of
//    fl(x*cf), fl(x*(-sf)) and atan2 only needs the quadrant added back as an angle.
//  * atan2(eQ,eI) = -(phi + delta): phi = r + j*pi/2 is the previous trigArg wrapped
//    to (-pi, pi] (j = quadrant count mod 4 in -2..2, shifted by 2 when the pilot
//    sample is negative), and delta is the tiny rotation caused by the float
//    roundings, obtained exactly from FMA residuals t_i = eI - x*cos r,
//    t_q = eQ + x*sin r:
//        cross = -(cos r*t_q + sin r*t_i)/x,  dot = (cos r*t_i - sin r*t_q)/x,
//        delta = cross*(1 - dot)            (|cross|,|dot| <~ 2^-23; O(2^-69) dropped)
//  * trigArg = fl32(w*toff + ph) is rounded inside double by the add-magic trick
//    for the binade trigArg currently lives in (2 DADD instead of 2 conversions).
//  * every assumption (x normal, roundings tiny, phi clear of the +-pi seam, binade
//    unchanged) is a guard evaluated OFF the dependent chain; a step whose guard
//    fails commits nothing and is redone by the generic step (the reference's
//    statements with the library atan2): ~25 of 2.4 million steps.
//
// Both formulations evaluate the same real functions the reference does, to ~1 ulp
// of double, and round to float where the reference rounds; they differ from
// glibc's result only when the exact value lies within ~1e-16 relative of a float
// rounding boundary (~1e-8 per sample), the same class of event as any other libm.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FMRX_HD __device__ __forceinline__
#else
#define FMRX_HD static inline
#endif

namespace pllcore {

// ---- exact-rounding primitives ---------------------------------------------
#if defined(__CUDACC__)
FMRX_HD double p_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
FMRX_HD double p_mul(double a, double b) { return __dmul_rn(a, b); }
FMRX_HD double p_add(double a, double b) { return __dadd_rn(a, b); }
FMRX_HD float p_fmulf(float a, float b) { return __fmul_rn(a, b); }
FMRX_HD float p_faddf(float a, float b) { return __fadd_rn(a, b); }
FMRX_HD float p_fmaf(float a, float b, float c) { return __fmaf_rn(a, b, c); }
FMRX_HD float p_d2f(double a) { return __double2float_rn(a); }
FMRX_HD int p_hi32(double a) { return __double2hiint(a); }
#else
// host build: compiled with -ffp-contract=off, so * and + are single IEEE ops
FMRX_HD double p_fma(double a, double b, double c) { return fma(a, b, c); }
FMRX_HD double p_mul(double a, double b) { return a * b; }
FMRX_HD double p_add(double a, double b) { return a + b; }
FMRX_HD float p_fmulf(float a, float b) { return a * b; }
FMRX_HD float p_faddf(float a, float b) { return a + b; }
FMRX_HD float p_fmaf(float a, float b, float c) { return fmaf(a, b, c); }
FMRX_HD float p_d2f(double a) { return (float)a; }
FMRX_HD int p_hi32(double a)
{
    uint64_t u;
    memcpy(&u, &a, sizeof(u));
    return (int)(uint32_t)(u >> 32);
}
#endif

// ---- constants ---------------------------------------------------------------
#define FMRX_FAST_TRIG_MAX 16777216.0f     /* |x| <= 2^24 */
#define FMRX_SEAM_PHI_MAX 3.1415916        /* pi - 1.05e-6: how close to the +-pi seam the atan2 shortcut goes */
#define FMRX_RINT_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

// pi/2 = P1 + P2 + P3: 29 + 29 + 53 bits (n*P1, n*P2 exact for |n| < 2^24, which
// covers every float argument |x| <= 2^24).  The constants travel as a struct so the
// device loop can pin them in registers once (k_pll) instead of re-loading 64-bit
// immediates / constant-bank words every step.
struct TrigK {
    double two_over_pi;          // 0x1.45f306dc9c883p-1
    double p1, p2, p3;           // 0x1.921fb54000000p+0, 0x1.10b4611000000p-30, 0x1.4c4c6628b80dcp-59
    double s1, s2, s3, s4, s5, s6;   // fdlibm __kernel_sin
    double c1, c2, c3, c4, c5, c6;   // fdlibm __kernel_cos
    double pio2_hi, pio2_lo;
};
#define FMRX_TRIGK_INIT                                                                  \
    {                                                                                    \
        0.6366197723675814, 1.570796325802803, 9.920935774287987e-10,                    \
            2.2517417741562176e-18, -1.66666666666666324348e-01,                         \
            8.33333333332248946124e-03, -1.98412698298579493134e-04,                     \
            2.75573137070700676789e-06, -2.50507602534068634195e-08,                     \
            1.58969099521155010221e-10, 4.16666666666666019037e-02,                      \
            -1.38888888888741095749e-03, 2.48015872894767294178e-05,                     \
            -2.75573143513906633035e-07, 2.08757232129817482790e-09,                     \
            -1.13596475577881948265e-11, 1.5707963267948966, 6.123233995736766e-17       \
    }
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
static inline
#endif
TrigK trig_constants()
{
    const TrigK k = FMRX_TRIGK_INIT;
    return k;
}

// sin r, cos r of the reduced argument and the quadrant count nd (an integer held
// in a double) of a float-valued argument |x| <= 2^24: x = r + nd*pi/2.
FMRX_HD void sincos_reduced(const TrigK &K, double x, double &sn_r, double &cs_r, double &r_out,
                            double &nd_out)
{
    // nd = rint(x * 2/pi) by the add-magic trick (fused: one rounding).  A
    // tie-adjacent miss only moves |r| a hair past pi/4.
    const double qm = p_fma(x, K.two_over_pi, FMRX_RINT_MAGIC);
    const double nd = p_add(qm, -FMRX_RINT_MAGIC);
    double r = p_fma(-nd, K.p1, x);         // exact
    r = p_fma(-nd, K.p2, r);                // exact product, one rounding
    r = p_fma(-nd, K.p3, r);
    const double z = p_mul(r, r);
    const double z2 = p_mul(z, z);
    const double z4 = p_mul(z2, z2);
    // sin r = r + r*z*(S1 + z*S2 + ... + z^5*S6), Estrin order
    const double s12 = p_fma(K.s2, z, K.s1);
    const double s34 = p_fma(K.s4, z, K.s3);
    const double s56 = p_fma(K.s6, z, K.s5);
    double ps = p_fma(s34, z2, s12);
    ps = p_fma(s56, z4, ps);
    sn_r = p_fma(p_mul(z, r), ps, r);
    // cos r = 1 - z/2 + z^2*(C1 + z*C2 + ... + z^5*C6)
    const double c12 = p_fma(K.c2, z, K.c1);
    const double c34 = p_fma(K.c4, z, K.c3);
    const double c56 = p_fma(K.c6, z, K.c5);
    double pc = p_fma(c34, z2, c12);
    pc = p_fma(c56, z4, pc);
    cs_r = p_fma(z2, pc, p_fma(-0.5, z, 1.0));
    r_out = r;
    nd_out = nd;
}

// Rotate (sin r, cos r) by n quadrants (used off the chain only).
template <class T> FMRX_HD void rotate_quadrant(int n, T sn_r, T cs_r, T &sn, T &cs)
{
    const int q = n & 3;
    const T a = (q & 1) ? cs_r : sn_r;
    const T b = (q & 1) ? sn_r : cs_r;
    sn = (q & 2) ? -a : a;
    cs = ((q + 1) & 2) ? -b : b;
}

// low word of (nd + 1.5 * 2^52) for an integer-valued double |nd| < 2^31: nd as an int, without a conversion instruction
FMRX_HD int quadrant_of(double nd)
{
    const double q = p_add(nd, FMRX_RINT_MAGIC);
#if defined(__CUDACC__)
    return __double2loint(q);
#else
    uint64_t u;
    memcpy(&u, &q, sizeof(u));
    return (int)(uint32_t)u;
#endif
}

// cos of a float, rounded to float, for the NCO output (src/filter.cpp:170): same
// reduction; arguments beyond its range take the library cos.
FMRX_HD float cos_of_float(float a)
{
    if (fabsf(a) <= FMRX_FAST_TRIG_MAX) {
        double sn_r, cs_r, r, nd, sn, cs;
        const TrigK K = trig_constants();
        sincos_reduced(K, (double)a, sn_r, cs_r, r, nd);
        // the quadrant count from the low word of nd + magic (an exact sum: nd is an integer below 2^24) instead of a
        // double->int conversion: k_audio evaluates this once per IF sample and its conversion pipe is what is full
        rotate_quadrant(quadrant_of(nd), sn_r, cs_r, sn, cs);
        return p_d2f(cs);
    }
    return p_d2f(cos((double)a));
}

// r + j*pi/2 for the quadrant count m (an integer in a double), with
// j = m - 4*rint(m/4) in -2..2: the argument wrapped to (-pi, pi].  For m = 2 (mod 4)
// the tie of the rint goes to even, i.e. to either side: the half turn is then taken with
// the sign that keeps the result inside the interval (+pi + r for r < 0, -pi + r for
// r > 0).  Only r == 0 there (trigArg an exact odd multiple of pi: the seam itself) is
// left to the guard of error_from_feedback.  An unlocked loop visits every quadrant; a
// locked one stays at m = 0.
FMRX_HD double wrapped_angle(const TrigK &K, double m, double r)
{
    const double t = p_add(p_fma(m, 0.25, FMRX_RINT_MAGIC), -FMRX_RINT_MAGIC);
    const double jd = p_fma(-4.0, t, m);        // exact
    // both signs of the half turn are evaluated beside each other and one select follows r (this
    // is on the dependent chain of the exact step); same sign bit of j and r = the wrong side
    const double same = p_fma(jd, K.pio2_hi, p_fma(jd, K.pio2_lo, r));
    const double other = p_fma(-jd, K.pio2_hi, p_fma(-jd, K.pio2_lo, r));
    const bool flip = fabs(jd) == 2.0 && (p_hi32(jd) ^ p_hi32(r)) >= 0;
    return flip ? other : same;
}

struct Consts {
    float kp, ki;
    double w;        // (2*PI)*(double)(freq/Fs)
};

// PLL state carried between steps: the reference's floats plus the double
// leftovers of the last sincos (a pure function of trigOffset and phaseEst: they
// are recomputed, never stored, when a state is loaded).
struct Chain {
    float integ, ph, toff;
    double tad;          // last trigArg (a float value, held in a double)
    float fi, fq;        // the reference's feedbackI/Q; maintained by the generic
                         // step only -- read them through chain_feedback()
    float cf, sf;        // fl32(cos r), fl32(sin r): the unrotated feedback pair
    double cr, sr;       // cos r, sin r
    double r, nd;        // ta = r + nd*pi/2
    // trigArg = fl32(w*toff + ph): while |s| stays in the binade whose exponent field
    // is `binade` (high word of 2^e) the float grid there has spacing ulp = 2^(e-23),
    // so the float rounding is G = rint(s/ulp) (add-magic trick, low word = G) and
    // trigArg = G*ulp, all exact in double.  FMRX_DISARMED switches the fast step off:
    // the generic step runs and re-arms it.
    double ulp, inv_ulp;
    unsigned binade;
};

#define FMRX_DISARMED 0x80000000u

FMRX_HD void chain_arm(Chain &c)
{
    const float at = fabsf(p_d2f(c.tad));
    if (at >= 0x1p-60f && at < FMRX_FAST_TRIG_MAX) {
        int e;
        (void)frexpf(at, &e);                 // at = m * 2^e, m in [0.5, 1)
        c.binade = (unsigned)(e - 1 + 1023) << 20;
        c.ulp = ldexp(1.0, e - 1 - 23);
        c.inv_ulp = ldexp(1.0, 23 - (e - 1));
    } else {
        c.binade = FMRX_DISARMED;
        c.ulp = c.inv_ulp = 0.0;
    }
}

// The reference's feedbackI, feedbackQ of the current state.
FMRX_HD void chain_feedback(const Chain &c, float &fi, float &fq)
{
    if (c.binade != FMRX_DISARMED) {
        rotate_quadrant((int)c.nd, c.sf, c.cf, fq, fi);
    } else {
        fi = c.fi;
        fq = c.fq;
    }
}

// Recompute the sincos leftovers for c.ta (|ta| <= 2^24) and arm the fast step.
FMRX_HD void chain_refresh(Chain &c)
{
    const TrigK K = trig_constants();
    sincos_reduced(K, c.tad, c.sr, c.cr, c.r, c.nd);
    c.sf = p_d2f(c.sr);
    c.cf = p_d2f(c.cr);
    rotate_quadrant((int)c.nd, c.sf, c.cf, c.fq, c.fi);              // :168-169
    chain_arm(c);
}

// Load (integ, ph, fi, fq, toff) and rebuild the derived fields.  A state written
// by this code (or the reference's initial state fi=1, fq=0, toff=ph=0) satisfies
// fi, fq = fl32(cos, sin)(fl32(w*toff + ph)); anything else disarms the fast step
// for one sample.
FMRX_HD void chain_load(Chain &c, const Consts &k)
{
    const float fi = c.fi, fq = c.fq;
    const float ta = p_d2f(p_add(p_mul(k.w, (double)c.toff), (double)c.ph));
    c.tad = (double)ta;
    c.cr = c.sr = c.r = c.nd = c.ulp = c.inv_ulp = 0.0;
    c.cf = c.sf = 0.0f;
    c.binade = FMRX_DISARMED;
    if (!(fabsf(ta) <= FMRX_FAST_TRIG_MAX))
        return;
    chain_refresh(c);
    if (!(c.fi == fi && c.fq == fq)) {
        c.fi = fi;
        c.fq = fq;
        c.binade = FMRX_DISARMED;
    }
}

// The generic step: the reference's statements one by one (library atan2).
FMRX_HD void chain_step_generic(Chain &c, const Consts &k, float x)
{
    float fi, fq;
    chain_feedback(c, fi, fq);
    const float ei = p_fmulf(x, fi);                                 // :159
    const float eq = p_fmulf(x, -fq);                                // :160
    const float ed = p_d2f(atan2((double)eq, (double)ei));           // :161
    c.integ = p_faddf(c.integ, p_fmulf(k.ki, ed));                   // :163
    c.ph = p_faddf(c.ph, p_faddf(p_fmulf(k.kp, ed), c.integ));       // :164
    c.toff = p_faddf(c.toff, 1.0f);                                  // :166
    const float ta = p_d2f(p_add(p_mul(k.w, (double)c.toff), (double)c.ph));   // :167
    c.tad = (double)ta;
    if (fabsf(ta) <= FMRX_FAST_TRIG_MAX) {
        chain_refresh(c);
    } else {   // beyond the exact-reduction range (or NaN): library sin/cos
        c.fi = p_d2f(cos(c.tad));                                    // :168
        c.fq = p_d2f(sin(c.tad));                                    // :169
        c.binade = FMRX_DISARMED;
    }
}

// Per-sample inputs that do not depend on the recurrence: prepared off the chain
// (on the device: one lane per sample, 32 at a time).
struct StepIn {
    float x;         // pilot sample
    double xd;       // (double)x
    double inv_x;    // 1.0 / (double)x, IEEE divide
    double turn;     // 2.0 if x < 0 else 0.0: half a turn, in quadrants
    double v;        // w * (double)trigOffset_after_this_step (:166-167)
};

FMRX_HD StepIn step_inputs(const Consts &k, float x, float toff_after)
{
    StepIn in;
    in.x = x;
    in.xd = (double)x;
    in.inv_x = 1.0 / in.xd;
    in.turn = (x < 0.0f) ? 2.0 : 0.0;
    in.v = p_mul(k.w, (double)toff_after);
    return in;
}

// What the atan2 shortcut and the float products of ONE sample need from the trigArg
// that precedes it: the unrotated feedback pair and its double leftovers, already
// combined with that sample's off-chain inputs (1/x and its half-turn).
struct Feedback {
    float cf, sf;        // fl32(cos r), fl32(sin r)
    double cr, sr;       // cos r, sin r
    double csx, snx;     // cos r / x, sin r / x
    double phi;          // trigArg wrapped to (-pi, pi], half a turn more when x < 0
};

// Everything a sample needs from the preceding trigArg `tad` (a float value in a
// double, |tad| <= 2^24).  On the device every lane evaluates this for a different
// CANDIDATE trigArg before the loop filter has decided which one it is.
FMRX_HD Feedback make_feedback(const TrigK &K, double tad, double turn, double inv_x, double *r_out,
                               double *nd_out)
{
    Feedback f;
    double r, nd;
    sincos_reduced(K, tad, f.sr, f.cr, r, nd);
    f.sf = p_d2f(f.sr);                                              // :168-169 (up to the quadrant)
    f.cf = p_d2f(f.cr);
    f.phi = wrapped_angle(K, p_add(nd, turn), r);
    f.csx = p_mul(f.cr, inv_x);
    f.snx = p_mul(f.sr, inv_x);
    if (r_out) {
        *r_out = r;
        *nd_out = nd;
    }
    return f;
}

// The phase-detector output errorD = fl32(atan2(eQ, eI)) of one sample (:159-161) from
// the feedback prepared for it.  `ok` is cleared if a guard of the shortcut fails.
FMRX_HD float error_from_feedback(const Feedback &f, float x, double xd, bool &ok)
{
    const float ei = p_fmulf(x, f.cf);                               // :159 (up to the quadrant)
    const float eq = p_fmulf(x, -f.sf);                              // :160
    const double ti = p_fma(-xd, f.cr, (double)ei);   // exact rounding residuals, rounded once
    const double tq = p_fma(xd, f.sr, (double)eq);
    const double m1 = p_mul(f.csx, tq);
    const double cross = -p_fma(f.snx, ti, m1);                // rotation by the roundings
    const double dotc = p_fma(-f.snx, tq, p_mul(f.csx, ti));   // radial part (second order)
    const double a1 = p_fma(-f.snx, ti, p_add(f.phi, -m1));    // phi + cross
    const double alpha = p_fma(-cross, dotc, a1);              // phi + cross*(1 - dotc)
    // the result must not wrap at the +-pi seam: |alpha| <= |phi| + |cross| (1 + |dotc|) < pi
    // with |cross| < 2^-20 = 9.54e-7 and |phi| < pi - 1.05e-6
    ok = ok && fabs(cross) < 0x1p-20 && fabs(dotc) < 0x1p-20 && fabs(f.phi) < FMRX_SEAM_PHI_MAX;
    return p_d2f(-alpha);                                            // :161
}

// The loop filter (:163-164) and the double sum the reference then stores to a float
// (:167): updates integ, ph; returns s = w*trigOffset + phaseEst.
FMRX_HD double filter_step(const Consts &k, float ed, double v, float &integ, float &ph, double &phd)
{
    integ = p_faddf(integ, p_fmulf(k.ki, ed));                       // :163
    ph = p_faddf(ph, p_faddf(p_fmulf(k.kp, ed), integ));             // :164
    phd = (double)ph;
    return p_add(v, phd);                                            // :167 in double
}

// atan2 + loop filter of one sample (:159-167).
FMRX_HD double step_front(const Consts &k, const Feedback &f, float x, double xd, double v, float &integ,
                          float &ph, double &phd, bool &ok)
{
    const float ed = error_from_feedback(f, x, xd, ok);
    return filter_step(k, ed, v, integ, ph, phd);
}

// rint(s/ulp) as a double holding an integer plus the magic constant: its low word is
// the grid index G (two's complement), (q - magic)*ulp is the float-rounded value.
FMRX_HD double grid_round(double s, double inv_ulp) { return p_fma(s, inv_ulp, FMRX_RINT_MAGIC); }
FMRX_HD int grid_index(double q)
{
#if defined(__CUDACC__)
    return __double2loint(q);
#else
    uint64_t u;
    memcpy(&u, &q, sizeof(u));
    return (int)(uint32_t)u;
#endif
}
FMRX_HD double grid_value(double q, double ulp) { return p_mul(p_add(q, -FMRX_RINT_MAGIC), ulp); }
FMRX_HD bool in_binade(double s, unsigned binade)
{
    return (((unsigned)p_hi32(s) & 0x7fffffffu) - binade) < 0x00100000u;
}

// The speculative fast step, composed sequentially: ALWAYS updates the state and
// returns whether every guard held.  When it returns false the state is garbage and
// the caller must restore a checkpoint.  (k_pll runs the same pieces with the
// make_feedback() of the NEXT sample evaluated by the 32 lanes for 32 candidate
// trigArgs ahead of time; the values it ends up using are exactly these.)
FMRX_HD bool chain_step_spec(Chain &c, const Consts &k, const TrigK &K, const StepIn &in)
{
    Feedback f;
    f.cf = c.cf; f.sf = c.sf; f.cr = c.cr; f.sr = c.sr;
    f.phi = wrapped_angle(K, p_add(c.nd, in.turn), c.r);
    f.csx = p_mul(c.cr, in.inv_x);
    f.snx = p_mul(c.sr, in.inv_x);
    bool ok = true;
    double phd;
    const double s = step_front(k, f, in.x, in.xd, in.v, c.integ, c.ph, phd, ok);
    c.toff = p_faddf(c.toff, 1.0f);                                  // :166
    ok = ok && in_binade(s, c.binade);
    c.tad = grid_value(grid_round(s, c.inv_ulp), c.ulp);             // :167 stored to float
    sincos_reduced(K, c.tad, c.sr, c.cr, c.r, c.nd);
    c.sf = p_d2f(c.sr);                                              // :168-169 (up to the quadrant)
    c.cf = p_d2f(c.cr);
    return ok;
}

// The checked fast step: commits only if every guard held.
FMRX_HD bool chain_step_fast(Chain &c, const Consts &k, const TrigK &K, const StepIn &in)
{
    Chain t = c;
    if (!chain_step_spec(t, k, K, in))
        return false;
    c = t;
    return true;
}

// One checked step; *slow counts steps that took the generic path.  Returns trigArg.
FMRX_HD float chain_step(Chain &c, const Consts &k, const TrigK &K, float x, unsigned *slow)
{
    const StepIn in = step_inputs(k, x, p_faddf(c.toff, 1.0f));
    if (!chain_step_fast(c, k, K, in)) {
        chain_step_generic(c, k, x);
        if (slow)
            ++*slow;
    }
    return p_d2f(c.tad);
}

// ---- the run-ahead predictor ------------------------------------------------------------
//
// The phase detector (:159-161) computes atan2(x*(-sin t), x*cos t) with t the previous
// trigArg: up to the float roundings of the feedback pair and of the two products that is
// wrap(pi*(x < 0) - t) onto (-pi, pi].  With t = w*trigOffset + phaseEst, the part that
// does not depend on the recurrence,
//     c = pi*(x < 0) - (w*trigOffset mod 2 pi),
// is prepared per sample, and one predictor step is a handful of float operations.  The predictor is never
// used for a result: it only says where to CENTRE the candidate table of a step, and it
// restarts from the exact (integrator, phaseEst) at every group.  While the loop is
// locked its phaseEst stays within a grid step of the exact one over a group
// (tests/test_pll_model.py::test_predictor_tracks_the_exact_recurrence).
FMRX_HD float predictor_c(const Consts &k, float x, float toff_before)
{
    const double vp = p_mul(k.w, (double)toff_before);
    const double kq = p_add(p_add(p_mul(vp, 0.15915494309189535), FMRX_RINT_MAGIC), -FMRX_RINT_MAGIC);
    const double rp = p_fma(-kq, 2.4492935982947064e-16, p_fma(-kq, 6.283185307179586, vp));   // vp mod 2 pi, in [-pi, pi]
    return p_d2f(p_add(x < 0.0f ? 3.141592653589793 : 0.0, -rp));
}

// integ' = integ + Ki e and ph' = ph + Kp e + integ' = (ph + integ) + (Kp + Ki) e, with
// e = d - 2 pi k, d = c - ph, k = rint(d / 2 pi): arranged so that only five float
// operations depend on each other from one phaseEst to the next (d, d/2pi, two for the
// rint, one FMA); the roundings differ from the reference's, which is all the same to a
// predictor.
FMRX_HD void predictor_step(const Consts &k, float c, float &integ, float &ph)
{
    const float kpi = p_faddf(k.kp, k.ki);
    const float d = p_faddf(c, -ph);
    const float phi = p_faddf(ph, integ);
    const float a = p_fmaf(kpi, d, phi);                             // beside the rint
    const float kq = p_faddf(p_faddf(p_fmulf(d, 0.15915494f), 12582912.0f), -12582912.0f);   // rint(d / 2 pi)
    ph = p_fmaf(-kq, p_fmulf(kpi, 6.2831855f), a);                   // :164
    integ = p_fmaf(k.ki, p_fmaf(-kq, 6.2831855f, d), integ);         // :163
}

// ---- the one-hypothesis ("1H") scheme: an exact-ORDER predictor --------------------------------
//
// Where phaseEst's float grid is coarse (|phaseEst| in the hundreds and beyond: modes 2/3, where the
// reference hands the PLL if_fs*interp and phaseEst cancels w*trigOffset; any loop that locks onto something
// other than the pilot, or onto nothing, and runs away) the loop filter (:163-164) rounds phaseEst so
// coarsely that an errorD good to ~1e-7 reproduces phaseEst BIT FOR BIT almost always.  So k_pll's predictor
// warp then runs the recurrence in the reference's own operation order with errorD = wrap(pi*(x < 0) - trigArg)
// -- what atan2(x*(-sin t), x*cos t) is up to the float roundings of the feedback pair -- in float; the
// candidate warps evaluate the exact phase detector for the trigArg that follows from each PREDICTED
// phaseEst; warp 0 runs the exact loop filter on those errorDs and accepts a block iff its phaseEst
// matched the prediction at every step.  By induction an accepted block IS the reference's trajectory.
//
// Large arguments never meet float arithmetic: per sample the I/O warps prepare (in double) a float P near
// the phaseEst to come, S = w*trigOffset + P split into B = fl32(S) and r = fl32(S - B), and
// c = fl32(wrap(pi*(x_next < 0) - B)).  Then with d = phaseEst (-) P (exact: the two are within a factor of two)
//     trigArg = fl32(w*trigOffset + phaseEst) = fl32(S + d) = B (+) (r (+) d)     (up to a double rounding)
//     z = trigArg (-) B                                    (exact; small while P tracks phaseEst)
//     errorD of the next sample ~ wrap(c - z).
struct OneHypIn {
    float P, r, B, c;
};

FMRX_HD double wrap_pm_pi(double a)
{
    const double kq = p_add(p_add(p_mul(a, 0.15915494309189535), FMRX_RINT_MAGIC), -FMRX_RINT_MAGIC);
    return p_fma(-kq, 2.4492935982947064e-16, p_fma(-kq, 6.283185307179586, a));
}

// v = w*trigOffset after the step; ph0 a phaseEst `ahead` steps before this one and slope its mean advance per
// step lately; x_next the pilot sample the resulting angle is for.
FMRX_HD OneHypIn onehyp_inputs(double v, float ph0, float slope, int ahead, float x_next)
{
    OneHypIn o;
    o.P = p_d2f(p_fma((double)slope, (double)ahead, (double)ph0));
    const double S = p_add(v, (double)o.P);
    o.B = p_d2f(S);
    o.r = p_d2f(p_add(S, -(double)o.B));
    o.c = p_d2f(wrap_pm_pi(p_add(x_next < 0.0f ? 3.141592653589793 : 0.0, -(double)o.B)));
    return o;
}

// the angle for the first sample of a group, from the exact trigArg before it
FMRX_HD float onehyp_first_angle(float x, double tad)
{
    return p_d2f(wrap_pm_pi(p_add(x < 0.0f ? 3.141592653589793 : 0.0, -tad)));
}

// One predictor step: `a` is the angle (the approximate errorD) for this sample; returns the one for the next,
// NOT reduced: while P tracks phaseEst and the loop is anywhere near lock it lies in [-pi, pi] as it is.  The
// caller keeps the largest |angle| of a block of steps and, if that exceeded pi, runs the block again with
// onehyp_predictor_step_reduced (a reduction on the dependent chain of every step would cost the common case
// a third of its speed: k_pll's predictor warp sets the pace of a one-hypothesis group).
FMRX_HD float onehyp_predictor_step(const Consts &k, const OneHypIn &in, float a, float &integ, float &ph)
{
    integ = p_faddf(integ, p_fmulf(k.ki, a));                        // :163
    ph = p_faddf(ph, p_faddf(p_fmulf(k.kp, a), integ));              // :164
    const float d = p_faddf(ph, -in.P);
    const float y = p_faddf(in.r, d);
    const float t = p_faddf(in.B, y);                                // :167 (trigArg)
    const float z = p_faddf(t, -in.B);
    return p_faddf(in.c, -z);
}

// The same step with the angle reduced to [-pi, pi]: c - z again with the whole turns taken out of z FIRST (z is
// exact; c - z as formed above has lost its low bits to the turns).  2 pi = 6.2831855f - 1.7484555e-07f.
FMRX_HD float onehyp_predictor_step_reduced(const Consts &k, const OneHypIn &in, float a, float &integ, float &ph)
{
    integ = p_faddf(integ, p_fmulf(k.ki, a));                        // :163
    ph = p_faddf(ph, p_faddf(p_fmulf(k.kp, a), integ));              // :164
    const float d = p_faddf(ph, -in.P);
    const float y = p_faddf(in.r, d);
    const float t = p_faddf(in.B, y);                                // :167 (trigArg)
    const float z = p_faddf(t, -in.B);
    const float an = p_faddf(in.c, -z);
    const float kq = p_faddf(p_faddf(p_fmulf(an, 0.15915494f), 12582912.0f), -12582912.0f);
    return p_fmaf(kq, 1.7484555e-07f, p_faddf(p_fmaf(-kq, 6.2831855f, -z), in.c));
}

// The short step, for groups in which trigArg is small and phaseEst's grid coarse (modes 2/3 once phaseEst has run
// away: trigArg stays near 1.7 while phaseEst is in the thousands): the predictor does not round trigArg to its
// float grid at all -- z = r + d instead of (B (+) (r (+) d)) (-) B, an error of at most ulp(trigArg)/2 = 2.4e-7 in
// the angle, 6e-9 in Kp*errorD, against a phaseEst spacing of 1.2e-4 and more -- which takes three of the nine
// dependent operations off its chain.  cr = c (-) r is formed off the chain.
FMRX_HD float onehyp_predictor_step_short(const Consts &k, float P, float cr, float a, float &integ, float &ph)
{
    integ = p_faddf(integ, p_fmulf(k.ki, a));                        // :163
    ph = p_faddf(ph, p_faddf(p_fmulf(k.kp, a), integ));              // :164
    return p_faddf(cr, -p_faddf(ph, -P));
}
// ... usable while |trigArg| < 4 (its float spacing at most 2.4e-7) and |phaseEst| >= 1024 (its spacing at least 1.2e-4)
FMRX_HD bool onehyp_short_ok(double tad, float ph) { return fabs(tad) < 4.0 && fabsf(ph) >= 1024.0f; }

#define FMRX_ONEHYP_PI 3.14159274f       /* fl32(pi): the largest |angle| the unreduced step may carry */

// trigArg as the reference forms it (:167): fl32(w*trigOffset + (double)phaseEst), held in a double
FMRX_HD double onehyp_trigarg(double v, float ph) { return (double)p_d2f(p_add(v, (double)ph)); }

// trigOffset advances by float additions of 1 (:166).  From an integer-valued start
// in [0, 2^24] the value after j steps is min(start + j, 2^24) exactly (2^24 + 1
// rounds back to 2^24: the counter saturates), which lets the device prepare
// StepIn::v for 32 steps at once.  Any other start takes the step-by-step path.
FMRX_HD bool toff_is_regular(float t) { return t >= 0.0f && t <= 16777216.0f && t == floorf(t); }
FMRX_HD float toff_after(float start, int steps)
{
    return fminf(start + (float)steps, 16777216.0f);
}

}  // namespace pllcore
