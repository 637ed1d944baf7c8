// pll_model.cpp -- HOST build of the device PLL step (csrc/fmrx_pll_core.h) for the
// CPU tests: lets the low-latency sincos / atan2 formulation be compared with the
// oracle bit for bit on long signals without a GPU.  Test infrastructure only; the
// product library never contains a host PLL.
// Build: g++ -O2 -ffp-contract=off [-mfma] -shared -fPIC (tests/conftest.py).
#include <cstddef>
#include <cstring>
#include "fmrx_pll_core.h"

using namespace pllcore;

extern "C" {

// state5 = {integrator, phaseEst, feedbackI, feedbackQ, trigOffset}
int pll_model_run(const float *pilot, int n, float freq, float Fs, float bw, float *state5,
                  float *trig_out, unsigned *slow_steps)
{
    Consts k;
    k.kp = bw * 2.666f;
    k.ki = (bw * bw) * 3.555f;
    k.w = (2.0 * 3.14159265358979323846) * (double)(freq / Fs);
    Chain c;
    memset(&c, 0, sizeof(c));
    c.integ = state5[0]; c.ph = state5[1]; c.fi = state5[2]; c.fq = state5[3]; c.toff = state5[4];
    chain_load(c, k);
    const TrigK K = trig_constants();
    unsigned slow = 0;
    // mirrors k_pll: groups of 32 steps run speculatively from a checkpoint; a group
    // with a failed guard is redone step by step
    for (int base = 0; base < n; base += 32) {
        const int cnt = n - base < 32 ? n - base : 32;
        const Chain ck = c;
        bool good = toff_is_regular(c.toff);
        if (good) {
            for (int t = 0; t < cnt; t++) {
                const StepIn in = step_inputs(k, pilot[base + t], toff_after(ck.toff, t + 1));
                good &= chain_step_spec(c, k, K, in);
                if (trig_out)
                    trig_out[base + t] = (float)c.tad;
            }
        }
        if (!good) {
            c = ck;
            for (int t = 0; t < cnt; t++) {
                const float ta = chain_step(c, k, K, pilot[base + t], &slow);
                if (trig_out)
                    trig_out[base + t] = ta;
            }
        }
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    if (slow_steps)
        *slow_steps = slow;
    return 0;
}

// float(sin), float(cos) of float arguments through the device formulation
void pll_model_sincos(const float *x, int n, float *s, float *c)
{
    for (int i = 0; i < n; i++) {
        double sr, cr, r, nd, sn, cs;
        sincos_reduced(trig_constants(), (double)x[i], sr, cr, r, nd);
        rotate_quadrant((int)nd, sr, cr, sn, cs);
        s[i] = (float)sn;
        c[i] = cos_of_float(x[i]);
    }
}

void pll_model_sincos_d(const float *x, int n, double *s, double *c)
{
    for (int i = 0; i < n; i++) {
        double sr, cr, r, nd;
        sincos_reduced(trig_constants(), (double)x[i], sr, cr, r, nd);
        rotate_quadrant((int)nd, sr, cr, s[i], c[i]);
    }
}


// The run-ahead predictor against the exact recurrence: from the exact state at the start
// of each group of `group` samples, step the predictor and the exact chain side by side
// and record, per group, the largest distance (in float grid steps of trigArg) between the
// exact trigArg and the grid point the predictor would centre the candidate table on.
// state5 as in pll_model_run; worst[n_groups].
int pll_model_predict(const float *pilot, int n, float freq, float Fs, float bw, float *state5, int group, int *worst, long long *hist)
{
    Consts k;
    k.kp = bw * 2.666f;
    k.ki = (bw * bw) * 3.555f;
    k.w = (2.0 * 3.14159265358979323846) * (double)(freq / Fs);
    const TrigK K = trig_constants();
    Chain c;
    memset(&c, 0, sizeof(c));
    c.integ = state5[0]; c.ph = state5[1]; c.fi = state5[2]; c.fq = state5[3]; c.toff = state5[4];
    chain_load(c, k);
    int ng = 0;
    for (int b = 0; b < n; b += group, ng++) {
        float pi = c.integ, pp = c.ph;          // the predictor restarts from the exact state
        int w = 0;
        for (int j = 0; j < group && b + j < n; j++) {
            const float x = pilot[b + j];
            predictor_step(k, predictor_c(k, x, c.toff), pi, pp);
            const float ta = chain_step(c, k, K, x, nullptr);
            const double v = k.w * (double)c.toff;      // trigOffset after the step
            int e;
            (void)frexpf(fabsf(ta), &e);
            const double ulp = ldexp(1.0, e - 1 - 23);
            const double d = fabs((double)ta / ulp - nearbyint((v + (double)pp) / ulp));
            if (ta != 0.0f && d > w)
                w = d > 1e9 ? 1000000000 : (int)d;
            if (hist && ng >= 1 && ta != 0.0f) {
                const double sd = (double)ta / ulp - nearbyint((v + (double)pp) / ulp);
                int b = (int)sd + 8;
                hist[b < 0 ? 0 : b > 16 ? 16 : b]++;
            }
        }
        worst[ng] = w;
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    return ng;
}

}  // extern "C"

// feedbackI/Q implied by (phaseEst, trigOffset): fl32(cos, sin)(fl32(w*toff + ph))
extern "C" void pll_model_feedback(float freq, float Fs, float ph, float toff, float *fi, float *fq)
{
    Consts k;
    k.kp = k.ki = 0.0f;
    k.w = (2.0 * 3.14159265358979323846) * (double)(freq / Fs);
    Chain c;
    memset(&c, 0, sizeof(c));
    c.ph = ph;
    c.toff = toff;
    c.fi = 1.0f;
    chain_load(c, k);
    if (fabsf((float)c.tad) <= FMRX_FAST_TRIG_MAX) {
        chain_refresh(c);
        *fi = c.fi;
        *fq = c.fq;
    } else {
        *fi = (float)cos(c.tad);
        *fq = (float)sin(c.tad);
    }
}
