// pll_model.cpp -- HOST build of the device PLL step (csrc/fmrx_pll_core.h) for the
// CPU tests: lets the low-latency sincos / atan2 formulation be compared with the
// oracle bit for bit on long signals without a GPU.  Test infrastructure only; the
// product library never contains a host PLL.
// Build: g++ -O2 -ffp-contract=off [-mfma] -shared -fPIC (tests/conftest.py).
#include <cstddef>
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
    c.integ = state5[0]; c.ph = state5[1]; c.fi = state5[2]; c.fq = state5[3]; c.toff = state5[4];
    bool valid = chain_load(c, k);
    unsigned slow = 0;
    for (int i = 0; i < n; i++) {
        const float x = pilot[i];
        const double inv_x = 1.0 / (double)x;
        const float ta = chain_step(c, k, x, inv_x, valid, &slow);
        if (trig_out)
            trig_out[i] = ta;
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[2] = c.fi; state5[3] = c.fq; state5[4] = c.toff;
    if (slow_steps)
        *slow_steps = slow;
    return 0;
}

// float(sin), float(cos) of float arguments through the device formulation
void pll_model_sincos(const float *x, int n, float *s, float *c)
{
    for (int i = 0; i < n; i++) {
        const Trig t = sincos_f32arg((double)x[i]);
        s[i] = (float)t.sn;
        c[i] = cos_of_float(x[i]);
    }
}

void pll_model_sincos_d(const float *x, int n, double *s, double *c)
{
    for (int i = 0; i < n; i++) {
        const Trig t = sincos_f32arg((double)x[i]);
        s[i] = t.sn;
        c[i] = t.cs;
    }
}

}  // extern "C"
