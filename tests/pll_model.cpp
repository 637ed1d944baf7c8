// pll_model.cpp -- HOST build of the device PLL step (csrc/fmrx_pll_core.h) for the
// CPU tests: lets the low-latency sincos / atan2 formulation be compared with the
// oracle bit for bit on long signals without a GPU.  Test infrastructure only; the
// product library never contains a host PLL.
// Build: g++ -O2 -ffp-contract=off [-mfma] -shared -fPIC (tests/conftest.py).
#include <cstddef>
#include <cstring>
#include <cstdio>
#include <cstdlib>
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

// ---- the candidate-table path of k_pll, sequentially ----------------------------------------
//
// Same arithmetic as the kernel's three kinds of warp (csrc/fmrx_kernels.cu: prepare(), the
// predictor, the candidate tables, pll_table_group, pll_block_exact), without the concurrency:
// groups of 1024 steps, the predictor restarted from the exact state at every group, pi per
// block of 16 from the predictor's phaseEst of the step before it, three hypotheses per step
// selected by comparing t = fma(phaseEst, 1/ulp, -pi) with the table's thresholds, the same
// guards, a block with a failed guard stepped again the exact way.  Lets a regime that the
// tables get wrong be found on the CPU.  stats: [0] table blocks, [1] exact blocks, [2] groups
// without tables.
// pll_model_stale_head = 1 reproduces a hazard k_pll had: the predictor and the candidate warps do the
// first 48 steps of the NEXT group at the end of a group; if that next group then does not continue its
// predecessor (it restarts from the exact state because the predecessor needed exact blocks), those
// tables -- centred and, worse, given their block pi by the OLD predictor run -- still carried valid
// stamps and were used with the pi warp 0 takes from the exact phaseEst.
extern "C" {
int pll_model_stale_head = 0;
}

extern "C" int pll_model_tables(const float *pilot, int n, float freq, float Fs, float bw, float *state5,
                                float *trig_out, long long *stats)
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
    const int GROUP = 1024;
    struct In { float x; double xd, inv_x, turn, v; int vi; float vr, cpred; };
    static In in[GROUP + 1];
    static float php[GROUP];            // the predictor's phaseEst after each step
    static float php_stale[48];         // ... of the first 48 steps, by the run that continued from the group before
    float stale_integ = 0.0f, stale_ph = 0.0f, stale_prev = 0.0f;
    bool have_stale = false;
    auto bits = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    auto fbits = [](int i) { float f; memcpy(&f, &i, 4); return f; };
    for (int base = 0; base < n; base += GROUP) {
        const int cnt = n - base < GROUP ? n - base : GROUP;
        bool spec = toff_is_regular(c.toff) && c.binade != FMRX_DISARMED && cnt == GROUP && base + GROUP < n;
        const double ulp = c.ulp, inv_ulp = c.inv_ulp;
        const float inv_ulp_f = (float)inv_ulp;
        const Chain ck = c;
        bool good = spec;
        if (spec) {
            // prepare(): per-sample inputs of this group (and the first of the next: the last table needs it)
            for (int j = 0; j <= GROUP; j++) {
                const int u = base + j;
                In &i = in[j];
                i.x = pilot[u];
                i.xd = (double)i.x;
                i.inv_x = 1.0 / i.xd;
                i.turn = i.x < 0.0f ? 2.0 : 0.0;
                const float toff = toff_after(ck.toff, j + 1);
                i.v = p_mul(k.w, (double)toff);
                const double qv = grid_round(i.v, inv_ulp);
                i.vi = grid_index(qv);
                i.vr = (float)p_fma(i.v, inv_ulp, -p_add(qv, -FMRX_RINT_MAGIC));
                i.cpred = predictor_c(k, i.x, toff_after(ck.toff, j));
            }
            // the predictor, from the exact state
            {
                if (have_stale) {
                    float pi = stale_integ, pp = stale_ph;
                    for (int j = 0; j < 48; j++) {
                        predictor_step(k, in[j].cpred, pi, pp);
                        php_stale[j] = pp;
                    }
                }
                float pi = ck.integ, pp = ck.ph;
                for (int j = 0; j < GROUP; j++) {
                    predictor_step(k, in[j].cpred, pi, pp);
                    php[j] = pp;
                }
                stale_integ = pi;
                stale_ph = pp;
            }
            const bool use_stale = pll_model_stale_head && have_stale;
            const float stale_prev_now = stale_prev;
            long long exact_before = stats[1];
            float integ = c.integ, ph = c.ph;
            // Kp*errorD, Ki*errorD of the first sample, from the known trigArg
            float kpe, kie;
            {
                const Feedback f0 = make_feedback(K, c.tad, in[0].turn, in[0].inv_x, nullptr, nullptr);
                const float ed = error_from_feedback(f0, in[0].x, in[0].xd, good);
                kpe = k.kp * ed;
                kie = k.ki * ed;
            }
            int gi = grid_index(grid_round(c.tad, inv_ulp));
            for (int t = 0; good && t < GROUP; t += 16) {
                // pi of the block
                const float ph_for_pi = t == 0 ? ph : php[t - 1];
                const float pm = ph_for_pi * inv_ulp_f + 12582912.0f;
                const float pi_f = pm - 12582912.0f;
                const int pi_i = bits(pm) - 0x4B400000;
                if (!(fabsf(pi_f) < 2097152.0f)) { good = false; break; }
                const float integ0 = integ, ph0 = ph;
                const int gi0 = gi;
                integ = integ + kie;
                ph = ph + (kpe + integ);
                bool bad = false;
                float cmax = 0.0f;
                int gidx[16];
                for (int j = 0; j < 16; j++) {
                    const int u = t + j;
                    // the candidate table of step u: hypotheses G_c - 1, G_c, G_c + 1 of trigArg(u) -> errorD of sample u + 1
                    const bool st = use_stale && u < 48;
                    const int gc = grid_index(grid_round(p_add(in[u].v, (double)(st ? php_stale[u] : php[u])), inv_ulp));
                    int pi_tab = pi_i;          // the pi the candidate warps used for this block
                    if (st) {
                        const float phs = t == 0 ? stale_prev_now : php_stale[t - 1];
                        pi_tab = bits(phs * inv_ulp_f + 12582912.0f) - 0x4B400000;
                    }
                    float kpe_h[3], kie_h[3];
                    bool ok = true;
                    for (int h = 0; h < 3; h++) {
                        const int gl = gc - 1 + h;
                        const double tad = p_mul((double)gl, ulp);
                        const Feedback f = make_feedback(K, tad, in[u + 1].turn, in[u + 1].inv_x, nullptr, nullptr);
                        const int ag = gl < 0 ? -gl : gl;
                        bool okh = ag > (1 << 23) && ag < (1 << 24);
                        const float ed = error_from_feedback(f, in[u + 1].x, in[u + 1].xd, okh);
                        kpe_h[h] = k.kp * ed;
                        kie_h[h] = k.ki * ed;
                        ok = ok && okh;
                    }
                    const int n1 = gc - (in[u].vi + pi_tab);
                    ok = ok && n1 >= -60 && n1 <= 60;
                    const float lp = ok ? ((float)n1 + -0.5f) + -in[u].vr : 0x1p100f;
                    // the chain step
                    const float hp = lp + 1.0f, lc = lp + 0.5f;
                    const float tt = fmaf(ph, inv_ulp_f, -pi_f);
                    const int sel = tt < lp ? 0 : tt > hp ? 2 : 1;
                    if (j < 15) {
                        const float i_s = integ + kie_h[sel];
                        const float p_s = ph + (kpe_h[sel] + i_s);
                        integ = i_s;
                        ph = p_s;
                    } else {
                        kpe = kpe_h[sel];
                        kie = kie_h[sel];
                    }
                    const float q = tt + -lc;
                    cmax = fmaxf(cmax, fabsf(fabsf(fabsf(q) + -0.5f) + -0.5f));
                    gidx[j] = in[u].vi + pi_i + (bits((tt + in[u].vr) + 12582912.0f) - 0x4B400000);
                }
                if (bad || !(cmax < 0.5f - 0x1p-15f)) {
                    // the block again, the exact way (pll_block_exact)
                    stats[1]++;
                    Chain e;
                    memset(&e, 0, sizeof(e));
                    e.integ = integ0;
                    e.ph = ph0;
                    e.toff = toff_after(ck.toff, t);
                    e.tad = p_mul((double)gi0, ulp);
                    chain_refresh(e);
                    for (int j = 0; j < 16; j++) {
                        if (e.binade == FMRX_DISARMED || e.ulp != ulp) { good = false; break; }
                        const In &i = in[t + j];
                        StepIn si;
                        si.x = i.x; si.xd = i.xd; si.inv_x = i.inv_x; si.turn = i.turn; si.v = i.v;
                        if (!chain_step_fast(e, k, K, si))
                            chain_step_generic(e, k, i.x);
                        gidx[j] = grid_index(grid_round(e.tad, e.inv_ulp));
                    }
                    if (!good || e.binade == FMRX_DISARMED || e.ulp != ulp) { good = false; break; }
                    const In &nx = in[t + 16];
                    const Feedback f = make_feedback(K, e.tad, nx.turn, nx.inv_x, nullptr, nullptr);
                    bool ok = true;
                    float ed = error_from_feedback(f, nx.x, nx.xd, ok);
                    if (!ok) {
                        float fi, fq;
                        chain_feedback(e, fi, fq);
                        ed = (float)atan2((double)(nx.x * -fq), (double)(nx.x * fi));
                    }
                    kpe = k.kp * ed;
                    kie = k.ki * ed;
                    integ = e.integ;
                    ph = e.ph;
                } else {
                    stats[0]++;
                }
                gi = gidx[15];
                for (int j = 0; j < 16; j++)
                    trig_out[base + t + j] = (float)p_mul((double)gidx[j], ulp);
            }
            have_stale = good && stats[1] - exact_before > 2;      // the next group will not "continue" this one
            stale_prev = php[GROUP - 1];
            if (good) {
                // (integ, ph) are the state after sample base + GROUP - 1 with the next sample's errorD pending:
                // rebuild the chain from the last trigArg
                c.integ = integ;
                c.ph = ph;
                c.toff = toff_after(ck.toff, GROUP);
                c.tad = p_mul((double)gi, ulp);
                chain_refresh(c);
                if (c.ulp != ulp)
                    ;                     // binade change: the next group starts on the new grid
            }
        }
        if (!good) {
            have_stale = false;
            stats[2]++;
            c = ck;
            for (int t = 0; t < cnt; t++)
                trig_out[base + t] = chain_step(c, k, K, pilot[base + t], nullptr);
        }
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    (void)fbits;
    return 0;
}

// ---- a predictor in the reference's own operation order (design study for modes 2/3) ---------
//
// Where phaseEst's float spacing exceeds trigArg's (the reference hands its PLL if_fs*interp in
// modes 2/3, phaseEst cancels w*trigOffset) the hypotheses of a step have to be neighbouring floats
// of PHASEEST.  That needs a predictor that follows phaseEst to the bit most of the time: the loop
// filter exactly as the reference orders it (:163-164), errorD from the ROUNDED trigArg by the
// identity atan2(x*(-sin t), x*cos t) = wrap(pi*(x < 0) - t) evaluated in double and rounded to
// float once (right to an ulp or so of the reference's float errorD).  Restarted from the exact state
// every `group` steps; hist[d + 8] counts steps whose predicted phaseEst is d float spacings from
// the exact one (clamped to +-8).
extern "C" int pll_model_predict_exact_order(const float *pilot, int n, float freq, float Fs, float bw, float *state5,
                                             int group, long long *hist)
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
    const double two_pi = 6.283185307179586476925287, pi = 3.14159265358979323846;
    float pi_ = 0, pp = 0;
    double pt = 0;                      // the predictor's (rounded) trigArg
    for (int u = 0; u < n; u++) {
        if (u % group == 0) {
            pi_ = c.integ;
            pp = c.ph;
            pt = c.tad;
        }
        const float x = pilot[u];
        // errorD of this sample from the predictor's previous trigArg
        double a = (x < 0.0f ? pi : 0.0) - pt;
        a -= two_pi * nearbyint(a / two_pi);
        const float e = (float)a;
        pi_ = pi_ + k.ki * e;                                  // :163
        pp = pp + (k.kp * e + pi_);                            // :164
        const float toff = toff_after(c.toff, 1);
        pt = (double)(float)(k.w * (double)toff + (double)pp); // :166-167
        chain_step(c, k, K, x, nullptr);
        const float sp = fabsf(c.ph) > 0 ? nextafterf(fabsf(c.ph), INFINITY) - fabsf(c.ph) : 1e-45f;
        double d = ((double)pp - (double)c.ph) / (double)sp;
        int b = (int)nearbyint(d);
        b = b < -8 ? -8 : b > 8 ? 8 : b;
        hist[b + 8]++;
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    return 0;
}

// The one-hypothesis scheme of DESIGN.md section 9 (modes 2/3), sequentially: the exact-order predictor
// runs a group ahead from the exact state; the exact phase detector is evaluated for the trigArg that
// follows from the PREDICTED phaseEst of the step before; the exact loop filter runs on those errorDs and a
// block of 16 is accepted iff its phaseEst matched the predictor's bit for bit at every step (otherwise it
// is stepped again the exact way).  stats: [0] accepted blocks, [1] exact blocks.
extern "C" int pll_model_one_hypothesis(const float *pilot, int n, float freq, float Fs, float bw, float *state5,
                                        float *trig_out, long long *stats)
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
    const double two_pi = 6.283185307179586476925287, pi = 3.14159265358979323846;
    const int GROUP = 1024;
    static float pph[GROUP];
    static double pt[GROUP];
    auto bits = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    for (int base = 0; base < n; base += GROUP) {
        const int cnt = n - base < GROUP ? n - base : GROUP;
        // the predictor, from the exact state
        {
            float pi_ = c.integ, pp = c.ph;
            double t = c.tad;
            for (int j = 0; j < cnt; j++) {
                const float x = pilot[base + j];
                double a = (x < 0.0f ? pi : 0.0) - t;
                a -= two_pi * nearbyint(a / two_pi);
                const float e = (float)a;
                pi_ = pi_ + k.ki * e;
                pp = pp + (k.kp * e + pi_);
                t = (double)(float)(k.w * (double)toff_after(c.toff, j + 1) + (double)pp);
                pph[j] = pp;
                pt[j] = t;
            }
        }
        for (int tb = 0; tb < cnt; tb += 16) {
            const int nb = cnt - tb < 16 ? cnt - tb : 16;
            const Chain ck = c;
            float integ = c.integ, ph = c.ph;
            int bad = toff_is_regular(c.toff) ? 0 : 1;
            for (int j = 0; j < nb && !bad; j++) {
                const int u = tb + j;
                const float x = pilot[base + u];
                const double tprev = u == 0 ? ck.tad : (j == 0 ? c.tad : pt[u - 1]);
                bool ok = fabs(tprev) <= FMRX_FAST_TRIG_MAX;
                const Feedback f = make_feedback(K, tprev, x < 0.0f ? 2.0 : 0.0, 1.0 / (double)x, nullptr, nullptr);
                const float ed = error_from_feedback(f, x, (double)x, ok);
                integ = integ + k.ki * ed;
                ph = ph + (k.kp * ed + integ);
                bad |= !ok;
                bad |= bits(ph) ^ bits(pph[u]);
            }
            if (!bad) {
                stats[0]++;
                c.integ = integ;
                c.ph = ph;
                c.toff = toff_after(ck.toff, nb);
                c.tad = pt[tb + nb - 1];
                chain_refresh(c);
                for (int j = 0; j < nb; j++)
                    trig_out[base + tb + j] = (float)pt[tb + j];
            } else {
                stats[1]++;
                c = ck;
                for (int j = 0; j < nb; j++)
                    trig_out[base + tb + j] = chain_step(c, k, K, pilot[base + tb + j], nullptr);
            }
        }
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    return 0;
}

// ---- the one-hypothesis scheme AS k_pll RUNS IT (float predictor), sequentially ------------------
//
// Same roles as the kernel (csrc/fmrx_kernels.cu, the "1H" groups), per group of 1024 steps:
//   prepare  : per sample, from (integrator, phaseEst) at the start of the group TWO BEFORE (the I/O warps
//              work two groups ahead) a float P near the phaseEst to come (linear extrapolation: the mean
//              slope of phaseEst is the integrator), then with S = w*trigOffset + P in double:
//              B = fl32(S), r = fl32(S - B), c' = fl32(wrap(pi*(x_next < 0) - B)).
//              trigArg = fl32(w*trigOffset + phaseEst) = fl32(S + d), d = phaseEst (-) P  (exact: Sterbenz),
//              is then B (+) (r (+) d), its distance from B is z = trigArg (-) B (small), and the phase
//              detector's angle for the next sample wrap(c' - z): no large-argument reduction anywhere.
//   predictor: FMUL + 8 dependent FADD per step in the reference's operation order (:163-164),
//              restarted from the exact state at every group.
//   candidates: the exact phase detector (make_feedback / error_from_feedback) for the trigArg the
//              reference forms from the PREDICTED phaseEst of the step before (:167, in double).
//   chain    : the exact loop filter on those errorDs; a block of 16 is accepted iff its phaseEst
//              matched the predictor's bit for bit at every step (and every candidate guard held).
// policy 0: a failed block is stepped the exact way and the comparison goes on (the predictor is not
// restarted inside a group); policy 1: after a failed block the rest of the group is stepped exactly.
// stats: [0] accepted blocks, [1] exact blocks, [2] groups with at least one failure.
// per group of the last pll_model_onehyp_float run: |phaseEst| at its start and whether it had a failed block
extern "C" {
float pll_model_grp_ph[1 << 17];
unsigned char pll_model_grp_fail[1 << 17];
}

extern "C" int pll_model_onehyp_float(const float *pilot, int n, float freq, float Fs, float bw, float *state5,
                                      float *trig_out, long long *stats, int policy)
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
    const int GROUP = 2048, PBLK = 16;       // PLL_GROUP, PLL_PBLK
    static float pph[GROUP];
    static OneHypIn in1[GROUP];
    static double v[GROUP];
    auto bits = [](float f) { int i; memcpy(&i, &f, 4); return i; };
    // (integ, ph) at the start of every group: the inputs of group g are prepared while group g - 1 runs, and
    // extrapolate from its start (the launch's initial state for the first two groups)
    static float h_slope[1 << 17], h_ph[1 << 17];
    for (int base = 0, g = 0; base < n; base += GROUP, g++) {
        const int cnt = n - base < GROUP ? n - base : GROUP;
        const bool regular = toff_is_regular(c.toff) && fabs(c.tad) <= FMRX_FAST_TRIG_MAX;
        bool group_failed = false;
        if (g < (1 << 17)) {
            pll_model_grp_ph[g] = fabsf(c.ph);
            pll_model_grp_fail[g] = 0;
            h_slope[g] = g > 0 ? (c.ph + -h_ph[g - 1]) * (1.0f / GROUP) : c.integ;      // as k_pll's header: the slope over the group before
            h_ph[g] = c.ph;
        }
        const int og = g >= 1 ? g - 1 : 0;
        const float e_integ = h_slope[og], e_ph = h_ph[og];
        const int e_age = (g - og) * GROUP;
        if (regular) {
            for (int j = 0; j < cnt; j++) {
                v[j] = p_mul(k.w, (double)toff_after(c.toff, j + 1));
                const float x_next = base + j + 1 < n ? pilot[base + j + 1] : 1.0f;
                in1[j] = onehyp_inputs(v[j], e_ph, e_integ, e_age + j + 1, x_next);
            }
            // predictor
            float pi_ = c.integ, pp = c.ph;
            const bool short_form = onehyp_short_ok(c.tad, c.ph);  // (decided per group, in the header)
            stats[3 + 0] += 0;
            if (short_form && g < (1 << 17))
                pll_model_grp_fail[g] |= 2;                        // (bit 1: the group ran the short predictor step)
            float a = onehyp_first_angle(pilot[base], c.tad);      // the angle from the exact trigArg before the group
            for (int j0 = 0; j0 < cnt; j0 += PBLK) {       // as the predictor warp: blocks of PBLK = 16 (PLL_PBLK), again with the reduction if an angle left [-pi, pi]
                const float a0 = a, i0 = pi_, p0 = pp;
                float amax = fabsf(a);
                for (int j = j0; j < j0 + PBLK && j < cnt; j++) {
                    a = short_form ? onehyp_predictor_step_short(k, in1[j].P, in1[j].c + -in1[j].r, a, pi_, pp)
                                   : onehyp_predictor_step(k, in1[j], a, pi_, pp);
                    pph[j] = pp;
                    if (j + 1 < j0 + PBLK)
                        amax = fmaxf(amax, fabsf(a));
                }
                if (!(amax <= FMRX_ONEHYP_PI)) {
                    stats[3]++;                  // blocks of 8 predictor steps that needed the angle reduction
                    a = a0; pi_ = i0; pp = p0;
                    if (!(fabsf(a) <= FMRX_ONEHYP_PI))
                        a = a - 6.2831855f * ((a * 0.15915494f + 12582912.0f) - 12582912.0f);
                    for (int j = j0; j < j0 + PBLK && j < cnt; j++) {
                        a = onehyp_predictor_step_reduced(k, in1[j], a, pi_, pp);
                        pph[j] = pp;
                    }
                }
            }
        }
        for (int tb = 0; tb < cnt; tb += 16) {
            const int nb = cnt - tb < 16 ? cnt - tb : 16;
            const Chain ck = c;
            float integ = c.integ, ph = c.ph;
            int bad = (regular && !(policy == 1 && group_failed)) ? 0 : 1;
            double tprev = c.tad;
            for (int j = 0; j < nb && !bad; j++) {
                const int u = tb + j;
                const float x = pilot[base + u];
                if (j > 0)
                    tprev = onehyp_trigarg(v[u - 1], pph[u - 1]);       // :167 from the predicted phaseEst
                bool ok = fabs(tprev) <= FMRX_FAST_TRIG_MAX;
                const Feedback f = make_feedback(K, tprev, x < 0.0f ? 2.0 : 0.0, 1.0 / (double)x, nullptr, nullptr);
                const float ed = error_from_feedback(f, x, (double)x, ok);
                integ = integ + k.ki * ed;
                ph = ph + (k.kp * ed + integ);
                bad |= !ok;
                bad |= bits(ph) ^ bits(pph[u]);
                if (bad && getenv("PLL_MODEL_DEBUG") && fabsf(c.ph) > 30000.0f)
                    fprintf(stderr, "fail g %d step %d ok %d ph %.9g pred %.9g (%d ulps) ed %.9g x %.4g tprev %.9g | in: P %.9g r %.9g B %.9g c %.9g\n", g, u, (int)ok, ph,
                            pph[u], bits(ph) - bits(pph[u]), ed, x, tprev, in1[u].P, in1[u].r, in1[u].B, in1[u].c);
            }
            if (!bad) {
                stats[0]++;
                c.integ = integ;
                c.ph = ph;
                c.toff = toff_after(ck.toff, nb);
                c.tad = onehyp_trigarg(v[tb + nb - 1], ph);
                chain_refresh(c);
                for (int j = 0; j < nb; j++)
                    trig_out[base + tb + j] = (float)onehyp_trigarg(v[tb + j], pph[tb + j]);
            } else {
                stats[1]++;
                if (!group_failed)
                    stats[2]++;
                group_failed = true;
                if (g < (1 << 17))
                    pll_model_grp_fail[g] = 1;
                c = ck;
                for (int j = 0; j < nb; j++)
                    trig_out[base + tb + j] = chain_step(c, k, K, pilot[base + tb + j], nullptr);
            }
        }
    }
    state5[0] = c.integ; state5[1] = c.ph; state5[4] = c.toff;
    chain_feedback(c, state5[2], state5[3]);
    return 0;
}
