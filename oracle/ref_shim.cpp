/*
 * ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference operators.
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles the reference's own
 * src/filter.cpp and src/iofunc.cpp from /root/reference (where they lie; no
 * copy is made) with the reference's flags (src/Makefile:3-8: g++ -O3, no
 * -march, no fast-math) and links them with this file into
 * oracle/_ref/libref_fm.so.  Nothing here re-implements DSP: each wrapper only
 * moves data between plain arrays and the std::vector signatures declared in
 * the reference's include/filter.h:15-27 and include/iofunc.h:28.
 *
 * ref_chain_* replays the reference's block loop single-threaded, calling the
 * reference functions in the order src/project.cpp:48-84 and :132-196 call
 * them, so that per-stage dumps can be compared.  (The reference binary itself
 * is also built, oracle/_ref/project, for end-to-end PCM-prefix checks.)
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "filter.h"
#include "fourier.h"   /* reference include/filter.h via -I */
#include "iofunc.h"   /* reference include/iofunc.h via -I */

typedef std::vector<float> vf;

static void out_copy(float *dst, const vf &v)
{
    if (dst && !v.empty())
        std::memcpy(dst, v.data(), v.size() * sizeof(float));
}

extern "C" {

void ref_lpf_taps(float *h, float Fs, float Fc, int taps, int gain)
{
    vf v;
    impulseResponseLPF(v, Fs, Fc, taps, gain);
    out_copy(h, v);
}

void ref_bpf_taps(float *h, float fs, float fb, float fe, int taps)
{
    vf v;
    impulseResponseBPF(v, fs, fb, fe, taps);
    out_copy(h, v);
}

int ref_resample(float *out, float *state, int state_len, const float *in,
                 int n_in, const float *coeff, int taps, int up, int down)
{
    vf o, s(state, state + state_len), i(in, in + n_in), c(coeff, coeff + taps);
    resample(o, s, i, c, up, down);
    out_copy(out, o);
    out_copy(state, s);
    return (int)o.size();
}

void ref_fmdemod(float *out, float *prev_i, float *prev_q, const float *i_ds,
                 const float *q_ds, int n)
{
    vf o, i(i_ds, i_ds + n), q(q_ds, q_ds + n);
    FMDemod(o, *prev_i, *prev_q, i, q);
    out_copy(out, o);
}

/* st = {integrator, phaseEst, feedbackI, feedbackQ, ncoOut_state, trigOffset} */
void ref_pll(float *inout, int n, float freq, float Fs, float scale,
             float phase_adjust, float norm_bw, float *st)
{
    vf v(inout, inout + n);
    PLL(v, freq, Fs, scale, phase_adjust, norm_bw, st[0], st[1], st[2], st[3],
        st[4], st[5]);
    out_copy(inout, v);
}

void ref_mixer(float *out, const float *a, const float *b, int n)
{
    vf o, x(a, a + n), y(b, b + n);
    mixer(o, x, y);
    out_copy(out, o);
}

void ref_lr_extract(float *left, float *right, const float *mono,
                    const float *stereo, int n)
{
    vf l, r, m(mono, mono + n), s(stereo, stereo + n);
    LRExtraction(l, r, m, s);
    out_copy(left, l);
    out_copy(right, r);
}

/* ---- block-loop replay over the reference operators -------------------- */

struct ref_dump {
    float *i_ds, *q_ds, *demod, *chan, *pilot, *trig, *nco, *mixer;
    float *mono, *mono_shift, *stereo, *left, *right;
};

struct ref_chain {
    int rf_decim, audio_decim, audio_interp, block_size, if_fs, bp_fs;
    vf rf_coeff, chan_coeff, pilot_coeff, audio_coeff;
    vf st_i, st_q, st_chan, st_pilot, st_audio, mono_state;
    float prev_i, prev_q;
    float integ, phase, fb_i, fb_q, nco_state, trig_off;
};

ref_chain *ref_chain_create(int mode, int taps)
{
    static const int k_rf_fs[4] = { 2400000, 1152000, 2400000, 2304000 };
    static const int k_rf_dec[4] = { 10, 4, 10, 9 };
    static const int k_if[4] = { 240000, 288000, 240000, 256000 };
    static const int k_ad[4] = { 5, 6, 800, 2560 };
    static const int k_ai[4] = { 1, 1, 147, 441 };
    if (mode < 0 || mode > 3)
        return nullptr;
    ref_chain *c = new ref_chain();
    c->rf_decim = k_rf_dec[mode];
    c->audio_decim = k_ad[mode];
    c->audio_interp = k_ai[mode];
    c->bp_fs = k_if[mode];
    c->if_fs = k_if[mode] * k_ai[mode];
    c->block_size = 256 * c->rf_decim * c->audio_decim;
    const int audio_taps = taps * c->audio_interp;
    impulseResponseLPF(c->rf_coeff, k_rf_fs[mode], 100000, taps, 1);
    impulseResponseBPF(c->chan_coeff, c->bp_fs, 22000.0, 54000.0, taps);
    impulseResponseBPF(c->pilot_coeff, c->bp_fs, 18500, 19500, taps);
    impulseResponseLPF(c->audio_coeff, c->if_fs, 16000, audio_taps, c->audio_interp);
    c->st_i.assign(taps - 1, 0.0f);
    c->st_q.assign(taps - 1, 0.0f);
    c->st_chan.assign(taps - 1, 0.0f);
    c->st_pilot.assign(taps - 1, 0.0f);
    c->st_audio.assign(audio_taps - 1, 0.0f);
    c->mono_state.assign(5, 0.0f);
    c->prev_i = c->prev_q = 0.0f;
    c->integ = 0.0f; c->phase = 0.0f; c->fb_i = 1.0f; c->fb_q = 0.0f;
    c->nco_state = 1.0f; c->trig_off = 0.0f;
    return c;
}

void ref_chain_destroy(ref_chain *c) { delete c; }

int ref_chain_block_size(const ref_chain *c) { return c->block_size; }

void ref_chain_block(ref_chain *c, const uint8_t *iq, int16_t *pcm, const ref_dump *d)
{
    const int n = c->block_size;
    /* same conversion expression as the reference's readStdinBlockData
     * (src/iofunc.cpp:67), applied to a memory buffer instead of std::cin */
    vf iqf(n);
    for (int k = 0; k < n; k++)
        iqf[k] = (((float)(unsigned char)iq[k]) - 128.0) / 128.0;
    vf ib(n / 2), qb(n / 2);
    for (int k = 0, j = 0; k < n; k += 2, j++) {
        ib[j] = iqf[k];
        qb[j] = iqf[k + 1];
    }
    vf i_ds, q_ds, demod, mono, mono_shift, chan, carrier, pilot, mix, stereo, left, right;
    resample(i_ds, c->st_i, ib, c->rf_coeff, 1, c->rf_decim);
    resample(q_ds, c->st_q, qb, c->rf_coeff, 1, c->rf_decim);
    FMDemod(demod, c->prev_i, c->prev_q, i_ds, q_ds);

    resample(mono, c->st_audio, demod, c->audio_coeff, c->audio_interp, c->audio_decim);
    mono_shift.insert(mono_shift.end(), c->mono_state.begin(), c->mono_state.end());
    mono_shift.insert(mono_shift.end(), mono.begin(), mono.end() - 5);
    c->mono_state.assign(mono.end() - 5, mono.end());

    resample(chan, c->st_chan, demod, c->chan_coeff, 1, 1);
    resample(carrier, c->st_pilot, demod, c->pilot_coeff, 1, 1);
    pilot = carrier;
    PLL(carrier, 19000, c->if_fs, 2, 0, 0.01, c->integ, c->phase, c->fb_i, c->fb_q,
        c->nco_state, c->trig_off);
    mixer(mix, chan, carrier);
    resample(stereo, c->st_audio, mix, c->audio_coeff, c->audio_interp, c->audio_decim);
    LRExtraction(left, right, mono_shift, stereo);

    for (size_t k = 0; k < left.size(); k++) {
        pcm[2 * k] = std::isnan(right[k]) ? 0 : static_cast<short int>(right[k] * 16384);
        pcm[2 * k + 1] = std::isnan(left[k]) ? 0 : static_cast<short int>(left[k] * 16384);
    }
    if (d) {
        out_copy(d->i_ds, i_ds); out_copy(d->q_ds, q_ds); out_copy(d->demod, demod);
        out_copy(d->chan, chan); out_copy(d->pilot, pilot); out_copy(d->nco, carrier);
        out_copy(d->mixer, mix); out_copy(d->mono, mono); out_copy(d->mono_shift, mono_shift);
        out_copy(d->stereo, stereo); out_copy(d->left, left); out_copy(d->right, right);
        /* the reference does not expose trigArg; d->trig is left untouched */
    }
}

void ref_chain_get_pll(const ref_chain *c, float *st)
{
    st[0] = c->integ; st[1] = c->phase; st[2] = c->fb_i; st[3] = c->fb_q;
    st[4] = c->nco_state; st[5] = c->trig_off;
}


/* ---- the reference's RDS sketch (src/project.cpp:200-271), replayed with the reference's OWN
 * operators in the order rds_thread calls them (the thread itself is never started and sits
 * in the same translation unit as main, so it cannot be linked from here) ---- */
struct ref_rds {
    float fs;
    int taps, delay;
    vf extract_coeff, carrier_coeff, channel_state, carrier_state, shift_state;
    float integ, phase, fb_i, fb_q, nco_state, trig_off;
};

ref_rds *ref_rds_create(float bp_fs, int taps, int channel_delay)
{
    ref_rds *r = new ref_rds();
    r->fs = bp_fs;
    r->taps = taps;
    r->delay = channel_delay;
    r->channel_state.assign(taps - 1, 0.0);                                      /* :209 */
    r->carrier_state.assign(taps - 1, 0.0);                                      /* :216 */
    r->shift_state.assign(channel_delay, 0.0);                                   /* :207 */
    impulseResponseBPF(r->extract_coeff, bp_fs, 54000, 60000, taps);             /* :210 */
    impulseResponseBPF(r->carrier_coeff, bp_fs, 113500, 114500, taps);           /* :217 */
    r->integ = 0.0; r->phase = 0.0; r->fb_i = 1.0; r->fb_q = 0.0; r->trig_off = 0.0; r->nco_state = 1.0;   /* :219-224 */
    return r;
}

void ref_rds_destroy(ref_rds *r) { delete r; }

void ref_rds_block(ref_rds *r, const float *demod, int n, float *mixer_out, float *channel, float *carrier_nco)
{
    vf demod_data(demod, demod + n), channel_data, channel_squared, carrier_data, channel_shift, mixer_data;
    resample(channel_data, r->channel_state, demod_data, r->extract_coeff, 1, 1);                  /* :244 */
    channel_squared.resize(channel_data.size(), 0.0);
    for (size_t i = 0; i < channel_data.size(); i++)
        channel_squared[i] = channel_data[i] * channel_data[i];                                    /* :249-251 */
    resample(carrier_data, r->carrier_state, channel_squared, r->carrier_coeff, 1, 1);             /* :254 */
    PLL(carrier_data, 114000, r->fs, 0.5, 0, 0.01, r->integ, r->phase, r->fb_i, r->fb_q, r->nco_state, r->trig_off);   /* :256 */
    channel_shift.insert(channel_shift.end(), r->shift_state.begin(), r->shift_state.end());       /* :259-263 */
    channel_shift.insert(channel_shift.end(), channel_data.begin(), channel_data.end() - r->delay);
    r->shift_state.assign(channel_data.end() - r->delay, channel_data.end());                      /* :265-266 */
    mixer(mixer_data, carrier_data, channel_shift);                                                /* :269 */
    out_copy(mixer_out, mixer_data);
    if (channel) out_copy(channel, channel_data);
    if (carrier_nco) out_copy(carrier_nco, carrier_data);
}


/* ---- the reference's spectrum estimate (src/fourier.cpp:35-117), its own code ---- */
int ref_estimate_psd(float *freq, float *psd, const float *samples, int n, int freq_bins, float Fs)
{
    vf f, p, x(samples, samples + n);
    estimatePSD(f, p, x, freq_bins, Fs);
    out_copy(freq, f);
    out_copy(psd, p);
    return (int)p.size();
}

} /* extern "C" */
