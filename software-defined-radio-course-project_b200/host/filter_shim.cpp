// filter_shim.cpp -- the reference's seven filter.h operators as thin adapters onto
// the C ABI of libfmrx_b200.so (include/fmrx.h).  No DSP happens here: each function
// sizes the caller's vectors the way the reference does and forwards raw pointers.
#include <cstdio>
#include <cstdlib>

#include "fmrx.h"
#include "fmrx_filter.hpp"

namespace {
void must(int status, const char *what)
{
    if (status == FMRX_OK)
        return;
    std::fprintf(stderr, "%s: %s (%s)\n", what, fmrx_strerror(status), fmrx_last_error());
    std::exit(1);
}
}  // namespace

void impulseResponseLPF(std::vector<float> &h, const float Fs, const float Fc, const int num_taps,
                        const int gain)
{
    h.assign(num_taps > 0 ? num_taps : 0, 0.0f);
    must(fmrx_impulse_response_lpf(h.data(), Fs, Fc, num_taps, gain), "impulseResponseLPF");
}

void impulseResponseBPF(std::vector<float> &h, const float fs, const float fb, const float fe,
                        const int num_taps)
{
    h.assign(num_taps > 0 ? num_taps : 0, 0.0f);
    must(fmrx_impulse_response_bpf(h.data(), fs, fb, fe, num_taps), "impulseResponseBPF");
}

void resample(std::vector<float> &output, std::vector<float> &state, const std::vector<float> &input,
              const std::vector<float> &coeff, const int up_factor, const int down_factor)
{
    const size_t taps = coeff.size();
    const size_t state_in = state.size();
    output.assign(input.size() * static_cast<size_t>(up_factor) / static_cast<size_t>(down_factor), 0.0f);
    if (state.size() < taps - 1)
        state.resize(taps - 1);                 // capacity for the new state
    size_t n_out = 0;
    must(fmrx_resample(output.data(), &n_out, state.data(), state_in, input.data(), input.size(),
                       coeff.data(), static_cast<int>(taps), up_factor, down_factor),
         "resample");
    output.resize(n_out);
    state.resize(taps - 1);
}

void FMDemod(std::vector<float> &fm_demod, float &prev_i, float &prev_q, const std::vector<float> &i_ds,
             const std::vector<float> &q_ds)
{
    fm_demod.assign(i_ds.size(), 0.0f);
    must(fmrx_fmdemod(fm_demod.data(), &prev_i, &prev_q, i_ds.data(), q_ds.data(), i_ds.size()), "FMDemod");
}

void PLL(std::vector<float> &ncoOut, const float freq, const float Fs, const float ncoScale,
         const float phaseAdjust, const float normBandwidth, float &integrator, float &phaseEst,
         float &feedbackI, float &feedbackQ, float &ncoOut_state, float &trigOffset)
{
    float st[6] = { integrator, phaseEst, feedbackI, feedbackQ, ncoOut_state, trigOffset };
    must(fmrx_pll(ncoOut.data(), ncoOut.size(), freq, Fs, ncoScale, phaseAdjust, normBandwidth, st), "PLL");
    integrator = st[0];
    phaseEst = st[1];
    feedbackI = st[2];
    feedbackQ = st[3];
    ncoOut_state = st[4];
    trigOffset = st[5];
}

void mixer(std::vector<float> &output, const std::vector<float> &arr1, const std::vector<float> &arr2)
{
    output.assign(arr1.size(), 0.0f);
    must(fmrx_mixer(output.data(), arr1.data(), arr2.data(), arr1.size()), "mixer");
}

void LRExtraction(std::vector<float> &left, std::vector<float> &right, const std::vector<float> &mono_data,
                  const std::vector<float> &stereo_data)
{
    left.assign(mono_data.size(), 0.0f);
    right.assign(mono_data.size(), 0.0f);
    must(fmrx_lr_extract(left.data(), right.data(), mono_data.data(), stereo_data.data(), mono_data.size()),
         "LRExtraction");
}
