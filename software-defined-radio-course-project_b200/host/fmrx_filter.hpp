// fmrx_filter.hpp -- the reference's operator surface, re-declared for the B200
// library.  Same names, argument order, ownership and (absent) error convention as
// the reference's include/filter.h:15-27 and include/iofunc.h:28, so a translation
// unit written against those headers (src/project.cpp) links against
// filter_shim.cpp + iofunc_shim.cpp + libfmrx_b200.so unchanged.
//
// Contract carried over from the reference:
//   * the caller owns every vector and every state scalar; outputs are resized by
//     the callee (filter.cpp:17,42,80-82,108,147-148,178,189-192);
//   * resample() leaves `state` holding the last taps-1 inputs (filter.cpp:95-102);
//   * PLL() overwrites its input vector with the NCO output (filter.cpp:144-148);
//   * all functions are void and never throw; a CUDA failure inside the library
//     (there is no CPU fallback) is reported on stderr and ends the process with
//     status 1 -- the reference's own failure mode (project.cpp:51-54,284-299).
#ifndef FMRX_FILTER_HPP
#define FMRX_FILTER_HPP

#include <vector>

void impulseResponseLPF(std::vector<float> &h, const float Fs, const float Fc, const int num_taps,
                        const int gain);
void impulseResponseBPF(std::vector<float> &h, const float fs, const float fb, const float fe,
                        const int num_taps);
void resample(std::vector<float> &output, std::vector<float> &state, const std::vector<float> &input,
              const std::vector<float> &coeff, const int up_factor, const int down_factor);
void FMDemod(std::vector<float> &fm_demod, float &prev_i, float &prev_q, const std::vector<float> &i_ds,
             const std::vector<float> &q_ds);
void PLL(std::vector<float> &ncoOut, const float freq, const float Fs, const float ncoScale,
         const float phaseAdjust, const float normBandwidth, float &integrator, float &phaseEst,
         float &feedbackI, float &feedbackQ, float &ncoOut_state, float &trigOffset);
void mixer(std::vector<float> &output, const std::vector<float> &arr1, const std::vector<float> &arr2);
void LRExtraction(std::vector<float> &left, std::vector<float> &right, const std::vector<float> &mono_data,
                  const std::vector<float> &stereo_data);
void readStdinBlockData(unsigned int num_samples, unsigned int block_id, std::vector<float> &block_data);

#endif
