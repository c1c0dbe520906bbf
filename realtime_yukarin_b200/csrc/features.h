// features.h -- declarations for features.cu / convert.cu / session.cu
#pragma once
#include "common.cuh"

namespace ryk {

struct Engine;

int stage1_prologue_run(Engine* e, const float* d_mc, const int* d_index, const int* d_count, int C, float* d_x, int Tp_capacity, cudaStream_t st);
int stage1_epilogue_run(Engine* e, const float* d_y, const int* d_index, const uint8_t* d_mask, const int* d_count, int T, int C,
                        const float* d_f0_in, const float* d_ap_in, const uint8_t* d_voiced_in, int nb, float silent_mc0,
                        float* d_mc_out, float* d_f0_out, float* d_ap_out, uint8_t* d_voiced_out, cudaStream_t st);
constexpr int kColminFloats = 64 * 512;      // column-minimum partials of the stage-2 prologue (one scratch per concurrent stream)
int sr_prologue_run(Engine* e, const float* d_sp, int T, int Tp, int nb, float* d_x, cudaStream_t st, float* d_colmin = nullptr);
int sr_epilogue_run(Engine* e, const float* d_y, int T, int nb, float* d_sp_out, cudaStream_t st);

constexpr float kSilentMc0 = -18.420680743952367f;   // ln(1e-8): silent template mel-cepstrum c0 (DECIDE)

// device buffers of one VoiceChanger.convert_from_acoustic_feature evaluation (window of T frames)
struct ConvertBuffers {
  float *d_wave, *d_f0, *d_ap, *d_mc; uint8_t* d_voiced;                 // inputs
  double* d_mse; uint8_t* d_mask; int* d_index; int* d_count;           // gate
  float *d_mc_out, *d_f0_out, *d_ap_out, *d_sp_mid, *d_sp_out; uint8_t* d_voiced_out;
};
int convert_buffers_get(Engine* e, int T, int n_wave, int nb, int C, ConvertBuffers* out);
// stream-ordered except for one 8-byte D2H of the effective-frame count (picks the stage-1 plan)
int convert_window_device(Engine* e, const ConvertBuffers& cb, int T, int n_wave, int frame_length, int hop, double threshold_db,
                          int order, int fftlen, cudaStream_t st);

void session_destroy_all(Engine* e);
int session_streams_fork(Engine* e, cudaEvent_t ev);
int session_streams_join(Engine* e);

}  // namespace ryk
