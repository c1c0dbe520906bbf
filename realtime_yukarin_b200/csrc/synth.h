// synth.h -- device-resident state of one WORLD realtime synthesizer (one per audio stream).
#pragma once
#include <vector>

#include "common.cuh"

namespace ryk {

struct Engine;
constexpr int kNoiseTile = 8192;

struct SynthState {            // lives in device memory
  long long cumulative_frame;  // index of the newest frame (-1 before the first AddParameters)
  int handoff;
  double handoff_phase, handoff_f0;
  long long last_location, synthesized_sample;
  long long n_pulses, next_pulse;
  long long rng_generated;     // (unused on device; the host tracks how far the noise ring is filled)
  uint32_t rng_state[2][4];    // xorshift128 state at the fill position, double-buffered across bulk launches
  int plan_blocks, plan_count; long long plan_first;
  int blocks_out;              // result of the last drain
  int carry_sel;
  int last_add_status;         // 1 ok, 0 ring full
};

struct SynthDev {              // passed by value to kernels
  int fs, fft_size, buffer_size, cap_frames, cap_pulses, cap_noise, max_pulses, carry_len, max_samples_per_add;
  double frame_period;         // seconds
  SynthState* state;
  double* f0; float* sp; float* ap;                 // frame ring
  long long* p_index; double* p_time; int* p_vuv;   // pulse ring
  uint32_t* noise;                                  // randn integer sums, by absolute position
  double *if0, *ivuv, *tp;                          // per-add scratch
  double* resp;                                     // [max_pulses][fft]
  double* carry[2];
  double* dc_remover;
};

struct Synth {
  SynthDev dev;
  std::vector<void*> allocs;
  long long host_cum_frames = -1;
  long long host_noise_generated = 0;
  int host_noise_slot = 0;
};

int synth_create(Engine* e, int fs, double frame_period_ms, int fft_size, int buffer_size, int ring_frames, Synth** out);
void synth_destroy(Synth* s);
int synth_host_advance(Engine* e, Synth* s, int n, cudaStream_t st);
int synth_add_kernel(Engine* e, Synth* s, const double* d_f0, int n, const float* d_sp, const float* d_ap, cudaStream_t st);
int synth_add_async(Engine* e, Synth* s, const double* d_f0, int n, const float* d_sp, const float* d_ap, cudaStream_t st);
int synth_drain_async(Engine* e, Synth* s, double* d_out, int max_blocks, cudaStream_t st);

}  // namespace ryk
