// session.cu -- one audio stream's encode -> convert -> decode chain kept resident in HBM.
//
// Semantics = the reference's three StreamWrapper-driven stages as the workers run them
// (realtime_voice_conversion/worker/{encode,convert,decode}_worker.py with stream/*.py):
// stage input chunk j is added at start_time = extra + j*T and step k processes [k*T - extra, k*T + T + extra),
// i.e. window item i of step k is item k*n - 2e + i of the stage's input sequence, silent where that is
// negative (verified against the reference's BaseStream.fetch).  Instead of Python segment
// lists, each stage keeps its last window on the device and slides it by one chunk per push:
//   wave window   (n_wave + 2 e_wave samples)                      -> WORLD analysis -> trim
//   feature window (n_feat + 2 e_conv frames of f0/ap/mc/voiced + the aligned samples) -> convert -> trim
//   converted window (n_feat + 2 e_dec frames of f0/ap/sp)         -> realtime synthesizer -> NaN scrub
// Only the chunk's samples go in and whole synthesizer blocks come out.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/ryk.h"
#include "engine.h"
#include "features.h"
#include "synth.h"
#include "unet.h"

namespace ryk {

struct Session {
  ryk_session_config cfg;
  int hop, rate, n_wave, n_feat, e_wave, e_enc_frames, e_conv, e_dec;
  int Lw, Tw, Td, nb, C;
  int flip = 0;
  long long step = 0;
  // sliding windows (double-buffered)
  float* wave_win[2];
  float *cw_f0[2], *cw_ap[2], *cw_mc[2], *cw_wave[2]; uint8_t* cw_voiced[2];
  float *dw_f0[2], *dw_ap[2], *dw_sp[2];
  // per-step scratch
  float *enc_f0, *enc_sp, *enc_ap, *enc_mc; uint8_t* enc_voiced;
  double* dec_f0_f64;
  double* out_blocks; int max_blocks;
  int* d_n_out;
  float* d_chunk;
  ConvertBuffers cb;
  double* d_mse; uint8_t* d_mask; int* d_index; int* d_count;
  float *cv_mc_out, *cv_f0_out, *cv_ap_out, *cv_sp_mid, *cv_sp_out; uint8_t* cv_voiced_out;
  Synth* synth = nullptr;
  DioPlan* dio = nullptr;
  std::vector<void*> allocs;
  // pinned staging
  float* h_in = nullptr; double* h_out = nullptr; int* h_n = nullptr;
};

// dst = [old[shift..L), new[0..shift)] row-wise (rows of `row` elements)
template <typename T>
__global__ void k_slide(const T* __restrict__ old_, const T* __restrict__ new_, T* __restrict__ dst, size_t L, size_t shift, size_t row) {
  size_t total = L * row;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / row;
    dst[i] = r + shift < L ? old_[i + shift * row] : new_[i - (L - shift) * row];
  }
}

template <typename T>
__global__ void k_fill_rows(T* __restrict__ dst, size_t rows, size_t row, T first, T rest) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows * row; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (i % row == 0) ? first : rest;
}

__global__ void k_f32_to_f64(const float* __restrict__ a, double* __restrict__ b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = (double)a[i];
}

// NaN -> 0 on the produced samples (decode_stream.py:38) and publish the sample count
__global__ void k_scrub(double* __restrict__ y, const SynthState* __restrict__ st, int block, int max_samples, int* __restrict__ n_out) {
  int n = st->blocks_out * block;
  if (n > max_samples) n = max_samples;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { double v = y[i]; if (v != v) y[i] = 0.0; }
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
}

template <typename T>
static int slide(const T* old_, const T* new_, T* dst, size_t L, size_t shift, size_t row, cudaStream_t st) {
  size_t total = L * row;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256); if (blocks > 2368) blocks = 2368;
  k_slide<T><<<blocks, 256, 0, st>>>(old_, new_, dst, L, shift, row);
  RYK_CUDA(cudaGetLastError());
  return 0;
}

static Session* get_session(Engine* e, int id) { return (id >= 0 && id < (int)e->sessions.size()) ? e->sessions[id] : nullptr; }

static void session_free(Session* s) {
  if (!s) return;
  for (void* p : s->allocs) cudaFree(p);
  if (s->h_in) cudaFreeHost(s->h_in);
  if (s->h_out) cudaFreeHost(s->h_out);
  if (s->h_n) cudaFreeHost(s->h_n);
  synth_destroy(s->synth);
  delete s;
}

void session_destroy_all(Engine* e) {
  for (Session* s : e->sessions) session_free(s);
  e->sessions.clear();
}

static int session_step(Engine* e, Session* s, const float* d_chunk, double* d_out, int out_capacity, int* d_n_out, cudaStream_t st) {
  const int f = s->flip, g = f ^ 1;
  const ryk_session_config& c = s->cfg;
  // ---- encode: slide the wave window, analyse it, keep the central n_feat frames ----
  if (slide<float>(s->wave_win[f], d_chunk, s->wave_win[g], s->Lw, s->n_wave, 1, st)) return -1;
  if (dio_stonemask_run(e, s->dio, s->wave_win[g], st)) return -1;
  const int n_enc = s->Lw / s->hop;
  if (spectral_analysis_run(e, s->wave_win[g], s->Lw, c.fs, c.frame_period_ms, dio_plan_f0(s->dio), n_enc, c.fft_length, c.order,
                            s->enc_sp, s->enc_ap, s->enc_mc, s->enc_f0, s->enc_voiced, st)) return -1;
  e->launches += 13;
  const int pe = s->e_enc_frames;
  // ---- convert window: append the new frames (and their aligned samples) ----
  if (slide<float>(s->cw_f0[f], s->enc_f0 + pe, s->cw_f0[g], s->Tw, s->n_feat, 1, st)) return -1;
  if (slide<float>(s->cw_ap[f], s->enc_ap + (size_t)pe * s->nb, s->cw_ap[g], s->Tw, s->n_feat, s->nb, st)) return -1;
  if (slide<float>(s->cw_mc[f], s->enc_mc + (size_t)pe * s->C, s->cw_mc[g], s->Tw, s->n_feat, s->C, st)) return -1;
  if (slide<uint8_t>(s->cw_voiced[f], s->enc_voiced + pe, s->cw_voiced[g], s->Tw, s->n_feat, 1, st)) return -1;
  if (slide<float>(s->cw_wave[f], s->wave_win[g] + (size_t)pe * s->hop, s->cw_wave[g], (size_t)s->Tw * s->hop, (size_t)s->n_feat * s->hop, 1, st)) return -1;
  e->launches += 6;
  ConvertBuffers cb;
  cb.d_wave = s->cw_wave[g]; cb.d_f0 = s->cw_f0[g]; cb.d_ap = s->cw_ap[g]; cb.d_mc = s->cw_mc[g]; cb.d_voiced = s->cw_voiced[g];
  cb.d_mse = s->d_mse; cb.d_mask = s->d_mask; cb.d_index = s->d_index; cb.d_count = s->d_count;
  cb.d_mc_out = s->cv_mc_out; cb.d_f0_out = s->cv_f0_out; cb.d_ap_out = s->cv_ap_out; cb.d_sp_mid = s->cv_sp_mid;
  cb.d_sp_out = s->cv_sp_out; cb.d_voiced_out = s->cv_voiced_out;
  if (convert_window_device(e, cb, s->Tw, s->Tw * s->hop, c.fft_length, s->hop, c.threshold_db, c.order, c.fft_length, st)) return -1;
  // ---- decode window: central n_feat converted frames ----
  const int pc = s->e_conv;
  if (slide<float>(s->dw_f0[f], s->cv_f0_out + pc, s->dw_f0[g], s->Td, s->n_feat, 1, st)) return -1;
  if (slide<float>(s->dw_ap[f], s->cv_ap_out + (size_t)pc * s->nb, s->dw_ap[g], s->Td, s->n_feat, s->nb, st)) return -1;
  if (slide<float>(s->dw_sp[f], s->cv_sp_out + (size_t)pc * s->nb, s->dw_sp[g], s->Td, s->n_feat, s->nb, st)) return -1;
  k_f32_to_f64<<<(s->Td + 127) / 128, 128, 0, st>>>(s->dw_f0[g], s->dec_f0_f64, s->Td);
  e->launches += 4;
  if (synth_add_async(e, s->synth, s->dec_f0_f64, s->Td, s->dw_sp[g], s->dw_ap[g], st)) return -1;
  int max_blocks = out_capacity / c.vocoder_buffer_size;
  if (max_blocks > s->max_blocks) max_blocks = s->max_blocks;
  RYK_CHECK(max_blocks > 0, "output capacity is smaller than one synthesizer block");
  if (synth_drain_async(e, s->synth, d_out, max_blocks, st)) return -1;
  k_scrub<<<8, 256, 0, st>>>(d_out, s->synth->dev.state, c.vocoder_buffer_size, max_blocks * c.vocoder_buffer_size, d_n_out);
  e->launches += 1;
  RYK_CUDA(cudaGetLastError());
  s->flip = g;
  s->step++;
  return 0;
}

}  // namespace ryk

using namespace ryk;
struct ryk_engine { Engine impl; };

extern "C" {

int ryk_session_create(ryk_engine* h, const ryk_session_config* cfg, int* session_id) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  RYK_CHECK(cfg && session_id, "null argument");
  RYK_CHECK(e->stage1 && e->stage2, "load both models before creating a session");
  Session* s = new Session();
  s->cfg = *cfg;
  s->hop = (int)(cfg->fs * cfg->frame_period_ms / 1000.0);
  s->rate = (int)lround(1000.0 / cfg->frame_period_ms);
  s->n_wave = (int)lrint(cfg->buffer_time * cfg->fs);
  s->n_feat = (int)lrint(cfg->buffer_time * s->rate);
  s->e_wave = (int)lrint(cfg->encode_extra_time * cfg->fs);
  s->e_enc_frames = (int)lrint(cfg->encode_extra_time * s->rate);
  s->e_conv = (int)lrint(cfg->convert_extra_time * s->rate);
  s->e_dec = (int)lrint(cfg->decode_extra_time * s->rate);
  s->Lw = s->n_wave + 2 * s->e_wave;
  s->Tw = s->n_feat + 2 * s->e_conv;
  s->Td = s->n_feat + 2 * s->e_dec;
  s->nb = cfg->fft_length / 2 + 1;
  s->C = cfg->order + 1;
  RYK_CHECK(s->n_wave == s->n_feat * s->hop && s->e_wave == s->e_enc_frames * s->hop, "buffer_time / encode_extra_time must be whole frames");
  RYK_CHECK(s->Lw / s->hop - 2 * s->e_enc_frames == s->n_feat, "encode window does not trim to one chunk of frames");
  RYK_CHECK(s->nb == 513 && e->stage1->in_ch == s->C, "session configuration does not match the loaded models");
  if (sptk_prepare(e, cfg->order, cfg->alpha, cfg->fft_length)) return -1;
  auto A = [&](void** p, size_t bytes) -> int { RYK_CUDA(cudaMalloc(p, bytes ? bytes : 16)); RYK_CUDA(cudaMemset(*p, 0, bytes ? bytes : 16)); s->allocs.push_back(*p); return 0; };
  const int n_enc = s->Lw / s->hop;
  for (int i = 0; i < 2; ++i) {
    if (A((void**)&s->wave_win[i], sizeof(float) * s->Lw)) return -1;
    if (A((void**)&s->cw_f0[i], sizeof(float) * s->Tw)) return -1;
    if (A((void**)&s->cw_ap[i], sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
    if (A((void**)&s->cw_mc[i], sizeof(float) * (size_t)s->Tw * s->C)) return -1;
    if (A((void**)&s->cw_voiced[i], (size_t)s->Tw)) return -1;
    if (A((void**)&s->cw_wave[i], sizeof(float) * (size_t)s->Tw * s->hop)) return -1;
    if (A((void**)&s->dw_f0[i], sizeof(float) * s->Td)) return -1;
    if (A((void**)&s->dw_ap[i], sizeof(float) * (size_t)s->Td * s->nb)) return -1;
    if (A((void**)&s->dw_sp[i], sizeof(float) * (size_t)s->Td * s->nb)) return -1;
    // silent template mel-cepstrum in the not-yet-filled part of the convert window
    k_fill_rows<float><<<64, 256, 0, e->stream>>>(s->cw_mc[i], s->Tw, s->C, kSilentMc0, 0.f);
  }
  if (A((void**)&s->enc_f0, sizeof(float) * n_enc)) return -1;
  if (A((void**)&s->enc_sp, sizeof(float) * (size_t)n_enc * s->nb)) return -1;
  if (A((void**)&s->enc_ap, sizeof(float) * (size_t)n_enc * s->nb)) return -1;
  if (A((void**)&s->enc_mc, sizeof(float) * (size_t)n_enc * s->C)) return -1;
  if (A((void**)&s->enc_voiced, (size_t)n_enc)) return -1;
  if (A((void**)&s->dec_f0_f64, sizeof(double) * s->Td)) return -1;
  s->max_blocks = (s->Td * s->hop) / cfg->vocoder_buffer_size + 4;
  if (A((void**)&s->out_blocks, sizeof(double) * (size_t)s->max_blocks * cfg->vocoder_buffer_size)) return -1;
  if (A((void**)&s->d_n_out, sizeof(int))) return -1;
  if (A((void**)&s->d_chunk, sizeof(float) * s->n_wave)) return -1;
  if (A((void**)&s->d_mse, sizeof(double) * s->Tw)) return -1;
  if (A((void**)&s->d_mask, (size_t)s->Tw)) return -1;
  if (A((void**)&s->d_index, sizeof(int) * s->Tw)) return -1;
  if (A((void**)&s->d_count, sizeof(int) * 2)) return -1;
  if (A((void**)&s->cv_mc_out, sizeof(float) * (size_t)s->Tw * s->C)) return -1;
  if (A((void**)&s->cv_f0_out, sizeof(float) * s->Tw)) return -1;
  if (A((void**)&s->cv_ap_out, sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
  if (A((void**)&s->cv_sp_mid, sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
  if (A((void**)&s->cv_sp_out, sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
  if (A((void**)&s->cv_voiced_out, (size_t)s->Tw)) return -1;
  RYK_CUDA(cudaMallocHost(&s->h_in, sizeof(float) * s->n_wave));
  RYK_CUDA(cudaMallocHost(&s->h_out, sizeof(double) * (size_t)s->max_blocks * cfg->vocoder_buffer_size));
  RYK_CUDA(cudaMallocHost(&s->h_n, sizeof(int)));
  if (dio_plan_create(e, s->Lw, cfg->fs, cfg->frame_period_ms, cfg->f0_floor, cfg->f0_ceil, &s->dio)) return -1;
  e->dio_plans[std::make_tuple(-(int)e->sessions.size() - 1, cfg->fs, 0, 0, 0)] = s->dio;   // owned by the engine's plan table
  if (synth_create(e, cfg->fs, cfg->frame_period_ms, cheaptrick_fft_size(cfg->fs, 71.0), cfg->vocoder_buffer_size, 4096, &s->synth)) return -1;
  // build the U-Net plans this session can need up front (allocation + tensor maps), not on the first chunk
  UNetPlan* p = nullptr;
  for (int Tp = 128; Tp <= s->Tw + 128; Tp += 128) if (unet_get_plan(e, e->stage1, 1, 1, Tp, e->precision, &p)) return -1;
  if (unet_get_plan(e, e->stage2, 1, s->Tw + (128 - s->Tw % 128), 512, e->precision, &p)) return -1;
  RYK_CUDA(cudaStreamSynchronize(e->stream));
  e->sessions.push_back(s);
  *session_id = (int)e->sessions.size() - 1;
  return 0;
}

int ryk_session_destroy(ryk_engine* h, int id) {
  Engine* e = &h->impl;
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CUDA(cudaStreamSynchronize(e->stream));
  session_free(s);
  e->sessions[id] = nullptr;
  return 0;
}

int ryk_session_push(ryk_engine* h, int id, const float* wave, int n, double* out, int out_capacity, int* n_out) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
  cudaStream_t st = e->stream;
  memcpy(s->h_in, wave, sizeof(float) * n);
  RYK_CUDA(cudaMemcpyAsync(s->d_chunk, s->h_in, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  int cap = out_capacity < s->max_blocks * s->cfg.vocoder_buffer_size ? out_capacity : s->max_blocks * s->cfg.vocoder_buffer_size;
  if (session_step(e, s, s->d_chunk, s->out_blocks, cap, s->d_n_out, st)) return -1;
  RYK_CUDA(cudaMemcpyAsync(s->h_n, s->d_n_out, sizeof(int), cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaMemcpyAsync(s->h_out, s->out_blocks, sizeof(double) * (size_t)(cap / s->cfg.vocoder_buffer_size) * s->cfg.vocoder_buffer_size,
                           cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaStreamSynchronize(st));
  int produced = *s->h_n;
  memcpy(out, s->h_out, sizeof(double) * produced);
  *n_out = produced;
  return 0;
}

int ryk_session_push_device(ryk_engine* h, int id, const float* wave_dev, int n, double* out_dev, int out_capacity, int* n_out_dev) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
  return session_step(e, s, wave_dev, out_dev, out_capacity, n_out_dev, e->stream);
}

}  // extern "C"
