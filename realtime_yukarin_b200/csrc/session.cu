// session.cu -- one audio stream's encode -> convert -> decode chain kept resident in HBM and run as a
// three-stage software pipeline on three CUDA streams.
//
// Semantics = the reference's three StreamWrapper-driven stages as the workers run them
// (realtime_voice_conversion/worker/{encode,convert,decode}_worker.py with stream/*.py): stage input chunk j
// is added at start_time = extra + j*T and step k processes [k*T - extra, k*T + T + extra), i.e. window item i
// of step k is item k*n - 2e + i of the stage's input sequence, silent where that is negative (SURVEY A.9a,
// verified against the reference's BaseStream.fetch).  Instead of Python segment lists each stage keeps its
// last window on the device and slides it by one chunk per step:
//   wave window      (n_wave + 2 e_wave samples)                                -> WORLD analysis -> trim
//   feature window   (n_feat + 2 e_conv frames of f0/ap/mc/voiced + aligned samples) -> gate, stage 1, stage 2 -> trim
//   converted window (n_feat + 2 e_dec frames of f0/ap/sp)                      -> realtime synthesizer -> NaN scrub
//
// Pipelining = what run.py does with three OS processes and queues (run.py:58-93), done with streams and events:
//   stream E: slide wave, silence gate (needs only the samples), DIO/StoneMask/CheapTrick/D4C          of chunk k+1
//   stream C: slide features, stage-1 U-Net (+f0 map), mc2sp                                          of chunk k
//   stream C2: stage-2 U-Net (the tcgen05 layers)                                                     of chunk k-1
//   stream D: slide converted features, synthesizer add/plan/pulse/overlap-add, NaN scrub             of chunk k-2
// Inter-stage buffers are double-buffered (index = step parity); events order producer/consumer and guard reuse.
// The only host<->device handshake inside a step is the 8-byte effective-frame count that selects the stage-1
// plan (T_eff + 128 - T_eff % 128); the gate runs first in stream E so the count is on the host long before
// stream C needs it.
#include <math.h>
#include <stdio.h>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <vector>

#include "../../include/ryk.h"
#include "engine.h"
#include "features.h"
#include "synth.h"
#include "unet.h"

namespace ryk {

constexpr int kRing = 8;          // event / output-slot ring (pipeline depth is bounded by the buffer guards below)

struct StageGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; };
struct Group;
struct Session {
  // optional per-stage device timing (RYK_STAGE_TIMES=1): [stage E1,E2,S1,S2,D][begin/end][ring]
  cudaEvent_t tev[5][2][kRing]; bool stage_times = false;
  int owner = 0;                               // plan-cache owner id (activation buffers are private to the session)
  float* d_colmin[2] = {nullptr, nullptr};
  Group* group = nullptr; int slot = 0;        // member of a batched stage-2 group (config 5), else nullptr
  cudaEvent_t ev_pro[kRing];                   // stage-2 prologue of step r done (group members only)
  ryk_session_config cfg;
  int hop, rate, n_wave, n_feat, e_wave, e_enc_frames, e_conv, e_dec;
  int Lw, Tw, Td, nb, C;
  long long step = 0;              // chunks submitted
  long long collected = 0;         // chunks collected through the host API
  cudaStream_t sE = nullptr, sC = nullptr, sD = nullptr;     // gate | stage 1 | decode
  // stage 2 of even / odd chunks on two streams with two activation plans (RYK_S2_ALT=0: one): the latency-bound bottleneck layers
  // (c4-d2: 30 % of a forward's time, a few CTAs each) of one chunk overlap the GPU-filling layers of its neighbour
  cudaStream_t sC2s[2] = {nullptr, nullptr}; bool two_s2 = false;
  cudaStream_t sA[2] = {nullptr, nullptr};   // WORLD analysis of even / odd chunks: two chunks' analyses may be in flight
  cudaEvent_t ev_gate[kRing];
  cudaEvent_t ev_count[kRing], ev_enc[kRing], ev_cslide[kRing], ev_s1[kRing], ev_conv[kRing], ev_dslide[kRing], ev_dec[kRing];
  // sliding windows, double-buffered by step parity
  float* wave_win[2];
  float *cw_f0[2], *cw_ap[2], *cw_mc[2], *cw_wave[2]; uint8_t* cw_voiced[2];
  float *dw_f0[2], *dw_ap[2], *dw_sp[2];
  // inter-stage buffers, double-buffered by step parity
  float *enc_f0[2], *enc_sp[2], *enc_ap[2], *enc_mc[2]; uint8_t* enc_voiced[2];
  double* d_mse; uint8_t* d_mask[2]; int* d_index[2]; int* d_count[2];
  float *cv_mc_out[2], *cv_f0_out[2], *cv_ap_out[2], *cv_sp_out[2]; uint8_t* cv_voiced_out[2];
  float* cv_sp_mid[2];
  double* dec_f0_f64;
  int max_blocks;
  // host-API staging rings (pinned host + device), slot = step % kRing
  float* d_chunk[kRing]; float* h_in[kRing];
  double* d_out[kRing]; double* h_out[kRing];
  int* d_n_out[kRing]; int* h_n[kRing];
  int* h_count[kRing];
  float* d_chunk_fixed = nullptr;      // the chunk the (captured) encode graph reads
  double* d_out_fixed[2];              // blocks written by the (captured) decode graph, by parity
  int* d_n_fixed[2];
  bool use_graphs = true;
  bool merge_s2 = false;           // this step: stage-2 prologue + 16 layers + epilogue replayed as ONE graph (no profiling events in between)
  bool host_prof = false; double host_wait_us = 0.0, host_total_us = 0.0; long long host_steps = 0;   // RYK_HOST_PROF=1
  std::map<int, StageGraph> graphs;
  // stage 1 with the padded-length bucket chosen ON THE DEVICE: one graph per chunk parity = {k_set_bucket -> SWITCH conditional node
  // whose body i is the stage-1 sequence for the padded length 128 i (0: no effective frame)}; no host sync on the submit path
  cudaGraphExec_t s1_switch[2] = {nullptr, nullptr}; long long s1_switch_launches[2][16]; int s1_buckets = 0; int last_bucket = 0;
  bool device_buckets = false;
  Synth* synth = nullptr;
  DioPlan* dio[2] = {nullptr, nullptr};     // one analysis plan per chunk parity
  std::vector<void*> allocs, pinned;
};

// Several sessions on one GPU sharing ONE batched stage-2 forward per step (BASELINE config 5: 8 streams per GPU,
// stage-2 input (B, 1, Tp, 512)).  Everything else (analysis, gate, stage 1, synthesis) stays per stream: those
// stages carry per-stream state and data-dependent lengths, and they are a small share of the SM time.
struct Group {
  std::vector<Session*> members;
  int Tp = 0, owner = 0;
  UNetPlan* p2 = nullptr;                      // stage-2 plan at batch = members.size()
  cudaStream_t sG = nullptr;
  cudaEvent_t ev_fwd[kRing];                   // batched forward of step r done
  long long step = 0, collected = 0;
  StageGraph fwd_graph;
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

// several windows slid by one launch (blockIdx.y = window): dst = [old[shift..L), new[0..shift)] in units of `elem` bytes
struct SlideDesc { const void* old_; const void* new_; void* dst; size_t L, shift, row; int elem; };
struct SlideBatch { SlideDesc d[5]; int n; };
__global__ void k_slide_multi(SlideBatch b) {
  const SlideDesc& d = b.d[blockIdx.y];
  size_t total = d.L * d.row;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / d.row;
    bool keep = r + d.shift < d.L;
    size_t src = keep ? i + d.shift * d.row : i - (d.L - d.shift) * d.row;
    if (d.elem == 4) ((uint32_t*)d.dst)[i] = keep ? ((const uint32_t*)d.old_)[src] : ((const uint32_t*)d.new_)[src];
    else ((uint8_t*)d.dst)[i] = keep ? ((const uint8_t*)d.old_)[src] : ((const uint8_t*)d.new_)[src];
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

// ---- output re-blocker + silence gate (realtime_voice_conversion/worker/decode_worker.py:38-59) ----
// The reference's decode worker concatenates the synthesizer blocks into `wave_fragment`, cuts one out_audio_chunk off its
// front whenever enough samples are queued (at most one per step) and drops the chunk when its mean STFT power is below
// -output_silent_threshold dB.  Here the fragment lives in HBM (ping-pong buffers), the sample count of a step is read from
// device memory (no host sync) and the gate kernels (world_synth.cu) are predicated on the device-side "chunk emitted" flag.
struct ReblockState { int len, sel, overflow; };
struct Reblock {
  int chunk = 0, max_in = 0, cap = 0, n_fft = 2048, hop = 512;
  double threshold_db = 80.0;
  ReblockState* d_state = nullptr;
  double* d_frag[2] = {nullptr, nullptr};
  double* d_scratch = nullptr;
  double* d_chunk[kRing]; int* d_nvalid[kRing]; int* d_status[kRing]; double* d_power[kRing];
  double* h_chunk[kRing]; int* h_status[kRing]; double* h_power[kRing]; int* h_overflow[kRing];
  cudaEvent_t ev[kRing];
  double* d_stage_in = nullptr; int* d_stage_n = nullptr;     // staging for the host-buffer entry point
  long long pushed = 0;
  std::vector<void*> allocs, pinned;
};

static void reblock_free(Reblock* R) {
  if (!R) return;
  for (int i = 0; i < kRing; ++i) if (R->ev[i]) cudaEventDestroy(R->ev[i]);
  for (void* p : R->allocs) cudaFree(p);
  for (void* p : R->pinned) cudaFreeHost(p);
  delete R;
}

__global__ void __launch_bounds__(1024) k_reblock(ReblockState* __restrict__ st, double* __restrict__ frag0, double* __restrict__ frag1, int cap,
                                                 const double* __restrict__ in, const int* __restrict__ n_in_p, int max_in, int chunk,
                                                 double* __restrict__ out, int* __restrict__ n_valid) {
  __shared__ int sh_len, sh_sel;
  if (threadIdx.x == 0) { sh_len = st->len; sh_sel = st->sel; }
  __syncthreads();
  const int len = sh_len, sel = sh_sel;
  int n_in = *n_in_p;
  if (n_in < 0) n_in = 0;
  int clamped = 0;
  if (n_in > max_in) { n_in = max_in; clamped = 1; }
  double* cur = sel ? frag1 : frag0;
  double* nxt = sel ? frag0 : frag1;
  const int new_len = len + n_in;
  if (new_len >= chunk) {
    int rest = new_len - chunk, over = 0;
    if (rest > cap) { rest = cap; over = 1; }
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) out[i] = i < len ? cur[i] : in[i - len];
    for (int j = threadIdx.x; j < rest; j += blockDim.x) { const int i = chunk + j; nxt[j] = i < len ? cur[i] : in[i - len]; }
    if (threadIdx.x == 0) { st->len = rest; st->sel = sel ^ 1; st->overflow |= over | clamped; *n_valid = chunk; }
  } else {
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) cur[len + i] = in[i];
    if (threadIdx.x == 0) { st->len = new_len; st->overflow |= clamped; *n_valid = 0; }
  }
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

static int slide_batch(SlideBatch& b, cudaStream_t st) {
  size_t mx = 0;
  for (int i = 0; i < b.n; ++i) { size_t t = b.d[i].L * b.d[i].row; if (t > mx) mx = t; }
  if (mx == 0) return 0;
  int blocks = (int)((mx + 255) / 256); if (blocks > 592) blocks = 592;
  k_slide_multi<<<dim3(blocks, b.n), 256, 0, st>>>(b);
  RYK_CUDA(cudaGetLastError());
  return 0;
}
template <typename T>
static void slide_add(SlideBatch& b, const T* old_, const T* new_, T* dst, size_t L, size_t shift, size_t row) {
  SlideDesc& d = b.d[b.n++];
  d.old_ = old_; d.new_ = new_; d.dst = dst; d.L = L; d.shift = shift; d.row = row; d.elem = (int)sizeof(T);
}

static Session* get_session(Engine* e, int id) { return (id >= 0 && id < (int)e->sessions.size()) ? e->sessions[id] : nullptr; }

static void session_free(Session* s) {
  if (!s) return;
  if (s->host_prof && s->host_steps > 0)
    fprintf(stderr, "[ryk host prof] %lld steps: %.1f us per step on the host, of which %.1f us waiting for the gate count\n", s->host_steps,
            s->host_total_us / s->host_steps, s->host_wait_us / s->host_steps);
  for (cudaStream_t st : {s->sE, s->sA[0], s->sA[1], s->sC, s->sC2s[0], s->sC2s[1], s->sD}) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (int i = 0; i < kRing; ++i)
    for (cudaEvent_t ev : {s->ev_gate[i], s->ev_pro[i], s->ev_count[i], s->ev_enc[i], s->ev_cslide[i], s->ev_s1[i], s->ev_conv[i], s->ev_dslide[i], s->ev_dec[i]}) if (ev) cudaEventDestroy(ev);
  for (auto& kv : s->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (int b = 0; b < 2; ++b) if (s->s1_switch[b]) cudaGraphExecDestroy(s->s1_switch[b]);
  if (s->stage_times) for (int a = 0; a < 5; ++a) for (int w = 0; w < 2; ++w) for (int i = 0; i < kRing; ++i) cudaEventDestroy(s->tev[a][w][i]);
  for (void* p : s->allocs) cudaFree(p);
  for (void* p : s->pinned) cudaFreeHost(p);
  synth_destroy(s->synth);
  delete s;
}

static void group_free(Group* G) {
  if (!G) return;
  if (G->sG) { cudaStreamSynchronize(G->sG); cudaStreamDestroy(G->sG); }
  for (int i = 0; i < kRing; ++i) if (G->ev_fwd[i]) cudaEventDestroy(G->ev_fwd[i]);
  if (G->fwd_graph.exec) cudaGraphExecDestroy(G->fwd_graph.exec);
  for (Session* m : G->members) {
    m->group = nullptr;
    for (int key = 40; key < 46; ++key) {          // the captured stage-2 prologue / epilogue graphs point into the group's plan
      auto it = m->graphs.find(key);
      if (it != m->graphs.end()) { if (it->second.exec) cudaGraphExecDestroy(it->second.exec); m->graphs.erase(it); }
    }
  }
  delete G;
}

void session_destroy_all(Engine* e) {
  for (Reblock* R : e->reblocks) reblock_free(R);
  e->reblocks.clear();
  for (Group* G : e->groups) group_free(G);
  e->groups.clear();
  for (Session* s : e->sessions) session_free(s);
  e->sessions.clear();
}

// make every session stream wait for what is already queued on the engine's main stream
int session_streams_fork(Engine* e, cudaEvent_t ev) {
  for (Session* s : e->sessions) {
    if (!s) continue;
    for (cudaStream_t st : {s->sE, s->sA[0], s->sA[1], s->sC, s->sC2s[0], s->sC2s[1], s->sD}) RYK_CUDA(cudaStreamWaitEvent(st, ev, 0));
  }
  for (Group* G : e->groups) if (G) RYK_CUDA(cudaStreamWaitEvent(G->sG, ev, 0));
  return 0;
}
// make the engine's main stream wait for everything queued on the session streams
int session_streams_join(Engine* e) {
  for (Session* s : e->sessions) {
    if (!s) continue;
    for (cudaStream_t st : {s->sE, s->sA[0], s->sA[1], s->sC, s->sC2s[0], s->sC2s[1], s->sD}) {
      cudaEvent_t ev;
      RYK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      RYK_CUDA(cudaEventRecord(ev, st));
      RYK_CUDA(cudaStreamWaitEvent(e->stream, ev, 0));
      RYK_CUDA(cudaEventDestroy(ev));
    }
  }
  for (Group* G : e->groups) {
    if (!G) continue;
    cudaEvent_t ev;
    RYK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RYK_CUDA(cudaEventRecord(ev, G->sG));
    RYK_CUDA(cudaStreamWaitEvent(e->stream, ev, 0));
    RYK_CUDA(cudaEventDestroy(ev));
  }
  return 0;
}

// ---- CUDA-graph cache ---------------------------------------------------------------------------------
// Every stage of a step is a fixed kernel sequence over fixed buffers (selected by chunk parity, and for stage 1 by the
// padded effective length), so each variant is stream-captured once and replayed: a step costs ~6 graph launches on
// the host instead of ~90 kernel launches (the host was the bottleneck at 0.75 ms of launch overhead per 0.78 ms step).
// RYK_SESSION_SKIP (timing experiments only; results are garbage): bit 0 analysis (E2), 1 stage 1, 2 stage-2 layers 1..14, 3 synthesis
// Compiled in only with -DRYK_DIAG (tools/gpu_skip_sweep.sh builds a separate diagnostics library): a release libryk.so has no
// knob that can turn the timed path into a partial one.
static int session_skip_mask() {
#ifdef RYK_DIAG
  static int m = -1;
  if (m < 0) { const char* v = getenv("RYK_SESSION_SKIP"); m = v ? atoi(v) : 0; }
  return m;
#else
  return 0;
#endif
}
template <typename F>
static int run_stage(Engine* e, Session* s, int key, cudaStream_t st, F&& body, bool capture_only = false) {
  if (const int m = session_skip_mask()) {
    if (((m & 1) && (key == 2 || key == 3)) || ((m & 2) && key >= 4 && key < 40) || ((m & 4) && (key == 42 || key == 43)) || ((m & 8) && key >= 46 && key <= 49))
      return 0;
  }
  if (!s->use_graphs) return body();
  StageGraph& g = s->graphs[key];
  if (!g.exec) {
    long long before = e->launches;
    cudaGraph_t graph = nullptr;
    RYK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = body();
    cudaError_t err = cudaStreamEndCapture(st, &graph);
    if (rc) return rc;
    RYK_CUDA(err);
    RYK_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
    RYK_CUDA(cudaGraphDestroy(graph));
    g.launches = e->launches - before;
    e->launches = before;
  }
  if (capture_only) return 0;
  RYK_CUDA(cudaGraphLaunch(g.exec, st));
  e->launches += g.launches;
  return 0;
}

enum { G_E1 = 0, G_E2 = 2, G_S1 = 4 /* + 2 * bucket + parity; bucket 0 = no effective frame, else padded length / 128 (1..15) */,
       G_S2A = 40, G_S2B = 42, G_S2C = 44, G_D = 46, G_D1 = 48, G_S2M = 50 };

// value of the SWITCH node = padded effective length / 128 (count[1] / 128), 0 when no frame is effective (count[0] == 0)
__global__ void k_set_bucket(cudaGraphConditionalHandle handle, const int* __restrict__ count, int n_buckets) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int v = count[0] > 0 ? count[1] / 128 : 0;
    if (v < 0 || v >= n_buckets) v = n_buckets - 1;        // cannot happen (count[1] <= Tw + 128); keeps the node in range
    cudaGraphSetConditional(handle, (unsigned)v);
  }
}

static int stage1_enqueue(Engine* e, Session* s, int b, int tp1, bool capture_only);

// Build the per-parity stage-1 graph with a device-side switch over the padded-length buckets (CUDA conditional nodes, 12.8+).
static int stage1_build_switch(Engine* e, Session* s, int b) {
  const int n_buckets = (s->Tw + (128 - s->Tw % 128)) / 128 + 1;
  RYK_CHECK(n_buckets <= 16, "window too long for the stage-1 graph table");
  s->s1_buckets = n_buckets;
  cudaGraph_t graph = nullptr;
  RYK_CUDA(cudaGraphCreate(&graph, 0));
  cudaGraphConditionalHandle handle;
  RYK_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
  // node 1: the setter
  cudaGraphNode_t set_node = nullptr;
  {
    cudaKernelNodeParams kp = {};
    const int* cnt = s->d_count[b];
    int nb_ = n_buckets;
    void* args[3] = {(void*)&handle, (void*)&cnt, (void*)&nb_};
    kp.func = (void*)k_set_bucket; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
    RYK_CUDA(cudaGraphAddKernelNode(&set_node, graph, nullptr, 0, &kp));
  }
  // node 2: SWITCH
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = handle;
  cp.conditional.type = cudaGraphCondTypeSwitch;
  cp.conditional.size = (unsigned)n_buckets;
  cudaGraphNode_t sw = nullptr;
  RYK_CUDA(cudaGraphAddNode(&sw, graph, &set_node, 1, &cp));
  const bool saved = s->use_graphs;
  for (int i = 0; i < n_buckets; ++i) {
    cudaGraph_t body = cp.conditional.phGraph_out[i];
    const long long before = e->launches;
    RYK_CUDA(cudaStreamBeginCaptureToGraph(s->sC, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    s->use_graphs = false;                               // run the body's kernels straight into the capture
    int rc = stage1_enqueue(e, s, b, i * 128, false);
    s->use_graphs = saved;
    cudaGraph_t out = nullptr;
    cudaError_t err = cudaStreamEndCapture(s->sC, &out);
    if (rc) return rc;
    RYK_CUDA(err);
    s->s1_switch_launches[b][i] = e->launches - before + 1;     // + the setter
    e->launches = before;
  }
  RYK_CUDA(cudaGraphInstantiate(&s->s1_switch[b], graph, 0));
  RYK_CUDA(cudaGraphDestroy(graph));
  return 0;
}

// Stage 1 of a chunk of parity b: slide the feature window, (gather ->) 1-D U-Net at padded length tp1 (0: no effective frame,
// voice_changer.py:32-35 skips the net) -> scatter into the silent template + f0 map, mc2sp.  One graph per (tp1 bucket, parity);
// all of them are captured when the session is created so that no chunk ever pays for a capture in the middle of a stream.
static int stage1_enqueue(Engine* e, Session* s, int b, int tp1, bool capture_only) {
  const int f = b, g = b ^ 1, pe = s->e_enc_frames;
  const ryk_session_config& c = s->cfg;
  return run_stage(e, s, G_S1 + 2 * (tp1 / 128) + b, s->sC, [&]() -> int {
        SlideBatch sb; sb.n = 0;
        slide_add<float>(sb, s->cw_f0[f], s->enc_f0[b] + pe, s->cw_f0[g], s->Tw, s->n_feat, 1);
        slide_add<float>(sb, s->cw_ap[f], s->enc_ap[b] + (size_t)pe * s->nb, s->cw_ap[g], s->Tw, s->n_feat, s->nb);
        slide_add<float>(sb, s->cw_mc[f], s->enc_mc[b] + (size_t)pe * s->C, s->cw_mc[g], s->Tw, s->n_feat, s->C);
        slide_add<uint8_t>(sb, s->cw_voiced[f], s->enc_voiced[b] + pe, s->cw_voiced[g], s->Tw, s->n_feat, 1);
        if (slide_batch(sb, s->sC)) return -1;
        e->launches += 1;
        const float* d_y = nullptr;
        if (tp1 > 0) {
          UNetPlan* p1 = nullptr;
          if (unet_get_plan(e, e->stage1, 1, 1, tp1, e->precision, &p1, s->owner)) return -1;
          if (stage1_prologue_run(e, s->cw_mc[g], s->d_index[b], s->d_count[b], s->C, (float*)p1->d_in, tp1, s->sC)) return -1;
          if (unet_forward(e, p1, s->sC)) return -1;
          d_y = (const float*)p1->d_out;
        }
        if (stage1_epilogue_run(e, d_y, s->d_index[b], s->d_mask[b], s->d_count[b], s->Tw, s->C, s->cw_f0[g], s->cw_ap[g], s->cw_voiced[g], s->nb,
                                kSilentMc0, s->cv_mc_out[b], s->cv_f0_out[b], s->cv_ap_out[b], s->cv_voiced_out[b], s->sC)) return -1;
        return mc2sp_run(e, s->cv_mc_out[b], s->Tw, c.order, c.fft_length, 1e-16, s->cv_sp_mid[b], nullptr, s->sC);
      }, capture_only);
}

// Step k = s->step is enqueued in three parts so that a group can interleave its members:
//   front: streams E and C (analysis, gate, stage 1, mc2sp) and the stage-2 prologue
//   mid:   the stage-2 U-Net forward (single session: on its own stream C2; group: one batched forward on the group stream)
//   back:  stage-2 epilogue and stream D (synthesizer); results land in s->d_out_fixed[b] / s->d_n_fixed[b]; s->step advances.
#define STEP_LOCALS                                                                                         \
  const long long k = s->step;                                                                              \
  const int b = (int)(k & 1), f = b, g = b ^ 1; /* windows: read [f], write [g]; inter-stage sets: [b] */    \
  const int r = (int)(k % kRing);                                                                           \
  const ryk_session_config& c = s->cfg;                                                                     \
  const int pe = s->e_enc_frames, pc = s->e_conv;                                                           \
  (void)f; (void)g; (void)r; (void)c; (void)pe; (void)pc;

#define TSTAMP(stage, which, stream) do { if (s->stage_times) RYK_CUDA(cudaEventRecord(s->tev[stage][which][r], stream)); } while (0)

#define S2_LOCALS cudaStream_t sC2 = s->sC2s[(s->two_s2 && !s->group) ? b : 0]; const int s2_owner = s->owner + ((s->two_s2 && !s->group && b) ? 3000000 : 0); float* d_colmin = s->d_colmin[(s->two_s2 && !s->group) ? b : 0]; (void)s2_owner; (void)d_colmin;

static int session_front(Engine* e, Session* s, const float* d_chunk_user) {
  STEP_LOCALS
  S2_LOCALS

  // ================= stream E: gate + WORLD analysis =================
  RYK_CUDA(cudaMemcpyAsync(s->d_chunk_fixed, d_chunk_user, sizeof(float) * s->n_wave, cudaMemcpyDeviceToDevice, s->sE));
  if (k >= 2) {
    RYK_CUDA(cudaStreamWaitEvent(s->sE, s->ev_s1[(k - 2) % kRing], 0));      // mask/index/count[b]: last read by stage 1 of k-2
    RYK_CUDA(cudaStreamWaitEvent(s->sE, s->ev_enc[(k - 2) % kRing], 0));     // wave_win[g]: last read by the analysis of k-2
  }
  TSTAMP(0, 0, s->sE);
  if (run_stage(e, s, G_E1 + b, s->sE, [&]() -> int {
        if (slide<float>(s->wave_win[f], s->d_chunk_fixed, s->wave_win[g], s->Lw, s->n_wave, 1, s->sE)) return -1;
        if (slide<float>(s->cw_wave[f], s->wave_win[g] + (size_t)pe * s->hop, s->cw_wave[g], (size_t)s->Tw * s->hop, (size_t)s->n_feat * s->hop, 1, s->sE)) return -1;
        e->launches += 2;
        return gate_mask_run(e, s->cw_wave[g], s->Tw * s->hop, c.fft_length, s->hop, c.threshold_db, s->Tw, s->d_mse, s->d_mask[b], s->d_index[b],
                             s->d_count[b], s->sE);
      })) return -1;
  TSTAMP(0, 1, s->sE);
  RYK_CUDA(cudaMemcpyAsync(s->h_count[r], s->d_count[b], sizeof(int) * 2, cudaMemcpyDeviceToHost, s->sE));
  RYK_CUDA(cudaEventRecord(s->ev_count[r], s->sE));
  RYK_CUDA(cudaEventRecord(s->ev_gate[r], s->sE));

  // ================= stream A[b]: WORLD analysis (the chunks of one parity share a plan and a stream) =================
  cudaStream_t sA = s->sA[b];
  RYK_CUDA(cudaStreamWaitEvent(sA, s->ev_gate[r], 0));
  if (k >= 2) RYK_CUDA(cudaStreamWaitEvent(sA, s->ev_cslide[(k - 2) % kRing], 0));   // enc_*[b] consumed by stage 1 of k-2
  TSTAMP(1, 0, sA);
  if (run_stage(e, s, G_E2 + b, sA, [&]() -> int {
        if (dio_stonemask_run(e, s->dio[b], s->wave_win[g], sA)) return -1;
        const int n_enc = s->Lw / s->hop;
        e->launches += 13;
        return spectral_analysis_run(e, s->wave_win[g], s->Lw, c.fs, c.frame_period_ms, dio_plan_f0(s->dio[b]), n_enc, c.fft_length, c.order,
                                     s->enc_sp[b], s->enc_ap[b], s->enc_mc[b], s->enc_f0[b], s->enc_voiced[b], sA);
      })) return -1;
  TSTAMP(1, 1, sA);
  RYK_CUDA(cudaEventRecord(s->ev_enc[r], sA));

  // ================= stream C: stage 1 (+ f0 map, mc2sp) =================
  RYK_CUDA(cudaStreamWaitEvent(s->sC, s->ev_enc[r], 0));
  if (k >= 2) {
    RYK_CUDA(cudaStreamWaitEvent(s->sC, s->ev_dslide[(k - 2) % kRing], 0));   // cv_{f0,ap,voiced}_out[b] consumed by decode k-2
    RYK_CUDA(cudaStreamWaitEvent(s->sC, s->ev_conv[(k - 2) % kRing], 0));     // cv_sp_mid[b] consumed by stage 2 of k-2
  }
  TSTAMP(2, 0, s->sC);
  if (s->device_buckets) {
    // launch-count bookkeeping only (never waits): the newest count that has already arrived tells which body ran last
    for (int back = 1; back <= 3 && k - back >= 0; ++back) {
      const int rr = (int)((k - back) % kRing);
      if (cudaEventQuery(s->ev_count[rr]) == cudaSuccess) { s->last_bucket = s->h_count[rr][0] > 0 ? s->h_count[rr][1] / 128 : 0; break; }
    }
    if (s->last_bucket < 0 || s->last_bucket >= s->s1_buckets) s->last_bucket = s->s1_buckets - 1;
    RYK_CUDA(cudaGraphLaunch(s->s1_switch[b], s->sC));
    e->launches += s->s1_switch_launches[b][s->last_bucket];
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    RYK_CUDA(cudaEventSynchronize(s->ev_count[r]));                            // effective-frame count of THIS step (RYK_NO_GRAPH / RYK_HOST_BUCKETS path)
    if (s->host_prof) s->host_wait_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    const int t_eff = s->h_count[r][0], tp1 = s->h_count[r][1];
    RYK_CHECK(tp1 / 128 < 16, "window too long for the stage-1 graph table");
    if (stage1_enqueue(e, s, b, t_eff > 0 ? tp1 : 0, false)) return -1;
  }
  // NB: enc_*[b] may be overwritten by encode k+2 once this stage's slides ran; the stage-1 graph is short, so the
  // guard event is simply the end of the stage.
  TSTAMP(2, 1, s->sC);
  RYK_CUDA(cudaEventRecord(s->ev_cslide[r], s->sC));
  RYK_CUDA(cudaEventRecord(s->ev_s1[r], s->sC));

  // ================= stream C2: stage-2 prologue =================
  RYK_CUDA(cudaStreamWaitEvent(sC2, s->ev_s1[r], 0));
  if (k >= 2) RYK_CUDA(cudaStreamWaitEvent(sC2, s->ev_dslide[(k - 2) % kRing], 0));  // cv_sp_out[b] consumed by decode k-2
  const int Tp = s->Tw + (128 - s->Tw % 128);
  TSTAMP(3, 0, sC2);
  if (s->group) {
    Group* G = s->group;
    if (G->step >= 1) RYK_CUDA(cudaStreamWaitEvent(sC2, G->ev_fwd[(G->step - 1) % kRing], 0));   // batched input read by forward k-1
    float* dst = (float*)G->p2->d_in + (size_t)s->slot * Tp * 512;
    if (run_stage(e, s, G_S2A + b, sC2, [&]() -> int { return sr_prologue_run(e, s->cv_sp_mid[b], s->Tw, Tp, s->nb, dst, sC2, d_colmin); })) return -1;
    RYK_CUDA(cudaEventRecord(s->ev_pro[r], sC2));
  } else {
    UNetPlan* p2 = nullptr;
    if (unet_get_plan(e, e->stage2, 1, Tp, 512, e->precision, &p2, s2_owner)) return -1;
    if (!s->merge_s2 && run_stage(e, s, G_S2A + b, sC2, [&]() -> int {
          if (sr_prologue_run(e, s->cv_sp_mid[b], s->Tw, Tp, s->nb, (float*)p2->d_in, sC2, d_colmin)) return -1;
          return unet_forward(e, p2, sC2, 0, 0);
        })) return -1;
  }
  return 0;
}

// single session: stage-2 layers 1..14 (the tcgen05 layers) on the session's own stream
static int session_mid_single(Engine* e, Session* s, bool was_profiling) {
  STEP_LOCALS
  S2_LOCALS
  const int Tp = s->Tw + (128 - s->Tw % 128);
  UNetPlan* p2 = nullptr;
  if (unet_get_plan(e, e->stage2, 1, Tp, 512, e->precision, &p2, s2_owner)) return -1;
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (was_profiling) { RYK_CUDA(cudaEventCreate(&pe0)); RYK_CUDA(cudaEventCreate(&pe1)); RYK_CUDA(cudaEventRecord(pe0, sC2)); }
  if (s->merge_s2) {
    // whole stage 2 as one graph: no launch gaps between prologue, the 16 layers and the epilogue (PDL chains through)
    return run_stage(e, s, G_S2M + b, sC2, [&]() -> int {
      if (sr_prologue_run(e, s->cv_sp_mid[b], s->Tw, Tp, s->nb, (float*)p2->d_in, sC2, d_colmin)) return -1;
      if (unet_forward(e, p2, sC2, 0, 15)) return -1;
      return sr_epilogue_run(e, (const float*)p2->d_out, s->Tw, s->nb, s->cv_sp_out[b], sC2);
    });
  }
  if (run_stage(e, s, G_S2B + b, sC2, [&]() -> int { return unet_forward(e, p2, sC2, 1, 14); })) return -1;
  if (was_profiling) { RYK_CUDA(cudaEventRecord(pe1, sC2)); e->prof_events.emplace_back(pe0, pe1); }
  return 0;
}

static int session_back(Engine* e, Session* s) {
  STEP_LOCALS
  S2_LOCALS
  const int Tp = s->Tw + (128 - s->Tw % 128);
  if (s->group) {
    Group* G = s->group;
    RYK_CUDA(cudaStreamWaitEvent(sC2, G->ev_fwd[G->step % kRing], 0));
    const float* src = (const float*)G->p2->d_out + (size_t)s->slot * Tp * 512;
    if (run_stage(e, s, G_S2C + b, sC2, [&]() -> int { return sr_epilogue_run(e, src, s->Tw, s->nb, s->cv_sp_out[b], sC2); })) return -1;
  } else {
    UNetPlan* p2 = nullptr;
    if (unet_get_plan(e, e->stage2, 1, Tp, 512, e->precision, &p2, s2_owner)) return -1;
    if (!s->merge_s2 && run_stage(e, s, G_S2C + b, sC2, [&]() -> int {
          if (unet_forward(e, p2, sC2, 15, 15)) return -1;
          return sr_epilogue_run(e, (const float*)p2->d_out, s->Tw, s->nb, s->cv_sp_out[b], sC2);
        })) return -1;
  }
  TSTAMP(3, 1, sC2);
  RYK_CUDA(cudaEventRecord(s->ev_conv[r], sC2));

  // ================= stream D: realtime synthesizer =================
  RYK_CUDA(cudaStreamWaitEvent(s->sD, s->ev_conv[r], 0));
  TSTAMP(4, 0, s->sD);
  if (synth_host_advance(e, s->synth, s->Td, s->sD)) return -1;
  const int max_blocks = s->max_blocks;
  if (run_stage(e, s, G_D1 + b, s->sD, [&]() -> int {
        SlideBatch sb; sb.n = 0;
        slide_add<float>(sb, s->dw_f0[f], s->cv_f0_out[b] + pc, s->dw_f0[g], s->Td, s->n_feat, 1);
        slide_add<float>(sb, s->dw_ap[f], s->cv_ap_out[b] + (size_t)pc * s->nb, s->dw_ap[g], s->Td, s->n_feat, s->nb);
        slide_add<float>(sb, s->dw_sp[f], s->cv_sp_out[b] + (size_t)pc * s->nb, s->dw_sp[g], s->Td, s->n_feat, s->nb);
        if (slide_batch(sb, s->sD)) return -1;
        k_f32_to_f64<<<(s->Td + 127) / 128, 128, 0, s->sD>>>(s->dw_f0[g], s->dec_f0_f64, s->Td);
        e->launches += 2;
        RYK_CUDA(cudaGetLastError());
        return 0;
      })) return -1;
  // the converted features of this parity are free again as soon as they sit in the decode window
  RYK_CUDA(cudaEventRecord(s->ev_dslide[r], s->sD));
  if (run_stage(e, s, G_D + b, s->sD, [&]() -> int {
        if (synth_add_kernel(e, s->synth, s->dec_f0_f64, s->Td, s->dw_sp[g], s->dw_ap[g], s->sD)) return -1;
        if (synth_drain_async(e, s->synth, s->d_out_fixed[b], max_blocks, s->sD)) return -1;
        k_scrub<<<8, 256, 0, s->sD>>>(s->d_out_fixed[b], s->synth->dev.state, c.vocoder_buffer_size, max_blocks * c.vocoder_buffer_size, s->d_n_fixed[b]);
        e->launches += 1;
        RYK_CUDA(cudaGetLastError());
        return 0;
      })) return -1;
  TSTAMP(4, 1, s->sD);
  // ev_dec[r] is recorded by the caller after the copies it appends to stream D
  s->step++;
  return 0;
}

static int session_enqueue(Engine* e, Session* s, const float* d_chunk_user) {
  const auto host_t0 = std::chrono::steady_clock::now();
  struct HostProf { Session* s; std::chrono::steady_clock::time_point t0;
    ~HostProf() { if (s->host_prof) { s->host_total_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); s->host_steps++; } } } host_prof_guard{s, host_t0};
  const bool was_profiling = e->profile;
  e->profile = false;                                      // the session places its own timing events (between graph launches)
  // RYK_S2_MERGE=1 (experiment, off by default: +1 % end to end): replay stage 2 as one graph on steps without profiling events
  { static int merge = -1; if (merge < 0) { const char* v = getenv("RYK_S2_MERGE"); merge = v && atoi(v) ? 1 : 0; }
    s->merge_s2 = merge && !was_profiling && !s->group && s->use_graphs && session_skip_mask() == 0; }
  int rc = session_front(e, s, d_chunk_user);
  if (!rc) rc = session_mid_single(e, s, was_profiling);
  if (!rc) rc = session_back(e, s);
  e->profile = was_profiling;
  return rc;
}

// One step of every member + the batched stage-2 forward between their front and back halves.
static int group_enqueue_impl(Engine* e, Group* G, const float* const* d_chunks, bool was_profiling) {
  const int r = (int)(G->step % kRing);
  for (size_t i = 0; i < G->members.size(); ++i)
    if (session_front(e, G->members[i], d_chunks[i])) return -1;
  for (Session* m : G->members) {
    RYK_CUDA(cudaStreamWaitEvent(G->sG, m->ev_pro[m->step % kRing], 0));
    if (m->step >= 1) RYK_CUDA(cudaStreamWaitEvent(G->sG, m->ev_conv[(m->step - 1) % kRing], 0));   // batched output read by epilogue k-1
  }
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (was_profiling) { RYK_CUDA(cudaEventCreate(&pe0)); RYK_CUDA(cudaEventCreate(&pe1)); RYK_CUDA(cudaEventRecord(pe0, G->sG)); }
  {
    StageGraph& g = G->fwd_graph;
    const bool use_graphs = G->members[0]->use_graphs;
    if (!use_graphs) {
      if (unet_forward(e, G->p2, G->sG, 0, 15)) return -1;
    } else {
      if (!g.exec) {
        long long before = e->launches;
        cudaGraph_t graph = nullptr;
        RYK_CUDA(cudaStreamBeginCapture(G->sG, cudaStreamCaptureModeThreadLocal));
        int rc = unet_forward(e, G->p2, G->sG, 0, 15);
        cudaError_t err = cudaStreamEndCapture(G->sG, &graph);
        if (rc) return rc;
        RYK_CUDA(err);
        RYK_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
        RYK_CUDA(cudaGraphDestroy(graph));
        g.launches = e->launches - before;
        e->launches = before;
      }
      RYK_CUDA(cudaGraphLaunch(g.exec, G->sG));
      e->launches += g.launches;
    }
  }
  if (was_profiling) { RYK_CUDA(cudaEventRecord(pe1, G->sG)); e->prof_events.emplace_back(pe0, pe1); }
  RYK_CUDA(cudaEventRecord(G->ev_fwd[r], G->sG));
  for (Session* m : G->members)
    if (session_back(e, m)) return -1;
  G->step++;
  return 0;
}
static int group_enqueue(Engine* e, Group* G, const float* const* d_chunks) {
  const bool was_profiling = e->profile;
  e->profile = false;
  int rc = group_enqueue_impl(e, G, d_chunks, was_profiling);
  e->profile = was_profiling;
  return rc;
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
  memset(s->ev_count, 0, sizeof(s->ev_count)); memset(s->ev_enc, 0, sizeof(s->ev_enc)); memset(s->ev_cslide, 0, sizeof(s->ev_cslide));
  memset(s->ev_s1, 0, sizeof(s->ev_s1)); memset(s->ev_gate, 0, sizeof(s->ev_gate)); memset(s->ev_pro, 0, sizeof(s->ev_pro)); memset(s->ev_conv, 0, sizeof(s->ev_conv)); memset(s->ev_dslide, 0, sizeof(s->ev_dslide)); memset(s->ev_dec, 0, sizeof(s->ev_dec));
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
  // The analysis, stage-1 and synthesis stages are chains of small, latency-bound kernels; stage 2 is bulk work that fills
  // every SM.  Higher stream priority for the former lets their CTAs take freed SM slots first, so their latency does not
  // inflate behind stage-2 waves (the analysis chain is what the host's per-step count sync waits on).
  int prio_lo = 0, prio_hi = 0;
  RYK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));      // lo = least (numerically largest), hi = greatest
  { const char* v = getenv("RYK_NO_PRIORITY"); if (v && atoi(v)) prio_hi = prio_lo; }
  // RYK_PRIO="E,A,C,C2,D" (tuning experiments): priority level of each stream as steps above the lowest (0 .. lo - hi)
  int lv[5] = {prio_lo - prio_hi, prio_lo - prio_hi, prio_lo - prio_hi, 0, prio_lo - prio_hi};
  if (const char* v = getenv("RYK_PRIO")) sscanf(v, "%d,%d,%d,%d,%d", &lv[0], &lv[1], &lv[2], &lv[3], &lv[4]);
  auto PR = [&](int i) { int l = lv[i] < 0 ? 0 : (lv[i] > prio_lo - prio_hi ? prio_lo - prio_hi : lv[i]); return prio_lo - l; };
  RYK_CUDA(cudaStreamCreateWithPriority(&s->sE, cudaStreamNonBlocking, PR(0)));
  for (int i = 0; i < 2; ++i) RYK_CUDA(cudaStreamCreateWithPriority(&s->sA[i], cudaStreamNonBlocking, PR(1)));
  RYK_CUDA(cudaStreamCreateWithPriority(&s->sC, cudaStreamNonBlocking, PR(2)));
  for (int i = 0; i < 2; ++i) RYK_CUDA(cudaStreamCreateWithPriority(&s->sC2s[i], cudaStreamNonBlocking, PR(3)));
  { const char* v = getenv("RYK_S2_ALT"); s->two_s2 = !(v && atoi(v) == 0); }
  RYK_CUDA(cudaStreamCreateWithPriority(&s->sD, cudaStreamNonBlocking, PR(4)));
  for (int i = 0; i < kRing; ++i) {
    cudaEvent_t* evs[] = {&s->ev_gate[i], &s->ev_pro[i], &s->ev_count[i], &s->ev_enc[i], &s->ev_cslide[i], &s->ev_s1[i], &s->ev_conv[i], &s->ev_dslide[i], &s->ev_dec[i]};
    for (cudaEvent_t* ev : evs) RYK_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  }
  { const char* v = getenv("RYK_STAGE_TIMES"); s->stage_times = v && atoi(v) != 0; }
  { const char* v = getenv("RYK_HOST_PROF"); s->host_prof = v && atoi(v) != 0; }
  if (s->stage_times) for (int a = 0; a < 5; ++a) for (int w = 0; w < 2; ++w) for (int i = 0; i < kRing; ++i) RYK_CUDA(cudaEventCreate(&s->tev[a][w][i]));
  // zero-fill on the ENGINE stream: the template fill below (k_fill_rows on e->stream, a non-blocking stream) must be ordered after
  // it -- a legacy-default-stream cudaMemset is not, and could land after the fill (seen once as a 4e-3 RMSE mismatch)
  auto A = [&](void** p, size_t bytes) -> int { RYK_CUDA(cudaMalloc(p, bytes ? bytes : 16)); RYK_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, e->stream)); s->allocs.push_back(*p); return 0; };
  auto P = [&](void** p, size_t bytes) -> int { RYK_CUDA(cudaMallocHost(p, bytes ? bytes : 16)); memset(*p, 0, bytes ? bytes : 16); s->pinned.push_back(*p); return 0; };
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
    if (A((void**)&s->enc_f0[i], sizeof(float) * n_enc)) return -1;
    if (A((void**)&s->enc_sp[i], sizeof(float) * (size_t)n_enc * s->nb)) return -1;
    if (A((void**)&s->enc_ap[i], sizeof(float) * (size_t)n_enc * s->nb)) return -1;
    if (A((void**)&s->enc_mc[i], sizeof(float) * (size_t)n_enc * s->C)) return -1;
    if (A((void**)&s->enc_voiced[i], (size_t)n_enc)) return -1;
    if (A((void**)&s->d_mask[i], (size_t)s->Tw)) return -1;
    if (A((void**)&s->d_index[i], sizeof(int) * s->Tw)) return -1;
    if (A((void**)&s->d_count[i], sizeof(int) * 2)) return -1;
    if (A((void**)&s->cv_mc_out[i], sizeof(float) * (size_t)s->Tw * s->C)) return -1;
    if (A((void**)&s->cv_f0_out[i], sizeof(float) * s->Tw)) return -1;
    if (A((void**)&s->cv_ap_out[i], sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
    if (A((void**)&s->cv_sp_out[i], sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
    if (A((void**)&s->cv_voiced_out[i], (size_t)s->Tw)) return -1;
    if (A((void**)&s->cv_sp_mid[i], sizeof(float) * (size_t)s->Tw * s->nb)) return -1;
  }
  if (A((void**)&s->d_mse, sizeof(double) * s->Tw)) return -1;
  for (int i = 0; i < 2; ++i) if (A((void**)&s->d_colmin[i], sizeof(float) * kColminFloats)) return -1;
  if (A((void**)&s->dec_f0_f64, sizeof(double) * s->Td)) return -1;
  if (A((void**)&s->d_chunk_fixed, sizeof(float) * s->n_wave)) return -1;
  { const char* ng = getenv("RYK_NO_GRAPH"); s->use_graphs = !(ng && atoi(ng) != 0); }
  s->max_blocks = (s->Td * s->hop) / cfg->vocoder_buffer_size + 4;
  const size_t out_samples = (size_t)s->max_blocks * cfg->vocoder_buffer_size;
  for (int i = 0; i < 2; ++i) {
    if (A((void**)&s->d_out_fixed[i], sizeof(double) * out_samples)) return -1;
    if (A((void**)&s->d_n_fixed[i], sizeof(int))) return -1;
  }
  for (int i = 0; i < kRing; ++i) {
    if (A((void**)&s->d_chunk[i], sizeof(float) * s->n_wave)) return -1;
    if (A((void**)&s->d_out[i], sizeof(double) * out_samples)) return -1;
    if (A((void**)&s->d_n_out[i], sizeof(int))) return -1;
    if (P((void**)&s->h_in[i], sizeof(float) * s->n_wave)) return -1;
    if (P((void**)&s->h_out[i], sizeof(double) * out_samples)) return -1;
    if (P((void**)&s->h_n[i], sizeof(int))) return -1;
    if (P((void**)&s->h_count[i], sizeof(int) * 2)) return -1;
  }
  for (int i = 0; i < 2; ++i) {
    if (dio_plan_create(e, s->Lw, cfg->fs, cfg->frame_period_ms, cfg->f0_floor, cfg->f0_ceil, &s->dio[i], e->f0_method)) return -1;
    e->dio_plans[std::make_tuple(-(int)e->sessions.size() - 1, cfg->fs, i, 0, 0)] = s->dio[i];   // owned by the engine's plan table
  }
  if (synth_create(e, cfg->fs, cfg->frame_period_ms, cheaptrick_fft_size(cfg->fs, 71.0), cfg->vocoder_buffer_size, 4096, &s->synth)) return -1;
  // build the U-Net plans this session can need up front (allocation + tensor maps), not on the first chunk
  UNetPlan* p = nullptr;
  s->owner = (int)e->sessions.size() + 1;
  for (int Tp = 128; Tp <= s->Tw + 128; Tp += 128) if (unet_get_plan(e, e->stage1, 1, 1, Tp, e->precision, &p, s->owner)) return -1;
  // (the stage-2 plan is created on first use: a session that joins a group never needs its own)
  RYK_CUDA(cudaStreamSynchronize(e->stream));
  RYK_CUDA(cudaDeviceSynchronize());
  { const char* hb = getenv("RYK_HOST_BUCKETS"); s->device_buckets = s->use_graphs && !(hb && atoi(hb) != 0); }
  if (s->device_buckets) {
    const char* sp = getenv("RYK_S1_PDL");
    const bool pdl_bodies = !(sp && atoi(sp) == 0);
    if (!pdl_bodies) tc_force_pdl(0);
    int rc = 0;
    for (int b = 0; b < 2 && !rc; ++b) rc = stage1_build_switch(e, s, b);
    if (!pdl_bodies) tc_force_pdl(-1);
    if (rc) return rc;
  } else if (s->use_graphs) {
    const long long before = e->launches;
    for (int b = 0; b < 2; ++b)
      for (int tp1 = 0; tp1 <= s->Tw + (128 - s->Tw % 128); tp1 += 128)
        if (stage1_enqueue(e, s, b, tp1, true)) return -1;
    e->launches = before;
  }
  e->sessions.push_back(s);
  *session_id = (int)e->sessions.size() - 1;
  return 0;
}

int ryk_session_destroy(ryk_engine* h, int id) {
  Engine* e = &h->impl;
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(s->group == nullptr, "session belongs to a group: destroy the group first");
  RYK_CUDA(cudaStreamSynchronize(e->stream));
  const int owner = s->owner;
  session_free(s);                         // synchronises the session's streams
  unet_release_owner(e->stage1, owner);
  unet_release_owner(e->stage2, owner);
  unet_release_owner(e->stage2, owner + 3000000);
  e->sessions[id] = nullptr;
  return 0;
}

// Queue one chunk (host samples) without waiting for its output; *ticket identifies it for ryk_session_collect.
int ryk_session_submit(ryk_engine* h, int id, const float* wave, int n, long long* ticket) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
  RYK_CHECK(s->group == nullptr, "session belongs to a group: use ryk_group_submit");
  RYK_CHECK(s->step - s->collected < kRing - 2, "too many chunks in flight: collect before submitting more");
  const long long k = s->step;
  const int r = (int)(k % kRing);
  memcpy(s->h_in[r], wave, sizeof(float) * n);
  RYK_CUDA(cudaMemcpyAsync(s->d_chunk[r], s->h_in[r], sizeof(float) * n, cudaMemcpyHostToDevice, s->sE));
  const int cap = s->max_blocks * s->cfg.vocoder_buffer_size;
  const int b = (int)(k & 1);
  if (session_enqueue(e, s, s->d_chunk[r])) return -1;
  RYK_CUDA(cudaMemcpyAsync(s->h_n[r], s->d_n_fixed[b], sizeof(int), cudaMemcpyDeviceToHost, s->sD));
  RYK_CUDA(cudaMemcpyAsync(s->h_out[r], s->d_out_fixed[b], sizeof(double) * (size_t)cap, cudaMemcpyDeviceToHost, s->sD));
  RYK_CUDA(cudaEventRecord(s->ev_dec[r], s->sD));
  if (ticket) *ticket = k;
  return 0;
}

// Wait for the chunk `ticket` (tickets must be collected in order) and copy its samples out.
int ryk_session_collect(ryk_engine* h, int id, long long ticket, double* out, int out_capacity, int* n_out) {
  Engine* e = &h->impl;
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(ticket == s->collected && ticket < s->step, "tickets are collected in submission order");
  const int r = (int)(ticket % kRing);
  RYK_CUDA(cudaEventSynchronize(s->ev_dec[r]));
  int produced = *s->h_n[r];
  RYK_CHECK(produced <= out_capacity, "output buffer too small for the produced blocks");
  memcpy(out, s->h_out[r], sizeof(double) * produced);
  *n_out = produced;
  s->collected++;
  return 0;
}

// Non-blocking completion query (cudaEventQuery of the step's last decode-stream event): *done = 1 when ryk_session_collect would
// not wait.  This is what queue_output_wave.get_nowait() needs (run.py:176-182).
int ryk_session_poll(ryk_engine* h, int id, long long ticket, int* done) {
  Engine* e = &h->impl;
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr && done != nullptr, "no such session");
  RYK_CHECK(ticket >= 0 && ticket < s->step && ticket + kRing > s->step, "ticket is not among the last 8 steps");
  cudaError_t q = cudaEventQuery(s->ev_dec[ticket % kRing]);
  if (q != cudaSuccess && q != cudaErrorNotReady) RYK_CUDA(q);
  *done = q == cudaSuccess ? 1 : 0;
  return 0;
}

int ryk_session_push(ryk_engine* h, int id, const float* wave, int n, double* out, int out_capacity, int* n_out) {
  long long ticket = 0;
  if (ryk_session_submit(h, id, wave, n, &ticket)) return -1;
  return ryk_session_collect(h, id, ticket, out, out_capacity, n_out);
}

// Device-resident step, asynchronous: returns as soon as the work is queued (output valid after ryk_engine_synchronize
// or any later stream-ordered work of this session).
int ryk_session_push_device(ryk_engine* h, int id, const float* wave_dev, int n, double* out_dev, int out_capacity, int* n_out_dev) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr, "no such session");
  RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
  RYK_CHECK(s->group == nullptr, "session belongs to a group: use ryk_group_push_device");
  const int r = (int)(s->step % kRing), b = (int)(s->step & 1);
  const int cap = s->max_blocks * s->cfg.vocoder_buffer_size;
  RYK_CHECK(out_capacity >= cap, "out_capacity must hold (frames * hop / block + 4) synthesizer blocks");
  if (session_enqueue(e, s, wave_dev)) return -1;
  RYK_CUDA(cudaMemcpyAsync(out_dev, s->d_out_fixed[b], sizeof(double) * (size_t)cap, cudaMemcpyDeviceToDevice, s->sD));
  RYK_CUDA(cudaMemcpyAsync(n_out_dev, s->d_n_fixed[b], sizeof(int), cudaMemcpyDeviceToDevice, s->sD));
  RYK_CUDA(cudaEventRecord(s->ev_dec[r], s->sD));
  s->collected = s->step;         // device-resident steps are not collected through the host API
  return 0;
}

// ---- groups: several sessions of one GPU sharing one batched stage-2 forward per step (BASELINE config 5) ----
static Group* get_group(Engine* e, int id) { return (id >= 0 && id < (int)e->groups.size()) ? e->groups[id] : nullptr; }

int ryk_group_create(ryk_engine* h, const int* session_ids, int n_sessions, int* group_id) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  RYK_CHECK(session_ids && group_id && n_sessions >= 1 && n_sessions <= 64, "a group holds 1..64 sessions");
  Group* G = new Group();
  G->owner = 1000000 + (int)e->groups.size();
  memset(G->ev_fwd, 0, sizeof(G->ev_fwd));
  for (int i = 0; i < n_sessions; ++i) {
    Session* s = get_session(e, session_ids[i]);
    if (!s || s->group || s->step != 0 || (i > 0 && s->Tw != G->members[0]->Tw)) {
      for (Session* m : G->members) m->group = nullptr;
      delete G;
      RYK_CHECK(false, "group members must be distinct fresh sessions (no chunk pushed yet) with the same window length");
    }
    s->group = G; s->slot = i;
    G->members.push_back(s);
  }
  G->Tp = G->members[0]->Tw + (128 - G->members[0]->Tw % 128);
  if (unet_get_plan(e, e->stage2, n_sessions, G->Tp, 512, e->precision, &G->p2, G->owner)) { group_free(G); return -1; }
  { int lo = 0, hi = 0; RYK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi)); RYK_CUDA(cudaStreamCreateWithPriority(&G->sG, cudaStreamNonBlocking, lo)); }
  for (int i = 0; i < kRing; ++i) RYK_CUDA(cudaEventCreateWithFlags(&G->ev_fwd[i], cudaEventDisableTiming));
  RYK_CUDA(cudaDeviceSynchronize());
  e->groups.push_back(G);
  *group_id = (int)e->groups.size() - 1;
  return 0;
}

int ryk_group_destroy(ryk_engine* h, int group_id) {
  Engine* e = &h->impl;
  Group* G = get_group(e, group_id);
  RYK_CHECK(G != nullptr, "no such group");
  RYK_CUDA(cudaDeviceSynchronize());
  const int owner = G->owner;
  group_free(G);                     // the member sessions survive (ungrouped) and are destroyed separately
  unet_release_owner(e->stage2, owner);
  e->groups[group_id] = nullptr;
  return 0;
}

int ryk_group_size(ryk_engine* h, int group_id) {
  Group* G = get_group(&h->impl, group_id);
  return G ? (int)G->members.size() : -1;
}

// Queue one chunk per member (host samples; waves[i] belongs to member i in creation order).
int ryk_group_submit(ryk_engine* h, int group_id, const float* const* waves, int n, long long* ticket) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Group* G = get_group(e, group_id);
  RYK_CHECK(G != nullptr, "no such group");
  RYK_CHECK(G->step - G->collected < kRing - 2, "too many chunks in flight: collect before submitting more");
  const long long k = G->step;
  const int r = (int)(k % kRing), b = (int)(k & 1);
  std::vector<const float*> d_chunks(G->members.size());
  for (size_t i = 0; i < G->members.size(); ++i) {
    Session* s = G->members[i];
    RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
    memcpy(s->h_in[r], waves[i], sizeof(float) * n);
    RYK_CUDA(cudaMemcpyAsync(s->d_chunk[r], s->h_in[r], sizeof(float) * n, cudaMemcpyHostToDevice, s->sE));
    d_chunks[i] = s->d_chunk[r];
  }
  if (group_enqueue(e, G, d_chunks.data())) return -1;
  for (Session* s : G->members) {
    const int cap = s->max_blocks * s->cfg.vocoder_buffer_size;
    RYK_CUDA(cudaMemcpyAsync(s->h_n[r], s->d_n_fixed[b], sizeof(int), cudaMemcpyDeviceToHost, s->sD));
    RYK_CUDA(cudaMemcpyAsync(s->h_out[r], s->d_out_fixed[b], sizeof(double) * (size_t)cap, cudaMemcpyDeviceToHost, s->sD));
    RYK_CUDA(cudaEventRecord(s->ev_dec[r], s->sD));
  }
  if (ticket) *ticket = k;
  return 0;
}

// Wait for step `ticket` of every member; outs[i] receives member i's samples, n_outs[i] their count.
int ryk_group_collect(ryk_engine* h, int group_id, long long ticket, double* const* outs, int out_capacity, int* n_outs) {
  Engine* e = &h->impl;
  Group* G = get_group(e, group_id);
  RYK_CHECK(G != nullptr, "no such group");
  RYK_CHECK(ticket == G->collected && ticket < G->step, "tickets are collected in submission order");
  const int r = (int)(ticket % kRing);
  for (size_t i = 0; i < G->members.size(); ++i) {
    Session* s = G->members[i];
    RYK_CUDA(cudaEventSynchronize(s->ev_dec[r]));
    const int produced = *s->h_n[r];
    RYK_CHECK(produced <= out_capacity, "output buffer too small for the produced blocks");
    memcpy(outs[i], s->h_out[r], sizeof(double) * produced);
    n_outs[i] = produced;
    s->collected++;
  }
  G->collected++;
  return 0;
}

// Device-resident group step, asynchronous (see ryk_session_push_device).
int ryk_group_push_device(ryk_engine* h, int group_id, const float* const* waves_dev, int n, double* const* outs_dev, int out_capacity,
                          int* const* n_outs_dev) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Group* G = get_group(e, group_id);
  RYK_CHECK(G != nullptr, "no such group");
  const int r = (int)(G->step % kRing), b = (int)(G->step & 1);
  for (Session* s : G->members) {
    RYK_CHECK(n == s->n_wave, "chunk length must be round(fs * buffer_time)");
    RYK_CHECK(out_capacity >= s->max_blocks * s->cfg.vocoder_buffer_size, "out_capacity must hold (frames * hop / block + 4) synthesizer blocks");
  }
  if (group_enqueue(e, G, waves_dev)) return -1;
  for (size_t i = 0; i < G->members.size(); ++i) {
    Session* s = G->members[i];
    const int cap = s->max_blocks * s->cfg.vocoder_buffer_size;
    RYK_CUDA(cudaMemcpyAsync(outs_dev[i], s->d_out_fixed[b], sizeof(double) * (size_t)cap, cudaMemcpyDeviceToDevice, s->sD));
    RYK_CUDA(cudaMemcpyAsync(n_outs_dev[i], s->d_n_fixed[b], sizeof(int), cudaMemcpyDeviceToDevice, s->sD));
    RYK_CUDA(cudaEventRecord(s->ev_dec[r], s->sD));
    s->collected = s->step;
  }
  G->collected = G->step;
  return 0;
}

// Diagnostics (RYK_STAGE_TIMES=1 at session creation): device timeline of the last min(steps, 8) steps.  start/end[i*5 + a] =
// ms since the oldest listed step began, for stage a in {gate+slides, WORLD analysis, stage 1 (+mc2sp), stage 2
// (prologue..epilogue), synthesis}; returns the number of steps listed (oldest first).
int ryk_session_stage_times(ryk_engine* h, int id, float* start, float* end) {
  Engine* e = &h->impl;
  Session* s = get_session(e, id);
  RYK_CHECK(s != nullptr && s->stage_times && s->step >= 1, "stage timing is not enabled for this session (RYK_STAGE_TIMES=1) or no step ran");
  RYK_CUDA(cudaDeviceSynchronize());
  const int n = s->step < kRing ? (int)s->step : kRing;
  const int r0 = (int)((s->step - n) % kRing);
  for (int i = 0; i < n; ++i) {
    const int r = (int)((s->step - n + i) % kRing);
    for (int a = 0; a < 5; ++a) {
      RYK_CUDA(cudaEventElapsedTime(&start[i * 5 + a], s->tev[0][0][r0], s->tev[a][0][r]));
      RYK_CUDA(cudaEventElapsedTime(&end[i * 5 + a], s->tev[0][0][r0], s->tev[a][1][r]));
    }
  }
  return n;
}


// ---- output re-blocker + silence gate -----------------------------------------------------------------
static Reblock* get_reblock(Engine* e, int id) { return (id >= 0 && id < (int)e->reblocks.size()) ? e->reblocks[id] : nullptr; }

int ryk_reblock_create(ryk_engine* h, int out_audio_chunk, int max_in, int n_fft, int hop, double threshold_db, int* reblock_id) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  RYK_CHECK(out_audio_chunk > n_fft / 2 && max_in > 0, "out_audio_chunk must exceed n_fft / 2 (reflect-centred STFT)");
  RYK_CHECK(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0 && hop > 0, "unsupported STFT geometry");
  Reblock* R = new Reblock();
  for (int i = 0; i < kRing; ++i) R->ev[i] = nullptr;
  R->chunk = out_audio_chunk; R->max_in = max_in; R->n_fft = n_fft; R->hop = hop; R->threshold_db = threshold_db;
  R->cap = 2 * out_audio_chunk + 2 * max_in;
  auto A = [&](void** p, size_t bytes) -> int { RYK_CUDA(cudaMalloc(p, bytes)); RYK_CUDA(cudaMemset(*p, 0, bytes)); R->allocs.push_back(*p); return 0; };
  auto H = [&](void** p, size_t bytes) -> int { RYK_CUDA(cudaMallocHost(p, bytes)); memset(*p, 0, bytes); R->pinned.push_back(*p); return 0; };
  int rc = 0;
  rc |= A((void**)&R->d_state, sizeof(ReblockState));
  for (int i = 0; i < 2; ++i) rc |= A((void**)&R->d_frag[i], sizeof(double) * R->cap);
  rc |= A((void**)&R->d_scratch, sizeof(double) * output_gate_scratch_doubles(out_audio_chunk, n_fft, hop));
  rc |= A((void**)&R->d_stage_in, sizeof(double) * max_in);
  rc |= A((void**)&R->d_stage_n, sizeof(int));
  for (int i = 0; i < kRing && !rc; ++i) {
    rc |= A((void**)&R->d_chunk[i], sizeof(double) * out_audio_chunk);
    rc |= A((void**)&R->d_nvalid[i], sizeof(int));
    rc |= A((void**)&R->d_status[i], sizeof(int));
    rc |= A((void**)&R->d_power[i], sizeof(double));
    rc |= H((void**)&R->h_chunk[i], sizeof(double) * out_audio_chunk);
    rc |= H((void**)&R->h_status[i], sizeof(int));
    rc |= H((void**)&R->h_power[i], sizeof(double));
    rc |= H((void**)&R->h_overflow[i], sizeof(int));
    if (!rc && cudaEventCreateWithFlags(&R->ev[i], cudaEventDisableTiming) != cudaSuccess) rc = -1;
  }
  if (rc) { reblock_free(R); return -1; }
  RYK_CUDA(cudaDeviceSynchronize());             // the zero-fills ran on the legacy default stream; pushes use other (non-blocking) streams
  e->reblocks.push_back(R);
  *reblock_id = (int)e->reblocks.size() - 1;
  return 0;
}

int ryk_reblock_destroy(ryk_engine* h, int id) {
  Engine* e = &h->impl;
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr, "no such re-blocker");
  RYK_CUDA(cudaDeviceSynchronize());
  reblock_free(R);
  e->reblocks[id] = nullptr;
  return 0;
}

// Append *n_dev samples (device memory) and emit at most one chunk + its gate decision into ring slot ticket % 8.
// session_id >= 0: the work is queued on that session's decode stream right behind its latest step; with wave_dev == NULL the
// step's own output (blocks + count) is consumed in place.  session_id < 0: engine stream, wave_dev / n_dev required.
int ryk_reblock_push_device(ryk_engine* h, int id, int session_id, const double* wave_dev, const int* n_dev, long long* ticket) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr, "no such re-blocker");
  cudaStream_t st = e->stream;
  if (session_id >= 0) {
    Session* s = get_session(e, session_id);
    RYK_CHECK(s != nullptr, "no such session");
    RYK_CHECK(s->step > 0, "the session has not processed a chunk yet");
    st = s->sD;
    if (!wave_dev) {
      const int b = (int)((s->step - 1) & 1);
      RYK_CHECK(s->max_blocks * s->cfg.vocoder_buffer_size <= R->max_in, "re-blocker max_in is smaller than the session's block capacity");
      wave_dev = s->d_out_fixed[b]; n_dev = s->d_n_fixed[b];
    }
  }
  RYK_CHECK(wave_dev != nullptr && n_dev != nullptr, "wave_dev / n_dev are required without an attached session");
  const long long k = R->pushed;
  const int r = (int)(k % kRing);
  k_reblock<<<1, 1024, 0, st>>>(R->d_state, R->d_frag[0], R->d_frag[1], R->cap, wave_dev, n_dev, R->max_in, R->chunk, R->d_chunk[r], R->d_nvalid[r]);
  e->launches += 1;
  RYK_CUDA(cudaGetLastError());
  if (output_gate_async(e, R->d_chunk[r], R->d_nvalid[r], R->chunk, R->n_fft, R->hop, R->threshold_db, R->d_scratch, R->d_power[r], R->d_status[r], st)) return -1;
  RYK_CUDA(cudaMemcpyAsync(R->h_status[r], R->d_status[r], sizeof(int), cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaMemcpyAsync(R->h_power[r], R->d_power[r], sizeof(double), cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaMemcpyAsync(R->h_chunk[r], R->d_chunk[r], sizeof(double) * R->chunk, cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaMemcpyAsync(R->h_overflow[r], &R->d_state->overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
  RYK_CUDA(cudaEventRecord(R->ev[r], st));
  R->pushed++;
  if (ticket) *ticket = k;
  return 0;
}

// Wait for push `ticket` (one of the last 8): *status 0 = no chunk this step, 1 = chunk written to chunk_out, 2 = chunk was
// silent (the reference forwards None); *power_db = mean STFT power of the chunk (0 when status is 0).
int ryk_reblock_collect(ryk_engine* h, int id, long long ticket, double* chunk_out, int* status, double* power_db) {
  Engine* e = &h->impl;
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr, "no such re-blocker");
  RYK_CHECK(ticket >= 0 && ticket < R->pushed && ticket + kRing > R->pushed, "ticket is not among the last 8 pushes");
  const int r = (int)(ticket % kRing);
  RYK_CUDA(cudaEventSynchronize(R->ev[r]));
  // the reference's wave_fragment grows without bound when a step yields more than one out_audio_chunk (decode_worker.py:47-52);
  // the device fragment is bounded, so that configuration is an error here instead of silently dropped samples
  RYK_CHECK(*R->h_overflow[r] == 0, "re-blocker fragment overflow: a step produced more samples than out_audio_chunk can drain");
  const int stt = *R->h_status[r];
  if (status) *status = stt;
  if (power_db) *power_db = *R->h_power[r];
  if (chunk_out && stt != 0) memcpy(chunk_out, R->h_chunk[r], sizeof(double) * R->chunk);
  return 0;
}

// Non-blocking: *done = 1 when ryk_reblock_collect(ticket) would not wait.
int ryk_reblock_poll(ryk_engine* h, int id, long long ticket, int* done) {
  Engine* e = &h->impl;
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr && done != nullptr, "no such re-blocker");
  RYK_CHECK(ticket >= 0 && ticket < R->pushed && ticket + kRing > R->pushed, "ticket is not among the last 8 pushes");
  cudaError_t q = cudaEventQuery(R->ev[ticket % kRing]);
  if (q != cudaSuccess && q != cudaErrorNotReady) RYK_CUDA(q);
  *done = q == cudaSuccess ? 1 : 0;
  return 0;
}

// Device pointers of ring slot ticket % 8 (valid until 8 further pushes; ordered after the push on its stream).
int ryk_reblock_result_device(ryk_engine* h, int id, long long ticket, const double** chunk_dev, const int** status_dev, const double** power_dev) {
  Engine* e = &h->impl;
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr, "no such re-blocker");
  RYK_CHECK(ticket >= 0 && ticket < R->pushed && ticket + kRing > R->pushed, "ticket is not among the last 8 pushes");
  const int r = (int)(ticket % kRing);
  if (chunk_dev) *chunk_dev = R->d_chunk[r];
  if (status_dev) *status_dev = R->d_status[r];
  if (power_dev) *power_dev = R->d_power[r];
  return 0;
}

// Host buffers: H2D + push + collect.
int ryk_reblock_push(ryk_engine* h, int id, const double* wave, int n, double* chunk_out, int* status, double* power_db) {
  Engine* e = &h->impl;
  RYK_CUDA(cudaSetDevice(e->device));
  Reblock* R = get_reblock(e, id);
  RYK_CHECK(R != nullptr, "no such re-blocker");
  RYK_CHECK(n >= 0 && n <= R->max_in, "more samples than the re-blocker's max_in");
  if (n > 0) RYK_CUDA(cudaMemcpyAsync(R->d_stage_in, wave, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
  RYK_CUDA(cudaMemcpyAsync(R->d_stage_n, &n, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  RYK_CUDA(cudaStreamSynchronize(e->stream));           // n lives on the caller's stack
  long long ticket = 0;
  if (ryk_reblock_push_device(h, id, -1, R->d_stage_in, R->d_stage_n, &ticket)) return -1;
  return ryk_reblock_collect(h, id, ticket, chunk_out, status, power_db);
}

}  // extern "C"
