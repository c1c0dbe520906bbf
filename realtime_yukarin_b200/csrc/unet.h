// unet.h -- U-Net containers (weights + per-shape plans).
#pragma once
#include <map>
#include <tuple>
#include <vector>

#include "conv.h"

namespace ryk {

struct Engine;

struct UNetLayerW {
  int transposed = 0, cin = 0, cout = 0, k = 3, s = 1, p = 1, act = 0;
  float* d_w_direct = nullptr;
  __half* d_w_tc = nullptr;
  __half* d_w_frag = nullptr;            // fused stage-1 kernel (1-D nets, k4 layers)
  float* d_scale = nullptr;
  float* d_shift = nullptr;
  float h_scale0 = 1.f, h_shift0 = 0.f;   // first channel's scale/shift (used by the Cout = 1 kernel)
  bool loaded = false;
};

struct UNetPlan {
  int B = 1, H = 1, W = 0, precision = 0;
  std::vector<ConvLayer> layers;
  std::vector<void*> buffers;
  void* d_in = nullptr;     // fp32 NHWC input  [B][H][W][in_ch]
  void* d_out = nullptr;    // fp32 NHWC output [B][H][W][out_ch]
  bool fused = false;       // 1-D FP16 plan that s1_fused.cu can run as one launch
};

struct UNet {
  int ndim = 2, in_ch = 1, out_ch = 1, base = 64;
  std::vector<UNetLayerW> layers;                                  // 0..7 encoder, 8..15 decoder
  std::map<std::tuple<int, int, int, int, int>, UNetPlan*> plans;  // (B, H, W, precision, owner)
};

UNet* unet_create(int ndim, int in_ch, int out_ch, int base);
void unet_destroy(UNet* n);
int unet_set_layer(Engine* e, UNet* n, int idx, const float* W, const float* scale, const float* shift);
// owner: 0 = the engine's shared plans (per-op host API, serialised on the engine stream); every session / group passes its own
// id so that concurrently running streams never share activation buffers.
int unet_get_plan(Engine* e, UNet* n, int B, int H, int W, int precision, UNetPlan** out, int owner = 0);
void unet_release_owner(UNet* n, int owner);
int unet_forward(Engine* e, UNetPlan* p, cudaStream_t st, int first_layer = 0, int last_layer = 15);
// s1_fused.cu
bool s1_fused_eligible(const UNet* n, const UNetPlan* p);
int s1_fused_run(Engine* e, const UNetPlan* p, cudaStream_t st);

}  // namespace ryk
