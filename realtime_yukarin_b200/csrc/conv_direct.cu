// conv_direct.cu -- FP32 CUDA-core implicit-GEMM convolution / transposed convolution over NHWC
// activations with fused folded-BN scale/shift + LeakyReLU/ReLU epilogue and skip-concat by pointer.
//
// Role on the hot path (components H and I):
//   * every layer of the stage-1 1-D U-Net (0.55 GFLOP, latency bound, M <= 192 pixels per layer);
//   * the first (Cin = 1) and last (Cout = 1) 3x3 layers of the stage-2 2-D U-Net, whose GEMM shape
//     has nothing for a tensor core to chew on;
//   * all layers when the engine runs in FP32 "bisect" precision (the numerics reference for the
//     tcgen05 path in conv_tc.cu).
// Tiling: 64 output pixels x 64 output channels per CTA, 16-channel K steps through shared memory,
// 4x4 register micro-tiles (classic SGEMM shape).  Transposed convs are evaluated per output-parity
// class so that every pixel of a tile shares the same set of contributing taps.
#include "conv.h"

namespace ryk {

template <typename T> __device__ inline float ld_act(const T* p);
template <> __device__ inline float ld_act<float>(const float* p) { return *p; }
template <> __device__ inline float ld_act<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ inline void st_act(T* p, float v);
template <> __device__ inline void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ inline void st_act<__half>(__half* p, float v) { *p = __float2half_rn(v); }

struct DirectParams {
  int transposed, B, Hin, Win, Hout, Wout, C0, C1, Cout, KH, KW, SH, SW, PH, PW, act;
  int classH, classW;            // transposed: SH, SW ; else 1, 1
  int Hc, Wc;                    // class-local output grid
  const void* in0; const void* in1; void* out;
  const float* w; const float* scale; const float* shift;
};

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_conv_direct(DirectParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int pix_b[BM], pix_y[BM], pix_x[BM];
  const int Cin = p.C0 + p.C1;
  const int cls = blockIdx.z;
  const int py = cls / p.classW, px = cls % p.classW;
  const int npix = p.B * p.Hc * p.Wc;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  if (tid < BM) {
    int q = m0 + tid;
    if (q < npix) {
      int b = q / (p.Hc * p.Wc);
      int r = q - b * p.Hc * p.Wc;
      int my = r / p.Wc, mx = r - my * p.Wc;
      pix_b[tid] = b; pix_y[tid] = my * p.classH + py; pix_x[tid] = mx * p.classW + px;
    } else {
      pix_b[tid] = -1; pix_y[tid] = 0; pix_x[tid] = 0;
    }
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_pix = tid >> 2, a_c4 = (tid & 3) * 4;     // A loader: pixel, first of 4 channels
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;      // B loader
  const TIn* in0 = (const TIn*)p.in0;
  const TIn* in1 = (const TIn*)p.in1;

  for (int ky = 0; ky < p.KH; ++ky) {
    if (p.transposed && ((py + p.PH - ky) % p.SH + p.SH) % p.SH != 0) continue;
    for (int kx = 0; kx < p.KW; ++kx) {
      if (p.transposed && ((px + p.PW - kx) % p.SW + p.SW) % p.SW != 0) continue;
      // input coordinate of this thread's A pixel for this tap
      long long abase = -1;
      {
        int b = pix_b[a_pix];
        if (b >= 0) {
          int iy, ix;
          if (!p.transposed) { iy = pix_y[a_pix] * p.SH - p.PH + ky; ix = pix_x[a_pix] * p.SW - p.PW + kx; }
          else { iy = (pix_y[a_pix] + p.PH - ky) / p.SH; ix = (pix_x[a_pix] + p.PW - kx) / p.SW; }
          if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) abase = ((long long)(b * p.Hin + iy) * p.Win + ix);
        }
      }
      const float* wtap = p.w + (size_t)(ky * p.KW + kx) * Cin * p.Cout;
      for (int c0 = 0; c0 < Cin; c0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + a_c4 + j;
          float v = 0.f;
          if (abase >= 0 && c < Cin) {
            if (c < p.C0) v = ld_act<TIn>(in0 + abase * p.C0 + c);
            else v = ld_act<TIn>(in1 + abase * p.C1 + (c - p.C0));
          }
          As[a_c4 + j][a_pix] = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = c0 + b_k, n = n0 + b_n4 + j;
          Bs[b_k][b_n4 + j] = (c < Cin && n < p.Cout) ? wtap[(size_t)c * p.Cout + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
  TOut* out = (TOut*)p.out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pl = ty * 4 + i;
    int b = pix_b[pl];
    if (b < 0) continue;
    size_t obase = ((size_t)(b * p.Hout + pix_y[pl]) * p.Wout + pix_x[pl]) * p.Cout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j] * p.scale[n] + p.shift[n];
      if (p.act == ACT_LEAKY) v = v > 0.f ? v : 0.2f * v;
      else if (p.act == ACT_RELU) v = v > 0.f ? v : 0.f;
      st_act<TOut>(out + obase + n, v);
    }
  }
}

static int conv_direct_generic(const ConvLayer& L, cudaStream_t st) {
  DirectParams p;
  p.transposed = L.transposed; p.B = L.B; p.Hin = L.Hin; p.Win = L.Win; p.Hout = L.Hout; p.Wout = L.Wout;
  p.C0 = L.C0; p.C1 = L.C1; p.Cout = L.Cout; p.KH = L.KH; p.KW = L.KW; p.SH = L.SH; p.SW = L.SW; p.PH = L.PH; p.PW = L.PW;
  p.act = L.act;
  p.classH = L.transposed ? L.SH : 1; p.classW = L.transposed ? L.SW : 1;
  RYK_CHECK(L.Hout % p.classH == 0 && L.Wout % p.classW == 0, "transposed conv output must be a multiple of the stride");
  p.Hc = L.Hout / p.classH; p.Wc = L.Wout / p.classW;
  p.in0 = L.in0; p.in1 = L.in1; p.out = L.out; p.w = L.w_direct; p.scale = L.scale; p.shift = L.shift;
  RYK_CHECK(L.w_direct && L.scale && L.shift && L.in0 && L.out, "direct conv layer is missing a device pointer");
  int npix = L.B * p.Hc * p.Wc;
  dim3 grid((npix + 63) / 64, (L.Cout + 63) / 64, p.classH * p.classW);
  if (L.in_dtype == DT_F32 && L.out_dtype == DT_F32) k_conv_direct<float, float><<<grid, 256, 0, st>>>(p);
  else if (L.in_dtype == DT_F32 && L.out_dtype == DT_F16) k_conv_direct<float, __half><<<grid, 256, 0, st>>>(p);
  else if (L.in_dtype == DT_F16 && L.out_dtype == DT_F16) k_conv_direct<__half, __half><<<grid, 256, 0, st>>>(p);
  else k_conv_direct<__half, float><<<grid, 256, 0, st>>>(p);
  RYK_CUDA(cudaGetLastError());
  return 0;
}


// ---- stage-2 edge layers -----------------------------------------------------------------------
// The first (Cin = 1) and last (Cout = 1) 3x3 layers of the 2-D U-Net are pure bandwidth problems
// (25 MB written / 50 MB read at 384x512); the generic 64x64 tile above would waste 16-64x of its math on
// them, so they get dedicated kernels: coalesced 16-byte accesses, weights in shared memory.

// Cin = 1, fp32 input -> Cout (multiple of 8, <= 64) channels. One thread = one pixel x 8 output channels.
template <typename TOut>
__global__ void __launch_bounds__(256) k_conv3x3_cin1(const float* __restrict__ in, const float* __restrict__ w /*[9][1][Cout]*/,
                                                     const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                     int B, int H, int W, int Cout, TOut* __restrict__ out) {
  __shared__ float ws[9 * 64], sc[64], sh[64];
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) { sc[i] = scale[i]; sh[i] = shift[i]; }
  __syncthreads();
  const int groups = Cout / 8;
  const long long total = (long long)B * H * W * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int g = idx % groups; long long pix = idx / groups;
    int x = pix % W; long long r = pix / W; int y = r % H; int b = r / H;
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      int iy = y + t / 3 - 1, ix = x + t % 3 - 1;
      v[t] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(in + ((long long)b * H + iy) * W + ix) : 0.f;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = g * 8 + j;
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(v[t], ws[t * Cout + n], a);
      a = a * sc[n] + sh[n];
      if (act == ACT_LEAKY) a = a > 0.f ? a : 0.2f * a; else if (act == ACT_RELU) a = fmaxf(a, 0.f);
      o[j] = a;
    }
    TOut* dst = out + pix * Cout + g * 8;
    if constexpr (sizeof(TOut) == 2) {
      __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]), h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
      *reinterpret_cast<uint4*>(dst) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                  *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// Cout = 1 from two fp16 sources of 64 channels each (skip concat) -> fp32. One warp = one pixel at a time:
// lanes 0-15 read 4 channels each of source 0, lanes 16-31 of source 1 (8-byte loads, 256 B per tap), warp-shuffle sum.
__global__ void __launch_bounds__(256) k_conv3x3_cout1_h(const __half* __restrict__ in0, const __half* __restrict__ in1,
                                                        const float* __restrict__ w /*[9][128][1]*/, float scale, float shift, int act,
                                                        int B, int H, int W, float* __restrict__ out) {
  __shared__ float ws[9 * 128];
  for (int i = threadIdx.x; i < 9 * 128; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // block tile: 4 rows x 16 columns of pixels, 8 warps x 8 pixels
  const int tiles_x = (W + 15) / 16, tiles_y = (H + 3) / 4;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; const int b = t / tiles_y;
  const __half* src = lane < 16 ? in0 : in1;
  const int c = (lane & 15) * 4;
  const float* wl = ws + (lane < 16 ? 0 : 64) + c;
  for (int q = 0; q < 8; ++q) {
    int pl = warp * 8 + q;
    int y = ty * 4 + pl / 16, x = tx * 16 + pl % 16;
    if (y >= H || x >= W) continue;
    float a = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      uint2 raw = __ldg(reinterpret_cast<const uint2*>(src + (((long long)b * H + iy) * W + ix) * 64 + c));
      __half2 h01 = *reinterpret_cast<__half2*>(&raw.x), h23 = *reinterpret_cast<__half2*>(&raw.y);
      float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const float* wt = wl + tap * 128;
      a = fmaf(f01.x, wt[0], a); a = fmaf(f01.y, wt[1], a); a = fmaf(f23.x, wt[2], a); a = fmaf(f23.y, wt[3], a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a = a * scale + shift;
      if (act == ACT_LEAKY) a = a > 0.f ? a : 0.2f * a; else if (act == ACT_RELU) a = fmaxf(a, 0.f);
      out[((long long)b * H + y) * W + x] = a;
    }
  }
}

__global__ void k_read2(const float* a, const float* b, float* out) { out[0] = a[0]; out[1] = b[0]; }

int conv_direct_run(const ConvLayer& L, cudaStream_t st) {
  // dedicated kernels for the stage-2 edge layers
  if (!L.transposed && L.KH == 3 && L.KW == 3 && L.SH == 1 && L.SW == 1 && L.PH == 1 && L.PW == 1) {
    if (L.C0 == 1 && L.C1 == 0 && L.in_dtype == DT_F32 && L.Cout % 8 == 0 && L.Cout <= 64) {
      long long total = (long long)L.B * L.Hin * L.Win * (L.Cout / 8);
      int blocks = (int)((total + 255) / 256);
      if (L.out_dtype == DT_F16)
        k_conv3x3_cin1<__half><<<blocks, 256, 0, st>>>((const float*)L.in0, L.w_direct, L.scale, L.shift, L.act, L.B, L.Hin, L.Win, L.Cout, (__half*)L.out);
      else
        k_conv3x3_cin1<float><<<blocks, 256, 0, st>>>((const float*)L.in0, L.w_direct, L.scale, L.shift, L.act, L.B, L.Hin, L.Win, L.Cout, (float*)L.out);
      RYK_CUDA(cudaGetLastError());
      return 0;
    }
    if (L.Cout == 1 && L.C0 == 64 && L.C1 == 64 && L.in_dtype == DT_F16 && L.out_dtype == DT_F32 && L.host_scale_valid) {
      int blocks = L.B * ((L.Hin + 3) / 4) * ((L.Win + 15) / 16);
      k_conv3x3_cout1_h<<<blocks, 256, 0, st>>>((const __half*)L.in0, (const __half*)L.in1, L.w_direct, L.host_scale, L.host_shift, L.act,
                                                L.B, L.Hin, L.Win, (float*)L.out);
      RYK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  return conv_direct_generic(L, st);
}

// ---- weight repacking ------------------------------------------------------------------------
__global__ void k_pack_direct(const float* __restrict__ w, int transposed, int Cin, int Cout, int KH, int KW, float* __restrict__ out) {
  size_t total = (size_t)KH * KW * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n = i % Cout; size_t r = i / Cout;
    int c = r % Cin; r /= Cin;
    int kx = r % KW; int ky = r / KW;
    size_t src = transposed ? (((size_t)c * Cout + n) * KH + ky) * KW + kx : (((size_t)n * Cin + c) * KH + ky) * KW + kx;
    out[i] = w[src];
  }
}

int pack_weights_direct(const float* d_w, int transposed, int Cin, int Cout, int KH, int KW, float* d_out, cudaStream_t st) {
  k_pack_direct<<<256, 256, 0, st>>>(d_w, transposed, Cin, Cout, KH, KW, d_out);
  RYK_CUDA(cudaGetLastError());
  return 0;
}

// tensor-core packing: fp16, K-major rows.
//   conv  : out[n][(ky*KW+kx)*Cin + c]                                    = W[n][c][ky][kx]
//   deconv: out[cls][n][(dy*(KW/SW)+dx)*Cin + c], cls = py*SW+px          = W[c][n][ky][kx]
//           with k = 3 - 2d - parity along a k4 s2 p1 dimension and k = 0 along a k1 s1 p0 dimension (1-D nets)
__global__ void k_pack_tc(const float* __restrict__ w, int transposed, int Cin, int Cout, int KH, int KW, int SH, int SW,
                          __half* __restrict__ out) {
  size_t total = (size_t)KH * KW * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (!transposed) {
      size_t K = (size_t)KH * KW * Cin;
      int n = i / K; size_t k = i % K;
      int tap = k / Cin, c = k % Cin;
      int ky = tap / KW, kx = tap % KW;
      out[i] = __float2half_rn(w[(((size_t)n * Cin + c) * KH + ky) * KW + kx]);
    } else {
      int th = KH / SH, tw = KW / SW;
      size_t K = (size_t)th * tw * Cin;
      size_t per_cls = K * Cout;
      int cls = i / per_cls; size_t r = i % per_cls;
      int n = r / K; size_t k = r % K;
      int tap = k / Cin, c = k % Cin;
      int dy = tap / tw, dx = tap % tw, py = cls / SW, px = cls % SW;
      int ky = SH == 2 ? 3 - 2 * dy - py : 0, kx = SW == 2 ? 3 - 2 * dx - px : 0;
      out[i] = __float2half_rn(w[(((size_t)c * Cout + n) * KH + ky) * KW + kx]);
    }
  }
}

int pack_weights_tc(const float* d_w, int transposed, int Cin, int Cout, int KH, int KW, int SH, int SW, __half* d_out, cudaStream_t st) {
  k_pack_tc<<<512, 256, 0, st>>>(d_w, transposed, Cin, Cout, KH, KW, SH, SW, d_out);
  RYK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ryk
