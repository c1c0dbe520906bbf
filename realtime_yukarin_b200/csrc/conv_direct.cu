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

int conv_direct_run(const ConvLayer& L, cudaStream_t st) {
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
//   conv  : out[n][(ky*KW+kx)*Cin + c]                       = W[n][c][ky][kx]
//   deconv: out[cls][n][(dy*2+dx)*Cin + c], cls = py*2+px    = W[c][n][3-2dy-py][3-2dx-px]   (k4 s2 p1 only)
__global__ void k_pack_tc(const float* __restrict__ w, int transposed, int Cin, int Cout, int KH, int KW, __half* __restrict__ out) {
  size_t total = (size_t)KH * KW * Cin * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (!transposed) {
      size_t K = (size_t)KH * KW * Cin;
      int n = i / K; size_t k = i % K;
      int tap = k / Cin, c = k % Cin;
      int ky = tap / KW, kx = tap % KW;
      out[i] = __float2half_rn(w[(((size_t)n * Cin + c) * KH + ky) * KW + kx]);
    } else {
      size_t K = (size_t)4 * Cin;
      size_t per_cls = K * Cout;
      int cls = i / per_cls; size_t r = i % per_cls;
      int n = r / K; size_t k = r % K;
      int tap = k / Cin, c = k % Cin;
      int dy = tap >> 1, dx = tap & 1, py = cls >> 1, px = cls & 1;
      int ky = 3 - 2 * dy - py, kx = 3 - 2 * dx - px;
      out[i] = __float2half_rn(w[(((size_t)c * Cout + n) * KH + ky) * KW + kx]);
    }
  }
}

int pack_weights_tc(const float* d_w, int transposed, int Cin, int Cout, int KH, int KW, __half* d_out, cudaStream_t st) {
  k_pack_tc<<<512, 256, 0, st>>>(d_w, transposed, Cin, Cout, KH, KW, d_out);
  RYK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ryk
