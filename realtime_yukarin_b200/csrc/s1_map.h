// s1_map.h -- index arithmetic of the fused stage-1 kernel (s1_fused.cu), shared between the device code, the host-side weight
// packer and the CPU emulation in tests/s1_fused_emulate.cpp (which replays the kernel's loops warp by warp with software
// ldmatrix / mma.sync m16n8k16 and checks them against a direct convolution: every function here is exercised without a GPU).
//
// Layers 1..14 of the 1-D U-Net (unet.cu: conv k4 s2 p1 / transposed conv k4 s2 p1 over NHWC rows [W][C]) as GEMMs
//   out[m][n] = sum_k A[m][k] * Wt[n][k]
//   conv   : m = output pixel,               k = tap * Cin + c, A[m][k] = in[2m - 1 + tap][c]          (tap 0..3,  K = 4 Cin)
//   deconv : m = input-aligned pixel, class r = output parity (out pixel 2m + r), k = j * Cin + c (j 0..1, K = 2 Cin):
//            r = 0: j = 0 -> (in[m],   tap 1), j = 1 -> (in[m-1], tap 3);   r = 1: j = 0 -> (in[m], tap 2), j = 1 -> (in[m+1], tap 0)
//            (out[o] += in[i] W[c][n][t] with o = 2 i - 1 + t, the Chainer / torch transposed convolution)
// Out-of-range input pixels are zero (padding 1).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define S1_HD __host__ __device__ __forceinline__
#else
#define S1_HD inline
#endif

namespace ryk {

constexpr int kS1Warps = 16;                 // warps per CTA
constexpr int kS1Threads = kS1Warps * 32;
constexpr int kS1ActBytes = 104 * 1024;      // staged input rows of one layer (one M slab): every layer of a 384-frame window is ONE slab
constexpr int kS1PartialBytes = kS1Warps * 16 * 32 * 4;   // split-K partial accumulators: [warp][16 regs][32 lanes] floats
constexpr int kS1RowPad = 8;                 // halfs of padding per staged row (keeps rows 16-byte aligned, spreads banks)

struct S1Geom {
  int transposed, Win, Cin, Cout;
};

S1_HD int s1_K(const S1Geom& g) { return (g.transposed ? 2 : 4) * g.Cin; }
S1_HD int s1_M(const S1Geom& g) { return g.transposed ? g.Win : g.Win / 2; }          // GEMM rows per class
S1_HD int s1_classes(const S1Geom& g) { return g.transposed ? 2 : 1; }
S1_HD int s1_Wout(const S1Geom& g) { return g.transposed ? g.Win * 2 : g.Win / 2; }
S1_HD int s1_tasks(const S1Geom& g) { return s1_classes(g) * (g.Cout / 16); }          // (class, 16-column group)
S1_HD size_t s1_task_halfs(const S1Geom& g) { return (size_t)s1_K(g) * 16; }           // fragment-packed weights of one task

// input pixel read by GEMM row m for K-slice j (= k / Cin) of class cls
S1_HD int s1_in_px(const S1Geom& g, int cls, int m, int j) {
  if (!g.transposed) return 2 * m - 1 + j;
  return j == 0 ? m : (cls == 0 ? m - 1 : m + 1);
}
S1_HD int s1_out_px(const S1Geom& g, int cls, int m) { return g.transposed ? 2 * m + cls : m; }
S1_HD int s1_tap(const S1Geom& g, int cls, int j) {
  if (!g.transposed) return j;
  return cls == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 2 : 0);
}
// element (cls, n, k) in the model file's Chainer layout: conv (Cout, Cin, 4), deconv (Cin, Cout, 4)
S1_HD size_t s1_w_src(const S1Geom& g, int cls, int n, int k) {
  const int j = k / g.Cin, c = k % g.Cin, t = s1_tap(g, cls, j);
  return g.transposed ? ((size_t)c * g.Cout + n) * 4 + t : ((size_t)n * g.Cin + c) * 4 + t;
}
// ... and in the fragment-packed array (index in halfs).  Per task (cls, ng = n / 16): [kp = k / 32][nt = (n / 8) % 2][lane][8 halfs],
// the 16 bytes of a lane being the mma.sync m16n8k16 B fragments {b0, b1} of k-tile 2 kp and {b0, b1} of k-tile 2 kp + 1:
//   lane = (n % 8) * 4 + (k % 8) / 2,   b0: k % 16 < 8, b1: k % 16 >= 8,   two consecutive k per 32-bit register.
S1_HD size_t s1_w_dst(const S1Geom& g, int cls, int n, int k) {
  const int K = s1_K(g), NG = g.Cout / 16;
  const int ng = n / 16, nt = (n / 8) % 2, nr = n % 8;
  const int kp = k / 32, kt = (k / 16) % 2, kk = k % 16;
  const int lane = nr * 4 + (kk % 8) / 2, reg = kt * 2 + kk / 8;
  const size_t task = (size_t)cls * NG + ng;
  return task * ((size_t)K * 16) + ((size_t)(kp * 2 + nt) * 32 + lane) * 8 + reg * 2 + (kk % 2);
}

// How a cluster of `nc` CTAs cuts one layer: MS slabs along M (so that a CTA stages at most kS1ActBytes of input rows) x nc / MS
// partitions of the (class, column group) tasks.
struct S1Cut {
  int MS;           // M slabs (power of two, <= nc)
  int slab;         // GEMM rows per slab (multiple of 16)
  int NP;           // task partitions = nc / MS
  int ms, ks;       // warps of a CTA: ms m-groups x ks K-splits (ms * ks <= kS1Warps)
  int RS;           // staged row stride in halfs
};
S1_HD int s1_rows_for(const S1Geom& g, int mrows) { return g.transposed ? mrows + 2 : 2 * mrows + 2; }   // input pixels touched by mrows GEMM rows
S1_HD S1Cut s1_cut(const S1Geom& g, int nc) {
  S1Cut c;
  const int M = s1_M(g);
  c.RS = g.Cin + kS1RowPad;
  c.MS = 1;
  for (;;) {
    int slab = ((M + c.MS - 1) / c.MS + 15) / 16 * 16;
    if ((size_t)s1_rows_for(g, slab < M ? slab : M) * c.RS * 2 <= (size_t)kS1ActBytes || c.MS >= nc) { c.slab = slab; break; }
    c.MS *= 2;
  }
  c.NP = nc / c.MS;
  const int mt = (c.slab + 15) / 16;                 // m-tiles of a slab
  int ms = 1;
  while (ms < kS1Warps && ms < (mt + 1) / 2) ms *= 2;   // smallest power of two >= ceil(mt / 2): two m-tiles per warp and pass
  int ks = kS1Warps / ms;
  const int KP = s1_K(g) / 32;
  while (ks > 1 && ks * 2 > KP) ks /= 2;              // every K-split gets an even number (>= 2) of k-tile pairs
  c.ms = ms; c.ks = ks;
  return c;
}
// first staged input pixel of the slab starting at GEMM row m0 (staged row index = pixel - s1_px0)
S1_HD int s1_px0(const S1Geom& g, int m0) { return g.transposed ? m0 - 1 : 2 * m0 - 1; }

// ldmatrix.x4 row address of `lane` for the 16x16 A tile at GEMM rows mbase.., k-tile channel offset ch0: lanes 0-7 -> rows 0-7 / k 0-7,
// 8-15 -> rows 8-15 / k 0-7, 16-23 -> rows 0-7 / k 8-15, 24-31 -> rows 8-15 / k 8-15  (= a0, a1, a2, a3 of mma.m16n8k16)
S1_HD int s1_ldm_row(int lane) { return (lane & 7) + ((lane >> 3) & 1) * 8; }
S1_HD int s1_ldm_kofs(int lane) { return (lane >> 4) * 8; }
// accumulator register r (0..3) of `lane` in a 16x8 C tile: row = lane / 4 + 8 * (r / 2), col = (lane % 4) * 2 + r % 2
S1_HD int s1_c_row(int lane, int r) { return (lane >> 2) + (r >> 1) * 8; }
S1_HD int s1_c_col(int lane, int r) { return (lane & 3) * 2 + (r & 1); }

}  // namespace ryk
