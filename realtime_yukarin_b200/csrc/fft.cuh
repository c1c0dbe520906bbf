// fft.cuh -- CTA-cooperative, shared-memory resident FP64 FFT used by the fused per-frame WORLD
// kernels (CheapTrick, D4C, StoneMask, per-pulse synthesis).  One CTA owns one frame; the whole
// transform (n <= 4096 complex points) lives in shared memory, so a frame's FFT -> elementwise ->
// FFT chains never touch HBM.  Radix-2 decimation-in-time, same butterfly algebra as the CPU
// oracle's wo_fft (keeps FP64 results within a few ulp of it).
#pragma once
#include "common.cuh"

namespace ryk {

constexpr int kTwiddleN = 4096;   // table holds exp(-2 pi i k / kTwiddleN), k < kTwiddleN/2

// The twiddle table (kTwiddleN/2 double2 in global memory, L1/L2 resident) is owned by the engine
// and passed to every kernel as `tw`; fft_fill_twiddles computes it on the host exactly like the
// oracle does (cos/sin of 2 pi k / n in double).
void fft_fill_twiddles(double2* host_table);

// In-place complex FFT of a[0..n) in shared memory. sign=-1 forward, +1 inverse (unnormalised).
// All threads of the CTA must call; contains __syncthreads(). On return data is visible to all.
__device__ inline void fft_smem(double2* a, int n, int log2n, int sign, const double2* __restrict__ tw) {
  const int T = blockDim.x;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += T) {
    int j = (int)(__brev((unsigned)i) >> (32 - log2n));
    if (j > i) { double2 t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  __syncthreads();
  // Stages are taken two at a time: a thread owns the quad {i, i + h, i + 2h, i + 3h} (h = half length of stage s), does stage s on
  // (i, i + h) and (i + 2h, i + 3h) and stage s + 1 on (i, i + 2h) and (i + h, i + 3h) in registers -- the SAME butterflies with the
  // SAME table twiddles in the same order as the one-stage-per-barrier form (results are bit-identical), at half the block barriers
  // and half the shared-memory round trips.  These kernels are latency chains of barriers: k_d4c ran ~300 of them per frame.
  auto bfly = [&](double2& va, double2& vb, const double2 w) {
    const double wi = sign < 0 ? w.y : -w.y;
    const double xr = vb.x * w.x - vb.y * wi;
    const double xi = vb.x * wi + vb.y * w.x;
    vb = make_double2(va.x - xr, va.y - xi);
    va = make_double2(va.x + xr, va.y + xi);
  };
  int s = 1;
  for (; s + 1 <= log2n; s += 2) {
    const int half = 1 << (s - 1);
    const int t1 = kTwiddleN >> s, t2 = kTwiddleN >> (s + 1);
    for (int j = threadIdx.x; j < (n >> 2); j += T) {
      const int k = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << (s + 1)) + k;
      const int i1 = i0 + half, i2 = i0 + 2 * half, i3 = i0 + 3 * half;
      double2 a0 = a[i0], a1 = a[i1], a2 = a[i2], a3 = a[i3];
      const double2 w1 = __ldg(&tw[k * t1]);
      bfly(a0, a1, w1);
      bfly(a2, a3, w1);
      bfly(a0, a2, __ldg(&tw[k * t2]));
      bfly(a1, a3, __ldg(&tw[(k + half) * t2]));
      a[i0] = a0; a[i1] = a1; a[i2] = a2; a[i3] = a3;
    }
    __syncthreads();
  }
  if (s == log2n) {
    const int half = 1 << (s - 1);
    const int tstep = kTwiddleN >> s;          // kTwiddleN / len
    for (int j = threadIdx.x; j < (n >> 1); j += T) {
      const int k = j & (half - 1);
      const int ia = ((j >> (s - 1)) << s) + k;
      const int ib = ia + half;
      double2 va = a[ia], vb = a[ib];
      bfly(va, vb, __ldg(&tw[k * tstep]));
      a[ia] = va; a[ib] = vb;
    }
    __syncthreads();
  }
}

__device__ inline int ilog2(int n) { return 31 - __clz(n); }

// WORLD's GetMinimumPhaseSpectrum on a log-amplitude half spectrum.
// in : a[i].x = log_spectrum[i] for i in 0..n/2 (a[i].y ignored);  out: a[0..n/2] = minimum-phase spectrum.
__device__ inline void min_phase_smem(double2* a, int n, int log2n, const double2* __restrict__ tw) {
  const int T = blockDim.x;
  __syncthreads();
  for (int i = threadIdx.x; i <= n / 2; i += T) a[i].y = 0.0;
  __syncthreads();
  for (int i = n / 2 + 1 + threadIdx.x; i < n; i += T) a[i] = make_double2(a[n - i].x, 0.0);
  fft_smem(a, n, log2n, -1, tw);
  for (int i = threadIdx.x; i < n; i += T) {
    double2 v = a[i];
    if (i == 0 || i == n / 2) v = make_double2(v.x, 0.0);
    else if (i < n / 2) v = make_double2(v.x * 2.0, 0.0);
    else v = make_double2(0.0, 0.0);
    a[i] = v;
  }
  fft_smem(a, n, log2n, -1, tw);
  for (int i = threadIdx.x; i <= n / 2; i += T) {
    double2 v = a[i];
    double m = exp(v.x / n);
    double s, c;
    sincos(v.y / n, &s, &c);
    a[i] = make_double2(m * c, m * s);
  }
  __syncthreads();
}

// Hermitian completion + unnormalised inverse: in a[0..n/2] half spectrum; out a[i].x real signal.
__device__ inline void irfft_smem(double2* a, int n, int log2n, const double2* __restrict__ tw) {
  const int T = blockDim.x;
  __syncthreads();
  if (threadIdx.x == 0) { a[0].y = 0.0; a[n / 2].y = 0.0; }
  for (int k = 1 + threadIdx.x; k < n / 2; k += T) a[n - k] = make_double2(a[k].x, -a[k].y);
  fft_smem(a, n, log2n, +1, tw);
}

}  // namespace ryk
