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
  for (int s = 1; s <= log2n; ++s) {
    int half = 1 << (s - 1);
    int tstep = kTwiddleN >> s;          // kTwiddleN / len
    for (int j = threadIdx.x; j < (n >> 1); j += T) {
      int k = j & (half - 1);
      int ia = ((j >> (s - 1)) << s) + k;
      int ib = ia + half;
      double2 w = __ldg(&tw[k * tstep]);
      double wi = sign < 0 ? w.y : -w.y;
      double2 vb = a[ib], va = a[ia];
      double xr = vb.x * w.x - vb.y * wi;
      double xi = vb.x * wi + vb.y * w.x;
      a[ib] = make_double2(va.x - xr, va.y - xi);
      a[ia] = make_double2(va.x + xr, va.y + xi);
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
