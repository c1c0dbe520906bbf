// analysis_shared.cuh -- kernels shared by the two f0 extractors (world_analysis.cu: DIO, world_harvest.cu: Harvest): band
// multiplication of the chunk spectrum with cached filter spectra and the ordered extraction of the four zero-crossing event trains
// (dio.cpp / harvest.cpp ZeroCrossingEngine + GetFourZeroCrossingIntervals).
#pragma once
#include <cufft.h>

#include "common.cuh"

namespace ryk {

static inline int cufft_ok(cufftResult r, const char* what) {
  if (r != CUFFT_SUCCESS) { set_error(std::string("cuFFT ") + what + " failed with code " + std::to_string((int)r)); return -1; }
  return 0;
}

// Z[b] = Y * LP[b]
static __global__ void k_band_mul(const cufftDoubleComplex* __restrict__ y, const cufftDoubleComplex* __restrict__ lp,
                           cufftDoubleComplex* __restrict__ z, int nbins) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int b = blockIdx.y;
  if (i < nbins) {
    cufftDoubleComplex a = y[i], w = lp[(size_t)b * nbins + i];
    z[(size_t)b * nbins + i] = make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
  }
}

// Ordered extraction of the four event trains of one band. grid = (4, nbands), block = 1024.
__device__ inline double dio_signal(const double* __restrict__ f, int type, int i) {
  switch (type) {
    case 0: return f[i];
    case 1: return -f[i];
    case 2: return (-f[i]) - (-f[i + 1]);
    default: return -((-f[i]) - (-f[i + 1]));
  }
}

static __global__ void __launch_bounds__(1024) k_dio_zero_cross(const double* __restrict__ filtered, int fft_size, int y_length,
                                                        const int* __restrict__ delay_src, int delay_mul, int delay_add, double fs,
                                                        int* __restrict__ edges, double* __restrict__ loc,
                                                        double* __restrict__ itv, int* __restrict__ counts) {
  int type = blockIdx.x, band = blockIdx.y;
  const double* f = filtered + (size_t)band * fft_size + delay_src[band] * delay_mul + delay_add;   // delay compensation
  int L = type < 2 ? y_length : y_length - 1;
  size_t slot = ((size_t)band * 4 + type) * y_length;
  int* e = edges + slot;
  __shared__ int wsum[32];
  __shared__ int total;
  int per = (L - 1 + blockDim.x - 1) / blockDim.x;   // candidates i in [0, L-1)
  int lo = threadIdx.x * per, hi = min(lo + per, L - 1);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) {
    double a = dio_signal(f, type, i), b = dio_signal(f, type, i + 1);
    cnt += (0.0 < a && b <= 0.0) ? 1 : 0;
  }
  // block exclusive scan of cnt
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane];
    int iv = v;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += u; }
    wsum[lane] = iv - v;
    if (lane == 31) total = iv;
  }
  __syncthreads();
  int pos = wsum[w] + inc - cnt;
  for (int i = lo; i < hi; ++i) {
    double a = dio_signal(f, type, i), b = dio_signal(f, type, i + 1);
    if (0.0 < a && b <= 0.0) e[pos++] = i + 1;
  }
  __syncthreads();
  int count = total;
  if (count < 2) { if (threadIdx.x == 0) counts[band * 4 + type] = 0; return; }
  for (int i = threadIdx.x; i < count - 1; i += blockDim.x) {
    int e0 = e[i], e1 = e[i + 1];
    double s0a = dio_signal(f, type, e0 - 1), s0b = dio_signal(f, type, e0);
    double s1a = dio_signal(f, type, e1 - 1), s1b = dio_signal(f, type, e1);
    double f0e = e0 - s0a / (s0b - s0a);
    double f1e = e1 - s1a / (s1b - s1a);
    itv[slot + i] = fs / (f1e - f0e);
    loc[slot + i] = (f0e + f1e) / 2.0 / fs;
  }
  if (threadIdx.x == 0) counts[band * 4 + type] = count - 1;
}


}  // namespace ryk
