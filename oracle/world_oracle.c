/*
 * oracle/world_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double precision) of the third-party arithmetic that
 * realtime-yukarin's hot path executes: WORLD analysis (DIO + StoneMask, CheapTrick, D4C),
 * SPTK mel-cepstrum conversion (freqt / sp2mc / mc2sp) and the WORLD *realtime* synthesizer.
 *
 * PARITY UNPINNED: none of these sources exist under /root/reference (pyworld, world4py, pysptk
 * are un-vendored, unpinned pip dependencies: requirements.txt:1-8, setup.py:12-19) and the
 * reference ships no golden vectors.  The algorithms below restate the published WORLD
 * (M. Morise, github.com/mmorise/World, v0.2.x: dio.cpp, stonemask.cpp, cheaptrick.cpp, d4c.cpp,
 * synthesisrealtime.cpp, common.cpp, matlabfunctions.cpp) and SPTK 3.x (freqt.c) / pysptk
 * conversion.py algorithms.  Reference call sites that reach this arithmetic:
 *   - analysis : realtime_voice_conversion/yukarin_wrapper/vocoder.py:26-48
 *                -> acoustic_feature_wrapper.py:28-33 -> yukarin.AcousticFeature.extract
 *   - mc2sp    : realtime_voice_conversion/yukarin_wrapper/voice_changer.py:38
 *   - synthesis: realtime_voice_conversion/yukarin_wrapper/vocoder.py:72-120
 *                (_InitializeSynthesizer / _AddParameters / _Synthesis2)
 * Every point where the upstream behaviour could not be recovered is marked DECIDE and is
 * frozen here; DESIGN.md lists them.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WO_PI 3.1415926535897932384
#define WO_LOG2 0.69314718055994529
#define WO_CUTOFF 50.0
#define WO_FLOOR_F0_STONEMASK 40.0
#define WO_SAFE_MIN 0.000000000001
#define WO_EPS 0.00000000000000022204460492503131
#define WO_DEFAULT_F0 500.0
#define WO_MAX_VALUE 100000.0
#define WO_FREQ_INTERVAL 3000.0
#define WO_UPPER_LIMIT 15000.0
#define WO_FLOOR_F0_D4C 47.0

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

int wo_matlab_round(double x) { return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5); }

int wo_suitable_fft_size(int sample) {
  return (int)pow(2.0, (int)(log((double)sample) / WO_LOG2) + 1.0);
}

int wo_cheaptrick_fft_size(int fs, double f0_floor) {
  return (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / f0_floor + 1) / WO_LOG2));
}

/* ------------------------------------------------------------------ FFT (radix-2, double) */
typedef struct { int n; double *wr, *wi; int *rev; } wo_plan;
static wo_plan g_plans[24];

static wo_plan *get_plan(int n) {
  int lg = 0;
  while ((1 << lg) < n) lg++;
  wo_plan *p = &g_plans[lg];
  if (p->n == n) return p;
  p->n = n;
  p->wr = (double *)malloc(sizeof(double) * (n / 2 + 1));
  p->wi = (double *)malloc(sizeof(double) * (n / 2 + 1));
  p->rev = (int *)malloc(sizeof(int) * n);
  for (int k = 0; k < n / 2; ++k) {
    p->wr[k] = cos(2.0 * WO_PI * k / n);
    p->wi[k] = -sin(2.0 * WO_PI * k / n);
  }
  for (int i = 0; i < n; ++i) {
    int r = 0;
    for (int b = 0; b < lg; ++b) if (i & (1 << b)) r |= 1 << (lg - 1 - b);
    p->rev[i] = r;
  }
  return p;
}

/* in-place complex FFT. sign=-1: forward (e^{-i}), sign=+1: inverse (unnormalised). */
void wo_fft(double *re, double *im, int n, int sign) {
  wo_plan *p = get_plan(n);
  for (int i = 0; i < n; ++i) {
    int j = p->rev[i];
    if (j > i) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1, step = n / len;
    for (int s = 0; s < n; s += len) {
      for (int k = 0; k < half; ++k) {
        double wr = p->wr[k * step];
        double wi = sign < 0 ? p->wi[k * step] : -p->wi[k * step];
        int a = s + k, b = s + k + half;
        double xr = re[b] * wr - im[b] * wi;
        double xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
    }
  }
}

/* real input (n) -> spectrum bins 0..n/2 (forward). work arrays are allocated by caller (n). */
static void rfft(const double *x, int n, double *re, double *im) {
  for (int i = 0; i < n; ++i) { re[i] = x[i]; im[i] = 0.0; }
  wo_fft(re, im, n, -1);
}

/* hermitian half spectrum (bins 0..n/2 in re/im, arrays of size n) -> real signal, unnormalised */
static void irfft_unnorm(double *re, double *im, int n, double *out) {
  for (int k = 1; k < n / 2; ++k) { re[n - k] = re[k]; im[n - k] = -im[k]; }
  im[0] = 0.0; im[n / 2] = 0.0;
  wo_fft(re, im, n, +1);
  for (int i = 0; i < n; ++i) out[i] = re[i];
}

/* ------------------------------------------------------------------ matlab helpers */
/* interp1: k = clamp(#{x[j] <= xi}, 1, n-1); linear (extrapolating) -- matlabfunctions.cpp histc+interp1 */
static int histc_index(const double *x, int n, double xi) {
  int lo = 0, hi = n;  /* first index with x[idx] > xi */
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (x[mid] <= xi) lo = mid + 1; else hi = mid;
  }
  if (lo < 1) lo = 1;
  if (lo > n - 1) lo = n - 1;
  return lo;
}

void wo_interp1(const double *x, const double *y, int n, const double *xi, int m, double *yi) {
  for (int i = 0; i < m; ++i) {
    int k = histc_index(x, n, xi[i]);
    double s = (xi[i] - x[k - 1]) / (x[k] - x[k - 1]);
    yi[i] = y[k - 1] + s * (y[k] - y[k - 1]);
  }
}

/* interp1Q: equally spaced x (origin x0, spacing dx) */
static void interp1Q(double x0, double dx, const double *y, int n, const double *xi, int m, double *yi) {
  for (int i = 0; i < m; ++i) {
    double pos = (xi[i] - x0) / dx;
    int base = (int)pos;
    double frac = pos - base;
    double dy = base + 1 < n ? y[base + 1] - y[base] : 0.0;   /* delta_y[n-1] = 0 */
    yi[i] = y[base] + dy * frac;
  }
}

static void nuttall_window(int n, double *y) {
  for (int i = 0; i < n; ++i) {
    double tmp = i / (n - 1.0);
    y[i] = 0.355768 - 0.487396 * cos(2.0 * WO_PI * tmp) + 0.144232 * cos(4.0 * WO_PI * tmp) -
           0.012604 * cos(6.0 * WO_PI * tmp);
  }
}

/* xorshift128 randn (matlabfunctions.cpp) -- explicit state so that streams are independent */
typedef struct { uint32_t x, y, z, w; } wo_rng;
void wo_rng_seed(wo_rng *r) { r->x = 123456789u; r->y = 362436069u; r->z = 521288629u; r->w = 88675123u; }
double wo_randn(wo_rng *r) {
  uint32_t t, tmp = 0;
  for (int i = 0; i < 12; ++i) {
    t = r->x ^ (r->x << 11);
    r->x = r->y; r->y = r->z; r->z = r->w;
    r->w = (r->w ^ (r->w >> 19)) ^ (t ^ (t >> 8));
    tmp += r->w >> 4;
  }
  return tmp / 268435456.0 - 6.0;
}
/* n consecutive draws starting at draw index `start` of the canonical stream (test helper) */
void wo_randn_stream(long long start, int n, double *out) {
  wo_rng r; wo_rng_seed(&r);
  for (long long i = 0; i < start; ++i) wo_randn(&r);
  for (int i = 0; i < n; ++i) out[i] = wo_randn(&r);
}

/* common.cpp DCCorrection */
static void dc_correction(const double *input, double f0, int fs, int fft_size, double *output) {
  int upper_limit = 2 + (int)(f0 * fft_size / fs);
  double *replica = (double *)malloc(sizeof(double) * upper_limit);
  double *axis = (double *)malloc(sizeof(double) * upper_limit);
  for (int i = 0; i < upper_limit; ++i) axis[i] = (double)i * fs / fft_size;
  int upper_limit_replica = upper_limit - 1;
  interp1Q(f0 - axis[0], -(double)fs / fft_size, input, upper_limit + 1, axis, upper_limit_replica, replica);
  for (int i = 0; i < upper_limit_replica; ++i) output[i] = input[i] + replica[i];
  free(replica); free(axis);
}

/* common.cpp LinearSmoothing (in-place safe) */
static void linear_smoothing(const double *input, double width, int fs, int fft_size, double *output) {
  int boundary = (int)(width * fft_size / fs) + 1;
  int half = fft_size / 2;
  int mlen = half + boundary * 2 + 1;
  double *mirror = (double *)malloc(sizeof(double) * mlen);
  double *segment = (double *)malloc(sizeof(double) * mlen);
  double *axis = (double *)malloc(sizeof(double) * (half + 1));
  double *low = (double *)malloc(sizeof(double) * (half + 1));
  double *high = (double *)malloc(sizeof(double) * (half + 1));
  for (int i = 0; i < boundary; ++i) mirror[i] = input[boundary - i];
  for (int i = boundary; i < half + boundary; ++i) mirror[i] = input[i - boundary];
  for (int i = half + boundary; i <= half + boundary * 2; ++i) mirror[i] = input[half - (i - (half + boundary))];
  segment[0] = mirror[0] * fs / fft_size;
  for (int i = 1; i < mlen; ++i) segment[i] = mirror[i] * fs / fft_size + segment[i - 1];
  for (int i = 0; i <= half; ++i) axis[i] = (double)i / fft_size * fs - width / 2.0;
  double origin = -(boundary - 0.5) * fs / fft_size;
  double interval = (double)fs / fft_size;
  interp1Q(origin, interval, segment, mlen, axis, half + 1, low);
  for (int i = 0; i <= half; ++i) axis[i] += width;
  interp1Q(origin, interval, segment, mlen, axis, half + 1, high);
  for (int i = 0; i <= half; ++i) output[i] = (high[i] - low[i]) / width;
  free(mirror); free(segment); free(axis); free(low); free(high);
}

/* ------------------------------------------------------------------ DIO (dio.cpp) */
int wo_dio_num_frames(int fs, int x_length, double frame_period) {
  return (int)(1000.0 * x_length / fs / frame_period) + 1;
}

static void design_low_cut_filter(int N, int fft_size, double *f) {
  for (int i = 1; i <= N; ++i) f[i - 1] = 0.5 - 0.5 * cos(i * 2.0 * WO_PI / (N + 1));
  for (int i = N; i < fft_size; ++i) f[i] = 0.0;
  double sum = 0.0;
  for (int i = 0; i < N; ++i) sum += f[i];
  for (int i = 0; i < N; ++i) f[i] = -f[i] / sum;
  for (int i = 0; i < (N - 1) / 2; ++i) f[fft_size - (N - 1) / 2 + i] = f[i];
  for (int i = 0; i < N; ++i) f[i] = f[i + (N - 1) / 2];
  f[0] += 1.0;
}

static int zero_crossing_engine(const double *sig, int y_length, double fs, double *locations, double *intervals) {
  int *edges = (int *)malloc(sizeof(int) * y_length);
  int count = 0;
  for (int i = 0; i < y_length - 1; ++i)
    if (0.0 < sig[i] && sig[i + 1] <= 0.0) edges[count++] = i + 1;
  if (count < 2) { free(edges); return 0; }
  double *fine = (double *)malloc(sizeof(double) * count);
  for (int i = 0; i < count; ++i)
    fine[i] = edges[i] - sig[edges[i] - 1] / (sig[edges[i]] - sig[edges[i] - 1]);
  for (int i = 0; i < count - 1; ++i) {
    intervals[i] = fs / (fine[i + 1] - fine[i]);
    locations[i] = (fine[i] + fine[i + 1]) / 2.0 / fs;
  }
  free(edges); free(fine);
  return count - 1;
}

/* out_cand/out_score: optional [bands][f0_length] dumps of raw candidates / normalised scores */
void wo_dio(const double *x, int x_length, int fs, double frame_period, double f0_floor, double f0_ceil,
            double *temporal_positions, double *f0, double *out_cand, double *out_score) {
  const double channels_in_octave = 2.0, allowed_range = 0.1;
  int nbands = 1 + (int)(log(f0_ceil / f0_floor) / WO_LOG2 * channels_in_octave);
  double *boundary = (double *)malloc(sizeof(double) * nbands);
  for (int i = 0; i < nbands; ++i) boundary[i] = f0_floor * pow(2.0, (i + 1) / channels_in_octave);
  int y_length = 1 + x_length;              /* decimation ratio 1 (pyworld speed=1) */
  double actual_fs = (double)fs;
  int fft_size = wo_suitable_fft_size(y_length + wo_matlab_round(actual_fs / WO_CUTOFF) * 2 + 1 +
                                      (4 * (int)(1.0 + actual_fs / boundary[0] / 2.0)));
  int f0_length = wo_dio_num_frames(fs, x_length, frame_period);
  for (int i = 0; i < f0_length; ++i) temporal_positions[i] = i * frame_period / 1000.0;

  /* GetSpectrumForEstimation */
  double *y = (double *)calloc(fft_size, sizeof(double));
  double *yr = (double *)malloc(sizeof(double) * fft_size), *yi = (double *)malloc(sizeof(double) * fft_size);
  double *fr = (double *)malloc(sizeof(double) * fft_size), *fi = (double *)malloc(sizeof(double) * fft_size);
  for (int i = 0; i < x_length; ++i) y[i] = x[i];
  double mean_y = 0.0;
  for (int i = 0; i < y_length; ++i) mean_y += y[i];
  mean_y /= y_length;
  for (int i = 0; i < y_length; ++i) y[i] -= mean_y;
  for (int i = y_length; i < fft_size; ++i) y[i] = 0.0;
  rfft(y, fft_size, yr, yi);
  int cutoff_in_sample = wo_matlab_round(actual_fs / WO_CUTOFF);
  design_low_cut_filter(cutoff_in_sample * 2 + 1, fft_size, y);
  rfft(y, fft_size, fr, fi);
  for (int i = 0; i <= fft_size / 2; ++i) {
    double tmp = yr[i] * fr[i] - yi[i] * fi[i];
    yi[i] = yr[i] * fi[i] + yi[i] * fr[i];
    yr[i] = tmp;
  }

  double *cand = (double *)malloc(sizeof(double) * nbands * f0_length);
  double *score = (double *)malloc(sizeof(double) * nbands * f0_length);
  double *lp = (double *)malloc(sizeof(double) * fft_size);
  double *filtered = (double *)malloc(sizeof(double) * fft_size);
  double *loc[4], *itv[4], *interp[4];
  for (int e = 0; e < 4; ++e) {
    loc[e] = (double *)malloc(sizeof(double) * y_length);
    itv[e] = (double *)malloc(sizeof(double) * y_length);
    interp[e] = (double *)malloc(sizeof(double) * f0_length);
  }
  for (int b = 0; b < nbands; ++b) {
    /* GetFilteredSignal */
    int half_average_length = wo_matlab_round(actual_fs / boundary[b] / 2.0);
    nuttall_window(half_average_length * 4, lp);
    for (int i = half_average_length * 4; i < fft_size; ++i) lp[i] = 0.0;
    rfft(lp, fft_size, fr, fi);
    for (int i = 0; i <= fft_size / 2; ++i) {
      double tmp = yr[i] * fr[i] - yi[i] * fi[i];
      fi[i] = yr[i] * fi[i] + yi[i] * fr[i];
      fr[i] = tmp;
    }
    irfft_unnorm(fr, fi, fft_size, filtered);
    int index_bias = half_average_length * 2;
    for (int i = 0; i < y_length; ++i) filtered[i] = filtered[i + index_bias];
    /* GetFourZeroCrossingIntervals */
    int cnt[4];
    cnt[0] = zero_crossing_engine(filtered, y_length, actual_fs, loc[0], itv[0]);
    for (int i = 0; i < y_length; ++i) filtered[i] = -filtered[i];
    cnt[1] = zero_crossing_engine(filtered, y_length, actual_fs, loc[1], itv[1]);
    for (int i = 0; i < y_length - 1; ++i) filtered[i] = filtered[i] - filtered[i + 1];
    cnt[2] = zero_crossing_engine(filtered, y_length - 1, actual_fs, loc[2], itv[2]);
    for (int i = 0; i < y_length - 1; ++i) filtered[i] = -filtered[i];
    cnt[3] = zero_crossing_engine(filtered, y_length - 1, actual_fs, loc[3], itv[3]);
    double *c = cand + (size_t)b * f0_length, *s = score + (size_t)b * f0_length;
    if (!(cnt[0] > 2 && cnt[1] > 2 && cnt[2] > 2 && cnt[3] > 2)) {
      for (int i = 0; i < f0_length; ++i) { s[i] = WO_MAX_VALUE; c[i] = 0.0; }
    } else {
      for (int e = 0; e < 4; ++e) wo_interp1(loc[e], itv[e], cnt[e], temporal_positions, f0_length, interp[e]);
      for (int i = 0; i < f0_length; ++i) {
        c[i] = (interp[0][i] + interp[1][i] + interp[2][i] + interp[3][i]) / 4.0;
        s[i] = sqrt(((interp[0][i] - c[i]) * (interp[0][i] - c[i]) + (interp[1][i] - c[i]) * (interp[1][i] - c[i]) +
                     (interp[2][i] - c[i]) * (interp[2][i] - c[i]) + (interp[3][i] - c[i]) * (interp[3][i] - c[i])) / 3.0);
        if (c[i] > boundary[b] || c[i] < boundary[b] / 2.0 || c[i] > f0_ceil || c[i] < f0_floor) {
          c[i] = 0.0; s[i] = WO_MAX_VALUE;
        }
      }
    }
    for (int i = 0; i < f0_length; ++i) s[i] = s[i] / (c[i] + WO_SAFE_MIN);
  }
  if (out_cand) memcpy(out_cand, cand, sizeof(double) * nbands * f0_length);
  if (out_score) memcpy(out_score, score, sizeof(double) * nbands * f0_length);

  /* GetBestF0Contour */
  double *best = (double *)malloc(sizeof(double) * f0_length);
  for (int i = 0; i < f0_length; ++i) {
    double tmp = score[i];
    best[i] = cand[i];
    for (int b = 1; b < nbands; ++b)
      if (tmp > score[(size_t)b * f0_length + i]) { tmp = score[(size_t)b * f0_length + i]; best[i] = cand[(size_t)b * f0_length + i]; }
  }

  /* FixF0Contour */
  for (int i = 0; i < f0_length; ++i) f0[i] = 0.0;
  int vrm = (int)(0.5 + 1000.0 / frame_period / f0_floor) * 2 + 1;
  if (f0_length > vrm) {
    double *t1 = (double *)malloc(sizeof(double) * f0_length), *t2 = (double *)malloc(sizeof(double) * f0_length);
    double *base = (double *)malloc(sizeof(double) * f0_length);
    /* step 1 */
    for (int i = 0; i < f0_length; ++i) base[i] = (i < vrm || i >= f0_length - vrm) ? 0.0 : best[i];
    for (int i = 0; i < vrm; ++i) t1[i] = 0.0;
    for (int i = vrm; i < f0_length; ++i)
      t1[i] = fabs((base[i] - base[i - 1]) / (WO_SAFE_MIN + base[i])) < allowed_range ? base[i] : 0.0;
    /* step 2 */
    for (int i = 0; i < f0_length; ++i) t2[i] = t1[i];
    int center = (vrm - 1) / 2;
    for (int i = center; i < f0_length - center; ++i)
      for (int j = -center; j <= center; ++j)
        if (t1[i + j] == 0) { t2[i] = 0.0; break; }
    /* voiced sections */
    int *pos = (int *)malloc(sizeof(int) * f0_length), *neg = (int *)malloc(sizeof(int) * f0_length);
    int pc = 0, nc = 0;
    for (int i = 1; i < f0_length; ++i) {
      if (t2[i] == 0 && t2[i - 1] != 0) neg[nc++] = i - 1;
      else if (t2[i - 1] == 0 && t2[i] != 0) pos[pc++] = i;
    }
    /* step 3 (forward), SelectBestF0 inlined */
    for (int i = 0; i < f0_length; ++i) t1[i] = t2[i];
    for (int i = 0; i < nc; ++i) {
      int limit = i == nc - 1 ? f0_length - 1 : neg[i + 1];
      for (int j = neg[i]; j < limit; ++j) {
        double ref = (t1[j] * 3.0 - t1[j - 1]) / 2.0;
        double minerr = fabs(ref - cand[j + 1]), bestf = cand[j + 1];
        for (int b = 1; b < nbands; ++b) {
          double err = fabs(ref - cand[(size_t)b * f0_length + j + 1]);
          if (err < minerr) { minerr = err; bestf = cand[(size_t)b * f0_length + j + 1]; }
        }
        if (fabs(1.0 - bestf / ref) > allowed_range) bestf = 0.0;
        t1[j + 1] = bestf;
        if (bestf == 0) break;
      }
    }
    /* step 4 (backward) */
    for (int i = 0; i < f0_length; ++i) f0[i] = t1[i];
    for (int i = pc - 1; i >= 0; --i) {
      int limit = i == 0 ? 1 : pos[i - 1];
      for (int j = pos[i]; j > limit; --j) {
        double ref = (f0[j] * 3.0 - f0[j + 1]) / 2.0;
        double minerr = fabs(ref - cand[j - 1]), bestf = cand[j - 1];
        for (int b = 1; b < nbands; ++b) {
          double err = fabs(ref - cand[(size_t)b * f0_length + j - 1]);
          if (err < minerr) { minerr = err; bestf = cand[(size_t)b * f0_length + j - 1]; }
        }
        if (fabs(1.0 - bestf / ref) > allowed_range) bestf = 0.0;
        f0[j - 1] = bestf;
        if (bestf == 0) break;
      }
    }
    free(t1); free(t2); free(base); free(pos); free(neg);
  }
  for (int e = 0; e < 4; ++e) { free(loc[e]); free(itv[e]); free(interp[e]); }
  free(boundary); free(y); free(yr); free(yi); free(fr); free(fi); free(cand); free(score);
  free(lp); free(filtered); free(best);
}

/* ------------------------------------------------------------------ StoneMask (stonemask.cpp) */
static double stonemask_fix_f0(const double *power, const double *numer, int fft_size, int fs, double f0_initial, int nh) {
  double numerator = 0.0, denominator = 0.0;
  for (int i = 0; i < nh; ++i) {
    int index = wo_matlab_round(f0_initial * fft_size / fs * (i + 1));
    double inst = power[index] == 0.0 ? 0.0
                  : (double)index * fs / fft_size + numer[index] / power[index] * fs / 2.0 / WO_PI;
    double amp = sqrt(power[index]);
    numerator += amp * inst;
    denominator += amp * (i + 1);
  }
  return numerator / (denominator + WO_SAFE_MIN);
}

static double stonemask_refine(const double *x, int x_length, int fs, double pos, double f0_initial) {
  if (f0_initial <= WO_FLOOR_F0_STONEMASK || f0_initial > fs / 12.0) return 0.0;
  int half = (int)(1.5 * fs / f0_initial + 1.0);
  double wlen_time = (2.0 * half + 1.0) / fs;
  int blen = half * 2 + 1;
  int fft_size = (int)pow(2.0, 2.0 + (int)(log(half * 2.0 + 1.0) / WO_LOG2));
  double *mainw = (double *)malloc(sizeof(double) * blen), *diffw = (double *)malloc(sizeof(double) * blen);
  double *buf = (double *)calloc(fft_size, sizeof(double));
  double *mr = (double *)malloc(sizeof(double) * fft_size), *mi = (double *)malloc(sizeof(double) * fft_size);
  double *dr = (double *)malloc(sizeof(double) * fft_size), *di = (double *)malloc(sizeof(double) * fft_size);
  int basic_index = wo_matlab_round((pos + (double)(-half) / fs) * fs + 0.001);
  for (int i = 0; i < blen; ++i) {
    double tmp = ((basic_index + i) - 1.0) / fs - pos;
    mainw[i] = 0.42 + 0.5 * cos(2.0 * WO_PI * tmp / wlen_time) + 0.08 * cos(4.0 * WO_PI * tmp / wlen_time);
  }
  diffw[0] = -mainw[1] / 2.0;
  for (int i = 1; i < blen - 1; ++i) diffw[i] = -(mainw[i + 1] - mainw[i - 1]) / 2.0;
  diffw[blen - 1] = mainw[blen - 2] / 2.0;
  for (int i = 0; i < blen; ++i) buf[i] = x[imax(0, imin(x_length - 1, basic_index + i - 1))] * mainw[i];
  rfft(buf, fft_size, mr, mi);
  for (int i = 0; i < blen; ++i) buf[i] = x[imax(0, imin(x_length - 1, basic_index + i - 1))] * diffw[i];
  rfft(buf, fft_size, dr, di);
  double *power = buf;   /* reuse */
  double *numer = (double *)malloc(sizeof(double) * (fft_size / 2 + 1));
  for (int j = 0; j <= fft_size / 2; ++j) {
    numer[j] = mr[j] * di[j] - mi[j] * dr[j];
    power[j] = mr[j] * mr[j] + mi[j] * mi[j];
  }
  double tentative = stonemask_fix_f0(power, numer, fft_size, fs, f0_initial, 2);
  double mean_f0;
  if (tentative <= 0.0 || tentative > f0_initial * 2) mean_f0 = 0.0;
  else mean_f0 = stonemask_fix_f0(power, numer, fft_size, fs, tentative, 6);
  if (fabs(mean_f0 - f0_initial) > f0_initial * 0.2) mean_f0 = f0_initial;
  free(mainw); free(diffw); free(buf); free(mr); free(mi); free(dr); free(di); free(numer);
  return mean_f0;
}

void wo_stonemask(const double *x, int x_length, int fs, const double *tpos, const double *f0, int f0_length, double *refined) {
  for (int i = 0; i < f0_length; ++i) refined[i] = stonemask_refine(x, x_length, fs, tpos[i], f0[i]);
}

/* ------------------------------------------------------------------ CheapTrick (cheaptrick.cpp)
 * DECIDE: the two randn() injections (1e-12 on the waveform, |randn|*eps on the spectrum) are
 * replaced by 0 and by the constant eps respectively, so that CPU and GPU agree sample-exactly
 * on what a silent (all-zero) frame produces. */
void wo_cheaptrick(const double *x, int x_length, int fs, const double *tpos, const double *f0, int f0_length,
                   int fft_size, double q1, double *spectrogram) {
  double f0_floor = 3.0 * fs / (fft_size - 3.0);
  int half_fft = fft_size / 2;
  double *wave = (double *)malloc(sizeof(double) * fft_size);
  double *re = (double *)malloc(sizeof(double) * fft_size), *im = (double *)malloc(sizeof(double) * fft_size);
  double *window = (double *)malloc(sizeof(double) * fft_size);
  for (int f = 0; f < f0_length; ++f) {
    double cf0 = f0[f] <= f0_floor ? WO_DEFAULT_F0 : f0[f];
    int half = wo_matlab_round(1.5 * fs / cf0);
    int origin = wo_matlab_round(tpos[f] * fs + 0.001);
    double average = 0.0;
    for (int i = 0; i <= half * 2; ++i) {
      double position = (i - half) / 1.5 / fs;
      window[i] = 0.5 * cos(WO_PI * position * cf0) + 0.5;
      average += window[i] * window[i];
    }
    average = sqrt(average);
    for (int i = 0; i <= half * 2; ++i) window[i] /= average;
    double w1 = 0, w2 = 0;
    for (int i = 0; i <= half * 2; ++i) {
      int idx = imin(x_length - 1, imax(0, origin + i - half));
      wave[i] = x[idx] * window[i];
      w1 += wave[i]; w2 += window[i];
    }
    double coef = w1 / w2;
    for (int i = 0; i <= half * 2; ++i) wave[i] -= window[i] * coef;
    for (int i = half * 2 + 1; i < fft_size; ++i) wave[i] = 0.0;
    rfft(wave, fft_size, re, im);
    for (int i = 0; i <= half_fft; ++i) wave[i] = re[i] * re[i] + im[i] * im[i];
    dc_correction(wave, cf0, fs, fft_size, wave);
    linear_smoothing(wave, cf0 * 2.0 / 3.0, fs, fft_size, wave);
    for (int i = 0; i <= half_fft; ++i) wave[i] = wave[i] + WO_EPS;
    /* SmoothingWithRecovery */
    for (int i = 0; i <= half_fft; ++i) wave[i] = log(wave[i]);
    for (int i = 1; i < half_fft; ++i) wave[fft_size - i] = wave[i];
    rfft(wave, fft_size, re, im);
    for (int i = 0; i <= half_fft; ++i) {
      double sl, cl;
      if (i == 0) { sl = 1.0; cl = (1.0 - 2.0 * q1) + 2.0 * q1; }
      else {
        double quef = (double)i / fs;
        sl = sin(WO_PI * cf0 * quef) / (WO_PI * cf0 * quef);
        cl = (1.0 - 2.0 * q1) + 2.0 * q1 * cos(2.0 * WO_PI * quef * cf0);
      }
      re[i] = re[i] * sl * cl / fft_size;
      im[i] = 0.0;
    }
    irfft_unnorm(re, im, fft_size, wave);
    for (int i = 0; i <= half_fft; ++i) spectrogram[(size_t)f * (half_fft + 1) + i] = exp(wave[i]);
  }
  free(wave); free(re); free(im); free(window);
}

/* ------------------------------------------------------------------ D4C (d4c.cpp)
 * DECIDE: the randn()*1e-12 waveform dither is dropped (see CheapTrick note). */
static void d4c_windowed_waveform(const double *x, int x_length, int fs, double cf0, double pos, int blackman,
                                  double ratio, double *waveform) {
  int half = wo_matlab_round(ratio * fs / cf0 / 2.0);
  int origin = wo_matlab_round(pos * fs + 0.001);
  double w1 = 0, w2 = 0;
  double *window = (double *)malloc(sizeof(double) * (half * 2 + 1));
  for (int i = 0; i <= half * 2; ++i) {
    double position = (2.0 * (i - half) / ratio) / fs;
    if (blackman)
      window[i] = 0.42 + 0.5 * cos(WO_PI * position * cf0) + 0.08 * cos(WO_PI * position * cf0 * 2);
    else
      window[i] = 0.5 * cos(WO_PI * position * cf0) + 0.5;
    int idx = imin(x_length - 1, imax(0, origin + i - half));
    waveform[i] = x[idx] * window[i];
    w1 += waveform[i]; w2 += window[i];
  }
  double coef = w1 / w2;
  for (int i = 0; i <= half * 2; ++i) waveform[i] -= window[i] * coef;
  free(window);
}

static void d4c_centroid(const double *x, int x_length, int fs, double cf0, int fft_size, double pos,
                         double *wave, double *re, double *im, double *centroid) {
  for (int i = 0; i < fft_size; ++i) wave[i] = 0.0;
  d4c_windowed_waveform(x, x_length, fs, cf0, pos, 1, 4.0, wave);
  int lim = wo_matlab_round(2.0 * fs / cf0) * 2;
  double power = 0.0;
  for (int i = 0; i <= lim; ++i) power += wave[i] * wave[i];
  for (int i = 0; i <= lim; ++i) wave[i] /= sqrt(power);
  rfft(wave, fft_size, re, im);
  double *tr = (double *)malloc(sizeof(double) * (fft_size / 2 + 1)), *ti = (double *)malloc(sizeof(double) * (fft_size / 2 + 1));
  for (int i = 0; i <= fft_size / 2; ++i) { tr[i] = re[i]; ti[i] = im[i]; }
  for (int i = 0; i < fft_size; ++i) wave[i] *= i + 1.0;
  rfft(wave, fft_size, re, im);
  for (int i = 0; i <= fft_size / 2; ++i) centroid[i] = re[i] * tr[i] + ti[i] * im[i];
  free(tr); free(ti);
}

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

void wo_d4c(const double *x, int x_length, int fs, const double *tpos, const double *f0, int f0_length,
            int fft_size, double threshold, double *aperiodicity, double *out_ap0) {
  int nb = fft_size / 2 + 1;
  for (size_t i = 0; i < (size_t)f0_length * nb; ++i) aperiodicity[i] = 1.0 - WO_SAFE_MIN;
  int fft_d4c = (int)pow(2.0, 1.0 + (int)(log(4.0 * fs / WO_FLOOR_F0_D4C + 1) / WO_LOG2));
  int nap = (int)(dmin(WO_UPPER_LIMIT, fs / 2.0 - WO_FREQ_INTERVAL) / WO_FREQ_INTERVAL);
  int window_length = (int)(WO_FREQ_INTERVAL * fft_d4c / fs) * 2 + 1;
  double *window = (double *)malloc(sizeof(double) * window_length);
  nuttall_window(window_length, window);

  /* D4CLoveTrain */
  double *ap0 = (double *)malloc(sizeof(double) * f0_length);
  {
    double lowest_f0 = 40.0;
    int lt_fft = (int)pow(2.0, 1.0 + (int)(log(3.0 * fs / lowest_f0 + 1) / WO_LOG2));
    int b0 = (int)ceil(100.0 * lt_fft / fs), b1 = (int)ceil(4000.0 * lt_fft / fs), b2 = (int)ceil(7900.0 * lt_fft / fs);
    double *wave = (double *)malloc(sizeof(double) * lt_fft);
    double *re = (double *)malloc(sizeof(double) * lt_fft), *im = (double *)malloc(sizeof(double) * lt_fft);
    double *ps = (double *)malloc(sizeof(double) * lt_fft);
    for (int f = 0; f < f0_length; ++f) {
      if (f0[f] == 0.0) { ap0[f] = 0.0; continue; }
      double cf0 = dmax(f0[f], lowest_f0);
      int wl = wo_matlab_round(1.5 * fs / cf0) * 2 + 1;
      d4c_windowed_waveform(x, x_length, fs, cf0, tpos[f], 1, 3.0, wave);
      for (int i = wl; i < lt_fft; ++i) wave[i] = 0.0;
      rfft(wave, lt_fft, re, im);
      for (int i = 0; i <= b0; ++i) ps[i] = 0.0;
      for (int i = b0 + 1; i < lt_fft / 2 + 1; ++i) ps[i] = re[i] * re[i] + im[i] * im[i];
      for (int i = b0; i <= b2; ++i) ps[i] += ps[i - 1];
      ap0[f] = ps[b1] / ps[b2];
    }
    free(wave); free(re); free(im); free(ps);
  }
  if (out_ap0) memcpy(out_ap0, ap0, sizeof(double) * f0_length);

  double *coarse = (double *)malloc(sizeof(double) * (nap + 2));
  double *coarse_axis = (double *)malloc(sizeof(double) * (nap + 2));
  coarse[0] = -60.0;
  coarse[nap + 1] = -WO_SAFE_MIN;
  for (int i = 0; i <= nap; ++i) coarse_axis[i] = i * WO_FREQ_INTERVAL;
  coarse_axis[nap + 1] = fs / 2.0;
  double *freq_axis = (double *)malloc(sizeof(double) * nb);
  for (int i = 0; i < nb; ++i) freq_axis[i] = (double)i * fs / fft_size;

  int hb = fft_d4c / 2 + 1;
  double *wave = (double *)malloc(sizeof(double) * fft_d4c);
  double *re = (double *)malloc(sizeof(double) * fft_d4c), *im = (double *)malloc(sizeof(double) * fft_d4c);
  double *c1 = (double *)malloc(sizeof(double) * hb), *c2 = (double *)malloc(sizeof(double) * hb);
  double *sc = (double *)malloc(sizeof(double) * hb), *sps = (double *)malloc(sizeof(double) * hb);
  double *gd = (double *)malloc(sizeof(double) * hb), *sgd = (double *)malloc(sizeof(double) * hb);
  double *ps = (double *)malloc(sizeof(double) * hb);
  for (int f = 0; f < f0_length; ++f) {
    if (f0[f] == 0 || ap0[f] <= threshold) continue;
    double cf0 = dmax(WO_FLOOR_F0_D4C, f0[f]);
    /* static centroid */
    d4c_centroid(x, x_length, fs, cf0, fft_d4c, tpos[f] - 0.25 / cf0, wave, re, im, c1);
    d4c_centroid(x, x_length, fs, cf0, fft_d4c, tpos[f] + 0.25 / cf0, wave, re, im, c2);
    for (int i = 0; i < hb; ++i) sc[i] = c1[i] + c2[i];
    dc_correction(sc, cf0, fs, fft_d4c, sc);
    /* smoothed power spectrum */
    for (int i = 0; i < fft_d4c; ++i) wave[i] = 0.0;
    d4c_windowed_waveform(x, x_length, fs, cf0, tpos[f], 0, 4.0, wave);
    rfft(wave, fft_d4c, re, im);
    for (int i = 0; i < hb; ++i) sps[i] = re[i] * re[i] + im[i] * im[i];
    dc_correction(sps, cf0, fs, fft_d4c, sps);
    linear_smoothing(sps, cf0, fs, fft_d4c, sps);
    /* static group delay */
    for (int i = 0; i < hb; ++i) gd[i] = sc[i] / sps[i];
    linear_smoothing(gd, cf0 / 2.0, fs, fft_d4c, gd);
    linear_smoothing(gd, cf0, fs, fft_d4c, sgd);
    for (int i = 0; i < hb; ++i) gd[i] -= sgd[i];
    /* coarse aperiodicity */
    int boundary = wo_matlab_round(fft_d4c * 8.0 / window_length);
    int half_wl = window_length / 2;
    for (int i = 0; i < fft_d4c; ++i) wave[i] = 0.0;
    for (int b = 0; b < nap; ++b) {
      int center = (int)(WO_FREQ_INTERVAL * (b + 1) * fft_d4c / fs);
      for (int j = 0; j <= half_wl * 2; ++j) wave[j] = gd[center - half_wl + j] * window[j];
      rfft(wave, fft_d4c, re, im);
      for (int j = 0; j < hb; ++j) ps[j] = re[j] * re[j] + im[j] * im[j];
      qsort(ps, hb, sizeof(double), cmp_double);
      for (int j = 1; j < hb; ++j) ps[j] += ps[j - 1];
      double ca = 10 * log10(ps[fft_d4c / 2 - boundary - 1] / ps[fft_d4c / 2]);
      coarse[1 + b] = dmin(0.0, ca + (cf0 - 100) / 50.0);
    }
    double *out = aperiodicity + (size_t)f * nb;
    wo_interp1(coarse_axis, coarse, nap + 2, freq_axis, nb, out);
    for (int i = 0; i < nb; ++i) out[i] = pow(10.0, out[i] / 20.0);
  }
  free(window); free(ap0); free(coarse); free(coarse_axis); free(freq_axis);
  free(wave); free(re); free(im); free(c1); free(c2); free(sc); free(sps); free(gd); free(sgd); free(ps);
}

/* ------------------------------------------------------------------ SPTK freqt / pysptk sp2mc, mc2sp */
void wo_freqt(const double *c1, int m1, double *c2, int m2, double a) {
  double b = 1 - a * a;
  double *d = (double *)calloc(m2 + 1, sizeof(double)), *g = (double *)calloc(m2 + 1, sizeof(double));
  for (int i = -m1; i <= 0; i++) {
    if (0 <= m2) { d[0] = g[0]; g[0] = c1[-i] + a * d[0]; }
    if (1 <= m2) { d[1] = g[1]; g[1] = b * d[0] + a * d[1]; }
    for (int j = 2; j <= m2; j++) { d[j] = g[j]; g[j] = d[j - 1] + a * (d[j] - g[j - 1]); }
  }
  memcpy(c2, g, sizeof(double) * (m2 + 1));
  free(d); free(g);
}

/* sp: (T, fftlen/2+1) power spectrum -> mc: (T, order+1). pysptk.sp2mc: c = irfft(log sp); c[0]/=2; freqt(c, order, alpha)
 * DECIDE: the full fftlen-long mirrored cepstrum is passed to freqt (m1 = fftlen-1), as pysptk does. */
void wo_sp2mc(const double *sp, int T, int fftlen, int order, double alpha, double *mc) {
  int nb = fftlen / 2 + 1;
  double *re = (double *)malloc(sizeof(double) * fftlen), *im = (double *)malloc(sizeof(double) * fftlen);
  double *c = (double *)malloc(sizeof(double) * fftlen);
  for (int t = 0; t < T; ++t) {
    for (int i = 0; i < nb; ++i) { re[i] = log(sp[(size_t)t * nb + i]); im[i] = 0.0; }
    irfft_unnorm(re, im, fftlen, c);
    for (int i = 0; i < fftlen; ++i) c[i] /= fftlen;
    c[0] /= 2.0;
    wo_freqt(c, fftlen - 1, mc + (size_t)t * (order + 1), order, alpha);
  }
  free(re); free(im); free(c);
}

/* pysptk.mc2sp: c = freqt(mc, fftlen/2, -alpha); c[0]*=2; symmetric extend; exp(real(rfft)) */
void wo_mc2sp(const double *mc, int T, int order, double alpha, int fftlen, double *sp) {
  int nb = fftlen / 2 + 1;
  double *c = (double *)malloc(sizeof(double) * nb);
  double *sym = (double *)malloc(sizeof(double) * fftlen);
  double *re = (double *)malloc(sizeof(double) * fftlen), *im = (double *)malloc(sizeof(double) * fftlen);
  for (int t = 0; t < T; ++t) {
    wo_freqt(mc + (size_t)t * (order + 1), order, c, fftlen / 2, -alpha);
    c[0] *= 2.0;
    for (int i = 0; i < fftlen; ++i) sym[i] = 0.0;
    sym[0] = c[0];
    for (int i = 1; i < nb; ++i) { sym[i] = c[i]; sym[fftlen - i] = c[i]; }
    rfft(sym, fftlen, re, im);
    for (int i = 0; i < nb; ++i) sp[(size_t)t * nb + i] = exp(re[i]);
  }
  free(c); free(sym); free(re); free(im);
}

/* ------------------------------------------------------------------ silence gate support
 * librosa.feature.rms(frame_length, hop, center=True, pad_mode='reflect')**2, fp64 accumulation
 * (DECIDE: fp64 instead of librosa's float32 so that the threshold decision is reproducible). */
void wo_frame_mse(const float *wave, int n, int frame_length, int hop, double *mse, int n_frames) {
  int pad = frame_length / 2;
  for (int f = 0; f < n_frames; ++f) {
    double acc = 0.0;
    for (int j = 0; j < frame_length; ++j) {
      int idx = f * hop + j - pad;
      if (idx < 0) idx = -idx;
      if (idx >= n) idx = 2 * (n - 1) - idx;
      if (idx < 0) idx = 0;           /* degenerate tiny inputs */
      if (idx >= n) idx = n - 1;
      double v = n > 0 ? (double)wave[idx] : 0.0;
      acc += v * v;
    }
    mse[f] = acc / frame_length;
  }
}

/* ------------------------------------------------------------------ WORLD realtime synthesizer
 * Restates synthesisrealtime.cpp (InitializeSynthesizer / AddParameters / Synthesis2) with flat
 * (global-index) bookkeeping instead of the ring of pointers: frame g lives at frames[g % cap],
 * pulse p at pulses[p % cap].  Behavioural contract kept: cumulative_frame starts at -1, the
 * hand-off (f0, phase) between AddParameters calls, pulse detection on |d fmod(phase,2pi)| > pi,
 * block emission iff synthesized_sample + buffer_size < last_location, noise_size = distance to
 * the next pulse, OLA at index - synthesized_sample - fft/2 + 1.
 * DECIDE: dc_remover = normalised Hanning over fft/2 taps, first half of the periodic response zeroed;
 *         frame indices beyond the newest frame clamp to it; AddParameters returns 0 (and does
 *         nothing) when more than `ring_frames` frames would be pending;
 *         the output buffer is shifted left by buffer_size and zero-filled at the tail;
 *         randn stream: one canonical xorshift128 stream per synthesizer, seeded at creation and
 *         addressed by ABSOLUTE SAMPLE POSITION: the pulse at sample index q uses draws
 *         [max(q,0), max(q,0)+noise_size).  Because noise_size = next_index - index this is
 *         WORLD's sequential consumption shifted by the first pulse's index (draws before it
 *         are discarded); it makes the noise data-independent and randomly addressable on GPU. */
typedef struct {
  int fs, fft_size, buffer_size;
  double frame_period;      /* seconds */
  int cap_frames;           /* frame ring capacity */
  double *f0;               /* [cap_frames] */
  float *sp, *ap;           /* [cap_frames][nb]  (features cross the boundary as float32: SURVEY A.9) */
  long long cumulative_frame;   /* index of newest frame, -1 initially */
  int handoff;
  double handoff_phase, handoff_f0;
  long long last_location;
  long long synthesized_sample;
  /* pulses */
  int cap_pulses;
  long long *p_index; double *p_time; int *p_vuv;
  long long n_pulses, next_pulse;
  double *buffer;           /* [2*buffer_size + fft_size] */
  double *dc_remover;       /* [fft/2] */
  wo_rng rng;
  long long rng_pos;        /* stream position of the next draw */
  int mode;                 /* WO_CANON_* bits: 0 = the GPU-friendly DECIDE 10 / 11 variants (default), see wo_synth_set_mode */
} wo_synth;

void *wo_synth_create(int fs, double frame_period_ms, int fft_size, int buffer_size, int ring_frames) {
  wo_synth *s = (wo_synth *)calloc(1, sizeof(wo_synth));
  int nb = fft_size / 2 + 1;
  s->fs = fs; s->fft_size = fft_size; s->buffer_size = buffer_size;
  s->frame_period = frame_period_ms / 1000.0;
  s->cap_frames = ring_frames;
  s->f0 = (double *)calloc(ring_frames, sizeof(double));
  s->sp = (float *)calloc((size_t)ring_frames * nb, sizeof(float));
  s->ap = (float *)calloc((size_t)ring_frames * nb, sizeof(float));
  s->cumulative_frame = -1;
  s->handoff = 0; s->handoff_phase = 0.0; s->handoff_f0 = 0.0;
  s->last_location = 0; s->synthesized_sample = 0;
  s->cap_pulses = 1 << 16;
  s->p_index = (long long *)calloc(s->cap_pulses, sizeof(long long));
  s->p_time = (double *)calloc(s->cap_pulses, sizeof(double));
  s->p_vuv = (int *)calloc(s->cap_pulses, sizeof(int));
  s->buffer = (double *)calloc(2 * buffer_size + fft_size, sizeof(double));
  s->dc_remover = (double *)calloc(fft_size / 2, sizeof(double));
  double dc = 0.0;
  for (int i = 0; i < fft_size / 2; ++i) {
    s->dc_remover[i] = 0.5 - 0.5 * cos(2.0 * WO_PI * (i + 1.0) / (1.0 + fft_size / 2));
    dc += s->dc_remover[i];
  }
  for (int i = 0; i < fft_size / 2; ++i) s->dc_remover[i] /= dc;
  wo_rng_seed(&s->rng);
  return s;
}

/* Canonical-WORLD switches, used ONLY to quantify how far DECIDE 10 / 11 move the oracle from synthesisrealtime.cpp
 * (tests/test_oracle_canonical.py):
 *   bit 0 (WO_CANON_RANDN): randn() consumed sequentially, noise_size draws per pulse and nothing skipped, as GetNoiseSpectrum does
 *                           (instead of addressing the stream by the pulse's absolute sample position);
 *   bit 1 (WO_CANON_PHASE): total phase as ONE running sum over the samples (instead of the fixed 256-sample blocked order). */
#define WO_CANON_RANDN 1
#define WO_CANON_PHASE 2
void wo_synth_set_mode(void *h, int mode) { ((wo_synth *)h)->mode = mode; }
/* discard n draws (tests: the position-addressed stream == the sequential stream advanced to the first pulse's sample index) */
void wo_synth_skip_randn(void *h, long long n) { wo_synth *s = (wo_synth *)h; for (long long i = 0; i < n; ++i) wo_randn(&s->rng); }

void wo_synth_destroy(void *h) {
  wo_synth *s = (wo_synth *)h;
  free(s->f0); free(s->sp); free(s->ap); free(s->p_index); free(s->p_time); free(s->p_vuv);
  free(s->buffer); free(s->dc_remover); free(s);
}

/* f0: [n] double; sp, ap: [n][nb] float */
int wo_synth_add(void *h, const double *f0, int n, const float *sp, const float *ap) {
  wo_synth *s = (wo_synth *)h;
  int nb = s->fft_size / 2 + 1;
  if (n <= 0) return 1;
  /* frames still needed: from floor(synthesized_sample / hop) - 1 */
  long long oldest_needed = (long long)(s->synthesized_sample / (s->frame_period * s->fs)) - 1;
  if (oldest_needed < 0) oldest_needed = 0;
  if (s->cumulative_frame + n - oldest_needed + 1 > s->cap_frames) return 0;   /* ring full */
  for (int i = 0; i < n; ++i) {
    long long g = s->cumulative_frame + 1 + i;
    int slot = (int)(g % s->cap_frames);
    s->f0[slot] = f0[i];
    memcpy(s->sp + (size_t)slot * nb, sp + (size_t)i * nb, sizeof(float) * nb);
    memcpy(s->ap + (size_t)slot * nb, ap + (size_t)i * nb, sizeof(float) * nb);
  }
  s->cumulative_frame += n;
  if (s->cumulative_frame < 1) {     /* first-ever single frame */
    s->handoff_f0 = f0[n - 1];
    s->handoff = 1;
    return 1;
  }
  long long first_frame = s->cumulative_frame - n;   /* may be -1 on the first call */
  long long start_sample = (long long)ceil((double)(first_frame) * s->frame_period * s->fs);
  if (start_sample < 0) start_sample = 0;
  long long end_sample = (long long)ceil((double)s->cumulative_frame * s->frame_period * s->fs);
  int ns = (int)(end_sample - start_sample);
  int hf = s->handoff;
  /* coarse axes */
  int nc = n + hf;
  double *ct = (double *)malloc(sizeof(double) * nc), *cf = (double *)malloc(sizeof(double) * nc), *cv = (double *)malloc(sizeof(double) * nc);
  long long cum0 = first_frame < 0 ? 0 : first_frame;
  ct[0] = cum0 * s->frame_period; cf[0] = s->handoff_f0; cv[0] = s->handoff_f0 == 0 ? 0.0 : 1.0;
  for (int i = 0; i < n; ++i) {
    ct[i + hf] = (double)(i + cum0 + hf) * s->frame_period;
    cf[i + hf] = f0[i];
    cv[i + hf] = f0[i] == 0.0 ? 0.0 : 1.0;
  }
  double *ta = (double *)malloc(sizeof(double) * ns), *if0 = (double *)malloc(sizeof(double) * ns), *ivuv = (double *)malloc(sizeof(double) * ns);
  for (int i = 0; i < ns; ++i) ta[i] = (double)(i + start_sample) / (double)s->fs;
  wo_interp1(ct, cf, nc, ta, ns, if0);
  wo_interp1(ct, cv, nc, ta, ns, ivuv);
  for (int i = 0; i < ns; ++i) {
    ivuv[i] = ivuv[i] > 0.5 ? 1.0 : 0.0;
    if0[i] = ivuv[i] == 0.0 ? WO_DEFAULT_F0 : if0[i];
  }
  /* pulse locations */
  int np_ = ns + hf;
  double *tp = (double *)malloc(sizeof(double) * (np_ + 1)), *wp = (double *)malloc(sizeof(double) * (np_ + 1));
  /* DECIDE: total phase = hand-off phase + prefix sum of the per-sample increments, summed in a FIXED
   * blocked order (256-sample blocks summed left to right, then block totals left to right) instead of
   * WORLD's single running sum.  The two differ only in the last bits, but the unvoiced default f0 (500 Hz
   * at 24 kHz) puts every pulse exactly on a 2*pi multiple, where those last bits decide the pulse's
   * sample; a fixed order makes the CPU and the parallel GPU scan bit-identical. */
  {
    double tp0 = hf == 1 ? s->handoff_phase : 2.0 * WO_PI * if0[0] / s->fs;
    const int BLK = 256;
    double base = tp0;
    if (s->mode & WO_CANON_PHASE) {      /* WORLD: total_phase[i] = total_phase[i - 1] + 2 pi f0 / fs, one running sum from the hand-off phase */
      tp[0] = tp0;
      for (int i = 1; i < np_; ++i) tp[i] = tp[i - 1] + 2.0 * WO_PI * if0[i - hf] / s->fs;
    } else
    for (int b0 = 0; b0 < np_; b0 += BLK) {
      int b1 = b0 + BLK < np_ ? b0 + BLK : np_;
      double local = 0.0;
      for (int i = b0; i < b1; ++i) {
        double inc = i == 0 ? 0.0 : 2.0 * WO_PI * if0[i - hf] / s->fs;
        local = local + inc;
        tp[i] = base + local;
      }
      base = base + local;
    }
  }
  s->handoff_phase = tp[np_ - 1];
  for (int i = 0; i < np_; ++i) wp[i] = fmod(tp[i], 2.0 * WO_PI);
  for (int i = 0; i < np_ - 1; ++i) {
    if (fabs(wp[i + 1] - wp[i]) > WO_PI) {
      double t = ta[i] - (double)hf / s->fs;
      long long idx = wo_matlab_round(t * s->fs);
      int slot = (int)(s->n_pulses % s->cap_pulses);
      s->p_time[slot] = t; s->p_index[slot] = idx;
      /* vuv at the pulse: interpolated_vuv indexed by the pulse's sample (clamped to this block) */
      long long li = idx - start_sample;
      if (li < 0) li = 0;
      if (li >= ns) li = ns - 1;
      s->p_vuv[slot] = ivuv[li] > 0.5 ? 1 : 0;
      s->n_pulses++;
      s->last_location = idx;
    }
  }
  s->handoff_f0 = f0[n - 1];
  s->handoff = 1;
  free(ct); free(cf); free(cv); free(ta); free(if0); free(ivuv); free(tp); free(wp);
  return 1;
}

static void min_phase(const double *logspec_half, int n, double *re, double *im) {
  /* logspec_half: bins 0..n/2. output re/im bins 0..n/2 of the minimum-phase spectrum. */
  for (int i = 0; i <= n / 2; ++i) { re[i] = logspec_half[i]; im[i] = 0.0; }
  for (int i = n / 2 + 1; i < n; ++i) { re[i] = logspec_half[n - i]; im[i] = 0.0; }
  wo_fft(re, im, n, -1);
  /* cepstrum is real (input real & even): fold */
  im[0] = 0.0;
  for (int i = 1; i < n / 2; ++i) { re[i] *= 2.0; im[i] = 0.0; }
  im[n / 2] = 0.0;
  for (int i = n / 2 + 1; i < n; ++i) { re[i] = 0.0; im[i] = 0.0; }
  wo_fft(re, im, n, -1);
  for (int i = 0; i <= n / 2; ++i) {
    double tmp = exp(re[i] / n);
    double ph = im[i] / n;
    re[i] = tmp * cos(ph);
    im[i] = tmp * sin(ph);
  }
}

static double safe_ap(double x) { return dmax(0.001, dmin(0.999999999999, x)); }

/* one pulse -> impulse response [fft_size] */
static void synth_one_pulse(wo_synth *s, long long p, int noise_size, double *response) {
  long long qpos = s->p_index[p % s->cap_pulses] < 0 ? 0 : s->p_index[p % s->cap_pulses];
  if (!(s->mode & WO_CANON_RANDN)) while (s->rng_pos < qpos) { wo_randn(&s->rng); s->rng_pos++; }
  int n = s->fft_size, nb = n / 2 + 1;
  int slot = (int)(p % s->cap_pulses);
  double t = s->p_time[slot];
  int vuv = s->p_vuv[slot];
  long long fl = (long long)(t / s->frame_period);
  long long ce = (long long)ceil(t / s->frame_period);
  double interp = t / s->frame_period - fl;
  if (fl > s->cumulative_frame) fl = s->cumulative_frame;
  if (ce > s->cumulative_frame) ce = s->cumulative_frame;
  const float *sp0 = s->sp + (size_t)(fl % s->cap_frames) * nb, *sp1 = s->sp + (size_t)(ce % s->cap_frames) * nb;
  const float *ap0 = s->ap + (size_t)(fl % s->cap_frames) * nb, *ap1 = s->ap + (size_t)(ce % s->cap_frames) * nb;
  double *spec = (double *)malloc(sizeof(double) * nb), *apr = (double *)malloc(sizeof(double) * nb);
  double *lg = (double *)malloc(sizeof(double) * nb);
  double *re = (double *)malloc(sizeof(double) * n), *im = (double *)malloc(sizeof(double) * n);
  double *nr = (double *)malloc(sizeof(double) * n), *ni = (double *)malloc(sizeof(double) * n);
  double *periodic = (double *)malloc(sizeof(double) * n), *aperiodic = (double *)malloc(sizeof(double) * n);
  double *tmp = (double *)malloc(sizeof(double) * n);
  for (int i = 0; i < nb; ++i) {
    if (fl == ce) {
      spec[i] = fabs((double)sp0[i]);
      apr[i] = pow(safe_ap((double)ap0[i]), 2.0);
    } else {
      spec[i] = (1.0 - interp) * fabs((double)sp0[i]) + interp * fabs((double)sp1[i]);
      apr[i] = pow((1.0 - interp) * safe_ap((double)ap0[i]) + interp * safe_ap((double)ap1[i]), 2.0);
    }
  }
  /* periodic */
  if (vuv == 0 || apr[0] > 0.999) {
    for (int i = 0; i < n; ++i) periodic[i] = 0.0;
  } else {
    for (int i = 0; i < nb; ++i) lg[i] = log(spec[i] * (1.0 - apr[i]) + WO_SAFE_MIN) / 2.0;
    min_phase(lg, n, re, im);
    irfft_unnorm(re, im, n, tmp);
    for (int i = 0; i < n / 2; ++i) { periodic[i] = tmp[i + n / 2]; periodic[i + n / 2] = tmp[i]; }
    double dc = 0.0;
    for (int i = n / 2; i < n; ++i) dc += periodic[i];
    for (int i = 0; i < n / 2; ++i) periodic[i] = 0.0;
    for (int i = n / 2; i < n; ++i) periodic[i] -= dc * s->dc_remover[i - n / 2];
  }
  /* aperiodic */
  {
    double avg = 0.0;
    for (int i = 0; i < noise_size; ++i) { tmp[i] = wo_randn(&s->rng); avg += tmp[i]; }
    s->rng_pos += noise_size;
    avg /= noise_size;
    for (int i = 0; i < noise_size; ++i) tmp[i] -= avg;
    for (int i = noise_size; i < n; ++i) tmp[i] = 0.0;
    rfft(tmp, n, nr, ni);
    if (vuv != 0) for (int i = 0; i < nb; ++i) lg[i] = log(spec[i] * apr[i]) / 2.0;
    else for (int i = 0; i < nb; ++i) lg[i] = log(spec[i]) / 2.0;
    min_phase(lg, n, re, im);
    for (int i = 0; i < nb; ++i) {
      double a = re[i] * nr[i] - im[i] * ni[i];
      double b = re[i] * ni[i] + im[i] * nr[i];
      re[i] = a; im[i] = b;
    }
    irfft_unnorm(re, im, n, tmp);
    for (int i = 0; i < n / 2; ++i) { aperiodic[i] = tmp[i + n / 2]; aperiodic[i + n / 2] = tmp[i]; }
  }
  double sq = sqrt((double)noise_size);
  for (int i = 0; i < n; ++i) response[i] = (periodic[i] * sq + aperiodic[i]) / n;
  free(spec); free(apr); free(lg); free(re); free(im); free(nr); free(ni); free(periodic); free(aperiodic); free(tmp);
}

/* returns 1 and writes buffer_size samples to out when a block could be produced, else 0 */
int wo_synth_synthesis2(void *h, double *out) {
  wo_synth *s = (wo_synth *)h;
  int B = s->buffer_size, n = s->fft_size;
  if (s->n_pulses == 0) return 0;
  if (s->synthesized_sample + B >= s->last_location) return 0;
  int total = 2 * B + n;
  for (int i = 0; i < total - B; ++i) s->buffer[i] = s->buffer[i + B];
  for (int i = total - B; i < total; ++i) s->buffer[i] = 0.0;
  double *resp = (double *)malloc(sizeof(double) * n);
  while (s->next_pulse < s->n_pulses) {
    long long cur = s->p_index[s->next_pulse % s->cap_pulses];
    if (cur >= s->synthesized_sample + B) break;
    long long nxt = s->p_index[(s->next_pulse + 1) % s->cap_pulses];   /* exists: cur < block end < last_location */
    int noise_size = (int)(nxt - cur);
    if (noise_size < 1) noise_size = 1;
    if (noise_size > n) noise_size = n;
    synth_one_pulse(s, s->next_pulse, noise_size, resp);
    long long offset = cur - s->synthesized_sample - n / 2 + 1;
    int index = (int)(offset < 0 ? -offset : 0);
    for (int i = index; i < n; ++i) s->buffer[i + offset] += resp[i];
    s->next_pulse++;
  }
  free(resp);
  s->synthesized_sample += B;
  for (int i = 0; i < B; ++i) out[i] = s->buffer[i];
  return 1;
}

long long wo_synth_pulse_count(void *h) { return ((wo_synth *)h)->n_pulses; }
void wo_synth_get_pulses(void *h, long long first, int count, long long *index, double *time, int *vuv) {
  wo_synth *s = (wo_synth *)h;
  for (int i = 0; i < count; ++i) {
    int slot = (int)((first + i) % s->cap_pulses);
    index[i] = s->p_index[slot]; time[i] = s->p_time[slot]; vuv[i] = s->p_vuv[slot];
  }
}

/* ------------------------------------------------------------------ WORLD offline Synthesis()
 * realtime_voice_conversion/yukarin_wrapper/vocoder.py:50-62 calls
 * pyworld.synthesize(f0, spectrogram, aperiodicity, fs, frame_period) = WORLD synthesis.cpp
 * Synthesis() with y_length = (int)(f0_length * frame_period * fs / 1000) (pyworld.pyx).
 * Restated from WORLD v0.2.3+ (the version with the fractional pulse time shift):
 * GetTimeBase / GetTemporalParametersForTimeBase / GetPulseLocationsForTimeBase, GetDCRemover
 * (normalised Hanning over the whole fft_size), GetPeriodicResponse (+ linear-phase fractional
 * shift) / GetAperiodicResponse / GetOneFrameSegment, overlap-add at index - fft/2 + 1.
 * DECIDE (shared with the realtime synthesizer above): total phase summed in the fixed blocked
 * order (256-sample blocks), randn addressed by absolute sample position (pulse at sample q
 * uses draws [q, q + noise_size) of the canonical xorshift128 stream), features enter as float32.
 * [UPSTREAM-UNVERIFIED] like the rest of this file. */
int wo_synthesize_length(int f0_length, double frame_period_ms, int fs) {
  return (int)((double)f0_length * frame_period_ms * fs / 1000.0);
}

int wo_synthesize(const double *f0, int f0_length, const float *sp, const float *ap, int fft_size,
                  double frame_period_ms, int fs, int y_length, double *y,
                  int max_pulses, long long *pulse_index, double *pulse_shift, int *pulse_vuv) {
  const int n = fft_size, nb = n / 2 + 1;
  const double frame_period = frame_period_ms / 1000.0;
  const double lowest_f0 = (double)fs / fft_size + 1.0;
  for (int i = 0; i < y_length; ++i) y[i] = 0.0;
  if (y_length < 2 || f0_length < 1) return 0;
  /* coarse axes with one extrapolated point */
  int nc = f0_length + 1;
  double *ct = (double *)malloc(sizeof(double) * nc), *cf = (double *)malloc(sizeof(double) * nc), *cv = (double *)malloc(sizeof(double) * nc);
  for (int i = 0; i < f0_length; ++i) {
    ct[i] = i * frame_period;
    cf[i] = f0[i] < lowest_f0 ? 0.0 : f0[i];
    cv[i] = cf[i] == 0.0 ? 0.0 : 1.0;
  }
  ct[f0_length] = f0_length * frame_period;
  if (f0_length >= 2) {
    cf[f0_length] = cf[f0_length - 1] * 2.0 - cf[f0_length - 2];
    cv[f0_length] = cv[f0_length - 1] * 2.0 - cv[f0_length - 2];
  } else { cf[f0_length] = cf[0]; cv[f0_length] = cv[0]; }
  double *ta = (double *)malloc(sizeof(double) * y_length), *if0 = (double *)malloc(sizeof(double) * y_length), *ivuv = (double *)malloc(sizeof(double) * y_length);
  for (int i = 0; i < y_length; ++i) ta[i] = (double)i / (double)fs;
  wo_interp1(ct, cf, nc, ta, y_length, if0);
  wo_interp1(ct, cv, nc, ta, y_length, ivuv);
  for (int i = 0; i < y_length; ++i) {
    ivuv[i] = ivuv[i] > 0.5 ? 1.0 : 0.0;
    if0[i] = ivuv[i] == 0.0 ? WO_DEFAULT_F0 : if0[i];
  }
  /* total phase: blocked fixed-order prefix sum of 2 pi f0 / fs (inclusive: tp[0] = first increment) */
  double *tp = (double *)malloc(sizeof(double) * y_length), *wp = (double *)malloc(sizeof(double) * y_length);
  {
    const int BLK = 256;
    double base = 0.0;
    for (int b0 = 0; b0 < y_length; b0 += BLK) {
      int b1 = b0 + BLK < y_length ? b0 + BLK : y_length;
      double local = 0.0;
      for (int i = b0; i < b1; ++i) { local = local + 2.0 * WO_PI * if0[i] / fs; tp[i] = base + local; }
      base = base + local;
    }
  }
  for (int i = 0; i < y_length; ++i) wp[i] = fmod(tp[i], 2.0 * WO_PI);
  int np_ = 0;
  long long *pidx = (long long *)malloc(sizeof(long long) * y_length);
  double *pshift = (double *)malloc(sizeof(double) * y_length);
  for (int i = 0; i < y_length - 1; ++i) {
    if (fabs(wp[i + 1] - wp[i]) > WO_PI) {
      double y1 = wp[i] - 2.0 * WO_PI, y2 = wp[i + 1];
      double x = -y1 / (y2 - y1);
      pidx[np_] = i; pshift[np_] = x / fs; ++np_;
    }
  }
  /* dc remover over the whole fft_size */
  double *dcr = (double *)malloc(sizeof(double) * n);
  {
    double dc = 0.0;
    for (int i = 0; i < n / 2; ++i) {
      dcr[i] = 0.5 - 0.5 * cos(2.0 * WO_PI * (i + 1.0) / (1.0 + n));
      dcr[n - i - 1] = dcr[i];
      dc += dcr[i] * 2.0;
    }
    for (int i = 0; i < n / 2; ++i) { dcr[i] /= dc; dcr[n - i - 1] = dcr[i]; }
  }
  double *spec = (double *)malloc(sizeof(double) * nb), *apr = (double *)malloc(sizeof(double) * nb), *lg = (double *)malloc(sizeof(double) * nb);
  double *re = (double *)malloc(sizeof(double) * n), *im = (double *)malloc(sizeof(double) * n);
  double *nr = (double *)malloc(sizeof(double) * n), *ni = (double *)malloc(sizeof(double) * n);
  double *periodic = (double *)malloc(sizeof(double) * n), *aperiodic = (double *)malloc(sizeof(double) * n), *tmp = (double *)malloc(sizeof(double) * n);
  double *noise = (double *)malloc(sizeof(double) * n);
  wo_rng rng; wo_rng_seed(&rng);
  long long rng_pos = 0;
  for (int p = 0; p < np_; ++p) {
    long long idx = pidx[p];
    int noise_size = (int)(pidx[p + 1 < np_ ? p + 1 : np_ - 1] - idx);
    if (noise_size > n) noise_size = n;
    double t = ta[idx];
    int vuv = ivuv[idx] > 0.5 ? 1 : 0;
    if (p < max_pulses) { if (pulse_index) pulse_index[p] = idx; if (pulse_shift) pulse_shift[p] = pshift[p]; if (pulse_vuv) pulse_vuv[p] = vuv; }
    int fl = (int)floor(t / frame_period), ce = (int)ceil(t / frame_period);
    if (fl > f0_length - 1) fl = f0_length - 1;
    if (ce > f0_length - 1) ce = f0_length - 1;
    double interp = t / frame_period - fl;
    const float *sp0 = sp + (size_t)fl * nb, *sp1 = sp + (size_t)ce * nb, *ap0 = ap + (size_t)fl * nb, *ap1 = ap + (size_t)ce * nb;
    for (int i = 0; i < nb; ++i) {
      if (fl == ce) { spec[i] = fabs((double)sp0[i]); apr[i] = pow(safe_ap((double)ap0[i]), 2.0); }
      else {
        spec[i] = (1.0 - interp) * fabs((double)sp0[i]) + interp * fabs((double)sp1[i]);
        apr[i] = pow((1.0 - interp) * safe_ap((double)ap0[i]) + interp * safe_ap((double)ap1[i]), 2.0);
      }
    }
    if (vuv == 0 || apr[0] > 0.999) {
      for (int i = 0; i < n; ++i) periodic[i] = 0.0;
    } else {
      for (int i = 0; i < nb; ++i) lg[i] = log(spec[i] * (1.0 - apr[i]) + WO_SAFE_MIN) / 2.0;
      min_phase(lg, n, re, im);
      double coef = 2.0 * WO_PI * pshift[p] * fs / n;
      for (int i = 0; i < nb; ++i) {
        double r = re[i], m = im[i];
        double re2 = cos(coef * i), im2 = sqrt(1.0 - re2 * re2);
        re[i] = r * re2 + m * im2;
        im[i] = m * re2 - r * im2;
      }
      irfft_unnorm(re, im, n, tmp);
      for (int i = 0; i < n / 2; ++i) { periodic[i] = tmp[i + n / 2]; periodic[i + n / 2] = tmp[i]; }
      double dc = 0.0;
      for (int i = n / 2; i < n; ++i) dc += periodic[i];
      for (int i = 0; i < n / 2; ++i) periodic[i] = -dc * dcr[i];
      for (int i = n / 2; i < n; ++i) periodic[i] -= dc * dcr[i];
    }
    if (noise_size > 0) {
      while (rng_pos < idx) { wo_randn(&rng); rng_pos++; }      /* windows [idx, idx + noise_size) ascend and never overlap */
      for (int i = 0; i < noise_size; ++i) noise[i] = wo_randn(&rng);
      rng_pos += noise_size;
      double avg = 0.0;
      for (int i = 0; i < noise_size; ++i) avg += noise[i];
      avg /= noise_size;
      for (int i = 0; i < noise_size; ++i) tmp[i] = noise[i] - avg;
    }
    for (int i = noise_size; i < n; ++i) tmp[i] = 0.0;
    rfft(tmp, n, nr, ni);
    if (vuv != 0) for (int i = 0; i < nb; ++i) lg[i] = log(spec[i] * apr[i]) / 2.0;
    else for (int i = 0; i < nb; ++i) lg[i] = log(spec[i]) / 2.0;
    min_phase(lg, n, re, im);
    for (int i = 0; i < nb; ++i) {
      double a = re[i] * nr[i] - im[i] * ni[i], b = re[i] * ni[i] + im[i] * nr[i];
      re[i] = a; im[i] = b;
    }
    irfft_unnorm(re, im, n, tmp);
    for (int i = 0; i < n / 2; ++i) { aperiodic[i] = tmp[i + n / 2]; aperiodic[i + n / 2] = tmp[i]; }
    double sq = sqrt((double)noise_size);
    long long offset = idx - n / 2 + 1;
    int lower = (int)(offset < 0 ? -offset : 0);
    int upper = (int)(n < y_length - offset ? n : y_length - offset);
    for (int i = lower; i < upper; ++i) y[i + offset] += (periodic[i] * sq + aperiodic[i]) / n;
  }
  free(ct); free(cf); free(cv); free(ta); free(if0); free(ivuv); free(tp); free(wp); free(pidx); free(pshift); free(dcr);
  free(spec); free(apr); free(lg); free(re); free(im); free(nr); free(ni); free(periodic); free(aperiodic); free(tmp); free(noise);
  return np_;
}

/* ------------------------------------------------------------------ output silence gate
 * realtime_voice_conversion/worker/decode_worker.py:53-59:
 *   power = librosa.core.power_to_db(numpy.abs(librosa.stft(wave)) ** 2).mean();  drop if power < -threshold
 * librosa (0.6/0.7 era) defaults: n_fft 2048, hop 512, periodic Hann window, center=True with reflect
 * padding, frames = 1 + len // hop; power_to_db(ref=1, amin=1e-10, top_db=80):
 *   db = 10 log10(max(amin, S));  db = max(db, max(db) - top_db);  mean over bins and frames.
 * DECIDE: fp64 throughout (librosa stores the STFT as complex64) so the decision is reproducible. */
int wo_stft_frames(int n, int hop) { return 1 + n / hop; }

double wo_stft_power_db_mean(const double *wave, int n, int n_fft, int hop, double amin, double top_db) {
  int frames = wo_stft_frames(n, hop), nb = n_fft / 2 + 1, pad = n_fft / 2;
  double *win = (double *)malloc(sizeof(double) * n_fft), *buf = (double *)malloc(sizeof(double) * n_fft);
  double *re = (double *)malloc(sizeof(double) * n_fft), *im = (double *)malloc(sizeof(double) * n_fft);
  double *db = (double *)malloc(sizeof(double) * (size_t)frames * nb);
  for (int i = 0; i < n_fft; ++i) win[i] = 0.5 - 0.5 * cos(2.0 * WO_PI * i / n_fft);
  double mx = -1e300;
  for (int f = 0; f < frames; ++f) {
    for (int j = 0; j < n_fft; ++j) {
      int idx = f * hop + j - pad;
      if (idx < 0) idx = -idx;
      if (idx >= n) idx = 2 * (n - 1) - idx;
      if (idx < 0) idx = 0;
      if (idx >= n) idx = n - 1;
      buf[j] = wave[idx] * win[j];
    }
    rfft(buf, n_fft, re, im);
    for (int k = 0; k < nb; ++k) {
      double pw = re[k] * re[k] + im[k] * im[k];
      double d = 10.0 * log10(pw > amin ? pw : amin);
      db[(size_t)f * nb + k] = d;
      if (d > mx) mx = d;
    }
  }
  double acc = 0.0;
  for (size_t i = 0; i < (size_t)frames * nb; ++i) acc += db[i] > mx - top_db ? db[i] : mx - top_db;
  free(win); free(buf); free(re); free(im); free(db);
  return acc / ((double)frames * nb);
}
