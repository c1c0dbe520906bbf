// CPU emulation of the fused stage-1 kernel's k4 layers (realtime_yukarin_b200/csrc/s1_fused.cu): the kernel's loops replayed CTA by
// CTA / warp by warp with software ldmatrix.x4 and mma.sync.m16n8k16 (PTX ISA fragment layouts), over the SAME index functions
// (csrc/s1_map.h) and the same fragment-packed weight array, checked against a direct (transposed) convolution in the model file's
// Chainer weight layout.  Test infrastructure: built and run by tests/test_s1_fused_map.py, no GPU involved.
//   usage: s1_fused_emulate <base> <tp1> <cluster size>      exit code 0 = every layer matches
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../realtime_yukarin_b200/csrc/s1_map.h"

using namespace ryk;

static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xffff) / 65536.0f - 0.5f; }

struct Warp {            // register state of 32 lanes
  float acc[32][2][2][4];
};

// ldmatrix.sync.aligned.m8n8.x4.b16: lane 8 i + r supplies the address of row r of matrix i; thread t receives, from matrix i,
// the two 16-bit elements (row t / 4, columns 2 (t % 4), 2 (t % 4) + 1).
static void ldmatrix_x4(const std::vector<float>& smem, const int (&addr)[32], float (&frag)[32][4][2]) {
  for (int t = 0; t < 32; ++t)
    for (int i = 0; i < 4; ++i) {
      const int a = addr[8 * i + t / 4];
      if (a < 0 || a + 8 > (int)smem.size()) { fprintf(stderr, "ldmatrix address %d outside shared memory\n", a); exit(2); }
      frag[t][i][0] = smem[a + 2 * (t % 4)];
      frag[t][i][1] = smem[a + 2 * (t % 4) + 1];
    }
}
// mma.sync.aligned.m16n8k16.row.col: A regs a0 (row g, k 2q..), a1 (row g + 8), a2 (row g, k + 8), a3 (row g + 8, k + 8);
// B regs b0 (k 2q.., n g), b1 (k + 8); C regs c0, c1 (row g, n 2q, 2q + 1), c2, c3 (row g + 8) with g = t / 4, q = t % 4.
static void mma_16816(float (&c)[32][4], const float (&a)[32][4][2], const float (&b)[32][2][2]) {
  float A[16][16], B[16][8];
  for (int t = 0; t < 32; ++t) {
    const int g = t / 4, q = t % 4;
    for (int e = 0; e < 2; ++e) {
      A[g][2 * q + e] = a[t][0][e]; A[g + 8][2 * q + e] = a[t][1][e]; A[g][2 * q + 8 + e] = a[t][2][e]; A[g + 8][2 * q + 8 + e] = a[t][3][e];
      B[2 * q + e][g] = b[t][0][e]; B[2 * q + 8 + e][g] = b[t][1][e];
    }
  }
  for (int t = 0; t < 32; ++t)
    for (int r = 0; r < 4; ++r) {
      const int row = s1_c_row(t, r), col = s1_c_col(t, r);
      float s = c[t][r];
      for (int k = 0; k < 16; ++k) s += A[row][k] * B[k][col];
      c[t][r] = s;
    }
}

struct Layer {
  S1Geom g; int C0, C1;
  std::vector<float> in0, in1, w_chainer, w_frag, out, ref;
};

static void reference(Layer& L) {
  const S1Geom& g = L.g;
  const int Wout = s1_Wout(g);
  L.ref.assign((size_t)Wout * g.Cout, 0.f);
  auto in = [&](int px, int c) { return c < L.C0 ? L.in0[(size_t)px * L.C0 + c] : L.in1[(size_t)px * L.C1 + c - L.C0]; };
  if (!g.transposed) {
    for (int o = 0; o < Wout; ++o)
      for (int n = 0; n < g.Cout; ++n) {
        double s = 0;
        for (int t = 0; t < 4; ++t) {
          const int px = 2 * o - 1 + t;
          if (px < 0 || px >= g.Win) continue;
          for (int c = 0; c < g.Cin; ++c) s += (double)in(px, c) * L.w_chainer[((size_t)n * g.Cin + c) * 4 + t];
        }
        L.ref[(size_t)o * g.Cout + n] = (float)s;
      }
  } else {
    std::vector<double> acc((size_t)Wout * g.Cout, 0.0);
    for (int i = 0; i < g.Win; ++i)
      for (int t = 0; t < 4; ++t) {
        const int o = 2 * i - 1 + t;
        if (o < 0 || o >= Wout) continue;
        for (int c = 0; c < g.Cin; ++c) {
          const double x = in(i, c);
          for (int n = 0; n < g.Cout; ++n) acc[(size_t)o * g.Cout + n] += x * L.w_chainer[((size_t)c * g.Cout + n) * 4 + t];
        }
      }
    for (size_t i = 0; i < acc.size(); ++i) L.ref[i] = (float)acc[i];
  }
}

static void pack(Layer& L) {
  const S1Geom& g = L.g;
  const int K = s1_K(g);
  L.w_frag.assign((size_t)s1_classes(g) * g.Cout * K, NAN);
  for (int cls = 0; cls < s1_classes(g); ++cls)
    for (int n = 0; n < g.Cout; ++n)
      for (int k = 0; k < K; ++k) {
        const size_t d = s1_w_dst(g, cls, n, k);
        if (d >= L.w_frag.size() || !isnan(L.w_frag[d])) { fprintf(stderr, "s1_w_dst is not a bijection at (%d,%d,%d)\n", cls, n, k); exit(2); }
        L.w_frag[d] = L.w_chainer[s1_w_src(g, cls, n, k)];
      }
}

// the kernel's s1_layer() for every CTA of the cluster
static void emulate(Layer& L, int nc) {
  const S1Geom& g = L.g;
  const S1Cut c = s1_cut(g, nc);
  const int M = s1_M(g), K = s1_K(g), KP = K / 32, NG = g.Cout / 16;
  L.out.assign((size_t)s1_Wout(g) * g.Cout, NAN);
  if (c.ms * c.ks > kS1Warps || c.MS * c.NP != nc || KP % c.ks || (KP / c.ks) % 2) { fprintf(stderr, "bad cut\n"); exit(2); }
  for (int rank = 0; rank < nc; ++rank) {
    const int mslab = rank % c.MS, np = rank / c.MS;
    const int m0 = mslab * c.slab, m1 = std::min(M, m0 + c.slab);
    if (m0 >= M) continue;
    const int px0 = s1_px0(g, m0), RS = c.RS;
    std::vector<float> act(kS1ActBytes / 2, NAN);
    const int nrows = s1_rows_for(g, m1 - m0), vpr = g.Cin / 8;
    if ((size_t)nrows * RS > act.size()) { fprintf(stderr, "staged rows exceed kS1ActBytes\n"); exit(2); }
    for (int i = 0; i < nrows * vpr; ++i) {
      const int r = i / vpr, ch = (i - r * vpr) * 8, px = px0 + r;
      for (int e = 0; e < 8; ++e) {
        float v = 0.f;
        if (px >= 0 && px < g.Win) v = ch < L.C0 ? L.in0[(size_t)px * L.C0 + ch + e] : L.in1[(size_t)px * L.C1 + (ch - L.C0) + e];
        act[(size_t)r * RS + ch + e] = v;
      }
    }
    const int kp_per = KP / c.ks;
    const int rows = m1 - m0, mt_slab = (rows + 15) / 16;
    const int npass = (mt_slab + 2 * c.ms - 1) / (2 * c.ms);
    const int ntasks = s1_tasks(g);
    std::vector<float> partial((size_t)kS1Warps * 16 * 32);
    for (int task = np; task < ntasks; task += c.NP) {
      const int cls = task / NG, ng = task - cls * NG;
      const float* wt = L.w_frag.data() + (size_t)task * s1_task_halfs(g);
      for (int p = 0; p < npass; ++p) {
        std::vector<Warp> W(kS1Warps);
        std::vector<char> work(kS1Warps, 0);
        for (int warp = 0; warp < kS1Warps; ++warp) {
          const int mg = warp / c.ks, kpart = warp - mg * c.ks;
          const bool active = warp < c.ms * c.ks;
          const int kp_lo = kpart * kp_per, kp_hi = kp_lo + kp_per;
          const int tA = mg + c.ms * (2 * p), tB = tA + c.ms;
          work[warp] = active && tA < mt_slab;
          memset(&W[warp], 0, sizeof(Warp));
          if (!work[warp]) continue;
          for (int kp = kp_lo; kp < kp_hi; ++kp)
            for (int kt = 0; kt < 2; ++kt) {
              const int k0 = kp * 32 + kt * 16;
              const int j = k0 / g.Cin;
              float aA[32][4][2], aB[32][4][2], b[2][32][2][2];
              int addrA[32], addrB[32];
              for (int lane = 0; lane < 32; ++lane) {
                const int lrow = s1_ldm_row(lane), kofs = s1_ldm_kofs(lane);
                const int mA = m0 + std::min(tA * 16 + lrow, rows - 1), mB = m0 + std::min(tB * 16 + lrow, rows - 1);
                const int ch = k0 - j * g.Cin + kofs;
                addrA[lane] = (s1_in_px(g, cls, mA, j) - px0) * RS + ch;
                addrB[lane] = (s1_in_px(g, cls, mB, j) - px0) * RS + ch;
                for (int nt = 0; nt < 2; ++nt) {
                  const float* v = wt + ((size_t)(kp * 2 + nt) * 32 + lane) * 8;      // the lane's uint4: regs x, y (k-tile 0), z, w (k-tile 1)
                  for (int e = 0; e < 2; ++e) { b[nt][lane][0][e] = v[(kt * 2 + 0) * 2 + e]; b[nt][lane][1][e] = v[(kt * 2 + 1) * 2 + e]; }
                }
              }
              ldmatrix_x4(act, addrA, aA);
              ldmatrix_x4(act, addrB, aB);
              float cA0[32][4], cA1[32][4], cB0[32][4], cB1[32][4];
              for (int t = 0; t < 32; ++t)
                for (int r = 0; r < 4; ++r) { cA0[t][r] = W[warp].acc[t][0][0][r]; cA1[t][r] = W[warp].acc[t][0][1][r]; cB0[t][r] = W[warp].acc[t][1][0][r]; cB1[t][r] = W[warp].acc[t][1][1][r]; }
              mma_16816(cA0, aA, b[0]); mma_16816(cA1, aA, b[1]); mma_16816(cB0, aB, b[0]); mma_16816(cB1, aB, b[1]);
              for (int t = 0; t < 32; ++t)
                for (int r = 0; r < 4; ++r) { W[warp].acc[t][0][0][r] = cA0[t][r]; W[warp].acc[t][0][1][r] = cA1[t][r]; W[warp].acc[t][1][0][r] = cB0[t][r]; W[warp].acc[t][1][1][r] = cB1[t][r]; }
            }
        }
        if (c.ks > 1) {
          for (int warp = 0; warp < kS1Warps; ++warp) {
            const int kpart = warp % c.ks;
            if (work[warp] && kpart > 0)
              for (int lane = 0; lane < 32; ++lane)
                for (int i = 0; i < 2; ++i) for (int nt = 0; nt < 2; ++nt) for (int r = 0; r < 4; ++r)
                  partial[((size_t)warp * 16 + (i * 2 + nt) * 4 + r) * 32 + lane] = W[warp].acc[lane][i][nt][r];
          }
          for (int warp = 0; warp < kS1Warps; ++warp) {
            const int mg = warp / c.ks, kpart = warp - mg * c.ks;
            if (work[warp] && kpart == 0)
              for (int kq = 1; kq < c.ks; ++kq) {
                const int w2 = mg * c.ks + kq;
                for (int lane = 0; lane < 32; ++lane)
                  for (int i = 0; i < 2; ++i) for (int nt = 0; nt < 2; ++nt) for (int r = 0; r < 4; ++r)
                    W[warp].acc[lane][i][nt][r] += partial[((size_t)w2 * 16 + (i * 2 + nt) * 4 + r) * 32 + lane];
              }
          }
        }
        for (int warp = 0; warp < kS1Warps; ++warp) {
          const int mg = warp / c.ks, kpart = warp - mg * c.ks;
          if (!(work[warp] && kpart == 0)) continue;
          const int tA = mg + c.ms * (2 * p), tB = tA + c.ms;
          for (int lane = 0; lane < 32; ++lane)
            for (int i = 0; i < 2; ++i) {
              const int tile = i ? tB : tA;
              for (int nt = 0; nt < 2; ++nt) {
                const int n = ng * 16 + nt * 8 + s1_c_col(lane, 0);
                for (int rr = 0; rr < 2; ++rr) {
                  const int mrel = tile * 16 + s1_c_row(lane, rr * 2);
                  if (mrel < rows) {
                    const int opx = s1_out_px(g, cls, m0 + mrel);
                    for (int e = 0; e < 2; ++e) {
                      float& o = L.out[(size_t)opx * g.Cout + n + e];
                      if (!isnan(o)) { fprintf(stderr, "output (%d,%d) written twice\n", opx, n + e); exit(2); }
                      o = W[warp].acc[lane][i][nt][rr * 2 + e];
                    }
                  }
                }
              }
            }
        }
      }
    }
  }
}

int main(int argc, char** argv) {
  const int base = argc > 1 ? atoi(argv[1]) : 64, tp1 = argc > 2 ? atoi(argv[2]) : 128, nc = argc > 3 ? atoi(argv[3]) : 16;
  static const int enc_mult[8] = {1, 2, 4, 8, 8, 8, 8, 8}, dec_mult[7] = {8, 8, 8, 8, 4, 2, 1};
  int bad = 0;
  for (int i = 1; i <= 14; ++i) {
    Layer L;
    if (i < 8) { L.g = S1Geom{0, tp1 >> (i - 1), base * enc_mult[i - 1], base * enc_mult[i]}; L.C0 = L.g.Cin; L.C1 = 0; }
    else {
      const int d = i - 8;
      L.C0 = d == 0 ? base * enc_mult[7] : base * dec_mult[d - 1];
      L.C1 = d == 0 ? 0 : base * enc_mult[7 - d];
      L.g = S1Geom{1, tp1 >> (7 - d), L.C0 + L.C1, base * dec_mult[d]};
    }
    const S1Geom& g = L.g;
    L.in0.resize((size_t)g.Win * L.C0); L.in1.resize((size_t)g.Win * L.C1);
    for (auto& v : L.in0) v = frand();
    for (auto& v : L.in1) v = frand();
    L.w_chainer.resize((size_t)g.Cin * g.Cout * 4);
    for (auto& v : L.w_chainer) v = frand();
    reference(L);
    pack(L);
    emulate(L, nc);
    double worst = 0; size_t missing = 0;
    for (size_t q = 0; q < L.ref.size(); ++q) {
      if (isnan(L.out[q])) { ++missing; continue; }
      worst = std::max(worst, (double)fabsf(L.out[q] - L.ref[q]));
    }
    const S1Cut c = s1_cut(g, nc);
    const double tol = 1e-3 * sqrt((double)s1_K(g));
    const bool ok = missing == 0 && worst < tol;
    printf("layer %2d %s Win %4d Cin %4d Cout %4d  MS %d slab %3d NP %2d ms %2d ks %2d  max |err| %.2e  missing %zu  %s\n", i, g.transposed ? "deconv" : "conv  ",
           g.Win, g.Cin, g.Cout, c.MS, c.slab, c.NP, c.ms, c.ks, worst, missing, ok ? "ok" : "FAIL");
    if (!ok) bad = 1;
  }
  return bad;
}
