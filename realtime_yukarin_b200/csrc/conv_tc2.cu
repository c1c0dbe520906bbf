// conv_tc2.cu -- CTA-pair (cta_group::2) tcgen05 implicit-GEMM convolution for the large stage-2 layers.
//
// Why: the big k4 layers of the stage-2 U-Net run at the L2 -> shared-memory operand-traffic limit with
// the one-CTA-per-tile kernel of conv_tc.cu (profiles/r01c_ncu_full_one_step.csv: 11-12 TB/s of l1tex<-xbar reads, the
// practical cap).  This kernel cuts the operand bytes per FLOP two ways:
//   1. CTA pairs.  A thread-block cluster of two CTAs (two SMs of one TPC) computes a 256-pixel x N tile with ONE
//      tcgen05.mma.cta_group::2 (UMMA M = 256): CTA r stages its own 128 pixels of A and only HALF of the weight tile
//      (N/2 rows of B); the tensor cores of both SMs read both halves.  Weight bytes per CTA halve.
//   2. Parity-class fusion for transposed convs.  A k4 s2 p1 transposed conv is four dense 2x2-tap convs, one per output
//      parity class (conv_tc.cu runs them as separate tiles).  The four classes read 9 distinct shifted input views
//      (offsets {-1,0,1}^2), each class 4 of them.  Here ONE CTA pair keeps G accumulator groups in TMEM (G = 4 classes of
//      N = 64 for the last decoder layer, G = 2 classes of N = 128 for the one before) and every staged A view feeds all
//      the (class, tap) pairs that use it: 9 A loads instead of 16 per channel chunk.
// Pipeline per CTA: two smem rings (A views: 16 KB stages; B half tiles: N/2 x 128 B stages), TMA producer thread, and --
// in the leader CTA only -- the MMA thread; `full` barriers live in the leader (both CTAs' TMA loads complete_tx on them
// through the .cta_group::2 form), `empty` barriers are signalled in both CTAs by tcgen05.commit ... .multicast::cluster.
// Epilogue as in conv_tc.cu (tcgen05.ld -> scale/shift/act -> fp16 -> swizzled staging in the idle A ring -> TMA store),
// each CTA for its own 128 pixels, once per accumulator group.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "conv.h"
#include "tc_ptx.cuh"

namespace ryk {

// shared::cluster address of this CTA-local shared-memory address in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// TMA loads of the pair kernel: data lands in THIS CTA's shared memory, the transaction bytes are counted on the LEADER's
// mbarrier (`leader_bar` = shared::cluster address of the barrier in CTA 0; .cta_group::2 allows the peer's barrier)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");      // non-.aligned forms: the role branches leave warps 4 / 5 divergent
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

#ifdef RYK_TC_TIMELINE
// diagnostics build only (RYK_NVCC_EXTRA=-DRYK_TC_TIMELINE): per-CTA phase timestamps, dumped by conv_tc2_run to RYK_TC_TIMELINE_FILE
constexpr int kTl2MaxCtas = 8192, kTl2Slots = 10;
__device__ unsigned long long g_tl2[kTl2MaxCtas * kTl2Slots];
__device__ __forceinline__ unsigned long long tl2_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TL2(slot) do { int c_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (c_ < kTl2MaxCtas) g_tl2[c_ * kTl2Slots + (slot)] = tl2_now(); } while (0)
#else
#define TL2(slot) do {} while (0)
#endif

// One staged A view and the (accumulator group, class-local tap) pairs it feeds.
struct Tc2View { int8_t dy, dx, npairs, group[4], tap[4]; int8_t pad_; };
struct Tc2Params {
  int transposed, B, Hout, Wout, Cout;
  int Hc, Wc, tile_w, tile_h, tiles_w, tiles_h;
  int chunks0, chunks1;            // 64-channel chunks of source 0 / 1
  int sh, sw, ph, pw;
  int n_groups;                    // accumulator groups per CTA pair (1, 2 or 4), N columns each
  int n_views[4];                  // per blockIdx.z
  Tc2View views[4][9];             // per blockIdx.z: conv: 16 > 9 is handled by the `conv16` flag below
  int8_t group_cls[4][4];          // per blockIdx.z: parity class (py * 2 + px) of each group; convs: 0
  int conv16;                      // plain conv: 16 views computed on the fly (dy = t / 4, dx = t % 4, one pair (group 0, tap t))
  int act;
  const float* scale; const float* shift;
};

template <int NG, int kSA, int kSB, int kTmemCols>
__global__ void __launch_bounds__(kTcThreads, 2)
k_conv_tc2(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Tc2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kABytes = kBlockM * kBlockK * 2;          // 16 KB: 128 pixels x 64 channels
  constexpr uint32_t kBBytes = (NG / 2) * kBlockK * 2;         // this CTA's half of an N-row weight tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kSA * kABytes;
  uint64_t* fullA = (uint64_t*)(smem_b + kSB * kBBytes);
  uint64_t* emptyA = fullA + kSA;
  uint64_t* fullB = emptyA + kSA;
  uint64_t* emptyB = fullB + kSB;
  uint64_t* tmem_full = emptyB + kSB;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_full + 1);
  float* s_scale = (float*)(((uintptr_t)(tmem_ptr_smem + 4) + 15) & ~(uintptr_t)15);
  float* s_shift = s_scale + NG;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0: leader (issues the MMAs), 1: peer
  pdl_trigger();
  if (threadIdx.x == 0) {
    TL2(0);
#ifdef RYK_TC_TIMELINE
    { unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); int c_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (c_ < kTl2MaxCtas) { g_tl2[c_ * kTl2Slots + 9] = sm; g_tl2[c_ * kTl2Slots + 8] = rank; } }
#endif
  }

  // tile coordinates: the pair owns pixel tiles blockIdx.x (even: leader, odd: peer) of the class-local output grid
  int mt = blockIdx.x;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int b = mt;                                     // >= B for the padding tile of an odd tile count: TMA zero-fills / clips
  const int n0 = blockIdx.y * NG;
  const int z = blockIdx.z;
  const int oy0 = th * p.tile_h, ox0 = tw * p.tile_w;
  const int chunks = p.chunks0 + p.chunks1;
  const int n_views = p.conv16 ? 16 : p.n_views[z];

  if (threadIdx.x == 128) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (p.chunks1 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (threadIdx.x == 160) {
    for (int i = 0; i < kSA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < kSB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp < 4) {
    for (int i = threadIdx.x; i < NG; i += 128) { s_scale[i] = __ldg(p.scale + n0 + i); s_shift[i] = __ldg(p.shift + n0 + i); }
  }
  if (warp == 4) {   // TMEM allocation for the pair: the same warp of both CTAs issues it
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                 // barriers of BOTH CTAs are initialised before any remote complete_tx / multicast arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) TL2(7);       // after the cluster barrier, before the grid dependency wait
  pdl_wait();
  if (threadIdx.x == 0) TL2(1);

  auto view_of = [&](int v, int& dy, int& dx, int& npairs) {
    if (p.conv16) { dy = v >> 2; dx = v & 3; npairs = 1; }
    else { const Tc2View& V = p.views[z][v]; dy = V.dy; dx = V.dx; npairs = V.npairs; }
  };

  if (warp == 4 && lane == 0) {
    // ===== TMA producer (both CTAs): own A view + own half of every B tile =====
    int ia = 0, ib = 0;
    for (int cc = 0; cc < chunks; ++cc) {
      for (int v = 0; v < n_views; ++v) {
        int dy, dx, npairs;
        view_of(v, dy, dx, npairs);
        {
          const int s = ia % kSA;
          mbar_wait(&emptyA[s], ((ia / kSA) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&fullA[s], 2 * kABytes);
          const int ix = p.transposed ? ox0 + dx : ox0 * p.sw + dx - p.pw;
          const int iy = p.transposed ? oy0 + dy : oy0 * p.sh + dy - p.ph;
          if (cc < p.chunks0) tma_load_4d_2sm(smem_a + s * kABytes, &tmA0, map_to_rank(smem_u32(&fullA[s]), 0), cc * kBlockK, ix, iy, b);
          else tma_load_4d_2sm(smem_a + s * kABytes, &tmA1, map_to_rank(smem_u32(&fullA[s]), 0), (cc - p.chunks0) * kBlockK, ix, iy, b);
          ++ia;
        }
        for (int q = 0; q < npairs; ++q) {
          int g, tap;
          if (p.conv16) { g = 0; tap = v; } else { g = p.views[z][v].group[q]; tap = p.views[z][v].tap[q]; }
          const int cls = p.conv16 ? 0 : p.group_cls[z][g];
          const int s = ib % kSB;
          mbar_wait(&emptyB[s], ((ib / kSB) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&fullB[s], 2 * kBBytes);
          tma_load_2d_2sm(smem_b + s * kBBytes, &tmB, map_to_rank(smem_u32(&fullB[s]), 0), (tap * chunks + cc) * kBlockK, cls * p.Cout + n0 + (int)rank * (NG / 2));
          ++ib;
        }
      }
    }
  } else if (warp == 5 && lane == 0 && rank == 0) {
    // ===== MMA issuer (leader CTA only): UMMA 256 x NG x 16, A / B halves from both CTAs' shared memory =====
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NG >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    uint32_t started = 0;              // bit g: accumulator group g already holds a partial sum
    int ia = 0, ib = 0;
    for (int cc = 0; cc < chunks; ++cc) {
      for (int v = 0; v < n_views; ++v) {
        int dy, dx, npairs;
        view_of(v, dy, dx, npairs);
        const int sa = ia % kSA;
        mbar_wait(&fullA[sa], (ia / kSA) & 1);
        if (ia == 0) TL2(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = make_sw128_desc(smem_u32(smem_a + sa * kABytes));
        for (int q = 0; q < npairs; ++q) {
          const int g = p.conv16 ? 0 : p.views[z][v].group[q];
          const int sb = ib % kSB;
          mbar_wait(&fullB[sb], (ib / kSB) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t bdesc = make_sw128_desc(smem_u32(smem_b + sb * kBBytes));
          const uint32_t tmem_d = tmem_base + (uint32_t)(g * NG);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_f16_2sm(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (((started >> g) & 1u) || k > 0) ? 1u : 0u);
          started |= 1u << g;
          umma_commit_2sm(&emptyB[sb]);
          ++ib;
        }
        umma_commit_2sm(&emptyA[sa]);
        ++ia;
      }
    }
    umma_commit_2sm(tmem_full);
    TL2(3);
  } else if (warp < 4) {
    // ===== epilogue (both CTAs, own 128 pixels): one pass per accumulator group =====
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 0) TL2(4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int g = 0; g < p.n_groups; ++g) {
      uint8_t* stage = smem_a + (size_t)g * (NG / 64) * kABytes;         // NG / 64 blocks of [128 pixels][64 channels] fp16
#pragma unroll 1
      for (int c0 = 0; c0 < NG; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * NG + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint8_t* blk = stage + (c0 >> 6) * kABytes + row * 128;
        const int cbase = (c0 & 32) >> 3;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + c0 + j), sc1 = *reinterpret_cast<const float4*>(s_scale + c0 + j + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + c0 + j), sh1 = *reinterpret_cast<const float4*>(s_shift + c0 + j + 4);
          const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
          const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v0 = fmaf(__uint_as_float(r[j + 2 * q]), sc[2 * q], sh[2 * q]);
            float v1 = fmaf(__uint_as_float(r[j + 2 * q + 1]), sc[2 * q + 1], sh[2 * q + 1]);
            if (p.act == ACT_LEAKY) { v0 = v0 > 0.f ? v0 : 0.2f * v0; v1 = v1 > 0.f ? v1 : 0.2f * v1; }
            else if (p.act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            __half2 h = __floats2half2_rn(v0, v1);
            pk[q] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int chunk = cbase + (j >> 3);
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0 && b < p.B) {
      for (int g = 0; g < p.n_groups; ++g) {
        const int cls = p.conv16 ? 0 : p.group_cls[z][g];
        const int py = cls >> 1, px = cls & 1;
        const int xs = p.transposed ? ox0 * p.sw + px : ox0;
        const int ys = p.transposed ? oy0 * p.sh + py : oy0;
#pragma unroll
        for (int jb = 0; jb < NG / 64; ++jb) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(&tmO), "r"(smem_u32(smem_a + (size_t)(g * (NG / 64) + jb) * kABytes)), "r"(n0 + jb * 64), "r"(xs), "r"(ys), "r"(b) : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) TL2(5);
  // both CTAs are done with TMEM and with each other's shared memory / barriers before either one deallocates or exits
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    if (lane == 0) TL2(6);
  }
}

// ------------------------------------------------------------------------------------ host side
template <int NG, int kSA, int kSB> static constexpr size_t tc2_smem_bytes() {
  return (size_t)kSA * (kBlockM * kBlockK * 2) + (size_t)kSB * ((NG / 2) * kBlockK * 2) + (2 * kSA + 2 * kSB + 1) * 8 + 16 + 1024 + 2 * NG * 4 + 32;
}

// instantiations: (N per group, A stages, B stages, TMEM columns); RYK_TC2_DEEP=1 selects the deeper B rings (tuning)
#define TC2_LIST(X)                                                                                       \
  X(128, 4, 4, 128) X(128, 4, 6, 128) X(128, 4, 4, 256) X(128, 4, 6, 256) X(64, 4, 6, 256) X(64, 4, 12, 256) X(64, 4, 8, 128)

int tc2_init() {
#define TC2_ATTR(NG, SA, SB, TC) RYK_CUDA(cudaFuncSetAttribute(k_conv_tc2<NG, SA, SB, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes<NG, SA, SB>()));
  TC2_LIST(TC2_ATTR)
#undef TC2_ATTR
  return 0;
}

static int tc2_env(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

// RYK_TC2: 0 (default) = never, 1 = where the pair kernel fills the GPU, 2 = wherever the shape allows (unit tests).
// Off by default: at batch 1 the pair pipeline is latency-bound, the one-CTA kernel is faster.
static int tc2_mode() { return tc2_env("RYK_TC2", 0); }

// Picks (groups, N per group) for the pair kernel; false: the layer stays on the one-CTA-per-tile kernel of conv_tc.cu.
bool tc2_layer_config(const ConvLayer& L, int num_sms, int* groups, int* ng) {
  const int mode = tc2_mode();
  if (mode == 0) return false;
  const bool k2d = L.KH == 4 && L.KW == 4 && L.SH == 2 && L.SW == 2 && L.PH == 1 && L.PW == 1;
  if (!k2d || L.C0 % kBlockK != 0 || L.C1 % kBlockK != 0 || L.C0 == 0) return false;
  if (L.in_dtype != DT_F16 || L.out_dtype != DT_F16) return false;
  if (L.tile_w * L.tile_h != kBlockM) return false;
  const int Wc = L.transposed ? L.Win : L.Wout, Hc = L.transposed ? L.Hin : L.Hout;
  const int tiles_mn = L.B * ((Wc + L.tile_w - 1) / L.tile_w) * ((Hc + L.tile_h - 1) / L.tile_h);
  const int tiles_x = (tiles_mn + 1) & ~1;
  int G = 0, NG = 0, ctas = 0;
  if (!L.transposed) {
    if (L.Cout % 128 != 0) return false;
    G = 1; NG = 128; ctas = tiles_x * (L.Cout / 128);
  } else if (L.Cout == 64) { G = tc2_env("RYK_TC2_G64", 4) == 2 ? 2 : 4; NG = 64; ctas = tiles_x * (4 / G); }
  else if (L.Cout == 128) { G = 2; NG = 128; ctas = tiles_x * 2; }
  else if (L.Cout % 128 == 0) { G = 1; NG = 128; ctas = tiles_x * (L.Cout / 128) * 4; }
  else return false;
  if (mode == 1 && ctas < num_sms) return false;           // too few tiles: split-K on the other kernel spreads the weight stream better
  *groups = G; *ng = NG;
  return true;
}

int tc2_layer_prepare(ConvLayer& L, PFN_cuTensorMapEncodeTiled_v12000 encode) {
  const int classes = L.transposed ? 4 : 1;
  const int ntaps = L.transposed ? 4 : 16;
  cuuint64_t dims[2] = {(cuuint64_t)ntaps * (L.C0 + L.C1), (cuuint64_t)classes * L.Cout};
  cuuint64_t strides[1] = {dims[0] * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(L.tc2_ng / 2)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&L.tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(L.w_tc), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(pair-kernel weights) failed: " + std::to_string((int)r)); return -1; }
  return 0;
}

int conv_tc2_run(const ConvLayer& L, cudaStream_t st, bool pdl) {
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.transposed = L.transposed; p.B = L.B; p.Hout = L.Hout; p.Wout = L.Wout; p.Cout = L.Cout;
  p.Hc = L.transposed ? L.Hin : L.Hout; p.Wc = L.transposed ? L.Win : L.Wout;
  p.tile_w = L.tile_w; p.tile_h = L.tile_h;
  p.tiles_w = (p.Wc + L.tile_w - 1) / L.tile_w; p.tiles_h = (p.Hc + L.tile_h - 1) / L.tile_h;
  p.chunks0 = L.C0 / kBlockK; p.chunks1 = L.C1 / kBlockK;
  p.sh = L.SH; p.sw = L.SW; p.ph = L.PH; p.pw = L.PW;
  p.n_groups = L.tc2_groups;
  p.act = L.act; p.scale = L.scale; p.shift = L.shift;
  const int G = L.tc2_groups, NG = L.tc2_ng;
  int grid_z = 1;
  if (!L.transposed) {
    p.conv16 = 1;
  } else {
    // class (py, px), class-local tap (ty, tx) reads the input at offset (ty - 1 + py, tx - 1 + px) (k_pack_tc's layout:
    // tap index ty * 2 + tx, class index py * 2 + px).  Group the classes of one CTA pair and merge equal offsets into views.
    grid_z = 4 / G;
    for (int z = 0; z < grid_z; ++z) {
      int nv = 0;
      for (int g = 0; g < G; ++g) {
        const int cls = G == 4 ? g : (G == 2 ? z * 2 + g : z);         // G = 2: the two classes of one output-row parity
        p.group_cls[z][g] = (int8_t)cls;
        const int py = cls >> 1, px = cls & 1;
        for (int ty = 0; ty < 2; ++ty) for (int tx = 0; tx < 2; ++tx) {
          const int dy = ty - 1 + py, dx = tx - 1 + px;
          int v = 0;
          for (; v < nv; ++v) if (p.views[z][v].dy == dy && p.views[z][v].dx == dx) break;
          if (v == nv) { p.views[z][v].dy = (int8_t)dy; p.views[z][v].dx = (int8_t)dx; p.views[z][v].npairs = 0; ++nv; }
          Tc2View& V = p.views[z][v];
          V.group[V.npairs] = (int8_t)g; V.tap[V.npairs] = (int8_t)(ty * 2 + tx); V.npairs++;
        }
      }
      p.n_views[z] = nv;
    }
  }
  const int tiles_mn = L.B * p.tiles_w * p.tiles_h;
  dim3 grid((tiles_mn + 1) & ~1, G == 1 ? L.Cout / NG : 1, grid_z);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kTcThreads); cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  const bool deep = tc2_env("RYK_TC2_DEEP", 0) != 0;
#define TC2_LAUNCH(NGv, SA, SB, TC) do { cfg.dynamicSmemBytes = tc2_smem_bytes<NGv, SA, SB>(); \
    RYK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc2<NGv, SA, SB, TC>, L.tmA0, L.tmA1, L.tmB2, L.tmO, p)); } while (0)
  if (NG == 128 && G == 1) { if (deep) TC2_LAUNCH(128, 4, 6, 128); else TC2_LAUNCH(128, 4, 4, 128); }
  else if (NG == 128 && G == 2) { if (deep) TC2_LAUNCH(128, 4, 6, 256); else TC2_LAUNCH(128, 4, 4, 256); }
  else if (NG == 64 && G == 4) { if (deep) TC2_LAUNCH(64, 4, 12, 256); else TC2_LAUNCH(64, 4, 6, 256); }
  else if (NG == 64 && G == 2) TC2_LAUNCH(64, 4, 8, 128);
  else { set_error("pair kernel: unsupported (groups, N) combination"); return -1; }
#undef TC2_LAUNCH
  RYK_CUDA(cudaGetLastError());
#ifdef RYK_TC_TIMELINE
  cudaStreamCaptureStatus cap_ = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap_);
  if (const char* path = cap_ == cudaStreamCaptureStatusNone ? getenv("RYK_TC_TIMELINE_FILE") : nullptr) {
    cudaStreamSynchronize(st);
    static std::vector<unsigned long long> h(kTl2MaxCtas * kTl2Slots);
    cudaMemcpyFromSymbol(h.data(), g_tl2, sizeof(unsigned long long) * h.size());
    int n = grid.x * grid.y * grid.z; if (n > kTl2MaxCtas) n = kTl2MaxCtas;
    if (FILE* f = fopen(path, "w")) {
      fprintf(f, "# pair kernel grid %d %d %d groups %d ng %d\n", grid.x, grid.y, grid.z, G, NG);
      for (int c = 0; c < n; ++c) { for (int k = 0; k < kTl2Slots; ++k) fprintf(f, "%llu ", h[c * kTl2Slots + k]); fprintf(f, "\n"); }
      fclose(f);
    }
  }
#endif
  return 0;
}

}  // namespace ryk
