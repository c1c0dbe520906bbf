// mma_rate_probe.cu -- measurement, not product code: what bounds a 1-CTA (cta_group::1) tcgen05 main loop on B200?
//   (1) tensor-pipe rate of kind::f16 128 x N x 16 MMAs issued back to back on operands ALREADY in shared memory (no refill),
//   (2) TMA load bandwidth per SM with every SM pulling L2-resident tiles and no MMA,
//   (3) both at once (the refill writes and the operand reads share the shared-memory port).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate_probe tools/mma_rate_probe.cu -lcuda
// Run  : gpurun_out/mma_rate_probe      (prints one table; DESIGN.md 4b quotes it)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../realtime_yukarin_b200/csrc/tc_ptx.cuh"

using namespace ryk;

struct ProbeOut { unsigned long long t_mma, t_tma; unsigned n_mma, n_tma; };

// mode bit 0: issue MMAs; bit 1: issue TMA loads (32 KB per "stage": 16 KB A box + 16 KB B tile) into a ring of `slots` stages
template <int N>
__global__ void __launch_bounds__(192, 1)
k_probe(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmA4, int mode, int iters, int iters_tma, int slots, int a_stride_bytes, ProbeOut* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ops = smem;                                   // 2 operand stages of 48 KB (A 16 KB + B up to 32 KB), never refilled
  uint8_t* ring = smem + ((mode & 1) ? 2 * 49152 : 0);   // TMA ring (independent of the MMAs)
  __shared__ uint64_t full[4][8];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode & 1) for (int i = threadIdx.x; i < 2 * 49152 / 4; i += blockDim.x) {
    // bit 10: random fp16 operands in (-1, 1) (a real workload's toggling) instead of the constant 1.0
    uint32_t v = 0x3c003c00u;
    if (mode & 1024) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83ff83ffu) | 0x38003800u;            // sign + 10 mantissa bits, exponent 14: |x| in [0.5, 1)
    }
    ((uint32_t*)ops)[i] = v;
  }
  if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) for (int i = 0; i < 8; ++i) mbar_init(&full[w][i], 1); mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 5 && lane == 0 && (mode & 1)) {
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t st = smem_u32(ops + (i & 1) * 49152);
      const uint64_t adesc = make_sw128_desc(st + (uint32_t)((i >> 2) & 1) * a_stride_bytes), bdesc = make_sw128_desc(st + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + (uint32_t)((i & 1) * N), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
    }
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    out[blockIdx.x].t_mma = clock64() - t0;
    out[blockIdx.x].n_mma = iters * 4;
  }
  const int issuers = ((mode >> 4) & 7) ? ((mode >> 4) & 7) : 1;      // TMA issuer threads (lane 0 of warps 0 .. issuers - 1), each with its own ring
  if (warp < issuers && lane == 0 && (mode & 2)) {
    unsigned long long t0 = clock64();
    const int total = iters_tma / issuers;                   // one 32 KB stage per iteration
    uint8_t* myring = ring + warp * slots * 32768;
    uint64_t* myfull = full[warp];
    for (int i = 0; i < total + slots; ++i) {
      if (i >= slots) { const int s = (i - slots) % slots; mbar_wait(&myfull[s], ((i - slots) / slots) & 1); }
      if (i < total) {
        const int s = i % slots;
        mbar_expect_tx(&myfull[s], 32768);
        const int cta = (mode & 8) ? 0 : blockIdx.x;        // bit 3: every CTA walks the SAME rows (the conv kernels' weight tiles)
        const int row = ((cta * 37 + (i * issuers + warp) * 5) % 120) * 128;
        if (mode & 256) {    // a halo box as conv_tc3.cu loads it: 64 ch x 8 px x 32 rows (32 KB) of an NHWC tensor, per-CTA position
          tma_load_4d(myring + s * 32768, &tmA4, &myfull[s], ((i & 1) * 64), (blockIdx.x % 32) * 8 + (i % 3) - 1, ((blockIdx.x / 32) * 32 + (i % 5)) % 160, 0);
        } else if (mode & 4) {      // one 32 KB box (64 K x 256 rows) instead of two 16 KB boxes
          tma_load_2d(myring + s * 32768, &tmB2, &myfull[s], (i % 16) * 64, row % (128 * 118));
        } else {
          tma_load_2d(myring + s * 32768, &tmB, &myfull[s], (i % 16) * 64, row);
          tma_load_2d(myring + s * 32768 + 16384, &tmB, &myfull[s], ((i + 7) % 16) * 64, (row + 128 * 60) % (128 * 120));
        }
      }
    }
    if (warp == 0) { out[blockIdx.x].t_tma = clock64() - t0; out[blockIdx.x].n_tma = total * issuers; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static CUtensorMap g_tm2, g_tmA[2];
template <int N> static void run(const CUtensorMap& tm, int mode, int slots, int a_stride, ProbeOut* d_out, int ctas, const char* label) {
  const int iters = 256000 / N, iters_tma = 1000;          // ~512k clocks of nominal MMA work; 32 MB of loads per CTA
  const int issuers = ((mode >> 4) & 7) ? ((mode >> 4) & 7) : 1;
  const size_t smem = (mode & 1 ? 2 * 49152 : 0) + 1024 + (mode & 2 ? issuers * slots * 32768 : 0);
  if (smem > 227 * 1024) { printf("%-44s skipped (smem)\n", label); return; }
  CK(cudaFuncSetAttribute(k_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * ctas));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_probe<N><<<ctas, 192, smem>>>(tm, g_tm2, g_tmA[(mode >> 9) & 1], mode, iters, iters_tma, slots, a_stride, d_out);       // warm-up
  CK(cudaEventRecord(e0));
  k_probe<N><<<ctas, 192, smem>>>(tm, g_tm2, g_tmA[(mode >> 9) & 1], mode, iters, iters_tma, slots, a_stride, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  ProbeOut* h = (ProbeOut*)malloc(sizeof(ProbeOut) * ctas);
  CK(cudaMemcpy(h, d_out, sizeof(ProbeOut) * ctas, cudaMemcpyDeviceToHost));
  double cm = 0, ct = 0; int nm = 0, nt = 0;
  for (int i = 0; i < ctas; ++i) { if (h[i].n_mma) { cm += (double)h[i].t_mma / h[i].n_mma; ++nm; } if (h[i].n_tma) { ct += (double)h[i].t_tma / h[i].n_tma; ++nt; } }
  printf("%-44s kernel %7.1f us |", label, ms * 1e3);
  if (nm) printf(" %6.1f clk per 128x%dx16 MMA (nominal %d)", cm / nm, N, N / 2);
  if (nt) printf(" | %7.1f clk per 32 KB stage = %5.1f B/clk/SM, chip %5.2f TB/s", ct / nt, 32768.0 / (ct / nt), (double)ctas * iters_tma * 32768.0 / (ms * 1e-3) / 1e12);
  printf("\n");
  free(h);
}

int main() {
  CK(cudaSetDevice(0));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  __half* d_w; const size_t K = 1024, rows = 128 * 120;     // 30 MB fp16: L2 resident
  CK(cudaMalloc(&d_w, K * rows * 2)); CK(cudaMemset(d_w, 0, K * rows * 2));
  CUtensorMap tm;
  cuuint64_t dims[2] = {K, rows}; cuuint64_t strides[1] = {K * 2}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  { cuuint32_t box2[2] = {64, 256};
    if (encode(&g_tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_w, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 2 failed\n"); return 1; } }
  for (int st = 1; st <= 2; ++st) {   // NHWC activation [1][192*st][256*st][128] fp16, halo box 64 ch x 8 px x 32 rows with element stride st
    __half* d_a; CK(cudaMalloc(&d_a, (size_t)192 * st * 256 * st * 128 * 2)); CK(cudaMemset(d_a, 0, (size_t)192 * st * 256 * st * 128 * 2));
    cuuint64_t ad[4] = {128, (cuuint64_t)256 * st, (cuuint64_t)192 * st, 1}; cuuint64_t as_[3] = {256, (cuuint64_t)256 * st * 256, (cuuint64_t)192 * st * 256 * st * 256};
    cuuint32_t ab[4] = {64, (cuuint32_t)(8 * st), (cuuint32_t)(32 * st), 1}; cuuint32_t ae[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
    if (encode(&g_tmA[st - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d_a, ad, as_, ab, ae, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
  }
  ProbeOut* d_out; CK(cudaMalloc(&d_out, sizeof(ProbeOut) * sms));
  printf("SMs %d\n", sms);
  run<64>(tm, 1, 0, 0, d_out, sms, "MMA only, N = 64");
  run<128>(tm, 1, 0, 0, d_out, sms, "MMA only, N = 128");
  run<256>(tm, 1, 0, 0, d_out, sms, "MMA only, N = 256");
  run<128>(tm, 1, 0, 1024, d_out, sms, "MMA only, N = 128, A start +1024 B on odd groups");
  run<128>(tm, 1 | 1024, 0, 0, d_out, sms, "MMA only, N = 128, RANDOM operands");
  run<256>(tm, 1 | 1024, 0, 0, d_out, sms, "MMA only, N = 256, RANDOM operands");
  run<128>(tm, 1 | 1024, 0, 0, d_out, 16, "MMA only, N = 128, RANDOM operands, 16 CTAs");
  run<128>(tm, 3 | 1024 | (2 << 4), 2, 0, d_out, sms, "MMA N = 128 RANDOM + TMA 2 issuers x 2 stages");
  run<128>(tm, 2, 2, 0, d_out, sms, "TMA only, 2 stages in flight");
  run<128>(tm, 2, 4, 0, d_out, sms, "TMA only, 4 stages in flight");
  run<128>(tm, 2, 1, 0, d_out, sms, "TMA only, 1 stage in flight (latency)");
  run<128>(tm, 2, 4, 0, d_out, 1, "TMA only, 4 stages, ONE CTA");
  run<128>(tm, 2 | 4, 4, 0, d_out, sms, "TMA only, one 32 KB box per stage");
  run<128>(tm, 2 | 8, 4, 0, d_out, sms, "TMA only, all CTAs SAME rows (hot lines)");
  run<128>(tm, 2 | 8 | (2 << 4), 3, 0, d_out, sms, "TMA only, same rows, 2 issuers");
  run<128>(tm, 2 | 256, 4, 0, d_out, sms, "TMA only, 4-D halo box 64x8x32, stride 1");
  run<128>(tm, 2 | 256 | 512, 4, 0, d_out, sms, "TMA only, 4-D halo box 64x8x32, stride 2");
  run<128>(tm, 2 | 256 | (2 << 4), 3, 0, d_out, sms, "TMA only, halo box stride 1, 2 issuers");
  run<128>(tm, 2 | (2 << 4), 3, 0, d_out, sms, "TMA only, 2 issuer threads x 3 stages");
  run<128>(tm, 2 | (4 << 4), 1, 0, d_out, sms, "TMA only, 4 issuer threads x 1 stage");
  run<128>(tm, 2 | (2 << 4), 3, 0, d_out, 1, "TMA only, 2 issuers x 3 stages, ONE CTA");
  run<128>(tm, 3 | (2 << 4), 2, 0, d_out, sms, "MMA N = 128 + TMA 2 issuers x 2 stages");
  run<64>(tm, 3, 4, 0, d_out, sms, "MMA N = 64 + TMA 4 stages in flight");
  run<128>(tm, 3, 4, 0, d_out, sms, "MMA N = 128 + TMA 4 stages in flight");
  run<256>(tm, 3, 4, 0, d_out, sms, "MMA N = 256 + TMA 4 stages in flight");
  run<128>(tm, 3, 2, 0, d_out, sms, "MMA N = 128 + TMA 2 stages in flight");
  run<128>(tm, 3, 4, 0, d_out, sms / 2, "MMA N = 128 + TMA 4 stages, half the SMs");
  return 0;
}
