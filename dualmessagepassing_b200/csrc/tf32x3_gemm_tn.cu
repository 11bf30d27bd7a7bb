// tf32x3_gemm_tn.cu -- weight-gradient GEMM  D[M,N] = sum_e (row_scale[e] * X[e,0:M])^T * G[e,0:N]
// on the tcgen05 tensor cores with fp32-level accuracy (3xTF32 split), M, N in {64, 128}, E = edges.
//
// Replaces the K = E long autograd reductions `X.t() @ G` of the DMPNN layer (dW_eloop, dW_src/dst, dW_in/out,
// MLP dW1/dW2: 5 edge-sized reductions per layer, 145 of the 413 ms step at config 5 on cuBLAS sgemm).
//
// The contraction index (the edge) is the SLOW index of both row-major operands, i.e. both are "MN-major" for
// the MMA; tcgen05 kind::tf32 accepts that directly, so no transpose is staged:
//   smem tile of 16 edges x 128 features = 4 column blocks (32 features = 128 B) x 16 edge rows; a block is
//   four 4-row / 512-byte atoms of the SWIZZLE_128B_BASE32B layout (SBO = 512 B between 4-edge atoms, LBO = 2048 B
//   between feature blocks); the descriptor start advances by 1024 B per MMA k-step (8 edges).
// Structure mirrors tf32x3_gemm.cu: 8 producer warps (cp.async of whole 512-byte rows into the swizzled hi tile, then
// lo = x - trunc_tf32(x), optional row scale and running column sums), 1 MMA warp (6 tcgen05.mma per 16-edge stage),
// 8 flush warps.  Each CTA owns a contiguous range
// of edges; to bound the length of any tensor-core accumulation chain the accumulator is double-buffered in TMEM
// and FLUSHED every kFlushStages stages into an fp32 partial in global memory (round-to-nearest adds, L2-resident),
// and a second kernel adds the per-CTA partials in a fixed order -> deterministic, no atomics.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dmp {
namespace gemm {

constexpr int kTnEdges = 16;            // edges per stage (2 MMA k-steps): small stages -> 6 of them fit, and the
constexpr int kTnStages = 6;            // asynchronous copies can run kTnCopyDepth stages (64 KB per SM) ahead
constexpr int kTnCopyDepth = kTnStages - 2;
constexpr int kTnProducerWarps = 8;
constexpr int kTnProducerThreads = kTnProducerWarps * 32;
constexpr int kTnFlushWarps = 8;
constexpr int kTnMmaWarp = 8;
constexpr int kTnThreads = (kTnFlushWarps + 1 + kTnProducerWarps) * 32;  // 544
constexpr int kFlushStages = 32;        // 512 edges per tensor-core accumulation chain

struct TnParams {
  const float* X; int64_t ldx;
  const float* row_scale;
  const float* G; int64_t ldg;
  float* partial;        // [grid][N][M]  (transposed: lanes = m are contiguous)
  float* part_sx;        // [grid][M] column sums of X (unscaled) or NULL
  float* part_sg;        // [grid][N] column sums of G or NULL
  int64_t E;
  uint32_t lbo, sbo, kadv;   // descriptor geometry (bytes); defaults set by the host wrapper
  uint32_t idesc_xor, ltype;
  int l2_prefetch;
};

// kind::tf32, fp32 accumulate, A and B MN-major
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

template <int M, int N>
struct TnSmem {
  static constexpr int kXBytes = kTnEdges * 128 * 4;        // one of hi / lo; always 4 feature blocks: the MMA runs
                                                            // with M = 128 (for M = 64 the upper two blocks stay zero)
  static constexpr int kGBytes = kTnEdges * N * 4;
  static constexpr int kStageBytes = 2 * kXBytes + 2 * kGBytes;
  static constexpr int kSumBytes = 2 * kTnProducerWarps * 128 * 4;   // per-warp column-sum scratch (X and G)
  static constexpr int kTotal = kTnStages * kStageBytes + 256 + kSumBytes + 1024;
};

// smem offset of 16-byte chunk c16 (4 features) of edge row k inside a [32 edges x F features] MN-major tile.
// MN-major tf32 operands must use the "128B swizzle with 32-byte base" layout (CUTLASS: SW128_32B is the only
// layout for mn-major tf32): rows of 128 B (32 features of one edge), 4-row / 512-byte atoms, and the 32-BYTE chunk
// index inside a row XOR-ed with (row & 3)  -- Swizzle<2,5,2> on byte addresses.
__device__ __forceinline__ uint32_t swz_mn(int k, int c16) {
  const int c = c16 & 7;
  return (uint32_t)((c16 >> 3) * (kTnEdges * 128) + k * 128 + ((((c >> 1) ^ (k & 3)) << 5) | ((c & 1) << 4)));
}
// descriptor for that layout: layout type 1 (SWIZZLE_128B_BASE32B), sm_100 version bit
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t type) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)type << 61);
}

template <int M, int N, bool SX, bool SG>
__global__ void __launch_bounds__(kTnThreads, 1) tf32x3_gemm_tn_kernel(const TnParams p) {
  using L = TnSmem<M, N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + kTnStages * L::kStageBytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * kTnStages;
  const uint32_t bar_acc_full = sBar + 16 * kTnStages, bar_acc_empty = bar_acc_full + 16;
  const uint32_t tmem_slot = bar_acc_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* sum_scratch = reinterpret_cast<float*>(smem_raw + (sBar + 256 - smem_u32(smem_raw)));   // [2][8 warps][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // this CTA's contiguous range of 32-edge stages
  const int64_t stages_total = (p.E + kTnEdges - 1) / kTnEdges;
  const int64_t s_begin = stages_total * blockIdx.x / gridDim.x;
  const int64_t s_end = stages_total * (blockIdx.x + 1) / gridDim.x;
  const int64_t n_stages = s_end - s_begin;
  const int64_t n_flush = (n_stages + kFlushStages - 1) / kFlushStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTnStages; ++s) {
      mbar_init(bar_full + 8 * s, kTnProducerWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kTnFlushWarps);
    }
    fence_barrier_init();
  }
  constexpr int kTmemCols = 2 * N;
  if (warp == kTnMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  if constexpr (M < 128) {   // feature blocks 2,3 of every X tile are never written by the producers: zero them once
    for (int i = threadIdx.x; i < kTnStages * 2 * (L::kXBytes / 16); i += kTnThreads) {
      const int st = i / (2 * (L::kXBytes / 16)), rem = i % (2 * (L::kXBytes / 16));
      const uint32_t addr = base + st * L::kStageBytes + rem * 16;
      asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "f"(0.0f) : "memory");
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp > kTnMmaWarp) {
    // =========================== PRODUCERS ===========================
    const int pt = threadIdx.x - (kTnMmaWarp + 1) * 32;   // 0..255
    // X tile: 16 edges x M/4 chunks of 16 B; thread handles chunk ids pt + 256*i.  A warp covers whole rows.
    constexpr int kXChunks = kTnEdges * M / 4, kGChunks = kTnEdges * N / 4;
    constexpr int kXPer = kXChunks / kTnProducerThreads, kGPer = kGChunks / kTnProducerThreads;   // 2 (or 1)
    // bias gradients for free: a thread always handles the same 4 columns (256 % (M/4) == 0), so it keeps running
    // column sums of everything it streams; rows past E are zero-filled and add nothing
    float4 sum_x = make_float4(0.f, 0.f, 0.f, 0.f), sum_g = sum_x;
    uint32_t offx[kXPer], offg[kGPer];
#pragma unroll
    for (int i = 0; i < kXPer; ++i) offx[i] = swz_mn((pt + kTnProducerThreads * i) / (M / 4), (pt + kTnProducerThreads * i) % (M / 4));
#pragma unroll
    for (int i = 0; i < kGPer; ++i) offg[i] = swz_mn((pt + kTnProducerThreads * i) / (N / 4), (pt + kTnProducerThreads * i) % (N / 4));
    int istage = 0;
    uint32_t iphase = 0;
    auto issue = [&](int64_t st) {
      mbar_wait(bar_empty + 8 * istage, iphase ^ 1);
      const int64_t e0 = (s_begin + st) * kTnEdges;
      const uint32_t x_hi = base + istage * L::kStageBytes, g_hi = x_hi + 2 * L::kXBytes;
#pragma unroll
      for (int i = 0; i < kXPer; ++i) {
        const int c = pt + kTnProducerThreads * i;
        const int64_t e = e0 + c / (M / 4);
        const bool ok = e < p.E;
        cp_async16(x_hi + offx[i], ok ? (const void*)(p.X + e * p.ldx + (c % (M / 4)) * 4) : (const void*)p.X, ok ? 16u : 0u);
      }
#pragma unroll
      for (int i = 0; i < kGPer; ++i) {
        const int c = pt + kTnProducerThreads * i;
        const int64_t e = e0 + c / (N / 4);
        const bool ok = e < p.E;
        cp_async16(g_hi + offg[i], ok ? (const void*)(p.G + e * p.ldg + (c % (N / 4)) * 4) : (const void*)p.G, ok ? 16u : 0u);
      }
      if (++istage == kTnStages) { istage = 0; iphase ^= 1; }
    };
#pragma unroll
    for (int d = 0; d < kTnCopyDepth; ++d) {
      if (d < n_stages) issue(d);
      cp_async_commit();
    }
    int stage = 0;
    for (int64_t st = 0; st < n_stages; ++st) {
      cp_async_wait<kTnCopyDepth - 1>();                       // this thread's copies of stage `st` have landed
      const uint32_t x_hi = base + stage * L::kStageBytes, x_lo = x_hi + L::kXBytes;
      const uint32_t g_hi = x_lo + L::kXBytes, g_lo = g_hi + L::kGBytes;
#pragma unroll
      for (int i = 0; i < kXPer; ++i) {
        float4 v = lds128(x_hi + offx[i]);
        if constexpr (SX) {
          sum_x.x = __fadd_rn(sum_x.x, v.x); sum_x.y = __fadd_rn(sum_x.y, v.y);
          sum_x.z = __fadd_rn(sum_x.z, v.z); sum_x.w = __fadd_rn(sum_x.w, v.w);
        }
        if (p.row_scale != nullptr) {
          const int64_t e = (s_begin + st) * kTnEdges + (pt + kTnProducerThreads * i) / (M / 4);
          const float sc = e < p.E ? __ldg(p.row_scale + e) : 1.0f;
          v.x = __fmul_rn(sc, v.x); v.y = __fmul_rn(sc, v.y); v.z = __fmul_rn(sc, v.z); v.w = __fmul_rn(sc, v.w);
          sts128(x_hi + offx[i], v);
        }
        sts128(x_lo + offx[i], make_float4(tf32_trunc_residual(v.x), tf32_trunc_residual(v.y),
                                           tf32_trunc_residual(v.z), tf32_trunc_residual(v.w)));
      }
#pragma unroll
      for (int i = 0; i < kGPer; ++i) {
        const float4 v = lds128(g_hi + offg[i]);
        if constexpr (SG) {
          sum_g.x = __fadd_rn(sum_g.x, v.x); sum_g.y = __fadd_rn(sum_g.y, v.y);
          sum_g.z = __fadd_rn(sum_g.z, v.z); sum_g.w = __fadd_rn(sum_g.w, v.w);
        }
        sts128(g_lo + offg[i], make_float4(tf32_trunc_residual(v.x), tf32_trunc_residual(v.y),
                                           tf32_trunc_residual(v.z), tf32_trunc_residual(v.w)));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      if (++stage == kTnStages) stage = 0;
      if (st + kTnCopyDepth < n_stages) issue(st + kTnCopyDepth);
      cp_async_commit();
    }
    if constexpr (SX || SG) {
      // threads with equal (pt % chunks-per-row) own the same columns: with 32 chunks per row that is one lane of each
      // of the 8 warps; with 16 chunks per row (64 features) lanes l and l+16 of every warp as well
      const int pw = pt >> 5;
      float* sx = sum_scratch + pw * 128;
      float* sg = sum_scratch + (kTnProducerWarps + pw) * 128;
      if (M == 64) {
        sum_x.x += __shfl_down_sync(0xffffffffu, sum_x.x, 16); sum_x.y += __shfl_down_sync(0xffffffffu, sum_x.y, 16);
        sum_x.z += __shfl_down_sync(0xffffffffu, sum_x.z, 16); sum_x.w += __shfl_down_sync(0xffffffffu, sum_x.w, 16);
      }
      if (N == 64) {
        sum_g.x += __shfl_down_sync(0xffffffffu, sum_g.x, 16); sum_g.y += __shfl_down_sync(0xffffffffu, sum_g.y, 16);
        sum_g.z += __shfl_down_sync(0xffffffffu, sum_g.z, 16); sum_g.w += __shfl_down_sync(0xffffffffu, sum_g.w, 16);
      }
      if (SX && lane < M / 4) *reinterpret_cast<float4*>(sx + 4 * lane) = sum_x;
      if (SG && lane < N / 4) *reinterpret_cast<float4*>(sg + 4 * lane) = sum_g;
      asm volatile("bar.sync 1, %0;" ::"n"(kTnProducerThreads) : "memory");
      if (SX && pt < M) {
        float t = 0.0f;
        for (int w = 0; w < kTnProducerWarps; ++w) t = __fadd_rn(t, sum_scratch[w * 128 + pt]);
        p.part_sx[(int64_t)blockIdx.x * M + pt] = t;
      }
      if (SG && pt < N) {
        float t = 0.0f;
        for (int w = 0; w < kTnProducerWarps; ++w) t = __fadd_rn(t, sum_scratch[(kTnProducerWarps + w) * 128 + pt]);
        p.part_sg[(int64_t)blockIdx.x * N + pt] = t;
      }
    }
  } else if (warp == kTnMmaWarp) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = make_idesc_mn(128, N) ^ p.idesc_xor;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int64_t st = 0;
    for (int64_t f = 0; f < n_flush; ++f) {
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
      const int64_t st_hi = (st + kFlushStages < n_stages) ? st + kFlushStages : n_stages;
      bool first = true;
      for (; st < st_hi; ++st) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t x_hi = base + stage * L::kStageBytes, x_lo = x_hi + L::kXBytes;
          const uint32_t g_hi = x_lo + L::kXBytes, g_lo = g_hi + L::kGBytes;
#pragma unroll
          for (int j = 0; j < kTnEdges / 8; ++j) {
            const uint64_t dxh = smem_desc_mn32(x_hi + j * p.kadv, p.lbo, p.sbo, p.ltype);
            const uint64_t dxl = smem_desc_mn32(x_lo + j * p.kadv, p.lbo, p.sbo, p.ltype);
            const uint64_t dgh = smem_desc_mn32(g_hi + j * p.kadv, p.lbo, p.sbo, p.ltype);
            const uint64_t dgl = smem_desc_mn32(g_lo + j * p.kadv, p.lbo, p.sbo, p.ltype);
            umma_tf32(d_tmem, dxl, dgh, idesc, (first && j == 0) ? 0u : 1u);
            umma_tf32(d_tmem, dxh, dgl, idesc, 1u);
            umma_tf32(d_tmem, dxh, dgh, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * stage);
          if (st == st_hi - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        first = false;
        if (++stage == kTnStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== FLUSH (TMEM -> fp32 partial in global) ===========================
    // TMEM lane = m (feature of X), column = n (feature of G).  Warp w: lane quadrant w%4, column half w/4.
    const int quad = warp & 3, half = warp >> 2;
    const int m = quad * 32 + lane;
    constexpr int kColsPerWarp = N / 2;
    float* part = p.partial + (int64_t)blockIdx.x * M * N + m;   // element (m, n) at n*M + m: lanes contiguous
    int acc = 0;
    uint32_t acc_phase = 0;
    {
      for (int64_t f = 0; f < n_flush; ++f) {
        mbar_wait(bar_acc_full + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * N + half * kColsPerWarp);
#pragma unroll 1
        for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + c0, v);
          if (m < M) {
            float* q = part + (int64_t)(half * kColsPerWarp + c0) * M;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float old = (f == 0) ? 0.0f : q[j * M];
              q[j * M] = __fadd_rn(old, v[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (n_flush == 0 && m < M) {                      // CTA without work still owns a (zero) partial
        for (int n = half * kColsPerWarp; n < (half + 1) * kColsPerWarp; ++n) part[(int64_t)n * M] = 0.0f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTnMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// D[m, n] (+)= sum over CTAs (ascending) of partial[cta][n][m]
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ partial, int parts, int M, int N,
                                                        float* __restrict__ D, int64_t ldd, int accumulate,
                                                        const float* __restrict__ part_sx, float* __restrict__ sum_x,
                                                        const float* __restrict__ part_sg, float* __restrict__ sum_g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over N*M, m fastest (coalesced partial reads)
  if (idx >= M * N) {
    const int j = idx - M * N;                               // tail threads: the column sums
    if (j < M && sum_x != nullptr) {
      float s = 0.0f;
      for (int c = 0; c < parts; ++c) s = __fadd_rn(s, part_sx[(int64_t)c * M + j]);
      sum_x[j] = s;
    } else if (j >= M && j < M + N && sum_g != nullptr) {
      float s = 0.0f;
      for (int c = 0; c < parts; ++c) s = __fadd_rn(s, part_sg[(int64_t)c * N + (j - M)]);
      sum_g[j - M] = s;
    }
    return;
  }
  const int n = idx / M, m = idx % M;
  float s = 0.0f;
  for (int c = 0; c < parts; ++c) s = __fadd_rn(s, partial[(int64_t)c * M * N + idx]);
  float* d = D + (int64_t)m * ldd + n;
  *d = accumulate ? __fadd_rn(*d, s) : s;
}

template <int M, int N, bool SX, bool SG>
static int launch_tn_s(const TnParams& p, unsigned grid, cudaStream_t stream) {
  using L = TnSmem<M, N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tf32x3_gemm_tn_kernel<M, N, SX, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm_tn_tf32x3: cannot reserve %d bytes of shared memory: %s", L::kTotal, cudaGetErrorString(e));
      return DMP_ERR_CUDA;
    }
    configured = true;
  }
  tf32x3_gemm_tn_kernel<M, N, SX, SG><<<grid, kTnThreads, L::kTotal, stream>>>(p);
  return launch_status("tf32x3_gemm_tn_kernel");
}

template <int M, int N>
static int launch_tn(const TnParams& p, unsigned grid, cudaStream_t stream) {
  const bool sx = p.part_sx != nullptr, sg = p.part_sg != nullptr;
  if (sx && sg) return launch_tn_s<M, N, true, true>(p, grid, stream);
  if (sx) return launch_tn_s<M, N, true, false>(p, grid, stream);
  if (sg) return launch_tn_s<M, N, false, true>(p, grid, stream);
  return launch_tn_s<M, N, false, false>(p, grid, stream);
}

}  // namespace gemm
}  // namespace dmp

extern "C" int dmp_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t* bytes_host) {
  using namespace dmp;
  DMP_CHECK_ARG(bytes_host != nullptr, "gemm_tn_workspace_bytes: null output");
  *bytes_host = (int64_t)kNumSMs * (M * N + M + N) * 4;
  return DMP_OK;
}

extern "C" int dmp_gemm_tn_tf32x3(const float* X, int64_t ldx, const float* row_scale, const float* G, int64_t ldg,
                                  float* D, int64_t ldd, float* colsum_x, float* colsum_g, int64_t E, int64_t M,
                                  int64_t N, int accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace dmp;
  using namespace dmp::gemm;
  DMP_CHECK_ARG(E >= 0, "gemm_tn_tf32x3: negative E");
  DMP_CHECK_ARG(D != nullptr && ldd >= N, "gemm_tn_tf32x3: bad output");
  DMP_CHECK_ARG((M == 64 || M == 128) && (N == 64 || N == 128), "gemm_tn_tf32x3: M and N must be 64 or 128");
  DMP_CHECK_ARG(E == 0 || (X && G && ldx >= M && ldg >= N && ldx % 4 == 0 && ldg % 4 == 0 && aligned_to(X, 16) &&
                           aligned_to(G, 16)),
                "gemm_tn_tf32x3: operands must have dense 16-byte aligned rows");
  const int64_t stages = (E + kTnEdges - 1) / kTnEdges;
  int64_t grid = stages < kNumSMs ? stages : kNumSMs;
  if (grid < 1) grid = 1;
  DMP_CHECK_ARG(workspace != nullptr && workspace_bytes >= grid * (M * N + M + N) * 4,
                "gemm_tn_tf32x3: workspace too small");
  TnParams p;
  p.X = X; p.ldx = ldx; p.row_scale = row_scale; p.G = G; p.ldg = ldg;
  p.partial = static_cast<float*>(workspace); p.E = E;
  p.part_sx = colsum_x ? p.partial + grid * M * N : nullptr;
  p.part_sg = colsum_g ? p.partial + grid * (M * N + M) : nullptr;
  // 4096 B between 32-feature blocks (LBO), 512 B between 4-edge atoms (SBO), 1024 B per MMA k-step (8 edges)
  p.lbo = kTnEdges * 128; p.sbo = 512; p.kadv = 1024; p.idesc_xor = 0; p.ltype = 1;
  p.l2_prefetch = 0;
  if (const char* dbg = getenv("DMP_TN_DBG"))
    sscanf(dbg, "%u,%u,%u,%u,%u", &p.lbo, &p.sbo, &p.kadv, &p.idesc_xor, &p.ltype);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (M == 128 && N == 128) rc = launch_tn<128, 128>(p, (unsigned)grid, s);
  else if (M == 128 && N == 64) rc = launch_tn<128, 64>(p, (unsigned)grid, s);
  else if (M == 64 && N == 128) rc = launch_tn<64, 128>(p, (unsigned)grid, s);
  else rc = launch_tn<64, 64>(p, (unsigned)grid, s);
  if (rc != DMP_OK) return rc;
  const int total = (int)(M * N + M + N);
  tn_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(p.partial, (int)grid, (int)M, (int)N, D, ldd, accumulate,
                                                      p.part_sx, colsum_x, p.part_sg, colsum_g);
  return launch_status("tn_reduce_kernel");
}
