// tf32x3_gemm_tn.cu -- weight-gradient GEMM  D[M,N] = sum_e (row_scale[e] * X[e,0:M])^T * G[e,0:N]
// on the tcgen05 tensor cores with fp32-level accuracy (3xTF32 split), M, N in {64, 128}, E = edges.
//
// Replaces the K = E long autograd reductions `X.t() @ G` of the DMPNN layer (dW_eloop, dW_src/dst, dW_in/out,
// MLP dW1/dW2: 4 edge-sized reductions per layer call, 145 of the 413 ms step at config 5 on cuBLAS sgemm).
//
// The contraction index (the edge) is the SLOW index of both row-major operands.
//   * G is the B operand, read from shared memory "MN-major": a tile of 16 edges x 128 features = 4 column blocks
//     (32 features = 128 B) x 16 edge rows; a block is four 4-row / 512-byte atoms of the SWIZZLE_128B_BASE32B layout
//     (SBO = 512 B between 4-edge atoms, LBO = 2048 B between feature blocks); the descriptor start advances by
//     1024 B per MMA k-step (8 edges).  cp.async lands whole rows directly in the swizzled hi tile (the tensor core
//     ignores the low 13 mantissa bits, so the raw data IS the hi operand); the split writes lo = g - trunc_tf32(g)
//     into a separate, shorter ring.
//   * X^T is the A operand, read from TENSOR MEMORY (lane = feature, column = edge).  A warp copies its own 32-feature
//     block of 16 rows into a private raw buffer, thread m then reads column m (conflict-free), applies the row scale,
//     splits and writes 16 hi + 16 lo columns with tcgen05.st.  With both operands in shared memory the six MMAs of a
//     stage read 48 KB of it and the kernel was bound by shared-memory bandwidth (MMA-only ablation 6.9 ms for 40 M
//     edges, 0.55 of HBM peak overall); the TMEM form halves that and needs no second smem copy of X.
// Warps: 0-7 X producers (TMEM lane quadrant = warp % 4; warps q and q+4 take even / odd stages, because the
// lds -> tcgen05.st -> wait::st -> arrive chain of one stage is ~0.45 us of latency per warp), 8-11 flush, 12 MMA issuer
// (6 tcgen05.mma per 16-edge stage), 13-16 G producers.  Ring sizing: a G hi slot is held from copy issue to MMA retirement, a lo slot only from split to
// retirement: kTnCopyDepth stages of copies in flight (96 KB per SM) AND kTnLoStages of slack for the
// split -> full barrier -> MMA -> commit -> done barrier round trip (measured ~1 us).  Holding the in-flight data in
// registers instead does not work: ptxas puts every LDG of the loop on one scoreboard, so waiting for the oldest
// load waits for all of them.
// Each CTA owns a contiguous range of edges; to bound the length of any tensor-core accumulation chain the accumulator
// is double-buffered in TMEM and FLUSHED every kFlushStages stages (256 edges) into an fp32 partial in global memory
// (round-to-nearest adds, L2-resident), and a second kernel adds the per-CTA partials in a fixed order ->
// deterministic, no atomics.
#include <stdlib.h>
#include <type_traits>

#include "tc_common.cuh"

namespace dmp {
namespace gemm {

constexpr int kTnEdges = 32;            // edges per stage (4 MMA k-steps): every barrier round trip, tcgen05.st wait
                                        // and commit is paid per stage, 16-edge stages spent 0.28 us on them alone
constexpr int kTnCopyDepth = 3;         // stages of asynchronous copies in flight (96 KB per SM)
constexpr int kTnLoStages = 2;          // ring of G residual tiles = slack between split and MMA retirement
constexpr int kTnHiStages = kTnCopyDepth + kTnLoStages;   // ring of raw (= hi) G tiles
constexpr int kTnXDepth = 2;                              // an X warp sees every second stage: copies in flight per warp
constexpr int kTnXRaw = kTnXDepth + 1;                    // per-warp ring of raw 16 x 32 X blocks
constexpr int kTnASlots = 4;            // TMEM ring of X^T: 32 hi + 32 lo columns per stage
constexpr int kTnAColsPerSlot = 2 * kTnEdges;
constexpr int kTnACol = 256;            // first TMEM column of that ring (accumulators: columns [0, 2N))
constexpr int kTnXWarps = 8, kTnFlushWarps = 4, kTnGWarps = 8;
constexpr int kTnMmaWarp = kTnXWarps + kTnFlushWarps;     // 12
constexpr int kTnGThreads = kTnGWarps * 32;
constexpr int kTnThreads = (kTnXWarps + kTnFlushWarps + 1 + kTnGWarps) * 32;  // 672
// Stages (x 32 edges) per tensor-core accumulation chain.  The accumulator TRUNCATES, so the error of a chain grows
// linearly with its length; measured on 4 M x 128 x 128 (max-norm vs fp64; cuBLAS sgemm 2.5e-6 .. 2.9e-6), 40 M-edge time:
//   64 stages 1.6e-5, 7.99 ms | 32: 7.2e-6, 7.90 ms | 16 (round 1): 3.8e-6, 8.35 ms | 8: 1.9e-6, 8.67 ms
// 8 puts the weight gradients below cuBLAS' own error for 4 % of this kernel's time.
#ifndef DMP_TN_FLUSH_STAGES
#define DMP_TN_FLUSH_STAGES 8
#endif
constexpr int kFlushStages = DMP_TN_FLUSH_STAGES;
constexpr int kTnTmemCols = 512;

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
  int ablate;               // debug only (env DMP_TN_ABLATE): 1 no proxy fence, 2 no MMA, 4 no loads, 8 no split, 16 no flush
};

// kind::tf32, fp32 accumulate, A from tensor memory, B MN-major
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int M, int N>
struct TnSmem {
  static constexpr int kGBytes = kTnEdges * N * 4;          // one of hi / lo
  static constexpr int kXBlockBytes = kTnEdges * 32 * 4;    // 16 edges x 32 features, row-major, one warp's block
  static constexpr int kGHiOff = 0;
  static constexpr int kGLoOff = kTnHiStages * kGBytes;
  static constexpr int kXOff = kGLoOff + kTnLoStages * kGBytes;
  static constexpr int kTileBytes = kXOff + kTnXWarps * kTnXRaw * kXBlockBytes;
  static constexpr int kSumBytes = (kTnGWarps + 1) * 128 * 4;   // per-warp column-sum scratch (G) + odd-stage sums of X
  static constexpr int kTotal = kTileBytes + 256 + kSumBytes + 1024;
};

// smem offset of 16-byte chunk c16 (4 features) of edge row k inside a [16 edges x F features] MN-major tile.
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
  const uint32_t sBar = base + L::kTileBytes;
  const uint32_t bar_full = sBar, bar_done = sBar + 8 * kTnASlots;      // both indexed by stage % kTnASlots
  const uint32_t bar_acc_full = sBar + 16 * kTnASlots, bar_acc_empty = bar_acc_full + 16;
  const uint32_t tmem_slot = bar_acc_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* sum_scratch = reinterpret_cast<float*>(smem_raw + (sBar + 256 - smem_u32(smem_raw)));   // [G warps + 1][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kXActive = M / 32;       // lane quadrants that hold real features (2 for M = 64)
  const int xq = warp & 3, xh = warp >> 2;   // X producers: quadrant and stage parity

  // this CTA's contiguous range of 16-edge stages
  const int64_t stages_total = (p.E + kTnEdges - 1) / kTnEdges;
#ifdef DMP_DEBUG
  const int ablate = p.ablate;
#else
  constexpr int ablate = 0;     // the ablation switches cost predicates and address recomputation in every inner loop
#endif
  const bool interleave = (ablate & 32) == 0;   // stage s of CTA b = b + s * grid: neighbouring SMs stream neighbouring rows
  const int64_t s_begin = interleave ? blockIdx.x : stages_total * blockIdx.x / gridDim.x;
  const int64_t s_end = stages_total * (blockIdx.x + 1) / gridDim.x;
  const int64_t s_step = interleave ? gridDim.x : 1;
  const int64_t n_stages = interleave ? (stages_total - blockIdx.x + gridDim.x - 1) / gridDim.x
                                      : s_end - stages_total * blockIdx.x / gridDim.x;
  const int64_t n_flush = (n_stages + kFlushStages - 1) / kFlushStages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTnASlots; ++s) {
      mbar_init(bar_full + 8 * s, kXActive + kTnGWarps);
      mbar_init(bar_done + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kTnFlushWarps);
    }
    fence_barrier_init();
  }
  if (warp == kTnMmaWarp) tmem_alloc(tmem_slot, kTnTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // =========================== X PRODUCERS: global -> raw smem block -> split -> tensor memory ===========================
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16) + kTnACol;
    if (warp >= kXActive) {
      // M = 64: the MMA still runs with 128 rows; lanes 64..127 of the A ring are zero for the whole kernel
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0.0f;
      for (int s = 0; s < kTnASlots * kTnAColsPerSlot / 32; ++s) tmem_st32(t_lane + s * 32, z);
      tmem_st_wait();
      tc_fence_before();
    }
  }
  if constexpr (kXActive < 4) {   // make those zeros visible to the MMA warp before its first instruction
    __syncthreads();
    tc_fence_after();
  }

  if (warp < kTnXWarps && xq < kXActive) {
    const uint32_t t_lane = tmem_base + ((uint32_t)(xq * 32) << 16) + kTnACol;
    const uint32_t xraw = base + L::kXOff + (uint32_t)warp * kTnXRaw * L::kXBlockBytes;
    const bool scaled = p.row_scale != nullptr;
    const int crow = lane >> 3, cch = lane & 7;       // this lane copies chunk cch of rows crow + 4 i
    float sum_x = 0.0f;
    int islot = 0;
    // source of this lane's first chunk of the stage being issued; advanced by two stages per issue
    const float* xsrc = p.X + ((s_begin + xh * s_step) * kTnEdges + crow) * p.ldx + xq * 32 + cch * 4;
    const int64_t xadv = 2 * s_step * kTnEdges * p.ldx;
    auto issue = [&](int64_t st) {
      const int64_t e0 = (s_begin + st * s_step) * kTnEdges;
      const uint32_t dst = xraw + (uint32_t)islot * L::kXBlockBytes + (uint32_t)(crow * 128 + cch * 16);
      if (!(ablate & (4 | 64))) {
        if (e0 + kTnEdges <= p.E) {
#pragma unroll
          for (int i = 0; i < kTnEdges / 4; ++i) cp_async16(dst + i * 512, xsrc + 4 * i * p.ldx, 16u);
        } else {
#pragma unroll
          for (int i = 0; i < kTnEdges / 4; ++i) {
            const bool ok = e0 + crow + 4 * i < p.E;
            cp_async16(dst + i * 512, ok ? (const void*)(xsrc + 4 * i * p.ldx) : (const void*)p.X, ok ? 16u : 0u);
          }
        }
      }
      xsrc += xadv;
      if (++islot == kTnXRaw) islot = 0;
    };
#pragma unroll
    for (int d = 0; d < kTnXDepth; ++d) {
      if (xh + 2 * d < n_stages) issue(xh + 2 * d);
      cp_async_commit();
    }
    int rslot = 0, aslot = xh;
    uint32_t aphase = 0;      // parity of the done-barrier phase that frees A slot `aslot` (valid from the second lap)
    for (int64_t st = xh; st < n_stages; st += 2) {
      float my_scale = 1.0f;
      if (scaled) {
        const int64_t e = (s_begin + st * s_step) * kTnEdges + lane;       // kTnEdges == 32: one edge per lane
        if (e < p.E) my_scale = __ldg(p.row_scale + e);
      }
      cp_async_wait<kTnXDepth - 1>();
      __syncwarp();                                            // the whole block of this warp has landed
      if (st >= kTnASlots) mbar_wait(bar_done + 8 * aslot, aphase ^ 1);   // MMA of stage st - kTnASlots has retired
      tc_fence_after();
      const uint32_t src = xraw + (uint32_t)rslot * L::kXBlockBytes + (uint32_t)lane * 4;
      const uint32_t t_slot = t_lane + aslot * kTnAColsPerSlot;
      // warp-uniform choice made once per stage, not per element
      auto split_stage = [&](auto with_scale) {
#pragma unroll
        for (int h16 = 0; h16 < kTnEdges / 16; ++h16) {
          float hi[16], lo[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            float v = (ablate & 8) ? 0.0f : lds32(src + (h16 * 16 + r) * 128);
            if constexpr (SX) sum_x = __fadd_rn(sum_x, v);
            if constexpr (decltype(with_scale)::value) v = __fmul_rn(__shfl_sync(0xffffffffu, my_scale, h16 * 16 + r), v);
            // the tensor core reads only the top 19 bits of a tf32 operand (TMEM as well as smem), so the raw word IS
            // the truncated hi part and lo = v - trunc(v) is exact: 2 instructions per element where cvt.rna.tf32
            // expands to 6 (the X warps were 65 % busy with the rounding form)
            hi[r] = v;
            lo[r] = tf32_trunc_residual(v);
          }
          tmem_st16(t_slot + h16 * 16, hi);
          tmem_st16(t_slot + kTnEdges + h16 * 16, lo);
        }
      };
      if (scaled) split_stage(std::true_type{}); else split_stage(std::false_type{});
      if (++rslot == kTnXRaw) rslot = 0;
      if (st + 2 * kTnXDepth < n_stages) issue(st + 2 * kTnXDepth);   // overlaps the tcgen05.st latency
      cp_async_commit();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * aslot);
      aslot += 2;
      if (aslot >= kTnASlots) { aslot -= kTnASlots; aphase ^= 1; }
    }
    if constexpr (SX) {
      // the two warps of a quadrant hold the even-stage and odd-stage halves of the same column sums
      float* sx = sum_scratch + kTnGWarps * 128;
      if (xh == 1) sx[xq * 32 + lane] = sum_x;
      asm volatile("bar.sync 2, %0;" ::"n"(kXActive * 64) : "memory");
      if (xh == 0) p.part_sx[(int64_t)blockIdx.x * M + xq * 32 + lane] = __fadd_rn(sum_x, sx[xq * 32 + lane]);
    }
  } else if (warp > kTnMmaWarp) {
    // =========================== G PRODUCERS: global -> swizzled hi tile, lo = g - trunc(g) ===========================
    const int gt = threadIdx.x - (kTnMmaWarp + 1) * 32;   // 0..kTnGThreads-1
    constexpr int kGChunks = kTnEdges * N / 4;
    constexpr int kGPer = kGChunks / kTnGThreads;          // 4 (N = 128) or 2
    // bias gradients for free: a thread always handles the same 4 columns (128 % (N/4) == 0), so it keeps running
    // column sums of everything it streams; rows past E are zero-filled and add nothing
    float4 sum_g = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t offg[kGPer];
#pragma unroll
    for (int i = 0; i < kGPer; ++i) offg[i] = swz_mn((gt + kTnGThreads * i) / (N / 4), (gt + kTnGThreads * i) % (N / 4));
    const uint32_t ghi = base + L::kGHiOff, glo = base + L::kGLoOff;
    int ihi = 0;
    // chunk i of this thread is row grow + (512 / N) i, column chunk gch: constant smem and global strides
    constexpr int kRowStep = kTnGThreads / (N / 4);          // 8 (N = 128) or 16 (N = 64)
    const int grow = gt / (N / 4), gch = gt % (N / 4);
    const float* gsrc = p.G + (s_begin * kTnEdges + grow) * p.ldg + gch * 4;
    const int64_t gadv = s_step * kTnEdges * p.ldg;
    auto issue = [&](int64_t st) {
      const int64_t e0 = (s_begin + st * s_step) * kTnEdges;
      const uint32_t g_hi = ghi + (uint32_t)ihi * L::kGBytes;
      if (!(ablate & (4 | 128))) {
        if (e0 + kTnEdges <= p.E) {
#pragma unroll
          for (int i = 0; i < kGPer; ++i) cp_async16(g_hi + offg[i], gsrc + kRowStep * i * p.ldg, 16u);
        } else {
#pragma unroll
          for (int i = 0; i < kGPer; ++i) {
            const bool ok = e0 + grow + kRowStep * i < p.E;
            cp_async16(g_hi + offg[i], ok ? (const void*)(gsrc + kRowStep * i * p.ldg) : (const void*)p.G, ok ? 16u : 0u);
          }
        }
      }
      gsrc += gadv;
      if (++ihi == kTnHiStages) ihi = 0;
    };
#pragma unroll
    for (int d = 0; d < kTnCopyDepth; ++d) {
      if (d < n_stages) issue(d);
      cp_async_commit();
    }
    int shi = 0, slo = 0, fslot = 0, dslot = kTnASlots - kTnLoStages;
    uint32_t dphase = 1;     // parity to wait for on bar_done[dslot]: stage st - kTnLoStages
    for (int64_t st = 0; st < n_stages; ++st) {
      cp_async_wait<kTnCopyDepth - 1>();                       // this thread's copies of stage `st` have landed
      // lo slot st % kTnLoStages and hi slot (st + kTnCopyDepth) % kTnHiStages were both last used by stage
      // st - kTnLoStages: one wait covers the split's output buffer and the next copy's landing buffer
      if (st >= kTnLoStages) mbar_wait(bar_done + 8 * dslot, dphase);
      const uint32_t g_hi = ghi + (uint32_t)shi * L::kGBytes, g_lo = glo + (uint32_t)slo * L::kGBytes;
      if (!(ablate & 8)) {
        // all loads first: ptxas otherwise chains lds -> residual -> sts one chunk at a time (the profile showed the
        // G warps 100 % busy, a third of it waiting for one LDS after the other, and everyone else waiting for them)
        float4 v[kGPer];
#pragma unroll
        for (int i = 0; i < kGPer; ++i) v[i] = lds128(g_hi + offg[i]);
#pragma unroll
        for (int i = 0; i < kGPer; ++i) {
          if constexpr (SG) {
            sum_g.x = __fadd_rn(sum_g.x, v[i].x); sum_g.y = __fadd_rn(sum_g.y, v[i].y);
            sum_g.z = __fadd_rn(sum_g.z, v[i].z); sum_g.w = __fadd_rn(sum_g.w, v[i].w);
          }
          sts128(g_lo + offg[i], make_float4(tf32_trunc_residual(v[i].x), tf32_trunc_residual(v[i].y),
                                             tf32_trunc_residual(v[i].z), tf32_trunc_residual(v[i].w)));
        }
      }
      if (!(ablate & 1)) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * fslot);
      if (++shi == kTnHiStages) shi = 0;
      if (++slo == kTnLoStages) slo = 0;
      if (++fslot == kTnASlots) fslot = 0;
      if (++dslot == kTnASlots) { dslot = 0; dphase ^= 1; }
      if (st + kTnCopyDepth < n_stages) issue(st + kTnCopyDepth);
      cp_async_commit();
    }
    if constexpr (SG) {
      // threads with equal (gt % chunks-per-row) own the same columns: with 32 chunks per row that is one lane of each
      // of the 4 warps; with 16 chunks per row (64 features) lanes l and l+16 of every warp as well
      const int gw = gt >> 5;
      float* sg = sum_scratch + gw * 128;
      if (N == 64) {
        sum_g.x += __shfl_down_sync(0xffffffffu, sum_g.x, 16); sum_g.y += __shfl_down_sync(0xffffffffu, sum_g.y, 16);
        sum_g.z += __shfl_down_sync(0xffffffffu, sum_g.z, 16); sum_g.w += __shfl_down_sync(0xffffffffu, sum_g.w, 16);
      }
      if (lane < N / 4) *reinterpret_cast<float4*>(sg + 4 * lane) = sum_g;
      asm volatile("bar.sync 1, %0;" ::"n"(kTnGThreads) : "memory");
      if (gt < N) {
        float t = 0.0f;
        for (int w = 0; w < kTnGWarps; ++w) t = __fadd_rn(t, sum_scratch[w * 128 + gt]);
        p.part_sg[(int64_t)blockIdx.x * N + gt] = t;
      }
    }
  } else if (warp == kTnMmaWarp) {
    // =========================== MMA ISSUER ===========================
    const uint32_t idesc = make_idesc_mn(128, N) ^ p.idesc_xor;
    const uint32_t ghi = base + L::kGHiOff, glo = base + L::kGLoOff;
    int shi = 0, slo = 0, aslot = 0;
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
        mbar_wait(bar_full + 8 * aslot, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t g_hi = ghi + (uint32_t)shi * L::kGBytes, g_lo = glo + (uint32_t)slo * L::kGBytes;
          const uint32_t a_hi = tmem_base + kTnACol + (uint32_t)(aslot * kTnAColsPerSlot), a_lo = a_hi + kTnEdges;
#pragma unroll
          for (int j = 0; j < kTnEdges / 8; ++j) {
            if (ablate & 2) break;
            const uint64_t dgh = smem_desc_mn32(g_hi + j * p.kadv, p.lbo, p.sbo, p.ltype);
            const uint64_t dgl = smem_desc_mn32(g_lo + j * p.kadv, p.lbo, p.sbo, p.ltype);
            // small terms first, the dominant hi*hi product last
            umma_tf32_ts(d_tmem, a_lo + j * 8, dgh, idesc, (first && j == 0) ? 0u : 1u);
            umma_tf32_ts(d_tmem, a_hi + j * 8, dgl, idesc, 1u);
            umma_tf32_ts(d_tmem, a_hi + j * 8, dgh, idesc, 1u);
          }
          umma_commit(bar_done + 8 * aslot);
          if (st == st_hi - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        first = false;
        if (++shi == kTnHiStages) shi = 0;
        if (++slo == kTnLoStages) slo = 0;
        if (++aslot == kTnASlots) { aslot = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kTnXWarps && warp < kTnMmaWarp) {
    // =========================== FLUSH (TMEM -> fp32 partial in global) ===========================
    // TMEM lane = m (feature of X), column = n (feature of G).  Warp w: lane quadrant w % 4, all N columns.
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    float* part = p.partial + (int64_t)blockIdx.x * M * N + m;   // element (m, n) at n*M + m: lanes contiguous
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t f = 0; f < n_flush; ++f) {
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(t_lane + c0, v);
        if (m < M && !(ablate & 16)) {
          float* q = part + (int64_t)c0 * M;
          // Each element of the partial is only ever touched by this thread, so a fire-and-forget fp32 reduction at L2
          // (round-to-nearest, program order per address) gives the same sum as load + add + store for half the L2
          // traffic and no round trip -- the read-modify-write form cost 0.9 of this kernel's 8.5 ms.
          if (f != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q + j * M), "f"(v[j]) : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j * M] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (n_flush == 0 && m < M) {                      // CTA without work still owns a (zero) partial
      for (int n = 0; n < N; ++n) part[(int64_t)n * M] = 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTnMmaWarp) tmem_dealloc(tmem_base, kTnTmemCols);
}

// D[m, n] (+)= sum over CTAs of partial[cta][n][m]   (and the column sums), in a FIXED order: lane j of a group of four
// adds the partials j, j+4, j+8, ... in ascending order, then (s0 + s1) + (s2 + s3).  The cost of this kernel is the
// chain of dependent L2 round trips, not the adds (148 partials one after the other: 9.6 us; batches of 8: 7 us, a
// tenth of a small-graph training step), so every output gets four lanes with 37 independent loads each.
constexpr int kRedLanes = 4;
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ partial, int parts, int M, int N,
                                                        float* __restrict__ D, int64_t ldd, int accumulate,
                                                        const float* __restrict__ part_sx, float* __restrict__ sum_x,
                                                        const float* __restrict__ part_sg, float* __restrict__ sum_g) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gid / kRedLanes, j = gid % kRedLanes;     // output (m fastest: neighbouring groups read one sector), lane
  const float* src = nullptr;
  float* dst = nullptr;
  int64_t step = 0;
  bool acc = false;
  if (idx < M * N) {
    src = partial + idx; step = (int64_t)M * N;
    dst = D + (int64_t)(idx % M) * ldd + idx / M;
    acc = accumulate != 0;
  } else {
    const int t = idx - M * N;                                // the column sums
    if (t < M && sum_x != nullptr) { src = part_sx + t; dst = sum_x + t; step = M; }
    else if (t >= M && t < M + N && sum_g != nullptr) { src = part_sg + (t - M); dst = sum_g + (t - M); step = N; }
  }
  float s = 0.0f;
  if (src != nullptr) {
    int c = j;
    for (; c + 7 * kRedLanes < parts; c += 8 * kRedLanes) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(c + u * kRedLanes) * step);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
    }
    for (; c < parts; c += kRedLanes) s = __fadd_rn(s, __ldcg(src + (int64_t)c * step));
  }
  s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
  s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
  if (src != nullptr && j == 0) *dst = acc ? __fadd_rn(*dst, s) : s;
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
  const int sms = persistent_sms();
  int64_t grid = stages < sms ? stages : sms;
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
  p.ablate = 0;
#ifdef DMP_DEBUG
  // debug overrides, read once per process: descriptor geometry (scripts/tn_probe.py) and ablation (scripts/tn_ablate.py);
  // not compiled into the release library
  static const char* const dbg = getenv("DMP_TN_DBG");
  static const int ablate = getenv("DMP_TN_ABLATE") ? atoi(getenv("DMP_TN_ABLATE")) : 0;
  if (dbg) sscanf(dbg, "%u,%u,%u,%u,%u", &p.lbo, &p.sbo, &p.kadv, &p.idesc_xor, &p.ltype);
  p.ablate = ablate;
#endif
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (M == 128 && N == 128) rc = launch_tn<128, 128>(p, (unsigned)grid, s);
  else if (M == 128 && N == 64) rc = launch_tn<128, 64>(p, (unsigned)grid, s);
  else if (M == 64 && N == 128) rc = launch_tn<64, 128>(p, (unsigned)grid, s);
  else rc = launch_tn<64, 64>(p, (unsigned)grid, s);
  if (rc != DMP_OK) return rc;
  const int total = (int)(M * N + M + N);
  tn_reduce_kernel<<<(total * kRedLanes + 255) / 256, 256, 0, s>>>(p.partial, (int)grid, (int)M, (int)N, D, ldd, accumulate,
                                                      p.part_sx, colsum_x, p.part_sg, colsum_g);
  return launch_status("tn_reduce_kernel");
}
