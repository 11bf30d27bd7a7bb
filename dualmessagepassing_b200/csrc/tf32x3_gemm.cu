// tf32x3_gemm.cu -- edge-sized projection GEMM on the 5th-gen tensor cores with fp32-level accuracy.
//
//   D[M,N] = epilogue( (row_scale ⊙ A)[M,K] · Bt[N,K]^T )          K, N in {64, 128}, M = number of edges/nodes
//
// Why: in strict fp32 the DMPNN layer is bound by its 15 edge-sized [E,H]x[H,H] projections (cuBLAS sgemm:
// 48.7 TFLOP/s on B200, 5.4 ms per 8 M rows), not by the sparse core.  A single-pass TF32 GEMM is HBM-bound
// (1.3 ms) but has 3e-4 error.  Here each fp32 operand is split as x = hi + lo with hi = tf32(x) and
// three tcgen05.mma (kind::tf32) products hi·hi + lo·hi + hi·lo accumulate in fp32 in TMEM: measured error
// vs fp64 equals cuBLAS sgemm's (~5e-7 max-norm relative), at tensor-core speed.
//
// Structure (one persistent CTA per SM, 17 warps, no TMA descriptors -- the operand transform needs the
// data in registers anyway):
//   warps 9..16  PRODUCERS  global fp32 (coalesced 128-bit, 3 k-blocks of loads in flight per thread)
//                           -> optional row scale (fuses `coef ⊙ gE`, dmpnn.py:146 backward)
//                           -> hi/lo split -> 128B-swizzled K-major smem tiles -> fence.proxy.async -> mbarrier
//   warp  8      MMA        one elected lane issues 3 x (32/8) tcgen05.mma per k-block, tcgen05.commit
//                           releases the smem stage / publishes the accumulator
//   warps 0..7   EPILOGUE   tcgen05.ld 32x32b -> bias / activation / act' / accumulate -> global.  For N = 128 the
//                           MMA is issued TRANSPOSED (weights as the M=128 operand, the edge tile as the N operand) so
//                           that a TMEM lane is an output FEATURE: the 32 lanes of a warp then store 32 consecutive
//                           floats of one output row (one 128-byte wavefront per store instead of 32)
// The weight matrix (<= 64 KB) is split once per CTA and stays resident in smem; accumulators are double
// buffered in TMEM (2 x N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "tc_common.cuh"

namespace dmp {
namespace gemm {

// tf32x3_gemm_v2.cu: 64-row tiles, cross-terms-first MMA order (lower truncation error); every product whose row scale
// (if any) rides on the accumulator goes there; this file keeps producer-side row scales and the gather epilogue
int launch_gemm_v2(const float* A, int64_t lda, const float* Wt, int64_t ldw, const float* scale, const float* bias,
                   const float* aux, int64_t ld_aux, float* D, int64_t ldd, int64_t M, int64_t N, int64_t K, int mode,
                   int act, float slope, cudaStream_t stream);

constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueWarps = 8;        // two warps per TMEM lane quadrant, each takes half of the columns
constexpr int kMmaWarp = 8;
constexpr int kThreadsGemm = (kEpilogueWarps + 1 + kProducerWarps) * 32;  // 544
constexpr bool kTmaDefault = true;      // streamed operand by TMA unless DMP_GEMM_TMA says otherwise
constexpr int kPrefetch = 3;            // k-blocks of global loads kept in flight per producer thread

// epilogue flags (low 4 bits = DMP_ACT_*)
constexpr int kEpiMulActGradFromOutput = 32;  // D = acc * act'(aux) with aux = activation OUTPUT
constexpr int kEpiAccumulate = 64;            // D += acc

struct GemmParams {
  const float* A; int64_t lda;
  const float* row_scale;
  int use_tma;              // N == 128: the streamed operand arrives by TMA (one cp.async.bulk.tensor per k-block) instead
                            // of 1 024 per-thread cp.async
  const float* epi_scale;   // accumulate mode, N == 128: the row scale is applied to the ACCUMULATOR (D += s_r * acc_r)
                            // instead of to the streamed operand -- same product, and the producers keep their fast path
  const float* Bt; int64_t ldb;
  const float* bias;
  const float* aux; int64_t ld_aux;
  float* D; int64_t ldd;
  int64_t M;
  int epilogue;
  float slope;
  // kModeAccGather: D += acc + sgn_r * norm_r * tab_{rev_r}[dst32[r], :]
  const int32_t* g_dst;
  const uint8_t* g_rev;
  const float* g_norm;
  const float* g_tab0;
  const float* g_tab1;
  int64_t ld_tab;
};

template <int N, int K>
struct Smem {
  static constexpr int kKBlocks = K / kKB;
  // N == 128 ("TS" form): the split weights live in TENSOR MEMORY as the MMA's A operand -- no smem copy, which
  // halves the tensor core's smem read traffic and leaves room for 7 instead of 3 operand stages
  static constexpr bool kTS = (N == 128);
  static constexpr int kNumStages = kTS ? 7 : 3;
  static constexpr int kBBlockBytes = N * 128;                 // one k-block of B (hi or lo)
  static constexpr int kBBytes = kTS ? 0 : 2 * kKBlocks * kBBlockBytes;  // hi + lo
  static constexpr int kABlockBytes = kTileM * 128;            // 16 KB
  static constexpr int kStageBytes = 2 * kABlockBytes;         // hi + lo
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kBBytes + kNumStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
};

// Epilogue modes (compile-time, so that the fully unrolled TMEM->global loop stays small and branch-free;
// the first version kept them as run-time flags and the 11 k-instruction kernel thrashed the I-cache):
//   piecewise-linear activations are ONE code path  y = x > 0 ? x : slope*x   (none: slope 1, relu: slope 0)
enum : int {
  kModeStore = 0,        // D = acc
  kModeAccumulate = 1,   // D += acc
  kModeBiasPwl = 2,      // D = pwl(acc + bias)
  kModeGradPwl = 3,      // D = acc * (aux > 0 ? 1 : slope)
  kModeBiasSmooth = 4,   // D = tanh|sigmoid(acc + bias)
  kModeGradSmooth = 5,   // D = acc * d tanh|sigmoid expressed through aux = activation output
  kModeAccGather = 6,    // D += acc + sgn_r * norm_r * table_{rev_r}[dst_r, :]   (backward of fn.sum folded into dX_e)
};

template <int MODE>
__device__ __forceinline__ float epilogue_op(float acc, float bias, float aux, float old, float slope, int act) {
  if constexpr (MODE == kModeStore) return acc;
  if constexpr (MODE == kModeAccumulate || MODE == kModeAccGather) return __fadd_rn(old, acc);
  if constexpr (MODE == kModeBiasPwl) {
    const float x = __fadd_rn(acc, bias);
    return x > 0.0f ? x : __fmul_rn(x, slope);
  }
  if constexpr (MODE == kModeGradPwl) return __fmul_rn(acc, aux > 0.0f ? 1.0f : slope);
  if constexpr (MODE == kModeBiasSmooth) return apply_act(__fadd_rn(acc, bias), act, slope);
  return __fmul_rn(acc, act_grad_from_output(aux, act, slope));
}

template <int N, int K, int MODE>
__global__ void __launch_bounds__(kThreadsGemm, 1) tf32x3_gemm_kernel(const GemmParams p,
                                                                      const __grid_constant__ CUtensorMap tmap) {
  using L = Smem<N, K>;
  constexpr int kKBlocks = L::kKBlocks;
  constexpr int kStages = L::kNumStages;
  constexpr bool kTS = L::kTS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = base;                                  // [hi kb0..][lo kb0..]
  const uint32_t sA = base + L::kBBytes;                     // stage s: [hi 16K][lo 16K]
  const uint32_t sBar = sA + kStages * L::kStageBytes;
  const uint32_t bar_full = sBar;                            // kStages x 8 B
  const uint32_t bar_empty = sBar + 8 * kStages;
  const uint32_t bar_acc_full = sBar + 16 * kStages;         // 2 x 8 B
  const uint32_t bar_acc_empty = bar_acc_full + 16;
  const uint32_t tmem_slot = bar_acc_empty + 16;
  const uint32_t bar_raw = tmem_slot + 8;                    // kStages x 8 B: TMA completion of the raw (hi) tile
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.M + kTileM - 1) / kTileM;
  // ablation switches for performance triage (scripts/gemm_ablate.py): compiled in only with -DDMP_DEBUG; the release
  // library rejects the corresponding epilogue bits (dmp_gemm_tf32x3) and these fold to constants
#ifdef DMP_DEBUG
  const bool dbg_no_ldg = (p.epilogue >> 8) & 1, dbg_no_sts = (p.epilogue >> 9) & 1;
  const bool dbg_no_mma = (p.epilogue >> 10) & 1, dbg_no_stg = (p.epilogue >> 11) & 1;
  const bool dbg_no_fence = (p.epilogue >> 12) & 1, dbg_spin = (p.epilogue >> 13) & 1, dbg_no_tld = (p.epilogue >> 14) & 1;
  long long* dbg_ts = ((p.epilogue >> 16) & 1) && blockIdx.x == 0 ? (long long*)p.aux : nullptr;  // [3][256][2]
#else
  constexpr bool dbg_no_ldg = false, dbg_no_sts = false, dbg_no_mma = false, dbg_no_stg = false, dbg_no_fence = false,
                 dbg_spin = false, dbg_no_tld = false;
  constexpr long long* dbg_ts = nullptr;
#endif
#define MBAR_WAIT(bar, par) do { if (dbg_spin) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)

  // ---- one-time setup: barriers, TMEM, resident split weights ---------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, kProducerWarps);        // one arrival per producer warp
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_raw + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kEpilogueWarps);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  // TMEM map (TS): [0,128) acc 0 | [128,256) acc 1 | [256,256+K) W_hi | [256+K,256+2K) W_lo   (lane = feature)
  constexpr int kTmemCols = kTS ? 512 : 2 * N;
  constexpr uint32_t kWhiCol = 256, kWloCol = 256 + K;
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  if constexpr (!kTS) {
    for (int c = threadIdx.x; c < N * (K / 4); c += kThreadsGemm) {  // 16-byte chunks of Bt[N,K]
      const int n = c / (K / 4);
      const int k4 = c % (K / 4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.Bt + (int64_t)n * p.ldb + k4 * 4));
      const int kb = k4 / 8, c16 = k4 % 8;
      const uint32_t off = (uint32_t)kb * L::kBBlockBytes + swz(n, c16);
      split_store(sB + off, sB + kKBlocks * L::kBBlockBytes + off, v);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if constexpr (kTS) {
    if (warp < 4) {   // thread = weight row (output feature) = TMEM lane; 32 k-values per tcgen05.st
      const int f = warp * 32 + lane;
      const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < K; c0 += 32) {
        float hi[32], lo[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p.Bt + (int64_t)f * p.ldb + c0 + 4 * q));
          hi[4 * q + 0] = tf32_rna(v.x); lo[4 * q + 0] = __fsub_rn(v.x, hi[4 * q + 0]);
          hi[4 * q + 1] = tf32_rna(v.y); lo[4 * q + 1] = __fsub_rn(v.y, hi[4 * q + 1]);
          hi[4 * q + 2] = tf32_rna(v.z); lo[4 * q + 2] = __fsub_rn(v.z, hi[4 * q + 2]);
          hi[4 * q + 3] = tf32_rna(v.w); lo[4 * q + 3] = __fsub_rn(v.w, hi[4 * q + 3]);
        }
        tmem_st32(t_lane + kWhiCol + c0, hi);
        tmem_st32(t_lane + kWloCol + c0, lo);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp > kMmaWarp) {
    // =========================== PRODUCERS ===========================
    const int pt = threadIdx.x - (kMmaWarp + 1) * 32;          // 0..255
    // thread handles chunks c = pt + 256*i (i<4): row = c/8, 16B-chunk = c%8  (a warp covers 4 full rows)
    const int c16 = pt & 7;
    const int row0 = pt >> 3;                                   // rows row0 + 32*i
    const int64_t total_kb = ((num_tiles - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x) * kKBlocks;
    if constexpr (kTS) {
      // ---- cp.async producer (N == 128): raw fp32 rows go global -> smem asynchronously (kCopyDepth k-blocks =
      // 80 KB per SM in flight, no register staging); the raw tile IS the hi operand (a kind::tf32 MMA ignores the low
      // 13 mantissa bits), the producer only adds lo = x - trunc_tf32(x).  ~2x fewer instructions per stage than the
      // register path, which was issue/latency bound (ncu: producers 58 % busy, 14 % waiting on loads).
      // (separate hi / lo rings, 9 + 4 slots, as in tf32x3_gemm_tn.cu were tried here: store and grad modes got 12 %
      // SLOWER -- with the bursty epilogue holding the accumulators, the 7 combined slots of split data ahead of the
      // MMA matter more than the extra round-trip slack)
      constexpr int kCopyDepth = kStages - 2;
      uint32_t offs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) offs[i] = swz(row0 + 32 * i, c16);
      int istage = 0;
      uint32_t iphase = 0;
      // per-row scales of the tile (same 4 rows for all k-blocks): fetched when the tile's first k-block is ISSUED
      // (kCopyDepth stages ahead of its use), double-buffered by tile parity so the fetch latency is never exposed
      float sc_a[4] = {1.f, 1.f, 1.f, 1.f}, sc_b[4] = {1.f, 1.f, 1.f, 1.f};
      // source of this thread's first chunk of the k-block being issued, advanced incrementally (no 64-bit multiplies
      // per copy): + kKB floats per k-block, + gridDim.x tiles after the last k-block of a tile
      const float* asrc = p.A + ((int64_t)blockIdx.x * kTileM + row0) * p.lda + c16 * 4;
      const int64_t lda32 = 32 * p.lda;
      const int64_t tile_adv = (int64_t)gridDim.x * kTileM * p.lda - (kKBlocks - 1) * kKB;
      int64_t irow0 = (int64_t)blockIdx.x * kTileM;    // first row of the tile being issued
      int ikb = 0;
      uint32_t itl = 0;                                // parity of the local tile index
      auto issue = [&](int64_t) {
        MBAR_WAIT(bar_empty + 8 * istage, iphase ^ 1);
        const uint32_t hi = sA + istage * L::kStageBytes;
        if (p.use_tma) {
          // one elected thread: expected bytes on the stage's mbarrier, then ONE bulk tensor copy of the 128 x 32 box
          // (rows past M are zero-filled by the TMA unit, the 128-byte swizzle is applied by it too)
          if (pt == 0) {
            mbar_expect_tx(bar_raw + 8 * istage, L::kABlockBytes);
            tma_load_2d(hi, &tmap, ikb * kKB, (int)irow0, bar_raw + 8 * istage);
          }
        } else if (irow0 + kTileM <= p.M) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async16(hi + offs[i], asrc + i * lda32, 16u);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = irow0 + row0 + 32 * i < p.M;
            cp_async16(hi + offs[i], ok ? (const void*)(asrc + i * lda32) : (const void*)p.A, ok ? 16u : 0u);
          }
        }
        if (ikb == 0 && p.row_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int64_t r = irow0 + row0 + 32 * i;
            const float v = r < p.M ? __ldg(p.row_scale + r) : 1.0f;
            if (itl & 1) sc_b[i] = v; else sc_a[i] = v;
          }
        }
        if (++ikb == kKBlocks) { ikb = 0; asrc += tile_adv; irow0 += (int64_t)gridDim.x * kTileM; itl ^= 1; }
        else asrc += kKB;
        if (++istage == kStages) { istage = 0; iphase ^= 1; }
      };
      auto l2_prefetch_tile = [&](int64_t local_tile) {
        const int64_t tile = (int64_t)blockIdx.x + local_tile * gridDim.x;
        if (pt == 0 && tile < num_tiles) {
          const int64_t r0 = tile * kTileM;
          const int64_t rows = (p.M - r0) < kTileM ? (p.M - r0) : kTileM;
          prefetch_l2_bulk(p.A + r0 * p.lda, (uint32_t)(((rows - 1) * p.lda + K) * 4));
        }
      };
      // ncu (profiles/r1_ncu_summary_cfg5.md): with the bulk L2 prefetch on, DRAM reads were 1.8x the operand size
      // (prefetched lines were fetched again by the cp.async that followed); off by default.
      constexpr int kL2Ahead = 4;
      constexpr bool kUseL2Prefetch = false;
      if (kUseL2Prefetch) for (int t = 2; t <= kL2Ahead; ++t) l2_prefetch_tile(t);
#pragma unroll
      for (int d = 0; d < kCopyDepth; ++d) {
        if (d < total_kb) issue(d);
        cp_async_commit();
      }
      int stage = 0;
      uint32_t rphase = 0;
      for (int64_t it = 0; it < total_kb; ++it) {
        if (kUseL2Prefetch && it % kKBlocks == 0) l2_prefetch_tile(it / kKBlocks + kL2Ahead + 1);
        if (p.use_tma) MBAR_WAIT(bar_raw + 8 * stage, rphase);   // the bulk copy of k-block `it` has landed
        else cp_async_wait<kCopyDepth - 1>();                 // this thread's copies of k-block `it` have landed
        const uint32_t hi = sA + stage * L::kStageBytes;
        const uint32_t lo = hi + L::kABlockBytes;
        if (p.row_scale != nullptr) {
          const bool odd = ((it / kKBlocks) & 1) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float sc = odd ? sc_b[i] : sc_a[i];
            float4 v = lds128(hi + offs[i]);
            v.x = __fmul_rn(sc, v.x); v.y = __fmul_rn(sc, v.y); v.z = __fmul_rn(sc, v.z); v.w = __fmul_rn(sc, v.w);
            sts128(hi + offs[i], v);
            sts128(lo + offs[i], make_float4(tf32_trunc_residual(v.x), tf32_trunc_residual(v.y),
                                             tf32_trunc_residual(v.z), tf32_trunc_residual(v.w)));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = lds128(hi + offs[i]);
            sts128(lo + offs[i], make_float4(tf32_trunc_residual(v.x), tf32_trunc_residual(v.y),
                                             tf32_trunc_residual(v.z), tf32_trunc_residual(v.w)));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * stage);
        if (++stage == kStages) { stage = 0; rphase ^= 1; }
        if (it + kCopyDepth < total_kb) issue(it + kCopyDepth);
        cp_async_commit();
      }
    } else {
      float4 buf[kPrefetch][4];
      float scale_buf[kPrefetch][4];
      int64_t it_load = 0;
      auto load_block = [&](int64_t it, float4 (&dst)[4], float (&sc)[4]) {
        const int64_t tile = (int64_t)blockIdx.x + (it / kKBlocks) * gridDim.x;
        const int kb = (int)(it % kKBlocks);
  #pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = tile * kTileM + row0 + 32 * i;
          if (r < p.M && !dbg_no_ldg) {
            const float* src = p.A + r * p.lda + kb * kKB + c16 * 4;
            float4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src));
            dst[i] = v;
            sc[i] = p.row_scale != nullptr ? __ldg(p.row_scale + r) : 1.0f;
          } else {
            dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            sc[i] = 1.0f;
          }
        }
      };
  #pragma unroll
      for (int slot = 0; slot < kPrefetch; ++slot)
        if (slot < total_kb) {
          load_block(slot, buf[slot], scale_buf[slot]);
          ++it_load;
        }
      int stage = 0;
      uint32_t phase = 0;
      constexpr int kL2Ahead = 3;  // tiles of this CTA kept on their way into L2
      auto l2_prefetch_tile = [&](int64_t local_tile) {
        // ONE thread, ONE bulk prefetch per tile (the tile's rows form one contiguous range of rows*lda floats).
        // UBLKPF takes uniform operands: issued from many lanes the compiler serialises it lane by lane.
        const int64_t tile = (int64_t)blockIdx.x + local_tile * gridDim.x;
        if (pt == 0 && tile < num_tiles) {
          const int64_t r0 = tile * kTileM;
          const int64_t rows = (p.M - r0) < kTileM ? (p.M - r0) : kTileM;
          prefetch_l2_bulk(p.A + r0 * p.lda, (uint32_t)(((rows - 1) * p.lda + K) * 4));
        }
      };
      for (int t = 1; t <= kL2Ahead; ++t) l2_prefetch_tile(t);
      for (int64_t it0 = 0; it0 < total_kb; it0 += kPrefetch) {
  #pragma unroll
        for (int slot = 0; slot < kPrefetch; ++slot) {  // compile-time slot: the prefetch buffers stay in registers
          if (it0 + slot < total_kb) {
            if ((it0 + slot) % kKBlocks == 0) l2_prefetch_tile((it0 + slot) / kKBlocks + kL2Ahead + 1);
            MBAR_WAIT(bar_empty + 8 * stage, phase ^ 1);
            if (dbg_ts && pt == 0 && it0 + slot < 256) dbg_ts[2 * (it0 + slot)] = clock64();
            const uint32_t hi = sA + stage * L::kStageBytes;
            const uint32_t lo = hi + L::kABlockBytes;
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 v = buf[slot][i];
              if (p.row_scale != nullptr) {
                const float s = scale_buf[slot][i];
                v.x = __fmul_rn(s, v.x); v.y = __fmul_rn(s, v.y); v.z = __fmul_rn(s, v.z); v.w = __fmul_rn(s, v.w);
              }
              const uint32_t off = swz(row0 + 32 * i, c16);
              if (!dbg_no_sts) split_store(hi + off, lo + off, v);
            }
            // every writer fences its own generic-proxy stores towards the async proxy, the warp converges, and
            // ONE lane arrives: 8 smem atomics per stage instead of 256 (the per-thread version cost 1.3 ms / 8 M rows)
            if (!dbg_no_fence) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * stage);
            if (dbg_ts && pt == 0 && it0 + slot < 256) dbg_ts[2 * (it0 + slot) + 1] = clock64();
            if (it_load < total_kb) {
              load_block(it_load, buf[slot], scale_buf[slot]);
              ++it_load;
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA ISSUER ===========================
    // TRANSPOSED (N == 128): D^T[feature, edge] = W[feature, k] * X[edge, k]^T  (M = 128 features, N = 128 edges)
    constexpr bool kT = (N == 128);
    constexpr uint32_t idesc = kT ? make_idesc(128, kTileM) : make_idesc(kTileM, N);
    constexpr int kAccCols = kT ? kTileM : N;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int dbg_it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      MBAR_WAIT(bar_acc_empty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
      for (int kb = 0; kb < kKBlocks; ++kb) {
        MBAR_WAIT(bar_full + 8 * stage, phase);
        if (dbg_ts && lane == 0 && dbg_it < 256) dbg_ts[512 + 2 * dbg_it] = clock64();
        tc_fence_after();
        if (lane == 0 && dbg_no_mma) {
          if ((p.epilogue >> 15) & 1) {   // plain arrivals instead of tcgen05.commit (measures commit latency)
            mbar_arrive(bar_empty + 8 * stage);
            if (kb == kKBlocks - 1) mbar_arrive(bar_acc_full + 8 * acc);
          } else {
            umma_commit(bar_empty + 8 * stage);
            if (kb == kKBlocks - 1) umma_commit(bar_acc_full + 8 * acc);
          }
        }
        if (lane == 0 && !dbg_no_mma) {
          const uint32_t a_hi = sA + stage * L::kStageBytes;
          const uint32_t a_lo = a_hi + L::kABlockBytes;
          const uint32_t b_hi = sB + kb * L::kBBlockBytes;
          const uint32_t b_lo = b_hi + kKBlocks * L::kBBlockBytes;
#pragma unroll
          for (int j = 0; j < kKB / 8; ++j) {
            const uint64_t dah = make_smem_desc(a_hi + j * 32);
            const uint64_t dal = make_smem_desc(a_lo + j * 32);
            // small terms first, the dominant hi*hi product last
            if constexpr (kT) {
              // weights = A operand read from TMEM (8 columns per k-step), edge tile = B operand from smem
              const uint32_t w_hi = tmem_base + kWhiCol + (uint32_t)(kb * kKB + j * 8);
              const uint32_t w_lo = tmem_base + kWloCol + (uint32_t)(kb * kKB + j * 8);
              umma_tf32_ts(d_tmem, w_hi, dal, idesc, (kb | j) != 0 ? 1u : 0u);
              umma_tf32_ts(d_tmem, w_lo, dah, idesc, 1u);
              umma_tf32_ts(d_tmem, w_hi, dah, idesc, 1u);
            } else {
              const uint64_t dbh = make_smem_desc(b_hi + j * 32);
              const uint64_t dbl = make_smem_desc(b_lo + j * 32);
              umma_tf32(d_tmem, dal, dbh, idesc, (kb | j) != 0 ? 1u : 0u);
              umma_tf32(d_tmem, dah, dbl, idesc, 1u);
              umma_tf32(d_tmem, dah, dbh, idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * stage);                 // smem stage free once these MMAs retire
          if (kb == kKBlocks - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        if (dbg_ts && lane == 0 && dbg_it < 256) dbg_ts[512 + 2 * dbg_it + 1] = clock64();
        ++dbg_it;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== EPILOGUE ===========================
    constexpr bool kNeedAux = (MODE == kModeGradPwl || MODE == kModeGradSmooth);
    constexpr bool kNeedBias = (MODE == kModeBiasPwl || MODE == kModeBiasSmooth);
    const int act = p.epilogue & 15;
    int acc = 0;
    uint32_t acc_phase = 0;
    int dbg_tile = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if constexpr (N == 128 && MODE == kModeAccGather) {
        // Gather epilogue, register-lean and latency-aware: 16 rows at a time; the operands that do not depend on
        // the accumulator (previous D, row metadata, gathered table rows) are requested BEFORE waiting for the MMAs
        // of this tile, so their DRAM latency hides behind the tensor-core work.
        const int quad = warp & 3, half = warp >> 2;
        const int f = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileM + half * 64);
        {
          // The epilogue's own loads keep only ~32 KB per SM in flight (register bound), too little for random
          // 128-byte table rows straight from DRAM.  So pull the NEXT tile's operands towards L2 now: lane j handles
          // row j of this warp's 64-row half, each lane prefetches the one line its warp will read.
          const int64_t nt = tile + gridDim.x;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int64_t r = nt * kTileM + half * 64 + hh * 32 + lane;
            if (nt < num_tiles && r < p.M) {
              const int d = __ldg(p.g_dst + r);
              const int rv = p.g_rev != nullptr ? (int)__ldg(p.g_rev + r) : 0;
              asm volatile("prefetch.global.L2 [%0];" ::"l"((rv ? p.g_tab1 : p.g_tab0) + (int64_t)d * p.ld_tab + quad * 32));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.D + r * p.ldd + quad * 32));
            }
          }
        }
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
          const int64_t r0 = tile * kTileM + half * 64 + sub * 16;
          const int nvalid = (int)((p.M - r0) < 16 ? (p.M - r0) : 16);     // warp-uniform, may be <= 0
          float* dst = p.D + r0 * p.ldd + f;
          float old[16], g[16];
          int d_l = 0, rv_l = 0;
          float w_l = 1.0f;
          if (lane < nvalid) {
            d_l = __ldg(p.g_dst + r0 + lane);
            if (p.g_rev != nullptr) rv_l = (int)__ldg(p.g_rev + r0 + lane);
            if (p.g_norm != nullptr) w_l = __ldg(p.g_norm + r0 + lane);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int d = __shfl_sync(0xffffffffu, d_l, j);
            const int rv = __shfl_sync(0xffffffffu, rv_l, j);
            if (j < nvalid) {
              old[j] = dst[j * p.ldd];
              g[j] = __ldg((rv ? p.g_tab1 : p.g_tab0) + (int64_t)d * p.ld_tab + f);
            }
          }
          if (sub == 0) {
            MBAR_WAIT(bar_acc_full + 8 * acc, acc_phase);
            tc_fence_after();
          }
          float v[16];
          tmem_ld16(t_lane + sub * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nvalid) {
              const int rv = __shfl_sync(0xffffffffu, rv_l, j);
              float x = g[j];
              if (p.g_norm != nullptr) x = __fmul_rn(x, __shfl_sync(0xffffffffu, w_l, j));
              dst[j * p.ldd] = __fadd_rn(__fadd_rn(old[j], v[j]), rv ? x : -x);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        ++dbg_tile;
        continue;
      }
      if constexpr (N == 128 && (kNeedAux || MODE == kModeAccumulate)) {
        // Epilogues with a streamed operand (previous D for accumulate, activation output for act'): one 32-row chunk
        // at a time (64 live registers instead of 96: no spills), the operand of the first chunk is requested BEFORE
        // waiting for this tile's MMAs so that its DRAM latency hides behind the tensor-core work.
        const int quad = warp & 3, half = warp >> 2;
        const int f = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileM + half * 64);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int64_t r0 = tile * kTileM + half * 64 + c * 32;
          const int nvalid = (int)((p.M - r0) < 32 ? (p.M - r0) : 32);     // warp-uniform, may be <= 0
          float* dst = p.D + r0 * p.ldd + f;
          const float* src = kNeedAux ? p.aux + r0 * p.ld_aux + f : dst;
          const int64_t lds = kNeedAux ? p.ld_aux : p.ldd;
          float t[32];
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = src[j * lds];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = j < nvalid ? src[j * lds] : 0.0f;
          }
          float sc_l = 1.0f;       // lane j holds the scale of row r0 + j; broadcast by shuffle below
          if (MODE == kModeAccumulate && p.epi_scale != nullptr && lane < nvalid) sc_l = __ldg(p.epi_scale + r0 + lane);
          if (c == 0) {
            MBAR_WAIT(bar_acc_full + 8 * acc, acc_phase);
            tc_fence_after();
          }
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          if (MODE == kModeAccumulate && p.epi_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(__shfl_sync(0xffffffffu, sc_l, j), v[j]);
          }
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * p.ldd] = epilogue_op<MODE>(v[j], 0.0f, t[j], t[j], p.slope, act);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) dst[j * p.ldd] = epilogue_op<MODE>(v[j], 0.0f, t[j], t[j], p.slope, act);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        ++dbg_tile;
        continue;
      }
      MBAR_WAIT(bar_acc_full + 8 * acc, acc_phase);
      if (dbg_ts && threadIdx.x == 0 && dbg_tile < 256) dbg_ts[1024 + 2 * dbg_tile] = clock64();
      tc_fence_after();
      if constexpr (N == 128) {
        // TMEM lane == output feature; the 32 registers of one tcgen05.ld == 32 consecutive edges (tile rows):
        // the 32 lanes of a warp store 32 consecutive floats of ONE output row -> a single 128-byte wavefront.
        // Warp w owns TMEM lanes 32*(w%4).. (hardware rule) and the column half (w/4); the two 32-column loads of
        // its half are issued back to back so the ~1k-cycle TMEM read latency is paid once per tile.
        const int quad = warp & 3, half = warp >> 2;
        const int f = quad * 32 + lane;
        const float bias_f = (kNeedBias && p.bias != nullptr) ? __ldg(p.bias + f) : 0.0f;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileM + half * 64);
        float v[2][32];
        if (!dbg_no_tld) {
          tmem_ld32_nowait(t_lane, v[0]);
          tmem_ld32_nowait(t_lane + 32, v[1]);
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int64_t r0 = tile * kTileM + half * 64 + c * 32;
          float* dst = p.D + r0 * p.ldd + f;
          const float* aux = kNeedAux ? p.aux + r0 * p.ld_aux + f : nullptr;
          if (r0 + 32 <= p.M) {                                   // full chunk: no per-row predicate
            if (!dbg_no_stg) {
              // operands of the epilogue (previous D for accumulate, activation output for act') are fetched as
              // 32 independent loads BEFORE the first store: one latency per chunk instead of one per row
              float t[32];
              if constexpr (kNeedAux || MODE == kModeAccumulate || MODE == kModeAccGather) {
                const float* src = kNeedAux ? aux : dst;
                const int64_t lds = kNeedAux ? p.ld_aux : p.ldd;
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = src[j * lds];
              }
              if constexpr (MODE == kModeAccGather) {
                // row metadata: lane j fetches row r0+j (one coalesced load each), broadcast by shuffle in the loop
                const int d_l = __ldg(p.g_dst + r0 + lane);
                const int rv_l = p.g_rev != nullptr ? (int)__ldg(p.g_rev + r0 + lane) : 0;
                const float w_l = p.g_norm != nullptr ? __ldg(p.g_norm + r0 + lane) : 1.0f;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[c][j] = __fadd_rn(t[j], v[c][j]);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int d = __shfl_sync(0xffffffffu, d_l, j);
                  const int rv = __shfl_sync(0xffffffffu, rv_l, j);
                  t[j] = __ldg((rv ? p.g_tab1 : p.g_tab0) + (int64_t)d * p.ld_tab + f);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int rv = __shfl_sync(0xffffffffu, rv_l, j);
                  float x = t[j];
                  if (p.g_norm != nullptr) x = __fmul_rn(x, __shfl_sync(0xffffffffu, w_l, j));
                  *dst = __fadd_rn(v[c][j], rv ? x : -x);
                  dst += p.ldd;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float y = kNeedAux ? t[j] : 0.0f;
                  const float old = (MODE == kModeAccumulate) ? t[j] : 0.0f;
                  *dst = epilogue_op<MODE>(v[c][j], bias_f, y, old, p.slope, act);
                  dst += p.ldd;
                }
              }
            }
          } else {
            const int nvalid = (int)(p.M - r0);                   // may be <= 0
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                const float y = kNeedAux ? aux[j * p.ld_aux] : 0.0f;
                const float old = (MODE == kModeAccumulate || MODE == kModeAccGather) ? dst[j * p.ldd] : 0.0f;
                float o = epilogue_op<MODE>(v[c][j], bias_f, y, old, p.slope, act);
                if constexpr (MODE == kModeAccGather) {
                  const int64_t r = r0 + j;
                  const int rv = p.g_rev != nullptr ? (int)__ldg(p.g_rev + r) : 0;
                  float x = __ldg((rv ? p.g_tab1 : p.g_tab0) + (int64_t)__ldg(p.g_dst + r) * p.ld_tab + f);
                  if (p.g_norm != nullptr) x = __fmul_rn(x, __ldg(p.g_norm + r));
                  o = __fadd_rn(o, rv ? x : -x);
                }
                dst[j * p.ldd] = o;
              }
            }
          }
        }
      } else {
        // TMEM lane == tile row (N = 64 features: the M = 64 transposed form would leave half the lanes idle)
        const int quad = warp & 3, half = warp >> 2;
        const int64_t r = tile * kTileM + quad * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll 1
        for (int c0 = half * (N / 2); c0 < (half + 1) * (N / 2); c0 += 32) {
          float v[32];
          tmem_ld32(t_row + c0, v);
          if (r < p.M) {
            float* drow = p.D + r * p.ldd + c0;
            const float* arow = kNeedAux ? p.aux + r * p.ld_aux + c0 : nullptr;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f), y = b, d = b, o;
              if (kNeedBias && p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * q));
              if (kNeedAux) y = *reinterpret_cast<const float4*>(arow + 4 * q);
              if (MODE == kModeAccumulate || MODE == kModeAccGather) d = *reinterpret_cast<const float4*>(drow + 4 * q);
              o.x = epilogue_op<MODE>(v[4 * q + 0], b.x, y.x, d.x, p.slope, act);
              o.y = epilogue_op<MODE>(v[4 * q + 1], b.y, y.y, d.y, p.slope, act);
              o.z = epilogue_op<MODE>(v[4 * q + 2], b.z, y.z, d.z, p.slope, act);
              o.w = epilogue_op<MODE>(v[4 * q + 3], b.w, y.w, d.w, p.slope, act);
              if constexpr (MODE == kModeAccGather) {
                const int rv = p.g_rev != nullptr ? (int)__ldg(p.g_rev + r) : 0;
                float4 x = __ldg(reinterpret_cast<const float4*>((rv ? p.g_tab1 : p.g_tab0) +
                                                                 (int64_t)__ldg(p.g_dst + r) * p.ld_tab + c0 + 4 * q));
                if (p.g_norm != nullptr) {
                  const float w = __ldg(p.g_norm + r);
                  x.x = __fmul_rn(x.x, w); x.y = __fmul_rn(x.y, w); x.z = __fmul_rn(x.z, w); x.w = __fmul_rn(x.w, w);
                }
                o.x = __fadd_rn(o.x, rv ? x.x : -x.x); o.y = __fadd_rn(o.y, rv ? x.y : -x.y);
                o.z = __fadd_rn(o.z, rv ? x.z : -x.z); o.w = __fadd_rn(o.w, rv ? x.w : -x.w);
              }
              *reinterpret_cast<float4*>(drow + 4 * q) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
      if (dbg_ts && threadIdx.x == 0 && dbg_tile < 256) dbg_ts[1024 + 2 * dbg_tile + 1] = clock64();
      ++dbg_tile;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

template <int N, int K, int MODE>
static int launch_gemm_mode(const GemmParams& p, cudaStream_t stream) {
  using L = Smem<N, K>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tf32x3_gemm_kernel<N, K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm_tf32x3: cannot reserve %d bytes of shared memory: %s", L::kTotal, cudaGetErrorString(e));
      return DMP_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t tiles = (p.M + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  GemmParams q = p;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  q.use_tma = (L::kTS && tma_enabled() && p.M >= kTileM && make_tmap_rows(&tmap, p.A, p.lda, p.M, K)) ? 1 : 0;
  tf32x3_gemm_kernel<N, K, MODE><<<grid, kThreadsGemm, L::kTotal, stream>>>(q, tmap);
  return launch_status("tf32x3_gemm_kernel");
}

template <int N, int K>
static int launch_gemm(const GemmParams& p, int mode, cudaStream_t stream) {
  switch (mode) {
    case kModeStore: return launch_gemm_mode<N, K, kModeStore>(p, stream);
    case kModeAccumulate: return launch_gemm_mode<N, K, kModeAccumulate>(p, stream);
    case kModeBiasPwl: return launch_gemm_mode<N, K, kModeBiasPwl>(p, stream);
    case kModeGradPwl: return launch_gemm_mode<N, K, kModeGradPwl>(p, stream);
    case kModeBiasSmooth: return launch_gemm_mode<N, K, kModeBiasSmooth>(p, stream);
    case kModeAccGather: return launch_gemm_mode<N, K, kModeAccGather>(p, stream);
    default: return launch_gemm_mode<N, K, kModeGradSmooth>(p, stream);
  }
}

}  // namespace gemm
}  // namespace dmp

extern "C" int dmp_gemm_tf32x3(const float* A, int64_t lda, const float* row_scale, const float* Bt, int64_t ldb,
                               const float* bias, const float* aux, int64_t ld_aux, float* D, int64_t ldd,
                               int64_t M, int64_t N, int64_t K, int epilogue, float slope, void* stream) {
  using namespace dmp;
  using namespace dmp::gemm;
  DMP_CHECK_ARG(M >= 0, "gemm_tf32x3: negative M");
  if (M == 0) return DMP_OK;
  DMP_CHECK_ARG(A && Bt && D, "gemm_tf32x3: null pointer");
  DMP_CHECK_ARG((N == 64 || N == 128) && (K == 64 || K == 128), "gemm_tf32x3: N and K must be 64 or 128 (got %lld, %lld)",
                (long long)N, (long long)K);
  DMP_CHECK_ARG(lda >= K && ldb >= K && ldd >= N && lda % 4 == 0 && ldb % 4 == 0 && ldd % 4 == 0,
                "gemm_tf32x3: leading dimensions must be >= the row length and multiples of 4");
  DMP_CHECK_ARG(aligned_to(A, 16) && aligned_to(Bt, 16) && aligned_to(D, 16) && aligned_to(bias, 16) && aligned_to(aux, 16),
                "gemm_tf32x3: operands must be 16-byte aligned");
  const int act = epilogue & 15;
  DMP_CHECK_ARG(act >= DMP_ACT_NONE && act <= DMP_ACT_SIGMOID, "gemm_tf32x3: bad activation");
#ifndef DMP_DEBUG
  DMP_CHECK_ARG((epilogue & ~(15 | kEpiMulActGradFromOutput | kEpiAccumulate)) == 0, "gemm_tf32x3: unknown epilogue bits 0x%x",
                epilogue);
#endif
  DMP_CHECK_ARG(!(epilogue & kEpiMulActGradFromOutput) || (aux != nullptr && ld_aux >= N && ld_aux % 4 == 0),
                "gemm_tf32x3: act' epilogue needs aux");
  DMP_CHECK_ARG(A != D, "gemm_tf32x3: D must not alias A");
  const bool mul_grad = (epilogue & kEpiMulActGradFromOutput) != 0;
  const bool accumulate = (epilogue & kEpiAccumulate) != 0;
  DMP_CHECK_ARG(!(accumulate && (mul_grad || bias != nullptr || act != DMP_ACT_NONE)),
                "gemm_tf32x3: accumulate cannot be combined with bias / activation epilogues");
  DMP_CHECK_ARG(!(mul_grad && bias != nullptr), "gemm_tf32x3: act' epilogue takes no bias");
  GemmParams p;
  p.A = A; p.lda = lda; p.row_scale = row_scale; p.Bt = Bt; p.ldb = ldb; p.bias = bias;
  p.aux = aux; p.ld_aux = ld_aux; p.D = D; p.ldd = ldd; p.M = M; p.epilogue = epilogue; p.slope = slope;
  p.g_dst = nullptr; p.g_rev = nullptr; p.g_norm = nullptr; p.g_tab0 = p.g_tab1 = nullptr; p.ld_tab = 0;
  const bool smooth = (act == DMP_ACT_TANH || act == DMP_ACT_SIGMOID);
  if (act == DMP_ACT_NONE) p.slope = 1.0f;   // piecewise-linear family: none = slope 1, relu = slope 0
  if (act == DMP_ACT_RELU) p.slope = 0.0f;
  int mode;
  if (accumulate) mode = kModeAccumulate;
  else if (mul_grad) mode = smooth ? kModeGradSmooth : (act == DMP_ACT_NONE ? kModeStore : kModeGradPwl);
  else if (smooth) mode = kModeBiasSmooth;
  else if (bias != nullptr || act != DMP_ACT_NONE) mode = kModeBiasPwl;
  else mode = kModeStore;
  p.epi_scale = nullptr;
  if (mode == kModeAccumulate && row_scale != nullptr) { p.epi_scale = row_scale; p.row_scale = nullptr; }
  cudaStream_t s = (cudaStream_t)stream;
  static const bool use_v2 = [] { const char* e = getenv("DMP_GEMM_V2"); return e ? atoi(e) != 0 : true; }();
  if (use_v2 && p.row_scale == nullptr)   // epilogue-mode ids coincide (kModeStore..kModeGradSmooth = 0..5)
    return launch_gemm_v2(A, lda, Bt, ldb, p.epi_scale, bias, aux, ld_aux, D, ldd, M, N, K, mode, act, p.slope, s);
  if (p.epi_scale != nullptr && N == 64) { p.row_scale = p.epi_scale; p.epi_scale = nullptr; }   // legacy N = 64: scale on A
  if (N == 128 && K == 128) return launch_gemm<128, 128>(p, mode, s);
  if (N == 128 && K == 64) return launch_gemm<128, 64>(p, mode, s);
  if (N == 64 && K == 128) return launch_gemm<64, 128>(p, mode, s);
  return launch_gemm<64, 64>(p, mode, s);
}

extern "C" int dmp_gemm_tf32x3_acc_gather(const float* A, int64_t lda, const float* row_scale, const float* Bt,
                                          int64_t ldb, float* D, int64_t ldd, int64_t M, int64_t N, int64_t K,
                                          const int32_t* dst32, const uint8_t* rev, const float* norm,
                                          const float* tab_fwd, const float* tab_rev, int64_t ld_tab, void* stream) {
  using namespace dmp;
  using namespace dmp::gemm;
  DMP_CHECK_ARG(M >= 0, "gemm_acc_gather: negative M");
  if (M == 0) return DMP_OK;
  DMP_CHECK_ARG(A && Bt && D && dst32 && tab_fwd, "gemm_acc_gather: null pointer");
  DMP_CHECK_ARG(rev == nullptr || tab_rev != nullptr, "gemm_acc_gather: reversed edges need tab_rev");
  DMP_CHECK_ARG((N == 64 || N == 128) && (K == 64 || K == 128), "gemm_acc_gather: N and K must be 64 or 128");
  DMP_CHECK_ARG(lda >= K && ldb >= K && ldd >= N && ld_tab >= N && lda % 4 == 0 && ldb % 4 == 0 && ldd % 4 == 0 &&
                    ld_tab % 4 == 0,
                "gemm_acc_gather: leading dimensions must be >= the row length and multiples of 4");
  DMP_CHECK_ARG(aligned_to(A, 16) && aligned_to(Bt, 16) && aligned_to(D, 16) && aligned_to(tab_fwd, 16) &&
                    aligned_to(tab_rev, 16),
                "gemm_acc_gather: operands must be 16-byte aligned");
  DMP_CHECK_ARG(A != D, "gemm_acc_gather: D must not alias A");
  GemmParams p;
  p.A = A; p.lda = lda; p.row_scale = row_scale; p.epi_scale = nullptr; p.Bt = Bt; p.ldb = ldb; p.bias = nullptr;
  p.aux = nullptr; p.ld_aux = 0; p.D = D; p.ldd = ldd; p.M = M; p.epilogue = 0; p.slope = 1.0f;
  p.g_dst = dst32; p.g_rev = rev; p.g_norm = norm; p.g_tab0 = tab_fwd; p.g_tab1 = tab_rev ? tab_rev : tab_fwd;
  p.ld_tab = ld_tab;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 128 && K == 128) return launch_gemm<128, 128>(p, kModeAccGather, s);
  if (N == 128 && K == 64) return launch_gemm<128, 64>(p, kModeAccGather, s);
  if (N == 64 && K == 128) return launch_gemm<64, 128>(p, kModeAccGather, s);
  return launch_gemm<64, 64>(p, kModeAccGather, s);
}
