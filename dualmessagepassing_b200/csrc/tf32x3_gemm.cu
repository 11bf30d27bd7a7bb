// tf32x3_gemm.cu -- row-streaming projections on the tcgen05 tensor cores, 3xTF32 split, CROSS-TERMS-FIRST.
//
//   single   D[M,128]  = epilogue( (A[M,K]) · Wt[128,K]^T )                         (dmp_gemm_tf32x3, N = 128)
//   dual     D[M,N] (+)= A·W1t^T + c ⊙ (A·W2t^T)   |   D1 = A·W1t^T, D2 = A·W2t^T    (dmp_gemm_tf32x3_dual, N in {64,128})
//
// Accuracy.  The tensor core accumulates in fp32 with TRUNCATION (round toward zero), not round-to-nearest.  Round 1
// issued the three products of the split (lo·hi, hi·lo, hi·hi) interleaved per k-step, so every one of the 48 MMAs of
// a K = 128 row tile truncated an accumulator that already held the large hi·hi partial sums: a systematic 1.2e-6
// (max-norm, vs fp64) against 4.6e-7 for an fp32 FMA GEMM.  Here a tile is 64 rows with ALL of K resident in shared
// memory, and the MMAs are issued cross terms first (32 MMAs into a still-small accumulator: their truncation is 2^-11
// smaller), the 16 hi·hi MMAs last: a bit-level model of the accumulator predicts 6.6e-7 (tests/test_gpu_gemm.py
// asserts <= 1.0e-6; measured values in DESIGN.md).  Two kernels share this file: the 128-row kernel every production
// call takes ("v3", second half of the file) and the 64-row kernel below it falls back to for M < 128.
//
// Structure (one persistent CTA per SM, 17 warps):
//   warps 9..16  PRODUCERS  one elected thread issues K/32 TMA boxes (64 rows x 32 floats, SWIZZLE_128B) per tile onto
//                           the stage's mbarrier (expect_tx); the raw tile IS the hi operand (kind::tf32 ignores the low
//                           13 mantissa bits); the warps then write lo = x - trunc_tf32(x).  Without TMA (M < 64 or
//                           encoder unavailable): per-thread cp.async.
//   warp  8      MMA        weights live in TENSOR MEMORY as the A operand ("TS" form, transposed product
//                           D^T[feature, row] = W[feature, k] · X[row, k]^T: a TMEM lane is an output feature, the 32
//                           lanes of an epilogue warp store 32 consecutive floats of one output row)
//   warps 0..7   EPILOGUE   4 accumulator buffers of 64 columns in TMEM; warp (quadrant q, half h) owns features
//                           32q.. and rows 32h.. of the tile.
// Dual form: lanes 0..63 = rows f0.. of W1, lanes 64..127 = the same rows of W2; acc1[f, r] sits in quadrant q, acc2[f, r]
// in quadrant q+2 -- different warps by the hardware's lane-quadrant rule -- so the two warps swap half of their rows
// through shared memory (named barrier per pair) and each finalises 16 rows.  For N = 128 two CTAs (blockIdx parity =
// feature half) walk the same tiles in the same order: the second read of a tile hits L2.
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace dmp {
namespace gemm {

constexpr int kV2Rows = 64;                         // rows per tile = MMA N
constexpr int kV2ProducerWarps = 8;
constexpr int kV2EpilogueWarps = 8;
constexpr int kV2MmaWarp = 8;
constexpr int kV2Threads = (kV2EpilogueWarps + 1 + kV2ProducerWarps) * 32;   // 544
constexpr int kV2AccBufs = 4;                       // TMEM accumulator ring: 4 x 64 columns
constexpr int kV2XchgBytes = 8 * 2048;              // dual: per epilogue warp 16 rows x 32 features

enum : int {
  kV2Store = 0, kV2Accumulate = 1, kV2BiasPwl = 2, kV2GradPwl = 3, kV2BiasSmooth = 4, kV2GradSmooth = 5,   // single
  kV2AccumulateScaled = 6,                         // D += s_r * acc_r (row scale applied to the accumulator)
  kV2DualStore = 8, kV2DualAccumulate = 9, kV2DualSeparate = 10,                                            // dual
};

template <int K>
struct V2Smem {
  static constexpr int kKBlocks = K / kKB;
  static constexpr int kBlockBytes = kV2Rows * 128;             // one k-block (64 rows x 32 floats): 8 KB
  static constexpr int kHalfBytes = kKBlocks * kBlockBytes;     // hi or lo of one tile
  static constexpr int kStageBytes = 2 * kHalfBytes;            // 64 KB (K = 128) / 32 KB (K = 64)
  static constexpr int kStages = (K == 128) ? 3 : 6;
  static constexpr int kTotal = kStages * kStageBytes + kV2XchgBytes + 256 + 1024;
};

struct V2Params {
  const float* A; int64_t lda;
  const float* W1; const float* W2; int64_t ldw;   // [N, K] (nn.Linear layout); W2 only in the dual forms
  const float* scale;                              // single/accumulate: D += s_r * acc_r; dual: c_r
  const float* pre_scale;                          // single, non-accumulate: rows of A are scaled first (rounded product)
  const float* bias;
  const float* aux; int64_t ld_aux;
  float* D; int64_t ldd;
  float* D2; int64_t ldd2;
  int64_t M;
  int act; float slope;
  int use_tma;
  unsigned int* tile_ctr;                          // v3: dynamic tile scheduler ([half 0, half 1, CTAs done]) or NULL = static
};

template <int MODE>
__device__ __forceinline__ float v2_epilogue_op(float acc, float bias, float aux, float old, float slope, int act) {
  if constexpr (MODE == kV2Store) return acc;
  if constexpr (MODE == kV2Accumulate || MODE == kV2AccumulateScaled) return __fadd_rn(old, acc);
  if constexpr (MODE == kV2BiasPwl) {
    const float x = __fadd_rn(acc, bias);
    return x > 0.0f ? x : __fmul_rn(x, slope);
  }
  if constexpr (MODE == kV2GradPwl) return __fmul_rn(acc, aux > 0.0f ? 1.0f : slope);
  if constexpr (MODE == kV2BiasSmooth) return apply_act(__fadd_rn(acc, bias), act, slope);
  return __fmul_rn(acc, act_grad_from_output(aux, act, slope));
}

// NOUT: output features (dual: 64 or 128; single: 128).  DUAL: two weights stacked on the TMEM lanes.
template <int NOUT, int K, int MODE>
__global__ void __launch_bounds__(kV2Threads, 1) tf32x3_gemm_v2_kernel(const V2Params p,
                                                                       const __grid_constant__ CUtensorMap tmap) {
  using L = V2Smem<K>;
  constexpr bool kDual = MODE >= kV2DualStore;
  constexpr int kKBlocks = L::kKBlocks;
  constexpr int kStages = L::kStages;
  constexpr int kHalves = kDual ? NOUT / 64 : 1;              // CTAs per row tile (dual, N = 128: feature halves)
  // single-weight form with NOUT = 64: lanes 64..127 carry zero weights (the MMA is still M = 128; at 64 features the
  // doubled tensor work stays hidden behind the HBM stream) and their epilogue warps only release the accumulator
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                                   // stage s: [hi: kKBlocks x 8 KB][lo: kKBlocks x 8 KB]
  const uint32_t sX = sA + kStages * L::kStageBytes;          // dual exchange buffer: warp w writes [w*2048, +2048)
  const uint32_t sBar = sX + kV2XchgBytes;
  const uint32_t bar_full = sBar;                             // kStages x 8 B   (lo written, hi landed)
  const uint32_t bar_empty = sBar + 8 * kStages;              // MMAs of the stage retired
  const uint32_t bar_raw = sBar + 16 * kStages;               // TMA completion of the raw (= hi) tile
  const uint32_t bar_acc_full = sBar + 24 * kStages;          // kV2AccBufs x 8 B
  const uint32_t bar_acc_empty = bar_acc_full + 8 * kV2AccBufs;
  const uint32_t tmem_slot = bar_acc_empty + 8 * kV2AccBufs;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int fh = (kHalves == 2) ? (int)(blockIdx.x & 1) : 0;
  const int64_t tile0 = (kHalves == 2) ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t tstep = (kHalves == 2) ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const int64_t num_tiles = (p.M + kV2Rows - 1) / kV2Rows;
  const int64_t my_tiles = tile0 < num_tiles ? (num_tiles - tile0 + tstep - 1) / tstep : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, kV2ProducerWarps);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_raw + 8 * s, 1);
    }
    for (int a = 0; a < kV2AccBufs; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kV2EpilogueWarps);
    }
    fence_barrier_init();
  }
  // TMEM map: [0,256) 4 accumulators of 64 columns | [256,256+K) W hi | [256+K,256+2K) W lo
  constexpr int kTmemCols = 512;
  constexpr uint32_t kWhiCol = 256, kWloCol = 256 + K;
  if (warp == kV2MmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp < 4) {   // thread = TMEM lane = weight row (single) | (weight, row) (dual); hi = tf32(w), lo = w - hi
    const int l = warp * 32 + lane;
    const float* wrow;
    if constexpr (kDual) wrow = ((l < 64) ? p.W1 : p.W2) + (int64_t)(fh * 64 + (l & 63)) * p.ldw;
    else wrow = p.W1 + (int64_t)(l < NOUT ? l : 0) * p.ldw;
    const bool zero_row = !kDual && l >= NOUT;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < K; c0 += 32) {
      float hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = __ldg(reinterpret_cast<const float4*>(wrow + c0 + 4 * q));
        if (zero_row) v = make_float4(0.f, 0.f, 0.f, 0.f);
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

  if (warp > kV2MmaWarp) {
    // =========================== PRODUCERS ===========================
    const int pt = threadIdx.x - (kV2MmaWarp + 1) * 32;        // 0..255
    // thread handles 16-byte chunks c = pt + 256 i of a k-block: row = c / 8 (0..63 over i < 2), chunk = c % 8
    const int c16 = pt & 7;
    const int row0 = pt >> 3;                                   // rows row0, row0 + 32
    uint32_t offs[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) offs[i] = swz(row0 + 32 * i, c16);
    constexpr int kAhead = kStages - 1;                         // tiles of copies in flight
    int istage = 0;
    uint32_t iphase = 0;
    int64_t itile = 0;
    auto issue = [&]() {                                        // start the copies of local tile `itile`
      mbar_wait(bar_empty + 8 * istage, iphase ^ 1);
      const uint32_t hi = sA + istage * L::kStageBytes;
      const int64_t r0 = (tile0 + itile * tstep) * kV2Rows;
      if (p.use_tma) {
        if (pt == 0) {
          mbar_expect_tx(bar_raw + 8 * istage, L::kHalfBytes);
#pragma unroll
          for (int kb = 0; kb < kKBlocks; ++kb)
            tma_load_2d(hi + kb * L::kBlockBytes, &tmap, kb * kKB, (int)r0, bar_raw + 8 * istage);
        }
      } else {
        const float* src = p.A + (r0 + row0) * p.lda + c16 * 4;
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const bool ok = r0 + row0 + 32 * i < p.M;
            cp_async16(hi + kb * L::kBlockBytes + offs[i],
                       ok ? (const void*)(src + (int64_t)(32 * i) * p.lda + kb * kKB) : (const void*)p.A, ok ? 16u : 0u);
          }
        }
      }
      ++itile;
      if (++istage == kStages) { istage = 0; iphase ^= 1; }
    };
#pragma unroll 1
    for (int d = 0; d < kAhead; ++d) {
      if (d < my_tiles) issue();
      cp_async_commit();
    }
    int stage = 0;
    uint32_t rphase = 0;
#pragma unroll 1
    for (int64_t t = 0; t < my_tiles; ++t) {
      if (p.use_tma) mbar_wait(bar_raw + 8 * stage, rphase);
      else cp_async_wait<kAhead - 1>();
      const uint32_t hi = sA + stage * L::kStageBytes;
      const uint32_t lo = hi + L::kHalfBytes;
      float sc[2] = {1.0f, 1.0f};
      if (p.pre_scale != nullptr) {
        const int64_t r0 = (tile0 + t * tstep) * kV2Rows + row0;
#pragma unroll
        for (int i = 0; i < 2; ++i) sc[i] = (r0 + 32 * i < p.M) ? __ldg(p.pre_scale + r0 + 32 * i) : 1.0f;
      }
#pragma unroll
      for (int kb = 0; kb < kKBlocks; ++kb) {
        float4 v[2];                         // loads before stores: the volatile asm keeps program order
#pragma unroll
        for (int i = 0; i < 2; ++i) v[i] = lds128(hi + kb * L::kBlockBytes + offs[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (p.pre_scale != nullptr) {      // (row_scale ⊙ A) as an individually rounded fp32 product, then the split
            v[i].x = __fmul_rn(sc[i], v[i].x); v[i].y = __fmul_rn(sc[i], v[i].y);
            v[i].z = __fmul_rn(sc[i], v[i].z); v[i].w = __fmul_rn(sc[i], v[i].w);
            sts128(hi + kb * L::kBlockBytes + offs[i], v[i]);
          }
          sts128(lo + kb * L::kBlockBytes + offs[i],
                 make_float4(tf32_trunc_residual(v[i].x), tf32_trunc_residual(v[i].y), tf32_trunc_residual(v[i].z),
                             tf32_trunc_residual(v[i].w)));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      if (++stage == kStages) { stage = 0; rphase ^= 1; }
      if (t + kAhead < my_tiles) issue();
      cp_async_commit();
    }
  } else if (warp == kV2MmaWarp) {
    // =========================== MMA ISSUER ===========================
    constexpr uint32_t idesc = make_idesc(128, kV2Rows);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
#pragma unroll 1
    for (int64_t t = 0; t < my_tiles; ++t) {
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kV2Rows);
        const uint32_t a_hi = sA + stage * L::kStageBytes;
        const uint32_t a_lo = a_hi + L::kHalfBytes;
        // cross terms first (the accumulator is still small: their truncation is harmless), dominant hi*hi last
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
#pragma unroll
          for (int j = 0; j < kKB / 8; ++j) {
            const uint64_t dah = make_smem_desc(a_hi + kb * L::kBlockBytes + j * 32);
            const uint64_t dal = make_smem_desc(a_lo + kb * L::kBlockBytes + j * 32);
            const uint32_t w_hi = tmem_base + kWhiCol + (uint32_t)(kb * kKB + j * 8);
            const uint32_t w_lo = tmem_base + kWloCol + (uint32_t)(kb * kKB + j * 8);
            umma_tf32_ts(d_tmem, w_hi, dal, idesc, (kb | j) != 0 ? 1u : 0u);
            umma_tf32_ts(d_tmem, w_lo, dah, idesc, 1u);
          }
        }
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
#pragma unroll
          for (int j = 0; j < kKB / 8; ++j) {
            const uint64_t dah = make_smem_desc(a_hi + kb * L::kBlockBytes + j * 32);
            const uint32_t w_hi = tmem_base + kWhiCol + (uint32_t)(kb * kKB + j * 8);
            umma_tf32_ts(d_tmem, w_hi, dah, idesc, 1u);
          }
        }
        umma_commit(bar_empty + 8 * stage);                   // smem stage free once these MMAs retire
        umma_commit(bar_acc_full + 8 * acc);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
      if (++acc == kV2AccBufs) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== EPILOGUE ===========================
    const int quad = warp & 3, half = warp >> 2;     // TMEM lane quadrant (hardware rule: warp % 4), row half (32 rows)
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (!kDual) {
      constexpr bool kNeedAux = (MODE == kV2GradPwl || MODE == kV2GradSmooth);
      constexpr bool kNeedBias = (MODE == kV2BiasPwl || MODE == kV2BiasSmooth);
      constexpr bool kNeedOld = (MODE == kV2Accumulate || MODE == kV2AccumulateScaled);
      constexpr bool kScaled = (MODE == kV2AccumulateScaled);
      const int f = quad * 32 + lane;
      const bool live = quad * 32 < NOUT;               // NOUT = 64: quadrants 2, 3 hold the zero rows
      const float bias_f = (kNeedBias && live && p.bias != nullptr) ? __ldg(p.bias + f) : 0.0f;
#pragma unroll 1
      for (int64_t t = 0; t < my_tiles; ++t) {
        if (!live) {                                    // warp-uniform
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
          if (++acc == kV2AccBufs) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        const int64_t r0 = (tile0 + t * tstep) * kV2Rows + half * 32;
        const int nvalid = (int)((p.M - r0) < 32 ? (p.M - r0) : 32);     // warp-uniform, may be <= 0
        float* dst = p.D + r0 * p.ldd + f;
        float tt[32];
        if constexpr (kNeedAux || kNeedOld) {
          // streamed epilogue operand requested BEFORE waiting for this tile's MMAs: its DRAM latency hides behind them
          const float* src = kNeedAux ? p.aux + r0 * p.ld_aux + f : dst;
          const int64_t lds = kNeedAux ? p.ld_aux : p.ldd;
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tt[j] = src[j * lds];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) tt[j] = j < nvalid ? src[j * lds] : 0.0f;
          }
        }
        float sc_l = 1.0f;
        if (kScaled && lane < nvalid) sc_l = __ldg(p.scale + r0 + lane);
        mbar_wait(bar_acc_full + 8 * acc, acc_phase);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kV2Rows + half * 32), v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        // accumulate reads and writes the same addresses: hide that from the compiler, or it keeps all 32 load addresses
        // (64 registers) alive across the accumulator wait for re-use by the stores and spills
        asm volatile("" : "+l"(dst));
        if constexpr (kScaled) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(__shfl_sync(0xffffffffu, sc_l, j), v[j]);
        }
        if (nvalid == 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            dst[j * p.ldd] = v2_epilogue_op<MODE>(v[j], bias_f, kNeedAux ? tt[j] : 0.0f, kNeedOld ? tt[j] : 0.0f, p.slope, p.act);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid)
              dst[j * p.ldd] = v2_epilogue_op<MODE>(v[j], bias_f, kNeedAux ? tt[j] : 0.0f, kNeedOld ? tt[j] : 0.0f, p.slope, p.act);
        }
        if (++acc == kV2AccBufs) { acc = 0; acc_phase ^= 1; }
      }
    } else {
      const int which = quad >> 1;                     // 0: this warp holds acc1 (W1), 1: acc2 (W2)
      const int f = fh * 64 + (quad & 1) * 32 + lane;  // output feature of this thread
      const uint32_t my_x = sX + (uint32_t)warp * 2048u + (uint32_t)lane * 4u;           // [row j][lane] floats
      const uint32_t peer_x = sX + (uint32_t)(warp ^ 2) * 2048u + (uint32_t)lane * 4u;   // written by quadrant q^2, same half
      const int bar_id = 1 + (quad & 1) * 2 + half;    // named barrier of the pair (ids 1..4), 64 threads
#pragma unroll 1
      for (int64_t t = 0; t < my_tiles; ++t) {
        const int64_t rt = (tile0 + t * tstep) * kV2Rows + half * 32;     // first of this warp's 32 rows
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kV2Rows + half * 32);
        if constexpr (MODE == kV2DualSeparate) {
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          tc_fence_after();
          float v[32];
          tmem_ld32(t_lane, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
          float* out = which ? p.D2 : p.D;
          const int64_t ldo = which ? p.ldd2 : p.ldd;
          float* dst = out + rt * ldo + f;
          const int nvalid = (int)((p.M - rt) < 32 ? (p.M - rt) : 32);
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * ldo] = v[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) dst[j * ldo] = v[j];
          }
        } else {
          // rows this warp finalises: 16 rows starting at r0 (acc1 warp: the first 16 of the pair's 32, acc2 warp: the rest)
          const int64_t r0 = rt + which * 16;
          const int nvalid = (int)((p.M - r0) < 16 ? (p.M - r0) : 16);     // warp-uniform, may be <= 0
          float* dst = p.D + r0 * p.ldd + f;
          float old[16];
          if constexpr (MODE == kV2DualAccumulate) {
            // previous D requested BEFORE waiting for this tile's MMAs
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = j < nvalid ? dst[j * p.ldd] : 0.0f;
          }
          float sc_l = 1.0f;       // lane j (< 16) holds the scale of row r0 + j; broadcast by shuffle below
          if (p.scale != nullptr && lane < nvalid) sc_l = __ldg(p.scale + r0 + lane);
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          tc_fence_after();
          float keep[16], send[16];
          tmem_ld16(t_lane + (which ? 16 : 0), keep);           // the rows this warp finalises
          tmem_ld16(t_lane + (which ? 0 : 16), send);           // the rows the partner finalises
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);  // everything needed is out of TMEM
          asm volatile("" : "+l"(dst));                         // see the single-weight accumulate epilogue
          // the partner has finished reading what I wrote for the previous tile
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_x + j * 128), "f"(send[j]) : "memory");
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float got = lds32(peer_x + j * 128);
            const float a1 = which ? got : keep[j];
            const float a2 = which ? keep[j] : got;
            const float c = __shfl_sync(0xffffffffu, sc_l, j);
            float r = (MODE == kV2DualAccumulate) ? __fadd_rn(old[j], a1) : a1;
            r = __fadd_rn(r, __fmul_rn(c, a2));
            if (j < nvalid) dst[j * p.ldd] = r;
          }
        }
        if (++acc == kV2AccBufs) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kV2MmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}


// =====================================================================================================================
// v3: the same products on 128-ROW tiles.  Measured on this B200 (scripts/micro/mma_rate.cu): a kind::tf32 MMA with
// M = 128 costs 64 cycles at N = 128 (2048 MAC/clk/SM, the pipe's full rate) but 45 cycles at N = 64 (an issue floor),
// so the 64-row tiles above pay 1.4x the tensor time per row.  Here N = 128, and cross-terms-first needs the whole K of
// a 128-row tile resident: 64 KB of raw (= hi) data per tile, so hi and lo live in SEPARATE rings of k-block slots:
//   hi ring  12 slots x 16 KB (three tiles at K = 128): TMA lands here; a tile's slots are held until its hi*hi pass retires
//   lo ring   2 slots x 16 KB: written by the producer warps, released k-block by k-block as the cross-term pass retires
// Warps: 0-7 epilogue, 8 MMA issuer, 9-16 lo producers, 17 TMA issuer (decoupled from the producers so that the loads of
// tile t+2 start the moment tile t retires).  MMA order per tile: for every k-block [lo*hi, hi*lo] x 4 k-steps (commit
// frees the lo slot), then for every k-block hi*hi x 4 (one commit frees the tile's hi slots and publishes the
// accumulator).  Needs TMA (M >= 128); anything else runs the 64-row kernel above.
constexpr int kV3Rows = 128;
constexpr int kV3Threads = 18 * 32;
constexpr int kV3TmaWarp = 17;
constexpr int kV3SlotBytes = kV3Rows * 128;         // 16 KB: 128 rows x 32 floats
// 224 KB of k-block slots (14) for every form.  The dual forms interleave the two weight matrices lane by lane
// (TMEM lane 2i = row f0+i of W1, lane 2i+1 = the same row of W2), so the two accumulators of a feature sit in
// NEIGHBOURING LANES OF ONE WARP and the epilogue combines them with shuffles -- the earlier layout (lanes 0..63 / 64..127)
// put them in different warps and cost a 32 KB exchange buffer, two named barriers and 64 KB of shared-memory traffic per tile.
constexpr int kV3Slots = 14;
constexpr int kV3Smem = kV3Slots * kV3SlotBytes + 512 + 1024;

template <int NOUT, int K, int MODE>
__global__ void __launch_bounds__(kV3Threads, 1) tf32x3_gemm_v3_kernel(const V2Params p,
                                                                       const __grid_constant__ CUtensorMap tmap) {
  constexpr bool kDual = MODE >= kV2DualStore;
  constexpr int kKBlocks = K / kKB;                           // k-blocks per tile: 4 or 2
  // ring split: 2 lo slots (they only bridge split -> cross-term MMA retirement), everything else is TMA prefetch depth.
  // COMPILE-TIME on purpose: as launch parameters (swept in profiles/r2_gemm_sweep.txt: 1 lo slot -12 %, 3-4 no gain) the
  // run-time modulo in the single MMA-issuing thread cost 15-25 % of every projection
  constexpr int kV3LoSlots = 2;
  constexpr int kV3HiSlots = kV3Slots - kV3LoSlots;          // 12: the barrier block below holds exactly that many
  constexpr int kHalves = kDual ? NOUT / 64 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sHi = base;
  const uint32_t sLo = sHi + kV3HiSlots * kV3SlotBytes;
  const uint32_t sBar = sLo + kV3LoSlots * kV3SlotBytes;      // 224 KB of tiles below the barriers
  const uint32_t bar_raw = sBar;                              // 12: TMA completion of a hi slot
  const uint32_t bar_empty_hi = sBar + 96;                    // 12: the tile that used the hi slot has retired
  const uint32_t bar_full_lo = sBar + 192;                    // <= 4: lo slot written
  const uint32_t bar_empty_lo = sBar + 224;                   // <= 4: cross-term MMAs of the k-block retired
  const uint32_t bar_acc_full = sBar + 256;                   // 2
  const uint32_t bar_acc_empty = sBar + 272;                  // 2
  const uint32_t tmem_slot = sBar + 288;
  // Tile queue: the TMA thread decides which tile comes next (statically strided, or from a global counter so that CTAs
  // which start late -- an NCCL collective was holding their SM -- take only what is left) and publishes it here;
  // every other role reads its tile ids from the queue.  Entry n is overwritten by entry n + kTileQ: by then the hi
  // ring (<= 6 tiles) and the two accumulators guarantee that every role is past entry n + kTileQ - 8.
  constexpr int kTileQ = 16;
  const uint32_t bar_tile = sBar + 304;                       // 16: queue entry written
  const uint32_t s_tile = sBar + 432;                         // 16 x int32: tile id, -1 = no more work
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int fh = (kHalves == 2) ? (int)(blockIdx.x & 1) : 0;
  const int64_t tile0 = (kHalves == 2) ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t tstep = (kHalves == 2) ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const int64_t num_tiles = (p.M + kV3Rows - 1) / kV3Rows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kV3HiSlots; ++s) {
      mbar_init(bar_raw + 8 * s, 1);
      mbar_init(bar_empty_hi + 8 * s, 1);
    }
    for (int s = 0; s < kV3LoSlots; ++s) {
      mbar_init(bar_full_lo + 8 * s, kV2ProducerWarps);
      mbar_init(bar_empty_lo + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kV2EpilogueWarps);
    }
    for (int q = 0; q < kTileQ; ++q) mbar_init(bar_tile + 8 * q, 1);
    fence_barrier_init();
  }
  // TMEM map: [0,128) acc 0 | [128,256) acc 1 | [256,256+K) W hi | [256+K,256+2K) W lo
  constexpr int kTmemCols = 512;
  constexpr uint32_t kWhiCol = 256, kWloCol = 256 + K;
  if (warp == kV2MmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp < 4) {
    const int l = warp * 32 + lane;
    const float* wrow;
    if constexpr (kDual) wrow = ((l & 1) ? p.W2 : p.W1) + (int64_t)(fh * 64 + (l >> 1)) * p.ldw;   // interleaved lanes
    else wrow = p.W1 + (int64_t)(l < NOUT ? l : 0) * p.ldw;
    const bool zero_row = !kDual && l >= NOUT;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < K; c0 += 32) {
      float hi[32], lo[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = __ldg(reinterpret_cast<const float4*>(wrow + c0 + 4 * q));
        if (zero_row) v = make_float4(0.f, 0.f, 0.f, 0.f);
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

  // consumer side of the tile queue: n-th tile of this CTA, or -1
  int tq = 0;
  uint32_t tq_phase = 0;
  auto next_tile = [&]() -> int64_t {
    mbar_wait(bar_tile + 8 * tq, tq_phase);
    const int32_t t = (int32_t)lds32u(s_tile + 4 * tq);
    if (++tq == kTileQ) { tq = 0; tq_phase ^= 1; }
    return (int64_t)t;
  };

  if (warp == kV3TmaWarp) {
    // =========================== TMA ISSUER ===========================
    if (elect_one()) {   // (the warp arrives here converged)
      int slot = 0, wq = 0;
      uint32_t sphase = 0;
      int64_t tile = tile0;
#pragma unroll 1
      for (;;) {
        if (p.tile_ctr != nullptr) tile = (int64_t)atomicAdd(p.tile_ctr + fh, 1u);
        const bool more = tile < num_tiles;
        sts32u(s_tile + 4 * wq, more ? (uint32_t)tile : 0xffffffffu);
        mbar_arrive(bar_tile + 8 * wq);                         // release: the entry is visible to whoever sees the phase
        if (++wq == kTileQ) wq = 0;
        if (!more) break;
        const int r0 = (int)(tile * kV3Rows);
#pragma unroll 1
        for (int kb = 0; kb < kKBlocks; ++kb) {
          mbar_wait(bar_empty_hi + 8 * slot, sphase ^ 1);       // the slot's previous tile has retired
          mbar_expect_tx(bar_raw + 8 * slot, kV3SlotBytes);
          tma_load_2d(sHi + slot * kV3SlotBytes, &tmap, kb * kKB, r0, bar_raw + 8 * slot);
          if (++slot == kV3HiSlots) { slot = 0; sphase ^= 1; }
        }
        tile += tstep;
      }
    }
  } else if (warp > kV2MmaWarp) {
    // =========================== LO PRODUCERS ===========================
    const int pt = threadIdx.x - (kV2MmaWarp + 1) * 32;        // 0..255
    const int c16 = pt & 7;
    const int row0 = pt >> 3;                                   // rows row0 + 32 i
    uint32_t offs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) offs[i] = swz(row0 + 32 * i, c16);
    int hs = 0, ls = 0;
    uint32_t hphase = 0, lphase = 0;
#pragma unroll 1
    for (;;) {
     const int64_t tile = next_tile();
     if (tile < 0) break;
#pragma unroll 1
     for (int kb = 0; kb < kKBlocks; ++kb) {
      mbar_wait(bar_raw + 8 * hs, hphase);                      // the raw k-block has landed
      mbar_wait(bar_empty_lo + 8 * ls, lphase ^ 1);             // the lo slot's previous user has retired
      const uint32_t hi = sHi + hs * kV3SlotBytes, lo = sLo + ls * kV3SlotBytes;
      float sc[4] = {1.0f, 1.0f, 1.0f, 1.0f};
      if (p.pre_scale != nullptr) {
        const int64_t r0 = tile * kV3Rows + row0;
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[i] = (r0 + 32 * i < p.M) ? __ldg(p.pre_scale + r0 + 32 * i) : 1.0f;
      }
      // all four loads first (the volatile asm keeps program order: interleaved with the stores the chain is
      // LDS -> residual -> STS four times over, and the MMA issuer waits for exactly this warp group k-block by k-block)
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = lds128(hi + offs[i]);
      if (p.pre_scale != nullptr) {          // (row_scale ⊙ A) as an individually rounded fp32 product, then the split
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i].x = __fmul_rn(sc[i], v[i].x); v[i].y = __fmul_rn(sc[i], v[i].y);
          v[i].z = __fmul_rn(sc[i], v[i].z); v[i].w = __fmul_rn(sc[i], v[i].w);
          sts128(hi + offs[i], v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        sts128(lo + offs[i], make_float4(tf32_trunc_residual(v[i].x), tf32_trunc_residual(v[i].y),
                                         tf32_trunc_residual(v[i].z), tf32_trunc_residual(v[i].w)));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full_lo + 8 * ls);
      if (++hs == kV3HiSlots) { hs = 0; hphase ^= 1; }
      if (++ls == kV3LoSlots) { ls = 0; lphase ^= 1; }
     }
    }
  } else if (warp == kV2MmaWarp) {
    // =========================== MMA ISSUER ===========================
    constexpr uint32_t idesc = make_idesc(128, kV3Rows);
    int acc = 0, ls = 0, hs0 = 0;                              // hs0: hi slot of the tile's first k-block
    uint32_t acc_phase = 0, lphase = 0;
#pragma unroll 1
    for (;;) {
      if (next_tile() < 0) break;
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kV3Rows);
      // pass 1: cross terms, k-block by k-block (the accumulator is still small: truncation costs 2^-11 of what it
      // would cost after the hi*hi terms)
#pragma unroll
      for (int kb = 0; kb < kKBlocks; ++kb) {
        mbar_wait(bar_full_lo + 8 * ls, lphase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = sHi + ((hs0 + kb) % kV3HiSlots) * kV3SlotBytes;
          const uint32_t a_lo = sLo + ls * kV3SlotBytes;
#pragma unroll
          for (int j = 0; j < kKB / 8; ++j) {
            const uint64_t dah = make_smem_desc(a_hi + j * 32);
            const uint64_t dal = make_smem_desc(a_lo + j * 32);
            const uint32_t w_hi = tmem_base + kWhiCol + (uint32_t)(kb * kKB + j * 8);
            const uint32_t w_lo = tmem_base + kWloCol + (uint32_t)(kb * kKB + j * 8);
            umma_tf32_ts(d_tmem, w_hi, dal, idesc, (kb | j) != 0 ? 1u : 0u);
            umma_tf32_ts(d_tmem, w_lo, dah, idesc, 1u);
          }
          umma_commit(bar_empty_lo + 8 * ls);
        }
        __syncwarp();
        if (++ls == kV3LoSlots) { ls = 0; lphase ^= 1; }
      }
      // pass 2: the dominant hi*hi terms
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) {
          const uint32_t a_hi = sHi + ((hs0 + kb) % kV3HiSlots) * kV3SlotBytes;
#pragma unroll
          for (int j = 0; j < kKB / 8; ++j) {
            const uint64_t dah = make_smem_desc(a_hi + j * 32);
            const uint32_t w_hi = tmem_base + kWhiCol + (uint32_t)(kb * kKB + j * 8);
            umma_tf32_ts(d_tmem, w_hi, dah, idesc, 1u);
          }
        }
#pragma unroll
        for (int kb = 0; kb < kKBlocks; ++kb) umma_commit(bar_empty_hi + 8 * ((hs0 + kb) % kV3HiSlots));
        umma_commit(bar_acc_full + 8 * acc);
      }
      __syncwarp();
      hs0 = (hs0 + kKBlocks) % kV3HiSlots;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================== EPILOGUE ===========================
    const int quad = warp & 3, half = warp >> 2;     // TMEM lane quadrant, row half (64 rows)
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (!kDual) {
      constexpr bool kNeedAux = (MODE == kV2GradPwl || MODE == kV2GradSmooth);
      constexpr bool kNeedBias = (MODE == kV2BiasPwl || MODE == kV2BiasSmooth);
      constexpr bool kNeedOld = (MODE == kV2Accumulate || MODE == kV2AccumulateScaled);
      constexpr bool kScaled = (MODE == kV2AccumulateScaled);
      const int f = quad * 32 + lane;
      const bool live = quad * 32 < NOUT;
      const float bias_f = (kNeedBias && live && p.bias != nullptr) ? __ldg(p.bias + f) : 0.0f;
#pragma unroll 1
      for (;;) {
        const int64_t tile = next_tile();
        if (tile < 0) break;
        if (!live) {
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kV3Rows + half * 64);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int64_t r0 = tile * kV3Rows + half * 64 + c * 32;
          const int nvalid = (int)((p.M - r0) < 32 ? (p.M - r0) : 32);     // warp-uniform, may be <= 0
          float* dst = p.D + r0 * p.ldd + f;
          float tt[32];
          if constexpr (kNeedAux || kNeedOld) {
            // streamed epilogue operand of the chunk requested before the accumulator is touched (first chunk: before
            // waiting for this tile's MMAs, so its DRAM latency hides behind them)
            const float* src = kNeedAux ? p.aux + r0 * p.ld_aux + f : dst;
            const int64_t lds = kNeedAux ? p.ld_aux : p.ldd;
            if (nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) tt[j] = src[j * lds];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) tt[j] = j < nvalid ? src[j * lds] : 0.0f;
            }
          }
          float sc_l = 1.0f;
          if (kScaled && lane < nvalid) sc_l = __ldg(p.scale + r0 + lane);
          if (c == 0) {
            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            tc_fence_after();
          }
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
          }
          asm volatile("" : "+l"(dst));      // accumulate: do not keep the 32 load addresses alive for the stores
          if constexpr (kScaled) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(__shfl_sync(0xffffffffu, sc_l, j), v[j]);
          }
          if (nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              dst[j * p.ldd] = v2_epilogue_op<MODE>(v[j], bias_f, kNeedAux ? tt[j] : 0.0f, kNeedOld ? tt[j] : 0.0f, p.slope, p.act);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid)
                dst[j * p.ldd] = v2_epilogue_op<MODE>(v[j], bias_f, kNeedAux ? tt[j] : 0.0f, kNeedOld ? tt[j] : 0.0f, p.slope, p.act);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
      // lane 2i: accumulator of W1 (a1) for feature f, lane 2i+1: accumulator of W2 (a2) for the same feature
      const int kind = lane & 1;
      const int f = fh * 64 + quad * 16 + (lane >> 1);
#pragma unroll 1
      for (;;) {
        const int64_t tile = next_tile();
        if (tile < 0) break;
        const int64_t rt = tile * kV3Rows + half * 64;                    // first of this warp's 64 rows
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kV3Rows + half * 64);
        if constexpr (MODE == kV2DualSeparate) {
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          tc_fence_after();
          float* out = kind ? p.D2 : p.D;
          const int64_t ldo = kind ? p.ldd2 : p.ldd;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld32(t_lane + c * 32, v);
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
            }
            const int64_t r0 = rt + c * 32;
            float* dst = out + r0 * ldo + f;
            const int nvalid = (int)((p.M - r0) < 32 ? (p.M - r0) : 32);
            if (nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[j * ldo] = v[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) dst[j * ldo] = v[j];
            }
          }
        } else {
          // Of every 32-row chunk the even lane finalises rows 0..15 and the odd lane rows 16..31 of feature f: one
          // shuffle per output swaps the half each lane does not finalise.  Same operations in the same order as the two
          // launches this kernel replaces: (S + c*P) resp. ((old + a1) + c*a2).
          const int64_t rmine = rt + kind * 16;                             // chunk c: rows rmine + 32 c + jj
          const int nv = (int)((p.M - rmine) < 64 ? (p.M - rmine) : 64);    // rows c*32 + jj < nv exist (may be <= 0)
          const bool full = rt + 64 <= p.M;                                 // warp-uniform fast path
          float* dst = p.D + rmine * p.ldd + f;
          float old[32];
          if constexpr (MODE == kV2DualAccumulate) {
            // requested before waiting for this tile's MMAs: the DRAM latency hides behind them
            if (full) {
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) old[c * 16 + jj] = dst[(int64_t)(c * 32 + jj) * p.ldd];
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                  old[c * 16 + jj] = (c * 32 + jj < nv) ? dst[(int64_t)(c * 32 + jj) * p.ldd] : 0.0f;
            }
          }
          float sc_l[2] = {1.0f, 1.0f};
          if (p.scale != nullptr) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if (rt + c * 32 + lane < p.M) sc_l[c] = __ldg(p.scale + rt + c * 32 + lane);
          }
          // The epilogue is the critical path of the accumulate form and has no register room to hold the next tile's
          // operands, so it asks L2 for them now (64-byte segment of rows lane, lane + 32; the row scales): the profile
          // showed a fifth of the epilogue's time waiting for exactly these loads at DRAM latency.  The next tile id is
          // taken from the queue only if the TMA thread has published it already (it runs 2-3 tiles ahead).
          if (mbar_test(bar_tile + 8 * tq, tq_phase)) {
            const int32_t tn = (int32_t)lds32u(s_tile + 4 * tq);
            if (tn >= 0) {
              const int64_t rn = (int64_t)tn * kV3Rows + half * 64;
              if constexpr (MODE == kV2DualAccumulate) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                  if (rn + c * 32 + lane < p.M)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.D + (rn + c * 32 + lane) * p.ldd + fh * 64 + quad * 16));
              }
              if (p.scale != nullptr && lane < 2 && rn + lane * 32 < p.M)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.scale + rn + lane * 32));
            }
          }
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          tc_fence_after();
          asm volatile("" : "+l"(dst));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld32(t_lane + c * 32, v);
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float got = __shfl_xor_sync(0xffffffffu, kind ? v[jj] : v[jj + 16], 1);
              const float a1 = kind ? got : v[jj];
              const float a2 = kind ? v[jj + 16] : got;
              const float cs = __shfl_sync(0xffffffffu, sc_l[c], jj + kind * 16);
              float r = (MODE == kV2DualAccumulate) ? __fadd_rn(old[c * 16 + jj], a1) : a1;
              r = __fadd_rn(r, __fmul_rn(cs, a2));
              if (full || c * 32 + jj < nv) dst[(int64_t)(c * 32 + jj) * p.ldd] = r;
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kV2MmaWarp) tmem_dealloc(tmem_base, kTmemCols);
  if (p.tile_ctr != nullptr && threadIdx.x == 0) {
    // every CTA has made its last fetch: the last one to get here leaves the counters zeroed for the next launch
    __threadfence();
    if (atomicAdd(p.tile_ctr + 2, 1u) == gridDim.x - 1) {
      p.tile_ctr[0] = 0u; p.tile_ctr[1] = 0u;
      __threadfence();
      p.tile_ctr[2] = 0u;
    }
  }
}

// TMA descriptor of the streamed operand: fp32 [M rows x K], box = 64 rows x 32 floats, 128-byte swizzle
static bool make_tmap_rows64(CUtensorMap* tmap, const float* A, int64_t lda, int64_t M, int K) {
  TmapEncodeFn enc = tmap_encoder();
  if (enc == nullptr || M > 0x7fffffffLL) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)kV2Rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Counters of the dynamic tile scheduler: [tiles handed out to feature half 0, half 1, CTAs finished, pad] per slot.
// A launch leaves its slot zeroed (the last CTA resets it), so a slot may be shared by launches that can never run
// CONCURRENTLY: eager launches get the slot of their STREAM (kernels of one stream are serialised); launches recorded
// into a CUDA graph keep their slot for the life of the graph and may be replayed on any stream, so each gets one of
// its own from a separate range.  When a range is used up the launch falls back to the static schedule.
constexpr int kCtrEager = 1024, kCtrPool = 4096;
__device__ unsigned int g_tile_ctr[kCtrPool][4];

static unsigned int* tile_counters(cudaStream_t stream) {
  static const bool on = [] { const char* e = getenv("DMP_GEMM_DYNAMIC"); return e ? atoi(e) != 0 : true; }();
  if (!on) return nullptr;
  static std::mutex mu;
  static unsigned int* base[64] = {};
  static std::unordered_map<cudaStream_t, unsigned> slot_of_stream[64];
  static unsigned next_captured = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (base[dev] == nullptr) {
    void* ptr = nullptr;
    if (cudaGetSymbolAddress(&ptr, g_tile_ctr) != cudaSuccess) return nullptr;
    base[dev] = static_cast<unsigned int*>(ptr);
  }
  unsigned slot;
  if (st == cudaStreamCaptureStatusNone) {
    auto& map = slot_of_stream[dev];
    auto it = map.find(stream);
    if (it == map.end()) {
      if (map.size() >= (size_t)kCtrEager) return nullptr;
      it = map.emplace(stream, (unsigned)map.size()).first;
    }
    slot = it->second;
  } else {
    if (next_captured >= (unsigned)(kCtrPool - kCtrEager)) return nullptr;
    slot = kCtrEager + next_captured++;
  }
  return base[dev] + 4 * slot;
}

static bool v3_enabled() {   // DMP_GEMM_V3=0: 64-row kernel everywhere (A/B runs)
  static const bool on = [] { const char* e = getenv("DMP_GEMM_V3"); return e ? atoi(e) != 0 : true; }();
  return on;
}

template <int NOUT, int K, int MODE>
static int launch_v2_mode(const V2Params& p, cudaStream_t stream) {
  using L = V2Smem<K>;
  constexpr int kHalves = (MODE >= kV2DualStore) ? NOUT / 64 : 1;
  const int64_t streams = persistent_sms() / kHalves;
  V2Params q = p;
  q.tile_ctr = nullptr;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (v3_enabled() && tma_enabled() && p.M >= kV3Rows && make_tmap_rows(&tmap, p.A, p.lda, p.M, K)) {
    static bool configured3 = false;
    if (!configured3) {
      cudaError_t e = cudaFuncSetAttribute(tf32x3_gemm_v3_kernel<NOUT, K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kV3Smem);
      if (e != cudaSuccess) {
        set_error("gemm_tf32x3 (v3): cannot reserve %d bytes of shared memory: %s", kV3Smem, cudaGetErrorString(e));
        return DMP_ERR_CUDA;
      }
      configured3 = true;
    }
    const int64_t tiles = (p.M + kV3Rows - 1) / kV3Rows;
    const unsigned grid = (unsigned)((tiles < streams ? tiles : streams) * kHalves);
    q.use_tma = 1;
    q.tile_ctr = tile_counters(stream);
    tf32x3_gemm_v3_kernel<NOUT, K, MODE><<<grid, kV3Threads, kV3Smem, stream>>>(q, tmap);
    return launch_status("tf32x3_gemm_v3_kernel");
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tf32x3_gemm_v2_kernel<NOUT, K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm_tf32x3 (v2): cannot reserve %d bytes of shared memory: %s", L::kTotal, cudaGetErrorString(e));
      return DMP_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t tiles = (p.M + kV2Rows - 1) / kV2Rows;
  const unsigned grid = (unsigned)((tiles < streams ? tiles : streams) * kHalves);
  q.use_tma = (tma_enabled() && p.M >= kV2Rows && make_tmap_rows64(&tmap, p.A, p.lda, p.M, K)) ? 1 : 0;
  tf32x3_gemm_v2_kernel<NOUT, K, MODE><<<grid, kV2Threads, L::kTotal, stream>>>(q, tmap);
  return launch_status("tf32x3_gemm_v2_kernel");
}

template <int NOUT, int K>
static int launch_v2_single(const V2Params& p, int mode, cudaStream_t s) {
  switch (mode) {
    case kV2Store: return launch_v2_mode<NOUT, K, kV2Store>(p, s);
    case kV2Accumulate: return launch_v2_mode<NOUT, K, kV2Accumulate>(p, s);
    case kV2BiasPwl: return launch_v2_mode<NOUT, K, kV2BiasPwl>(p, s);
    case kV2GradPwl: return launch_v2_mode<NOUT, K, kV2GradPwl>(p, s);
    case kV2BiasSmooth: return launch_v2_mode<NOUT, K, kV2BiasSmooth>(p, s);
    case kV2AccumulateScaled: return launch_v2_mode<NOUT, K, kV2AccumulateScaled>(p, s);
    default: return launch_v2_mode<NOUT, K, kV2GradSmooth>(p, s);
  }
}

template <int NOUT, int K>
static int launch_v2_dual(const V2Params& p, int mode, cudaStream_t s) {
  if (mode == DMP_DUAL_STORE) return launch_v2_mode<NOUT, K, kV2DualStore>(p, s);
  if (mode == DMP_DUAL_ACCUMULATE) return launch_v2_mode<NOUT, K, kV2DualAccumulate>(p, s);
  return launch_v2_mode<NOUT, K, kV2DualSeparate>(p, s);
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
  DMP_CHECK_ARG((epilogue & ~(15 | DMP_EPI_MUL_ACT_GRAD | DMP_EPI_ACCUMULATE)) == 0, "gemm_tf32x3: unknown epilogue bits 0x%x",
                epilogue);
  const bool mul_grad = (epilogue & DMP_EPI_MUL_ACT_GRAD) != 0;
  const bool accumulate = (epilogue & DMP_EPI_ACCUMULATE) != 0;
  DMP_CHECK_ARG(!mul_grad || (aux != nullptr && ld_aux >= N && ld_aux % 4 == 0), "gemm_tf32x3: act' epilogue needs aux");
  DMP_CHECK_ARG(A != D, "gemm_tf32x3: D must not alias A");
  DMP_CHECK_ARG(!(accumulate && (mul_grad || bias != nullptr || act != DMP_ACT_NONE)),
                "gemm_tf32x3: accumulate cannot be combined with bias / activation epilogues");
  DMP_CHECK_ARG(!(mul_grad && bias != nullptr), "gemm_tf32x3: act' epilogue takes no bias");
  const bool smooth = (act == DMP_ACT_TANH || act == DMP_ACT_SIGMOID);
  if (act == DMP_ACT_NONE) slope = 1.0f;   // piecewise-linear family: none = slope 1, relu = slope 0
  if (act == DMP_ACT_RELU) slope = 0.0f;
  int mode;
  if (accumulate) mode = row_scale != nullptr ? kV2AccumulateScaled : kV2Accumulate;
  else if (mul_grad) mode = smooth ? kV2GradSmooth : (act == DMP_ACT_NONE ? kV2Store : kV2GradPwl);
  else if (smooth) mode = kV2BiasSmooth;
  else if (bias != nullptr || act != DMP_ACT_NONE) mode = kV2BiasPwl;
  else mode = kV2Store;
  V2Params p;
  p.A = A; p.lda = lda; p.W1 = Bt; p.W2 = nullptr; p.ldw = ldb; p.bias = bias; p.aux = aux; p.ld_aux = ld_aux;
  p.D = D; p.ldd = ldd; p.D2 = nullptr; p.ldd2 = 0; p.M = M; p.act = act; p.slope = slope; p.use_tma = 0;
  // accumulate: the row scale rides on the accumulated row (D += s_r * acc_r); otherwise on the rows of A
  p.scale = accumulate ? row_scale : nullptr;
  p.pre_scale = accumulate ? nullptr : row_scale;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 128) return K == 128 ? launch_v2_single<128, 128>(p, mode, s) : launch_v2_single<128, 64>(p, mode, s);
  return K == 128 ? launch_v2_single<64, 128>(p, mode, s) : launch_v2_single<64, 64>(p, mode, s);
}

extern "C" int dmp_gemm_tf32x3_dual(const float* A, int64_t lda, const float* W1t, const float* W2t, int64_t ldw,
                                    const float* row_scale, float* D, int64_t ldd, float* D2, int64_t ldd2,
                                    int64_t M, int64_t N, int64_t K, int mode, void* stream) {
  using namespace dmp;
  using namespace dmp::gemm;
  DMP_CHECK_ARG(M >= 0, "gemm_tf32x3_dual: negative M");
  if (M == 0) return DMP_OK;
  DMP_CHECK_ARG(A && W1t && W2t && D, "gemm_tf32x3_dual: null pointer");
  DMP_CHECK_ARG(mode == DMP_DUAL_STORE || mode == DMP_DUAL_ACCUMULATE || mode == DMP_DUAL_SEPARATE,
                "gemm_tf32x3_dual: bad mode %d", mode);
  DMP_CHECK_ARG((N == 64 || N == 128) && (K == 64 || K == 128), "gemm_tf32x3_dual: N and K must be 64 or 128 (got %lld, %lld)",
                (long long)N, (long long)K);
  DMP_CHECK_ARG(lda >= K && ldw >= K && ldd >= N && lda % 4 == 0 && ldw % 4 == 0,
                "gemm_tf32x3_dual: leading dimensions must be >= the row length (A, W: multiples of 4)");
  DMP_CHECK_ARG(aligned_to(A, 16) && aligned_to(W1t, 16) && aligned_to(W2t, 16),
                "gemm_tf32x3_dual: A and the weights must be 16-byte aligned");
  DMP_CHECK_ARG(mode != DMP_DUAL_SEPARATE || (D2 != nullptr && ldd2 >= N && row_scale == nullptr),
                "gemm_tf32x3_dual: separate mode needs D2 and takes no row scale");
  DMP_CHECK_ARG(A != D && A != D2, "gemm_tf32x3_dual: outputs must not alias A");
  V2Params p;
  p.A = A; p.lda = lda; p.W1 = W1t; p.W2 = W2t; p.ldw = ldw; p.scale = row_scale; p.pre_scale = nullptr; p.bias = nullptr;
  p.aux = nullptr;
  p.ld_aux = 0; p.D = D; p.ldd = ldd; p.D2 = D2; p.ldd2 = ldd2; p.M = M; p.act = 0; p.slope = 1.0f; p.use_tma = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 128 && K == 128) return launch_v2_dual<128, 128>(p, mode, s);
  if (N == 128 && K == 64) return launch_v2_dual<128, 64>(p, mode, s);
  if (N == 64 && K == 128) return launch_v2_dual<64, 128>(p, mode, s);
  return launch_v2_dual<64, 64>(p, mode, s);
}
