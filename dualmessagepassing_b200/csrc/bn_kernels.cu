// bn_kernels.cu -- BatchNorm1d inside the DMPNN MLPs (rows A4/A5/A8: dmpnn.py:45-52 with batch_norm=True,
// UNC model.py:145-156), as deterministic column reductions + elementwise kernels for sm_100a.
//
// The MLP is Linear -> BatchNorm1d -> act -> Linear; the statistics run over ALL node rows / ALL edge rows, so
// nothing after the first Linear can be fused into the producing GEMM.  What is left is HBM-bound column work:
//
//   bn_stats            two passes over x [rows,H]: column sums -> mean, centred squares -> biased variance
//                       (two-pass, not E[x^2]-mean^2: no cancellation when |mean| >> std)
//   bn_act              y = act(((x - mean) * invstd) * gamma + beta)
//   bn_backward_reduce  sum_g[c] = sum_r g[r,c],  sum_gx[c] = sum_r g[r,c] * xhat[r,c]      (= dbeta, dgamma)
//   bn_backward_apply   gx = gamma*invstd * (g - sum_g/R - xhat*sum_gx/R)   (training)   |   g*gamma*invstd  (eval)
//
// Reductions: every CTA owns a contiguous block of rows, thread (tx, ty) accumulates rows ty, ty+8, ... of columns
// tx, tx+32, ... in ascending order, the 8 row-lanes are combined in a fixed order, the per-CTA partials go to
// global memory and the LAST CTA to finish (integer ticket, no float atomics) adds them in CTA order ->
// bit-reproducible run to run.
#include "common.cuh"

namespace dmp {

constexpr int kBnTx = 32, kBnTy = 8;
constexpr int kBnMaxCols = 4;                 // columns per thread per pass: H <= 128 in one pass, wider rows loop
constexpr int kBnRowBatchMax = 8;             // rows whose loads are in flight together per thread and column

struct BnReduceParams {
  const float* x; int64_t ldx;                // reduced matrix (stats: x; backward: g)
  const float* y; int64_t ldy;                // second operand (backward: pre-BN x) or NULL
  const float* mean;                          // [H] centre (pass 2 of stats; backward) or NULL
  const float* invstd;                        // [H] (backward) or NULL
  float* partial;                             // [grid][2][H]
  unsigned int* ticket;
  float* out0;                                // [H] final result 0
  float* out1;                                // [H] final result 1 or NULL
  int64_t rows;
  int H;
  int rows_per_cta;
  float scale0;                               // out0 = scale0 * sum
};

// MODE 0: sum x          MODE 1: sum (x - mean)^2          MODE 2: sum g  and  sum g * (y - mean) * invstd
template <int MODE>
__global__ void __launch_bounds__(kBnTx * kBnTy) bn_reduce_kernel(const BnReduceParams p) {
  __shared__ float red[2][kBnTy][kBnTx * kBnMaxCols];
  __shared__ bool is_last;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r1 = (r0 + p.rows_per_cta < p.rows) ? r0 + p.rows_per_cta : p.rows;
  for (int c0 = 0; c0 < p.H; c0 += kBnTx * kBnMaxCols) {
    float a0[kBnMaxCols], a1[kBnMaxCols], mu[kBnMaxCols], is[kBnMaxCols];
#pragma unroll
    for (int k = 0; k < kBnMaxCols; ++k) {
      const int c = c0 + tx + kBnTx * k;
      a0[k] = a1[k] = 0.0f;
      mu[k] = (MODE != 0 && c < p.H) ? __ldg(p.mean + c) : 0.0f;
      is[k] = (MODE == 2 && c < p.H) ? __ldg(p.invstd + c) : 1.0f;
    }
    // kBnRowBatch rows per trip: all their loads are issued before the first add (one row per trip left a warp with
    // 2-4 loads in flight: 1.3 TB/s on [180 k, 64]); the adds keep the ascending row order
    constexpr int kBnRowBatch = (MODE == 2) ? kBnRowBatchMax / 2 : kBnRowBatchMax;   // MODE 2 streams two matrices
    auto accumulate = [&](float v, float y, int k) {
      if (MODE == 0) {
        a0[k] = __fadd_rn(a0[k], v);
      } else if (MODE == 1) {
        const float d = __fsub_rn(v, mu[k]);
        a0[k] = __fadd_rn(a0[k], __fmul_rn(d, d));
      } else {
        const float xh = __fmul_rn(__fsub_rn(y, mu[k]), is[k]);
        a0[k] = __fadd_rn(a0[k], v);
        a1[k] = __fadd_rn(a1[k], __fmul_rn(v, xh));
      }
    };
    int64_t r = r0 + ty;
    for (; r + (kBnRowBatch - 1) * kBnTy < r1; r += kBnRowBatch * kBnTy) {
#pragma unroll
      for (int k = 0; k < kBnMaxCols; ++k) {
        const int c = c0 + tx + kBnTx * k;
        if (c < p.H) {
          float v[kBnRowBatch], y[kBnRowBatch];
#pragma unroll
          for (int u = 0; u < kBnRowBatch; ++u) {
            v[u] = __ldg(p.x + (r + u * kBnTy) * p.ldx + c);
            y[u] = (MODE == 2) ? __ldg(p.y + (r + u * kBnTy) * p.ldy + c) : 0.0f;
          }
#pragma unroll
          for (int u = 0; u < kBnRowBatch; ++u) accumulate(v[u], y[u], k);
        }
      }
    }
    for (; r < r1; r += kBnTy) {
#pragma unroll
      for (int k = 0; k < kBnMaxCols; ++k) {
        const int c = c0 + tx + kBnTx * k;
        if (c < p.H) accumulate(__ldg(p.x + r * p.ldx + c), (MODE == 2) ? __ldg(p.y + r * p.ldy + c) : 0.0f, k);
      }
    }
#pragma unroll
    for (int k = 0; k < kBnMaxCols; ++k) {
      red[0][ty][tx + kBnTx * k] = a0[k];
      red[1][ty][tx + kBnTx * k] = a1[k];
    }
    __syncthreads();
    if (ty < (MODE == 2 ? 2 : 1)) {
#pragma unroll
      for (int k = 0; k < kBnMaxCols; ++k) {
        const int c = c0 + tx + kBnTx * k;
        if (c < p.H) {
          float s = 0.0f;
          for (int j = 0; j < kBnTy; ++j) s = __fadd_rn(s, red[ty][j][tx + kBnTx * k]);
          p.partial[((int64_t)blockIdx.x * 2 + ty) * p.H + c] = s;
        }
      }
    }
    __syncthreads();
  }
  // last CTA to arrive adds the partials in CTA order
  __threadfence();
  __syncthreads();
  if (tx == 0 && ty == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // four lanes per output: lane j adds the partials j, j+4, ... in ascending CTA order (8 independent L2 reads per
  // trip), then (s0 + s1) + (s2 + s3) -- a fixed order; one thread per output was a chain of ~600 dependent reads
  const int t = ty * kBnTx + tx;
  const int nout = p.H * (MODE == 2 ? 2 : 1);
  for (int o0 = 0; o0 < nout; o0 += kBnTx * kBnTy / 4) {
    const int o = o0 + t / 4, j = t % 4;
    const bool live = o < nout;
    const int which = live ? o / p.H : 0, col = live ? o % p.H : 0;
    const float* src = p.partial + (int64_t)which * p.H + col;
    float s = 0.0f;
    if (live) {
      unsigned b = j;
      for (; b + 28 < gridDim.x; b += 32) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(b + 4 * u) * 2 * p.H);
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
      }
      for (; b < gridDim.x; b += 4) s = __fadd_rn(s, __ldcg(src + (int64_t)b * 2 * p.H));
    }
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    if (live && j == 0) {
      if (which == 0) p.out0[col] = __fmul_rn(s, p.scale0);
      else p.out1[col] = s;
    }
  }
  if (t == 0) *p.ticket = 0u;   // self-cleaning: the next launch on this workspace starts from zero
}

struct BnApplyParams {
  const float* x; int64_t ldx;     // forward: pre-BN x;           backward: g (dL/dy)
  const float* y; int64_t ldy;     // backward: pre-BN x
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  const float* sum_g; const float* sum_gx;
  float* out; int64_t ld_out;
  int64_t rows; int H; int act; float slope; float inv_rows; int training;
};

template <bool BWD>
__global__ void __launch_bounds__(kBnTx * kBnTy) bn_apply_kernel(const BnApplyParams p) {
  // warp = one row at a time, lanes along the columns (128-byte coalesced segments); per-column constants hit L1
  for (int64_t r = (int64_t)blockIdx.x * kBnTy + threadIdx.y; r < p.rows; r += (int64_t)gridDim.x * kBnTy) {
    for (int c = threadIdx.x; c < p.H; c += kBnTx) {
      const float mu = __ldg(p.mean + c), is = __ldg(p.invstd + c);
      const float ga = p.gamma ? __ldg(p.gamma + c) : 1.0f;
      if (!BWD) {
        float v = __fmul_rn(__fmul_rn(__fsub_rn(p.x[r * p.ldx + c], mu), is), ga);
        if (p.beta) v = __fadd_rn(v, __ldg(p.beta + c));
        p.out[r * p.ld_out + c] = apply_act(v, p.act, p.slope);
      } else {
        const float g = p.x[r * p.ldx + c];
        float v = g;
        if (p.training) {
          const float xh = __fmul_rn(__fsub_rn(__ldg(p.y + r * p.ldy + c), mu), is);
          v = __fsub_rn(__fsub_rn(g, __fmul_rn(__ldg(p.sum_g + c), p.inv_rows)),
                        __fmul_rn(xh, __fmul_rn(__ldg(p.sum_gx + c), p.inv_rows)));
        }
        p.out[r * p.ld_out + c] = __fmul_rn(__fmul_rn(ga, is), v);
      }
    }
  }
}

// Vector form of the two elementwise kernels (H % 4 == 0, 16-byte aligned rows, H <= 128): a thread owns four adjacent
// columns -- their constants are loaded once -- and walks rows in batches of four whose loads are all issued before the
// first store.  Same operations per element as bn_apply_kernel, so the results are bit-identical; the scalar kernel
// kept one 4-byte load in flight per thread (0.6 of HBM on an L2-sized matrix, far less on a 20 GB one).
constexpr int kBnVecRows = 4;
template <bool BWD>
__global__ void __launch_bounds__(256) bn_apply_vec_kernel(const BnApplyParams p) {
  const int groups = p.H >> 2;                       // float4 column groups per row (<= 32)
  const int gx = threadIdx.x % groups, gy = threadIdx.x / groups, ny = 256 / groups;
  if (gy >= ny) return;
  const int c = gx * 4;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(p.mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(p.invstd + c));
  const float4 ga = p.gamma ? __ldg(reinterpret_cast<const float4*>(p.gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 be = (!BWD && p.beta) ? __ldg(reinterpret_cast<const float4*>(p.beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sgx = sg;
  if (BWD && p.training) {
    sg = __ldg(reinterpret_cast<const float4*>(p.sum_g + c));
    sgx = __ldg(reinterpret_cast<const float4*>(p.sum_gx + c));
  }
  const bool has_beta = p.beta != nullptr;
  auto one = [&](float x, float y, float m, float i, float g, float b, float s0, float s1) -> float {
    if (!BWD) {
      float v = __fmul_rn(__fmul_rn(__fsub_rn(x, m), i), g);
      if (has_beta) v = __fadd_rn(v, b);
      return apply_act(v, p.act, p.slope);
    }
    float v = x;
    if (p.training) {
      const float xh = __fmul_rn(__fsub_rn(y, m), i);
      v = __fsub_rn(__fsub_rn(x, __fmul_rn(s0, p.inv_rows)), __fmul_rn(xh, __fmul_rn(s1, p.inv_rows)));
    }
    return __fmul_rn(__fmul_rn(g, i), v);
  };
  const int64_t stride = (int64_t)gridDim.x * ny;
  for (int64_t r0 = (int64_t)blockIdx.x * ny + gy; r0 < p.rows; r0 += stride * kBnVecRows) {
    float4 x[kBnVecRows], y[kBnVecRows];
#pragma unroll
    for (int u = 0; u < kBnVecRows; ++u) {
      const int64_t r = r0 + u * stride;
      x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.rows) {
        x[u] = *reinterpret_cast<const float4*>(p.x + r * p.ldx + c);
        if (BWD && p.training) y[u] = __ldg(reinterpret_cast<const float4*>(p.y + r * p.ldy + c));
      }
    }
#pragma unroll
    for (int u = 0; u < kBnVecRows; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < p.rows) {
        float4 o;
        o.x = one(x[u].x, y[u].x, mu.x, is.x, ga.x, be.x, sg.x, sgx.x);
        o.y = one(x[u].y, y[u].y, mu.y, is.y, ga.y, be.y, sg.y, sgx.y);
        o.z = one(x[u].z, y[u].z, mu.z, is.z, ga.z, be.z, sg.z, sgx.z);
        o.w = one(x[u].w, y[u].w, mu.w, is.w, ga.w, be.w, sg.w, sgx.w);
        *reinterpret_cast<float4*>(p.out + r * p.ld_out + c) = o;
      }
    }
  }
}

static bool apply_vec_ok(const BnApplyParams& p) {
  const bool al = aligned_to(p.x, 16) && aligned_to(p.y, 16) && aligned_to(p.out, 16) && aligned_to(p.mean, 16) &&
                  aligned_to(p.invstd, 16) && aligned_to(p.gamma, 16) && aligned_to(p.beta, 16) &&
                  aligned_to(p.sum_g, 16) && aligned_to(p.sum_gx, 16);
  return al && p.H % 4 == 0 && p.H <= 128 && (256 % (p.H / 4)) == 0 && p.ldx % 4 == 0 && p.ld_out % 4 == 0 &&
         (p.y == nullptr || p.ldy % 4 == 0);
}

template <bool BWD>
static void launch_apply(const BnApplyParams& p, unsigned grid_rows8, cudaStream_t stream) {
  if (apply_vec_ok(p)) {
    const int ny = 256 / (p.H / 4);
    int64_t need = (p.rows + (int64_t)ny * kBnVecRows - 1) / ((int64_t)ny * kBnVecRows);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (need < 1) need = 1;
    bn_apply_vec_kernel<BWD><<<(unsigned)(need < cap ? need : cap), 256, 0, stream>>>(p);
  } else {
    bn_apply_kernel<BWD><<<grid_rows8, dim3(kBnTx, kBnTy), 0, stream>>>(p);
  }
}

static int reduce_grid(int64_t rows, int* rows_per_cta) {
  int64_t per = (rows + (int64_t)kNumSMs * 4 - 1) / ((int64_t)kNumSMs * 4);
  if (per < 64) per = 64;
  *rows_per_cta = (int)per;
  return (int)((rows + per - 1) / per);
}

static unsigned apply_grid(int64_t rows) {
  int64_t need = (rows + kBnTy - 1) / kBnTy;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (unsigned)(need < cap ? need : cap);
}

}  // namespace dmp

extern "C" int dmp_bn_workspace_bytes(int64_t H, int64_t* bytes_host) {
  using namespace dmp;
  DMP_CHECK_ARG(bytes_host != nullptr && H >= 0, "bn_workspace_bytes: bad arguments");
  *bytes_host = 256 + (int64_t)kNumSMs * 4 * 2 * H * 4;   // ticket (zero-initialised by the caller ONCE) + partials
  return DMP_OK;
}

extern "C" int dmp_bn_stats(const float* x, int64_t ldx, int64_t rows, int64_t H, float* mean, float* var,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(rows > 0 && H > 0 && H < (1 << 20), "bn_stats: needs at least one row and one column");
  DMP_CHECK_ARG(x && mean && var && ldx >= H, "bn_stats: bad operands");
  int64_t need = 0;
  dmp_bn_workspace_bytes(H, &need);
  DMP_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "bn_stats: workspace too small");
  BnReduceParams p;
  p.x = x; p.ldx = ldx; p.y = nullptr; p.ldy = 0; p.mean = nullptr; p.invstd = nullptr;
  p.ticket = static_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  p.rows = rows; p.H = (int)H; p.out1 = nullptr;
  const int grid = reduce_grid(rows, &p.rows_per_cta);
  const dim3 block(kBnTx, kBnTy);
  p.out0 = mean; p.scale0 = 1.0f / (float)rows;
  bn_reduce_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(p);
  int rc = launch_status("bn_reduce_kernel<0>");
  if (rc != DMP_OK) return rc;
  p.mean = mean; p.out0 = var;
  bn_reduce_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(p);
  return launch_status("bn_reduce_kernel<1>");
}

extern "C" int dmp_bn_act(const float* x, int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, float* out, int64_t ld_out, int64_t rows, int64_t H, int act, float slope,
                          void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(rows >= 0 && H >= 0, "bn_act: negative size");
  if (rows == 0 || H == 0) return DMP_OK;
  DMP_CHECK_ARG(x && mean && invstd && out && ldx >= H && ld_out >= H, "bn_act: bad operands");
  DMP_CHECK_ARG(act >= DMP_ACT_NONE && act <= DMP_ACT_SIGMOID, "bn_act: bad activation %d", act);
  BnApplyParams p;
  p.x = x; p.ldx = ldx; p.y = nullptr; p.ldy = 0; p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.beta = beta;
  p.sum_g = p.sum_gx = nullptr; p.out = out; p.ld_out = ld_out; p.rows = rows; p.H = (int)H; p.act = act;
  p.slope = slope; p.inv_rows = 0.0f; p.training = 0;
  launch_apply<false>(p, apply_grid(rows), (cudaStream_t)stream);
  return launch_status("bn_apply_kernel<fwd>");
}

extern "C" int dmp_bn_backward(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* mean,
                               const float* invstd, const float* gamma, float* gx, int64_t ld_gx, float* dgamma,
                               float* dbeta, int64_t rows, int64_t H, int training, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(rows > 0 && H > 0 && H < (1 << 20), "bn_backward: needs at least one row and one column");
  DMP_CHECK_ARG(g && x && mean && invstd && gx && dgamma && dbeta && ldg >= H && ldx >= H && ld_gx >= H,
                "bn_backward: bad operands");
  int64_t need = 0;
  dmp_bn_workspace_bytes(H, &need);
  DMP_CHECK_ARG(workspace != nullptr && workspace_bytes >= need, "bn_backward: workspace too small");
  BnReduceParams q;
  q.x = g; q.ldx = ldg; q.y = x; q.ldy = ldx; q.mean = mean; q.invstd = invstd;
  q.ticket = static_cast<unsigned int*>(workspace);
  q.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  q.rows = rows; q.H = (int)H; q.out0 = dbeta; q.out1 = dgamma; q.scale0 = 1.0f;
  const int grid = reduce_grid(rows, &q.rows_per_cta);
  bn_reduce_kernel<2><<<grid, dim3(kBnTx, kBnTy), 0, (cudaStream_t)stream>>>(q);
  int rc = launch_status("bn_reduce_kernel<2>");
  if (rc != DMP_OK) return rc;
  BnApplyParams p;
  p.x = g; p.ldx = ldg; p.y = x; p.ldy = ldx; p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.beta = nullptr;
  p.sum_g = dbeta; p.sum_gx = dgamma; p.out = gx; p.ld_out = ld_gx; p.rows = rows; p.H = (int)H; p.act = 0;
  p.slope = 0.0f; p.inv_rows = 1.0f / (float)rows; p.training = training;
  launch_apply<true>(p, apply_grid(rows), (cudaStream_t)stream);
  return launch_status("bn_apply_kernel<bwd>");
}
