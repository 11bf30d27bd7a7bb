// edge_kernels.cu -- per-edge streaming kernels of the DMPNN sparse core (sm_100a):
//   dmp_edge_update            forward edge state:   S + coef*P + (Qd[a] - Qs[b]) + ebias   (dmpnn.py:112-123,142-149)
//   dmp_edge_backward          backward of fn.sum + degree term:  T = sgn*gN[dst]*norm,  CG = coef*gE
//   dmp_gate_residual(+bwd)    act / gate / residual epilogue of the rep-net loop           (dmpnn.py:236-241,266-275)
//   dmp_permute_edge_scalar    per-edge scalar -> per-segment-position scalar
//
// Work decomposition: a group of G lanes owns one edge row (G*ITER*VEC >= H; a 128-wide fp32 row is one
// float4 per lane of a warp -> every access is a fully used 512-byte burst).  Persistent grid
// (a multiple of 148 SMs x resident CTAs), grid-stride over rows.  Edge-sized operands are touched once
// and use no-allocate / streaming accesses; the two node-table gathers use the default path so re-used
// rows can hit in L2.  Index/coef words are read once per row by all lanes of the group (broadcast).
#include "common.cuh"

namespace dmp {

static inline unsigned persistent_grid(int64_t rows, int groups_per_cta, int ctas_per_sm) {
  int64_t need = (rows + groups_per_cta - 1) / groups_per_cta;
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// ---- forward edge update ---------------------------------------------------------------------------
struct EdgeFwdParams {
  const int32_t* a32;
  const int32_t* b32;
  const float* coef;
  const float* S; int64_t ldS;
  const float* P; int64_t ldP;
  const float* Qd; int64_t ldQd;
  const float* Qs; int64_t ldQs;
  const float* ebias;
  float* out; int64_t ld_out;
  float* agg; int64_t ld_agg;
  int64_t E;
  int H;
  int order;
};

template <int VEC, int G, int ITER>
__global__ void __launch_bounds__(kThreads) edge_update_kernel(const EdgeFwdParams p) {
  constexpr int kGroups = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * kGroups;
  int col[ITER];
  bool ok[ITER];
  Row<VEC> bias[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    col[it] = (lane + it * G) * VEC;
    ok[it] = col[it] < p.H;
    if (ok[it] && p.ebias != nullptr) bias[it] = ld_row<VEC>(p.ebias + col[it]);
  }
  for (int64_t e = (int64_t)blockIdx.x * kGroups + threadIdx.x / G; e < p.E; e += stride) {
    const int a = __ldg(p.a32 + e);
    const int b = __ldg(p.b32 + e);
    const float c = __ldg(p.coef + e);
    const float* s_row = p.S + e * p.ldS;
    const bool has_p = p.P != nullptr;   // NULL: S already holds eloop + coef*P (fused dual projection, SCM order)
    const float* p_row = has_p ? p.P + e * p.ldP : nullptr;
    const float* qd_row = p.Qd + (int64_t)a * p.ldQd;
    const float* qs_row = p.Qs + (int64_t)b * p.ldQs;
    Row<VEC> s[ITER], pp[ITER], qd[ITER], qs[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      if (ok[it]) {
        qd[it] = ld_row<VEC>(qd_row + col[it]);
        qs[it] = ld_row<VEC>(qs_row + col[it]);
        s[it] = ld_plain<VEC>(s_row + col[it]);  // may alias out
        if (has_p) pp[it] = ld_stream<VEC>(p_row + col[it]);
      }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      if (!ok[it]) continue;
      Row<VEC> o, m;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float msg = __fsub_rn(qd[it].v[k], qs[it].v[k]);
        float t;
        if (!has_p) {
          t = __fadd_rn(s[it].v[k], msg);
        } else {
          const float add = __fmul_rn(c, pp[it].v[k]);
          if (p.order == DMP_ORDER_SCM) t = __fadd_rn(__fadd_rn(s[it].v[k], add), msg);
          else t = __fadd_rn(__fadd_rn(s[it].v[k], msg), add);
        }
        if (p.ebias != nullptr) t = __fadd_rn(t, bias[it].v[k]);
        o.v[k] = t;
        m.v[k] = msg;
      }
      st_stream<VEC>(p.out + e * p.ld_out + col[it], o);
      if (p.agg != nullptr) st_stream<VEC>(p.agg + e * p.ld_agg + col[it], m);
    }
  }
}

// Mirrored halves: after `add_reversed_edges` on one graph (SCM/train.py:299-327) edge e + E/2 is the reverse of edge
// e, and its endpoint ROLES coincide (a = dst of the original, b = its src), so both rows gather the same Q_d / Q_s
// rows.  One group handles the pair and fetches the two table rows once: 4 instead of 5 row-reads+writes per edge.  The
// indices of both rows are compared, a pair that does not match simply loads its own rows -> always correct.
template <int VEC, int G, int ITER>
__global__ void __launch_bounds__(kThreads, 4) edge_update_pair_kernel(const EdgeFwdParams p) {
  constexpr int kGroups = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * kGroups;
  const int64_t half = p.E / 2;
  const bool has_p = p.P != nullptr;
  int col[ITER];
  bool ok[ITER];
  Row<VEC> bias[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    col[it] = (lane + it * G) * VEC;
    ok[it] = col[it] < p.H;
    if (ok[it] && p.ebias != nullptr) bias[it] = ld_row<VEC>(p.ebias + col[it]);
  }
#pragma unroll 1
  for (int64_t e = (int64_t)blockIdx.x * kGroups + threadIdx.x / G; e < half; e += stride) {
    const int64_t e2 = e + half;
    const int a = __ldg(p.a32 + e), b = __ldg(p.b32 + e);
    const int a2 = __ldg(p.a32 + e2), b2 = __ldg(p.b32 + e2);
    const float c = __ldg(p.coef + e), c2 = __ldg(p.coef + e2);
    const float* qd_row = p.Qd + (int64_t)a * p.ldQd;
    const float* qs_row = p.Qs + (int64_t)b * p.ldQs;
    Row<VEC> s[ITER], pp[ITER], s2[ITER], pp2[ITER], qd[ITER], qs[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      if (ok[it]) {
        qd[it] = ld_row<VEC>(qd_row + col[it]);
        qs[it] = ld_row<VEC>(qs_row + col[it]);
        s[it] = ld_plain<VEC>(p.S + e * p.ldS + col[it]);     // may alias out
        s2[it] = ld_plain<VEC>(p.S + e2 * p.ldS + col[it]);
        if (has_p) {
          pp[it] = ld_stream<VEC>(p.P + e * p.ldP + col[it]);
          pp2[it] = ld_stream<VEC>(p.P + e2 * p.ldP + col[it]);
        }
      }
    }
    auto emit = [&](int64_t row, float cc, const Row<VEC> (&sv)[ITER], const Row<VEC> (&pv)[ITER]) {
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        if (!ok[it]) continue;
        Row<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float msg = __fsub_rn(qd[it].v[k], qs[it].v[k]);
          float t;
          if (!has_p) {
            t = __fadd_rn(sv[it].v[k], msg);
          } else {
            const float add = __fmul_rn(cc, pv[it].v[k]);
            if (p.order == DMP_ORDER_SCM) t = __fadd_rn(__fadd_rn(sv[it].v[k], add), msg);
            else t = __fadd_rn(__fadd_rn(sv[it].v[k], msg), add);
          }
          if (p.ebias != nullptr) t = __fadd_rn(t, bias[it].v[k]);
          o.v[k] = t;
        }
        st_stream<VEC>(p.out + row * p.ld_out + col[it], o);
      }
    };
    emit(e, c, s, pp);
    if (a2 != a || b2 != b) {     // not a mirrored pair (group-uniform): fetch the second row's own table rows
#pragma unroll
      for (int it = 0; it < ITER; ++it) {
        if (ok[it]) {
          qd[it] = ld_row<VEC>(p.Qd + (int64_t)a2 * p.ldQd + col[it]);
          qs[it] = ld_row<VEC>(p.Qs + (int64_t)b2 * p.ldQs + col[it]);
        }
      }
    }
    emit(e2, c2, s2, pp2);
  }
}

// ---- backward edge gather ----------------------------------------------------------------------------
struct EdgeBwdParams {
  const int32_t* dst32;
  const uint8_t* rev;
  const float* norm;
  const float* coef;
  const float* gN; const float* gN_rev; int64_t ld_gN;
  const float* gE; int64_t ld_gE;
  float* T; int64_t ldT; int64_t T_rev_off;
  float* CG; int64_t ldCG;
  int64_t E;
  int H;
};

template <int VEC, int G, int ITER>
__global__ void __launch_bounds__(kThreads) edge_backward_kernel(const EdgeBwdParams p) {
  constexpr int kGroups = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * kGroups;
  int col[ITER];
  bool ok[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    col[it] = (lane + it * G) * VEC;
    ok[it] = col[it] < p.H;
  }
  const bool do_t = p.T != nullptr;
  const bool do_cg = p.CG != nullptr;
  for (int64_t e = (int64_t)blockIdx.x * kGroups + threadIdx.x / G; e < p.E; e += stride) {
    Row<VEC> g[ITER], ge[ITER];
    float w = 1.0f, c = 0.0f;
    bool neg = true;
    int64_t t_off = 0;
    if (do_t) {
      const int d = __ldg(p.dst32 + e);
      if (p.rev != nullptr) neg = __ldg(p.rev + e) == 0;
      t_off = neg ? 0 : p.T_rev_off;
      if (p.norm != nullptr) w = __ldg(p.norm + e);
      const float* row = (neg ? p.gN : p.gN_rev) + (int64_t)d * p.ld_gN;
#pragma unroll
      for (int it = 0; it < ITER; ++it)
        if (ok[it]) g[it] = ld_row<VEC>(row + col[it]);
    }
    if (do_cg) {
      c = __ldg(p.coef + e);
      const float* row = p.gE + e * p.ld_gE;
#pragma unroll
      for (int it = 0; it < ITER; ++it)
        if (ok[it]) ge[it] = ld_stream<VEC>(row + col[it]);
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      if (!ok[it]) continue;
      if (do_t) {
        Row<VEC> t;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float x = g[it].v[k];
          if (p.norm != nullptr) x = __fmul_rn(x, w);
          t.v[k] = neg ? -x : x;
        }
        st_stream<VEC>(p.T + e * p.ldT + t_off + col[it], t);
      }
      if (do_cg) {
        Row<VEC> t;
#pragma unroll
        for (int k = 0; k < VEC; ++k) t.v[k] = __fmul_rn(c, ge[it].v[k]);
        st_stream<VEC>(p.CG + e * p.ldCG + col[it], t);
      }
    }
  }
}

// ---- act / gate / residual epilogue ------------------------------------------------------------------
struct GateParams {
  const float* x; int64_t ldx;
  const float* gate;
  const float* prev; int64_t ld_prev;
  float* out; int64_t ld_out;
  int64_t rows;
  int H;
  int act;
  float slope;
};

template <int VEC, int G, int ITER, bool BWD>
__global__ void __launch_bounds__(kThreads) gate_residual_kernel(const GateParams p) {
  // BWD: x=gout? no -- see launcher: for BWD, `prev` carries the forward pre-activation x and `x` carries gout.
  constexpr int kGroups = kThreads / G;
  const int lane = threadIdx.x % G;
  const int64_t stride = (int64_t)gridDim.x * kGroups;
  for (int64_t r = (int64_t)blockIdx.x * kGroups + threadIdx.x / G; r < p.rows; r += stride) {
    const float gate = p.gate != nullptr ? __ldg(p.gate + r) : 1.0f;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int col = (lane + it * G) * VEC;
      if (col >= p.H) continue;
      Row<VEC> x = ld_plain<VEC>(p.x + r * p.ldx + col);
      Row<VEC> o;
      if constexpr (!BWD) {
        Row<VEC> pv;
        if (p.prev != nullptr) pv = ld_plain<VEC>(p.prev + r * p.ld_prev + col);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float y = apply_act(x.v[k], p.act, p.slope);
          if (p.gate != nullptr) y = __fmul_rn(y, gate);
          if (p.prev != nullptr) y = __fadd_rn(pv.v[k], y);
          o.v[k] = y;
        }
      } else {
        Row<VEC> pre;
        if (p.act != DMP_ACT_NONE) pre = ld_plain<VEC>(p.prev + r * p.ld_prev + col);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          float y = x.v[k];
          if (p.gate != nullptr) y = __fmul_rn(y, gate);
          if (p.act != DMP_ACT_NONE)
            y = __fmul_rn(y, (p.act & DMP_ACT_FROM_OUTPUT)
                                 ? act_grad_from_output(pre.v[k], p.act & ~DMP_ACT_FROM_OUTPUT, p.slope)
                                 : act_grad(pre.v[k], p.act, p.slope));
          o.v[k] = y;
        }
      }
      st_row<VEC>(p.out + r * p.ld_out + col, o);
    }
  }
}

__global__ void __launch_bounds__(kThreads) permute_scalar_kernel(const uint32_t* __restrict__ eid,
                                                                  const float* __restrict__ values,
                                                                  float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    out[j] = __ldg(values + (__ldg(eid + j) & DMP_EID_MASK));
}

// ---- dispatch helpers -----------------------------------------------------------------------------------
#define DMP_DISPATCH_SHAPE(KERNEL, PARAMS, ROWS, STREAM, ...)                                         \
  do {                                                                                                \
    const Shape s_ = pick_shape((PARAMS).H, vec);                                                     \
    const int ctas_per_sm_ = 2048 / kThreads;                                                         \
    const unsigned grid_ = persistent_grid((ROWS), kThreads / s_.g, ctas_per_sm_);                    \
    if (vec == 4) {                                                                                   \
      if (s_.g == 8) KERNEL<4, 8, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);             \
      else if (s_.g == 16) KERNEL<4, 16, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);      \
      else if (s_.iter == 1) KERNEL<4, 32, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else if (s_.iter == 2) KERNEL<4, 32, 2 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else KERNEL<4, 32, 4 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);                      \
    } else if (vec == 2) {                                                                            \
      if (s_.g == 8) KERNEL<2, 8, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);             \
      else if (s_.g == 16) KERNEL<2, 16, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);      \
      else if (s_.iter == 1) KERNEL<2, 32, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else if (s_.iter == 2) KERNEL<2, 32, 2 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else KERNEL<2, 32, 4 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);                      \
    } else {                                                                                          \
      if (s_.g == 8) KERNEL<1, 8, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);             \
      else if (s_.g == 16) KERNEL<1, 16, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);      \
      else if (s_.iter == 1) KERNEL<1, 32, 1 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else if (s_.iter == 2) KERNEL<1, 32, 2 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);    \
      else KERNEL<1, 32, 4 __VA_ARGS__><<<grid_, kThreads, 0, STREAM>>>(PARAMS);                      \
    }                                                                                                 \
  } while (0)

#define DMP_COMMA_TRUE , true
#define DMP_COMMA_FALSE , false

}  // namespace dmp

extern "C" int dmp_edge_update(const int32_t* a32, const int32_t* b32, const float* coef,
                               const float* S, int64_t ldS, const float* P, int64_t ldP,
                               const float* Qd, int64_t ldQd, const float* Qs, int64_t ldQs,
                               const float* ebias, float* out, int64_t ld_out, float* edge_agg,
                               int64_t ld_agg, int64_t num_edges, int64_t H, int order, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_edges >= 0 && H >= 0, "edge_update: negative size");
  if (num_edges == 0 || H == 0) return DMP_OK;
  DMP_CHECK_ARG(a32 && b32 && coef && S && Qd && Qs && out, "edge_update: null pointer");   // P may be NULL
  const bool pair = (order & DMP_EDGE_MIRRORED_HALVES) != 0 && num_edges % 2 == 0;
  order &= ~DMP_EDGE_MIRRORED_HALVES;
  DMP_CHECK_ARG(order == DMP_ORDER_SCM || order == DMP_ORDER_UNC, "edge_update: bad order %d", order);
  DMP_CHECK_ARG(ldS >= H && (!P || ldP >= H) && ldQd >= H && ldQs >= H && ld_out >= H && (!edge_agg || ld_agg >= H),
                "edge_update: leading dimension smaller than H");
  DMP_CHECK_ARG((P == nullptr || P != out) && Qd != out && Qs != out, "edge_update: only S may alias out");
  const int vec = pick_vec(H, {ldS, P ? ldP : 0, ldQd, ldQs, ld_out, edge_agg ? ld_agg : 0},
                           {S, P, Qd, Qs, ebias, out, edge_agg});
  const int64_t chunk = max_chunk(vec);
  for (int64_t c0 = 0; c0 < H; c0 += chunk) {
    EdgeFwdParams p;
    p.a32 = a32; p.b32 = b32; p.coef = coef;
    p.S = S + c0; p.ldS = ldS; p.P = P ? P + c0 : nullptr; p.ldP = ldP;
    p.Qd = Qd + c0; p.ldQd = ldQd; p.Qs = Qs + c0; p.ldQs = ldQs;
    p.ebias = ebias ? ebias + c0 : nullptr;
    p.out = out + c0; p.ld_out = ld_out;
    p.agg = edge_agg ? edge_agg + c0 : nullptr; p.ld_agg = ld_agg;
    p.E = num_edges; p.H = (int)((H - c0 < chunk) ? (H - c0) : chunk); p.order = order;
    const Shape sh = pick_shape(p.H, vec);
    if (pair && sh.iter == 1 && edge_agg == nullptr) {   // one column vector per lane (H <= 128 at VEC 4), no edge_agg
                                                         // output: the pair fits 64 registers
      const unsigned grid = persistent_grid(num_edges / 2, kThreads / sh.g, 4);
      cudaStream_t st = (cudaStream_t)stream;
#define DMP_PAIR_G(V)                                                                       \
      do {                                                                                  \
        if (sh.g == 8) edge_update_pair_kernel<V, 8, 1><<<grid, kThreads, 0, st>>>(p);        \
        else if (sh.g == 16) edge_update_pair_kernel<V, 16, 1><<<grid, kThreads, 0, st>>>(p); \
        else edge_update_pair_kernel<V, 32, 1><<<grid, kThreads, 0, st>>>(p);                 \
      } while (0)
      if (vec == 4) DMP_PAIR_G(4);
      else if (vec == 2) DMP_PAIR_G(2);
      else DMP_PAIR_G(1);
#undef DMP_PAIR_G
    } else {
      DMP_DISPATCH_SHAPE(edge_update_kernel, p, num_edges, (cudaStream_t)stream);
    }
    int rc = launch_status("edge_update_kernel");
    if (rc != DMP_OK) return rc;
  }
  return DMP_OK;
}

extern "C" int dmp_edge_backward(const int32_t* dst32, const uint8_t* rev, const float* norm,
                                 const float* coef, const float* gN, const float* gN_rev, int64_t ld_gN, const float* gE,
                                 int64_t ld_gE, float* T, int64_t ldT, int64_t T_rev_col_offset, float* CG,
                                 int64_t ldCG, int64_t num_edges, int64_t H, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_edges >= 0 && H >= 0, "edge_backward: negative size");
  if (num_edges == 0 || H == 0 || (T == nullptr && CG == nullptr)) return DMP_OK;
  DMP_CHECK_ARG(T == nullptr || (dst32 && gN && ld_gN >= H && ldT >= H && T_rev_col_offset >= 0 &&
                                 ldT >= T_rev_col_offset + H),
                "edge_backward: bad T operands");
  DMP_CHECK_ARG(CG == nullptr || (coef && gE && ld_gE >= H && ldCG >= H), "edge_backward: bad CG operands");
  DMP_CHECK_ARG(CG == nullptr || CG != gE, "edge_backward: CG must not alias gE");
  const int vec = pick_vec(H, {T ? ld_gN : 0, T ? ldT : 0, T ? T_rev_col_offset : 0, CG ? ld_gE : 0, CG ? ldCG : 0},
                           {T ? gN : nullptr, T, CG ? gE : nullptr, CG});
  const int64_t chunk = max_chunk(vec);
  for (int64_t c0 = 0; c0 < H; c0 += chunk) {
    EdgeBwdParams p;
    p.dst32 = dst32; p.rev = rev; p.norm = norm; p.coef = coef;
    p.gN = gN ? gN + c0 : nullptr; p.gN_rev = gN_rev ? gN_rev + c0 : p.gN; p.ld_gN = ld_gN;
    p.gE = gE ? gE + c0 : nullptr; p.ld_gE = ld_gE;
    p.T = T ? T + c0 : nullptr; p.ldT = ldT; p.T_rev_off = T_rev_col_offset;
    p.CG = CG ? CG + c0 : nullptr; p.ldCG = ldCG;
    p.E = num_edges; p.H = (int)((H - c0 < chunk) ? (H - c0) : chunk);
    DMP_DISPATCH_SHAPE(edge_backward_kernel, p, num_edges, (cudaStream_t)stream);
    int rc = launch_status("edge_backward_kernel");
    if (rc != DMP_OK) return rc;
  }
  return DMP_OK;
}

static int gate_common(bool bwd, const float* x, int64_t ldx, const float* gate, const float* prev,
                       int64_t ld_prev, float* out, int64_t ld_out, int64_t rows, int64_t H, int act,
                       float slope, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(rows >= 0 && H >= 0, "gate_residual: negative size");
  if (rows == 0 || H == 0) return DMP_OK;
  DMP_CHECK_ARG(x && out && ldx >= H && ld_out >= H, "gate_residual: bad operands");
  DMP_CHECK_ARG((act & ~DMP_ACT_FROM_OUTPUT) >= DMP_ACT_NONE && (act & ~DMP_ACT_FROM_OUTPUT) <= DMP_ACT_SIGMOID &&
                    (bwd || !(act & DMP_ACT_FROM_OUTPUT)),
                "gate_residual: bad activation %d", act);
  DMP_CHECK_ARG(prev == nullptr || ld_prev >= H, "gate_residual: bad prev leading dimension");
  const int vec = pick_vec(H, {ldx, ld_out, prev ? ld_prev : 0}, {x, prev, out});
  const int64_t chunk = max_chunk(vec);
  for (int64_t c0 = 0; c0 < H; c0 += chunk) {
    GateParams p;
    p.x = x + c0; p.ldx = ldx; p.gate = gate;
    p.prev = prev ? prev + c0 : nullptr; p.ld_prev = ld_prev;
    p.out = out + c0; p.ld_out = ld_out; p.rows = rows;
    p.H = (int)((H - c0 < chunk) ? (H - c0) : chunk); p.act = act; p.slope = slope;
    if (bwd) DMP_DISPATCH_SHAPE(gate_residual_kernel, p, rows, (cudaStream_t)stream, DMP_COMMA_TRUE);
    else DMP_DISPATCH_SHAPE(gate_residual_kernel, p, rows, (cudaStream_t)stream, DMP_COMMA_FALSE);
    int rc = launch_status("gate_residual_kernel");
    if (rc != DMP_OK) return rc;
  }
  return DMP_OK;
}

extern "C" int dmp_gate_residual(const float* x, int64_t ldx, const float* gate, const float* prev,
                                 int64_t ld_prev, float* out, int64_t ld_out, int64_t rows, int64_t H,
                                 int act, float slope, void* stream) {
  return gate_common(false, x, ldx, gate, prev, ld_prev, out, ld_out, rows, H, act, slope, stream);
}

extern "C" int dmp_gate_residual_backward(const float* gout, int64_t ld_gout, const float* x, int64_t ldx,
                                          const float* gate, float* gx, int64_t ld_gx, int64_t rows,
                                          int64_t H, int act, float slope, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG((act & ~DMP_ACT_FROM_OUTPUT) == DMP_ACT_NONE || x != nullptr,
                "gate_residual_backward: x required for act'");
  // kernel reads `x` slot = gout, `prev` slot = forward pre-activation
  if ((act & ~DMP_ACT_FROM_OUTPUT) == DMP_ACT_NONE) act = DMP_ACT_NONE;
  return gate_common(true, gout, ld_gout, gate, act == DMP_ACT_NONE ? nullptr : x, ldx, gx, ld_gx, rows, H,
                     act, slope, stream);
}

extern "C" int dmp_permute_edge_scalar(const int32_t* eid, const float* values, float* out,
                                       int64_t num_edges, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_edges >= 0, "permute_edge_scalar: negative size");
  if (num_edges == 0) return DMP_OK;
  DMP_CHECK_ARG(eid && values && out, "permute_edge_scalar: null pointer");
  int64_t need = (num_edges + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)kNumSMs * 8;
  permute_scalar_kernel<<<(unsigned)(need < cap ? need : cap), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(eid), values, out, num_edges);
  return launch_status("permute_scalar_kernel");
}
