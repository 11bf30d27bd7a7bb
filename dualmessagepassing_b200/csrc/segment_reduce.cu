// segment_reduce.cu -- deterministic sorted-segment sum of edge rows (forward node aggregation and
// the backward scatter of the endpoint gathers).  Replaces DGL's `fn.sum` gSpMM (dmpnn.py:92,163)
// and autograd's atomic `index_add_`; see include/dmp_b200.h for the contract.
//
// Work decomposition: one group of G lanes per segment (G = 8/16/32 so that a 128-wide fp32 row is
// exactly one float4 per lane of a full warp; narrower rows pack several segments per warp).  A lane
// owns ITER column vectors and walks the segment's edge list in ascending position, keeping U rows
// in flight (U independent 128-bit loads per lane) and adding them in order with __fadd_rn, so the
// result equals a sequential CPU loop bit for bit.  The edge-id list is read G entries at a time with one
// coalesced load per group and broadcast by shuffle; bit 31 of each entry is the reversed flag, so neither
// the flag nor the sign needs a second gather.
#include "common.cuh"

namespace dmp {

struct SegParams {
  const int32_t* indptr;
  const uint32_t* eid;
  const float* w_perm;
  const float* V;
  int64_t ldV;
  int64_t rev_off;
  const float* base;
  int64_t ld_base;
  const float* bias;
  float* out;
  int64_t ld_out;
  int64_t nseg;
  int H;
  int mode;
  int64_t split_off;   // SPLIT: column offset of the reversed-edge sums in `out` (= full H, also when H is chunked)
};

// KIND: 0 plain, 1 FILTER (keep only forward / only reversed edges), 2 SPLIT (two sums per segment: forward edges into
// out[:, 0:H], reversed edges into out[:, H:2H] -- the aggregate-first form of the node update reads every row once)
// MINB: resident CTAs per SM the register allocation must allow.  Long segments want U = 8 rows in flight per lane
// (80 registers, 3 CTAs); SHORT segments (a destination-range partition leaves ~2.5 local edges per global a-/b-segment)
// are bound by the indptr -> eid -> row latency chain of one segment per warp, so the DMP_SEG_SHORT variant keeps U = 2
// rows in flight and fits twice the warps (same adds in the same order: identical bits).
template <int VEC, int G, int ITER, int U, int KIND, int MINB = 3>
__global__ void __launch_bounds__(kThreads, MINB) segment_reduce_kernel(const SegParams p) {
  constexpr int kGroups = kThreads / G;
  constexpr bool FILTER = (KIND == 1), SPLIT = (KIND == 2);
  const int lane = threadIdx.x % G;
  const int64_t seg = (int64_t)blockIdx.x * kGroups + threadIdx.x / G;
  if (seg >= p.nseg) return;

  const int beg = __ldg(p.indptr + seg);
  const int end = __ldg(p.indptr + seg + 1);
  const bool sign_by_rev = (p.mode & DMP_SEG_SIGN_BY_REV) != 0;
  const bool has_w = p.w_perm != nullptr;
  // edge filter by the reversed flag (bit 31): keep = flag XOR-matches; filtered rows are never loaded
  const uint32_t filt = (p.mode & (DMP_SEG_ONLY_FWD | DMP_SEG_ONLY_REV));

  int col[ITER];
  bool ok[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    col[it] = (lane + it * G) * VEC;
    ok[it] = col[it] < p.H;
  }

  float acc[ITER][VEC], acc_rev[SPLIT ? ITER : 1][VEC];
#pragma unroll
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      acc[it][k] = 0.0f;
      if (SPLIT) acc_rev[it][k] = 0.0f;
    }

  // The segment's (edge id | flag) words and weights are fetched G at a time with ONE coalesced load per group and
  // handed out by shuffle: the row addresses of a whole chunk are known up front, so no row load ever waits behind a
  // dependent index load (was: one broadcast index load per U rows in the critical path of every iteration).
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x % 32) / G * G));
  for (int j0 = beg; j0 < end; j0 += G) {
    const int cnt = (end - j0 < G) ? (end - j0) : G;
    const uint32_t my_ef = lane < cnt ? __ldg(p.eid + j0 + lane) : 0u;
    const float my_w = (has_w && lane < cnt) ? __ldg(p.w_perm + j0 + lane) : 1.0f;
    for (int j = 0; j < cnt; j += U) {
      uint32_t ef[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ef[u] = __shfl_sync(gmask, my_ef, (j + u) & (G - 1), G);
        w[u] = __shfl_sync(gmask, my_w, (j + u) & (G - 1), G);
      }
      Row<VEC> v[U][ITER];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bool keep = j + u < cnt;
        if constexpr (FILTER) {
          keep = keep && !((filt & DMP_SEG_ONLY_FWD) && (ef[u] >> 31)) && !((filt & DMP_SEG_ONLY_REV) && !(ef[u] >> 31));
          if (!keep) ef[u] = 0xffffffffu;   // sentinel: (id mask, rev) can never both be all-ones for a real edge
        }
        if (keep) {
          const uint32_t r = ef[u] >> 31;
          const float* row = p.V + (int64_t)(ef[u] & DMP_EID_MASK) * p.ldV + (r ? p.rev_off : 0);
#pragma unroll
          for (int it = 0; it < ITER; ++it)
            if (ok[it]) v[u][it] = ld_stream<VEC>(row + col[it]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < cnt && (!FILTER || ef[u] != 0xffffffffu)) {
          const bool neg = sign_by_rev && (ef[u] >> 31) == 0;
#pragma unroll
          for (int it = 0; it < ITER; ++it) {
            if (ok[it]) {
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                float t = v[u][it].v[k];
                if (neg) t = -t;
                if (has_w) t = __fmul_rn(t, w[u]);
                if (SPLIT && (ef[u] >> 31)) acc_rev[it][k] = __fadd_rn(acc_rev[it][k], t);
                else acc[it][k] = __fadd_rn(acc[it][k], t);
              }
            }
          }
        }
      }
    }
  }

  const bool negate = (p.mode & DMP_SEG_NEGATE_OUT) != 0;
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    if (!ok[it]) continue;
    Row<VEC> r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = negate ? -acc[it][k] : acc[it][k];
    if (p.base != nullptr) {
      Row<VEC> b = ld_plain<VEC>(p.base + seg * p.ld_base + col[it]);  // base may alias out
#pragma unroll
      for (int k = 0; k < VEC; ++k) r.v[k] = __fadd_rn(b.v[k], r.v[k]);
    }
    if (p.bias != nullptr) {
      Row<VEC> b = ld_row<VEC>(p.bias + col[it]);
#pragma unroll
      for (int k = 0; k < VEC; ++k) r.v[k] = __fadd_rn(r.v[k], b.v[k]);
    }
    st_row<VEC>(p.out + seg * p.ld_out + col[it], r);
    if constexpr (SPLIT) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) r.v[k] = negate ? -acc_rev[it][k] : acc_rev[it][k];
      st_row<VEC>(p.out + seg * p.ld_out + p.split_off + col[it], r);
    }
  }
}

template <int VEC, int G, int ITER, int U>
static int launch(const SegParams& p, cudaStream_t stream) {
  const bool filter = (p.mode & (DMP_SEG_ONLY_FWD | DMP_SEG_ONLY_REV)) != 0;
  constexpr int kGroups = kThreads / G;
  const int64_t blocks = (p.nseg + kGroups - 1) / kGroups;
  if (blocks > 0x7fffffffLL) {
    set_error("segment_reduce: too many segments (%lld)", (long long)p.nseg);
    return DMP_ERR_UNSUPPORTED;
  }
  if constexpr (VEC == 4 && G == 32 && ITER == 1) {
    if ((p.mode & DMP_SEG_SHORT) && !(p.mode & DMP_SEG_SPLIT_BY_REV) && !filter) {
      segment_reduce_kernel<VEC, G, ITER, 2, 0, 6><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
      return launch_status("segment_reduce_kernel<short>");
    }
  }
  if (p.mode & DMP_SEG_SPLIT_BY_REV) segment_reduce_kernel<VEC, G, ITER, U, 2><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  else if (filter) segment_reduce_kernel<VEC, G, ITER, U, 1><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  else segment_reduce_kernel<VEC, G, ITER, U, 0><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  return launch_status("segment_reduce_kernel");
}

template <int VEC>
static int dispatch(const SegParams& p, const Shape& s, cudaStream_t stream) {
  if (s.g == 8) return launch<VEC, 8, 1, 8>(p, stream);
  if (s.g == 16) return launch<VEC, 16, 1, 8>(p, stream);
  if (s.iter == 1) return launch<VEC, 32, 1, 8>(p, stream);
  if (s.iter == 2) return launch<VEC, 32, 2, 4>(p, stream);
  return launch<VEC, 32, 4, 2>(p, stream);
}

}  // namespace dmp

extern "C" int dmp_segment_reduce(const int32_t* indptr, const int32_t* eid, const float* w_perm,
                                  const float* V, int64_t ldV, int64_t rev_col_offset,
                                  const float* base, int64_t ld_base, const float* bias,
                                  float* out, int64_t ld_out, int64_t num_segments, int64_t H, int mode,
                                  void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_segments >= 0 && H >= 0, "segment_reduce: negative size");
  if (num_segments == 0 || H == 0) return DMP_OK;
  // eid / V may be NULL only for an edgeless graph (every segment empty: they are never dereferenced)
  DMP_CHECK_ARG(indptr && out, "segment_reduce: null pointer");
  DMP_CHECK_ARG(ldV >= H && ld_out >= H && (base == nullptr || ld_base >= H),
                "segment_reduce: leading dimension smaller than H");
  const bool split = (mode & DMP_SEG_SPLIT_BY_REV) != 0;
  DMP_CHECK_ARG(!split || (ld_out >= 2 * H && base == nullptr && bias == nullptr &&
                           !(mode & (DMP_SEG_ONLY_FWD | DMP_SEG_ONLY_REV))),
                "segment_reduce: SPLIT_BY_REV writes [fwd | rev] rows of 2H floats and takes no base / bias / filter");
  const int vec = pick_vec(H, {ldV, ld_out, base ? ld_base : 0, rev_col_offset, split ? H : 0}, {V, out, base, bias});
  const int64_t chunk = max_chunk(vec);
  for (int64_t c0 = 0; c0 < H; c0 += chunk) {
    const int64_t Hc = (H - c0 < chunk) ? (H - c0) : chunk;
    SegParams p;
    p.indptr = indptr;
    p.eid = reinterpret_cast<const uint32_t*>(eid);
    p.w_perm = w_perm;
    p.V = V + c0;
    p.ldV = ldV;
    p.rev_off = rev_col_offset;
    p.base = base ? base + c0 : nullptr;
    p.ld_base = ld_base;
    p.bias = bias ? bias + c0 : nullptr;
    p.out = out + c0;
    p.ld_out = ld_out;
    p.nseg = num_segments;
    p.H = (int)Hc;
    p.mode = mode;
    p.split_off = H;
    const Shape s = pick_shape(Hc, vec);
    int rc;
    if (vec == 4) rc = dispatch<4>(p, s, (cudaStream_t)stream);
    else if (vec == 2) rc = dispatch<2>(p, s, (cudaStream_t)stream);
    else rc = dispatch<1>(p, s, (cudaStream_t)stream);
    if (rc != DMP_OK) return rc;
  }
  return DMP_OK;
}
