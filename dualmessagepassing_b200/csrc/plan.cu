// plan.cu -- graph plan construction on device (row A0 of SURVEY.md section 8):
// COO (int64, edge-id order) -> dst/a/b endpoint roles, out-degrees, the degree coefficient of
// dmpnn.py:144-146 and three stable edge-id segmentations (by dst = CSC, by a, by b).
//
// Replaces DGL's lazy COO->CSC conversion behind fn.sum (dmpnn.py:92,163), graph.out_degrees()
// (dmpnn.py:100-101) and the index side of the endpoint gathers.  Integer work only; results are
// bit-exact against numpy's stable argsort / bincount (oracle/graph_oracle.py).
//
// Steps (all on the caller's stream, no host sync):
//   1. roles_kernel     : narrow to int32, pick a/b by the reversed flag, histogram dst/a/b/src with
//                         integer atomics (order-independent), flag out-of-range endpoints.
//   2. exclusive scan   : three histograms -> three indptr arrays (scan_* kernels below: block scan, scan of the
//                         block sums, add back).
//   3. stable sort      : (key, edge id | rev<<31) pairs by key with a hand-written LSD radix sort restricted to
//                         ceil(log2 N) bits, 8 bits per pass: per-tile digit histogram -> scan over (digit, tile) ->
//                         scatter with a STABLE in-tile rank (__match_any_sync peer masks, warps own contiguous
//                         chunks of the tile) -- stable, so edge ids ascend inside a segment.
//   4. coef_kernel      : coef[e] = lut[deg] for deg < lut_len else 2*(1+log2f(1+deg)).
// No library primitives: everything in the plan build is written here.
#include "common.cuh"

namespace dmp {

__global__ void __launch_bounds__(kThreads) roles_kernel(
    const int64_t* __restrict__ src, const int64_t* __restrict__ dst, const uint8_t* __restrict__ rev,
    int64_t N, int64_t E, int32_t* __restrict__ dst32, int32_t* __restrict__ a32, int32_t* __restrict__ b32,
    uint32_t* __restrict__ eid_flag, int32_t* __restrict__ cnt_dst, int32_t* __restrict__ cnt_a,
    int32_t* __restrict__ cnt_b, unsigned long long* __restrict__ cnt_src, int32_t* __restrict__ status) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t s = src[e], d = dst[e];
    const uint32_t r = (rev != nullptr && rev[e] != 0) ? 1u : 0u;
    eid_flag[e] = (uint32_t)e | (r << 31);
    if (s < 0 || s >= N || d < 0 || d >= N) {
      atomicExch(status, 1);
      dst32[e] = 0; a32[e] = 0; b32[e] = 0;
      continue;
    }
    const int32_t a = r ? (int32_t)s : (int32_t)d;
    const int32_t b = r ? (int32_t)d : (int32_t)s;
    dst32[e] = (int32_t)d;
    a32[e] = a;
    b32[e] = b;
    atomicAdd(cnt_dst + d, 1);
    atomicAdd(cnt_a + a, 1);
    atomicAdd(cnt_b + b, 1);
    if (cnt_src != nullptr) atomicAdd(cnt_src + s, 1ULL);
  }
}

__global__ void __launch_bounds__(kThreads) coef_kernel(const int32_t* __restrict__ dst32,
                                                        const int64_t* __restrict__ deg,
                                                        const float* __restrict__ lut, int64_t lut_len,
                                                        float* __restrict__ coef, int64_t E) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t d = deg[dst32[e]];
    float c;
    if (lut != nullptr && d >= 0 && d < lut_len) {
      c = __ldg(lut + d);
    } else {
      // 2 * (1 + log2(1 + float(d))), same association as dmpnn.py:144-146
      const float t = log2f(__fadd_rn(1.0f, (float)d));
      c = __fmul_rn(2.0f, __fadd_rn(1.0f, t));
    }
    coef[e] = c;
  }
}

// ---- exclusive scan of int32 (three phases, deterministic) ------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                         // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 per block

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* smem /*[8]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 8 ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < 8) smem[lane] = w;  // inclusive warp totals
  }
  __syncthreads();
  const int warp_off = warp > 0 ? smem[warp - 1] : 0;
  *total = smem[7];
  return warp_off + x - v;  // exclusive
}

// phase 1: per-block exclusive scan, block totals to `sums`
__global__ void __launch_bounds__(kScanThreads) scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                                  int32_t* __restrict__ sums, int64_t n) {
  __shared__ int sm[8];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems], t = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    t += v[i];
  }
  int total;
  int off = block_exclusive_scan_256(t, sm, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// phase 3: add the scanned block sums back
__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ sums,
                                                                int64_t n) {
  const int add = sums[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += add;
}

// exclusive scan of n ints; `tmp` needs scan_tmp_ints(n) ints.  Recursion depth <= 3 for n < 2^33.
static int64_t scan_tmp_ints(int64_t n) {
  int64_t total = 0;
  while (n > 1) {
    n = (n + kScanTile - 1) / kScanTile;
    total += n;
    if (n == 1) break;
  }
  return total + 1;
}
static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tmp, cudaStream_t stream) {
  if (n <= 0) return DMP_OK;
  const int64_t blocks = (n + kScanTile - 1) / kScanTile;
  scan_block_kernel<<<(unsigned)blocks, kScanThreads, 0, stream>>>(in, out, tmp, n);
  int rc = launch_status("scan_block_kernel");
  if (rc != DMP_OK || blocks == 1) return rc;
  rc = exclusive_scan(tmp, tmp, blocks, tmp + blocks, stream);   // in place on the block sums
  if (rc != DMP_OK) return rc;
  scan_add_kernel<<<(unsigned)blocks, kScanThreads, 0, stream>>>(out, tmp, n);
  return launch_status("scan_add_kernel");
}

// ---- stable LSD radix sort of (key, value) pairs, 8 bits per pass -----------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 8;                             // items per thread
constexpr int kSortTile = kSortThreads * kSortRounds;      // 2048 keys per block
// element order inside a tile: warp w owns the contiguous chunk [w*256, (w+1)*256), round k covers
// indices w*256 + k*32 + lane -> (warp, round, lane) order == index order, which is what stability needs.
__device__ __forceinline__ int64_t sort_index(int64_t tile, int warp, int round, int lane) {
  return tile * kSortTile + warp * (32 * kSortRounds) + round * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift,
                                                                  int32_t* __restrict__ hist /*[256][tiles]*/,
                                                                  int64_t tiles) {
  __shared__ int cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kSortRounds; ++k) {
    const int64_t i = sort_index(blockIdx.x, warp, k, lane);
    if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 255], 1);   // integer atomics: order-independent result
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(
    const int32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
    const int32_t* __restrict__ offs /*[256][tiles], exclusive scan of hist*/, int64_t tiles,
    int32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ int cnt[8][256];     // running per-warp digit counts
  __shared__ int wbase[8][256];   // digits of earlier warps of this tile
  for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  int32_t key[kSortRounds];
  uint32_t val[kSortRounds];
  int pre[kSortRounds];
#pragma unroll
  for (int k = 0; k < kSortRounds; ++k) {
    const int64_t i = sort_index(blockIdx.x, warp, k, lane);
    const bool ok = i < n;
    key[k] = ok ? keys[i] : 0;
    val[k] = ok ? vals[i] : 0u;
    const int d = (key[k] >> shift) & 255;
    const uint32_t active = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t peers = __match_any_sync(active, d);
      const int before = cnt[warp][d];              // count from this warp's earlier rounds
      pre[k] = before + __popc(peers & lt);
      __syncwarp(active);
      if ((peers & lt) == 0) cnt[warp][d] = before + __popc(peers);   // lowest peer lane updates the counter
    }
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix over the 8 warps for digit = threadIdx.x
    int run = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      wbase[w][threadIdx.x] = run;
      run += cnt[w][threadIdx.x];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortRounds; ++k) {
    const int64_t i = sort_index(blockIdx.x, warp, k, lane);
    if (i < n) {
      const int d = (key[k] >> shift) & 255;
      const int64_t dst = (int64_t)offs[(int64_t)d * tiles + blockIdx.x] + wbase[warp][d] + pre[k];
      keys_out[dst] = key[k];
      vals_out[dst] = val[k];
    }
  }
}

// Sort pairs by the low `bits` bits of the key.  Result lands in (keys_out, vals_out); (keys_tmp, vals_tmp) and the
// inputs are used as ping-pong buffers only when more than one pass is needed (the inputs are never written).
static int stable_sort_pairs(const int32_t* keys_in, const uint32_t* vals_in, int32_t* keys_out, uint32_t* vals_out,
                             int32_t* keys_tmp, uint32_t* vals_tmp, int64_t n, int bits, int32_t* hist,
                             int32_t* scan_tmp, cudaStream_t stream) {
  const int passes = (bits + 7) / 8;
  const int64_t tiles = (n + kSortTile - 1) / kSortTile;
  const int32_t* kin = keys_in;
  const uint32_t* vin = vals_in;
  for (int p = 0; p < passes; ++p) {
    // the last pass must write (keys_out, vals_out): alternate so that it does
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    int32_t* ko = to_out ? keys_out : keys_tmp;
    uint32_t* vo = to_out ? vals_out : vals_tmp;
    radix_hist_kernel<<<(unsigned)tiles, kSortThreads, 0, stream>>>(kin, n, 8 * p, hist, tiles);
    int rc = launch_status("radix_hist_kernel");
    if (rc != DMP_OK) return rc;
    rc = exclusive_scan(hist, hist, 256 * tiles, scan_tmp, stream);
    if (rc != DMP_OK) return rc;
    radix_scatter_kernel<<<(unsigned)tiles, kSortThreads, 0, stream>>>(kin, vin, n, 8 * p, hist, tiles, ko, vo);
    rc = launch_status("radix_scatter_kernel");
    if (rc != DMP_OK) return rc;
    kin = ko;
    vin = vo;
  }
  return DMP_OK;
}

static inline int64_t align_up(int64_t x) { return (x + 255) & ~(int64_t)255; }

static int num_key_bits(int64_t N) {
  int bits = 1;
  while (((int64_t)1 << bits) < N) ++bits;
  return bits;
}

struct PlanWs {
  int64_t off_cnt;       // 3 x (N+1) int32 histograms
  int64_t off_keys_out;  // E int32 sorted keys (discarded)
  int64_t off_vals_in;   // E uint32 edge id | rev flag
  int64_t off_keys_tmp;  // E int32   radix ping-pong
  int64_t off_vals_tmp;  // E uint32  radix ping-pong
  int64_t off_hist;      // 256 x tiles int32 digit histogram / offsets
  int64_t off_scan;      // scratch of the exclusive scans
  int64_t total;
};

static int plan_ws_layout(int64_t N, int64_t E, PlanWs* w) {
  const int64_t tiles = (E + kSortTile - 1) / kSortTile;
  const int64_t hist_ints = 256 * (tiles > 0 ? tiles : 1);
  const int64_t scan_ints = scan_tmp_ints(hist_ints > N + 1 ? hist_ints : N + 1) + 8;
  int64_t off = 0;
  w->off_cnt = off; off += align_up(3 * (N + 1) * 4);
  w->off_keys_out = off; off += align_up(E * 4);
  w->off_vals_in = off; off += align_up(E * 4);
  w->off_keys_tmp = off; off += align_up(E * 4);
  w->off_vals_tmp = off; off += align_up(E * 4);
  w->off_hist = off; off += align_up(hist_ints * 4);
  w->off_scan = off; off += align_up(scan_ints * 4);
  w->total = off;
  return DMP_OK;
}

}  // namespace dmp

extern "C" int dmp_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t* bytes_host) {
  using namespace dmp;
  DMP_CHECK_ARG(bytes_host != nullptr, "plan_workspace_bytes: null output");
  DMP_CHECK_ARG(num_nodes >= 0 && num_edges >= 0 && num_nodes < 0x7fffffffLL && num_edges < 0x7fffffffLL,
                "plan: N and E must be in [0, 2^31)");
  PlanWs w;
  int rc = plan_ws_layout(num_nodes, num_edges, &w);
  if (rc != DMP_OK) return rc;
  *bytes_host = w.total;
  return DMP_OK;
}

extern "C" int dmp_plan_build(const int64_t* src, const int64_t* dst, const uint8_t* rev,
                              const int64_t* out_deg, int64_t N, int64_t E, const float* coef_lut,
                              int64_t lut_len, int32_t* dst32, int32_t* a32, int32_t* b32,
                              int32_t* csc_indptr, int32_t* csc_eid, int32_t* a_indptr, int32_t* a_eid,
                              int32_t* b_indptr, int32_t* b_eid, int64_t* out_deg_out, float* coef,
                              int32_t* status, void* ws, int64_t ws_bytes, void* stream_) {
  using namespace dmp;
  cudaStream_t stream = (cudaStream_t)stream_;
  DMP_CHECK_ARG(N >= 0 && E >= 0 && N < 0x7fffffffLL && E < 0x7fffffffLL, "plan: N and E must be in [0, 2^31)");
  DMP_CHECK_ARG(csc_indptr && a_indptr && b_indptr && status, "plan: null indptr/status output");
  DMP_CHECK_ARG(N == 0 || out_deg_out != nullptr, "plan: null out_deg_out");
  DMP_CHECK_ARG(E == 0 || (src && dst && dst32 && a32 && b32 && csc_eid && a_eid && b_eid && coef),
                "plan: null edge array");
  PlanWs w;
  int rc = plan_ws_layout(N, E, &w);
  if (rc != DMP_OK) return rc;
  DMP_CHECK_ARG(ws != nullptr && ws_bytes >= w.total, "plan: workspace too small (%lld < %lld)",
                (long long)ws_bytes, (long long)w.total);
  char* base = static_cast<char*>(ws);
  int32_t* cnt = reinterpret_cast<int32_t*>(base + w.off_cnt);
  int32_t* cnt_dst = cnt;
  int32_t* cnt_a = cnt + (N + 1);
  int32_t* cnt_b = cnt + 2 * (N + 1);
  int32_t* keys_out = reinterpret_cast<int32_t*>(base + w.off_keys_out);
  uint32_t* vals_in = reinterpret_cast<uint32_t*>(base + w.off_vals_in);
  int32_t* keys_tmp = reinterpret_cast<int32_t*>(base + w.off_keys_tmp);
  uint32_t* vals_tmp = reinterpret_cast<uint32_t*>(base + w.off_vals_tmp);
  int32_t* hist = reinterpret_cast<int32_t*>(base + w.off_hist);
  int32_t* scan_tmp = reinterpret_cast<int32_t*>(base + w.off_scan);

  DMP_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
  DMP_CUDA_OK(cudaMemsetAsync(cnt, 0, 3 * (N + 1) * sizeof(int32_t), stream));
  if (out_deg == nullptr && N > 0) DMP_CUDA_OK(cudaMemsetAsync(out_deg_out, 0, N * sizeof(int64_t), stream));
  if (out_deg != nullptr && N > 0)
    DMP_CUDA_OK(cudaMemcpyAsync(out_deg_out, out_deg, N * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));

  const bool has_rev = rev != nullptr;
  if (E > 0) {
    const int64_t need = (E + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    const unsigned grid = (unsigned)(need < cap ? need : cap);
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "int64 atomics");
    roles_kernel<<<grid, kThreads, 0, stream>>>(
        src, dst, rev, N, E, dst32, a32, b32, vals_in, cnt_dst, cnt_a, cnt_b,
        out_deg == nullptr ? reinterpret_cast<unsigned long long*>(out_deg_out) : nullptr, status);
    rc = launch_status("roles_kernel");
    if (rc != DMP_OK) return rc;
  }

  // indptr = exclusive scan of the (N+1)-long histograms (last slot is 0 -> indptr[N] = E)
  rc = exclusive_scan(cnt_dst, csc_indptr, N + 1, scan_tmp, stream);
  if (rc != DMP_OK) return rc;
  const int bits = num_key_bits(N);
  if (E > 0) {
    rc = stable_sort_pairs(dst32, vals_in, keys_out, reinterpret_cast<uint32_t*>(csc_eid), keys_tmp, vals_tmp, E, bits,
                           hist, scan_tmp, stream);
    if (rc != DMP_OK) return rc;
  }
  rc = exclusive_scan(cnt_a, a_indptr, N + 1, scan_tmp, stream);
  if (rc != DMP_OK) return rc;
  rc = exclusive_scan(cnt_b, b_indptr, N + 1, scan_tmp, stream);
  if (rc != DMP_OK) return rc;
  if (has_rev) {
    if (E > 0) {
      rc = stable_sort_pairs(a32, vals_in, keys_out, reinterpret_cast<uint32_t*>(a_eid), keys_tmp, vals_tmp, E, bits,
                             hist, scan_tmp, stream);
      if (rc != DMP_OK) return rc;
      rc = stable_sort_pairs(b32, vals_in, keys_out, reinterpret_cast<uint32_t*>(b_eid), keys_tmp, vals_tmp, E, bits,
                             hist, scan_tmp, stream);
      if (rc != DMP_OK) return rc;
    }
  } else {
    // a == dst: the a-structure is the CSC (copied, not re-sorted); b == src: the CSR.
    if (E > 0) {
      DMP_CUDA_OK(cudaMemcpyAsync(a_eid, csc_eid, E * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
      rc = stable_sort_pairs(b32, vals_in, keys_out, reinterpret_cast<uint32_t*>(b_eid), keys_tmp, vals_tmp, E, bits,
                             hist, scan_tmp, stream);
      if (rc != DMP_OK) return rc;
    }
  }
  if (E > 0) {
    const int64_t need = (E + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    coef_kernel<<<(unsigned)(need < cap ? need : cap), kThreads, 0, stream>>>(dst32, out_deg_out, coef_lut,
                                                                              lut_len, coef, E);
    rc = launch_status("coef_kernel");
    if (rc != DMP_OK) return rc;
  }
  return DMP_OK;
}
