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
//   2. ExclusiveSum     : three histograms -> three indptr arrays (cub::DeviceScan).
//   3. stable sort      : (key, edge id | rev<<31) pairs by key with cub::DeviceRadixSort restricted to
//                         ceil(log2 N) bits -- LSD radix sort is stable, so ids ascend inside a segment.
//   4. coef_kernel      : coef[e] = lut[deg] for deg < lut_len else 2*(1+log2f(1+deg)).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

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
  int64_t off_cub;
  int64_t cub_bytes;
  int64_t total;
};

static int plan_ws_layout(int64_t N, int64_t E, PlanWs* w) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                   (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)E, 0,
                                                   num_key_bits(N));
  cudaError_t e2 = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                 (int)(N + 1));
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("plan: cub size query failed");
    return DMP_ERR_CUDA;
  }
  int64_t off = 0;
  w->off_cnt = off; off += align_up(3 * (N + 1) * 4);
  w->off_keys_out = off; off += align_up(E * 4);
  w->off_vals_in = off; off += align_up(E * 4);
  w->off_cub = off;
  w->cub_bytes = (int64_t)(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
  off += align_up(w->cub_bytes);
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
  void* cub_ws = base + w.off_cub;
  size_t cub_bytes = (size_t)w.cub_bytes;

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
  DMP_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, cnt_dst, csc_indptr, (int)(N + 1), stream));
  const int bits = num_key_bits(N);
  if (E > 0)
    DMP_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, (const int32_t*)dst32, keys_out,
                                                (const uint32_t*)vals_in, reinterpret_cast<uint32_t*>(csc_eid),
                                                (int)E, 0, bits, stream));
  DMP_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, cnt_a, a_indptr, (int)(N + 1), stream));
  DMP_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, cnt_b, b_indptr, (int)(N + 1), stream));
  if (has_rev) {
    if (E > 0) {
      DMP_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, (const int32_t*)a32, keys_out,
                                                  (const uint32_t*)vals_in, reinterpret_cast<uint32_t*>(a_eid),
                                                  (int)E, 0, bits, stream));
      DMP_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, (const int32_t*)b32, keys_out,
                                                  (const uint32_t*)vals_in, reinterpret_cast<uint32_t*>(b_eid),
                                                  (int)E, 0, bits, stream));
    }
  } else {
    // a == dst: the a-structure is the CSC (copied, not re-sorted); b == src: the CSR.
    if (E > 0) {
      DMP_CUDA_OK(cudaMemcpyAsync(a_eid, csc_eid, E * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
      DMP_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, (const int32_t*)b32, keys_out,
                                                  (const uint32_t*)vals_in, reinterpret_cast<uint32_t*>(b_eid),
                                                  (int)E, 0, bits, stream));
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
