// collate.cu -- device-side batching of graphs (row N1 of SURVEY.md section 8f): the disjoint union the reference
// builds on the CPU main process every step (`GraphAdjDataset.batchify` -> `dgl.batch`, SCM/dataset.py:1604-1611,
// 1321-1328) over graphs that carry their reversed edges (`add_reversed_edges`, SCM/train.py:299-327: per graph
// [E0 forward edges | E0 reversed edges (v,u), is_reversed = 1]), written directly in HBM from a device-resident
// dataset.  Integer-only, bit-exact against the numpy restatement (tests/test_gpu_collate.py).
//
//   dmp_batch_offsets   one CTA: per selected graph its node / doubled-edge counts and their exclusive prefix sums
//   dmp_batch_fill      one thread per output node / output edge: owner graph by binary search in the prefix sums,
//                       then gather + renumber (node offset = prefix sum of node counts, dgl.batch semantics)
#include "common.cuh"

namespace dmp {

__global__ void __launch_bounds__(1024) batch_offsets_kernel(const int64_t* __restrict__ sel, int64_t B,
                                                             const int64_t* __restrict__ noff,
                                                             const int64_t* __restrict__ eoff, int reversed,
                                                             int64_t* __restrict__ new_noff, int64_t* __restrict__ new_eoff) {
  // B is small (<= a few thousand pairs): a single CTA, serial carry between 1024-wide chunks
  __shared__ int64_t sn[1024], se[1024];
  __shared__ int64_t carry_n, carry_e;
  if (threadIdx.x == 0) { carry_n = 0; carry_e = 0; new_noff[0] = 0; new_eoff[0] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < B; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int64_t n = 0, e = 0;
    if (i < B) {
      const int64_t g = sel[i];
      n = noff[g + 1] - noff[g];
      e = (eoff[g + 1] - eoff[g]) * (reversed ? 2 : 1);
    }
    sn[threadIdx.x] = n; se[threadIdx.x] = e;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {          // Hillis-Steele inclusive scan
      int64_t a = 0, b = 0;
      if ((int)threadIdx.x >= d) { a = sn[threadIdx.x - d]; b = se[threadIdx.x - d]; }
      __syncthreads();
      sn[threadIdx.x] += a; se[threadIdx.x] += b;
      __syncthreads();
    }
    if (i < B) { new_noff[i + 1] = carry_n + sn[threadIdx.x]; new_eoff[i + 1] = carry_e + se[threadIdx.x]; }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_n += sn[1023]; carry_e += se[1023]; }
    __syncthreads();
  }
}

__device__ __forceinline__ int64_t owner_of(const int64_t* __restrict__ off, int64_t B, int64_t x) {
  int64_t lo = 0, hi = B;                       // largest i with off[i] <= x  (off[0] = 0, off[B] = total > x)
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= x) lo = mid; else hi = mid;
  }
  return lo;
}

struct BatchFillParams {
  const int64_t* sel; int64_t B;
  const int64_t* noff; const int64_t* eoff;
  const int64_t* u; const int64_t* v; const int64_t* vlabel; const int64_t* elabel;
  const int64_t* new_noff; const int64_t* new_eoff;
  int64_t total_nodes, total_edges;
  int reversed;
  int padded;     // total_* are PADDED sizes; the real totals are new_noff[B] / new_eoff[B] (device): rows past them are dummies
  int64_t* src; int64_t* dst; uint8_t* rev; int64_t* vlabel_out; int64_t* elabel_out; int64_t* node_graph;
  int64_t* edge_graph;
};

__global__ void __launch_bounds__(kThreads) batch_fill_kernel(const BatchFillParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = p.total_nodes + p.total_edges;
  const int64_t real_nodes = p.padded ? __ldg(p.new_noff + p.B) : p.total_nodes;
  const int64_t real_edges = p.padded ? __ldg(p.new_eoff + p.B) : p.total_edges;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    if (t < p.total_nodes) {
      if (t >= real_nodes) {          // padding: isolated dummy node of a dummy graph B (dropped by every pooling)
        if (p.vlabel_out) p.vlabel_out[t] = 0;
        if (p.node_graph) p.node_graph[t] = p.B;
        continue;
      }
      const int64_t i = owner_of(p.new_noff, p.B, t);
      const int64_t g = __ldg(p.sel + i);
      const int64_t from = __ldg(p.noff + g) + (t - __ldg(p.new_noff + i));
      if (p.vlabel_out) p.vlabel_out[t] = __ldg(p.vlabel + from);
      if (p.node_graph) p.node_graph[t] = i;
    } else {
      const int64_t k = t - p.total_nodes;
      if (k >= real_edges) {          // padding: forward self-loops spread round-robin over the dummy nodes -- they touch
                                      // no real row, and no dummy node becomes a hub (a 15 k-edge segment would
                                      // serialise one lane group of every segment reduce for ~1 ms)
        const int64_t dummies = p.total_nodes - real_nodes;      // >= 1 whenever an edge is padded (caller's contract)
        const int64_t node = real_nodes + (k - real_edges) % (dummies > 0 ? dummies : 1);
        p.src[k] = node;
        p.dst[k] = node;
        if (p.rev) p.rev[k] = 0;
        if (p.elabel_out) p.elabel_out[k] = 0;
        if (p.edge_graph) p.edge_graph[k] = p.B;
        continue;
      }
      const int64_t i = owner_of(p.new_eoff, p.B, k);
      const int64_t g = __ldg(p.sel + i);
      const int64_t e0 = __ldg(p.eoff + g + 1) - __ldg(p.eoff + g);
      const int64_t pos = k - __ldg(p.new_eoff + i);
      const bool r = p.reversed && pos >= e0;
      const int64_t orig = __ldg(p.eoff + g) + (r ? pos - e0 : pos);
      const int64_t uu = __ldg(p.u + orig), vv = __ldg(p.v + orig);
      const int64_t off = __ldg(p.new_noff + i);
      p.src[k] = (r ? vv : uu) + off;
      p.dst[k] = (r ? uu : vv) + off;
      if (p.rev) p.rev[k] = r ? 1 : 0;
      if (p.elabel_out) p.elabel_out[k] = __ldg(p.elabel + orig);
      if (p.edge_graph) p.edge_graph[k] = i;
    }
  }
}

// ---- ragged <-> padded (row N2): `split_and_batchify_graph_feats` (SCM/utils/dl.py:51-81) without the Python loop over
// the batch and without the `.tolist()` device->host synchronisation.  One warp per padded row, lanes along the features.
struct RaggedParams {
  const float* x; int64_t ldx;
  const int64_t* off; int64_t B; int64_t max_len; int H; int pre_pad;
  float* out; int64_t ld_out; uint8_t* mask; int64_t total_rows;
};

template <bool UNPAD>
__global__ void __launch_bounds__(kThreads) ragged_kernel(const RaggedParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = UNPAD ? p.total_rows : p.B * p.max_len;
  for (int64_t r = warp; r < rows; r += nwarps) {
    if (!UNPAD) {
      const int64_t b = r / p.max_len, j = r - b * p.max_len;
      const int64_t o = __ldg(p.off + b);
      int64_t l = __ldg(p.off + b + 1) - o;
      if (l > p.max_len) l = p.max_len;
      const int64_t start = p.pre_pad ? p.max_len - l : 0;
      const bool valid = j >= start && j < start + l;
      if (lane == 0 && p.mask != nullptr) p.mask[r] = valid ? 1 : 0;
      const float* src = p.x + (o + (j - start)) * p.ldx;
      float* dst = p.out + r * p.ld_out;
      for (int c = lane; c < p.H; c += 32) dst[c] = valid ? __ldg(src + c) : 0.0f;
    } else {
      const int64_t b = owner_of(p.off, p.B, r);
      const int64_t o = __ldg(p.off + b);
      int64_t l = __ldg(p.off + b + 1) - o;
      const int64_t k = r - o;                        // position inside the graph
      const bool valid = k < p.max_len;               // rows truncated by max_len get a zero gradient
      if (l > p.max_len) l = p.max_len;
      const int64_t start = p.pre_pad ? p.max_len - l : 0;
      const float* src = p.x + (b * p.max_len + start + k) * p.ldx;
      float* dst = p.out + r * p.ld_out;
      for (int c = lane; c < p.H; c += 32) dst[c] = valid ? __ldg(src + c) : 0.0f;
    }
  }
}

}  // namespace dmp

static int ragged_launch(bool unpad, const float* x, int64_t ldx, const int64_t* offsets, int64_t B, int64_t max_len,
                         int64_t H, int pre_pad, float* out, int64_t ld_out, uint8_t* mask, int64_t total_rows,
                         void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(B >= 0 && max_len >= 0 && H >= 0 && total_rows >= 0 && H < (1ll << 31), "ragged: negative size");
  const int64_t rows = unpad ? total_rows : B * max_len;
  if (rows == 0) return DMP_OK;
  DMP_CHECK_ARG(offsets && out && (H == 0 || x) && ldx >= H && ld_out >= H, "ragged: bad operands");
  RaggedParams p;
  p.x = x; p.ldx = ldx; p.off = offsets; p.B = B; p.max_len = max_len; p.H = (int)H; p.pre_pad = pre_pad;
  p.out = out; p.ld_out = ld_out; p.mask = mask; p.total_rows = total_rows;
  const int64_t need = (rows * 32 + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 8;
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  if (unpad) ragged_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(p);
  else ragged_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(p);
  return launch_status("ragged_kernel");
}

extern "C" int dmp_ragged_pad(const float* x, int64_t ldx, const int64_t* offsets, int64_t num_graphs, int64_t max_len,
                              int64_t H, int pre_pad, float* out, int64_t ld_out, uint8_t* mask, void* stream) {
  return ragged_launch(false, x, ldx, offsets, num_graphs, max_len, H, pre_pad, out, ld_out, mask, 0, stream);
}

extern "C" int dmp_ragged_unpad(const float* padded, int64_t ld, const int64_t* offsets, int64_t num_graphs,
                                int64_t max_len, int64_t H, int pre_pad, float* out, int64_t ld_out, int64_t total_rows,
                                void* stream) {
  return ragged_launch(true, padded, ld, offsets, num_graphs, max_len, H, pre_pad, out, ld_out, nullptr, total_rows,
                       stream);
}

extern "C" int dmp_batch_offsets(const int64_t* sel, int64_t num_selected, const int64_t* node_offsets,
                                 const int64_t* edge_offsets, int add_reversed, int64_t* batch_node_offsets,
                                 int64_t* batch_edge_offsets, void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_selected >= 0, "batch_offsets: negative count");
  DMP_CHECK_ARG(batch_node_offsets && batch_edge_offsets && (num_selected == 0 || (sel && node_offsets && edge_offsets)),
                "batch_offsets: null pointer");
  batch_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sel, num_selected, node_offsets, edge_offsets, add_reversed,
                                                             batch_node_offsets, batch_edge_offsets);
  return launch_status("batch_offsets_kernel");
}

extern "C" int dmp_batch_fill(const int64_t* sel, int64_t num_selected, const int64_t* node_offsets,
                              const int64_t* edge_offsets, const int64_t* u, const int64_t* v, const int64_t* node_label,
                              const int64_t* edge_label, const int64_t* batch_node_offsets,
                              const int64_t* batch_edge_offsets, int64_t total_nodes, int64_t total_edges,
                              int add_reversed, int padded, int64_t* src, int64_t* dst, uint8_t* rev,
                              int64_t* node_label_out, int64_t* edge_label_out, int64_t* node_graph, int64_t* edge_graph,
                              void* stream) {
  using namespace dmp;
  DMP_CHECK_ARG(num_selected >= 0 && total_nodes >= 0 && total_edges >= 0, "batch_fill: negative size");
  if (total_nodes + total_edges == 0) return DMP_OK;
  DMP_CHECK_ARG(sel && node_offsets && edge_offsets && batch_node_offsets && batch_edge_offsets, "batch_fill: null pointer");
  DMP_CHECK_ARG(total_edges == 0 || (u && v && src && dst), "batch_fill: null edge arrays");
  DMP_CHECK_ARG((node_label_out == nullptr || node_label != nullptr) && (edge_label_out == nullptr || edge_label != nullptr),
                "batch_fill: label output without label input");
  BatchFillParams p;
  p.sel = sel; p.B = num_selected; p.noff = node_offsets; p.eoff = edge_offsets; p.u = u; p.v = v;
  p.vlabel = node_label; p.elabel = edge_label; p.new_noff = batch_node_offsets; p.new_eoff = batch_edge_offsets;
  p.total_nodes = total_nodes; p.total_edges = total_edges; p.reversed = add_reversed; p.padded = padded;
  p.src = src; p.dst = dst; p.rev = rev; p.vlabel_out = node_label_out; p.elabel_out = edge_label_out;
  p.node_graph = node_graph; p.edge_graph = edge_graph;
  const int64_t need = (total_nodes + total_edges + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 8;
  batch_fill_kernel<<<(unsigned)(need < cap ? need : cap), kThreads, 0, (cudaStream_t)stream>>>(p);
  return launch_status("batch_fill_kernel");
}
