/* dmp_b200.h -- C ABI of libdmp_b200.so: the sparse core of the DMPNN dual message-passing layer,
 * hand-written CUDA for sm_100a (B200).
 *
 * The reference (HKUST-KnowComp/DualMessagePassing) has no FFI: its boundary is the Python module
 * `DMPLayer.forward(graph, node_feat, edge_feat)` (SubgraphCountingMatching/models/dmpnn.py:158-166)
 * and `DualGraphConv.forward(graph, node_feat, edge_feat, edge_norm)`
 * (UnsupervisedNodeClassification/Model/DMPNN/src/model.py:267-273), which reach native code only
 * through DGL (`update_all` / `apply_edges` / `fn.sum`) and ATen.  The entry points below are what a
 * binding for that path would call instead of DGL; each one cites the reference lines it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - sizes are int64_t, indices handed in by the caller are int64 (DGL's id type), indices the
 *     library produces are int32 (E and N must be < 2^31);
 *   - feature matrices are row-major fp32 with an explicit leading dimension `ld*` (in floats);
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing allocates;
 *   - return value: 0 = ok, negative = error, message via dmp_last_error() (thread-local);
 *   - all floating-point reductions run in a fixed order (ascending edge id inside a segment,
 *     round-to-nearest adds/multiplies, no FMA contraction, no atomics): results are bit-reproducible
 *     and equal to a sequential CPU loop in edge-id order.
 */
#ifndef DMP_B200_H_
#define DMP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DMP_API __attribute__((visibility("default")))
#else
#define DMP_API
#endif

#define DMP_OK 0
#define DMP_ERR_INVALID (-1)
#define DMP_ERR_CUDA (-2)
#define DMP_ERR_UNSUPPORTED (-3)

/* bit 31 of an entry of a segment's edge-id list carries the edge's reversed flag */
#define DMP_EID_MASK 0x7fffffff
#define DMP_REV_BIT 0x80000000u

/* dmp_segment_reduce mode bits */
#define DMP_SEG_SIGN_BY_REV 1 /* message sign: -1 on forward edges, +1 on reversed edges (dmpnn.py:113,121) */
#define DMP_SEG_NEGATE_OUT 2  /* out = -(sum)  (used for dQ_s = -SB in backward)                         */
#define DMP_SEG_ONLY_FWD 4    /* skip reversed edges (their rows are not even loaded)                     */
#define DMP_SEG_ONLY_REV 8    /* skip forward edges                                                        */
#define DMP_SEG_SPLIT_BY_REV 16 /* two sums per segment in one pass: forward edges -> out[:, 0:H], reversed edges ->
                                   out[:, H:2H] (ld_out >= 2H; no base / bias / filter).  Each half equals the
                                   ONLY_FWD / ONLY_REV result bit for bit.  By linearity of the node update
                                   sum_e s_e n_e (X_e W) = (sum_e s_e n_e X_e) W  (dmpnn.py:113-133) this replaces the
                                   edge-sized projection by two node-sized ones. */

#define DMP_SEG_SHORT 32 /* hint: segments hold only a few rows on average (destination-range partition: ~2.5 local edges
                            per global segment): a higher-occupancy variant with fewer rows in flight per lane.  Same
                            additions in the same order -- results are bit-identical with or without the hint. */

/* dmp_edge_update order */
#define DMP_ORDER_SCM 0 /* ((eloop + add) + agg) + ebias   dmpnn.py:147-149  */
#define DMP_ORDER_UNC 1 /* ((eloop + agg) + add) + ebias   model.py:257-259  */
#define DMP_EDGE_MIRRORED_HALVES 16 /* OR-ed into `order`: hint that edge e + E/2 is the reverse of edge e (one graph after
                                       the reversed-edge append, train.py:299-327), so both gather the same Q_d / Q_s rows:
                                       the pair is processed together and the rows are fetched once.  Results are
                                       identical with or without the hint (pairs whose endpoints differ fall back). */

/* activation ids for the fused epilogues (dmpnn.py:138,154 when num_mlp_layers == 0) */
#define DMP_ACT_NONE 0
#define DMP_ACT_RELU 1
#define DMP_ACT_LEAKY_RELU 2 /* slope passed explicitly (reference: 1/5.5, constants.py:10) */
#define DMP_ACT_TANH 3
#define DMP_ACT_SIGMOID 4
/* dmp_gate_residual_backward only: OR-ed into `act` when `x` holds the activation OUTPUT y = act(pre)
 * instead of the pre-activation (lets the forward apply the activation in place and keep one tensor) */
#define DMP_ACT_FROM_OUTPUT 16

DMP_API const char* dmp_last_error(void);
DMP_API int dmp_version(void);
/* The tensor-core kernels are persistent, one CTA per SM, with a static tile schedule: when another kernel (an NCCL
 * collective overlapped with the layer, parallel.py) occupies some SMs, the CTAs that cannot become resident start late
 * and stretch the launch to up to twice its time.  dmp_set_sm_reserve(n) (thread-local, default 0) makes the following
 * launches of this thread use 148 - n CTAs, leaving n SMs to the collective. */
DMP_API int dmp_set_sm_reserve(int sms);

/* ------------------------------------------------------------------------------------------------
 * Graph plan (row A0): replaces DGL's COO->CSC conversion behind `fn.sum` (dmpnn.py:92,163),
 * `graph.out_degrees()` (dmpnn.py:100-101), the per-edge endpoint gathers' index set-up
 * (`edges.src/dst`, dmpnn.py:112,120) and the degree term of dmpnn.py:144-146.
 *
 *   src,dst   int64 [E]   COO in edge-id order          rev      uint8 [E] or NULL (is_reversed flag)
 *   out_deg   int64 [N] or NULL (NULL => computed as bincount(src), written to out_deg_out)
 *   coef_lut  fp32 [lut_len] or NULL: coef_lut[d] = 2*(1+log2(1+d)) precomputed by the host maths
 *             library so that small degrees are bit-identical to the reference's CPU log2; degrees
 *             >= lut_len use the device log2f.
 * outputs (caller-allocated):
 *   dst32,a32,b32 int32 [E]: destination; endpoint meeting W_dst (a = rev ? src : dst); endpoint
 *             meeting W_src (b = rev ? dst : src)
 *   csc_indptr/a_indptr/b_indptr int32 [N+1]; csc_eid/a_eid/b_eid int32 [E]: stable counting sort
 *             of edge ids by dst / a / b (ascending edge id inside a segment), bit 31 = rev flag
 *   out_deg_out int64 [N]; coef fp32 [E] = 2*(1+log2(1+out_deg[dst[e]]))
 *   status    int32 [1]: set non-zero on device if an endpoint is outside [0,N)
 *   ws        workspace of at least dmp_plan_workspace_bytes(N,E) bytes
 * If rev == NULL the a-structure equals the CSC and the b-structure is the CSR; they are still
 * written so that callers need no special case.
 */
DMP_API int dmp_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t* bytes_host);
DMP_API int dmp_plan_build(const int64_t* src, const int64_t* dst, const uint8_t* rev, const int64_t* out_deg,
                   int64_t num_nodes, int64_t num_edges, const float* coef_lut, int64_t lut_len,
                   int32_t* dst32, int32_t* a32, int32_t* b32,
                   int32_t* csc_indptr, int32_t* csc_eid, int32_t* a_indptr, int32_t* a_eid,
                   int32_t* b_indptr, int32_t* b_eid, int64_t* out_deg_out, float* coef,
                   int32_t* status, void* ws, int64_t ws_bytes, void* stream);

/* out[j] = values[eid[j] & DMP_EID_MASK]  (per-segment-position copy of a per-edge scalar, e.g. the
 * UNC `norm`, model.py:234-235), so the reduce kernel streams it instead of gathering it. */
DMP_API int dmp_permute_edge_scalar(const int32_t* eid, const float* values, float* out, int64_t num_edges,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Segment reduce (rows A3 + A4 bias part; backward: autograd index_add_ of the endpoint gathers):
 *
 *   out[x,:] = ((base[x,:] + sum_{j in [indptr[x], indptr[x+1])} w_j * sgn_j * V[e_j, off_j : off_j+H]) + bias)
 *
 * with e_j = eid[j] & DMP_EID_MASK, r_j = eid[j] >> 31, sgn_j = (mode & SIGN_BY_REV) ? (r_j ? +1 : -1) : +1,
 * off_j = r_j ? rev_col_offset : 0, w_j = w_perm ? w_perm[j] : 1 (applied as a separate rounding, like
 * `node_msg * norm`), accumulated sequentially from 0.0f in ascending j.  base/bias/w_perm may be NULL.
 * Replaces `fn.sum(node_msg -> node_agg)` + `matmul(nloop) + agg + nbias` (dmpnn.py:92,131-133) and, in
 * backward, the scatter-adds of d(edge_msg) onto the endpoint rows.
 */
DMP_API int dmp_segment_reduce(const int32_t* indptr, const int32_t* eid, const float* w_perm,
                       const float* V, int64_t ldV, int64_t rev_col_offset,
                       const float* base, int64_t ld_base, const float* bias,
                       float* out, int64_t ld_out, int64_t num_segments, int64_t H, int mode,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Edge update (rows A2 edge_msg + A5): per edge e
 *   msg = Qd[a32[e],:] - Qs[b32[e],:]                      (dmpnn.py:112,120,123)
 *   add = coef[e] * P[e,:]                                  (dmpnn.py:146)
 *   out[e,:] = ((S[e,:] + add) + msg) + ebias   (SCM)   or  ((S[e,:] + msg) + add) + ebias   (UNC)
 * S = X_e*W_eloop, P = X_e*(W_src-W_dst), Qd = X_v*W_dst, Qs = X_v*W_src are produced by the dense
 * stage.  edge_agg (optional, may be NULL) receives msg, mirroring the reference's frame side effect
 * `edata["edge_agg"]` (dmpnn.py:126).  out may alias S.
 * P may be NULL: S then already holds `eloop + add` (DMP_DUAL_STORE output of dmp_gemm_tf32x3_dual, rounded like the
 * reference's `matmul(h, eloop) + add`), and out = (S + msg) + ebias -- the SCM association, bit for bit.
 */
DMP_API int dmp_edge_update(const int32_t* a32, const int32_t* b32, const float* coef,
                    const float* S, int64_t ldS, const float* P, int64_t ldP,
                    const float* Qd, int64_t ldQd, const float* Qs, int64_t ldQs, const float* ebias,
                    float* out, int64_t ld_out, float* edge_agg, int64_t ld_agg,
                    int64_t num_edges, int64_t H, int order, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward edge-side gather (gSpMM backward of `fn.sum` + derivative of the degree term):
 *   T[e, off_e : off_e+H] = sgn_e * gN[dst32[e],:] (* norm[e])   sgn_e = rev[e] ? +1 : -1   (d node_msg)
 *   CG[e,:] = coef[e] * gE[e,:]                            (d P)          -- skipped if CG == NULL
 * off_e = rev[e] ? T_rev_col_offset : 0 (mirror of dmp_segment_reduce's rev_col_offset: with the
 * two-branch [E,2H] message buffer the gradient lands in the half the edge's branch read from).
 * rev may be NULL (all forward), norm may be NULL.  T may be NULL to produce CG only.
 * gN_rev (may be NULL = gN): table gathered for REVERSED edges.  With gN = dL/dnode_pre it is the plain
 * gSpMM backward; with gN = gN·W_in^T and gN_rev = gN·W_out^T (node-sized tables) T is directly the
 * contribution of the node aggregation to dL/dX_e, which the dense backward then accumulates onto.
 */
DMP_API int dmp_edge_backward(const int32_t* dst32, const uint8_t* rev, const float* norm, const float* coef,
                      const float* gN, const float* gN_rev, int64_t ld_gN, const float* gE, int64_t ld_gE,
                      float* T, int64_t ldT, int64_t T_rev_col_offset, float* CG, int64_t ldCG,
                      int64_t num_edges, int64_t H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused elementwise epilogue of the rep-net loop (row A7, dmpnn.py:236-241,266-275):
 *   y = act(x)            (act only when the layer has no MLP, dmpnn.py:138,154)
 *   y = y * gate[row]     (gate [rows] or NULL; pattern-side masked_fill is the 0/1 gate)
 *   out = prev + y        (prev NULL => no residual)
 * and its backward: gx = gout * gate[row] * act'(x).
 */
DMP_API int dmp_gate_residual(const float* x, int64_t ldx, const float* gate, const float* prev, int64_t ld_prev,
                      float* out, int64_t ld_out, int64_t rows, int64_t H, int act, float slope,
                      void* stream);
DMP_API int dmp_gate_residual_backward(const float* gout, int64_t ld_gout, const float* x, int64_t ldx,
                               const float* gate, float* gx, int64_t ld_gx, int64_t rows, int64_t H,
                               int act, float slope, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Edge-sized projection on the tcgen05 tensor cores with fp32-level accuracy (3xTF32 split):
 *
 *   D[M,N] = epilogue( (row_scale ⊙ A)[M,K] · Bt[N,K]^T )        N, K in {64,128}; Bt is the weight in
 *                                                                 nn.Linear layout ([out, in], row-major)
 * Replaces the per-edge `th.matmul(...)` calls of dmpnn.py:112-113,120-121,146-147 and the MLP Linears
 * (dmpnn.py:45-52) and their autograd transposes.  row_scale [M] (may be NULL) multiplies each row of A
 * first, as a separately rounded fp32 product (fuses `coef ⊙ gE`); with DMP_EPI_ACCUMULATE the same factor is applied
 * to the accumulated row instead (D[r,:] += row_scale[r] * (A[r,:]·Bt^T): equal up to fp32 rounding).  The three
 * products of the split are issued cross terms first (the tensor core's fp32 accumulator truncates: DESIGN.md 4.1).  epilogue = DMP_ACT_* id, optionally
 * OR-ed with DMP_EPI_MUL_ACT_GRAD (D = acc * act'(aux), aux = activation OUTPUT [M,N]) and/or
 * DMP_EPI_ACCUMULATE (D += result).  bias [N] may be NULL.  D must not alias A.
 */
#define DMP_EPI_MUL_ACT_GRAD 32
#define DMP_EPI_ACCUMULATE 64
DMP_API int dmp_gemm_tf32x3(const float* A, int64_t lda, const float* row_scale, const float* Bt, int64_t ldb,
                            const float* bias, const float* aux, int64_t ld_aux, float* D, int64_t ldd,
                            int64_t M, int64_t N, int64_t K, int epilogue, float slope, void* stream);

/* Two projections of the SAME streamed operand in one pass (the operand is read from HBM once, both products stay in
 * tensor memory):
 *   DMP_DUAL_STORE        D[r,:]  = A[r,:]·W1t^T + row_scale[r] * (A[r,:]·W2t^T)
 *   DMP_DUAL_ACCUMULATE   D[r,:]  = (D[r,:] + A[r,:]·W1t^T) + row_scale[r] * (A[r,:]·W2t^T)
 *   DMP_DUAL_SEPARATE     D = A·W1t^T ,  D2 = A·W2t^T                (row_scale must be NULL)
 * W1t, W2t are [N,K] (nn.Linear layout) with a common leading dimension; N, K in {64,128}.  Replaces the pair
 * `matmul(h, eloop_weight) + 2*(1+d) * matmul(h, src_weight - dst_weight)` of dmpnn.py:146-147 (STORE: the sum is
 * rounded exactly like `eloop + add`, so dmp_edge_update takes it as S with P = NULL; SEPARATE for the UNC association
 * order of model.py:257) and its autograd transpose `dX_e += gE·W_eloop^T + (coef ⊙ gE)·(W_src-W_dst)^T` (ACCUMULATE).
 * row_scale may be NULL (= 1). */
#define DMP_DUAL_STORE 0
#define DMP_DUAL_ACCUMULATE 1
#define DMP_DUAL_SEPARATE 2
DMP_API int dmp_gemm_tf32x3_dual(const float* A, int64_t lda, const float* W1t, const float* W2t, int64_t ldw,
                                 const float* row_scale, float* D, int64_t ldd, float* D2, int64_t ldd2,
                                 int64_t M, int64_t N, int64_t K, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight-gradient reduction on the tensor cores (3xTF32 split, fp32-level accuracy):
 *
 *   D[M,N] (+)= sum_e (row_scale[e] * X[e,0:M])^T * G[e,0:N]          M, N in {64,128}
 *
 * Replaces autograd's `X.t() @ G` for dW_eloop, dW_src/dst, dW_in/out and the MLP weight gradients
 * (transposes of the matmuls at dmpnn.py:112-113,120-121,146-147,45-52).  Deterministic: every CTA owns
 * a contiguous edge range, partials are added in CTA order.  workspace >= dmp_gemm_tn_workspace_bytes(M,N).
 * colsum_x [M] / colsum_g [N] (either may be NULL) additionally receive sum_e X[e,:] (unscaled) / sum_e G[e,:]
 * -- the bias gradients (dnbias, debias, MLP db) come for free from the rows the kernel streams anyway.
 */
DMP_API int dmp_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t* bytes_host);
DMP_API int dmp_gemm_tn_tf32x3(const float* X, int64_t ldx, const float* row_scale, const float* G, int64_t ldg,
                               float* D, int64_t ldd, float* colsum_x, float* colsum_g, int64_t E, int64_t M,
                               int64_t N, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm1d between the two Linears of the node / edge MLP (dmpnn.py:45-52 with batch_norm=True -- the class
 * default; UNC model.py:145-156 always).  The statistics run over all node rows / all edge rows, so this cannot be a
 * GEMM epilogue; it is column-reduction + elementwise work, deterministic (fixed-order partial sums, no float atomics).
 *   dmp_bn_stats      mean[c] = sum_r x[r,c] / rows ;  var[c] = sum_r (x[r,c] - mean[c])^2 / rows      (two passes)
 *   dmp_bn_act        out = act(((x - mean) * invstd) * gamma + beta)              gamma / beta may be NULL
 *   dmp_bn_backward   dbeta = sum_r g, dgamma = sum_r g * xhat, and
 *                     gx = gamma*invstd * (g - dbeta/rows - xhat*dgamma/rows)   (training != 0: batch statistics)
 *                     gx = gamma*invstd * g                                     (training == 0: running statistics)
 *                     gx may alias g.
 * workspace: dmp_bn_workspace_bytes(H) bytes, ZERO-INITIALISED ONCE by the caller (the kernels leave it zeroed).
 */
DMP_API int dmp_bn_workspace_bytes(int64_t H, int64_t* bytes_host);
DMP_API int dmp_bn_stats(const float* x, int64_t ldx, int64_t rows, int64_t H, float* mean, float* var,
                         void* workspace, int64_t workspace_bytes, void* stream);
DMP_API int dmp_bn_act(const float* x, int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, float* out, int64_t ld_out, int64_t rows, int64_t H, int act, float slope,
                       void* stream);
DMP_API int dmp_bn_backward(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* mean,
                            const float* invstd, const float* gamma, float* gx, int64_t ld_gx, float* dgamma,
                            float* dbeta, int64_t rows, int64_t H, int training, void* workspace,
                            int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side batching (row N1): the disjoint union `GraphAdjDataset.batchify` -> `dgl.batch` builds on the CPU every
 * step (SCM/dataset.py:1604-1611,1321-1328), over graphs that carry their reversed edges (`add_reversed_edges`,
 * SCM/train.py:299-327), from a dataset that lives in HBM as flat arrays:
 *   node_offsets / edge_offsets  int64 [G+1]  graph g owns nodes [no[g], no[g+1]) and ORIGINAL edges [eo[g], eo[g+1])
 *   u, v                         int64        endpoints LOCAL to the graph;  node_label / edge_label int64 (may be NULL)
 *   sel                          int64 [B]    the graphs of this batch, in batch order
 * dmp_batch_offsets writes the exclusive prefix sums of the selected graphs' node counts and (doubled if add_reversed)
 * edge counts ([B+1] each; DGL: node / edge id offsets of graph i in the batch).  dmp_batch_fill then writes, per batch
 * graph i: its E0 forward edges (u+off, v+off), then (add_reversed) its E0 reversed edges (v+off, u+off) with rev = 1
 * -- per-graph edge order preserved, exactly dgl.batch of add_reversed_edges'ed graphs; labels gathered alongside
 * (the caller applies `label += max_ngel` on reversed edges, train.py:310); node_graph / edge_graph [total] = owning
 * batch index (may be NULL).  total_nodes / total_edges are host integers (the host keeps the per-graph sizes).
 * padded != 0: total_nodes / total_edges are FIXED bucket sizes >= the real totals (which the kernel reads from the
 * offsets, on the device) -- the rest is filled with dummy nodes (graph id B, label 0) and forward self-loops spread
 * round-robin over them (total_nodes must exceed the real node count when any edge is padded): a batch of any composition then has the same shapes and addresses, which is what lets the whole
 * training step be captured in one CUDA graph (train_step.GraphedTrainStep).
 */
DMP_API int dmp_batch_offsets(const int64_t* sel, int64_t num_selected, const int64_t* node_offsets,
                              const int64_t* edge_offsets, int add_reversed, int64_t* batch_node_offsets,
                              int64_t* batch_edge_offsets, void* stream);
DMP_API int dmp_batch_fill(const int64_t* sel, int64_t num_selected, const int64_t* node_offsets,
                           const int64_t* edge_offsets, const int64_t* u, const int64_t* v, const int64_t* node_label,
                           const int64_t* edge_label, const int64_t* batch_node_offsets,
                           const int64_t* batch_edge_offsets, int64_t total_nodes, int64_t total_edges,
                           int add_reversed, int padded, int64_t* src, int64_t* dst, uint8_t* rev,
                           int64_t* node_label_out, int64_t* edge_label_out, int64_t* node_graph, int64_t* edge_graph,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ragged <-> padded (row N2): `split_and_batchify_graph_feats(batched_graph_feats, graph_sizes, pre_pad)`
 * (SCM/utils/dl.py:51-81; called on the rep-net outputs, basemodel.py:1579-1590) without the Python loop over the batch
 * and without its `.tolist()` synchronisation.  offsets int64 [B+1] = exclusive prefix sums of the graph sizes.
 *   dmp_ragged_pad     out[b, j, :] = x[offsets[b] + j - start_b, :] inside the graph's window, 0 outside;
 *                      mask[b, j] = 1 inside (start_b = pre_pad ? max_len - len_b : 0); out is [B*max_len, H]
 *   dmp_ragged_unpad   the inverse gather (its autograd transpose): out[offsets[b] + k, :] = padded[b, start_b + k, :]
 */
DMP_API int dmp_ragged_pad(const float* x, int64_t ldx, const int64_t* offsets, int64_t num_graphs, int64_t max_len,
                           int64_t H, int pre_pad, float* out, int64_t ld_out, uint8_t* mask, void* stream);
DMP_API int dmp_ragged_unpad(const float* padded, int64_t ld, const int64_t* offsets, int64_t num_graphs,
                             int64_t max_len, int64_t H, int pre_pad, float* out, int64_t ld_out, int64_t total_rows,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMP_B200_H_ */
