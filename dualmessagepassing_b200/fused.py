"""Whole-layer autograd function with explicit buffer management (the path every supported configuration takes).

`torch.autograd` over the composed ops of `layers.dual_message_passing` keeps ~14 edge-sized
temporaries alive (the reference keeps even more); at config 5
(E = 40 M, H = 128) one [E,H] fp32 tensor is 20.5 GB, so that composition does not fit 180 GB of HBM.
This function computes the same quantities but owns every edge-sized buffer: forward keeps only
{X_e (input), edge_pre, h1 (+ the pre-BatchNorm z when the MLP has BatchNorm)}; backward re-uses the saved
buffers in place ("Memory plan").

Supported: SCM flavour (dmpnn.py:111-156) and UNC flavour (model.py:222-265) -- association order, `norm`,
post-activation -- with an MLP of any depth (0 = activation only), with or without BatchNorm1d between the
Linears (training or eval statistics), activation in {none, relu, leaky_relu, tanh, sigmoid}.  Widths that are
not 64 / 128 are zero-padded by the caller (layers.py).  Anything else runs the composed path.
Backward: SURVEY.md Appendix A.2.
"""
import torch

from . import _lib
from .functional import (bn_act, bn_backward, bn_stats, edge_backward, edge_update, gemm_tf32x3, gemm_tf32x3_dual,
                         gemm_tn_tf32x3, segment_reduce_two_level)
from .functional import segment_reduce as _segment_reduce_seq


def segment_reduce(indptr, eid, V, H, *, plan=None, **kw):
    """Sequential-order reduce (DGL's fn.sum order, bit for bit) unless the plan asks for chunked long segments."""
    if plan is not None and plan.long_chunk:
        return segment_reduce_two_level(indptr, eid, V, H, chunk=int(plan.long_chunk), **kw)
    return _segment_reduce_seq(indptr, eid, V, H, **kw)

_ACT = {"none": _lib.ACT_NONE, "relu": _lib.ACT_RELU, "leaky_relu": _lib.ACT_LEAKY_RELU,
        "tanh": _lib.ACT_TANH, "sigmoid": _lib.ACT_SIGMOID}


def supported_activation(name):
    return name in _ACT


def _act_inplace(x, act, slope):
    """x <- act(x) with the sm_100a elementwise kernel (in place)."""
    if act == _lib.ACT_NONE or x.numel() == 0:
        return x
    rows, H = x.shape
    _lib.call("dmp_gate_residual", x.device, _lib.ptr(x), x.stride(0), None, None, 0, _lib.ptr(x), x.stride(0), rows, H,
              act, slope, _lib.stream_ptr(x.device), tag="act_inplace")
    return x


def _act_backward_inplace(g, y, act, slope):
    """g <- g * act'(.) expressed through the activation OUTPUT y (in place on g)."""
    if act == _lib.ACT_NONE or g.numel() == 0:
        return g
    rows, H = g.shape
    _lib.call("dmp_gate_residual_backward", g.device,
              _lib.ptr(g), g.stride(0), _lib.ptr(y), y.stride(0), None, _lib.ptr(g), g.stride(0), rows, H,
              act | _lib.ACT_FROM_OUTPUT, slope, _lib.stream_ptr(g.device), tag="act_bwd_inplace")
    return g


# Dense backend of the edge-/node-sized projections:
#   "auto"    tcgen05 3xTF32 kernels (fp32-level accuracy) when N, K in {64,128}, else cuBLAS
#   "cublas"  always torch.mm (cuBLAS sgemm)
DENSE_BACKEND = "auto"
_ACT_NAME = {v: k for k, v in _ACT.items()}

# below this many rows the persistent tcgen05 kernels' fixed cost (weight split per CTA, TMEM allocation, second
# reduce launch) exceeds what cuBLAS sgemm needs for the whole product
TC_MIN_ROWS = 16384
# The K = rows reductions pay off much earlier: a 64 x 64 result over 1 856 rows takes cuBLAS sgemm 15 us (one CTA walks the
# whole K), 24 us over 2 771 rows, 82 us over 20 516, and the bias gradients need a 13 us column-sum kernel on top; the
# reduction kernel spreads K over the SMs (3-7 us + a 4 us fixed-order second pass) and returns the column sums for free.
TN_MIN_ROWS = 512
# same-operand projection pairs in ONE launch (dmp_gemm_tf32x3_dual); False = two dmp_gemm_tf32x3 launches (A/B runs)
DUAL_GEMM = __import__("os").environ.get("DMP_DUAL_GEMM", "1") != "0"


def _dense_ok(t):
    return t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0


def _use_tc(A, Wt):
    return (DENSE_BACKEND == "auto" and A.dtype == torch.float32 and A.shape[0] >= TC_MIN_ROWS
            and Wt.shape[0] in (64, 128) and Wt.shape[1] in (64, 128) and _dense_ok(A))


def _rowmm(A, Wt, *, bias=None, act=_lib.ACT_NONE, slope=0.0, aux=None, mul_act_grad=False, accumulate=False,
           out=None, row_scale=None):
    """out (+)= epilogue(A @ Wt.T); Wt is [N,K] (nn.Linear layout).  One launch on the tensor-core path."""
    if _use_tc(A, Wt) and (out is None or _dense_ok(out)):
        return gemm_tf32x3(A, Wt.contiguous(), bias=bias, act=_ACT_NAME[act], slope=slope, aux=aux,
                           mul_act_grad=mul_act_grad, accumulate=accumulate, out=out, row_scale=row_scale)
    if row_scale is not None:
        A = A * row_scale.reshape(-1, 1)
    if accumulate:
        out.addmm_(A, Wt.t())
        return out
    if out is not None and bias is None and act == _lib.ACT_NONE and not mul_act_grad:
        return torch.mm(A, Wt.t(), out=out)
    r = torch.addmm(bias, A, Wt.t()) if bias is not None else torch.mm(A, Wt.t())
    if mul_act_grad:
        _act_backward_inplace(r, aux, act, slope)
    else:
        _act_inplace(r, act, slope)
    if out is not None:
        out.copy_(r)
        return out
    return r


def _tnmm(X, G, *, row_scale=None, colsum_x=False, colsum_g=False):
    """(row_scale ⊙ X).T @ G -- the K = rows long weight-gradient reduction; tensor cores when both widths allow.
    colsum_x / colsum_g: also return X.sum(0) / G.sum(0) (bias gradients) -> (D, sum_x, sum_g)."""
    want_sums = colsum_x or colsum_g
    if (DENSE_BACKEND == "auto" and X.shape[0] >= min(TN_MIN_ROWS, TC_MIN_ROWS) and X.shape[1] in (64, 128) and G.shape[1] in (64, 128)
            and _dense_ok(X) and _dense_ok(G)):
        return gemm_tn_tf32x3(X, G, row_scale=row_scale, colsum_x=colsum_x, colsum_g=colsum_g)
    Xs = X * row_scale.reshape(-1, 1) if row_scale is not None else X
    D = Xs.t() @ G
    if want_sums:
        return D, (X.sum(0) if colsum_x else None), (G.sum(0) if colsum_g else None)
    return D


# ---- MLP = Linear [-> BatchNorm1d] -> act -> ... -> Linear  (dmpnn.py:45-52, model.py:145-156) ---------------------
class MLPSpec:
    """Structure of one MLP (no tensors that need gradients: those travel as Function arguments).

    n_lin   number of Linear layers (0 = the layer applies `act` directly, dmpnn.py:136-138)
    bn      list of the nn.BatchNorm1d modules between the Linears (len n_lin - 1) or None
    Tensor layout in the argument list, per Linear i: W_i, b_i[, gamma_i, beta_i if bn and i < n_lin - 1]."""

    def __init__(self, n_lin, bn=None):
        self.n_lin, self.bn = n_lin, bn

    def tensors_per(self, i):
        return 4 if (self.bn is not None and i < self.n_lin - 1) else 2

    def num_tensors(self):
        return sum(self.tensors_per(i) for i in range(self.n_lin))

    def split(self, tensors):
        out, k = [], 0
        for i in range(self.n_lin):
            n = self.tensors_per(i)
            t = tuple(tensors[k:k + n]) + ((None, None) if n == 2 else ())
            out.append(t)
            k += n
        return out   # [(W, b, gamma|None, beta|None)]


def mlp_spec_and_tensors(seq):
    """(MLPSpec, [tensors]) of an nn.Sequential built like dmpnn.py:45-52 / model.py:145-156, or None if its structure
    is something else (then the layer takes the composed path)."""
    mods = list(seq)
    lin_idx = [i for i, m in enumerate(mods) if isinstance(m, torch.nn.Linear)]
    if not lin_idx:
        return (MLPSpec(0), []) if not mods else None
    bns, tensors = [], []
    for k, i in enumerate(lin_idx):
        last = k == len(lin_idx) - 1
        between = mods[i + 1:(lin_idx[k + 1] if not last else len(mods))]
        lin = mods[i]
        tensors += [lin.weight, lin.bias]
        if last:
            if between:
                return None
            continue
        bn = [m for m in between if isinstance(m, torch.nn.BatchNorm1d)]
        rest = [m for m in between if not isinstance(m, torch.nn.BatchNorm1d)]
        if len(bn) > 1 or len(rest) != 1 or (bn and between[0] is not bn[0]):
            return None
        if bn:
            b = bn[0]
            if not b.affine:
                return None
            bns.append(b)
            tensors += [b.weight, b.bias]
        else:
            bns.append(None)
    has_bn = [b is not None for b in bns]
    if any(has_bn) and not all(has_bn):
        return None
    return MLPSpec(len(lin_idx), bns if any(has_bn) else None), tensors


def _bn_forward_stats(z, bn, training):
    """(mean, invstd) used to normalise z, with nn.BatchNorm1d's running-statistics side effects."""
    use_batch = training or not bn.track_running_stats or bn.running_mean is None
    if use_batch:
        mean, var = bn_stats(z)
        if training and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                rows = z.shape[0]
                bn.running_mean.mul_(1 - m).add_(mean[:bn.num_features], alpha=m)
                unbiased = var[:bn.num_features] * (rows / max(rows - 1, 1))
                bn.running_var.mul_(1 - m).add_(unbiased, alpha=m)
    else:
        pad = z.shape[1] - bn.num_features        # zero-padded width (layers.py): padded columns stay 0
        mean, var = bn.running_mean, bn.running_var
        if pad:
            mean = torch.nn.functional.pad(mean, (0, pad))
            var = torch.nn.functional.pad(var, (0, pad), value=1.0)
    return mean, torch.rsqrt(var + bn.eps), use_batch


def _mlp_forward(pre, spec, layers, act, slope, post_act, training):
    """Returns (out, saved) with saved[i] = (z_i or None, mean, invstd, h_i, batch_stats) for every hidden Linear i."""
    x, saved = pre, []
    for i, (W, b, gamma, beta) in enumerate(layers):
        if i == spec.n_lin - 1:
            out = _rowmm(x, W, bias=b, act=post_act, slope=slope)     # bias + post-activation in the GEMM epilogue
            return out, saved
        if spec.bn is None:
            h = _rowmm(x, W, bias=b, act=act, slope=slope)            # bias + activation in the GEMM epilogue
            saved.append((None, None, None, h, False))
        else:
            z = _rowmm(x, W, bias=b)
            mean, invstd, batch_stats = _bn_forward_stats(z, spec.bn[i], training)
            h = bn_act(z, mean, invstd, gamma, beta, act, slope)
            saved.append((z, mean, invstd, h, batch_stats))
        x = h
    raise AssertionError("unreachable")


def _mlp_backward(g_out, pre, spec, layers, saved, act, slope, need_w, training):
    """Returns (g_pre, scratch buffer or None, [grads in the tensor-list order]).  Consumes the saved activations:
    g_pre is written into the storage of the first hidden activation."""
    grads = [None] * spec.num_tensors()
    pos = [0]
    for i in range(spec.n_lin):
        pos.append(pos[-1] + spec.tensors_per(i))
    g, scratch = g_out, None
    for i in range(spec.n_lin - 1, -1, -1):
        W, b, gamma, beta = layers[i]
        x_in = pre if i == 0 else saved[i - 1][3]
        if need_w:
            dW, db, _ = _tnmm(g, x_in, colsum_x=True)
            grads[pos[i]] = dW
            grads[pos[i] + 1] = db if b is not None else None
        if i == 0:
            if saved:                                   # the first hidden activation is dead: re-use its storage
                g_pre = _rowmm(g, W.t(), out=saved[0][3])
            else:
                g_pre = _rowmm(g, W.t())
            return g_pre, scratch, grads
        z, mean, invstd, h, batch_stats = saved[i - 1]
        # g_y = (g @ W) * act'(h): new edge-sized buffer, act' folded into the GEMM epilogue
        gy = _rowmm(g, W.t(), act=act, slope=slope, aux=h, mul_act_grad=True)
        if spec.bn is not None:
            _, dgamma, dbeta = bn_backward(gy, z, mean, invstd, layers[i - 1][2], batch_stats)
            grads[pos[i - 1] + 2], grads[pos[i - 1] + 3] = dgamma, dbeta
        g = scratch = gy
    raise AssertionError("unreachable")


class LayerCfg:
    """Non-tensor configuration of one fused layer call."""

    def __init__(self, order, act, slope, nmlp, emlp, post_act=_lib.ACT_NONE, training=True):
        self.order, self.act, self.slope, self.nmlp, self.emlp = order, act, float(slope), nmlp, emlp
        self.post_act, self.training = post_act, training


class _FusedDMPLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, cfg, part, X_v, X_e, norm, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nbias, ebias,
                *mlp_tensors):
        order, act, slope = cfg.order, cfg.act, cfg.slope
        nt = cfg.nmlp.num_tensors()
        nl = cfg.nmlp.split(mlp_tensors[:nt])
        el = cfg.emlp.split(mlp_tensors[nt:])
        has_mlp = cfg.nmlp.n_lin > 0
        H = nloop_w.shape[1]
        E, N = plan.E, plan.N
        norm_flat = norm_perm = None
        if norm is not None:
            norm_flat, norm_perm = plan.norm_permuted(norm)
        # destination-range partition (parallel.py): X_v holds only the owned node rows; the all-gather of the other
        # ranks' rows runs on NCCL's stream WHILE this rank aggregates its own edges (the node side needs no remote row)
        csc_indptr, X_v_full, gather_work = plan.csc_indptr, X_v, None
        if part is not None:
            _lib.sm_reserve(0)     # (a previous call that raised between reserve and release must not leak its reserve)
            from .parallel import all_gather_rows_async
            n_lo, n_hi, group = part
            X_v_full, gather_work = all_gather_rows_async(X_v, group)
            csc_indptr = plan.csc_indptr[n_lo:n_hi + 1]
            _lib.sm_reserve(_lib.SM_RESERVE)     # persistent kernels leave SMs to the collective while it is in flight

        # ---- node side (dmpnn.py:113-133): project, aggregate incident edge messages, self loop, bias
        in_t, out_t = in_w.t(), out_w.t()
        Din = in_w.shape[0]
        # Tensor-core path: AGGREGATE FIRST.  sum_e s_e n_e (X_e W) = (sum_e s_e n_e X_e) W, so one pass over X_e
        # produces the forward-edge and reversed-edge sums per destination ([N, 2 Din]) and the projections become
        # node-sized; the edge-sized product X_e W_in|out (41 GB of traffic at config 5) is never formed, and backward
        # gets dW_in = A_fwd^T gN, dW_out = A_rev^T gN from the saved sums.  Same terms, different association: within
        # the fp32 tolerance of the oracle, not bit-identical to the per-edge order (which the cuBLAS path keeps).
        agg_first = _use_tc(X_e, in_t) and Din in (64, 128)
        A2 = None
        m_off = 0 if plan.rev_layout in ("none", "halves") else H   # column offset of the reversed branch in [E, 2H]
        if agg_first:
            split = plan.rev is not None
            A2 = segment_reduce(csc_indptr, plan.csc_eid, X_e, Din, plan=plan, w_perm=norm_perm,
                                mode=_lib.SEG_SIGN_BY_REV | (_lib.SEG_SPLIT_BY_REV if split else 0),
                                tag="segment_reduce.node_fwd")
            node_pre = _rowmm(X_v, nloop_w.t(), bias=nbias)
            _rowmm(A2[:, :Din], in_t, out=node_pre, accumulate=True)
            if split:
                _rowmm(A2[:, Din:], out_t, out=node_pre, accumulate=True)
        else:
            Ln = _rowmm(X_v, nloop_w.t())
            if plan.rev_layout == "none":
                M = _rowmm(X_e, in_t)
            elif plan.rev_layout == "halves":
                h = plan.rev_split
                M = torch.empty((E, H), dtype=X_e.dtype, device=X_e.device)
                _rowmm(X_e[:h], in_t, out=M[:h])
                _rowmm(X_e[h:], out_t, out=M[h:])
            else:
                # two-branch buffer [E, 2H]: both projections, the kernel picks the half by the edge's flag
                M = torch.empty((E, 2 * H), dtype=X_e.dtype, device=X_e.device)
                _rowmm(X_e, in_t, out=M[:, :H])
                _rowmm(X_e, out_t, out=M[:, H:])
            node_pre = segment_reduce(csc_indptr, plan.csc_eid, M, H, w_perm=norm_perm, rev_col_offset=m_off,
                                      base=Ln, bias=nbias, mode=_lib.SEG_SIGN_BY_REV, out=Ln,
                                      tag="segment_reduce.node_fwd")
            del M

        # ---- edge side (dmpnn.py:112-123,142-149): endpoint gather, degree term, self loop, bias
        w_sd = src_w - dst_w
        dual = DUAL_GEMM and _use_tc(X_e, eloop_w.t())
        if dual and order == _lib.ORDER_SCM:
            # one pass over X_e: both projections in tensor memory, U = S + coef*P rounded like `eloop + add` (dmpnn.py:147)
            S = gemm_tf32x3_dual(X_e, eloop_w.t().contiguous(), w_sd.t().contiguous(), row_scale=plan.coef, mode="store")
            P = None
        elif dual:
            # UNC association ((eloop + agg) + add, model.py:257) needs both terms: still one pass over X_e
            S, P = gemm_tf32x3_dual(X_e, eloop_w.t().contiguous(), w_sd.t().contiguous(), mode="separate")
        else:
            P = _rowmm(X_e, w_sd.t())
            S = _rowmm(X_e, eloop_w.t())
        if gather_work is not None:
            gather_work.wait()
            _lib.sm_reserve(0)
        if DUAL_GEMM and _use_tc(X_v_full, dst_w.t()):
            # both endpoint tables in ONE pass over X_v (same operand, two weights)
            Qd, Qs = gemm_tf32x3_dual(X_v_full, dst_w.t().contiguous(), src_w.t().contiguous(), mode="separate")
        else:
            Qd = _rowmm(X_v_full, dst_w.t())
            Qs = _rowmm(X_v_full, src_w.t())
        edge_pre = edge_update(plan, S, P, Qd, Qs, ebias, order, out=S)
        del P, Qd, Qs

        nsaved = esaved = ()
        if has_mlp:
            node_out, nsaved = _mlp_forward(node_pre, cfg.nmlp, nl, act, slope, cfg.post_act, cfg.training)
            edge_out, esaved = _mlp_forward(edge_pre, cfg.emlp, el, act, slope, cfg.post_act, cfg.training)
        else:
            # act(pre) in place: backward only needs the output (act' from output)
            node_out, edge_out = _act_inplace(node_pre, act, slope), _act_inplace(edge_pre, act, slope)
            node_pre = edge_pre = None
        ctx.plan, ctx.cfg, ctx.norm_flat, ctx.m_off, ctx.part = plan, cfg, norm_flat, m_off, part
        ctx.norm_perm = norm_perm
        ctx.X_v_full = X_v_full if part is not None else None
        ctx.nsaved, ctx.esaved = nsaved, esaved          # intermediates (never returned): plain references
        ctx.node_pre, ctx.edge_pre, ctx.A2 = node_pre, edge_pre, A2
        need_out = (not has_mlp) or cfg.post_act != _lib.ACT_NONE
        ctx.save_for_backward(X_v, X_e, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nbias, ebias,
                              node_out if need_out else None, edge_out if need_out else None, *mlp_tensors)
        return node_out, edge_out

    @staticmethod
    def backward(ctx, g_node_out, g_edge_out):
        (X_v, X_e, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nbias, ebias, node_act, edge_act,
         *mlp_tensors) = ctx.saved_tensors
        plan, cfg = ctx.plan, ctx.cfg
        if getattr(ctx, "consumed", False):
            raise RuntimeError("the fused DMPNN layer re-uses its saved buffers in backward and cannot be "
                               "back-propagated twice (retain_graph); set layer.fused = False for that")
        ctx.consumed = True
        order, act, slope = cfg.order, cfg.act, cfg.slope
        nt = cfg.nmlp.num_tensors()
        nl = cfg.nmlp.split(mlp_tensors[:nt])
        el = cfg.emlp.split(mlp_tensors[nt:])
        has_mlp = cfg.nmlp.n_lin > 0
        node_pre, edge_pre, A2 = ctx.node_pre, ctx.edge_pre, ctx.A2
        nsaved, esaved = ctx.nsaved, ctx.esaved
        ctx.node_pre = ctx.edge_pre = ctx.A2 = ctx.nsaved = ctx.esaved = None
        H = nloop_w.shape[1]
        E = plan.E
        need = ctx.needs_input_grad
        need_xv, need_xe = need[3], need[4]
        need_w = any(need[6:])
        g_node_out = g_node_out.contiguous()
        g_edge_out = g_edge_out.contiguous()
        nmlp_grads = emlp_grads = []

        # ---- through the MLP / activation: gN = dL/dnode_pre, gE = dL/dedge_pre -------------------------
        if has_mlp:
            if cfg.post_act != _lib.ACT_NONE:     # UNC post-activation (model.py:247-248,262-263), derivative from the output
                g_node_out = _act_backward_inplace(g_node_out.clone(), node_act, cfg.post_act, slope)
                g_edge_out = _act_backward_inplace(g_edge_out.clone(), edge_act, cfg.post_act, slope)
            gN, _, nmlp_grads = _mlp_backward(g_node_out, node_pre, cfg.nmlp, nl, nsaved, act, slope, need_w, cfg.training)
            gE, buf, emlp_grads = _mlp_backward(g_edge_out, edge_pre, cfg.emlp, el, esaved, act, slope, need_w, cfg.training)
        else:
            gN = _act_backward_inplace(g_node_out.clone(), node_act, act, slope)
            gE = _act_backward_inplace(g_edge_out.clone(), edge_act, act, slope)
            buf = None
        del node_pre, edge_pre, nsaved, esaved
        part = ctx.part
        X_v_full = X_v
        n_lo = 0
        if part is not None:
            n_lo, n_hi, group = part
            X_v_full = ctx.X_v_full

        # ---- sparse core backward: two sorted-segment sums of gE, one gather of gN ------
        short = _lib.SEG_SHORT if E < 6 * plan.N else 0     # few rows per segment (partitioned graph): high-occupancy variant
        dQd = segment_reduce(plan.a_indptr, plan.a_eid, gE, H, plan=plan, mode=short, tag="segment_reduce.dQd_bwd")
        dQs = segment_reduce(plan.b_indptr, plan.b_eid, gE, H, plan=plan, mode=_lib.SEG_NEGATE_OUT | short,
                             tag="segment_reduce.dQs_bwd")
        w_sd = src_w - dst_w
        Din = in_w.shape[0]
        # Gradient of the node aggregation w.r.t. the edge side.  On the tensor-core path nothing edge-sized is
        # materialised for it: dX_e receives  sgn*norm*(gN W_n^T)[dst]  from node-sized tables and dW_in / dW_out
        # come from the aggregate-first sums A2 saved by forward (a node-sized reduction each).  Otherwise
        # T = sgn*norm*gN[dst] is written out and fed to plain GEMMs.
        gather = (need_xe or need_w) and _use_tc(gE, w_sd) and Din in (64, 128)
        T = None
        if not gather:
            m_cols = H + ctx.m_off
            if buf is not None and ctx.m_off == 0 and buf.shape[1] == H:
                T = buf
            else:
                T = torch.zeros((E, m_cols), dtype=gE.dtype, device=gE.device) if ctx.m_off else \
                    torch.empty((E, H), dtype=gE.dtype, device=gE.device)
            # every local edge's destination is owned: index the owned slice of gN by (dst - n_lo)
            edge_backward(plan, ctx.norm_flat, gN, gE, want_CG=False, t_rev_col_offset=ctx.m_off, T=T, row_offset=n_lo)
        del buf

        # ---- dense backward --------------------------------------------------------------------------------
        dX_v = dX_e = None
        scatter_work = None
        if need_xv:
            if part is None:
                dX_v = _rowmm(gN, nloop_w)
                _rowmm(dQd, dst_w, out=dX_v, accumulate=True)
                _rowmm(dQs, src_w, out=dX_v, accumulate=True)
            else:
                # partial sums over this rank's edges for EVERY node -> owners: the reduce-scatter runs on NCCL's stream
                # while this rank computes dX_e and the weight gradients below
                from .parallel import reduce_scatter_rows_async
                partial = _rowmm(dQd, dst_w)
                _rowmm(dQs, src_w, out=partial, accumulate=True)
                dX_v, scatter_work = reduce_scatter_rows_async(partial, group)
                _lib.sm_reserve(_lib.SM_RESERVE)
        if need_xe:
            if gather:
                # 1. dX_e <- sgn*norm*(gN W_n^T)[dst]: streaming gather from node-sized tables (high-occupancy kernel:
                #    a GEMM epilogue cannot keep enough random 128-byte loads in flight, measured 37 ms vs 7 ms here)
                # 2. both projections of gE accumulate onto it in ONE pass (coef scales the second product's rows)
                if plan.rev is not None and DUAL_GEMM and _use_tc(gN, in_w):
                    tab_in, tab_out = gemm_tf32x3_dual(gN, in_w, out_w, mode="separate")    # one pass over gN
                else:
                    tab_in = _rowmm(gN, in_w)
                    tab_out = _rowmm(gN, out_w) if plan.rev is not None else None
                dX_e = torch.empty((E, Din), dtype=gE.dtype, device=gE.device)
                edge_backward(plan, ctx.norm_flat, tab_in, None, want_CG=False, T=dX_e, gN_rev=tab_out, row_offset=n_lo)
                del tab_in, tab_out
                if DUAL_GEMM and _use_tc(gE, eloop_w):
                    gemm_tf32x3_dual(gE, eloop_w, w_sd, row_scale=plan.coef, mode="accumulate", out=dX_e)
                else:
                    _rowmm(gE, eloop_w, out=dX_e, accumulate=True)
                    _rowmm(gE, w_sd, out=dX_e, accumulate=True, row_scale=plan.coef)
            else:
                dX_e = _rowmm(gE, eloop_w)
                _rowmm(gE, w_sd, out=dX_e, accumulate=True, row_scale=plan.coef)
                if plan.rev_layout == "none":
                    _rowmm(T, in_w, out=dX_e, accumulate=True)
                elif plan.rev_layout == "halves":
                    h = plan.rev_split
                    _rowmm(T[:h], in_w, out=dX_e[:h], accumulate=True)
                    _rowmm(T[h:], out_w, out=dX_e[h:], accumulate=True)
                else:
                    _rowmm(T[:, :H], in_w, out=dX_e, accumulate=True)
                    _rowmm(T[:, H:], out_w, out=dX_e, accumulate=True)
        d_in = d_out = d_src = d_dst = d_nloop = d_eloop = d_nb = d_eb = None
        if need_w:
            d_nloop, _, d_nb = _tnmm(X_v, gN, colsum_g=True)
            d_eloop, _, d_eb = _tnmm(X_e, gE, colsum_g=True)
            if nbias is None:
                d_nb = None
            if ebias is None:
                d_eb = None
            d_sd = _tnmm(gE, X_e, row_scale=plan.coef).t()   # ((coef ⊙ gE)^T X_e)^T: the scale rides on gE, as in autograd
            d_dst = _tnmm(X_v_full, dQd)
            d_dst.sub_(d_sd)
            d_src = _tnmm(X_v_full, dQs)
            d_src.add_(d_sd)
            if gather:
                if A2 is None:   # forward did not run aggregate-first: rebuild the per-destination sums of X_e
                    seg_ptr = plan.csc_indptr if part is None else plan.csc_indptr[part[0]:part[1] + 1]
                    A2 = segment_reduce(seg_ptr, plan.csc_eid, X_e, Din, w_perm=ctx.norm_perm,
                                        mode=_lib.SEG_SIGN_BY_REV | (_lib.SEG_SPLIT_BY_REV if plan.rev is not None else 0),
                                        tag="segment_reduce.dW_bwd")
                d_in = _tnmm(A2[:, :Din], gN)
                d_out = _tnmm(A2[:, Din:], gN) if plan.rev is not None else torch.zeros_like(out_w)
            elif plan.rev_layout == "none":
                d_in = _tnmm(X_e, T)
                d_out = torch.zeros_like(out_w)
            elif plan.rev_layout == "halves":
                h = plan.rev_split
                d_in = _tnmm(X_e[:h], T[:h])
                d_out = _tnmm(X_e[h:], T[h:])
            else:
                d_in = _tnmm(X_e, T[:, :H])
                d_out = _tnmm(X_e, T[:, H:])
        if scatter_work is not None:
            scatter_work.wait()
            _lib.sm_reserve(0)
            _rowmm(gN, nloop_w, out=dX_v, accumulate=True)
        if part is not None and need_w:
            # every weight gradient above is a partial sum over this rank's nodes/edges
            from .parallel import allreduce_tensors_
            allreduce_tensors_([d_in, d_out, d_src, d_dst, d_nloop, d_eloop, d_nb, d_eb] + list(nmlp_grads)
                               + list(emlp_grads), group)
        return (None, None, None, dX_v, dX_e, None, d_in, d_out, d_src, d_dst, d_nloop, d_eloop, d_nb, d_eb,
                *nmlp_grads, *emlp_grads)


def fused_dmp_layer(plan, X_v, X_e, weights, nbias, ebias, nmlp, emlp, *, act_func, slope, order, norm=None,
                    post_act="none", training=True, part=None):
    """weights = (in, out, src, dst, nloop, eloop); nmlp / emlp = (MLPSpec, [tensors]) from `mlp_spec_and_tensors`.

    Raises ValueError (like the reference's DGL frame-size check) when the feature rows do not match the graph."""
    _lib.require_cuda(X_v, X_e)
    n_rows = plan.N if part is None else part[1] - part[0]
    if X_v.dim() != 2 or X_e.dim() != 2 or X_v.shape[0] != n_rows or X_e.shape[0] != plan.E:
        raise ValueError("feature rows (%s nodes, %s edges) do not match the graph (%d, %d)"
                         % (tuple(X_v.shape), tuple(X_e.shape), n_rows, plan.E))
    Din = weights[0].shape[0]
    if X_v.shape[1] != Din or X_e.shape[1] != Din:
        raise ValueError("feature width (%d nodes, %d edges) does not match the layer's input_dim %d"
                         % (X_v.shape[1], X_e.shape[1], Din))
    if norm is not None and norm.numel() != plan.E:
        raise ValueError("edge_norm must have one entry per edge")
    cfg = LayerCfg(order, _ACT[act_func], slope, nmlp[0], emlp[0], _ACT[post_act], training)
    return _FusedDMPLayer.apply(plan, cfg, part, X_v.contiguous(), X_e.contiguous(), norm, *weights, nbias, ebias,
                                *nmlp[1], *emlp[1])
