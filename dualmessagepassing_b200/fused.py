"""Whole-layer autograd function with explicit buffer management (the path used at BASELINE scale).

`torch.autograd` over the composed ops of `layers.dual_message_passing` keeps ~14 edge-sized
temporaries alive (the reference keeps even more); at config 5
(E = 40 M, H = 128) one [E,H] fp32 tensor is 20.5 GB, so that composition does not fit 180 GB of HBM.
This function computes exactly the same quantities in the same order but owns every edge-sized
buffer: forward keeps only {X_e (input), edge_pre, h1}; backward re-uses the saved buffers in place.
Peak is 5 edge-sized tensors in forward and 7 in backward including the caller's input, output and
upstream gradient ("Memory plan").

Supported: SCM flavour or UNC flavour without BatchNorm, num_mlp_layers in {0, 2}, activation in
{none, relu, leaky_relu, tanh, sigmoid}, dropout inactive.  Anything else runs the composed path.
Reference lines: SubgraphCountingMatching/models/dmpnn.py:111-156 (forward), SURVEY.md Appendix A.2
(backward).
"""
import torch

from . import _lib
from .functional import (edge_backward, edge_update, gemm_tf32x3, gemm_tf32x3_acc_gather, gemm_tn_tf32x3,
                         segment_reduce)

_ACT = {"none": _lib.ACT_NONE, "relu": _lib.ACT_RELU, "leaky_relu": _lib.ACT_LEAKY_RELU,
        "tanh": _lib.ACT_TANH, "sigmoid": _lib.ACT_SIGMOID}


def supported_activation(name):
    return name in _ACT


def _act_inplace(x, act, slope):
    """x <- act(x) with the sm_100a elementwise kernel (in place)."""
    if act == _lib.ACT_NONE or x.numel() == 0:
        return x
    rows, H = x.shape
    _lib.call("dmp_gate_residual", x.device, _lib.ptr(x), H, None, None, 0, _lib.ptr(x), H, rows, H, act,
              slope, _lib.stream_ptr(x.device), tag="act_inplace")
    return x


def _act_backward_inplace(g, y, act, slope):
    """g <- g * act'(.) expressed through the activation OUTPUT y (in place on g)."""
    if act == _lib.ACT_NONE or g.numel() == 0:
        return g
    rows, H = g.shape
    _lib.call("dmp_gate_residual_backward", g.device,
              _lib.ptr(g), H, _lib.ptr(y), H, None, _lib.ptr(g), H, rows, H, act | _lib.ACT_FROM_OUTPUT, slope,
              _lib.stream_ptr(g.device), tag="act_bwd_inplace")
    return g


# Dense backend of the edge-/node-sized projections:
#   "auto"    tcgen05 3xTF32 kernel (dmp_gemm_tf32x3, fp32-level accuracy) when N, K in {64,128}, else cuBLAS
#   "cublas"  always torch.mm (cuBLAS sgemm)
DENSE_BACKEND = "auto"
_ACT_NAME = {v: k for k, v in _ACT.items()}


# below this many rows the persistent tcgen05 kernels' fixed cost (weight split per CTA, TMEM allocation, second
# reduce launch) exceeds what cuBLAS sgemm needs for the whole product
TC_MIN_ROWS = 16384


def _use_tc(A, Wt):
    return (DENSE_BACKEND == "auto" and A.dtype == torch.float32 and A.shape[0] >= TC_MIN_ROWS
            and Wt.shape[0] in (64, 128) and Wt.shape[1] in (64, 128)
            and A.stride(1) == 1 and A.stride(0) % 4 == 0 and A.data_ptr() % 16 == 0)


def _rowmm(A, Wt, *, bias=None, act=_lib.ACT_NONE, slope=0.0, aux=None, mul_act_grad=False, accumulate=False,
           out=None, row_scale=None):
    """out (+)= epilogue(A @ Wt.T); Wt is [N,K] (nn.Linear layout).  One launch on the tensor-core path."""
    if _use_tc(A, Wt) and (out is None or (out.stride(1) == 1 and out.stride(0) % 4 == 0 and out.data_ptr() % 16 == 0)):
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
    if (DENSE_BACKEND == "auto" and X.shape[0] >= TC_MIN_ROWS and X.shape[1] in (64, 128) and G.shape[1] in (64, 128)
            and X.stride(1) == 1 and G.stride(1) == 1 and X.stride(0) % 4 == 0 and G.stride(0) % 4 == 0
            and X.data_ptr() % 16 == 0 and G.data_ptr() % 16 == 0):
        return gemm_tn_tf32x3(X, G, row_scale=row_scale, colsum_x=colsum_x, colsum_g=colsum_g)
    Xs = X * row_scale.reshape(-1, 1) if row_scale is not None else X
    D = Xs.t() @ G
    if want_sums:
        return D, (X.sum(0) if colsum_x else None), (G.sum(0) if colsum_g else None)
    return D


def _mlp_forward(pre, W1, b1, W2, b2, act, slope):
    """Linear -> act -> Linear (dmpnn.py:45-52 without BN). Returns (out, h1) with h1 = act(lin1)."""
    h1 = _rowmm(pre, W1, bias=b1, act=act, slope=slope)      # bias + activation in the GEMM epilogue
    out = _rowmm(h1, W2, bias=b2)
    return out, h1


def _mlp_backward(g_out, pre, h1, W1, W2, act, slope, need_w):
    """Returns (g_pre written into h1's storage, dW1, db1, dW2, db2). Consumes h1."""
    dW2 = db2 = dW1 = db1 = None
    if need_w:
        dW2, db2, _ = _tnmm(g_out, h1, colsum_x=True)
    # g1 = (g_out @ W2) * act'(h1): new edge-sized buffer, act' folded into the GEMM epilogue
    g1 = _rowmm(g_out, W2.t(), act=act, slope=slope, aux=h1, mul_act_grad=True)
    if need_w:
        dW1, db1, _ = _tnmm(g1, pre, colsum_x=True)
    g_pre = _rowmm(g1, W1.t(), out=h1)   # h1 is dead: re-use its storage
    return g_pre, g1, dW1, db1, dW2, db2


class _FusedDMPLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, cfg, X_v, X_e, norm, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nbias, ebias,
                nW1, nb1, nW2, nb2, eW1, eb1, eW2, eb2, part=None):
        order, act, slope, has_mlp = cfg
        H = nloop_w.shape[1]
        E, N = plan.E, plan.N
        norm_flat = norm_perm = None
        if norm is not None:
            norm_flat, norm_perm = plan.norm_permuted(norm)
        # destination-range partition (parallel.py): X_v holds only the owned node rows; gather the rest
        csc_indptr, X_v_full = plan.csc_indptr, X_v
        if part is not None:
            from .parallel import all_gather_rows
            n_lo, n_hi, group = part
            X_v_full = all_gather_rows(X_v, group)
            csc_indptr = plan.csc_indptr[n_lo:n_hi + 1]

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
            A2 = segment_reduce(csc_indptr, plan.csc_eid, X_e, Din, w_perm=norm_perm,
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
        Qd = _rowmm(X_v_full, dst_w.t())
        Qs = _rowmm(X_v_full, src_w.t())
        P = _rowmm(X_e, (src_w - dst_w).t())
        S = _rowmm(X_e, eloop_w.t())
        edge_pre = edge_update(plan, S, P, Qd, Qs, ebias, order, out=S)
        del P, Qd, Qs

        if has_mlp:
            node_out, nh1 = _mlp_forward(node_pre, nW1, nb1, nW2, nb2, act, slope)
            edge_out, eh1 = _mlp_forward(edge_pre, eW1, eb1, eW2, eb2, act, slope)
        else:
            # act(pre) in place: backward only needs the output (act' from output)
            node_out, nh1 = _act_inplace(node_pre, act, slope), None
            edge_out, eh1 = _act_inplace(edge_pre, act, slope), None
            node_pre = edge_pre = None
        ctx.plan, ctx.cfg, ctx.norm_flat, ctx.m_off, ctx.part = plan, cfg, norm_flat, m_off, part
        ctx.norm_perm = norm_perm
        ctx.X_v_full = X_v_full if part is not None else None
        ctx.save_for_backward(X_v, X_e, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nW1, nW2, eW1, eW2,
                              node_pre, nh1, edge_pre, eh1,
                              node_out if not has_mlp else None, edge_out if not has_mlp else None, A2)
        ctx.has_bias = (nbias is not None, ebias is not None)
        ctx.mlp_bias = (nb1 is not None, nb2 is not None, eb1 is not None, eb2 is not None)
        return node_out, edge_out

    @staticmethod
    def backward(ctx, g_node_out, g_edge_out):
        (X_v, X_e, in_w, out_w, src_w, dst_w, nloop_w, eloop_w, nW1, nW2, eW1, eW2,
         node_pre, nh1, edge_pre, eh1, node_act, edge_act, A2) = ctx.saved_tensors
        plan = ctx.plan
        if getattr(ctx, "consumed", False):
            raise RuntimeError("the fused DMPNN layer re-uses its saved buffers in backward and cannot be "
                               "back-propagated twice (retain_graph); set layer.fused = False for that")
        ctx.consumed = True
        order, act, slope, has_mlp = ctx.cfg
        H = nloop_w.shape[1]
        E = plan.E
        need = ctx.needs_input_grad
        need_xv, need_xe = need[2], need[3]
        need_w = any(need[5:])
        g_node_out = g_node_out.contiguous()
        g_edge_out = g_edge_out.contiguous()
        dnW1 = dnb1 = dnW2 = dnb2 = deW1 = deb1 = deW2 = deb2 = None

        # ---- through the MLP / activation: gN = dL/dnode_pre, gE = dL/dedge_pre -------------------------
        if has_mlp:
            gN, _, dnW1, dnb1, dnW2, dnb2 = _mlp_backward(g_node_out, node_pre, nh1, nW1, nW2, act, slope, need_w)
            gE, buf, deW1, deb1, deW2, deb2 = _mlp_backward(g_edge_out, edge_pre, eh1, eW1, eW2, act, slope, need_w)
            buf2 = edge_pre  # dead after dW1: second scratch buffer
        else:
            gN = _act_backward_inplace(g_node_out.clone(), node_act, act, slope)
            gE = _act_backward_inplace(g_edge_out.clone(), edge_act, act, slope)
            buf = buf2 = None
        del node_pre, nh1, edge_pre, eh1
        part = ctx.part
        X_v_full, gN_full = X_v, gN
        if part is not None:
            from .parallel import allreduce_tensors_, reduce_scatter_rows
            n_lo, n_hi, group = part
            X_v_full = ctx.X_v_full
            gN_full = torch.zeros((plan.N, H), dtype=gN.dtype, device=gN.device)
            gN_full[n_lo:n_hi] = gN  # edge_backward indexes gN by GLOBAL destination id

        # ---- sparse core backward: two sorted-segment sums of gE, one gather of gN ------
        dQd = segment_reduce(plan.a_indptr, plan.a_eid, gE, H, tag="segment_reduce.dQd_bwd")
        dQs = segment_reduce(plan.b_indptr, plan.b_eid, gE, H, mode=_lib.SEG_NEGATE_OUT,
                             tag="segment_reduce.dQs_bwd")
        w_sd = src_w - dst_w
        Din = in_w.shape[0]
        # Gradient of the node aggregation w.r.t. the edge side.  On the tensor-core path nothing edge-sized is
        # materialised for it: dX_e receives  sgn*norm*(gN W_n^T)[dst]  from node-sized tables inside the GEMM epilogue
        # and dW_in / dW_out come from the aggregate-first sums A2 saved by forward (a node-sized reduction each).  Otherwise T = sgn*norm*gN[dst] is written out and fed to plain GEMMs.
        gather = (need_xe or need_w) and _use_tc(gE, w_sd) and Din in (64, 128)
        T = None
        if not gather:
            m_cols = H + ctx.m_off
            if buf is not None and ctx.m_off == 0:
                T = buf
            else:
                T = torch.zeros((E, m_cols), dtype=gE.dtype, device=gE.device) if ctx.m_off else \
                    torch.empty((E, H), dtype=gE.dtype, device=gE.device)
            edge_backward(plan, ctx.norm_flat, gN_full, gE, want_CG=False, t_rev_col_offset=ctx.m_off, T=T)
        del buf, buf2

        # ---- dense backward --------------------------------------------------------------------------------
        dX_v = dX_e = None
        if need_xv:
            if part is None:
                dX_v = _rowmm(gN, nloop_w)
                _rowmm(dQd, dst_w, out=dX_v, accumulate=True)
                _rowmm(dQs, src_w, out=dX_v, accumulate=True)
            else:
                # partial sums over this rank's edges for EVERY node -> owners (reduce-scatter over NVLink)
                partial = _rowmm(dQd, dst_w)
                _rowmm(dQs, src_w, out=partial, accumulate=True)
                dX_v = reduce_scatter_rows(partial, group)
                _rowmm(gN, nloop_w, out=dX_v, accumulate=True)
                del partial
        if need_xe:
            if gather:
                # 1. dX_e <- sgn*norm*(gN W_n^T)[dst]: streaming gather from node-sized tables (high-occupancy kernel:
                #    a GEMM epilogue cannot keep enough random 128-byte loads in flight, measured 37 ms vs 7 ms here)
                # 2./3. both projections accumulate onto it; coef ⊙ gE is a per-row scale of the streamed operand
                tab_in = _rowmm(gN_full, in_w)
                tab_out = _rowmm(gN_full, out_w) if plan.rev is not None else None
                dX_e = torch.empty((E, Din), dtype=gE.dtype, device=gE.device)
                edge_backward(plan, ctx.norm_flat, tab_in, None, want_CG=False, T=dX_e, gN_rev=tab_out)
                del tab_in, tab_out
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
            if not ctx.has_bias[0]:
                d_nb = None
            if not ctx.has_bias[1]:
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
        mb = ctx.mlp_bias
        if part is not None and need_w:
            # every weight gradient above is a partial sum over this rank's nodes/edges
            allreduce_tensors_([d_in, d_out, d_src, d_dst, d_nloop, d_eloop, d_nb, d_eb, dnW1, dnb1, dnW2, dnb2,
                                deW1, deb1, deW2, deb2], group)
        return (None, None, dX_v, dX_e, None, d_in, d_out, d_src, d_dst, d_nloop, d_eloop, d_nb, d_eb,
                dnW1, dnb1 if mb[0] else None, dnW2, dnb2 if mb[1] else None,
                deW1, deb1 if mb[2] else None, deW2, deb2 if mb[3] else None, None)


def fused_dmp_layer(plan, X_v, X_e, weights, nbias, ebias, nmlp, emlp, *, act_func, slope, order, norm=None):
    """weights = (in, out, src, dst, nloop, eloop); nmlp/emlp = (W1, b1, W2, b2) or None."""
    _lib.require_cuda(X_v, X_e)
    has_mlp = nmlp is not None
    cfg = (order, _ACT[act_func], float(slope), has_mlp)
    n = nmlp if has_mlp else (None,) * 4
    e = emlp if has_mlp else (None,) * 4
    return _FusedDMPLayer.apply(plan, cfg, X_v.contiguous(), X_e.contiguous(), norm, *weights, nbias, ebias,
                                *n, *e)
