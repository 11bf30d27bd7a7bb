"""Autograd functions over the C ABI (include/dmp_b200.h): the DMPNN sparse core.

`sparse_core` is the differentiable unit that replaces the DGL `update_all` + `apply_edges` pair of
SubgraphCountingMatching/models/dmpnn.py:158-166 once the dense projections are done:

    node_pre = L_n + segsum_dst(sgn * M * norm) + nbias          (dmpnn.py:113-124,92,131-133)
    edge_pre = S + coef * P + (Q_d[a] - Q_s[b]) + ebias          (dmpnn.py:112-123,142-149)

Backward is the deterministic sorted-segment form: two segment sums of gE
(keys a and b) and one gather of gN -- no float atomics anywhere.
"""
import torch

from . import _lib


def _stream(t):
    return _lib.stream_ptr(t.device)


def segment_reduce(indptr, eid, V, H, *, w_perm=None, rev_col_offset=0, base=None, bias=None, mode=0, out=None,
                   tag=None):
    """Raw (non-differentiable) call of dmp_segment_reduce. V: [*, ldV] fp32, returns [nseg, H]
    ([nseg, 2H] = [forward-edge sums | reversed-edge sums] with SEG_SPLIT_BY_REV)."""
    _lib.require_cuda(indptr, eid, V, w_perm, base, bias)
    V, ldV = _lib.row_major(V)
    nseg = indptr.numel() - 1
    if out is None:
        out = torch.empty((nseg, 2 * H if mode & _lib.SEG_SPLIT_BY_REV else H), dtype=torch.float32, device=V.device)
    ld_base = 0
    if base is not None:
        base, ld_base = _lib.row_major(base)
    if bias is not None:
        bias = bias.contiguous()
    out_m, ld_out = _lib.row_major(out)
    assert out_m.data_ptr() == out.data_ptr(), "out must have dense rows"
    _lib.call("dmp_segment_reduce", V.device,
              _lib.ptr(indptr), _lib.ptr(eid), _lib.ptr(w_perm), _lib.ptr(V), ldV, rev_col_offset,
              _lib.ptr(base), ld_base, _lib.ptr(bias), _lib.ptr(out), ld_out, nseg, H, mode, _stream(V),
              tag=tag or "segment_reduce")
    return out


def segment_reduce_two_level(indptr, eid, V, H, *, chunk=1024, w_perm=None, rev_col_offset=0, base=None, bias=None, mode=0,
                             out=None, tag=None):
    """Same reduction for graphs with very long segments (hubs: Yelp has 30.5 M links on 82 k nodes): every segment is
    cut into chunks of `chunk` positions, the chunks are reduced in parallel (sequentially inside a chunk, ascending
    position) and the chunk partials of a segment are then added in ascending chunk order.  Deterministic and run-to-run
    bit-stable, but NOT the strictly sequential order of `segment_reduce` for segments longer than `chunk` (fp32 addition
    is not associative): use it where the reference's own order is not sequential either (pooling by `sum(dim=0)`,
    model.py:319-325) or as an explicit throughput option; shorter segments are bit-identical.  No host synchronisation:
    the chunk table is sized by its upper bound nseg + E/chunk."""
    nseg = indptr.numel() - 1
    E = eid.numel()
    t_max = nseg + E // chunk + 1
    ip = indptr.to(torch.int64)
    nch = (ip[1:] - ip[:-1] + (chunk - 1)) // chunk                       # chunks per segment (0 for empty segments)
    ch_off = torch.zeros(nseg + 1, dtype=torch.int64, device=indptr.device)
    ch_off[1:] = torch.cumsum(nch, 0)
    t = torch.arange(t_max + 1, device=indptr.device)
    owner = torch.searchsorted(ch_off, t, right=True) - 1                  # segment of chunk t (nseg for the padding)
    inside = owner < nseg
    own = owner.clamp(max=max(nseg - 1, 0))
    start = ip[own] + (t - ch_off[own]) * chunk
    fine = torch.where(inside, start, torch.full_like(start, E)).to(torch.int32)   # [t_max + 1]: refined indptr
    split = bool(mode & _lib.SEG_SPLIT_BY_REV)
    partial = segment_reduce(fine, eid, V, H, w_perm=w_perm, rev_col_offset=rev_col_offset, mode=mode & ~_lib.SEG_NEGATE_OUT,
                             tag=(tag or "segment_reduce") + ".chunks")
    ident = torch.arange(t_max, dtype=torch.int32, device=indptr.device)
    return segment_reduce(ch_off.to(torch.int32), ident, partial, 2 * H if split else H, base=base, bias=bias,
                          mode=mode & _lib.SEG_NEGATE_OUT, out=out, tag=(tag or "segment_reduce") + ".combine")


def edge_update(plan, S, P, Qd, Qs, ebias, order, out=None, edge_agg=None):
    """Raw call of dmp_edge_update; `out` may be S itself (in place)."""
    if S.shape[0] != plan.E or (P is not None and P.shape[0] != plan.E):
        raise ValueError("edge_update: operands have %d rows, the graph has %d edges" % (S.shape[0], plan.E))
    _lib.require_cuda(S, P, Qd, Qs, ebias)
    S, ldS = _lib.row_major(S)
    ldP = 0
    if P is not None:   # None: S already holds eloop + coef*P (gemm_tf32x3_dual, DUAL_STORE)
        P, ldP = _lib.row_major(P)
    Qd, ldQd = _lib.row_major(Qd)
    Qs, ldQs = _lib.row_major(Qs)
    E, H = S.shape
    if out is None:
        out = torch.empty((E, H), dtype=torch.float32, device=S.device)
    _, ld_out = _lib.row_major(out)
    ld_agg = 0
    if edge_agg is not None:
        _, ld_agg = _lib.row_major(edge_agg)
    if ebias is not None:
        ebias = ebias.contiguous()
    if plan.mirrored_halves:
        order |= _lib.EDGE_MIRRORED_HALVES
    _lib.call("dmp_edge_update", S.device,
              _lib.ptr(plan.a32), _lib.ptr(plan.b32), _lib.ptr(plan.coef), _lib.ptr(S), ldS, _lib.ptr(P), ldP,
              _lib.ptr(Qd), ldQd, _lib.ptr(Qs), ldQs, _lib.ptr(ebias), _lib.ptr(out), ld_out,
              _lib.ptr(edge_agg), ld_agg, E, H, order, _stream(S), tag="edge_update")
    return out


def edge_backward(plan, norm_flat, gN, gE, *, want_T=True, want_CG=True, t_rev_col_offset=0, T=None, CG=None,
                  gN_rev=None, row_offset=0):
    """Raw call of dmp_edge_backward: T = sgn*gN[dst]*norm (optionally into the rev half), CG = coef*gE.
    row_offset: gN / gN_rev hold rows [row_offset, row_offset + gN.shape[0]) of the table the plan's destination ids
    index (destination-range partition: every local edge's destination is owned, so only the owned slice exists)."""
    H = gN.shape[1] if gN is not None else gE.shape[1]
    dev = gN.device if gN is not None else gE.device
    ld_gN = ld_gE = ldT = ldCG = 0
    if want_T:
        gN, ld_gN = _lib.row_major(gN)
        if gN_rev is not None:
            gN_rev, ld2 = _lib.row_major(gN_rev)
            if ld2 != ld_gN:
                raise ValueError("gN and gN_rev must share a leading dimension")
        if T is None:
            if t_rev_col_offset:
                T = torch.zeros((plan.E, t_rev_col_offset + H), dtype=torch.float32, device=dev)
            else:
                T = torch.empty((plan.E, H), dtype=torch.float32, device=dev)
        _, ldT = _lib.row_major(T)
    if want_CG:
        gE, ld_gE = _lib.row_major(gE)
        if CG is None:
            CG = torch.empty((plan.E, H), dtype=torch.float32, device=dev)
        _, ldCG = _lib.row_major(CG)
    _lib.call("dmp_edge_backward", dev,
              _lib.ptr(plan.dst32), _lib.ptr(plan.rev), _lib.ptr(norm_flat), _lib.ptr(plan.coef),
              _shift(gN if want_T else None, row_offset, ld_gN), _shift(gN_rev if want_T else None, row_offset, ld_gN),
              ld_gN,
              _lib.ptr(gE if want_CG else None), ld_gE,
              _lib.ptr(T if want_T else None), ldT, t_rev_col_offset, _lib.ptr(CG if want_CG else None), ldCG,
              plan.E, H, _lib.stream_ptr(dev), tag="edge_backward")
    return (T if want_T else None), (CG if want_CG else None)


def _shift(t, row_offset, ld):
    """Device pointer of the (virtual) row 0 of a table whose first stored row is `row_offset`."""
    if t is None:
        return None
    return t.data_ptr() - int(row_offset) * int(ld) * 4


class _SparseCore(torch.autograd.Function):
    """node_pre, edge_pre = f(M, S, P, L_n, Q_d, Q_s, nbias, ebias)."""

    @staticmethod
    def forward(ctx, plan, norm, order, m_rev_off, M, S, P, Ln, Qd, Qs, nbias, ebias):
        H = S.shape[1]
        norm_flat = norm_perm = None
        if norm is not None:
            norm_flat, norm_perm = plan.norm_permuted(norm)
        node_pre = segment_reduce(plan.csc_indptr, plan.csc_eid, M, H, w_perm=norm_perm,
                                  rev_col_offset=m_rev_off, base=Ln, bias=nbias, mode=_lib.SEG_SIGN_BY_REV,
                                  tag="segment_reduce.node_fwd")
        edge_pre = edge_update(plan, S, P, Qd, Qs, ebias, order)
        ctx.plan, ctx.norm_flat, ctx.m_rev_off, ctx.H = plan, norm_flat, m_rev_off, H
        ctx.m_cols = M.shape[1]
        ctx.has_bias = (nbias is not None, ebias is not None)
        return node_pre, edge_pre

    @staticmethod
    def backward(ctx, gN, gE):
        plan, H = ctx.plan, ctx.H
        need = ctx.needs_input_grad  # (plan, norm, order, m_rev_off, M, S, P, Ln, Qd, Qs, nbias, ebias)
        gN = gN.contiguous()
        gE = gE.contiguous()
        dM = dP = dQd = dQs = dnb = deb = None
        if need[4] or need[6]:
            T = None
            if need[4] and ctx.m_rev_off:
                T = torch.zeros((plan.E, ctx.m_cols), dtype=torch.float32, device=gE.device)
            dM, dP = edge_backward(plan, ctx.norm_flat, gN, gE, want_T=need[4], want_CG=need[6],
                                   t_rev_col_offset=ctx.m_rev_off, T=T)
        if need[8]:
            dQd = segment_reduce(plan.a_indptr, plan.a_eid, gE, H, tag="segment_reduce.dQd_bwd")
        if need[9]:
            dQs = segment_reduce(plan.b_indptr, plan.b_eid, gE, H, mode=_lib.SEG_NEGATE_OUT,
                                 tag="segment_reduce.dQs_bwd")
        if ctx.has_bias[0] and need[10]:
            dnb = gN.sum(0)
        if ctx.has_bias[1] and need[11]:
            deb = gE.sum(0)
        return (None, None, None, None, dM, gE if need[5] else None, dP, gN if need[7] else None,
                dQd, dQs, dnb, deb)


def sparse_core(plan, M, S, P, Ln, Qd, Qs, nbias=None, ebias=None, *, norm=None, order=_lib.ORDER_SCM,
                m_rev_off=0):
    _lib.require_cuda(M, S, P, Ln, Qd, Qs)
    return _SparseCore.apply(plan, norm, order, m_rev_off, M, S, P, Ln, Qd, Qs, nbias, ebias)


_ACT_IDS = {"none": _lib.ACT_NONE, "relu": _lib.ACT_RELU, "leaky_relu": _lib.ACT_LEAKY_RELU,
            "tanh": _lib.ACT_TANH, "sigmoid": _lib.ACT_SIGMOID}


class _GateResidual(torch.autograd.Function):
    """out = prev + gate * act(x)   (row A7: dmpnn.py:236-241,266-275; act only for MLP-less layers)."""

    @staticmethod
    def forward(ctx, x, gate, prev, act, slope):
        _lib.require_cuda(x, gate, prev)
        x, ldx = _lib.row_major(x)
        rows, H = x.shape
        out = torch.empty_like(x)
        ld_prev = 0
        if prev is not None:
            prev, ld_prev = _lib.row_major(prev)
        g = None
        if gate is not None:
            g = gate.reshape(-1).contiguous().float()
            if g.numel() != rows:
                raise ValueError("gate_residual: gate has %d entries, x has %d rows" % (g.numel(), rows))
        if prev is not None and tuple(prev.shape) != (rows, H):
            raise ValueError("gate_residual: prev is %s, x is %s" % (tuple(prev.shape), (rows, H)))
        _lib.call("dmp_gate_residual", x.device, _lib.ptr(x), ldx, _lib.ptr(g), _lib.ptr(prev), ld_prev,
                  _lib.ptr(out), H, rows, H, act, slope, _stream(x), tag="gate_residual")
        ctx.save_for_backward(x if act != _lib.ACT_NONE else None, g)
        ctx.act, ctx.slope, ctx.has_prev = act, slope, prev is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        x, g = ctx.saved_tensors
        gout, ldg = _lib.row_major(gout.contiguous())
        rows, H = gout.shape
        gx = None
        if ctx.needs_input_grad[0]:
            if g is None and ctx.act == _lib.ACT_NONE:
                gx = gout
            else:
                gx = torch.empty_like(gout)
                _lib.call("dmp_gate_residual_backward", gout.device,
                          _lib.ptr(gout), ldg, _lib.ptr(x), H, _lib.ptr(g), _lib.ptr(gx), H, rows, H, ctx.act,
                          ctx.slope, _stream(gout), tag="gate_residual_bwd")
        gprev = gout if (ctx.has_prev and ctx.needs_input_grad[2]) else None
        return gx, None, gprev, None, None


def gate_residual(x, gate=None, prev=None, act="none", slope=0.0):
    """Fused `prev + gate * act(x)`; gate is [rows] or [rows,1] (0/1 mask or soft gate), no grad to it.

    Differences from the reference's torch ops, by construction: a 0 gate MULTIPLIES (0 * NaN = NaN, where
    `masked_fill` would give 0), and a gate that requires grad is refused (the reference's gates are label-match
    masks computed without grad, basemodel.py:1394-1423)."""
    if gate is not None and gate.requires_grad:
        raise NotImplementedError("gate_residual does not differentiate with respect to the gate")
    return _GateResidual.apply(x, gate, prev, _ACT_IDS[act], float(slope))


def gemm_tf32x3_supported(M, N, K):
    return N in (64, 128) and K in (64, 128)


def gemm_tf32x3(A, Wt, *, row_scale=None, bias=None, act="none", slope=0.0, aux=None, mul_act_grad=False,
                accumulate=False, out=None):
    """D = epilogue((row_scale ⊙ A) @ Wt.T) on the tcgen05 tensor cores with a 3xTF32 split (fp32-level accuracy).

    A [M,K] fp32 (dense rows), Wt [N,K] (nn.Linear layout).  Raw, non-differentiable (the fused layer calls it)."""
    _lib.require_cuda(A, Wt, row_scale, bias, aux)
    A, lda = _lib.row_major(A)
    Wt, ldb = _lib.row_major(Wt)
    M, K = A.shape
    N = Wt.shape[0]
    if Wt.shape[1] != K:
        raise ValueError("inner dimensions differ: A is %s, Wt is %s" % (tuple(A.shape), tuple(Wt.shape)))
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs `out`")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _, ldd = _lib.row_major(out)
    epi = _ACT_IDS[act]
    ld_aux = 0
    if mul_act_grad:
        aux, ld_aux = _lib.row_major(aux)
        epi |= _lib.EPI_MUL_ACT_GRAD
    if accumulate:
        epi |= _lib.EPI_ACCUMULATE
    if row_scale is not None:
        row_scale = row_scale.reshape(-1).contiguous()
    if bias is not None:
        bias = bias.contiguous()
    _lib.call("dmp_gemm_tf32x3", A.device, _lib.ptr(A), lda, _lib.ptr(row_scale), _lib.ptr(Wt), ldb,
              _lib.ptr(bias), _lib.ptr(aux if mul_act_grad else None), ld_aux, _lib.ptr(out), ldd, M, N, K, epi,
              float(slope), _stream(A),
              tag="gemm_tf32x3." + ("acc" if accumulate else "grad" if mul_act_grad else
                                    "bias_act" if (bias is not None or act != "none") else "store") +
                  ("_scaled" if row_scale is not None else ""),
              nbytes=4 * (M * K + M * N * (2 if accumulate else 1) + (M * N if mul_act_grad else 0) + N * K))
    return out


_tn_ws = {}


def gemm_tn_tf32x3(X, G, *, row_scale=None, out=None, accumulate=False, colsum_x=False, colsum_g=False):
    """D (+)= (row_scale ⊙ X).T @ G on the tensor cores (3xTF32), X [E,M], G [E,N], M,N in {64,128}. Raw call.
    With colsum_x / colsum_g also returns X.sum(0) (unscaled) / G.sum(0): (D, sum_x or None, sum_g or None)."""
    _lib.require_cuda(X, G, row_scale)
    X, ldx = _lib.row_major(X)
    G, ldg = _lib.row_major(G)
    E, M = X.shape
    N = G.shape[1]
    if G.shape[0] != E:
        raise ValueError("row counts differ: %s vs %s" % (tuple(X.shape), tuple(G.shape)))
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs `out`")
        out = torch.empty((M, N), dtype=torch.float32, device=X.device)
    _, ldd = _lib.row_major(out)
    key = (str(X.device), _stream(X), M, N)   # one workspace per stream: concurrent calls must not share partials
    ws = _tn_ws.get(key)
    if ws is None:
        import ctypes
        nb = ctypes.c_int64(0)
        _lib.check(_lib.load().dmp_gemm_tn_workspace_bytes(M, N, ctypes.byref(nb)), "dmp_gemm_tn_workspace_bytes")
        ws = _tn_ws[key] = torch.empty(nb.value, dtype=torch.uint8, device=X.device)
    if row_scale is not None:
        row_scale = row_scale.reshape(-1).contiguous()
    sx = torch.empty(M, dtype=torch.float32, device=X.device) if colsum_x else None
    sg = torch.empty(N, dtype=torch.float32, device=X.device) if colsum_g else None
    _lib.call("dmp_gemm_tn_tf32x3", X.device, _lib.ptr(X), ldx, _lib.ptr(row_scale), _lib.ptr(G), ldg, _lib.ptr(out),
              ldd, _lib.ptr(sx), _lib.ptr(sg), E, M, N, int(accumulate), _lib.ptr(ws), ws.numel(), _stream(X),
              tag="gemm_tn_tf32x3", nbytes=4 * (E * M + E * N + M * N) + (4 * E if row_scale is not None else 0))
    if colsum_x or colsum_g:
        return out, sx, sg
    return out


def gemm_tf32x3_dual(A, W1t, W2t, *, row_scale=None, mode="store", out=None, out2=None):
    """Two projections of the same streamed operand in ONE pass (dmp_gemm_tf32x3_dual):
        "store"       out  = A @ W1t.T + row_scale ⊙ (A @ W2t.T)
        "accumulate"  out  = (out + A @ W1t.T) + row_scale ⊙ (A @ W2t.T)
        "separate"    out, out2 = A @ W1t.T, A @ W2t.T
    A [M,K] fp32 dense rows; W1t, W2t [N,K] (nn.Linear layout); N, K in {64,128}.  Raw, non-differentiable."""
    _lib.require_cuda(A, W1t, W2t, row_scale, out, out2)
    A, lda = _lib.row_major(A)
    W = torch.stack([W1t, W2t]).contiguous()       # common leading dimension, 16-byte aligned
    M, K = A.shape
    N = W1t.shape[0]
    if tuple(W1t.shape) != (N, K) or tuple(W2t.shape) != (N, K):
        raise ValueError("weights must both be [N,K]=[%d,%d]: got %s, %s" % (N, K, tuple(W1t.shape), tuple(W2t.shape)))
    imode = {"store": _lib.DUAL_STORE, "accumulate": _lib.DUAL_ACCUMULATE, "separate": _lib.DUAL_SEPARATE}[mode]
    if out is None:
        if mode == "accumulate":
            raise ValueError("accumulate needs `out`")
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _, ldd = _lib.row_major(out)
    ldd2 = 0
    if mode == "separate":
        if out2 is None:
            out2 = torch.empty((M, N), dtype=torch.float32, device=A.device)
        _, ldd2 = _lib.row_major(out2)
    if row_scale is not None:
        row_scale = row_scale.reshape(-1).contiguous()
        if row_scale.numel() != M:
            raise ValueError("row_scale has %d entries for %d rows" % (row_scale.numel(), M))
    _lib.call("dmp_gemm_tf32x3_dual", A.device, _lib.ptr(A), lda, _lib.ptr(W[0]), _lib.ptr(W[1]), K,
              _lib.ptr(row_scale), _lib.ptr(out), ldd, _lib.ptr(out2 if mode == "separate" else None), ldd2, M, N, K,
              imode, _stream(A), tag="gemm_tf32x3_dual." + mode,
              nbytes=4 * (M * K + M * N * {"store": 1, "accumulate": 2, "separate": 2}[mode] + 2 * N * K)
              + (4 * M if row_scale is not None else 0))
    return (out, out2) if mode == "separate" else out


_bn_ws = {}


def _bn_workspace(device, H):
    import ctypes
    key = (str(device), _lib.stream_ptr(device), H)
    ws = _bn_ws.get(key)
    if ws is None:
        nb = ctypes.c_int64(0)
        _lib.check(_lib.load().dmp_bn_workspace_bytes(H, ctypes.byref(nb)), "dmp_bn_workspace_bytes")
        ws = _bn_ws[key] = torch.zeros(nb.value, dtype=torch.uint8, device=device)   # zeroed once; kernels keep it so
    return ws


def bn_stats(x):
    """(mean, biased variance) over the rows of x [rows,H]: two deterministic passes (dmp_bn_stats)."""
    _lib.require_cuda(x)
    x, ldx = _lib.row_major(x)
    rows, H = x.shape
    mean = torch.empty(H, dtype=torch.float32, device=x.device)
    var = torch.empty(H, dtype=torch.float32, device=x.device)
    ws = _bn_workspace(x.device, H)
    _lib.call("dmp_bn_stats", x.device, _lib.ptr(x), ldx, rows, H, _lib.ptr(mean), _lib.ptr(var), _lib.ptr(ws),
              ws.numel(), _stream(x), tag="bn_stats", nbytes=2 * 4 * rows * H)
    return mean, var


def bn_act(x, mean, invstd, gamma, beta, act, slope, out=None):
    """act(((x - mean) * invstd) * gamma + beta), elementwise (dmp_bn_act)."""
    _lib.require_cuda(x, mean, invstd, gamma, beta)
    x, ldx = _lib.row_major(x)
    rows, H = x.shape
    if out is None:
        out = torch.empty((rows, H), dtype=torch.float32, device=x.device)
    _, ldo = _lib.row_major(out)
    _lib.call("dmp_bn_act", x.device, _lib.ptr(x), ldx, _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(gamma),
              _lib.ptr(beta), _lib.ptr(out), ldo, rows, H, act, float(slope), _stream(x), tag="bn_act",
              nbytes=2 * 4 * rows * H)
    return out


def bn_backward(g, x, mean, invstd, gamma, training, out=None):
    """(gx, dgamma, dbeta) of y = ((x - mean) * invstd) * gamma + beta given g = dL/dy; gx may be g (in place)."""
    _lib.require_cuda(g, x, mean, invstd, gamma)
    g, ldg = _lib.row_major(g)
    x, ldx = _lib.row_major(x)
    rows, H = g.shape
    if out is None:
        out = g
    _, ldo = _lib.row_major(out)
    dgamma = torch.empty(H, dtype=torch.float32, device=g.device)
    dbeta = torch.empty(H, dtype=torch.float32, device=g.device)
    ws = _bn_workspace(g.device, H)
    _lib.call("dmp_bn_backward", g.device, _lib.ptr(g), ldg, _lib.ptr(x), ldx, _lib.ptr(mean), _lib.ptr(invstd),
              _lib.ptr(gamma), _lib.ptr(out), ldo, _lib.ptr(dgamma), _lib.ptr(dbeta), rows, H, int(bool(training)),
              _lib.ptr(ws), ws.numel(), _stream(g), tag="bn_backward", nbytes=5 * 4 * rows * H)
    return out, dgamma, dbeta


# ---- sparse matrix x dense rows as a weighted segment reduce (row N4: the perm-pooling of DMPLRPPoolLayer) ----------------
class SparseRows:
    """CSR of a torch sparse matrix S and of its transpose, in the layout dmp_segment_reduce takes (int32 indptr, int32
    column list, fp32 weights in position order).  Built once per matrix (the LRP pooling matrices are preprocessing
    outputs, fixed per batch) and cached on the tensor."""

    def __init__(self, S):
        S = S.coalesce() if S.layout == torch.sparse_coo else S.to_sparse_coo().coalesce()
        _lib.require_cuda(S)
        rows, cols = S.indices()
        vals = S.values().float()
        self.shape = tuple(S.shape)

        def csr(r, c, v, n, width):
            order = torch.argsort(r * width + c)              # (row, col) order = torch's coalesced storage order
            r, c, v = r[order], c[order], v[order]
            indptr = torch.zeros(n + 1, dtype=torch.int32, device=r.device)
            indptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
            return indptr, c.to(torch.int32).contiguous(), v.contiguous()

        self.fwd = csr(rows, cols, vals, self.shape[0], self.shape[1])
        self.bwd = csr(cols, rows, vals, self.shape[1], self.shape[0])


def _sparse_rows(S):
    cached = getattr(S, "_dmp_rows", None)
    if cached is None:
        cached = SparseRows(S)
        try:
            S._dmp_rows = cached
        except AttributeError:
            pass
    return cached


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, X):
        indptr, idx, w = rows.fwd
        ctx.rows = rows
        return segment_reduce(indptr, idx, X.contiguous(), X.shape[1], w_perm=w, tag="segment_reduce.spmm")

    @staticmethod
    def backward(ctx, g):
        indptr, idx, w = ctx.rows.bwd
        return None, segment_reduce(indptr, idx, g.contiguous(), g.shape[1], w_perm=w, tag="segment_reduce.spmm_bwd")


def spmm(S, X):
    """`torch.sparse.mm(S, X)` (dmplrp.py:180,185) as one deterministic weighted segment reduce: row i of the result is
    sum_j S[i,j] X[j,:] accumulated sequentially in column order; backward = the same kernel over S^T.  No gradient
    with respect to S (the reference's pooling matrices are constants)."""
    return _SpMM.apply(_sparse_rows(S), X.float())
