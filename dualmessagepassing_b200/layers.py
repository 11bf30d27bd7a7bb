"""Drop-in DMPNN layers: same constructors, parameter names and call signatures as the reference.

  DMPLayer        <- SubgraphCountingMatching/models/dmpnn.py:16-176
  DualGraphConv   <- UnsupervisedNodeClassification/Model/DMPNN/src/model.py:117-280

`forward(graph, node_feat, edge_feat[, edge_norm])` takes a (batched) DGLGraph or a `DMPGraph` whose
tensors live on a CUDA device.  The DGL `update_all` / `apply_edges` path is replaced by the sm_100a
sparse core behind the C ABI (`functional.sparse_core`); the per-layer projections stay dense GEMMs
(cuBLAS through torch.mm).  Node-side projections are done BEFORE the endpoint gather
(`Q = X_v W`, then `Q[a] - Q[b]`), which is bit-identical per row to the reference's gather-then-project
 and removes E-sized GEMMs; only the selected branch of the reversed / forward
message is computed instead of both followed by masked_fill.

There is no CPU path: CPU tensors raise.
"""
import torch
import torch.nn as nn

from . import _lib
from .act import map_activation_str_to_layer
from .constants import (EDGEFEAT, LEAKY_RELU_A, NODEFEAT, OUTDEGREE, REVFLAG, UNC_FEAT, UNC_NORM,
                        UNC_OUTDEGREE, UNC_REVFLAG)
from .functional import sparse_core
from . import fused as _fused
from .fused import fused_dmp_layer, mlp_spec_and_tensors, supported_activation
from .init import init_module, init_weight
from .plan import get_plan


class _SplitMM(torch.autograd.Function):
    """M[:h] = X[:h] @ W0 ; M[h:] = X[h:] @ W1 written into one buffer (the [forward | reversed] edge layout
    produced by add_reversed_edges on a single graph): each edge only pays for its own branch."""

    @staticmethod
    def forward(ctx, X, W0, W1, h):
        M = torch.empty((X.shape[0], W0.shape[1]), dtype=X.dtype, device=X.device)
        torch.mm(X[:h], W0, out=M[:h])
        torch.mm(X[h:], W1, out=M[h:])
        ctx.save_for_backward(X, W0, W1)
        ctx.h = h
        return M

    @staticmethod
    def backward(ctx, gM):
        X, W0, W1 = ctx.saved_tensors
        h = ctx.h
        gX = gW0 = gW1 = None
        if ctx.needs_input_grad[0]:
            gX = torch.empty_like(X)
            torch.mm(gM[:h], W0.t(), out=gX[:h])
            torch.mm(gM[h:], W1.t(), out=gX[h:])
        if ctx.needs_input_grad[1]:
            gW0 = X[:h].t() @ gM[:h]
        if ctx.needs_input_grad[2]:
            gW1 = X[h:].t() @ gM[h:]
        return gX, gW0, gW1, None


def dual_message_passing(plan, node_feat, edge_feat, in_weight, out_weight, src_weight, dst_weight,
                         nloop_weight, eloop_weight, nbias, ebias, norm=None, order=_lib.ORDER_SCM):
    """(node_pre, edge_pre) of one dual message-passing step: the part of the reference layer between the
    frame initialisation and the MLP/activation (dmpnn.py:111-133,142-149)."""
    _lib.require_cuda(node_feat, edge_feat)
    if node_feat.shape[0] != plan.N or edge_feat.shape[0] != plan.E:
        raise ValueError("feature rows (%d nodes, %d edges) do not match the graph (%d, %d)"
                         % (node_feat.shape[0], edge_feat.shape[0], plan.N, plan.E))
    X_v = node_feat.float()
    X_e = edge_feat.float()
    H = nloop_weight.shape[1]
    # node-sized projections (project-then-gather)
    Ln = X_v @ nloop_weight
    Qd = X_v @ dst_weight
    Qs = X_v @ src_weight
    # edge-sized projections
    S = X_e @ eloop_weight
    P = X_e @ (src_weight - dst_weight)
    m_rev_off = 0
    if plan.rev_layout == "none":
        M = X_e @ in_weight
    elif plan.rev_layout == "halves":
        M = _SplitMM.apply(X_e, in_weight, out_weight, plan.rev_split)
    else:
        M = X_e @ torch.cat([in_weight, out_weight], dim=1)  # [E, 2H]: branch picked inside the kernel
        m_rev_off = H
    return sparse_core(plan, M, S, P, Ln, Qd, Qs, nbias, ebias, norm=norm, order=order, m_rev_off=m_rev_off)


def _padded_width(d):
    """Width the tcgen05 kernels take (64 or 128) for a feature dimension d, or None when padding would more than
    double the row (d <= 32) or d > 128: those widths stay on cuBLAS."""
    if 32 < d <= 64:
        return 64
    return 128 if 64 < d <= 128 else None


def _pad_cols(t, width):
    """Zero-pad the last dimension to `width` (differentiable: autograd slices the gradient back)."""
    return t if t is None or t.shape[-1] == width else torch.nn.functional.pad(t, (0, width - t.shape[-1]))


def _pad_mat(w, rows, cols):
    return torch.nn.functional.pad(w, (0, cols - w.shape[1], 0, rows - w.shape[0]))


_pad_maps = {}


def _pad_map(shapes, targets, device):
    """(gather index padded <- [sources | 0], gather index sources <- padded) for zero-padding a list of small tensors
    into their target shapes; cached per shape signature (the shapes of a layer's parameters never change)."""
    key = (tuple(shapes), tuple(targets), str(device))
    hit = _pad_maps.get(key)
    if hit is None:
        n_src = sum(int(torch.Size(s).numel()) for s in shapes)
        fwd, inv, src_off, dst_off = [], [], 0, 0
        for s, t in zip(shapes, targets):
            idx = torch.full(tuple(t), n_src, dtype=torch.int64)             # n_src = the appended zero
            src = torch.arange(src_off, src_off + int(torch.Size(s).numel()), dtype=torch.int64).view(tuple(s))
            window = tuple(slice(0, d) for d in s)
            idx[window] = src
            pos = torch.arange(dst_off, dst_off + idx.numel(), dtype=torch.int64).view(tuple(t))
            fwd.append(idx.reshape(-1))
            inv.append(pos[window].reshape(-1))
            src_off += src.numel()
            dst_off += idx.numel()
        hit = _pad_maps[key] = (torch.cat(fwd).to(device), torch.cat(inv).to(device))
    return hit


class _PadMany(torch.autograd.Function):
    """Zero-pad ~30 parameter tensors of a layer in two kernels (concatenate, gather) instead of a fill and a copy each,
    forward and backward: hidden 50 -> 64 spent a fifth of the UNC encoder step in `F.pad` and its slice gradients."""

    @staticmethod
    def forward(ctx, targets, *ts):
        shapes = [tuple(t.shape) for t in ts]
        fwd, inv = _pad_map(shapes, targets, ts[0].device)
        flat = torch.cat([t.reshape(-1) for t in ts] + [ts[0].new_zeros(1)])[fwd]
        ctx.inv, ctx.shapes, ctx.targets = inv, shapes, targets
        return tuple(flat.split([int(torch.Size(t).numel()) for t in targets]))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g.reshape(-1) if g is not None else torch.zeros(int(torch.Size(t).numel()), device=ctx.inv.device)
              for g, t in zip(gs, ctx.targets)]
        src = torch.cat(gs)[ctx.inv]
        out = src.split([int(torch.Size(s).numel()) for s in ctx.shapes])
        return (None, *[o.view(s) for o, s in zip(out, ctx.shapes)])


def _pad_many(tensors, targets):
    """Zero-padded copies of `tensors` (None entries pass through) with shapes `targets`."""
    live = [(i, t) for i, t in enumerate(tensors) if t is not None]
    if not live:
        return list(tensors)
    flat = _PadMany.apply(tuple(tuple(targets[i]) for i, _ in live), *[t.float() for _, t in live])
    out = list(tensors)
    for (i, _), f in zip(live, flat):
        out[i] = f.view(tuple(targets[i]))
    return out


def _run_fused(layer, plan, node_feat, edge_feat, nmlp, emlp, *, act_func, slope, order, norm=None, post_act="none"):
    """Whole-layer function (fused.py) with the widths the tensor-core kernels take.

    UNC/run.sh trains with hidden 50: when the graph is large enough for the tcgen05 kernels and a width is not
    64 / 128, features, weights, biases and BatchNorm affine parameters are zero-padded to the next supported
    width (exact: padded inputs meet zero weights, padded outputs are exactly 0 and are sliced off; the gradient of
    the padding is the slice)."""
    Din, H = layer.input_dim, layer.hidden_dim
    weights = (layer.in_weight, layer.out_weight, layer.src_weight, layer.dst_weight, layer.nloop_weight,
               layer.eloop_weight)
    nbias, ebias = layer.nbias, layer.ebias
    pd, ph = _padded_width(Din), _padded_width(H)
    pad = (_fused.DENSE_BACKEND == "auto" and edge_feat.shape[0] >= _fused.TC_MIN_ROWS and pd is not None
           and ph is not None and (pd != Din or ph != H))
    xv, xe = node_feat.float(), edge_feat.float()
    if pad:
        small = list(weights) + [nbias, ebias] + list(nmlp[1]) + list(emlp[1])
        targets = [(pd, ph)] * 6 + [(ph,), (ph,)] + [None if t is None else ((ph, ph) if t.dim() == 2 else (ph,))
                                                     for t in list(nmlp[1]) + list(emlp[1])]
        small = _pad_many(small, targets)
        weights, nbias, ebias = tuple(small[:6]), small[6], small[7]
        n_n = len(nmlp[1])
        nmlp, emlp = (nmlp[0], small[8:8 + n_n]), (emlp[0], small[8 + n_n:])
        xv, xe = _pad_cols(xv, pd), _pad_cols(xe, pd)
    nv, ne = fused_dmp_layer(plan, xv, xe, weights, nbias, ebias, nmlp, emlp, act_func=act_func, slope=slope,
                             order=order, norm=norm, post_act=post_act, training=layer.training)
    if pad and ph != H:
        nv, ne = nv[:, :H], ne[:, :H]
    return nv, ne


def _act_of_module(m):
    """(name, slope) of an activation module the fused kernels implement, else None."""
    if m is None or isinstance(m, nn.Identity):
        return "none", 0.0
    if isinstance(m, nn.LeakyReLU):
        return "leaky_relu", float(m.negative_slope)
    if isinstance(m, nn.ReLU):
        return "relu", 0.0
    if isinstance(m, nn.Tanh):
        return "tanh", 0.0
    if isinstance(m, nn.Sigmoid):
        return "sigmoid", 0.0
    return None


class DMPLayer(nn.Module):
    """SubgraphCountingMatching/models/dmpnn.py:16-176 -- same constructor, state_dict and forward."""

    def __init__(self, input_dim, hidden_dim, init_neigenv=4.0, init_eeigenv=4.0, bias=True,
                 num_mlp_layers=2, batch_norm=True, act_func="relu", dropout=0.0):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.act_func = act_func
        self.num_mlp_layers = num_mlp_layers
        self.batch_norm = batch_norm
        # "auto": whole-layer function with explicit buffer re-use (fused.py) whenever the configuration
        # allows it; False forces the composed autograd path (same kernels, torch-managed memory)
        self.fused = "auto"

        def weight():
            return nn.Parameter(torch.empty(input_dim, hidden_dim))

        self.in_weight = weight()
        self.out_weight = weight()
        self.src_weight = weight()
        self.dst_weight = weight()
        self.nloop_weight = weight()
        self.eloop_weight = weight()
        if bias:
            self.nbias = nn.Parameter(torch.empty(hidden_dim))
            self.ebias = nn.Parameter(torch.empty(hidden_dim))
        else:
            self.register_parameter("nbias", None)
            self.register_parameter("ebias", None)

        def mlp():
            mods = []
            for i in range(num_mlp_layers):
                mods.append(nn.Linear(hidden_dim, hidden_dim))
                if i != num_mlp_layers - 1:
                    if batch_norm:
                        mods.append(nn.BatchNorm1d(hidden_dim))
                    mods.append(map_activation_str_to_layer(act_func))
            return nn.Sequential(*mods)

        self.nmlp = mlp()
        self.emlp = mlp()
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)

        # same RNG consumption order as the reference (dmpnn.py:64-77): six weights, then the MLPs
        for w in (self.in_weight, self.out_weight, self.src_weight, self.dst_weight, self.nloop_weight,
                  self.eloop_weight):
            init_weight(w, activation=act_func, init="uniform")
        for m in self.nmlp.modules():
            init_module(m, activation=act_func, init="uniform")
        for m in self.emlp.modules():
            init_module(m, activation=act_func, init="uniform")
        if bias:
            nn.init.zeros_(self.nbias)
            nn.init.zeros_(self.ebias)
        with torch.no_grad():  # "reparametrisation trick" dmpnn.py:79-86
            self.in_weight.div_(init_neigenv)
            self.out_weight.div_(init_neigenv)
            self.nloop_weight.div_(init_neigenv)
            self.src_weight.div_(init_eeigenv)
            self.dst_weight.div_(init_eeigenv)
            self.eloop_weight.div_(init_eeigenv)

    def _fused_mlps(self):
        """(nmlp, emlp) as (MLPSpec, tensors) when the whole-layer function covers this configuration, else None."""
        if not self.fused or not supported_activation(self.act_func):
            return None
        n, e = mlp_spec_and_tensors(self.nmlp), mlp_spec_and_tensors(self.emlp)
        return None if n is None or e is None else (n, e)

    def forward(self, graph, node_feat, edge_feat):
        plan = get_plan(graph, REVFLAG, OUTDEGREE)
        # frame side effects of dmpnn.py:96-109 (inputs stay bound to the graph)
        graph.ndata[NODEFEAT] = node_feat
        graph.edata[EDGEFEAT] = edge_feat
        mlps = self._fused_mlps()
        if mlps is not None:
            node_out, edge_out = _run_fused(self, plan, node_feat, edge_feat, mlps[0], mlps[1], act_func=self.act_func,
                                            slope=LEAKY_RELU_A, order=_lib.ORDER_SCM)
            return self.drop(node_out), self.drop(edge_out)       # dmpnn.py:138,154 (identity unless training with p > 0)
        node_pre, edge_pre = dual_message_passing(
            plan, node_feat, edge_feat, self.in_weight, self.out_weight, self.src_weight, self.dst_weight,
            self.nloop_weight, self.eloop_weight, self.nbias, self.ebias, order=_lib.ORDER_SCM)
        if len(self.nmlp) > 0:
            node_out = self.nmlp(node_pre)
        else:
            node_out = self.act(node_pre)
        if len(self.emlp) > 0:
            edge_out = self.emlp(edge_pre)
        else:
            edge_out = self.act(edge_pre)
        return self.drop(node_out), self.drop(edge_out)

    def extra_repr(self):
        return "in=%s, out=%s" % (self.input_dim, self.hidden_dim)

    def get_output_dim(self):
        return self.hidden_dim


class DMPLRPPoolLayer(DMPLayer):
    """SubgraphCountingMatching/models/dmplrp.py:19-198 -- the dual message-passing step of DMPLayer (its body is a
    verbatim copy in the reference, dmplrp.py:123-168) followed by local relational pooling over permutation
    sequences (dmplrp.py:180-185):

        z = node_to_perm @ node_out + edge_to_perm @ edge_out          two sparse products  -> [D * L^2, H]
        y[d, c] = sum_{a, b} z[d, a, b] * lrp_weight[b, c, a] + lrp_bias                     one dense contraction
        node_out = pooling_matrix @ y                                   one sparse product   -> [N, H]

    The three sparse products run as weighted segment reduces on the sm_100a kernel (`functional.spmm`), sequential in
    column order like torch's CPU `sparse.mm`.  Same constructor, state_dict (`lrp_weight` [in, hid, L^2], `lrp_bias`)
    and forward signature / 5-tuple result as the reference."""

    def __init__(self, input_dim, hidden_dim, init_neigenv=4.0, init_eeigenv=4.0, lrp_seq_len=4, bias=True,
                 num_mlp_layers=2, batch_norm=True, act_func="relu", dropout=0.0):
        nn.Module.__init__(self)
        self.input_dim, self.hidden_dim, self.lrp_seq_len = input_dim, hidden_dim, lrp_seq_len
        self.act_func, self.num_mlp_layers, self.batch_norm, self.num_rels = act_func, num_mlp_layers, batch_norm, 3
        self.fused = "auto"

        def weight(*extra):
            return nn.Parameter(torch.empty(input_dim, hidden_dim, *extra))

        # parameter registration and RNG consumption order of dmplrp.py:38-85
        self.in_weight, self.out_weight = weight(), weight()
        self.src_weight, self.dst_weight = weight(), weight()
        self.nloop_weight, self.eloop_weight = weight(), weight()
        self.lrp_weight = weight(lrp_seq_len * lrp_seq_len)
        if bias:
            self.nbias = nn.Parameter(torch.empty(hidden_dim))
            self.ebias = nn.Parameter(torch.empty(hidden_dim))
            self.lrp_bias = nn.Parameter(torch.empty(hidden_dim))
        else:
            for k in ("nbias", "ebias", "lrp_bias"):
                self.register_parameter(k, None)

        def mlp():
            mods = []
            for i in range(num_mlp_layers):
                mods.append(nn.Linear(hidden_dim, hidden_dim))
                if i != num_mlp_layers - 1:
                    if batch_norm:
                        mods.append(nn.BatchNorm1d(hidden_dim))
                    mods.append(map_activation_str_to_layer(act_func))
            return nn.Sequential(*mods)

        self.nmlp, self.emlp = mlp(), mlp()
        self.act = map_activation_str_to_layer(act_func)
        self.drop = nn.Dropout(dropout)
        for w in (self.in_weight, self.out_weight, self.src_weight, self.dst_weight, self.nloop_weight,
                  self.eloop_weight):
            init_weight(w, activation=act_func, init="uniform")
        init_weight(self.lrp_weight, init="uniform")
        for m in self.nmlp.modules():
            init_module(m, activation=act_func, init="uniform")
        for m in self.emlp.modules():
            init_module(m, activation=act_func, init="uniform")
        if bias:
            for b in (self.nbias, self.ebias, self.lrp_bias):
                nn.init.zeros_(b)
        with torch.no_grad():
            for w in (self.in_weight, self.out_weight, self.nloop_weight):
                w.div_(init_neigenv)
            for w in (self.src_weight, self.dst_weight, self.eloop_weight):
                w.div_(init_eeigenv)

    def forward(self, graph, node_feat, edge_feat, pooling_matrix, node_to_perm_matrix, edge_to_perm_matrix):
        from .functional import spmm
        node_out, edge_out = DMPLayer.forward(self, graph, node_feat, edge_feat)
        z = spmm(node_to_perm_matrix, node_out) + spmm(edge_to_perm_matrix, edge_out)
        z = z.view(-1, self.lrp_seq_len * self.lrp_seq_len, self.input_dim)
        y = torch.einsum("dab,bca->dc", z, self.lrp_weight)
        if self.lrp_bias is not None:
            y = y + self.lrp_bias
        return spmm(pooling_matrix, y), edge_out, pooling_matrix, node_to_perm_matrix, edge_to_perm_matrix

    def extra_repr(self):
        return "in=%s, out=%s\nlrp_seq_len=%s" % (self.input_dim, self.hidden_dim, self.lrp_seq_len)


class DualGraphConv(nn.Module):
    """UnsupervisedNodeClassification/Model/DMPNN/src/model.py:117-280 -- same constructor, state_dict
    (including the unused `nfc` / `efc`, model.py:137-138) and forward(graph, node_feat, edge_feat, edge_norm)."""

    def __init__(self, input_dim, hidden_dim, init_neigenv=4.0, init_eeigenv=4.0, bias=True, batch_norm=True,
                 activation=None, dropout=0.0):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim

        def weight():
            return nn.Parameter(torch.empty(input_dim, hidden_dim))

        self.in_weight = weight()
        self.out_weight = weight()
        self.src_weight = weight()
        self.dst_weight = weight()
        self.nloop_weight = weight()
        self.eloop_weight = weight()
        self.nfc = nn.Linear(hidden_dim, hidden_dim)
        self.efc = nn.Linear(hidden_dim, hidden_dim)
        if bias:
            self.nbias = nn.Parameter(torch.zeros(hidden_dim))
            self.ebias = nn.Parameter(torch.zeros(hidden_dim))
        else:
            self.register_parameter("nbias", None)
            self.register_parameter("ebias", None)

        def mlp():
            mods = [nn.Linear(hidden_dim, hidden_dim)]
            if batch_norm:
                mods.append(nn.BatchNorm1d(hidden_dim))
            mods.append(nn.LeakyReLU(LEAKY_RELU_A) if activation is None else activation)
            mods.append(nn.Linear(hidden_dim, hidden_dim))
            return nn.Sequential(*mods)

        self.nmlp = mlp()
        self.emlp = mlp()
        self.act = activation
        self.drop = nn.Dropout(dropout)
        self.fused = "auto"   # False forces the composed autograd path (same kernels, torch-managed memory)

        for w in (self.in_weight, self.out_weight, self.src_weight, self.dst_weight, self.nloop_weight,
                  self.eloop_weight, self.nmlp[0].weight, self.nmlp[-1].weight, self.emlp[0].weight,
                  self.emlp[-1].weight):
            nn.init.xavier_uniform_(w)
        for b in (self.nmlp[0].bias, self.nmlp[-1].bias, self.emlp[0].bias, self.emlp[-1].bias):
            nn.init.zeros_(b)
        with torch.no_grad():
            self.in_weight.div_(init_neigenv)
            self.out_weight.div_(init_neigenv)
            self.nloop_weight.div_(init_neigenv)
            self.src_weight.div_(init_eeigenv)
            self.dst_weight.div_(init_eeigenv)
            self.eloop_weight.div_(init_eeigenv)

    def _fused_cfg(self):
        """(nmlp, emlp, mlp activation, slope, post activation) when the whole-layer function covers this module."""
        if not self.fused:
            return None
        n, e = mlp_spec_and_tensors(self.nmlp), mlp_spec_and_tensors(self.emlp)
        if n is None or e is None or n[0].n_lin != 2:
            return None
        inner = _act_of_module(self.nmlp[-2])                 # model.py:148,154: LeakyReLU(1/5.5) or `activation`
        post = _act_of_module(self.act)                        # model.py:247-248
        if inner is None or post is None or self.emlp[-2].__class__ is not self.nmlp[-2].__class__:
            return None
        if inner[0] == "leaky_relu" and post[0] == "leaky_relu" and inner[1] != post[1]:
            return None
        slope = inner[1] if inner[0] == "leaky_relu" else post[1]
        return n, e, inner[0], slope, post[0]

    def forward(self, graph, node_feat, edge_feat, edge_norm=None):
        plan = get_plan(graph, UNC_REVFLAG, UNC_OUTDEGREE)
        graph.ndata[UNC_FEAT] = node_feat
        graph.edata[UNC_FEAT] = edge_feat
        if edge_norm is not None:
            graph.edata[UNC_NORM] = edge_norm
        norm = graph.edata[UNC_NORM] if UNC_NORM in graph.edata else None  # model.py:234: key presence decides
        cfg = self._fused_cfg()
        if cfg is not None:
            # model.py:245,260: `self.drop(out)` is called and its result DISCARDED -- a numerical no-op that only
            # advances the RNG in training mode; mirrored for RNG-stream parity
            if self.training and self.drop.p > 0:
                self.drop(torch.empty((plan.N, self.hidden_dim), device=node_feat.device))
                self.drop(torch.empty((plan.E, self.hidden_dim), device=node_feat.device))
            nmlp, emlp, act_name, slope, post = cfg
            return _run_fused(self, plan, node_feat, edge_feat, nmlp, emlp, act_func=act_name, slope=slope,
                              order=_lib.ORDER_UNC, norm=norm, post_act=post)
        node_pre, edge_pre = dual_message_passing(
            plan, node_feat, edge_feat, self.in_weight, self.out_weight, self.src_weight, self.dst_weight,
            self.nloop_weight, self.eloop_weight, self.nbias, self.ebias, norm=norm, order=_lib.ORDER_UNC)
        # model.py:245,260: the reference calls self.drop(out) and discards the result -- a numerical
        # no-op that only advances the RNG in training mode; mirrored for RNG-stream parity.
        if self.training and self.drop.p > 0:
            self.drop(node_pre)
            self.drop(edge_pre)
        node_out = self.nmlp(node_pre)
        edge_out = self.emlp(edge_pre)
        if self.act:
            node_out = self.act(node_out)
            edge_out = self.act(edge_out)
        return node_out, edge_out

    def extra_repr(self):
        return "in=%s, out=%s" % (self.input_dim, self.hidden_dim)
