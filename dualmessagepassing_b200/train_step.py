"""End-to-end training step around the DMPNN layers, for the graphs/sec measurement.

The reference's trainer and `GraphAdjModelV2` (SubgraphCountingMatching/train.py:449-844,
models/basemodel.py:965-1663) are callers of the hot path and are NOT rebuilt.  What the e2e metric needs is
one step of the same shape: label embeddings -> label-match filter gate -> 3 shared DMP layers on the pattern
batch and on the graph batch (mask / gate / residual) -> sum-pool prediction head -> MSE -> backward -> clip ->
AdamW(amsgrad).  This module is that step, written for throughput: the per-batch host work is one vectorised
collate, the device work has no host synchronisation, ragged pooling is a sorted-segment reduce.

  SyntheticPairDataset   directed ER pattern/graph pairs of the BASELINE configs (numpy PCG64), stored flat
  collate                `dgl.batch` + `add_reversed_edges` semantics on flat arrays (dataset.py:1321-1328,
                         train.py:299-327): per graph [forward block | reversed block], offsets = prefix sums
  SubgraphCountingModel  embedding / gate / DMPNNRepNet / SumPredictNet-shaped head (pred.py:87-156,198-216)
"""
import numpy as np
import torch
import torch.nn as nn

from .constants import EDGELABEL, NODELABEL, REVFLAG
from .functional import segment_reduce
from .graph import DMPGraph
from .models import DMPNNRepNet

CONFIGS = {
    # name: pairs/batch, graph n range, E0 per node, pattern n range, pattern E0 range, labels (v,e) graph, pattern, H
    "cfg1": dict(pairs=64, gn=(8, 32), ge_per_n=3, pn=(3, 4), pe=(2, 4), labels=(1, 1), hidden=64),
    "cfg2": dict(pairs=512, gn=(16, 64), ge_per_n=4, pn=(3, 8), pe=(2, 8), labels=(16, 16), hidden=64),
    "cfg3": dict(pairs=64, gn=(10, 28), ge_per_n=2, pn=(4, 4), pe=(3, 3), labels=(7, 4), hidden=128),
}


class SyntheticPairDataset:
    """`num` pattern/graph pairs as flat arrays: offsets[i]..offsets[i+1] index graph i's nodes / edges."""

    def __init__(self, cfg, num, seed):
        c = CONFIGS[cfg]
        rng = np.random.Generator(np.random.PCG64(seed))
        self.cfg, self.num = c, num

        def build(n_range, e_fn, lv, le):
            n = rng.integers(n_range[0], n_range[1] + 1, size=num)
            e = np.asarray([e_fn(int(k)) for k in n], dtype=np.int64)
            noff = np.concatenate([[0], np.cumsum(n)])
            eoff = np.concatenate([[0], np.cumsum(e)])
            owner = np.repeat(np.arange(num), e)
            nn_ = n[owner]
            u = (rng.random(eoff[-1]) * nn_).astype(np.int64)
            v = (rng.random(eoff[-1]) * np.maximum(nn_ - 1, 1)).astype(np.int64)
            v = np.where(nn_ > 1, v + (v >= u), v)  # no self loops
            return dict(n=n, e=e, noff=noff, eoff=eoff, u=u, v=v,
                        vl=rng.integers(0, lv, size=noff[-1]), el=rng.integers(0, le, size=eoff[-1]))

        lv, le = c["labels"]
        self.g = build(c["gn"], lambda k: c["ge_per_n"] * k, lv, le)
        self.p = build(c["pn"], lambda k: int(rng.integers(c["pe"][0], c["pe"][1] + 1)), lv, le)
        self.counts = rng.poisson(3.0, size=num).astype(np.float32)


def _collate_side(d, idx):
    """Disjoint union of graphs `idx`, reversed edges appended per graph (per-graph [fwd | rev] blocks)."""
    n, e = d["n"][idx], d["e"][idx]
    new_noff = np.concatenate([[0], np.cumsum(n)])
    # gather node ranges
    node_src = np.repeat(d["noff"][idx] - new_noff[:-1], n) + np.arange(new_noff[-1])
    e2 = 2 * e
    new_eoff = np.concatenate([[0], np.cumsum(e2)])
    owner = np.repeat(np.arange(len(idx)), e2)
    pos = np.arange(new_eoff[-1]) - new_eoff[:-1][owner]        # position inside the graph's doubled edge list
    is_rev = pos >= e[owner]
    orig = d["eoff"][idx][owner] + np.where(is_rev, pos - e[owner], pos)
    u, v = d["u"][orig], d["v"][orig]
    off = new_noff[:-1][owner]
    src = np.where(is_rev, v, u) + off
    dst = np.where(is_rev, u, v) + off
    return dict(src=src, dst=dst, rev=is_rev, vl=d["vl"][node_src], el=d["el"][orig], n=n, e=e2,
                num_nodes=int(new_noff[-1]), node_graph=np.repeat(np.arange(len(idx)), n))


def collate(ds, idx):
    return dict(p=_collate_side(ds.p, idx), g=_collate_side(ds.g, idx), y=ds.counts[idx])


def to_device(batch, device, pinned=None):
    """One H2D copy per array (non-blocking from pinned staging); returns (pattern, graph, target, bytes)."""
    nbytes = 0

    def put(a, dtype):
        nonlocal nbytes
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        t = t.pin_memory()
        nbytes += t.numel() * t.element_size()
        return t.to(device, non_blocking=True)

    out = []
    for side in ("p", "g"):
        b = batch[side]
        g = DMPGraph(put(b["src"], torch.int64), put(b["dst"], torch.int64), b["num_nodes"])
        g.edata[REVFLAG] = put(b["rev"], torch.bool)
        g.ndata[NODELABEL] = put(b["vl"], torch.int64)
        g.edata[EDGELABEL] = put(b["el"], torch.int64)
        g.ndata["graph_id"] = put(b["node_graph"], torch.int64)
        g._batch_num_nodes = put(b["n"], torch.int64)
        g._batch_num_edges = put(b["e"], torch.int64)
        g.rev_layout_hint = "general"   # per-graph [fwd|rev] blocks
        g.validate_plan = False         # synthetic ids are in range: skip the one D2H check per plan
        out.append(g)
    y = put(batch["y"], torch.float32)
    return out[0], out[1], y, nbytes


class DevicePairDataset:
    """The flat arrays of a SyntheticPairDataset resident in HBM: batches are then built ON the device
    (`batch_on_device`, row N1) instead of collating on the host and copying every step."""

    def __init__(self, ds, device):
        self.device, self.num, self.cfg = torch.device(device), ds.num, ds.cfg

        def put(a, dtype=torch.int64):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(device)

        self.sides = {}
        for side in ("p", "g"):
            d = getattr(ds, side)
            self.sides[side] = dict(noff=put(d["noff"]), eoff=put(d["eoff"]), u=put(d["u"]), v=put(d["v"]),
                                    vl=put(d["vl"]), el=put(d["el"]), n_host=np.asarray(d["n"], np.int64),
                                    e_host=np.asarray(d["e"], np.int64))
        self.counts = put(ds.counts, torch.float32)


def batch_on_device(dev_side, sel, sel_host, add_reversed=True, pad_to=None):
    """Disjoint union of graphs `sel` (device int64 [B]; `sel_host` = the same ids on the host, used only to size
    the outputs) with per-graph reversed-edge blocks, written by two kernels (dmp_batch_offsets + dmp_batch_fill):
    `dgl.batch` + `add_reversed_edges` semantics (dataset.py:1321-1328, train.py:299-327), bit-identical to the host
    `_collate_side`.  Returns a DMPGraph with the same frames `to_device` produces.

    pad_to = (nodes, edges): fixed output sizes (>= the real totals, nodes > real nodes when any edge is padded); the
    tail is dummy nodes (graph id B) and self-loops spread round-robin over them.  `sel_host` may then be None: the
    call needs nothing from the host and can be captured in a CUDA graph."""
    from . import _lib
    d = dev_side
    dev = d["u"].device
    B = int(sel.numel())
    if pad_to is None:
        tn = int(d["n_host"][sel_host].sum())
        te = int(d["e_host"][sel_host].sum()) * (2 if add_reversed else 1)
    else:
        tn, te = int(pad_to[0]), int(pad_to[1])
    i64 = dict(dtype=torch.int64, device=dev)
    new_noff, new_eoff = torch.empty(B + 1, **i64), torch.empty(B + 1, **i64)
    st = _lib.stream_ptr(dev)
    _lib.call("dmp_batch_offsets", dev, _lib.ptr(sel), B, _lib.ptr(d["noff"]), _lib.ptr(d["eoff"]), int(add_reversed),
              _lib.ptr(new_noff), _lib.ptr(new_eoff), st, tag="batch_offsets")
    src, dst = torch.empty(te, **i64), torch.empty(te, **i64)
    rev = torch.empty(te, dtype=torch.uint8, device=dev)
    vl, el, ng = torch.empty(tn, **i64), torch.empty(te, **i64), torch.empty(tn, **i64)
    _lib.call("dmp_batch_fill", dev, _lib.ptr(sel), B, _lib.ptr(d["noff"]), _lib.ptr(d["eoff"]), _lib.ptr(d["u"]),
              _lib.ptr(d["v"]), _lib.ptr(d["vl"]), _lib.ptr(d["el"]), _lib.ptr(new_noff), _lib.ptr(new_eoff), tn, te,
              int(add_reversed), int(pad_to is not None), _lib.ptr(src), _lib.ptr(dst), _lib.ptr(rev), _lib.ptr(vl),
              _lib.ptr(el), _lib.ptr(ng), None, st, tag="batch_fill")
    g = DMPGraph(src, dst, tn)
    g.edata[REVFLAG] = rev.view(torch.bool)
    g.ndata[NODELABEL] = vl
    g.edata[EDGELABEL] = el
    g.ndata["graph_id"] = ng
    g._batch_num_nodes = new_noff[1:] - new_noff[:-1]
    g._batch_num_edges = new_eoff[1:] - new_eoff[:-1]
    g.rev_layout_hint = "general"   # per-graph [fwd|rev] blocks
    g.validate_plan = False         # ids are in range by construction: skip the one D2H check per plan
    return g


def collate_on_device(dds, idx):
    """(pattern, graph, target, h2d bytes) for pairs `idx` (host array): the only host->device traffic is the id list."""
    sel_host = np.asarray(idx, np.int64)
    sel = torch.from_numpy(sel_host).pin_memory().to(dds.device, non_blocking=True)
    p = batch_on_device(dds.sides["p"], sel, sel_host)
    g = batch_on_device(dds.sides["g"], sel, sel_host)
    return p, g, dds.counts[sel], sel_host.nbytes


class SubgraphCountingModel(nn.Module):
    def __init__(self, hidden, num_vlabels, num_elabels, num_layers=3, act_func="leaky_relu"):
        super().__init__()
        self.hidden, self.num_vlabels = hidden, num_vlabels
        self.vl_emb = nn.Embedding(num_vlabels, hidden)
        self.el_emb = nn.Embedding(2 * num_elabels, hidden)   # reversed edges: label += max label (train.py:310)
        self.num_elabels = num_elabels
        self.rep = DMPNNRepNet(hidden, num_layers=num_layers, rep_act_func=act_func, rep_dmpnn_num_mlp_layers=2,
                               rep_dmpnn_batch_norm=False)
        self.p_fc = nn.Linear(hidden, hidden)
        self.g_fc = nn.Linear(hidden, hidden)
        self.pred_fc1 = nn.Linear(4 * hidden + 4, hidden)
        self.pred_fc2 = nn.Linear(hidden + 4, 1)
        self.act = nn.LeakyReLU(1 / 5.5)

    def _labels(self, g):
        return g.ndata[NODELABEL], g.edata[EDGELABEL] + g.edata[REVFLAG].long() * self.num_elabels

    def _embed(self, g):
        vl, el = self._labels(g)
        return label_embedding(self.vl_emb.weight, vl), label_embedding(self.el_emb.weight, el)

    def _pool(self, x, g):
        """Per-graph sum of node rows: nodes of a graph are contiguous -> sorted segments, no atomics."""
        n = g.batch_num_nodes()
        indptr = torch.zeros(n.numel() + 1, dtype=torch.int32, device=x.device)
        indptr[1:] = torch.cumsum(n, 0)
        return _SegSum.apply(x, indptr, g.ndata["graph_id"])

    def forward(self, pattern, graph, union=None):
        """`union` (optional, from `union_graph(pattern, graph)`): run the SHARED layers once on the disjoint union of
        the pattern batch and the graph batch instead of twice (share_rep_net, dmpnn.py:187-188) -- same arithmetic
        per node/edge (the pattern side simply has gate 1), half the kernel launches."""
        bsz = pattern.batch_num_nodes().numel()
        # ScalarFilter-style gate (filter.py:6-16, basemodel.py:1394-1423): a graph node/edge passes if its
        # label occurs in the paired pattern
        # (+1 row: padded batches park their dummy nodes in graph id `bsz`, train_step.batch_on_device(pad_to=...))
        pres_v = torch.zeros((bsz + 1, self.num_vlabels), device=graph.device)
        pres_v.index_put_((pattern.ndata["graph_id"], pattern.ndata[NODELABEL]),
                          torch.ones((), device=graph.device))     # (a device scalar: capturable in a CUDA graph)
        v_gate = pres_v[graph.ndata["graph_id"], graph.ndata[NODELABEL]]
        if union is None:
            p_v, p_e = self._embed(pattern)
            g_v, g_e = self._embed(graph)
            p_v, p_e = self.rep.get_pattern_rep(pattern, p_v, p_e)
            g_v, g_e = self.rep.get_graph_rep(graph, g_v, g_e, v_gate=v_gate)
        else:
            # one lookup per table over the concatenated labels: the union's feature rows come out in place
            (p_vl, p_el), (g_vl, g_el) = self._labels(pattern), self._labels(graph)
            np_ = p_vl.shape[0]
            gate = torch.cat([torch.ones(np_, device=v_gate.device), v_gate])
            u_v, u_e = self.rep.get_graph_rep(union, label_embedding(self.vl_emb.weight, torch.cat([p_vl, g_vl])),
                                              label_embedding(self.el_emb.weight, torch.cat([p_el, g_el])), v_gate=gate)
            p_v, g_v = u_v[:np_], u_v[np_:]
        p = self._pool(row_linear(p_v, self.p_fc), pattern)
        g = self._pool(row_linear(g_v, self.g_fc), graph)
        pl = pattern.batch_num_nodes().float().view(-1, 1)
        gl = graph.batch_num_nodes().float().view(-1, 1)
        extra = [pl, gl, 1.0 / pl, 1.0 / gl]
        y = self.act(self.pred_fc1(torch.cat([p, g, g - p, g * p] + extra, dim=1)))
        return self.pred_fc2(torch.cat([y] + extra, dim=1)).view(-1)


class _LabelEmbedding(torch.autograd.Function):
    """`weight[labels]` (nn.Embedding of node / edge labels, train.py:299-350) with a backward that suits a vocabulary of
    a few dozen labels and 10^5 rows: dW = onehot(labels)^T @ g as ONE K = rows reduction on the tensor cores
    (exact products with 0/1, deterministic) instead of torch's sort + unique-by-key + segment kernels (8 radix passes per
    lookup: 13 % of the small-graph training step)."""

    @staticmethod
    def forward(ctx, weight, labels):
        ctx.save_for_backward(labels)
        ctx.vocab = weight.shape[0]
        return weight.index_select(0, labels)

    @staticmethod
    def backward(ctx, g):
        from .fused import _tnmm
        (labels,) = ctx.saved_tensors
        L = ctx.vocab
        width = 64 if L <= 64 else 128 if L <= 128 else L      # the reduction kernel takes 64 / 128 columns
        onehot = torch.zeros((labels.numel(), width), dtype=g.dtype, device=g.device)
        onehot.scatter_(1, labels.view(-1, 1), 1.0)
        dW = _tnmm(onehot, g.contiguous()) if g.is_cuda else onehot.t() @ g
        return dW[:L], None


class _RowLinear(torch.autograd.Function):
    """nn.Linear over node rows (the predict net's p_layer / g_layer, pred.py:95-112) on the layer's own projection
    kernels: bias in the GEMM epilogue, dW and db from ONE K = rows reduction (torch: addmm + mm + a 64 x 64 sgemm whose
    single CTA walks all 20 k rows -- 82 us -- + a column sum)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from .fused import _rowmm
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        return _rowmm(x, weight, bias=bias)

    @staticmethod
    def backward(ctx, g):
        from .fused import _rowmm, _tnmm
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        dx = _rowmm(g, weight.t().contiguous()) if ctx.needs_input_grad[0] else None
        dw, db, _ = _tnmm(g, x, colsum_x=True)
        return dx, dw, (db if ctx.needs_input_grad[2] else None)


def row_linear(x, linear):
    if not x.is_cuda:
        return linear(x)
    return _RowLinear.apply(x, linear.weight, linear.bias)


def label_embedding(weight, labels):
    return _LabelEmbedding.apply(weight, labels)


def union_graph(pattern, graph):
    """Disjoint union [pattern batch | graph batch] for the shared representation layers (one plan, one call)."""
    np_ = pattern.number_of_nodes()
    ps, pd = pattern.all_edges()
    gs, gd = graph.all_edges()
    u = DMPGraph(torch.cat([ps, gs + np_]), torch.cat([pd, gd + np_]), np_ + graph.number_of_nodes())
    u.edata[REVFLAG] = torch.cat([pattern.edata[REVFLAG], graph.edata[REVFLAG]])
    u.rev_layout_hint = "general"
    u.validate_plan = False
    return u


class _SegSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, indptr, row_segment):
        eid = torch.arange(x.shape[0], dtype=torch.int32, device=x.device)
        ctx.save_for_backward(row_segment)
        return segment_reduce(indptr, eid, x.contiguous(), x.shape[1], tag="segment_reduce.pool")

    @staticmethod
    def backward(ctx, g):
        (row_segment,) = ctx.saved_tensors
        g = torch.cat([g, g.new_zeros((1, g.shape[1]))])     # dummy rows of a padded batch (segment id = #graphs): no gradient
        return g[row_segment], None, None


def train_step(model, optimizer, pattern, graph, target, *, world=1, clip=10.0, fuse_batches=True):
    """forward -> MSE -> backward -> (gradient all-reduce) -> clip -> optimizer.  Returns the loss tensor (device)."""
    optimizer.zero_grad(set_to_none=True)
    pred = model(pattern, graph, union=union_graph(pattern, graph) if fuse_batches else None)
    loss = torch.mean((pred - target) ** 2)
    loss.backward()
    if world > 1:
        from .parallel import allreduce_gradients
        allreduce_gradients(model.parameters(), average=True)
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip, foreach=True)
    optimizer.step()
    return loss


class GraphedTrainStep:
    """The whole training step -- batch construction on the device (row N1), plan build, 3 shared DMP layers on the union
    of pattern and graph batch, head, MSE, backward, clip, AdamW -- captured ONCE in a CUDA graph and replayed per step:
    the only host work per step is drawing the pair ids and one 4 KB copy.

    Batches differ in their node / edge totals, so the union is padded to a fixed bucket (mean + 6 sigma of the dataset's
    per-batch totals, at least one dummy node) with dummy nodes and self-loop dummy edges that no real row ever
    sees and no pooling includes; a batch that does not fit the bucket (never, at 6 sigma) runs the eager step instead.
    Needs an optimizer built with `capturable=True`."""

    def __init__(self, model, optimizer, dds, pairs, clip=10.0, warmup=3, world=1):
        from . import _lib
        self.model, self.opt, self.dds, self.pairs, self.clip = model, optimizer, dds, int(pairs), clip
        self.world = world            # > 1: the flat-buffer gradient all-reduce (NCCL) is captured with the step
        dev = dds.device
        self.pad = {}
        for side in ("p", "g"):
            d = dds.sides[side]
            n, e2 = d["n_host"].astype(np.float64), 2.0 * d["e_host"].astype(np.float64)
            pn = pairs * n.mean() + 6.0 * np.sqrt(pairs) * n.std() + 1
            pe = pairs * e2.mean() + 6.0 * np.sqrt(pairs) * e2.std()
            self.pad[side] = (int(np.ceil(pn / 64.0)) * 64, int(np.ceil(pe / 64.0)) * 64)
        self.sel = torch.zeros(self.pairs, dtype=torch.int64, device=dev)
        self.sel_pinned = torch.zeros(self.pairs, dtype=torch.int64).pin_memory()
        self.graph = None
        self.loss = None
        self.fallbacks = 0
        # warm-up on a side stream (lazy initialisations: kernel attributes, workspaces, optimizer state), then capture
        rng = np.random.Generator(np.random.PCG64(0))
        side_stream = torch.cuda.Stream(device=dev)
        side_stream.wait_stream(torch.cuda.current_stream(dev))
        self.warmup_ids = []          # the warm-up steps are real optimizer steps: recorded so that a run can be reproduced
        with torch.cuda.stream(side_stream):
            for _ in range(max(int(warmup), 1)):     # >= 1: optimizer state must exist before capture, or its
                ids = np.sort(rng.choice(dds.num, size=self.pairs, replace=False))   # zero-initialisation is replayed
                self.warmup_ids.append(ids)
                self._load(ids)
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side_stream)
        torch.cuda.synchronize(dev)
        self._load(np.sort(rng.choice(dds.num, size=self.pairs, replace=False)))
        g = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(g):
            self.loss = self._body()
        self.graph = g
        self.dmp_kernels_in_graph = _lib.LAUNCHES - l0     # C-ABI launches recorded into the graph (replayed, not re-issued)

    def _load(self, idx):
        self.sel_pinned.copy_(torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)))
        self.sel.copy_(self.sel_pinned, non_blocking=True)

    def _body(self):
        p = batch_on_device(self.dds.sides["p"], self.sel, None, pad_to=self.pad["p"])
        g = batch_on_device(self.dds.sides["g"], self.sel, None, pad_to=self.pad["g"])
        y = self.dds.counts[self.sel]
        self.opt.zero_grad(set_to_none=True)
        pred = self.model(p, g, union=union_graph(p, g))
        loss = torch.mean((pred - y) ** 2)
        loss.backward()
        if self.world > 1:
            from .parallel import allreduce_gradients
            allreduce_gradients(self.model.parameters(), average=True)
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip, foreach=True)
        self.opt.step()
        return loss

    def fits(self, idx):
        for side in ("p", "g"):
            d = self.dds.sides[side]
            if int(d["n_host"][idx].sum()) + 1 > self.pad[side][0] or 2 * int(d["e_host"][idx].sum()) > self.pad[side][1]:
                return False
        return True

    def __call__(self, idx):
        """One training step on pairs `idx` (host array of length `pairs`); returns the loss tensor (device)."""
        if not self.fits(idx):
            self.fallbacks += 1
            p, g, y, _ = collate_on_device(self.dds, idx)
            return train_step(self.model, self.opt, p, g, y, world=self.world, clip=self.clip)
        self._load(idx)
        self.graph.replay()
        return self.loss
