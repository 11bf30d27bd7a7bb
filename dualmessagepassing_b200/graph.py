"""Graph container the DMPNN layers consume, plus the reference's graph-construction helpers.

`DMPGraph` exposes the slice of the DGLGraph API that the reference layers and their callers use
(`ndata` / `edata` frames, `all_edges(form, order)`, `out_degrees()`, `add_edges`, `batch_num_*`),
so `layer(graph, node_feat, edge_feat)` reads the same whether `graph` is a real DGLGraph or a
`DMPGraph`.  Real DGL graphs are accepted by the layers through the same duck-typed calls.

Semantics follow:
  * `batch`               dgl.batch as used by SubgraphCountingMatching/dataset.py:1321-1328
  * `add_reversed_edges`  SubgraphCountingMatching/train.py:299-327, dataset.py:1522-1563
  * `build_graph_from_triplets`, `compute_edgenorm`
                          UnsupervisedNodeClassification/Model/DMPNN/src/utils.py:437-453,473-491
Everything here is torch tensor code and runs on whichever device the tensors live on.
"""
import torch

from .constants import EDGEID, EDGELABEL, REVFLAG


class _EdgeFrame(dict):
    """`edata` of a DMPGraph: a dict that forgets the graph's reversed-flag layout hint when a flag tensor is assigned,
    replaced or removed by the caller (a hint that outlives its flags would mis-plan the graph)."""
    _FLAG_KEYS = (REVFLAG, "is_rev")

    def __init__(self, graph):
        super().__init__()
        self._graph = graph

    def __setitem__(self, key, value):
        if key in self._FLAG_KEYS and self.get(key) is not value:
            self._graph.rev_layout_hint = None
        super().__setitem__(key, value)

    def __delitem__(self, key):
        if key in self._FLAG_KEYS:
            self._graph.rev_layout_hint = None
        super().__delitem__(key)

    def pop(self, key, *default):
        if key in self._FLAG_KEYS:
            self._graph.rev_layout_hint = None
        return super().pop(key, *default)


class DMPGraph:
    def __init__(self, src, dst, num_nodes, device=None):
        src = torch.as_tensor(src, dtype=torch.int64, device=device)
        dst = torch.as_tensor(dst, dtype=torch.int64, device=device)
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src and dst must be 1-D and of equal length")
        self._src, self._dst = src, dst
        self._n = int(num_nodes)
        self.ndata = {}
        self.rev_layout_hint = None  # "halves" when add_reversed_edges built the edge list; reset by _EdgeFrame
        self.edata = _EdgeFrame(self)
        self._batch_num_nodes = None
        self._batch_num_edges = None
        self._dmp_plans = {}      # plan cache, see plan.get_plan

    # ---- structure (DGL-compatible spellings) ------------------------------------------------------
    @property
    def device(self):
        return self._src.device

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.numel())

    num_edges = number_of_edges

    def all_edges(self, form="uv", order="eid"):
        if order not in ("eid", None):
            raise NotImplementedError("only edge-id order is kept")
        if form == "uv":
            return self._src, self._dst
        eid = torch.arange(self.number_of_edges(), device=self.device)
        if form == "eid":
            return eid
        return self._src, self._dst, eid

    edges = all_edges

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def batch_num_nodes(self):
        if self._batch_num_nodes is None:
            return torch.tensor([self._n], device=self.device)
        return self._batch_num_nodes

    def batch_num_edges(self):
        if self._batch_num_edges is None:
            return torch.tensor([self.number_of_edges()], device=self.device)
        return self._batch_num_edges

    @property
    def batch_size(self):
        return int(self.batch_num_nodes().numel())

    def local_var(self):
        return self

    def add_edges(self, u, v, data=None):
        """Append edges at ids E..E+k-1; features absent on either side are zero-filled (DGL)."""
        u = torch.as_tensor(u, dtype=torch.int64, device=self.device)
        v = torch.as_tensor(v, dtype=torch.int64, device=self.device)
        e_old, k = self.number_of_edges(), int(u.numel())
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        data = data or {}
        for key in set(self.edata) | set(data):
            new = data.get(key)
            old = self.edata.get(key)
            if old is None:
                old = torch.zeros((e_old,) + tuple(new.shape[1:]), dtype=new.dtype, device=self.device)
            if new is None:
                new = torch.zeros((k,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)
            self.edata[key] = torch.cat([old, new.to(self.device)])
        self._dmp_plans.clear()
        self.rev_layout_hint = None

    def to(self, device, non_blocking=False):
        g = DMPGraph(self._src.to(device, non_blocking=non_blocking),
                     self._dst.to(device, non_blocking=non_blocking), self._n)
        g.ndata = {k: v.to(device, non_blocking=non_blocking) for k, v in self.ndata.items()}
        for k, v in self.edata.items():
            g.edata[k] = v.to(device, non_blocking=non_blocking)
        if self._batch_num_nodes is not None:
            g._batch_num_nodes = self._batch_num_nodes.to(device)
            g._batch_num_edges = self._batch_num_edges.to(device)
        g.rev_layout_hint = self.rev_layout_hint
        return g

    def __repr__(self):
        return "DMPGraph(num_nodes=%d, num_edges=%d, ndata=%s, edata=%s)" % (
            self._n, self.number_of_edges(), sorted(self.ndata), sorted(self.edata))


def batch(graphs):
    """Disjoint union with prefix-sum node/edge offsets, frames concatenated in list order."""
    if not graphs:
        raise ValueError("batch() of an empty list")
    dev = graphs[0].device
    nn = torch.tensor([g.number_of_nodes() for g in graphs], dtype=torch.int64)
    ne = torch.tensor([g.number_of_edges() for g in graphs], dtype=torch.int64)
    n_off = torch.cumsum(nn, 0) - nn
    src = torch.cat([g._src + int(o) for g, o in zip(graphs, n_off)])
    dst = torch.cat([g._dst + int(o) for g, o in zip(graphs, n_off)])
    out = DMPGraph(src, dst, int(nn.sum()), device=dev)
    for key in graphs[0].ndata:
        out.ndata[key] = torch.cat([g.ndata[key] for g in graphs])
    for key in graphs[0].edata:
        out.edata[key] = torch.cat([g.edata[key] for g in graphs])
    out._batch_num_nodes = nn.to(dev)
    out._batch_num_edges = ne.to(dev)
    return out


def add_reversed_edges(graph, max_num_edges=None, max_edge_label=None):
    """Append (v,u) for every (u,v) with is_reversed=1 (train.py:299-313). In place; returns graph."""
    if REVFLAG in graph.edata:
        return graph
    e0 = graph.number_of_edges()
    u, v = graph.all_edges(form="uv", order="eid")
    data = {REVFLAG: torch.ones((e0,), dtype=torch.bool, device=graph.device)}
    if max_num_edges is not None:
        data[EDGEID] = torch.arange(max_num_edges, max_num_edges + e0, device=graph.device)
    if max_edge_label is not None and EDGELABEL in graph.edata:
        data[EDGELABEL] = graph.edata[EDGELABEL] + max_edge_label
    graph.add_edges(v, u, data=data)
    graph.rev_layout_hint = "halves"
    return graph


def compute_edgenorm(graph, norm="in"):
    """Per-edge normaliser of the UNC pipeline (what `utils.py:437-453` hands to the layers as `edge_norm`): the
    reciprocal of the destination's in-degree ("in"), of the source's out-degree ("out"), or of the geometric mean of
    the two ("both"), shape [E,1].  A degree of 0 cannot occur on an existing edge's own endpoint for "in"/"out"; for
    caller-supplied degree frames that do contain zeros the non-finite entries take the smallest finite value."""
    u, v = graph.all_edges(form="uv", order="eid")
    deg_in = graph.ndata["in_deg"] if "in_deg" in graph.ndata else graph.ndata.setdefault("in_deg", graph.in_degrees())
    deg_out = graph.ndata["out_deg"] if "out_deg" in graph.ndata else graph.ndata.setdefault("out_deg", graph.out_degrees())
    if norm == "in":
        scale = deg_in[v].float()
    elif norm == "out":
        scale = deg_out[u].float()
    elif norm == "both":
        scale = (deg_out[u].float() * deg_in[v].float()).sqrt()
    else:
        raise ValueError("norm must be 'in', 'out' or 'both', got %r" % (norm,))
    w = (1.0 / scale).unsqueeze(-1)
    bad = ~torch.isfinite(w)
    if bool(bad.any()):
        w = torch.where(bad, w[~bad].min() if bool((~bad).any()) else torch.zeros((), device=w.device), w)
    return w


def build_graph_from_triplets(num_nodes, num_rels, triplets, device=None):
    """utils.py:473-491: sort triplets by (src, dst, rel); forward block, then reversed block with
    `type += num_rels`; `norm` = 1 / in-degree of the destination."""
    t = torch.as_tensor(triplets, dtype=torch.int64)
    key = (t[:, 0] * num_nodes + t[:, 2]) * max(int(num_rels), 1) + t[:, 1]
    t = t[torch.argsort(key, stable=True)]
    src = torch.cat([t[:, 0], t[:, 2]])
    dst = torch.cat([t[:, 2], t[:, 0]])
    g = DMPGraph(src, dst, num_nodes, device=device)
    g.edata["type"] = torch.cat([t[:, 1], t[:, 1] + num_rels]).to(g.device)
    g.edata["norm"] = compute_edgenorm(g)
    return g
