"""Minimal stand-in for the slice of DGL that the reference DMPNN layers touch.

TEST INFRASTRUCTURE ONLY.  `dgl` is not installable in this image (no network), but the
reference's own layer classes (`SubgraphCountingMatching/models/dmpnn.py:16-176`,
`UnsupervisedNodeClassification/Model/DMPNN/src/model.py:117-280`) are plain PyTorch
apart from five DGL entry points.  This shim restates those five entry points with the
semantics listed in SURVEY.md Appendix B so that the UNMODIFIED reference classes can be
imported from /root/reference and executed on CPU to produce golden vectors
(`make_golden.py`).  It is never imported by the product package.

Semantics restated (DGL >= 0.6 public behaviour):
  * `g.update_all(msg_udf, fn.sum(msg, out), node_udf)`: the message UDF sees every edge in
    edge-id order; `edges.src[k]` / `edges.dst[k]` are row gathers of `ndata[k]` by the
    edge's source / destination; `edges.data` IS the edge frame (writes persist).  The sum
    reducer adds each destination's messages sequentially in ascending edge id (stable
    COO->CSC order); nodes without in-edges get zeros.  The node UDF then sees all nodes.
  * `g.apply_edges(udf)`: same EdgeBatch; returned dict is written to `edata`.
  * `g.out_degrees()` / `g.in_degrees()`: int64 counts over the current edge set.
  * `g.add_edges(u, v, data)`: append at ids E..E+k-1, features missing on old edges are
    zero-filled (this is how `is_reversed` becomes False on the original edges).
  * `dgl.batch(list)`: disjoint union with node/edge offsets = prefix sums, frames
    concatenated in list order.
"""
import sys
import types
import collections.abc

import torch


class _Frame(dict):
    pass


class _Gather:
    """edges.src / edges.dst: lazy row gather of a node frame."""

    def __init__(self, frame, index):
        self._frame = frame
        self._index = index

    def __getitem__(self, key):
        return self._frame[key][self._index]

    def __contains__(self, key):
        return key in self._frame


class _EdgeBatch:
    def __init__(self, g):
        self.src = _Gather(g.ndata, g._src)
        self.dst = _Gather(g.ndata, g._dst)
        self.data = g.edata  # the frame itself: UDF writes persist


class _NodeBatch:
    def __init__(self, g):
        self.data = g.ndata


class ShimGraph:
    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src, dtype=torch.int64)
        self._dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self.ndata = _Frame()
        self.edata = _Frame()
        self.batch_num_nodes_ = [self._n]
        self.batch_num_edges_ = [int(self._src.numel())]

    # ---- structure -------------------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.numel())

    num_edges = number_of_edges

    def all_edges(self, form="uv", order="eid"):
        assert order == "eid"
        if form == "uv":
            return self._src, self._dst
        eid = torch.arange(self.number_of_edges())
        return self._src, self._dst, eid

    def edges(self, form="uv", order="eid"):
        return self.all_edges(form, order)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def add_edges(self, u, v, data=None):
        u = torch.as_tensor(u, dtype=torch.int64)
        v = torch.as_tensor(v, dtype=torch.int64)
        e_old = self.number_of_edges()
        k = int(u.numel())
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        data = data or {}
        for key in set(self.edata) | set(data):
            if key in self.edata:
                old = self.edata[key]
            else:
                new = data[key]
                old = torch.zeros((e_old,) + tuple(new.shape[1:]), dtype=new.dtype)
            if key in data:
                new = data[key]
            else:
                new = torch.zeros((k,) + tuple(old.shape[1:]), dtype=old.dtype)
            self.edata[key] = torch.cat([old, new])
        self.batch_num_edges_ = [self.number_of_edges()]

    def batch_num_nodes(self):
        return torch.tensor(self.batch_num_nodes_)

    def batch_num_edges(self):
        return torch.tensor(self.batch_num_edges_)

    # ---- message passing -------------------------------------------------------------
    def update_all(self, message_func, reduce_func, apply_node_func=None):
        kind, msg_key, out_key = reduce_func
        assert kind == "sum"
        msgs = message_func(_EdgeBatch(self))
        m = msgs[msg_key]
        agg = torch.zeros((self._n,) + tuple(m.shape[1:]), dtype=m.dtype)
        # sequential accumulate in edge-id order == per-destination CSC order
        agg = agg.index_add(0, self._dst, m)
        self.ndata[out_key] = agg
        if apply_node_func is not None:
            self.ndata.update(apply_node_func(_NodeBatch(self)))

    def apply_edges(self, func):
        self.edata.update(func(_EdgeBatch(self)))


def batch(graphs):
    n_off, srcs, dsts = 0, [], []
    for g in graphs:
        srcs.append(g._src + n_off)
        dsts.append(g._dst + n_off)
        n_off += g._n
    out = ShimGraph(torch.cat(srcs), torch.cat(dsts), n_off)
    for key in graphs[0].ndata:
        out.ndata[key] = torch.cat([g.ndata[key] for g in graphs])
    for key in graphs[0].edata:
        out.edata[key] = torch.cat([g.edata[key] for g in graphs])
    out.batch_num_nodes_ = [g._n for g in graphs]
    out.batch_num_edges_ = [g.number_of_edges() for g in graphs]
    return out


def install():
    """Register stub modules so the reference sources import in this image."""

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "dgl" in sys.modules and getattr(sys.modules["dgl"], "_is_dmp_shim", False):
        return sys.modules["dgl"]
    dgl = stub("dgl", DGLGraph=ShimGraph, batch=batch, __version__="0.6.1", _is_dmp_shim=True)
    dgl.function = stub("dgl.function", sum=lambda msg, out: ("sum", msg, out))
    dgl.nn = stub("dgl.nn")
    dgl.nn.pytorch = stub("dgl.nn.pytorch", RelGraphConv=object)
    stub("igraph", Graph=object)
    stub("tensorboardX", SummaryWriter=object)
    # removed from torch >= 2.0; only the container helpers import it
    stub("torch._six", container_abcs=collections.abc, string_classes=(str,), int_classes=(int,))
    return dgl
