"""Device-side graph plan: everything the sparse kernels need to know about one (batched) graph.

Built once per graph by `dmp_plan_build` (include/dmp_b200.h) and cached on the graph object, because
the same graph passes through every layer and through forward and backward.
Replaces DGL's lazy COO->CSC conversion behind `fn.sum` and `graph.out_degrees()`
(SubgraphCountingMatching/models/dmpnn.py:92,100-101,163).
"""
import ctypes

import torch

from . import _lib

_LUT_LEN = 4096
# A segment longer than LONG_SEGMENT rows switches the plan's reductions to parallel chunks of LONG_CHUNK rows
# (functional.segment_reduce_two_level: deterministic; shorter segments keep the strictly sequential, DGL-identical order).
# Detected from the validation read (one D2H per plan); with validate=False set `graph.long_segment_chunk` yourself.
LONG_SEGMENT, LONG_CHUNK = 65536, 4096
_lut_cache = {}


def _coef_lut(device):
    """coef_lut[d] = 2*(1+log2(1+d)) evaluated with the host maths library (torch CPU), exactly the
    expression of dmpnn.py:144-146, so that small degrees carry the reference's bits."""
    key = str(device)
    if key not in _lut_cache:
        d = torch.arange(_LUT_LEN, dtype=torch.int64).float()
        lut = 2 * (1 + (1 + d).log2())
        _lut_cache[key] = lut.to(device)
    return _lut_cache[key]


class DMPPlan:
    """Index structures of one graph (all int32 on device; layout in DESIGN.md)."""

    def __init__(self, src, dst, num_nodes, rev=None, out_deg=None, validate=True, rev_layout=None, rev_split=None):
        _lib.require_cuda(src, dst, rev, out_deg)
        lib = _lib.load()
        dev = src.device
        self.device = dev
        self.N = int(num_nodes)
        self.E = int(src.numel())
        N, E = self.N, self.E
        src = src.contiguous().to(torch.int64)
        dst = dst.contiguous().to(torch.int64)
        self.rev = None
        if rev is not None:
            self.rev = rev.contiguous().view(-1).to(torch.uint8)
            if self.rev.numel() != E:
                raise ValueError("rev flag must have one entry per edge")
        if out_deg is not None:
            out_deg = out_deg.contiguous().view(-1).to(torch.int64)
            if out_deg.numel() != N:
                raise ValueError("out_deg must have one entry per node")
        i32 = dict(dtype=torch.int32, device=dev)
        self.dst32 = torch.empty(E, **i32)
        self.a32 = torch.empty(E, **i32)
        self.b32 = torch.empty(E, **i32)
        self.csc_indptr = torch.empty(N + 1, **i32)
        self.csc_eid = torch.empty(E, **i32)
        self.a_indptr = torch.empty(N + 1, **i32)
        self.a_eid = torch.empty(E, **i32)
        self.b_indptr = torch.empty(N + 1, **i32)
        self.b_eid = torch.empty(E, **i32)
        self.out_deg = torch.empty(N, dtype=torch.int64, device=dev)
        self.coef = torch.empty(E, dtype=torch.float32, device=dev)
        status = torch.zeros(3, **i32)
        nbytes = ctypes.c_int64(0)
        _lib.check(lib.dmp_plan_workspace_bytes(N, E, ctypes.byref(nbytes)), "dmp_plan_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        lut = _coef_lut(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.dmp_plan_build(
                _lib.ptr(src), _lib.ptr(dst), _lib.ptr(self.rev), _lib.ptr(out_deg), N, E,
                _lib.ptr(lut), lut.numel(),
                _lib.ptr(self.dst32), _lib.ptr(self.a32), _lib.ptr(self.b32),
                _lib.ptr(self.csc_indptr), _lib.ptr(self.csc_eid), _lib.ptr(self.a_indptr), _lib.ptr(self.a_eid),
                _lib.ptr(self.b_indptr), _lib.ptr(self.b_eid), _lib.ptr(self.out_deg), _lib.ptr(self.coef),
                _lib.ptr(status), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "dmp_plan_build")
        # Layout of the reversed flags decides how the node-message projection is issued (layers.py):
        #   "none"    no flags                    -> one GEMM with W_in
        #   "halves"  [forward block | reversed block], split at `rev_split` (E/2 after add_reversed_edges on a
        #             single graph; any split point for a partition of such a graph) -> two GEMMs
        #   "general" anything else (e.g. batched graphs: per-graph blocks) -> one [E, 2H] two-branch GEMM
        if self.rev is None:
            self.rev_layout = "none"
        elif rev_layout is not None:
            self.rev_layout = rev_layout
        else:
            self.rev_layout = "general"
        self.rev_split = E // 2 if rev_split is None else int(rev_split)
        if validate:
            # one small D2H read per plan: endpoint range check (+ layout detection when not hinted)
            if self.rev is not None and rev_layout is None and E > 0:
                # [forward block | reversed block] <=> the flags are sorted; the split point is the forward count
                n_fwd = (self.rev == 0).sum()
                is_sorted = (self.rev[1:] >= self.rev[:-1]).all() if E > 1 else torch.ones((), dtype=torch.bool, device=dev)
                status[1] = torch.where(is_sorted, n_fwd + 1, torch.zeros_like(n_fwd)).to(torch.int32)
            if E > 0 and N > 0:
                # longest segment of the three segmentations: a hub of > LONG_SEGMENT rows serialises one lane group for
                # milliseconds (measured: 176 ms vs 2.6 ms on a Yelp-sized power-law graph) -> chunked reduction
                status[2] = torch.stack([(p[1:] - p[:-1]).max() for p in (self.csc_indptr, self.a_indptr, self.b_indptr)]).max()
            st = status.tolist()
            if st[0] != 0:
                raise ValueError("graph has an edge endpoint outside [0, num_nodes)")
            if st[1] != 0:
                self.rev_layout = "halves"
                self.rev_split = st[1] - 1
            self.max_segment = st[2]
        self._norm_perm = {}
        self._mirrored = None
        # graphs with hub nodes (power-law degree): the layer's three segment reductions cut segments longer than
        # `long_chunk` into parallel chunks (functional.segment_reduce_two_level: deterministic, but not DGL's strictly
        # sequential order on those hubs).  Automatic above LONG_SEGMENT rows, or `graph.long_segment_chunk = ...`
        self.long_chunk = LONG_CHUNK if getattr(self, "max_segment", 0) > LONG_SEGMENT else None
        del ws

    @property
    def mirrored_halves(self):
        """True when edge e + E/2 is the reverse of edge e for every e (one graph after the reversed-edge append,
        train.py:299-327): both rows then gather the same endpoint rows and `dmp_edge_update` fetches them once.
        Decided once per plan with two device-side comparisons (one small D2H read)."""
        if self._mirrored is None:
            h = self.rev_split
            self._mirrored = bool(self.rev_layout == "halves" and self.E > 0 and 2 * h == self.E
                                  and torch.equal(self.a32[:h], self.a32[h:]) and torch.equal(self.b32[:h], self.b32[h:]))
        return self._mirrored

    def norm_permuted(self, norm):
        """`norm` ([E] or [E,1]) re-ordered to CSC position order, cached per tensor version."""
        if norm is None:
            return None
        key = (norm.data_ptr(), norm._version, tuple(norm.shape))
        hit = self._norm_perm.get(key)
        if hit is None:
            flat = norm.detach().reshape(-1).contiguous().float()
            if flat.numel() != self.E:
                raise ValueError("edge_norm must have one entry per edge")
            out = torch.empty_like(flat)
            _lib.call("dmp_permute_edge_scalar", self.device, _lib.ptr(self.csc_eid), _lib.ptr(flat),
                      _lib.ptr(out), self.E, _lib.stream_ptr(self.device))
            self._norm_perm = {key: (flat, out)}
            hit = self._norm_perm[key]
        return hit

    def nbytes(self):
        t = [self.dst32, self.a32, self.b32, self.csc_indptr, self.csc_eid, self.a_indptr, self.a_eid,
             self.b_indptr, self.b_eid, self.out_deg, self.coef]
        return sum(x.numel() * x.element_size() for x in t)


def graph_arrays(graph):
    """(src, dst, num_nodes) of a DMPGraph or a DGLGraph, edge-id order."""
    src, dst = graph.all_edges(form="uv", order="eid")
    n = graph.num_nodes() if hasattr(graph, "num_nodes") else graph.number_of_nodes()
    return src, dst, int(n)


def get_plan(graph, rev_key, deg_key, validate=True):
    """Plan of `graph`, cached on the graph object and keyed by the identity of the tensors it was built
    from (edge list, reversed flag, caller-supplied out-degree: dmpnn.py:100-101 honours a pre-existing
    `ndata["out_deg"]` instead of recomputing it)."""
    src, dst, n = graph_arrays(graph)
    rev = graph.edata[rev_key] if rev_key in graph.edata else None
    deg = graph.ndata[deg_key] if deg_key in graph.ndata else None
    key = (src.data_ptr(), dst.data_ptr(), int(src.numel()), n,
           None if rev is None else (rev.data_ptr(), rev._version),
           None if deg is None else (deg.data_ptr(), deg._version))
    cache = getattr(graph, "_dmp_plans", None)
    if cache is None:
        cache = {}
        try:
            graph._dmp_plans = cache
        except AttributeError:  # foreign graph type that forbids attributes: no caching
            pass
    plan = cache.get(key)
    if plan is None:
        hint = getattr(graph, "rev_layout_hint", None)
        validate = getattr(graph, "validate_plan", validate)
        plan = DMPPlan(src, dst, n, rev=rev, out_deg=deg, validate=validate, rev_layout=hint)
        plan.long_chunk = getattr(graph, "long_segment_chunk", plan.long_chunk)
        plan._key_refs = (src, dst, rev, deg)   # keep the key tensors alive: a recycled address must not hit this plan
        cache.clear()
        cache[key] = plan
        if deg is None:
            # same side effect as the reference: the computed degrees stay in the node frame
            graph.ndata[deg_key] = plan.out_deg
            key2 = key[:5] + ((plan.out_deg.data_ptr(), plan.out_deg._version),)
            cache[key2] = plan
    return plan
