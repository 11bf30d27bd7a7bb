"""Multi-GPU execution of the DMPNN layer. One process per GPU, torch.distributed.

Two cases, matching how the path shards:

1. Batches of pattern/graph pairs (configs 1-3) are independent units: every rank runs the layers on its
   own batch and the only exchange is the gradient all-reduce -> `allreduce_gradients` (one flat fp32
   buffer per step: the whole model is < 1 M parameters, so this is latency-bound by design).

2. One large graph (config 5) is partitioned by DESTINATION-node range: rank r owns nodes
   [r*N/P, (r+1)*N/P) and every edge whose destination it owns, in global edge-id order -- so the
   node aggregation (`fn.sum`, dmpnn.py:92,163) is purely local and keeps the single-GPU summation
   order bit for bit.  What an edge needs from elsewhere is the state of its SOURCE node: for an
   Erdos-Renyi graph a rank's edges touch almost every node, so the halo is the whole table and the
   exchange is an all-gather of the owned node rows forward and a reduce-scatter of the node-gradient
   partial sums backward (NCCL over NVLink/NVSwitch; edge states and their gradients never move).
   Weight gradients are partial sums over the local edges/nodes and are all-reduced once per layer call.

The reference has no distributed code at all; this module has no counterpart there.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .constants import LEAKY_RELU_A
from .fused import _ACT, fused_dmp_layer, mlp_spec_and_tensors
from .plan import DMPPlan


# ---- collectives on row-partitioned matrices (work on NCCL; gloo fallbacks keep the host logic testable)
def all_gather_rows(x, group=None):
    world = dist.get_world_size(group)
    out = torch.empty((x.shape[0] * world,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def reduce_scatter_rows(x, group=None):
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = x.shape[0] // world
    if dist.get_backend(group) == "nccl":
        out = torch.empty((rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y[rank * rows:(rank + 1) * rows].clone()


class _Work:
    """Handle of an asynchronous collective that keeps its buffers alive until it has been waited for."""

    def __init__(self, work=None, keep=None):
        self.work, self.keep = work, keep

    def wait(self):
        if self.work is not None:
            self.work.wait()
        self.work = self.keep = None
        return True


def all_gather_rows_async(x, group=None):
    """(out, work): the all-gather is enqueued on the communicator's own stream; `work.wait()` makes the CURRENT stream
    wait for it.  Whatever the caller launches in between overlaps the transfer (NVLink/NVSwitch: no SM contention)."""
    world = dist.get_world_size(group)
    out = torch.empty((x.shape[0] * world,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    x = x.contiguous()
    work = dist.all_gather_into_tensor(out, x, group=group, async_op=True)
    return out, _Work(work, x)


def reduce_scatter_rows_async(x, group=None):
    """(out, work) -- see all_gather_rows_async; `x` must stay alive (and unmodified) until work.wait()."""
    world = dist.get_world_size(group)
    rows = x.shape[0] // world
    if dist.get_backend(group) == "nccl":
        out = torch.empty((rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = x.contiguous()
        work = dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return out, _Work(work, x)    # the input buffer must outlive the collective
    return reduce_scatter_rows(x, group), _Work()


def allreduce_gradients(params, group=None, average=True):
    """Data-parallel gradient exchange: ONE all-reduce of a flat fp32 buffer, then scatter back."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def allreduce_tensors_(tensors, group=None):
    """Sum a list of small tensors across ranks with one flat all-reduce (in place)."""
    ts = [t for t in tensors if t is not None]
    if not ts:
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


# ---- destination-range partition of one graph ---------------------------------------------------------------
def padded_num_nodes(num_nodes, world):
    return ((num_nodes + world - 1) // world) * world


def partition_by_destination(src, dst, rev, num_nodes, rank, world):
    """Edges owned by `rank` (destination in its node range), in global edge-id order.

    Returns dict(eids, src, dst, rev, n_lo, n_hi, num_nodes_padded, out_deg): ids stay GLOBAL; out_deg is
    the global out-degree (the degree term of dmpnn.py:144-146 counts edges owned by other ranks too)."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    npad = padded_num_nodes(num_nodes, world)
    per = npad // world
    n_lo, n_hi = rank * per, (rank + 1) * per
    mine = np.nonzero((dst >= n_lo) & (dst < n_hi))[0]
    out = dict(eids=mine, src=src[mine], dst=dst[mine], rev=None if rev is None else np.asarray(rev)[mine],
               n_lo=n_lo, n_hi=n_hi, num_nodes_padded=npad,
               out_deg=np.bincount(src, minlength=npad).astype(np.int64))
    return out


class PartitionedDMPLayer:
    """Runs a DMPLayer (2-layer MLP or MLP-less, no BatchNorm) on this rank's partition of one large graph.

    `runner(node_feat_local, edge_feat_local) -> (node_out_local, edge_out_local)`; differentiable; the
    parameter gradients it leaves in `.grad` are already summed over ranks."""

    def __init__(self, layer, src, dst, rev, num_nodes, rank, world, device, group=None):
        if layer.num_mlp_layers > 1 and layer.batch_norm:
            raise NotImplementedError("partitioned execution needs a BatchNorm-free DMPLayer (the batch statistics "
                                      "would have to be all-reduced across ranks)")
        if layer.act_func not in _ACT:
            raise NotImplementedError("activation %r" % layer.act_func)
        self.layer, self.rank, self.world, self.group, self.device = layer, rank, world, group, device
        part = partition_by_destination(src, dst, rev, num_nodes, rank, world)
        self.part = part
        self.n_lo, self.n_hi, self.N = part["n_lo"], part["n_hi"], part["num_nodes_padded"]
        self.local_N = self.n_hi - self.n_lo
        self.local_E = int(part["eids"].size)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        r = part["rev"]
        layout, split = None, None
        if r is not None and self.local_E > 0:
            # a destination-range slice of a [forward | reversed] edge list is again [forward | reversed]
            n_fwd = int((~r.astype(bool)).sum())
            if (not r[:n_fwd].any()) and r[n_fwd:].all():
                layout, split = "halves", n_fwd
        self.plan = DMPPlan(t(part["src"]), t(part["dst"]), self.N,
                            rev=None if r is None else t(r.astype(np.uint8)), out_deg=t(part["out_deg"]),
                            rev_layout=layout, rev_split=split)

    def __call__(self, node_feat_local, edge_feat_local):
        L = self.layer
        weights = (L.in_weight, L.out_weight, L.src_weight, L.dst_weight, L.nloop_weight, L.eloop_weight)
        return fused_dmp_layer(self.plan, node_feat_local, edge_feat_local, weights, L.nbias, L.ebias,
                               mlp_spec_and_tensors(L.nmlp), mlp_spec_and_tensors(L.emlp), act_func=L.act_func,
                               slope=LEAKY_RELU_A, order=_lib.ORDER_SCM, training=L.training,
                               part=(self.n_lo, self.n_hi, self.group))
