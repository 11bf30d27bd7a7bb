"""N>1 host logic on CPU: world_size-2 gloo. The CUDA kernels cannot run here, so the partitioned
ALGORITHM (destination-range partition + all-gather fwd + reduce-scatter bwd + weight-grad all-reduce,
dualmessagepassing_b200/parallel.py) is executed with the CPU oracle layer standing in for the kernels
and compared with the single-graph oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dmp_oracle, graph_oracle as go
from tests._cases import make_graph


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _collectives_worker(rank, world, port, q):
    from dualmessagepassing_b200 import parallel as par
    _setup(rank, world, port)
    x = torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * rank
    g = par.all_gather_rows(x)
    ok = g.shape == (6, 2) and torch.equal(g[3 * rank:3 * rank + 3], x) and float(g[3 * (1 - rank), 0]) == 100 * (1 - rank)
    full = torch.ones(4, 3) * (rank + 1)
    rs = par.reduce_scatter_rows(full)
    ok = ok and rs.shape == (2, 3) and torch.all(rs == 3)
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(rank + 1))
    q_ = torch.nn.Parameter(torch.zeros(2, 2))
    q_.grad = torch.full((2, 2), 10.0 * (rank + 1))
    par.allreduce_gradients([p, q_], average=True)
    ok = ok and torch.all(p.grad == 1.5) and torch.all(q_.grad == 15.0)
    # asynchronous forms used by the fused layer (the collective overlaps local work until .wait())
    g2, w = par.all_gather_rows_async(x)
    w.wait()
    ok = ok and torch.equal(g2, g)
    rs2, w = par.reduce_scatter_rows_async(full)
    w.wait()
    ok = ok and torch.equal(rs2, rs)
    a, b = torch.ones(3) * rank, torch.ones(2, 2)
    par.allreduce_tensors_([a, None, b])
    ok = ok and torch.all(a == 1) and torch.all(b == 2)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_collective_wrappers_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_collectives_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == {0: True, 1: True}


def test_partition_covers_edges_once_in_order():
    from dualmessagepassing_b200.parallel import partition_by_destination
    s, d, r = make_graph(seed=3, n=101, e0=700, rev="shuffled")
    seen = []
    for rank in range(4):
        part = partition_by_destination(s, d, r, 101, rank, 4)
        assert part["num_nodes_padded"] == 104 and part["n_hi"] - part["n_lo"] == 26
        assert np.all(np.diff(part["eids"]) > 0)  # global edge-id order is kept
        assert np.all((part["dst"] >= part["n_lo"]) & (part["dst"] < part["n_hi"]))
        assert np.array_equal(part["out_deg"][:101], go.out_degrees(s, 101))
        seen.append(part["eids"])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(s)))


def _partition_worker(rank, world, port, q):
    from dualmessagepassing_b200 import parallel as par
    _setup(rank, world, port)
    n, h = 40, 8
    s, d, r = make_graph(seed=11, n=n, e0=150, rev="halves", isolated=0)
    E = len(s)
    torch.manual_seed(1)
    import dualmessagepassing_b200 as dmp
    layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu")
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    xv, xe, gv, ge = torch.randn(n, h), torch.randn(E, h), torch.randn(n, h), torch.randn(E, h)

    part = par.partition_by_destination(s, d, r, n, rank, world)
    lo, hi, ids = part["n_lo"], part["n_hi"], torch.from_numpy(part["eids"])
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xv_l = xv[lo:hi].clone().requires_grad_(True)
    xe_l = xe[ids].clone().requires_grad_(True)

    class Gather(torch.autograd.Function):  # all-gather fwd / reduce-scatter bwd, as in fused.py
        @staticmethod
        def forward(ctx, x):
            return par.all_gather_rows(x)

        @staticmethod
        def backward(ctx, g):
            return par.reduce_scatter_rows(g)

    xv_full = Gather.apply(xv_l)
    nv, ne = dmp_oracle.dmp_layer(P, part["src"], part["dst"], n, xv_full, xe_l, rev=part["rev"],
                                  out_deg=torch.from_numpy(part["out_deg"][:n]), flavour="scm", act_func="leaky_relu")
    nv_l = nv[lo:hi]  # rows of other ranks' nodes are not this rank's output
    torch.autograd.backward((nv_l, ne), (gv[lo:hi], ge[ids]))
    grads = [P[k].grad for k in sorted(P)]
    par.allreduce_tensors_(grads)
    # numpy payloads: torch tensors in a Queue are passed by shared-memory handle and die with the worker
    q.put((rank, lo, hi, part["eids"], nv_l.detach().numpy(), ne.detach().numpy(), xv_l.grad.numpy(),
           xe_l.grad.numpy(), {k: P[k].grad.numpy() for k in sorted(P)}))
    dist.destroy_process_group()


def test_partitioned_algorithm_equals_single_graph_oracle():
    world, n, h = 2, 40, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_partition_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]

    s, d, r = make_graph(seed=11, n=n, e0=150, rev="halves", isolated=0)
    E = len(s)
    torch.manual_seed(1)
    import dualmessagepassing_b200 as dmp
    layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu")
    P = {k: v.clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xv, xe, gv, ge = torch.randn(n, h), torch.randn(E, h), torch.randn(n, h), torch.randn(E, h)
    xv.requires_grad_(True)
    xe.requires_grad_(True)
    nv, ne = dmp_oracle.dmp_layer(P, s, d, n, xv, xe, rev=r, flavour="scm", act_func="leaky_relu")
    torch.autograd.backward((nv, ne), (gv, ge))
    for rank, lo, hi, eids, nv_l, ne_l, gxv, gxe, gP in res:
        ids = torch.from_numpy(eids)
        nv_l, ne_l, gxv, gxe = (torch.from_numpy(x) for x in (nv_l, ne_l, gxv, gxe))
        gP = {k: torch.from_numpy(v) for k, v in gP.items()}
        # forward: same per-destination summation order => bit-identical to the single-graph oracle
        assert torch.equal(nv_l, nv.detach()[lo:hi]) and torch.equal(ne_l, ne.detach()[ids])
        torch.testing.assert_close(gxe, xe.grad[ids], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gxv, xv.grad[lo:hi], rtol=1e-5, atol=2e-6)
        for k, g in gP.items():
            # weight gradients are fp32 sums over all edges, split differently across ranks
            torch.testing.assert_close(g, P[k].grad, rtol=1e-4, atol=1e-5 * max(1.0, float(P[k].grad.abs().max())),
                                       msg=lambda m: k + m)


def test_partitioned_split_aggregation_is_local_and_bit_identical():
    """Aggregate-first node update under the destination-range partition: every rank reduces only the edges it owns
    (global edge-id order kept) and gets exactly the rows [n_lo, n_hi) of the single-graph [A_fwd | A_rev] sums."""
    from dualmessagepassing_b200.parallel import partition_by_destination
    from oracle import sparse_core as sc
    n, H, world = 203, 24, 4
    s, d, r = make_graph(seed=31, n=n, e0=1500, rev="halves")
    E = len(s)
    g = torch.Generator().manual_seed(1)
    X, norm = torch.randn(E, H, generator=g), torch.rand(E, generator=g)
    r8 = torch.from_numpy(r.astype(np.uint8))
    indptr, eid = sc.stable_segments(torch.from_numpy(d.astype(np.int32)), r8, n)
    want = sc.seg_reduce(indptr, eid, X, H, w_perm=norm[(eid.long() & 0x7FFFFFFF)], mode=1 | 16)
    seen = 0
    for rank in range(world):
        part = partition_by_destination(s, d, r, n, rank, world)
        ids = torch.from_numpy(part["eids"])
        npad = part["num_nodes_padded"]
        lp, le = sc.stable_segments(torch.from_numpy(part["dst"].astype(np.int32)),
                                    torch.from_numpy(part["rev"].astype(np.uint8)), npad)
        lw = norm[ids][(le.long() & 0x7FFFFFFF)]
        got = sc.seg_reduce(lp, le, X[ids], H, w_perm=lw, mode=1 | 16)
        lo, hi = part["n_lo"], min(part["n_hi"], n)
        assert torch.equal(got[lo:hi], want[lo:hi])
        assert float(got[:lo].abs().sum()) == 0 and float(got[part["n_hi"]:].abs().sum()) == 0
        seen += ids.numel()
    assert seen == E
