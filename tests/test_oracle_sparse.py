"""The sequential C oracle of the sparse core against independent CPU formulations (torch / numpy)."""
import numpy as np
import torch

from oracle import graph_oracle as go
from oracle import sparse_core as sc


def _graph(seed, n, e0, rev=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    u, v = go.erdos_renyi(rng, n, e0, self_loops=True)
    if rev:
        s, d, r = go.add_reversed_edges(u, v)
        perm = rng.permutation(len(s))  # arbitrary flag pattern, not just halves
        return s[perm], d[perm], r[perm]
    return u, v, None


def test_stable_segments_equal_numpy_stable_argsort():
    s, d, r = _graph(3, 50, 400)
    plan = go.build_plan(s, d, 50, r)
    for key, ip, eid in (("dst32", "csc_indptr", "csc_eid"), ("a32", "a_indptr", "a_eid"), ("b32", "b_indptr", "b_eid")):
        indptr, e = sc.stable_segments(torch.from_numpy(plan[key]), torch.from_numpy(r.astype(np.uint8)), 50)
        assert np.array_equal(indptr.numpy(), plan[ip])
        assert np.array_equal(e.numpy() & 0x7FFFFFFF, plan[eid])
        assert np.array_equal((e.numpy().view(np.uint32) >> 31).astype(bool), r[plan[eid]])


def test_seg_reduce_equals_torch_index_add_in_edge_order():
    s, d, r = _graph(4, 40, 300)
    E, H = len(s), 24
    g = torch.Generator().manual_seed(0)
    M = torch.randn(E, H, generator=g)
    base = torch.randn(40, H, generator=g)
    bias = torch.randn(H, generator=g)
    norm = torch.rand(E, generator=g)
    r8 = torch.from_numpy(r.astype(np.uint8))
    indptr, eid = sc.stable_segments(torch.from_numpy(d.astype(np.int32)), r8, 40)
    w_perm = norm[(eid.long() & 0x7FFFFFFF)]
    got = sc.seg_reduce(indptr, eid, M, H, w_perm=w_perm, base=base, bias=bias, mode=1)
    sgn = torch.where(torch.from_numpy(r), 1.0, -1.0).unsqueeze(1)
    msg = (sgn * M) * norm.unsqueeze(1)
    want = (base + torch.zeros(40, H).index_add(0, torch.from_numpy(d), msg)) + bias
    assert torch.equal(got, want)
    # SPLIT_BY_REV (16): [sum over forward edges | sum over reversed edges], each an index_add over its subset
    split = sc.seg_reduce(indptr, eid, M, H, w_perm=w_perm, mode=1 | 16)
    rt, dt = torch.from_numpy(r), torch.from_numpy(d)
    for half, keep in ((split[:, :H], ~rt), (split[:, H:], rt)):
        assert torch.equal(half, torch.zeros(40, H).index_add(0, dt[keep], msg[keep]))
    assert torch.equal(split[:, :H], sc.seg_reduce(indptr, eid, M, H, w_perm=w_perm, mode=1 | 4))
    assert torch.equal(split[:, H:], sc.seg_reduce(indptr, eid, M, H, w_perm=w_perm, mode=1 | 8))


def test_edge_update_and_backward_equal_torch_expressions():
    s, d, r = _graph(5, 30, 200)
    plan = go.build_plan(s, d, 30, r)
    E, H = len(s), 10
    g = torch.Generator().manual_seed(1)
    S, P = torch.randn(E, H, generator=g), torch.randn(E, H, generator=g)
    Qd, Qs = torch.randn(30, H, generator=g), torch.randn(30, H, generator=g)
    eb = torch.randn(H, generator=g)
    coef = torch.from_numpy(plan["coef"])
    a, b = torch.from_numpy(plan["a32"]), torch.from_numpy(plan["b32"])
    msg = Qd[a.long()] - Qs[b.long()]
    add = coef.unsqueeze(1) * P
    assert torch.equal(sc.edge_update(a, b, coef, S, P, Qd, Qs, eb, 0), ((S + add) + msg) + eb)
    out, agg = sc.edge_update(a, b, coef, S, P, Qd, Qs, eb, 1, want_agg=True)
    assert torch.equal(out, ((S + msg) + add) + eb) and torch.equal(agg, msg)
    gN, gE = torch.randn(30, H, generator=g), torch.randn(E, H, generator=g)
    norm = torch.rand(E, generator=g)
    r8 = torch.from_numpy(r.astype(np.uint8))
    T, CG = sc.edge_backward(torch.from_numpy(plan["dst32"]), r8, norm, coef, gN, gE)
    sgn = torch.where(torch.from_numpy(r), 1.0, -1.0).unsqueeze(1)
    assert torch.equal(T, sgn * (gN[torch.from_numpy(d)] * norm.unsqueeze(1)))
    assert torch.equal(CG, coef.unsqueeze(1) * gE)


def test_degree_coef_is_reference_expression():
    deg = np.array([0, 1, 2, 3, 7, 100, 4095, 100000])
    dst = np.arange(len(deg))
    c = go.degree_coef(deg, dst)
    d = torch.from_numpy(deg).float()
    assert np.array_equal(c, (2 * (1 + (1 + d).log2())).numpy())
    assert c[0] == 2.0 and c[1] == 4.0 and c[3] == 6.0
