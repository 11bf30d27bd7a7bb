"""dmp_plan_build on the GPU must be BIT-EXACT against the numpy restatement of the DGL index semantics."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from tests._cases import hub_graph, make_graph, t

pytestmark = pytest.mark.gpu

CASES = [
    ("tiny_rev", dict(seed=1, n=5, e0=7, rev="halves")),
    ("norev", dict(seed=2, n=64, e0=300, rev=None)),
    ("shuffled_flags", dict(seed=3, n=200, e0=1500, rev="shuffled")),
    ("many_isolated", dict(seed=4, n=500, e0=100, rev="halves", isolated=400)),
    ("big", dict(seed=5, n=100_000, e0=1_000_000, rev="halves")),
    ("pow2_nodes", dict(seed=6, n=1024, e0=5000, rev="shuffled")),
]


def _check(plan, want, rev):
    for k in ("dst32", "a32", "b32", "csc_indptr", "a_indptr", "b_indptr", "out_deg"):
        assert np.array_equal(getattr(plan, k).cpu().numpy(), want[k]), k
    for k in ("csc_eid", "a_eid", "b_eid"):
        got = getattr(plan, k).cpu().numpy().view(np.uint32)
        assert np.array_equal(got & 0x7FFFFFFF, want[k].view(np.uint32)), k
        flag = (got >> 31).astype(bool)
        exp = np.zeros_like(flag) if rev is None else np.asarray(rev, bool)[want[k]]
        assert np.array_equal(flag, exp), k + " flag"
    # degrees below the LUT length carry the host log2 bits exactly
    assert np.array_equal(plan.coef.cpu().numpy(), want["coef"])


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_plan_bit_exact(name, kw):
    from dualmessagepassing_b200.plan import DMPPlan
    s, d, r = make_graph(**kw)
    n = kw["n"]
    plan = DMPPlan(t(s), t(d), n, rev=t(None if r is None else r.astype(np.uint8)))
    _check(plan, go.build_plan(s, d, n, r), r)
    if kw["rev"] == "halves":
        assert plan.rev_layout == "halves" and plan.rev_split == len(s) // 2
    elif kw["rev"] == "shuffled":
        assert plan.rev_layout == "general"
    else:
        assert plan.rev_layout == "none"


def test_plan_honours_supplied_out_degree_and_large_degrees():
    from dualmessagepassing_b200.plan import DMPPlan
    s, d, r = make_graph(seed=7, n=50, e0=200, rev="halves")
    rng = np.random.Generator(np.random.PCG64(7))
    deg = rng.integers(0, 20000, size=50).astype(np.int64)  # beyond the LUT: device log2f branch
    plan = DMPPlan(t(s), t(d), 50, rev=t(r.astype(np.uint8)), out_deg=t(deg))
    want = go.build_plan(s, d, 50, r, out_deg=deg)
    assert np.array_equal(plan.out_deg.cpu().numpy(), deg)
    got, exp = plan.coef.cpu().numpy(), want["coef"]
    small = deg[d] < 4096
    assert np.array_equal(got[small], exp[small])
    np.testing.assert_allclose(got, exp, rtol=3e-7, atol=0)  # log2f vs host log2: <= 2 ulp


def test_plan_hub_and_empty_graphs():
    from dualmessagepassing_b200.plan import DMPPlan
    s, d, r = hub_graph(8, 300, 1000, 5000)
    plan = DMPPlan(t(s), t(d), 300, rev=t(r.astype(np.uint8)))
    _check(plan, go.build_plan(s, d, 300, r), r)
    empty = DMPPlan(torch.zeros(0, dtype=torch.int64, device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"), 4)
    assert empty.csc_indptr.tolist() == [0] * 5 and empty.out_deg.tolist() == [0] * 4


def test_plan_rejects_out_of_range_endpoint():
    from dualmessagepassing_b200.plan import DMPPlan
    with pytest.raises(ValueError, match="outside"):
        DMPPlan(torch.tensor([0, 9], device="cuda"), torch.tensor([1, 0], device="cuda"), 3)


def test_plan_is_cached_on_graph_and_invalidated():
    import dualmessagepassing_b200 as dmp
    from dualmessagepassing_b200.constants import OUTDEGREE, REVFLAG
    s, d, r = make_graph(seed=9, n=20, e0=40, rev="halves")
    g = dmp.DMPGraph(s[:40], d[:40], 20, device="cuda")
    dmp.add_reversed_edges(g)
    p1 = dmp.get_plan(g, REVFLAG, OUTDEGREE)
    assert dmp.get_plan(g, REVFLAG, OUTDEGREE) is p1
    assert torch.equal(g.ndata[OUTDEGREE], g.out_degrees())  # reference side effect: degrees cached in the frame
    g.ndata[OUTDEGREE] = torch.full((20,), 3, device="cuda")  # caller override must trigger a rebuild
    p2 = dmp.get_plan(g, REVFLAG, OUTDEGREE)
    assert p2 is not p1 and torch.all(p2.out_deg == 3)
