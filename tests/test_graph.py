"""DMPGraph helpers against the numpy graph oracle and the reference-generated batched graph."""
import numpy as np
import torch

import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200.constants import REVFLAG
from oracle import graph_oracle as go
from tests import _golden


def _rand_graphs(seed, k):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(k):
        n = int(rng.integers(1, 9))
        u, v = go.erdos_renyi(rng, n, int(rng.integers(0, 3 * n)), self_loops=True)
        out.append((u, v, n))
    return out


def test_batch_matches_oracle_and_preserves_order():
    gs = _rand_graphs(1, 7)
    src, dst, n, bn, be = go.batch_graphs(gs)
    bg = dmp.batch([dmp.DMPGraph(u, v, k) for u, v, k in gs])
    s, d = bg.all_edges()
    assert np.array_equal(s.numpy(), src) and np.array_equal(d.numpy(), dst)
    assert bg.number_of_nodes() == n
    assert np.array_equal(bg.batch_num_nodes().numpy(), bn) and np.array_equal(bg.batch_num_edges().numpy(), be)


def test_add_reversed_edges_layout():
    u, v, n = _rand_graphs(2, 1)[0]
    g = dmp.DMPGraph(u, v, n)
    g.edata["label"] = torch.arange(len(u))
    dmp.add_reversed_edges(g, max_num_edges=100, max_edge_label=10)
    s, d, r = go.add_reversed_edges(u, v)
    gs, gd = g.all_edges()
    assert np.array_equal(gs.numpy(), s) and np.array_equal(gd.numpy(), d)
    assert np.array_equal(g.edata[REVFLAG].numpy(), r)
    assert torch.equal(g.edata["label"][len(u):], torch.arange(len(u)) + 10)
    assert torch.equal(g.edata["id"][len(u):], torch.arange(100, 100 + len(u)))
    assert g.rev_layout_hint == "halves"
    assert torch.equal(g.out_degrees(), torch.from_numpy(go.out_degrees(s, n)))


def test_batched_reversed_graphs_reproduce_reference_golden():
    """Per-graph add_reversed_edges then batch == what the reference pipeline (shim) produced."""
    case = _golden.load("scm_graph_rep_3layers")
    bn = case["batch_num_nodes"].tolist()
    be = case["batch_num_edges"].tolist()
    ps, pd = case["per_graph_src"], case["per_graph_dst"]
    graphs, off = [], 0
    for n, e in zip(bn, be):
        g = dmp.DMPGraph(ps[off:off + e // 2], pd[off:off + e // 2], n)  # original half of each graph
        dmp.add_reversed_edges(g)
        graphs.append(g)
        off += e
    bg = dmp.batch(graphs)
    s, d = bg.all_edges()
    assert torch.equal(s, case["src"]) and torch.equal(d, case["dst"])
    assert torch.equal(bg.edata[REVFLAG], case["rev"])
    assert torch.equal(bg.out_degrees(), case["out_deg"])


def test_build_graph_from_triplets_matches_oracle_and_golden():
    case = _golden.load("unc_norm_bn_tanh")
    n, R = case["num_nodes"], case["num_rels"] // 2
    g = dmp.build_graph_from_triplets(n, R, case["triplets"])
    s, d = g.all_edges()
    assert torch.equal(s, case["src"]) and torch.equal(d, case["dst"])
    assert torch.equal(g.edata["type"], case["rel"])
    assert torch.equal(g.edata["norm"], case["norm"])
    os_, od, orel, onorm = go.build_graph_from_triplets(n, R, case["triplets"].numpy())
    assert np.array_equal(os_, s.numpy()) and np.array_equal(orel, case["rel"].numpy())
    assert np.array_equal(onorm, case["norm"].numpy())


def test_layout_hint_is_dropped_when_the_flags_are_replaced():
    """a `rev_layout_hint` must not outlive the flag tensor it described."""
    import dualmessagepassing_b200 as dmp
    from dualmessagepassing_b200.constants import REVFLAG
    g = dmp.add_reversed_edges(dmp.DMPGraph([0, 1, 2], [1, 2, 0], 3))
    assert g.rev_layout_hint == "halves"
    same = g.edata[REVFLAG]
    g.edata[REVFLAG] = same                       # re-binding the same tensor keeps the hint
    assert g.rev_layout_hint == "halves"
    g.edata[REVFLAG] = torch.tensor([1, 0, 1, 0, 1, 0], dtype=torch.bool)
    assert g.rev_layout_hint is None
    g.rev_layout_hint = "general"
    g.edata.pop(REVFLAG)
    assert g.rev_layout_hint is None
