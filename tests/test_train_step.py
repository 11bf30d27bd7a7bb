"""Host-side collate of the e2e training harness against the graph oracle (CPU), and the step itself (GPU)."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go


def test_collate_equals_per_graph_reverse_then_batch():
    from dualmessagepassing_b200 import train_step as ts
    ds = ts.SyntheticPairDataset("cfg1", num=40, seed=1)
    idx = np.array([3, 7, 8, 21, 39])
    b = ts.collate(ds, idx)
    for side, d in (("p", ds.p), ("g", ds.g)):
        graphs = []
        for i in idx:
            u = d["u"][d["eoff"][i]:d["eoff"][i + 1]]
            v = d["v"][d["eoff"][i]:d["eoff"][i + 1]]
            s, t, r = go.add_reversed_edges(u, v)       # train.py:299-313 on each graph ...
            graphs.append((s, t, int(d["n"][i]), r))
        src, dst, n, bn, be = go.batch_graphs([(s, t, k) for s, t, k, _ in graphs])   # ... then dgl.batch
        rev = np.concatenate([r for *_, r in graphs])
        assert np.array_equal(b[side]["src"], src) and np.array_equal(b[side]["dst"], dst)
        assert np.array_equal(b[side]["rev"], rev) and b[side]["num_nodes"] == n
        assert np.array_equal(b[side]["n"], bn) and np.array_equal(b[side]["e"], be)
        assert np.all(b[side]["src"] != b[side]["dst"])  # generator draws no self loops
        assert np.all(d["u"] < np.repeat(d["n"], d["e"])) and np.all(d["v"] < np.repeat(d["n"], d["e"]))


@pytest.mark.gpu
def test_train_step_runs_and_learns():
    from dualmessagepassing_b200 import train_step as ts
    ds = ts.SyntheticPairDataset("cfg1", num=64, seed=3)
    torch.manual_seed(0)
    model = ts.SubgraphCountingModel(64, 1, 1).cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3, amsgrad=True)
    p, g, y, nb = ts.to_device(ts.collate(ds, np.arange(64)), torch.device("cuda"))
    losses = [float(ts.train_step(model, opt, p, g, y)) for _ in range(30)]
    assert nb > 0 and all(np.isfinite(losses))
    assert losses[-1] < 0.7 * losses[0], losses


@pytest.mark.gpu
def test_union_of_pattern_and_graph_batches_equals_separate_calls():
    """Shared layers on the disjoint union == the two separate calls (same per-row arithmetic; fp32 tolerance because
    the dense projections see different row counts)."""
    from dualmessagepassing_b200 import train_step as ts
    ds = ts.SyntheticPairDataset("cfg3", num=32, seed=5)
    torch.manual_seed(1)
    model = ts.SubgraphCountingModel(128, 7, 4).cuda()
    p, g, y, _ = ts.to_device(ts.collate(ds, np.arange(32)), torch.device("cuda"))
    a = model(p, g)
    b = model(p, g, union=ts.union_graph(p, g))
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
