"""Rows N1 / N2: device-side batching (dmp_batch_offsets / dmp_batch_fill) bit-exact against the host collate (itself
checked against the graph oracle in test_train_step.py), and the ragged <-> padded kernels against the restated
reference loop (oracle/graph_oracle.py: dl.py:51-81,113-127)."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,num,bsz", [("cfg1", 200, 64), ("cfg2", 700, 512), ("cfg3", 100, 64), ("cfg1", 2600, 2500)])
def test_device_batch_equals_host_collate(cfg, num, bsz):
    from dualmessagepassing_b200 import train_step as ts
    from dualmessagepassing_b200.constants import EDGELABEL, NODELABEL, REVFLAG
    ds = ts.SyntheticPairDataset(cfg, num=num, seed=11)
    dds = ts.DevicePairDataset(ds, "cuda")
    rng = np.random.Generator(np.random.PCG64(5))
    idx = np.sort(rng.choice(num, size=bsz, replace=False))
    host = ts.collate(ds, idx)
    p, g, y, nbytes = ts.collate_on_device(dds, idx)
    assert nbytes == idx.nbytes
    for side, dg in (("p", p), ("g", g)):
        want = host[side]
        src, dst = dg.all_edges()
        assert np.array_equal(src.cpu().numpy(), want["src"]) and np.array_equal(dst.cpu().numpy(), want["dst"])
        assert np.array_equal(dg.edata[REVFLAG].cpu().numpy(), want["rev"])
        assert np.array_equal(dg.ndata[NODELABEL].cpu().numpy(), want["vl"])
        assert np.array_equal(dg.edata[EDGELABEL].cpu().numpy(), want["el"])
        assert np.array_equal(dg.ndata["graph_id"].cpu().numpy(), want["node_graph"])
        assert np.array_equal(dg.batch_num_nodes().cpu().numpy(), want["n"])
        assert np.array_equal(dg.batch_num_edges().cpu().numpy(), want["e"])
        assert dg.number_of_nodes() == want["num_nodes"]
    assert np.array_equal(y.cpu().numpy(), host["y"])
    # without reversed edges: plain dgl.batch
    sel = torch.from_numpy(idx).cuda()
    g0 = ts.batch_on_device(dds.sides["g"], sel, idx, add_reversed=False)
    graphs = [(ds.g["u"][ds.g["eoff"][i]:ds.g["eoff"][i + 1]], ds.g["v"][ds.g["eoff"][i]:ds.g["eoff"][i + 1]],
               int(ds.g["n"][i])) for i in idx]
    s0, d0, n0, bn, be = go.batch_graphs(graphs)
    src, dst = g0.all_edges()
    assert np.array_equal(src.cpu().numpy(), s0) and np.array_equal(dst.cpu().numpy(), d0) and g0.number_of_nodes() == n0
    assert not bool(g0.edata[REVFLAG].any())


@pytest.mark.gpu
def test_train_step_on_device_batches_matches_host_batches():
    """Same model, same pairs: the loss from device-built batches equals the loss from host-collated ones bit for bit."""
    from dualmessagepassing_b200 import train_step as ts
    ds = ts.SyntheticPairDataset("cfg1", num=64, seed=3)
    dds = ts.DevicePairDataset(ds, "cuda")
    torch.manual_seed(0)
    model = ts.SubgraphCountingModel(64, 1, 1).cuda()
    idx = np.arange(0, 64, 2)
    p, g, y, _ = ts.to_device(ts.collate(ds, idx), torch.device("cuda"))
    p2, g2, y2, _ = ts.collate_on_device(dds, idx)
    a = model(p, g, union=ts.union_graph(p, g))
    b = model(p2, g2, union=ts.union_graph(p2, g2))
    assert torch.equal(a, b) and torch.equal(y, y2)


@pytest.mark.gpu
def test_padded_device_batch_has_fixed_shapes_and_inert_padding():
    """pad_to: the real part equals the unpadded batch, the tail is dummy nodes (graph id B) + self-loops spread over them; the model's prediction on a padded batch equals the unpadded one."""
    from dualmessagepassing_b200 import train_step as ts
    from dualmessagepassing_b200.constants import EDGELABEL, NODELABEL, REVFLAG
    ds = ts.SyntheticPairDataset("cfg1", num=120, seed=4)
    dds = ts.DevicePairDataset(ds, "cuda")
    idx = np.arange(10, 74)
    sel = torch.from_numpy(idx).cuda()
    real = ts.batch_on_device(dds.sides["g"], sel, idx)
    N, E = real.number_of_nodes(), real.number_of_edges()
    padded = ts.batch_on_device(dds.sides["g"], sel, None, pad_to=(N + 37, E + 200))
    assert padded.number_of_nodes() == N + 37 and padded.number_of_edges() == E + 200
    (s0, d0), (s1, d1) = real.all_edges(), padded.all_edges()
    assert torch.equal(s1[:E], s0) and torch.equal(d1[:E], d0)
    assert torch.equal(s1[E:], d1[E:]) and bool((s1[E:] >= N).all()) and not bool(padded.edata[REVFLAG][E:].any())
    assert int(torch.bincount(s1[E:] - N, minlength=37).max()) <= -(-200 // 37)       # spread round-robin: no dummy hub
    for key, frame in ((NODELABEL, "ndata"), ("graph_id", "ndata"), (EDGELABEL, "edata")):
        a, b = getattr(real, frame)[key], getattr(padded, frame)[key]
        assert torch.equal(b[:a.numel()], a)
    assert bool((padded.ndata["graph_id"][N:] == 64).all())
    assert torch.equal(padded.batch_num_nodes(), real.batch_num_nodes())
    torch.manual_seed(0)
    model = ts.SubgraphCountingModel(64, 1, 1).cuda()
    p = ts.batch_on_device(dds.sides["p"], sel, idx)
    p_pad = ts.batch_on_device(dds.sides["p"], sel, None, pad_to=(p.number_of_nodes() + 5, p.number_of_edges() + 64))
    a = model(p, real, union=ts.union_graph(p, real))
    b = model(p_pad, padded, union=ts.union_graph(p_pad, padded))
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_graphed_train_step_matches_eager_step():
    """The CUDA-graphed step (device batching + fwd + bwd + clip + AdamW in one replay) follows the eager step's loss
    trajectory on the same pair ids."""
    import copy
    from dualmessagepassing_b200 import _lib, train_step as ts
    ds = ts.SyntheticPairDataset("cfg1", num=256, seed=6)
    dds = ts.DevicePairDataset(ds, "cuda")
    torch.manual_seed(0)
    model = ts.SubgraphCountingModel(64, 1, 1).cuda()
    ref = copy.deepcopy(model)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True, capturable=True)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, amsgrad=True, capturable=True)
    rng = np.random.Generator(np.random.PCG64(1))
    batches = [np.sort(rng.choice(256, size=64, replace=False)) for _ in range(6)]
    graphed = ts.GraphedTrainStep(model, opt, dds, 64, warmup=2)
    l0 = _lib.LAUNCHES
    got = [float(graphed(b)) for b in batches]
    assert _lib.LAUNCHES == l0 and graphed.fallbacks == 0            # replays only: no C-ABI call from the host
    want = []
    for b in graphed.warmup_ids:                                     # the warm-up steps were real optimizer steps
        p, g, y, _ = ts.collate_on_device(dds, b)
        ts.train_step(ref, opt_ref, p, g, y)
    for b in batches:
        p, g, y, _ = ts.collate_on_device(dds, b)
        want.append(float(ts.train_step(ref, opt_ref, p, g, y)))
    np.testing.assert_allclose(got, want, rtol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("pre_pad", [False, True])
@pytest.mark.parametrize("H", [1, 50, 128])
def test_ragged_pad_matches_reference_loop(pre_pad, H):
    from dualmessagepassing_b200 import ragged
    rng = np.random.Generator(np.random.PCG64(H))
    sizes = rng.integers(1, 40, size=97)
    x = rng.standard_normal((int(sizes.sum()), H)).astype(np.float32)
    want, wmask = go.split_and_batchify(x, sizes, pre_pad)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    got, mask = ragged.split_and_batchify_graph_feats(xt, torch.from_numpy(sizes).cuda(), pre_pad=pre_pad)
    assert np.array_equal(got.detach().cpu().numpy(), want) and np.array_equal(mask.cpu().numpy(), wmask)
    # caller-supplied max_size (no synchronisation) larger than the largest graph
    got2, mask2 = ragged.split_and_batchify_graph_feats(xt, torch.from_numpy(sizes).cuda(), pre_pad=pre_pad, max_size=48)
    assert got2.shape == (97, 48, H) and int(mask2.sum()) == int(sizes.sum())
    sl = slice(48 - want.shape[1], None) if pre_pad else slice(0, want.shape[1])
    assert np.array_equal(got2[:, sl].detach().cpu().numpy(), want)
    # backward = the inverse gather
    w = torch.randn_like(got)
    (got * w).sum().backward()
    rows = []
    for i, l in enumerate(sizes):
        st = want.shape[1] - l if pre_pad else 0
        rows.append(w[i, st:st + l])
    assert torch.equal(xt.grad, torch.cat(rows))
    # equal sizes: the reference's .view fast path
    eq, m = ragged.split_and_batchify_graph_feats(xt[:90].detach(), torch.full((9,), 10).cuda())
    assert torch.equal(eq, xt[:90].detach().view(9, 10, H)) and bool(m.all())


def test_len_to_mask_and_deferred_scalars_cpu():
    from dualmessagepassing_b200 import ragged
    lens = [3, 1, 4, 4, 2]
    for pre in (False, True):
        for mx in (-1, 6):
            got = ragged.batch_convert_len_to_mask(torch.tensor(lens), mx, pre)
            assert np.array_equal(got.numpy(), go.len_to_mask(lens, mx, pre))
    d = ragged.DeferredScalars()
    for i in range(3):
        d.add(loss=torch.tensor(float(i)), reg=torch.tensor(2.0 * i))
    out = d.fetch()
    assert out == {"loss": [0.0, 1.0, 2.0], "reg": [0.0, 2.0, 4.0]} and d.fetch() == {}


@pytest.mark.gpu
@pytest.mark.parametrize("rows,vocab,width", [(50_000, 32, 64), (3_000, 7, 128), (20_000, 100, 64)])
def test_label_embedding_matches_nn_embedding(rows, vocab, width):
    """train_step.label_embedding: same rows forward, dW = onehot^T g (tensor-core reduction for >= 16384 rows) backward."""
    from dualmessagepassing_b200 import _lib
    from dualmessagepassing_b200.train_step import label_embedding
    g = torch.Generator(device="cuda").manual_seed(rows + vocab)
    labels = torch.randint(0, vocab, (rows,), device="cuda", generator=g)
    w = torch.randn(vocab, width, device="cuda", generator=g)
    up = torch.randn(rows, width, device="cuda", generator=g)
    w1, w2 = w.clone().requires_grad_(True), w.clone().requires_grad_(True)
    _lib.PROFILE = []
    out = label_embedding(w1, labels)
    out.backward(up)
    tags, _lib.PROFILE = {t[0] for t in _lib.PROFILE}, None
    ref = torch.nn.functional.embedding(labels, w2)
    ref.backward(up)
    assert torch.equal(out, ref)
    want = torch.zeros(vocab, width, dtype=torch.float64, device="cuda").index_add_(0, labels, up.double())
    scale = float(want.abs().max())
    assert float((w1.grad.double() - want).abs().max()) <= 1e-5 * scale
    assert float((w1.grad.double() - want).abs().max()) <= 4 * float((w2.grad.double() - want).abs().max()) + 2e-6 * scale
    assert ("gemm_tn_tf32x3" in tags) == (rows >= 512)
