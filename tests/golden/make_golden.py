"""Generate golden vectors by executing the UNMODIFIED reference layer classes.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

`dgl` cannot be installed offline, so the reference classes are imported over the DGL shim in
`_dgl_shim.py` (five entry points restated).  All arithmetic inside the layers -- the part the
oracle and the CUDA path must reproduce -- is the reference's own code:
  SubgraphCountingMatching/models/dmpnn.py:16-176 (DMPLayer), 215-277 (rep loops)
  UnsupervisedNodeClassification/Model/DMPNN/src/model.py:117-280 (DualGraphConv), 310-328 (DMPNN.forward pooling)
Output: tests/golden/<case>.npz with inputs, the reference-initialised state_dict, outputs and
all gradients (fp32, CPU, torch RNG seeded per case).
"""
import os
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _dgl_shim  # noqa: E402

REF = "/root/reference"


def er_graph(rng, n, e0, self_loops):
    from oracle.graph_oracle import erdos_renyi
    return erdos_renyi(rng, n, e0, self_loops)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, sum(a.nbytes for a in out.values()) // 1024, "KiB")


def state(module):
    return {"param:" + k: v.detach().clone() for k, v in module.state_dict().items()}


def grads(module):
    return {"grad:" + k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
            for k, p in module.named_parameters()}


# ------------------------------------------------------------------------------------------------
def run_scm():
    _dgl_shim.install()
    sys.path.insert(0, os.path.join(REF, "SubgraphCountingMatching"))
    from models.dmpnn import DMPLayer, DMPNN  # the reference classes, unmodified
    import constants as C

    cases = [
        # name, N, E0, Din, H, rev, loops, mlp, bn, act, bias, supplied out_deg
        ("scm_rev_mlp2_lrelu", 40, 70, 16, 16, True, False, 2, False, "leaky_relu", True, False),
        ("scm_rev_mlp0_relu", 33, 50, 12, 12, True, True, 0, False, "relu", True, False),
        ("scm_norev_mlp2_bn_tanh", 25, 90, 8, 20, False, True, 2, True, "tanh", True, False),
        ("scm_rev_din_ne_h_nobias", 30, 45, 10, 24, True, False, 2, False, "leaky_relu", False, True),
        ("scm_rev_h64_medium", 200, 800, 64, 64, True, False, 2, False, "leaky_relu", True, False),
        ("scm_rev_mlp1_h50", 21, 64, 50, 50, True, True, 1, False, "relu", True, False),
    ]
    for ci, (name, n, e0, din, h, rev, loops, mlp, bn, act, bias, given_deg) in enumerate(cases):
        rng = np.random.Generator(np.random.PCG64(7000 + ci))
        torch.manual_seed(7000 + ci)
        u, v = er_graph(rng, n, e0, loops)
        # leave a few isolated nodes: remap endpoints into [0, n-3)
        u, v = u % max(n - 3, 1), v % max(n - 3, 1)
        g = _dgl_shim.ShimGraph(u, v, n)
        if rev:  # exactly SubgraphCountingMatching/train.py:299-313
            num_ge = g.number_of_edges()
            uu, vv = g.all_edges(form="uv", order="eid")
            g.add_edges(vv, uu, data={C.REVFLAG: torch.ones((num_ge,), dtype=torch.bool)})
        if given_deg:  # caller-supplied out_deg must be honoured (dmpnn.py:100-101)
            g.ndata[C.OUTDEGREE] = torch.from_numpy(rng.integers(0, 9, size=n)).long()
        layer = DMPLayer(din, h, bias=bias, num_mlp_layers=mlp, batch_norm=bn, act_func=act, dropout=0.0)
        layer.train()
        E = g.number_of_edges()
        xv = torch.randn(n, din, requires_grad=True)
        xe = torch.randn(E, din, requires_grad=True)
        gv = torch.randn(n, h)
        ge = torch.randn(E, h)
        st = state(layer)  # before forward (BN running stats)
        nv, ne = layer(g, xv, xe)
        edge_agg = g.edata[C.EDGEAGG].detach().clone()
        ((nv * gv).sum() + (ne * ge).sum()).backward()
        src, dst = g.all_edges()
        save(name, src=src, dst=dst, num_nodes=n,
             rev=g.edata[C.REVFLAG] if rev else None,
             out_deg=g.ndata[C.OUTDEGREE], out_deg_given=np.asarray(given_deg),
             node_feat=xv, edge_feat=xe, grad_node_out=gv, grad_edge_out=ge,
             node_out=nv, edge_out=ne, edge_agg=edge_agg,
             grad_node_feat=xv.grad, grad_edge_feat=xe.grad,
             meta=np.asarray([din, h, mlp, int(bn), int(bias)]), act=np.asarray(act),
             **st, **grads(layer))

    # ---- three shared layers on a batch of graphs through DMPNN.get_graph_rep / get_pattern_rep ---
    rng = np.random.Generator(np.random.PCG64(7100))
    torch.manual_seed(7100)
    h = 16
    gs = []
    for _ in range(5):
        n = int(rng.integers(4, 12))
        u, v = er_graph(rng, n, 2 * n, False)
        g = _dgl_shim.ShimGraph(u, v, n)
        uu, vv = g.all_edges()
        g.add_edges(vv, uu, data={C.REVFLAG: torch.ones((len(u),), dtype=torch.bool)})
        g.ndata[C.OUTDEGREE] = g.out_degrees()
        gs.append(g)
    bg = _dgl_shim.batch(gs)  # flags interleave per graph: [fwd0, rev0, fwd1, rev1, ...]
    layers = torch.nn.ModuleList([DMPLayer(h, h, num_mlp_layers=2, batch_norm=False,
                                           act_func="leaky_relu") for _ in range(3)])
    fake = types.SimpleNamespace(g_rep_net={"dmpnn": layers}, p_rep_net={"dmpnn": layers}, rep_residual=True)
    N, E = bg.number_of_nodes(), bg.number_of_edges()
    xv = torch.randn(N, h, requires_grad=True)
    xe = torch.randn(E, h, requires_grad=True)
    v_gate = (torch.rand(N, 1) > 0.3).float()
    e_gate = (torch.rand(E, 1) > 0.3).float()
    gv, ge = torch.randn(N, h), torch.randn(E, h)
    st = {("param:%d." % i) + k: v.detach().clone() for i, l in enumerate(layers) for k, v in l.state_dict().items()}
    ov, oe = DMPNN.get_graph_rep(fake, bg, xv, xe, v_gate=v_gate, e_gate=e_gate)
    ((ov * gv).sum() + (oe * ge).sum()).backward()
    gr = {("grad:%d." % i) + k: p.grad.detach().clone() for i, l in enumerate(layers) for k, p in l.named_parameters()}
    src, dst = bg.all_edges()
    extra = dict(src=src, dst=dst, num_nodes=N, rev=bg.edata[C.REVFLAG], out_deg=bg.ndata[C.OUTDEGREE],
                 batch_num_nodes=bg.batch_num_nodes(), batch_num_edges=bg.batch_num_edges(),
                 node_feat=xv, edge_feat=xe, v_gate=v_gate, e_gate=e_gate,
                 grad_node_out=gv, grad_edge_out=ge, node_out=ov, edge_out=oe,
                 grad_node_feat=xv.grad.clone(), grad_edge_feat=xe.grad.clone(),
                 per_graph_src=np.concatenate([g._src.numpy() for g in gs]),
                 per_graph_dst=np.concatenate([g._dst.numpy() for g in gs]))
    save("scm_graph_rep_3layers", **extra, **st, **gr)

    # pattern side: masked_fill + residual (dmpnn.py:215-243)
    for l in layers:
        l.zero_grad()
    xv2 = xv.detach().clone().requires_grad_(True)
    xe2 = xe.detach().clone().requires_grad_(True)
    v_mask, e_mask = v_gate.bool(), e_gate.bool()
    ov, oe = DMPNN.get_pattern_rep(fake, bg, xv2, xe2, v_mask=v_mask, e_mask=e_mask)
    ((ov * gv).sum() + (oe * ge).sum()).backward()
    gr = {("grad:%d." % i) + k: p.grad.detach().clone() for i, l in enumerate(layers) for k, p in l.named_parameters()}
    extra.update(node_out=ov, edge_out=oe, grad_node_feat=xv2.grad, grad_edge_feat=xe2.grad)
    save("scm_pattern_rep_3layers", **extra, **st, **gr)


# ------------------------------------------------------------------------------------------------
def run_unc():
    _dgl_shim.install()
    sys.path.insert(0, os.path.join(REF, "UnsupervisedNodeClassification/Model/DMPNN/src"))
    import model as M  # the reference module, unmodified

    cases = [
        # name, N, triplets, R, Din, H, bn, post_act, is_rev key present
        ("unc_norm_bn_tanh", 30, 60, 4, 10, 10, True, "tanh", False),
        ("unc_norm_bn_last", 30, 60, 4, 10, 14, True, None, False),
        ("unc_isrev_nobn", 24, 40, 3, 8, 8, False, "tanh", True),
        ("unc_h50_medium", 120, 500, 10, 50, 50, True, "tanh", False),
    ]
    from oracle.graph_oracle import build_graph_from_triplets
    for ci, (name, n, t, R, din, h, bn, post, isrev) in enumerate(cases):
        rng = np.random.Generator(np.random.PCG64(7200 + ci))
        torch.manual_seed(7200 + ci)
        trip = np.stack([rng.integers(0, n, t), rng.integers(0, R, t), rng.integers(0, n, t)], 1).astype(np.int64)
        src, dst, rel, _ = build_graph_from_triplets(n, R, trip)
        g = _dgl_shim.ShimGraph(src, dst, n)
        # UnsupervisedNodeClassification/Model/DMPNN/src/utils.py:437-453 semantics, via torch
        indeg = g.in_degrees().float()
        norm = indeg[g._dst].reciprocal().unsqueeze(-1)
        norm.masked_fill_(torch.isnan(norm), norm.min())
        norm.masked_fill_(torch.isinf(norm), norm.min())
        if isrev:
            g.edata["is_rev"] = torch.cat([torch.zeros(t, dtype=torch.bool), torch.ones(t, dtype=torch.bool)])
        act = torch.nn.Tanh() if post == "tanh" else None
        layer = M.DualGraphConv(din, h, batch_norm=bn, activation=act, dropout=0.0)
        layer.train()
        E = 2 * t
        xv = torch.randn(n, din, requires_grad=True)
        xe = torch.randn(E, din, requires_grad=True)
        gv, ge = torch.randn(n, h), torch.randn(E, h)
        st = state(layer)
        nv, ne = layer(g, xv, xe, norm)
        # relation pooling of DMPNN.forward (model.py:319-325), same expression on the layer output
        r = torch.from_numpy(rel)
        pooled = torch.cat([ne.masked_fill((r != i).view(-1, 1), 0.0).sum(dim=0, keepdim=True)
                            / ((r == i).sum().float() + 1e-8) for i in range(2 * R)], dim=0)
        ((nv * gv).sum() + (ne * ge).sum()).backward()
        save(name, src=src, dst=dst, num_nodes=n, rel=rel, num_rels=2 * R, triplets=trip,
             rev=g.edata["is_rev"] if isrev else None, norm=norm, out_deg=g.ndata["out_deg"],
             node_feat=xv, edge_feat=xe, grad_node_out=gv, grad_edge_out=ge,
             node_out=nv, edge_out=ne, rel_pooled=pooled,
             grad_node_feat=xv.grad, grad_edge_feat=xe.grad,
             meta=np.asarray([din, h, 2, int(bn), 1]), act=np.asarray(post or "none"),
             **st, **grads(layer))


def run_dmplrp():
    """DMPLRPPoolLayer (SubgraphCountingMatching/models/dmplrp.py:19-198, unmodified): the dual message-passing step +
    local relational pooling through three torch.sparse products and one einsum.  The pooling matrices are random
    sparse matrices of the shapes the LRP preprocessing produces ([D*L^2, N], [D*L^2, E], [N, D])."""
    _dgl_shim.install()
    sys.path.insert(0, os.path.join(REF, "SubgraphCountingMatching"))
    from models.dmplrp import DMPLRPPoolLayer  # the reference class, unmodified
    import constants as C
    for ci, (name, n, e0, h, L, D, mlp, bn, act) in enumerate([
            ("lrp_h16_mlp2", 30, 50, 16, 4, 22, 2, False, "leaky_relu"),
            ("lrp_h12_mlp0_bn", 18, 40, 12, 3, 15, 0, True, "tanh")]):
        rng = np.random.Generator(np.random.PCG64(7300 + ci))
        torch.manual_seed(7300 + ci)
        u, v = er_graph(rng, n, e0, False)
        g = _dgl_shim.ShimGraph(u, v, n)
        uu, vv = g.all_edges(form="uv", order="eid")
        g.add_edges(vv, uu, data={C.REVFLAG: torch.ones((len(u),), dtype=torch.bool)})
        E = g.number_of_edges()
        layer = DMPLRPPoolLayer(h, h, lrp_seq_len=L, num_mlp_layers=mlp, batch_norm=bn, act_func=act)
        layer.train()

        def sparse(rows, cols, nnz):
            idx = np.unique(np.stack([rng.integers(0, rows, nnz), rng.integers(0, cols, nnz)]), axis=1)
            val = torch.from_numpy(rng.standard_normal(idx.shape[1]).astype(np.float32))
            return torch.sparse_coo_tensor(torch.from_numpy(idx), val, (rows, cols)).coalesce()

        n2p, e2p, pool = sparse(D * L * L, n, 3 * D * L), sparse(D * L * L, E, 3 * D * L), sparse(n, D, 4 * n)
        xv = torch.randn(n, h, requires_grad=True)
        xe = torch.randn(E, h, requires_grad=True)
        gv, ge = torch.randn(n, h), torch.randn(E, h)
        st = state(layer)
        nv, ne, *_ = layer(g, xv, xe, pool, n2p, e2p)
        ((nv * gv).sum() + (ne * ge).sum()).backward()
        src, dst = g.all_edges()
        mats = {}
        for k, m in (("n2p", n2p), ("e2p", e2p), ("pool", pool)):
            mats[k + "_idx"], mats[k + "_val"], mats[k + "_shape"] = m.indices(), m.values(), np.asarray(m.shape)
        save(name, src=src, dst=dst, num_nodes=n, rev=g.edata[C.REVFLAG], out_deg=g.ndata[C.OUTDEGREE],
             node_feat=xv, edge_feat=xe, grad_node_out=gv, grad_edge_out=ge, node_out=nv, edge_out=ne,
             grad_node_feat=xv.grad, grad_edge_feat=xe.grad, meta=np.asarray([h, L, D, mlp, int(bn)]),
             act=np.asarray(act), **mats, **st, **grads(layer))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "unc":
        run_unc()
    elif len(sys.argv) > 1 and sys.argv[1] == "lrp":
        run_dmplrp()
    else:
        run_scm()
        run_dmplrp()
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "unc"])
