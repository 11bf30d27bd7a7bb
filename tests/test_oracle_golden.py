"""The oracle restatement must reproduce the reference classes' golden vectors (CPU, fp32).

Golden vectors come from the unmodified reference `DMPLayer` / `DualGraphConv` / `DMPNN.get_*_rep`
executed over the DGL shim (tests/golden/make_golden.py).  Same ops in the same order on the same
BLAS => bit-identical forward; gradients are compared at 1e-6/1e-5 because autograd's accumulation
order of the several X_e / X_v contributions is not part of the contract.
"""
import pytest
import torch

from oracle import dmp_oracle
from tests import _golden

SCM = [c for c in _golden.case_names("scm_") if "rep_3layers" not in c]
UNC = _golden.case_names("unc_")


def _float_params(params, dtype=torch.float32):
    out = {}
    for k, v in params.items():
        if v.dtype.is_floating_point:
            out[k] = v.clone().to(dtype).requires_grad_(True)
        else:
            out[k] = v.clone()
    return out


def _run(case, flavour):
    P = _float_params(case["params"])
    xv = case["node_feat"].clone().requires_grad_(True)
    xe = case["edge_feat"].clone().requires_grad_(True)
    kw = {}
    if flavour == "scm":
        kw = dict(act_func=case["act"])
    else:
        # model.py:144-167: the constructor's `activation` is used INSIDE the MLP too (LeakyReLU if None)
        post = None if case["act"] == "none" else case["act"]
        kw = dict(act_func="leaky_relu", mlp_act=post or "leaky_relu", post_act=post, norm=case.get("norm"))
    nv, ne = dmp_oracle.dmp_layer(P, case["src"], case["dst"], case["num_nodes"], xv, xe,
                                  rev=case.get("rev"), out_deg=case["out_deg"], flavour=flavour, **kw)
    ((nv * case["grad_node_out"]).sum() + (ne * case["grad_edge_out"]).sum()).backward()
    return P, xv, xe, nv, ne


@pytest.mark.parametrize("name", SCM)
def test_scm_layer_matches_reference(name):
    case = _golden.load(name)
    P, xv, xe, nv, ne = _run(case, "scm")
    assert torch.equal(nv, case["node_out"])
    assert torch.equal(ne, case["edge_out"])
    torch.testing.assert_close(xv.grad, case["grad_node_feat"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xe.grad, case["grad_edge_feat"], rtol=1e-5, atol=1e-6)
    for k, g in case["grads"].items():
        got = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        torch.testing.assert_close(got, g, rtol=1e-5, atol=2e-6, msg=lambda m: k + ": " + m)


@pytest.mark.parametrize("name", UNC)
def test_unc_layer_matches_reference(name):
    case = _golden.load(name)
    P, xv, xe, nv, ne = _run(case, "unc")
    assert torch.equal(nv, case["node_out"])
    assert torch.equal(ne, case["edge_out"])
    torch.testing.assert_close(xv.grad, case["grad_node_feat"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xe.grad, case["grad_edge_feat"], rtol=1e-5, atol=1e-6)
    pooled = dmp_oracle.relation_mean_pool(ne.detach(), case["rel"], case["num_rels"])
    assert torch.equal(pooled, case["rel_pooled"])
    for k, g in case["grads"].items():
        if k.startswith(("nfc", "efc")):
            continue  # constructed but unused by the reference (model.py:137-138)
        got = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        torch.testing.assert_close(got, g, rtol=1e-5, atol=2e-6, msg=lambda m: k + ": " + m)


@pytest.mark.parametrize("side", ["graph", "pattern"])
def test_rep_loop_matches_reference(side):
    case = _golden.load("scm_%s_rep_3layers" % side)
    layers = [_float_params(p) for p in _golden.split_layers(case["params"])]
    xv = case["node_feat"].clone().requires_grad_(True)
    xe = case["edge_feat"].clone().requires_grad_(True)
    kw = dict(rev=case["rev"], out_deg=case["out_deg"], act_func="leaky_relu")
    if side == "graph":
        ov, oe = dmp_oracle.graph_rep(layers, case["src"], case["dst"], case["num_nodes"], xv, xe,
                                      v_gate=case["v_gate"], e_gate=case["e_gate"], **kw)
    else:
        ov, oe = dmp_oracle.pattern_rep(layers, case["src"], case["dst"], case["num_nodes"], xv, xe,
                                        v_mask=case["v_gate"].bool(), e_mask=case["e_gate"].bool(), **kw)
    assert torch.equal(ov, case["node_out"])
    assert torch.equal(oe, case["edge_out"])
    ((ov * case["grad_node_out"]).sum() + (oe * case["grad_edge_out"]).sum()).backward()
    # forward is bit-exact; through three layers the autograd accumulation order of the oracle vs the
    # reference already moves single elements by ~6e-6 abs (fp32 cancellation), hence 1e-4 / 2e-5 here
    torch.testing.assert_close(xv.grad, case["grad_node_feat"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(xe.grad, case["grad_edge_feat"], rtol=1e-4, atol=2e-5)
    for i, g in enumerate(_golden.split_layers(case["grads"])):
        for k, v in g.items():
            torch.testing.assert_close(layers[i][k].grad, v, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", _golden.case_names("lrp_"))
def test_dmplrp_pool_layer_matches_reference(name):
    """dmplrp.py:19-198: the oracle's DMPLayer body + lrp_pool reproduce the reference class bit for bit."""
    case = _golden.load(name)
    h, L, D, mlp, bn = [int(x) for x in case["meta"]]
    P = _float_params(case["params"])
    xv = case["node_feat"].clone().requires_grad_(True)
    xe = case["edge_feat"].clone().requires_grad_(True)
    mats = {k: torch.sparse_coo_tensor(case[k + "_idx"], case[k + "_val"], tuple(int(x) for x in case[k + "_shape"])).coalesce()
            for k in ("n2p", "e2p", "pool")}
    nv, ne = dmp_oracle.dmp_layer(P, case["src"], case["dst"], case["num_nodes"], xv, xe, rev=case["rev"],
                                  out_deg=case["out_deg"], flavour="scm", act_func=case["act"])
    out = dmp_oracle.lrp_pool(nv, ne, P["lrp_weight"], P.get("lrp_bias"), mats["pool"], mats["n2p"], mats["e2p"], L)
    assert torch.equal(out, case["node_out"]) and torch.equal(ne, case["edge_out"])
    ((out * case["grad_node_out"]).sum() + (ne * case["grad_edge_out"]).sum()).backward()
    torch.testing.assert_close(xv.grad, case["grad_node_feat"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xe.grad, case["grad_edge_feat"], rtol=1e-5, atol=1e-6)
    for k, g in case["grads"].items():
        got = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        torch.testing.assert_close(got, g, rtol=1e-5, atol=2e-6, msg=lambda m: k + ": " + m)
