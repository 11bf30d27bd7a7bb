"""CPU oracle for the DMPNN dual message-passing layer.

TEST INFRASTRUCTURE ONLY -- never imported by `dualmessagepassing_b200`; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it.

This is a restatement (plain PyTorch on CPU, fp32 or fp64, autograd for gradients) of what the
reference executes through DGL, written in the reference's own operation order:

  * SCM flavour  : SubgraphCountingMatching/models/dmpnn.py:96-166   (DMPLayer)
  * UNC flavour  : UnsupervisedNodeClassification/Model/DMPNN/src/model.py:206-273 (DualGraphConv)
  * rep-net loop : SubgraphCountingMatching/models/dmpnn.py:215-277  (mask / gate / residual)
  * UNC pooling  : UnsupervisedNodeClassification/Model/DMPNN/src/model.py:319-325

DGL semantics assumed: UDFs see edges in edge-id order, `fn.sum`
accumulates each destination's messages sequentially in ascending edge id.

Parity status: PINNED against the unmodified reference classes executed over a DGL shim
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).
The DGL runtime itself is not available offline, so the five DGL entry points are restated in
`tests/golden/_dgl_shim.py`; that part of the chain is unpinned (stated in DESIGN.md).
"""
import math

import torch
import torch.nn.functional as F

LEAKY_RELU_A = 1 / 5.5  # SubgraphCountingMatching/constants.py:10


def activation(name):
    """Subset of SubgraphCountingMatching/utils/act.py:457-474 that is elementwise."""
    table = {
        "none": lambda x: x,
        "relu": F.relu,
        "relu6": F.relu6,
        "tanh": torch.tanh,
        "sigmoid": torch.sigmoid,
        "leaky_relu": lambda x: F.leaky_relu(x, LEAKY_RELU_A),
        "elu": F.elu,
        "celu": F.celu,
        "selu": F.selu,
        "gelu": F.gelu,
    }
    return table[name]


def _mlp(x, params, prefix, act, training=True):
    """nn.Sequential(Linear, [BatchNorm1d], act, Linear) or empty; keys as in the state_dict.

    dmpnn.py:45-60 (SCM), model.py:144-167 (UNC).
    """
    idx = sorted({int(k.split(".")[1]) for k in params if k.startswith(prefix + ".")})
    if not idx:
        return None
    n_lin = 0
    out = x
    last = idx[-1]
    for i in idx:
        w = params["%s.%d.weight" % (prefix, i)]
        b = params.get("%s.%d.bias" % (prefix, i))
        if w.dim() == 2:  # Linear
            out = F.linear(out, w, b)
            n_lin += 1
            if i != last and ("%s.%d.weight" % (prefix, i + 1)) not in params:
                out = act(out)  # no BN between this Linear and the activation
        else:  # BatchNorm1d followed by the activation
            rm = params.get("%s.%d.running_mean" % (prefix, i))
            rv = params.get("%s.%d.running_var" % (prefix, i))
            if training:
                out = F.batch_norm(out, None, None, w, b, True, 0.1, 1e-5)
            else:
                out = F.batch_norm(out, rm, rv, w, b, False, 0.1, 1e-5)
            out = act(out)
    return out


def dmp_layer(params, src, dst, num_nodes, node_feat, edge_feat, *, rev=None, out_deg=None,
              norm=None, flavour="scm", act_func="relu", mlp_act=None, post_act=None,
              training=True, return_pre=False):
    """One dual message-passing layer, reference op order.

    params   : dict of tensors keyed like the reference state_dict (in_weight, ..., nmlp.0.weight, ...)
    src, dst : int64 [E] (DGL edge-id order);  rev: bool [E] or None;  out_deg: int64 [N] or None
    norm     : [E,1] or None (UNC only, model.py:234-235)
    flavour  : "scm" -> dmpnn.py association `(eloop + add) + agg`, act when no MLP
               "unc" -> model.py:257 association `(eloop + agg) + add`, MLP always, optional post_act
    """
    X_v, X_e = node_feat, edge_feat
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    N = int(num_nodes)
    if out_deg is None:  # dmpnn.py:100-101
        out_deg = torch.bincount(src, minlength=N)
    W = params
    hs, hd = X_v[src], X_v[dst]

    # ---- message UDF: dmpnn.py:111-127 / model.py:222-238 --------------------------------
    edge_msg = hd @ W["dst_weight"] - hs @ W["src_weight"]
    node_msg = -(X_e @ W["in_weight"])
    if rev is not None:
        rmask = torch.as_tensor(rev, dtype=torch.bool).view(-1, 1)
        mask = ~rmask
        rev_edge_msg = hs @ W["dst_weight"] - hd @ W["src_weight"]
        rev_node_msg = X_e @ W["out_weight"]
        edge_msg = edge_msg.masked_fill(rmask, 0.0) + rev_edge_msg.masked_fill(mask, 0.0)
        node_msg = node_msg.masked_fill(rmask, 0.0) + rev_node_msg.masked_fill(mask, 0.0)
    if norm is not None:
        node_msg = node_msg * norm.view(-1, 1)

    # ---- fn.sum over CSC, sequential in edge-id order: dmpnn.py:92,163 ---------------------
    node_agg = torch.zeros((N, node_msg.shape[1]), dtype=node_msg.dtype).index_add(0, dst, node_msg)

    # ---- node update: dmpnn.py:129-140 / model.py:240-250 ----------------------------------
    node_pre = X_v @ W["nloop_weight"] + node_agg
    if W.get("nbias") is not None:
        node_pre = node_pre + W["nbias"]

    # ---- edge update: dmpnn.py:142-156 / model.py:252-265 ----------------------------------
    d = out_deg[dst].unsqueeze(-1).to(X_e.dtype)
    d = (1 + d).log2()
    add = 2 * (1 + d) * (X_e @ (W["src_weight"] - W["dst_weight"]))
    if flavour == "scm":
        edge_pre = X_e @ W["eloop_weight"] + add + edge_msg
    else:
        edge_pre = X_e @ W["eloop_weight"] + edge_msg + add
    if W.get("ebias") is not None:
        edge_pre = edge_pre + W["ebias"]

    act = activation(act_func) if isinstance(act_func, str) else act_func
    m_act = act if mlp_act is None else (activation(mlp_act) if isinstance(mlp_act, str) else mlp_act)
    node_out = _mlp(node_pre, W, "nmlp", m_act, training)
    edge_out = _mlp(edge_pre, W, "emlp", m_act, training)
    if flavour == "scm":
        if node_out is None:
            node_out = act(node_pre)
        if edge_out is None:
            edge_out = act(edge_pre)
    else:
        if post_act is not None:
            p = activation(post_act) if isinstance(post_act, str) else post_act
            node_out, edge_out = p(node_out), p(edge_out)
    if return_pre:
        return node_out, edge_out, node_pre, edge_pre
    return node_out, edge_out


def graph_rep(layer_params, src, dst, num_nodes, v_emb, e_emb, *, rev=None, out_deg=None,
              v_gate=None, e_gate=None, residual=True, **layer_kw):
    """dmpnn.py:245-277 -- gate-multiply + residual loop over layers (graph side)."""
    v = v_emb * v_gate if v_gate is not None else v_emb
    e = e_emb * e_gate if e_gate is not None else e_emb
    for params in layer_params:
        nv, ne = dmp_layer(params, src, dst, num_nodes, v, e, rev=rev, out_deg=out_deg, **layer_kw)
        if v_gate is not None:
            nv = nv * v_gate
        if e_gate is not None:
            ne = ne * e_gate
        if residual and nv.shape == v.shape and ne.shape == e.shape:
            v, e = v + nv, e + ne
        else:
            v, e = nv, ne
    return v, e


def pattern_rep(layer_params, src, dst, num_nodes, v_emb, e_emb, *, rev=None, out_deg=None,
                v_mask=None, e_mask=None, residual=True, **layer_kw):
    """dmpnn.py:215-243 -- masked_fill + residual loop over layers (pattern side)."""
    vz = ~v_mask if v_mask is not None else None
    ez = ~e_mask if e_mask is not None else None
    v = v_emb.masked_fill(vz, 0.0) if vz is not None else v_emb
    e = e_emb.masked_fill(ez, 0.0) if ez is not None else e_emb
    for params in layer_params:
        nv, ne = dmp_layer(params, src, dst, num_nodes, v, e, rev=rev, out_deg=out_deg, **layer_kw)
        if vz is not None:
            nv = nv.masked_fill(vz, 0.0)
        if ez is not None:
            ne = ne.masked_fill(ez, 0.0)
        if residual and nv.shape == v.shape and ne.shape == e.shape:
            v, e = v + nv, e + ne
        else:
            v, e = nv, ne
    return v, e


def lrp_pool(node_out, edge_out, lrp_weight, lrp_bias, pooling_matrix, node_to_perm, edge_to_perm, seq_len):
    """SubgraphCountingMatching/models/dmplrp.py:180-185 -- local relational pooling after the dual update: three
    torch.sparse products and one contraction over (sequence position, input feature)."""
    z = torch.sparse.mm(node_to_perm, node_out) + torch.sparse.mm(edge_to_perm, edge_out)
    z = z.view(-1, seq_len * seq_len, lrp_weight.shape[0])
    y = torch.einsum("dab,bca->dc", z, lrp_weight)
    if lrp_bias is not None:
        y = y + lrp_bias
    return torch.sparse.mm(pooling_matrix, y)


def relation_mean_pool(z, rel, num_rels):
    """model.py:319-325 -- per-relation masked mean of edge states."""
    rows = []
    for i in range(num_rels):
        rows.append(z.masked_fill((rel != i).view(-1, 1), 0.0).sum(dim=0, keepdim=True)
                    / ((rel == i).sum().to(z.dtype) + 1e-8))
    return torch.cat(rows, dim=0)


# ---- constructor parity: SubgraphCountingMatching/utils/init.py:17-75,125-193 -------------------
def scm_gain(act_func):
    if act_func in ("none", "maximum", "minimum"):
        nl = "linear"
    elif act_func in ("relu", "relu6", "elu", "selu", "celu", "gelu"):
        nl = "relu"
    elif act_func in ("leaky_relu", "prelu"):
        nl = "leaky_relu"
    elif act_func in ("softmax", "sparsemax", "gumbel_softmax"):
        nl = "sigmoid"
    else:
        nl = act_func
    return torch.nn.init.calculate_gain(nl, LEAKY_RELU_A)


def scm_uniform_bound(shape, act_func):
    fan = shape[0] + shape[1]
    return 1.7320508075688772 * scm_gain(act_func) * math.sqrt(2.0 / float(fan))
