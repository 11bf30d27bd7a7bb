"""Layer stacks around the DMPNN convolution (rows A7 and A9 of SURVEY.md section 8).

  DMPNNRepNet   the representation loop of SubgraphCountingMatching/models/dmpnn.py:183-277
                (`create_rep_net`, `get_pattern_rep`, `get_graph_rep`): shared DMPLayers applied to the
                pattern batch (mask-fill + residual) and the graph batch (gate-multiply + residual).
  relation_mean_pool
                per-relation mean of edge states, UnsupervisedNodeClassification/Model/DMPNN/src/model.py:319-325,
                as ONE typed segment reduction instead of `num_rels` masked passes.

The mask/gate multiply and the residual add are one fused kernel per tensor (`functional.gate_residual`).
The rest of `GraphAdjModelV2` (encoders, filter net, prediction heads) is the caller and stays PyTorch.
"""
import torch
import torch.nn as nn

from . import _lib
from .functional import gate_residual, segment_reduce_two_level
from .layers import DMPLayer


class DMPNNRepNet(nn.Module):
    """`rep_net` of the reference DMPNN: `num_layers` DMPLayers of width hid_dim, keyword names as in
    dmpnn.py:183-213 (`init_neigenv`, `init_eeigenv`, `rep_dmpnn_num_mlp_layers`, `rep_dmpnn_batch_norm`,
    `rep_act_func`, `rep_dropout`, `rep_residual`)."""

    def __init__(self, hid_dim, num_layers=3, rep_residual=True, **kw):
        super().__init__()
        self.hid_dim = hid_dim
        self.rep_residual = rep_residual
        self.dmpnn = nn.ModuleList([
            DMPLayer(hid_dim, hid_dim,
                     init_neigenv=kw.get("init_neigenv", 4.0), init_eeigenv=kw.get("init_eeigenv", 4.0),
                     num_mlp_layers=kw.get("rep_dmpnn_num_mlp_layers", 2),
                     batch_norm=kw.get("rep_dmpnn_batch_norm", False),
                     act_func=kw.get("rep_act_func", "relu"), dropout=kw.get("rep_dropout", 0.0))
            for _ in range(num_layers)])

    def _loop(self, graph, v, e, v_scale, e_scale):
        for layer in self.dmpnn:
            nv, ne = layer(graph, v, e)
            res = self.rep_residual and nv.shape == v.shape and ne.shape == e.shape
            if v_scale is not None or res:
                nv = gate_residual(nv, v_scale, v if res else None)
            if e_scale is not None or res:
                ne = gate_residual(ne, e_scale, e if res else None)
            v, e = nv, ne
        return v, e

    def get_pattern_rep(self, pattern, p_v_emb, p_e_emb, v_mask=None, e_mask=None):
        """dmpnn.py:215-243: masked positions are zero-filled before and after every layer."""
        vm = v_mask.reshape(-1).float() if v_mask is not None else None
        em = e_mask.reshape(-1).float() if e_mask is not None else None
        v = gate_residual(p_v_emb, vm) if vm is not None else p_v_emb
        e = gate_residual(p_e_emb, em) if em is not None else p_e_emb
        return self._loop(pattern, v, e, vm, em)

    def get_graph_rep(self, graph, g_v_emb, g_e_emb, v_mask=None, e_mask=None, v_gate=None, e_gate=None):
        """dmpnn.py:245-277: gate = mask.float() * gate; multiply before and after every layer."""
        def combine(mask, gate):
            if mask is None and gate is None:
                return None
            if gate is None:
                return mask.reshape(-1).float()
            if mask is not None:
                return (mask.float() * gate).reshape(-1)
            return gate.reshape(-1).float()

        vg, eg = combine(v_mask, v_gate), combine(e_mask, e_gate)
        v = gate_residual(g_v_emb, vg) if vg is not None else g_v_emb
        e = gate_residual(g_e_emb, eg) if eg is not None else g_e_emb
        return self._loop(graph, v, e, vg, eg)


_rel_index_cache = {}


def _relation_index(rel, num_rels):
    """(position order, int32 indptr, counts) of the relation ids, cached per tensor: the edge types of a graph do not
    change between steps, and keeping the sort / bincount (which synchronise) out of the step makes it capturable."""
    key = (rel.data_ptr(), rel._version, int(rel.numel()), int(num_rels), str(rel.device))
    hit = _rel_index_cache.get(key)
    if hit is None:
        r = rel.reshape(-1).to(torch.int64)
        order = torch.argsort(r, stable=True)
        counts = torch.bincount(r, minlength=num_rels)
        indptr = torch.zeros(num_rels + 1, dtype=torch.int32, device=rel.device)
        indptr[1:] = torch.cumsum(counts, 0)
        if len(_rel_index_cache) > 8:
            _rel_index_cache.clear()
        hit = _rel_index_cache[key] = (order.to(torch.int32), indptr, counts, rel)   # rel kept alive: the key is its address
    return hit


class _RelationPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, rel, num_rels):
        _lib.require_cuda(z, rel)
        order, indptr, counts, _ = _relation_index(rel, num_rels)
        # num_rels segments of ~E/num_rels rows each: chunked so that the whole GPU works on them (the reference's
        # `masked_fill(...).sum(dim=0)` has no sequential order to preserve)
        sums = segment_reduce_two_level(indptr, order, z, z.shape[1], chunk=512, tag="segment_reduce.rel_pool")
        denom = counts.to(z.dtype) + 1e-8
        ctx.save_for_backward(rel.reshape(-1).to(torch.int64), denom)
        return sums / denom.unsqueeze(-1)

    @staticmethod
    def backward(ctx, g):
        rel, denom = ctx.saved_tensors
        return (g / denom.unsqueeze(-1))[rel], None, None


def relation_mean_pool(z, rel, num_rels):
    """r_i = sum_{e: type(e)=i} z_e / (count_i + 1e-8)   (model.py:319-325)."""
    return _RelationPool.apply(z, rel, num_rels)
