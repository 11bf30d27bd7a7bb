"""Ragged <-> padded helpers of the model skeleton around the layer (row N2 of SURVEY.md section 8f).

The reference pads the per-graph slices of the rep-net outputs with a Python loop over the batch and a `.tolist()`
device->host synchronisation (`split_and_batchify_graph_feats`, SubgraphCountingMatching/utils/dl.py:51-81), builds
length masks with another loop (`batch_convert_len_to_mask`, dl.py:113-127) and reads seven loss scalars back with
`.item()` every step (train.py:663-669).  Same results here from one kernel / one vectorised op / one deferred copy:

  split_and_batchify_graph_feats(feats, graph_sizes, pre_pad=False, max_size=None)   -> (padded [B,max,H], mask [B,max])
  batch_convert_len_to_mask(batch_lens, max_seq_len=-1, pre_pad=False)               -> mask [B,max]
  DeferredScalars                                                                    -> one D2H copy per logging interval

`max_size` / `max_seq_len` given by the caller (the dataset knows its largest graph) keeps the call free of host
synchronisation; left at None / -1 the maximum is read back once, as the reference does.
"""
import torch

from . import _lib


class _RaggedPad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offsets, max_len, pre_pad):
        _lib.require_cuda(x, offsets)
        x2, ldx = _lib.row_major(x)
        B, H = offsets.numel() - 1, x2.shape[1]
        out = torch.empty((B * max_len, H), dtype=torch.float32, device=x.device)
        mask = torch.empty(B * max_len, dtype=torch.uint8, device=x.device)
        _lib.call("dmp_ragged_pad", x.device, _lib.ptr(x2), ldx, _lib.ptr(offsets), B, max_len, H, int(pre_pad),
                  _lib.ptr(out), H, _lib.ptr(mask), _lib.stream_ptr(x.device), tag="ragged_pad")
        ctx.save_for_backward(offsets)
        ctx.args = (B, max_len, H, int(pre_pad), x.shape[0])
        ctx.mark_non_differentiable(mask)
        return out.view(B, max_len, H), mask.view(B, max_len).view(torch.bool)

    @staticmethod
    def backward(ctx, g, _):
        (offsets,) = ctx.saved_tensors
        B, max_len, H, pre_pad, rows = ctx.args
        g2 = g.contiguous().view(B * max_len, H)
        gx = torch.empty((rows, H), dtype=torch.float32, device=g.device)
        _lib.call("dmp_ragged_unpad", g.device, _lib.ptr(g2), H, _lib.ptr(offsets), B, max_len, H, pre_pad,
                  _lib.ptr(gx), H, rows, _lib.stream_ptr(g.device), tag="ragged_unpad")
        return gx, None, None, None


def split_and_batchify_graph_feats(batched_graph_feats, graph_sizes, pre_pad=False, max_size=None):
    """dl.py:51-81: [sum(sizes), H] -> ([B, max_size, H] zero-padded, mask [B, max_size] bool).  Differentiable."""
    sizes = graph_sizes.reshape(-1).to(torch.int64)
    if max_size is None:
        max_size = int(sizes.max().item())          # the reference's one synchronisation; pass max_size to avoid it
    offsets = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=sizes.device)
    offsets[1:] = torch.cumsum(sizes, 0)
    return _RaggedPad.apply(batched_graph_feats.float(), offsets, int(max_size), bool(pre_pad))


def batch_convert_len_to_mask(batch_lens, max_seq_len=-1, pre_pad=False):
    """dl.py:113-127 as one vectorised comparison (no per-sample loop)."""
    lens = torch.as_tensor(batch_lens).reshape(-1)
    if max_seq_len == -1:
        max_seq_len = int(lens.max().item())
    pos = torch.arange(max_seq_len, device=lens.device).unsqueeze(0)
    if pre_pad:
        return pos >= (max_seq_len - lens).unsqueeze(1)
    return pos < lens.unsqueeze(1)


class DeferredScalars:
    """The reference reads 7 scalars per step with `.item()` (train.py:663-669): 7 synchronisations.  Collect device
    scalars here and read them back with ONE copy whenever they are actually logged."""

    def __init__(self):
        self._names, self._vals = [], []

    def add(self, **scalars):
        for k, v in scalars.items():
            self._names.append(k)
            self._vals.append(v.detach().reshape(()).float())

    def fetch(self):
        """{name: [values...]} of everything added since the last fetch (one device->host copy)."""
        if not self._vals:
            return {}
        host = torch.stack(self._vals).cpu().tolist()
        out = {}
        for k, v in zip(self._names, host):
            out.setdefault(k, []).append(v)
        self._names, self._vals = [], []
        return out
