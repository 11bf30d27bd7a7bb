"""Host-side helpers around the layer that have a CPU path (no GPU needed): batched zero-padding of a layer's parameter
tensors, the label-embedding function and the row Linear of the training harness."""
import torch

from dualmessagepassing_b200.layers import _pad_cols, _pad_many, _pad_mat, _padded_width
from dualmessagepassing_b200.train_step import label_embedding, row_linear


def test_padded_width_rule():
    assert [_padded_width(d) for d in (16, 32, 33, 50, 64, 65, 100, 128, 129)] == [None, None, 64, 64, 64, 128, 128, 128, None]


def test_pad_many_equals_per_tensor_padding_forward_and_backward():
    torch.manual_seed(0)
    ts = [torch.randn(50, 50, requires_grad=True), None, torch.randn(50, requires_grad=True),
          torch.randn(50, 50, requires_grad=True), torch.randn(50, requires_grad=True)]
    tg = [(64, 64), None, (64,), (64, 64), (64,)]
    out = _pad_many(ts, tg)
    ref = [_pad_mat(ts[0], 64, 64), None, _pad_cols(ts[2], 64), _pad_mat(ts[3], 64, 64), _pad_cols(ts[4], 64)]
    for o, r in zip(out, ref):
        assert (o is None and r is None) or torch.equal(o, r)
    w0, w2 = torch.randn(64, 64), torch.randn(64)
    ((out[0] * w0).sum() + (out[2] * w2).sum()).backward()        # out[3], out[4] unused: zero gradients
    assert torch.equal(ts[0].grad, w0[:50, :50]) and torch.equal(ts[2].grad, w2[:50])
    assert torch.equal(ts[3].grad, torch.zeros(50, 50)) and torch.equal(ts[4].grad, torch.zeros(50))
    # the index maps are cached per shape signature and reused
    again = _pad_many([t.detach() if t is not None else None for t in ts], tg)
    assert torch.equal(again[0], ref[0].detach())


def test_label_embedding_cpu_matches_nn_embedding():
    torch.manual_seed(1)
    labels = torch.randint(0, 7, (300,))
    w1 = torch.randn(7, 16, requires_grad=True)
    w2 = w1.detach().clone().requires_grad_(True)
    up = torch.randn(300, 16)
    label_embedding(w1, labels).backward(up)
    torch.nn.functional.embedding(labels, w2).backward(up)
    torch.testing.assert_close(w1.grad, w2.grad, rtol=1e-5, atol=1e-5)


def test_row_linear_cpu_is_the_module():
    torch.manual_seed(2)
    lin = torch.nn.Linear(8, 8)
    x = torch.randn(5, 8)
    assert torch.equal(row_linear(x, lin), lin(x))
