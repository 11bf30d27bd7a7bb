"""dmp_gemm_tf32x3_dual (two projections of one streamed operand), dmp_bn_* (BatchNorm1d of the MLPs) and the
P = NULL form of dmp_edge_update, through the C ABI.

Bars: the dual kernel must equal the two-launch composition it replaces BIT FOR BIT (same MMA sequence per
output element, same epilogue roundings) and stay within the 3xTF32 error bound vs fp64 everywhere; the BatchNorm
kernels are compared with torch's fp32 batch_norm (forward, running statistics, all gradients) at rtol 1e-5 /
atol 1e-6 and must be run-to-run bit-stable."""
import pytest
import torch

from oracle import sparse_core as sc
from tests._cases import make_graph, t

pytestmark = pytest.mark.gpu


def _err(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M", [1, 127, 128, 129, 1000, 74 * 128 + 5, 148 * 128 + 77, 300_000])
@pytest.mark.parametrize("N,K", [(128, 128), (64, 64), (128, 64), (64, 128)])
def test_dual_store_and_accumulate(M, N, K):
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(3 * M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W1 = torch.randn(N, K, device="cuda", generator=g) / 4
    W2 = torch.randn(N, K, device="cuda", generator=g) / 4
    c = 2 + 6 * torch.rand(M, device="cuda", generator=g)
    ref = A.double() @ W1.double().t() + c.double().unsqueeze(1) * (A.double() @ W2.double().t())
    got = F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store")
    assert _err(got, ref) <= 1.5e-6, _err(got, ref)
    p1, p2 = F.gemm_tf32x3(A, W1), F.gemm_tf32x3(A, W2)
    want = p1 + c.unsqueeze(1) * p2
    assert torch.equal(got, want)        # same accumulators, same roundings as the two launches it replaces
    # accumulate: (old + acc1) + c * acc2
    D0 = torch.randn(M, N, device="cuda", generator=g)
    D = D0.clone()
    F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="accumulate", out=D)
    want = (D0 + p1) + c.unsqueeze(1) * p2
    assert torch.equal(D, want)
    # no scale = scale 1
    got1 = F.gemm_tf32x3_dual(A, W1, W2, mode="store")
    assert torch.equal(got1, p1 + p2)
    # separate outputs
    s1, s2 = F.gemm_tf32x3_dual(A, W1, W2, mode="separate")
    assert torch.equal(s1, p1) and torch.equal(s2, p2)


def test_dual_strided_operands_and_determinism():
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(9)
    M = 20_001
    big = torch.randn(M, 256, device="cuda", generator=g)
    A = big[:, 64:192]                      # lda = 256
    W1 = torch.randn(128, 128, device="cuda", generator=g) / 4
    W2 = torch.randn(128, 128, device="cuda", generator=g) / 4
    c = torch.rand(M, device="cuda", generator=g)
    out = torch.zeros(M, 256, device="cuda")
    r = [F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store") for _ in range(3)]
    assert torch.equal(r[0], r[1]) and torch.equal(r[0], r[2])
    F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store", out=out[:, 128:])      # ldd = 256
    assert torch.equal(out[:, 128:], r[0]) and float(out[:, :128].abs().max()) == 0.0
    assert torch.equal(r[0], F.gemm_tf32x3_dual(A.contiguous(), W1, W2, row_scale=c, mode="store"))


@pytest.mark.parametrize("H", [128, 64, 50])
@pytest.mark.parametrize("rev", ["halves", "shuffled", None])
def test_edge_update_without_P_equals_precombined(H, rev):
    """P = NULL: S already holds eloop + coef*P -> out = (S + msg) + ebias, equal to the SCM three-operand form."""
    import dualmessagepassing_b200 as dmp
    from dualmessagepassing_b200 import _lib, functional as F
    from dualmessagepassing_b200.constants import REVFLAG
    n, e0 = 700, 5000
    s, d, r = make_graph(seed=H, n=n, e0=e0, rev=rev)
    E = len(s)
    g = dmp.DMPGraph(s, d, n, device="cuda")
    if r is not None:
        g.edata[REVFLAG] = t(r)
    plan = dmp.get_plan(g, REVFLAG, "out_deg")
    gen = torch.Generator(device="cuda").manual_seed(1)
    S, P = torch.randn(E, H, device="cuda", generator=gen), torch.randn(E, H, device="cuda", generator=gen)
    Qd, Qs = torch.randn(n, H, device="cuda", generator=gen), torch.randn(n, H, device="cuda", generator=gen)
    b = torch.randn(H, device="cuda", generator=gen)
    want = F.edge_update(plan, S, P, Qd, Qs, b, _lib.ORDER_SCM)
    U = S + plan.coef.unsqueeze(1) * P                       # fl(S + fl(c*P)): what DUAL_STORE writes
    got = F.edge_update(plan, U, None, Qd, Qs, b, _lib.ORDER_SCM)
    assert torch.equal(got, want)
    # against the sequential C oracle too
    ref = sc.edge_update(plan.a32.cpu(), plan.b32.cpu(), plan.coef.cpu(), S.cpu(), P.cpu(), Qd.cpu(), Qs.cpu(), b.cpu(),
                         order=0)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("rows,H", [(2, 8), (63, 50), (4097, 64), (100_000, 128), (20_000, 300)])
@pytest.mark.parametrize("act", ["leaky_relu", "tanh", "none"])
def test_bn_kernels_match_torch(rows, H, act):
    from dualmessagepassing_b200 import _lib, functional as F
    g = torch.Generator(device="cuda").manual_seed(rows + H)
    x = torch.randn(rows, H, device="cuda", generator=g) * 3 + 5          # |mean| > std: two-pass variance matters
    gamma = torch.rand(H, device="cuda", generator=g) + 0.5
    beta = torch.randn(H, device="cuda", generator=g)
    gy = torch.randn(rows, H, device="cuda", generator=g)
    slope = 1 / 5.5
    fn = {"none": lambda v: v, "tanh": torch.tanh, "leaky_relu": lambda v: torch.nn.functional.leaky_relu(v, slope)}[act]
    mean, var = F.bn_stats(x)
    xd = x.double()
    torch.testing.assert_close(mean.double(), xd.mean(0), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(var.double(), xd.var(0, unbiased=False), rtol=2e-6, atol=1e-7)
    m2, v2 = F.bn_stats(x)
    assert torch.equal(mean, m2) and torch.equal(var, v2)               # deterministic
    invstd = torch.rsqrt(var + 1e-5)
    got = F.bn_act(x, mean, invstd, gamma, beta, {"none": 0, "leaky_relu": 2, "tanh": 3}[act], slope)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    want = torch.nn.functional.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    torch.testing.assert_close(got, fn(want).detach(), rtol=1e-5, atol=5e-6)   # |x| up to ~15: 3 ulp
    # backward of the normalisation alone (the activation's derivative rides on the producing GEMM's epilogue)
    if rows > 1:
        want.backward(gy)
        gx, dgamma, dbeta = F.bn_backward(gy.clone(), x, mean, invstd, gamma, True)
        sx = float(xr.grad.abs().max())
        torch.testing.assert_close(gx, xr.grad, rtol=1e-4, atol=2e-6 * max(1.0, sx))
        torch.testing.assert_close(dgamma, gr.grad, rtol=1e-5, atol=2e-5 * max(1.0, float(gr.grad.abs().max())))
        torch.testing.assert_close(dbeta, br.grad, rtol=1e-5, atol=2e-5 * max(1.0, float(br.grad.abs().max())))
        # eval mode: fixed statistics
        gx2, _, _ = F.bn_backward(gy.clone(), x, mean, invstd, gamma, False)
        torch.testing.assert_close(gx2, gy * (gamma * invstd), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("flavour", ["scm", "unc"])
@pytest.mark.parametrize("width", [50, 64])
def test_bn_layer_updates_running_stats_like_torch_and_eval_mode(flavour, width):
    """The fused BatchNorm path keeps nn.BatchNorm1d's side effects (running_mean / running_var / num_batches_tracked)
    and its eval-mode arithmetic; compared with the composed path (torch's own BatchNorm) on the same module."""
    import copy
    import dualmessagepassing_b200 as dmp
    from dualmessagepassing_b200 import fused
    from dualmessagepassing_b200.constants import REVFLAG
    n, e0 = 900, 4000
    s, d, r = make_graph(seed=width, n=n, e0=e0, rev="halves")
    E = len(s)
    torch.manual_seed(width)
    if flavour == "scm":
        layer = dmp.DMPLayer(width, width, num_mlp_layers=2, batch_norm=True, act_func="leaky_relu").cuda()
    else:
        layer = dmp.DualGraphConv(width, width, batch_norm=True, activation=torch.nn.Tanh()).cuda()
    ref = copy.deepcopy(layer)
    ref.fused = False
    g = dmp.DMPGraph(s, d, n, device="cuda")
    g.edata[REVFLAG if flavour == "scm" else "is_rev"] = t(r)
    xv, xe = torch.randn(n, width, device="cuda"), torch.randn(E, width, device="cuda")
    old = fused.TC_MIN_ROWS
    fused.TC_MIN_ROWS = 1
    try:
        for mod in (layer, ref):
            mod.train()
        for step in range(2):
            a = layer(g, xv + step, xe - step)
            b = ref(g, xv + step, xe - step)
            for x, y in zip(a, b):
                torch.testing.assert_close(x, y, rtol=2e-5, atol=2e-5 * float(y.abs().max()))
        for seq_a, seq_b in ((layer.nmlp, ref.nmlp), (layer.emlp, ref.emlp)):
            assert int(seq_a[1].num_batches_tracked) == int(seq_b[1].num_batches_tracked) == 2
            torch.testing.assert_close(seq_a[1].running_mean, seq_b[1].running_mean, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(seq_a[1].running_var, seq_b[1].running_var, rtol=1e-5, atol=1e-6)
        layer.eval()
        ref.eval()
        a0, b0 = xv.clone().requires_grad_(True), xe.clone().requires_grad_(True)
        a1, b1 = xv.clone().requires_grad_(True), xe.clone().requires_grad_(True)
        oa, ob = layer(g, a0, b0), ref(g, a1, b1)
        (oa[0].sum() + (oa[1] ** 2).sum()).backward()
        (ob[0].sum() + (ob[1] ** 2).sum()).backward()
        for x, y in list(zip(oa, ob)) + [(a0.grad, a1.grad), (b0.grad, b1.grad)]:
            torch.testing.assert_close(x, y, rtol=2e-5, atol=2e-5 * float(y.abs().max()))
    finally:
        fused.TC_MIN_ROWS = old


def test_fused_layer_rejects_mismatched_feature_rows():
    """the whole-layer path must raise (like DGL's frame-size check), not read out of bounds."""
    import dualmessagepassing_b200 as dmp
    from dualmessagepassing_b200.constants import REVFLAG
    s, d, r = make_graph(seed=3, n=100, e0=400, rev="halves")
    g = dmp.DMPGraph(s, d, 100, device="cuda")
    g.edata[REVFLAG] = t(r)
    layer = dmp.DMPLayer(16, 16, num_mlp_layers=2, batch_norm=False, act_func="relu").cuda()
    xv, xe = torch.randn(100, 16, device="cuda"), torch.randn(len(s), 16, device="cuda")
    with pytest.raises(ValueError):
        layer(g, xv, xe[: len(s) // 2])          # edge features not doubled after add_reversed_edges
    with pytest.raises(ValueError):
        layer(g, xv[:50], xe)
    with pytest.raises(ValueError):
        layer(g, torch.randn(100, 8, device="cuda"), xe)
