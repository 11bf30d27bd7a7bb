"""tcgen05 3xTF32 projection GEMM (dmp_gemm_tf32x3) against fp64 and against cuBLAS sgemm.

Floating point: the bar is fp32-level accuracy -- max-norm relative error vs an fp64 product <= 2.5e-6 (north_star
tolerance is rel 1e-5) and within 4x of cuBLAS sgemm's own error (+5e-7): measured 1.2e-6 vs 5.7e-7, the tensor
core's fp32 accumulator truncates where SIMT FMA rounds.  Epilogues are compared with torch fp32 ops."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _err(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M", [1, 127, 128, 129, 1000, 148 * 128 + 5, 300_000])
@pytest.mark.parametrize("N,K", [(128, 128), (64, 64), (128, 64), (64, 128)])
def test_gemm_matches_fp64(M, N, K):
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Wt = torch.randn(N, K, device="cuda", generator=g) / 4
    ref = A.double() @ Wt.double().t()
    got = F.gemm_tf32x3(A, Wt)
    e_ours, e_cublas = _err(got, ref), _err(A @ Wt.t(), ref)
    # cross-terms-first kernel (tf32x3_gemm_v2.cu): the truncating fp32 accumulator of the tensor core
    # then costs about what an FMA GEMM's rounding costs (model: 6.6e-7 at K = 128)
    assert e_ours <= 1.0e-6, e_ours
    if M >= 128:  # a single row's sgemm error is luck; compare the two backends on real tiles only
        assert e_ours <= 4 * e_cublas + 5e-7, (e_ours, e_cublas)


def test_gemm_wide_dynamic_range_and_special_values():
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(5000, 128, device="cuda", generator=g) * torch.logspace(-6, 6, 128, device="cuda")
    A[7] = 0.0
    Wt = torch.randn(128, 128, device="cuda", generator=g)
    ref = A.double() @ Wt.double().t()
    got = F.gemm_tf32x3(A, Wt)
    rowscale = ref.abs().max(dim=1, keepdim=True).values.clamp_min(1e-30)
    assert float(((got.double() - ref).abs() / rowscale).max()) <= 2.5e-6
    assert torch.all(got[7] == 0)


@pytest.mark.parametrize("act,slope", [("none", 0.0), ("relu", 0.0), ("leaky_relu", 1 / 5.5), ("tanh", 0.0)])
def test_gemm_epilogues(act, slope):
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 3001, 128, 128
    A = torch.randn(M, K, device="cuda", generator=g)
    Wt = torch.randn(N, K, device="cuda", generator=g) / 4
    bias = torch.randn(N, device="cuda", generator=g)
    scale = torch.rand(M, device="cuda", generator=g) * 10
    plain = F.gemm_tf32x3(A, Wt)
    fn = {"none": lambda x: x, "relu": torch.relu, "tanh": torch.tanh,
          "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, slope)}[act]
    # bias + activation: same accumulator, then the same fp32 ops as torch
    got = F.gemm_tf32x3(A, Wt, bias=bias, act=act, slope=slope)
    torch.testing.assert_close(got, fn(plain + bias), rtol=1e-6, atol=1e-6)
    # row scale is applied to A as an individually rounded fp32 product
    got = F.gemm_tf32x3(A, Wt, row_scale=scale)
    assert torch.equal(got, F.gemm_tf32x3(scale.unsqueeze(1) * A, Wt))
    # accumulate into an existing D
    D0 = torch.randn(M, N, device="cuda", generator=g)
    D = D0.clone()
    F.gemm_tf32x3(A, Wt, out=D, accumulate=True)
    assert torch.equal(D, D0 + plain)
    # accumulate with a row scale: the accumulated row is scaled (exactly D0 + s * (A W^T))
    D = D0.clone()
    F.gemm_tf32x3(A, Wt, out=D, accumulate=True, row_scale=scale)
    want = D0 + scale.unsqueeze(1) * plain
    assert torch.equal(D, want)
    # act'(output) multiply (MLP backward)
    y = fn(torch.randn(M, N, device="cuda", generator=g))
    got = F.gemm_tf32x3(A, Wt, act=act, slope=slope, aux=y, mul_act_grad=True)
    grad = {"none": torch.ones_like(y), "relu": (y > 0).float(), "tanh": 1 - y * y,
            "leaky_relu": torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))}[act]
    torch.testing.assert_close(got, plain * grad, rtol=1e-6, atol=1e-6)


def test_gemm_strided_and_deterministic():
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(2)
    big = torch.randn(4000, 256, device="cuda", generator=g)
    A = big[:, 128:]                      # lda = 256
    Wt = torch.randn(128, 128, device="cuda", generator=g)
    out = torch.zeros(4000, 256, device="cuda")
    F.gemm_tf32x3(A, Wt, out=out[:, :128])  # ldd = 256
    ref = A.contiguous().double() @ Wt.double().t()
    assert _err(out[:, :128], ref) <= 2.5e-6 and torch.all(out[:, 128:] == 0)
    again = F.gemm_tf32x3(A, Wt)
    assert torch.equal(again, out[:, :128])


@pytest.mark.parametrize("E", [1, 31, 32, 33, 5000, 148 * 32 * 64 + 77, 1_000_000])
@pytest.mark.parametrize("M,N", [(128, 128), (64, 64), (128, 64), (64, 128)])
def test_gemm_tn_matches_fp64(E, M, N):
    """Weight-gradient reduction X^T G (K = E long): tensor-core 3xTF32 vs fp64, with cuBLAS sgemm as yardstick."""
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(E + M + N)
    X = torch.randn(E, M, device="cuda", generator=g)
    G = torch.randn(E, N, device="cuda", generator=g)
    ref = X.double().t() @ G.double()
    got = F.gemm_tn_tf32x3(X, G)
    scale = float(ref.abs().max()) + 1e-30
    e_ours = float((got.double() - ref).abs().max()) / scale
    e_cublas = float(((X.t() @ G).double() - ref).abs().max()) / scale
    # random +-1-ish data: the result is a sqrt(E)-sized sum, both backends sit at a few 1e-6 of max|ref|
    assert e_ours <= 1e-5, (e_ours, e_cublas)
    assert e_ours <= 3 * e_cublas + 1e-6, (e_ours, e_cublas)
    assert torch.equal(got, F.gemm_tn_tf32x3(X, G))  # deterministic


def test_gemm_tn_row_scale_accumulate_strided():
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(5)
    E = 70001
    X = torch.randn(E, 256, device="cuda", generator=g)[:, 128:]
    G = torch.randn(E, 128, device="cuda", generator=g)
    c = torch.rand(E, device="cuda", generator=g) * 8
    D0 = torch.randn(128, 128, device="cuda", generator=g)
    D = D0.clone()
    F.gemm_tn_tf32x3(X, G, row_scale=c, out=D, accumulate=True)
    ref = D0.double() + (c.double().unsqueeze(1) * X.double()).t() @ G.double()
    assert float((D.double() - ref).abs().max() / ref.abs().max()) <= 1e-5


@pytest.mark.parametrize("M,N", [(128, 128), (64, 64), (128, 64)])
def test_gemm_tn_column_sums(M, N):
    """Bias gradients ride along: column sums of both streamed operands (unscaled X, G)."""
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(M + N)
    E = 123457
    X = torch.randn(E, M, device="cuda", generator=g) + 0.5
    G = torch.randn(E, N, device="cuda", generator=g) - 0.25
    c = torch.rand(E, device="cuda", generator=g)
    D, sx, sg = F.gemm_tn_tf32x3(X, G, row_scale=c, colsum_x=True, colsum_g=True)
    torch.testing.assert_close(sx.double(), X.double().sum(0), rtol=1e-6, atol=1e-6 * float(X.double().sum(0).abs().max()))
    torch.testing.assert_close(sg.double(), G.double().sum(0), rtol=1e-6, atol=1e-6 * float(G.double().sum(0).abs().max()))
    ref = (c.double().unsqueeze(1) * X.double()).t() @ G.double()
    assert float((D.double() - ref).abs().max() / ref.abs().max()) <= 1e-5


def test_tile_scheduler_static_and_dynamic_agree_bitwise():
    """The projection kernels draw their tiles from a global counter (DMP_GEMM_DYNAMIC, default on); which CTA computes
    which tile must not change a single bit.  The switch is read once per process, hence the subprocess."""
    import os, subprocess, sys
    code = (
        "import torch, sys; sys.path.insert(0, '.');"
        "from dualmessagepassing_b200 import functional as F;"
        "g = torch.Generator(device='cuda').manual_seed(7);"
        "A = torch.randn(200_003, 128, device='cuda', generator=g); W = torch.randn(128, 128, device='cuda', generator=g);"
        "W2 = torch.randn(128, 128, device='cuda', generator=g); c = torch.rand(200_003, device='cuda', generator=g);"
        "a = F.gemm_tf32x3(A, W, bias=W[0].contiguous(), act='leaky_relu', slope=0.2);"
        "b = F.gemm_tf32x3_dual(A, W, W2, row_scale=c, mode='store');"
        "print(float(a.double().sum()), float(a.double().abs().sum()), float(b.double().sum()), float(b.double().abs().sum()))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for dyn in ("0", "1"):
        env = dict(os.environ, DMP_GEMM_DYNAMIC=dyn)
        outs.append(subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True,
                                   timeout=300).stdout.strip())
    assert outs[0] and outs[0] == outs[1], outs


def test_tile_counters_reset_across_graph_replays_and_eager_launches():
    """A launch leaves its tile counters zeroed: the same captured launches replayed many times, interleaved with eager
    launches on the same and on another stream, keep giving the eager result."""
    from dualmessagepassing_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(70_001, 64, device="cuda", generator=g)
    W1 = torch.randn(64, 64, device="cuda", generator=g)
    W2 = torch.randn(64, 64, device="cuda", generator=g)
    c = torch.rand(70_001, device="cuda", generator=g)
    want1, want2 = F.gemm_tf32x3(A, W1), F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store")
    out1, out2 = torch.empty_like(want1), torch.empty_like(want2)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):      # warm-up outside capture (lazy module state)
            F.gemm_tf32x3(A, W1, out=out1)
            F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store", out=out2)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        F.gemm_tf32x3(A, W1, out=out1)
        F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store", out=out2)
        F.gemm_tf32x3(A, W1, out=out1)
    for i in range(6):
        out1.zero_(); out2.zero_()
        graph.replay()
        if i % 2:
            with torch.cuda.stream(side):
                other = F.gemm_tf32x3(A, W2)
            eager = F.gemm_tf32x3(A, W1)
            assert torch.equal(eager, want1)
        torch.cuda.synchronize()
        assert torch.equal(out1, want1) and torch.equal(out2, want2), i
        if i % 2:
            assert torch.equal(other, F.gemm_tf32x3(A, W2))
