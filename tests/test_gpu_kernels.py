"""Sparse-core kernels through the C ABI: fp32 BIT-EXACT against the sequential C oracle."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import sparse_core as sc
from tests._cases import hub_graph, make_graph, t

pytestmark = pytest.mark.gpu

SHAPES = [  # (n, e0, H) -- H covers every (VEC, G, ITER) dispatch: 4..512+, odd, 50 (UNC run.sh), 64/128
    (30, 100, 4), (30, 100, 7), (50, 300, 32), (64, 500, 50), (200, 2000, 64), (300, 3000, 128),
    (40, 200, 130), (40, 200, 256), (20, 100, 516), (20, 60, 1030),
]


def _plan(s, d, n, r):
    from dualmessagepassing_b200.plan import DMPPlan
    return DMPPlan(t(s), t(d), n, rev=t(None if r is None else r.astype(np.uint8)))


def _cpu_plan(plan):
    return {k: getattr(plan, k).cpu() for k in ("dst32", "a32", "b32", "csc_indptr", "csc_eid", "a_indptr", "a_eid",
                                                "b_indptr", "b_eid", "coef")}


@pytest.mark.parametrize("n,e0,H", SHAPES)
@pytest.mark.parametrize("rev", ["halves", "shuffled", None])
def test_segment_reduce_bit_exact(n, e0, H, rev):
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=n + H, n=n, e0=e0, rev=rev)
    plan = _plan(s, d, n, r)
    cp = _cpu_plan(plan)
    E = len(s)
    g = torch.Generator().manual_seed(H)
    two_branch = rev == "shuffled"
    M = torch.randn(E, 2 * H if two_branch else H, generator=g)
    base, bias, norm = torch.randn(n, H, generator=g), torch.randn(H, generator=g), torch.rand(E, generator=g)
    off = H if two_branch else 0
    for use_norm in (False, True):
        w_perm_cpu = norm[(cp["csc_eid"].long() & 0x7FFFFFFF)] if use_norm else None
        want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], M, H, w_perm=w_perm_cpu, rev_off=off, base=base,
                             bias=bias, mode=_lib.SEG_SIGN_BY_REV)
        w_perm = plan.norm_permuted(norm.cuda())[1] if use_norm else None
        got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, M.cuda(), H, w_perm=w_perm, rev_col_offset=off,
                               base=base.cuda(), bias=bias.cuda(), mode=_lib.SEG_SIGN_BY_REV)
        assert torch.equal(got.cpu(), want)
    # backward flavour: plain / negated sums over the a- and b-keyed segments
    gE = torch.randn(E, H, generator=g)
    for ip, ei, mode in (("a_indptr", "a_eid", 0), ("b_indptr", "b_eid", _lib.SEG_NEGATE_OUT)):
        want = sc.seg_reduce(cp[ip], cp[ei], gE, H, mode=mode)
        got = F.segment_reduce(getattr(plan, ip), getattr(plan, ei), gE.cuda(), H, mode=mode)
        assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("n,e0,H", SHAPES)
@pytest.mark.parametrize("order", [0, 1])
def test_edge_update_bit_exact(n, e0, H, order):
    from dualmessagepassing_b200 import functional as F
    s, d, r = make_graph(seed=7 * n + H, n=n, e0=e0, rev="shuffled")
    plan = _plan(s, d, n, r)
    cp = _cpu_plan(plan)
    E = len(s)
    g = torch.Generator().manual_seed(H + order)
    S, P = torch.randn(E, H, generator=g), torch.randn(E, H, generator=g)
    Qd, Qs, eb = torch.randn(n, H, generator=g), torch.randn(n, H, generator=g), torch.randn(H, generator=g)
    want, want_agg = sc.edge_update(cp["a32"], cp["b32"], cp["coef"], S, P, Qd, Qs, eb, order, want_agg=True)
    agg = torch.empty(E, H, device="cuda")
    got = F.edge_update(plan, S.cuda(), P.cuda(), Qd.cuda(), Qs.cuda(), eb.cuda(), order, edge_agg=agg)
    assert torch.equal(got.cpu(), want) and torch.equal(agg.cpu(), want_agg)
    # in place over S, no bias
    Sg = S.cuda()
    out = F.edge_update(plan, Sg, P.cuda(), Qd.cuda(), Qs.cuda(), None, order, out=Sg)
    assert out.data_ptr() == Sg.data_ptr()
    assert torch.equal(Sg.cpu(), sc.edge_update(cp["a32"], cp["b32"], cp["coef"], S, P, Qd, Qs, None, order))


@pytest.mark.parametrize("n,e0,H", SHAPES)
@pytest.mark.parametrize("order", [0, 1])
def test_edge_update_mirrored_halves_bit_exact(n, e0, H, order):
    """[forward | reversed] layout of one graph: edge e + E/2 mirrors edge e, the kernel handles the pair together and
    fetches the two endpoint rows once (DMP_EDGE_MIRRORED_HALVES).  Same bits as the oracle; a plan whose halves do NOT
    mirror each other (second half permuted) must not take the hint, and the raw hint on such data still is correct."""
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=3 * n + H, n=n, e0=e0, rev="halves")
    E = len(s)
    g = torch.Generator().manual_seed(H + order)
    S, P = torch.randn(E, H, generator=g), torch.randn(E, H, generator=g)
    Qd, Qs, eb = torch.randn(n, H, generator=g), torch.randn(n, H, generator=g), torch.randn(H, generator=g)
    plan = _plan(s, d, n, r)
    assert plan.rev_layout == "halves" and plan.mirrored_halves
    cp = _cpu_plan(plan)
    want = sc.edge_update(cp["a32"], cp["b32"], cp["coef"], S, P, Qd, Qs, eb, order)
    got = F.edge_update(plan, S.cuda(), P.cuda(), Qd.cuda(), Qs.cuda(), eb.cuda(), order)
    assert torch.equal(got.cpu(), want)
    Sg = S.cuda()                                                    # in place over S
    F.edge_update(plan, Sg, P.cuda(), Qd.cuda(), Qs.cuda(), eb.cuda(), order, out=Sg)
    assert torch.equal(Sg.cpu(), want)
    # halves that are not mirrors of each other
    h = E // 2
    perm = np.concatenate([np.arange(h), h + np.random.Generator(np.random.PCG64(n)).permutation(h)])
    plan2 = _plan(s[perm], d[perm], n, r[perm])
    assert plan2.rev_layout == "halves" and not plan2.mirrored_halves
    cp2 = _cpu_plan(plan2)
    want2 = sc.edge_update(cp2["a32"], cp2["b32"], cp2["coef"], S, P, Qd, Qs, eb, order)
    assert torch.equal(F.edge_update(plan2, S.cuda(), P.cuda(), Qd.cuda(), Qs.cuda(), eb.cuda(), order).cpu(), want2)
    plan2._mirrored = True                                           # force the hint onto non-mirrored data
    assert torch.equal(F.edge_update(plan2, S.cuda(), P.cuda(), Qd.cuda(), Qs.cuda(), eb.cuda(), order).cpu(), want2)


@pytest.mark.parametrize("n,e0,H", SHAPES[:8])
def test_edge_backward_bit_exact(n, e0, H):
    from dualmessagepassing_b200 import functional as F
    s, d, r = make_graph(seed=11 * n + H, n=n, e0=e0, rev="shuffled")
    plan = _plan(s, d, n, r)
    cp = _cpu_plan(plan)
    E = len(s)
    g = torch.Generator().manual_seed(H)
    gN, gE, norm = torch.randn(n, H, generator=g), torch.randn(E, H, generator=g), torch.rand(E, generator=g)
    r8 = torch.from_numpy(r.astype(np.uint8))
    for nm in (None, norm):
        for off in (0, H):
            wT, wCG = sc.edge_backward(cp["dst32"], r8, nm, cp["coef"], gN, gE, t_rev_off=off)
            T, CG = F.edge_backward(plan, None if nm is None else nm.cuda(), gN.cuda(), gE.cuda(),
                                    t_rev_col_offset=off)
            assert torch.equal(T.cpu(), wT) and torch.equal(CG.cpu(), wCG)
    # two gather tables (forward edges read gN, reversed edges gN_rev): the dX_e initialisation of the fused backward
    gN2 = torch.randn(n, H, generator=g)
    wT, _ = sc.edge_backward(cp["dst32"], r8, norm, cp["coef"], gN, gE, gN_rev=gN2)
    T, _ = F.edge_backward(plan, norm.cuda(), gN.cuda(), None, want_CG=False, gN_rev=gN2.cuda())
    assert torch.equal(T.cpu(), wT)


def test_long_segment_hub_bit_exact_and_deterministic():
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = hub_graph(21, 400, 2000, 20000)
    plan = _plan(s, d, 400, r)
    cp = _cpu_plan(plan)
    E, H = len(s), 128
    M = torch.randn(E, H, generator=torch.Generator().manual_seed(3))
    want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], M, H, mode=_lib.SEG_SIGN_BY_REV)
    Mg = M.cuda()
    outs = [F.segment_reduce(plan.csc_indptr, plan.csc_eid, Mg, H, mode=_lib.SEG_SIGN_BY_REV) for _ in range(3)]
    assert torch.equal(outs[0].cpu(), want)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])  # run-to-run bit stability


def test_million_edge_hub_sequential_bit_exact_and_two_level_close():
    """A 1 M-edge segment: the default path stays the strictly sequential sum (bit-exact vs the C oracle); the chunked
    two-level option is deterministic, bit-identical on segments no longer than a chunk, and within fp32 reassociation
    error of an fp64 sum on the hub."""
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = hub_graph(22, 3000, 20000, 1_000_000)
    plan = _plan(s, d, 3000, r)
    cp = _cpu_plan(plan)
    E, H = len(s), 64
    M = torch.randn(E, H, generator=torch.Generator().manual_seed(4))
    want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], M, H, mode=_lib.SEG_SIGN_BY_REV)
    Mg = M.cuda()
    got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, Mg, H, mode=_lib.SEG_SIGN_BY_REV)
    assert torch.equal(got.cpu(), want)
    for mode in (_lib.SEG_SIGN_BY_REV, _lib.SEG_SIGN_BY_REV | _lib.SEG_SPLIT_BY_REV, _lib.SEG_NEGATE_OUT):
        seq = F.segment_reduce(plan.csc_indptr, plan.csc_eid, Mg, H, mode=mode)
        two = [F.segment_reduce_two_level(plan.csc_indptr, plan.csc_eid, Mg, H, chunk=1024, mode=mode) for _ in range(2)]
        assert torch.equal(two[0], two[1])
        lens = (plan.csc_indptr[1:] - plan.csc_indptr[:-1]).long()
        short = lens <= 1024
        assert torch.equal(two[0][short], seq[short])              # one chunk per segment: the same sequence of adds
        # hub rows: compare both orders with an fp64 sum of the same terms
        hub = int(torch.argmax(lens))
        lo, hi = int(plan.csc_indptr[hub]), int(plan.csc_indptr[hub + 1])
        e = plan.csc_eid[lo:hi].long()
        rows, rv = (e & 0x7FFFFFFF), (e >> 31) & 1
        x = Mg[rows].double()
        if mode & _lib.SEG_SIGN_BY_REV:
            x = torch.where(rv.bool().unsqueeze(1), x, -x)
        if mode & _lib.SEG_SPLIT_BY_REV:
            ref = torch.cat([(x * (rv == 0).unsqueeze(1)).sum(0), (x * (rv == 1).unsqueeze(1)).sum(0)])
        else:
            ref = x.sum(0) * (-1 if mode & _lib.SEG_NEGATE_OUT else 1)
        scale = float(x.abs().sum(0).max())
        assert float((two[0][hub].double() - ref).abs().max()) <= 2e-7 * scale
        assert float((seq[hub].double() - ref).abs().max()) <= 2e-6 * scale   # 1 M sequential fp32 adds drift further


def test_short_segment_hint_is_bit_identical():
    """DMP_SEG_SHORT (many segments of ~2 rows: the partitioned graph's a-/b-keyed reductions) only changes occupancy."""
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=8, n=50_000, e0=60_000, rev="halves")
    plan = _plan(s, d, 50_000, r)
    G = torch.randn(len(s), 128, device="cuda")
    for mode in (0, _lib.SEG_NEGATE_OUT):
        a = F.segment_reduce(plan.b_indptr, plan.b_eid, G, 128, mode=mode)
        b = F.segment_reduce(plan.b_indptr, plan.b_eid, G, 128, mode=mode | _lib.SEG_SHORT)
        assert torch.equal(a, b)
    cp = _cpu_plan(plan)
    want = sc.seg_reduce(cp["a_indptr"], cp["a_eid"], G.cpu(), 128)
    assert torch.equal(F.segment_reduce(plan.a_indptr, plan.a_eid, G, 128, mode=_lib.SEG_SHORT).cpu(), want)


def test_strided_operands_use_leading_dimension():
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=5, n=40, e0=200, rev="halves")
    plan = _plan(s, d, 40, r)
    cp = _cpu_plan(plan)
    E, H = len(s), 32
    big = torch.randn(E, 3 * H, generator=torch.Generator().manual_seed(9))
    view = big[:, H:2 * H]  # ld = 3H, offset H: vector width must fall back correctly if misaligned
    want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], view.contiguous(), H, mode=_lib.SEG_SIGN_BY_REV)
    got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, big.cuda()[:, H:2 * H], H, mode=_lib.SEG_SIGN_BY_REV)
    assert torch.equal(got.cpu(), want)
    odd = torch.randn(E, H + 1)[:, 1:]  # rows start 4 bytes off 16-byte alignment -> scalar path
    want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], odd.contiguous(), H)
    got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, odd.cuda(), H)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("act,slope", [("none", 0.0), ("relu", 0.0), ("leaky_relu", 1 / 5.5)])
def test_gate_residual_bit_exact(act, slope):
    from dualmessagepassing_b200 import functional as F
    ids = {"none": 0, "relu": 1, "leaky_relu": 2}
    g = torch.Generator().manual_seed(4)
    x, prev, gout = torch.randn(333, 64, generator=g), torch.randn(333, 64, generator=g), torch.randn(333, 64, generator=g)
    gate = (torch.rand(333, generator=g) > 0.4).float()
    xg = x.cuda().requires_grad_(True)
    pg = prev.cuda().requires_grad_(True)
    out = F.gate_residual(xg, gate.cuda(), pg, act=act, slope=slope)
    assert torch.equal(out.cpu(), sc.gate_residual(x, gate, prev, ids[act], slope))
    out.backward(gout.cuda())
    assert torch.equal(xg.grad.cpu(), sc.gate_residual_backward(gout, x, gate, ids[act], slope))
    assert torch.equal(pg.grad.cpu(), gout)


def test_gate_residual_tanh_close():
    from dualmessagepassing_b200 import functional as F
    x = torch.randn(100, 50, generator=torch.Generator().manual_seed(5))
    out = F.gate_residual(x.cuda(), None, None, act="tanh")
    torch.testing.assert_close(out.cpu(), torch.tanh(x), rtol=1e-6, atol=1e-6)


def test_full_size_properties_config5_slice():
    """Size-independent properties at BASELINE scale (N=2M nodes; E reduced to fit test time is NOT done:
    this runs the real 40M-edge index build and one reduce): linearity of the segment sum and the
    checksum identity sum_x out[x] == sum_e sgn_e M[e] in fp64."""
    from dualmessagepassing_b200 import _lib, functional as F
    from dualmessagepassing_b200.plan import DMPPlan
    N, E0, H = 2_000_000, 20_000_000, 32
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randint(0, N, (E0,), device="cuda", generator=g)
    v = torch.randint(0, N, (E0,), device="cuda", generator=g)
    src, dst = torch.cat([u, v]), torch.cat([v, u])
    rev = torch.cat([torch.zeros(E0, dtype=torch.uint8, device="cuda"), torch.ones(E0, dtype=torch.uint8, device="cuda")])
    plan = DMPPlan(src, dst, N, rev=rev)
    assert plan.rev_layout == "halves"
    ip = plan.csc_indptr.long()
    assert int(ip[-1]) == 2 * E0 and bool((ip[1:] >= ip[:-1]).all())
    assert torch.equal(ip[1:] - ip[:-1], torch.bincount(dst, minlength=N))
    eid = plan.csc_eid.long() & 0x7FFFFFFF
    assert torch.equal(torch.sort(eid).values, torch.arange(2 * E0, device="cuda"))  # a permutation
    assert bool((dst[eid][1:] >= dst[eid][:-1]).all())                                # sorted by destination
    same = dst[eid][1:] == dst[eid][:-1]
    assert bool((eid[1:][same] > eid[:-1][same]).all())                               # stable inside a segment
    M = torch.randn(2 * E0, H, device="cuda", generator=g)
    out = F.segment_reduce(plan.csc_indptr, plan.csc_eid, M, H, mode=_lib.SEG_SIGN_BY_REV)
    sgn = rev.double() * 2 - 1
    want = (M.double() * sgn.unsqueeze(1)).sum(0)
    torch.testing.assert_close(out.double().sum(0), want, rtol=1e-6, atol=1e-2)
    out2 = F.segment_reduce(plan.csc_indptr, plan.csc_eid, M * 2, H, mode=_lib.SEG_SIGN_BY_REV)
    assert torch.equal(out2, out * 2)  # scaling by a power of two commutes with every rounding


@pytest.mark.parametrize("H", [32, 128])
def test_segment_reduce_flag_filters_bit_exact(H):
    """ONLY_FWD / ONLY_REV (aggregate-first dW_in / dW_out in backward) against the C oracle."""
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=99, n=300, e0=2500, rev="shuffled")
    plan = _plan(s, d, 300, r)
    cp = _cpu_plan(plan)
    E = len(s)
    g = torch.Generator().manual_seed(H)
    X, norm = torch.randn(E, H, generator=g), torch.rand(E, generator=g)
    w_cpu = norm[(cp["csc_eid"].long() & 0x7FFFFFFF)]
    w_gpu = plan.norm_permuted(norm.cuda())[1]
    for filt in (_lib.SEG_ONLY_FWD, _lib.SEG_ONLY_REV):
        mode = _lib.SEG_SIGN_BY_REV | filt
        want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], X, H, w_perm=w_cpu, mode=mode)
        got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, X.cuda(), H, w_perm=w_gpu, mode=mode)
        assert torch.equal(got.cpu(), want)
    both = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], X, H, w_perm=w_cpu, mode=_lib.SEG_SIGN_BY_REV)
    assert float(both.abs().sum()) > 0


@pytest.mark.parametrize("H", [8, 32, 64, 128, 200])
def test_segment_reduce_split_by_rev_bit_exact(H):
    """SPLIT_BY_REV (aggregate-first node update): one pass, [forward sums | reversed sums]; each half equals the
    filtered reduce and the C oracle bit for bit, also with strided rows."""
    from dualmessagepassing_b200 import _lib, functional as F
    s, d, r = make_graph(seed=7 + H, n=257, e0=3000, rev="shuffled")
    plan = _plan(s, d, 257, r)
    cp = _cpu_plan(plan)
    E = len(s)
    g = torch.Generator().manual_seed(H)
    Xw, norm = torch.randn(E, H + 8, generator=g), torch.rand(E, generator=g)
    X = Xw[:, :H]                                             # ldV = H + 8
    w_cpu = norm[(cp["csc_eid"].long() & 0x7FFFFFFF)]
    w_gpu = plan.norm_permuted(norm.cuda())[1]
    mode = _lib.SEG_SIGN_BY_REV | _lib.SEG_SPLIT_BY_REV
    want = sc.seg_reduce(cp["csc_indptr"], cp["csc_eid"], X, H, w_perm=w_cpu, mode=mode)
    got = F.segment_reduce(plan.csc_indptr, plan.csc_eid, Xw.cuda()[:, :H], H, w_perm=w_gpu, mode=mode)
    assert got.shape == (257, 2 * H) and torch.equal(got.cpu(), want)
    for filt, half in ((_lib.SEG_ONLY_FWD, got[:, :H]), (_lib.SEG_ONLY_REV, got[:, H:])):
        one = F.segment_reduce(plan.csc_indptr, plan.csc_eid, Xw.cuda()[:, :H], H, w_perm=w_gpu,
                               mode=_lib.SEG_SIGN_BY_REV | filt)
        assert torch.equal(one, half)
    with pytest.raises(RuntimeError):
        F.segment_reduce(plan.csc_indptr, plan.csc_eid, X.cuda(), H, mode=mode, bias=torch.zeros(H, device="cuda"))
