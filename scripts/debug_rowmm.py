import sys, torch, numpy as np
sys.path.insert(0, ".")
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200 import fused, _lib
from dualmessagepassing_b200.constants import REVFLAG
from tests._cases import make_graph
n, e0, h, rev = 1500, 6000, 128, sys.argv[1] if len(sys.argv) > 1 else "shuffled"
s, d, r = make_graph(seed=n, n=n, e0=e0, rev=rev, isolated=3)
E = len(s)
torch.manual_seed(n)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu").cuda()
xv, xe, gv, ge = torch.randn(n, h), torch.randn(E, h), torch.randn(n, h), torch.randn(E, h)
orig = fused._rowmm
count = [0]
def checked(A, Wt, *, bias=None, act=_lib.ACT_NONE, slope=0.0, aux=None, mul_act_grad=False, accumulate=False, out=None):
    count[0] += 1
    A64, W64 = A.double(), Wt.double()
    ref = A64 @ W64.t()
    if bias is not None: ref = ref + bias.double()
    if mul_act_grad:
        y = aux.double()
        ref = ref * torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    elif act == _lib.ACT_LEAKY_RELU:
        ref = torch.where(ref > 0, ref, ref * slope)
    if accumulate: ref = ref + out.double()
    res = orig(A, Wt, bias=bias, act=act, slope=slope, aux=aux, mul_act_grad=mul_act_grad, accumulate=accumulate, out=out)
    err = float((res.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    flag = "  <<<<<<" if err > 1e-5 else ""
    print("call %2d A%s lda=%d ptr%%512=%d Wt%s contig=%s acc=%d mulgrad=%d bias=%d out=%s err=%.3g%s" % (
        count[0], tuple(A.shape), A.stride(0), A.data_ptr() % 512, tuple(Wt.shape), Wt.is_contiguous(), accumulate, mul_act_grad,
        bias is not None, None if out is None else (tuple(out.shape), out.stride(0)), err, flag))
    return res
fused._rowmm = checked
g = dmp.DMPGraph(s, d, n, device="cuda")
g.edata[REVFLAG] = torch.from_numpy(r).cuda()
a, b = xv.cuda().requires_grad_(True), xe.cuda().requires_grad_(True)
nv, ne = layer(g, a, b)
((nv * gv.cuda()).sum() + (ne * ge.cuda()).sum()).backward()
