import os, sys, torch
sys.path.insert(0, ".")
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200 import fused
from dualmessagepassing_b200.constants import REVFLAG
from tests._cases import make_graph
n, e0, h = 1500, 6000, 128
s, d, r = make_graph(seed=n, n=n, e0=e0, rev="shuffled", isolated=3)
E = len(s)
torch.manual_seed(n)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu").cuda()
xv, xe, gv, ge = torch.randn(n, h).cuda(), torch.randn(E, h).cuda(), torch.randn(n, h).cuda(), torch.randn(E, h).cuda()
orig = fused._rowmm
log = {}
def rec(A, Wt, **k):
    ins = (A.detach().clone(), Wt.detach().clone(), None if k.get("aux") is None else k["aux"].detach().clone(),
           None if not k.get("accumulate") else k["out"].detach().clone())
    res = orig(A, Wt, **k)
    log[fused.DENSE_BACKEND].append((ins, res.detach().clone(), {kk: (vv if not torch.is_tensor(vv) else "T") for kk, vv in k.items()}))
    return res
fused._rowmm = rec
for backend in ("cublas", "auto"):
    fused.DENSE_BACKEND = backend
    log[backend] = []
    layer.zero_grad()
    g = dmp.DMPGraph(s, d, n, device="cuda"); g.edata[REVFLAG] = torch.from_numpy(r).cuda()
    a, b = xv.clone().requires_grad_(True), xe.clone().requires_grad_(True)
    nv, ne = layer(g, a, b)
    torch.autograd.backward((nv, ne), (gv, ge))
def rel(x, y):
    if x is None: return 0.0
    return float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
for i, (c, t) in enumerate(zip(log["cublas"], log["auto"])):
    (ci, co, ck), (ti, to, tk) = c, t
    print("call %2d out %-9.2g A %-9.2g W %-9.2g aux %-9.2g acc_in %-9.2g  %s" % (
        i + 1, rel(to, co), rel(ti[0], ci[0]), rel(ti[1], ci[1]), rel(ti[2], ci[2]) if ci[2] is not None else 0,
        rel(ti[3], ci[3]) if ci[3] is not None else 0, ck))
