import os, sys, torch
sys.path.insert(0, ".")
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200 import fused
from dualmessagepassing_b200.constants import REVFLAG
from tests._cases import make_graph
mode = sys.argv[1]
n, e0, h = 1500, 6000, 128
s, d, r = make_graph(seed=n, n=n, e0=e0, rev="shuffled", isolated=3)
E = len(s)
torch.manual_seed(n)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu").cuda()
xv, xe, gv, ge = torch.randn(n, h).cuda(), torch.randn(E, h).cuda(), torch.randn(n, h).cuda(), torch.randn(E, h).cuda()
if mode == "sync":
    orig = fused._rowmm
    def synced(*a, **k):
        torch.cuda.synchronize(); r_ = orig(*a, **k); torch.cuda.synchronize(); return r_
    fused._rowmm = synced
res = {}
for backend in ("cublas", "auto", "auto", "auto"):
    fused.DENSE_BACKEND = backend
    layer.zero_grad()
    g = dmp.DMPGraph(s, d, n, device="cuda"); g.edata[REVFLAG] = torch.from_numpy(r).cuda()
    a, b = xv.clone().requires_grad_(True), xe.clone().requires_grad_(True)
    nv, ne = layer(g, a, b)
    torch.autograd.backward((nv, ne), (gv, ge))
    cur = {"dXv": a.grad.clone(), "dXe": b.grad.clone(), "nmlp0": layer.nmlp[0].weight.grad.clone(), "nv": nv.detach().clone()}
    if backend == "cublas": ref = cur
    else:
        print(mode, {k: "%.2g" % float((cur[k] - ref[k]).abs().max() / ref[k].abs().max()) for k in cur})
