import sys, torch, numpy as np
sys.path.insert(0, ".")
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200 import fused
from dualmessagepassing_b200.constants import REVFLAG
from tests._cases import make_graph
from tests.test_gpu_layer import _oracle_run
n, e0, h, rev = 1500, 6000, 128, sys.argv[1] if len(sys.argv) > 1 else "shuffled"
s, d, r = make_graph(seed=n, n=n, e0=e0, rev=rev, isolated=3)
E = len(s)
torch.manual_seed(n)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu")
sd = {k: v.clone() for k, v in layer.state_dict().items()}
xv, xe, gv, ge = torch.randn(n, h), torch.randn(E, h), torch.randn(n, h), torch.randn(E, h)
ref64 = _oracle_run(sd, s, d, n, r, xv, xe, gv, ge, torch.float64, flavour="scm", act_func="leaky_relu")
layer.cuda().train()
for backend in ("cublas", "auto"):
    fused.DENSE_BACKEND = backend
    layer.zero_grad()
    g = dmp.DMPGraph(s, d, n, device="cuda")
    g.edata[REVFLAG] = torch.from_numpy(r).cuda()
    a, b = xv.cuda().requires_grad_(True), xe.cuda().requires_grad_(True)
    nv, ne = layer(g, a, b)
    ((nv * gv.cuda()).sum() + (ne * ge.cuda()).sum()).backward()
    ours = {"node_out": nv, "edge_out": ne, "grad_node_feat": a.grad, "grad_edge_feat": b.grad}
    ours.update({"grad " + k: p.grad for k, p in layer.named_parameters() if p.grad is not None})
    print("backend", backend)
    for k, v64 in ref64.items():
        got = ours[k].detach().cpu().double()
        print("   %-24s %.3g" % (k, float((got - v64).abs().max()) / float(v64.abs().max())))
