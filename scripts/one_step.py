"""Two DMPLayer fwd+bwd steps at a bench workload, nothing else (target of the ncu captures in scripts/ncu_capture.sh)."""
import sys, torch
sys.path.insert(0, ".")
import bench
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200.constants import REVFLAG
n, e0, h, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
src, dst, rev = bench.make_graph(n, e0, 5000)
dev = torch.device("cuda")
torch.manual_seed(0)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu").to(dev)
g = dmp.DMPGraph(torch.from_numpy(src), torch.from_numpy(dst), n).to(dev)
g.edata[REVFLAG] = torch.from_numpy(rev).to(dev).bool(); g.rev_layout_hint = "halves"
E = 2 * e0
xv, xe = torch.randn(n, h, device=dev), torch.randn(E, h, device=dev)
gv, ge = torch.randn(n, h, device=dev), torch.randn(E, h, device=dev)
for _ in range(steps):
    layer.zero_grad(set_to_none=True)
    a, b = xv.requires_grad_(True), xe.requires_grad_(True)
    nv, ne = layer(g, a, b)
    torch.autograd.backward((nv, ne), (gv, ge))
    a.grad = None; b.grad = None
    del nv, ne
torch.cuda.synchronize()
print("done")
