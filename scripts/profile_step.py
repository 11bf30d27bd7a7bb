"""Kernel-time breakdown of one DMPLayer fwd+bwd step at config 5 (torch.profiler, CUDA activities)."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
import dualmessagepassing_b200 as dmp
from dualmessagepassing_b200.constants import REVFLAG
n, e0, h, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg5"]
src, dst, rev = bench.make_graph(n, e0, 5000)
dev = torch.device("cuda")
torch.manual_seed(0)
layer = dmp.DMPLayer(h, h, num_mlp_layers=2, batch_norm=False, act_func="leaky_relu").to(dev)
g = dmp.DMPGraph(torch.from_numpy(src), torch.from_numpy(dst), n).to(dev)
g.edata[REVFLAG] = torch.from_numpy(rev).to(dev).bool(); g.rev_layout_hint = "halves"
E = 2 * e0
xv, xe = torch.randn(n, h, device=dev), torch.randn(E, h, device=dev)
gv, ge = torch.randn(n, h, device=dev), torch.randn(E, h, device=dev)
def step():
    layer.zero_grad(set_to_none=True)
    a, b = xv.requires_grad_(True), xe.requires_grad_(True)
    nv, ne = layer(g, a, b)
    torch.autograd.backward((nv, ne), (gv, ge))
    a.grad = None; b.grad = None
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print("total device ms %.1f" % (tot / 1e3))
for r in rows[:22]:
    print("%8.2f ms %5.1f%% x%-3d %s" % (r.device_time_total / 1e3, 100 * r.device_time_total / tot, r.count, r.key[:110]))
