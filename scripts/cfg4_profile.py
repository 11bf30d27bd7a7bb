"""Kernel table of one fwd+bwd of the UNC encoder body (BASELINE configs[3]; bench.run_cfg4's workload).
    python scripts/cfg4_profile.py [rows]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dualmessagepassing_b200 as dmp
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
n, nt, R, h = 20_000, 90_000, 10, 50
rng = np.random.Generator(np.random.PCG64(4000))
trip = np.stack([rng.integers(0, n, nt), rng.integers(0, R, nt), rng.integers(0, n, nt)], 1)
g = dmp.build_graph_from_triplets(n, R, trip).to(dev)
E = g.number_of_edges()
torch.manual_seed(4000)
layers = [dmp.DualGraphConv(h, h, activation=torch.nn.Tanh()).to(dev).train(),
          dmp.DualGraphConv(h, h, activation=None).to(dev).train()]
gen = torch.Generator(device=dev).manual_seed(4000)
h0, z0 = torch.randn(n, h, device=dev, generator=gen), torch.randn(E, h, device=dev, generator=gen)
gh, gz = torch.randn(n, h, device=dev, generator=gen), torch.randn(E, h, device=dev, generator=gen)
gr = torch.randn(2 * R, h, device=dev, generator=gen)
params = [p for L in layers for p in L.parameters()]


def step():
    for p in params:
        p.grad = None
    a, b = h0.requires_grad_(True), z0.requires_grad_(True)
    x, y = a, b
    for L in layers:
        x, y = L(g, x, y, g.edata["norm"])
    pooled = dmp.relation_mean_pool(y, g.edata["type"], 2 * R)
    torch.autograd.backward((x, y, pooled), (gh, gz, gr))
    a.grad = b.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print("kernels in one step: %d, summed device time %.3f ms" % (sum(e.count for e in ev), tot / 1e3))
for e in sorted(ev, key=lambda r: -r.device_time_total)[:int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
    print("%8.1f us x%-4d %5.1f%%  %s" % (e.device_time_total, e.count, 100 * e.device_time_total / tot, e.key[:110]))
