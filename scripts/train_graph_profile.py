"""Kernel table of ONE CUDA-graph replay of the training step (torch.profiler / CUPTI sees the kernels inside a replay).
    python scripts/train_graph_profile.py [cfg2]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import train_step as ts
from torch.profiler import profile, ProfilerActivity
cfgname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = ts.CONFIGS[cfgname]
ds = ts.SyntheticPairDataset(cfgname, num=4 * cfg["pairs"], seed=2000)
dds = ts.DevicePairDataset(ds, "cuda")
torch.manual_seed(2000)
model = ts.SubgraphCountingModel(cfg["hidden"], cfg["labels"][0], cfg["labels"][1]).cuda()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True, capturable=True, fused=True)
step = ts.GraphedTrainStep(model, opt, dds, cfg["pairs"])
rng = np.random.Generator(np.random.PCG64(7))
idx = [np.sort(rng.choice(ds.num, size=cfg["pairs"], replace=False)) for _ in range(8)]
for i in idx[:4]:
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in idx:
    step(i)
e1.record(); torch.cuda.synchronize()
print("ms per replay", e0.elapsed_time(e1) / len(idx), "padded union", step.pad)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(idx[0]); torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print("kernels in one replay: %d, summed device time %.3f ms" % (sum(e.count for e in ev), tot / 1e3))
for e in sorted(ev, key=lambda r: -r.device_time_total)[:int(sys.argv[2]) if len(sys.argv) > 2 else 32]:
    print("%8.1f us x%-4d %5.1f%%  %s" % (e.device_time_total, e.count, 100 * e.device_time_total / tot, e.key[:110]))
