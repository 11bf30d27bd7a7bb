"""Host-side phase breakdown of the e2e training step (cfg2)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import train_step as ts, _lib
cfgname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = ts.CONFIGS[cfgname]
dev = torch.device("cuda")
ds = ts.SyntheticPairDataset(cfgname, num=4 * cfg["pairs"], seed=2000)
torch.manual_seed(0)
model = ts.SubgraphCountingModel(cfg["hidden"], cfg["labels"][0], cfg["labels"][1]).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True)
rng = np.random.Generator(np.random.PCG64(7))
T = {k: 0.0 for k in ("collate", "h2d", "fwd", "bwd", "clip+opt")}
def sync(): torch.cuda.synchronize()
for it in range(25):
    t0 = time.perf_counter()
    idx = np.sort(rng.choice(ds.num, size=cfg["pairs"], replace=False)); b = ts.collate(ds, idx)
    t1 = time.perf_counter()
    p, g, y, nb = ts.to_device(b, dev); sync()
    t2 = time.perf_counter()
    opt.zero_grad(set_to_none=True); pred = model(p, g, union=ts.union_graph(p, g)); loss = torch.mean((pred - y) ** 2); sync()
    t3 = time.perf_counter()
    loss.backward(); sync()
    t4 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0, foreach=True); opt.step(); sync()
    t5 = time.perf_counter()
    if it >= 5:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): T[k] += v / 20
print({k: round(v * 1e3, 2) for k, v in T.items()}, "total ms", round(sum(T.values()) * 1e3, 2))
l0 = _lib.LAUNCHES
pred = model(p, g, union=ts.union_graph(p, g)); torch.mean((pred - y) ** 2).backward()
print("dmp launches per fwd+bwd", _lib.LAUNCHES - l0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    opt.zero_grad(); pred = model(p, g, union=ts.union_graph(p, g)); torch.mean((pred - y) ** 2).backward(); sync()
ev = prof.key_averages()
cuda_ms = sum(e.device_time_total for e in ev) / 1e3
print("device busy ms in fwd+bwd", round(cuda_ms, 2), "kernels", sum(e.count for e in ev if e.device_time_total > 0))
rows = sorted(ev, key=lambda r: -r.device_time_total)[:12]
for r in rows:
    print("%8.3f ms x%-4d %s" % (r.device_time_total / 1e3, r.count, r.key[:90]))
cpu = sorted(ev, key=lambda r: -r.self_cpu_time_total)[:14]
print("-- host self time")
for r in cpu:
    print("%8.3f ms x%-4d %s" % (r.self_cpu_time_total / 1e3, r.count, r.key[:90]))
