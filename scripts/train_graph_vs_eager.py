import sys, copy, numpy as np, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import train_step as ts
for cfgname in ("cfg1", "cfg2"):
    cfg = ts.CONFIGS[cfgname]
    ds = ts.SyntheticPairDataset(cfgname, num=4 * cfg["pairs"], seed=2000)
    dds = ts.DevicePairDataset(ds, "cuda")
    torch.manual_seed(2000)
    model = ts.SubgraphCountingModel(cfg["hidden"], cfg["labels"][0], cfg["labels"][1]).cuda()
    ref = copy.deepcopy(model)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, amsgrad=True, capturable=True)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, amsgrad=True, capturable=True)
    rng = np.random.Generator(np.random.PCG64(7))
    batches = [np.sort(rng.choice(ds.num, size=cfg["pairs"], replace=False)) for _ in range(40)]
    graphed = ts.GraphedTrainStep(model, opt, dds, cfg["pairs"], warmup=3)
    got = [float(graphed(b).detach()) for b in batches]
    for b in graphed.warmup_ids:
        p, g, y, _ = ts.collate_on_device(dds, b); ts.train_step(ref, opt_ref, p, g, y)
    want = []
    for b in batches:
        p, g, y, _ = ts.collate_on_device(dds, b); want.append(float(ts.train_step(ref, opt_ref, p, g, y).detach()))
    print(cfgname, "graphed", [round(x, 2) for x in got[::4]])
    print(cfgname, "eager  ", [round(x, 2) for x in want[::4]])
    print(cfgname, "max rel diff", max(abs(a - b) / max(abs(b), 1e-6) for a, b in zip(got, want)))
