"""Segment reduce on a skewed (power-law) graph of Yelp's size (UNC/Data/README.md: 82 465 nodes, 30.5 M links): the
strictly sequential default against the chunked two-level option (functional.segment_reduce_two_level).
    python scripts/skew_bench.py  -> one JSON line"""
import json, sys, numpy as np, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import _lib, functional as F
from dualmessagepassing_b200.plan import DMPPlan

n, e, H = 82_465, 30_000_000, 128
rng = np.random.Generator(np.random.PCG64(9))
w = (1.0 / np.arange(1, n + 1) ** 0.9)
w /= w.sum()
dst = rng.choice(n, size=e, p=w).astype(np.int64)          # Zipf-like in-degree: the top node gets ~7 % of all edges
src = rng.integers(0, n, size=e, dtype=np.int64)
dev = torch.device("cuda")
plan = DMPPlan(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), n, validate=False)
lens = (plan.csc_indptr[1:] - plan.csc_indptr[:-1]).long()
V = torch.randn(e, H, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
alg = 4 * H * (e + n) + 4 * e + 4 * (n + 1)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


seq = timed(lambda: F.segment_reduce(plan.csc_indptr, plan.csc_eid, V, H), 2)
out = {"graph": "power-law in-degree, %d nodes / %d edges, H=%d" % (n, e, H), "max_in_degree": int(lens.max()),
       "median_in_degree": int(lens.median()), "alg_bytes": alg,
       "sequential": {"ms": seq, "gbs": alg / seq / 1e6, "frac": alg / seq / 1e6 / peak}}
for chunk in (256, 1024, 4096):
    ms = timed(lambda: F.segment_reduce_two_level(plan.csc_indptr, plan.csc_eid, V, H, chunk=chunk), 5)
    out["two_level_%d" % chunk] = {"ms": ms, "gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak}
a = F.segment_reduce(plan.csc_indptr, plan.csc_eid, V, H)
b = F.segment_reduce_two_level(plan.csc_indptr, plan.csc_eid, V, H, chunk=1024)
out["max_rel_diff_two_level_vs_sequential"] = float((a - b).abs().max() / a.abs().max())
print(json.dumps(out))
