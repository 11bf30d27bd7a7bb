"""a-/b-keyed backward reductions of ONE RANK of an 8-way destination-range partition of config 5: 2 M global segments,
5 M local edge rows (~2.5 rows per segment), H = 128.  Times dmp_segment_reduce with and without DMP_SEG_SHORT and checks
that the two agree bit for bit.   python scripts/short_seg_bench.py -> one JSON line"""
import json, sys, numpy as np, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import _lib, functional as F

N, E, H = 2_000_000, 5_000_000, 128
rng = np.random.Generator(np.random.PCG64(3))
key = np.sort(rng.integers(0, N, size=E, dtype=np.int64))
indptr = np.zeros(N + 1, dtype=np.int32)
np.cumsum(np.bincount(key, minlength=N), out=indptr[1:])
eid = rng.permutation(E).astype(np.int32)
eid[::2] |= np.int32(-2**31)                      # reversed flag on half of the entries (ignored without SIGN_BY_REV)
dev = torch.device("cuda")
ip, ei = torch.from_numpy(indptr).to(dev), torch.from_numpy(eid).to(dev)
V = torch.randn(E, H, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
alg = 4 * H * (E + N) + 4 * E + 4 * (N + 1)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"segments": N, "rows": E, "alg_bytes": alg}
for name, mode in (("long_variant", 0), ("short_variant", _lib.SEG_SHORT)):
    ms = timed(lambda: F.segment_reduce(ip, ei, V, H, mode=mode | _lib.SEG_NEGATE_OUT))
    out[name] = {"ms": round(ms, 4), "frac": round(alg / ms / 1e6 / peak, 3)}
a = F.segment_reduce(ip, ei, V, H, mode=_lib.SEG_SIGN_BY_REV)
b = F.segment_reduce(ip, ei, V, H, mode=_lib.SEG_SIGN_BY_REV | _lib.SEG_SHORT)
out["bit_identical"] = bool(torch.equal(a, b))
print(json.dumps(out))
