"""Projection kernels at config-5 size (E = 40 M rows, 128 x 128): CUDA-event time per launch, fraction of the measured HBM
peak, error vs fp64 on a sample.  (profiles/r2_gemm_sweep.txt was taken with a build whose hi/lo ring split was a launch
parameter, DMP_V3_LO; the split is compile-time again: the run-time modulo slowed the MMA-issuing thread.)"""
import json, os, sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F

E, H = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000, 128
dev = torch.device("cuda")
A, G = torch.randn(E, H, device=dev), torch.randn(E, H, device=dev)
D = torch.empty(E, H, device=dev)
W1, W2 = torch.randn(H, H, device=dev) / 8, torch.randn(H, H, device=dev) / 8
c = torch.rand(E, device=dev) * 6
bias = torch.randn(H, device=dev)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
row = 4 * H * E


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = {
    "store": (lambda: F.gemm_tf32x3(A, W1, out=D), 2 * row),
    "bias_act": (lambda: F.gemm_tf32x3(A, W1, bias=bias, act="leaky_relu", slope=0.18, out=D), 2 * row),
    "grad": (lambda: F.gemm_tf32x3(A, W1, act="leaky_relu", slope=0.18, aux=G, mul_act_grad=True, out=D), 3 * row),
    "dual.store": (lambda: F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store", out=D), 2 * row),
    "dual.accumulate": (lambda: F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="accumulate", out=D), 3 * row),
    "tn": (lambda: F.gemm_tn_tf32x3(A, G), 2 * row),
}
out = {"rows": E, "DMP_V3_LO": os.environ.get("DMP_V3_LO", "default")}
for name, (fn, nbytes) in cases.items():
    ms = timed(fn)
    out[name] = {"ms": round(ms, 3), "frac": round(nbytes / ms / 1e6 / peak, 3)}
s = slice(0, 200_000)
ref = A[s].double() @ W1.double().t()
out["err_vs_fp64"] = float((F.gemm_tf32x3(A[s].contiguous(), W1).double() - ref).abs().max() / ref.abs().max())
out["err_cublas"] = float(((A[s] @ W1.t()).double() - ref).abs().max() / ref.abs().max())
print(json.dumps(out))
