"""Probe: does cuBLAS 12.9 BF16x9 fp32 emulation (CUBLAS_EMULATE_SINGLE_PRECISION=1) speed up the
edge-sized [E,128]x[128,128] sgemm, and what is its error vs fp64?  Run twice: plain, and with
LD_PRELOAD of the system cuBLAS 12.9 (torch bundles 12.8)."""
import os, sys, time, torch
E, H = 8_000_000, 128
torch.manual_seed(0)
A = torch.randn(E, H, device="cuda"); W = torch.randn(H, H, device="cuda") / 4
G = torch.randn(E, H, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("cublas version", torch.backends.cuda.cublas_version() if hasattr(torch.backends.cuda, "cublas_version") else "?",
      "EMULATE", os.environ.get("CUBLAS_EMULATE_SINGLE_PRECISION"), "PRELOAD", bool(os.environ.get("LD_PRELOAD")))
ms_row = t(lambda: A @ W); ms_red = t(lambda: A.t() @ G)
fl = 2.0 * E * H * H
print("row gemm  [E,128]x[128,128]: %.2f ms  %.1f TFLOP/s" % (ms_row, fl / ms_row / 1e9))
print("red gemm  [128,E]x[E,128]  : %.2f ms  %.1f TFLOP/s" % (ms_red, fl / ms_red / 1e9))
sub = slice(0, 200000)
ref = (A[sub].double() @ W.double())
err = ((A[sub] @ W).double() - ref).abs().max().item() / ref.abs().max().item()
print("row gemm max-norm rel err vs fp64: %.3g" % err)
for mode in ("tf32",):
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = t(lambda: A @ W)
    err = ((A[sub] @ W).double() - ref).abs().max().item() / ref.abs().max().item()
    print("allow_tf32 row gemm: %.2f ms %.1f TFLOP/s err %.3g" % (ms, fl / ms / 1e9, err))
    torch.backends.cuda.matmul.allow_tf32 = False
