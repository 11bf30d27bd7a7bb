"""Ablation timing of dmp_gemm_tf32x3 (debug bits 8..11 of `epilogue`): which role bounds the tile time?"""
import sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import _lib
E, N, K = 8_000_000, 128, 128
A = torch.randn(E, K, device="cuda"); Wt = torch.randn(N, K, device="cuda") / 4; D = torch.empty(E, N, device="cuda")
lib = _lib.load()
def run(flags):
    return lib.dmp_gemm_tf32x3(A.data_ptr(), K, None, Wt.data_ptr(), K, None, None, 0, D.data_ptr(), N, E, N, K, flags << 8, 0.0,
                               torch.cuda.current_stream().cuda_stream)
def t(flags, n=5):
    run(flags); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): run(flags)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
names = {15 + 128: "barriers, plain arrive instead of commit", 15: "barriers only", 15 + 16: "barriers, no fence", 15 + 32: "barriers, spin-wait", 15 + 64: "barriers, no tmem ld",
         15 + 16 + 32 + 64: "barriers, none of the three", 32: "full, spin-wait"}
names0 = {0: "full", 1: "no global loads", 2: "no smem stores", 4: "no MMA", 8: "no global stores", 3: "no ldg+sts", 12: "no mma+stg",
         7: "no ldg+sts+mma", 15: "barriers only", 9: "no ldg+stg", 6: "no sts+mma", 14: "only ldg"}
for f, nm in names.items():
    print("%-20s %.3f ms" % (nm, t(f)))
