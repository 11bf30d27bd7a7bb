"""Micro-benchmark: dmp_gemm_tf32x3 vs cuBLAS sgemm on the edge-sized projection shape."""
import sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
for (N, K) in [(128, 128), (64, 64)]:
    A = torch.randn(E, K, device="cuda"); Wt = torch.randn(N, K, device="cuda") / 4
    out = torch.empty(E, N, device="cuda")
    def t(fn, n=10):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    G = torch.randn(E, N, device="cuda")
    ms_rc = t(lambda: A.t() @ G); ms_ro = t(lambda: F.gemm_tn_tf32x3(A, G))
    refr = A.double().t() @ G.double()
    er = lambda x: float((x.double() - refr).abs().max() / refr.abs().max())
    print("   reduction X^T G: cuBLAS %.3f ms (err %.2g) | tf32x3 %.3f ms (%.0f GB/s, err %.2g)" % (
        ms_rc, er(A.t() @ G), ms_ro, (E * K + E * N) * 4 / 1e6 / ms_ro, er(F.gemm_tn_tf32x3(A, G))))
    ms_c = t(lambda: torch.mm(A, Wt.t(), out=out))
    ms_o = t(lambda: F.gemm_tf32x3(A, Wt, out=out))
    gb = (E * K + E * N) * 4 / 1e9
    ref = A[:100000].double() @ Wt.double().t()
    err = lambda x: float((x.double() - ref).abs().max() / ref.abs().max())
    print("N=%d K=%d E=%d  cuBLAS sgemm %.3f ms (%.0f GB/s, %.1f TF/s, err %.2g) | tf32x3 %.3f ms (%.0f GB/s, %.1f eff TF/s, err %.2g)"
          % (N, K, E, ms_c, gb / ms_c * 1e3, 2.0 * E * N * K / ms_c / 1e9, err(torch.mm(A[:100000], Wt.t())),
             ms_o, gb / ms_o * 1e3, 2.0 * E * N * K / ms_o / 1e9, err(F.gemm_tf32x3(A[:100000], Wt))))
