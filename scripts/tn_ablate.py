"""Ablation timing of the weight-gradient kernel (results are WRONG under ablation; timing only)."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from dualmessagepassing_b200 import functional as F
    E = 40_000_000
    for H in (128, 64):
        A = torch.randn(E, H, device="cuda"); G = torch.randn(E, H, device="cuda")
        F.gemm_tn_tf32x3(A, G); F.gemm_tn_tf32x3(A, G); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): F.gemm_tn_tf32x3(A, G)
        e1.record(); torch.cuda.synchronize()
        print("ablate=%s H=%d %.3f ms" % (os.environ.get("DMP_TN_ABLATE", "0"), H, e0.elapsed_time(e1) / 5), flush=True)
        del A, G
else:
    for ab in sys.argv[1:] or ["0", "1", "2", "4", "8", "16", "3", "12", "31"]:
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DMP_TN_ABLATE=ab))
