"""One full-size (config 5) launch set of the K = E reduction kernel -- the target of `ncu -k regex:tn_kernel`."""
import sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
X, G = torch.randn(E, 128, device="cuda"), torch.randn(E, 128, device="cuda")
for _ in range(2):
    F.gemm_tn_tf32x3(X, G)
torch.cuda.synchronize()
