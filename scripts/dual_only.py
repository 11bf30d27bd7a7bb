"""One full-size (config 5) launch of the dual store / dual accumulate projection kernels -- target of `ncu -k regex:v3_kernel`."""
import sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000), 128
A = torch.randn(E, H, device="cuda")
D = torch.empty(E, H, device="cuda")
W1, W2 = torch.randn(H, H, device="cuda") / 8, torch.randn(H, H, device="cuda") / 8
c = torch.rand(E, device="cuda") * 6
F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="store", out=D)
F.gemm_tf32x3_dual(A, W1, W2, row_scale=c, mode="accumulate", out=D)
torch.cuda.synchronize()
