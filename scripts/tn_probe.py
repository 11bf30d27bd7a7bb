"""Probe the MN-major operand layout of the TN kernel with one-hot inputs (E = 32, a single stage)."""
import os, sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E, M, N = 32, 128, 128
def probe(e0, m0, n0):
    X = torch.zeros(E, M, device="cuda"); G = torch.zeros(E, N, device="cuda")
    X[e0, m0] = 1.0; G[e0, n0] = 1.0
    D = F.gemm_tn_tf32x3(X, G)
    nz = torch.nonzero(D).tolist()
    return nz[:6], float(D.sum())
for (e0, m0, n0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 4, 0), (0, 32, 0), (0, 0, 32), (1, 0, 0), (1, 3, 5), (8, 0, 0), (9, 33, 70), (31, 127, 127)]:
    print(os.environ.get("DMP_TN_DBG", "default"), (e0, m0, n0), "->", probe(e0, m0, n0))
# random check
X = torch.randn(E, M, device="cuda"); G = torch.randn(E, N, device="cuda")
ref = X.double().t() @ G.double()
print("random err", float((F.gemm_tn_tf32x3(X, G).double() - ref).abs().max() / ref.abs().max()))
