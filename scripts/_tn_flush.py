import sys, os, json, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E, H = 40_000_000, 128
dev = torch.device("cuda")
X, G = torch.randn(E, H, device=dev), torch.randn(E, H, device=dev)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timed(lambda: F.gemm_tn_tf32x3(X, G))
n = 4_000_000
D = F.gemm_tn_tf32x3(X[:n], G[:n])
ref = X[:n].double().t() @ G[:n].double()
cub = (X[:n].t() @ G[:n])
print(os.environ.get("DMP_B200_LIB", "default")[-16:], "tn %.2f ms" % ms, "err %.2e" % float((D.double() - ref).abs().max() / ref.abs().max()),
      "cublas %.2e" % float((cub.double() - ref).abs().max() / ref.abs().max()))
