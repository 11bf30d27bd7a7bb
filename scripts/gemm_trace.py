"""Per-stage clock64 trace of CTA 0 of dmp_gemm_tf32x3 (debug bit 16): producer / MMA / epilogue timelines."""
import sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import _lib
E, N, K = 2_000_000, 128, 128
A = torch.randn(E, K, device="cuda"); Wt = torch.randn(N, K, device="cuda") / 4; D = torch.empty(E, N, device="cuda")
ts = torch.zeros(3 * 512, dtype=torch.int64, device="cuda")
lib = _lib.load()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(2):
    rc = lib.dmp_gemm_tf32x3(A.data_ptr(), K, None, Wt.data_ptr(), K, None, ts.data_ptr(), N, D.data_ptr(), N, E, N, K,
                             (flags << 8) | (1 << 16), 0.0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
t = ts.cpu().view(3, 256, 2)
t0 = int(t[0, 0, 0])
def show(name, arr, lo, hi):
    print(name)
    for i in range(lo, hi):
        print("  %3d  start %7d  end %7d  (busy %5d, since prev start %5d)" % (
            i, int(arr[i, 0]) - t0, int(arr[i, 1]) - t0, int(arr[i, 1] - arr[i, 0]), int(arr[i, 0] - arr[i - 1, 0]) if i else 0))
show("producer warp5 lane0: [after wait(empty) .. after arrive(full)] per stage", t[0], 100, 116)
show("MMA warp: [after wait(full) .. after issue+commit] per stage", t[1], 100, 116)
show("epilogue warp0: [after wait(acc_full) .. after arrive(acc_empty)] per tile", t[2], 25, 33)
