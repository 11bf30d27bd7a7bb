"""One launch of every hand-written hot kernel at config-5 size (E = 40 M edges, N = 2 M nodes, H = 128), in a fixed
order, for `ncu --set full` (scripts/ncu_capture.sh).  Order = the TAGS list below; scripts/summarise_ncu.py relies on it."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from dualmessagepassing_b200 import _lib, functional as F
from dualmessagepassing_b200.plan import DMPPlan

TAGS = ["segment_reduce.node_fwd", "segment_reduce.dQd_bwd", "edge_update", "edge_backward", "gemm_tf32x3.store",
        "gemm_tf32x3.bias_act", "gemm_tf32x3.grad", "gemm_tf32x3_dual.store", "gemm_tf32x3_dual.accumulate",
        "gemm_tn_tf32x3"]

if __name__ == "__main__":
    n, e0, h, _ = bench.WORKLOADS["cfg5"]
    src, dst, rev = bench.make_graph(n, e0, 5000)
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a).to(dev)
    plan = DMPPlan(t(src), t(dst), n, rev=t(rev.astype(np.uint8)), rev_layout="halves", rev_split=e0)
    E = 2 * e0
    A, G = torch.randn(E, h, device=dev), torch.randn(E, h, device=dev)
    D = torch.empty(E, h, device=dev)
    W = torch.randn(h, h, device=dev) / 8
    W2 = torch.randn(h, h, device=dev) / 8
    bias = torch.randn(h, device=dev)
    Q = torch.randn(n, h, device=dev)
    torch.cuda.synchronize()
    F.segment_reduce(plan.csc_indptr, plan.csc_eid, A, h, mode=_lib.SEG_SIGN_BY_REV | _lib.SEG_SPLIT_BY_REV)
    F.segment_reduce(plan.a_indptr, plan.a_eid, G, h)
    F.edge_update(plan, A, None, Q, Q, bias, _lib.ORDER_SCM, out=D)
    F.edge_backward(plan, None, Q, None, want_CG=False, T=D, gN_rev=Q)
    F.gemm_tf32x3(A, W, out=D)
    F.gemm_tf32x3(A, W, bias=bias, act="leaky_relu", slope=0.18, out=D)
    F.gemm_tf32x3(A, W, act="leaky_relu", slope=0.18, aux=G, mul_act_grad=True, out=D)
    F.gemm_tf32x3_dual(A, W, W2, row_scale=plan.coef, mode="store", out=D)
    F.gemm_tf32x3_dual(A, W, W2, row_scale=plan.coef, mode="accumulate", out=D)
    F.gemm_tn_tf32x3(A, G)
    torch.cuda.synchronize()
    print("done")
