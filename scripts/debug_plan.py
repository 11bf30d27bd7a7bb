import numpy as np, torch, sys
sys.path.insert(0, ".")
from oracle import graph_oracle as go
from tests._cases import make_graph, t
from dualmessagepassing_b200.plan import DMPPlan
for rev in ["halves", None]:
    for (n, e0) in [(30, 100), (64, 300)]:
        s, d, r = make_graph(seed=n + 4, n=n, e0=e0, rev=rev)
        plan = DMPPlan(t(s), t(d), n, rev=t(None if r is None else r.astype(np.uint8)))
        want = go.build_plan(s, d, n, r)
        torch.cuda.synchronize()
        for k in ("dst32", "a32", "b32", "csc_indptr", "a_indptr", "b_indptr", "out_deg", "csc_eid", "a_eid", "b_eid"):
            got = getattr(plan, k).cpu().numpy()
            w = want[k]
            if k.endswith("eid"):
                got = got.view(np.uint32) & 0x7FFFFFFF
                w = w.view(np.uint32)
            ok = np.array_equal(got, w)
            print(rev, n, e0, k, "OK" if ok else "MISMATCH got[:8]=%s want[:8]=%s" % (got[:8], w[:8]))
