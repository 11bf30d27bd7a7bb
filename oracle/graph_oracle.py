"""CPU oracle (numpy) for the integer graph structures the DMPNN layer consumes.

TEST INFRASTRUCTURE ONLY -- never imported by `dualmessagepassing_b200`.

Restates, with numpy primitives, the index semantics of:
  * `dgl.batch`                      SubgraphCountingMatching/dataset.py:1321-1328
  * reversed-edge append             SubgraphCountingMatching/train.py:299-327, dataset.py:1522-1563
  * `graph.out_degrees()`            SubgraphCountingMatching/models/dmpnn.py:100-101
  * stable COO->CSC/CSR              DGL-internal (stable counting sort by key, edge ids ascending)
  * UNC bidirectional graph + norm   UnsupervisedNodeClassification/Model/DMPNN/src/utils.py:437-491
  * degree coefficient               dmpnn.py:144-146  c_e = 2*(1+log2(1+deg(dst)))

Parity status: unpinned by DGL itself (not installable offline); consistent with the DGL shim that
the golden generator runs the unmodified reference layers on.
"""
import numpy as np


def batch_graphs(graphs):
    """graphs: list of (src, dst, num_nodes). Disjoint union, per-graph edge order preserved."""
    srcs, dsts, n_off = [], [], 0
    bn, be = [], []
    for s, d, n in graphs:
        srcs.append(np.asarray(s, dtype=np.int64) + n_off)
        dsts.append(np.asarray(d, dtype=np.int64) + n_off)
        n_off += int(n)
        bn.append(int(n))
        be.append(len(s))
    src = np.concatenate(srcs) if srcs else np.zeros(0, np.int64)
    dst = np.concatenate(dsts) if dsts else np.zeros(0, np.int64)
    return src, dst, n_off, np.asarray(bn, np.int64), np.asarray(be, np.int64)


def add_reversed_edges(src, dst):
    """Append (v,u) for every (u,v); flag = 1 on the appended half (train.py:299-313)."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    e0 = len(src)
    rev = np.concatenate([np.zeros(e0, bool), np.ones(e0, bool)])
    return np.concatenate([src, dst]), np.concatenate([dst, src]), rev


def out_degrees(src, num_nodes):
    return np.bincount(np.asarray(src, np.int64), minlength=num_nodes).astype(np.int64)


def in_degrees(dst, num_nodes):
    return np.bincount(np.asarray(dst, np.int64), minlength=num_nodes).astype(np.int64)


def stable_segments(key, num_nodes):
    """Stable counting sort of edge ids by `key`: (indptr[N+1], eid[E]) int32."""
    key = np.asarray(key, np.int64)
    eid = np.argsort(key, kind="stable").astype(np.int32)
    cnt = np.bincount(key, minlength=num_nodes)
    indptr = np.zeros(num_nodes + 1, np.int64)
    np.cumsum(cnt, out=indptr[1:])
    return indptr.astype(np.int32), eid


def endpoint_roles(src, dst, rev=None):
    """a_e = endpoint that meets W_dst, b_e = endpoint that meets W_src (dmpnn.py:112,120-123)."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    if rev is None:
        return dst.copy(), src.copy()
    r = np.asarray(rev, bool)
    return np.where(r, src, dst), np.where(r, dst, src)


def degree_coef(out_deg, dst):
    """c_e = 2*(1+log2(1+float(out_deg[dst_e]))) in fp32, torch CPU log2 (dmpnn.py:144-146)."""
    import torch
    d = torch.from_numpy(np.asarray(out_deg, np.int64))[torch.from_numpy(np.asarray(dst, np.int64))]
    d = d.float()
    c = 2 * (1 + (1 + d).log2())
    return c.numpy()


def build_plan(src, dst, num_nodes, rev=None, out_deg=None):
    """Everything `dmp_plan_build` must reproduce bit-exactly (coef: see degree_coef)."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    a, b = endpoint_roles(src, dst, rev)
    deg = out_degrees(src, num_nodes) if out_deg is None else np.asarray(out_deg, np.int64)
    csc_indptr, csc_eid = stable_segments(dst, num_nodes)
    a_indptr, a_eid = stable_segments(a, num_nodes)
    b_indptr, b_eid = stable_segments(b, num_nodes)
    return dict(dst32=dst.astype(np.int32), a32=a.astype(np.int32), b32=b.astype(np.int32),
                csc_indptr=csc_indptr, csc_eid=csc_eid, a_indptr=a_indptr, a_eid=a_eid,
                b_indptr=b_indptr, b_eid=b_eid, out_deg=deg, coef=degree_coef(deg, dst))


def compute_edgenorm_in(dst, num_nodes):
    """utils.py:437-453 with norm="in": 1/in_deg(dst), nan/inf replaced by the minimum."""
    import torch
    indeg = torch.from_numpy(in_degrees(dst, num_nodes)).float()
    norm = indeg[torch.from_numpy(np.asarray(dst, np.int64))].reciprocal().unsqueeze(-1)
    norm.masked_fill_(torch.isnan(norm), norm.min())
    norm.masked_fill_(torch.isinf(norm), norm.min())
    return norm.numpy()


def build_graph_from_triplets(num_nodes, num_rels, triplets):
    """utils.py:473-491: sort by (src,dst,rel); forward block then reversed block; type += R."""
    t = np.asarray(triplets, np.int64)
    order = np.lexsort((t[:, 1], t[:, 2], t[:, 0]))  # primary src, then dst, then rel
    t = t[order]
    src = np.concatenate([t[:, 0], t[:, 2]])
    dst = np.concatenate([t[:, 2], t[:, 0]])
    rel = np.concatenate([t[:, 1], t[:, 1] + num_rels])
    return src, dst, rel, compute_edgenorm_in(dst, num_nodes)


def erdos_renyi(rng, n, e0, self_loops=False):
    """Directed ER multigraph: e0 ordered pairs uniformly with replacement."""
    u = rng.integers(0, n, size=e0, dtype=np.int64)
    if self_loops or n < 2:
        v = rng.integers(0, n, size=e0, dtype=np.int64)
    else:
        v = rng.integers(0, n - 1, size=e0, dtype=np.int64)
        v = v + (v >= u)
    return u, v


def split_and_batchify(feats, sizes, pre_pad=False):
    """SubgraphCountingMatching/utils/dl.py:51-81 restated with the reference's own loop: ragged [sum(sizes), H] ->
    (padded [B, max, H], mask [B, max])."""
    feats = np.asarray(feats)
    sizes = [int(x) for x in np.asarray(sizes).reshape(-1)]
    bsz, mx = len(sizes), max(sizes)
    out = np.zeros((bsz, mx) + feats.shape[1:], feats.dtype)
    mask = np.zeros((bsz, mx), bool)
    idx = 0
    for i, l in enumerate(sizes):
        if pre_pad:
            out[i, mx - l:] = feats[idx:idx + l]
            mask[i, mx - l:] = True
        else:
            out[i, :l] = feats[idx:idx + l]
            mask[i, :l] = True
        idx += l
    return out, mask


def len_to_mask(lens, max_len=-1, pre_pad=False):
    """dl.py:113-127 restated (loop over the batch)."""
    lens = [int(x) for x in lens]
    if max_len == -1:
        max_len = max(lens)
    mask = np.ones((len(lens), max_len), bool)
    for i, l in enumerate(lens):
        if pre_pad:
            mask[i, :max_len - l] = False
        else:
            mask[i, l:] = False
    return mask
