"""ctypes wrapper over oracle/libdmp_oracle.so (sequential C restatement of the sparse core).

TEST INFRASTRUCTURE ONLY -- see oracle/sparse_core.c.  Works on contiguous CPU torch tensors.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdmp_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _i64(x):
    return ctypes.c_int64(int(x))


def seg_reduce(indptr, eid, V, H, *, w_perm=None, rev_off=0, base=None, bias=None, mode=0, ldV=None):
    nseg = indptr.numel() - 1
    out = torch.empty((nseg, 2 * H if mode & 16 else H), dtype=torch.float32)   # 16 = SPLIT_BY_REV: [fwd | rev]
    ldV = V.stride(0) if ldV is None else ldV
    lib().oracle_seg_reduce(_p(indptr), _p(eid), _p(w_perm), _p(V), _i64(ldV), _i64(rev_off), _p(base),
                            _i64(base.stride(0) if base is not None else 0), _p(bias), _p(out), _i64(out.stride(0)),
                            _i64(nseg), _i64(H), ctypes.c_int(mode))
    return out


def edge_update(a32, b32, coef, S, P, Qd, Qs, ebias, order, want_agg=False):
    E, H = S.shape
    out = torch.empty((E, H), dtype=torch.float32)
    agg = torch.empty((E, H), dtype=torch.float32) if want_agg else None
    lib().oracle_edge_update(_p(a32), _p(b32), _p(coef), _p(S), _i64(S.stride(0)), _p(P), _i64(P.stride(0)),
                             _p(Qd), _i64(Qd.stride(0)), _p(Qs), _i64(Qs.stride(0)), _p(ebias), _p(out),
                             _i64(H), _p(agg), _i64(H), _i64(E), _i64(H), ctypes.c_int(order))
    return (out, agg) if want_agg else out


def edge_backward(dst32, rev, norm, coef, gN, gE, t_rev_off=0, gN_rev=None):
    E, H = gE.shape
    T = torch.zeros((E, H + t_rev_off), dtype=torch.float32)
    CG = torch.empty((E, H), dtype=torch.float32)
    lib().oracle_edge_backward(_p(dst32), _p(rev), _p(norm), _p(coef), _p(gN), _p(gN_rev), _i64(gN.stride(0)), _p(gE),
                               _i64(gE.stride(0)), _p(T), _i64(T.stride(0)), _i64(t_rev_off), _p(CG), _i64(H),
                               _i64(E), _i64(H))
    return T, CG


def stable_segments(key32, rev, num_nodes):
    E = key32.numel()
    indptr = torch.empty(num_nodes + 1, dtype=torch.int32)
    eid = torch.empty(E, dtype=torch.int32)
    lib().oracle_stable_segments(_p(key32), _p(rev), _i64(num_nodes), _i64(E), _p(indptr), _p(eid))
    return indptr, eid


def gate_residual(x, gate, prev, act, slope):
    rows, H = x.shape
    out = torch.empty((rows, H), dtype=torch.float32)
    lib().oracle_gate_residual(_p(x), _i64(x.stride(0)), _p(gate), _p(prev),
                               _i64(prev.stride(0) if prev is not None else 0), _p(out), _i64(H), _i64(rows),
                               _i64(H), ctypes.c_int(act), ctypes.c_float(slope))
    return out


def gate_residual_backward(gout, x, gate, act, slope):
    rows, H = gout.shape
    gx = torch.empty((rows, H), dtype=torch.float32)
    lib().oracle_gate_residual_backward(_p(gout), _i64(gout.stride(0)), _p(x),
                                        _i64(x.stride(0) if x is not None else 0), _p(gate), _p(gx), _i64(H),
                                        _i64(rows), _i64(H), ctypes.c_int(act), ctypes.c_float(slope))
    return gx
