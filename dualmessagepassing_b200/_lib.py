"""ctypes binding of libdmp_b200.so (C ABI declared in include/dmp_b200.h).

The library is built in-tree (`dualmessagepassing_b200/csrc/Makefile`, sm_100a only).  There is no
CPU or PyTorch fallback: if the shared object is missing, or a CUDA tensor is not supplied, the
calls raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMP_B200_LIB") or os.path.join(_HERE, "libdmp_b200.so")   # override: A/B builds

# mirrors of the #defines in include/dmp_b200.h
EID_MASK = 0x7FFFFFFF
SEG_SIGN_BY_REV = 1
SEG_NEGATE_OUT = 2
SEG_ONLY_FWD = 4
SEG_ONLY_REV = 8
SEG_SPLIT_BY_REV = 16
SEG_SHORT = 32
ORDER_SCM = 0
ORDER_UNC = 1
EDGE_MIRRORED_HALVES = 16
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
ACT_FROM_OUTPUT = 16
EPI_MUL_ACT_GRAD = 32
EPI_ACCUMULATE = 64
DUAL_STORE, DUAL_ACCUMULATE, DUAL_SEPARATE = 0, 1, 2

_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float

# name -> argtypes; every entry must be exported by the library (tests/test_abi.py checks the header too)
SIGNATURES = {
    "dmp_version": [],
    "dmp_set_sm_reserve": [_i32],
    "dmp_plan_workspace_bytes": [_i64, _i64, ctypes.POINTER(ctypes.c_int64)],
    "dmp_plan_build": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp, _vp, _vp, _vp, _i64, _vp],
    "dmp_permute_edge_scalar": [_vp, _vp, _vp, _i64, _vp],
    "dmp_segment_reduce": [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _vp],
    "dmp_edge_update": [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64,
                        _i64, _i64, _i32, _vp],
    "dmp_edge_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp],
    "dmp_gate_residual": [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _f32, _vp],
    "dmp_gate_residual_backward": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _vp],
    "dmp_gemm_tn_workspace_bytes": [_i64, _i64, ctypes.POINTER(ctypes.c_int64)],
    "dmp_gemm_tn_tf32x3": [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp],
    "dmp_gemm_tf32x3": [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _vp],
    "dmp_gemm_tf32x3_dual": [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "dmp_batch_offsets": [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp],
    "dmp_batch_fill": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                       _vp, _vp, _vp],
    "dmp_ragged_pad": [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp, _vp],
    "dmp_ragged_unpad": [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _i64, _vp],
    "dmp_bn_workspace_bytes": [_i64, ctypes.POINTER(ctypes.c_int64)],
    "dmp_bn_stats": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp],
    "dmp_bn_act": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _vp],
    "dmp_bn_backward": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp],
}

_lib = None


def load():
    """Load the shared object once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdmp_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dualmessagepassing_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.dmp_last_error.restype = ctypes.c_char_p
    lib.dmp_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dmp_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))


# ---- optional per-launch timing (bench.py): CUDA events on the launching stream around every C-ABI call
PROFILE = None   # None = off; else list of (tag, start_event, end_event, algorithmic bytes or None)
LAUNCHES = 0     # number of C-ABI kernel-launching calls made so far (bench.py reports the delta)


def call(name, device, *args, tag=None, nbytes=None):
    """Invoke one entry point on `device`'s current stream; raise on a non-zero status."""
    global LAUNCHES
    fn = getattr(load(), name)
    if PROFILE is not None:
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn(*args)
            e1.record(stream)
            PROFILE.append((tag or name, e0, e1, nbytes))
    elif device.index is None or device.index == torch.cuda.current_device():
        rc = fn(*args)          # common case: no device switch (the guard costs more than the launch on small graphs)
    else:
        with torch.cuda.device(device):
            rc = fn(*args)
    LAUNCHES += 1
    check(rc, name)


# SMs left to an overlapped NCCL collective while it is in flight (parallel.py / fused.py); 0 disables
SM_RESERVE = int(os.environ.get("DMP_SM_RESERVE", "0"))


def sm_reserve(n):
    check(load().dmp_set_sm_reserve(int(n)), "dmp_set_sm_reserve")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dualmessagepassing_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor"
                               % t.device)


def row_major(t):
    """fp32 matrix whose rows are dense (stride(1) == 1); returns (tensor, leading dimension)."""
    if t.dtype != torch.float32:
        raise TypeError("expected float32, got %s" % t.dtype)
    if t.dim() != 2:
        raise ValueError("expected a matrix, got shape %s" % (tuple(t.shape),))
    if t.shape[1] > 1 and t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, max(ld, t.shape[1])
