#!/bin/bash
# Ablation of the K = E reduction kernel at config-5 size: builds a -DDMP_DEBUG copy of the library next to the release
# one (scripts/micro/libdmp_dbg.so, git-ignored) and times dmp_gemm_tn_tf32x3 with parts of the kernel switched off
# (DMP_TN_ABLATE bits: 1 no proxy fence, 2 no MMA, 4 no loads, 8 no split, 16 no flush, 32 blocked instead of interleaved
# stages, 64 no X loads, 128 no G loads).  usage: scripts/tn_ablate.sh [build|run]
set -e
cd "$(dirname "$0")/.."
C=dualmessagepassing_b200/csrc
if [ "$1" != "run" ]; then
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    -Xcompiler -fvisibility=hidden --fmad=false -Wno-deprecated-gpu-targets -DDMP_DEBUG $TN_DEFS -c $C/tf32x3_gemm_tn.cu -o /tmp/tn_dbg.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scripts/micro/libdmp_dbg.so /tmp/tn_dbg.o \
    $C/api.o $C/collate.o $C/plan.o $C/segment_reduce.o $C/edge_kernels.o $C/bn_kernels.o $C/tf32x3_gemm.o -cudart static
fi
if [ "$1" != "build" ]; then
  for a in 0 2 4 8 16 64 128 6 10 1 32; do
    DMP_B200_LIB=$PWD/scripts/micro/libdmp_dbg.so DMP_TN_ABLATE=$a python - <<PY
import os, sys, torch
sys.path.insert(0, ".")
from dualmessagepassing_b200 import functional as F
E = 40_000_000
X, G = torch.randn(E, 128, device="cuda"), torch.randn(E, 128, device="cuda")
F.gemm_tn_tf32x3(X, G); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): F.gemm_tn_tf32x3(X, G)
b.record(); torch.cuda.synchronize()
print("ablate %3s: %.3f ms" % (os.environ["DMP_TN_ABLATE"], a.elapsed_time(b) / 5))
PY
  done
fi
