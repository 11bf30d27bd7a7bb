# usage: scripts/ncu_full_only.sh <tag>   -> gpurun_out/prof_<tag>.ncu-rep (+ raw csv)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:"segment_reduce_kernel|edge_update|edge_backward_kernel|tf32x3_gemm" \
  -f -o gpurun_out/prof_$1 python scripts/ncu_kernels_fullsize.py > gpurun_out/ncu_full_$1.log 2>&1
tail -2 gpurun_out/ncu_full_$1.log
ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
