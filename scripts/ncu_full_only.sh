mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"segment_reduce_kernel|edge_update|edge_backward_kernel|tf32x3_gemm_kernel|tf32x3_gemm_tn_kernel" \
  -f -o gpurun_out/prof_final2_r1 python scripts/ncu_kernels_fullsize.py > gpurun_out/ncu_full_final2.log 2>&1
tail -2 gpurun_out/ncu_full_final2.log
