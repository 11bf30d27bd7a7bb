#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) --set full of one full-size launch of every
# hand-written hot kernel.  Single GPU only.  Outputs under gpurun_out/; summarise here with scripts/summarise_ncu.py.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_cfg5_r1_final2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train --no-mlp0 > gpurun_out/ncu_list_final2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"segment_reduce_kernel|edge_update|edge_backward_kernel|tf32x3_gemm_kernel|tf32x3_gemm_tn_kernel" \
  -f -o gpurun_out/prof_final2_r1 python scripts/ncu_kernels_fullsize.py > gpurun_out/ncu_full_final2.log 2>&1
tail -3 gpurun_out/ncu_full_final2.log
ls -la gpurun_out/prof_final2_r1.ncu-rep
