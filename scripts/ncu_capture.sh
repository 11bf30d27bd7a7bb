# usage: scripts/ncu_capture.sh <tag>   (on the GPU box; ~5 min)
#  1. launch list of the bench command (per-launch gpu__time_duration, cold cache, serialised: compare SHARES)
#  2. --set full of one full-size launch of every hot kernel (scripts/ncu_kernels_fullsize.py)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train --no-mlp0 --no-cfg4 --no-gpu-baseline \
  > gpurun_out/ncu_list_$1.log 2>&1
tail -1 gpurun_out/ncu_list_$1.log | head -c 300
bash scripts/ncu_full_only.sh $1
