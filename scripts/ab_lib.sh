#!/bin/bash
# A/B two builds of the library on the same box: DMP_B200_LIB selects the .so (see _lib.py)
for rep in 1 2; do
for lib in dualmessagepassing_b200/libdmp_b200_old.so dualmessagepassing_b200/libdmp_b200.so; do
DMP_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-train --no-mlp0 --steps 5 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
ks=j['kernels']
print('$lib', round(j['ms_per_step'],1), ' '.join('%s=%.2f' % (k.replace('gemm_tf32x3.','').replace('gemm_','').replace('segment_reduce.',''), v['avg_ms']) for k,v in ks.items() if '@N' not in k))
"
done; done
