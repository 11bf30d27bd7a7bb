DMP_GEMM_TMA=1 timeout 150 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2
for t in 0 1; do DMP_GEMM_TMA=$t timeout 100 python scripts/gemm_bench.py 40000000 2>&1 | grep "N=128" | sed "s/^/TMA=$t /"; done
for rep in 1 2; do for t in 0 1; do
DMP_GEMM_TMA=$t timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-train --no-mlp0 --steps 5 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read())
ks=j['kernels']
print('TMA=$t', round(j['ms_per_step'],1), j['clocks']['sm_mhz'], ' '.join('%s=%.2f' % (k.replace('gemm_tf32x3.',''), v['avg_ms']) for k,v in ks.items() if 'gemm_tf32x3' in k and '@E' in k))
"
done; done
