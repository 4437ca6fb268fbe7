mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/gputests.log | cut -c1-300
PEA_GEMM_PLAIN_EPI=0 PEA_GEMM_BN=128 timeout 200 python tools/microbench.py --only gemm --iters 30 --out gpurun_out/r2/mb_gemm_old.jsonl > /dev/null 2>&1
PEA_GEMM_BN=128 timeout 200 python tools/microbench.py --only gemm --iters 30 --out gpurun_out/r2/mb_gemm_plain128.jsonl > /dev/null 2>&1
timeout 200 python tools/microbench.py --only gemm --iters 30 --out gpurun_out/r2/mb_gemm_new.jsonl > /dev/null 2>&1
python - <<'PY'
import json
def load(f): return {d['name']:d for d in map(json.loads,open(f)) if d.get('op')=='gemm'}
a,b,c=[load('gpurun_out/r2/mb_gemm_%s.jsonl'%v) for v in ('old','plain128','new')]
for k in a: print('%-26s old(bn128,general epi) %.1f | bn128+plain epi %.1f | new %.1f | cublas %.1f'%(k,a[k]['us'],b[k]['us'],c[k]['us'],c[k]['cublas_us']))
PY
for cfg in "1024 1280" "4096 640"; do set -- $cfg; timeout 100 python tools/chain_probe.py --mode fwd --N $1 --C $2 --own-gemm 1 --pdl 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('own-gemm chain',d['N'],d['C'],'gemm pair per rep %.1f us, attn marginal %.2f'%(d['graph_without_us']/20,d['marginal_us_per_launch']))"; done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench3.json 2> gpurun_out/r2/bench3.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench3.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dp_params_identical')})
    r=d['roofline']; print('roof',r['frac'],r['traffic_source']); print(r['in_chain_marginal'])
    print({k:v['us_per_launch'] for k,v in r['per_kernel'].items()})
    for row in r.get('sweep',[]): print(row['N'],row['C'],row['L'],row['fwd_us'],row['bwd_us'],row['fwd_frac'],row['bwd_frac'],row['fwdbwd_frac'])
    for row in r.get('sweep_gemm',[]): print(row)
except Exception as e: print('parse failed',e)
PY
