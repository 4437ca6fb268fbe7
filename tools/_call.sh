mkdir -p gpurun_out/r2
T=build/variants
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attn2" 2>&1 | tail -2
PEA_B200_LIB=$T/libpea_nopf.so timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_nopf.jsonl > /dev/null 2>&1
timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_pf.jsonl > /dev/null 2>&1
python - <<'PY'
import json
def load(f): return [d for d in map(json.loads,open(f)) if d.get('op')=='attn2']
a,b=load('gpurun_out/r2/mb_nopf.jsonl'),load('gpurun_out/r2/mb_pf.jsonl')
for x,y in zip(a,b): print(x['N'],x['C'],x['L'],'fwd nopf %.1f pf %.1f | bwd nopf %.1f pf %.1f | fb %.3f -> %.3f'%(x['fwd_us'],y['fwd_us'],x['bwd_us'],y['bwd_us'],x['fwdbwd_frac'],y['fwdbwd_frac']))
PY
for lib in nopf product; do for m in fwd bwd; do for cfg in "1024 1280" "4096 640"; do set -- $cfg
  if [ $lib = nopf ]; then export PEA_B200_LIB=$T/libpea_nopf.so; else unset PEA_B200_LIB; fi
  timeout 100 python tools/chain_probe.py --mode $m --N $1 --C $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib',d['mode'],d['N'],d['C'],'marginal %.2f bracketed %.2f'%(d['marginal_us_per_launch'],d['bracketed_us_median']))"
done; done; done
