mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "attn2 or group_projection" 2>&1 | tail -2
for v in base product base product; do
  if [ $v = product ]; then unset PEA_B200_LIB; else export PEA_B200_LIB=build/variants/libpea_$v.so; fi
  timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_$v.jsonl > /dev/null 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[d for d in map(json.loads,open(f'gpurun_out/r2/mb_{v}.jsonl')) if d.get('op')=='attn2']
print(v, ' | '.join('%d/%d b %.1f (%.3f)'%(x['N'],x['C'],x['bwd_us'],x['bwd_frac']) for x in rows if x['L']==77))
PY
done
for v in base product; do
  if [ $v = product ]; then unset PEA_B200_LIB; else export PEA_B200_LIB=build/variants/libpea_$v.so; fi
  for cfg in "1024 1280" "4096 640"; do set -- $cfg
  timeout 100 python tools/chain_probe.py --mode bwd --N $1 --C $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v',d['mode'],d['N'],d['C'],'marginal %.2f bracketed %.2f'%(d['marginal_us_per_launch'],d['bracketed_us_median']))"
done; done
