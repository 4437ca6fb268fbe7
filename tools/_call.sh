mkdir -p gpurun_out/r2
T=build/variants
for v in product q6 q7 product q7; do
  if [ $v = product ]; then unset PEA_B200_LIB; else export PEA_B200_LIB=$T/libpea_$v.so; fi
  timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_$v.jsonl > /dev/null 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[d for d in map(json.loads,open(f'gpurun_out/r2/mb_{v}.jsonl')) if d.get('op')=='attn2']
print(v, ' | '.join('%d/%d %.1f (%.3f)'%(x['N'],x['C'],x['fwd_us'],x['fwd_frac']) for x in rows))
PY
done
PEA_B200_LIB=$T/libpea_q7.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attn2_core or full_size_rows" 2>&1 | tail -2
for v in product q7; do
  if [ $v = product ]; then unset PEA_B200_LIB; else export PEA_B200_LIB=$T/libpea_$v.so; fi
  for cfg in "1024 1280" "4096 640"; do set -- $cfg
  timeout 100 python tools/chain_probe.py --mode fwd --N $1 --C $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v',d['mode'],d['N'],d['C'],'marginal %.2f bracketed %.2f'%(d['marginal_us_per_launch'],d['bracketed_us_median']))"
done; done
