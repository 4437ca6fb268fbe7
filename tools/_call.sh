mkdir -p gpurun_out/r2
for v in 256 128 64 0 256 128; do
  PEA_ATTN2_L2PROMO=$v timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_promo$v.jsonl > /dev/null 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
rows=[d for d in map(json.loads,open(f'gpurun_out/r2/mb_promo{v}.jsonl')) if d.get('op')=='attn2']
print('promo',v, ' | '.join('%d/%d f %.1f b %.1f (%.3f)'%(x['N'],x['C'],x['fwd_us'],x['bwd_us'],x['fwdbwd_frac']) for x in rows if x['L']==77))
PY
done
for v in 256 128; do for m in fwd bwd; do
  PEA_ATTN2_L2PROMO=$v timeout 100 python tools/chain_probe.py --mode $m --N 1024 --C 1280 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('promo $v',d['mode'],d['N'],d['C'],'marginal %.2f bracketed %.2f'%(d['marginal_us_per_launch'],d['bracketed_us_median']))"
done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attn2" 2>&1 | tail -2
