mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2/final_gputests.txt; cat gpurun_out/r2/final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_1gpu.json 2> gpurun_out/r2/final_bench_1gpu.err; echo bench rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn2_.*_tc_kernel -c 4 -f -o gpurun_out/r2/ncu_attn2_big_after python tools/microbench.py --only attn --filter 16384:1280 --iters 1 --warm 1 --out gpurun_out/r2/mb_under_ncu3.jsonl > gpurun_out/r2/ncu_attn2_big_after.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn2_.*_tc_kernel -c 12 -f -o gpurun_out/r2/ncu_attn2_final3 python tools/microbench.py --only attn --filter 1024:1280,4096:640 --iters 2 --warm 1 --out gpurun_out/r2/mb_under_ncu4.jsonl > gpurun_out/r2/ncu_attn2_final3.log 2>&1; echo ncu2 rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], 'frac',r['frac'], r['in_chain_marginal']['step_mix']['frac'])
for s in r['sweep']: print(s['N'],s['C'],s['L'],s['fwd_us'],s['bwd_us'],s['fwd_frac'],s['bwd_frac'],s['fwdbwd_frac'])
PY
