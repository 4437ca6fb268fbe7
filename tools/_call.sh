mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2/final_gputests.txt; cat gpurun_out/r2/final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_1gpu.json 2> gpurun_out/r2/final_bench_1gpu.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2/final_bench_ref.json 2> gpurun_out/r2/final_bench_ref.err; echo ref rc=$?; tail -c 600 gpurun_out/r2/final_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], 'frac',r['frac'], r['in_chain_marginal']['step_mix']['frac'], {k:round(v,1) for k,v in r['adapter_fwd_bwd'].items() if k.endswith('_us')})
PY
