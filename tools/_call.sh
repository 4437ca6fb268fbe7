mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2/final_gputests.txt; cat gpurun_out/r2/final_gputests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_1gpu.json 2> gpurun_out/r2/final_bench_1gpu.err; echo bench rc=$?; tail -c 600 gpurun_out/r2/final_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], 'frac',r['frac'], r.get('in_chain_marginal'), r.get('adapter_fwd_bwd'))
PY
