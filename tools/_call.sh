mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/gputests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench5.json 2> gpurun_out/r2/bench5.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench5.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dp_params_identical')})
    r=d['roofline']; print('roof',r['frac']); print(r['in_chain_marginal']['step_mix']); print(r.get('adapter_fwd_bwd'))
    print({k:v['us_per_launch'] for k,v in r['per_kernel'].items()})
    print(r.get('sweep_error'))
except Exception as e: print('parse failed',e)
PY
