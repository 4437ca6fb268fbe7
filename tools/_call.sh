mkdir -p gpurun_out/r2
timeout 100 python tools/capture_probe.py 1 1 0 2>&1 | grep PROBE
timeout 600 python -m pytest tests -m gpu -x -q -rP > gpurun_out/r2/gputests.log 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2/gputests.log | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench1.json 2> gpurun_out/r2/bench1.err; echo "bench rc=$?"
grep "bench \|Error" gpurun_out/r2/bench1.err | tail -16
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dp_params_identical','loss')})
    print('e2e',d['e2e']); print('strong',d['strong']); print('infer',d.get('infer')); print('cpu',d.get('cpu_baseline'))
    r=d['roofline']; print('roof',r['frac'],r.get('in_step_marginal')); print(r['per_kernel'])
    for row in r.get('sweep',[]): print(row)
    for row in r.get('sweep_gemm',[]): print(row)
    print(r.get('sweep_error'), r.get('sweep_clocks'))
except Exception as e: print('parse failed',e)
PY
