mkdir -p gpurun_out/r2
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_step.py -x -q -k "dp_two" > gpurun_out/r2/gputests_2gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2/gputests_2gpu.log | cut -c1-300
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err; echo "bench2 rc=$?"
grep "bench \|Error\|error" gpurun_out/r2/bench_2gpu.err | tail -12 | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','dp_params_identical','dp_param_checksums','loss')})
    print('e2e',d['e2e']); print('strong',d['strong']); print(d['execution'])
except Exception as e: print('parse failed',e)
PY
