mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -k "two_gpus and graphcap" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_2gpu.json 2> gpurun_out/r2/final_bench_2gpu.err; echo bench rc=$?; tail -c 500 gpurun_out/r2/final_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','dp_params_identical')}, d.get('comm_in_graph'))
PY
