mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_step.py -q -m gpu -k "nccl or two_gpu or 2gpu or dp" 2>&1 | tail -4 > gpurun_out/r2/final_gputests_2gpu.txt; cat gpurun_out/r2/final_gputests_2gpu.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_2gpu.json 2> gpurun_out/r2/final_bench_2gpu.err; echo bench rc=$?; tail -c 400 gpurun_out/r2/final_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','dp_params_identical')}, d['e2e'], d['strong'])
PY
