mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/final_bench_8gpu.json 2> gpurun_out/r2/final_bench_8gpu.err; echo bench rc=$?; tail -c 300 gpurun_out/r2/final_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/final_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','dp_params_identical')}, d['e2e']['value'], d['strong']['value'], d.get('comm_in_graph'))
PY
