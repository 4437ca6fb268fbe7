mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "adapter or layernorm or rowop or split or mean or colsum" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_step.py -x -q -k "sink or accumulate or fp16 or kd" 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tail -5
import json, torch, bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for i in range(3):
    r = bench.adapter_subrun(dev, replays=30)
    print({k: round(v, 1) for k, v in r.items() if k.endswith("_us")})
PY
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
