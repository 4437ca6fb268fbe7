mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_step.py -x -q -k "dp_two" > gpurun_out/r2/gputests_2gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2/gputests_2gpu.log | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err; echo "bench2 rc=$?"
grep "bench \|Error\|error" gpurun_out/r2/bench_2gpu.err | tail -8 | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','dp_params_identical','dp_param_checksums','loss')})
    print('e2e',d['e2e']); print('strong',d['strong']); print(d['execution'])
except Exception as e: print('parse failed',e)
PY
# experimental: collectives captured in the graph (short leash)
python - <<'PY'
import os, subprocess, sys, re
src=open('tests/test_gpu_step.py').read()
m=re.search(r"_DP_NCCL_WORKER = r'''(.*?)'''", src, re.S)
open('/tmp/dpw.py','w').write(m.group(1))
procs=[]
for r in range(2):
    env=dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT='29533')
    procs.append(subprocess.Popen([sys.executable,'/tmp/dpw.py',os.getcwd(),'graphcap','fp32'],env=env,stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True))
for p in procs:
    try:
        out=p.communicate(timeout=150)[0]
        print('graphcap rc',p.returncode, out[-400:].replace('\n',' | '))
    except subprocess.TimeoutExpired:
        for q in procs: q.kill()
        print('graphcap TIMEOUT'); break
PY
