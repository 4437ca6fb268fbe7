mkdir -p gpurun_out/r2
for cfg in "0 0 0" "0 1 0" "1 0 0" "1 1 0" "1 0 1" "1 1 1" "0 1 1"; do
  timeout 100 python tools/capture_probe.py $cfg 2>&1 | grep PROBE
done
T=build/variants
PEA_B200_LIB=$T/libpea_trace.so timeout 100 python tools/trace_fwd.py 1024 1280 warm > gpurun_out/r2/trace_fwd_1024_warm_b.txt 2>&1
head -4 gpurun_out/r2/trace_fwd_1024_warm_b.txt | cut -c1-600
PEA_B200_LIB=$T/libpea_split.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "attn2" 2>&1 | tail -2
PEA_B200_LIB=$T/libpea_split_trace.so timeout 100 python tools/trace_bwd.py 1024 1280 > gpurun_out/r2/trace_bwd_1024_split.txt 2>&1
tail -3 gpurun_out/r2/trace_bwd_1024_split.txt | cut -c1-400
timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_base3.jsonl > /dev/null 2>&1
PEA_B200_LIB=$T/libpea_split.so timeout 200 python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_split.jsonl > /dev/null 2>&1
python - <<'PY'
import json
def load(f): return [json.loads(l) for l in open(f)]
a,b=load('gpurun_out/r2/mb_base3.jsonl'),load('gpurun_out/r2/mb_split.jsonl')
for x,y in zip(a,b): print(x['N'],x['C'],x['L'],'bwd base %.1f split %.1f | fwd %.1f %.1f | fb frac %.3f -> %.3f'%(x['bwd_us'],y['bwd_us'],x['fwd_us'],y['fwd_us'],x['fwdbwd_frac'],y['fwdbwd_frac']))
PY
