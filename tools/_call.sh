mkdir -p gpurun_out/r2
T=build/variants
PEA_B200_LIB=$T/libpea_trace.so python tools/trace_fwd.py 1024 1280 > gpurun_out/r2/trace_fwd_1024_cold.txt 2>&1
PEA_B200_LIB=$T/libpea_trace.so python tools/trace_fwd.py 1024 1280 warm > gpurun_out/r2/trace_fwd_1024_warm.txt 2>&1
PEA_B200_LIB=$T/libpea_trace.so python tools/trace_bwd.py 1024 1280 > gpurun_out/r2/trace_bwd_1024.txt 2>&1
for m in fwd bwd; do
  python tools/chain_probe.py --mode $m >> gpurun_out/r2/chain.jsonl 2>> gpurun_out/r2/chain.err
  python tools/chain_probe.py --mode $m --N 4096 --C 640 >> gpurun_out/r2/chain.jsonl 2>> gpurun_out/r2/chain.err
done
python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_base.jsonl > /dev/null 2> gpurun_out/r2/mb_base.err
PEA_B200_LIB=$T/libpea_waitread.so python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_waitread.jsonl > /dev/null 2>&1
PEA_B200_LIB=$T/libpea_parinit.so python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_parinit.jsonl > /dev/null 2>&1
python tools/microbench.py --only attn --iters 30 --out gpurun_out/r2/mb_base2.jsonl > /dev/null 2>&1
cat gpurun_out/r2/chain.jsonl
