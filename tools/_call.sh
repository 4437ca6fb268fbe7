export PEA_B200_LIB=build/variants/libpea_trace.so
timeout 120 python tools/trace_bwd.py 16384 1280 2>&1 | cut -c1-330
echo; timeout 120 python tools/trace_bwd.py 1024 1280 2>&1 | cut -c1-330
