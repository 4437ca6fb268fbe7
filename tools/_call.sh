timeout 170 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -4
