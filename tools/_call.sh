timeout 120 python tools/_dbg_lse.py 2>&1 | tail -12
