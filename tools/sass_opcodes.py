#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths: disassemble the built libpea_b200.so (cuobjdump -sass) and count,
per kernel, the opcodes that only tcgen05 / TMEM / TMA code contains -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit), SYNCS (mbarrier), ACQBULK / griddepcontrol (PDL) -- next to
the legacy HMMA (mma.sync) count.  Runs on the CPU-only build box.

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pea_diffusion_b200", "libpea_b200.so")
PATTERNS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
            ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"),
            ("UTCATOMSWS", r"\bUTCATOMSWS"), ("PDL(ACQBULK/PREEXIT)", r"\b(ACQBULK|PREEXIT)"), ("HMMA", r"\bHMMA"),
            ("MUFU.EX2", r"\bMUFU\.EX2"), ("RED/ATOM", r"^(RED|REDG|ATOM|ATOMG)\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        kernels[cur]["instructions"] += 1
        for name, pat in PATTERNS:
            if re.search(pat, m.group(1)):
                kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    cols = ["instructions"] + [n for n, _ in PATTERNS]
    print(f"# {os.path.relpath(LIB, ROOT)}: cuobjdump -sass opcode counts per kernel (sm_100a).  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
    print("# UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync.")
    print("# " + " | ".join(f"{c:>12s}" for c in cols) + " | kernel")
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        name = name.replace("pea::(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        print("  " + " | ".join(f"{cnt.get(c, 0):12d}" for c in cols) + " | " + name)


if __name__ == "__main__":
    main()
