#!/usr/bin/env python
"""Summarise an ncu report (read on the GPU-less build box) into a small text file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_attn_r1.ncu-rep profiles/r1_attn2_mma_sync.txt
    python tools/ncu_summary.py --launches gpurun_out/launches_r1.csv profiles/r1_launches_step_forward.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep} (per launch; read with ncu -i ... --page raw --csv)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\nkernel: {name[:150]}\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"  {k:80s} {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
            try:
                t = float(r[hdr.index('gpu__time_duration.sum')])
                tu = units[hdr.index('gpu__time_duration.sum')]
                t_us = t / 1e3 if tu in ("ns", "nsecond") else t if tu in ("us", "usecond") else t * 1e3
                rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
                sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tb = rd * sc[units[hdr.index("dram__bytes_read.sum")]] + wr * sc[units[hdr.index("dram__bytes_write.sum")]]
                f.write(f"  {'derived: dram traffic (read+write) bytes':80s} {tb:.0f}\n")
                f.write(f"  {'derived: dram GB/s over gpu__time_duration':80s} {tb / t_us / 1e3:.1f}\n")
            except Exception as e:  # noqa
                f.write(f"  derived: n/a ({e})\n")


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:110]
        tot[short] += v
        cnt[short] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of {path}: {sum(cnt.values())} launches, {T/1e3:.2f} ms serialised\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'total_us':>12s} {'share':>7s} {'n':>6s} {'avg_us':>9s}  kernel\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{v:12.1f} {100*v/T:6.2f}% {cnt[k]:6d} {v/cnt[k]:9.1f}  {k}\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[1], sys.argv[2])
