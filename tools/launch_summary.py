"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into one training step:
the launches between the last two `ema_kernel` launches (the EMA opens every step).
usage: python tools/launch_summary.py gpurun_out/launches.csv "header line" > profiles/<name>.txt"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "us")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
    rows.append((r["Kernel Name"], us, r["Grid Size"], r["Block Size"]))
emas = [i for i, r in enumerate(rows) if r[0].startswith("ema_kernel") or "ema_kernel" in r[0]]
assert len(emas) >= 2, "need at least two steps in the capture"
step = rows[emas[-2]:emas[-1]]
short = lambda n: n.split("(")[0].replace("void ", "").replace("rsp::", "")[:78]
agg = defaultdict(lambda: [0.0, 0])
for n, us, g, b in step:
    agg[short(n)][0] += us
    agg[short(n)][1] += 1
total = sum(us for _, us, _, _ in step)
print(sys.argv[2] if len(sys.argv) > 2 else "# launch list")
print(f"# launches in the step: {len(step)}; sum of kernel durations {total:.1f} us (serialised by ncu: streams do not overlap here)")
print("        us  share count  kernel")
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} {100 * us / total:5.1f}% {c:5d}  {n}")
print("\n# every launch in order: duration_us grid block kernel")
for n, us, g, b in step:
    print(f"{us:9.1f} {g:>16s} {b:>14s} {short(n)}")
