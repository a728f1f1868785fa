"""Per-instruction stall samples from an `ncu --page source --csv` export: prints the instructions around the tcgen05.mma
(UTCHMMA) issue region with their sample counts and dominant stall reasons.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > src.csv; python tools/ncu_source_stalls.py src.csv [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
total = 0
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = int(r[col["# Samples"]] or 0)
    total += n
    data.append((r, n))
print(f"total samples {total}")
for i, (r, n) in enumerate(data):
    if n >= mins or "UTCHMMA" in r[col["Source"]] or "UTCBAR" in r[col["Source"]]:
        st = sorted(((int(r[col[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
        st = [(c, v) for v, c in st if v]
        print(f"{i:5d} {n:6d} {r[col['Instructions Executed']]:>9} {r[col['Source']][:90]:<90} {st}")
