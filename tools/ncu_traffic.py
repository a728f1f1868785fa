"""Summarises an `ncu --set full` report of conv launches (read here, without a GPU) and records the DRAM traffic of the
dominant launch for bench.py's `roofline.traffic`.

    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_conv_pair python tools/conv_bench.py ...
    python tools/ncu_traffic.py gpurun_out/r02_conv_pair.ncu-rep resnet18_b64_112 conv_stem_kernel \
        profiles/r02_ncu_conv_kernels.txt

writes the per-kernel table to the last argument and merges {key: {dram_bytes, note}} into profiles/r02_traffic.json.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main():
    rep, key, dominant, out = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    lines = [f"# {Path(rep).name}: ncu --set full --clock-control none (cold caches, serialised launches)",
             f"{'kernel':58s}" + "".join(f"{t:>10s}" for _, t in COLS)]
    traffic = None
    for r in body:
        vals = []
        for m, _ in COLS:
            i = hdr.index(m)
            v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            vals.append(v)
        lines.append(f"{r[name_i][:58]:58s}" + "".join(f"{v:10.1f}" for v in vals))
        if traffic is None and dominant in r[name_i]:
            traffic = (vals[1] + vals[2]) * 1e6
            us = vals[0]
    Path(out).write_text("\n".join(lines) + "\n")
    tfile = ROOT / "profiles" / "r02_traffic.json"
    table = json.loads(tfile.read_text()) if tfile.exists() else {}
    table[key] = {"dram_bytes": traffic,
                  "note": f"dram__bytes_read.sum + dram__bytes_write.sum of one {dominant} launch ({us:.0f} us under ncu), "
                          f"from {Path(out).name} (ncu --set full capture of this round)"}
    tfile.write_text(json.dumps(table, indent=1) + "\n")
    print("\n".join(lines))
    print(table[key])


if __name__ == "__main__":
    main()
