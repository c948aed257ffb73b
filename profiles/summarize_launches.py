"""Summarise an `ncu --csv` launch list: per-kernel totals, and (when dram byte metrics were collected) bytes and GB/s per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [--detail substring]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    detail = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--detail" else None
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per_launch = defaultdict(dict)
    for r in rd:
        per_launch[int(r["ID"])]["name"] = r["Kernel Name"]
        per_launch[int(r["ID"])]["grid"] = r["Grid Size"]
        per_launch[int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in sorted(per_launch):
        d = per_launch[i]
        name = re.sub(r"^void (gmd::)?(<unnamed>::|unnamed>::)?", "", d["name"]).split("(")[0]
        t = d.get("gpu__time_duration.sum", 0.0)
        rb, wb = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        a = agg[name]
        a[0] += 1; a[1] += t; a[2] += rb; a[3] += wb
        if detail and detail in name:
            print(f"  id {i:4d} {t / 1e3:8.1f} us  grid {d['grid']:>16s}  dram rd {rb / 1e6:7.1f} MB wr {wb / 1e6:7.1f} MB  {name}")
    total = sum(a[1] for a in agg.values())
    print(f"total {total / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f"  dram {(a[2] + a[3]) / 1e9:6.2f} GB  {(a[2] + a[3]) / max(a[1], 1):6.0f} GB/s" if a[2] + a[3] > 0 else ""
        print(f"  {a[1] / 1e6:8.3f} ms {100 * a[1] / total:5.1f}%  x{a[0]:4d}  {name}{extra}")


if __name__ == "__main__":
    main()
