"""roofline_traffic.json from an ncu launch list of profiles/prof_step.py (time + DRAM bytes per launch of one warm denoise step):
DRAM bytes per launch of the tcgen05 GEMM / conv family, which bench.py quotes as `roofline.traffic`.
usage: python profiles/make_roofline_traffic.py profiles/launches_step_r02c_warm.csv > profiles/roofline_traffic.json"""
import csv
import json
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = defaultdict(dict)
    unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in csv.DictReader(lines):
        d = per[int(r["ID"])]
        d["name"] = r["Kernel Name"]
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * unit.get(r["Metric Unit"], 1)
    fam = [d for d in per.values() if "gemm_kernel" in d["name"]]
    tot_ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    fam_ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in fam)
    fam_b = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in fam)
    out = {"kernel": "gmd::gemm_kernel<...> family (tcgen05 GEMM + implicit-GEMM conv), all launches of one warm eager denoise step (SDR UNet 16 samples + GM UNet 8 samples)",
           "launches": len(fam), "dram_bytes_per_denoise_step": fam_b, "dram_bytes_per_launch": fam_b / max(len(fam), 1),
           "ncu_time_ms_per_denoise_step": fam_ms, "ncu_share_of_step": fam_ms / max(tot_ms, 1e-9),
           "source": f"{path}: ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum on profiles/prof_step.py (warm: the launch list of the step itself, no cache flush between kernels)",
           "note": "compare with roofline.algorithmic_bytes_per_launch (operands + outputs of each call once): activations written by one kernel are partly still in the 126 MB L2 when the next reads them, weights (1.7 GB per UNet) always come from HBM"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
