"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file X.csv``) per kernel and write the JSON that bench.py reads for ``roofline.traffic``.

    python tools/ncu_summary.py profiles/r2_launches.csv profiles/r2_traffic.json [steps_in_capture]
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = [l for l in open(src) if l.startswith('"')]
    rd = csv.DictReader(rows)
    per = defaultdict(lambda: defaultdict(float))
    for r in rd:
        key = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1])
        per[key][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ker = defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    for (_, name), m in per.items():
        k = ker[name]
        k["launches"] += 1
        k["time_us"] += m.get("gpu__time_duration.sum", 0.0) / 1e3
        k["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
        k["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
    tot_t = sum(k["time_us"] for k in ker.values())
    out = {"capture": src, "kernels": {}}
    for name, k in sorted(ker.items(), key=lambda kv: -kv[1]["time_us"]):
        n = k["launches"]
        out["kernels"][name] = {"launches": n, "avg_us": k["time_us"] / n, "share_of_kernel_time": k["time_us"] / tot_t,
                                "dram_bytes_per_launch": (k["dram_read"] + k["dram_write"]) / n,
                                "dram_read_per_launch": k["dram_read"] / n, "dram_write_per_launch": k["dram_write"] / n,
                                "achieved_dram_gbs": (k["dram_read"] + k["dram_write"]) / (k["time_us"] * 1e-6) / 1e9}
    g = next((v for k, v in out["kernels"].items() if "gemm_tc" in k or "mft_fused" in k), None)
    if g:
        out["dram_bytes_per_gemm_launch"] = g["dram_bytes_per_launch"]
    out["dram_bytes_total"] = sum(k["dram_read"] + k["dram_write"] for k in ker.values())
    if not steps:   # one pupil_kernel launch per PSF + gradient step
        steps = float(next((v["launches"] for k, v in out["kernels"].items() if k.startswith("pupil_kernel")), 0))
    if steps:
        out["steps_in_capture"] = steps
        out["dram_bytes_per_step"] = out["dram_bytes_total"] / steps
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
