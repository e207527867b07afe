"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:bn_ --csv`
launch list of tools/bn_bench.py: per kernel launches, mean duration, DRAM traffic and achieved HBM GB/s, next to the
algorithmic bytes (tensor elements x 4 B declared in DESIGN.md section 5).

    python tools/ncu_bn_summary.py gpurun_out/r01_s3/bn_kernels.csv B L C [peak_GBs]"""
import collections
import csv
import json
import sys


def main():
    path, B, L, C = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    peak = float(sys.argv[5]) if len(sys.argv) > 5 else None
    rows = collections.defaultdict(dict)
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows[(int(r["ID"]), r["Kernel Name"].split("(")[0])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    agg = collections.defaultdict(lambda: {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for (_, name), m in rows.items():
        a = agg[name]
        a["launches"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.0)
        a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
    tensor = B * L * C * 4.0            # one [B*T, C] fp32 tensor (DNA_default: T = L)
    algo = {"bn_col_stats_kernel": tensor, "bn_apply_kernel": 2 * tensor}       # read | read + write (lower bound)
    out = {"B": B, "L": L, "C": C, "tensor_bytes": tensor, "kernels": {}}
    for name, a in sorted(agg.items()):
        n = a["launches"]
        us = a["ns"] / n / 1e3
        traffic = (a["rd"] + a["wr"]) / n
        k = {"launches": n, "mean_us": round(us, 2), "dram_read_per_launch": a["rd"] / n, "dram_write_per_launch": a["wr"] / n,
             "dram_GBs": round(traffic / (us * 1e-6) / 1e9, 1)}
        if name in algo:
            k["algorithmic_bytes_per_launch"] = algo[name]
            k["algorithmic_GBs"] = round(algo[name] / (us * 1e-6) / 1e9, 1)
            if peak:
                k["frac_of_peak"] = round(k["algorithmic_GBs"] / peak, 3)
        out["kernels"][name] = k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
