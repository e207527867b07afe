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
    tensor = B * L * C * 4.0            # one [B*T, C] fp32 tensor (DNA_default: T = L)
    agg = collections.defaultdict(lambda: {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "algo": 0.0})
    for (_, name), m in rows.items():
        a = agg[name]
        a["launches"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.0)
        rd = m.get("dram__bytes_read.sum", 0.0)
        a["rd"] += rd
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
        if name == "bn_col_stats_kernel":
            a["algo"] += tensor                               # reads the tensor once
        elif name == "bn_apply_kernel":                      # reads one tensor (or two: conv2c + a materialised branch1,
            a["algo"] += tensor * (2 if rd > 1.5 * tensor else 1) + tensor      # told apart by the measured reads), writes one
    out = {"B": B, "L": L, "C": C, "tensor_bytes": tensor, "kernels": {}}
    for name, a in sorted(agg.items()):
        n = a["launches"]
        us = a["ns"] / n / 1e3
        traffic = (a["rd"] + a["wr"]) / n
        k = {"launches": n, "mean_us": round(us, 2), "dram_read_per_launch": a["rd"] / n, "dram_write_per_launch": a["wr"] / n,
             "dram_GBs": round(traffic / (us * 1e-6) / 1e9, 1)}
        if a["algo"]:
            k["algorithmic_bytes_per_launch"] = a["algo"] / n
            k["algorithmic_GBs"] = round(a["algo"] / n / (us * 1e-6) / 1e9, 1)
            if peak:
                k["frac_of_peak"] = round(k["algorithmic_GBs"] / peak, 3)
        out["kernels"][name] = k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
