"""profiles/r0N_traffic.json from the `ncu --set full` captures of one evidence pass (run here: ncu is installed, no GPU needed).

    python tools/ncu_traffic.py gpurun_out/<tag> [precision] [batch]

Per kernel category: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured
launches), the ncu durations and pipe utilisations.  bench.py copies `dram_bytes_per_launch` of the dominant category into
roofline.traffic."""
import csv, json, subprocess, sys, os

def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]

def f(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return None

def main():
    tag = sys.argv[1]
    prec = sys.argv[2] if len(sys.argv) > 2 else "tc"
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    cats = {}
    for rep, name in (("prof_gemm.ncu-rep", None), ("prof_lstm.ncu-rep", "lstm_rec")):
        path = os.path.join(tag, rep)
        if not os.path.exists(path):
            continue
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        u = dict(zip(hdr, units))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            cat = name
            if cat is None:   # gemm_tc_kernel: the resident-weight launches (N = 8H) are the LSTM input projection
                kname = d.get("Kernel Name", "")      # the CTA-pair kernel = convolutions, the single-CTA kernel = input projection
                cat = "conv" if "pair" in kname else "lstm_in"
            rd = f(d, "dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]]
            wr = f(d, "dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
            c = cats.setdefault(cat, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ms": 0.0, "tensor_pct": 0.0,
                                      "issue_pct": 0.0, "dram_pct": 0.0})
            c["launches"] += 1
            c["dram_read"] += rd; c["dram_write"] += wr
            c["ms"] += f(d, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0)
            c["tensor_pct"] += f(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0.0
            c["issue_pct"] += f(d, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0
            c["dram_pct"] += f(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
    per = {}
    for k, c in cats.items():
        n = c["launches"]
        per[k] = {"captured_launches": n, "dram_bytes_per_launch": (c["dram_read"] + c["dram_write"]) / n,
                  "dram_read_per_launch": c["dram_read"] / n, "dram_write_per_launch": c["dram_write"] / n,
                  "ncu_ms_per_launch": c["ms"] / n, "tensor_pipe_active_pct": c["tensor_pct"] / n,
                  "issue_active_pct": c["issue_pct"] / n, "dram_throughput_pct": c["dram_pct"] / n}
    json.dump({"precision": prec, "batch": batch, "source": "ncu --set full --clock-control none, %s" % tag,
               "per_category": per}, sys.stdout, indent=1)
    print()

main()
