#!/usr/bin/env python
"""Turn `ncu --set full` reports (gpurun_out/<tag>_full_<workload>.ncu-rep) into the small CSV summaries committed under
profiles/ (one row per captured launch, the metrics the roofline discussion in DESIGN.md uses) and refresh
profiles/traffic.json (DRAM bytes per launch of each workload's dominant kernel).

  python profiles/summarise.py r01b            # reads gpurun_out/r01b_full_*.ncu-rep, writes profiles/r01b_ncu_full_*.csv
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__cycles_active.avg",
]
# which launch rows matter for traffic.json: (workload, kernel substring, pick) -- pick = "max" takes the longest launch
DOMINANT = {"knn_cosine_1q": "pdx_scan_kernel", "hamming": "hamming_kernel", "u8": "u8_scan_kernel",
            "maxsim": "maxsim_tc_kernel", "knn_cosine_multi": "knn_tc_filter_kernel", "batch_demo": "pdx_scan_kernel"}
ALGORITHMIC = {"knn_cosine_1q": 30_720_000_000, "hamming": 12_800_000_000, "u8": 19_200_000_000,
               "maxsim": 92_160_000_000, "knn_cosine_multi": 10_000_000 * 768 * 2,
               "batch_demo": 10_000 * 128 * 4 * 25}  # 25 groups of 4 queries each read the (L2-resident) corpus once
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(rep):
    if rep.endswith(".csv"):  # already exported on the GPU box (profiles/capture.sh)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    return rd[0], rd[1], rd[2:]  # header, units, launches


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for w, kern in DOMINANT.items():
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full_{w}.ncu-rep")
        if not os.path.exists(rep):
            rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full_{w}.raw.csv")
        if not os.path.exists(rep):
            continue
        hdr, units, rows = raw_rows(rep)
        col = {h: i for i, h in enumerate(hdr)}
        keep = [m for m in METRICS if m in col]
        dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_{w}.csv")
        with open(dst, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["ID", "Kernel Name"] + keep)
            wr.writerow(["", ""] + [units[col[m]] for m in keep])
            for r in rows:
                wr.writerow([r[col["ID"]], r[col["Kernel Name"]]] + [r[col[m]] for m in keep])
        # dominant launch = the longest one of the matching kernel
        best = None
        for r in rows:
            if kern not in r[col["Kernel Name"]]:
                continue
            t = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            if best is None or t > best[0]:
                rb = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT_SCALE[units[col["dram__bytes_read.sum"]]]
                wb = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT_SCALE[units[col["dram__bytes_write.sum"]]]
                best = (t, rb + wb, r[col["Kernel Name"]], units[col["gpu__time_duration.sum"]])
        if best:
            traffic[w] = {"kernel": best[2], "bytes": int(best[1]), "algorithmic": ALGORITHMIC[w],
                          "ncu_duration": f"{best[0]} {best[3]}", "capture": os.path.basename(dst)}
            print(w, traffic[w])
    traffic["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (the longest captured "
                           "launch), from ncu --set full at scale 1.0 on 1 GPU (profiles/<tag>_ncu_full_*.csv; produced by "
                           "profiles/capture.sh + profiles/summarise.py). bench.py reports these as roofline.traffic.")
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
