#!/usr/bin/env python
"""Summarise an .ncu-rep (brought back under gpurun_out/) into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_prof.md
Reads the report HERE (no GPU needed) through `ncu -i ... --page raw/source --csv`.
"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
       "smsp__thread_inst_executed_per_inst_executed.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warp" in h and "per_issue_active" in h and "not_issued" not in h]
    md = [f"# ncu summary of `{rep}`", "", "`ncu --set full --clock-control none --import-source on` (cold caches, serialised replays:",
          "compare shares and ratios, not absolute times with the bench).", ""]
    for r in rows[2:]:
        md.append(f"## {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for m in RAW:
            if m in idx:
                md.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
        vals = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
        md.append("")
        md.append("Top warp-stall reasons (cycles per issued instruction): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
            for v, h in vals))
        md.append("")
    open(out, "w").write("\n".join(md) + "\n")
    print("wrote", out)
    if len(sys.argv) > 3:
        # per-launch DRAM traffic by kernel (bench.py's roofline.traffic reads this file)
        import json
        import re
        traffic = {}
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = re.sub(r"<.*", "", r[idx["Kernel Name"]].split("::")[-1].split("(")[0]).strip()
            b = sum(float(r[idx[m]]) * unit.get(units[idx[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic.setdefault(name, []).append({"grid": r[idx["Grid Size"]], "dram_bytes": b,
                                                 "time_us": float(r[idx["gpu__time_duration.sum"]])})
        json.dump({"source": rep, "how": "ncu --set full --clock-control none, one frame of configs[1]", "kernels": traffic},
                  open(sys.argv[3], "w"), indent=1)
        print("wrote", sys.argv[3])


if __name__ == "__main__":
    main()
