#!/usr/bin/env python
"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`) into
profiles/<tag>_ncu_full.md and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py's roofline.traffic).

    python tools/ncu_summary.py raw.csv r01e "command line that was profiled"
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
GB = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    raw, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full, round %s" % tag, "", "Command: `%s`" % cmd, ""]
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
        out += ["", "## %s" % name, "", "| metric | unit | value |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                out.append("| %s | %s | %s |" % (k, units[hdr.index(k)], r[hdr.index(k)]))
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[hdr.index(k)]) * GB[units[hdr.index(k)]]
        traffic.setdefault(name.split("<")[0], []).append(b)
    with open(os.path.join(ROOT, "profiles", "%s_ncu_full.md" % tag), "w") as fh:
        fh.write("\n".join(out) + "\n")
    tj = {"source": "profiles/%s_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches "
                    "of that kernel in one step)" % tag, "unit": "bytes per launch"}
    for k, v in traffic.items():
        tj[k] = sum(v) / len(v)
        tj[k + "_launches"] = v
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump(tj, fh, indent=1)
    print(json.dumps(tj))


if __name__ == "__main__":
    main()
