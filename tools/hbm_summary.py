#!/usr/bin/env python
"""ncu raw CSV (`ncu -i X.ncu-rep --page raw --csv`) -> markdown table of the HBM-bound kernels: duration, DRAM bytes
moved, achieved GB/s against the measured copy peak (MEASURED_PEAKS.json).

    python tools/hbm_summary.py raw.csv "profiled command" > profiles/r02_hbm_kernels.md
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    raw, cmd = sys.argv[1], sys.argv[2]
    peak = 6575.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        for k in ("hbm_gbs", "hbm_gbps"):
            if k in mp:
                peak = float(mp[k])
                break
    except (OSError, ValueError):
        pass
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    agg = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
        t = val(r, "gpu__time_duration.sum")
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        grid = r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "?"
        pct = r[hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")] \
            if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in hdr else "?"
        a = agg.setdefault(name, [])
        a.append((t, b, grid, pct))
    print("# HBM-bound kernels under ncu --set full (round 2)\n")
    print("Command: `%s`\n" % cmd)
    print("Peak: %.0f GB/s (MEASURED_PEAKS.json copy bandwidth; durations under ncu are cold-cache, serialised "
          "launches, so the GB/s here is a lower bound of what the kernel reaches inside a pipeline).\n" % peak)
    print("| kernel | launches | largest launch: grid | duration us | DRAM read+write MB | achieved GB/s | frac of peak | "
          "ncu dram throughput % |")
    print("|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items()):
        t, b, grid, pct = max(a, key=lambda x: x[1])
        gbps = b / t / 1e9 if t > 0 else 0.0
        print("| %s | %d | %s | %.1f | %.1f | %.0f | %.2f | %s |" % (name, len(a), grid, t * 1e6, b / 1e6, gbps, gbps / peak, pct))


if __name__ == "__main__":
    main()
