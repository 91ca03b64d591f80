#!/usr/bin/env python
"""Read-level network (LatentSpaceLSTM forward) rate on the GPU against the torch CPU restatement on the same tensor.

    python tools/rl_bench.py [--windows 16] [--positions 1000] [--reads 50] [--cpu-windows 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=16)
    ap.add_argument("--positions", type=int, default=1000)
    ap.add_argument("--reads", type=int, default=50)
    ap.add_argument("--cpu-windows", type=int, default=2, help="0 = skip the CPU arm")
    ap.add_argument("--tc-only", action="store_true", help="skip the fp32-convolution pass (for profiler runs)")
    args = ap.parse_args()
    from medaka_b200 import read_level
    from oracle import rl_oracle
    sd = rl_oracle.synth_rl_state_dict(0)
    x = rl_oracle.synth_rl_features(args.windows, args.positions, args.reads, seed=1, empty_rows=0, ragged=False)
    m = read_level.LatentSpaceLSTM()
    m.load_state_dict(sd)
    m.forward_arrays(x[:2])
    def best():
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = m.forward_arrays(x)
            ts.append(time.perf_counter() - t0)
        return min(ts), out
    if args.tc_only:
        t_fp32, probs_fp32 = float("nan"), None
    else:
        m.set_conv(False)
        t_fp32, probs_fp32 = best()
    m.set_conv(True)
    m.forward_arrays(x[:2])
    t, probs = best()
    cells = args.windows * args.positions * args.reads
    flop_conv = cells * 2.0 * 17 * 128 * 128
    res = {"windows": args.windows, "positions": args.positions, "reads": args.reads, "gpu_call_ms": t * 1e3,
           "positions_per_s": args.windows * args.positions / t, "read_cells_per_s": cells / t,
           "conv17_TFLOPs_algorithmic_whole_call": flop_conv / t / 1e12, "gpu_call_ms_fp32_conv": t_fp32 * 1e3,
           "max_abs_prob_diff_tc_vs_fp32_conv": None if probs_fp32 is None else float(np.abs(probs - probs_fp32).max())}
    if args.cpu_windows <= 0:
        print(json.dumps(res))
        return
    import torch
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    threads = min(threads, 16)
    xc = x[:args.cpu_windows]
    model = rl_oracle.build(sd)
    rl_oracle.predict(model, xc[:1], threads=threads)
    t0 = time.perf_counter()
    want = rl_oracle.predict(model, xc, threads=threads)
    tc = time.perf_counter() - t0
    res["cpu"] = {"windows": args.cpu_windows, "threads": threads, "seconds": tc,
                  "positions_per_s": args.cpu_windows * args.positions / tc, "torch": torch.__version__}
    res["max_abs_prob_diff_vs_cpu"] = float(np.abs(probs[:args.cpu_windows] - want).max())
    res["gpu_over_cpu"] = res["positions_per_s"] / res["cpu"]["positions_per_s"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
