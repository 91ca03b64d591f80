#!/usr/bin/env python
"""bench.py - pileup positions/sec through the consensus-inference hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Workload (config.workload): BASELINE.json configs[1] - r1041_e82_400bps_sup_v5 consensus on a
synthetic 10 Mb draft: 1111 windows x 10000 pileup columns x 10 features (chunk_len 10000,
overlap 1000; the reference's six 200-window batches coalesced into ONE device batch - the
engine takes any batch size and a B200 holds the whole draft).  One step = one pass of the
hot path over that batch (features in -> probabilities + labels out).  Weights are seeded
synthetic (random-init, the archives in the reference are Git-LFS stubs).

`value`   : positions/s with inputs resident in HBM when the timed region starts
            (mdk_engine_forward_dev), K steps bracketed by CUDA events on the engine stream.
`e2e`     : the same metric through the reference-facing call with HOST buffers
            (mdk_engine_forward: pinned H2D of the features, D2H of probabilities + labels
            inside the timed region).
`roofline`: the dominant kernel, tensor-core bound: algorithmic GRU-gate FLOPs of that kernel per
            launch / its mean launch duration (CUDA events per stage, recorded every step).
N > 1: one process per GPU (torchrun), weights broadcast once over NCCL from rank 0, each rank
runs the same per-GPU workload on its own windows (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOWS, COLS, FEATS = 1111, 10000, 10
# algorithmic FLOPs per position: H=128, F=10, 2 layers, bidirectional
FLOP_REC_PER_LAYER = 2 * (2 * 384 * 128)          # 196 608  (both directions, one layer)
FLOP_INPROJ1 = 2 * (2 * 384 * 256)                # 393 216
FLOP_INPROJ0 = 2 * (2 * 384 * 10)                 # 15 360
FLOP_GRU_TOTAL = 2 * FLOP_REC_PER_LAYER + FLOP_INPROJ1 + FLOP_INPROJ0   # 801 792


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    """Host threads this process may really use: scheduler affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


# stdout carries exactly ONE line (the JSON): libraries that chat on fd 1 (NCCL prints its version there under torchrun)
# are sent to stderr for the whole run, and the line is written to the saved descriptor at the end
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(msg):
    print("[bench %.1fs] %s" % (time.perf_counter() - T_START, msg), file=sys.stderr, flush=True)


T_START = time.perf_counter()


def fill_features(x, seed):
    """Fill a float32 [B,T,F] array IN PLACE with normalised-count-like values (rows sum to 1)."""
    rng = np.random.default_rng(seed)
    rng.random(out=x.reshape(-1), dtype=np.float32)
    step = max(1, (1 << 22) // (x.shape[1] * x.shape[2]))
    for i in range(0, x.shape[0], step):
        b = x[i:i + step]
        b *= b * b
        b /= b.sum(axis=-1, keepdims=True)
    return x


def cpu_reference_rate(threads, sample_windows, cols, feats, steps=1, warmup=0, budget_s=12.0):
    """positions/s of the reference's CPU arithmetic (torch fp32 nn.GRU + Linear + softmax, the
    oracle restatement of medaka/architectures/gru.py + models.py:303-313) on a bounded sample.

    ``cols`` <= 0 picks the number of columns so that one pass takes about ``budget_s`` seconds
    (calibrated on a 20-column probe); per-position cost does not depend on the window length.
    Returns (positions/s, seconds per step, cols used)."""
    import torch
    from oracle import gru_oracle, synth
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict(0, num_features=feats)
    model = gru_oracle.build(sd, num_features=feats)
    if cols <= 0:
        probe = fill_features(np.empty((sample_windows, 20, feats), dtype=np.float32), 2)
        gru_oracle.predict_on_batch(model, probe)
        t0 = time.perf_counter()
        gru_oracle.predict_on_batch(model, probe)
        r0 = sample_windows * 20 / (time.perf_counter() - t0)
        cols = int(min(2000, max(40, r0 * budget_s / sample_windows)))
    x = fill_features(np.empty((sample_windows, cols, feats), dtype=np.float32), 1)
    for _ in range(warmup):
        gru_oracle.predict_on_batch(model, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        gru_oracle.predict_on_batch(model, x)
    dt = time.perf_counter() - t0
    return steps * sample_windows * cols / dt, dt / steps, cols


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = host_cores()
    sample_windows = args.cpu_windows
    rate, sec_per_step, cols = cpu_reference_rate(cores, sample_windows, args.cpu_cols, FEATS, steps=args.steps,
                                                  warmup=min(args.warmup, 1))
    sample = "%d windows x %d cols per step (the reference's 200-window batch, truncated in time)" % (
        sample_windows, cols)
    line = {
        "impl": "reference", "metric": "pileup positions/sec (consensus inference)", "value": rate,
        "unit": "positions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "timing": "host wall clock, CPU only", "threads": cores},
        "cpu_baseline": {"value": rate, "unit": "positions/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name():
    return ("r1041_e82_400bps_sup_v5 consensus, synthetic 10 Mb draft: %d windows x %d cols x %d feats "
            "(chunk_len 10000, overlap 1000), six 200-window batches coalesced into one device batch"
            % (WINDOWS, COLS, FEATS))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--windows", type=int, default=WINDOWS, help="windows per step per GPU")
    ap.add_argument("--cols", type=int, default=COLS)
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--cpu-windows", type=int, default=200, help="windows in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-cols", type=int, default=0,
                    help="columns per window in the CPU-baseline sample (0 = sized for ~12 s per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from medaka_b200 import libmedaka as lm
    from medaka_b200 import models
    from oracle import synth   # seeded synthetic weights/inputs + the cpu_baseline leg only

    lib = lm.load()
    ffi = lm.ffi
    dev = local_rank if world > 1 else 0
    info = lm.require_gpu(dev)

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    # ---- weights: rank 0 owns them, one NCCL broadcast of the packed fp32 blob (1.62 MB) ----
    sd = synth.synth_state_dict(0, num_features=FEATS)
    keys = sorted(sd)
    if world > 1:
        blob = np.concatenate([sd[k].ravel() for k in keys])
        t = torch.from_numpy(blob if rank == 0 else np.zeros_like(blob)).cuda()
        dist.broadcast(t, src=0)
        flat = t.cpu().numpy()
        off = 0
        for k in keys:
            n = sd[k].size
            sd[k] = flat[off:off + n].reshape(sd[k].shape).copy()
            off += n
    model = models.GRUModel(num_features=FEATS, device=dev)
    model.load_state_dict(sd)
    model.set_precision(args.precision)
    eng = model.engine

    B, T, F = args.windows, args.cols, FEATS
    P = B * T
    log("generating %d x %d x %d synthetic features into pinned host memory" % (B, T, F))
    feats = fill_features(model.pinned("bench_feats", (B, T, F), np.float32), 1000 + rank)
    log("reserving workspace")
    lm.check(lib.mdk_engine_reserve(eng, B, T))

    # ---- device-resident leg ("value") ----
    def dalloc(nbytes):
        pp = ffi.new("void **")
        lm.check(lib.mdk_dev_alloc(dev, nbytes, pp))
        return pp[0]

    d_feats = dalloc(feats.nbytes)
    d_probs = dalloc(P * 5 * 4)
    d_labels = dalloc(P)
    lm.check(lib.mdk_memcpy_h2d(dev, d_feats, ffi.from_buffer(feats), feats.nbytes))

    def step_dev():
        lm.check(lib.mdk_engine_forward_dev(eng, ffi.cast("const float *", d_feats), B, T,
                                            ffi.cast("float *", d_probs), ffi.NULL,
                                            ffi.cast("uint8_t *", d_labels)))

    def barrier():
        if dist is not None:
            dist.barrier()
        lm.check(lib.mdk_engine_sync(eng))

    log("device-resident leg: warm-up")
    for _ in range(args.warmup):
        step_dev()
    barrier()
    log("device-resident leg: timing %d steps" % args.steps)
    launches0 = model.launch_count()
    sampler = ClockSampler(dev)
    sampler.start()
    ms = ffi.new("float *")
    lm.check(lib.mdk_engine_timer_start(eng))
    for _ in range(args.steps):
        step_dev()
    lm.check(lib.mdk_engine_timer_stop(eng, ms))     # records + synchronises the end event
    barrier()
    clocks = sampler.stop()
    dev_ms = float(ms[0])
    launches = model.launch_count() - launches0
    tm = ffi.new("mdk_timings *")
    lm.check(lib.mdk_engine_mean_timings(eng, min(args.steps, 32), tm))
    stage = {k: float(getattr(tm, k)) for k in ("inproj0_ms", "rec0_ms", "inproj1_ms", "rec1_ms", "head_ms")}

    # sanity: the timed path produced real outputs (labels consistent with probabilities)
    chk = np.empty((min(B, 4), T, 5), dtype=np.float32)
    lm.check(lib.mdk_memcpy_d2h(dev, ffi.from_buffer(chk), d_probs, chk.nbytes))
    assert np.isfinite(chk).all() and abs(float(chk.sum(-1).mean()) - 1.0) < 1e-4

    # ---- host-buffer leg ("e2e"): pinned H2D + forward + D2H of probs and labels EVERY step, through the
    # reference-facing C-ABI call with host buffers (mdk_engine_submit / mdk_engine_wait, the asynchronous form of
    # mdk_engine_forward that medaka_b200.prediction.run_prediction uses: two calls in flight, so the copies of
    # neighbouring steps hide under the compute of the current one) ----
    h_feats = feats
    h_probs = [model.pinned("bench_probs%d" % i, (B, T, 5), np.float32) for i in range(2)]
    h_labels = [model.pinned("bench_labels%d" % i, (B, T), np.uint8) for i in range(2)]

    def run_host(n):
        tickets = []
        for k in range(n):
            tickets.append(model.submit_arrays(h_feats, h_probs[k % 2], h_labels[k % 2]))
            if k >= 1:
                model.wait(tickets[k - 1])
        model.wait(tickets[-1])

    log("host-buffer leg")
    run_host(max(1, min(args.warmup, 2)))
    barrier()
    lm.check(lib.mdk_engine_timer_start(eng))
    run_host(args.steps)
    lm.check(lib.mdk_engine_timer_stop(eng, ms))
    barrier()
    e2e_ms = float(ms[0])
    assert np.isfinite(h_probs[(args.steps - 1) % 2][:2]).all()

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks, device-timed
        dev_ms, e2e_ms = float(t[0]), float(t[1])

    total_positions = world * args.steps * P
    value = total_positions / (dev_ms * 1e-3)
    e2e = total_positions / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel ----
    peaks = measured_peaks()
    kernels = {
        "rec_tc_kernel (GRU recurrence, layer 0 and layer 1 launches)":
            (0.5 * (stage["rec0_ms"] + stage["rec1_ms"]), P * FLOP_REC_PER_LAYER,
             stage["rec0_ms"] + stage["rec1_ms"]),
        "gemm_tc_kernel (layer-1 input projection)": (stage["inproj1_ms"], P * FLOP_INPROJ1, stage["inproj1_ms"]),
    }
    dom = max(kernels, key=lambda k: kernels[k][2])
    k_ms, k_flop, k_share_ms = kernels[dom]
    achieved = k_flop / (k_ms * 1e-3) / 1e12
    # DRAM traffic per launch of that kernel: from the committed ncu --set full capture of this same workload
    # (profiles/ncu_traffic.json, written by tools/ncu_summary.py); null for any other shape
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if (B, T) == (WINDOWS, COLS) and args.precision == "tc" and os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic = tj.get("rec_tc_kernel" if dom.startswith("rec_tc") else "gemm_tc_kernel")
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tflops_sustained"],
        "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
        "peak_source": "MEASURED_PEAKS.json bf16 sustained (kernel timed inside a long step)"
        if peaks["source"] == "measured" else "fallback (B200_PROFILING.md)",
        "note": "algorithmic FLOPs; operands are fp16 hi/lo pairs so the kernel issues 3 MMAs per product "
                "(fp32-faithful parity), i.e. executed tensor FLOPs are 3x this figure",
        "kernel_share_of_step": k_share_ms / max(sum(stage.values()), 1e-9),
        "whole_pipeline_achieved": value * FLOP_GRU_TOTAL / 1e12,
        "whole_pipeline_frac": value * FLOP_GRU_TOTAL / 1e12 / peaks["tflops_sustained"] / world,
        "stage_ms": stage,
    }

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = host_cores()
        log("cpu baseline on %d threads" % cores)
        rate, sec, ccols = cpu_reference_rate(cores, args.cpu_windows, args.cpu_cols, F, steps=1, warmup=0)
        cpu_baseline = {"value": rate, "unit": "positions/s", "cores": cores, "kind": "port",
                        "sample": "%d windows x %d cols, 1 pass (%.1f s), torch %s fp32 nn.GRU oracle" % (
                            args.cpu_windows, ccols, sec, torch.__version__)}
    log("done")

    line = {
        "metric": "pileup positions/sec (consensus inference)", "value": value, "unit": "positions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 gate math; fp16 hi/lo split tensor-core operands, fp32 accumulate"
        if args.precision == "tc" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name() if (B, T) == (WINDOWS, COLS) else
                   "synthetic %d windows x %d cols x %d feats" % (B, T, F),
                   "windows_per_gpu": B, "cols": T, "precision": args.precision,
                   "l2_policy": "inputs larger than L2 (444 MB of features, >3 GB of activations per step)",
                   "sm_count": info["sm_count"]},
        "e2e": {"value": e2e, "unit": "positions/s", "h2d_bytes_per_step": int(feats.nbytes),
                "d2h_bytes_per_step": int(P * 5 * 4 + P), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
