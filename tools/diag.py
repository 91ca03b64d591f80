#!/usr/bin/env python
"""GPU bring-up diagnostics: each check runs in its own process (a trapped kernel poisons the CUDA
context) under a timeout, and reports into gpurun_out/diag.json.  Not part of the product."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def check_selftest(arg):
    from medaka_b200 import libmedaka as lm
    lm.load()
    ffi = lm.ffi
    res = {}
    variant = int(arg)
    for N, K in [(16, 16), (16, 128), (32, 64), (128, 128), (64, 256)]:
        rs = np.random.RandomState(N * 1000 + K)
        A = rs.uniform(-1, 1, (128, K)).astype(np.float32)
        B = rs.uniform(-1, 1, (N, K)).astype(np.float32)
        D = np.full((128, N), np.nan, dtype=np.float32)
        lm.check(lm.lib.mdk_selftest_umma(0, ffi.cast("const float *", ffi.from_buffer(A)),
                                          ffi.cast("const float *", ffi.from_buffer(B)),
                                          ffi.cast("float *", ffi.from_buffer(D)), N, K, variant))
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        res["N%d_K%d" % (N, K)] = float(np.nanmax(np.abs(D - ref))) if np.isfinite(D).any() else "all-nan"
    return res


def _forward(precision, B, T, seed=0, F=10):
    from medaka_b200 import models
    from oracle import gru_oracle, synth
    sd = synth.synth_state_dict(seed, num_features=F)
    feats = synth.synth_features(B, T, F, seed=42)
    man = gru_oracle.manual_forward(sd, feats)
    m = models.GRUModel(num_features=F)
    m.load_state_dict(sd)
    m.set_precision(precision)
    m.keep_activations(True)
    t0 = time.time()
    out = m.forward_arrays(feats, want_logits=True)
    dt = time.time() - t0
    h0, h1 = m.read_activation(0), m.read_activation(1)
    scale = np.abs(man["logits"]).max(-1, keepdims=True)
    res = {
        "dh0": float(np.abs(h0 - man["h0"]).max()), "dh1": float(np.abs(h1 - man["h1"]).max()),
        "dh0_fwd": float(np.abs(h0[..., :128] - man["h0"][..., :128]).max()),
        "dh0_rev": float(np.abs(h0[..., 128:] - man["h0"][..., 128:]).max()),
        "dh0_t0": float(np.abs(h0[:, 0, :128] - man["h0"][:, 0, :128]).max()),
        "dh0_t1": float(np.abs(h0[:, min(1, T - 1), :128] - man["h0"][:, min(1, T - 1), :128]).max()),
        "dlogit_scaled": float((np.abs(out.logits - man["logits"]) / scale).max()),
        "label_mismatch": int((out.labels != np.argmax(man["probs"], -1)).sum()),
        "n": int(out.labels.size), "wall_s": dt, "timings": m.last_timings(),
        "finite": bool(np.isfinite(out.logits).all()),
    }
    return res


def check_forward(arg):
    precision, B, T = arg.split(",")
    return _forward(precision, int(B), int(T))


def check_misc(arg):
    from medaka_b200 import features, labels, common
    from oracle import features_oracle, labels_oracle, synth
    res = {}
    counts, pos = synth.synth_counts(200000, seed=3)
    for norm in ("total", "fwd_rev", None):
        for sym in (False, True):
            ef, ed = features_oracle.post_process_pileup(counts.copy(), pos, norm, ("",), sym)
            s = features.CountsFeatureEncoder(normalise=norm, sym_indels=sym)._post_process_pileup(
                counts.copy(), pos, common.Region("r", 0, int(pos["major"][-1]) + 1))
            res["norm_%s_%d" % (norm, sym)] = [int((s.features != ef).sum()), int((np.asarray(s.depth) != ed.astype(np.int64)).sum())]
    rs = np.random.RandomState(1)
    lg = rs.normal(0, 5, (500000, 5)).astype(np.float32)
    e = np.exp(lg - lg.max(-1, keepdims=True))
    p = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    lab, q = labels.decode_arrays(p)
    el, eq = labels_oracle.decode_arrays(p)
    res["decode"] = [int((lab != el).sum()), int((q != eq).sum())]
    return res


def check_determinism(arg):
    """Same input through fresh engines / repeated calls must be bit-identical; locate oracle deviations."""
    from medaka_b200 import models
    from oracle import gru_oracle, synth
    precision, B, T = arg.split(",")
    B, T = int(B), int(T)
    sd = synth.synth_state_dict(0)
    feats = synth.synth_features(B, T, 10, seed=42)
    man = gru_oracle.manual_forward(sd, feats)
    res = {"runs": []}
    first = None
    for rep in range(4):
        m = models.GRUModel(num_features=10)
        m.load_state_dict(sd)
        m.set_precision(precision)
        m.keep_activations(True)
        for call in range(2):
            out = m.forward_arrays(feats, want_logits=True)
            h0 = m.read_activation(0)
            d = np.abs(h0 - man["h0"])
            idx = np.unravel_index(np.argmax(d), d.shape)
            entry = {"rep": rep, "call": call, "dh0": float(d.max()), "argmax": [int(i) for i in idx],
                     "n_bad": int((d > 1e-6).sum()),
                     "bad_t0_fwd": int((d[:, 0, :128] > 1e-6).sum()), "bad_rows": sorted(set(np.where(d > 1e-6)[0].tolist()))[:10],
                     "bad_units": sorted(set(np.where(d > 1e-6)[2].tolist()))[:16]}
            if first is None:
                first = h0.copy()
            entry["same_as_first"] = bool(np.array_equal(first, h0))
            res["runs"].append(entry)
        m.close()
    man2 = gru_oracle.manual_forward(sd, feats)
    res["oracle_stable"] = bool(np.array_equal(man["h0"], man2["h0"]))
    res["oracle_drift"] = float(np.abs(man["h0"] - man2["h0"]).max())
    import torch
    res["torch_threads"] = torch.get_num_threads()
    return res


TRACE_SLOTS = ["issuer: h_ready seen", "issuer: r MMAs queued + commit", "issuer: z MMAs queued + commit",
               "issuer: n MMAs queued + commit", "gate: r tcgen05.ld done", "gate: sigmoid(r) done",
               "gate: z tcgen05.ld done", "gate: sigmoid(z) done", "gate: n tcgen05.ld done", "gate: h tile written",
               "gate: proxy fence done", "gate: arrived on h_ready", "gate: end of step (global stores queued)",
               "relay: r commit seen, arrived", "relay: z commit seen, arrived", "relay: n commit seen, arrived"]


def check_rec_trace(arg):
    """Cycle stamps of the recurrent kernel's hand-off points on CTA (0,0), steps 512..527 (NT = 1 shapes only)."""
    from medaka_b200 import libmedaka as lm, models
    from oracle import synth
    B, T = (int(x) for x in arg.split(","))
    lib, ffi = lm.load(), lm.ffi
    m = models.GRUModel()
    m.load_state_dict(synth.synth_state_dict(0))
    feats = synth.synth_features_fast(B, T, 10, seed=3)
    m.forward_arrays(feats)                                        # warm-up, untraced
    base = m.last_timings()
    lm.check(lib.mdk_debug_rec_trace(0, 1, ffi.NULL))
    m.forward_arrays(feats)
    traced = m.last_timings()
    buf = np.zeros((2, 16, 40), dtype=np.uint64)
    lm.check(lib.mdk_debug_rec_trace(0, 0, ffi.cast("uint64_t *", ffi.from_buffer(buf))))
    res = {"untraced_ms": base, "traced_ms": traced}
    for layer in (0, 1):
        t = buf[layer].astype(np.int64)
        # every stamp relative to the moment the r issuer saw h_ready of the same step
        rel = t[:, :16] - t[:, :1]
        period = np.diff(t[:, 0])
        res["layer%d" % layer] = {
            "step_cycles_median": float(np.median(period)),
            "median_offset_from_h_ready_seen": {TRACE_SLOTS[k]: float(np.median(rel[:, k])) for k in range(16)},
            "raw_first_step": [int(x) for x in rel[1]],
            "gate_warp_arrivals_median": [float(x) for x in np.median(t[:, 16:32] - t[:, :1], axis=0)],
            "aux_warp (staging issued, prefetch issued, tile copy read, gi landed + arrived)": [float(x) for x in np.median(t[:, 32:36] - t[:, :1], axis=0)],
        }
    return res


PP_SLOTS = ["issuer: H seen", "issuer: r queued", "issuer: z queued", "issuer: n queued", "relay: r arrived",
            "relay: z arrived", "relay: n arrived", "gate: r ld done", "gate: r math done", "gate: z ld done",
            "gate: z math done", "gate: n ld done", "gate: h written", "gate: arrived H", "issuer: logits queued",
            "issuer: r guard passed", "issuer: z guard passed", "issuer: n guard passed", "issuer: logits guard passed",
            "issuer: r commit seen by issuer (L0 only)"]


def check_pp(arg):
    """Ping-pong recurrent kernel (gru_pp.cu) against the oracle: 'B,T,F[,mode]'."""
    from medaka_b200 import models
    from oracle import gru_oracle, synth
    parts = arg.split(",")
    B, T, F = int(parts[0]), int(parts[1]), int(parts[2])
    mode = parts[3] if len(parts) > 3 else "pp"
    sd = synth.synth_state_dict(7, num_features=F)
    feats = synth.synth_features(B, T, F, seed=B * 3 + T)
    ref_probs, ref_logits = gru_oracle.predict_on_batch(gru_oracle.build(sd, num_features=F), feats)
    m = models.GRUModel(num_features=F)
    m.load_state_dict(sd)
    m.set_rec_mode(mode)
    out = m.forward_arrays(feats, want_logits=True)
    res = {"finite": bool(np.isfinite(out.logits).all())}
    scale = np.abs(ref_logits).max(-1, keepdims=True)
    err = np.abs(out.logits - ref_logits) / scale
    res["dlogit_scaled"] = float(np.nanmax(err))
    res["label_mismatch"] = int((out.labels != np.argmax(ref_probs, -1)).sum())
    res["n"] = int(out.labels.size)
    bad = np.argwhere(~(err.max(-1) < 1e-3))
    res["n_bad_pos"] = int(len(bad))
    res["bad_windows"] = sorted(set(bad[:, 0].tolist()))[:20]
    res["bad_t_first"] = sorted(set(bad[:, 1].tolist()))[:10]
    if T * B <= 400000:
        man = gru_oracle.manual_forward(sd, feats)
        h0 = m.read_activation(0)
        d = np.abs(h0 - man["h0"])
        res["dh0_fwd"] = float(d[..., :128].max())
        res["dh0_rev"] = float(d[..., 128:].max())
        res["dh0_bad_windows"] = sorted(set(np.argwhere(d.max(-1) > 1e-4)[:, 0].tolist()))[:20]
    if T * B <= 400000:
        # per-direction partial logits against W_lin[:, half] . h1[half] of the loop-level oracle
        from medaka_b200 import libmedaka as lm
        tiles = (B + 15) // 16
        plog = np.zeros((2, tiles, T, 5, 16), dtype=np.float32)
        lm.check(lm.lib.mdk_debug_read_plog(m.engine, lm.ffi.cast("float *", lm.ffi.from_buffer(plog)), plog.size))
        W = sd["linear.weight"].astype(np.float64)
        for d, name in ((0, "fwd"), (1, "rev")):
            exp = man["h1"][..., d * 128:(d + 1) * 128].astype(np.float64) @ W[:, d * 128:(d + 1) * 128].T   # [B,T,5]
            got = np.zeros((tiles * 16, T, 5), dtype=np.float32)
            got[:] = plog[d].transpose(0, 3, 1, 2).reshape(tiles * 16, T, 5)
            e = np.abs(got[:B] - exp)
            bad = np.argwhere(e.max(-1) > 1e-3)
            res["plog_%s_maxerr" % name] = float(e.max())
            res["plog_%s_bad_tiles" % name] = sorted(set((bad[:, 0] // 16).tolist()))
            res["plog_%s_bad_t" % name] = sorted(set(bad[:, 1].tolist()))[:12]
            if len(bad):
                w, t = bad[0]
                # does the bad row match the expected row of a neighbouring time step?
                cands = {dt: float(np.abs(got[w, t] - exp[w, t + dt]).max()) for dt in (-2, -1, 1, 2) if 0 <= t + dt < T}
                res["plog_%s_neighbour_match" % name] = cands
    res["timings"] = m.last_timings()
    m.close()
    return res


def check_rec_timing(arg):
    """'mode,B,T[,reps]' -> mean stage times of forward_dev-style forwards (host buffers, one at a time)."""
    from medaka_b200 import models
    from oracle import synth
    parts = arg.split(",")
    mode, B, T = parts[0], int(parts[1]), int(parts[2])
    reps = int(parts[3]) if len(parts) > 3 else 3
    m = models.GRUModel()
    m.load_state_dict(synth.synth_state_dict(0))
    m.set_rec_mode(mode)
    feats = synth.synth_features_fast(B, T, 10, seed=3)
    m.forward_arrays(feats)
    ts = []
    for _ in range(reps):
        m.forward_arrays(feats)
        ts.append(m.last_timings())
    m.close()
    return {k: float(np.mean([t[k] for t in ts])) for k in ts[0]}


def check_pp_trace(arg):
    """Cycle stamps of the ping-pong kernel's hand-off points on CTA (0,0), steps 512..527: 'B,T'."""
    from medaka_b200 import libmedaka as lm, models
    from oracle import synth
    parts = [int(x) for x in arg.split(",")]
    B, T = parts[0], parts[1]
    lib, ffi = lm.load(), lm.ffi
    if len(parts) > 2:
        lm.check(lib.mdk_debug_pp_flags(parts[2]))
    m = models.GRUModel()
    m.load_state_dict(synth.synth_state_dict(0))
    m.set_rec_mode("pp")
    feats = synth.synth_features_fast(B, T, 10, seed=3)
    m.forward_arrays(feats)
    base = m.last_timings()
    lm.check(lib.mdk_debug_rec_trace(0, 1, ffi.NULL))
    m.forward_arrays(feats)
    traced = m.last_timings()
    buf = np.zeros((2, 16, 40), dtype=np.uint64)
    lm.check(lib.mdk_debug_rec_trace(0, 0, ffi.cast("uint64_t *", ffi.from_buffer(buf))))
    res = {"untraced_ms": base, "traced_ms": traced}
    for layer in (0, 1):
        t = buf[layer].astype(np.int64)
        out = {}
        for X in (0, 1):
            tt = t[:, X * 20:X * 20 + 20]
            rel = tt - tt[:, :1]
            out["tile%d" % X] = {"period": float(np.median(np.diff(tt[:, 0]))),
                                 "offsets": {PP_SLOTS[k]: float(np.median(rel[:, k])) for k in range(20)}}
        out["B_minus_A_h_seen"] = float(np.median(t[:, 20] - t[:, 0]))
        res["layer%d" % layer] = out
    m.close()
    return res


def check_timeline(arg):
    """'mode,B,T,n' -> stage completion times (ms) of n device-resident forwards queued back to back on alternating
    lanes: the schedule the two lanes really ran."""
    from medaka_b200 import libmedaka as lm, models
    from oracle import synth
    parts = arg.split(",")
    mode, B, T, n = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
    stagger_ms = float(parts[4]) if len(parts) > 4 else 0.0      # host-side delay before the second lane's first forward
    lib, ffi = lm.load(), lm.ffi
    m = models.GRUModel()
    m.load_state_dict(synth.synth_state_dict(0))
    m.set_rec_mode(mode)
    m.reserve(B, T)
    eng, dev = m.engine, 0
    feats = synth.synth_features_fast(B, T, 10, seed=3)

    def dalloc(nbytes):
        pp = ffi.new("void **")
        lm.check(lib.mdk_dev_alloc(dev, nbytes, pp))
        return pp[0]
    d_feats = dalloc(feats.nbytes)
    lm.check(lib.mdk_memcpy_h2d(dev, d_feats, ffi.from_buffer(feats), feats.nbytes))
    d_probs = [dalloc(B * T * 20), dalloc(B * T * 20)]

    def fwd(i):
        lm.check(lib.mdk_engine_forward_dev(eng, ffi.cast("const float *", d_feats), B, T,
                                            ffi.cast("float *", d_probs[i & 1]), ffi.NULL, ffi.NULL))
    for i in range(4):
        fwd(i)
    lm.check(lib.mdk_engine_sync(eng))
    ms = ffi.new("float *")
    lm.check(lib.mdk_engine_timer_start(eng))
    for i in range(n):
        fwd(i)
        if i == 0 and stagger_ms > 0:
            time.sleep(stagger_ms * 1e-3)
    lm.check(lib.mdk_engine_timer_stop(eng, ms))
    out = np.zeros((n, 8), dtype=np.float32)
    lm.check(lib.mdk_debug_timeline(eng, n, ffi.cast("float *", ffi.from_buffer(out))))
    rows = [[round(float(v), 2) for v in r[[0, 3, 4, 5, 6]]] for r in out]     # start, rec0, gemm, rec1, head done
    return {"total_ms": float(ms[0]), "ms_per_forward": float(ms[0]) / n, "start_rec0_gemm_rec1_head": rows}


def check_e2e_timeline(arg):
    """'B,T,bw,steps' -> the bench's host-buffer leg (batches of bw windows submitted with the look-ahead) and the stage
    completion times of its last groups."""
    from medaka_b200 import libmedaka as lm, models
    from oracle import synth
    B, T, bw, steps = [int(x) for x in arg.split(",")]
    lib, ffi = lm.load(), lm.ffi
    m = models.GRUModel()
    m.load_state_dict(synth.synth_state_dict(0))
    m.reserve(min(m.preferred_batch_size(), B), T)
    feats = m.pinned("f", (B, T, 10), np.float32)
    feats[...] = synth.synth_features_fast(min(B, 64), T, 10, seed=3)[np.arange(B) % min(B, 64)]
    probs = m.pinned("p", (2, B, T, 5), np.float32)
    labels = m.pinned("l", (2, B, T), np.uint8)
    depth = m.lookahead(bw, T)
    batches = [(a, min(B, a + bw)) for a in range(0, B, bw)]
    waits = []

    def run(n):
        pending = []
        for k in range(n):
            for a, b in batches:
                while len(pending) >= depth:
                    t0 = time.perf_counter()
                    m.wait(pending.pop(0))
                    waits.append(time.perf_counter() - t0)
                pending.append(m.submit_arrays(feats[a:b], probs[k % 2, a:b], labels[k % 2, a:b]))
        while pending:
            m.wait(pending.pop(0))
    run(2)
    lm.check(lib.mdk_engine_sync(m.engine))
    del waits[:]
    ms = ffi.new("float *")
    lm.check(lib.mdk_engine_timer_start(m.engine))
    t0 = time.perf_counter()
    run(steps)
    host_s = time.perf_counter() - t0
    lm.check(lib.mdk_engine_timer_stop(m.engine, ms))
    n = min(16, steps * ((B + 1183) // 1184))
    out = np.zeros((n, 8), dtype=np.float32)
    lm.check(lib.mdk_debug_timeline(m.engine, n, ffi.cast("float *", ffi.from_buffer(out))))
    rows = [[round(float(v), 1) for v in r[[0, 1, 3, 4, 5, 6, 7]]] for r in out]
    return {"ms_per_step": float(ms[0]) / steps, "host_ms_per_step": host_s * 1e3 / steps, "depth": depth,
            "wait_ms_total": sum(waits) * 1e3, "n_waits": len(waits),
            "start_in_rec0_gemm_rec1_head_end": rows}


CHECKS = {"e2e_timeline": check_e2e_timeline, "timeline": check_timeline, "pp": check_pp, "rec_timing": check_rec_timing, "pp_trace": check_pp_trace, "selftest": check_selftest, "forward": check_forward, "misc": check_misc, "determinism": check_determinism,
          "rec_trace": check_rec_trace}

PLAN = [
    ("determinism", "fp32,20,64"), ("determinism", "fp32,37,130"), ("determinism", "tc,200,300"),
    ("selftest", "0"), ("selftest", "3"),
    ("forward", "tc,20,64"), ("forward", "tc,37,130"), ("forward", "tc,1200,24"), ("forward", "tc,3,1500"),
    ("forward", "tc,200,2000"), ("forward", "fp32,20,64"), ("determinism", "fp32,20,64"), ("forward", "fp32,20,64"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check")
    ap.add_argument("--arg", default="")
    args = ap.parse_args()
    if args.check:
        print("RESULT " + json.dumps(CHECKS[args.check](args.arg)))
        return
    os.makedirs(OUT, exist_ok=True)
    report = {}
    for name, arg in PLAN:
        key = "%s[%s]" % (name, arg)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--check", name, "--arg", arg],
                               capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                report[key] = json.loads(lines[-1][7:])
            else:
                report[key] = {"error": (r.stdout[-1500:] + "\n" + r.stderr[-2500:])}
        except subprocess.TimeoutExpired:
            report[key] = {"error": "timeout"}
        report[key + ".s"] = round(time.time() - t0, 1)
        print(key, json.dumps(report[key])[:3000], flush=True)
        with open(os.path.join(OUT, "diag.json"), "w") as fh:
            json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
