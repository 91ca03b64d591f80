#!/usr/bin/env python
"""Whole-pipeline rate of `medaka inference`'s body (medaka/prediction.py:83-222) on synthetic pileups: region triage ->
DataLoader threads (synthetic counts per region -> GPU count normalisation -> windows -> Batch.collate) -> engine
(200-window batches, look-ahead, coalesced groups) -> output store (.npzstore), one process per GPU.

    python tools/pipeline_bench.py [--mb 20] [--batch-size 200]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/pipeline_bench.py --mb 40

Prints one JSON line per run (rank 0): pileup columns / s through the whole pipeline, wall clock (this is a host
pipeline: threads, numpy, file writes), max over ranks.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=20.0, help="draft megabases (all ranks together)")
    ap.add_argument("--region-mb", type=float, default=1.0)
    ap.add_argument("--batch-size", type=int, default=200)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--profile", action="store_true", help="cProfile every thread of the timed run, print the top entries")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from medaka_b200 import common, features, libmedaka as lm, models, prediction
    from oracle import synth      # seeded synthetic weights / counts only
    dev = local_rank if world > 1 else 0
    lm.require_gpu(dev)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    base_counts, base_pos = synth.synth_counts(int(args.region_mb * 1e6 * 1.18) + 8, seed=11)

    def pileup_source(region, bam, encoder):
        # one synthetic count table re-based per region (~15 % insertion columns): the generator itself is not the
        # subject of the measurement
        n_ref = region.end - region.start
        pos = base_pos.copy()
        keep = pos["major"] < n_ref
        pos = pos[keep]
        pos["major"] += region.start
        return [(base_counts[keep], pos)]

    model = models.GRUModel(num_features=10, device=dev)
    model.load_state_dict(synth.synth_state_dict(0))
    enc = features.CountsFeatureEncoder(normalise="total", pileup_source=pileup_source)
    n_regions = max(world, int(round(args.mb / args.region_mb)))
    regions = [common.Region("ctg%d" % i, 0, int(args.region_mb * 1e6)) for i in range(n_regions)]
    mine = prediction.shard_regions(regions, world)[rank] if world > 1 else regions
    tmp = tempfile.mkdtemp(prefix="mdk_pipe_")
    out = args.out or os.path.join(tmp, "probs_rank%d.npzstore" % rank)
    # warm-up on one small region (allocations, first-touch of pinned buffers)
    warm = os.path.join(tmp, "warm.npzstore")
    prediction.predict_regions(warm, None, [common.Region("warm", 0, 200000)], model, enc, chunk_len=10000, chunk_ovlp=1000,
                               batch_size=args.batch_size, bam_chunk=1000000, bam_workers=args.workers)
    if dist is not None:
        dist.barrier()
    acc = {}
    if args.profile:
        # wall-clock accumulators around the stages (summed over the threads that run them); cProfile cannot follow
        # several threads at once under Python 3.12
        import threading
        from medaka_b200 import datastore, torch_ext
        lock = threading.Lock()

        def timed(owner, name, label):
            fn = getattr(owner, name)

            def wrapper(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    d = time.perf_counter() - t
                    with lock:
                        acc[label] = acc.get(label, 0.0) + d
                        acc[label + " calls"] = acc.get(label + " calls", 0) + 1
            setattr(owner, name, wrapper)
        enc.pileup_source = (lambda f: (lambda *a: f(*a)))(pileup_source)
        timed(enc, "pileup_source", "synthetic counts (loader threads)")
        timed(enc, "_post_process_pileup", "normalise on GPU incl. copies (loader threads)")
        timed(torch_ext.Batch, "collate", "Batch.collate (batcher thread)")
        timed(model, "predict_async", "predict_async: copy into pinned + submit (main)")
        timed(model, "wait", "engine wait (main)")
        timed(datastore._NpzBackend, "write_fields", "store write_fields (writer thread)")
        timed(datastore.DataStore, "write_sample", "write_sample submit (main)")
    t0 = time.perf_counter()
    prediction.predict_regions(out, None, mine, model, enc, chunk_len=10000, chunk_ovlp=1000,
                               batch_size=args.batch_size, bam_chunk=1000000, bam_workers=args.workers)
    dt = time.perf_counter() - t0
    if args.profile:
        sys.stderr.write("stage seconds (wall, summed over threads): %s\n" % json.dumps({k: round(v, 3) for k, v in acc.items()}))
    cols = sum(int((base_pos["major"] < (r.end - r.start)).sum()) for r in mine)
    if dist is not None:
        import torch
        t = torch.tensor([dt, float(cols)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, cols = float(tmax[0]), float(t[1])
    if rank == 0:
        print(json.dumps({"metric": "pileup columns/s through predict_regions (synthetic pileups -> store)",
                          "value": cols / dt, "unit": "columns/s", "n_gpus": world, "seconds": dt, "columns": cols,
                          "regions": n_regions, "batch_size": args.batch_size, "workers": args.workers,
                          "timing": "host wall clock, max over ranks"}))
    model.close()
    shutil.rmtree(tmp, ignore_errors=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
