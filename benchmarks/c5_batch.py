#!/usr/bin/env python
"""BASELINE config C5: batched, independent S2S registrations (64x1024 scan pairs) sharded over GPUs.

    python benchmarks/c5_batch.py --pairs 4096 [--lanes 8]                          # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/c5_batch.py ...  # N GPUs, one rank each

Every unit is a full registration from raw clouds: two index builds, two covariance passes (k = 20) and the LM align,
exactly what `NanoGICP::align` does for a fresh pair.  Units are independent, so ranks get contiguous shards
(sharding.shard_range) and no collective touches the data.  Inside a rank the C++ batch driver (ddlo_batch_*,
csrc/batch.cu) owns S lanes (CUDA stream + engine, align kernel limited to 148 // S SMs) and deals the units to them;
Python stages the inputs in HBM before the clock starts and hands over the job table.  Prints one JSON
line with registrations/s over all ranks (max-over-ranks time, strong scaling: the pair count is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

B_REG_C1 = 100.9e6  # algorithmic bytes of one C1 registration (2 indexes, 2 covariance passes, L = E = 5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096, help="registrations over ALL ranks")
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic scan pairs (frames f, f+1) that are cycled")
    ap.add_argument("--lanes", type=int, default=8, help="lanes (stream + engine) per GPU")
    ap.add_argument("--align-blocks", type=int, default=0, help="SMs one align may use (0 = 148 // lanes)")
    ap.add_argument("--host-threads", type=int, default=2, help="C++ threads that enqueue the units")
    ap.add_argument("--mode", choices=["waves", "lanes"], default="waves", help="align stage: batched round kernels over waves of units, or one cooperative launch per unit on its lane")
    ap.add_argument("--wave", type=int, default=64, help="units per wave (waves mode)")
    ap.add_argument("--out", type=str, default="", help="also write the JSON line to this file (rank 0)")
    args = ap.parse_args()

    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng
    from dynamic_direct_lidar_odometry_b200 import synth
    from dynamic_direct_lidar_odometry_b200.sharding import max_over_ranks, shard_range, sum_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo")  # host-side barrier / max only: the shards never exchange data

    begin, end = shard_range(args.pairs, rank, world)
    w = synth.make_world()
    scans = [synth.scan(f, 64, 1024, w) for f in range(args.unique + 1)]

    batch = ng.Batch(local_rank, lanes=args.lanes, align_blocks=args.align_blocks, host_threads=args.host_threads, mode=args.mode, wave_units=args.wave)
    ids = [batch.stage(s) for s in scans]
    units = [(ids[p % args.unique + 1], ids[p % args.unique], None) for p in range(begin, end)]  # source = frame f+1, target = frame f
    jobs = ng.Batch.jobs(units)
    warm = max(4 * args.lanes, 8 * args.wave if args.mode == "waves" else 0)
    batch.run(ng.Batch.jobs([units[i % len(units)] for i in range(warm)]))  # warm-up (allocator pools of every lane and wave slot, first touch)

    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    res = batch.run(jobs, raw=True)
    dt = time.perf_counter() - t0
    (tmax,) = max_over_ranks([dt], dist)
    (total,) = sum_over_ranks([end - begin], dist)
    ok = all(res[i].flags & 1 for i in range(end - begin))
    err = 0.0
    for i, p in enumerate(range(begin, end)):
        u = p % args.unique
        gt = np.linalg.inv(synth.pose(u)) @ synth.pose(u + 1)
        T = np.array(res[i].final_transformation, dtype=np.float64).reshape(4, 4).T
        err = max(err, float(np.abs(T[:3, 3] - gt[:3, 3]).max()))
    (err, bad) = max_over_ranks([err, 0.0 if ok else 1.0], dist)
    if rank == 0:
        try:
            peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        value = total / tmax
        line = {"metric": "gicp_s2s_batched_registrations_per_s", "value": value, "unit": "registrations/s", "n_gpus": world,
                "pairs": int(total), "seconds": tmax, "mode": batch.mode, "wave_units": args.wave if batch.mode == "waves" else None,
                "lanes_per_gpu": batch.lanes, "align_sms_per_lane": batch.align_blocks if batch.mode == "lanes" else None, "lm_rounds_and_polls": batch.stats(),
                "host_threads_per_gpu": batch.host_threads, "all_converged": bad == 0.0, "max_translation_error_vs_truth_m": err,
                "scaling": "strong (fixed pair count)", "driver": "ddlo_batch_submit / ddlo_batch_wait (C++)",
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_registration": B_REG_C1, "achieved": value / world * B_REG_C1 / 1e9,
                             "peak": peak, "unit": "GB/s per GPU", "frac": value / world * B_REG_C1 / 1e9 / peak},
                "config": {"workload": "C5: independent S2S registrations of 64x1024 synthetic scan pairs, full pipeline per pair",
                           "unique_pairs_cycled": args.unique}}
        print(json.dumps(line))
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(json.dumps(line) + "\n")
    batch.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
