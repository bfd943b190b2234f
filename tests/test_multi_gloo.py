"""The N>1 path of bench.py on CPU: two gloo ranks shard independent registrations (run on the oracle
here, there is no GPU), no data-path collective, device time reduced with max-over-ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamic_direct_lidar_odometry_b200 import synth
    from dynamic_direct_lidar_odometry_b200.sharding import max_over_ranks, shard_range, sum_over_ranks
    from oracle import pyoracle as po

    n_pairs = 5
    b, e = shard_range(n_pairs, rank, world)
    w = synth.make_world()
    poses = []
    for p in range(b, e):  # pair p = frames (p, p+1), an independent unit
        eng = po.NanoGICP(threads=1)
        eng.setInputSource(po.Cloud(synth.scan(p + 1, 8, 128, w)))
        eng.setInputTarget(po.Cloud(synth.scan(p, 8, 128, w)))
        poses.append(eng.align().T)
    fake_ms = 10.0 * (rank + 1)
    (tmax,) = max_over_ranks([fake_ms], dist)
    (total,) = sum_over_ranks([e - b], dist)
    np.save(os.path.join(out_dir, f"poses_{rank}.npy"), np.array(poses).reshape(-1, 4, 4))
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([tmax, total, b, e]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    from dynamic_direct_lidar_odometry_b200 import synth
    from oracle import pyoracle as po

    po.build()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    metas = [np.load(tmp_path / f"meta_{r}.npy") for r in range(world)]
    assert all(m[0] == 20.0 for m in metas)      # max over ranks of the per-rank time
    assert all(m[1] == 5 for m in metas)         # every unit processed exactly once
    assert metas[0][3] == metas[1][2]            # contiguous, disjoint shards
    got = np.concatenate([np.load(tmp_path / f"poses_{r}.npy") for r in range(world)])
    w = synth.make_world()
    for p in range(5):                           # identical to the single-process answer
        eng = po.NanoGICP(threads=1)
        eng.setInputSource(po.Cloud(synth.scan(p + 1, 8, 128, w)))
        eng.setInputTarget(po.Cloud(synth.scan(p, 8, 128, w)))
        assert np.array_equal(eng.align().T, got[p])
