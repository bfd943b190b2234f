"""Host logic of the odometry-loop harness (odometry_loop.py) on the CPU oracle: no GPU needed.

The harness replays OdomNode's protocol around two NanoGICP engines (odom.cc:480-532, 745-793,
1067-1150, 1215-1315).  Here it is driven with the oracle backend on a short, small synthetic
sequence and checked against the generator's ground truth: the sensor moves 0.1 m / 0.5 deg per frame,
so the estimated trajectory must follow it to a few millimetres, keyframes must appear about every
metre, and the submap must only be re-injected when the keyframe selection changed.
"""
import numpy as np

from dynamic_direct_lidar_odometry_b200 import odometry_loop as ol
from oracle_backend import OracleBackend
from dynamic_direct_lidar_odometry_b200 import synth


def _truth(frame):
    return np.linalg.inv(synth.pose(0)) @ synth.pose(frame)


def test_oracle_loop_tracks_ground_truth(oracle):
    w = synth.make_world()
    frames = 14
    scans = [synth.scan(f, 16, 256, w) for f in range(frames)]
    cfg = ol.LoopConfig(k_correspondences_s2s=10, k_correspondences_s2m=10, keyframe_thresh_dist=0.5, submap_knn=3)
    loop = ol.run_sequence(OracleBackend(oracle), scans, cfg)
    assert len(loop.records) == frames - 1
    for f, r in enumerate(loop.records, start=1):
        assert r.s2s_converged and r.s2m_converged
        gt = _truth(f)
        assert np.abs(r.T[:3, 3].astype(np.float64) - gt[:3, 3]).max() < 0.02, (f, r.T[:3, 3], gt[:3, 3])
    # 1.3 m travelled with a 0.5 m threshold: first keyframe + at least two more
    assert 3 <= len(loop.keyframes) <= 4
    # the submap is re-injected exactly when the selection changed: on the first frame and after new keyframes
    changed = [r.submap_changed for r in loop.records]
    assert changed[0]
    assert sum(changed) <= 1 + sum(r.new_keyframe for r in loop.records)
    assert all(r.submap_points > 0 for r in loop.records)


def test_oracle_transform_matches_reference_float_order(oracle):
    """OracleBackend.transform restates pcl::transformPointCloud in float: (r0 x + r1 y) + (r2 z + t)."""
    rng = np.random.default_rng(3)
    p = np.ones((100, 4), dtype=np.float32)
    p[:, :3] = rng.uniform(-50, 50, (100, 3)).astype(np.float32)
    T = synth.perturbed_guess(synth.pose(7))
    out = OracleBackend(oracle).transform(oracle.Cloud(p), T).points
    want = (T.astype(np.float64) @ p.astype(np.float64).T).T
    assert np.abs(out[:, :3] - want[:, :3]).max() < 1e-4
    assert np.all(out[:, 3] == 1.0)
