"""The device-resident keyframe store (ddlo_keyframes_*) and the C3 loop written in C++.  B200 box."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng
from dynamic_direct_lidar_odometry_b200 import odometry_loop as ol
from dynamic_direct_lidar_odometry_b200 import synth
from oracle_backend import OracleBackend, qhull_concave, qhull_convex

pytestmark = pytest.mark.gpu

POSE_T, POSE_R = 1e-5, 1e-6


def rot_angle(Ra, Rb):
    R = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    return float(np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0)


def yaw(deg):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def test_keyframe_store_selection_and_assembly(rt):
    """ddlo_keyframes_get_submap against OdomNode::getSubmapKeyframes restated in Python with Qhull (scipy) for the hulls:
    same indices (k nearest + convex-hull + concave-hull keyframes, ties included), `changed` only when the selection
    differs from the previous call, cloud and covariances concatenated on the device in index order."""
    rng = np.random.default_rng(4)
    store = ng.KeyframeStore(rt)
    clouds, covs, pos = [], [], []
    prev = None
    for i in range(28):
        p = np.array([0.8 * i + rng.normal(0, 0.2), 6 * np.sin(0.35 * i) + rng.normal(0, 0.2), 1.5 + rng.normal(0, 0.005)], dtype=np.float32)
        pts = np.ones((40 + i, 4), dtype=np.float32)
        pts[:, :3] = (rng.normal(0, 1, (40 + i, 3)) + p).astype(np.float32)
        c = ng.PointCloud(rt, pts)
        v = ng.Covariances.compute(c, 10)
        store.add(p, ng.rotation_to_wxyz(yaw(3 * i)), c, v)
        clouds.append(pts)
        covs.append(v.to_host())
        pos.append(p)
        assert len(store) == i + 1
        query = (p + rng.normal(0, 0.5, 3)).astype(np.float32)
        for knn, kcv, kcc, alpha in ((3, 2, 2, 1.5), (1, 0, 3, 4.0)):
            changed, idx, cloud, cv = store.get_submap(query, knn, kcv, kcc, alpha)
            # the restatement
            P = np.array(pos, dtype=np.float64)
            d = [float(np.float32(np.sqrt(np.sum((query - q).astype(np.float64) ** 2)))) for q in pos]
            sel = []
            ol.OdometryLoop._push_submap_indices(d, knn, list(range(len(d))), sel)
            convex = qhull_convex(P) if len(pos) >= 4 else []
            concave = qhull_concave(P, alpha) if len(pos) >= 5 else []
            ol.OdometryLoop._push_submap_indices([d[j] for j in convex], kcv, convex, sel)
            ol.OdometryLoop._push_submap_indices([d[j] for j in concave], kcc, concave, sel)
            want = sorted(set(sel))
            assert idx == want, (i, knn, kcv, kcc)
            got_cv, got_cc, dim = store.hulls()
            assert dim == 2 and got_cv == convex and (got_cc == concave or len(pos) < 5)
            assert changed == (want != prev)
            prev = want
            if changed:
                assert np.array_equal(cloud.download(), np.concatenate([clouds[j] for j in want]))
                assert np.array_equal(cv.to_host(), np.concatenate([covs[j] for j in want]))
            else:
                assert cloud is None and cv is None


def test_new_keyframe_decision(rt):
    """updateKeyframes' decision (odom.cc:1067-1126): distance / rotation thresholds and the `num_nearby <= 1` rule"""
    store = ng.KeyframeStore(rt)
    pts = synth.scan(0, 8, 64)
    c, v = ng.PointCloud(rt, pts), ng.Covariances.compute(ng.PointCloud(rt, pts), 5)
    store.add([0, 0, 0], ng.rotation_to_wxyz(np.eye(3)), c, v)
    q0 = ng.rotation_to_wxyz(np.eye(3))
    new, idx, d, th = store.is_new([0.4, 0, 0], q0, 1.0, 15.0)
    assert (new, idx) == (False, 0) and abs(d - 0.4) < 1e-6 and abs(th) < 1e-3
    assert store.is_new([1.2, 0, 0], q0, 1.0, 15.0)[0]                      # moved further than the threshold
    new, _, _, th = store.is_new([0.2, 0, 0], ng.rotation_to_wxyz(yaw(20)), 1.0, 15.0)
    assert new and abs(th - 20.0) < 1e-3                                     # turned on the spot, only one keyframe nearby
    store.add([0.5, 0, 0], q0, c, v)
    assert not store.is_new([0.2, 0, 0], ng.rotation_to_wxyz(yaw(20)), 1.0, 15.0)[0]   # ... but two nearby: no keyframe
    assert store.is_new([0.2, 0, 0], ng.rotation_to_wxyz(yaw(20)), 0.1, 15.0)[0]       # beyond the distance threshold again
    with pytest.raises(ng.DdloError):
        ng.KeyframeStore(rt).is_new([0, 0, 0], q0, 1.0, 15.0)


def sequence(frames, beams, cols):
    w = synth.make_world()
    return [synth.scan(f, beams, cols, w) for f in range(frames)]


def test_sequence_loop_with_hull_selection_matches_oracle(rt, oracle):
    """the odometry loop with the full keyframe selection of getSubmapKeyframes (k nearest + convex + concave hull;
    C++ hulls on the GPU side, Qhull through scipy on the oracle side): same decisions, same iteration counts, poses
    within the bar frame by frame"""
    scans = sequence(34, 16, 256)
    cfg = ol.LoopConfig(k_correspondences_s2s=10, k_correspondences_s2m=10, keyframe_thresh_dist=0.25, submap_knn=2, submap_kcv=2, submap_kcc=2)
    got = ol.run_sequence(ol.GpuBackend(rt), scans, cfg)
    want = ol.run_sequence(OracleBackend(oracle), scans, cfg)
    assert len(got.keyframes) == len(want.keyframes) >= 8  # hulls need 4 / 5 keyframes to come into play
    assert got._convex == want._convex and got._concave == want._concave and len(got._convex) >= 3
    for g, o in zip(got.records, want.records):
        assert (g.s2s_iterations, g.s2m_iterations, g.new_keyframe, g.submap_changed, g.submap_points) == (
            o.s2s_iterations, o.s2m_iterations, o.new_keyframe, o.submap_changed, o.submap_points)
        assert np.abs(g.T[:3, 3].astype(np.float64) - o.T[:3, 3]).max() < POSE_T
        assert rot_angle(g.T[:3, :3], o.T[:3, :3]) < POSE_R


@pytest.mark.parametrize("voxel", [0.0, 0.4])
def test_cpp_odometry_sequence_matches_python_loop(rt, tmp_path, voxel):
    """tests/cpp/odometry_sequence.cpp - the C3 frame loop in C++ on the C ABI and the keyframe store, no Python in the
    frame - against the Python harness on the same library: identical decisions and iteration counts, poses equal to
    float rounding of the pose products (the two hosts multiply 4x4 float matrices in different orders)."""
    exe = Path(__file__).resolve().parent / "cpp" / "_build" / "odometry_sequence"
    if not exe.exists():
        import __graft_entry__ as ge

        ge.build_cpp_tests()
    scans = sequence(30, 16, 256)
    path = tmp_path / "scans.bin"
    with open(path, "wb") as fh:
        fh.write(np.int32(len(scans)).tobytes())
        for s in scans:
            fh.write(np.int32(len(s)).tobytes())
            fh.write(np.ascontiguousarray(s, dtype=np.float32).tobytes())
    out = subprocess.run([str(exe), str(path), "10", "0.25", "15", "2", "2", "2", str(voxel), str(voxel)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [line.split() for line in out.stdout.splitlines() if line.startswith("frame")]
    cfg = ol.LoopConfig(k_correspondences_s2s=10, k_correspondences_s2m=10, keyframe_thresh_dist=0.25, submap_knn=2, submap_kcv=2, submap_kcc=2,
                        voxel_leaf_scan=voxel or None, voxel_leaf_submap=voxel or None)
    loop = ol.run_sequence(ol.GpuBackend(rt), scans, cfg)
    assert len(rows) == len(loop.records) == len(scans) - 1
    summary = [line.split() for line in out.stdout.splitlines() if line.startswith("summary")][0]
    assert int(summary[4]) == len(loop.keyframes) >= 6
    for row, r in zip(rows, loop.records):
        assert [int(x) for x in row[2:9]] == [r.s2s_iterations, r.s2m_iterations, int(r.s2s_converged), int(r.s2m_converged), int(r.new_keyframe),
                                              int(r.submap_changed), r.submap_points]
        T = np.array(row[11:27], dtype=np.float32).reshape(4, 4)
        assert np.abs(T[:3, 3] - r.T[:3, 3]).max() < POSE_T and rot_angle(T[:3, :3], r.T[:3, :3]) < 2 * POSE_R
        assert abs(float(row[10]) - r.residual_mean) <= 1e-4 * max(1.0, abs(r.residual_mean))  # the guesses differ by float rounding


def test_keyframe_store_three_dimensional_positions(rt):
    """positions that are three-dimensional in PCL's sense: the convex hull is the 3-D one (checked against Qhull), the
    concave hull is reported as not computed and contributes nothing to the selection"""
    rng = np.random.default_rng(11)
    store = ng.KeyframeStore(rt)
    pts = synth.scan(0, 8, 64)
    c, v = ng.PointCloud(rt, pts), ng.Covariances.compute(ng.PointCloud(rt, pts), 5)
    pos = rng.normal(0, 5, (30, 3)).astype(np.float32)
    for p in pos:
        store.add(p, ng.rotation_to_wxyz(np.eye(3)), c, v)
    query = np.zeros(3, np.float32)
    changed, idx, cloud, cv = store.get_submap(query, 3, 4, 4, 2.0)
    convex, concave, dim = store.hulls()
    assert dim == 3 and concave == [] and convex == qhull_convex(pos.astype(np.float64))
    d = [float(np.float32(np.sqrt(np.sum((query - q).astype(np.float64) ** 2)))) for q in pos]
    sel = []
    ol.OdometryLoop._push_submap_indices(d, 3, list(range(30)), sel)
    ol.OdometryLoop._push_submap_indices([d[j] for j in convex], 4, convex, sel)
    assert changed and idx == sorted(set(sel)) and len(cloud) == len(idx) * len(pts)
