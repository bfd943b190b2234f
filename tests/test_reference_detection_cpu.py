"""Pins the segmentation oracle (oracle/oracle_segmentation.cpp) to the REFERENCE'S OWN DetectionModule code: the class
declaration of include/detection/detection.h unmodified, and the definitions of loadParams, allocateMemory,
resetParameters, projectResiduals, projectScan, groundRemoval, cloudSegmentation and labelComponents extracted from
src/detection/detection.cpp at build time (oracle/extract_detection.py), compiled over stand-ins for ROS / OpenCV / PCL /
Eigen (oracle/stub_include/tracking/tracking.h) into oracle/_ref/libdetection_ref.so (oracle/refdet.py).  Every output of
this stage is integer or exact-float work, so the two must agree bit for bit.

The library is prebuilt where /root/reference exists and travels to the GPU box; where it is absent the comparison
tests skip and the committed fixture of its outputs (tests/golden/segmentation_reference.npz) still pins the oracle.
"""
from pathlib import Path

import numpy as np
import pytest

import segmentation_cases as cases
from dynamic_direct_lidar_odometry_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden" / "segmentation_reference.npz"


@pytest.fixture(scope="module")
def refdet(oracle):
    from oracle import refdet as rd

    if not rd.available():
        pytest.skip("oracle/_ref/libdetection_ref.so not built (needs /root/reference)")
    rd.lib()
    return rd


def same(a, b):
    assert a["label_count"] == b["label_count"]
    assert np.array_equal(a["label_mat"], b["label_mat"])
    assert np.array_equal(a["ground_mat"], b["ground_mat"])
    assert np.array_equal(a["range_mat"].view(np.uint32), b["range_mat"].view(np.uint32))
    assert np.array_equal(a["avg_residuals"][1:].view(np.uint64), b["avg_residuals"][1:].view(np.uint64))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("with_residuals", [True, False])
def test_oracle_matches_reference_detection_code(oracle, refdet, variant, with_residuals):
    params, st, T, res = cases.reference_case(variant)
    r = refdet.segment(params, st, T, res if with_residuals else None)
    o = oracle.segment_scan(oracle.SegParams(**params), st, T, res if with_residuals else None)
    same(o, r)
    if variant > 0:
        assert r["label_count"] > 4 and (r["label_mat"] == 999999).any() and ((r["ground_mat"] == 1).any() or params["ground_rows"] == 0)
    if not with_residuals:
        assert not r["avg_residuals"].any()


@pytest.mark.parametrize("frame,dropout", [(2, 0.0), (35, 0.2), (60, 0.05)])
def test_oracle_matches_reference_detection_code_other_frames(oracle, refdet, frame, dropout):
    params, st, T, res = cases.reference_case(1, frame=frame, size=400, dropout=dropout)
    same(oracle.segment_scan(oracle.SegParams(**params), st, T, res), refdet.segment(params, st, T, res))


def test_reference_detection_defaults_are_the_documented_ones(refdet):
    # no overrides but the image size: loadParams' own defaults (detection.cpp:76-105) drive the result; an empty scan
    # gives an all -1 label image and label_count_ 1
    params = dict(rows=360, cols=360)
    s = np.full((360, 360, 4), np.nan, dtype=np.float32)
    r = refdet.segment(params, s, np.eye(4, dtype=np.float32))
    assert r["label_count"] == 1 and (r["label_mat"] == -1).all() and not r["range_mat"].any()


def test_golden_fixture_is_current(oracle, refdet):
    params, st, T, res = cases.golden_case()
    g = np.load(GOLDEN)
    r = refdet.segment(params, st, T, res)
    assert int(g["label_count"]) == r["label_count"] and np.array_equal(g["label_mat"], r["label_mat"])


def test_oracle_matches_golden_fixture_of_the_reference(oracle):
    """runs everywhere: the stored answers of the reference's code for the stored inputs"""
    g = np.load(GOLDEN)
    params = {k: (float(v) if k in ("theta", "max_delta_z", "max_elevation") else int(v)) for k, v in zip(g["param_names"], g["param_values"])}
    o = oracle.segment_scan(oracle.SegParams(**params), g["scan_t"], g["T"], g["residuals"])
    same(o, dict(label_count=int(g["label_count"]), label_mat=g["label_mat"], ground_mat=g["ground_mat"], range_mat=g["range_mat"],
                 avg_residuals=g["avg_residuals"]))
    assert int(g["label_count"]) > 3


def test_residual_cloud_matches_reference_loop(oracle, refdet):
    """oracle_residual_image against the reference's own loop of OdomNode::scanMatching (odom.cc:804-827), extracted and
    compiled like the DetectionModule functions: every cell bit for bit, including which point wins a shared cell and
    the float atan2 / sqrt overloads the reference's `using namespace std;` selects"""
    rng = np.random.default_rng(0)
    n = 150_000
    th, ph, r = rng.uniform(-1.2, 1.2, n), rng.uniform(-1.2, 1.2, n), rng.uniform(0.5, 30.0, n)
    pts = np.stack([r * np.sin(th) * np.cos(ph), r * np.sin(ph), r * np.cos(th) * np.cos(ph), np.ones(n)], 1).astype(np.float32)
    res = rng.uniform(0.0, 1.0, n)
    want = refdet.residual_cloud(pts, res)
    got = oracle.residual_image(pts, res)
    assert (want[..., 3] != 0).sum() > 50_000
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a spinning-LiDAR scan: most points outside the +-60 degree image, still identical
    scan = synth.scan(3, 32, 512)
    res = np.linalg.norm(scan[:, :3], axis=1) * 1e-3
    assert np.array_equal(oracle.residual_image(scan, res).view(np.uint32), refdet.residual_cloud(scan, res).view(np.uint32))


def test_oracle_matches_reference_detection_code_fuzz(oracle, refdet):
    """random 360 x 360 range images (steps, holes, exact zeros, repeated heights) and random parameters within what the
    reference can express (its window, integer values for its int-typed parameters)"""
    rng = np.random.default_rng(77)
    accepted = 0
    for trial in range(12):
        H = W = 360
        el = np.linspace(0.5, -0.5, H)[:, None]
        az = np.linspace(-1.0, 1.0, W)[None, :]
        depth = 7.0 + 3.0 * np.sin(az * rng.uniform(1, 6) + rng.uniform(0, 6)) + 2.0 * np.cos(el * rng.uniform(2, 9))
        for _ in range(int(rng.integers(2, 14))):
            r0, c0 = int(rng.integers(140, 350)), int(rng.integers(140, 350))
            depth[r0:r0 + int(rng.integers(3, 60)), c0:c0 + int(rng.integers(3, 60))] -= rng.uniform(1.0, 4.0)
        depth = np.maximum(depth, 0.6) + rng.normal(0.0, 0.004, (H, W))
        s = np.empty((H, W, 4), dtype=np.float32)
        s[..., 0] = depth * np.cos(el) * np.cos(az)
        s[..., 1] = depth * np.cos(el) * np.sin(az)
        s[..., 2] = depth * np.sin(el)
        s[..., 3] = 1.0
        s[rng.random((H, W)) < rng.uniform(0.0, 0.1)] = np.nan
        s[rng.random((H, W)) < 0.01, 2] = 0.0
        if trial % 2:
            s[..., 2] = np.round(s[..., 2], 1)
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
        params = dict(rows=H, cols=W, ground_rows=int(rng.integers(0, 120)), ang_bottom=int(rng.integers(10, 40)), minimum_range=int(rng.integers(0, 3)),
                      sensor_mount_angle=int(rng.integers(-5, 6)), ground_angle_threshold=int(rng.integers(2, 15)), max_distance=int(rng.integers(6, 14)),
                      theta=float(rng.uniform(0.05, 1.2)), valid_point_num=int(rng.integers(1, 12)), valid_line_num=int(rng.integers(0, 4)),
                      min_line_num=int(rng.integers(0, 4)), min_delta_z=float(rng.uniform(0.0, 0.3)), max_delta_z=float(rng.uniform(0.3, 4.0)),
                      max_elevation=float(rng.uniform(-0.5, 3.0)), **cases.WINDOW)
        res = (rng.random((H, W)) * (rng.random((H, W)) < 0.8)).astype(np.float32) if trial % 4 else None
        r = refdet.segment(params, s, T, res)
        same(oracle.segment_scan(oracle.SegParams(**params), s, T, res), r)
        accepted += r["label_count"] - 1
    assert accepted >= 30
