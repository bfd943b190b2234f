"""Parity of the range-image segmentation stage with the CPU oracle, through the C ABI.

Everything here is integer / exact-float work: label images, ground flags and the segment count must be identical;
the range image and the per-segment average residuals bit-exact.  The oracle counts slope tests that land within a few
ulp of their threshold (`borderline`): a different libm could flip those, so a case is only compared while that count
is zero (it is for every seed used here).
"""
import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import binding as B
from dynamic_direct_lidar_odometry_b200 import synth
from dynamic_direct_lidar_odometry_b200.detection import DetectionModule, INVALID_SEGMENT

pytestmark = pytest.mark.gpu


def run_both(rt, oracle, params, scan_t, T, residuals):
    det = DetectionModule(rt, **params)
    det.projectScan(None, scan_t, T)
    if residuals is not None:
        det.projectResiduals(residuals)
    det.applySegmentation()
    o = oracle.segment_scan(oracle.SegParams(**params), scan_t, T, residuals)
    return det, o


def assert_same(det, o):
    assert o["borderline"] == 0, "pick another seed: a slope test sits on its threshold"
    assert np.array_equal(det.range_mat.view(np.uint32), o["range_mat"].view(np.uint32))
    assert np.array_equal(det.ground_mat, o["ground_mat"])
    assert det.label_count_ == o["label_count"]
    assert np.array_equal(det.label_mat, o["label_mat"])
    assert np.array_equal(det.avg_residuals[1:].view(np.uint64), o["avg_residuals"][1:].view(np.uint64))


def lidar_case(frame, beams, cols, dropout, **over):
    sc = synth.organized_scan(frame, beams, cols, dropout=dropout)
    T = synth.pose(frame).astype(np.float32)
    st = synth.organized_transform(sc, T)
    params = dict(rows=beams, cols=cols, ground_rows=beams * 3 // 8, window_row_min=0, window_row_max=beams - 1, window_col_min=0,
                  window_col_max=cols - 1, ang_bottom=22.5, minimum_range=1.0, sensor_mount_angle=0.0, max_distance=40.0)
    params.update(over)
    res = np.abs(np.random.default_rng(frame).normal(0.0, 0.05, (beams, cols))).astype(np.float32)
    res[np.random.default_rng(frame + 1).random((beams, cols)) < 0.3] = 0.0
    return params, st, T, res


@pytest.mark.parametrize("frame,beams,cols,dropout", [(3, 64, 1024, 0.02), (40, 64, 1024, 0.0), (7, 16, 256, 0.1), (12, 128, 2048, 0.01),
                                                      (5, 33, 255, 0.05)])
def test_segmentation_matches_oracle(rt, oracle, frame, beams, cols, dropout):
    params, st, T, res = lidar_case(frame, beams, cols, dropout)
    det, o = run_both(rt, oracle, params, st, T, res)
    assert_same(det, o)
    assert det.getSegmentsCount() > 3                      # the case really has accepted segments,
    assert (det.label_mat == INVALID_SEGMENT).any()        # rejected ones
    assert (det.ground_mat == 1).sum() > beams * cols // 20  # and ground


def test_segmentation_reference_window_and_defaults(rt, oracle):
    # the fork's own geometry: a 512 x 512 image and the hard-coded window 156..356 (detection.cpp:520-522)
    params, st, T, res = lidar_case(9, 512, 512, 0.02, window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356,
                                    ground_rows=30)
    det, o = run_both(rt, oracle, params, st, T, res)
    assert_same(det, o)
    lm = det.label_mat
    outside = np.ones_like(lm, dtype=bool)
    outside[156:357, 156:357] = False
    assert not ((lm[outside] > 0)).any()  # nothing outside the window is ever labelled


def test_segmentation_without_residuals(rt, oracle):
    params, st, T, _ = lidar_case(3, 64, 1024, 0.02)
    det, o = run_both(rt, oracle, params, st, T, None)
    assert_same(det, o)
    assert not det.avg_residuals.any()


def test_segmentation_residual_cloud_input(rt, oracle):
    # projectResiduals takes the residual cloud itself: intensity where the point is finite, zero elsewhere (:240-249)
    params, st, T, res = lidar_case(3, 32, 512, 0.05)
    cloud = np.zeros((32, 512, 4), dtype=np.float32)
    cloud[..., 3] = res
    cloud[::7, ::5, 0] = np.nan
    expect = res.copy()
    expect[::7, ::5] = 0.0
    det = DetectionModule(rt, **params)
    det.projectScan(None, st, T)
    det.projectResiduals(cloud)
    det.applySegmentation()
    o = oracle.segment_scan(oracle.SegParams(**params), st, T, expect)
    assert_same(det, o)


def sphere_image(H, W, radius=10.0):
    """every pixel at the same range: one component covering the whole image, the widest flood-fill fronts"""
    el = np.linspace(0.6, -0.6, H)[:, None]
    az = np.linspace(-1.2, 1.2, W)[None, :]
    s = np.empty((H, W, 4), dtype=np.float32)
    s[..., 0] = radius * np.cos(el) * np.cos(az)
    s[..., 1] = radius * np.cos(el) * np.sin(az)
    s[..., 2] = radius * np.sin(el) * np.ones_like(az)
    s[..., 3] = 1.0
    return s


@pytest.mark.parametrize("ring", [None, "64"])
def test_segmentation_one_huge_component(rt, oracle, monkeypatch, ring):
    # 540 x 540 pixels in a single segment; with a 64-entry ring most queue reads come back from global memory
    if ring:
        monkeypatch.setenv("DDLO_SEG_RING", ring)
    H = W = 540
    s = sphere_image(H, W)
    rng = np.random.default_rng(1)
    s[..., :3] *= (1.0 + 0.001 * rng.standard_normal((H, W, 1))).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    params = dict(rows=H, cols=W, ground_rows=0, window_row_min=0, window_row_max=H - 1, window_col_min=0, window_col_max=W - 1,
                  ang_bottom=34.0, minimum_range=1.0, theta=0.2, max_distance=50.0, max_delta_z=20.0, max_elevation=20.0)
    res = rng.random((H, W)).astype(np.float32)
    det, o = run_both(rt, oracle, params, s, T, res)
    assert_same(det, o)
    assert det.getSegmentsCount() == 1 and (det.label_mat == 1).sum() > H * W * 0.99


def test_segmentation_edge_cases(rt, oracle):
    T = np.eye(4, dtype=np.float32)
    base = dict(rows=8, cols=16, ground_rows=3, window_row_min=0, window_row_max=7, window_col_min=0, window_col_max=15, ang_bottom=10.0,
                minimum_range=1.0, valid_point_num=2, valid_line_num=1, min_line_num=1, min_delta_z=-10.0, max_delta_z=10.0)
    # no return at all
    s = np.full((8, 16, 4), np.nan, dtype=np.float32)
    det, o = run_both(rt, oracle, base, s, T, None)
    assert_same(det, o)
    assert det.label_count_ == 1 and (det.label_mat == -1).all()
    # a point with x == 0 exactly is "no info" for the ground test (:476-481); a z of exactly 0 never becomes min_z (:612)
    s = sphere_image(8, 16, 5.0)
    s[6, 3, 0] = 0.0
    s[2, 5:9, 2] = 0.0
    det, o = run_both(rt, oracle, base, s, T, np.ones((8, 16), dtype=np.float32))
    assert_same(det, o)
    assert (det.ground_mat == -1).any()
    # everything closer than minimum_range
    det, o = run_both(rt, oracle, {**base, "minimum_range": 100.0}, s, T, None)
    assert_same(det, o)
    assert (det.range_mat == 0).all()
    # an empty window
    det, o = run_both(rt, oracle, {**base, "window_row_min": 156, "window_row_max": 356}, s, T, None)
    assert_same(det, o)
    assert det.label_count_ == 1
    # 32-byte points (pcl::PointXYZI) give the same result as packed ones
    wide = np.zeros((8, 16, 8), dtype=np.float32)
    wide[..., :4] = s
    det2, _ = run_both(rt, oracle, base, wide, T, None)
    det1, _ = run_both(rt, oracle, base, s, T, None)
    assert np.array_equal(det1.label_mat, det2.label_mat)


def test_segmentation_argument_errors(rt):
    s = np.zeros((8, 16, 4), dtype=np.float32)
    det = DetectionModule(rt, rows=8, cols=16, ground_rows=3, window_col_max=16)
    det.projectScan(None, s, np.eye(4))
    with pytest.raises(B.DdloError) as e:
        det.applySegmentation()
    assert e.value.code == -8  # DDLO_E_UNSUPPORTED
    det = DetectionModule(rt, rows=8, cols=16, ground_rows=8, window_col_max=15)
    det.projectScan(None, s, np.eye(4))
    with pytest.raises(B.DdloError) as e:
        det.applySegmentation()
    assert e.value.code == -1  # DDLO_E_INVALID
    with pytest.raises(ValueError):
        det.projectScan(None, np.zeros((5, 4), dtype=np.float32), np.eye(4))


def test_segmentation_is_deterministic(rt):
    params, st, T, res = lidar_case(3, 64, 1024, 0.02)
    outs = []
    for _ in range(3):
        det = DetectionModule(rt, **params)
        det.projectScan(None, st, T)
        det.projectResiduals(res)
        det.applySegmentation()
        outs.append((det.label_mat.copy(), det.avg_residuals.copy()))
    for lm, avg in outs[1:]:
        assert np.array_equal(lm, outs[0][0]) and np.array_equal(avg.view(np.uint64), outs[0][1].view(np.uint64))


def test_segmentation_residuals_straight_from_the_engine(rt, oracle):
    # projectResidualsFrom(engine): the residual cloud of the last align is built on the device and never read back;
    # it must give exactly what residualImage() -> host -> projectResiduals() gives, and what the oracle gives for it
    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng

    w = synth.make_world()
    src, tgt = synth.scan(1, 16, 256, w), synth.scan(0, 16, 256, w)
    g = ng.NanoGICP(rt)
    g.setInputSource(ng.PointCloud(rt, src))
    g.setInputTarget(ng.PointCloud(rt, tgt))
    g.align()
    params, st, T, _ = lidar_case(3, 64, 1024, 0.02)
    a0, a1 = -np.pi / 2, np.pi / 2
    img = g.residualImage(1024, 64, a0, a1)
    assert (img[..., 3] > 0).sum() > 100  # the test really feeds residuals
    two_step = DetectionModule(rt, **params)
    two_step.projectScan(None, st, T)
    two_step.projectResiduals(img)
    two_step.applySegmentation()
    fused = DetectionModule(rt, **params)
    fused.projectScan(None, st, T)
    fused.projectResidualsFrom(g, a0, a1)
    fused.applySegmentation()
    assert np.array_equal(fused.label_mat, two_step.label_mat)
    assert np.array_equal(fused.avg_residuals.view(np.uint64), two_step.avg_residuals.view(np.uint64))
    assert fused.avg_residuals.any()
    o = oracle.segment_scan(oracle.SegParams(**params), st, T, img[..., 3])
    assert_same(fused, o)
    # before any align the engine has no residuals to give
    fresh = ng.NanoGICP(rt)
    det = DetectionModule(rt, **params)
    det.projectScan(None, st, T)
    det.projectResidualsFrom(fresh)
    with pytest.raises(B.DdloError) as e:
        det.applySegmentation()
    assert e.value.code == -5  # DDLO_E_NOT_READY


def test_cpp_detection_shim_matches(rt, oracle, tmp_path):
    """tests/cpp/segmentation_protocol.cpp drives ddlo_shim::DetectionModule (the C++ header with the reference's
    method and member names) through OdomNode::applySegmentation's calls; it must write exactly what the Python mirror
    and the oracle give for the same scan."""
    import subprocess
    from pathlib import Path

    exe = Path(__file__).resolve().parent / "cpp" / "_build" / "segmentation_protocol"
    if not exe.exists():
        import __graft_entry__ as ge

        ge.build_cpp_tests()
    params, st, T, res = lidar_case(3, 64, 1024, 0.02)
    H, W = params["rows"], params["cols"]
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(src, "wb") as fh:
        fh.write(np.array([H, W, params["ground_rows"]], dtype=np.int32).tobytes())
        fh.write(np.array([params["ang_bottom"], params["minimum_range"], params["sensor_mount_angle"], params["max_distance"]],
                          dtype=np.float32).tobytes())
        fh.write(np.ascontiguousarray(T.T, dtype=np.float32).tobytes())  # column-major
        fh.write(np.ascontiguousarray(st, dtype=np.float32).tobytes())
        fh.write(np.ascontiguousarray(res, dtype=np.float32).tobytes())
    out = subprocess.run([str(exe), str(src), str(dst)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    raw = dst.read_bytes()
    count = int(np.frombuffer(raw, dtype=np.int32, count=1)[0])
    labels = np.frombuffer(raw, dtype=np.int32, count=H * W, offset=4).reshape(H, W)
    ground = np.frombuffer(raw, dtype=np.int8, count=H * W, offset=4 + 4 * H * W).reshape(H, W)
    avg = np.frombuffer(raw, dtype=np.float64, count=count, offset=4 + 5 * H * W)
    det, o = run_both(rt, oracle, params, st, T, res)
    assert_same(det, o)
    assert count == det.label_count_ and np.array_equal(labels, det.label_mat) and np.array_equal(ground, det.ground_mat)
    assert np.array_equal(avg[1:].view(np.uint64), det.avg_residuals[1:].view(np.uint64))
    f = out.stdout.split()
    assert int(f[1]) == det.getSegmentsCount() and int(f[3]) == len(det.getGroundIndices())
    assert int(f[5]) == int(((det.label_mat > 0) & (det.label_mat != INVALID_SEGMENT)).sum())


def test_segmentation_transforms_sensor_frame_scan_on_device(rt, oracle):
    # projectScan(cloud_in, None, T): OdomNode::transformScans (odom.cc:957-963) happens on the device, in the float
    # arithmetic of pcl::transformPointCloud; NaN pixels stay NaN
    frame = 3
    sc = synth.organized_scan(frame, 64, 1024, dropout=0.02)
    T = synth.pose(frame).astype(np.float32)
    params, _, _, res = lidar_case(frame, 64, 1024, 0.02)
    det = DetectionModule(rt, **params)
    det.projectScan(sc, None, T)
    det.projectResiduals(res)
    det.applySegmentation()
    o = oracle.segment_scan(oracle.SegParams(**{**params, "scan_in_sensor_frame": 1}), sc, T, res)
    assert_same(det, o)
    assert det.getSegmentsCount() > 3
    # and it is the same scene as the host-transformed one: ranges agree to float rounding of the transform
    _, st, _, _ = lidar_case(frame, 64, 1024, 0.02)
    det2 = DetectionModule(rt, **params)
    det2.projectScan(None, st, T)
    det2.applySegmentation()
    assert np.allclose(det.range_mat, det2.range_mat, rtol=0, atol=1e-4)


# ------------------------------------------------------------------ against the reference's own DetectionModule code
@pytest.fixture(scope="module")
def refdet(oracle):
    from oracle import refdet as rd

    if not rd.available():
        pytest.skip("oracle/_ref/libdetection_ref.so not built (needs /root/reference)")
    rd.lib()
    return rd


def assert_same_as_reference(det, r):
    assert det.label_count_ == r["label_count"]
    assert np.array_equal(det.label_mat, r["label_mat"])
    assert np.array_equal(det.ground_mat, r["ground_mat"])
    assert np.array_equal(det.range_mat.view(np.uint32), r["range_mat"].view(np.uint32))
    assert np.array_equal(det.avg_residuals[1:].view(np.uint64), r["avg_residuals"][1:].view(np.uint64))


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_segmentation_matches_reference_detection_code(rt, oracle, refdet, variant):
    """the CUDA stage against the reference's own groundRemoval / cloudSegmentation / labelComponents (oracle/refdet.py),
    bit for bit; the oracle's borderline count guards against a slope test that a different atan2 could flip"""
    import segmentation_cases as cases

    params, st, T, res = cases.reference_case(variant)
    assert oracle.segment_scan(oracle.SegParams(**params), st, T, res)["borderline"] == 0 or variant == 1
    det = DetectionModule(rt, **params)
    det.projectScan(None, st, T)
    det.projectResiduals(res)
    det.applySegmentation()
    assert_same_as_reference(det, refdet.segment(params, st, T, res))


def test_segmentation_matches_reference_golden_fixture(rt):
    """the committed answers of the reference's code (tests/golden/segmentation_reference.npz): needs neither the oracle
    nor /root/reference"""
    from pathlib import Path

    g = np.load(Path(__file__).resolve().parent / "golden" / "segmentation_reference.npz")
    params = {k: (float(v) if k in ("theta", "max_delta_z", "max_elevation") else int(v)) for k, v in zip(g["param_names"], g["param_values"])}
    det = DetectionModule(rt, **params)
    det.projectScan(None, g["scan_t"], g["T"])
    det.projectResiduals(g["residuals"])
    det.applySegmentation()
    assert_same_as_reference(det, dict(label_count=int(g["label_count"]), label_mat=g["label_mat"], ground_mat=g["ground_mat"],
                                       range_mat=g["range_mat"], avg_residuals=g["avg_residuals"]))


@pytest.mark.parametrize("with_residuals", [True, False])
def test_segmentation_early_decisions_equal_full_replay(rt, oracle, monkeypatch, with_residuals):
    """Segments whose fate cannot depend on the push order are decided without replaying the queue; with
    DDLO_SEG_NO_SHORTCUT every segment is replayed.  Both must give the same images, on inputs that are hard for the
    shortcut: exact zeros in z, equal maxima, tiny segments, thresholds that split the segments about evenly."""
    rng = np.random.default_rng(17)
    cases = []
    for frame, over in ((3, {}), (40, dict(max_delta_z=1.0, max_distance=15.0)), (21, dict(max_elevation=0.3, min_delta_z=0.5)),
                        (8, dict(valid_point_num=3, valid_line_num=1, min_line_num=1, theta=0.4))):
        params, st, T, res = lidar_case(frame, 64, 1024, 0.05, **over)
        st = st.copy()
        flat = st.reshape(-1, 4)
        pick = rng.choice(len(flat), 3000, replace=False)
        flat[pick[:1500], 2] = 0.0                                   # exact zeros: never a minimum, always a maximum candidate
        flat[pick[1500:], 2] = np.round(flat[pick[1500:], 2], 1)     # many exactly equal heights
        cases.append((params, st, T, res))
    for params, st, T, res in cases:
        outs = []
        for no_shortcut in (False, True):
            if no_shortcut:
                monkeypatch.setenv("DDLO_SEG_NO_SHORTCUT", "1")
            else:
                monkeypatch.delenv("DDLO_SEG_NO_SHORTCUT", raising=False)
            det = DetectionModule(rt, **params)
            det.projectScan(None, st, T)
            if with_residuals:
                det.projectResiduals(res)
            det.applySegmentation()
            outs.append(det)
        monkeypatch.delenv("DDLO_SEG_NO_SHORTCUT", raising=False)
        assert outs[0].label_count_ == outs[1].label_count_ and np.array_equal(outs[0].label_mat, outs[1].label_mat)
        assert np.array_equal(outs[0].avg_residuals.view(np.uint64), outs[1].avg_residuals.view(np.uint64))
        o = oracle.segment_scan(oracle.SegParams(**params), st, T, res if with_residuals else None)
        if o["borderline"] == 0:
            assert_same(outs[0], o)


def test_segmentation_unordered_residual_sums_mode(rt, oracle):
    """opt-in unordered_residual_sums: same label image, averages equal to the push-order float sums up to the rounding
    of those sums (1e-5 relative is generous for a few thousand terms)"""
    for frame in (3, 40):
        params, st, T, res = lidar_case(frame, 64, 1024, 0.02)
        exact = DetectionModule(rt, **params)
        exact.projectScan(None, st, T)
        exact.projectResiduals(res)
        exact.applySegmentation()
        fast = DetectionModule(rt, **{**params, "unordered_residual_sums": 1})
        fast.projectScan(None, st, T)
        fast.projectResiduals(res)
        fast.applySegmentation()
        assert fast.label_count_ == exact.label_count_ > 4 and np.array_equal(fast.label_mat, exact.label_mat)
        assert np.allclose(fast.avg_residuals, exact.avg_residuals, rtol=1e-5, atol=0.0)
        assert fast.avg_residuals[1:].all()


def test_segmentation_fuzz_against_oracle(rt, oracle):
    """random small range images (piecewise-smooth depth with steps, holes, exact zeros and repeated heights) and random
    parameters: the early decisions, the queue replay and the statistics must agree with the oracle on every one"""
    rng = np.random.default_rng(2024)
    compared = accepted_total = 0
    for trial in range(40):
        H, W = int(rng.integers(6, 48)), int(rng.integers(12, 160))
        el = np.linspace(0.35, -0.35, H)[:, None]
        az = np.linspace(-1.0, 1.0, W)[None, :]
        depth = 6.0 + 3.0 * np.sin(az * rng.uniform(1, 5) + rng.uniform(0, 6)) + 2.0 * np.cos(el * rng.uniform(2, 9))
        for _ in range(int(rng.integers(0, 6))):  # foreground blobs: range steps
            r0, c0 = int(rng.integers(0, H)), int(rng.integers(0, W))
            hh, ww = int(rng.integers(2, max(3, H // 2))), int(rng.integers(2, max(3, W // 3)))
            depth[r0:r0 + hh, c0:c0 + ww] -= rng.uniform(1.0, 4.0)
        depth = np.maximum(depth, 0.6) + rng.normal(0.0, 0.004, (H, W))
        s = np.empty((H, W, 4), dtype=np.float32)
        s[..., 0] = depth * np.cos(el) * np.cos(az)
        s[..., 1] = depth * np.cos(el) * np.sin(az)
        s[..., 2] = depth * np.sin(el)
        s[..., 3] = 1.0
        s[rng.random((H, W)) < rng.uniform(0.0, 0.15)] = np.nan
        zero = rng.random((H, W)) < 0.02
        s[zero, 2] = 0.0
        if rng.random() < 0.5:
            s[..., 2] = np.round(s[..., 2], 1)  # many equal heights: ties at the maximum and at the minimum
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
        params = dict(rows=H, cols=W, ground_rows=int(rng.integers(0, H)), window_row_min=int(rng.integers(0, 3)), window_row_max=H - 1 - int(rng.integers(0, 3)),
                      window_col_min=int(rng.integers(0, 4)), window_col_max=W - 1 - int(rng.integers(0, 4)), ang_bottom=float(rng.uniform(5, 30)),
                      minimum_range=float(rng.uniform(0.3, 2.0)), sensor_mount_angle=float(rng.uniform(-5, 5)), ground_angle_threshold=float(rng.uniform(2, 15)),
                      theta=float(rng.uniform(0.05, 1.2)), valid_point_num=int(rng.integers(1, 12)), valid_line_num=int(rng.integers(0, 4)),
                      min_line_num=int(rng.integers(0, 4)), min_delta_z=float(rng.uniform(-0.5, 0.3)), max_delta_z=float(rng.uniform(0.2, 4.0)),
                      max_distance=float(rng.uniform(4.0, 12.0)), max_elevation=float(rng.uniform(-0.5, 3.0)))
        res = (rng.random((H, W)) * (rng.random((H, W)) < 0.8)).astype(np.float32) if rng.random() < 0.8 else None
        o = oracle.segment_scan(oracle.SegParams(**params), s, T, res)
        if o["borderline"]:
            continue
        det = DetectionModule(rt, **params)
        det.projectScan(None, s, T)
        if res is not None:
            det.projectResiduals(res)
        det.applySegmentation()
        assert_same(det, o)
        compared += 1
        accepted_total += det.getSegmentsCount()
    assert compared >= 25 and accepted_total >= 40
