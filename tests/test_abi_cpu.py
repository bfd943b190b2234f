"""CPU-only checks of the product's C-ABI library: it loads, exports every symbol the headers
declare, reports errors instead of falling back when there is no CUDA device, and the host-callable
copies of the DEVICE arithmetic (the same __host__ __device__ functions the kernels run) agree with
numpy and with the oracle's restatement."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ddlo_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(ddlo_lib):
    names = _declared(ROOT / "include" / "ddlo_gicp.h") + _declared(ROOT / "include" / "ddlo_gicp_testing.h")
    assert len(names) > 55
    missing = [n for n in names if not hasattr(ddlo_lib, n)]
    assert not missing, missing
    assert ddlo_lib.ddlo_abi_version() == 2


def test_binding_covers_header():
    from dynamic_direct_lidar_odometry_b200 import binding

    declared = set(_declared(ROOT / "include" / "ddlo_gicp.h"))
    bound = set(binding.SIGNATURES) | set(binding._SPECIAL)
    assert declared == bound, declared ^ bound


def test_defaults_are_the_references(ddlo_lib):
    from dynamic_direct_lidar_odometry_b200 import binding

    p = binding.Params()
    assert ddlo_lib.ddlo_params_default(C.byref(p)) == 0
    # nano_gicp_impl.hpp:58-62, lsq_registration_impl.hpp:53-61
    assert (p.k_correspondences, p.regularization_method, p.max_iterations, p.optimizer, p.lm_max_iterations) == (20, 3, 64, 1, 10)
    assert p.max_correspondence_distance == float(np.finfo(np.float32).max)
    assert (p.transformation_epsilon, p.rotation_epsilon, p.lm_init_lambda_factor) == (5e-4, 2e-3, 1e-9)


def test_no_cpu_fallback(ddlo_lib):
    """Without a CUDA device the library refuses to work (no CPU path hides behind the ABI)."""
    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng

    n = C.c_int(-1)
    rc = ddlo_lib.ddlo_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ng.DdloError) as e:
        ng.Runtime(0)
    assert e.value.code == -2 and "CUDA" in str(e.value)
    assert ddlo_lib.ddlo_gicp_align(None, None, None) == -1
    assert b"null" in ddlo_lib.ddlo_last_error()


def test_product_does_not_touch_the_oracle():
    """only tests/, smoke() and bench.py may import or load anything under oracle/"""
    pkg = ROOT / "dynamic_direct_lidar_odometry_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        text = f.read_text()
        assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text, f


def _sym6(A):
    return np.array([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]])


def _unsym6(s):
    return np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])


def test_device_math_on_host(ddlo_lib, oracle):
    from dynamic_direct_lidar_odometry_b200 import binding as B

    rng = np.random.default_rng(2)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * rng.uniform(1e-3, 10)
        A = a @ a.T
        w, V = np.empty(3), np.empty(9)
        ddlo_lib.ddlo_math_sym3_eig(B.ptr(_sym6(A)), B.ptr(w), B.ptr(V))
        V = V.reshape(3, 3)
        assert np.allclose(w, np.linalg.eigvalsh(A)[::-1], rtol=1e-12, atol=1e-14 * w[0])
        assert np.allclose(V @ np.diag(w) @ V.T, A, rtol=0, atol=1e-13 * np.abs(A).max())
        out = np.empty(6)
        Bm = A + 0.05 * np.trace(A) * np.eye(3)
        ddlo_lib.ddlo_math_sym3_inverse(B.ptr(_sym6(Bm)), B.ptr(out))
        assert np.allclose(_unsym6(out), np.linalg.inv(Bm), rtol=1e-11, atol=1e-13 / np.trace(A))
        # regularisation: same numbers as the oracle's restatement for every mode
        for m in range(5):
            ddlo_lib.ddlo_math_regularize(B.ptr(_sym6(A)), m, B.ptr(out))
            if m == 3:
                U, s, Vt = np.linalg.svd(A)
                if s[1] - s[2] > 1e-6 * s[0]:
                    assert np.allclose(_unsym6(out), U @ np.diag([1, 1, 1e-3]) @ Vt, atol=1e-9)
            elif m == 0:
                assert np.allclose(_unsym6(out), A, rtol=0, atol=0)
            elif m == 4:
                # ((C + 1e-3 I)^-1 / |.|_F)^-1 inverts twice: compare on the well-conditioned Bm
                ddlo_lib.ddlo_math_regularize(B.ptr(_sym6(Bm)), m, B.ptr(out))
                Ci = np.linalg.inv(Bm + 1e-3 * np.eye(3))
                assert np.allclose(_unsym6(out), np.linalg.inv(Ci / np.linalg.norm(Ci)), rtol=1e-10)
            elif m == 1:
                ww, VV = np.linalg.eigh(A)
                assert np.allclose(_unsym6(out), VV @ np.diag(np.maximum(ww, 1e-3)) @ VV.T, rtol=1e-9, atol=1e-12)
            else:
                ww, VV = np.linalg.eigh(A)
                assert np.allclose(_unsym6(out), VV @ np.diag(np.maximum(ww / ww.max(), 1e-3)) @ VV.T, rtol=1e-9, atol=1e-12)
        j = rng.normal(size=(12, 6)) * np.array([5, 5, 5, 1, 1, 1])
        H = np.ascontiguousarray(j.T @ j + 1e-6 * np.eye(6))
        b = rng.normal(size=6)
        x1, x2 = np.empty(6), np.empty(6)
        ddlo_lib.ddlo_math_ldlt6_solve(B.ptr(H), B.ptr(b), B.ptr(x1))
        ddlo_lib.ddlo_math_ldlt6_solve_fast(B.ptr(H), B.ptr(b), B.ptr(x2))
        xn = np.linalg.solve(H, b)
        tol = 1e-12 * np.linalg.cond(H) * np.linalg.norm(xn)
        assert np.linalg.norm(x1 - xn) <= tol and np.linalg.norm(x2 - xn) <= tol
        assert np.allclose(x1, oracle.math_ldlt6_solve(H, b), rtol=1e-9, atol=1e-15)
        om = rng.normal(size=3) * rng.choice([1e-7, 1e-2, 1.0])
        R = np.empty(9)
        ddlo_lib.ddlo_math_so3_exp(B.ptr(om), B.ptr(R))
        assert np.allclose(R.reshape(3, 3), oracle.math_so3_exp(om), atol=1e-15)
    # an indefinite matrix takes the pivoting fall-back of the fast path
    H = np.diag([1.0, -2.0, 3.0, 4.0, 5.0, 6.0])
    b = np.arange(6.0)
    x = np.empty(6)
    ddlo_lib.ddlo_math_ldlt6_solve_fast(B.ptr(H), B.ptr(b), B.ptr(x))
    assert np.allclose(x, np.linalg.solve(H, b))


def test_synthetic_data_is_deterministic():
    from dynamic_direct_lidar_odometry_b200 import synth

    a, b = synth.scan(3, 8, 64), synth.scan(3, 8, 64)
    assert np.array_equal(a, b) and a.dtype == np.float32 and (a[:, 3] == 1).all()
    assert not np.array_equal(a, synth.scan(4, 8, 64))
    r = np.linalg.norm(a[:, :3], axis=1)
    assert r.min() > 0.5 and r.max() < 100 and np.isfinite(a).all()
    v = synth.voxel_filter(synth.scan(0, 16, 256), 0.25)
    assert 0 < len(v) < 16 * 256


def test_shard_range():
    from dynamic_direct_lidar_odometry_b200.sharding import shard_range

    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_segmentation_params_layout_matches_header():
    """ddlo_segmentation_params is mirrored three times (ctypes for the product, ctypes for the oracle, the oracle's own
    C struct): all of them must list the header's fields in the header's order and types."""
    import ctypes as C
    import re

    from dynamic_direct_lidar_odometry_b200 import binding as B
    from oracle import pyoracle

    def fields_of(text, name_open, name_close):
        body = text[text.index(name_open) + len(name_open): text.index(name_close)]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        body = re.sub(r"//[^\n]*", "", body)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ctype, names = decl.split(None, 1)
            out += [(n.strip(), ctype) for n in names.split(",")]
        return out

    header = fields_of((ROOT / "include" / "ddlo_gicp.h").read_text(), "typedef struct ddlo_segmentation_params {", "} ddlo_segmentation_params;")
    oracle_c = fields_of((ROOT / "oracle" / "oracle_segmentation.cpp").read_text(), "struct oracle_seg_params {", "};\n\n// scan_t")
    assert header == oracle_c and len(header) == 21
    ctype = {"int": C.c_int, "float": C.c_float}
    want = [(n, ctype[t]) for n, t in header]
    assert list(B.SegmentationParams._fields_) == want
    assert list(pyoracle.SegParams._fields_) == want
    assert set(B.SegmentationParams.DEFAULTS) == {n for n, _ in header} == set(pyoracle.SegParams.DEFAULTS)
    assert B.SegmentationParams.DEFAULTS.keys() == pyoracle.SegParams.DEFAULTS.keys()
    for k, v in B.SegmentationParams.DEFAULTS.items():
        assert abs(v - pyoracle.SegParams.DEFAULTS[k]) < 1e-12


def test_detection_module_argument_checks_need_no_gpu():
    # the host mirror validates shapes before anything reaches the library
    from dynamic_direct_lidar_odometry_b200.detection import DetectionModule

    det = DetectionModule(None, rows=8, cols=16, ground_rows=3)
    with pytest.raises(ValueError):
        det.projectScan(None, np.zeros((7, 16, 4), dtype=np.float32), np.eye(4))
    with pytest.raises(ValueError):
        det.projectResiduals(np.zeros((8, 15), dtype=np.float32))
    with pytest.raises(RuntimeError):
        det.applySegmentation()
    with pytest.raises(TypeError):
        DetectionModule(None, rowz=8)
    det.projectScan(np.zeros((8, 16, 4), dtype=np.float32), None, np.eye(4))
    assert det.params.scan_in_sensor_frame == 1
    det.projectScan(None, np.zeros((8, 16, 4), dtype=np.float32), np.eye(4))
    assert det.params.scan_in_sensor_frame == 0


def test_align_traffic_capture_belongs_to_this_build():
    """`roofline.traffic` of the bench line comes from an ncu capture stored in profiles/align_traffic.json; the file is
    stamped with a fingerprint of the kernel's sources and bench.py reports null when it does not match.  The committed
    capture must be the one of the committed sources."""
    import json
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root))
    import bench

    stored = json.loads((root / "profiles" / "align_traffic.json").read_text())
    assert stored["kernel_source_sha16"] == bench.kernel_source_sha16()
    assert stored["dram_bytes_per_launch"] == stored["dram_bytes_read"] + stored["dram_bytes_write"]
