"""Pins the oracle (oracle/oracle_gicp.cpp, the CPU restatement every parity test is checked against) to the
REFERENCE'S OWN nano_gicp engine: nano_gicp.hpp / lsq_registration.hpp / nanoflann.hpp and their impl/ and gicp/
headers compiled unmodified from /root/reference into oracle/_ref/libnano_gicp_ref.so (oracle/refgicp.py,
oracle/ref_nano_gicp_shim.cpp).  Eigen, PCL and Boost do not exist in this image, so that build uses the stand-in
headers of oracle/stub_include/: the reference's control flow and formulas run exactly as written
(update_correspondences, linearize, compute_error, calculate_covariances with its five regularisations, step_lm /
step_gn / is_converged, swapSourceAndTarget, the covariance hand-over), and only Eigen's dense arithmetic is
restated.  Hence integer results must be identical and floating-point results agree to rounding.

The library is prebuilt here (where /root/reference exists) and travels to the GPU box; where it is absent these
tests skip.
"""
import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import synth


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle import refgicp

    if not refgicp.available():
        pytest.skip("oracle/_ref/libnano_gicp_ref.so not built (needs /root/reference)")
    refgicp.lib()
    return refgicp


@pytest.fixture(scope="module")
def pair():
    w = synth.make_world()
    return synth.scan(1, 16, 256, w), synth.scan(0, 16, 256, w)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])  # NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS
def test_covariances_match_reference_engine(oracle, ref, pair, method):
    src, _ = pair
    k = 10
    r = ref.NanoGICP()
    r.setCorrespondenceRandomness(k)
    r.setRegularizationMethod(method)
    r.setInputSource(ref.Cloud(src))
    assert r.calculateSourceCovariances()
    want = r.getSourceCovariances()
    got = oracle.Cloud(src).build_tree(oracle.BACKEND_NANOFLANN_REF if oracle.load_reference_nanoflann() else oracle.BACKEND_CANONICAL).covariances(k, method)
    assert got.shape == want.shape == (len(src), 4, 4)
    assert np.all(want[:, 3, :] == 0) and np.all(want[:, :, 3] == 0)
    if method in (1, 2, 3):
        # U diag V^T of a symmetric PSD matrix: the null-space basis of an exactly rank-deficient neighbourhood is
        # not unique; such points (two smallest singular values within 1e-6 relative) are left out, SURVEY.md §7
        raw = oracle.Cloud(src).build_tree().covariances(k, 0)[:, :3, :3]
        w = np.linalg.eigvalsh(raw)
        ok = (w[:, 1] - w[:, 0]) > 1e-6 * np.maximum(w[:, 2], 1e-30)
        assert ok.mean() > 0.9
    else:
        ok = np.ones(len(src), bool)
    err = np.linalg.norm((got - want)[ok].reshape(ok.sum(), -1), axis=1) / np.maximum(np.linalg.norm(want[ok].reshape(ok.sum(), -1), axis=1), 1e-300)
    assert err.max() < 1e-9, err.max()


@pytest.mark.parametrize("corr_dist", [None, 0.5])
def test_linearize_matches_reference_engine(oracle, ref, pair, corr_dist):
    src, tgt = pair
    o, r = oracle.NanoGICP(), ref.NanoGICP()
    for e, mod in ((o, oracle), (r, ref)):
        e.setCorrespondenceRandomness(10)
        e.setInputSource(mod.Cloud(src))
        e.setInputTarget(mod.Cloud(tgt))
        if corr_dist is not None:
            e.setMaxCorrespondenceDistance(corr_dist)
    r.calculateSourceCovariances(); r.calculateTargetCovariances()
    o.setSourceCovariances(r.getSourceCovariances()); o.setTargetCovariances(r.getTargetCovariances())
    T = np.linalg.inv(synth.pose(0)) @ synth.pose(1)
    T[:3, 3] += [0.05, -0.03, 0.01]
    eo, Ho, bo = o.linearize(T)
    er, Hr, br = r.linearize(T)
    co, do = o.correspondences()
    cr, dr = r.correspondences()
    assert np.array_equal(co, cr) and np.array_equal(do.view(np.uint32), dr.view(np.uint32))
    if corr_dist is not None:
        assert (cr < 0).any() and (cr >= 0).any()
    assert rel(Ho, Hr) < 1e-10 and rel(bo, br) < 1e-10 and abs(eo - er) <= 1e-10 * abs(er)
    v = cr >= 0
    assert rel(o.mahalanobis()[v], r.mahalanobis()[v]) < 1e-10
    T2 = T.copy(); T2[:3, 3] += [0.01, 0.02, -0.01]
    assert abs(o.compute_error(T2) - r.compute_error(T2)) <= 1e-10 * abs(r.compute_error(T2))
    assert np.array_equal(o.getResiduals(), r.getResiduals())
    Tf = T.astype(np.float32)
    assert np.abs(o.getResidualVectors(Tf) - r.getResidualVectors(Tf)).max() < 1e-5


@pytest.mark.parametrize("optimizer", [1, 0])  # LevenbergMarquardt, GaussNewton
def test_align_matches_reference_engine(oracle, ref, pair, optimizer):
    src, tgt = pair
    o, r = oracle.NanoGICP(), ref.NanoGICP()
    for e, mod in ((o, oracle), (r, ref)):
        e.setCorrespondenceRandomness(10)
        e.setOptimizer(optimizer)
        e.setInputSource(mod.Cloud(src))
        e.setInputTarget(mod.Cloud(tgt))
    ro, rr = o.align(), r.align()
    assert (ro.converged, ro.iterations) == (rr.converged, rr.iterations)
    assert rr.converged and rr.iterations >= 1
    assert np.abs(ro.T.astype(np.float64) - rr.T).max() < 1e-6
    assert rel(ro.hessian, rr.hessian) < 1e-8
    co, do = o.correspondences()
    cr, dr = r.correspondences()
    assert np.array_equal(co, cr) and np.array_equal(do.view(np.uint32), dr.view(np.uint32))
    # a perturbed guess and an iteration cap
    guess = np.eye(4, dtype=np.float32); guess[:3, 3] = [0.2, -0.1, 0.05]
    for e in (o, r):
        e.setMaximumIterations(2)
    ro, rr = o.align(guess), r.align(guess)
    assert (ro.converged, ro.iterations) == (rr.converged, rr.iterations)
    assert np.abs(ro.T.astype(np.float64) - rr.T).max() < 1e-6


def test_odomnode_protocol_matches_reference_engine(oracle, ref):
    """The S2S -> S2M call sequence of OdomNode (odom.cc:480-532, 745-793) on both engines: swapSourceAndTarget,
    the shared source tree, covariance hand-over, injected submap covariances."""
    w = synth.make_world()
    scans = [synth.scan(f, 16, 256, w) for f in range(4)]

    def run(mod, is_ref):
        s2s, s2m = mod.NanoGICP(), mod.NanoGICP()
        for e in (s2s, s2m):
            e.setCorrespondenceRandomness(10)
        first = mod.Cloud(scans[0])
        s2s.setInputTarget(first)
        s2s.calculateTargetCovariances()
        s2s.setInputSource(first)
        s2s.calculateSourceCovariances()
        s2m.setInputTarget(first)
        s2m.setTargetCovariances(s2s.getSourceCovariances())
        T_world = np.eye(4)
        out = []
        for f in range(1, 4):
            cur = mod.Cloud(scans[f])
            s2s.setInputSource(cur)
            s2m.registerInputSource(cur)
            if is_ref:
                s2m.shareSourceTreeOf(s2s)
            else:
                s2m.clearSourceCovariances()
            r1 = s2s.align()
            guess = (T_world @ r1.T.astype(np.float64)).astype(np.float32)
            s2m.setSourceCovariances(s2s.getSourceCovariances())
            s2s.swapSourceAndTarget()
            r2 = s2m.align(guess)
            T_world = r2.T.astype(np.float64)
            out.append((r1, r2, s2m.getResiduals()))
        return out

    for (o1, o2, ores), (r1, r2, rres) in zip(run(oracle, False), run(ref, True)):
        assert (o1.converged, o1.iterations) == (r1.converged, r1.iterations)
        assert (o2.converged, o2.iterations) == (r2.converged, r2.iterations)
        assert np.abs(o1.T.astype(np.float64) - r1.T).max() < 1e-6
        assert np.abs(o2.T.astype(np.float64) - r2.T).max() < 1e-6
        assert np.allclose(ores, rres, rtol=1e-6, atol=1e-9)


def test_oracle_reproduces_reference_engine_fixture(oracle):
    """tests/golden/gicp_reference_engine.npz holds outputs of the reference's own engine (written by
    tests/golden/make_golden.py where /root/reference exists); the restatement must reproduce them anywhere."""
    from pathlib import Path

    f = Path(__file__).resolve().parent / "golden" / "gicp_reference_engine.npz"
    s = np.load(f)
    o = oracle.NanoGICP()
    o.setInputSource(oracle.Cloud(s["src"]))
    o.setInputTarget(oracle.Cloud(s["tgt"]))
    o.calculateSourceCovariances(); o.calculateTargetCovariances()
    raw = oracle.Cloud(s["tgt"]).build_tree().covariances(20, 0)[:, :3, :3]
    w = np.linalg.eigvalsh(raw)
    ok = (w[:, 1] - w[:, 0]) > 1e-6 * np.maximum(w[:, 2], 1e-30)
    for m in range(5):
        got = oracle.Cloud(s["tgt"]).build_tree().covariances(20, m)
        sel = ok if m in (1, 2, 3) else np.ones(len(ok), bool)
        assert rel(got[sel], s[f"cov_method{m}"][sel]) < 1e-9, m
    o.setSourceCovariances(s["src_covs"]); o.setTargetCovariances(s["tgt_covs"])
    e, H, b = o.linearize(s["T"])
    corr, sqd = o.correspondences()
    assert np.array_equal(corr, s["corr"]) and np.array_equal(sqd.view(np.uint32), s["sqd"].view(np.uint32))
    assert rel(H, s["H"]) < 1e-10 and rel(b, s["b"]) < 1e-10 and abs(e - float(s["err"])) <= 1e-10 * abs(float(s["err"]))
    assert abs(o.compute_error(s["T2"]) - float(s["err2"])) <= 1e-10 * abs(float(s["err2"]))
    for name, opt in (("lm", 1), ("gn", 0)):
        o.setOptimizer(opt)
        r = o.align()
        assert (int(r.converged), r.iterations) == tuple(int(v) for v in s[f"{name}_meta"])
        assert np.abs(r.T.astype(np.float64) - s[f"{name}_T"]).max() < 1e-6
