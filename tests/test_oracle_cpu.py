"""CPU-only checks of the oracle (test infrastructure): against brute force, against the reference's
own vendored nanoflann where it compiled (oracle/_ref), against numpy for the restated Eigen
arithmetic, against a literal numpy restatement of the reference's 4x4 formulas, and against the
committed golden fixtures."""
from pathlib import Path

import numpy as np
import pytest

from dynamic_direct_lidar_odometry_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def pair():
    w = synth.make_world()
    return synth.scan(1, 16, 128, w), synth.scan(0, 16, 128, w)


# ------------------------------------------------------------------------------------------------ kNN
def test_kdtree_equals_bruteforce(oracle, pair):
    src, tgt = pair
    for k in (1, 7, 20):
        idx, d2 = oracle.Cloud(tgt).build_tree().knn(src, k)
        bidx, bd2 = oracle.knn_bruteforce(tgt, src, k)
        assert np.array_equal(idx, bidx)
        assert np.array_equal(d2.view(np.uint32), bd2.view(np.uint32))


def test_kdtree_ties_and_duplicates(oracle):
    g = np.arange(5, dtype=np.float32)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([lat, lat[::2]], 0)[np.random.default_rng(0).permutation(125 + 63)]
    idx, d2 = oracle.Cloud(pts).build_tree().knn(pts, 9)
    bidx, bd2 = oracle.knn_bruteforce(pts, pts, 9)
    assert np.array_equal(idx, bidx) and np.array_equal(d2, bd2)
    # ties are broken by index: among equal distances the indices ascend
    same = d2[:, 1:] == d2[:, :-1]
    assert (idx[:, 1:][same] > idx[:, :-1][same]).all()


def test_canonical_vs_reference_nanoflann(oracle):
    """The reference's own kd-tree (nanoflann 1.3.2, leaf 100, compiled from /root/reference) returns
    bit-identical squared distances; indices agree wherever the k+1 smallest distances are distinct
    (its order among exact ties is traversal order, nanoflann_impl.hpp:207-231)."""
    if not oracle.load_reference_nanoflann():
        pytest.skip("oracle/_ref/libnanoflann_ref.so not built (needs /root/reference)")
    w = synth.make_world()
    tgt = synth.scan(0, 32, 512, w)
    src = synth.scan(1, 32, 512, w)[::3]
    k = 20
    cidx, cd2 = oracle.Cloud(tgt).build_tree(oracle.BACKEND_CANONICAL).knn(src, k + 1)
    ridx, rd2 = oracle.Cloud(tgt).build_tree(oracle.BACKEND_NANOFLANN_REF).knn(src, k)
    assert np.array_equal(rd2.view(np.uint32), cd2[:, :k].view(np.uint32))
    distinct = (np.diff(cd2, axis=1) > 0).all(axis=1)
    assert distinct.mean() > 0.9
    assert np.array_equal(ridx[distinct], cidx[distinct, :k])
    # and on the rows with ties the index SETS still agree up to the tied entries
    for r in np.flatnonzero(~distinct)[:200]:
        assert sorted(rd2[r]) == sorted(cd2[r, :k])


def test_knn_golden(oracle):
    g = np.load(GOLDEN / "knn_small.npz")
    idx, d2 = oracle.Cloud(g["tgt"]).build_tree().knn(g["src"], 20)
    assert np.array_equal(idx, g["idx20"]) and np.array_equal(d2, g["d20"])
    if "ref_idx20" in g:  # what the reference's nanoflann answered when the fixture was made
        assert np.array_equal(g["ref_d20"], g["d20"])
        distinct = (np.diff(g["d20"], axis=1) > 0).all(axis=1)
        assert np.array_equal(g["ref_idx20"][distinct], g["idx20"][distinct])


# ------------------------------------------------------------------------- restated Eigen arithmetic
def test_math_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * rng.uniform(1e-3, 10)
        A = a @ a.T
        w, V = oracle.math_sym_eig3(A)
        wn = np.linalg.eigvalsh(A)[::-1]
        assert np.allclose(w, wn, rtol=1e-12, atol=1e-14 * wn[0])
        assert np.allclose(V @ np.diag(w) @ V.T, A, rtol=0, atol=1e-13 * np.abs(A).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-14)
        B = A + 0.05 * np.trace(A) * np.eye(3)  # condition number <= ~20: the cofactor inverse is accurate to ~1e-14
        assert np.allclose(oracle.math_inverse3(B), np.linalg.inv(B), rtol=1e-11, atol=1e-13 / np.trace(A))
        j = rng.normal(size=(12, 6)) * np.array([5, 5, 5, 1, 1, 1])
        H = j.T @ j + 1e-6 * np.eye(6)
        b = rng.normal(size=6)
        xs, xn = oracle.math_ldlt6_solve(H, b), np.linalg.solve(H, b)
        assert np.linalg.norm(xs - xn) <= 1e-12 * np.linalg.cond(H) * np.linalg.norm(xn)
        om = rng.normal(size=3) * rng.choice([1e-7, 1e-2, 1.0])
        th = np.linalg.norm(om)
        K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        Rr = np.eye(3) + (np.sin(th) / th) * K + ((1 - np.cos(th)) / th**2) * K @ K if th > 0 else np.eye(3)
        assert np.allclose(oracle.math_so3_exp(om), Rr, atol=1e-14)


def _plane_cov_numpy(nb):
    """nano_gicp_impl.hpp:398-436 with numpy: centre, N N^T / k, SVD, U diag(1,1,1e-3) V^T."""
    n = nb.astype(np.float64).T  # 3 x k
    n = n - n.mean(axis=1, keepdims=True)
    cov = n @ n.T / n.shape[1]
    U, s, Vt = np.linalg.svd(cov)
    return U @ np.diag([1.0, 1.0, 1e-3]) @ Vt, cov


def test_covariances_vs_numpy(oracle, pair):
    _, tgt = pair
    k = 20
    c = oracle.Cloud(tgt).build_tree()
    covs = c.covariances(k, oracle.REG_PLANE)
    raw = c.covariances(k, oracle.REG_NONE)
    idx, _ = c.knn(tgt, k)
    checked = 0
    for i in range(0, len(tgt), 7):
        want, cov = _plane_cov_numpy(tgt[idx[i], :3])
        assert np.allclose(raw[i, :3, :3], cov, rtol=1e-12, atol=1e-18)
        w = np.linalg.eigvalsh(cov)
        if (w[1] - w[0]) > 1e-6 * w[2]:  # smallest eigenvector well defined
            assert np.allclose(covs[i, :3, :3], want, atol=1e-9)
            checked += 1
        assert (covs[i, 3, :] == 0).all() and (covs[i, :, 3] == 0).all()
    assert checked > 100
    g = np.load(GOLDEN / "cov_small.npz")
    for m in range(5):
        assert np.allclose(c.covariances(k, m), g[f"method{m}"], rtol=1e-12, atol=1e-15)


def _linearize_numpy(src, tgt, cov_a, cov_b, T, corr):
    """nano_gicp_impl.hpp:262-339 spelled out with 4x4 matrices, exactly as the reference writes it."""
    H = np.zeros((6, 6))
    b = np.zeros(6)
    err = 0.0
    M_all = np.zeros((len(src), 4, 4))
    for i, j in enumerate(corr):
        if j < 0:
            continue
        RCR = cov_b[j] + T @ cov_a[i] @ T.T
        RCR[3, 3] = 1.0
        M = np.linalg.inv(RCR)
        M[3, 3] = 0.0
        M_all[i] = M
        mean_a = np.array([*src[i, :3].astype(np.float64), 1.0])
        mean_b = np.array([*tgt[j, :3].astype(np.float64), 1.0])
        ta = T @ mean_a
        e = mean_b - ta
        err += e @ M @ e
        J = np.zeros((4, 6))
        x = ta[:3]
        J[:3, :3] = [[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]]
        J[:3, 3:] = -np.eye(3)
        H += J.T @ M @ J
        b += J.T @ M @ e
    return err, H, b, M_all


def test_linearize_vs_numpy(oracle):
    g = np.load(GOLDEN / "gicp_small.npz")
    src, tgt = g["src"], g["tgt"]
    eng = oracle.NanoGICP()
    eng.setInputSource(oracle.Cloud(src))
    eng.setInputTarget(oracle.Cloud(tgt))
    eng.setSourceCovariances(g["src_covs"])
    eng.setTargetCovariances(g["tgt_covs"])
    e, H, b = eng.linearize(g["T"])
    corr, sqd = eng.correspondences()
    # correspondences: float32 transform (pairwise order) then nearest neighbour
    Tf = g["T"].astype(np.float32)
    q = np.stack([(Tf[r, 0] * src[:, 0] + Tf[r, 1] * src[:, 1]) + (Tf[r, 2] * src[:, 2] + Tf[r, 3]) for r in range(3)], 1).astype(np.float32)
    bidx, bd2 = oracle.knn_bruteforce(tgt, q, 1)
    assert np.array_equal(corr, bidx[:, 0]) and np.array_equal(sqd, bd2[:, 0])
    ne, nH, nb, nM = _linearize_numpy(src, tgt, g["src_covs"], g["tgt_covs"], g["T"], corr)
    assert abs(e - ne) < 1e-9 * abs(ne)
    assert np.linalg.norm(H - nH) < 1e-9 * np.linalg.norm(nH)
    assert np.linalg.norm(b - nb) < 1e-9 * np.linalg.norm(nb)
    assert np.allclose(eng.mahalanobis(), nM, rtol=1e-8, atol=1e-10)
    assert abs(e - float(g["err"])) <= 1e-12 * abs(e) and np.allclose(H, g["H"], rtol=1e-12) and np.allclose(b, g["b"], rtol=1e-11, atol=1e-9)
    e2 = eng.compute_error(g["T2"])
    assert abs(e2 - float(g["err2"])) <= 1e-12 * abs(e2)


def _lm_numpy(eng, guess, max_iter=64, rot_eps=2e-3, trans_eps=5e-4, factor=1e-9):
    """lsq_registration_impl.hpp:96-232 in numpy on top of the oracle's linearize / compute_error."""
    def conv(d):
        return max(np.abs(d[:3, :3] - np.eye(3)).max() / rot_eps, np.abs(d[:3, 3]).max() / trans_eps) < 1

    x0 = guess.astype(np.float64).copy()
    lam, converged, it_last, nl, ne = -1.0, False, 0, 0, 0
    for it in range(max_iter):
        if converged:
            break
        it_last = it
        y0, H, b = eng.linearize(x0)
        nl += 1
        if lam < 0:
            lam = factor * np.abs(np.diag(H)).max()
        nu, ok = 2.0, False
        for _ in range(10):
            d = np.linalg.solve(H + lam * np.eye(6), -b)
            th2 = d[:3] @ d[:3]
            th = np.sqrt(th2)
            imag, real = (0.5 - th2 / 48 + th2 * th2 / 3840, 1 - th2 / 8 + th2 * th2 / 384) if th2 < 1e-10 else (np.sin(th / 2) / th, np.cos(th / 2))
            w, (x, y, z) = real, imag * d[:3]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            delta = np.eye(4)
            delta[:3, :3], delta[:3, 3] = R, d[3:]
            xi = delta @ x0
            yi = eng.compute_error(xi)
            ne += 1
            rho = (y0 - yi) / (d @ (lam * d - b))
            if rho < 0:
                if conv(delta):
                    ok = True
                    break
                lam, nu = nu * lam, 2 * nu
                continue
            x0, lam, ok = xi, lam * max(1 / 3, 1 - (2 * rho - 1) ** 3), True
            break
        if not ok:
            break
        converged = conv(delta)
    return x0, converged, it_last, nl, ne


def test_align_vs_numpy_lm_and_golden(oracle):
    g = np.load(GOLDEN / "gicp_small.npz")
    eng = oracle.NanoGICP()
    eng.setInputSource(oracle.Cloud(g["src"]))
    eng.setInputTarget(oracle.Cloud(g["tgt"]))
    r = eng.align()
    ref = oracle.NanoGICP()
    ref.setInputSource(oracle.Cloud(g["src"]))
    ref.setInputTarget(oracle.Cloud(g["tgt"]))
    ref.calculateSourceCovariances()
    ref.calculateTargetCovariances()
    x, conv, it, nl, ne = _lm_numpy(ref, np.eye(4))
    assert (r.converged, r.iterations, r.n_linearize, r.n_compute_error) == (conv, it, nl, ne)
    assert np.allclose(r.T, x.astype(np.float32), atol=1e-6)
    assert np.array_equal(r.T, g["lm_T"]) and tuple(g["lm_meta"]) == (r.converged, r.iterations, r.n_linearize, r.n_compute_error)
    gt = np.linalg.inv(synth.pose(0)) @ synth.pose(1)
    assert np.abs(r.T[:3, 3] - gt[:3, 3]).max() < 0.05
    eng.setOptimizer(oracle.OPT_GAUSS_NEWTON)
    rg = eng.align()
    assert np.array_equal(rg.T, g["gn_T"])


def test_swap_and_covariance_reuse(oracle, pair):
    """swapSourceAndTarget keeps trees and covariances with their clouds (nano_gicp_impl.hpp:98-106)."""
    src, tgt = pair
    e = oracle.NanoGICP()
    S, T = oracle.Cloud(src), oracle.Cloud(tgt)
    e.setInputSource(S)
    e.setInputTarget(T)
    e.align()
    cs, ct = e.getSourceCovariances(), e.getTargetCovariances()
    e.swapSourceAndTarget()
    assert np.array_equal(e.getSourceCovariances(), ct) and np.array_equal(e.getTargetCovariances(), cs)
    r = e.align()  # now registers tgt onto src, covariances are not recomputed
    fwd = oracle.NanoGICP()
    fwd.setInputSource(oracle.Cloud(tgt))
    fwd.setInputTarget(oracle.Cloud(src))
    assert np.array_equal(r.T, fwd.align().T)
    # setInputSource with a new cloud clears only the source covariances
    e.setInputSource(oracle.Cloud(src[::2]))
    assert len(e.getSourceCovariances()) == 0 and len(e.getTargetCovariances()) == len(src)
