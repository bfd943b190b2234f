"""Regenerate tests/golden/*.npz from the CPU oracle.

The reference ships no tests, fixtures or golden vectors, so these are pins made by
this repository: the oracle's answers on small seeded inputs, frozen so that (a) a change of the
oracle shows up as a diff and (b) the CUDA path can be checked on the GPU box without trusting a
freshly built oracle.  Where the reference's own code could run (the vendored nanoflann and the nano_gicp
engine itself, oracle/_ref) its answers are stored too: knn_small.npz ref_* and gicp_reference_engine.npz.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    po.build()
    w = synth.make_world()
    src = synth.scan(1, 16, 128, w)
    tgt = synth.scan(0, 16, 128, w)

    # ---- kNN (a2)
    T = po.Cloud(tgt).build_tree(po.BACKEND_CANONICAL)
    idx20, d20 = T.knn(src, 20)
    idx1, d1 = T.knn(src, 1)
    knn = dict(src=src, tgt=tgt, idx20=idx20, d20=d20, idx1=idx1, d1=d1)
    if po.load_reference_nanoflann():
        R = po.Cloud(tgt).build_tree(po.BACKEND_NANOFLANN_REF)
        ridx20, rd20 = R.knn(src, 20)
        knn.update(ref_idx20=ridx20, ref_d20=rd20)
    np.savez_compressed(OUT / "knn_small.npz", **knn)

    # ---- covariances (a3), all regularisation modes
    covs = {f"method{m}": T.covariances(20, m) for m in range(5)}
    np.savez_compressed(OUT / "cov_small.npz", tgt=tgt, **covs)

    # ---- linearize / compute_error (a4-a6) and align (a7)
    eng = po.NanoGICP()
    S, Tg = po.Cloud(src), po.Cloud(tgt)
    eng.setInputSource(S)
    eng.setInputTarget(Tg)
    eng.calculateSourceCovariances()
    eng.calculateTargetCovariances()
    Tq = np.linalg.inv(synth.pose(0)) @ synth.pose(1)
    Tq[:3, 3] += [0.05, -0.03, 0.01]
    e, H, b = eng.linearize(Tq)
    corr, sqd = eng.correspondences()
    T2 = Tq.copy()
    T2[:3, 3] += [0.01, 0.02, -0.01]
    e2 = eng.compute_error(T2)
    res = {}
    for name, opt in (("lm", po.OPT_LEVENBERG_MARQUARDT), ("gn", po.OPT_GAUSS_NEWTON)):
        eng.setOptimizer(opt)
        r = eng.align()
        res.update({f"{name}_T": r.T, f"{name}_H": r.hessian, f"{name}_meta": np.array([r.converged, r.iterations, r.n_linearize, r.n_compute_error])})
    np.savez_compressed(OUT / "gicp_small.npz", src=src, tgt=tgt, src_covs=eng.getSourceCovariances(), tgt_covs=eng.getTargetCovariances(),
                        T=Tq, err=e, H=H, b=b, corr=corr, sqd=sqd, T2=T2, err2=e2, **res)
    # ---- the same quantities from the REFERENCE'S OWN ENGINE (oracle/refgicp.py: the reference's nano_gicp headers
    # compiled unmodified over Eigen/PCL stand-ins), where /root/reference is available to build it
    from oracle import refgicp as rg

    if rg.available():
        ref = rg.NanoGICP()
        ref.setInputSource(rg.Cloud(src))
        ref.setInputTarget(rg.Cloud(tgt))
        ref.calculateSourceCovariances()
        ref.calculateTargetCovariances()
        re_, rH, rb = ref.linearize(Tq)
        rcorr, rsqd = ref.correspondences()
        out = dict(src=src, tgt=tgt, src_covs=ref.getSourceCovariances(), tgt_covs=ref.getTargetCovariances(), T=Tq, err=re_, H=rH, b=rb,
                   corr=rcorr, sqd=rsqd, T2=T2, err2=ref.compute_error(T2))
        for m in range(5):
            rc = rg.NanoGICP()
            rc.setRegularizationMethod(m)
            rc.setInputSource(rg.Cloud(tgt))
            rc.calculateSourceCovariances()
            out[f"cov_method{m}"] = rc.getSourceCovariances()
        for name, opt in (("lm", po.OPT_LEVENBERG_MARQUARDT), ("gn", po.OPT_GAUSS_NEWTON)):
            ref.setOptimizer(opt)
            r = ref.align()
            out.update({f"{name}_T": r.T, f"{name}_H": r.hessian, f"{name}_meta": np.array([r.converged, r.iterations])})
        np.savez_compressed(OUT / "gicp_reference_engine.npz", **out)
    # ---- the segmentation stage from the REFERENCE'S OWN DetectionModule code (oracle/refdet.py), inputs included
    from oracle import refdet

    if refdet.available():
        sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
        import segmentation_cases as cases

        params, st, T, res = cases.golden_case()
        r = refdet.segment(params, st, T, res)
        np.savez_compressed(OUT / "segmentation_reference.npz", scan_t=st[..., :3].astype(np.float32), T=T, residuals=res,
                            param_names=np.array(list(params.keys())), param_values=np.array([float(v) for v in params.values()]),
                            label_mat=r["label_mat"], ground_mat=r["ground_mat"], range_mat=r["range_mat"], avg_residuals=r["avg_residuals"],
                            label_count=np.int32(r["label_count"]))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
