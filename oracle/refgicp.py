"""ctypes binding of oracle/_ref/libnano_gicp_ref.so — ORACLE / TEST INFRASTRUCTURE ONLY.

That library is the reference's OWN nano_gicp engine (nano_gicp.hpp, lsq_registration.hpp, nanoflann.hpp and their
impl/ and gicp/ headers, compiled unmodified from /root/reference by `make -C oracle ref`) on top of stand-in
headers for Eigen / PCL / Boost (oracle/stub_include/), which are not installed in this image.  It exists to pin
the oracle's restatement (oracle_gicp.cpp) to the reference's control flow and formulas as written: tests compare
the two on seeded inputs (tests/test_reference_engine_cpu.py).  The prebuilt .so travels to the GPU box; the
sources it is built from do not.

Same method names and matrix conventions as pyoracle.NanoGICP (numpy row-major, C side column-major).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import pyoracle as po

_f32p, _f64p, _i32p = po._f32p, po._f64p, po._i32p
_lib = None


def available() -> bool:
    po.build()
    return po.REF_GICP_LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libnano_gicp_ref.so is not available (needs /root/reference at build time)")
    L = C.CDLL(str(po.REF_GICP_LIB_PATH))
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    sig = {
        "refgicp_cloud_create": (vp, [_f32p, ci, ci]), "refgicp_cloud_free": (None, [vp]),
        "refgicp_create": (vp, []), "refgicp_free": (None, [vp]),
        "refgicp_set_num_threads": (None, [vp, ci]), "refgicp_set_correspondence_randomness": (None, [vp, ci]),
        "refgicp_set_regularization_method": (None, [vp, ci]), "refgicp_set_max_correspondence_distance": (None, [vp, cd]),
        "refgicp_set_maximum_iterations": (None, [vp, ci]), "refgicp_set_transformation_epsilon": (None, [vp, cd]),
        "refgicp_set_rotation_epsilon": (None, [vp, cd]), "refgicp_set_initial_lambda_factor": (None, [vp, cd]),
        "refgicp_set_lm_max_iterations": (None, [vp, ci]), "refgicp_set_optimizer": (None, [vp, ci]),
        "refgicp_set_input_source": (None, [vp, vp]), "refgicp_set_input_target": (None, [vp, vp]),
        "refgicp_register_input_source": (None, [vp, vp]), "refgicp_clear_source": (None, [vp]), "refgicp_clear_target": (None, [vp]),
        "refgicp_swap_source_and_target": (None, [vp]), "refgicp_share_source_tree": (None, [vp, vp]),
        "refgicp_calculate_source_covariances": (ci, [vp]), "refgicp_calculate_target_covariances": (ci, [vp]),
        "refgicp_source_covs_size": (ci, [vp]), "refgicp_target_covs_size": (ci, [vp]),
        "refgicp_get_source_covariances": (None, [vp, _f64p]), "refgicp_get_target_covariances": (None, [vp, _f64p]),
        "refgicp_set_source_covariances": (None, [vp, _f64p, ci]), "refgicp_set_target_covariances": (None, [vp, _f64p, ci]),
        "refgicp_align": (ci, [vp, _f32p, _f32p, C.POINTER(ci), C.POINTER(ci), _f64p]),
        "refgicp_linearize": (ci, [vp, _f64p, _f64p, _f64p, C.POINTER(cd)]), "refgicp_compute_error": (ci, [vp, _f64p, C.POINTER(cd)]),
        "refgicp_get_correspondences": (ci, [vp, _i32p, _f32p]), "refgicp_get_mahalanobis": (ci, [vp, _f64p]),
        "refgicp_get_residuals": (ci, [vp, _f64p]), "refgicp_get_residual_vectors": (ci, [vp, _f32p, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


class Cloud:
    """pcl::PointCloud<pcl::PointXYZI>::Ptr (of the stand-in PCL) filled from an (n, 3|4) float32 array"""

    def __init__(self, points):
        self.points = po._as_points(points)
        self.n = self.points.shape[0]
        self._h = lib().refgicp_cloud_create(self.points, self.n, self.points.shape[1])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().refgicp_cloud_free(self._h)
            self._h = None


class NanoGICP:
    """nano_gicp::NanoGICP<pcl::PointXYZI, pcl::PointXYZI> — the reference's class itself"""

    def __init__(self):
        self._g = lib().refgicp_create()
        self._src = self._tgt = None
        self._last = None

    def __del__(self):
        if getattr(self, "_g", None):
            lib().refgicp_free(self._g)
            self._g = None

    def setNumThreads(self, n): lib().refgicp_set_num_threads(self._g, n)
    def setCorrespondenceRandomness(self, k): lib().refgicp_set_correspondence_randomness(self._g, k)
    def setRegularizationMethod(self, m): lib().refgicp_set_regularization_method(self._g, m)
    def setMaxCorrespondenceDistance(self, d): lib().refgicp_set_max_correspondence_distance(self._g, d)
    def setMaximumIterations(self, n): lib().refgicp_set_maximum_iterations(self._g, n)
    def setTransformationEpsilon(self, e): lib().refgicp_set_transformation_epsilon(self._g, e)
    def setRotationEpsilon(self, e): lib().refgicp_set_rotation_epsilon(self._g, e)
    def setInitialLambdaFactor(self, f): lib().refgicp_set_initial_lambda_factor(self._g, f)
    def setLMMaxIterations(self, n): lib().refgicp_set_lm_max_iterations(self._g, n)
    def setOptimizer(self, t): lib().refgicp_set_optimizer(self._g, t)

    def setInputSource(self, cloud: Cloud):
        self._src = cloud
        lib().refgicp_set_input_source(self._g, cloud._h)

    def setInputTarget(self, cloud: Cloud):
        self._tgt = cloud
        lib().refgicp_set_input_target(self._g, cloud._h)

    def registerInputSource(self, cloud: Cloud):
        self._src = cloud
        lib().refgicp_register_input_source(self._g, cloud._h)

    def shareSourceTreeOf(self, other: "NanoGICP"):
        """`this.source_kdtree_ = other.source_kdtree_; this.source_covs_.clear();` (odom.cc:530-531)"""
        lib().refgicp_share_source_tree(self._g, other._g)

    def clearSource(self): lib().refgicp_clear_source(self._g)
    def clearTarget(self): lib().refgicp_clear_target(self._g)
    def swapSourceAndTarget(self):
        lib().refgicp_swap_source_and_target(self._g)
        self._src, self._tgt = self._tgt, self._src

    def calculateSourceCovariances(self) -> bool: return lib().refgicp_calculate_source_covariances(self._g) == 0
    def calculateTargetCovariances(self) -> bool: return lib().refgicp_calculate_target_covariances(self._g) == 0

    def _covs(self, size_fn, get_fn):
        n = size_fn(self._g)
        out = np.empty((n, 4, 4), dtype=np.float64)
        if n:
            get_fn(self._g, out.reshape(-1))
        return out

    def getSourceCovariances(self): return self._covs(lib().refgicp_source_covs_size, lib().refgicp_get_source_covariances)
    def getTargetCovariances(self): return self._covs(lib().refgicp_target_covs_size, lib().refgicp_get_target_covariances)

    def setSourceCovariances(self, covs):
        c = np.ascontiguousarray(covs, dtype=np.float64)
        lib().refgicp_set_source_covariances(self._g, c.reshape(-1), c.shape[0])

    def setTargetCovariances(self, covs):
        c = np.ascontiguousarray(covs, dtype=np.float64)
        lib().refgicp_set_target_covariances(self._g, c.reshape(-1), c.shape[0])

    def align(self, guess=None):
        g = po._cm(np.eye(4) if guess is None else guess, np.float32).reshape(-1)
        T = np.empty(16, dtype=np.float32)
        H = np.empty(36, dtype=np.float64)
        conv, it = C.c_int(), C.c_int()
        lib().refgicp_align(self._g, g, T, C.byref(conv), C.byref(it), H)
        self._last = po.AlignResult(T.reshape(4, 4).T.copy(), bool(conv.value), it.value, H.reshape(6, 6).T.copy(), -1, -1, False)
        return self._last

    def linearize(self, T):
        H, b, e = np.empty(36), np.empty(6), C.c_double()
        lib().refgicp_linearize(self._g, po._cm(T, np.float64).reshape(-1), H, b, C.byref(e))
        return e.value, H.reshape(6, 6).T.copy(), b

    def compute_error(self, T) -> float:
        e = C.c_double()
        lib().refgicp_compute_error(self._g, po._cm(T, np.float64).reshape(-1), C.byref(e))
        return e.value

    def correspondences(self):
        n = self._src.n
        corr, sqd = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.float32)
        lib().refgicp_get_correspondences(self._g, corr, sqd)
        return corr, sqd

    def mahalanobis(self) -> np.ndarray:
        out = np.empty((self._src.n, 4, 4), dtype=np.float64)
        lib().refgicp_get_mahalanobis(self._g, out.reshape(-1))
        return out

    def getResiduals(self, T=None) -> np.ndarray:
        out = np.empty(self._src.n, dtype=np.float64)
        lib().refgicp_get_residuals(self._g, out)
        return out

    def getResidualVectors(self, T) -> np.ndarray:
        out = np.empty((self._src.n, 3), dtype=np.float32)
        lib().refgicp_get_residual_vectors(self._g, po._cm(T, np.float32).reshape(-1), out.reshape(-1))
        return out
