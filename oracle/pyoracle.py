"""ctypes binding of oracle/_build/liboracle_gicp.so — ORACLE / TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (dynamic_direct_lidar_odometry_b200/) never does.

Matrix convention at this boundary: numpy arrays are ordinary row-major `M[i, j]`; the C side is
column-major like Eigen, so 4x4 / 6x6 matrices are transposed on the way in and out.  Covariances
travel as (n, 4, 4) float64 (`Eigen::Matrix4d` per point; symmetric, so no transpose is needed).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle_gicp.so"
REF_LIB_PATH = HERE / "_ref" / "libnanoflann_ref.so"
REF_GICP_LIB_PATH = HERE / "_ref" / "libnano_gicp_ref.so"  # the reference's own engine, see oracle/refgicp.py
REF_DETECTION_LIB_PATH = HERE / "_ref" / "libdetection_ref.so"  # the reference's own segmentation code, see oracle/refdet.py
REFERENCE_ROOT = Path("/root/reference")

BACKEND_CANONICAL = 0
BACKEND_NANOFLANN_REF = 1

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT = 0, 1

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class SegParams(C.Structure):
    """struct oracle_seg_params (oracle_segmentation.cpp); defaults are the reference's (detection.cpp:76-105, :520-522)."""

    _fields_ = [(n, C.c_int) for n in ("rows", "cols", "ground_rows", "valid_point_num", "min_line_num", "valid_line_num",
                                       "window_row_min", "window_row_max", "window_col_min", "window_col_max", "scan_in_sensor_frame",
                                       "unordered_residual_sums")] + \
               [(n, C.c_float) for n in ("ang_bottom", "ground_angle_threshold", "minimum_range", "sensor_mount_angle", "theta",
                                         "min_delta_z", "max_delta_z", "max_distance", "max_elevation")]

    DEFAULTS = dict(rows=128, cols=1024, ground_rows=30, valid_point_num=15, min_line_num=5, valid_line_num=5,
                    window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356, scan_in_sensor_frame=0, unordered_residual_sums=0,
                    ang_bottom=45.0, ground_angle_threshold=10.0, minimum_range=10.0, sensor_mount_angle=10.0,
                    theta=60.0 / 180.0 * np.pi, min_delta_z=0.1, max_delta_z=3.0, max_distance=20.0, max_elevation=2.0)

    def __init__(self, **kw):
        super().__init__()
        for k, v in {**self.DEFAULTS, **kw}.items():
            setattr(self, k, v)


def build(force: bool = False) -> None:
    """Compile the restatement, and the reference nanoflann when /root/reference is present."""
    newest = max((HERE / f).stat().st_mtime for f in ("oracle_gicp.cpp", "oracle_segmentation.cpp"))
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(HERE), "-B", "all"], check=True, capture_output=True)
    ref_hdr = REFERENCE_ROOT / "dynamic_direct_lidar_odometry/include/nano_gicp/impl/nanoflann_impl.hpp"
    shim = HERE / "ref_nanoflann_shim.cpp"
    gicp_shim = HERE / "ref_nano_gicp_shim.cpp"
    det_newest = max((HERE / f).stat().st_mtime for f in ("ref_detection_shim.cpp", "extract_detection.py", "stub_include/tracking/tracking.h"))
    stale = (not REF_LIB_PATH.exists() or REF_LIB_PATH.stat().st_mtime < shim.stat().st_mtime
             or not REF_GICP_LIB_PATH.exists() or REF_GICP_LIB_PATH.stat().st_mtime < gicp_shim.stat().st_mtime
             or not REF_DETECTION_LIB_PATH.exists() or REF_DETECTION_LIB_PATH.stat().st_mtime < det_newest)
    if ref_hdr.exists() and (force or stale):
        subprocess.run(["make", "-C", str(HERE), "-B", "ref"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    sig = {
        "oracle_load_reference_nanoflann": (ci, [C.c_char_p]),
        "oracle_max_threads": (ci, []),
        "oracle_cloud_create": (vp, [_f32p, ci, ci]),
        "oracle_cloud_free": (None, [vp]),
        "oracle_cloud_size": (ci, [vp]),
        "oracle_cloud_build_tree": (ci, [vp, ci]),
        "oracle_cloud_knn": (ci, [vp, _f32p, ci, ci, ci, _i32p, _f32p, ci]),
        "oracle_knn_bruteforce": (ci, [_f32p, ci, ci, _f32p, ci, ci, ci, _i32p, _f32p]),
        "oracle_cloud_covariances": (ci, [vp, ci, ci, _f64p, ci]),
        "oracle_math_sym_eig3": (None, [_f64p, _f64p, _f64p]),
        "oracle_math_inverse3": (None, [_f64p, _f64p]),
        "oracle_math_ldlt6_solve": (None, [_f64p, _f64p, _f64p]),
        "oracle_math_so3_exp": (None, [_f64p, _f64p]),
        "oracle_gicp_create": (vp, []),
        "oracle_gicp_free": (None, [vp]),
        "oracle_gicp_set_num_threads": (None, [vp, ci]),
        "oracle_gicp_set_knn_backend": (None, [vp, ci]),
        "oracle_gicp_set_correspondence_randomness": (None, [vp, ci]),
        "oracle_gicp_set_regularization_method": (None, [vp, ci]),
        "oracle_gicp_set_max_correspondence_distance": (None, [vp, cd]),
        "oracle_gicp_set_maximum_iterations": (None, [vp, ci]),
        "oracle_gicp_set_transformation_epsilon": (None, [vp, cd]),
        "oracle_gicp_set_rotation_epsilon": (None, [vp, cd]),
        "oracle_gicp_set_initial_lambda_factor": (None, [vp, cd]),
        "oracle_gicp_set_lm_max_iterations": (None, [vp, ci]),
        "oracle_gicp_set_optimizer": (None, [vp, ci]),
        "oracle_gicp_set_input_source": (None, [vp, vp]),
        "oracle_gicp_set_input_target": (None, [vp, vp]),
        "oracle_gicp_register_input_source": (None, [vp, vp]),
        "oracle_gicp_clear_source": (None, [vp]),
        "oracle_gicp_clear_target": (None, [vp]),
        "oracle_gicp_clear_source_covs": (None, [vp]),
        "oracle_gicp_clear_target_covs": (None, [vp]),
        "oracle_gicp_set_source_covariances": (None, [vp, _f64p, ci]),
        "oracle_gicp_set_target_covariances": (None, [vp, _f64p, ci]),
        "oracle_gicp_source_covs_size": (ci, [vp]),
        "oracle_gicp_target_covs_size": (ci, [vp]),
        "oracle_gicp_get_source_covariances": (None, [vp, _f64p]),
        "oracle_gicp_get_target_covariances": (None, [vp, _f64p]),
        "oracle_gicp_calculate_source_covariances": (ci, [vp]),
        "oracle_gicp_calculate_target_covariances": (ci, [vp]),
        "oracle_gicp_swap_source_and_target": (None, [vp]),
        "oracle_gicp_align": (ci, [vp, _f32p, _f32p, C.POINTER(ci), C.POINTER(ci), _f64p, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]),
        "oracle_gicp_linearize": (ci, [vp, _f64p, _f64p, _f64p, C.POINTER(cd)]),
        "oracle_gicp_compute_error": (ci, [vp, _f64p, C.POINTER(cd)]),
        "oracle_gicp_get_correspondences": (ci, [vp, _i32p, _f32p]),
        "oracle_gicp_get_mahalanobis": (ci, [vp, _f64p]),
        "oracle_gicp_get_residuals": (ci, [vp, _f64p]),
        "oracle_gicp_get_residual_vectors": (ci, [vp, _f32p, _f32p]),
        "oracle_voxel_filter": (ci, [_f32p, ci, ci, C.c_float, C.c_float, C.c_float, _f32p]),
        "oracle_crop_box": (ci, [_f32p, ci, ci, _f32p, _f32p, ci, ci, _f32p]),
        "oracle_residual_image": (None, [_f32p, ci, ci, _f64p, ci, ci, cd, cd, _f32p]),
        "oracle_segment_scan": (ci, [C.POINTER(SegParams), _f32p, ci, _f32p, vp, _i32p, _f32p, vp, _f64p, C.POINTER(ci)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def have_reference_nanoflann() -> bool:
    return REF_LIB_PATH.exists()


_ref_loaded = False


def load_reference_nanoflann() -> bool:
    global _ref_loaded
    if _ref_loaded:
        return True
    if not REF_LIB_PATH.exists():
        return False
    _ref_loaded = lib().oracle_load_reference_nanoflann(str(REF_LIB_PATH).encode()) == 0
    return _ref_loaded


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _as_points(points) -> np.ndarray:
    p = np.ascontiguousarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be (n,3) or (n,4) float32")
    return p


class Cloud:
    """A point cloud plus (optionally) a kd-tree over it — `pcl::PointCloud` + `KdTreeFLANN`."""

    def __init__(self, points):
        self.points = _as_points(points)
        self.n = self.points.shape[0]
        self._h = lib().oracle_cloud_create(self.points, self.n, self.points.shape[1])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_cloud_free(self._h)
            self._h = None

    def build_tree(self, backend: int = BACKEND_CANONICAL) -> "Cloud":
        if backend == BACKEND_NANOFLANN_REF and not load_reference_nanoflann():
            raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not available")
        if lib().oracle_cloud_build_tree(self._h, backend) != 0:
            raise RuntimeError("tree build failed")
        return self

    def knn(self, queries, k: int, threads: int = 0):
        q = _as_points(queries)
        idx = np.empty((q.shape[0], k), dtype=np.int32)
        d2 = np.empty((q.shape[0], k), dtype=np.float32)
        if lib().oracle_cloud_knn(self._h, q, q.shape[0], q.shape[1], k, idx, d2, threads) != 0:
            raise RuntimeError("kNN before build_tree")
        return idx, d2

    def covariances(self, k: int = 20, method: int = REG_PLANE, threads: int = 0) -> np.ndarray:
        out = np.empty((self.n, 4, 4), dtype=np.float64)
        if lib().oracle_cloud_covariances(self._h, k, method, out.reshape(-1), threads) != 0:
            raise RuntimeError("covariances failed (no tree, or fewer than k points)")
        return out


def voxel_filter(points, leaf) -> np.ndarray:
    """pcl::VoxelGrid restated (oracle_gicp.cpp oracle_voxel_filter): (m, 4) float32 centroids, ascending voxel index."""
    p = _as_points(points)
    lx, ly, lz = (leaf, leaf, leaf) if np.isscalar(leaf) else leaf
    out = np.empty((max(p.shape[0], 1), 4), dtype=np.float32)
    m = lib().oracle_voxel_filter(p, p.shape[0], p.shape[1], float(lx), float(ly), float(lz), out)
    if m < 0:
        raise OverflowError("leaf size too small for the input dataset" if m == -1 else "bad arguments")
    return out[:m].copy()


def crop_box(points, box_min, box_max, negative: bool = False, keep_organized: bool = False) -> np.ndarray:
    """pcl::CropBox restated (oracle_gicp.cpp oracle_crop_box)."""
    p = _as_points(points)
    lo = np.ascontiguousarray(box_min, dtype=np.float32)
    hi = np.ascontiguousarray(box_max, dtype=np.float32)
    out = np.empty((max(p.shape[0], 1), 4), dtype=np.float32)
    m = lib().oracle_crop_box(p, p.shape[0], p.shape[1], lo, hi, int(negative), int(keep_organized), out)
    return out[:m].copy()


def extract_stride(points, width: int, height: int, row_stride: int, col_stride: int) -> np.ndarray:
    """OdomNode's strided organised down-sample restated (TEST INFRASTRUCTURE).  The mask is built exactly as
    odom.cc:124-130 builds downsample_filter_indices_ (row * cloud_width_ + col for row = 0, row_stride, ... and
    col = 0, col_stride, ...); pcl::ExtractIndices with setNegative(false) + setKeepOrganized(true) (odom.cc:445-455;
    PCL 1.10 filters/impl/extract_indices.hpp: output = input, then every index NOT in the mask gets
    user_filter_value_ = NaN in each field).  PCL is not under /root/reference: parity unpinned, like the other filters."""
    p = _as_points(points)
    keep = np.zeros(p.shape[0], dtype=bool)
    for row in range(0, height, row_stride):
        for col in range(0, width, col_stride):
            keep[row * width + col] = True
    out = np.ones((p.shape[0], 4), dtype=np.float32)
    out[:, :3] = p[:, :3]
    out[~keep, :3] = np.nan
    return out


def residual_image(points, residuals, width: int = 512, height: int = 512, angle_min: float = -np.pi / 3, angle_max: float = np.pi / 3) -> np.ndarray:
    """odom.cc:804-827 restated (oracle_gicp.cpp oracle_residual_image): (height, width, 4) float32."""
    p = _as_points(points)
    r = np.ascontiguousarray(residuals, dtype=np.float64)
    out = np.empty((height, width, 4), dtype=np.float32)
    lib().oracle_residual_image(p, p.shape[0], p.shape[1], r, width, height, float(angle_min), float(angle_max), out.reshape(-1))
    return out


def segment_scan(params: "SegParams", scan_t, T, residuals=None):
    """DetectionModule projectScan + projectResiduals + groundRemoval + cloudSegmentation restated
    (oracle_segmentation.cpp).  scan_t: (rows, cols, >=3) float32 organised world-frame scan, NaN = no return;
    T: 4x4 pose; residuals: (rows, cols) float32 or None.  Returns a dict of label_mat, range_mat, ground_mat,
    label_count, avg_residuals (indexed by label) and borderline (threshold tests a different libm might flip)."""
    H, W = params.rows, params.cols
    s = np.ascontiguousarray(scan_t, dtype=np.float32).reshape(H * W, -1)
    T16 = _cm(T, np.float32)
    r = None if residuals is None else np.ascontiguousarray(residuals, dtype=np.float32).reshape(-1)
    label = np.empty(H * W, dtype=np.int32)
    rng = np.empty(H * W, dtype=np.float32)
    ground = np.empty(H * W, dtype=np.int8)
    avg = np.zeros(H * W, dtype=np.float64)
    border = C.c_int(0)
    n = lib().oracle_segment_scan(C.byref(params), s.reshape(-1), s.shape[1], T16, None if r is None else r.ctypes.data, label, rng,
                                  ground.ctypes.data, avg, C.byref(border))
    return dict(label_mat=label.reshape(H, W), range_mat=rng.reshape(H, W), ground_mat=ground.reshape(H, W), label_count=n,
                avg_residuals=avg[:n].copy(), borderline=border.value)


def knn_bruteforce(points, queries, k: int):
    p, q = _as_points(points), _as_points(queries)
    idx = np.empty((q.shape[0], k), dtype=np.int32)
    d2 = np.empty((q.shape[0], k), dtype=np.float32)
    lib().oracle_knn_bruteforce(p, p.shape[0], p.shape[1], q, q.shape[0], q.shape[1], k, idx, d2)
    return idx, d2


def _cm(M: np.ndarray, dtype) -> np.ndarray:
    """row-major numpy matrix -> flat column-major buffer"""
    return np.ascontiguousarray(np.asarray(M, dtype=dtype).T).reshape(-1)


class AlignResult:
    def __init__(self, T, converged, iterations, hessian, n_linearize, n_compute_error, lm_failed):
        self.T = T
        self.converged = converged
        self.iterations = iterations
        self.hessian = hessian
        self.n_linearize = n_linearize
        self.n_compute_error = n_compute_error
        self.lm_failed = lm_failed


class NanoGICP:
    """The reference's `nano_gicp::NanoGICP` method surface over the CPU restatement."""

    def __init__(self, backend: int = BACKEND_CANONICAL, threads: int = 0):
        self._g = lib().oracle_gicp_create()
        self._clouds = {}
        if backend == BACKEND_NANOFLANN_REF and not load_reference_nanoflann():
            raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not available")
        lib().oracle_gicp_set_knn_backend(self._g, backend)
        lib().oracle_gicp_set_num_threads(self._g, threads)

    def __del__(self):
        if getattr(self, "_g", None):
            lib().oracle_gicp_free(self._g)
            self._g = None

    # knobs (nano_gicp.hpp:83-85, lsq_registration.hpp:89-93, pcl::Registration setters)
    def setNumThreads(self, n): lib().oracle_gicp_set_num_threads(self._g, n)
    def setCorrespondenceRandomness(self, k): lib().oracle_gicp_set_correspondence_randomness(self._g, k)
    def setRegularizationMethod(self, m): lib().oracle_gicp_set_regularization_method(self._g, m)
    def setMaxCorrespondenceDistance(self, d): lib().oracle_gicp_set_max_correspondence_distance(self._g, d)
    def setMaximumIterations(self, n): lib().oracle_gicp_set_maximum_iterations(self._g, n)
    def setTransformationEpsilon(self, e): lib().oracle_gicp_set_transformation_epsilon(self._g, e)
    def setRotationEpsilon(self, e): lib().oracle_gicp_set_rotation_epsilon(self._g, e)
    def setInitialLambdaFactor(self, f): lib().oracle_gicp_set_initial_lambda_factor(self._g, f)
    def setLMMaxIterations(self, n): lib().oracle_gicp_set_lm_max_iterations(self._g, n)
    def setOptimizer(self, t): lib().oracle_gicp_set_optimizer(self._g, t)

    # state plumbing (nano_gicp_impl.hpp:98-181)
    def setInputSource(self, cloud: Cloud):
        self._src = cloud
        lib().oracle_gicp_set_input_source(self._g, cloud._h)

    def setInputTarget(self, cloud: Cloud):
        self._tgt = cloud
        lib().oracle_gicp_set_input_target(self._g, cloud._h)

    def registerInputSource(self, cloud: Cloud):
        self._src = cloud
        lib().oracle_gicp_register_input_source(self._g, cloud._h)

    def clearSource(self): lib().oracle_gicp_clear_source(self._g)
    def clearTarget(self): lib().oracle_gicp_clear_target(self._g)
    def clearSourceCovariances(self): lib().oracle_gicp_clear_source_covs(self._g)
    def clearTargetCovariances(self): lib().oracle_gicp_clear_target_covs(self._g)

    def setSourceCovariances(self, covs):
        c = np.ascontiguousarray(covs, dtype=np.float64)
        lib().oracle_gicp_set_source_covariances(self._g, c.reshape(-1), c.shape[0])

    def setTargetCovariances(self, covs):
        c = np.ascontiguousarray(covs, dtype=np.float64)
        lib().oracle_gicp_set_target_covariances(self._g, c.reshape(-1), c.shape[0])

    def getSourceCovariances(self) -> np.ndarray:
        n = lib().oracle_gicp_source_covs_size(self._g)
        out = np.empty((n, 4, 4), dtype=np.float64)
        if n:
            lib().oracle_gicp_get_source_covariances(self._g, out.reshape(-1))
        return out

    def getTargetCovariances(self) -> np.ndarray:
        n = lib().oracle_gicp_target_covs_size(self._g)
        out = np.empty((n, 4, 4), dtype=np.float64)
        if n:
            lib().oracle_gicp_get_target_covariances(self._g, out.reshape(-1))
        return out

    def calculateSourceCovariances(self) -> bool:
        return lib().oracle_gicp_calculate_source_covariances(self._g) == 0

    def calculateTargetCovariances(self) -> bool:
        return lib().oracle_gicp_calculate_target_covariances(self._g) == 0

    def swapSourceAndTarget(self):
        lib().oracle_gicp_swap_source_and_target(self._g)

    # registration (lsq_registration_impl.hpp:96-126)
    def align(self, guess=None) -> AlignResult:
        g = np.eye(4, dtype=np.float32) if guess is None else np.asarray(guess, dtype=np.float32)
        out = np.empty(16, dtype=np.float32)
        H = np.empty(36, dtype=np.float64)
        conv, it, nl, ne, fail = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = lib().oracle_gicp_align(self._g, _cm(g, np.float32), out, C.byref(conv), C.byref(it), H, C.byref(nl), C.byref(ne), C.byref(fail))
        if rc != 0:
            raise RuntimeError("oracle align failed (missing clouds or too few points)")
        self._last = AlignResult(out.reshape(4, 4).T.copy(), bool(conv.value), it.value, H.reshape(6, 6).T.copy(), nl.value, ne.value, bool(fail.value))
        return self._last

    def getFinalTransformation(self): return self._last.T
    def hasConverged(self): return self._last.converged
    def getFinalHessian(self): return self._last.hessian

    # cost-function hooks (protected in the reference; exposed for parity checks)
    def linearize(self, T):
        H = np.empty(36, dtype=np.float64)
        b = np.empty(6, dtype=np.float64)
        err = C.c_double()
        rc = lib().oracle_gicp_linearize(self._g, _cm(T, np.float64), H, b, C.byref(err))
        if rc != 0:
            raise RuntimeError(f"oracle linearize failed ({rc})")
        return err.value, H.reshape(6, 6).T.copy(), b

    def compute_error(self, T) -> float:
        err = C.c_double()
        if lib().oracle_gicp_compute_error(self._g, _cm(T, np.float64), C.byref(err)) != 0:
            raise RuntimeError("compute_error before linearize")
        return err.value

    def correspondences(self):
        n = self._src.n
        corr = np.empty(n, dtype=np.int32)
        sqd = np.empty(n, dtype=np.float32)
        m = lib().oracle_gicp_get_correspondences(self._g, corr, sqd)
        return corr[:m], sqd[:m]

    def mahalanobis(self) -> np.ndarray:
        out = np.empty((self._src.n, 4, 4), dtype=np.float64)
        lib().oracle_gicp_get_mahalanobis(self._g, out.reshape(-1))
        return out

    def getResiduals(self, T=None) -> np.ndarray:
        out = np.empty(self._src.n, dtype=np.float64)
        m = lib().oracle_gicp_get_residuals(self._g, out)
        return out[:m]

    def getResidualVectors(self, T) -> np.ndarray:
        out = np.empty((self._src.n, 3), dtype=np.float32)
        if lib().oracle_gicp_get_residual_vectors(self._g, _cm(T, np.float32), out.reshape(-1)) < 0:
            raise RuntimeError("getResiduals before align")
        return out


# ---- restated linear algebra, exposed for the numpy cross-checks --------------------------------
def math_sym_eig3(A):
    w = np.empty(3)
    V = np.empty(9)
    lib().oracle_math_sym_eig3(np.ascontiguousarray(A, dtype=np.float64).reshape(-1), w, V)
    return w, V.reshape(3, 3)


def math_inverse3(A):
    out = np.empty(9)
    lib().oracle_math_inverse3(np.ascontiguousarray(A, dtype=np.float64).reshape(-1), out)
    return out.reshape(3, 3)


def math_ldlt6_solve(A, rhs):
    x = np.empty(6)
    lib().oracle_math_ldlt6_solve(np.ascontiguousarray(A, dtype=np.float64).reshape(-1), np.ascontiguousarray(rhs, dtype=np.float64), x)
    return x


def math_so3_exp(omega):
    R = np.empty(9)
    lib().oracle_math_so3_exp(np.ascontiguousarray(omega, dtype=np.float64), R)
    return R.reshape(3, 3)
