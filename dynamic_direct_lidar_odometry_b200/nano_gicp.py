"""Host-side mirror of the reference's `nano_gicp::NanoGICP` over the C ABI (include/ddlo_gicp.h).

Same method names, argument meaning and state rules as
/root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/nano_gicp.hpp:80-136 and
lsq_registration.hpp:86-101, so tests and benchmarks read like calls into the reference.  The C++
equivalent for OdomNode is include/nano_gicp/nano_gicp.hpp in this package; this Python class exists
because the test and benchmark harness of this repo is Python.  Everything numeric happens in
libddlo_gicp_b200.so on the GPU.

numpy convention: 4x4 / 6x6 matrices are ordinary `M[row, col]` arrays; they are transposed into the
ABI's column-major (Eigen) layout here.  Covariances are (n, 4, 4) float64 (`Eigen::Matrix4d`).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import numpy as np

from . import binding as B
from .binding import (FLAG_CONVERGED, FLAG_COVS_COMPUTED, FLAG_LM_FAILED, OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT, REG_FROBENIUS,
                      REG_MIN_EIG, REG_NONE, REG_NORMALIZED_MIN_EIG, REG_PLANE, DdloError)

__all__ = ["Runtime", "PointCloud", "Covariances", "NanoGICP", "AlignInfo", "Batch", "DdloError", "REG_NONE", "REG_MIN_EIG",
           "REG_NORMALIZED_MIN_EIG", "REG_PLANE", "REG_FROBENIUS", "OPT_GAUSS_NEWTON", "OPT_LEVENBERG_MARQUARDT"]


def device_count() -> int:
    n = C.c_int()
    B.check(B.load().ddlo_device_count(C.byref(n)))
    return n.value


class Runtime:
    """One device + one CUDA stream; all handles made from it are ordered on that stream."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        B.check(B.load().ddlo_runtime_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            B.load().ddlo_runtime_destroy(self._h)
            self._h = None

    def set_align_blocks(self, max_blocks: int):
        """limit the align kernel of this runtime to max_blocks SMs (0 = all): lets several runtimes align concurrently"""
        B.check(B.load().ddlo_runtime_set_align_blocks(self._h, int(max_blocks)))

    def synchronize(self):
        B.check(B.load().ddlo_runtime_synchronize(self._h))

    def timer_begin(self):
        B.check(B.load().ddlo_runtime_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        B.check(B.load().ddlo_runtime_timer_end(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_longlong()
        B.check(B.load().ddlo_runtime_launch_count(self._h, C.byref(n)))
        return n.value

    def flush_l2(self, nbytes: int = 256 << 20):
        B.check(B.load().ddlo_runtime_flush_l2(self._h, nbytes))

    def event_record(self, slot: int):
        B.check(B.load().ddlo_runtime_event_record(self._h, slot))

    def event_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float()
        B.check(B.load().ddlo_runtime_event_elapsed(self._h, a, b, C.byref(ms)))
        return ms.value


def pinned_array(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory (cudaMallocHost through the ABI); freed with the array."""
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    p = C.c_void_p()
    B.check(B.load().ddlo_host_alloc(nbytes, C.byref(p)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    # the page-locked block lives exactly as long as some array still views it: ddlo_host_free runs when the ctypes
    # buffer (the base object of every view) is collected
    weakref.finalize(buf, _free_pinned, p.value)
    return arr


def _free_pinned(address: int) -> None:
    try:
        B.load().ddlo_host_free(C.c_void_p(address))
    except Exception:  # interpreter shutdown
        pass


class PointCloud:
    """`pcl::PointCloud<PointXYZI>::Ptr` living on the device, plus its kNN index once built
    (what `nanoflann::KdTreeFLANN::setInputCloud` creates in the reference)."""

    def __init__(self, rt: Runtime, points=None, _handle=None):
        self.rt = rt
        if _handle is not None:
            self._h = _handle
            return
        p = np.ascontiguousarray(points, dtype=np.float32)
        if p.ndim != 2 or p.shape[1] < 3:
            raise ValueError("points must be (n, >=3) float32")
        self._h = C.c_void_p()
        B.check(B.load().ddlo_cloud_create(rt._h, B.ptr(p), p.shape[0], p.strides[0] if p.shape[0] else 4 * p.shape[1], C.byref(self._h)))

    def __del__(self):
        # safe in any order: the C runtime is reference counted and outlives Runtime.close() while handles exist
        if getattr(self, "_h", None):
            try:
                B.load().ddlo_cloud_release(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    def size(self) -> int:
        n = C.c_int()
        B.check(B.load().ddlo_cloud_size(self._h, C.byref(n)))
        return n.value

    __len__ = size

    def download(self) -> np.ndarray:
        out = np.empty((self.size(), 4), dtype=np.float32)
        B.check(B.load().ddlo_cloud_download(self._h, B.ptr(out)))
        return out

    def share(self) -> "PointCloud":
        """ddlo_cloud_share: index built, owner stream synchronised, handle usable from other runtimes of the device"""
        B.check(B.load().ddlo_cloud_share(self._h))
        return self

    def build_index(self) -> "PointCloud":
        B.check(B.load().ddlo_cloud_build_index(self._h))
        return self

    def has_index(self) -> bool:
        v = C.c_int()
        B.check(B.load().ddlo_cloud_has_index(self._h, C.byref(v)))
        return bool(v.value)

    def nearestKSearch(self, queries, k: int):
        """KdTreeFLANN::nearestKSearch for a batch: (idx (nq,k) int32, sqdist (nq,k) float32)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        idx = np.empty((q.shape[0], k), dtype=np.int32)
        d2 = np.empty((q.shape[0], k), dtype=np.float32)
        B.check(B.load().ddlo_cloud_knn(self._h, B.ptr(q), q.shape[0], q.strides[0] if q.shape[0] else 12, k, B.ptr(idx), B.ptr(d2), None))
        return idx, d2

    def transformed(self, T) -> "PointCloud":
        t = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T)
        h = C.c_void_p()
        B.check(B.load().ddlo_cloud_transform(self._h, B.ptr(t), C.byref(h)))
        return PointCloud(self.rt, _handle=h)

    def voxel_filtered(self, leaf) -> "PointCloud":
        """pcl::VoxelGrid with setLeafSize(leaf, leaf, leaf) (or a 3-tuple), on the device (odom.cc:469-474)."""
        lx, ly, lz = (leaf, leaf, leaf) if np.isscalar(leaf) else leaf
        h = C.c_void_p()
        B.check(B.load().ddlo_cloud_voxel_filter(self._h, float(lx), float(ly), float(lz), C.byref(h)))
        return PointCloud(self.rt, _handle=h)

    def cropped(self, box_min, box_max, negative: bool = False, keep_organized: bool = False) -> "PointCloud":
        """pcl::CropBox with setMin/setMax (and setNegative / setKeepOrganized), on the device (odom.cc:459-465)."""
        lo = np.ascontiguousarray(box_min, dtype=np.float32)
        hi = np.ascontiguousarray(box_max, dtype=np.float32)
        h = C.c_void_p()
        B.check(B.load().ddlo_cloud_crop_box(self._h, B.ptr(lo), B.ptr(hi), int(negative), int(keep_organized), C.byref(h)))
        return PointCloud(self.rt, _handle=h)

    def stride_filtered(self, width: int, height: int, row_stride: int, col_stride: int) -> "PointCloud":
        """pcl::ExtractIndices with the strided mask of odom.cc:124-130, keep-organised (odom.cc:445-455): same size, NaN elsewhere."""
        h = C.c_void_p()
        B.check(B.load().ddlo_cloud_extract_stride(self._h, width, height, row_stride, col_stride, C.byref(h)))
        return PointCloud(self.rt, _handle=h)

    @staticmethod
    def concat(rt: Runtime, parts: Sequence["PointCloud"]) -> "PointCloud":
        arr = (C.c_void_p * len(parts))(*[p._h for p in parts])
        h = C.c_void_p()
        B.check(B.load().ddlo_cloud_concat(rt._h, arr, len(parts), C.byref(h)))
        return PointCloud(rt, _handle=h)


class Covariances:
    """`std::vector<Eigen::Matrix4d>` on the device; assignment shares the buffer (no copy)."""

    def __init__(self, rt: Runtime, matrices=None, _handle=None):
        self.rt = rt
        if _handle is not None:
            self._h = _handle
            return
        m = np.ascontiguousarray(matrices, dtype=np.float64)
        if m.ndim != 3 or m.shape[1:] != (4, 4):
            raise ValueError("covariances must be (n, 4, 4) float64")
        self._h = C.c_void_p()
        B.check(B.load().ddlo_covs_from_host(rt._h, B.ptr(m), m.shape[0], C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                B.load().ddlo_covs_release(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    @staticmethod
    def compute(cloud: PointCloud, k: int = 20, method: int = REG_PLANE) -> "Covariances":
        h = C.c_void_p()
        B.check(B.load().ddlo_covs_compute(cloud._h, k, method, C.byref(h)))
        return Covariances(cloud.rt, _handle=h)

    def share(self, target: Optional[PointCloud] = None) -> "Covariances":
        """ddlo_covs_share: usable from other runtimes of the device; `target` = the cloud these are the target covariances of"""
        B.check(B.load().ddlo_covs_share(self._h, target._h if target is not None else None))
        return self

    def size(self) -> int:
        n = C.c_int()
        B.check(B.load().ddlo_covs_size(self._h, C.byref(n)))
        return n.value

    __len__ = size

    def to_host(self) -> np.ndarray:
        out = np.empty((self.size(), 4, 4), dtype=np.float64)
        B.check(B.load().ddlo_covs_to_host(self._h, B.ptr(out)))
        return out

    @staticmethod
    def concat(rt: Runtime, parts: Sequence["Covariances"]) -> "Covariances":
        arr = (C.c_void_p * len(parts))(*[p._h for p in parts])
        h = C.c_void_p()
        B.check(B.load().ddlo_covs_concat(rt._h, arr, len(parts), C.byref(h)))
        return Covariances(rt, _handle=h)


class AlignInfo:
    def __init__(self, r: B.AlignResult):
        self.T = np.array(r.final_transformation, dtype=np.float32).reshape(4, 4).T.copy()
        self.hessian = np.array(r.final_hessian, dtype=np.float64).reshape(6, 6).T.copy()
        self.flags = r.flags
        self.converged = bool(r.flags & FLAG_CONVERGED)
        self.lm_failed = bool(r.flags & FLAG_LM_FAILED)
        self.covs_computed = bool(r.flags & FLAG_COVS_COMPUTED)
        self.iterations = r.nr_iterations
        self.n_linearize = r.n_linearize
        self.n_compute_error = r.n_compute_error
        self.final_error = r.final_error
        self.lm_lambda = r.lm_lambda


class NanoGICP:
    """nano_gicp::NanoGICP<PointXYZI, PointXYZI> (nano_gicp.hpp:58-148)."""

    def __init__(self, rt: Runtime):
        self.rt = rt
        self._g = C.c_void_p()
        B.check(B.load().ddlo_gicp_create(rt._h, C.byref(self._g)))
        self._p = B.Params()
        B.check(B.load().ddlo_gicp_get_params(self._g, C.byref(self._p)))
        self._src: Optional[PointCloud] = None
        self._tgt: Optional[PointCloud] = None
        self._last: Optional[AlignInfo] = None

    def __del__(self):
        if getattr(self, "_g", None):
            try:
                B.load().ddlo_gicp_destroy(self._g)
            except Exception:  # interpreter shutdown
                pass
            self._g = None

    def _push(self):
        B.check(B.load().ddlo_gicp_set_params(self._g, C.byref(self._p)))

    # -- knobs: nano_gicp.hpp:83-85, lsq_registration.hpp:89-91, pcl::Registration ---------------
    def setNumThreads(self, n: int):  # OpenMP thread count: meaningless on the device, accepted for source compatibility
        pass

    def setCorrespondenceRandomness(self, k: int):
        self._p.k_correspondences = k
        self._push()

    def setRegularizationMethod(self, method: int):
        self._p.regularization_method = method
        self._push()

    def setMaxCorrespondenceDistance(self, d: float):
        self._p.max_correspondence_distance = d
        self._push()

    def setMaximumIterations(self, n: int):
        self._p.max_iterations = n
        self._push()

    def setTransformationEpsilon(self, eps: float):
        self._p.transformation_epsilon = eps
        self._push()

    def setRotationEpsilon(self, eps: float):
        self._p.rotation_epsilon = eps
        self._push()

    def setInitialLambdaFactor(self, f: float):
        self._p.lm_init_lambda_factor = f
        self._push()

    def setLMMaxIterations(self, n: int):
        self._p.lm_max_iterations = n
        self._push()

    def setOptimizer(self, t: int):
        self._p.optimizer = t
        self._push()

    def setDebugPrint(self, flag: bool):
        pass

    # accepted and ignored, exactly like nano_gicp ignores them (odom.cc:96-98,104-112)
    def setEuclideanFitnessEpsilon(self, eps): pass
    def setRANSACIterations(self, n): pass
    def setRANSACOutlierRejectionThreshold(self, t): pass
    def setSearchMethodSource(self, tree, force_no_recompute=False): pass
    def setSearchMethodTarget(self, tree, force_no_recompute=False): pass

    # -- state plumbing: nano_gicp_impl.hpp:98-181 ---------------------------------------------------
    def setInputSource(self, cloud: PointCloud):
        B.check(B.load().ddlo_gicp_set_input_source(self._g, cloud._h, 1))
        self._src = cloud

    def registerInputSource(self, cloud: PointCloud):
        B.check(B.load().ddlo_gicp_set_input_source(self._g, cloud._h, 0))
        self._src = cloud

    def setInputTarget(self, cloud: PointCloud):
        B.check(B.load().ddlo_gicp_set_input_target(self._g, cloud._h))
        self._tgt = cloud

    def clearSource(self):
        B.check(B.load().ddlo_gicp_clear_source(self._g))
        self._src = None

    def clearTarget(self):
        B.check(B.load().ddlo_gicp_clear_target(self._g))
        self._tgt = None

    def setSourceCovariances(self, covs):
        c = covs if isinstance(covs, Covariances) else Covariances(self.rt, covs)
        B.check(B.load().ddlo_gicp_set_source_covariances(self._g, c._h))

    def setTargetCovariances(self, covs):
        c = covs if isinstance(covs, Covariances) else Covariances(self.rt, covs)
        B.check(B.load().ddlo_gicp_set_target_covariances(self._g, c._h))

    def _get_covs(self, fn) -> Optional[Covariances]:
        h = C.c_void_p()
        B.check(fn(self._g, C.byref(h)))
        return Covariances(self.rt, _handle=h) if h.value else None

    def getSourceCovariances(self) -> Optional[Covariances]:
        return self._get_covs(B.load().ddlo_gicp_get_source_covariances)

    def getTargetCovariances(self) -> Optional[Covariances]:
        return self._get_covs(B.load().ddlo_gicp_get_target_covariances)

    # the reference's public data members source_covs_ / target_covs_ (nano_gicp.hpp:135-136):
    # reading gives the shared device vector, assigning None is `.clear()`
    @property
    def source_covs_(self): return self.getSourceCovariances()

    @source_covs_.setter
    def source_covs_(self, v):
        B.check(B.load().ddlo_gicp_set_source_covariances(self._g, v._h if v is not None else None))

    @property
    def target_covs_(self): return self.getTargetCovariances()

    @target_covs_.setter
    def target_covs_(self, v):
        B.check(B.load().ddlo_gicp_set_target_covariances(self._g, v._h if v is not None else None))

    # source_kdtree_ / target_kdtree_ (nano_gicp.hpp:132-133): the index lives in the cloud handle, so
    # `s2m.source_kdtree_ = s2s.source_kdtree_` (odom.cc:530) is sharing the cloud that carries it
    @property
    def source_kdtree_(self): return self._src

    @source_kdtree_.setter
    def source_kdtree_(self, cloud: PointCloud):
        B.check(B.load().ddlo_gicp_set_input_source(self._g, cloud._h, 0))
        self._src = cloud

    @property
    def target_kdtree_(self): return self._tgt

    def calculateSourceCovariances(self) -> bool:
        B.check(B.load().ddlo_gicp_calculate_source_covariances(self._g))
        return True

    def calculateTargetCovariances(self) -> bool:
        B.check(B.load().ddlo_gicp_calculate_target_covariances(self._g))
        return True

    def swapSourceAndTarget(self):
        B.check(B.load().ddlo_gicp_swap_source_and_target(self._g))
        self._src, self._tgt = self._tgt, self._src

    # -- registration --------------------------------------------------------------------------------------
    def align(self, guess=None) -> AlignInfo:
        r = B.AlignResult()
        g = None if guess is None else np.ascontiguousarray(np.asarray(guess, dtype=np.float32).T)
        B.check(B.load().ddlo_gicp_align(self._g, None if g is None else B.ptr(g), C.byref(r)))
        self._last = AlignInfo(r)
        return self._last

    def align_async(self, guess=None) -> None:
        """enqueue align() on the runtime's stream without waiting (see ddlo_gicp_align_async)"""
        g = None if guess is None else np.ascontiguousarray(np.asarray(guess, dtype=np.float32).T)
        B.check(B.load().ddlo_gicp_align_async(self._g, None if g is None else B.ptr(g)))

    def align_finish(self) -> AlignInfo:
        r = B.AlignResult()
        B.check(B.load().ddlo_gicp_align_finish(self._g, C.byref(r)))
        self._last = AlignInfo(r)
        return self._last

    TIMELINE_TAGS = {1: "start", 2: "lin_done", 3: "lin_synced", 4: "lin_summed", 5: "solved", 6: "err_done", 7: "err_synced", 8: "decided", 9: "end", 10: "search_done"}

    def debug_enable(self, on: bool = True):
        """profiling of the align kernel (timeline, block times) is off unless enabled here"""
        B.check(B.load().ddlo_gicp_debug_enable(self._g, int(on)))

    def debug_timeline(self):
        """[(tag, microseconds since kernel start)] recorded by block 0 during the last align (profiling aid)."""
        buf = np.zeros(128, dtype=np.uint64)
        n = B.load().ddlo_gicp_debug_timeline(self._g, B.ptr(buf), 128)
        if n < 0:
            B.check(n)
        t = [(int(v >> np.uint64(56)), int(v & np.uint64(0x00FFFFFFFFFFFFFF))) for v in buf[:n]]
        t0 = t[0][1] if t else 0
        return [(self.TIMELINE_TAGS.get(tag, str(tag)), (ns - t0) / 1e3) for tag, ns in t]

    def debug_visits(self):
        """(4, ns, 4) ints {node visits, leaf scans, warp steps, 0} per source point for the first 4 linearize passes;
        None unless the library was built with -DDDLO_VISIT_STATS."""
        n = self._src.size() if self._src is not None else 0
        buf = np.zeros((4, max(n, 1), 4), dtype=np.int32)
        got = B.load().ddlo_gicp_debug_visits(self._g, B.ptr(buf), n)
        if got < 0:
            B.check(got)
        return buf if got > 0 else None

    def debug_block_times(self):
        """(passes<=8, blocks, 6) microseconds: pass start / search done / phase B done / after the grid sync / slowest and fastest lin_point call."""
        cap = 1024
        buf = np.zeros((8, cap, 8), dtype=np.uint64)
        nb = B.load().ddlo_gicp_debug_block_times(self._g, B.ptr(buf), cap)
        if nb < 0:
            B.check(nb)
        t = buf[:, :nb, :].astype(np.float64)
        out = np.zeros((8, nb, 8))
        out[:, :, :4] = (t[:, :, :4] - t[0, :, 0].min()) / 1e3
        out[:, :, 6:8] = (t[:, :, 6:8] - t[0, :, 0].min()) / 1e3
        out[:, :, 4] = t[:, :, 4] / 1e3                      # slowest lin_point call of the block
        out[:, :, 5] = (4294967295.0 - t[:, :, 5]) / 1e3     # fastest
        return out

    def getFinalTransformation(self) -> np.ndarray: return self._last.T
    def hasConverged(self) -> bool: return self._last.converged
    def getFinalHessian(self) -> np.ndarray: return self._last.hessian

    def alignedCloud(self) -> PointCloud:
        h = C.c_void_p()
        B.check(B.load().ddlo_gicp_aligned_cloud(self._g, C.byref(h)))
        return PointCloud(self.rt, _handle=h)

    # -- cost-function hooks (protected in the reference, exposed for parity checks) -------------------------
    def linearize(self, T):
        t = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T)
        H = np.empty((6, 6), dtype=np.float64)
        b = np.empty(6, dtype=np.float64)
        e = C.c_double()
        B.check(B.load().ddlo_gicp_linearize(self._g, B.ptr(t), B.ptr(H), B.ptr(b), C.byref(e)))
        return e.value, H.T.copy(), b

    def compute_error(self, T) -> float:
        t = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T)
        e = C.c_double()
        B.check(B.load().ddlo_gicp_compute_error(self._g, B.ptr(t), C.byref(e)))
        return e.value

    def correspondences(self):
        n = self._src.size()
        corr = np.empty(n, dtype=np.int32)
        sqd = np.empty(n, dtype=np.float32)
        B.check(B.load().ddlo_gicp_get_correspondences(self._g, B.ptr(corr), B.ptr(sqd), n))
        return corr, sqd

    def mahalanobis(self) -> np.ndarray:
        n = self._src.size()
        out = np.empty((n, 4, 4), dtype=np.float64)
        B.check(B.load().ddlo_gicp_get_mahalanobis(self._g, B.ptr(out), n))
        return out

    def getResiduals(self, T=None) -> np.ndarray:
        """getResiduals(std::vector<double>&, trans): sqrt(sq_distances_) of the last linearize."""
        n = self._src.size()
        out = np.empty(n, dtype=np.float64)
        B.check(B.load().ddlo_gicp_get_residuals(self._g, B.ptr(out), n))
        return out

    def getResidualsAsync(self, out: np.ndarray) -> None:
        """Enqueue getResiduals into `out` (float64, ideally from pinned_array); complete after the next
        align_finish() / Runtime.synchronize().  Lets pose and residuals share one host round trip."""
        if out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array")
        B.check(B.load().ddlo_gicp_get_residuals_async(self._g, B.ptr(out), out.size))

    def residualImage(self, width: int = 512, height: int = 512, angle_min: float = -np.pi / 3, angle_max: float = np.pi / 3) -> np.ndarray:
        """The residual cloud of odom.cc:804-827 as an (height, width, 4) float32 array (x, y, z, residual)."""
        out = np.empty((height, width, 4), dtype=np.float32)
        B.check(B.load().ddlo_gicp_residual_image(self._g, width, height, float(angle_min), float(angle_max), B.ptr(out)))
        return out

    def getResidualVectors(self, T) -> np.ndarray:
        """getResiduals(std::vector<Eigen::Vector3f>&, trans)."""
        n = self._src.size()
        t = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T)
        out = np.empty((n, 3), dtype=np.float32)
        B.check(B.load().ddlo_gicp_get_residual_vectors(self._g, B.ptr(t), B.ptr(out), n))
        return out


def align_batch(engines: Sequence[NanoGICP], guesses=None):
    """ddlo_gicp_align_batch: independent registrations queued back to back, one host sync."""
    m = len(engines)
    arr = (C.c_void_p * m)(*[e._g for e in engines])
    res = (B.AlignResult * m)()
    g = None
    if guesses is not None:
        g = np.ascontiguousarray(np.transpose(np.asarray(guesses, dtype=np.float32), (0, 2, 1)))
    B.check(B.load().ddlo_gicp_align_batch(arr, m, None if g is None else B.ptr(g), res))
    out = [AlignInfo(r) for r in res]
    for e, o in zip(engines, out):
        e._last = o
    return out


class Batch:
    """ddlo_batch_*: many independent registrations on one device, driven from C++ (BASELINE config C5).

    S lanes (stream + engine each, align kernels limited to num_SMs / S blocks) work through the submitted units
    concurrently; Python only hands over the job table.  Units are (source id, target id or -1, guess) over clouds
    staged in HBM with `stage`."""

    def __init__(self, device: int = 0, lanes: int = 4, align_blocks: int = 0, host_threads: int = 1, mode: str = "waves", wave_units: int = 0):
        self._b = C.c_void_p()
        B.check(B.load().ddlo_batch_create(device, lanes, align_blocks, host_threads, C.byref(self._b)))
        self.device = device
        self.mode = mode
        B.check(B.load().ddlo_batch_set_mode(self._b, {"lanes": B.BATCH_LANES, "waves": B.BATCH_WAVES}[mode], wave_units))
        a, c, t = C.c_int(), C.c_int(), C.c_int()
        B.check(B.load().ddlo_batch_info(self._b, C.byref(a), C.byref(c), C.byref(t)))
        self.lanes, self.align_blocks, self.host_threads = a.value, c.value, t.value
        self._res = None
        self._m = 0

    def close(self):
        if getattr(self, "_b", None):
            B.load().ddlo_batch_destroy(self._b)
            self._b = None

    __del__ = close

    def set_params(self, **kw):
        p = B.Params()
        B.check(B.load().ddlo_params_default(C.byref(p)))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown engine parameter {k}")
            setattr(p, k, v)
        B.check(B.load().ddlo_batch_set_params(self._b, C.byref(p)))

    def stage(self, points) -> int:
        p = np.ascontiguousarray(points, dtype=np.float32)
        if p.ndim != 2 or p.shape[1] < 3:
            raise ValueError("points must be (n, >=3) float32")
        i = C.c_int()
        B.check(B.load().ddlo_batch_stage_cloud(self._b, B.ptr(p), p.shape[0], p.strides[0] if p.shape[0] else 4 * p.shape[1], C.byref(i)))
        return i.value

    def set_shared_target(self, cloud_id: int, covariances=None):
        m = None if covariances is None else np.ascontiguousarray(covariances, dtype=np.float64)
        B.check(B.load().ddlo_batch_set_shared_target(self._b, cloud_id, None if m is None else B.ptr(m)))

    @staticmethod
    def jobs(units) -> "C.Array":
        """units: iterable of (source id, target id or -1, guess 4x4 or None) -> ddlo_batch_job[]"""
        units = list(units)
        arr = (B.BatchJob * len(units))()
        eye = np.eye(4, dtype=np.float32)
        for j, (s, t, g) in zip(arr, units):
            j.source, j.target = int(s), int(t)
            gm = eye if g is None else np.asarray(g, dtype=np.float32)
            j.guess[:] = np.ascontiguousarray(gm.T).ravel().tolist()
        return arr

    def submit(self, jobs) -> None:
        if not isinstance(jobs, C.Array):
            jobs = self.jobs(jobs)
        self._m = len(jobs)
        self._res = (B.AlignResult * max(self._m, 1))()
        self._jobs = jobs  # the job table must outlive the submission
        rc = B.load().ddlo_batch_submit(self._b, jobs, self._m, self._res)
        if rc != B.OK:
            text = B.load().ddlo_last_error().decode(errors="replace")
            B.load().ddlo_batch_wait(self._b)  # whatever was enqueued completes; the batch stays usable
            raise DdloError(rc, text)

    def submit_host(self, jobs, sources) -> None:
        """ddlo_batch_submit_host: unit i uploads sources[i] ((n, 4) float32 arrays, ideally from pinned_array) as its source scan"""
        if not isinstance(jobs, C.Array):
            jobs = self.jobs(jobs)
        self._m = len(jobs)
        self._res = (B.AlignResult * max(self._m, 1))()
        self._jobs = jobs
        self._host = [np.ascontiguousarray(s, dtype=np.float32) for s in sources]  # kept alive until wait()
        ptrs = (C.c_void_p * max(self._m, 1))(*[a.ctypes.data for a in self._host])
        ns = (C.c_int * max(self._m, 1))(*[a.shape[0] for a in self._host])
        self._host_tables = (ptrs, ns)
        stride = self._host[0].strides[0] if self._host else 16
        rc = B.load().ddlo_batch_submit_host(self._b, jobs, self._m, ptrs, ns, stride, self._res)
        if rc != B.OK:
            text = B.load().ddlo_last_error().decode(errors="replace")
            B.load().ddlo_batch_wait(self._b)
            raise DdloError(rc, text)

    def stats(self):
        """(LM rounds launched, completion polls) by the waves driver since creation"""
        r, p = C.c_longlong(), C.c_longlong()
        B.check(B.load().ddlo_batch_stats(self._b, C.byref(r), C.byref(p)))
        return r.value, p.value

    def wait(self, raw: bool = False):
        B.check(B.load().ddlo_batch_wait(self._b))
        if raw:
            return self._res
        return [AlignInfo(self._res[i]) for i in range(self._m)]

    def run(self, jobs, raw: bool = False):
        self.submit(jobs)
        return self.wait(raw)

    def launch_count(self) -> int:
        n = C.c_longlong()
        B.check(B.load().ddlo_batch_launch_count(self._b, C.byref(n)))
        return n.value


def rotation_to_wxyz(R) -> np.ndarray:
    """unit quaternion (w, x, y, z) of a 3x3 rotation (what OdomNode keeps as rotq_)"""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return np.asarray(q, dtype=np.float32)


class KeyframeStore:
    """ddlo_keyframes_*: OdomNode's keyframes_ / keyframe_normals_ on the device, the new-keyframe decision of
    updateKeyframes and the submap selection + assembly of getSubmapKeyframes (odom.cc:1067-1150, 1215-1315)."""

    def __init__(self, rt: Runtime):
        self.rt = rt
        self._k = C.c_void_p()
        B.check(B.load().ddlo_keyframes_create(rt._h, C.byref(self._k)))

    def __del__(self):
        if getattr(self, "_k", None):
            try:
                B.load().ddlo_keyframes_destroy(self._k)
            except Exception:
                pass
            self._k = None

    def __len__(self) -> int:
        n = C.c_int()
        B.check(B.load().ddlo_keyframes_count(self._k, C.byref(n)))
        return n.value

    def add(self, position, rotation_wxyz, cloud: PointCloud, covs: Covariances):
        p = np.ascontiguousarray(position, dtype=np.float32)
        q = np.ascontiguousarray(rotation_wxyz, dtype=np.float32)
        B.check(B.load().ddlo_keyframes_add(self._k, B.ptr(p), B.ptr(q), cloud._h, covs._h))

    def is_new(self, position, rotation_wxyz, thresh_dist: float, thresh_rot_deg: float):
        """(new keyframe?, closest index, distance to it, rotation against it in degrees)"""
        p = np.ascontiguousarray(position, dtype=np.float32)
        q = np.ascontiguousarray(rotation_wxyz, dtype=np.float32)
        new, idx, d, th = C.c_int(), C.c_int(), C.c_float(), C.c_float()
        B.check(B.load().ddlo_keyframes_is_new(self._k, B.ptr(p), B.ptr(q), thresh_dist, thresh_rot_deg, C.byref(new), C.byref(idx), C.byref(d), C.byref(th)))
        return bool(new.value), idx.value, d.value, th.value

    def get_submap(self, position, knn: int = 10, kcv: int = 10, kcc: int = 10, alpha: float = 1.0):
        """(changed, indices, cloud or None, covariances or None) - the cloud and covariances are built on the device"""
        p = np.ascontiguousarray(position, dtype=np.float32)
        changed, n = C.c_int(), C.c_int()
        cap = max(len(self), 1)
        idx = np.zeros(cap, dtype=np.int32)
        hc, hv = C.c_void_p(), C.c_void_p()
        B.check(B.load().ddlo_keyframes_get_submap(self._k, B.ptr(p), knn, kcv, kcc, float(alpha), C.byref(changed), C.byref(hc), C.byref(hv),
                                                   B.ptr(idx), cap, C.byref(n)))
        cloud = PointCloud(self.rt, _handle=hc) if hc.value else None
        covs = Covariances(self.rt, _handle=hv) if hv.value else None
        return bool(changed.value), idx[: n.value].tolist(), cloud, covs

    def hulls(self):
        """(convex hull vertex indices, concave hull vertex indices, dimension of the concave hull) of the last get_submap"""
        cap = max(len(self), 1)
        cv, cc = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        ncv, ncc, dim = C.c_int(), C.c_int(), C.c_int()
        B.check(B.load().ddlo_keyframes_hulls(self._k, B.ptr(cv), C.byref(ncv), B.ptr(cc), C.byref(ncc), cap, C.byref(dim)))
        return cv[: ncv.value].tolist(), cc[: ncc.value].tolist(), dim.value


def hull_convex(points) -> list:
    """convex-hull vertex indices of (n, 3) positions as the keyframe store computes them (csrc/hull.hpp; host only)"""
    p = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros(max(len(p), 1), dtype=np.int32)
    m = B.load().ddlo_hull_convex(B.ptr(p), len(p), B.ptr(out), len(out))
    return out[:m].tolist()


def hull_concave(points, alpha: float):
    """concave-hull (alpha shape) vertex indices, or None when the positions are 3-dimensional in PCL's sense"""
    p = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros(max(len(p), 1), dtype=np.int32)
    m = B.load().ddlo_hull_concave(B.ptr(p), len(p), float(alpha), B.ptr(out), len(out))
    return None if m == -3 else out[:m].tolist()
