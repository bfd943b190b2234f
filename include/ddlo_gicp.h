/*
 * ddlo_gicp.h — C ABI of the B200-native nano_gicp scan-registration path.
 *
 * This is the drop-in boundary for the reference's registration engine
 *   R = /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp
 *   nano_gicp::NanoGICP<PointXYZI,PointXYZI>        R/nano_gicp.hpp:58-148
 *   nano_gicp::LsqRegistration                      R/lsq_registration.hpp:60-128
 *   nanoflann::KdTreeFLANN                          R/nanoflann.hpp:53-203
 * Every entry point names the reference member it replaces.  The C++ class that keeps the
 * reference's method names on top of this ABI is
 *   dynamic_direct_lidar_odometry_b200/include/nano_gicp/nano_gicp.hpp
 * and INTEGRATION.md shows how OdomNode (R/../src/odometry/odom.cc) binds to it.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (DDLO_OK == 0,
 *     negative = error, text via ddlo_last_error()); nothing throws across this boundary.
 *   - 4x4 / 6x6 matrices are COLUMN-MAJOR, exactly the memory of Eigen::Matrix4f / Matrix4d /
 *     Matrix<double,6,6>, so `final_transformation_.data()` etc. can be passed straight through.
 *   - covariances cross the boundary as Eigen::Matrix4d per point (16 doubles, rows/cols 3 zero),
 *     the element type of `source_covs_` / `target_covs_` (R/nano_gicp.hpp:135-136).
 *   - handles are opaque, reference counted and bound to the device of the runtime that made them.
 *     A runtime owns one CUDA stream; all work of its handles is ordered on that stream.
 *   - there is no CPU fallback: without a CUDA device every call fails with DDLO_E_CUDA.
 */
#ifndef DDLO_GICP_H
#define DDLO_GICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDLO_ABI_VERSION 2

/* status codes */
enum {
  DDLO_OK = 0,
  DDLO_E_INVALID = -1,      /* null / malformed argument */
  DDLO_E_CUDA = -2,         /* CUDA runtime error, or no device */
  DDLO_E_EMPTY = -3,        /* empty cloud (nanoflann throws here, nanoflann_impl.hpp:1470) */
  DDLO_E_TOO_FEW = -4,      /* fewer points than k (undefined behaviour in the reference) */
  DDLO_E_NOT_READY = -5,    /* source/target/covariances/correspondences missing */
  DDLO_E_SIZE = -6,         /* covariance count does not match the cloud */
  DDLO_E_NONFINITE = -7,    /* NaN/Inf coordinate in an input cloud */
  DDLO_E_UNSUPPORTED = -8
};

/* RegularizationMethod, same order as R/gicp/gicp_settings.hpp:47-54 */
enum { DDLO_REG_NONE = 0, DDLO_REG_MIN_EIG = 1, DDLO_REG_NORMALIZED_MIN_EIG = 2, DDLO_REG_PLANE = 3, DDLO_REG_FROBENIUS = 4 };
/* LSQ_OPTIMIZER_TYPE, same order as R/lsq_registration.hpp:54-58 */
enum { DDLO_OPT_GAUSS_NEWTON = 0, DDLO_OPT_LEVENBERG_MARQUARDT = 1 };

/* align() status bits (ddlo_align_result.flags) */
enum {
  DDLO_FLAG_CONVERGED = 1,      /* converged_ (lsq_registration_impl.hpp:121) */
  DDLO_FLAG_LM_FAILED = 2,      /* step_lm ran out of trials: the reference prints "lm not converged!!" (:117) */
  DDLO_FLAG_COVS_COMPUTED = 4,  /* align had to compute missing covariances (nano_gicp_impl.hpp:186-193) */
  DDLO_FLAG_NONFINITE = 8       /* an input cloud holds NaN/Inf coordinates: align returns DDLO_E_NONFINITE */
};

typedef struct ddlo_runtime ddlo_runtime; /* device + stream + scratch */
typedef struct ddlo_cloud ddlo_cloud;     /* pcl::PointCloud<PointXYZI> on the device, with an optional kNN index */
typedef struct ddlo_covs ddlo_covs;       /* std::vector<Eigen::Matrix4d> on the device (6 doubles per point) */
typedef struct ddlo_gicp ddlo_gicp;       /* one NanoGICP instance */

/* Knobs of NanoGICP / LsqRegistration / pcl::Registration that the engine honours; defaults are the
 * reference's (nano_gicp_impl.hpp:58-62, lsq_registration_impl.hpp:53-61). */
typedef struct ddlo_params {
  int k_correspondences;              /* setCorrespondenceRandomness        (20) */
  int regularization_method;          /* setRegularizationMethod            (DDLO_REG_PLANE) */
  int max_iterations;                 /* setMaximumIterations               (64) */
  int optimizer;                      /* lsq_optimizer_type_                (DDLO_OPT_LEVENBERG_MARQUARDT) */
  int lm_max_iterations;              /* lm_max_iterations_                 (10) */
  int reserved_;
  double max_correspondence_distance; /* setMaxCorrespondenceDistance       (FLT_MAX) */
  double transformation_epsilon;      /* setTransformationEpsilon           (5e-4) */
  double rotation_epsilon;            /* setRotationEpsilon                 (2e-3) */
  double lm_init_lambda_factor;       /* setInitialLambdaFactor             (1e-9) */
} ddlo_params;

typedef struct ddlo_align_result {
  float final_transformation[16]; /* getFinalTransformation(), column-major */
  double final_hessian[36];       /* getFinalHessian(), column-major */
  int flags;                      /* DDLO_FLAG_* */
  int nr_iterations;              /* nr_iterations_ (index of the last outer iteration) */
  int n_linearize;                /* linearize() calls executed   (for the traffic model) */
  int n_compute_error;            /* compute_error() calls executed */
  double final_error;             /* last accepted sum of e^T M e */
  double lm_lambda;               /* lm_lambda_ on exit */
} ddlo_align_result;

/* ---- library ------------------------------------------------------------------------------- */
int ddlo_abi_version(void);
const char* ddlo_last_error(void); /* thread-local text of the last failure */
int ddlo_device_count(int* count);

/* ---- runtime --------------------------------------------------------------------------------- */
int ddlo_runtime_create(int device, ddlo_runtime** out);
int ddlo_runtime_destroy(ddlo_runtime* rt);
int ddlo_runtime_synchronize(ddlo_runtime* rt);
/* Batched workloads (BASELINE config C5): several runtimes (streams) of one device register independent pairs
 * concurrently.  The align kernel is a cooperative launch of one block per SM; limiting each runtime to
 * num_SMs / n_runtimes blocks lets n_runtimes aligns be resident at once instead of queueing behind each other.
 * max_blocks <= 0 restores the default (all SMs).  The fp64 sums are taken per block and then over the blocks in a
 * fixed order, so results are reproducible for a given block count and agree to fp64 rounding between counts. */
int ddlo_runtime_set_align_blocks(ddlo_runtime* rt, int max_blocks);
/* CUDA-event timer on the runtime's stream (bench.py times kernels with these) */
int ddlo_runtime_timer_begin(ddlo_runtime* rt);
int ddlo_runtime_timer_end(ddlo_runtime* rt, float* elapsed_ms); /* synchronises */
/* number of this library's kernels launched on this runtime since creation */
int ddlo_runtime_launch_count(ddlo_runtime* rt, long long* count);
/* write `bytes` of device scratch (L2 flush between timed iterations) */
int ddlo_runtime_flush_l2(ddlo_runtime* rt, size_t bytes);
/* 16 CUDA-event slots on the runtime's stream, for per-stage timing inside one step */
int ddlo_runtime_event_record(ddlo_runtime* rt, int slot);
int ddlo_runtime_event_elapsed(ddlo_runtime* rt, int slot_begin, int slot_end, float* elapsed_ms); /* synchronises on slot_end */
/* page-locked host memory for callers that want true asynchronous uploads */
int ddlo_host_alloc(size_t bytes, void** out);
int ddlo_host_free(void* p);

/* ---- clouds: pcl::PointCloud + nanoflann::KdTreeFLANN ------------------------------------------ */
/* Upload n points from HOST memory; point i starts at (const char*)xyz + i*stride_bytes and holds
 * x,y,z as float (a pcl::PointXYZI array is stride 32).  data[3] is taken as 1.0f. */
int ddlo_cloud_create(ddlo_runtime* rt, const float* xyz, int n, int stride_bytes, ddlo_cloud** out);
/* Same from DEVICE memory (float4 per point, w ignored); copies. */
int ddlo_cloud_create_from_device(ddlo_runtime* rt, const void* d_xyzw, int n, ddlo_cloud** out);
int ddlo_cloud_retain(ddlo_cloud* c);
int ddlo_cloud_release(ddlo_cloud* c);
int ddlo_cloud_size(const ddlo_cloud* c, int* n);
int ddlo_cloud_download(ddlo_cloud* c, float* xyzw_out /* n*4 */);
/* KdTreeFLANN::setInputCloud -> buildIndex (nanoflann.hpp:137-143). Idempotent. */
int ddlo_cloud_build_index(ddlo_cloud* c);
int ddlo_cloud_has_index(const ddlo_cloud* c, int* has);
/* KdTreeFLANN::nearestKSearch (nanoflann.hpp:146-156) for nq HOST queries (x,y,z float, stride in
 * bytes).  Output, HOST: idx[nq*k] (int32, -1 padded), sqdist[nq*k] (float, +inf padded), ascending
 * by (sqdist, idx) — ties are broken by index.  counts (optional) receives min(k, n) per query. */
int ddlo_cloud_knn(ddlo_cloud* c, const float* queries, int nq, int qstride_bytes, int k, int* idx, float* sqdist, int* counts);
/* rigidly transformed copy (pcl::transformPointCloud, float arithmetic), T column-major 4x4 float */
int ddlo_cloud_transform(ddlo_cloud* c, const float* T16, ddlo_cloud** out);
/* concatenation of m clouds (device-side `*submap_cloud_ += *keyframe`, odom.cc:1298-1313) */
int ddlo_cloud_concat(ddlo_runtime* rt, ddlo_cloud* const* parts, int m, ddlo_cloud** out);

/* Handles are bound to their runtime (stream).  ddlo_cloud_share builds the index if missing, synchronises the owner's
 * stream once and marks the cloud immutable; engines of OTHER runtimes of the same device may then take it as input
 * (the lanes of a batch share one submap this way).  ddlo_covs_share does the same for a covariance vector and, given
 * the cloud it is the target covariances of, prepares the Morton-ordered copy the align kernel reads. */
int ddlo_cloud_share(ddlo_cloud* c);
int ddlo_covs_share(ddlo_covs* v, ddlo_cloud* target /* or NULL */);

/* Scan preprocessing on the device (the filters OdomNode::preprocessPoints runs right
 * before this path, odom.cc:442-478, and on every new keyframe, odom.cc:494-499, 1133-1137).
 * pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(lx, ly, lz): one centroid per occupied voxel, voxels in
 * ascending index order, non-finite points dropped.  DDLO_E_UNSUPPORTED when the voxel index space does not fit
 * an int (PCL prints "Leaf size is too small for the input dataset" and passes the cloud through). */
int ddlo_cloud_voxel_filter(ddlo_cloud* c, float leaf_x, float leaf_y, float leaf_z, ddlo_cloud** out);
/* The strided organised down-sample OdomNode::preprocessPoints runs first (odom.cc:445-455): pcl::ExtractIndices with
 * the mask of odom.cc:124-130 - rows 0, row_stride, 2 row_stride, ... and columns 0, col_stride, ... of the
 * height x width scan (row-major) - setNegative(false), setKeepOrganized(true): same size and order, every other
 * point becomes NaN.  The cloud must hold at least width * height points; points behind that are removed too. */
int ddlo_cloud_extract_stride(ddlo_cloud* c, int width, int height, int row_stride, int col_stride, ddlo_cloud** out);
/* pcl::CropBox<PointXYZI>::filter with setMin/setMax (identity box pose): keeps the points inside the box, or
 * outside with negative != 0 (setNegative); keep_organized != 0 (setKeepOrganized) keeps the size and order and
 * overwrites removed points with NaN. */
int ddlo_cloud_crop_box(ddlo_cloud* c, const float* min_xyz, const float* max_xyz, int negative, int keep_organized, ddlo_cloud** out);

/* ---- covariances: std::vector<Eigen::Matrix4d> ---------------------------------------------------- */
/* NanoGICP::calculate_covariances (nano_gicp_impl.hpp:374-441); builds the index if missing. */
int ddlo_covs_compute(ddlo_cloud* c, int k, int regularization_method, ddlo_covs** out);
int ddlo_covs_from_host(ddlo_runtime* rt, const double* mat4x4, int n, ddlo_covs** out);
int ddlo_covs_to_host(ddlo_covs* v, double* mat4x4_out /* n*16 */);
int ddlo_covs_size(const ddlo_covs* v, int* n);
int ddlo_covs_retain(ddlo_covs* v);
int ddlo_covs_release(ddlo_covs* v);
int ddlo_covs_concat(ddlo_runtime* rt, ddlo_covs* const* parts, int m, ddlo_covs** out);

/* ---- the engine: NanoGICP ---------------------------------------------------------------------------- */
int ddlo_gicp_create(ddlo_runtime* rt, ddlo_gicp** out);
int ddlo_gicp_destroy(ddlo_gicp* g);
int ddlo_params_default(ddlo_params* p);
int ddlo_gicp_set_params(ddlo_gicp* g, const ddlo_params* p);
int ddlo_gicp_get_params(const ddlo_gicp* g, ddlo_params* p);

/* setInputSource (nano_gicp_impl.hpp:133-143): no-op for the same handle; else store, build the
 * index, drop source covariances.  build_index = 0 gives registerInputSource (:123-130). */
int ddlo_gicp_set_input_source(ddlo_gicp* g, ddlo_cloud* c, int build_index);
/* setInputTarget (:146-155) */
int ddlo_gicp_set_input_target(ddlo_gicp* g, ddlo_cloud* c);
/* clearSource / clearTarget (:109-120) */
int ddlo_gicp_clear_source(ddlo_gicp* g);
int ddlo_gicp_clear_target(ddlo_gicp* g);
/* setSourceCovariances / setTargetCovariances (:158-169); v == NULL is `source_covs_.clear()`. The
 * handle is shared, not copied: `s2m.source_covs_ = s2s.source_covs_` (odom.cc:765) costs nothing. */
int ddlo_gicp_set_source_covariances(ddlo_gicp* g, ddlo_covs* v);
int ddlo_gicp_set_target_covariances(ddlo_gicp* g, ddlo_covs* v);
/* getSourceCovariances / getTargetCovariances (nano_gicp.hpp:106-114): *out is retained, may be NULL */
int ddlo_gicp_get_source_covariances(ddlo_gicp* g, ddlo_covs** out);
int ddlo_gicp_get_target_covariances(ddlo_gicp* g, ddlo_covs** out);
int ddlo_gicp_get_input_source(ddlo_gicp* g, ddlo_cloud** out);
int ddlo_gicp_get_input_target(ddlo_gicp* g, ddlo_cloud** out);
/* calculateSourceCovariances / calculateTargetCovariances (:172-181) */
int ddlo_gicp_calculate_source_covariances(ddlo_gicp* g);
int ddlo_gicp_calculate_target_covariances(ddlo_gicp* g);
/* swapSourceAndTarget (:98-106) */
int ddlo_gicp_swap_source_and_target(ddlo_gicp* g);

/* pcl::Registration::align -> NanoGICP::computeTransformation -> LsqRegistration::
 * computeTransformation (nano_gicp_impl.hpp:184-196, lsq_registration_impl.hpp:96-232).  The whole
 * LM / GN loop runs on the device; the host sees one launch and one small read-back.
 * guess16 == NULL means identity. */
int ddlo_gicp_align(ddlo_gicp* g, const float* guess16, ddlo_align_result* result);
/* The same in two halves: _async only enqueues the work on the runtime's stream (missing covariances,
 * then the single align kernel); _finish copies the result back and synchronises.  Between the two
 * the host is free, and CUDA events recorded around _async time the device work alone. */
int ddlo_gicp_align_async(ddlo_gicp* g, const float* guess16);
int ddlo_gicp_align_finish(ddlo_gicp* g, ddlo_align_result* result);
/* the `output` cloud of align(): the source moved by the final transformation (:125) */
int ddlo_gicp_aligned_cloud(ddlo_gicp* g, ddlo_cloud** out);

/* Cost-function hooks (protected in the reference; exported so parity tests can compare H, b and
 * the error with the oracle).  T16 is a column-major 4x4 double (Eigen::Isometry3d::matrix()). */
int ddlo_gicp_linearize(ddlo_gicp* g, const double* T16, double* H36, double* b6, double* error);          /* :278-342 */
int ddlo_gicp_compute_error(ddlo_gicp* g, const double* T16, double* error);                                /* :345-371 */
int ddlo_gicp_get_correspondences(ddlo_gicp* g, int* correspondences, float* sq_distances, int capacity);   /* :235-275 */
int ddlo_gicp_get_mahalanobis(ddlo_gicp* g, double* mat4x4_out, int capacity);
/* getResiduals(std::vector<double>&, trans) (:225-232): sqrt(sq_distances_) of the last linearize */
int ddlo_gicp_get_residuals(ddlo_gicp* g, double* out, int capacity);
/* The same without a synchronisation of its own: the conversion and the copy are only enqueued on the runtime's
 * stream, and `out` is complete after the next synchronising call (ddlo_gicp_align_finish, ddlo_runtime_synchronize).
 * Between ddlo_gicp_align_async and ddlo_gicp_align_finish this makes pose and residuals cost one host round trip.
 * Meant for page-locked `out` (ddlo_host_alloc); with pageable memory the copy itself blocks. */
int ddlo_gicp_get_residuals_async(ddlo_gicp* g, double* out, int capacity);
/* The residual cloud OdomNode builds from getResiduals right after the S2M align (odom.cc:804-827; SURVEY.md §8f
 * row 3): the source scan projected into a width x height angular image, theta = atan2(x, z) and
 * phi = atan2(y, sqrt(x^2 + z^2)) both mapped from [angle_min, angle_max) (the fork uses 512 x 512, -60..+60 deg);
 * cell (v, u) holds x, y, z and the residual of the last scan point that falls into it, other cells are zero.
 * out_xyzi: HOST, height * width * 4 floats, row-major. */
int ddlo_gicp_residual_image(ddlo_gicp* g, int width, int height, double angle_min, double angle_max, float* out_xyzi);
/* Range-image segmentation of the organised scan: the stage DetectionModule runs on every frame right after the
 * residual image (OdomNode::applySegmentation, odom.cc:853-857).  One call covers
 *   projectScan        src/detection/detection.cpp:254-329   range of every return from the sensor position T(:,3)
 *   projectResiduals   src/detection/detection.cpp:203-252   residual image (the intensity plane of the residual cloud)
 *   groundRemoval      src/detection/detection.cpp:448-512   slope test between vertical neighbours on the lowest rows
 *   cloudSegmentation  src/detection/detection.cpp:514-546   raster-order seeds inside the `valid_range` window
 *   labelComponents    src/detection/detection.cpp:548-724   4-neighbour flood fill on the beam-angle criterion, then the
 *                                                            segment tests (size, scan lines, distance, height, elevation)
 * and reproduces the reference's results exactly: segment numbering in raster order of the seeds, the flood fill's
 * push order (the min_z / max_z update and the float residual sum of :612-634 depend on it), 999999 for rejected
 * segments, -1 for ground and empty pixels.  Defaults in comments are the reference's (detection.cpp:76-105). */
typedef struct ddlo_segmentation_params {
  int rows, cols;                                                     /* H_, W_: 128, 1024 */
  int ground_rows;                                                    /* groundRows: 30; must be < rows */
  int valid_point_num, min_line_num, valid_line_num;                  /* 15, 5, 5 */
  int window_row_min, window_row_max, window_col_min, window_col_max; /* the hard-coded valid_range lambda (:520-522):
                                                                         156, 356, 156, 356; window_col_max must be < cols
                                                                         (beyond that the reference wraps one-directionally) */
  int scan_in_sensor_frame;                                           /* 0: scan_t is segmentation_scan_t_ (world frame), as projectScan gets it.
                                                                         1: scan_t is segmentation_scan_ (sensor frame) and is first moved by T16
                                                                         on the device like pcl::transformPointCloud in OdomNode::transformScans
                                                                         (odom.cc:957-963): float (r0 x + r1 y) + (r2 z + t), non-finite points kept */
  int unordered_residual_sums;                                        /* 0: avg_residuals_ are float sums in the flood fill's push order, bit for
                                                                         bit the reference's (:632-634, :697).  1 (opt-in): the residuals of a segment
                                                                         are summed in double in no particular order; labels are unchanged, the
                                                                         averages agree with the reference's to float rounding of its own sum
                                                                         (~1e-6 relative), and almost no segment has to replay its queue */
  float ang_bottom;                                                   /* 45: vertical resolution = 2 ang_bottom / (rows - 1) */
  float ground_angle_threshold, minimum_range, sensor_mount_angle, theta; /* 10, 10, 10, 60 deg in rad */
  float min_delta_z, max_delta_z, max_distance, max_elevation;        /* 0.1, 3.0, 20, 2.0 */
} ddlo_segmentation_params;
/* scan_t: HOST, rows * cols points of stride_bytes each (x, y, z floats first; 16 or 32 as for ddlo_cloud_upload), the
 *         segmentation scan in the world frame, row 0 = top; a non-finite coordinate marks "no return".
 * T16: column-major 4x4 float pose.  residuals: HOST rows * cols floats or NULL (projectResiduals not called: every
 *      avg_residuals_ entry is 0, :703-707).
 * Outputs (HOST): label_mat int32, range_mat float, ground_mat int8 (1 ground, -1 no information, 0), each
 * rows * cols and any of them NULL to skip; avg_residuals[label] for label < min(*label_count, avg_capacity) (entries
 * behind *label_count, up to 1024, may be overwritten with zeros);
 * *label_count = label_count_ (labels 1 .. *label_count - 1 are the accepted segments); *device_ms (optional) = time of
 * the kernels alone, between the uploads and the read-back. */
int ddlo_segment_scan(ddlo_runtime* rt, const ddlo_segmentation_params* params, const float* scan_t, int stride_bytes, const float* T16,
                      const float* residuals, int* label_mat, float* range_mat, signed char* ground_mat, double* avg_residuals,
                      int avg_capacity, int* label_count, float* device_ms);

/* The same with projectResiduals fed on the device: the residual cloud of the engine's last align (odom.cc:804-827, as
 * ddlo_gicp_residual_image builds it, width = cols and height = rows, angles in [angle_min, angle_max)) never leaves
 * the GPU; its intensity channel is the residual plane (detection.cpp:240-249).  DDLO_E_NOT_READY before an align. */
int ddlo_gicp_segment_scan(ddlo_gicp* g, const ddlo_segmentation_params* params, const float* scan_t, int stride_bytes, const float* T16,
                           double angle_min, double angle_max, int* label_mat, float* range_mat, signed char* ground_mat,
                           double* avg_residuals, int avg_capacity, int* label_count, float* device_ms);

/* getResiduals(std::vector<Eigen::Vector3f>&, trans) (:199-222) */
int ddlo_gicp_get_residual_vectors(ddlo_gicp* g, const float* T16, float* out_xyz, int capacity);

/* m engines of one runtime aligned back to back on that runtime's stream with a single host synchronisation at the
 * end (kept from ABI 1; the batched workload proper is ddlo_batch_* below). */
int ddlo_gicp_align_batch(ddlo_gicp* const* engines, int m, const float* guesses16 /* m*16 or NULL */, ddlo_align_result* results);

/* ---- batched registrations: BASELINE.json config C5 ------------------------------------------
 * The reference has no batched entry point (OdomNode drives one engine from one callback, odom.cc:745-793); this is
 * the contract of SURVEY.md §8e: many independent (source, target, guess) units on one device, no collective.
 * A batch owns S lanes = S runtimes (CUDA streams) + S engines; each lane's align kernel is limited to
 * align_blocks_per_lane SMs (0: num_SMs / S) so that the lanes' cooperative launches are resident side by side.
 * Units are dealt to the lanes round-robin and run entirely on the device: fresh handles, index build(s),
 * covariances, LM align, result copy; the host only enqueues.  host_threads C++ threads (0 or 1: the calling thread)
 * share the enqueueing, each owning a group of lanes.  One process per GPU (or one batch per device in one process)
 * shards a workload over several GPUs; there is nothing to exchange between them. */
typedef struct ddlo_batch ddlo_batch;
typedef struct ddlo_batch_job {
  int source;      /* id of a staged cloud */
  int target;      /* id of a staged cloud (scan-to-scan unit: its index and covariances are built per unit), or -1 = the
                      batch's shared target (scan-to-map unit) */
  float guess[16]; /* column-major initial guess, as align(output, guess) */
} ddlo_batch_job;
int ddlo_batch_create(int device, int n_lanes, int align_blocks_per_lane, int host_threads, ddlo_batch** out);
int ddlo_batch_destroy(ddlo_batch* b);
int ddlo_batch_info(const ddlo_batch* b, int* n_lanes, int* align_blocks_per_lane, int* host_threads);
int ddlo_batch_set_params(ddlo_batch* b, const ddlo_params* p); /* all lanes */
/* How the align stage of the units runs.
 *   DDLO_BATCH_WAVES (default): the units are taken wave_units at a time (<= 0: keep, default 64).  While the lanes
 *     prepare one wave (handles, indexes, covariances), the previous one is aligned by batched kernels that advance ALL
 *     its problems by one LM round per three ordinary launches (no cooperative launch, no grid barrier; problems at
 *     different iterations share the launches).  A C++ thread owned by the batch drives the rounds between
 *     ddlo_batch_submit and ddlo_batch_wait.  Results are bit-identical to ddlo_gicp_align on an ordinary engine.
 *   DDLO_BATCH_LANES: every unit's align is the single-registration kernel on its lane's stream, limited to
 *     align_blocks_per_lane SMs; results are bit-identical to ddlo_gicp_align with that block limit. */
enum { DDLO_BATCH_LANES = 0, DDLO_BATCH_WAVES = 1 };
int ddlo_batch_set_mode(ddlo_batch* b, int mode, int wave_units);
/* LM rounds launched and completion polls made by the WAVES driver since creation */
int ddlo_batch_stats(ddlo_batch* b, long long* wave_rounds, long long* wave_polls);
/* Stage an input cloud in HBM (as ddlo_cloud_create; SURVEY.md §8d: inputs are pre-staged per GPU before timing). */
int ddlo_batch_stage_cloud(ddlo_batch* b, const float* xyz, int n, int stride_bytes, int* id);
int ddlo_batch_staged_count(const ddlo_batch* b, int* count);
/* One target for all units with target == -1 (a keyframe submap): index and covariances are prepared once
 * (covs_mat4x4 == NULL: computed with the batch's k and regularisation) and shared by the lanes. */
int ddlo_batch_set_shared_target(ddlo_batch* b, int cloud_id, const double* covs_mat4x4);
/* Start m units; returns without waiting for the device.  jobs and results (HOST, m entries each) must stay valid
 * until ddlo_batch_wait, which synchronises everything, fills results and returns the first error.
 * _run = submit + wait. */
int ddlo_batch_submit(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results);
/* The same with the SOURCE scans coming from host memory inside the submission: unit i uploads source_xyz[i]
 * (source_n[i] points of stride_bytes each, as ddlo_cloud_create; page-locked memory makes the copies asynchronous)
 * on its lane's stream instead of taking jobs[i].source; a NULL entry falls back to the staged id.  All arrays must
 * stay valid until ddlo_batch_wait. */
int ddlo_batch_submit_host(ddlo_batch* b, const ddlo_batch_job* jobs, int m, const float* const* source_xyz, const int* source_n, int stride_bytes,
                           ddlo_align_result* results);
int ddlo_batch_wait(ddlo_batch* b);
int ddlo_batch_run(ddlo_batch* b, const ddlo_batch_job* jobs, int m, ddlo_align_result* results);
int ddlo_batch_launch_count(ddlo_batch* b, long long* count); /* kernels launched by all lanes since creation */

/* ---- keyframe store and submap builder -------------------------------------------------------
 * What OdomNode keeps in host vectors around the path - keyframes_ (pose + world-frame cloud, odom.h:105-106) and
 * keyframe_normals_ (their covariances, :107-108) - lives on the device here, and getSubmapKeyframes
 * (odom.cc:1215-1315) becomes: selection on the host (a few scalar operations per keyframe), concatenation on the
 * device.  One store belongs to one runtime. */
typedef struct ddlo_keyframes ddlo_keyframes;
int ddlo_keyframes_create(ddlo_runtime* rt, ddlo_keyframes** out);
int ddlo_keyframes_destroy(ddlo_keyframes* k);
int ddlo_keyframes_count(const ddlo_keyframes* k, int* n);
/* keyframes_.push_back(...) + keyframe_normals_.push_back(getSourceCovariances()) (odom.cc:502-510, 1139-1149):
 * position = pose_, rotation = rotq_ as (w, x, y, z); the handles are shared (retained), nothing is copied. */
int ddlo_keyframes_add(ddlo_keyframes* k, const float* position_xyz, const float* rotation_wxyz, ddlo_cloud* cloud_world, ddlo_covs* covs);
int ddlo_keyframes_get(ddlo_keyframes* k, int index, float* position_xyz, float* rotation_wxyz, ddlo_cloud** cloud, ddlo_covs** covs); /* outputs optional; handles retained */
/* The decision of updateKeyframes (odom.cc:1067-1126) for the current pose: distance to the closest keyframe, rotation
 * against it, the count of keyframes within 1.5 thresh_dist.  Optional outputs may be NULL. */
int ddlo_keyframes_is_new(const ddlo_keyframes* k, const float* position_xyz, const float* rotation_wxyz, float thresh_dist, float thresh_rot_deg,
                          int* is_new, int* closest_index, float* closest_distance, float* rotation_deg);
/* getSubmapKeyframes (odom.cc:1215-1315) for the pose current_xyz (the translation of T_s2s_): the submap_knn nearest
 * keyframes, the submap_kcv nearest among the convex-hull vertices of all keyframe positions and the submap_kcc nearest
 * among the concave-hull vertices (alpha = concave_alpha, the reference sets keyframe_thresh_dist_, odom.cc:1175), ties
 * included as pushSubmapIndices includes them (:1178-1213), sorted and made unique.  *changed = 0: the selection equals
 * the previous one (submap_hasChanged_ = false) and no cloud is built; *changed = 1: *cloud and *covs (new handles, the
 * caller releases them) are the selected keyframes' clouds and covariances concatenated ON THE DEVICE in index order
 * (:1298-1313).  indices / n_indices (optional) receive the selection.  Hulls: pcl::ConvexHull / pcl::ConcaveHull
 * restated on the host (csrc/hull.hpp); for keyframe positions that are three-dimensional in PCL's sense (smallest /
 * largest covariance eigenvalue >= 1e-3) the concave hull is not implemented and contributes nothing. */
int ddlo_keyframes_get_submap(ddlo_keyframes* k, const float* current_xyz, int submap_knn, int submap_kcv, int submap_kcc, double concave_alpha,
                              int* changed, ddlo_cloud** cloud, ddlo_covs** covs, int* indices, int capacity, int* n_indices);
/* keyframe_convex_ / keyframe_concave_ of the last call, and the dimension (2 or 3) the concave hull was taken in */
int ddlo_keyframes_hulls(const ddlo_keyframes* k, int* convex, int* n_convex, int* concave, int* n_concave, int capacity, int* concave_dimension);

#ifdef __cplusplus
}
#endif
#endif /* DDLO_GICP_H */
