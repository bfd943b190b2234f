/*
 * ddlo_gicp_testing.h — host-callable copies of the device arithmetic, exported by
 * libddlo_gicp_b200.so for CPU-side unit tests (they compile the very same __host__ __device__
 * functions the kernels run, csrc/math.cuh).  Not part of the drop-in boundary.
 */
#ifndef DDLO_GICP_TESTING_H
#define DDLO_GICP_TESTING_H
#ifdef __cplusplus
extern "C" {
#endif
/* symmetric 3x3 as 6 doubles xx,xy,xz,yy,yz,zz; V row-major, eigenvalues descending */
void ddlo_math_sym3_eig(const double* sym6, double* w3, double* V9);
void ddlo_math_regularize(const double* sym6, int method, double* out6); /* nano_gicp_impl.hpp:401-437 */
void ddlo_math_ldlt6_solve(const double* A36, const double* rhs6, double* x6); /* Eigen::LDLT, lsq_registration_impl.hpp:190 */
void ddlo_math_ldlt6_solve_fast(const double* A36, const double* rhs6, double* x6); /* register-resident SPD path of the kernel */
void ddlo_math_so3_exp(const double* omega3, double* R9);                       /* gicp/so3.hpp:101-124 */
void ddlo_math_sym3_inverse(const double* sym6, double* out6);
/* the hull functions of the keyframe store (csrc/hull.hpp) on n points (x, y, z doubles): number of hull vertices,
 * indices (sorted) written up to capacity; ddlo_hull_concave returns -3 for points that are 3-dimensional in PCL's sense */
int ddlo_hull_convex(const double* xyz, int n, int* out_indices, int capacity);
int ddlo_hull_concave(const double* xyz, int n, double alpha, int* out_indices, int capacity);
int ddlo_hull_dimension(const double* xyz, int n);
struct ddlo_gicp;
/* bytes of the result record one align copies device -> host */
int ddlo_align_d2h_bytes(void);
/* Profiling of the align kernel is OFF by default (no timestamps taken, no profiling memory touched); on != 0 makes
 * the following aligns of this engine record the two tables below. */
int ddlo_gicp_debug_enable(struct ddlo_gicp* g, int on);
/* phase timeline of the last align (block 0): entries are tag << 56 | %globaltimer ns; returns the count */
/* per-block phase times (ns) of the first 8 linearize passes: out[pass][capacity_blocks][8]; returns block count */
int ddlo_gicp_debug_block_times(struct ddlo_gicp* g, unsigned long long* out, int capacity_blocks);
int ddlo_gicp_debug_timeline(struct ddlo_gicp* g, unsigned long long* out, int capacity);
/* builds with -DDDLO_VISIT_STATS only: search statistics of the first 4 linearize passes, out[4][ns][4] ints
 * {node visits, leaf scans, lock-step warp steps, 0}; returns ns, or 0 in a regular build */
int ddlo_gicp_debug_visits(struct ddlo_gicp* g, int* out, int capacity_points);
#ifdef __cplusplus
}
#endif
#endif
