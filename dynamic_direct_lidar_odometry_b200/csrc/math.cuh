// Small fp64 linear algebra used by the covariance, Mahalanobis and LM code (device + host).
// These restate what the reference takes from Eigen 3.3 (JacobiSVD of a symmetric 3x3, fixed-size
// inverse, LDLT, Quaternion::toRotationMatrix, Isometry products); see DESIGN.md "arithmetic".
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/ddlo_gicp.h"

namespace ddlo {

#define DDLO_HD __host__ __device__ __forceinline__

// symmetric 3x3 stored as 6 doubles: xx xy xz yy yz zz
struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

// rigid transform: R row-major + t, the affine part of Eigen::Isometry3d
struct Iso3 {
  double r[9];
  double t[3];
};

DDLO_HD Sym3 sym3_inverse(const Sym3& a) {
  const double c00 = a.yy * a.zz - a.yz * a.yz;
  const double c01 = a.xz * a.yz - a.xy * a.zz;
  const double c02 = a.xy * a.yz - a.xz * a.yy;
  const double det = a.xx * c00 + a.xy * c01 + a.xz * c02;
  const double inv = 1.0 / det;
  Sym3 r;
  r.xx = c00 * inv;
  r.xy = c01 * inv;
  r.xz = c02 * inv;
  r.yy = (a.xx * a.zz - a.xz * a.xz) * inv;
  r.yz = (a.xy * a.xz - a.xx * a.yz) * inv;
  r.zz = (a.xx * a.yy - a.xy * a.xy) * inv;
  return r;
}

// R * C * R^T for symmetric C
DDLO_HD Sym3 sym3_rotate(const double* R, const Sym3& c) {
  // A = R * C
  const double a00 = R[0] * c.xx + R[1] * c.xy + R[2] * c.xz, a01 = R[0] * c.xy + R[1] * c.yy + R[2] * c.yz,
               a02 = R[0] * c.xz + R[1] * c.yz + R[2] * c.zz;
  const double a10 = R[3] * c.xx + R[4] * c.xy + R[5] * c.xz, a11 = R[3] * c.xy + R[4] * c.yy + R[5] * c.yz,
               a12 = R[3] * c.xz + R[4] * c.yz + R[5] * c.zz;
  const double a20 = R[6] * c.xx + R[7] * c.xy + R[8] * c.xz, a21 = R[6] * c.xy + R[7] * c.yy + R[8] * c.yz,
               a22 = R[6] * c.xz + R[7] * c.yz + R[8] * c.zz;
  Sym3 r;
  r.xx = a00 * R[0] + a01 * R[1] + a02 * R[2];
  r.xy = a00 * R[3] + a01 * R[4] + a02 * R[5];
  r.xz = a00 * R[6] + a01 * R[7] + a02 * R[8];
  r.yy = a10 * R[3] + a11 * R[4] + a12 * R[5];
  r.yz = a10 * R[6] + a11 * R[7] + a12 * R[8];
  r.zz = a20 * R[6] + a21 * R[7] + a22 * R[8];
  return r;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3. w descending, V columns = eigenvectors
// (row-major V[3*i+j]).  For a symmetric PSD matrix this is JacobiSVD's U, S, V with U == V.
DDLO_HD void sym3_eig(const Sym3& s, double w[3], double V[9]) {
  double a00 = s.xx, a01 = s.xy, a02 = s.xz, a11 = s.yy, a12 = s.yz, a22 = s.zz;
  V[0] = 1, V[1] = 0, V[2] = 0, V[3] = 0, V[4] = 1, V[5] = 0, V[6] = 0, V[7] = 0, V[8] = 1;
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double diag = fabs(a00) + fabs(a11) + fabs(a22);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    // (p,q) = (0,1)
    if (a01 != 0.0) {
      const double theta = (a11 - a00) / (2.0 * a01);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      const double n00 = a00 - t * a01, n11 = a11 + t * a01;
      const double n02 = c * a02 - sn * a12, n12 = sn * a02 + c * a12;
      a00 = n00, a11 = n11, a01 = 0.0, a02 = n02, a12 = n12;
      for (int k = 0; k < 3; ++k) {
        const double vp = V[3 * k + 0], vq = V[3 * k + 1];
        V[3 * k + 0] = c * vp - sn * vq;
        V[3 * k + 1] = sn * vp + c * vq;
      }
    }
    // (0,2)
    if (a02 != 0.0) {
      const double theta = (a22 - a00) / (2.0 * a02);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      const double n00 = a00 - t * a02, n22 = a22 + t * a02;
      const double n01 = c * a01 - sn * a12, n12 = sn * a01 + c * a12;
      a00 = n00, a22 = n22, a02 = 0.0, a01 = n01, a12 = n12;
      for (int k = 0; k < 3; ++k) {
        const double vp = V[3 * k + 0], vq = V[3 * k + 2];
        V[3 * k + 0] = c * vp - sn * vq;
        V[3 * k + 2] = sn * vp + c * vq;
      }
    }
    // (1,2)
    if (a12 != 0.0) {
      const double theta = (a22 - a11) / (2.0 * a12);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
      const double n11 = a11 - t * a12, n22 = a22 + t * a12;
      const double n01 = c * a01 - sn * a02, n02 = sn * a01 + c * a02;
      a11 = n11, a22 = n22, a12 = 0.0, a01 = n01, a02 = n02;
      for (int k = 0; k < 3; ++k) {
        const double vp = V[3 * k + 1], vq = V[3 * k + 2];
        V[3 * k + 1] = c * vp - sn * vq;
        V[3 * k + 2] = sn * vp + c * vq;
      }
    }
  }
  w[0] = a00, w[1] = a11, w[2] = a22;
  // sort descending (3-element network), permuting columns of V
#define DDLO_SWAPCOL(i, j)                       \
  if (w[j] > w[i]) {                             \
    double tw = w[i];                            \
    w[i] = w[j];                                 \
    w[j] = tw;                                   \
    for (int k = 0; k < 3; ++k) {                \
      double tv = V[3 * k + i];                  \
      V[3 * k + i] = V[3 * k + j];               \
      V[3 * k + j] = tv;                         \
    }                                            \
  }
  DDLO_SWAPCOL(0, 1)
  DDLO_SWAPCOL(0, 2)
  DDLO_SWAPCOL(1, 2)
#undef DDLO_SWAPCOL
}

// V diag(values) V^T
DDLO_HD Sym3 sym3_recompose(const double V[9], const double v[3]) {
  Sym3 r;
  r.xx = V[0] * v[0] * V[0] + V[1] * v[1] * V[1] + V[2] * v[2] * V[2];
  r.xy = V[0] * v[0] * V[3] + V[1] * v[1] * V[4] + V[2] * v[2] * V[5];
  r.xz = V[0] * v[0] * V[6] + V[1] * v[1] * V[7] + V[2] * v[2] * V[8];
  r.yy = V[3] * v[0] * V[3] + V[4] * v[1] * V[4] + V[5] * v[2] * V[5];
  r.yz = V[3] * v[0] * V[6] + V[4] * v[1] * V[7] + V[5] * v[2] * V[8];
  r.zz = V[6] * v[0] * V[6] + V[7] * v[1] * V[7] + V[8] * v[2] * V[8];
  return r;
}

// regularisation of a raw neighbourhood covariance (nano_gicp_impl.hpp:401-437)
DDLO_HD Sym3 regularize_cov(const Sym3& cov, int method) {
  if (method == DDLO_REG_NONE) return cov;
  if (method == DDLO_REG_FROBENIUS) {
    Sym3 C = cov;
    C.xx += 1e-3;
    C.yy += 1e-3;
    C.zz += 1e-3;
    Sym3 Ci = sym3_inverse(C);
    const double nrm = sqrt(Ci.xx * Ci.xx + Ci.yy * Ci.yy + Ci.zz * Ci.zz + 2.0 * (Ci.xy * Ci.xy + Ci.xz * Ci.xz + Ci.yz * Ci.yz));
    Ci.xx /= nrm, Ci.xy /= nrm, Ci.xz /= nrm, Ci.yy /= nrm, Ci.yz /= nrm, Ci.zz /= nrm;
    return sym3_inverse(Ci);
  }
  double w[3], V[9], v[3];
  sym3_eig(cov, w, V);
  if (method == DDLO_REG_PLANE) {
    v[0] = 1.0, v[1] = 1.0, v[2] = 1e-3;
  } else if (method == DDLO_REG_MIN_EIG) {
    v[0] = fmax(w[0], 1e-3), v[1] = fmax(w[1], 1e-3), v[2] = fmax(w[2], 1e-3);
  } else {  // NORMALIZED_MIN_EIG
    const double wmax = fmax(w[0], fmax(w[1], w[2]));
    v[0] = fmax(w[0] / wmax, 1e-3), v[1] = fmax(w[1] / wmax, 1e-3), v[2] = fmax(w[2] / wmax, 1e-3);
  }
  return sym3_recompose(V, v);
}

// ---- Isometry helpers ---------------------------------------------------------------------------
DDLO_HD Iso3 iso_mul(const Iso3& a, const Iso3& b) {
  Iso3 r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
    r.t[i] = a.r[3 * i] * b.t[0] + a.r[3 * i + 1] * b.t[1] + a.r[3 * i + 2] * b.t[2] + a.t[i];
  }
  return r;
}

// so3.hpp:101-124 then Quaterniond::toRotationMatrix
DDLO_HD void so3_exp_matrix(const double* omega, double* R) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq);
    const double half_theta = 0.5 * theta;
    imag_factor = sin(half_theta) / theta;
    real_factor = cos(half_theta);
  }
  const double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

// LDL^T with symmetric pivoting (the algorithm of Eigen::LDLT), solve A x = rhs, A 6x6 row-major
DDLO_HD void ldlt6_solve(const double* A_in, const double* rhs, double* x) {
  double A[36];
  int perm[6];
  for (int i = 0; i < 36; ++i) A[i] = A_in[i];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double big = fabs(A[7 * k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[7 * i]) > big) {
        big = fabs(A[7 * i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < 6; ++j) {
        const double tmp = A[6 * k + j];
        A[6 * k + j] = A[6 * piv + j];
        A[6 * piv + j] = tmp;
      }
      for (int i = 0; i < 6; ++i) {
        const double tmp = A[6 * i + k];
        A[6 * i + k] = A[6 * i + piv];
        A[6 * i + piv] = tmp;
      }
      const int tp = perm[k];
      perm[k] = perm[piv];
      perm[piv] = tp;
    }
    const double d = A[7 * k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < 6; ++i) A[6 * i + k] /= d;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) {
        A[6 * i + j] -= A[6 * i + k] * d * A[6 * j + k];
        A[6 * j + i] = A[6 * i + j];
      }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = rhs[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) y[i] -= A[6 * i + j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] = A[7 * i] != 0.0 ? y[i] / A[7 * i] : 0.0;
  for (int i = 5; i >= 0; --i)
    for (int j = i + 1; j < 6; ++j) y[i] -= A[6 * j + i] * y[j];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

// The same solve for a symmetric POSITIVE DEFINITE matrix (what H + lambda I is in practice), with
// every index known at compile time so that the whole factorisation lives in registers: LDL^T in
// natural order, which is backward stable for SPD input.  Any non-positive pivot hands over to
// the pivoting routine above.
DDLO_HD void ldlt6_solve_fast(const double* A_in, const double* rhs, double* x) {
  double L[6][6], D[6];
  bool spd = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A_in[7 * j];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < j) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    spd = spd && (d > 0.0);
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i > j) {
        double v = A_in[6 * i + j];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < j) v -= L[i][k] * L[j][k] * D[k];
        L[i][j] = v * inv;
      }
  }
  if (!spd) {
    ldlt6_solve(A_in, rhs, x);
    return;
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = rhs[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < i) v -= L[i][k] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > i) v -= L[k][i] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

}  // namespace ddlo
