// hd.h -- scalar math shared by every kernel of the GICP / VoxelGrid path.
//
// Everything here is `__host__ __device__` so that the per-point logic the
// kernels run can also be compiled by g++ into the CPU test harness
// (tests/hd_harness.cpp) and checked against the oracle without a GPU.  The
// product library only ever calls these from CUDA kernels and from the small
// host-side driver; there is no CPU compute path in the product.
//
// Precision contract (mirrors the reference, see DESIGN.md "Numerics"):
//   * points are float32; T*p is evaluated in float32 with the association
//     ((c0*x + c1*y) + c2*z) + c3 and NO fused multiply-add
//     (reference: Matrix4f * Vector4f, gicp.hpp:307,341,382,469);
//   * residuals, covariances, Mahalanobis matrices and all sums are double
//     (gicp.hpp:310-314, 484-493);
//   * the whole library is compiled with -fmad=false so float/double products
//     and sums round exactly like the reference's x86-64 build.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#define LB_D __device__ __forceinline__
#else
#define LB_HD inline
#define LB_D inline
#endif

namespace lb {

struct f4 { float x, y, z, w; };  // same layout as CUDA float4

LB_HD float bits_to_float(int32_t i) { union { int32_t i; float f; } u; u.i = i; return u.f; }
LB_HD int32_t float_to_bits(float f) { union { int32_t i; float f; } u; u.f = f; return u.i; }

// ---------------------------------------------------------------- rigid transform
// Row-major 3x4 float [R|t].  ((r0*x + r1*y) + r2*z) + t  (Eigen coefficient order).
LB_HD void xform(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  float v;
  v = T[0] * x; v = v + T[1] * y; v = v + T[2] * z; ox = v + T[3];
  v = T[4] * x; v = v + T[5] * y; v = v + T[6] * z; oy = v + T[7];
  v = T[8] * x; v = v + T[9] * y; v = v + T[10] * z; oz = v + T[11];
}
// pcl::transformPointCloud (PCL 1.10 Transformer::se3): p0 + (p1 + (p2 + t)).
LB_HD void xform_pcl(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = T[0] * x + (T[1] * y + (T[2] * z + T[3]));
  oy = T[4] * x + (T[5] * y + (T[6] * z + T[7]));
  oz = T[8] * x + (T[9] * y + (T[10] * z + T[11]));
}

// Trigonometry of one optimiser state x: half-angle cos/sin (float, for applyState) and
// full-angle cos/sin (double, for the rotation derivatives).  Kept separate from its users so the
// persistent kernel can evaluate the 12 values on 12 lanes of the leader warp in parallel.
// index 0,1,2 = roll(x3), pitch(x4), yaw(x5).
struct Trig {
  float ch[3], sh[3];
  double c[3], s[3];
};
LB_HD float half_angle(const double* x, int k) { return 0.5f * (float)x[3 + k]; }
// sin/cos of the half angles are taken in double and rounded to float: that is the correctly
// rounded float result on host and device alike (glibc sinf/cosf used by the reference are
// correctly rounded in all but vanishingly few cases).
LB_HD void trig_compute(const double* x, Trig& t) {
  for (int k = 0; k < 3; k++) {
    double h = (double)half_angle(x, k);
    t.ch[k] = (float)cos(h); t.sh[k] = (float)sin(h);
    t.c[k] = cos(x[3 + k]); t.s[k] = sin(x[3 + k]);
  }
}

// applyState on identity (gicp.hpp:619-634): R = Rz(x5) Ry(x4) Rx(x3) built in
// float32 through Eigen's AngleAxisf -> Quaternionf products -> toRotationMatrix.
LB_HD void apply_state_trig(const double* x, const Trig& t, float* T /*12: row-major 3x4*/) {
  float cz = t.ch[2], sz = t.sh[2];
  float cy = t.ch[1], sy = t.sh[1];
  float cx = t.ch[0], sx = t.sh[0];
  float w1 = cz * cy, x1 = -(sz * sy), y1 = cz * sy, z1 = sz * cy;
  float qw = w1 * cx - x1 * sx;
  float qx = w1 * sx + x1 * cx;
  float qy = y1 * cx + z1 * sx;
  float qz = z1 * cx - y1 * sx;
  float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
  float twx = tx * qw, twy = ty * qw, twz = tz * qw;
  float txx = tx * qx, txy = ty * qx, txz = tz * qx;
  float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  T[0] = 1.0f - (tyy + tzz); T[1] = txy - twz;          T[2] = txz + twy;           T[3] = (float)x[0];
  T[4] = txy + twz;          T[5] = 1.0f - (txx + tzz); T[6] = tyz - twx;           T[7] = (float)x[1];
  T[8] = txz - twy;          T[9] = tyz + twx;          T[10] = 1.0f - (txx + tyy); T[11] = (float)x[2];
}
LB_HD void apply_state(const double* x, float* T) {
  Trig t;
  trig_compute(x, t);
  apply_state_trig(x, t, T);
}

// State from a transform (gicp.hpp:235-241), float entries promoted to double.
LB_HD void state_from_transform(const float* T, double* x) {
  x[0] = T[3]; x[1] = T[7]; x[2] = T[11];
  x[3] = atan2((double)T[9], (double)T[10]);
  x[4] = asin(-(double)T[8]);
  x[5] = atan2((double)T[4], (double)T[0]);
}

// dR/d(phi,theta,psi) closed forms (gicp.hpp:175-209), row-major 3x3 each.
LB_HD void r_derivatives_trig(const Trig& t, double* dP, double* dT, double* dS) {
  double cphi = t.c[0], sphi = t.s[0];
  double ctheta = t.c[1], stheta = t.s[1];
  double cpsi = t.c[2], spsi = t.s[2];
  dP[0] = 0.; dP[3] = 0.; dP[6] = 0.;
  dP[1] = sphi * spsi + cphi * cpsi * stheta;
  dP[4] = -cpsi * sphi + cphi * spsi * stheta;
  dP[7] = cphi * ctheta;
  dP[2] = cphi * spsi - cpsi * sphi * stheta;
  dP[5] = -cphi * cpsi - sphi * spsi * stheta;
  dP[8] = -ctheta * sphi;

  dT[0] = -cpsi * stheta; dT[3] = -spsi * stheta; dT[6] = -ctheta;
  dT[1] = cpsi * ctheta * sphi; dT[4] = ctheta * sphi * spsi; dT[7] = -sphi * stheta;
  dT[2] = cphi * cpsi * ctheta; dT[5] = cphi * ctheta * spsi; dT[8] = -cphi * stheta;

  dS[0] = -ctheta * spsi; dS[3] = cpsi * ctheta; dS[6] = 0.;
  dS[1] = -cphi * cpsi - sphi * spsi * stheta;
  dS[4] = -cphi * spsi + cpsi * sphi * stheta;
  dS[7] = 0.;
  dS[2] = cpsi * sphi - cphi * spsi * stheta;
  dS[5] = sphi * spsi + cphi * cpsi * stheta;
  dS[8] = 0.;
}

LB_HD void r_derivatives(const double* x, double* dP, double* dT, double* dS) {
  Trig t;
  trig_compute(x, t);
  r_derivatives_trig(t, dP, dT, dS);
}

// g[3..5] = tr(dR_k * Rhat)  (computeRDerivative + matricesInnerProd, gicp.hpp:211-213, gicp.h:361-370)
LB_HD void rotation_gradient_trig(const Trig& t, const double* Rhat /*row-major 3x3*/, double* g) {
  double dP[9], dT[9], dS[9];
  r_derivatives_trig(t, dP, dT, dS);
  double r0 = 0., r1 = 0., r2 = 0.;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      r0 += dP[j * 3 + i] * Rhat[i * 3 + j];
      r1 += dT[j * 3 + i] * Rhat[i * 3 + j];
      r2 += dS[j * 3 + i] * Rhat[i * 3 + j];
    }
  g[3] = r0; g[4] = r1; g[5] = r2;
}
LB_HD void rotation_gradient(const double* x, const double* Rhat, double* g) {
  Trig t;
  trig_compute(x, t);
  rotation_gradient_trig(t, Rhat, g);
}

// ---------------------------------------------------------------- symmetric 3x3
// Symmetric matrices are stored as 6 doubles: [xx, xy, xz, yy, yz, zz].
enum { SXX = 0, SXY = 1, SXZ = 2, SYY = 3, SYZ = 4, SZZ = 5 };

template <int P, int Q, int R>
LB_HD void jacobi_rotate(double (&A)[3][3], double (&V)[3][3], int sweep) {
  double apq = A[P][Q];
  if (apq == 0.0) return;
  double g = 100.0 * fabs(apq);
  if (sweep > 3 && fabs(A[P][P]) + g == fabs(A[P][P]) && fabs(A[Q][Q]) + g == fabs(A[Q][Q])) {
    A[P][Q] = A[Q][P] = 0.0;
    return;
  }
  double h = A[Q][Q] - A[P][P];
  double t;
  if (fabs(h) + g == fabs(h)) {
    t = apq / h;
  } else {
    double theta = 0.5 * h / apq;
    t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
    if (theta < 0.0) t = -t;
  }
  double c = 1.0 / sqrt(1.0 + t * t);
  double s = t * c;
  double tau = s / (1.0 + c);
  double hh = t * apq;
  A[P][P] -= hh;
  A[Q][Q] += hh;
  A[P][Q] = A[Q][P] = 0.0;
  double arp = A[R][P], arq = A[R][Q];
  A[R][P] = A[P][R] = arp - s * (arq + arp * tau);
  A[R][Q] = A[Q][R] = arq + s * (arp - arq * tau);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double vkp = V[k][P], vkq = V[k][Q];
    V[k][P] = vkp - s * (vkq + vkp * tau);
    V[k][Q] = vkq + s * (vkp - vkq * tau);
  }
}

// GICP covariance regularisation (gicp.hpp:139-153): SVD of the symmetric
// sample covariance, singular values replaced by (1, 1, eps).  cov is the full
// symmetric 3x3; out is [xx xy xz yy yz zz] of  sum_k v_k u_k u_k'.
LB_HD void regularise_cov(double (&A)[3][3], double eps, double* out6) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
    jacobi_rotate<0, 1, 2>(A, V, sweep);
    jacobi_rotate<0, 2, 1>(A, V, sweep);
    jacobi_rotate<1, 2, 0>(A, V, sweep);
  }
  double d0 = fabs(A[0][0]), d1 = fabs(A[1][1]), d2 = fabs(A[2][2]);
  // stable descending order of |d|: find order (o0,o1,o2)
  int o0 = 0, o1 = 1, o2 = 2;
  // insertion sort, identical tie behaviour to the oracle (strict '<' moves)
  if (d0 < d1) { int t = o0; o0 = o1; o1 = t; double td = d0; d0 = d1; d1 = td; }
  if (d1 < d2) {
    int t = o1; o1 = o2; o2 = t; double td = d1; d1 = d2; d2 = td;
    if (d0 < d1) { t = o0; o0 = o1; o1 = t; td = d0; d0 = d1; d1 = td; }
  }
  double u0[3], u1[3], u2[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    u0[r] = (o0 == 0) ? V[r][0] : ((o0 == 1) ? V[r][1] : V[r][2]);
    u1[r] = (o1 == 0) ? V[r][0] : ((o1 == 1) ? V[r][1] : V[r][2]);
    u2[r] = (o2 == 0) ? V[r][0] : ((o2 == 1) ? V[r][1] : V[r][2]);
  }
  // out(r,c) = ((0 + (1*u0r)*u0c) + (1*u1r)*u1c) + (eps*u2r)*u2c   for r <= c
  const int RR[6] = {0, 0, 0, 1, 1, 2};
  const int CC[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) {
    int r = RR[e], c = CC[e];
    double acc = 0.0;
    acc += (1.0 * u0[r]) * u0[c];
    acc += (1.0 * u1[r]) * u1[c];
    acc += (eps * u2[r]) * u2[c];
    out6[e] = acc;
  }
}

// Sample covariance from accumulated moments (gicp.hpp:129-137), then regularise.
// mean[3] = sum p ; m2 = [xx, yx, yy, zx, zy, zz] sums of float products.
LB_HD void cov_from_moments(const double* sum, const double* m2, int k, double eps, double* out6) {
  double kk = (double)k;
  double mx = sum[0] / kk, my = sum[1] / kk, mz = sum[2] / kk;
  double A[3][3];
  A[0][0] = m2[0] / kk - mx * mx;
  A[1][0] = m2[1] / kk - my * mx;
  A[1][1] = m2[2] / kk - my * my;
  A[2][0] = m2[3] / kk - mz * mx;
  A[2][1] = m2[4] / kk - mz * my;
  A[2][2] = m2[5] / kk - mz * mz;
  A[0][1] = A[1][0]; A[0][2] = A[2][0]; A[1][2] = A[2][1];
  regularise_cov(A, eps, out6);
}

// From-normals covariance (gicp.hpp:81-82; external CalculateCovarianceFromNormals,
// parity unpinned): C = I - (1-eps) n n'.
LB_HD void cov_from_normal(float nx, float ny, float nz, double eps, double* out6) {
  double n[3] = {(double)nx, (double)ny, (double)nz};
  double s = 1.0 - eps;
  out6[SXX] = 1.0 - s * n[0] * n[0];
  out6[SXY] = 0.0 - s * n[0] * n[1];
  out6[SXZ] = 0.0 - s * n[0] * n[2];
  out6[SYY] = 1.0 - s * n[1] * n[1];
  out6[SYZ] = 0.0 - s * n[1] * n[2];
  out6[SZZ] = 1.0 - s * n[2] * n[2];
}

// the BodyFilter nodelet (point_cloud_filter/src/body_filter.cc:28-56 = pcl::CropBox<PointXYZI> with
// setNegative(true) and a rotation about z) folded into the VoxelGrid's load predicate.  ia, ib = cos, sin of the
// rotation divided by cos^2 + sin^2 (the 3x3 inverse PCL takes of its float32 rotation matrix); a point whose local
// coordinates lie inside [min, max] is removed.  Parity unpinned in the last bit (PCL absent).
struct BodyBox {
  int enabled;
  float ia, ib;
  float mn[3], mx[3];
};
LB_HD bool body_box_drops(const BodyBox& b, float x, float y, float z) {
  if (!b.enabled) return false;
  float lx = b.ia * x + b.ib * y, ly = b.ia * y - b.ib * x, lz = z;
  bool outside = (lx < b.mn[0] || ly < b.mn[1] || lz < b.mn[2]) || (lx > b.mx[0] || ly > b.mx[1] || lz > b.mx[2]);
  return !outside;
}

// ---------------------------------------------------------------------------------------------------------------
// per-point surface normal the way point_cloud_filter::NormalComputation obtains it
// (normal_computation.cc:26-59 -> pcl::NormalEstimationOMP<PointXYZI, Normal>, k-NN mode).  PCL is not in the
// reference tree; this restates PCL 1.10's published algorithm in float32 like PCL runs it:
//   computeMeanAndCovarianceMatrix (common/impl/centroid.hpp, dense branch): nine float accumulators over the
//     neighbours in search order, divided by the count, covariance = E[xx'] - mu mu';
//   solvePlaneParameters (features/impl/feature.hpp) -> eigen33 (common/impl/eigen.hpp): smallest eigenvalue by the
//     closed-form cubic (computeRoots / computeRoots2), eigenvector = largest of the three row cross products of
//     (A - lambda I), curvature = |lambda / trace|;
//   flipNormalTowardsViewpoint (features/normal_3d.h): negate when (vp - p) . n < 0.
// Parity unpinned in the last bits (atan2f / cosf / sinf of the platform's libm); everything else is IEEE-exact.
struct NormalAccum {       // accu[0..8] of computeMeanAndCovarianceMatrix
  float a[9];
  LB_HD void reset() { for (int i = 0; i < 9; i++) a[i] = 0.f; }
  LB_HD void add(float x, float y, float z) {
    a[0] = a[0] + x * x; a[1] = a[1] + x * y; a[2] = a[2] + x * z;
    a[3] = a[3] + y * y; a[4] = a[4] + y * z; a[5] = a[5] + z * z;
    a[6] = a[6] + x; a[7] = a[7] + y; a[8] = a[8] + z;
  }
};

LB_HD void pcl_compute_roots2(float b, float c, float* roots) {
  roots[0] = 0.f;
  float d = b * b - 4.0f * c;
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

LB_HD void pcl_compute_roots(const float m[3][3], float* roots) {
  // characteristic polynomial det(x I - A) = x^3 - c2 x^2 + c1 x - c0
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
             m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
             m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (fabsf(c0) < 1.1920929e-07f) {          // std::numeric_limits<float>::epsilon(): one root is 0
    pcl_compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = sqrtf(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  float rho = sqrtf(-a_over_3);
  float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  float cos_theta = cosf(theta);
  float sin_theta = sinf(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) { float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  if (roots[1] >= roots[2]) {
    float t = roots[1]; roots[1] = roots[2]; roots[2] = t;
    if (roots[0] >= roots[1]) { float u = roots[0]; roots[0] = roots[1]; roots[1] = u; }
  }
  if (roots[0] <= 0.0f) pcl_compute_roots2(c2, c1, roots);   // a PSD matrix has no negative eigenvalue: drop to the quadratic
}

// out4 = (nx, ny, nz, curvature) of the neighbourhood accumulated in `acc` (count points), flipped towards `vp`
// as seen from the query point (px, py, pz).
LB_HD void pcl_normal_from_accum(const NormalAccum& acc, int count, float px, float py, float pz, const float* vp,
                                 float* out4) {
  float a[9];
  const float cnt = (float)count;
  for (int i = 0; i < 9; i++) a[i] = acc.a[i] / cnt;
  float C[3][3];
  C[0][0] = a[0] - a[6] * a[6];
  C[0][1] = a[1] - a[6] * a[7];
  C[0][2] = a[2] - a[6] * a[8];
  C[1][1] = a[3] - a[7] * a[7];
  C[1][2] = a[4] - a[7] * a[8];
  C[2][2] = a[5] - a[8] * a[8];
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  // eigen33: scale to [-1, 1]
  float scale = 0.f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { float v = fabsf(C[r][c]); if (v > scale) scale = v; }
  if (scale <= 1.17549435e-38f) scale = 1.0f;           // std::numeric_limits<float>::min()
  float S[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) S[r][c] = C[r][c] / scale;
  float roots[3];
  pcl_compute_roots(S, roots);
  const float eigenvalue = roots[0] * scale;
  S[0][0] = S[0][0] - roots[0]; S[1][1] = S[1][1] - roots[0]; S[2][2] = S[2][2] - roots[0];
  float v1[3] = {S[0][1] * S[1][2] - S[0][2] * S[1][1], S[0][2] * S[1][0] - S[0][0] * S[1][2], S[0][0] * S[1][1] - S[0][1] * S[1][0]};
  float v2[3] = {S[0][1] * S[2][2] - S[0][2] * S[2][1], S[0][2] * S[2][0] - S[0][0] * S[2][2], S[0][0] * S[2][1] - S[0][1] * S[2][0]};
  float v3[3] = {S[1][1] * S[2][2] - S[1][2] * S[2][1], S[1][2] * S[2][0] - S[1][0] * S[2][2], S[1][0] * S[2][1] - S[1][1] * S[2][0]};
  float l1 = (v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2];
  float l2 = (v2[0] * v2[0] + v2[1] * v2[1]) + v2[2] * v2[2];
  float l3 = (v3[0] * v3[0] + v3[1] * v3[1]) + v3[2] * v3[2];
  float n[3];
  if (l1 >= l2 && l1 >= l3) { float s = sqrtf(l1); n[0] = v1[0] / s; n[1] = v1[1] / s; n[2] = v1[2] / s; }
  else if (l2 >= l1 && l2 >= l3) { float s = sqrtf(l2); n[0] = v2[0] / s; n[1] = v2[1] / s; n[2] = v2[2] / s; }
  else { float s = sqrtf(l3); n[0] = v3[0] / s; n[1] = v3[1] / s; n[2] = v3[2] / s; }
  const float eig_sum = (C[0][0] + C[1][1]) + C[2][2];
  float curvature = (eig_sum != 0.f) ? fabsf(eigenvalue / eig_sum) : 0.f;
  // flipNormalTowardsViewpoint
  float dx = vp[0] - px, dy = vp[1] - py, dz = vp[2] - pz;
  float cos_theta = (dx * n[0] + dy * n[1]) + dz * n[2];
  if (cos_theta < 0.f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  out4[0] = n[0]; out4[1] = n[1]; out4[2] = n[2]; out4[3] = curvature;
}

// Mahalanobis matrix M = (R C1 R' + C2)^-1  (gicp.hpp:484-493).  R row-major
// 3x3 double; C1, C2, M symmetric-6.  Products follow the reference order
// (R*C1 first, then *R'), inverse by cofactors / determinant like
// Eigen::Matrix3d::inverse().
LB_HD void mahalanobis(const double* R, const double* C1, const double* C2, double* M) {
  double c[3][3] = {{C1[SXX], C1[SXY], C1[SXZ]}, {C1[SXY], C1[SYY], C1[SYZ]}, {C1[SXZ], C1[SYZ], C1[SZZ]}};
  double rc[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) rc[i][j] = (R[i * 3 + 0] * c[0][j] + R[i * 3 + 1] * c[1][j]) + R[i * 3 + 2] * c[2][j];
  double t[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++)
      t[i][j] = (rc[i][0] * R[j * 3 + 0] + rc[i][1] * R[j * 3 + 1]) + rc[i][2] * R[j * 3 + 2];
  double a = t[0][0] + C2[SXX], b = t[0][1] + C2[SXY], cc = t[0][2] + C2[SXZ];
  double d = t[1][1] + C2[SYY], e = t[1][2] + C2[SYZ], f = t[2][2] + C2[SZZ];
  // symmetric matrix [[a b cc],[b d e],[cc e f]]
  double k00 = d * f - e * e;
  double k01 = e * cc - b * f;   // cofactor(1,0) = m(2,1)m(0,2) - m(2,2)m(0,1)
  double k02 = b * e - cc * d;   // cofactor(2,0) = m(0,1)m(1,2) - m(0,2)m(1,1)
  double det = (k00 * a + k01 * b) + k02 * cc;
  double inv = 1.0 / det;
  M[SXX] = k00 * inv;
  M[SXY] = k01 * inv;
  M[SXZ] = k02 * inv;
  M[SYY] = (f * a - cc * cc) * inv;
  M[SYZ] = (cc * b - e * a) * inv;  // cofactor(2,1) = m(0,2)m(1,0) - m(0,0)m(1,2)
  M[SZZ] = (a * d - b * b) * inv;
}

// ---------------------------------------------------------------- objective
// Per-correspondence contribution to the 13 sums of the GICP objective
// (gicp.hpp:373-397): f += r'Mr ; gt += Mr ; Rhat += p (Mr)'.
// T: row-major 3x4 float; p: source point (already guess-transformed);
// q: matched target point; M symmetric-6 (all-zero M => exact zero contribution).
LB_HD void objective_terms(const float* T, float px, float py, float pz, float qx, float qy, float qz,
                           const double* M, double* acc /*13*/) {
  float tx, ty, tz;
  xform(T, px, py, pz, tx, ty, tz);
  double r0 = (double)(tx - qx), r1 = (double)(ty - qy), r2 = (double)(tz - qz);
  double t0 = (M[SXX] * r0 + M[SXY] * r1) + M[SXZ] * r2;
  double t1 = (M[SXY] * r0 + M[SYY] * r1) + M[SYZ] * r2;
  double t2 = (M[SXZ] * r0 + M[SYZ] * r1) + M[SZZ] * r2;
  acc[0] += (r0 * t0 + r1 * t1) + r2 * t2;
  acc[1] += t0; acc[2] += t1; acc[3] += t2;
  double dx = (double)px, dy = (double)py, dz = (double)pz;
  acc[4] += dx * t0; acc[5] += dx * t1; acc[6] += dx * t2;
  acc[7] += dy * t0; acc[8] += dy * t1; acc[9] += dy * t2;
  acc[10] += dz * t0; acc[11] += dz * t1; acc[12] += dz * t2;
}

// Turn the 13 reduced sums into f and the 6-gradient (gicp.hpp:398-401).
LB_HD void objective_finish_trig(const double* sums /*13*/, int m, const Trig& t, double* f, double* g /*6*/) {
  *f = sums[0] / (double)m;
  double sc = 2.0 / m;
  g[0] = sums[1] * sc; g[1] = sums[2] * sc; g[2] = sums[3] * sc;
  double Rhat[9];
  for (int i = 0; i < 9; i++) Rhat[i] = sums[4 + i] * sc;
  rotation_gradient_trig(t, Rhat, g);
}
LB_HD void objective_finish(const double* sums /*13*/, int m, const double* x, double* f, double* g /*6*/) {
  Trig t;
  trig_compute(x, t);
  objective_finish_trig(sums, m, t, f, g);
}

// ---------------------------------------------------------------- moment form of the objective
// For FIXED correspondences (one inner solve, gicp.hpp:518-524) the residual r_k = A p~_k - q_k, p~ = (p, 1), is
// linear in the 12 entries of A = [R | t], so f = sum_k r_k' M_k r_k is an exact quadratic form in A.  Expanded
// about A0 (the transform the correspondences were taken with): r_k = r0_k + D p~_k, D = A - A0, r0_k = A0 p~_k - q_k:
//     f(A)           = c0 + sum_ia D[i][a] (2 B[a][i] + W[i][a])
//     sum_k M_k r_k p~_k'   = B' + W                                       (3x4; what the 13 sums of gicp.hpp:373-397 hold)
//     W[i][a]        = sum_jb S[ab][ij] D[j][b]
// with the MOM_N = 74 moments, reduced over the correspondences ONCE per outer iteration:
//     S[ab][ij] = sum_k (p~ p~')_ab (M_k)_ij    10 x 6   (Q = sum p~p~' (x) M has only 60 distinct entries)
//     B[a][i]   = sum_k p~_a (M_k r0_k)_i       4 x 3
//     c0        = sum_k r0_k' M_k r0_k,   count = number of correspondences.
// Every objective / gradient evaluation of the BFGS line search (gicp.hpp:290-402) is then O(1) scalar work: no pass
// over the points and no grid-wide reduction.  Expanding about A0 keeps every term at its natural size (c0 is the
// objective at A0 itself, D ~ 1e-2): no cancellation between |p|^2-sized terms.  The only arithmetic difference to
// the reference: A p~ is evaluated in double here, in float32 there (gicp.hpp:307,341,382: ~1e-6 m per point).
constexpr int MOM_S = 0, MOM_B = 60, MOM_C = 72, MOM_CNT = 73, MOM_N = 74;
// index of the pair (a <= b) of p~ components: xx xy xz x1 yy yz y1 zz z1 11
LB_HD int mom_pair(int a, int b) {
  const int base[4] = {0, 4, 7, 9};
  return a <= b ? base[a] + (b - a) : base[b] + (a - b);
}
LB_HD int sym6(int i, int j) {
  const int base[3] = {0, 3, 5};
  return i <= j ? base[i] + (j - i) : base[j] + (i - j);
}

// contribution of one correspondence.  A0: row-major 3x4 float; p: source point; q: matched target point.
LB_HD void moment_terms(const float* A0, float px, float py, float pz, float qx, float qy, float qz, const double* M,
                        double* acc /*MOM_N*/) {
  const double p[4] = {(double)px, (double)py, (double)pz, 1.0};
  double r0[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    r0[i] = (((double)A0[4 * i] * p[0] + (double)A0[4 * i + 1] * p[1]) + (double)A0[4 * i + 2] * p[2]) + (double)A0[4 * i + 3];
  r0[0] -= (double)qx; r0[1] -= (double)qy; r0[2] -= (double)qz;
  const double t0 = (M[SXX] * r0[0] + M[SXY] * r0[1]) + M[SXZ] * r0[2];
  const double t1 = (M[SXY] * r0[0] + M[SYY] * r0[1]) + M[SYZ] * r0[2];
  const double t2 = (M[SXZ] * r0[0] + M[SYZ] * r0[1]) + M[SZZ] * r0[2];
  acc[MOM_C] += (r0[0] * t0 + r0[1] * t1) + r0[2] * t2;
  acc[MOM_CNT] += 1.0;
  int e = 0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    acc[MOM_B + 3 * a + 0] += p[a] * t0;
    acc[MOM_B + 3 * a + 1] += p[a] * t1;
    acc[MOM_B + 3 * a + 2] += p[a] * t2;
#pragma unroll
    for (int b = a; b < 4; b++) {
      const double pab = p[a] * p[b];
#pragma unroll
      for (int ij = 0; ij < 6; ij++) acc[MOM_S + 6 * e + ij] += pab * M[ij];
      e++;
    }
  }
}

// W = Q D for a 3x4 matrix D (row-major), from the S moments.
LB_HD void moment_apply(const double* mom, const double* D /*12*/, double* W /*12*/) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      double w = 0.0;
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int j = 0; j < 3; j++) w += mom[MOM_S + 6 * mom_pair(a, b) + sym6(i, j)] * D[4 * j + b];
      W[4 * i + a] = w;
    }
}

// The 13 sums of objective_terms (f, sum M r, sum p (M r)') at transform A, from the moments taken at A0.
LB_HD void moment_sums13(const double* mom, const float* A0, const float* A, double* sums /*13*/) {
  double D[12], W[12];
#pragma unroll
  for (int e = 0; e < 12; e++) D[e] = (double)A[e] - (double)A0[e];
  moment_apply(mom, D, W);
  double f = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double bia = mom[MOM_B + 3 * a + i];
      f += D[4 * i + a] * (2.0 * bia + W[4 * i + a]);
      const double g = bia + W[4 * i + a];          // (sum_k M_k r_k p~_k')[i][a]
      if (a == 3) sums[1 + i] = g;
      else sums[4 + 3 * a + i] = g;
    }
  sums[0] = mom[MOM_C] + f;
}

// Gauss-Newton normal equations at state x from the moments: J_k = [dA/dx_c p~_k]_c, H = sum J'MJ, b = sum J'Mr.
// out: f (sum, not divided), b[6], H[21] upper-triangular row-major -- the layout of gn_terms' 28 sums.
LB_HD void moment_gn28(const double* mom, const float* A0, const float* A, const double* dP, const double* dT,
                       const double* dS, double* sums /*28*/) {
  double s13[13];
  moment_sums13(mom, A0, A, s13);
  double G[12];                                          // sum_k M_k r_k p~_k' (3x4)
#pragma unroll
  for (int i = 0; i < 3; i++) {
    G[4 * i + 3] = s13[1 + i];
#pragma unroll
    for (int a = 0; a < 3; a++) G[4 * i + a] = s13[4 + 3 * a + i];
  }
  double Dc[6][12], QD[6][12];
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int e = 0; e < 12; e++) Dc[c][e] = 0.0;
  Dc[0][3] = 1.0; Dc[1][7] = 1.0; Dc[2][11] = 1.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int a = 0; a < 3; a++) { Dc[3][4 * i + a] = dP[3 * i + a]; Dc[4][4 * i + a] = dT[3 * i + a]; Dc[5][4 * i + a] = dS[3 * i + a]; }
#pragma unroll
  for (int c = 0; c < 6; c++) moment_apply(mom, Dc[c], QD[c]);
  sums[0] = s13[0];
  int e = 7;
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double bc = 0.0;
#pragma unroll
    for (int k = 0; k < 12; k++) bc += Dc[c][k] * G[k];
    sums[1 + c] = bc;
#pragma unroll
    for (int d = c; d < 6; d++) {
      double h = 0.0;
#pragma unroll
      for (int k = 0; k < 12; k++) h += Dc[c][k] * QD[d][k];
      sums[e++] = h;
    }
  }
}

// Gauss-Newton terms (not in the reference): J = [I | dP p, dT p, dS p],
// H += J'MJ (21 upper-tri), b += J'Mr (6), f += r'Mr.  acc: 28 doubles [f, b0..5, H00,H01,..H55 upper].
LB_HD void gn_terms(const float* T, const double* dP, const double* dT, const double* dS,
                    float px, float py, float pz, float qx, float qy, float qz,
                    const double* M, double* acc /*28*/) {
  float tx, ty, tz;
  xform(T, px, py, pz, tx, ty, tz);
  double r[3] = {(double)(tx - qx), (double)(ty - qy), (double)(tz - qz)};
  double p[3] = {(double)px, (double)py, (double)pz};
  double J[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    J[i][0] = (i == 0); J[i][1] = (i == 1); J[i][2] = (i == 2);
    J[i][3] = dP[i * 3 + 0] * p[0] + dP[i * 3 + 1] * p[1] + dP[i * 3 + 2] * p[2];
    J[i][4] = dT[i * 3 + 0] * p[0] + dT[i * 3 + 1] * p[1] + dT[i * 3 + 2] * p[2];
    J[i][5] = dS[i * 3 + 0] * p[0] + dS[i * 3 + 1] * p[1] + dS[i * 3 + 2] * p[2];
  }
  double Mm[3][3] = {{M[SXX], M[SXY], M[SXZ]}, {M[SXY], M[SYY], M[SYZ]}, {M[SXZ], M[SYZ], M[SZZ]}};
  double Mr[3], MJ[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Mr[i] = Mm[i][0] * r[0] + Mm[i][1] * r[1] + Mm[i][2] * r[2];
#pragma unroll
    for (int a = 0; a < 6; a++) MJ[i][a] = Mm[i][0] * J[0][a] + Mm[i][1] * J[1][a] + Mm[i][2] * J[2][a];
  }
  acc[0] += r[0] * Mr[0] + r[1] * Mr[1] + r[2] * Mr[2];
  int e = 7;
#pragma unroll
  for (int a = 0; a < 6; a++) {
    acc[1 + a] += J[0][a] * Mr[0] + J[1][a] * Mr[1] + J[2][a] * Mr[2];
#pragma unroll
    for (int b = a; b < 6; b++) {
      acc[e] += J[0][a] * MJ[0][b] + J[1][a] * MJ[1][b] + J[2][a] * MJ[2][b];
      e++;
    }
  }
}

// Solve the 6x6 SPD system H d = -b (H given as 21 upper-tri row-major).  Gaussian
// elimination with partial pivoting in double; returns 0 on success.
LB_HD int solve6_neg(const double* Hu, const double* b, double* d) {
  double A[6][7];
  int e = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) { A[i][j] = Hu[e]; A[j][i] = Hu[e]; e++; }
  for (int i = 0; i < 6; i++) A[i][6] = -b[i];
  for (int c = 0; c < 6; c++) {
    int piv = c;
    for (int r = c + 1; r < 6; r++) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (fabs(A[piv][c]) < 1e-300) return -1;
    if (piv != c) for (int j = 0; j < 7; j++) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
    for (int r = c + 1; r < 6; r++) {
      double f = A[r][c] / A[c][c];
      for (int j = c; j < 7; j++) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = 5; i >= 0; i--) {
    double s = A[i][6];
    for (int j = i + 1; j < 6; j++) s -= A[i][j] * d[j];
    d[i] = s / A[i][i];
  }
  return 0;
}

}  // namespace lb
