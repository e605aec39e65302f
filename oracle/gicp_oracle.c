/*
 * gicp_oracle.c -- CPU restatement of the reference GICP (TEST INFRASTRUCTURE ONLY).
 *
 * Follows multithreaded_gicp/include/multithreaded_gicp/gicp.hpp line by line:
 *   computeCovariances            gicp.hpp:64-156
 *   computeRDerivative            gicp.hpp:159-214
 *   estimateRigidTransformationBFGS gicp.hpp:217-287
 *   OptimizationFunctorWithIndices gicp.hpp:290-402
 *   computeTransformation         gicp.hpp:405-617
 *   applyState                    gicp.hpp:619-634
 * plus the pcl::Registration::align() behaviour around it.
 *
 * Mixed precision is mirrored: points float32, T*p evaluated in float32
 * (Matrix4f * Vector4f: ((c0*x + c1*y) + c2*z) + c3*w, no FMA), residual /
 * Mahalanobis / sums in double, applyState builds R in float32 from
 * AngleAxisf products (quaternion path of Eigen).
 * Compile with -ffp-contract=off.
 */
#include "lb_oracle.h"

#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void og_gicp_default_params(og_gicp_params* p) {
  p->k_correspondences = 20;        /* gicp.h:112 */
  p->gicp_epsilon = 0.001;          /* gicp.h:118 */
  p->rotation_epsilon = 2e-3;       /* gicp.h:119 */
  p->transformation_epsilon = 5e-4; /* gicp.h:126 */
  p->corr_dist_threshold = 5.0;     /* gicp.h:127 */
  p->max_iterations = 200;          /* gicp.h:125 */
  p->max_inner_iterations = 20;     /* gicp.h:121 */
  p->num_threads = 1;               /* gicp.h:117 */
  p->source_cov_from_normals = 0;
  p->target_cov_from_normals = 0;
  p->optimizer = 0;
}

/* ------------------------------------------------------------ small linear algebra */

/* Symmetric 3x3 eigen decomposition by cyclic Jacobi (double).  A is
 * overwritten; d = eigenvalues, V columns = eigenvectors.  Stands in for
 * Eigen::JacobiSVD<Matrix3d> at gicp.hpp:140: for a symmetric PSD matrix the
 * left singular vectors are the eigenvectors and the singular values |lambda|. */
static void sym3_jacobi(double A[3][3], double d[3], double V[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
    for (int e = 0; e < 3; e++) {
      int p = PQ[e][0], q = PQ[e][1];
      double apq = A[p][q];
      if (apq == 0.0) continue;
      double g = 100.0 * fabs(apq);
      /* negligible off-diagonal element (Numerical Recipes criterion) */
      if (sweep > 3 && fabs(A[p][p]) + g == fabs(A[p][p]) && fabs(A[q][q]) + g == fabs(A[q][q])) {
        A[p][q] = A[q][p] = 0.0;
        continue;
      }
      double h = A[q][q] - A[p][p];
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
      A[p][p] -= hh;
      A[q][q] += hh;
      A[p][q] = A[q][p] = 0.0;
      int r = 3 - p - q; /* the remaining index */
      double arp = A[r][p], arq = A[r][q];
      A[r][p] = A[p][r] = arp - s * (arq + arp * tau);
      A[r][q] = A[q][r] = arq + s * (arp - arq * tau);
      for (int k = 0; k < 3; k++) {
        double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = vkp - s * (vkq + vkp * tau);
        V[k][q] = vkq + s * (vkp - vkq * tau);
      }
    }
  }
  d[0] = A[0][0]; d[1] = A[1][1]; d[2] = A[2][2];
}

/* the same decomposition for ndt_oracle.c (SelfAdjointEigenSolver<Matrix3d>) */
void og_sym3_jacobi(double A[3][3], double d[3], double V[3][3]) { sym3_jacobi(A, d, V); }

/* gicp.hpp:139-153: U from SVD, singular values descending; rebuild with (1,1,eps) */
static void regularise_cov(double cov[3][3], double eps, double out[9]) {
  double d[3], V[3][3];
  sym3_jacobi(cov, d, V);
  int ord[3] = {0, 1, 2};
  /* stable sort by |d| descending */
  for (int i = 1; i < 3; i++) {
    int o = ord[i]; int j = i;
    while (j > 0 && fabs(d[ord[j - 1]]) < fabs(d[o])) { ord[j] = ord[j - 1]; j--; }
    ord[j] = o;
  }
  for (int i = 0; i < 9; i++) out[i] = 0.0;
  for (int k = 0; k < 3; k++) {
    double v = (k == 2) ? eps : 1.0;
    int c = ord[k];
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) out[r * 3 + cc] += (v * V[r][c]) * V[cc][c];
  }
}

/* Eigen Matrix3d::inverse() (cofactor form) */
static void inv3(const double m[9], double out[9]) {
#define MM(r, c) m[(r) * 3 + (c)]
  /* cofactor(i,j) = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1), i1=(i+1)%3, i2=(i+2)%3 (same for j);
   * Eigen: det = cofactors_col0 . col(0), result(j,i) = cofactor(i,j) / det */
  double cof[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      cof[i][j] = MM(i1, j1) * MM(i2, j2) - MM(i1, j2) * MM(i2, j1);
    }
  double det = (cof[0][0] * MM(0, 0) + cof[1][0] * MM(1, 0)) + cof[2][0] * MM(2, 0);
  double invdet = 1.0 / det;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[j * 3 + i] = cof[i][j] * invdet;
#undef MM
}

static void mat3_mul(const double a[9], const double b[9], double o[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      o[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}

/* Matrix4f * (x,y,z,1) in float, Eigen coefficient order, no FMA */
static inline void xform4f(const float T[16], const float p[3], float out[3]) {
  for (int r = 0; r < 3; r++) {
    float v = T[r * 4 + 0] * p[0];
    v = v + T[r * 4 + 1] * p[1];
    v = v + T[r * 4 + 2] * p[2];
    v = v + T[r * 4 + 3] * 1.0f;
    out[r] = v;
  }
}

/* pcl::transformPointCloud (PCL 1.10 detail::Transformer<float>::se3):
 * p0 + (p1 + (p2 + c3)) with pk = ck * src[k] */
static inline void pcl_transform_pt(const float T[16], const float p[3], float out[3]) {
  for (int r = 0; r < 3; r++) {
    float p0 = T[r * 4 + 0] * p[0];
    float p1 = T[r * 4 + 1] * p[1];
    float p2 = T[r * 4 + 2] * p[2];
    float v = p2 + T[r * 4 + 3];
    v = p1 + v;
    v = p0 + v;
    out[r] = v;
  }
}

static void mat4f_mul(const float a[16], const float b[16], float o[16]) {
  /* Eigen Matrix4f * Matrix4f, coefficient order k = 0..3 */
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float v = a[i * 4 + 0] * b[0 * 4 + j];
      v = v + a[i * 4 + 1] * b[1 * 4 + j];
      v = v + a[i * 4 + 2] * b[2 * 4 + j];
      v = v + a[i * 4 + 3] * b[3 * 4 + j];
      o[i * 4 + j] = v;
    }
}

static void mat4f_identity(float T[16]) {
  memset(T, 0, sizeof(float) * 16);
  T[0] = T[5] = T[10] = T[15] = 1.0f;
}

/* gicp.hpp:619-634 applied on identity.  Eigen: AngleAxisf*AngleAxisf*AngleAxisf
 * -> float quaternion product -> toRotationMatrix(). */
void og_gicp_apply_state(const double x[6], float T[16]) {
  float az = (float)x[5], ay = (float)x[4], ax = (float)x[3];
  float hz = 0.5f * az, hy = 0.5f * ay, hx = 0.5f * ax;
  float cz = cosf(hz), sz = sinf(hz);
  float cy = cosf(hy), sy = sinf(hy);
  float cx = cosf(hx), sx = sinf(hx);
  /* q1 = qz * qy   (qz = (w=cz, 0,0,sz), qy = (w=cy, 0,sy,0)) */
  float w1 = cz * cy, x1 = -(sz * sy), y1 = cz * sy, z1 = sz * cy;
  /* q = q1 * qx    (qx = (w=cx, sx,0,0)) */
  float qw = w1 * cx - x1 * sx;
  float qx = w1 * sx + x1 * cx;
  float qy = y1 * cx + z1 * sx;
  float qz = z1 * cx - y1 * sx;
  float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
  float twx = tx * qw, twy = ty * qw, twz = tz * qw;
  float txx = tx * qx, txy = ty * qx, txz = tz * qx;
  float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  mat4f_identity(T);
  T[0] = 1.0f - (tyy + tzz); T[1] = txy - twz;          T[2] = txz + twy;
  T[4] = txy + twz;          T[5] = 1.0f - (txx + tzz); T[6] = tyz - twx;
  T[8] = txz - twy;          T[9] = tyz + twx;          T[10] = 1.0f - (txx + tyy);
  T[3] = (float)x[0]; T[7] = (float)x[1]; T[11] = (float)x[2];
}

/* gicp.hpp:159-214 + matricesInnerProd gicp.h:361-370: g[3..5] = tr(dR' ... ) */
static void compute_r_derivative(const double x[6], const double R[9], double g[6]) {
  double dP[9], dT[9], dS[9];
  double phi = x[3], theta = x[4], psi = x[5];
  double cphi = cos(phi), sphi = sin(phi);
  double ctheta = cos(theta), stheta = sin(theta);
  double cpsi = cos(psi), spsi = sin(psi);
#define S(m, r, c) m[(r) * 3 + (c)]
  S(dP, 0, 0) = 0.; S(dP, 1, 0) = 0.; S(dP, 2, 0) = 0.;
  S(dP, 0, 1) = sphi * spsi + cphi * cpsi * stheta;
  S(dP, 1, 1) = -cpsi * sphi + cphi * spsi * stheta;
  S(dP, 2, 1) = cphi * ctheta;
  S(dP, 0, 2) = cphi * spsi - cpsi * sphi * stheta;
  S(dP, 1, 2) = -cphi * cpsi - sphi * spsi * stheta;
  S(dP, 2, 2) = -ctheta * sphi;

  S(dT, 0, 0) = -cpsi * stheta; S(dT, 1, 0) = -spsi * stheta; S(dT, 2, 0) = -ctheta;
  S(dT, 0, 1) = cpsi * ctheta * sphi; S(dT, 1, 1) = ctheta * sphi * spsi; S(dT, 2, 1) = -sphi * stheta;
  S(dT, 0, 2) = cphi * cpsi * ctheta; S(dT, 1, 2) = cphi * ctheta * spsi; S(dT, 2, 2) = -cphi * stheta;

  S(dS, 0, 0) = -ctheta * spsi; S(dS, 1, 0) = cpsi * ctheta; S(dS, 2, 0) = 0.;
  S(dS, 0, 1) = -cphi * cpsi - sphi * spsi * stheta;
  S(dS, 1, 1) = -cphi * spsi + cpsi * sphi * stheta;
  S(dS, 2, 1) = 0.;
  S(dS, 0, 2) = cpsi * sphi - cphi * spsi * stheta;
  S(dS, 1, 2) = sphi * spsi + cphi * cpsi * stheta;
  S(dS, 2, 2) = 0.;
  const double* D[3] = {dP, dT, dS};
  for (int a = 0; a < 3; a++) {
    double r = 0.;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r += S(D[a], j, i) * S(R, i, j);
    g[3 + a] = r;
  }
#undef S
}

/* d(R)/d(phi,theta,psi) matrices, needed by the Gauss-Newton variant */
static void r_derivative_mats(const double x[6], double dP[9], double dT[9], double dS[9]) {
  /* tr(dA * E_ij-ish) trick: recover each matrix by probing compute_r_derivative
   * would be wasteful; restate the closed forms instead (same as above). */
  double phi = x[3], theta = x[4], psi = x[5];
  double cphi = cos(phi), sphi = sin(phi);
  double ctheta = cos(theta), stheta = sin(theta);
  double cpsi = cos(psi), spsi = sin(psi);
#define S(m, r, c) m[(r) * 3 + (c)]
  S(dP, 0, 0) = 0.; S(dP, 1, 0) = 0.; S(dP, 2, 0) = 0.;
  S(dP, 0, 1) = sphi * spsi + cphi * cpsi * stheta;
  S(dP, 1, 1) = -cpsi * sphi + cphi * spsi * stheta;
  S(dP, 2, 1) = cphi * ctheta;
  S(dP, 0, 2) = cphi * spsi - cpsi * sphi * stheta;
  S(dP, 1, 2) = -cphi * cpsi - sphi * spsi * stheta;
  S(dP, 2, 2) = -ctheta * sphi;
  S(dT, 0, 0) = -cpsi * stheta; S(dT, 1, 0) = -spsi * stheta; S(dT, 2, 0) = -ctheta;
  S(dT, 0, 1) = cpsi * ctheta * sphi; S(dT, 1, 1) = ctheta * sphi * spsi; S(dT, 2, 1) = -sphi * stheta;
  S(dT, 0, 2) = cphi * cpsi * ctheta; S(dT, 1, 2) = cphi * ctheta * spsi; S(dT, 2, 2) = -cphi * stheta;
  S(dS, 0, 0) = -ctheta * spsi; S(dS, 1, 0) = cpsi * ctheta; S(dS, 2, 0) = 0.;
  S(dS, 0, 1) = -cphi * cpsi - sphi * spsi * stheta;
  S(dS, 1, 1) = -cphi * spsi + cpsi * sphi * stheta;
  S(dS, 2, 1) = 0.;
  S(dS, 0, 2) = cpsi * sphi - cphi * spsi * stheta;
  S(dS, 1, 2) = sphi * spsi + cphi * cpsi * stheta;
  S(dS, 2, 2) = 0.;
#undef S
}

/* ------------------------------------------------------------ objective (gicp.hpp:290-402) */
typedef struct {
  const float* src4;   /* m x 4, compacted correspondences: output[src_idx] */
  const float* tgt4;   /* m x 4: target[tgt_idx] */
  const double* M;     /* m x 9 */
  int m;
  long n_evals;
} functor_ctx;

/* Summation-order probe (tests / studies only; 0 = the reference's order, one serial loop, gicp.hpp:298,331,373).
 * With chunk = c > 0 the 13 sums are formed as partial sums over blocks of c consecutive correspondences which are
 * then added in block order: the SAME terms, the SAME arithmetic, only the association of the additions differs (what
 * any vectorised or parallel build of the reference does).  It exists to measure how far the reference's own result
 * moves under such a reassociation: where its BFGS line search stalls on the float32 noise floor of the objective,
 * the last bits of f decide the branch and the final pose moves by up to millimetres (DESIGN.md "Numerics"). */
static int g_sum_chunk = 0;
void og_set_sum_chunk(int c) { g_sum_chunk = c > 0 ? c : 0; }

static void functor_core(const functor_ctx* c, const double* x, double* f_out, double* g_out) {
  float T[16];
  og_gicp_apply_state(x, T); /* base_transformation_ = I (gicp.hpp:435) */
  double f = 0;
  double gt[3] = {0, 0, 0};
  double R[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int m = c->m;
  if (g_sum_chunk > 0) {
    for (int b = 0; b < m; b += g_sum_chunk) {
      double fb = 0, gb[3] = {0, 0, 0}, Rb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const int e = b + g_sum_chunk < m ? b + g_sum_chunk : m;
      for (int i = b; i < e; i++) {
        const float* ps = &c->src4[4 * (size_t)i];
        const float* pt = &c->tgt4[4 * (size_t)i];
        const double* M = &c->M[9 * (size_t)i];
        float pp[3];
        xform4f(T, ps, pp);
        double res[3] = {(double)(pp[0] - pt[0]), (double)(pp[1] - pt[1]), (double)(pp[2] - pt[2])};
        double temp[3];
        for (int r = 0; r < 3; r++) temp[r] = (M[r * 3 + 0] * res[0] + M[r * 3 + 1] * res[1]) + M[r * 3 + 2] * res[2];
        fb += (res[0] * temp[0] + res[1] * temp[1]) + res[2] * temp[2];
        gb[0] += temp[0]; gb[1] += temp[1]; gb[2] += temp[2];
        for (int r = 0; r < 3; r++)
          for (int cc = 0; cc < 3; cc++) Rb[r * 3 + cc] += (double)ps[r] * temp[cc];
      }
      f += fb; gt[0] += gb[0]; gt[1] += gb[1]; gt[2] += gb[2];
      for (int k = 0; k < 9; k++) R[k] += Rb[k];
    }
  } else
  for (int i = 0; i < m; i++) {
    const float* ps = &c->src4[4 * (size_t)i];
    const float* pt = &c->tgt4[4 * (size_t)i];
    const double* M = &c->M[9 * (size_t)i];
    float pp[3];
    xform4f(T, ps, pp);
    double res[3] = {(double)(pp[0] - pt[0]), (double)(pp[1] - pt[1]), (double)(pp[2] - pt[2])};
    double temp[3];
    for (int r = 0; r < 3; r++) temp[r] = (M[r * 3 + 0] * res[0] + M[r * 3 + 1] * res[1]) + M[r * 3 + 2] * res[2];
    f += (res[0] * temp[0] + res[1] * temp[1]) + res[2] * temp[2];
    if (g_out) {
      gt[0] += temp[0]; gt[1] += temp[1]; gt[2] += temp[2];
      /* pp = base_transformation_ * p_src = p_src (gicp.hpp:351,393) */
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) R[r * 3 + cc] += (double)ps[r] * temp[cc];
    }
  }
  if (f_out) *f_out = f / (double)m;
  if (g_out) {
    double sc = 2.0 / m;
    g_out[0] = gt[0] * sc; g_out[1] = gt[1] * sc; g_out[2] = gt[2] * sc;
    for (int i = 0; i < 9; i++) R[i] *= sc;
    compute_r_derivative(x, R, g_out);
  }
}
static double functor_f(void* c, const double* x) {
  double f; ((functor_ctx*)c)->n_evals++; functor_core((functor_ctx*)c, x, &f, NULL); return f;
}
static void functor_df(void* c, const double* x, double* g) {
  ((functor_ctx*)c)->n_evals++; functor_core((functor_ctx*)c, x, NULL, g);
}
static void functor_fdf(void* c, const double* x, double* f, double* g) {
  ((functor_ctx*)c)->n_evals++; functor_core((functor_ctx*)c, x, f, g);
}

void og_gicp_fdf(const float* src4, const float* tgt4, const double* M, int m,
                 const double x[6], double* f, double g[6]) {
  functor_ctx c = {src4, tgt4, M, m, 0};
  functor_core(&c, x, f, g);
}

/* ------------------------------------------------------------ covariances */
static int cov_knn(const float* pts, int n, int stride, const og_kdtree* tree, int k,
                   double eps, int num_threads, double* cov_out) {
  if (k > n) return -1; /* gicp.hpp:72-79: error, return (covariances stay unset) */
#pragma omp parallel for schedule(dynamic, 16) num_threads(num_threads)
  for (int i = 0; i < n; i++) {
    int nn_idx[64];
    float nn_d2[64];
    const float* q = &pts[(size_t)i * stride];
    int found = og_kdtree_knn(tree, q, k, nn_idx, nn_d2);
    double mean[3] = {0, 0, 0};
    double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < found; j++) {
      const float* p = &pts[(size_t)nn_idx[j] * stride];
      /* pt.x * pt.x is a float product (gicp.hpp:119-126) accumulated into double */
      mean[0] += p[0]; mean[1] += p[1]; mean[2] += p[2];
      cov[0][0] += p[0] * p[0];
      cov[1][0] += p[1] * p[0];
      cov[1][1] += p[1] * p[1];
      cov[2][0] += p[2] * p[0];
      cov[2][1] += p[2] * p[1];
      cov[2][2] += p[2] * p[2];
    }
    mean[0] /= (double)k; mean[1] /= (double)k; mean[2] /= (double)k;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b <= a; b++) {
        cov[a][b] /= (double)k;
        cov[a][b] -= mean[a] * mean[b];
        cov[b][a] = cov[a][b];
      }
    regularise_cov(cov, eps, &cov_out[9 * (size_t)i]);
  }
  return 0;
}

/* gicp.hpp:81-82: CalculateCovarianceFromNormals lives in the un-vendored
 * frontend_utils package.  PARITY UNPINNED.  The only basis-independent
 * definition consistent with the k-NN branch is C = I - (1-eps) n n'. */
static void cov_from_normals(const float* pts, int n, int stride, int noff, double eps, double* cov_out) {
  for (int i = 0; i < n; i++) {
    double nx = pts[(size_t)i * stride + noff], ny = pts[(size_t)i * stride + noff + 1],
           nz = pts[(size_t)i * stride + noff + 2];
    double nv[3] = {nx, ny, nz};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        cov_out[9 * (size_t)i + r * 3 + c] = ((r == c) ? 1.0 : 0.0) - (1.0 - eps) * nv[r] * nv[c];
  }
}

int og_gicp_covariances(const float* pts, int n, int stride_f, int k, double gicp_epsilon,
                        int num_threads, double* cov_out) {
  og_kdtree* t = og_kdtree_build(pts, n, stride_f);
  int rc = cov_knn(pts, n, stride_f, t, k, gicp_epsilon, num_threads, cov_out);
  og_kdtree_free(t);
  return rc;
}

/* ------------------------------------------------------------ 6x6 solve for GN */
static int solve6(double H[36], double b[6], double x[6]) {
  /* Gaussian elimination with partial pivoting */
  double A[6][7];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) A[i][j] = H[i * 6 + j]; A[i][6] = b[i]; }
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
    for (int j = i + 1; j < 6; j++) s -= A[i][j] * x[j];
    x[i] = s / A[i][i];
  }
  return 0;
}

/* Gauss-Newton inner solve (NOT in the reference). */
static int gn_solve(functor_ctx* c, double x[6], int max_inner, long* n_inner) {
  for (int it = 0; it < max_inner; it++) {
    float T[16];
    og_gicp_apply_state(x, T);
    double dP[9], dT[9], dS[9];
    r_derivative_mats(x, dP, dT, dS);
    double H[36], b[6];
    memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
    for (int i = 0; i < c->m; i++) {
      const float* ps = &c->src4[4 * (size_t)i];
      const float* pt = &c->tgt4[4 * (size_t)i];
      const double* M = &c->M[9 * (size_t)i];
      float pp[3];
      xform4f(T, ps, pp);
      double res[3] = {(double)(pp[0] - pt[0]), (double)(pp[1] - pt[1]), (double)(pp[2] - pt[2])};
      double J[3][6];
      double p[3] = {ps[0], ps[1], ps[2]};
      for (int r = 0; r < 3; r++) {
        J[r][0] = (r == 0); J[r][1] = (r == 1); J[r][2] = (r == 2);
        J[r][3] = dP[r * 3 + 0] * p[0] + dP[r * 3 + 1] * p[1] + dP[r * 3 + 2] * p[2];
        J[r][4] = dT[r * 3 + 0] * p[0] + dT[r * 3 + 1] * p[1] + dT[r * 3 + 2] * p[2];
        J[r][5] = dS[r * 3 + 0] * p[0] + dS[r * 3 + 1] * p[1] + dS[r * 3 + 2] * p[2];
      }
      double MJ[3][6], Mr[3];
      for (int r = 0; r < 3; r++) {
        Mr[r] = M[r * 3 + 0] * res[0] + M[r * 3 + 1] * res[1] + M[r * 3 + 2] * res[2];
        for (int a = 0; a < 6; a++)
          MJ[r][a] = M[r * 3 + 0] * J[0][a] + M[r * 3 + 1] * J[1][a] + M[r * 3 + 2] * J[2][a];
      }
      for (int a = 0; a < 6; a++) {
        b[a] += J[0][a] * Mr[0] + J[1][a] * Mr[1] + J[2][a] * Mr[2];
        for (int bb = 0; bb < 6; bb++)
          H[a * 6 + bb] += J[0][a] * MJ[0][bb] + J[1][a] * MJ[1][bb] + J[2][a] * MJ[2][bb];
      }
    }
    c->n_evals++;
    (*n_inner)++;
    double nb[6], dx[6];
    for (int a = 0; a < 6; a++) nb[a] = -b[a];
    if (solve6(H, nb, dx) != 0) return -1;
    double mx = 0;
    for (int a = 0; a < 6; a++) { x[a] += dx[a]; if (fabs(dx[a]) > mx) mx = fabs(dx[a]); }
    if (mx < 1e-6) break; /* below the float32 resolution of T(x) */
  }
  return 0;
}

/* ------------------------------------------------------------ align */
static double now_s(void) { return omp_get_wtime(); }

/* A target whose kd-tree and covariances are kept between aligns.  The reference keeps both for as long as the
 * caller does not call setInputTarget again (pcl::Registration rebuilds the tree only when the target changed, and
 * gicp.hpp:422-426 computes target covariances only "if unset"): that is the scan-to-submap case in which the
 * submap is unchanged between scans (BASELINE configs[2]). */
struct og_gicp_target {
  const float* pts; int n, stride_f;
  og_kdtree* tree;
  double* cov;       /* n x 9 */
};

static int align_impl(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                      const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                      const og_gicp_params* P, const float* guess_in, og_gicp_result* res,
                      double* src_cov_out, double* tgt_cov_out, float* aligned_out, const og_gicp_target* prepared);

og_gicp_target* og_gicp_target_prepare(const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                                       const og_gicp_params* P) {
  if (n_tgt <= 0 || P->k_correspondences > n_tgt) return NULL;
  og_gicp_target* t = (og_gicp_target*)calloc(1, sizeof(*t));
  t->pts = tgt; t->n = n_tgt; t->stride_f = tgt_stride_f;
  t->tree = og_kdtree_build(tgt, n_tgt, tgt_stride_f);
  t->cov = (double*)malloc(sizeof(double) * 9 * (size_t)n_tgt);
  if (P->target_cov_from_normals && tgt_normal_off_f >= 0)
    cov_from_normals(tgt, n_tgt, tgt_stride_f, tgt_normal_off_f, P->gicp_epsilon, t->cov);
  else
    cov_knn(tgt, n_tgt, tgt_stride_f, t->tree, P->k_correspondences, P->gicp_epsilon,
            P->num_threads > 0 ? P->num_threads : 1, t->cov);
  return t;
}

/* covariances supplied by the caller (n x 9 doubles) instead of computed from the target as it is now: the rolling
 * submap caches each point's covariance from the time of its insertion (oracle/submap_oracle.py) */
void og_gicp_target_set_covariances(og_gicp_target* t, const double* cov9) {
  if (t && cov9) memcpy(t->cov, cov9, sizeof(double) * 9 * (size_t)t->n);
}

void og_gicp_target_free(og_gicp_target* t) {
  if (!t) return;
  og_kdtree_free(t->tree);
  free(t->cov);
  free(t);
}

int og_gicp_align_prepared(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                           const og_gicp_target* target, const og_gicp_params* P, const float* guess,
                           og_gicp_result* res) {
  if (!target) { memset(res, 0, sizeof(*res)); res->status = -2; return -2; }
  return align_impl(src, n_src, src_stride_f, src_normal_off_f, target->pts, target->n, target->stride_f, -1, P, guess,
                    res, NULL, NULL, NULL, target);
}

int og_gicp_align(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                  const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                  const og_gicp_params* P, const float* guess_in, og_gicp_result* res,
                  double* src_cov_out, double* tgt_cov_out, float* aligned_out) {
  return align_impl(src, n_src, src_stride_f, src_normal_off_f, tgt, n_tgt, tgt_stride_f, tgt_normal_off_f, P, guess_in,
                    res, src_cov_out, tgt_cov_out, aligned_out, NULL);
}

static int align_impl(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                      const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                      const og_gicp_params* P, const float* guess_in, og_gicp_result* res,
                      double* src_cov_out, double* tgt_cov_out, float* aligned_out, const og_gicp_target* prepared) {
  memset(res, 0, sizeof(*res));
  mat4f_identity(res->final_transformation);
  if (n_src <= 0) { res->status = -1; return -1; } /* gicp.h:164-171: empty source -> error, no-op */
  if (n_tgt <= 0) { res->status = -2; return -2; }
  const int nt = P->num_threads > 0 ? P->num_threads : 1;
  double t_start = now_s();

  float guess[16];
  if (guess_in) memcpy(guess, guess_in, sizeof(guess)); else mat4f_identity(guess);

  /* align(): rebuild target tree; initComputeReciprocal(): source tree (gicp.hpp:412) */
  og_kdtree* tree = prepared ? prepared->tree : og_kdtree_build(tgt, n_tgt, tgt_stride_f);
  og_kdtree* tree_src = og_kdtree_build(src, n_src, src_stride_f);

  const size_t N = (size_t)n_src;
  double* mahal = (double*)malloc(sizeof(double) * 9 * N);       /* gicp.hpp:418 */
  double* Csrc = (double*)malloc(sizeof(double) * 9 * N);
  double* Ctgt = prepared ? prepared->cov : (double*)malloc(sizeof(double) * 9 * (size_t)n_tgt);
  float* output = (float*)malloc(sizeof(float) * 4 * N);         /* copy of input, w=1 */
  int* source_indices = (int*)malloc(sizeof(int) * N);
  int* target_indices = (int*)malloc(sizeof(int) * N);
  float* csrc4 = (float*)malloc(sizeof(float) * 4 * N);
  float* ctgt4 = (float*)malloc(sizeof(float) * 4 * N);
  double* cM = (double*)malloc(sizeof(double) * 9 * N);
  int rc = 0;

  for (size_t i = 0; i < N; i++) {
    for (int d = 0; d < 9; d++) mahal[9 * i + d] = (d % 4 == 0) ? 1.0 : 0.0;
    output[4 * i + 0] = src[i * src_stride_f + 0];
    output[4 * i + 1] = src[i * src_stride_f + 1];
    output[4 * i + 2] = src[i * src_stride_f + 2];
    output[4 * i + 3] = 1.0f;
  }

  /* covariances: target then source (gicp.hpp:420-432) */
  double t_cov0 = now_s();
  if (prepared) {
    /* target covariances already set: gicp.hpp:422 skips them */
  } else if (P->target_cov_from_normals && tgt_normal_off_f >= 0) {
    if (P->k_correspondences > n_tgt) rc = -3;
    else cov_from_normals(tgt, n_tgt, tgt_stride_f, tgt_normal_off_f, P->gicp_epsilon, Ctgt);
  } else {
    if (cov_knn(tgt, n_tgt, tgt_stride_f, tree, P->k_correspondences, P->gicp_epsilon, nt, Ctgt)) rc = -3;
  }
  if (!rc) {
    if (P->source_cov_from_normals && src_normal_off_f >= 0) {
      if (P->k_correspondences > n_src) rc = -3;
      else cov_from_normals(src, n_src, src_stride_f, src_normal_off_f, P->gicp_epsilon, Csrc);
    } else {
      if (cov_knn(src, n_src, src_stride_f, tree_src, P->k_correspondences, P->gicp_epsilon, nt, Csrc)) rc = -3;
    }
  }
  double t_cov1 = now_s();
  if (rc) { res->status = rc; goto done; }
  if (src_cov_out) memcpy(src_cov_out, Csrc, sizeof(double) * 9 * N);
  if (tgt_cov_out) memcpy(tgt_cov_out, Ctgt, sizeof(double) * 9 * (size_t)n_tgt);

  float transformation[16], previous[16];
  mat4f_identity(transformation); /* align() resets transformation_ = previous_ = final_ = I */
  mat4f_identity(previous);
  int nr_iterations = 0, converged = 0;
  const double dist_threshold = P->corr_dist_threshold * P->corr_dist_threshold;

  /* pcl::transformPointCloud(output, output, guess)  gicp.hpp:440 */
  for (size_t i = 0; i < N; i++) {
    float o[3];
    pcl_transform_pt(guess, &output[4 * i], o);
    output[4 * i + 0] = o[0]; output[4 * i + 1] = o[1]; output[4 * i + 2] = o[2];
  }

  double delta = 0.;
  long n_evals = 0, n_inner_total = 0;
  int m_last = 0;
  double t_it0 = now_s();
  while (!converged) {
    /* gicp.hpp:450-460: R = double(transformation_) * double(guess), 3x3 block */
    double tR[16];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double acc = 0.0;
        for (int k = 0; k < 4; k++) acc += (double)transformation[i * 4 + k] * (double)guess[k * 4 + j];
        tR[i * 4 + j] = acc;
      }
    double R[9], Rt[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { R[i * 3 + j] = tR[i * 4 + j]; Rt[j * 3 + i] = tR[i * 4 + j]; }

    double t_l0 = now_s();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (size_t i = 0; i < N; i++) {
      source_indices[i] = -1; target_indices[i] = -1;
      float q[3];
      xform4f(transformation, &output[4 * i], q); /* gicp.hpp:469 */
      int nn = -1; float d2 = 0.f;
      if (og_kdtree_knn(tree, q, 1, &nn, &d2) == 0) continue;
      if (d2 < dist_threshold) { /* float vs double compare, gicp.hpp:483 */
        const double* C1 = &Csrc[9 * i];
        const double* C2 = &Ctgt[9 * (size_t)nn];
        double RC1[9], temp[9];
        mat3_mul(R, C1, RC1);
        mat3_mul(RC1, Rt, temp);
        for (int d = 0; d < 9; d++) temp[d] += C2[d];
        inv3(temp, &mahal[9 * i]);
        source_indices[i] = (int)i;
        target_indices[i] = nn;
      }
    }
    res->t_lookups_s += now_s() - t_l0;

    /* compact (gicp.hpp:509-514) */
    int m = 0;
    for (size_t i = 0; i < N; i++) {
      if (source_indices[i] < 0) continue;
      memcpy(&csrc4[4 * (size_t)m], &output[4 * i], sizeof(float) * 4);
      const float* tp = &tgt[(size_t)target_indices[i] * tgt_stride_f];
      ctgt4[4 * (size_t)m + 0] = tp[0]; ctgt4[4 * (size_t)m + 1] = tp[1];
      ctgt4[4 * (size_t)m + 2] = tp[2]; ctgt4[4 * (size_t)m + 3] = 1.0f;
      memcpy(&cM[9 * (size_t)m], &mahal[9 * i], sizeof(double) * 9);
      m++;
    }
    m_last = m;
    memcpy(previous, transformation, sizeof(previous)); /* gicp.hpp:518 */

    double t_o0 = now_s();
    /* estimateRigidTransformationBFGS gicp.hpp:217-287 */
    if (m < 4) break; /* NotEnoughPointsException -> caught -> break (gicp.hpp:542-547) */
    {
      double x[6];
      x[0] = transformation[3]; x[1] = transformation[7]; x[2] = transformation[11];
      x[3] = atan2((double)transformation[9], (double)transformation[10]);
      x[4] = asin(-(double)transformation[8]);
      x[5] = atan2((double)transformation[4], (double)transformation[0]);
      functor_ctx ctx = {csrc4, ctgt4, cM, m, 0};
      if (P->optimizer == 1) {
        if (gn_solve(&ctx, x, P->max_inner_iterations, &n_inner_total) != 0) { n_evals += ctx.n_evals; break; }
      } else {
        og_functor fn = {functor_f, functor_df, functor_fdf, &ctx, 6};
        og_bfgs bf;
        og_bfgs_init_params(&bf, fn);
        const double gradient_tol = 1e-2;
        int inner = 0;
        int result = og_bfgs_minimize_init(&bf, x);
        result = OG_BFGS_RUNNING;
        do {
          inner++;
          result = og_bfgs_minimize_one_step(&bf, x);
          if (result) break;
          result = og_bfgs_test_gradient(&bf, gradient_tol);
        } while (result == OG_BFGS_RUNNING && inner < P->max_inner_iterations);
        n_inner_total += inner;
        if (!(result == OG_BFGS_NO_PROGRESS || result == OG_BFGS_SUCCESS || inner == P->max_inner_iterations)) {
          n_evals += ctx.n_evals;
          break; /* SolverDidntConvergeException */
        }
      }
      n_evals += ctx.n_evals;
      og_gicp_apply_state(x, transformation); /* setIdentity + applyState, gicp.hpp:277-278 */
    }
    /* delta gicp.hpp:526-541 */
    delta = 0.;
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) {
        double ratio = (k < 3 && l < 3) ? 1. / P->rotation_epsilon : 1. / P->transformation_epsilon;
        double c_delta = ratio * fabs((double)(previous[k * 4 + l] - transformation[k * 4 + l]));
        if (c_delta > delta) delta = c_delta;
      }
    res->t_optimization_s += now_s() - t_o0;

    nr_iterations++;
    if (nr_iterations >= P->max_iterations || delta < 1) { /* gicp.hpp:566-568 */
      converged = 1;
      memcpy(previous, transformation, sizeof(previous));
    }
  }
  double t_it1 = now_s();

  /* final_transformation_ = previous_transformation_ * guess  gicp.hpp:583 */
  mat4f_mul(previous, guess, res->final_transformation);
  if (aligned_out) {
    for (size_t i = 0; i < N; i++) {
      float p[3] = {src[i * src_stride_f], src[i * src_stride_f + 1], src[i * src_stride_f + 2]};
      pcl_transform_pt(res->final_transformation, p, &aligned_out[3 * i]);
    }
  }
  res->nr_iterations = nr_iterations;
  res->converged = converged;
  res->n_correspondences = m_last;
  res->delta = delta;
  res->n_fdf_evals = n_evals;
  res->n_inner_iterations = n_inner_total;
  res->t_covariances_s = t_cov1 - t_cov0;
  res->t_iterations_s = t_it1 - t_it0;
  res->t_total_s = now_s() - t_start;
  res->status = 0;

done:
  if (!prepared) { og_kdtree_free(tree); free(Ctgt); }
  og_kdtree_free(tree_src);
  free(mahal); free(Csrc); free(output); free(source_indices); free(target_indices);
  free(csrc4); free(ctgt4); free(cM);
  return res->status;
}

double og_gicp_fitness(const float* src, int n_src, int src_stride_f,
                       const float* tgt, int n_tgt, int tgt_stride_f,
                       const float* T, double max_range, int num_threads) {
  og_kdtree* tree = og_kdtree_build(tgt, n_tgt, tgt_stride_f);
  double sum = 0; long nr = 0;
  /* serial accumulation in index order, like pcl::Registration::getFitnessScore */
  float* d2 = (float*)malloc(sizeof(float) * (size_t)n_src);
  int* ok = (int*)malloc(sizeof(int) * (size_t)n_src);
#pragma omp parallel for schedule(dynamic, 64) num_threads(num_threads > 0 ? num_threads : 1)
  for (int i = 0; i < n_src; i++) {
    float p[3] = {src[(size_t)i * src_stride_f], src[(size_t)i * src_stride_f + 1], src[(size_t)i * src_stride_f + 2]};
    float q[3];
    pcl_transform_pt(T, p, q);
    int nn; float dd;
    ok[i] = og_kdtree_knn(tree, q, 1, &nn, &dd);
    d2[i] = dd;
  }
  for (int i = 0; i < n_src; i++)
    if (ok[i] && d2[i] <= max_range) { sum += d2[i]; nr++; }
  free(d2); free(ok);
  og_kdtree_free(tree);
  return nr > 0 ? sum / (double)nr : DBL_MAX;
}
