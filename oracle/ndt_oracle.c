/*
 * ndt_oracle.c -- CPU restatement of LOCUS's alternative registration method
 * (`registration_method: ndt`): the OpenMP NDT fork under
 * multithreaded_gicp/include/multithreaded_ndt/.
 *
 * TEST INFRASTRUCTURE ONLY (see lb_oracle.h).
 *
 * What follows which reference lines:
 *   target voxel Gaussians   voxel_grid_covariance_omp_impl.hpp:48-370 (applyFilter), defaults
 *                            voxel_grid_covariance_omp.h:186-187 (6 points per voxel, eigenvalue floor 0.01)
 *   neighbour search         voxel_grid_covariance_omp.h:433-466 (radius search over the voxel centroids, the
 *                            default KDTREE method), voxel_grid_covariance_omp_impl.hpp:373-440 (DIRECT7 / DIRECT1)
 *   derivatives              ndt_omp_impl.hpp:221-343 (computeDerivatives), :350-476 (angle derivatives),
 *                            :478-526 (point derivatives, float), :574-638 (updateDerivatives, float),
 *                            :641-756 (computeHessian / updateHessian, double)
 *   Newton + line search     ndt_omp_impl.hpp:100-208 (computeTransformation), :758-885 (More-Thuente helpers),
 *                            :887-1063 (computeStepLengthMT)
 *   call sites               PointCloudOdometry.cc:182-195, PointCloudLocalization.cc:267-280 (defaults of the
 *                            class: resolution 1.0, step 0.1, outlier ratio 0.55; epsilon / iterations from yaml)
 *
 * PARITY UNPINNED.  The reference holds no test for NDT at all
 * (test_point_cloud_odometry.cpp:19 "TODO: add tests for ndt") and PCL / Eigen /
 * FLANN are absent here, so nothing can pin this file beyond its own sanity
 * checks (tests/test_ndt_cpu.py: recovers a known offset; score / gradient /
 * Hessian agree with an independent float64 restatement of the NDT score and
 * its finite differences to 6e-8 / 4e-7 / 6e-5; voxel Gaussians agree with a
 * numpy restatement; pose conventions agree with scipy; the lattice neighbour
 * search of the product agrees with this file's kd-tree search).  Choices that live in those libraries and are made
 * here: Eigen's fixed-size float products are summed left to right; `exp` of
 * the float argument is expf; SelfAdjointEigenSolver<Matrix3d> is a cyclic
 * Jacobi; JacobiSVD<6x6>::solve is a one-sided Jacobi SVD with Eigen's rank
 * threshold; Matrix3d::inverse is the cofactor formula;
 * Transform::rotation() is the linear part itself; eulerAngles(0,1,2) follows
 * Eigen 3.3's published formula; FLANN's radius search is strict (d2 < r2) and
 * ordered by (d2, index); std::map iteration = ascending voxel index;
 * pcl::getAllNeighborCellIndices (DIRECT26) is restated from PCL 1.10.
 */
#include "lb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void og_sym3_jacobi(double A[3][3], double d[3], double V[3][3]); /* gicp_oracle.c */

typedef struct {
  int leaf_idx;
  int nr_points; /* -1 after a failed eigen / inverse check (voxel_grid_covariance_omp_impl.hpp:326-351) */
  int in_centroids; /* 1: was pushed to voxel_centroids_ (>= min points at that time) */
  double mean[3];
  double cov[9];
  double icov[9];
  float centroid[3];
} ndt_leaf;

struct og_ndt_target {
  og_ndt_params P;
  int n_all;        /* every occupied voxel, ascending leaf index (std::map order) */
  ndt_leaf* all;
  int n_valid;      /* voxel_centroids_: leaves that had >= min_points_per_voxel */
  int* valid;       /* index into all[] */
  float* centroids; /* n_valid x 3 */
  og_kdtree* tree;
  int min_b[3], max_b[3], div_b[3], divb_mul[3];
  float leaf, inv_leaf;
  int status;
};

void og_ndt_default_params(og_ndt_params* p) {
  p->resolution = 1.0f;           /* ndt_omp_impl.hpp:50 */
  p->step_size = 0.1;             /* :51 */
  p->outlier_ratio = 0.55;        /* :52 */
  p->transformation_epsilon = 0.1;/* :93 */
  p->max_iterations = 35;         /* :94 */
  p->min_points_per_voxel = 6;    /* voxel_grid_covariance_omp.h:186 */
  p->min_covar_eigvalue_mult = 0.01; /* :187 */
  p->search_method = 0;           /* KDTREE, ndt_omp_impl.hpp:96 */
  p->num_threads = 1;
}

/* ------------------------------------------------------------------ small linear algebra */
static void inv3(const double m[9], double out[9]) {
  /* cofactor formula (Eigen compute_inverse_size3) */
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[5] * m[6] - m[3] * m[8];
  double c20 = m[3] * m[7] - m[4] * m[6];
  double det = (m[0] * c00 + m[1] * c10) + m[2] * c20;
  double id = 1.0 / det;
  out[0] = c00 * id;
  out[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  out[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  out[3] = c10 * id;
  out[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  out[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  out[6] = c20 * id;
  out[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  out[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

static void mul3(const double a[9], const double b[9], double o[9]) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
}

/* x = pinv(A) b through a one-sided (Hestenes) Jacobi SVD; stands in for JacobiSVD<Matrix6d>(A, FullU|FullV).solve(b) */
static void svd6_solve(const double A[36], const double b[6], double x[6]) {
  double U[6][6], V[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { U[i][j] = A[6 * i + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  /* column pairs in round-robin order (five rounds of three disjoint pairs per sweep): any cyclic order converges to the
   * same decomposition; this one lets an implementation rotate the three pairs of a round side by side */
  static const int PP[15] = {0, 1, 2, 0, 3, 1, 0, 2, 1, 0, 1, 4, 0, 2, 3};
  static const int QQ[15] = {5, 4, 3, 4, 5, 2, 3, 4, 5, 2, 3, 5, 1, 5, 4};
  for (int sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (int e = 0; e < 15; e++) {
      const int p = PP[e], q = QQ[e];
      double alpha = 0, beta = 0, gamma = 0;
      for (int k = 0; k < 6; k++) { alpha += U[k][p] * U[k][p]; beta += U[k][q] * U[k][q]; gamma += U[k][p] * U[k][q]; }
      if (gamma == 0.0 || fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta)) continue;
      rotated = 1;
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      if (zeta < 0.0) t = -t;
      double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
      for (int k = 0; k < 6; k++) {
        double up = U[k][p], uq = U[k][q];
        U[k][p] = c * up - s * uq; U[k][q] = s * up + c * uq;
        double vp = V[k][p], vq = V[k][q];
        V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  double sig[6]; int ord[6];
  for (int j = 0; j < 6; j++) {
    double s2 = 0;
    for (int k = 0; k < 6; k++) s2 += U[k][j] * U[k][j];
    sig[j] = sqrt(s2); ord[j] = j;
  }
  for (int i = 1; i < 6; i++) { /* descending singular values, stable */
    int o = ord[i], j = i;
    while (j > 0 && sig[ord[j - 1]] < sig[o]) { ord[j] = ord[j - 1]; j--; }
    ord[j] = o;
  }
  double thr = sig[ord[0]] * (6.0 * DBL_EPSILON);
  if (thr < DBL_MIN) thr = DBL_MIN;
  for (int r = 0; r < 6; r++) x[r] = 0.0;
  for (int jj = 0; jj < 6; jj++) {
    int j = ord[jj];
    if (!(sig[j] > thr)) break;
    double ub = 0;
    for (int k = 0; k < 6; k++) ub += (U[k][j] / sig[j]) * b[k];
    double w = ub / sig[j];
    for (int r = 0; r < 6; r++) x[r] += V[r][j] * w;
  }
}

/* ------------------------------------------------------------------ target voxel Gaussians */
typedef struct { int key; int i; } key_i;
static int cmp_key_i(const void* a, const void* b) {
  const key_i* x = (const key_i*)a; const key_i* y = (const key_i*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->i < y->i ? -1 : (x->i > y->i);
}

static void finish_leaf(ndt_leaf* L, const og_ndt_params* P, const float csum[3]) {
  const int n = L->nr_points;
  for (int a = 0; a < 3; a++) L->centroid[a] = csum[a] / (float)n;   /* :284 */
  double pt_sum[3] = {L->mean[0], L->mean[1], L->mean[2]};
  for (int a = 0; a < 3; a++) L->mean[a] = L->mean[a] / n;           /* :286 */
  L->in_centroids = 0;
  if (n < P->min_points_per_voxel) return;
  L->in_centroids = 1;
  /* :319-320 */
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++)
      L->cov[3 * a + b] = (L->cov[3 * a + b] - 2 * (pt_sum[a] * L->mean[b])) / n + L->mean[a] * L->mean[b];
  const double scale = (n - 1.0) / n;
  for (int e = 0; e < 9; e++) L->cov[e] *= scale;
  /* :323-325 SelfAdjointEigenSolver reads the lower triangle */
  double A[3][3], d[3], V[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) A[a][b] = (a >= b) ? L->cov[3 * a + b] : L->cov[3 * b + a];
  og_sym3_jacobi(A, d, V);
  int o[3] = {0, 1, 2};   /* ascending eigenvalues */
  for (int i = 1; i < 3; i++) { int t = o[i], j = i; while (j > 0 && d[o[j - 1]] > d[t]) { o[j] = o[j - 1]; j--; } o[j] = t; }
  double ev[3] = {d[o[0]], d[o[1]], d[o[2]]};
  double E[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) E[3 * r + c] = V[r][o[c]];
  if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { L->nr_points = -1; return; }   /* :328-332 */
  const double floor_ev = P->min_covar_eigvalue_mult * ev[2];              /* :336 */
  if (ev[0] < floor_ev) {
    ev[0] = floor_ev;
    if (ev[1] < floor_ev) ev[1] = floor_ev;
    double ED[9], Ei[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) ED[3 * r + c] = E[3 * r + c] * ev[c];
    inv3(E, Ei);
    mul3(ED, Ei, L->cov);                                                  /* :346 */
  }
  inv3(L->cov, L->icov);                                                   /* :350 */
  double mx = L->icov[0], mn = L->icov[0];
  for (int e = 1; e < 9; e++) { if (L->icov[e] > mx) mx = L->icov[e]; if (L->icov[e] < mn) mn = L->icov[e]; }
  if (mx == (double)INFINITY || mn == -(double)INFINITY) L->nr_points = -1; /* :351-355 */
}

og_ndt_target* og_ndt_target_build(const float* pts, int n, int stride_f, const og_ndt_params* P) {
  og_ndt_target* t = (og_ndt_target*)calloc(1, sizeof(*t));
  t->P = *P;
  t->leaf = P->resolution;
  t->inv_leaf = 1.0f / P->resolution;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int n_fin = 0;
  for (int i = 0; i < n; i++) {
    const float* p = pts + (size_t)i * stride_f;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    for (int a = 0; a < 3; a++) { if (p[a] < mn[a]) mn[a] = p[a]; if (p[a] > mx[a]) mx[a] = p[a]; }
    n_fin++;
  }
  if (n_fin == 0) { t->status = -1; return t; }
  int64_t d[3];
  for (int a = 0; a < 3; a++) d[a] = (int64_t)((mx[a] - mn[a]) * t->inv_leaf) + 1;   /* :71-73 */
  if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) { t->status = -2; return t; }           /* :75-80 */
  for (int a = 0; a < 3; a++) {
    t->min_b[a] = (int)floorf(mn[a] * t->inv_leaf);
    t->max_b[a] = (int)floorf(mx[a] * t->inv_leaf);
    t->div_b[a] = t->max_b[a] - t->min_b[a] + 1;
  }
  t->divb_mul[0] = 1; t->divb_mul[1] = t->div_b[0]; t->divb_mul[2] = t->div_b[0] * t->div_b[1];
  key_i* ks = (key_i*)malloc(sizeof(key_i) * (size_t)n_fin);
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float* p = pts + (size_t)i * stride_f;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    int ijk[3];
    for (int a = 0; a < 3; a++) ijk[a] = (int)(floorf(p[a] * t->inv_leaf) - (float)t->min_b[a]);   /* :204-206 */
    ks[m].key = ijk[0] * t->divb_mul[0] + ijk[1] * t->divb_mul[1] + ijk[2] * t->divb_mul[2];
    ks[m].i = i; m++;
  }
  qsort(ks, (size_t)m, sizeof(key_i), cmp_key_i);
  int n_all = 0;
  for (int i = 0; i < m; i++) if (i == 0 || ks[i].key != ks[i - 1].key) n_all++;
  t->n_all = n_all;
  t->all = (ndt_leaf*)calloc((size_t)n_all, sizeof(ndt_leaf));
  int li = -1;
  float csum[3] = {0, 0, 0};
  for (int i = 0; i <= m; i++) {
    if (i == m || i == 0 || ks[i].key != ks[i - 1].key) {
      if (li >= 0) finish_leaf(&t->all[li], P, csum);
      if (i == m) break;
      li++;
      t->all[li].leaf_idx = ks[i].key;
      csum[0] = csum[1] = csum[2] = 0.f;
    }
    const float* p = pts + (size_t)ks[i].i * stride_f;
    ndt_leaf* L = &t->all[li];
    double pd[3] = {p[0], p[1], p[2]};
    for (int a = 0; a < 3; a++) { L->mean[a] += pd[a]; csum[a] += p[a]; }     /* :218-225 */
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) L->cov[3 * a + b] += pd[a] * pd[b];
    L->nr_points++;
  }
  free(ks);
  t->valid = (int*)malloc(sizeof(int) * (size_t)(n_all + 1));
  for (int i = 0; i < n_all; i++) if (t->all[i].in_centroids) t->valid[t->n_valid++] = i;
  t->centroids = (float*)malloc(sizeof(float) * 3 * (size_t)(t->n_valid + 1));
  for (int v = 0; v < t->n_valid; v++) memcpy(&t->centroids[3 * v], t->all[t->valid[v]].centroid, 3 * sizeof(float));
  if (t->n_valid > 0) t->tree = og_kdtree_build(t->centroids, t->n_valid, 3);
  return t;
}

void og_ndt_target_free(og_ndt_target* t) {
  if (!t) return;
  if (t->tree) og_kdtree_free(t->tree);
  free(t->centroids); free(t->valid); free(t->all); free(t);
}

int og_ndt_target_status(const og_ndt_target* t) { return t->status; }

int og_ndt_target_info(const og_ndt_target* t, int* n_valid, int* n_all, int min_b[3], int div_b[3]) {
  if (n_valid) *n_valid = t->n_valid;
  if (n_all) *n_all = t->n_all;
  for (int a = 0; a < 3; a++) { if (min_b) min_b[a] = t->min_b[a]; if (div_b) div_b[a] = t->div_b[a]; }
  return t->status;
}

void og_ndt_target_leaves(const og_ndt_target* t, int* leaf_idx, int* nr_points, double* mean3, double* icov9, float* centroid3) {
  for (int v = 0; v < t->n_valid; v++) {
    const ndt_leaf* L = &t->all[t->valid[v]];
    if (leaf_idx) leaf_idx[v] = L->leaf_idx;
    if (nr_points) nr_points[v] = L->nr_points;
    if (mean3) memcpy(&mean3[3 * v], L->mean, 3 * sizeof(double));
    if (icov9) memcpy(&icov9[9 * v], L->icov, 9 * sizeof(double));
    if (centroid3) memcpy(&centroid3[3 * v], L->centroid, 3 * sizeof(float));
  }
}

/* ------------------------------------------------------------------ neighbourhoods */
#define NDT_MAX_NB 64
static const ndt_leaf* find_leaf(const og_ndt_target* t, int idx) {
  int lo = 0, hi = t->n_all - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    if (t->all[mid].leaf_idx == idx) return &t->all[mid];
    if (t->all[mid].leaf_idx < idx) lo = mid + 1; else hi = mid - 1;
  }
  return NULL;
}

static int neighbourhood(const og_ndt_target* t, const float q[3], const ndt_leaf** out) {
  int k = 0;
  if (t->P.search_method == 0) {                    /* KDTREE: radiusSearch(point, resolution_) */
    if (!t->tree) return 0;
    int idx[NDT_MAX_NB]; float d2[NDT_MAX_NB];
    const double radius = (double)t->P.resolution;
    int found = og_kdtree_radius(t->tree, q, (float)(radius * radius), idx, d2, NDT_MAX_NB);
    if (found > NDT_MAX_NB) found = NDT_MAX_NB;
    for (int i = 0; i < found; i++) out[k++] = &t->all[t->valid[idx[i]]];
    return k;
  }
  static const int REL7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  /* DIRECT26 = pcl::getAllNeighborCellIndices() (pcl/filters/voxel_grid.h, NOT in the reference tree; restated from PCL 1.10):
   * the 13 "half" offsets -- (i, j, -1) for i, j in -1..1, then (i, -1, 0) for i in -1..1, then (-1, 0, 0) -- followed by
   * their negatives.  The point's own cell is not among them (voxel_grid_covariance_omp_impl.hpp:403-411). */
  int rel26[26][3];
  if (t->P.search_method == 1) {
    int n = 0;
    for (int i = -1; i < 2; i++) for (int j = -1; j < 2; j++) { rel26[n][0] = i; rel26[n][1] = j; rel26[n][2] = -1; n++; }
    for (int i = -1; i < 2; i++) { rel26[n][0] = i; rel26[n][1] = -1; rel26[n][2] = 0; n++; }
    rel26[n][0] = -1; rel26[n][1] = 0; rel26[n][2] = 0; n++;
    for (int m = 0; m < 13; m++) for (int a = 0; a < 3; a++) rel26[13 + m][a] = -rel26[m][a];
  }
  const int nrel = t->P.search_method == 3 ? 1 : (t->P.search_method == 1 ? 26 : 7);
  int ijk[3];
  for (int a = 0; a < 3; a++) ijk[a] = (int)floorf(q[a] / t->leaf);   /* voxel_grid_covariance_omp_impl.hpp:378-380 */
  for (int r = 0; r < nrel; r++) {
    int ok = 1, idx = 0;
    for (int a = 0; a < 3; a++) {
      int c = ijk[a] + (t->P.search_method == 1 ? rel26[r][a] : REL7[r][a]);
      if (c < t->min_b[a] || c > t->max_b[a]) ok = 0;
      idx += (c - t->min_b[a]) * t->divb_mul[a];
    }
    if (!ok) continue;
    const ndt_leaf* L = find_leaf(t, idx);
    if (L && L->nr_points >= t->P.min_points_per_voxel) out[k++] = L;
  }
  return k;
}

/* ------------------------------------------------------------------ derivatives */
typedef struct {
  double jd[8][3]; float jf[8][3];
  double hd[15][3]; float hf[15][3];
} ang_t;

static void angle_derivatives(const double p[6], ang_t* A) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
#define ROW(M, r, a, b, c) do { M[r][0] = (a); M[r][1] = (b); M[r][2] = (c); } while (0)
  /* Magnusson 2009 eq. 6.19 (rows a..h) */
  ROW(A->jd, 0, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
  ROW(A->jd, 1, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
  ROW(A->jd, 2, (-sy * cz), sy * sz, cy);
  ROW(A->jd, 3, sx * cy * cz, (-sx * cy * sz), sx * sy);
  ROW(A->jd, 4, (-cx * cy * cz), cx * cy * sz, (-cx * sy));
  ROW(A->jd, 5, (-cy * sz), (-cy * cz), 0.0);
  ROW(A->jd, 6, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0.0);
  ROW(A->jd, 7, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0.0);
  /* eq. 6.21 (a2 a3 b2 b3 c2 c3 d1 d2 d3 e1 e2 e3 f1 f2 f3) */
  ROW(A->hd, 0, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
  ROW(A->hd, 1, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
  ROW(A->hd, 2, (cx * cy * cz), (-cx * cy * sz), (cx * sy));
  ROW(A->hd, 3, (sx * cy * cz), (-sx * cy * sz), (sx * sy));
  ROW(A->hd, 4, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0.0);
  ROW(A->hd, 5, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0.0);
  ROW(A->hd, 6, (-cy * cz), (cy * sz), (sy));
  ROW(A->hd, 7, (-sx * sy * cz), (sx * sy * sz), (sx * cy));
  ROW(A->hd, 8, (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
  ROW(A->hd, 9, (sy * sz), (sy * cz), 0.0);
  ROW(A->hd, 10, (-sx * cy * sz), (-sx * cy * cz), 0.0);
  ROW(A->hd, 11, (cx * cy * sz), (cx * cy * cz), 0.0);
  ROW(A->hd, 12, (-cy * cz), (cy * sz), 0.0);
  ROW(A->hd, 13, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0.0);
  ROW(A->hd, 14, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0.0);
#undef ROW
  for (int r = 0; r < 8; r++) for (int c = 0; c < 3; c++) A->jf[r][c] = (float)A->jd[r][c];
  for (int r = 0; r < 15; r++) for (int c = 0; c < 3; c++) A->hf[r][c] = (float)A->hd[r][c];
}

/* which of the six second-derivative vectors a..f sits at block (i, j), i, j in 3..5 (ndt_omp_impl.hpp:513-521) */
static const int HBLK[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};

typedef struct { double d1, d2, d3; } gauss_t;
static void gauss_constants(const og_ndt_params* P, gauss_t* G) {
  double c1 = 10 * (1 - P->outlier_ratio);
  double c2 = P->outlier_ratio / pow((double)P->resolution, 3);
  G->d3 = -log(c2);
  G->d1 = -log(c1 + c2) - G->d3;
  G->d2 = -2 * log((-log(c1 * exp(-0.5) + c2) - G->d3) / G->d1);
}

/* updateDerivatives (ndt_omp_impl.hpp:574-638): float arithmetic, double accumulators */
static double update_derivatives(double g[6], double H[36], const float pg[3][6], const float ph[6][3],
                                 const double xt_d[3], const double cinv[9], const gauss_t* G, int compute_hessian) {
  float xt[3] = {(float)xt_d[0], (float)xt_d[1], (float)xt_d[2]};
  float cf[3][3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) cf[r][c] = (float)cinv[3 * r + c];
  const float d2f = (float)G->d2;
  float xc[3];
  for (int c = 0; c < 3; c++) xc[c] = (xt[0] * cf[0][c] + xt[1] * cf[1][c]) + xt[2] * cf[2][c];
  float e = expf(-d2f * ((xt[0] * xc[0] + xt[1] * xc[1]) + xt[2] * xc[2]) * 0.5f);
  float score_inc = (float)(-G->d1 * (double)e);
  e = d2f * e;
  if (e > 1 || e < 0 || e != e) return 0;
  e = (float)((double)e * G->d1);
  float CP[3][6];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 6; c++) CP[r][c] = (cf[r][0] * pg[0][c] + cf[r][1] * pg[1][c]) + cf[r][2] * pg[2][c];
  float gq[6];
  for (int c = 0; c < 6; c++) gq[c] = (xt[0] * CP[0][c] + xt[1] * CP[1][c]) + xt[2] * CP[2][c];
  for (int c = 0; c < 6; c++) g[c] += (double)(e * gq[c]);
  if (compute_hessian) {
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        float xH = 0.0f;
        if (i >= 3 && j >= 3) {
          const float* v = ph[HBLK[i - 3][j - 3]];
          xH = (xc[0] * v[0] + xc[1] * v[1]) + xc[2] * v[2];
        }
        float JCJ = (pg[0][j] * CP[0][i] + pg[1][j] * CP[1][i]) + pg[2][j] * CP[2][i];
        H[6 * i + j] += (double)(e * ((-d2f * gq[i] * gq[j] + xH) + JCJ));
      }
  }
  return (double)score_inc;
}

static void point_derivatives_f(const float x[3], const ang_t* A, float pg[3][6], float ph[6][3]) {
  float xj[8], xh[15];
  for (int r = 0; r < 8; r++) xj[r] = (A->jf[r][0] * x[0] + A->jf[r][1] * x[1]) + A->jf[r][2] * x[2];
  for (int r = 0; r < 15; r++) xh[r] = (A->hf[r][0] * x[0] + A->hf[r][1] * x[1]) + A->hf[r][2] * x[2];
  memset(pg, 0, sizeof(float) * 18);
  pg[0][0] = pg[1][1] = pg[2][2] = 1.0f;
  pg[1][3] = xj[0]; pg[2][3] = xj[1];
  pg[0][4] = xj[2]; pg[1][4] = xj[3]; pg[2][4] = xj[4];
  pg[0][5] = xj[5]; pg[1][5] = xj[6]; pg[2][5] = xj[7];
  ph[0][0] = 0; ph[0][1] = xh[0]; ph[0][2] = xh[1];      /* a */
  ph[1][0] = 0; ph[1][1] = xh[2]; ph[1][2] = xh[3];      /* b */
  ph[2][0] = 0; ph[2][1] = xh[4]; ph[2][2] = xh[5];      /* c */
  ph[3][0] = xh[6]; ph[3][1] = xh[7]; ph[3][2] = xh[8];  /* d */
  ph[4][0] = xh[9]; ph[4][1] = xh[10]; ph[4][2] = xh[11];/* e */
  ph[5][0] = xh[12]; ph[5][1] = xh[13]; ph[5][2] = xh[14];/* f */
}

static inline double dot3d(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void mv3d(const double m[9], const double v[3], double o[3]) {
  for (int r = 0; r < 3; r++) o[r] = (m[3 * r] * v[0] + m[3 * r + 1] * v[1]) + m[3 * r + 2] * v[2];
}

/* computeHessian / updateHessian (ndt_omp_impl.hpp:641-756): double arithmetic, serial over the points */
static void hessian_point_d(double H[36], const float xf[3], const double xt[3], const double cinv[9], const ang_t* A, const gauss_t* G) {
  double x[3] = {xf[0], xf[1], xf[2]};
  double pg[3][6]; memset(pg, 0, sizeof(pg));
  pg[0][0] = pg[1][1] = pg[2][2] = 1.0;
  pg[1][3] = dot3d(x, A->jd[0]); pg[2][3] = dot3d(x, A->jd[1]);
  pg[0][4] = dot3d(x, A->jd[2]); pg[1][4] = dot3d(x, A->jd[3]); pg[2][4] = dot3d(x, A->jd[4]);
  pg[0][5] = dot3d(x, A->jd[5]); pg[1][5] = dot3d(x, A->jd[6]); pg[2][5] = dot3d(x, A->jd[7]);
  double ph[6][3];
  ph[0][0] = 0; ph[0][1] = dot3d(x, A->hd[0]); ph[0][2] = dot3d(x, A->hd[1]);
  ph[1][0] = 0; ph[1][1] = dot3d(x, A->hd[2]); ph[1][2] = dot3d(x, A->hd[3]);
  ph[2][0] = 0; ph[2][1] = dot3d(x, A->hd[4]); ph[2][2] = dot3d(x, A->hd[5]);
  for (int k = 0; k < 3; k++) { ph[3][k] = dot3d(x, A->hd[6 + k]); ph[4][k] = dot3d(x, A->hd[9 + k]); ph[5][k] = dot3d(x, A->hd[12 + k]); }
  double cx[3]; mv3d(cinv, xt, cx);
  double e = G->d2 * exp(-G->d2 * dot3d(xt, cx) / 2);
  if (e > 1 || e < 0 || e != e) return;
  e *= G->d1;
  for (int i = 0; i < 6; i++) {
    double ci[3] = {pg[0][i], pg[1][i], pg[2][i]}, cov_dxd_pi[3];
    mv3d(cinv, ci, cov_dxd_pi);
    for (int j = 0; j < 6; j++) {
      double cj[3] = {pg[0][j], pg[1][j], pg[2][j]}, ccj[3], chv[3] = {0, 0, 0}, hv[3] = {0, 0, 0};
      mv3d(cinv, cj, ccj);
      if (i >= 3 && j >= 3) memcpy(hv, ph[HBLK[i - 3][j - 3]], sizeof(hv));
      mv3d(cinv, hv, chv);
      H[6 * i + j] += e * ((-G->d2 * dot3d(xt, cov_dxd_pi) * dot3d(xt, ccj) + dot3d(xt, chv)) + dot3d(cj, cov_dxd_pi));
    }
  }
}

/* pcl::transformPointCloud, PCL 1.10 se3 association */
static inline void se3(const float* T, const float* p, float* o) {
  o[0] = T[0] * p[0] + (T[1] * p[1] + (T[2] * p[2] + T[3]));
  o[1] = T[4] * p[0] + (T[5] * p[1] + (T[6] * p[2] + T[7]));
  o[2] = T[8] * p[0] + (T[9] * p[1] + (T[10] * p[2] + T[11]));
}

/* computeDerivatives on trans = T * src.  Returns the score. */
static double compute_derivatives(const og_ndt_target* t, const float* src, int n, int stride_f, const float* trans /* n x 3 */,
                                  const double p[6], int compute_hessian, double g[6], double H[36], ang_t* A_out) {
  gauss_t G; gauss_constants(&t->P, &G);
  ang_t A; angle_derivatives(p, &A);
  if (A_out) *A_out = A;
  double* per = (double*)calloc((size_t)n * 43 + 1, sizeof(double));
#pragma omp parallel for num_threads(t->P.num_threads > 0 ? t->P.num_threads : 1) schedule(guided, 8)
  for (int idx = 0; idx < n; idx++) {
    const float* x = src + (size_t)idx * stride_f;
    const float* xt = trans + 3 * (size_t)idx;
    const ndt_leaf* nb[NDT_MAX_NB];
    int k = neighbourhood(t, xt, nb);
    double* o = per + (size_t)idx * 43;
    float pg[3][6], ph[6][3];
    for (int c = 0; c < k; c++) {
      double xd[3] = {(double)xt[0] - nb[c]->mean[0], (double)xt[1] - nb[c]->mean[1], (double)xt[2] - nb[c]->mean[2]};
      point_derivatives_f(x, &A, pg, ph);
      o[0] += update_derivatives(o + 1, o + 7, pg, ph, xd, nb[c]->icov, &G, compute_hessian);
    }
  }
  double score = 0;
  memset(g, 0, 6 * sizeof(double)); memset(H, 0, 36 * sizeof(double));
  for (int i = 0; i < n; i++) {             /* ndt_omp_impl.hpp:336-340: summed in point order */
    const double* o = per + (size_t)i * 43;
    score += o[0];
    for (int c = 0; c < 6; c++) g[c] += o[1 + c];
    for (int c = 0; c < 36; c++) H[c] += o[7 + c];
  }
  free(per);
  return score;
}

static void compute_hessian_d(const og_ndt_target* t, const float* src, int n, int stride_f, const float* trans, const ang_t* A, double H[36]) {
  gauss_t G; gauss_constants(&t->P, &G);
  memset(H, 0, 36 * sizeof(double));
  for (int idx = 0; idx < n; idx++) {
    const float* xt = trans + 3 * (size_t)idx;
    const ndt_leaf* nb[NDT_MAX_NB];
    int k = neighbourhood(t, xt, nb);
    for (int c = 0; c < k; c++) {
      double xd[3] = {(double)xt[0] - nb[c]->mean[0], (double)xt[1] - nb[c]->mean[1], (double)xt[2] - nb[c]->mean[2]};
      hessian_point_d(H, src + (size_t)idx * stride_f, xd, nb[c]->icov, A, &G);
    }
  }
}

static void transform_cloud(const float* src, int n, int stride_f, const float* T, float* out) {
  for (int i = 0; i < n; i++) se3(T, src + (size_t)i * stride_f, out + 3 * (size_t)i);
}

int og_ndt_derivatives(const og_ndt_target* t, const float* src, int n, int stride_f, const float* T16, const double p[6],
                       int compute_hessian, double* score, double g[6], double H[36]) {
  float* trans = (float*)malloc(sizeof(float) * 3 * (size_t)(n + 1));
  transform_cloud(src, n, stride_f, T16, trans);
  *score = compute_derivatives(t, src, n, stride_f, trans, p, compute_hessian, g, H, NULL);
  free(trans);
  return 0;
}

int og_ndt_hessian(const og_ndt_target* t, const float* src, int n, int stride_f, const float* T16, const double p[6], double H[36]) {
  float* trans = (float*)malloc(sizeof(float) * 3 * (size_t)(n + 1));
  transform_cloud(src, n, stride_f, T16, trans);
  ang_t A; angle_derivatives(p, &A);
  compute_hessian_d(t, src, n, stride_f, trans, &A, H);
  free(trans);
  return 0;
}

/* ------------------------------------------------------------------ pose <-> matrix */
/* (Translation3f(p0..2) * AngleAxisf(p3, X) * AngleAxisf(p4, Y) * AngleAxisf(p5, Z)).matrix(), row-major 4x4 */
static void axis_rotation(float angle, int axis, float R[9]) {
  /* Eigen AngleAxis::toRotationMatrix with a unit axis vector */
  float ax[3] = {0, 0, 0}; ax[axis] = 1.0f;
  float s = sinf(angle), c = cosf(angle);
  float sin_axis[3] = {s * ax[0], s * ax[1], s * ax[2]};
  float cos1_axis[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float tmp;
  tmp = cos1_axis[0] * ax[1]; R[1] = tmp - sin_axis[2]; R[3] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * ax[2]; R[2] = tmp + sin_axis[1]; R[6] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * ax[2]; R[5] = tmp - sin_axis[0]; R[7] = tmp + sin_axis[0];
  R[0] = cos1_axis[0] * ax[0] + c; R[4] = cos1_axis[1] * ax[1] + c; R[8] = cos1_axis[2] * ax[2] + c;
}
static void mul3f(const float a[9], const float b[9], float o[9]) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
}
void og_ndt_pose_to_matrix(const double p[6], float T[16]) {
  float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
  axis_rotation((float)p[3], 0, Rx); axis_rotation((float)p[4], 1, Ry); axis_rotation((float)p[5], 2, Rz);
  mul3f(Rx, Ry, Rxy); mul3f(Rxy, Rz, R);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[4 * r + c] = R[3 * r + c]; T[4 * r + 3] = (float)p[r]; }
  T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

/* Matrix3f::eulerAngles(0, 1, 2) (Eigen 3.3), float */
void og_ndt_euler_xyz(const float T[16], float out[3]) {
#define M(r, c) T[4 * (r) + (c)]
  const float pi = 3.14159265358979323846f;
  float r0 = atan2f(M(1, 2), M(2, 2));
  float c2 = sqrtf(M(0, 0) * M(0, 0) + M(0, 1) * M(0, 1));
  float r1;
  if (r0 > 0.f) {
    r0 -= pi;          /* (!odd && res[0] > 0) and res[0] > 0 */
    r1 = atan2f(-M(0, 2), -c2);
  } else {
    r1 = atan2f(-M(0, 2), c2);
  }
  float s1 = sinf(r0), c1 = cosf(r0);
  float r2 = atan2f(s1 * M(2, 0) - c1 * M(1, 0), c1 * M(1, 1) - s1 * M(2, 1));
  out[0] = -r0; out[1] = -r1; out[2] = -r2;
#undef M
}

/* ------------------------------------------------------------------ More-Thuente (ndt_omp_impl.hpp:758-885) */
static int update_interval(double* a_l, double* f_l, double* g_l, double* a_u, double* f_u, double* g_u, double a_t, double f_t, double g_t) {
  if (f_t > *f_l) { *a_u = a_t; *f_u = f_t; *g_u = g_t; return 0; }
  else if (g_t * (*a_l - a_t) > 0) { *a_l = a_t; *f_l = f_t; *g_l = g_t; return 0; }
  else if (g_t * (*a_l - a_t) < 0) { *a_u = *a_l; *f_u = *f_l; *g_u = *g_l; *a_l = a_t; *f_l = f_t; *g_l = g_t; return 0; }
  return 1;
}

static double trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {                                   /* case 1 */
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return fabs(a_c - a_l) < fabs(a_q - a_l) ? a_c : 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {                        /* case 2 */
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return fabs(a_c - a_t) >= fabs(a_s - a_t) ? a_c : a_s;
  } else if (fabs(g_t) <= fabs(g_l)) {               /* case 3 */
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_n = fabs(a_c - a_t) < fabs(a_s - a_t) ? a_c : a_s;
    double lim = a_t + 0.66 * (a_u - a_t);
    if (a_t > a_l) return lim < a_n ? lim : a_n;
    return lim > a_n ? lim : a_n;
  }
  double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;   /* case 4 */
  double w = sqrt(z * z - g_t * g_u);
  return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

typedef struct {
  const og_ndt_target* t; const float* src; int n, stride_f;
  float* trans; float final[16]; long n_evals;
} align_ctx;

static double dot6(const double* a, const double* b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }

static double step_length_mt(align_ctx* c, const double x[6], double step_dir[6], double step_init, double step_max, double step_min,
                             double* score, double g[6], double H[36]) {
  double phi_0 = -*score;
  double d_phi_0 = -dot6(g, step_dir);
  double x_t[6];
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int i = 0; i < 6; i++) step_dir[i] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = phi_0 - phi_0 - mu * d_phi_0 * a_l, g_l = d_phi_0 - mu * d_phi_0;
  double f_u = phi_0 - phi_0 - mu * d_phi_0 * a_u, g_u = d_phi_0 - mu * d_phi_0;
  int interval_converged = (step_max - step_min) < 0, open_interval = 1;
  double a_t = step_init;
  a_t = a_t < step_max ? a_t : step_max;
  a_t = a_t > step_min ? a_t : step_min;
  for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
  og_ndt_pose_to_matrix(x_t, c->final);
  transform_cloud(c->src, c->n, c->stride_f, c->final, c->trans);
  ang_t A;
  *score = compute_derivatives(c->t, c->src, c->n, c->stride_f, c->trans, x_t, 1, g, H, &A); c->n_evals++;
  double phi_t = -*score, d_phi_t = -dot6(g, step_dir);
  double psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t, d_psi_t = d_phi_t - mu * d_phi_0;
  while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval) a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = a_t < step_max ? a_t : step_max;
    a_t = a_t > step_min ? a_t : step_min;
    for (int i = 0; i < 6; i++) x_t[i] = x[i] + step_dir[i] * a_t;
    og_ndt_pose_to_matrix(x_t, c->final);
    transform_cloud(c->src, c->n, c->stride_f, c->final, c->trans);
    *score = compute_derivatives(c->t, c->src, c->n, c->stride_f, c->trans, x_t, 0, g, H, &A); c->n_evals++;
    phi_t = -*score; d_phi_t = -dot6(g, step_dir);
    psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t; d_psi_t = d_phi_t - mu * d_phi_0;
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = 0;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
    }
    if (open_interval) interval_converged = update_interval(&a_l, &f_l, &g_l, &a_u, &f_u, &g_u, a_t, psi_t, d_psi_t);
    else interval_converged = update_interval(&a_l, &f_l, &g_l, &a_u, &f_u, &g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) compute_hessian_d(c->t, c->src, c->n, c->stride_f, c->trans, &A, H);
  return a_t;
}

static int is_identity16(const float* T) {
  for (int i = 0; i < 16; i++) if (T[i] != ((i % 5 == 0) ? 1.0f : 0.0f)) return 0;
  return 1;
}

int og_ndt_align(const og_ndt_target* t, const float* src, int n, int stride_f, const float* guess, og_ndt_result* R) {
  memset(R, 0, sizeof(*R));
  if (!t || t->status != 0) { R->status = -5; return R->status; }
  if (n <= 0) { R->status = -4; return R->status; }
  align_ctx c; memset(&c, 0, sizeof(c));
  c.t = t; c.src = src; c.n = n; c.stride_f = stride_f;
  c.trans = (float*)malloc(sizeof(float) * 3 * (size_t)n);
  for (int i = 0; i < 16; i++) c.final[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  if (guess && !is_identity16(guess)) memcpy(c.final, guess, 16 * sizeof(float));   /* ndt_omp_impl.hpp:116-121 */
  transform_cloud(src, n, stride_f, c.final, c.trans);
  float eul[3]; og_ndt_euler_xyz(c.final, eul);
  double p[6] = {c.final[3], c.final[7], c.final[11], eul[0], eul[1], eul[2]}, delta_p[6], g[6], H[36];
  double score = compute_derivatives(t, src, n, stride_f, c.trans, p, 1, g, H, NULL); c.n_evals++;
  int nr_iterations = 0, converged = 0;
  while (!converged) {
    double ng[6]; for (int i = 0; i < 6; i++) ng[i] = -g[i];
    svd6_solve(H, ng, delta_p);                                               /* :152-155 */
    double delta_p_norm = sqrt(dot6(delta_p, delta_p));
    if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {                  /* :161-165 */
      converged = delta_p_norm == delta_p_norm;
      break;
    }
    for (int i = 0; i < 6; i++) delta_p[i] /= delta_p_norm;
    delta_p_norm = step_length_mt(&c, p, delta_p, delta_p_norm, t->P.step_size, t->P.transformation_epsilon / 2, &score, g, H);
    for (int i = 0; i < 6; i++) { delta_p[i] *= delta_p_norm; p[i] = p[i] + delta_p[i]; }
    if (nr_iterations > t->P.max_iterations || (nr_iterations && (fabs(delta_p_norm) < t->P.transformation_epsilon))) converged = 1;   /* :196-200 */
    nr_iterations++;
  }
  memcpy(R->final_transformation, c.final, sizeof(c.final));
  R->converged = converged; R->nr_iterations = nr_iterations;
  R->trans_probability = score / (double)n;
  R->n_evaluations = c.n_evals;
  memcpy(R->pose, p, sizeof(p));
  free(c.trans);
  return 0;
}
