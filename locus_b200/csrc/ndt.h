// ndt.h -- scalar logic of the NDT registration (LOCUS's `registration_method: ndt`,
// multithreaded_gicp/include/multithreaded_ndt/).
//
// Like hd.h / bfgs.h everything here is `__host__ __device__`: the kernels of ndt.cu call it on the device, the CPU
// test harness (tests/ndt_harness.cpp) compiles it with g++ and runs it with a serial backend against the oracle.
// The product has no CPU compute path.
//
//   voxel Gaussians      voxel_grid_covariance_omp_impl.hpp:284-356   ndt_finish_voxel
//   neighbour search     voxel_grid_covariance_omp.h:433-466 (radius search over the voxel centroids) and
//                        voxel_grid_covariance_omp_impl.hpp:373-440 (DIRECT7 / DIRECT1)            ndt_neighbours
//   per-pair terms       ndt_omp_impl.hpp:478-526, 574-638 (float), 721-756 (double)                ndt_point_eval
//   angle derivatives    ndt_omp_impl.hpp:350-476                                                   ndt_angles
//   Newton step, More-Thuente line search, convergence   ndt_omp_impl.hpp:100-208, 758-1063        NdtCtl / ndt_ctl_*
//
// Where the reference nests loops around computeDerivatives(), this file has a resumable state machine: the controller
// (`ndt_ctl_advance`) consumes the sums of one evaluation and either finishes or posts the next request (transform,
// angles, which sums).  On the device it runs in one thread of a one-block kernel between two evaluation grids, so an
// align() never returns to the host between evaluations.
//
// Arithmetic follows the reference: float32 per-pair terms, double accumulators, double Newton / line search.
#pragma once

#include <float.h>

#include "hd.h"

namespace lb {

enum { NDT_KDTREE = 0, NDT_DIRECT26 = 1, NDT_DIRECT7 = 2, NDT_DIRECT1 = 3 };   // pclomp::NeighborSearchMethod
enum { NDT_WANT_NONE = 0, NDT_WANT_DERIV_H = 1, NDT_WANT_DERIV = 2, NDT_WANT_HESSIAN = 3 };
constexpr int NDT_NSUM = 43;                                        // score, gradient (6), Hessian (36)
constexpr int NDT_MAX_NB = 32;                                      // a ball of one voxel side meets at most 27 voxels

struct NdtGauss { double d1, d2, d3; };

struct NdtVoxel {          // 96 bytes, one per voxel of the centroid list
  double mean[3];
  double icov[9];
};

struct NdtAngles {
  float jf[8][3], hf[15][3];
  double jd[8][3], hd[15][3];
};

// Gaussian fitting constants (Magnusson 2009 eq. 6.8; ndt_omp_impl.hpp:107-113).  Evaluated on the host only.
LB_HD void ndt_gauss_constants(double outlier_ratio, float resolution, NdtGauss& G) {
  double c1 = 10 * (1 - outlier_ratio);
  double c2 = outlier_ratio / pow((double)resolution, 3);
  G.d3 = -log(c2);
  G.d1 = -log(c1 + c2) - G.d3;
  G.d2 = -2 * log((-log(c1 * exp(-0.5) + c2) - G.d3) / G.d1);
}

// Voxel lattice of the target (voxel_grid_covariance_omp_impl.hpp:67-103): bounding box of the finite points ->
// min_b / max_b / div_b.  Returns false when the int32 voxel index would overflow (the reference refuses such a grid).
struct NdtLattice {
  int min_b[3], max_b[3], div_b[3];
  float leaf, inv_leaf;
};
LB_HD bool ndt_lattice(const float* mn, const float* mx, float leaf, NdtLattice& L) {
  L.leaf = leaf;
  L.inv_leaf = 1.0f / leaf;
  long long d[3];
  for (int a = 0; a < 3; a++) d[a] = (long long)((mx[a] - mn[a]) * L.inv_leaf) + 1;
  if (d[0] * d[1] * d[2] > 2147483647LL) return false;
  for (int a = 0; a < 3; a++) {
    L.min_b[a] = (int)floorf(mn[a] * L.inv_leaf);
    L.max_b[a] = (int)floorf(mx[a] * L.inv_leaf);
    L.div_b[a] = L.max_b[a] - L.min_b[a] + 1;
  }
  return true;
}
LB_HD int ndt_voxel_key(const NdtLattice& L, float x, float y, float z) {
  int i0 = (int)(floorf(x * L.inv_leaf) - (float)L.min_b[0]);
  int i1 = (int)(floorf(y * L.inv_leaf) - (float)L.min_b[1]);
  int i2 = (int)(floorf(z * L.inv_leaf) - (float)L.min_b[2]);
  return i0 + i1 * L.div_b[0] + i2 * (L.div_b[0] * L.div_b[1]);
}

// ------------------------------------------------------------------ 3x3 helpers (double, row-major)
LB_HD void ndt_inv3(const double* m, double* out) {
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
LB_HD double ndt_dot3(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
LB_HD void ndt_mv3(const double* m, const double* v, double* o) {
#pragma unroll
  for (int r = 0; r < 3; r++) o[r] = (m[3 * r] * v[0] + m[3 * r + 1] * v[1]) + m[3 * r + 2] * v[2];
}

// ------------------------------------------------------------------ voxel Gaussian
// n points of one voxel: sum[3] = sum of the points (double), m2 = [xx xy xz yy yz zz] sums of products (double),
// csum = float sum in input order.  Returns nr_points as the reference leaves it (n, or -1 when the eigenvalue or
// inverse check fails); fills mean / icov / centroid.  Voxels with n < min_pts never reach this function.
LB_HD int ndt_finish_voxel(int n, const double* sum, const double* m2, const float* csum, double eig_mult, NdtVoxel& out,
                           float* centroid) {
#pragma unroll
  for (int a = 0; a < 3; a++) centroid[a] = csum[a] / (float)n;
  double mean[3];
#pragma unroll
  for (int a = 0; a < 3; a++) { mean[a] = sum[a] / n; out.mean[a] = mean[a]; }
  const int S6[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  double cov[9];
  const double scale = (n - 1.0) / n;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double v = (m2[S6[a][b]] - 2 * (sum[a] * mean[b])) / n + mean[a] * mean[b];
      cov[3 * a + b] = v * scale;
    }
#pragma unroll
  for (int e = 0; e < 9; e++) out.icov[e] = 0.0;
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) A[a][b] = (a >= b) ? cov[3 * a + b] : cov[3 * b + a];   // lower triangle, as the eigen solver reads it
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
    jacobi_rotate<0, 1, 2>(A, V, sweep);
    jacobi_rotate<0, 2, 1>(A, V, sweep);
    jacobi_rotate<1, 2, 0>(A, V, sweep);
  }
  double d[3] = {A[0][0], A[1][1], A[2][2]};
  int o0 = 0, o1 = 1, o2 = 2;                              // ascending, stable
  if (d[o0] > d[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (d[o1] > d[o2]) { int t = o1; o1 = o2; o2 = t; if (d[o0] > d[o1]) { t = o0; o0 = o1; o1 = t; } }
  double ev[3] = {d[o0], d[o1], d[o2]};
  if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) return -1;
  const double floor_ev = eig_mult * ev[2];
  if (ev[0] < floor_ev) {
    ev[0] = floor_ev;
    if (ev[1] < floor_ev) ev[1] = floor_ev;
    double E[9], ED[9], Ei[9];
#pragma unroll
    for (int r = 0; r < 3; r++) { E[3 * r] = V[r][o0]; E[3 * r + 1] = V[r][o1]; E[3 * r + 2] = V[r][o2]; }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) ED[3 * r + c] = E[3 * r + c] * ev[c];
    ndt_inv3(E, Ei);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) cov[3 * r + c] = (ED[3 * r] * Ei[c] + ED[3 * r + 1] * Ei[3 + c]) + ED[3 * r + 2] * Ei[6 + c];
  }
  ndt_inv3(cov, out.icov);
  double mx = out.icov[0], mn = out.icov[0];
#pragma unroll
  for (int e = 1; e < 9; e++) { if (out.icov[e] > mx) mx = out.icov[e]; if (out.icov[e] < mn) mn = out.icov[e]; }
  if (mx == (double)INFINITY || mn == -(double)INFINITY) return -1;
  return n;
}

// ------------------------------------------------------------------ the target as the evaluators see it
struct NdtTargetView {
  const NdtVoxel* vox;       // [n_valid] ascending voxel index = the reference's voxel_centroids_ order
  const f4* cen;             // [n_valid] float centroid, .w = nr_points as float bits (int)
  const uint32_t* hkey;      // open-addressing hash: voxel index -> slot
  const int32_t* hval;
  uint32_t hmask;
  int n_valid;
  int min_b[3], max_b[3], div_b[3];
  float leaf, inv_leaf, r2;  // r2 = float(resolution^2)
  int method, min_pts;
};

LB_HD uint32_t ndt_hash(uint32_t key) { key *= 2654435761u; return key ^ (key >> 15); }
LB_HD int ndt_hash_find(const NdtTargetView& tv, uint32_t key) {
  uint32_t h = ndt_hash(key) & tv.hmask;
  for (;;) {
    uint32_t k = tv.hkey[h];
    if (k == key) return tv.hval[h];
    if (k == 0xffffffffu) return -1;
    h = (h + 1) & tv.hmask;
  }
}

// KDTREE method, step 1: the lattice range that can hold a voxel whose float centroid lies within one voxel side of q
// (margin: float rounding of the centroid and of this arithmetic).  lo / hi are relative to min_b, clamped to the grid.
LB_HD void ndt_kd_range(const NdtTargetView& tv, const float* q, int* lo, int* hi) {
  const float r = tv.leaf;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float m = 1e-3f * r + 1e-5f * fabsf(q[a]);
    lo[a] = (int)floorf((q[a] - r - m) * tv.inv_leaf) - tv.min_b[a];
    hi[a] = (int)floorf((q[a] + r + m) * tv.inv_leaf) - tv.min_b[a];
    if (lo[a] < 0) lo[a] = 0;
    if (hi[a] > tv.div_b[a] - 1) hi[a] = tv.div_b[a] - 1;
  }
}
// step 2: one lattice cell -> its searchable voxel's slot and float d2 when the centroid is within the radius (strict)
LB_HD bool ndt_kd_probe(const NdtTargetView& tv, int cx, int cy, int cz, float qx, float qy, float qz, int& s, float& d2) {
  const uint32_t key = (uint32_t)(cx + cy * tv.div_b[0] + cz * tv.div_b[0] * tv.div_b[1]);
  s = ndt_hash_find(tv, key);
  if (s < 0) return false;
  const f4 c = tv.cen[s];
  const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
  d2 = (dx * dx + dy * dy) + dz * dz;
  return d2 < tv.r2;
}
// Relative cells of the DIRECT methods in the order the reference visits them.  DIRECT7 / DIRECT1: own cell, +x, -x, +y, -y,
// +z, -z (voxel_grid_covariance_omp_impl.hpp:413-440).  DIRECT26: pcl::getAllNeighborCellIndices() -- the 13 "half" offsets
// ((i, j, -1) for i, j in -1..1; (i, -1, 0) for i in -1..1; (-1, 0, 0)) followed by their negatives; the own cell is NOT among them.
LB_HD int ndt_direct_count(int method) { return method == NDT_DIRECT1 ? 1 : (method == NDT_DIRECT7 ? 7 : 26); }
LB_HD void ndt_direct_rel(int method, int r, int* d) {
  if (method != NDT_DIRECT26) {
    const int REL[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    d[0] = REL[r][0]; d[1] = REL[r][1]; d[2] = REL[r][2];
    return;
  }
  const int h = r % 13, sgn = r < 13 ? 1 : -1;
  int i, j, k;
  if (h < 9) { i = h / 3 - 1; j = h % 3 - 1; k = -1; }
  else if (h < 12) { i = h - 10; j = -1; k = 0; }
  else { i = -1; j = 0; k = 0; }
  d[0] = sgn * i; d[1] = sgn * j; d[2] = sgn * k;
}
// the r-th relative cell -> slot of a usable voxel, or -1
LB_HD int ndt_direct_probe(const NdtTargetView& tv, int r, float qx, float qy, float qz) {
  int rel[3];
  ndt_direct_rel(tv.method, r, rel);
  const int ijk[3] = {(int)floorf(qx / tv.leaf), (int)floorf(qy / tv.leaf), (int)floorf(qz / tv.leaf)};
  int idx = 0, mul = 1;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int c = ijk[a] + rel[a];
    if (c < tv.min_b[a] || c > tv.max_b[a]) return -1;
    idx += (c - tv.min_b[a]) * mul;
    mul *= tv.div_b[a];
  }
  const int s = ndt_hash_find(tv, (uint32_t)idx);
  if (s < 0) return -1;
  return float_to_bits(tv.cen[s].w) >= tv.min_pts ? s : -1;      // an invalidated voxel carries nr_points = -1
}

// Neighbourhood of the transformed point q, in the order the reference visits it.  Returns the count; slot[] = index
// into vox / cen.  KDTREE: every voxel whose float centroid lies within d2 < r2, ascending (d2, slot).
LB_HD int ndt_neighbours(const NdtTargetView& tv, float qx, float qy, float qz, int* slot) {
  int k = 0;
  if (tv.method == NDT_KDTREE) {
    float d2s[NDT_MAX_NB];
    const float q[3] = {qx, qy, qz};
    int lo[3], hi[3];
    ndt_kd_range(tv, q, lo, hi);
    for (int cz = lo[2]; cz <= hi[2]; cz++)
      for (int cy = lo[1]; cy <= hi[1]; cy++)
        for (int cx = lo[0]; cx <= hi[0]; cx++) {
          int s;
          float d2;
          if (!ndt_kd_probe(tv, cx, cy, cz, qx, qy, qz, s, d2) || k >= NDT_MAX_NB) continue;
          int j = k++;                                      // insertion by (d2, slot)
          while (j > 0 && (d2s[j - 1] > d2 || (d2s[j - 1] == d2 && slot[j - 1] > s))) { d2s[j] = d2s[j - 1]; slot[j] = slot[j - 1]; j--; }
          d2s[j] = d2; slot[j] = s;
        }
    return k;
  }
  const int nrel = ndt_direct_count(tv.method);
  for (int r = 0; r < nrel; r++) {
    const int s = ndt_direct_probe(tv, r, qx, qy, qz);
    if (s >= 0) slot[k++] = s;
  }
  return k;
}

// ------------------------------------------------------------------ angle derivatives (Magnusson 2009, eq. 6.19 / 6.21)
LB_HD void ndt_angles(const double* p, NdtAngles& A) {
  double cx, cy, cz, sx, sy, sz;
  if (fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = cos(p[3]); sx = sin(p[3]); }
  if (fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = cos(p[4]); sy = sin(p[4]); }
  if (fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = cos(p[5]); sz = sin(p[5]); }
#define LB_ROW(M, r, a, b, c) M[r][0] = (a); M[r][1] = (b); M[r][2] = (c)
  LB_ROW(A.jd, 0, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
  LB_ROW(A.jd, 1, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
  LB_ROW(A.jd, 2, (-sy * cz), sy * sz, cy);
  LB_ROW(A.jd, 3, sx * cy * cz, (-sx * cy * sz), sx * sy);
  LB_ROW(A.jd, 4, (-cx * cy * cz), cx * cy * sz, (-cx * sy));
  LB_ROW(A.jd, 5, (-cy * sz), (-cy * cz), 0.0);
  LB_ROW(A.jd, 6, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0.0);
  LB_ROW(A.jd, 7, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0.0);
  LB_ROW(A.hd, 0, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
  LB_ROW(A.hd, 1, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
  LB_ROW(A.hd, 2, (cx * cy * cz), (-cx * cy * sz), (cx * sy));
  LB_ROW(A.hd, 3, (sx * cy * cz), (-sx * cy * sz), (sx * sy));
  LB_ROW(A.hd, 4, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0.0);
  LB_ROW(A.hd, 5, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0.0);
  LB_ROW(A.hd, 6, (-cy * cz), (cy * sz), (sy));
  LB_ROW(A.hd, 7, (-sx * sy * cz), (sx * sy * sz), (sx * cy));
  LB_ROW(A.hd, 8, (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
  LB_ROW(A.hd, 9, (sy * sz), (sy * cz), 0.0);
  LB_ROW(A.hd, 10, (-sx * cy * sz), (-sx * cy * cz), 0.0);
  LB_ROW(A.hd, 11, (cx * cy * sz), (cx * cy * cz), 0.0);
  LB_ROW(A.hd, 12, (-cy * cz), (cy * sz), 0.0);
  LB_ROW(A.hd, 13, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0.0);
  LB_ROW(A.hd, 14, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0.0);
#undef LB_ROW
  for (int r = 0; r < 8; r++) for (int c = 0; c < 3; c++) A.jf[r][c] = (float)A.jd[r][c];
  for (int r = 0; r < 15; r++) for (int c = 0; c < 3; c++) A.hf[r][c] = (float)A.hd[r][c];
}

// which second-derivative vector (a b c d e f) sits at block (i, j), i, j in 3..5 (ndt_omp_impl.hpp:513-521)
LB_HD int ndt_hblk(int i, int j) {
  const int B[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  return B[i - 3][j - 3];
}

// e^x of a float argument.  The reference calls exp() on a float (glibc expf, correctly rounded in all but rare cases);
// a double exp rounded to float gives the same bits on the host and on the device.
LB_HD float ndt_expf(float x) { return (float)exp((double)x); }

// computePointDerivatives, float (ndt_omp_impl.hpp:478-526): the 3x6 point gradient and (hess) the six second-derivative
// vectors a..f of one source point.  Row r of an angle table times the point, then the rows placed as the reference does.
LB_HD float ndt_row_dot(const float (*M)[3], int r, float x0, float x1, float x2) { return (M[r][0] * x0 + M[r][1] * x1) + M[r][2] * x2; }
LB_HD void ndt_point_derivs_place(const float* xj /*8*/, const float* xh /*15*/, bool hess, float (*pg)[6], float (*ph)[3]) {
  pg[0][0] = 1.f; pg[0][1] = 0.f; pg[0][2] = 0.f; pg[0][3] = 0.f; pg[0][4] = xj[2]; pg[0][5] = xj[5];
  pg[1][0] = 0.f; pg[1][1] = 1.f; pg[1][2] = 0.f; pg[1][3] = xj[0]; pg[1][4] = xj[3]; pg[1][5] = xj[6];
  pg[2][0] = 0.f; pg[2][1] = 0.f; pg[2][2] = 1.f; pg[2][3] = xj[1]; pg[2][4] = xj[4]; pg[2][5] = xj[7];
  if (hess) {
    ph[0][0] = 0.f; ph[0][1] = xh[0]; ph[0][2] = xh[1];
    ph[1][0] = 0.f; ph[1][1] = xh[2]; ph[1][2] = xh[3];
    ph[2][0] = 0.f; ph[2][1] = xh[4]; ph[2][2] = xh[5];
#pragma unroll
    for (int c = 0; c < 3; c++) { ph[3][c] = xh[6 + c]; ph[4][c] = xh[9 + c]; ph[5][c] = xh[12 + c]; }
  }
}
LB_HD void ndt_point_derivs_f(const NdtAngles& A, float x0, float x1, float x2, bool hess, float (*pg)[6], float (*ph)[3]) {
  float xj[8], xh[15];
#pragma unroll
  for (int r = 0; r < 8; r++) xj[r] = ndt_row_dot(A.jf, r, x0, x1, x2);
  if (hess) {
#pragma unroll
    for (int r = 0; r < 15; r++) xh[r] = ndt_row_dot(A.hf, r, x0, x1, x2);
  }
  ndt_point_derivs_place(xj, xh, hess, pg, ph);
}

// updateDerivatives (ndt_omp_impl.hpp:574-638) for one (transformed point q, voxel) pair, float arithmetic:
// t[0] = score increment, t[1..6] = gradient terms, t[7..42] = Hessian terms (HESS only).  Returns false when the
// reference's validity check drops the pair (nothing is added then).
template <bool HESS>
LB_HD bool ndt_pair_terms_f(const NdtGauss& G, const float (*pg)[6], const float (*ph)[3], float q0, float q1, float q2, const NdtVoxel& vx,
                            float* t) {
  const float d2f = (float)G.d2;
  const float xt[3] = {(float)((double)q0 - vx.mean[0]), (float)((double)q1 - vx.mean[1]), (float)((double)q2 - vx.mean[2])};
  float cf[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) cf[r][cc] = (float)vx.icov[3 * r + cc];
  float xc[3];
#pragma unroll
  for (int cc = 0; cc < 3; cc++) xc[cc] = (xt[0] * cf[0][cc] + xt[1] * cf[1][cc]) + xt[2] * cf[2][cc];
  float e = ndt_expf(-d2f * ((xt[0] * xc[0] + xt[1] * xc[1]) + xt[2] * xc[2]) * 0.5f);
  const float score_inc = (float)(-G.d1 * (double)e);
  e = d2f * e;
  if (e > 1 || e < 0 || e != e) return false;
  e = (float)((double)e * G.d1);
  float CP[3][6], gq[6];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 6; cc++) CP[r][cc] = (cf[r][0] * pg[0][cc] + cf[r][1] * pg[1][cc]) + cf[r][2] * pg[2][cc];
#pragma unroll
  for (int cc = 0; cc < 6; cc++) gq[cc] = (xt[0] * CP[0][cc] + xt[1] * CP[1][cc]) + xt[2] * CP[2][cc];
  t[0] = score_inc;
#pragma unroll
  for (int cc = 0; cc < 6; cc++) t[1 + cc] = e * gq[cc];
  if (HESS) {
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 6; j++) {
        float xH = 0.0f;
        if (i >= 3 && j >= 3) {
          const float* v = ph[ndt_hblk(i, j)];
          xH = (xc[0] * v[0] + xc[1] * v[1]) + xc[2] * v[2];
        }
        float JCJ = (pg[0][j] * CP[0][i] + pg[1][j] * CP[1][i]) + pg[2][j] * CP[2][i];
        t[7 + 6 * i + j] = e * ((-d2f * gq[i] * gq[j] + xH) + JCJ);
      }
  }
  return true;
}

// One source point against its neighbourhood: adds to acc[43] = {score, gradient, Hessian (row-major)}.
// want = NDT_WANT_DERIV_H / NDT_WANT_DERIV: computeDerivatives' float path (Hessian only for _H);
// want = NDT_WANT_HESSIAN: computeHessian's double path (Hessian only).
LB_HD void ndt_point_eval(const NdtTargetView& tv, const NdtGauss& G, const NdtAngles& A, const float* T /*3x4*/, float x0, float x1,
                          float x2, int want, double* acc) {
  float q0, q1, q2;
  xform_pcl(T, x0, x1, x2, q0, q1, q2);
  int slot[NDT_MAX_NB];
  const int k = ndt_neighbours(tv, q0, q1, q2, slot);
  if (k == 0) return;
  if (want != NDT_WANT_HESSIAN) {
    float pg[3][6], ph[6][3], t[NDT_NSUM];
    ndt_point_derivs_f(A, x0, x1, x2, want == NDT_WANT_DERIV_H, pg, ph);
    for (int c = 0; c < k; c++) {
      const NdtVoxel& vx = tv.vox[slot[c]];
      if (want == NDT_WANT_DERIV_H) {
        if (!ndt_pair_terms_f<true>(G, pg, ph, q0, q1, q2, vx, t)) continue;
#pragma unroll
        for (int e = 0; e < NDT_NSUM; e++) acc[e] += (double)t[e];
      } else {
        if (!ndt_pair_terms_f<false>(G, pg, ph, q0, q1, q2, vx, t)) continue;
#pragma unroll
        for (int e = 0; e < 7; e++) acc[e] += (double)t[e];
      }
    }
    return;
  }
  // computeHessian / updateHessian, double (ndt_omp_impl.hpp:641-756)
  const double x[3] = {x0, x1, x2};
  double pg[3][6] = {{1, 0, 0, 0, ndt_dot3(x, A.jd[2]), ndt_dot3(x, A.jd[5])},
                     {0, 1, 0, ndt_dot3(x, A.jd[0]), ndt_dot3(x, A.jd[3]), ndt_dot3(x, A.jd[6])},
                     {0, 0, 1, ndt_dot3(x, A.jd[1]), ndt_dot3(x, A.jd[4]), ndt_dot3(x, A.jd[7])}};
  double ph[6][3];
  ph[0][0] = 0; ph[0][1] = ndt_dot3(x, A.hd[0]); ph[0][2] = ndt_dot3(x, A.hd[1]);
  ph[1][0] = 0; ph[1][1] = ndt_dot3(x, A.hd[2]); ph[1][2] = ndt_dot3(x, A.hd[3]);
  ph[2][0] = 0; ph[2][1] = ndt_dot3(x, A.hd[4]); ph[2][2] = ndt_dot3(x, A.hd[5]);
#pragma unroll
  for (int c = 0; c < 3; c++) { ph[3][c] = ndt_dot3(x, A.hd[6 + c]); ph[4][c] = ndt_dot3(x, A.hd[9 + c]); ph[5][c] = ndt_dot3(x, A.hd[12 + c]); }
  for (int c = 0; c < k; c++) {
    const NdtVoxel& vx = tv.vox[slot[c]];
    const double xt[3] = {(double)q0 - vx.mean[0], (double)q1 - vx.mean[1], (double)q2 - vx.mean[2]};
    double cx[3];
    ndt_mv3(vx.icov, xt, cx);
    double e = G.d2 * exp(-G.d2 * ndt_dot3(xt, cx) / 2);
    if (e > 1 || e < 0 || e != e) continue;
    e *= G.d1;
    double cc[6][3], xcc[6];      // c_inv * point_gradient.col(j) and x_trans . that
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const double cj[3] = {pg[0][j], pg[1][j], pg[2][j]};
      ndt_mv3(vx.icov, cj, cc[j]);
      xcc[j] = ndt_dot3(xt, cc[j]);
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double xch = 0.0;
        if (i >= 3 && j >= 3) {
          double chv[3];
          ndt_mv3(vx.icov, ph[ndt_hblk(i, j)], chv);
          xch = ndt_dot3(xt, chv);
        } else {
          const double z[3] = {0, 0, 0};
          double chv[3];
          ndt_mv3(vx.icov, z, chv);          // c_inv * 0: keeps the sign-of-zero / NaN behaviour of the reference's expression
          xch = ndt_dot3(xt, chv);
        }
        const double cj[3] = {pg[0][j], pg[1][j], pg[2][j]};
        acc[7 + 6 * i + j] += e * ((-G.d2 * xcc[i] * xcc[j] + xch) + ndt_dot3(cj, cc[i]));
      }
  }
}

// ------------------------------------------------------------------ pose <-> matrix
LB_HD float ndt_sinf(float a) { return (float)sin((double)a); }
LB_HD float ndt_cosf(float a) { return (float)cos((double)a); }

// Eigen AngleAxisf(angle, unit axis).toRotationMatrix(), row-major 3x3
LB_HD void ndt_axis_rotation(float angle, int axis, float* R) {
  float ax[3] = {0.f, 0.f, 0.f};
  ax[axis] = 1.0f;
  const float s = ndt_sinf(angle), c = ndt_cosf(angle);
  const float sa[3] = {s * ax[0], s * ax[1], s * ax[2]};
  const float ca[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float tmp;
  tmp = ca[0] * ax[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
  tmp = ca[0] * ax[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
  tmp = ca[1] * ax[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
  R[0] = ca[0] * ax[0] + c; R[4] = ca[1] * ax[1] + c; R[8] = ca[2] * ax[2] + c;
}
LB_HD void ndt_mul3f(const float* a, const float* b, float* o) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
}
// (Translation3f * AngleAxisf(X) * AngleAxisf(Y) * AngleAxisf(Z)).matrix() -> row-major 3x4 (ndt_omp_impl.hpp:177-186)
LB_HD void ndt_pose_to_matrix(const double* p, float* T) {
  float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
  ndt_axis_rotation((float)p[3], 0, Rx);
  ndt_axis_rotation((float)p[4], 1, Ry);
  ndt_axis_rotation((float)p[5], 2, Rz);
  ndt_mul3f(Rx, Ry, Rxy);
  ndt_mul3f(Rxy, Rz, R);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2];
    T[4 * r + 3] = (float)p[r];
  }
}
// Matrix3f::eulerAngles(0, 1, 2) (Eigen 3.3), T row-major with row stride 4
LB_HD void ndt_euler_xyz(const float* T, float* out) {
#define LB_M(r, c) T[4 * (r) + (c)]
  const float pi = 3.14159265358979323846f;
  float r0 = (float)atan2((double)LB_M(1, 2), (double)LB_M(2, 2));
  float c2 = sqrtf(LB_M(0, 0) * LB_M(0, 0) + LB_M(0, 1) * LB_M(0, 1));
  float r1;
  if (r0 > 0.f) {
    r0 -= pi;
    r1 = (float)atan2((double)-LB_M(0, 2), (double)-c2);
  } else {
    r1 = (float)atan2((double)-LB_M(0, 2), (double)c2);
  }
  const float s1 = ndt_sinf(r0), c1 = ndt_cosf(r0);
  float r2 = (float)atan2((double)(s1 * LB_M(2, 0) - c1 * LB_M(1, 0)), (double)(c1 * LB_M(1, 1) - s1 * LB_M(2, 1)));
  out[0] = -r0; out[1] = -r1; out[2] = -r2;
#undef LB_M
}

// ------------------------------------------------------------------ 6x6 solve through a one-sided Jacobi SVD
// x = pinv(A) b with Eigen's default rank threshold; stands in for JacobiSVD<Matrix6d>(A, FullU | FullV).solve(b).
// The column pairs of a sweep are visited in round-robin order: five rounds of three DISJOINT pairs.  Rotations of
// disjoint pairs touch disjoint columns, so a round's three rotations may run one after the other (here, and in the
// oracle) or side by side (the controller warp on the device, ndt.cu) with identical bits.
LB_HD void ndt_svd6_pair(int idx, int& p, int& q) {
  const int P[15] = {0, 1, 2, 0, 3, 1, 0, 2, 1, 0, 1, 4, 0, 2, 3};
  const int Q[15] = {5, 4, 3, 4, 5, 2, 3, 4, 5, 2, 3, 5, 1, 5, 4};
  p = P[idx]; q = Q[idx];
}
// One rotation: columns p, q of U (and V) from alpha = |u_p|^2, beta = |u_q|^2, gamma = u_p . u_q.  Returns false when
// the columns already are orthogonal to working precision (no rotation).
LB_HD bool ndt_svd6_angle(double alpha, double beta, double gamma, double& c, double& s) {
  if (gamma == 0.0 || fabs(gamma) <= DBL_EPSILON * sqrt(alpha * beta)) return false;
  const double zeta = (beta - alpha) / (2.0 * gamma);
  double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  if (zeta < 0.0) t = -t;
  c = 1.0 / sqrt(1.0 + t * t);
  s = c * t;
  return true;
}
// singular values from the rotated columns, Eigen's rank threshold, x = V S^+ U' b
LB_HD void ndt_svd6_finish(const double (*U)[6], const double (*V)[6], const double* b, double* x) {
  double sig[6];
  int ord[6];
  for (int j = 0; j < 6; j++) {
    double s2 = 0;
    for (int k = 0; k < 6; k++) s2 += U[k][j] * U[k][j];
    sig[j] = sqrt(s2); ord[j] = j;
  }
  for (int i = 1; i < 6; i++) {
    int o = ord[i], j = i;
    while (j > 0 && sig[ord[j - 1]] < sig[o]) { ord[j] = ord[j - 1]; j--; }
    ord[j] = o;
  }
  double thr = sig[ord[0]] * (6.0 * DBL_EPSILON);
  if (thr < DBL_MIN) thr = DBL_MIN;
  for (int r = 0; r < 6; r++) x[r] = 0.0;
  for (int jj = 0; jj < 6; jj++) {
    const int j = ord[jj];
    if (!(sig[j] > thr)) break;
    double ub = 0;
    for (int k = 0; k < 6; k++) ub += (U[k][j] / sig[j]) * b[k];
    const double w = ub / sig[j];
    for (int r = 0; r < 6; r++) x[r] += V[r][j] * w;
  }
}
LB_HD void ndt_svd6_solve(const double* A, const double* b, double* x) {
  double U[6][6], V[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { U[i][j] = A[6 * i + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (int e = 0; e < 15; e++) {
      int p, q;
      ndt_svd6_pair(e, p, q);
      double alpha = 0, beta = 0, gamma = 0;
      for (int k = 0; k < 6; k++) { alpha += U[k][p] * U[k][p]; beta += U[k][q] * U[k][q]; gamma += U[k][p] * U[k][q]; }
      double c, s;
      if (!ndt_svd6_angle(alpha, beta, gamma, c, s)) continue;
      rotated = 1;
      for (int k = 0; k < 6; k++) {
        double up = U[k][p], uq = U[k][q];
        U[k][p] = c * up - s * uq; U[k][q] = s * up + c * uq;
        double vp = V[k][p], vq = V[k][q];
        V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  ndt_svd6_finish(U, V, b, x);
}
// ------------------------------------------------------------------ More-Thuente helpers (ndt_omp_impl.hpp:758-885)
LB_HD bool ndt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t,
                               double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}
LB_HD double ndt_cubic(double a_l, double f_l, double g_l, double a_t, double f_t, double g_t) {
  double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
  double w = sqrt(z * z - g_t * g_l);
  return a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
}
LB_HD double ndt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double a_c = ndt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return fabs(a_c - a_l) < fabs(a_q - a_l) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    double a_c = ndt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return fabs(a_c - a_t) >= fabs(a_s - a_t) ? a_c : a_s;
  }
  if (fabs(g_t) <= fabs(g_l)) {
    double a_c = ndt_cubic(a_l, f_l, g_l, a_t, f_t, g_t);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_n = fabs(a_c - a_t) < fabs(a_s - a_t) ? a_c : a_s;
    double lim = a_t + 0.66 * (a_u - a_t);
    if (a_t > a_l) return lim < a_n ? lim : a_n;
    return lim > a_n ? lim : a_n;
  }
  return ndt_cubic(a_u, f_u, g_u, a_t, f_t, g_t);
}

// ------------------------------------------------------------------ the controller
struct NdtCtl {
  // parameters
  double step_size, tf_eps;
  int max_iterations;
  // request to the evaluator
  int want;
  float T[12];              // transform of the source for this evaluation
  NdtAngles ang;
  // state
  int phase;                // 0 initial derivatives, 1 first trial of a line search, 2 later trials, 3 final Hessian, 9 done
  double p[6], dir[6], x_t[6];
  double score, g[6], H[36];
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t, step_min, step_max;
  int step_iterations, open_interval, interval_converged;
  int nr_iterations, converged, n_evals;
  float final_T[12];        // final_transformation_ (top three rows)
};

LB_HD double ndt_dot6(const double* a, const double* b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}

LB_HD void ndt_post_eval(NdtCtl& c, int want) {
  for (int i = 0; i < 6; i++) c.x_t[i] = c.p[i] + c.dir[i] * c.a_t;
  ndt_pose_to_matrix(c.x_t, c.T);
  for (int i = 0; i < 12; i++) c.final_T[i] = c.T[i];
  ndt_angles(c.x_t, c.ang);
  c.want = want;
}

// align(output, guess): final_transformation_ = guess, first derivatives on guess * source (ndt_omp_impl.hpp:100-141)
LB_HD void ndt_ctl_begin(NdtCtl& c, const float* guess16 /*row-major 4x4*/, double step_size, double tf_eps, int max_iterations) {
  c.step_size = step_size; c.tf_eps = tf_eps; c.max_iterations = max_iterations;
  for (int i = 0; i < 12; i++) { c.T[i] = guess16[i]; c.final_T[i] = guess16[i]; }
  float e[3];
  ndt_euler_xyz(c.T, e);
  c.p[0] = c.T[3]; c.p[1] = c.T[7]; c.p[2] = c.T[11];
  c.p[3] = e[0]; c.p[4] = e[1]; c.p[5] = e[2];
  ndt_angles(c.p, c.ang);
  c.phase = 0; c.want = NDT_WANT_DERIV_H;
  c.nr_iterations = 0; c.converged = 0; c.n_evals = 0;
  c.score = 0;
  for (int i = 0; i < 6; i++) { c.g[i] = 0; c.dir[i] = 0; c.x_t[i] = c.p[i]; }
  for (int i = 0; i < 36; i++) c.H[i] = 0;
  c.a_t = 0; c.step_iterations = 0; c.open_interval = 1; c.interval_converged = 0;
}

// Consumes the sums of the evaluation that was requested (sums[43]) and moves on to the next request, or finishes
// (want = NDT_WANT_NONE, phase = 9).  The Newton direction is computed OUTSIDE: the function returns true when it needs
// delta_p = pinv(H) (-g) for the c.H / c.g it holds, and is called again with it (sums are ignored then) -- on the host
// the serial solve above answers, on the device the controller thread hands the solve to its whole warp (ndt.cu).
// Returns false once a request is posted or the align has finished.
LB_HD bool ndt_ctl_run(NdtCtl& c, const double* sums, const double* delta_p) {
  const double mu = 1.e-4, nu = 0.9;
  enum { NEWTON, LS_CHECK, LS_END, POST_LS } at;
  if (delta_p) {
    at = NEWTON;
  } else if (c.phase == 3) {
    for (int i = 0; i < 36; i++) c.H[i] = sums[7 + i];
    at = POST_LS;
  } else {
    c.n_evals++;                      // computeDerivatives calls; the closing Hessian pass is not one
    c.score = sums[0];
    for (int i = 0; i < 6; i++) c.g[i] = sums[1 + i];
    for (int i = 0; i < 36; i++) c.H[i] = sums[7 + i];
    if (c.phase == 0) {
      at = NEWTON;
    } else {
      c.phi_t = -c.score; c.d_phi_t = -ndt_dot6(c.g, c.dir);
      c.psi_t = c.phi_t - c.phi_0 - mu * c.d_phi_0 * c.a_t;
      c.d_psi_t = c.d_phi_t - mu * c.d_phi_0;
      if (c.phase == 2) {
        if (c.open_interval && (c.psi_t <= 0 && c.d_psi_t >= 0)) {
          c.open_interval = 0;
          c.f_l = c.f_l + c.phi_0 - mu * c.d_phi_0 * c.a_l; c.g_l = c.g_l + mu * c.d_phi_0;
          c.f_u = c.f_u + c.phi_0 - mu * c.d_phi_0 * c.a_u; c.g_u = c.g_u + mu * c.d_phi_0;
        }
        if (c.open_interval) c.interval_converged = ndt_update_interval(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.psi_t, c.d_psi_t);
        else c.interval_converged = ndt_update_interval(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.phi_t, c.d_phi_t);
        c.step_iterations++;
      }
      at = LS_CHECK;
    }
  }
  for (;;) {
    if (at == NEWTON) {
      if (!delta_p) return true;
      const double* dp = delta_p;
      delta_p = nullptr;                 // consumed: a second Newton step within this call asks again
      double nrm = sqrt(ndt_dot6(dp, dp));
      if (nrm == 0 || nrm != nrm) { c.converged = nrm == nrm; c.want = NDT_WANT_NONE; c.phase = 9; return false; }
      for (int i = 0; i < 6; i++) c.dir[i] = dp[i] / nrm;
      // computeStepLengthMT(p, dir, nrm, step_size, tf_eps / 2, ...)
      c.step_max = c.step_size; c.step_min = c.tf_eps / 2;
      c.phi_0 = -c.score;
      c.d_phi_0 = -ndt_dot6(c.g, c.dir);
      c.step_iterations = 0;
      if (c.d_phi_0 >= 0) {
        if (c.d_phi_0 == 0) { c.a_t = 0; at = POST_LS; continue; }
        c.d_phi_0 *= -1;
        for (int i = 0; i < 6; i++) c.dir[i] *= -1;
      }
      c.a_l = 0; c.a_u = 0;
      c.f_l = c.phi_0 - c.phi_0 - mu * c.d_phi_0 * c.a_l; c.g_l = c.d_phi_0 - mu * c.d_phi_0;
      c.f_u = c.phi_0 - c.phi_0 - mu * c.d_phi_0 * c.a_u; c.g_u = c.d_phi_0 - mu * c.d_phi_0;
      c.interval_converged = (c.step_max - c.step_min) < 0;
      c.open_interval = 1;
      c.a_t = nrm;
      c.a_t = c.a_t < c.step_max ? c.a_t : c.step_max;
      c.a_t = c.a_t > c.step_min ? c.a_t : c.step_min;
      ndt_post_eval(c, NDT_WANT_DERIV_H);
      c.phase = 1;
      return false;
    }
    if (at == LS_CHECK) {
      if (!c.interval_converged && c.step_iterations < 10 && !(c.psi_t <= 0 && c.d_phi_t <= -nu * c.d_phi_0)) {
        if (c.open_interval) c.a_t = ndt_trial_value(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.psi_t, c.d_psi_t);
        else c.a_t = ndt_trial_value(c.a_l, c.f_l, c.g_l, c.a_u, c.f_u, c.g_u, c.a_t, c.phi_t, c.d_phi_t);
        c.a_t = c.a_t < c.step_max ? c.a_t : c.step_max;
        c.a_t = c.a_t > c.step_min ? c.a_t : c.step_min;
        ndt_post_eval(c, NDT_WANT_DERIV);
        c.phase = 2;
        return false;
      }
      at = LS_END;
    }
    if (at == LS_END) {
      if (c.step_iterations) { c.want = NDT_WANT_HESSIAN; c.phase = 3; return false; }   // same transform and angles as the last trial
      at = POST_LS;
    }
    if (at == POST_LS) {
      const double delta_p_norm = c.a_t;
      for (int i = 0; i < 6; i++) c.p[i] = c.p[i] + c.dir[i] * delta_p_norm;
      if (c.nr_iterations > c.max_iterations || (c.nr_iterations && (fabs(delta_p_norm) < c.tf_eps))) c.converged = 1;
      c.nr_iterations++;
      if (c.converged) { c.want = NDT_WANT_NONE; c.phase = 9; return false; }
      at = NEWTON;
    }
  }
}

LB_HD void ndt_ctl_advance(NdtCtl& c, const double* sums) {
  double dp[6];
  const double* have = nullptr;
  while (ndt_ctl_run(c, sums, have)) {
    double ng[6];
    for (int i = 0; i < 6; i++) ng[i] = -c.g[i];
    ndt_svd6_solve(c.H, ng, dp);
    have = dp;
  }
}

}  // namespace lb
