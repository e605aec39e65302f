// ndt_harness.cpp -- TEST-ONLY CPU build of the product's __host__ __device__ NDT logic (locus_b200/csrc/ndt.h).
//
// The product has no CPU compute path.  This file runs the per-voxel, per-point and controller code the CUDA kernels
// of ndt.cu execute, with a serial backend (points in order, sums in point order -- the reference's own order), so it
// can be compared with oracle/ndt_oracle.c in a container without a GPU.  Compiled by tests/conftest.py into
// tests/_build/libndt_harness.so; never linked into liblocus_b200.so.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../locus_b200/csrc/ndt.h"

using namespace lb;

struct HNdt {
  std::vector<NdtVoxel> vox;
  std::vector<f4> cen;
  std::vector<int> leaf_idx;
  std::vector<uint32_t> hkey;
  std::vector<int32_t> hval;
  NdtTargetView tv;
  NdtLattice L;
  NdtGauss G;
  int status = 0;
};

extern "C" {

void* hn_target_build(const float* pts, int n, int stride_f, float resolution, int min_pts, double eig_mult, int method,
                      double outlier_ratio) {
  HNdt* h = new HNdt;
  ndt_gauss_constants(outlier_ratio, resolution, h->G);
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  std::vector<std::pair<int, int>> ks;
  for (int i = 0; i < n; i++) {
    const float* p = pts + (size_t)i * stride_f;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
    ks.push_back({0, i});
  }
  if (ks.empty()) { h->status = -1; return h; }
  if (!ndt_lattice(mn, mx, resolution, h->L)) { h->status = -2; return h; }
  for (auto& k : ks) { const float* p = pts + (size_t)k.second * stride_f; k.first = ndt_voxel_key(h->L, p[0], p[1], p[2]); }
  std::sort(ks.begin(), ks.end());                      // (key, input index): what the device's stable radix sort yields
  for (size_t i = 0; i < ks.size();) {
    size_t j = i;
    double sum[3] = {0, 0, 0}, m2[6] = {0, 0, 0, 0, 0, 0};
    float csum[3] = {0, 0, 0};
    for (; j < ks.size() && ks[j].first == ks[i].first; j++) {
      const float* p = pts + (size_t)ks[j].second * stride_f;
      const double d[3] = {p[0], p[1], p[2]};
      for (int a = 0; a < 3; a++) { sum[a] += d[a]; csum[a] += p[a]; }
      m2[0] += d[0] * d[0]; m2[1] += d[0] * d[1]; m2[2] += d[0] * d[2]; m2[3] += d[1] * d[1]; m2[4] += d[1] * d[2]; m2[5] += d[2] * d[2];
    }
    const int cnt = (int)(j - i);
    if (cnt >= min_pts) {
      NdtVoxel v; float c[3];
      int nr = ndt_finish_voxel(cnt, sum, m2, csum, eig_mult, v, c);
      h->vox.push_back(v);
      h->cen.push_back(f4{c[0], c[1], c[2], bits_to_float(nr)});
      h->leaf_idx.push_back(ks[i].first);
    }
    i = j;
  }
  const int nv = (int)h->vox.size();
  uint32_t cap = 16;
  while (cap < 2u * (uint32_t)nv) cap <<= 1;
  h->hkey.assign(cap, 0xffffffffu); h->hval.assign(cap, -1);
  for (int s = 0; s < nv; s++) {
    uint32_t key = (uint32_t)h->leaf_idx[s], hh = ndt_hash(key) & (cap - 1);
    while (h->hkey[hh] != 0xffffffffu) hh = (hh + 1) & (cap - 1);
    h->hkey[hh] = key; h->hval[hh] = s;
  }
  NdtTargetView& tv = h->tv;
  tv.vox = h->vox.data(); tv.cen = h->cen.data(); tv.hkey = h->hkey.data(); tv.hval = h->hval.data(); tv.hmask = cap - 1;
  tv.n_valid = nv;
  for (int a = 0; a < 3; a++) { tv.min_b[a] = h->L.min_b[a]; tv.max_b[a] = h->L.max_b[a]; tv.div_b[a] = h->L.div_b[a]; }
  tv.leaf = resolution; tv.inv_leaf = h->L.inv_leaf;
  const double radius = (double)resolution;
  tv.r2 = (float)(radius * radius);
  tv.method = method; tv.min_pts = min_pts;
  return h;
}

void hn_target_free(void* hp) { delete (HNdt*)hp; }

int hn_target_info(void* hp, int* n_valid, int* min_b, int* div_b) {
  HNdt* h = (HNdt*)hp;
  *n_valid = (int)h->vox.size();
  for (int a = 0; a < 3; a++) { min_b[a] = h->L.min_b[a]; div_b[a] = h->L.div_b[a]; }
  return h->status;
}

void hn_target_leaves(void* hp, int* leaf_idx, int* nr, double* mean, double* icov, float* centroid) {
  HNdt* h = (HNdt*)hp;
  for (size_t v = 0; v < h->vox.size(); v++) {
    leaf_idx[v] = h->leaf_idx[v];
    nr[v] = float_to_bits(h->cen[v].w);
    memcpy(mean + 3 * v, h->vox[v].mean, 24); memcpy(icov + 9 * v, h->vox[v].icov, 72);
    centroid[3 * v] = h->cen[v].x; centroid[3 * v + 1] = h->cen[v].y; centroid[3 * v + 2] = h->cen[v].z;
  }
}

// ndt_neighbours for a batch of (already transformed) query points: counts[i] and slots[i * 32 ..]
void hn_neighbours(void* hp, const float* q, int n, int* counts, int* slots) {
  HNdt* h = (HNdt*)hp;
  for (int i = 0; i < n; i++) counts[i] = ndt_neighbours(h->tv, q[3 * i], q[3 * i + 1], q[3 * i + 2], slots + (size_t)i * NDT_MAX_NB);
}

// The Newton solve as the controller warp runs it (ndt.cu, ndt_svd6_solve_warp), emulated serially: the three column pairs of a
// round are rotated from the SAME pre-round state (all alpha / beta / gamma and angles first, then all updates), which is what
// lanes working side by side do.  Must equal ndt_svd6_solve bit for bit because the pairs of a round are disjoint.
void hn_svd6_rounds(const double* A, const double* b, double* x) {
  double U[6][6], V[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) { U[i][j] = A[6 * i + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; sweep++) {
    bool any = false;
    for (int r = 0; r < 5; r++) {
      int p[3], q[3]; double c[3], s[3]; bool rot[3];
      for (int g = 0; g < 3; g++) {
        ndt_svd6_pair(3 * r + g, p[g], q[g]);
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 6; k++) { alpha += U[k][p[g]] * U[k][p[g]]; beta += U[k][q[g]] * U[k][q[g]]; gamma += U[k][p[g]] * U[k][q[g]]; }
        rot[g] = ndt_svd6_angle(alpha, beta, gamma, c[g], s[g]);
      }
      for (int g = 0; g < 3; g++) {
        if (!rot[g]) continue;
        any = true;
        for (int k = 0; k < 6; k++) {
          double up = U[k][p[g]], uq = U[k][q[g]], vp = V[k][p[g]], vq = V[k][q[g]];
          U[k][p[g]] = c[g] * up - s[g] * uq; U[k][q[g]] = s[g] * up + c[g] * uq;
          V[k][p[g]] = c[g] * vp - s[g] * vq; V[k][q[g]] = s[g] * vp + c[g] * vq;
        }
      }
    }
    if (!any) break;
  }
  ndt_svd6_finish(U, V, b, x);
}
void hn_svd6_serial(const double* A, const double* b, double* x) { ndt_svd6_solve(A, b, x); }

static void eval_serial(HNdt* h, const float* src, int n, int stride_f, const float* T, const NdtAngles& A, int want, double* sums) {
  for (int k = 0; k < NDT_NSUM; k++) sums[k] = 0;
  for (int i = 0; i < n; i++) {
    const float* p = src + (size_t)i * stride_f;
    if (want == NDT_WANT_HESSIAN) {
      // computeHessian adds every (point, voxel) pair straight into the 6x6 (ndt_omp_impl.hpp:658-716)
      ndt_point_eval(h->tv, h->G, A, T, p[0], p[1], p[2], want, sums);
      continue;
    }
    double acc[NDT_NSUM];                     // computeDerivatives: per-point subtotals, summed in point order (:316-340)
    for (int k = 0; k < NDT_NSUM; k++) acc[k] = 0;
    ndt_point_eval(h->tv, h->G, A, T, p[0], p[1], p[2], want, acc);
    for (int k = 0; k < NDT_NSUM; k++) sums[k] += acc[k];
  }
}

// The float passes as ndt_eval_group_kernel organises them (ndt.cu), emulated serially: candidate cells probed "eight at a
// time" in lane order (so hits arrive in an order unrelated to their distance), ranked by (key, slot), pairs computed in
// rounds of eight, and each sum accumulated over the pairs in rank order; per-point subtotals added in point order.
// Must equal eval_serial bit for bit: it is the same arithmetic in the same association.
void hn_eval_grouped(void* hp, const float* src, int n, int stride_f, const float* T, const double* p6, int want, double* sums) {
  HNdt* h = (HNdt*)hp;
  const NdtTargetView& tv = h->tv;
  NdtAngles A;
  ndt_angles(p6, A);
  const bool hess = want == NDT_WANT_DERIV_H;
  const int NT = hess ? NDT_NSUM : 7;
  for (int k = 0; k < NDT_NSUM; k++) sums[k] = 0;
  for (int i = 0; i < n; i++) {
    const float* x = src + (size_t)i * stride_f;
    float q0, q1, q2;
    xform_pcl(T, x[0], x[1], x[2], q0, q1, q2);
    int hit_slot[64]; float hit_key[64]; int nh = 0;
    if (tv.method == NDT_KDTREE) {
      const float q[3] = {q0, q1, q2};
      int lo[3], hi[3];
      ndt_kd_range(tv, q, lo, hi);
      const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
      if (nx > 0 && ny > 0 && nz > 0) {
        const int ncell = nx * ny * nz, nxy = nx * ny;
        const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)nx;
        for (int sub = 7; sub >= 0; sub--)                         // lanes in some order
          for (int c = sub; c < ncell; c += 8) {
            const int iz = (int)(((float)c + 0.5f) * inv_nxy), rem = c - iz * nxy;
            const int iy = (int)(((float)rem + 0.5f) * inv_nx);
            if (iz != c / nxy || iy != rem / nx) { sums[0] = NAN; return; }    // the float-reciprocal decode must be exact
            int sl; float d2;
            if (ndt_kd_probe(tv, lo[0] + (rem - iy * nx), lo[1] + iy, lo[2] + iz, q0, q1, q2, sl, d2)) { hit_slot[nh] = sl; hit_key[nh] = d2; nh++; }
          }
      }
    } else {
      const int nrel = ndt_direct_count(tv.method);
      for (int sub = 7; sub >= 0; sub--)
        for (int r = sub; r < nrel; r += 8) {
          const int sl = ndt_direct_probe(tv, r, q0, q1, q2);
          if (sl >= 0) { hit_slot[nh] = sl; hit_key[nh] = (float)r; nh++; }
        }
    }
    int sorted[64];
    for (int a = 0; a < nh; a++) {
      int rank = 0;
      for (int o = 0; o < nh; o++) rank += (hit_key[o] < hit_key[a] || (hit_key[o] == hit_key[a] && hit_slot[o] < hit_slot[a])) ? 1 : 0;
      sorted[rank] = hit_slot[a];
    }
    double acc[NDT_NSUM];
    for (int k = 0; k < NDT_NSUM; k++) acc[k] = 0;
    if (nh > 0) {
      float xj[8], xh[16], pg[3][6], ph[6][3];
      for (int r = 0; r < 8; r++) xj[r] = ndt_row_dot(A.jf, r, x[0], x[1], x[2]);
      if (hess) for (int r = 0; r < 15; r++) xh[r] = ndt_row_dot(A.hf, r, x[0], x[1], x[2]);
      ndt_point_derivs_place(xj, xh, hess, pg, ph);
      for (int base = 0; base < nh; base += 8) {
        float term[8][NDT_NSUM];
        const int m = std::min(8, nh - base);
        for (int sub = 0; sub < m; sub++) {
          float t[NDT_NSUM];
          const bool ok = hess ? ndt_pair_terms_f<true>(h->G, pg, ph, q0, q1, q2, tv.vox[sorted[base + sub]], t)
                               : ndt_pair_terms_f<false>(h->G, pg, ph, q0, q1, q2, tv.vox[sorted[base + sub]], t);
          for (int e = 0; e < NT; e++) term[sub][e] = ok ? t[e] : 0.0f;
        }
        for (int c = 0; c < NT; c++)
          for (int jj = 0; jj < m; jj++) acc[c] += (double)term[jj][c];
      }
    }
    for (int k = 0; k < NDT_NSUM; k++) sums[k] += acc[k];
  }
}

void hn_eval(void* hp, const float* src, int n, int stride_f, const float* T16, const double* p6, int want, double* sums43) {
  HNdt* h = (HNdt*)hp;
  NdtAngles A;
  ndt_angles(p6, A);
  eval_serial(h, src, n, stride_f, T16, A, want, sums43);
}

int hn_align(void* hp, const float* src, int n, int stride_f, const float* guess16, double step_size, double tf_eps, int max_iterations,
             float* final16, int* converged, int* iterations, int* evaluations, double* pose, double* trans_probability) {
  HNdt* h = (HNdt*)hp;
  static const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  NdtCtl c;
  ndt_ctl_begin(c, guess16 ? guess16 : I16, step_size, tf_eps, max_iterations);
  double sums[NDT_NSUM];
  while (c.want != NDT_WANT_NONE) {
    eval_serial(h, src, n, stride_f, c.T, c.ang, c.want, sums);
    ndt_ctl_advance(c, sums);
  }
  for (int i = 0; i < 12; i++) final16[i] = c.final_T[i];
  final16[12] = final16[13] = final16[14] = 0.f; final16[15] = 1.f;
  *converged = c.converged; *iterations = c.nr_iterations; *evaluations = c.n_evals;
  for (int i = 0; i < 6; i++) pose[i] = c.p[i];
  *trans_probability = c.score / (double)n;
  return 0;
}

}  // extern "C"
