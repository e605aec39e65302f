// hd_harness.cpp -- TEST-ONLY CPU build of the product's __host__ __device__
// per-point logic (locus_b200/csrc/hd.h, grid.h, bfgs.h).
//
// The product has no CPU compute path.  This file exists so that the scalar
// code the CUDA kernels execute per thread (ring searches, covariance
// regularisation, Mahalanobis matrices, objective terms, the BFGS / outer-loop
// state machine) can be checked against the oracle in the authoring container,
// which has no GPU.  It is compiled by tests/conftest.py into
// tests/_build/libhd_harness.so and never shipped or linked into
// liblocus_b200.so.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../locus_b200/csrc/bfgs.h"
#include "../locus_b200/csrc/grid.h"
#include "../locus_b200/csrc/hd.h"

using namespace lb;

struct HGrid {
  std::vector<f4> pts;
  std::vector<uint32_t> cell_start;
  GridView v;
};

extern "C" {
// Row f4: the product's BodyFilter predicate (hd.h body_box_drops), set up like lb_voxel_set_body_filter does.
void hh_body_box(const float* xyz, int n, const float* mn, const float* mx, float rot, unsigned char* dropped) {
  BodyBox b;
  b.enabled = 1;
  float A = cosf(rot), B = sinf(rot), det = A * A + B * B;
  b.ia = A / det; b.ib = B / det;
  for (int d = 0; d < 3; d++) { b.mn[d] = mn[d]; b.mx[d] = mx[d]; }
  for (int i = 0; i < n; i++) dropped[i] = body_box_drops(b, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]) ? 1 : 0;
}

// CPU model of knn_cov_quadreg_kernel's candidate handling: the candidates of one query are dealt round-robin to
// four RegLists (the quad's lanes); every `refresh` candidates the shared gate is tightened to the maximum of the
// lanes' (K/4)-th best keys (quad_row_done); at the end the four lists are merged.  The merged top-k must equal
// the plain sorted top-k whatever the refresh cadence -- i.e. the gate never rejects a true neighbour.
int hh_quad_gate_model(const float* d2, const int* orig, int n, int k, int refresh, int* out_orig) {
  RegList<20> L[4];
  for (int l = 0; l < 4; l++) L[l].init();
  for (int i = 0; i < n; i++) {
    L[i & 3].push(d2[i], orig[i]);
    if (refresh > 0 && (i % refresh) == refresh - 1) {
      unsigned long long t = 0;
      for (int l = 0; l < 4; l++) t = L[l].key[20 / 4 - 1] > t ? L[l].key[20 / 4 - 1] : t;
      for (int l = 0; l < 4; l++) L[l].gate = t;
    }
  }
  int p[4] = {0, 0, 0, 0}, found = 0;
  for (int round = 0; round < k; round++) {
    int best = -1;
    unsigned long long bk = ~0ull;
    for (int l = 0; l < 4; l++)
      if (p[l] < 20 && L[l].key[p[l]] < bk) { bk = L[l].key[p[l]]; best = l; }
    if (best < 0) break;
    out_orig[found++] = (int)(uint32_t)bk;
    p[best]++;
  }
  return found;
}

// RegList (the register-resident candidate list of knn_cov_quadreg_kernel): after n pushes of (d2, orig) it must
// hold the K best in ascending (d2, orig) order.  gate_d2 >= 0: candidates not better than (gate_d2, gate_orig)
// are rejected on top of that.  out_*: K entries (-1 when unfilled).
void hh_reglist_topk(const float* d2, const int* orig, int n, float gate_d2, int gate_orig, int* out_orig, float* out_d2, int* out_cnt) {
  RegList<20> L;
  L.init();
  if (gate_d2 >= 0.f) L.gate = RegList<20>::make_key(gate_d2, gate_orig);
  for (int i = 0; i < n; i++) L.push(d2[i], orig[i]);
  for (int j = 0; j < 20; j++) {
    bool filled = L.key[j] != ~0ull;
    out_orig[j] = filled ? (int)(uint32_t)L.key[j] : -1;
    out_d2[j] = filled ? bits_to_float((int32_t)(L.key[j] >> 32)) : -1.0f;
  }
  *out_cnt = L.cnt;
}


// CPU stand-in for the GPU index build (min/max, keys, stable sort, CSR).
void* hh_grid_build(const float* xyz, int n, int stride_f, float h) {
  HGrid* g = new HGrid;
  float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      float c = xyz[(size_t)i * stride_f + d];
      mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c);
    }
  GridView& v = g->v;
  v.ox = mn[0]; v.oy = mn[1]; v.oz = mn[2];
  v.h = h; v.inv_h = 1.0f / h;
  v.nx = (int)floorf((mx[0] - mn[0]) * v.inv_h) + 1;
  v.ny = (int)floorf((mx[1] - mn[1]) * v.inv_h) + 1;
  v.nz = (int)floorf((mx[2] - mn[2]) * v.inv_h) + 1;
  v.n = n;
  size_t nc = (size_t)v.nx * v.ny * v.nz;
  std::vector<uint32_t> key(n);
  for (int i = 0; i < n; i++) {
    const float* p = &xyz[(size_t)i * stride_f];
    int cx = std::min(std::max((int)floorf((p[0] - v.ox) * v.inv_h), 0), v.nx - 1);
    int cy = std::min(std::max((int)floorf((p[1] - v.oy) * v.inv_h), 0), v.ny - 1);
    int cz = std::min(std::max((int)floorf((p[2] - v.oz) * v.inv_h), 0), v.nz - 1);
    key[i] = (uint32_t)((cz * v.ny + cy) * v.nx + cx);
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  g->pts.resize(n);
  g->cell_start.assign(nc + 1, 0);
  for (int i = 0; i < n; i++) g->cell_start[key[i] + 1]++;
  for (size_t c = 0; c < nc; c++) g->cell_start[c + 1] += g->cell_start[c];
  for (int s = 0; s < n; s++) {
    int i = order[s];
    const float* p = &xyz[(size_t)i * stride_f];
    g->pts[s] = f4{p[0], p[1], p[2], bits_to_float(i)};
  }
  v.pts = g->pts.data();
  v.cell_start = g->cell_start.data();
  return g;
}
void hh_grid_free(void* g) { delete (HGrid*)g; }
void hh_grid_dims(void* g, int* dims) { HGrid* G = (HGrid*)g; dims[0] = G->v.nx; dims[1] = G->v.ny; dims[2] = G->v.nz; }

void hh_nn1_batch(void* g, const float* q, int nq, int stride_f, float max_d2, int* idx, float* d2) {
  HGrid* G = (HGrid*)g;
  for (int i = 0; i < nq; i++) {
    int bo; float bd;
    int s = nn1(G->v, q[(size_t)i * stride_f], q[(size_t)i * stride_f + 1], q[(size_t)i * stride_f + 2], max_d2, bo, bd);
    idx[i] = (s >= 0) ? bo : -1;
    d2[i] = bd;
  }
}

void hh_nn1_block_batch(void* g, const float* q, int nq, int stride_f, float max_d2, int* idx, float* d2) {
  HGrid* G = (HGrid*)g;
  for (int i = 0; i < nq; i++) {
    int bo; float bd;
    int s = nn1_pruned(G->v, q[(size_t)i * stride_f], q[(size_t)i * stride_f + 1], q[(size_t)i * stride_f + 2], max_d2, bo, bd);
    idx[i] = (s >= 0) ? bo : -1;
    d2[i] = bd;
  }
}

// the serial restatement of the staged search (grid.h nn1_ball_serial): ub2 (nullable) = per-query squared distance to a
// known target point (the bound the previous outer iteration's match gives), r0cut = first-look radius in cells
void hh_nn1_ball_batch(void* g, const float* q, int nq, int stride_f, float max_d2, const float* ub2, float r0cut, int* idx, float* d2) {
  HGrid* G = (HGrid*)g;
  for (int i = 0; i < nq; i++) {
    int bo; float bd;
    int s = nn1_ball_serial(G->v, q[(size_t)i * stride_f], q[(size_t)i * stride_f + 1], q[(size_t)i * stride_f + 2], max_d2,
                            ub2 != nullptr, ub2 ? ub2[i] : 0.f, r0cut, bo, bd);
    idx[i] = (s >= 0) ? bo : -1;
    d2[i] = bd;
  }
}

void hh_knn_batch(void* g, const float* q, int nq, int stride_f, int k, int* idx, float* d2) {
  HGrid* G = (HGrid*)g;
  for (int i = 0; i < nq; i++) {
    KnnList<32> L;
    int c = knn<32>(G->v, q[(size_t)i * stride_f], q[(size_t)i * stride_f + 1], q[(size_t)i * stride_f + 2], k, L);
    for (int j = 0; j < k; j++) { idx[(size_t)i * k + j] = j < c ? L.oi[j] : -1; d2[(size_t)i * k + j] = L.d2[j]; }
  }
}

// k-NN covariances exactly as the cov kernel does per thread; out: n x 6 (sym) in ORIGINAL order
// Row f2: the product's NormalAccum / pcl_normal_from_accum (hd.h) driven by the product's exact k-NN (grid.h),
// original point order.  Same libm as the oracle on this machine -> bit-identical to og_normals_knn.
void hh_normals_knn(void* g, int k, const float* vp, float* out4) {
  HGrid* G = (HGrid*)g;
  for (int s = 0; s < G->v.n; s++) {
    f4 q = G->pts[s];
    KnnList<32> L;
    int c = knn<32>(G->v, q.x, q.y, q.z, k, L);
    NormalAccum acc;
    acc.reset();
    for (int j = 0; j < c; j++) {
      f4 p = G->pts[L.si[j]];
      acc.add(p.x, p.y, p.z);
    }
    pcl_normal_from_accum(acc, c, q.x, q.y, q.z, vp, &out4[4 * (size_t)float_to_bits(q.w)]);
  }
}

void hh_cov_knn(void* g, int k, double eps, double* out6) {
  HGrid* G = (HGrid*)g;
  for (int s = 0; s < G->v.n; s++) {
    f4 q = G->pts[s];
    KnnList<32> L;
    int c = knn<32>(G->v, q.x, q.y, q.z, k, L);
    double sum[3] = {0, 0, 0}, m2[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < c; j++) {
      f4 p = G->pts[L.si[j]];
      sum[0] += p.x; sum[1] += p.y; sum[2] += p.z;
      m2[0] += p.x * p.x; m2[1] += p.y * p.x; m2[2] += p.y * p.y;
      m2[3] += p.z * p.x; m2[4] += p.z * p.y; m2[5] += p.z * p.z;
    }
    cov_from_moments(sum, m2, k, eps, &out6[6 * (size_t)float_to_bits(q.w)]);
  }
}

void hh_mahalanobis(const double* R, const double* C1, const double* C2, double* M) { mahalanobis(R, C1, C2, M); }
void hh_apply_state(const double* x, float* T12) { apply_state(x, T12); }

// ---- full align with a serial CPU backend (control-flow check of bfgs.h) ----
struct CpuBackend {
  const GridView* tg;            // target index
  const double* tgt_cov;         // sorted-target order, sym6
  const f4* src;                 // source points (guess-transformed), any order
  const double* src_cov;         // same order, sym6
  int n_src;
  float max_d2;
  std::vector<f4> corr;          // matched target point per source (w = sorted idx or -1)
  std::vector<double> M;         // sym6 per source
  int m;
  int correspond(const float* T, const double* R) {
    corr.resize(n_src); M.assign((size_t)n_src * 6, 0.0);
    m = 0;
    for (int i = 0; i < n_src; i++) {
      float qx, qy, qz;
      xform(T, src[i].x, src[i].y, src[i].z, qx, qy, qz);
      int bo; float bd;
      int s = nn1_pruned(*tg, qx, qy, qz, max_d2, bo, bd);
      if (s >= 0) {
        f4 t = tg->pts[s];
        corr[i] = f4{t.x, t.y, t.z, bits_to_float(s)};
        mahalanobis(R, &src_cov[6 * (size_t)i], &tgt_cov[6 * (size_t)s], &M[6 * (size_t)i]);
        m++;
      } else {
        corr[i] = f4{0, 0, 0, bits_to_float(-1)};
      }
    }
    return m;
  }
  void fdf(const double* x, double* f, double* g) {
    float T[12];
    apply_state(x, T);
    double acc[13] = {0};
    for (int i = 0; i < n_src; i++) {
      if (float_to_bits(corr[i].w) < 0) continue;
      objective_terms(T, src[i].x, src[i].y, src[i].z, corr[i].x, corr[i].y, corr[i].z, &M[6 * (size_t)i], acc);
    }
    objective_finish(acc, m, x, f, g);
  }
  int gn(const double* x, double* f, double* b, double* H) {
    float T[12];
    apply_state(x, T);
    double dP[9], dT[9], dS[9];
    r_derivatives(x, dP, dT, dS);
    double acc[28] = {0};
    for (int i = 0; i < n_src; i++) {
      if (float_to_bits(corr[i].w) < 0) continue;
      gn_terms(T, dP, dT, dS, src[i].x, src[i].y, src[i].z, corr[i].x, corr[i].y, corr[i].z, &M[6 * (size_t)i], acc);
    }
    *f = acc[0] / m;
    for (int a = 0; a < 6; a++) b[a] = acc[1 + a];
    for (int e = 0; e < 21; e++) H[e] = acc[7 + e];
    return 0;
  }
};

// Moment-form backend (LB_EXEC_MOMENTS): correspond() also reduces the 74 moments (serially, in source order);
// fdf / gn evaluate them -- the scalar code the solve kernel's leader warp runs.
struct CpuMomentBackend : CpuBackend {
  MomentObjective mo;
  int correspond(const float* T, const double* R) {
    int mm = CpuBackend::correspond(T, R);
    for (int e = 0; e < MOM_N; e++) mo.mom[e] = 0.0;
    for (int e = 0; e < 12; e++) mo.A0[e] = T[e];
    for (int i = 0; i < n_src; i++) {
      if (float_to_bits(corr[i].w) < 0) continue;
      moment_terms(T, src[i].x, src[i].y, src[i].z, corr[i].x, corr[i].y, corr[i].z, &M[6 * (size_t)i], mo.mom);
    }
    mo.m = mm;
    return mm;
  }
  void fdf(const double* x, double* f, double* g) { mo.fdf(x, f, g); }
  int gn(const double* x, double* f, double* b, double* H) { return mo.gn(x, f, b, H); }
};

static int hh_align_impl(const float* src_xyz, int n_src, const float* tgt_xyz, int n_tgt, float h_src, float h_tgt,
             int k, double eps, double rot_eps, double tf_eps, double corr_dist, int max_it, int max_inner,
             int optimizer, const float* guess, float* T_out, int* info, double* delta_out, bool moments);

// src/tgt: n x 3 float.  returns final T (16), iterations etc.  Covariances by k-NN.
int hh_align(const float* src_xyz, int n_src, const float* tgt_xyz, int n_tgt, float h_src, float h_tgt,
             int k, double eps, double rot_eps, double tf_eps, double corr_dist, int max_it, int max_inner,
             int optimizer, const float* guess, float* T_out, int* info, double* delta_out) {
  return hh_align_impl(src_xyz, n_src, tgt_xyz, n_tgt, h_src, h_tgt, k, eps, rot_eps, tf_eps, corr_dist, max_it, max_inner,
                       optimizer, guess, T_out, info, delta_out, false);
}
int hh_align_moments(const float* src_xyz, int n_src, const float* tgt_xyz, int n_tgt, float h_src, float h_tgt,
             int k, double eps, double rot_eps, double tf_eps, double corr_dist, int max_it, int max_inner,
             int optimizer, const float* guess, float* T_out, int* info, double* delta_out) {
  return hh_align_impl(src_xyz, n_src, tgt_xyz, n_tgt, h_src, h_tgt, k, eps, rot_eps, tf_eps, corr_dist, max_it, max_inner,
                       optimizer, guess, T_out, info, delta_out, true);
}

// objective + gradient at state x: exact pass over the pairs vs the moment form (moments taken about A0 = T(x0))
void hh_moment_fdf(const float* src4, const float* tgt4, const double* M6, int m, const double* x0, const double* x,
                   double* f_exact, double* g_exact, double* f_mom, double* g_mom) {
  float T[12], A0[12];
  apply_state(x, T);
  double acc[13] = {0};
  for (int i = 0; i < m; i++)
    objective_terms(T, src4[4 * i], src4[4 * i + 1], src4[4 * i + 2], tgt4[4 * i], tgt4[4 * i + 1], tgt4[4 * i + 2], &M6[6 * (size_t)i], acc);
  objective_finish(acc, m, x, f_exact, g_exact);
  MomentObjective mo;
  apply_state(x0, A0);
  for (int e = 0; e < MOM_N; e++) mo.mom[e] = 0.0;
  for (int e = 0; e < 12; e++) mo.A0[e] = A0[e];
  for (int i = 0; i < m; i++)
    moment_terms(A0, src4[4 * i], src4[4 * i + 1], src4[4 * i + 2], tgt4[4 * i], tgt4[4 * i + 1], tgt4[4 * i + 2], &M6[6 * (size_t)i], mo.mom);
  mo.m = m;
  mo.fdf(x, f_mom, g_mom);
}

static int hh_align_impl(const float* src_xyz, int n_src, const float* tgt_xyz, int n_tgt, float h_src, float h_tgt,
             int k, double eps, double rot_eps, double tf_eps, double corr_dist, int max_it, int max_inner,
             int optimizer, const float* guess, float* T_out, int* info /*iters, converged, n_corr, evals, inner*/,
             double* delta_out, bool moments) {
  HGrid* gs = (HGrid*)hh_grid_build(src_xyz, n_src, 3, h_src);
  HGrid* gt = (HGrid*)hh_grid_build(tgt_xyz, n_tgt, 3, h_tgt);
  std::vector<double> cs((size_t)n_src * 6), ct_orig((size_t)n_tgt * 6), ct((size_t)n_tgt * 6);
  hh_cov_knn(gs, k, eps, cs.data());       // original order
  hh_cov_knn(gt, k, eps, ct_orig.data());
  for (int s = 0; s < n_tgt; s++) {
    int oi = float_to_bits(gt->pts[s].w);
    for (int e = 0; e < 6; e++) ct[6 * (size_t)s + e] = ct_orig[6 * (size_t)oi + e];
  }
  float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const float* G = guess ? guess : I16;
  std::vector<f4> src(n_src);
  for (int i = 0; i < n_src; i++) {
    float x, y, z;
    xform_pcl(G, src_xyz[3 * i], src_xyz[3 * i + 1], src_xyz[3 * i + 2], x, y, z);
    src[i] = f4{x, y, z, 1.0f};
  }
  CpuMomentBackend be;
  be.tg = &gt->v; be.tgt_cov = ct.data(); be.src = src.data(); be.src_cov = cs.data(); be.n_src = n_src;
  be.max_d2 = (float)(corr_dist * corr_dist);
  OuterParams P{rot_eps, tf_eps, max_it, max_inner, optimizer};
  OuterResult R;
  if (moments) gicp_outer_loop(be, P, G, R);
  else gicp_outer_loop(static_cast<CpuBackend&>(be), P, G, R);
  memcpy(T_out, R.final_T, sizeof(float) * 16);
  info[0] = R.nr_iterations; info[1] = R.converged; info[2] = R.n_corr; info[3] = R.st.n_evals; info[4] = R.st.n_inner;
  *delta_out = R.delta;
  hh_grid_free(gs); hh_grid_free(gt);
  return 0;
}

}  // extern "C"
