/*
 * kdtree.c -- exact k-NN kd-tree for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Stands in for pcl::search::KdTree -> pcl::KdTreeFLANN -> flann::KDTreeSingleIndex
 * (L2_Simple<float>, leaf size 15, exact search, results sorted ascending),
 * which the reference uses at gicp.hpp:108 (k-NN for covariances) and
 * gicp.h:385-391 (1-NN correspondences).  FLANN is not in the reference tree;
 * this is a restatement of "exact k-NN under float32 squared L2", with one
 * documented choice: equidistant neighbours are ordered by ascending index
 * (FLANN's tie order is an implementation detail -- SURVEY H4).
 */
#include "lb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LEAF_SIZE 15

typedef struct {
  int left, right;   /* children (node ids) or -1 for leaf */
  int lo, hi;        /* range in perm[] for leaves */
  int dim;
  float split_lo, split_hi; /* max of left side / min of right side along dim */
} kd_node;

struct og_kdtree {
  const float* pts;
  int n, stride;
  int* perm;
  kd_node* nodes;
  int n_nodes, cap_nodes;
  float* xyz; /* packed copy in perm order for locality: 3 floats per point */
};

static int new_node(og_kdtree* t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (kd_node*)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
  }
  return t->n_nodes++;
}

static inline float coord(const og_kdtree* t, int i, int d) {
  return t->pts[(size_t)i * t->stride + d];
}

/* nth_element on perm[lo..hi) by coordinate d (ties by index for determinism) */
static inline int less_pt(const og_kdtree* t, int a, int b, int d) {
  float ca = coord(t, a, d), cb = coord(t, b, d);
  return (ca < cb) || (ca == cb && a < b);
}

static void nth_element(og_kdtree* t, int lo, int hi, int nth, int d) {
  int* p = t->perm;
  while (hi - lo > 1) {
    /* median of three pivot */
    int mid = lo + (hi - lo) / 2;
    int a = p[lo], b = p[mid], c = p[hi - 1];
    int piv;
    if (less_pt(t, a, b, d)) {
      if (less_pt(t, b, c, d)) piv = b; else piv = less_pt(t, a, c, d) ? c : a;
    } else {
      if (less_pt(t, a, c, d)) piv = a; else piv = less_pt(t, b, c, d) ? c : b;
    }
    int i = lo, j = hi - 1;
    while (i <= j) {
      while (less_pt(t, p[i], piv, d)) i++;
      while (less_pt(t, piv, p[j], d)) j--;
      if (i <= j) { int tmp = p[i]; p[i] = p[j]; p[j] = tmp; i++; j--; }
    }
    if (nth <= j) hi = j + 1;
    else if (nth >= i) lo = i;
    else return;
  }
}

static int build_rec(og_kdtree* t, int lo, int hi) {
  int id = new_node(t);
  kd_node nd;
  nd.left = nd.right = -1; nd.lo = lo; nd.hi = hi; nd.dim = 0; nd.split_lo = nd.split_hi = 0.f;
  if (hi - lo <= LEAF_SIZE) { t->nodes[id] = nd; return id; }
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = lo; i < hi; i++)
    for (int d = 0; d < 3; d++) {
      float c = coord(t, t->perm[i], d);
      if (c < mn[d]) mn[d] = c;
      if (c > mx[d]) mx[d] = c;
    }
  int dim = 0;
  float ext = mx[0] - mn[0];
  for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > ext) { ext = mx[d] - mn[d]; dim = d; }
  if (!(ext > 0.f)) { t->nodes[id] = nd; return id; } /* all identical: big leaf */
  int mid = lo + (hi - lo) / 2;
  nth_element(t, lo, hi, mid, dim);
  nd.dim = dim;
  float slo = -FLT_MAX, shi = FLT_MAX;
  for (int i = lo; i < mid; i++) { float c = coord(t, t->perm[i], dim); if (c > slo) slo = c; }
  for (int i = mid; i < hi; i++) { float c = coord(t, t->perm[i], dim); if (c < shi) shi = c; }
  nd.split_lo = slo; nd.split_hi = shi;
  t->nodes[id] = nd;
  int l = build_rec(t, lo, mid);
  int r = build_rec(t, mid, hi);
  t->nodes[id].left = l;
  t->nodes[id].right = r;
  return id;
}

og_kdtree* og_kdtree_build(const float* pts, int n, int stride_f) {
  og_kdtree* t = (og_kdtree*)calloc(1, sizeof(og_kdtree));
  t->pts = pts; t->n = n; t->stride = stride_f;
  t->perm = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) t->perm[i] = i;
  if (n > 0) build_rec(t, 0, n);
  t->xyz = (float*)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) t->xyz[3 * (size_t)i + d] = coord(t, t->perm[i], d);
  return t;
}

void og_kdtree_free(og_kdtree* t) {
  if (!t) return;
  free(t->perm); free(t->nodes); free(t->xyz); free(t);
}

typedef struct {
  int k, cnt;
  int* idx;
  float* d2;
} knn_heap; /* kept as a sorted array ascending by (d2, idx): k <= ~32 */

static inline int better(float d2a, int ia, float d2b, int ib) {
  return (d2a < d2b) || (d2a == d2b && ia < ib);
}

static inline void knn_push(knn_heap* h, float d2, int idx) {
  if (h->cnt == h->k) {
    if (!better(d2, idx, h->d2[h->k - 1], h->idx[h->k - 1])) return;
  } else {
    h->cnt++;
  }
  int j = h->cnt - 1;
  while (j > 0 && better(d2, idx, h->d2[j - 1], h->idx[j - 1])) {
    h->d2[j] = h->d2[j - 1]; h->idx[j] = h->idx[j - 1]; j--;
  }
  h->d2[j] = d2; h->idx[j] = idx;
}

static void search_rec(const og_kdtree* t, int id, const float q[3], knn_heap* h) {
  const kd_node* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; i++) {
      const float* p = &t->xyz[3 * (size_t)i];
      /* FLANN L2_Simple<float>: result += diff*diff in float, x then y then z */
      float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
      float d2 = dx * dx;
      d2 = d2 + dy * dy;
      d2 = d2 + dz * dz;
      knn_push(h, d2, t->perm[i]);
    }
    return;
  }
  float qd = q[nd->dim];
  /* distance (along dim) to each child's slab; 0 if inside */
  float dl = qd - nd->split_lo; if (dl < 0.f) dl = 0.f;   /* to reach left side need qd <= split_lo */
  float dr = nd->split_hi - qd; if (dr < 0.f) dr = 0.f;
  int first, second; float dsecond;
  if (dl <= dr) { first = nd->left; second = nd->right; dsecond = dr; }
  else { first = nd->right; second = nd->left; dsecond = dl; }
  float dfirst = (first == nd->left) ? dl : dr;
  if (h->cnt < h->k || dfirst * dfirst <= h->d2[h->k - 1]) search_rec(t, first, q, h);
  /* '<=' so that equidistant lower-index points on the far side are still found */
  if (h->cnt < h->k || dsecond * dsecond <= h->d2[h->k - 1]) search_rec(t, second, q, h);
}

int og_kdtree_knn(const og_kdtree* t, const float q[3], int k, int* idx, float* d2) {
  knn_heap h; h.k = k; h.cnt = 0; h.idx = idx; h.d2 = d2;
  if (t->n == 0 || k <= 0) return 0;
  search_rec(t, 0, q, &h);
  return h.cnt;
}

/* radius search (pcl::KdTreeFLANN::radiusSearch -> flann radiusSearch with RadiusResultSet: every point with
 * d2 < radius2, strictly; sorted ascending by (d2, index) like the k-NN results).  Returns the number found; writes at
 * most cap of them (the nearest ones are kept when the list is longer, callers size cap = n). */
typedef struct { float r2; int cnt, cap; int* idx; float* d2; } rad_set;

static void radius_rec(const og_kdtree* t, int id, const float q[3], rad_set* h) {
  const kd_node* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; i++) {
      const float* p = &t->xyz[3 * (size_t)i];
      float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
      float d2 = dx * dx;
      d2 = d2 + dy * dy;
      d2 = d2 + dz * dz;
      if (d2 < h->r2) {
        if (h->cnt < h->cap) { h->idx[h->cnt] = t->perm[i]; h->d2[h->cnt] = d2; }
        h->cnt++;
      }
    }
    return;
  }
  float qd = q[nd->dim];
  float dl = qd - nd->split_lo; if (dl < 0.f) dl = 0.f;
  float dr = nd->split_hi - qd; if (dr < 0.f) dr = 0.f;
  if (dl * dl <= h->r2) radius_rec(t, nd->left, q, h);
  if (dr * dr <= h->r2) radius_rec(t, nd->right, q, h);
}

int og_kdtree_radius(const og_kdtree* t, const float q[3], float radius2, int* idx, float* d2, int cap) {
  rad_set h; h.r2 = radius2; h.cnt = 0; h.cap = cap; h.idx = idx; h.d2 = d2;
  if (t->n == 0) return 0;
  radius_rec(t, 0, q, &h);
  int m = h.cnt < cap ? h.cnt : cap;
  /* insertion sort by (d2, index): neighbourhoods are small */
  for (int i = 1; i < m; i++) {
    float dd = d2[i]; int ii = idx[i]; int j = i;
    while (j > 0 && better(dd, ii, d2[j - 1], idx[j - 1])) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; j--; }
    d2[j] = dd; idx[j] = ii;
  }
  return h.cnt;
}

void og_kdtree_nn_batch(const og_kdtree* t, const float* q, int nq, int stride_f,
                        int* idx, float* d2, int num_threads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(num_threads)
  for (int i = 0; i < nq; i++) {
    int ii = -1; float dd = FLT_MAX;
    int c = og_kdtree_knn(t, &q[(size_t)i * stride_f], 1, &ii, &dd);
    idx[i] = c ? ii : -1;
    d2[i] = dd;
  }
}
