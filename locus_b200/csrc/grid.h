// grid.h -- dense voxel-hash index over a point cloud and exact ring searches.
//
// Replaces the FLANN KD-tree the reference reaches through
// pcl::search::KdTree (gicp.hpp:108 k-NN for covariances, gicp.h:385-391 1-NN
// correspondences, PointCloudLocalization.cc:327-336 post-align 1-NN).
//
// Layout in HBM (built once per cloud by gicp.cu):
//   pts[n]        float4, sorted by cell (x fastest, then y, then z); .w holds the
//                 ORIGINAL index of the point (int bits).
//   cell_start[ncells+1]  uint32 CSR offsets into pts.
// Because x is the fastest-varying cell coordinate, the cells (cx-r..cx+r, y, z)
// of one row are ONE contiguous run of float4 in memory: a 3x3x3 probe is nine
// contiguous runs, which is what makes the loads coalesce / hit L2 sectors.
//
// Exactness: a candidate is better iff (d2, original index) is lexicographically
// smaller; d2 = ((dx*dx)+(dy*dy))+(dz*dz) in float32 without FMA (FLANN
// L2_Simple<float>).  A search stops after ring r only when the current worst
// kept d2 is strictly below the squared distance from the query to the border
// of the scanned block, shrunk by a safety margin that covers float rounding of
// the cell assignment -- so results equal an exhaustive scan (tests/test_hd_grid).
#pragma once

#include "hd.h"

namespace lb {

struct GridView {
  const f4* pts;               // sorted by cell; w = original index bits
  const uint32_t* cell_start;  // ncells + 1
  float ox, oy, oz;            // origin (min corner)
  float inv_h, h;              // cell size
  int nx, ny, nz;
  int n;                       // number of points
};

LB_HD float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  float d = dx * dx;
  d = d + dy * dy;
  d = d + dz * dz;
  return d;
}

LB_HD bool better(float d2a, int ia, float d2b, int ib) { return (d2a < d2b) || (d2a == d2b && ia < ib); }

LB_HD int iabs_(int a) { return a < 0 ? -a : a; }
LB_HD int imax_(int a, int b) { return a > b ? a : b; }
LB_HD int imin_(int a, int b) { return a < b ? a : b; }

// cell coordinate of a query (may lie outside the grid); frac = distance (in cells)
// from the query to the nearest face of its cell.
LB_HD void query_cell(const GridView& g, float qx, float qy, float qz, int& cx, int& cy, int& cz, float& minfrac) {
  float ux = (qx - g.ox) * g.inv_h, uy = (qy - g.oy) * g.inv_h, uz = (qz - g.oz) * g.inv_h;
  // clamp to a sane range so the float->int conversion is defined for far-away queries
  const float LIM = 1.0e9f;
  ux = fminf(fmaxf(ux, -LIM), LIM); uy = fminf(fmaxf(uy, -LIM), LIM); uz = fminf(fmaxf(uz, -LIM), LIM);
  float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  cx = (int)fx; cy = (int)fy; cz = (int)fz;
  float ax = ux - fx, ay = uy - fy, az = uz - fz;
  float m = fminf(fminf(fminf(ax, 1.0f - ax), fminf(ay, 1.0f - ay)), fminf(az, 1.0f - az));
  minfrac = m;
}

// squared lower bound on the distance from the query to any point outside the
// scanned block of Chebyshev radius r (conservative: 0.01 cell safety margin).
LB_HD float ring_bound2(const GridView& g, int r, float minfrac) {
  float b = ((float)r + minfrac - 0.01f) * g.h;
  if (b <= 0.0f) return 0.0f;
  b = b * 0.9999f;
  return b * b;
}

// Visit every point in the shell of Chebyshev radius r around (cx,cy,cz).
// F(float x, float y, float z, int orig_index, int sorted_index)
template <class F>
LB_HD void visit_shell(const GridView& g, int cx, int cy, int cz, int r, F&& f) {
  int z0 = imax_(cz - r, 0), z1 = imin_(cz + r, g.nz - 1);
  int y0 = imax_(cy - r, 0), y1 = imin_(cy + r, g.ny - 1);
  for (int z = z0; z <= z1; z++) {
    bool zface = (iabs_(z - cz) == r);
    for (int y = y0; y <= y1; y++) {
      bool face = zface || (iabs_(y - cy) == r);
      int base = (z * g.ny + y) * g.nx;
      if (face) {
        int xa = imax_(cx - r, 0), xb = imin_(cx + r, g.nx - 1);
        if (xa > xb) continue;
        uint32_t s = g.cell_start[base + xa], e = g.cell_start[base + xb + 1];
        for (uint32_t i = s; i < e; i++) {
          f4 p = g.pts[i];
          f(p.x, p.y, p.z, float_to_bits(p.w), (int)i);
        }
      } else {
        // interior row of the shell: only the two end cells x = cx-r and x = cx+r
        int xs[2] = {cx - r, cx + r};
        for (int k = 0; k < 2; k++) {
          int x = xs[k];
          if (x < 0 || x >= g.nx) continue;
          uint32_t s = g.cell_start[base + x], e = g.cell_start[base + x + 1];
          for (uint32_t i = s; i < e; i++) {
            f4 p = g.pts[i];
            f(p.x, p.y, p.z, float_to_bits(p.w), (int)i);
          }
        }
      }
    }
  }
}

// first ring that can contain grid cells, and the ring after which the whole grid is covered
LB_HD void ring_range(const GridView& g, int cx, int cy, int cz, int& r_first, int& r_last) {
  int fx = imax_(imax_(-cx, cx - (g.nx - 1)), 0);
  int fy = imax_(imax_(-cy, cy - (g.ny - 1)), 0);
  int fz = imax_(imax_(-cz, cz - (g.nz - 1)), 0);
  r_first = imax_(fx, imax_(fy, fz));
  int lx = imax_(iabs_(cx), iabs_(g.nx - 1 - cx));
  int ly = imax_(iabs_(cy), iabs_(g.ny - 1 - cy));
  int lz = imax_(iabs_(cz), iabs_(g.nz - 1 - cz));
  r_last = imax_(lx, imax_(ly, lz));
}

// Exact nearest neighbour with a strict gate d2 < max_d2 (gicp.hpp:483).
// Returns the SORTED index of the neighbour (or -1); orig index and d2 by reference.
LB_HD int nn1(const GridView& g, float qx, float qy, float qz, float max_d2, int& best_orig, float& best_d2) {
  int cx, cy, cz; float minfrac;
  query_cell(g, qx, qy, qz, cx, cy, cz, minfrac);
  int r0, r1;
  ring_range(g, cx, cy, cz, r0, r1);
  float bd2 = max_d2; int bi = -1; int bs = -1;
  bool found = false;
  for (int r = r0; r <= r1; r++) {
    if (r > r0 || r0 > 0) {
      // everything not yet scanned is at least this far away
      float lb2 = ring_bound2(g, r - 1, minfrac);
      if (lb2 >= max_d2) break;
      if (found && bd2 < lb2) break;
    }
    visit_shell(g, cx, cy, cz, r, [&](float x, float y, float z, int oi, int si) {
      float d = dist2(qx, qy, qz, x, y, z);
      if (!found) {
        if (d < max_d2) { found = true; bd2 = d; bi = oi; bs = si; }
      } else if (better(d, oi, bd2, bi)) {
        bd2 = d; bi = oi; bs = si;
      }
    });
  }
  best_orig = bi; best_d2 = bd2;
  return bs;
}

// Exact k nearest neighbours (unbounded radius, like FLANN nearestKSearch).
// Kept ascending by (d2, orig index) in d2s/idx/sidx (arrays of length >= K).
// Returns the number found (min(K, n)).
template <int KMAX>
struct KnnList {
  float d2[KMAX];
  int oi[KMAX];   // original index
  int si[KMAX];   // sorted index
  int k, cnt;
  LB_HD void init(int k_) {
    k = k_; cnt = 0;
#pragma unroll
    for (int i = 0; i < KMAX; i++) { d2[i] = 3.0e38f; oi[i] = 0x7fffffff; si[i] = -1; }
  }
  LB_HD float worst() const { return d2[k - 1]; }
  LB_HD void push(float d, int o, int s) {
    // list is padded with (+huge, INT_MAX): a plain "better than last" test works while filling
    if (!better(d, o, d2[k - 1], oi[k - 1])) return;
    if (cnt < k) cnt++;
    // insert by bubbling up from the last slot (static indices when unrolled)
    d2[k - 1] = d; oi[k - 1] = o; si[k - 1] = s;
#pragma unroll
    for (int j = KMAX - 1; j > 0; j--) {
      if (j <= k - 1 && better(d2[j], oi[j], d2[j - 1], oi[j - 1])) {
        float td = d2[j]; d2[j] = d2[j - 1]; d2[j - 1] = td;
        int to = oi[j]; oi[j] = oi[j - 1]; oi[j - 1] = to;
        int ts = si[j]; si[j] = si[j - 1]; si[j - 1] = ts;
      }
    }
  }
};

// Sorted candidate list with STATIC indexing only, so that a fully unrolled build keeps it in registers (a list
// indexed by a loop-carried position lives in local memory and every insertion becomes a chain of dependent
// loads and stores).  Order = better(): ascending (d2, original index), packed into one 64-bit key -- d2 >= 0, so
// its float bits are monotone as an unsigned integer.  Unfilled entries hold the all-ones sentinel.  Only the key
// is kept: the point itself is fetched again through its original index when it is needed.
// `gate`: an extra rejection threshold the caller may tighten (see knn_cov_quadreg_kernel); candidates with
// key >= gate are known not to belong to the result.
template <int K>
struct RegList {
  unsigned long long key[K];
  unsigned long long gate;
  int cnt;
  LB_HD static unsigned long long make_key(float d2, int orig) {
    return ((unsigned long long)(uint32_t)float_to_bits(d2) << 32) | (uint32_t)orig;
  }
  LB_HD void init() {
    cnt = 0; gate = ~0ull;
#pragma unroll
    for (int i = 0; i < K; i++) key[i] = ~0ull;
  }
  LB_HD void push(float d2, int orig, int /*sorted position: not kept*/ = 0) {
    const unsigned long long k = make_key(d2, orig);
    if (k >= key[K - 1] || k >= gate) return;
    cnt += (cnt < K) ? 1 : 0;
    bool here = true;                       // k < (old) key[j]
#pragma unroll
    for (int j = K - 1; j > 0; j--) {
      const bool up = k < key[j - 1];       // the new entry belongs below j: slot j takes its lower neighbour
      key[j] = up ? key[j - 1] : (here ? k : key[j]);
      here = up;
    }
    if (here) key[0] = k;
  }
};

template <int KMAX>
LB_HD int knn(const GridView& g, float qx, float qy, float qz, int k, KnnList<KMAX>& L) {
  L.init(k);
  int cx, cy, cz; float minfrac;
  query_cell(g, qx, qy, qz, cx, cy, cz, minfrac);
  int r0, r1;
  ring_range(g, cx, cy, cz, r0, r1);
  for (int r = r0; r <= r1; r++) {
    if (L.cnt == k && r > 0) {
      float lb2 = ring_bound2(g, r - 1, minfrac);
      if (L.worst() < lb2) break;
    }
    visit_shell(g, cx, cy, cz, r, [&](float x, float y, float z, int oi, int si) {
      L.push(dist2(qx, qy, qz, x, y, z), oi, si);
    });
  }
  return L.cnt;
}

}  // namespace lb
