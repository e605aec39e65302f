// nn_staged.cuh -- exact 1-NN in the voxel hash for the 32 queries of a warp, candidates staged through shared memory.
//
// nn1_pruned() (grid.h) is what one thread does on its own: open a cell, wait for its CSR offsets, wait for its points,
// decide about the next cell -- a chain of 15-50 DEPENDENT L2 round trips per query, and a warp is as slow as the union
// of its lanes' chains.  That chain, not bandwidth or arithmetic, was the ~90 us correspondence step of align()
// (30 k queries x ~70 candidates is nothing).  Here the chain is cut to two or three round trips per warp:
//
//   1. every lane works out which cells can hold its nearest neighbour BEFORE it loads anything:
//        * when the previous outer iteration matched the point (gicp.hpp:463-498 runs once per outer iteration, and
//          the target does not change inside align()), the distance to that previous match under the new transform is
//          an exact upper bound -- the candidates are the cells a ball of that radius touches: usually 1-4 cells;
//        * otherwise (first outer iteration, unmatched before) the ball of the gate;
//        * either ball is cut to half a cell: most nearest neighbours are much closer than a cell;
//   2. the CSR offsets of all rows of that set are fetched together (one round trip),
//   3. all candidate points are copied into the warp's shared-memory stage asynchronously -- either one 16-byte
//      cp.async per point or (TMA = true) one cp.async.bulk per contiguous row run, completion on an mbarrier --
//      (one round trip), and
//   4. every lane scans its own candidates out of shared memory.
//
// Queries the staged pass cannot decide (a cut ball whose best lies outside it, a huge ball, more candidates than the
// stage holds) are "far".  They come in clusters -- a part of the scan the target does not cover -- so a warp or a CTA
// that finished them itself would be the long pole of the step.  Instead they are appended to a device-wide queue
// (NnsFarQueue) and, after a grid-wide barrier / in the next kernel, taken one per warp by ALL warps of the grid
// (nn1_ball_warp: the rows of the remaining ball spread over the lanes, candidates fetched flattened, one 64-bit
// warp-min).  Without a queue the warp finishes its own far queries one after the other.  Only queries outside the grid
// or with a huge ball fall back to nn1_pruned().  The result is the exact nearest neighbour under the same total order
// (d2, original index) as nn1() / nn1_pruned() -- a cell is left out only when its lower bound (same conservative
// margins as nn1_gap) exceeds a distance at which a target point is known to exist -- so every execution mode keeps
// producing identical bits.
#pragma once

#include "grid.h"

namespace lb {

constexpr int NNS_ROWS = 9;            // rows (y, z) of one lane's candidate set handled by the staged path

struct NnsFarItem {                    // an undecided query handed to the grid
  int s;                               // source point
  int bs;                              // sorted index of the best candidate so far, or -1
  unsigned long long best;             // its packed key (d2 bits << 32 | original index), or the gate key
  float ball2;                         // squared radius within which everything that can still matter lies
  int pad;
};
struct NnsFarQueue { NnsFarItem* items; int* count; };

struct NnsWarp {                       // per-warp view of the staging memory
  f4* stage;                           // cap points
  int cap;
  unsigned long long* mbar;            // TMA variant: completion barrier (count 1)
  unsigned phase;                      // its parity
};

__device__ __forceinline__ void nns_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void nns_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void nns_mbar_init(unsigned long long* mbar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(a) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void nns_mbar_expect(unsigned long long* mbar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nns_mbar_wait(unsigned long long* mbar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(mbar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
// one contiguous run of points: global -> shared through the TMA unit (1-D bulk copy), bytes counted on the barrier
__device__ __forceinline__ void nns_bulk_copy(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* mbar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(mbar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d), "l"(gsrc),
               "r"(bytes), "r"(b) : "memory");
}

// warp-wide minimum of a packed (d2 bits << 32 | original index) key
__device__ __forceinline__ unsigned long long nns_warp_min(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  return v;
}

// One query, the whole warp (all lanes pass the same arguments): everything within sqrt(b2) of the query; the search
// starts from the best candidate known so far (found0: bd2_0, bi0, bs0).  The rows of the ball are looked up 32 at a time, one per lane,
// and the candidates of those rows are fetched flattened -- lane l takes candidates l, l + 32, ... of the concatenated
// runs.  `scratch` = 66 words of the warp's stage.  Returns false when the ball spans too many cells (caller falls back).
__device__ __forceinline__ bool nn1_ball_warp(const GridView& g, float qx, float qy, float qz, float max_d2, float b2, bool found0,
                                              float bd2_0, int bi0, int bs0, uint32_t* scratch, int& out_bs, int& out_bi, float& out_bd2) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  BallGeom b;
  ball_geom(g, qx, qy, qz, b);
  int ylo, yhi, zlo, zhi;
  const float S = ball_rows(b, b2, ylo, yhi, zlo, zhi);
  if (!(S < 24.0f)) return false;
  ylo = imax_(ylo, -b.cy); yhi = imin_(yhi, g.ny - 1 - b.cy);
  zlo = imax_(zlo, -b.cz); zhi = imin_(zhi, g.nz - 1 - b.cz);
  const int wy = yhi - ylo + 1, wz = zhi - zlo + 1;
  const float inv_wy = 1.0f / (float)(wy > 0 ? wy : 1);
  const int nrows = (wy > 0 && wz > 0) ? wy * wz : 0;
  unsigned long long best = found0 ? (((unsigned long long)__float_as_uint(bd2_0) << 32) | (unsigned)bi0)
                                   : ((unsigned long long)__float_as_uint(max_d2) << 32);      // keys >= the gate key fail d2 < max_d2
  int best_si = found0 ? bs0 : -1;
  uint32_t* s_a0 = scratch;            // [32] first point of the lane's row window
  uint32_t* s_off = scratch + 32;      // [33] exclusive prefix of the window lengths
  float e2 = b2;                        // shrinks to the best distance after every batch of rows
  for (int jb = 0; jb < nrows; jb += 32) {
    if (jb > 0) {
      const unsigned long long wb = nns_warp_min(best);         // every lane continues from the warp's best and its position
      best_si = __shfl_sync(FULL, best_si, __ffs((int)__ballot_sync(FULL, best == wb)) - 1);
      best = wb;
      e2 = fminf(e2, __uint_as_float((unsigned)(best >> 32)));
    }
    const int j = jb + lane;
    uint32_t a0 = 0; int n0 = 0;
    if (j < nrows) {
      const int jz = (int)(((float)j + 0.5f) * inv_wy);        // j / wy for 0 <= j < 2401, wy <= 49 (exact: the quotient's
      const int dz = zlo + jz, dy = ylo + (j - jz * wy);      // fractional part is at least 0.5 / 49 away from an integer)
      int xlo, xhi;
      if (ball_window(g, b, dy, dz, e2, xlo, xhi)) {
        const int base = ((b.cz + dz) * g.ny + (b.cy + dy)) * g.nx;
        a0 = g.cell_start[base + xlo];
        n0 = (int)(g.cell_start[base + xhi + 1] - a0);
      }
    }
    int incl = n0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) continue;
    __syncwarp();
    s_a0[lane] = a0; s_off[lane] = (uint32_t)(incl - n0);
    if (lane == 31) s_off[32] = (uint32_t)total;
    __syncwarp();
    int cur = 0;
    for (int idx = lane; idx < total; idx += 64) {
      while ((uint32_t)idx >= s_off[cur + 1]) cur++;
      const uint32_t pi = s_a0[cur] + ((uint32_t)idx - s_off[cur]);
      const int idx2 = idx + 32;
      uint32_t pj = pi;
      if (idx2 < total) {
        while ((uint32_t)idx2 >= s_off[cur + 1]) cur++;
        pj = s_a0[cur] + ((uint32_t)idx2 - s_off[cur]);
      }
      const f4 p = g.pts[pi];
      const f4 p2 = g.pts[pj];
      const unsigned long long k1 = ((unsigned long long)__float_as_uint(dist2(qx, qy, qz, p.x, p.y, p.z)) << 32) | (unsigned)float_to_bits(p.w);
      const unsigned long long k2 = ((unsigned long long)__float_as_uint(dist2(qx, qy, qz, p2.x, p2.y, p2.z)) << 32) | (unsigned)float_to_bits(p2.w);
      if (k1 < best) { best = k1; best_si = (int)pi; }
      if (k2 < best) { best = k2; best_si = (int)pj; }       // pj == pi when there is no second candidate: harmless
    }
  }
  const unsigned long long wbest = nns_warp_min(best);
  // the winning key is unique (original indices are), unless nobody improved on the start value
  const unsigned holders = __ballot_sync(FULL, best == wbest);
  const int src = __ffs((int)holders) - 1;
  const int si = __shfl_sync(FULL, best_si, src);
  const bool any = found0 || (wbest >> 32) < (unsigned long long)__float_as_uint(max_d2);
  out_bs = any ? si : -1;
  out_bi = any ? (int)(unsigned)(wbest & 0xffffffffull) : -1;
  out_bd2 = any ? __uint_as_float((unsigned)(wbest >> 32)) : max_d2;
  return true;
}

// All 32 lanes of a warp call this together (active = this lane has a query).  have_ub: a target point is known at
// squared distance ub2 from the query (same float32 dist2 as the search).  Returns the SORTED index of the nearest
// target point with d2 < max_d2, or -1; original index and d2 by reference (like nn1_pruned).  With a queue, an
// undecided query (source point src_index) is appended to it and NNS_DEFERRED is returned.
//
// The staged pass searches the ball that guarantees completeness -- of the known bound, else of the gate -- cut to half
// a cell when it is larger.  A complete ball is final by construction; a cut one is final when its best candidate
// lies inside it (every point that close was among the candidates), else the query is undecided: its remaining ball
// (best so far, bound, or gate) goes to the queue.
constexpr int NNS_DEFERRED = -2;
constexpr float NNS_R0 = 0.5f;         // radius of the first look, in cells, when no tighter bound is known

template <bool TMA>
__device__ __forceinline__ int nn1_staged(const GridView& g, bool active, float qx, float qy, float qz, float max_d2, bool have_ub,
                                          float ub2, NnsWarp& w, int& best_orig, float& best_d2, const NnsFarQueue* fq = nullptr,
                                          int src_index = 0, long long* wprof = nullptr, float r0cut = NNS_R0) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long tp0 = wprof ? clock64() : 0;
  BallGeom b;
  ball_geom(g, qx, qy, qz, b);
  int r0, r1;
  ring_range(g, b.cx, b.cy, b.cz, r0, r1);
  (void)r1;
  const unsigned long long gate_key = (unsigned long long)__float_as_uint(max_d2) << 32;   // keys >= this fail d2 < max_d2
  unsigned long long best = gate_key;                            // (d2 bits << 32 | original index) of the best candidate
  int bs = -1;
  const bool outside = active && (r0 > 1 || !(max_d2 > 0.f));    // query more than a cell outside the grid: serial search
  const bool hub = have_ub && ub2 < max_d2;
  bool far = false;                                              // not decided yet after a pass
  long long prof_cand = 0, pt_rows = 0, pt_copy = 0, pt_scan = 0, pt_mark = tp0;
  int prof_chunks = 0;

  // ---- this lane's ball: what guarantees completeness (the known bound, else the gate); with r0cut > 0 cut to that
  // many cells (half a cell: at most 3 x 3 rows, usually 2 x 2 cells), without a cut up to 9 x 9 rows in batches of 9
  bool inpass = active && !outside;
  const float need2 = hub ? ub2 : max_d2;
  const float cap2 = r0cut > 0.f ? (r0cut * r0cut) * b.hh : need2;
  const float b2 = fminf(need2, cap2);
  const bool complete = need2 <= cap2;
  int ylo = 0, yhi = 0, zlo = 0, zhi = 0;
  if (!(ball_rows(b, b2, ylo, yhi, zlo, zhi) < 4.0f)) {           // a huge ball: nn1_ball_warp / the serial search
    ylo = yhi = zlo = zhi = 0;
    if (inpass) { far = true; inpass = false; }
  }

  {
    const int wy = yhi - ylo + 1, wz = zhi - zlo + 1;
    const int nrows = inpass ? wy * wz : 0;
    const int nbatch = __reduce_max_sync(FULL, (nrows + NNS_ROWS - 1) / NNS_ROWS);
    for (int batch = 0; batch < nbatch; batch++) {
      // ---- rows of this batch; a ball shrinks to the best distance found so far
      const float e2 = fminf(b2, __uint_as_float((unsigned)(best >> 32)));   // best >> 32 = the gate while nothing is found
      int a_lo[NNS_ROWS], a_hi[NNS_ROWS];
      int dy = ylo + (batch * NNS_ROWS) % wy, dz = zlo + (batch * NNS_ROWS) / wy;      // raster walk over the rows
#pragma unroll
      for (int r = 0; r < NNS_ROWS; r++, dy++) {
        a_lo[r] = 0; a_hi[r] = 0;
        if (dy > yhi) { dy = ylo; dz++; }
        const int j = batch * NNS_ROWS + r;
        int xlo, xhi;
        if (j < nrows && ball_window(g, b, dy, dz, e2, xlo, xhi)) {
          const int base = ((b.cz + dz) * g.ny + (b.cy + dy)) * g.nx;
          a_lo[r] = base + xlo; a_hi[r] = base + xhi + 1;
        }
      }
      // ---- CSR offsets of all rows, issued back to back (index 0 for rows that do not exist: both loads hit the same word)
      uint32_t rs[NNS_ROWS], re[NNS_ROWS];
#pragma unroll
      for (int r = 0; r < NNS_ROWS; r++) { rs[r] = g.cell_start[a_lo[r]]; re[r] = g.cell_start[a_hi[r]]; }
      int total = 0;
#pragma unroll
      for (int r = 0; r < NNS_ROWS; r++) { re[r] -= rs[r]; total += (int)re[r]; }      // re = run length from here on
      if (inpass && total > w.cap) { far = true; inpass = false; }                     // dense cells: nn1_ball_warp streams them
      if (!inpass) total = 0;
      if (wprof) { prof_cand += total; const long long tn = clock64(); pt_rows += tn - pt_mark; pt_mark = tn; }
      // ---- chunks of lanes whose candidates fit into the stage together: copy, wait, every lane scans its own
      bool todo = total > 0;
      while (__any_sync(FULL, todo)) {
        const int t = todo ? total : 0;
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += v;
        }
        const int off = incl - t;
        const bool in = todo && (incl <= w.cap);
        if (TMA) {
          const unsigned inmask = __ballot_sync(FULL, in);
          const int last = 31 - __clz((int)inmask);                  // the chunk is a prefix of the todo lanes
          const int chunk_pts = __shfl_sync(FULL, incl, last);
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // earlier generic reads of the stage vs the async writes
          if (lane == 0) nns_mbar_expect(w.mbar, (unsigned)chunk_pts * 16u);
          __syncwarp();
          if (in) {
            int k = off;
#pragma unroll
            for (int r = 0; r < NNS_ROWS; r++) {
              if (re[r]) nns_bulk_copy(&w.stage[k], &g.pts[rs[r]], re[r] * 16u, w.mbar);
              k += (int)re[r];
            }
          }
          nns_mbar_wait(w.mbar, w.phase);
          w.phase ^= 1u;
        } else {
          if (in) {
            int k = off;
#pragma unroll
            for (int r = 0; r < NNS_ROWS; r++)
              for (uint32_t i = 0; i < re[r]; i++, k++) nns_cp_async16(&w.stage[k], &g.pts[rs[r] + i]);
          }
          nns_cp_async_wait_all();
        }
        if (wprof) { const long long tn = clock64(); pt_copy += tn - pt_mark; pt_mark = tn; prof_chunks++; }
        if (in) {
          // eight candidates per trip, branch-free: packed keys (d2 bits << 32 | original index) order exactly like
          // better(); a tree of selects finds the trip's minimum and its position, one compare merges it into the best
          const f4* sp = w.stage + off;
          int bk = -1;
          for (int k = 0; k < t; k += 8) {
            unsigned long long key[8]; int pos[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const int kk = (k + u < t) ? (k + u) : (t - 1);        // the tail repeats the last candidate: harmless
              const f4 p = sp[kk];
              key[u] = ((unsigned long long)__float_as_uint(dist2(qx, qy, qz, p.x, p.y, p.z)) << 32) | (unsigned)float_to_bits(p.w);
              pos[u] = kk;
            }
#pragma unroll
            for (int st = 1; st < 8; st <<= 1)
#pragma unroll
              for (int u = 0; u < 8; u += 2 * st) {
                const bool lt = key[u + st] < key[u];
                key[u] = lt ? key[u + st] : key[u];
                pos[u] = lt ? pos[u + st] : pos[u];
              }
            if (key[0] < best) { best = key[0]; bk = pos[0]; }
          }
          if (bk >= 0) {                       // position in the stage -> sorted index of the cloud
            int acc = 0;
#pragma unroll
            for (int r = 0; r < NNS_ROWS; r++) {
              if (bk >= acc && bk < acc + (int)re[r]) bs = (int)rs[r] + (bk - acc);
              acc += (int)re[r];
            }
          }
          todo = false;
        }
        __syncwarp();
        if (wprof) { const long long tn = clock64(); pt_scan += tn - pt_mark; pt_mark = tn; }
      }
    }
    // ---- decision: a complete ball is final; a cut one only when its best lies inside it
    if (inpass && !complete && !(best < gate_key && __uint_as_float((unsigned)(best >> 32)) <= b2)) far = true;
  }
  // ---- undecided queries: to the grid's queue, or one after the other with the whole warp on each
  unsigned farmask = __ballot_sync(FULL, far);
  bool serial = outside;
  bool deferred = false;
  const long long tp1 = wprof ? clock64() : 0;
  if (wprof) {
    int tt = (int)prof_cand;
    for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(FULL, tt, o);
    if (lane == 0) { wprof[0] = tp1 - tp0; wprof[2] = __popc(farmask); wprof[5] = tt; wprof[4] = pt_rows; wprof[6] = pt_copy; wprof[7] = pt_scan + ((long long)prof_chunks << 40); }
  }
  const bool found = best < gate_key;
  float bd2 = __uint_as_float((unsigned)(best >> 32));           // == max_d2 while nothing is found
  int bi = found ? (int)(unsigned)(best & 0xffffffffull) : -1;
  const float ball2 = fminf(bd2, hub ? ub2 : max_d2);
  if (fq && farmask) {
    int base = 0;
    const int leader = __ffs((int)farmask) - 1;
    if (lane == leader) base = atomicAdd(fq->count, __popc(farmask));
    base = __shfl_sync(FULL, base, leader);
    if (far) {
      NnsFarItem it;
      it.s = src_index; it.bs = bs; it.best = best; it.ball2 = ball2; it.pad = 0;
      fq->items[base + __popc(farmask & ((1u << lane) - 1u))] = it;
      deferred = true;
    }
    farmask = 0;
  }
  while (farmask) {
    const int l = __ffs((int)farmask) - 1;
    farmask &= farmask - 1;
    const float x = __shfl_sync(FULL, qx, l), y = __shfl_sync(FULL, qy, l), z = __shfl_sync(FULL, qz, l);
    const bool f0 = __shfl_sync(FULL, (int)found, l) != 0;
    const float d0 = __shfl_sync(FULL, bd2, l), r2 = __shfl_sync(FULL, ball2, l);
    const int i0 = __shfl_sync(FULL, bi, l), s0 = __shfl_sync(FULL, bs, l);
    int rbs, rbi; float rbd;
    const bool ok = nn1_ball_warp(g, x, y, z, max_d2, r2, f0, d0, i0, s0, reinterpret_cast<uint32_t*>(w.stage), rbs, rbi, rbd);
    if (lane == l) {
      if (ok) { bs = rbs; bi = rbi; bd2 = rbd; }
      else serial = true;
    }
    __syncwarp();
  }
  if (wprof && lane == 0) wprof[1] = clock64() - tp1;
  if (serial) bs = nn1_pruned(g, qx, qy, qz, max_d2, bi, bd2);
  best_orig = bi; best_d2 = bd2;
  return deferred ? NNS_DEFERRED : bs;
}

// One queued query, the whole warp (every lane passes the same item and query): sorted index of its nearest neighbour or -1.
__device__ __forceinline__ int nn1_far_item(const GridView& g, const NnsFarItem& it, float qx, float qy, float qz, float max_d2,
                                            uint32_t* scratch) {
  const unsigned long long gate_key = (unsigned long long)__float_as_uint(max_d2) << 32;
  const bool f0 = it.best < gate_key;
  int rbs, rbi; float rbd;
  const bool ok = nn1_ball_warp(g, qx, qy, qz, max_d2, it.ball2, f0, __uint_as_float((unsigned)(it.best >> 32)),
                                f0 ? (int)(unsigned)(it.best & 0xffffffffull) : -1, it.bs, scratch, rbs, rbi, rbd);
  if (!ok) rbs = nn1_pruned(g, qx, qy, qz, max_d2, rbi, rbd);      // a huge ball: every lane computes the same serial search
  return rbs;
}

}  // namespace lb
