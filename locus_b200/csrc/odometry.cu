// Scan-to-scan odometry pipeline (lb_odometry_*): the per-scan chain of the reference's lidar callback
//   filtered scan -> PointCloudOdometry::SetLidar -> UpdateEstimate -> UpdateICP
//   (locus/src/Locus.cc:451-453, point_cloud_odometry/src/PointCloudOdometry.cc:221-274)
// as a two-stage host pipeline over the library's own C ABI:
//
//   stage V  one thread, one lb_voxel handle (own stream): VoxelGrid of scan k into slot k % R of a ring of
//            device-resident filtered clouds
//   stage G  `depth` threads, one lb_gicp handle each (own streams): set_source(ring[k]), set_target(ring[k-1]),
//            align -> result k
//
// Nothing here touches a point: every byte of device work is done by voxel.cu / gicp.cu kernels.  The pipeline
// exists because one align() is a chain of latency-bound kernels on 30-60 of 148 SMs: the rest of the GPU filters
// and indexes the following scans, and several aligns (each on its own CTAs) are in flight at once.  Results are
// those of the sequential calls: registration k reads only filtered clouds k and k-1 and its caller-given prior.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#ifdef LB_ODOMETRY_STUB          // TEST-ONLY build without CUDA: tests/odometry_harness.cpp supplies fake stages
#include "odometry_stub.h"
#else
#include "prims.cuh"
#endif

using namespace lb;

namespace {

struct Job {
  uint64_t ticket = 0;
  const uint8_t* scan = nullptr;
  size_t n_pts = 0;
  uint32_t point_step = 0;
  std::vector<lb_field> fields;
  int mem = LB_MEM_HOST;
  bool has_guess = false;
  float guess[16];
  uint8_t* filtered_out = nullptr;
  int mem_filtered = LB_MEM_HOST;
  int xyz_off = 0;
};

struct Slot {            // one filtered cloud of the ring
  uint8_t* d = nullptr;
  size_t n = 0;          // points; 0 = unusable (stage V failed)
  uint32_t point_step = 0;
  int xyz_off = 0;
  int status = LB_OK;
  char error[160] = "";
};

struct Prepared {        // cloud-sharing mode: scan t's prepared cloud, published by the worker of registration t
  uint64_t ticket = ~0ull;
  lb_cloud* cloud = nullptr;     // null when the preparation failed
  bool ready = false;
};

struct Entry {           // per in-flight ticket
  Job job;
  lb_odometry_result res;
  bool filtered = false; // stage V done
  bool done = false;     // stage G done (result complete)
};

}  // namespace

struct lb_odometry {
  int device = 0;
  int depth = 1;
  size_t max_points = 0;
  uint32_t max_step = 0;
  lb_voxel* vg = nullptr;
  std::vector<lb_gicp*> gicp;
  std::vector<Slot> ring;
  std::vector<Prepared> prep;          // same indexing as ring (ticket % R); used when share is on
  bool share = true;                   // lb_odometry_set_cloud_sharing (on by default)
  cudaStream_t copy_stream = nullptr;
  cudaStream_t voxel_stream = nullptr;

  std::mutex mu;
  std::condition_variable cv;          // one condition variable for every state change (few threads, short waits)
  std::deque<Entry*> inflight;         // tickets [next_return, next_ticket), in order
  uint64_t next_ticket = 0;            // next submission number
  uint64_t next_filter = 0;            // next ticket stage V takes
  uint64_t next_align = 0;             // next ticket stage G hands to a worker
  uint64_t next_return = 0;            // next ticket lb_odometry_next returns
  uint64_t done_floor = 0;             // every ticket < done_floor has finished stage G
  bool stop = false;
  // stage accounting (seconds, host wall clock; mu held when updated): which stage bounds the throughput
  double t_voxel_busy = 0, t_voxel_wait = 0, t_worker_busy = 0, t_worker_wait = 0;
  uint64_t n_filtered = 0, n_registered = 0;
  std::thread vthread;
  std::vector<std::thread> gthreads;

  Entry* entry(uint64_t t) { return inflight[(size_t)(t - next_return)]; }   // mu held
};

namespace {

void copy_err(char* dst, size_t cap) {
  const char* e = lb_last_error_string();
  strncpy(dst, e ? e : "", cap - 1);
  dst[cap - 1] = 0;
}

void advance_done_floor(lb_odometry* h) {   // mu held
  while (h->done_floor < h->next_ticket) {
    uint64_t t = h->done_floor;
    if (t < h->next_return) { h->done_floor++; continue; }
    if (!h->entry(t)->done) break;
    h->done_floor++;
  }
}

void voxel_stage(lb_odometry* h) {
  cudaSetDevice(h->device);
  const uint64_t R = h->ring.size();
  for (;;) {
    Entry* e = nullptr;
    uint64_t t = 0;
    const auto tw0 = std::chrono::steady_clock::now();
    {
      std::unique_lock<std::mutex> lk(h->mu);
      // slot t % R was last read by registrations t-R (source) and t-R+1 (target): both must have finished
      h->cv.wait(lk, [&] {
        return h->stop || (h->next_filter < h->next_ticket && (h->next_filter + 2 <= R + h->done_floor));
      });
      if (h->stop) return;
      t = h->next_filter;
      e = h->entry(t);
    }
    const auto tb0 = std::chrono::steady_clock::now();
    Slot& s = h->ring[t % R];
    const Job& j = e->job;
    size_t n_out = 0;
    int st = lb_voxel_filter(h->vg, j.scan, j.n_pts, j.point_step, j.fields.data(), (int)j.fields.size(), nullptr, 0, s.d,
                             h->max_points, &n_out, nullptr, j.mem, LB_MEM_DEVICE);
    s.status = st; s.n = (st == LB_OK) ? n_out : 0; s.point_step = j.point_step; s.xyz_off = j.xyz_off;
    if (st != LB_OK) copy_err(s.error, sizeof(s.error));
    if (st == LB_OK && j.filtered_out && n_out) {
      cudaError_t ce = cudaMemcpyAsync(j.filtered_out, s.d, n_out * j.point_step,
                                       j.mem_filtered == LB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                       h->copy_stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->copy_stream);
      if (ce != cudaSuccess) {
        s.status = LB_ERR_CUDA; s.n = 0;
        snprintf(s.error, sizeof(s.error), "lb_odometry: filtered cloud copy failed: %s", cudaGetErrorString(ce));
      }
    }
    {
      std::lock_guard<std::mutex> lk(h->mu);
      e->filtered = true;
      e->res.n_filtered = s.n;
      h->next_filter = t + 1;
      h->t_voxel_wait += std::chrono::duration<double>(tb0 - tw0).count();
      h->t_voxel_busy += std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();
      h->n_filtered++;
    }
    h->cv.notify_all();
  }
}

void align_stage(lb_odometry* h, int w) {
  cudaSetDevice(h->device);
  lb_gicp* g = h->gicp[(size_t)w];
  const uint64_t R = h->ring.size();
  for (;;) {
    Entry* e = nullptr;
    uint64_t t = 0;
    const auto tw0 = std::chrono::steady_clock::now();
    {
      std::unique_lock<std::mutex> lk(h->mu);
      h->cv.wait(lk, [&] { return h->stop || h->next_align < h->next_filter; });
      if (h->stop) return;
      t = h->next_align++;
      e = h->entry(t);
    }
    const auto tb0 = std::chrono::steady_clock::now();
    lb_odometry_result& r = e->res;
    const Slot& cur = h->ring[t % R];
    r.status = cur.status;
    if (cur.status != LB_OK) memcpy(r.error, cur.error, sizeof(r.error));
    if (h->share) {
      // this scan's cloud is prepared here, once, and published for the worker of the next registration
      lb_cloud* mine = nullptr;
      int st = cur.status;
      if (st == LB_OK && cur.n > 0) {
        st = lb_gicp_set_source(g, cur.d, cur.n, cur.point_step, (size_t)cur.xyz_off, LB_NO_NORMALS, LB_MEM_DEVICE);
        if (st == LB_OK) st = lb_gicp_prepare_source(g);
        if (st == LB_OK) st = lb_gicp_share_source(g, &mine);
        if (st != LB_OK) { r.status = st; copy_err(r.error, sizeof(r.error)); }
      }
      lb_cloud* prev = nullptr;
      {
        std::unique_lock<std::mutex> lk(h->mu);
        Prepared& p = h->prep[t % R];
        if (p.cloud) lb_cloud_release(p.cloud);       // slot of ticket t-R: both of its users finished long ago
        p.ticket = t; p.cloud = mine; p.ready = true;
        h->cv.notify_all();
        if (t > 0) {
          Prepared& q = h->prep[(t - 1) % R];
          h->cv.wait(lk, [&] { return h->stop || (q.ticket == t - 1 && q.ready); });
          if (!h->stop) prev = q.cloud;
        }
      }
      const bool adopt = t > 0 && mine && prev;
      if (adopt) st = lb_gicp_set_target_cloud(g, prev);       // the handle now holds its own reference
      if (t > 0) {
        // this worker is the only consumer of scan t-1's published reference: give it back now, so that the
        // owner can recycle the object two jobs later instead of R tickets later (no allocation in steady state)
        std::lock_guard<std::mutex> lk(h->mu);
        Prepared& q = h->prep[(t - 1) % R];
        if (q.ticket == t - 1 && q.cloud) { lb_cloud_release(q.cloud); q.cloud = nullptr; }
      }
      if (adopt) {
        if (st == LB_OK) st = lb_gicp_align(g, e->job.has_guess ? e->job.guess : nullptr, &r.gicp);
        r.status = st;
        if (st == LB_OK) r.has_pose = 1; else copy_err(r.error, sizeof(r.error));
      }
    } else if (t > 0 && cur.status == LB_OK) {
      const Slot& prv = h->ring[(t - 1) % R];
      if (prv.status == LB_OK && prv.n > 0 && cur.n > 0) {
        int st = lb_gicp_set_source(g, cur.d, cur.n, cur.point_step, (size_t)cur.xyz_off, LB_NO_NORMALS, LB_MEM_DEVICE);
        if (st == LB_OK)
          st = lb_gicp_set_target(g, prv.d, prv.n, prv.point_step, (size_t)prv.xyz_off, LB_NO_NORMALS, LB_MEM_DEVICE, nullptr);
        if (st == LB_OK) st = lb_gicp_align(g, e->job.has_guess ? e->job.guess : nullptr, &r.gicp);
        r.status = st;
        if (st == LB_OK) r.has_pose = 1; else copy_err(r.error, sizeof(r.error));
      }
    }
    {
      std::lock_guard<std::mutex> lk(h->mu);
      e->done = true;
      advance_done_floor(h);
      h->t_worker_wait += std::chrono::duration<double>(tb0 - tw0).count();
      h->t_worker_busy += std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();
      h->n_registered++;
    }
    h->cv.notify_all();
  }
}

}  // namespace

extern "C" {

int lb_odometry_create(int device, int depth, size_t max_points, uint32_t max_point_step, lb_odometry** out) {
  if (!out || depth < 1 || depth > 16 || max_points == 0 || max_point_step < 12) {
    set_error("lb_odometry_create: need depth in 1..16, max_points > 0, max_point_step >= 12");
    return LB_ERR_INVALID_ARG;
  }
  lb_odometry* h = new lb_odometry;
  h->device = device; h->depth = depth; h->max_points = max_points; h->max_step = max_point_step;
  // The VoxelGrid stage is the one serial stage of the pipeline and a chain of ~12 short dependent kernels: on its
  // own high-priority stream its CTAs are placed before the pending CTAs of the workers' k-NN / index kernels.
  int st = LB_OK;
  {
    int lo = 0, hi = 0;
    const char* e = getenv("LB_VOXEL_PRIO");
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->voxel_stream, cudaStreamNonBlocking, (e && atoi(e) == 0) ? lo : hi) != cudaSuccess) {
      set_error("lb_odometry_create: stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
      st = LB_ERR_CUDA;
    }
  }
  if (st == LB_OK) st = lb_voxel_create_on_stream(device, h->voxel_stream, &h->vg);
  for (int i = 0; st == LB_OK && i < depth; i++) {
    lb_gicp* g = nullptr;
    st = lb_gicp_create(device, &g);
    // no device allocation in steady state: clouds of up to max_points points, and enough spare cloud objects for the
    // sharing mode (a prepared cloud stays referenced until the next worker has adopted and used it)
    if (st == LB_OK) h->gicp.push_back(g);
    if (st == LB_OK) st = lb_gicp_reserve(g, max_points, 8);
  }
  if (st == LB_OK) {
    h->ring.resize((size_t)depth + 3);
    h->prep.resize((size_t)depth + 3);
    for (auto& s : h->ring) {
      if (cudaMalloc((void**)&s.d, max_points * (size_t)max_point_step) != cudaSuccess) {
        set_error("lb_odometry_create: ring allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        st = LB_ERR_CUDA;
        break;
      }
    }
  }
  if (st == LB_OK && cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("lb_odometry_create: stream creation failed");
    st = LB_ERR_CUDA;
  }
  if (st != LB_OK) {
    for (auto& s : h->ring) if (s.d) cudaFree(s.d);
    for (auto g : h->gicp) lb_gicp_destroy(g);
    if (h->vg) lb_voxel_destroy(h->vg);
    if (h->voxel_stream) cudaStreamDestroy(h->voxel_stream);
    delete h;
    return st;
  }
  h->vthread = std::thread(voxel_stage, h);
  for (int i = 0; i < depth; i++) h->gthreads.emplace_back(align_stage, h, i);
  *out = h;
  return LB_OK;
}

int lb_odometry_destroy(lb_odometry* h) {
  if (!h) return LB_OK;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    h->stop = true;
  }
  h->cv.notify_all();
  if (h->vthread.joinable()) h->vthread.join();
  for (auto& t : h->gthreads) if (t.joinable()) t.join();
  cudaSetDevice(h->device);
  for (auto& p : h->prep) if (p.cloud) { lb_cloud_release(p.cloud); p.cloud = nullptr; }
  for (auto g : h->gicp) lb_gicp_destroy(g);
  lb_voxel_destroy(h->vg);
  for (auto& s : h->ring) if (s.d) cudaFree(s.d);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->voxel_stream) cudaStreamDestroy(h->voxel_stream);
  for (auto e : h->inflight) delete e;
  delete h;
  return LB_OK;
}

lb_voxel* lb_odometry_voxel(lb_odometry* h) { return h ? h->vg : nullptr; }
lb_gicp* lb_odometry_gicp(lb_odometry* h, int i) { return (h && i >= 0 && i < h->depth) ? h->gicp[(size_t)i] : nullptr; }
int lb_odometry_depth(lb_odometry* h) { return h ? h->depth : 0; }

int lb_odometry_set_gicp_params(lb_odometry* h, const lb_gicp_params* p) {
  if (!h || !p) { set_error("lb_odometry_set_gicp_params: null argument"); return LB_ERR_INVALID_ARG; }
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->next_return != h->next_ticket) { set_error("lb_odometry_set_gicp_params: pipeline not idle"); return LB_ERR_INVALID_ARG; }
  }
  for (auto g : h->gicp) LB_TRY(lb_gicp_set_params(g, p));
  return LB_OK;
}

int lb_odometry_set_cloud_sharing(lb_odometry* h, int on) {
  if (!h) { set_error("lb_odometry_set_cloud_sharing: null handle"); return LB_ERR_INVALID_ARG; }
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->next_return != h->next_ticket) { set_error("lb_odometry_set_cloud_sharing: pipeline not idle"); return LB_ERR_INVALID_ARG; }
  if (h->next_ticket != 0 && (on != 0) != h->share) {
    // the previous scan's cloud exists in one form only (ring slot or prepared cloud): switch before the first scan
    set_error("lb_odometry_set_cloud_sharing: switch before the first scan is submitted");
    return LB_ERR_UNSUPPORTED;
  }
  h->share = on != 0;
  return LB_OK;
}

int lb_odometry_submit(lb_odometry* h, const uint8_t* scan, size_t n_pts, uint32_t point_step, const lb_field* fields,
                       int n_fields, int mem, const float* guess, uint8_t* filtered_out, int mem_filtered,
                       uint64_t* ticket) {
  if (!h || !scan || !fields || n_fields <= 0) { set_error("lb_odometry_submit: null argument"); return LB_ERR_INVALID_ARG; }
  if (n_pts > h->max_points || point_step > h->max_step) {
    set_error("lb_odometry_submit: scan of %zu x %u bytes exceeds the pipeline's ring (%zu x %u)", n_pts, point_step,
              h->max_points, h->max_step);
    return LB_ERR_CAPACITY;
  }
  int xo = -1;
  for (int f = 0; f < n_fields; f++)
    if (!strcmp(fields[f].name, "x") && fields[f].datatype == LB_FLOAT32) xo = (int)fields[f].offset;
  if (xo < 0) { set_error("lb_odometry_submit: FLOAT32 field 'x' required"); return LB_ERR_INVALID_ARG; }
  Entry* e = new Entry;
  Job& j = e->job;
  j.scan = scan; j.n_pts = n_pts; j.point_step = point_step; j.fields.assign(fields, fields + n_fields); j.mem = mem;
  j.has_guess = guess != nullptr;
  if (guess) memcpy(j.guess, guess, sizeof(j.guess));
  j.filtered_out = filtered_out; j.mem_filtered = mem_filtered; j.xyz_off = xo;
  memset(&e->res, 0, sizeof(e->res));
  {
    std::unique_lock<std::mutex> lk(h->mu);
    const uint64_t limit = 2 * (uint64_t)h->depth + 2;
    h->cv.wait(lk, [&] { return h->next_ticket - h->done_floor < limit; });
    j.ticket = e->res.ticket = h->next_ticket++;
    h->inflight.push_back(e);
    if (ticket) *ticket = j.ticket;
  }
  h->cv.notify_all();
  return LB_OK;
}

int lb_odometry_next(lb_odometry* h, lb_odometry_result* r, int block) {
  if (!h || !r) { set_error("lb_odometry_next: null argument"); return LB_ERR_INVALID_ARG; }
  Entry* e = nullptr;
  {
    std::unique_lock<std::mutex> lk(h->mu);
    if (h->next_return == h->next_ticket) { set_error("lb_odometry_next: nothing pending"); return LB_ERR_NO_ALIGN; }
    e = h->inflight.front();
    if (!e->done) {
      if (!block) return 1;
      h->cv.wait(lk, [&] { return e->done; });
    }
    h->inflight.pop_front();
    h->next_return++;
    advance_done_floor(h);
  }
  h->cv.notify_all();
  *r = e->res;
  delete e;
  return LB_OK;
}

int lb_odometry_pending(lb_odometry* h, size_t* n) {
  if (!h || !n) return LB_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  *n = (size_t)(h->next_ticket - h->next_return);
  return LB_OK;
}

int lb_odometry_stage_times(lb_odometry* h, double* out6) {
  if (!h || !out6) return LB_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->mu);
  out6[0] = (double)h->n_filtered; out6[1] = h->t_voxel_busy; out6[2] = h->t_voxel_wait;
  out6[3] = (double)h->n_registered; out6[4] = h->t_worker_busy; out6[5] = h->t_worker_wait;
  return LB_OK;
}

int lb_odometry_launch_count(lb_odometry* h, uint64_t* n) {
  if (!h || !n) return LB_ERR_INVALID_ARG;
  uint64_t tot = 0, v = 0;
  lb_voxel_launch_count(h->vg, &v); tot += v;
  for (auto g : h->gicp) { lb_gicp_launch_count(g, &v); tot += v; }
  *n = tot;
  return LB_OK;
}

}  // extern "C"
