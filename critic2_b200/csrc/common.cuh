// common.cuh -- context, resident grids, error handling and profiling shared by the kernels.
#pragma once

#include <cuda_runtime.h>

#include "nccl_shim.h"

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/critic2_gpu.h"

#define C2G_NSM_FALLBACK 148

struct c2g_grid {
  int n[3] = {0, 0, 0};
  long long nn = 0;
  double* d = nullptr;  // device, f(n1,n2,n3) index 1 fastest
  bool used = false;
  cudaEvent_t ready = nullptr;  // c2g_grid_upload_async: recorded on the copy stream after the H2D copy
  bool pending = false;         // the compute stream has not been ordered after `ready` yet
};

struct c2g_prof_entry {
  std::string name;
  double ms = 0.0;
  int launches = 0;
};

struct c2g_group;  // group.cu: the devices and worker threads of a c2g_init_devices context
struct c2g_context {
  c2g_group* group = nullptr;  // non-null: this context drives several devices of one process (group.cu)
  c2g_context* parent = nullptr;  // non-null: this is the per-device context of a multi-device context
  int device = 0;
  int nsm = C2G_NSM_FALLBACK;
  cudaStream_t stream = nullptr;
  std::vector<c2g_grid> grids;
  std::string err;
  std::string desc;
  // profiling
  bool prof_on = false;
  std::vector<c2g_prof_entry> prof;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;  // entry index, (start, stop)
  std::vector<cudaEvent_t> event_pool;
  long long launches = 0;
  // L2 flush buffer
  void* flushbuf = nullptr;
  size_t flushbytes = 0;
  // multi-GPU
  int rank = 0, nranks = 1;
  void* nccl = nullptr;  // ncclComm_t
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // c2g_timer_*
  // asynchronous host <-> device copies (c2g_grid_upload_async, c2g_basins_labels_async)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_order = nullptr;
  std::vector<void*> deferred;  // device blocks still read by copy_out: released by c2g_synchronize
  void* fft_cache = nullptr;    // cuFFT plans per grid shape (fft.cu)
  int* hpin = nullptr;          // page-locked scratch (64 ints) for the small counter read-backs between launches:
                                // a copy into pageable memory is staged by the driver and costs tens of microseconds more
  // device-memory cache (c2g_alloc / c2g_release below)
  std::multimap<size_t, void*> mem_free;           // size -> cached block
  std::unordered_map<void*, size_t> mem_live;      // block handed out -> its size
  size_t mem_cached_bytes = 0;

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  cudaEvent_t get_event() {
    if (!event_pool.empty()) {
      cudaEvent_t e = event_pool.back();
      event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  int prof_index(const char* name) {
    for (size_t i = 0; i < prof.size(); i++)
      if (prof[i].name == name) return (int)i;
    prof.push_back({name, 0.0, 0});
    return (int)prof.size() - 1;
  }
  // bracket a kernel launch (or a group) with events when profiling is on
  void prof_begin(const char* name) {
    if (!prof_on) return;
    int idx = prof_index(name);
    cudaEvent_t a = get_event(), b = get_event();
    cudaEventRecord(a, stream);
    pending.push_back({idx, {a, b}});
  }
  void prof_end(int nlaunch = 1) {
    launches += nlaunch;
    if (!prof_on) return;
    auto& p = pending.back();
    cudaEventRecord(p.second.second, stream);
    prof[p.first].launches += nlaunch;
  }
  // resolve pending events (after a synchronize)
  void prof_collect() {
    for (auto& p : pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.second.first, p.second.second) == cudaSuccess) prof[p.first].ms += ms;
      event_pool.push_back(p.second.first);
      event_pool.push_back(p.second.second);
    }
    pending.clear();
  }
};

#define C2G_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return (ctx)->fail(e__ == cudaErrorMemoryAllocation ? C2G_ERR_NOMEM : C2G_ERR_CUDA,           \
                         "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));      \
  } while (0)

#define C2G_KERNEL_CHECK(ctx)                                                                       \
  do {                                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                           \
    if (e__ != cudaSuccess)                                                                         \
      return (ctx)->fail(C2G_ERR_CUDA, "%s:%d: kernel launch: %s", __FILE__, __LINE__,              \
                         cudaGetErrorString(e__));                                                  \
  } while (0)

// result of an assignment (BADER or YT), device resident
struct c2g_basins {
  std::vector<c2g_basins*> parts;  // multi-device context: the z-slab result of every device (group.cu)
  c2g_context* ctx = nullptr;
  int kind = 0;  // 0 = bader, 1 = yt, 2 = isosurface regions (plain labels like bader, -1 = below the contour value)
  int gridh = -1;
  int n[3] = {0, 0, 0};
  long long nn = 0;
  int nmax = 0;
  // Bader: label[i] = index (0..nmax-1) into the ordered maxima list
  // YT   : label[i] = index of the basin for interior points, -1 for IAS points
  int* d_label = nullptr;  // owned planes [zlo, zhi) (points into d_lbuf for Bader slabs)
  int* d_lbuf = nullptr;   // Bader: label buffer with one halo plane below and above
  int zlo = 0, zhi = 0;    // owned z range (single GPU: 0..n3)
  std::vector<int> max_lin;        // linear id of each maximum, in the returned order
  std::vector<long long> counts;   // points per maximum
  std::vector<int> map;            // maximum -> basin id (1-based; 0 = discarded)
  int nattr = 0;
  bool has_map = false;
  int* d_map = nullptr;
  long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // YT extras
  int nvec = 0;
  std::vector<int> vec;
  std::vector<double> area;
  int* d_vec = nullptr;
  double* d_area = nullptr;
  void* yt = nullptr;  // YtState (yt.cu)
  long long n_ias = 0;
};
void c2g_yt_free_state(c2g_basins* res);
int c2g_yt_apply_map(c2g_basins* res);  // yt.cu: IAS closure for the current maximum -> basin map
int c2g_yt_weights_device(c2g_basins* res, int idb, double* d_w);  // yt.cu: weights of one basin into a device array
void c2g_fft_free_plans(c2g_context* ctx);
void c2g_slab_bounds(int n3, int nranks, int rank, int* zlo, int* zhi);

// integrate.cu: one streaming pass of per-maximum sums (sums[p*nmax+m]) and counts
// label_mask: 0x7fffffff for Bader labels (bit 31 = filled flag), -1 for YT labels (negative = no basin)
int c2g_launch_basin_reduce(c2g_context* ctx, long long nn, const int* label, int label_mask, int np, const double* const* f,
                            int nmax, double* sums, unsigned long long* counts);

// Device memory of the work buffers.  Every allocation of the library is used on ONE stream (ctx->stream), so a
// released block can be handed out again at once: program order on the stream is the only ordering needed.  The
// blocks are therefore cached per context instead of going back to the driver's stream-ordered pool: with
// cudaFreeAsync + cudaMallocAsync back to back (no host synchronisation between two API calls) the driver was
// measured to map fresh physical memory for multi-GB requests instead of reusing the pending frees (a 31 ms
// c2g_bader_assign took 50..3700 ms).  A request takes the smallest cached block of at least its size and at
// most 1.25 x its size; on an allocation failure the cache is dropped and the request retried.
static inline void c2g_mem_trim(c2g_context* ctx) {
  for (auto& kv : ctx->mem_free) cudaFreeAsync(kv.second, ctx->stream);
  ctx->mem_free.clear();
  ctx->mem_cached_bytes = 0;
}
static inline cudaError_t c2g_alloc(c2g_context* ctx, void** p, size_t bytes) {
  if (bytes == 0) bytes = 1;
  auto it = ctx->mem_free.lower_bound(bytes);
  if (it != ctx->mem_free.end() && it->first <= bytes + bytes / 4 + 256) {
    *p = it->second;
    ctx->mem_live[*p] = it->first;
    ctx->mem_cached_bytes -= it->first;
    ctx->mem_free.erase(it);
    return cudaSuccess;
  }
  cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
  if (e != cudaSuccess) {  // make room and retry once
    cudaGetLastError();
    c2g_mem_trim(ctx);
    cudaStreamSynchronize(ctx->stream);
    e = cudaMallocAsync(p, bytes, ctx->stream);
  }
  if (e == cudaSuccess) ctx->mem_live[*p] = bytes;
  return e;
}
static inline void c2g_release(c2g_context* ctx, void* p) {
  if (!p) return;
  auto it = ctx->mem_live.find(p);
  if (it == ctx->mem_live.end()) { cudaFreeAsync(p, ctx->stream); return; }  // not ours (should not happen)
  ctx->mem_free.emplace(it->second, p);
  ctx->mem_cached_bytes += it->second;
  ctx->mem_live.erase(it);
}
struct DevBuf {
  c2g_context* ctx = nullptr;
  void* p = nullptr;
  DevBuf() {}
  explicit DevBuf(c2g_context* c) : ctx(c) {}
  ~DevBuf() { reset(); }
  void reset() {
    if (p) { if (ctx) c2g_release(ctx, p); else cudaFree(p); }
    p = nullptr;
  }
  cudaError_t alloc(c2g_context* c, size_t bytes) {
    reset();
    ctx = c;
    return c2g_alloc(c, &p, bytes);
  }
  template <class T> T* as() { return (T*)p; }
};

// order the compute stream after the asynchronous upload of a grid (no-op for synchronously uploaded grids)
static inline void c2g_grid_ready(c2g_context* ctx, int h) {
  if (h < 0 || h >= (int)ctx->grids.size()) return;
  c2g_grid& g = ctx->grids[h];
  if (g.pending) { cudaStreamWaitEvent(ctx->stream, g.ready, 0); g.pending = false; }
}
static inline void c2g_grids_ready_all(c2g_context* ctx) {
  for (int h = 0; h < (int)ctx->grids.size(); h++) c2g_grid_ready(ctx, h);
}

static inline int c2g_blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }
