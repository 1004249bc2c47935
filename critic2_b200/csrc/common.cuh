// common.cuh -- context, resident grids, error handling and profiling shared by the kernels.
#pragma once

#include <cuda_runtime.h>

#include "nccl_shim.h"

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/critic2_gpu.h"

#define C2G_NSM_FALLBACK 148

struct c2g_grid {
  int n[3] = {0, 0, 0};
  long long nn = 0;
  double* d = nullptr;  // device, f(n1,n2,n3) index 1 fastest
  bool used = false;
};

struct c2g_prof_entry {
  std::string name;
  double ms = 0.0;
  int launches = 0;
};

struct c2g_context {
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
  c2g_context* ctx = nullptr;
  int kind = 0;  // 0 = bader, 1 = yt
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
void c2g_slab_bounds(int n3, int nranks, int rank, int* zlo, int* zhi);

// integrate.cu: one streaming pass of per-maximum sums (sums[p*nmax+m]) and counts
// label_mask: 0x7fffffff for Bader labels (bit 31 = filled flag), -1 for YT labels (negative = no basin)
int c2g_launch_basin_reduce(c2g_context* ctx, long long nn, const int* label, int label_mask, int np, const double* const* f,
                            int nmax, double* sums, unsigned long long* counts);

// stream-ordered pool allocations (cudaMallocAsync with an unbounded release threshold, set in c2g_init):
// repeated calls reuse the same device memory instead of paying cudaMalloc/cudaFree every time
static inline cudaError_t c2g_alloc(c2g_context* ctx, void** p, size_t bytes) {
  return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
}
static inline void c2g_release(c2g_context* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}
struct DevBuf {
  c2g_context* ctx = nullptr;
  void* p = nullptr;
  DevBuf() {}
  explicit DevBuf(c2g_context* c) : ctx(c) {}
  ~DevBuf() { reset(); }
  void reset() {
    if (p) { if (ctx) cudaFreeAsync(p, ctx->stream); else cudaFree(p); }
    p = nullptr;
  }
  cudaError_t alloc(c2g_context* c, size_t bytes) {
    reset();
    ctx = c;
    return c2g_alloc(c, &p, bytes);
  }
  template <class T> T* as() { return (T*)p; }
};

static inline int c2g_blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }
