// capi.cu -- context, resident grids, basin bookkeeping and profiling entry points of the C ABI
// declared in include/critic2_gpu.h.
#include "common.cuh"
#include "group.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>

c2g_nccl_api& c2g_nccl() {
  static c2g_nccl_api api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      api.err = "libnccl.so.2 not found";
      return api;
    }
#define C2G_SYM(name)                                             \
  api.name = (decltype(api.name))dlsym(h, "nccl" #name);          \
  if (!api.name) { api.err = "missing NCCL symbol nccl" #name; return api; }
    C2G_SYM(GetUniqueId) C2G_SYM(CommInitRank) C2G_SYM(CommDestroy) C2G_SYM(GetErrorString) C2G_SYM(GroupStart)
    C2G_SYM(GroupEnd) C2G_SYM(Send) C2G_SYM(Recv) C2G_SYM(AllReduce) C2G_SYM(AllGather) C2G_SYM(Broadcast)
#undef C2G_SYM
    api.ok = true;
  }
  return api;
}

namespace {

__global__ void k_flush(float4* buf, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

// labels (index into maxima list, or <0) -> basin ids through map
__global__ void k_map_labels(long long nn, const int* __restrict__ label, const int* __restrict__ map,
                             int* __restrict__ out, int mask) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int l = label[i] & mask;  // Bader labels carry a flag in bit 31; YT uses -1 for IAS points
    out[i] = (l >= 0) ? __ldg(map + l) : 0;
  }
}

// synthetic promolecular-like density; atoms binned on the host into a coarse cell list when rc > 0
struct PromolParams {
  int n1, n2, n3;
  double x2c[9];
  int nat, nimg;
  double rc;
};
__global__ void __launch_bounds__(256) k_promolecular(const __grid_constant__ PromolParams P, const double* __restrict__ xat,
                                                      const double* __restrict__ zat, const double* __restrict__ alpha,
                                                      double* __restrict__ f) {
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int x = (int)(i % P.n1), y = (int)((i / P.n1) % P.n2), z = (int)(i / ((long long)P.n1 * P.n2));
  const double xf0 = (double)x / P.n1, xf1 = (double)y / P.n2, xf2 = (double)z / P.n3;
  double s = 0.0;
  for (int a = 0; a < P.nat; a++) {
    const double za = zat[a], al = alpha[a];
    double b0 = xf0 - xat[3 * a], b1 = xf1 - xat[3 * a + 1], b2 = xf2 - xat[3 * a + 2];
    if (P.rc > 0.0) {  // minimum-image centred sum: only the nearest images can be inside rc
      b0 -= rint(b0); b1 -= rint(b1); b2 -= rint(b2);
    }
    for (int ta = -P.nimg; ta <= P.nimg; ta++)
      for (int tb = -P.nimg; tb <= P.nimg; tb++)
        for (int tc = -P.nimg; tc <= P.nimg; tc++) {
          const double d0 = b0 + ta, d1 = b1 + tb, d2 = b2 + tc;
          const double c0 = P.x2c[0] * d0 + P.x2c[3] * d1 + P.x2c[6] * d2;
          const double c1 = P.x2c[1] * d0 + P.x2c[4] * d1 + P.x2c[7] * d2;
          const double c2 = P.x2c[2] * d0 + P.x2c[5] * d1 + P.x2c[8] * d2;
          const double r = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
          if (P.rc > 0.0) {
            if (r >= P.rc) continue;
            const double u = 1.0 - (r / P.rc) * (r / P.rc);
            s += za * exp(-al * r) * u * u * u;
          } else {
            s += za * exp(-al * r);
          }
        }
  }
  f[i] = s;
}

// cell-list variant for many atoms with a cutoff: atoms sorted by bin, bins are nb^3 fractional boxes
struct PromolBinParams {
  int n1, n2, n3;
  double x2c[9];
  int nb;      // bins per axis
  int reach;   // bins to search on each side
  double rc;
};
__global__ void __launch_bounds__(256) k_promolecular_bins(const __grid_constant__ PromolBinParams P,
                                                           const int* __restrict__ binstart, const double* __restrict__ xat,
                                                           const double* __restrict__ zat, const double* __restrict__ alpha,
                                                           double* __restrict__ f) {
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int x = (int)(i % P.n1), y = (int)((i / P.n1) % P.n2), z = (int)(i / ((long long)P.n1 * P.n2));
  const double xf0 = (double)x / P.n1, xf1 = (double)y / P.n2, xf2 = (double)z / P.n3;
  const int nb = P.nb;
  const int bx = min((int)(xf0 * nb), nb - 1), by = min((int)(xf1 * nb), nb - 1), bz = min((int)(xf2 * nb), nb - 1);
  double s = 0.0;
  for (int dz = -P.reach; dz <= P.reach; dz++)
    for (int dy = -P.reach; dy <= P.reach; dy++)
      for (int dx = -P.reach; dx <= P.reach; dx++) {
        const int cx = ((bx + dx) % nb + nb) % nb, cy = ((by + dy) % nb + nb) % nb, cz = ((bz + dz) % nb + nb) % nb;
        const int b = cx + nb * (cy + nb * cz);
        for (int a = binstart[b]; a < binstart[b + 1]; a++) {
          double d0 = xf0 - xat[3 * a], d1 = xf1 - xat[3 * a + 1], d2 = xf2 - xat[3 * a + 2];
          d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);
          const double c0 = P.x2c[0] * d0 + P.x2c[3] * d1 + P.x2c[6] * d2;
          const double c1 = P.x2c[1] * d0 + P.x2c[4] * d1 + P.x2c[7] * d2;
          const double c2 = P.x2c[2] * d0 + P.x2c[5] * d1 + P.x2c[8] * d2;
          const double r = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
          if (r >= P.rc) continue;
          const double u = 1.0 - (r / P.rc) * (r / P.rc);
          s += zat[a] * exp(-alpha[a] * r) * u * u * u;
        }
      }
  f[i] = s;
}

int check_handle(c2g_context* ctx, int h, const char* who) {
  if (h < 0 || h >= (int)ctx->grids.size() || !ctx->grids[h].used)
    return ctx->fail(C2G_ERR_ARG, "%s: invalid grid handle %d", who, h);
  c2g_grid_ready(ctx, h);  // order the compute stream after an asynchronous upload of this grid
  return C2G_OK;
}

// copy streams and the ordering event, created on first use
int ensure_copy_streams(c2g_context* ctx) {
  if (!ctx->copy_in) {
    C2G_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    C2G_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    C2G_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
  }
  return C2G_OK;
}

}  // namespace

extern "C" {

int c2g_init(int device, c2g_context** out) {
  if (!out) return C2G_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return C2G_ERR_CUDA;  // no CPU fallback: fail loudly
  if (device < 0 || device >= ndev) return C2G_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return C2G_ERR_CUDA;
  c2g_context* ctx = new c2g_context();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return C2G_ERR_CUDA; }
  ctx->nsm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return C2G_ERR_CUDA; }
  {  // keep freed blocks in the stream-ordered pool (no trimming at synchronisation points)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  char buf[256];
  snprintf(buf, sizeof(buf), "critic2_gpu 0.1 sm_%d%d %s %d SMs %.1f GB", prop.major, prop.minor, prop.name,
           prop.multiProcessorCount, prop.totalGlobalMem / 1e9);
  ctx->desc = buf;
  *out = ctx;
  return C2G_OK;
}

int c2g_nccl_unique_id(void* uid128) {
  if (!uid128) return C2G_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!c2g_nccl().ok) return C2G_ERR_NCCL;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return C2G_ERR_NCCL;
  memcpy(uid128, &id, 128);
  return C2G_OK;
}

int c2g_init_multi(int device, int rank, int nranks, const void* uid128, c2g_context** out) {
  int rc = c2g_init(device, out);
  if (rc != C2G_OK) return rc;
  c2g_context* ctx = *out;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks > 1) {
    if (!uid128) return ctx->fail(C2G_ERR_ARG, "c2g_init_multi: null nccl id");
    if (!c2g_nccl().ok) return ctx->fail(C2G_ERR_NCCL, "c2g_init_multi: %s", c2g_nccl().err);
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    ncclComm_t comm;
    ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return ctx->fail(C2G_ERR_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
    ctx->nccl = (void*)comm;
  }
  return C2G_OK;
}

void c2g_finalize(c2g_context* ctx) {
  if (!ctx) return;
  if (ctx->group) { c2g_group_finalize(ctx); return; }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_in) {
    cudaStreamSynchronize(ctx->copy_in);
    cudaStreamSynchronize(ctx->copy_out);
    cudaStreamDestroy(ctx->copy_in);
    cudaStreamDestroy(ctx->copy_out);
    cudaEventDestroy(ctx->ev_order);
  }
  for (void* p : ctx->deferred) c2g_release(ctx, p);
  ctx->deferred.clear();
  c2g_fft_free_plans(ctx);
  if (ctx->hpin) { cudaFreeHost(ctx->hpin); ctx->hpin = nullptr; }
  for (auto& g : ctx->grids) {
    if (g.ready) cudaEventDestroy(g.ready);
    if (g.used && g.d) c2g_release(ctx, g.d);
  }
  c2g_mem_trim(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->flushbuf) cudaFree(ctx->flushbuf);
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  for (auto& p : ctx->pending) { cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second); }
  if (ctx->nccl) ncclCommDestroy((ncclComm_t)ctx->nccl);
  if (ctx->t0) { cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* c2g_last_error(const c2g_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
const char* c2g_describe(c2g_context* ctx) { return ctx ? ctx->desc.c_str() : ""; }

int c2g_grid_alloc(c2g_context* ctx, const int n[3], int* handle) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_grid_alloc(ctx, n, handle);
  if (!n || !handle || n[0] < 1 || n[1] < 1 || n[2] < 1) return ctx->fail(C2G_ERR_ARG, "c2g_grid_alloc: bad shape");
  c2g_grid g;
  g.n[0] = n[0]; g.n[1] = n[1]; g.n[2] = n[2];
  g.nn = (long long)n[0] * n[1] * n[2];
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&g.d, sizeof(double) * g.nn));
  g.used = true;
  int h = -1;
  for (size_t i = 0; i < ctx->grids.size(); i++)
    if (!ctx->grids[i].used) { h = (int)i; break; }
  if (h < 0) { ctx->grids.push_back(g); h = (int)ctx->grids.size() - 1; }
  else ctx->grids[h] = g;
  *handle = h;
  return C2G_OK;
}

int c2g_grid_upload(c2g_context* ctx, const double* f, const int n[3], int* handle) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return f ? grp_grid_upload(ctx, f, n, handle) : ctx->fail(C2G_ERR_ARG, "c2g_grid_upload: null field");
  if (!f) return ctx->fail(C2G_ERR_ARG, "c2g_grid_upload: null field");
  int rc = c2g_grid_alloc(ctx, n, handle);
  if (rc != C2G_OK) return rc;
  c2g_grid& g = ctx->grids[*handle];
  C2G_CUDA(ctx, cudaMemcpyAsync(g.d, f, sizeof(double) * g.nn, cudaMemcpyHostToDevice, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C2G_OK;
}

// Asynchronous upload: the copy runs on a separate stream and overlaps the kernels of other calls; every call that
// reads the grid is ordered after it on the device.  `f` must be page-locked for a real overlap and must stay
// valid and unchanged until c2g_synchronize (or any call that returns results computed from this grid).
int c2g_grid_upload_async(c2g_context* ctx, const double* f, const int n[3], int* handle) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return c2g_grid_upload(ctx, f, n, handle);  // multi-device: the slab scatter is synchronous
  if (!f) return ctx->fail(C2G_ERR_ARG, "c2g_grid_upload_async: null field");
  int rc = ensure_copy_streams(ctx);
  if (rc != C2G_OK) return rc;
  rc = c2g_grid_alloc(ctx, n, handle);
  if (rc != C2G_OK) return rc;
  c2g_grid& g = ctx->grids[*handle];
  // the block may come from the cache and still be read by work queued on the compute stream
  C2G_CUDA(ctx, cudaEventRecord(ctx->ev_order, ctx->stream));
  C2G_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_order, 0));
  C2G_CUDA(ctx, cudaMemcpyAsync(g.d, f, sizeof(double) * g.nn, cudaMemcpyHostToDevice, ctx->copy_in));
  if (!g.ready) C2G_CUDA(ctx, cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming));
  C2G_CUDA(ctx, cudaEventRecord(g.ready, ctx->copy_in));
  g.pending = true;
  return C2G_OK;
}

int c2g_slab_bounds_query(int n3, int nranks, int rank, int* zlo, int* zhi) {
  if (!zlo || !zhi || n3 < 1 || nranks < 1 || rank < 0 || rank >= nranks) return C2G_ERR_ARG;
  c2g_slab_bounds(n3, nranks, rank, zlo, zhi);
  return C2G_OK;
}

int c2g_slab_range(c2g_context* ctx, int n3, int* zlo, int* zhi) {
  if (!ctx || !zlo || !zhi || n3 < 1) return C2G_ERR_ARG;
  c2g_slab_bounds(n3, ctx->nranks, ctx->rank, zlo, zhi);  // a multi-device context has nranks = 1: the caller holds whole arrays
  return C2G_OK;
}

int c2g_grid_upload_slab(c2g_context* ctx, const double* fslab, const int n[3], int* handle) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return c2g_grid_upload(ctx, fslab, n, handle);  // one process: the caller holds the whole array
  int rc = c2g_grid_alloc(ctx, n, handle);
  if (rc != C2G_OK) return rc;
  c2g_grid& g = ctx->grids[*handle];
  const size_t plane = (size_t)n[0] * n[1];
  int zlo, zhi;
  c2g_slab_bounds(n[2], ctx->nranks, ctx->rank, &zlo, &zhi);
  if (zhi > zlo) {
    if (!fslab) return ctx->fail(C2G_ERR_ARG, "c2g_grid_upload_slab: null slab");
    C2G_CUDA(ctx, cudaMemcpyAsync(g.d + plane * zlo, fslab, sizeof(double) * plane * (zhi - zlo), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (ctx->nranks > 1) {
    // replicate: every rank broadcasts its slab (NVLink); rho is read-only afterwards
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    ctx->prof_begin("grid_allgather_nccl");
    ncclGroupStart();
    for (int r = 0; r < ctx->nranks; r++) {
      int a, b;
      c2g_slab_bounds(n[2], ctx->nranks, r, &a, &b);
      if (b > a) ncclBroadcast(g.d + plane * a, g.d + plane * a, plane * (b - a), ncclDouble, r, comm, ctx->stream);
    }
    ncclResult_t nr = ncclGroupEnd();
    ctx->prof_end();
    if (nr != ncclSuccess) return ctx->fail(C2G_ERR_NCCL, "c2g_grid_upload_slab: %s", ncclGetErrorString(nr));
  }
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_collect();
  return C2G_OK;
}

int c2g_grid_download(c2g_context* ctx, int handle, double* f) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return f ? grp_grid_download(ctx, handle, f) : ctx->fail(C2G_ERR_ARG, "c2g_grid_download: null output");
  int rc = check_handle(ctx, handle, "c2g_grid_download");
  if (rc) return rc;
  if (!f) return ctx->fail(C2G_ERR_ARG, "c2g_grid_download: null output");
  c2g_grid& g = ctx->grids[handle];
  C2G_CUDA(ctx, cudaMemcpyAsync(f, g.d, sizeof(double) * g.nn, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C2G_OK;
}

int c2g_grid_download_slab(c2g_context* ctx, int handle, double* fslab) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return c2g_grid_download(ctx, handle, fslab);
  int rc = check_handle(ctx, handle, "c2g_grid_download_slab");
  if (rc) return rc;
  c2g_grid& g = ctx->grids[handle];
  int zlo, zhi;
  c2g_slab_bounds(g.n[2], ctx->nranks, ctx->rank, &zlo, &zhi);
  if (zhi <= zlo) return C2G_OK;
  if (!fslab) return ctx->fail(C2G_ERR_ARG, "c2g_grid_download_slab: null output");
  const size_t plane = (size_t)g.n[0] * g.n[1];
  C2G_CUDA(ctx, cudaMemcpyAsync(fslab, g.d + plane * zlo, sizeof(double) * plane * (zhi - zlo), cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return C2G_OK;
}

int c2g_grid_free(c2g_context* ctx, int handle) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_grid_free(ctx, handle);
  int rc = check_handle(ctx, handle, "c2g_grid_free");
  if (rc) return rc;
  c2g_grid& g = ctx->grids[handle];
  c2g_release(ctx, g.d);  // (check_handle has ordered the compute stream after a pending upload)
  if (g.ready) cudaEventDestroy(g.ready);
  g = c2g_grid();
  return C2G_OK;
}

int c2g_grid_promolecular(c2g_context* ctx, int handle, const double x2c[9], int nat, const double* xat,
                          const double* zat, const double* alpha, int nimg, double rc) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_grid_promolecular(ctx, handle, x2c, nat, xat, zat, alpha, nimg, rc);
  int r = check_handle(ctx, handle, "c2g_grid_promolecular");
  if (r) return r;
  if (!x2c || nat < 1 || !xat || !zat || !alpha || nimg < 0) return ctx->fail(C2G_ERR_ARG, "c2g_grid_promolecular: bad argument");
  c2g_grid& g = ctx->grids[handle];
  cudaStream_t st = ctx->stream;
  double *dx = nullptr, *dz = nullptr, *da = nullptr;
  int* dbin = nullptr;
  // decide on bins: only with a cutoff and an (approximately) orthogonal cell large against rc
  int nb = 0, reach = 0;
  if (rc > 0.0 && nat >= 64) {
    // perpendicular widths of the cell
    const double a[3] = {x2c[0], x2c[1], x2c[2]}, b[3] = {x2c[3], x2c[4], x2c[5]}, c[3] = {x2c[6], x2c[7], x2c[8]};
    auto cross = [](const double* u, const double* v, double* w) {
      w[0] = u[1] * v[2] - u[2] * v[1]; w[1] = u[2] * v[0] - u[0] * v[2]; w[2] = u[0] * v[1] - u[1] * v[0];
    };
    auto nrm = [](const double* u) { return std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); };
    double bc[3], ca[3], ab[3];
    cross(b, c, bc); cross(c, a, ca); cross(a, b, ab);
    const double vol = std::fabs(a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2]);
    const double wmin = std::min(vol / nrm(bc), std::min(vol / nrm(ca), vol / nrm(ab)));
    nb = (int)std::floor(wmin / rc * 2.0);  // bin width >= rc/2
    if (nb >= 5) reach = 2 + 0;  // bin width w >= rc/2  => atoms within rc are at most ceil(rc/w) = 2 bins away
    else nb = 0;
    if (nb > 64) { nb = 64; }
    if (nb) {
      const double w = wmin / nb;
      reach = (int)std::ceil(rc / w);
      if (2 * reach + 1 > nb) nb = 0;
    }
  }
  std::vector<double> sx(3 * (size_t)nat), sz(nat), sa(nat);
  std::vector<int> binstart;
  if (nb) {
    std::vector<int> bin(nat), order(nat);
    for (int i = 0; i < nat; i++) {
      int q[3];
      for (int d = 0; d < 3; d++) {
        double v = xat[3 * i + d] - std::floor(xat[3 * i + d]);
        q[d] = std::min((int)(v * nb), nb - 1);
      }
      bin[i] = q[0] + nb * (q[1] + nb * q[2]);
      order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int u, int v) { return bin[u] < bin[v]; });
    binstart.assign((size_t)nb * nb * nb + 1, 0);
    for (int i = 0; i < nat; i++) binstart[bin[i] + 1]++;
    for (size_t i = 1; i < binstart.size(); i++) binstart[i] += binstart[i - 1];
    for (int k = 0; k < nat; k++) {
      const int i = order[k];
      for (int d = 0; d < 3; d++) sx[3 * k + d] = xat[3 * i + d] - std::floor(xat[3 * i + d]);
      sz[k] = zat[i]; sa[k] = alpha[i];
    }
  } else {
    std::copy(xat, xat + 3 * (size_t)nat, sx.begin());
    std::copy(zat, zat + nat, sz.begin());
    std::copy(alpha, alpha + nat, sa.begin());
  }
  C2G_CUDA(ctx, cudaMalloc(&dx, sizeof(double) * 3 * nat));
  C2G_CUDA(ctx, cudaMalloc(&dz, sizeof(double) * nat));
  C2G_CUDA(ctx, cudaMalloc(&da, sizeof(double) * nat));
  C2G_CUDA(ctx, cudaMemcpyAsync(dx, sx.data(), sizeof(double) * 3 * nat, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(dz, sz.data(), sizeof(double) * nat, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(da, sa.data(), sizeof(double) * nat, cudaMemcpyHostToDevice, st));
  ctx->prof_begin("promolecular");
  if (nb) {
    C2G_CUDA(ctx, cudaMalloc(&dbin, sizeof(int) * binstart.size()));
    C2G_CUDA(ctx, cudaMemcpyAsync(dbin, binstart.data(), sizeof(int) * binstart.size(), cudaMemcpyHostToDevice, st));
    PromolBinParams P;
    P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
    memcpy(P.x2c, x2c, sizeof(P.x2c));
    P.nb = nb; P.reach = reach; P.rc = rc;
    k_promolecular_bins<<<c2g_blocks_for(g.nn, 256), 256, 0, st>>>(P, dbin, dx, dz, da, g.d);
  } else {
    PromolParams P;
    P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
    memcpy(P.x2c, x2c, sizeof(P.x2c));
    P.nat = nat; P.nimg = nimg; P.rc = rc;
    k_promolecular<<<c2g_blocks_for(g.nn, 256), 256, 0, st>>>(P, dx, dz, da, g.d);
  }
  ctx->prof_end();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(dx); cudaFree(dz); cudaFree(da);
  if (dbin) cudaFree(dbin);
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "c2g_grid_promolecular: %s", cudaGetErrorString(e));
  ctx->prof_collect();
  return C2G_OK;
}

// ---- basins bookkeeping ----
int c2g_basins_maxima(c2g_basins* res, int* pmax) {
  if (!res || !pmax) return C2G_ERR_ARG;
  for (int i = 0; i < res->nmax; i++) {
    const int id = res->max_lin[i];
    pmax[3 * i + 0] = id % res->n[0] + 1;
    pmax[3 * i + 1] = (id / res->n[0]) % res->n[1] + 1;
    pmax[3 * i + 2] = id / (res->n[0] * res->n[1]) + 1;
  }
  return C2G_OK;
}

int c2g_basins_counts(c2g_basins* res, long long* counts) {
  if (!res || !counts) return C2G_ERR_ARG;
  if (!res->parts.empty()) return grp_basins_counts(res, counts);
  c2g_context* ctx = res->ctx;
  if (res->kind != 0) return ctx->fail(C2G_ERR_STATE, "c2g_basins_counts: Bader results only");
  if (res->counts.empty() && res->nmax > 0) {  // one 4 B/pt pass, on demand
    unsigned long long* d_counts = nullptr;
    C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_counts, sizeof(unsigned long long) * res->nmax));
    C2G_CUDA(ctx, cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * res->nmax, ctx->stream));
    const long long nnl = (long long)res->n[0] * res->n[1] * (res->zhi - res->zlo);
    int rc = nnl > 0 ? c2g_launch_basin_reduce(ctx, nnl, res->d_label, 0x7fffffff, 0, nullptr, res->nmax, nullptr, d_counts) : C2G_OK;
    if (rc == C2G_OK && ctx->nranks > 1 &&
        ncclAllReduce(d_counts, d_counts, res->nmax, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
      rc = ctx->fail(C2G_ERR_NCCL, "c2g_basins_counts: ncclAllReduce failed");
    std::vector<unsigned long long> hc(res->nmax);
    cudaError_t e = cudaMemcpyAsync(hc.data(), d_counts, sizeof(unsigned long long) * res->nmax, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    c2g_release(ctx, d_counts);
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "c2g_basins_counts: %s", cudaGetErrorString(e));
    res->counts.assign(hc.begin(), hc.end());
    ctx->prof_collect();
  }
  for (int i = 0; i < res->nmax; i++) counts[i] = res->counts[i];
  return C2G_OK;
}

// reclose: YT only -- the map is the attractor identification of the sweep itself (yt@proc.f90:129-168), so the
// interior / IAS classification follows it; false for the later relabelling of int_reorder_gridout, which renames
// basins and leaves the IAS points alone (integration@proc.f90:1139-1144)
static int set_map_impl(c2g_basins* res, int nattr, const int* map, bool reclose);
int c2g_basins_set_map(c2g_basins* res, int nattr, const int* map) { return set_map_impl(res, nattr, map, true); }
static int set_map_impl(c2g_basins* res, int nattr, const int* map, bool reclose) {
  if (!res) return C2G_ERR_ARG;
  if (!res->parts.empty()) return map ? grp_basins_set_map(res, nattr, map, false, 0) : C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (!map || nattr < 0) return ctx->fail(C2G_ERR_ARG, "c2g_basins_set_map: bad argument");
  for (int i = 0; i < res->nmax; i++)
    if (map[i] < 0 || map[i] > nattr) return ctx->fail(C2G_ERR_ARG, "c2g_basins_set_map: map(%d)=%d out of range", i + 1, map[i]);
  res->map.assign(map, map + res->nmax);
  res->nattr = nattr;
  if (!res->d_map) C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&res->d_map, sizeof(int) * std::max(res->nmax, 1)));
  C2G_CUDA(ctx, cudaMemcpyAsync(res->d_map, res->map.data(), sizeof(int) * res->nmax, cudaMemcpyHostToDevice, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  res->has_map = true;
  if (res->kind == 1 && reclose) return c2g_yt_apply_map(res);
  return C2G_OK;
}

int c2g_basins_nattr(c2g_basins* res, int* nattr) {
  if (!res || !nattr) return C2G_ERR_ARG;
  if (!res->has_map) return res->ctx->fail(C2G_ERR_STATE, "c2g_basins_nattr: no map set");
  *nattr = res->nattr;
  return C2G_OK;
}

int c2g_basins_relabel(c2g_basins* res, int nattr0, const int* assigned, int nattr_new) {
  if (!res) return C2G_ERR_ARG;
  if (!res->parts.empty()) return assigned ? grp_basins_set_map(res, nattr_new, assigned, true, nattr0) : C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_basins_relabel: no map set");
  if (nattr0 != res->nattr || !assigned) return ctx->fail(C2G_ERR_ARG, "c2g_basins_relabel: nattr0 mismatch");
  std::vector<int> m(res->nmax);
  for (int i = 0; i < res->nmax; i++) {
    const int old = res->map[i];
    m[i] = old > 0 ? assigned[old - 1] : 0;
  }
  return set_map_impl(res, nattr_new, m.data(), false);
}

int c2g_basins_labels(c2g_basins* res, int* idg) {
  if (!res) return C2G_ERR_ARG;
  if (!res->parts.empty()) return idg ? grp_basins_labels(res, idg) : C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (!idg) return ctx->fail(C2G_ERR_ARG, "c2g_basins_labels: null output");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_basins_labels: call c2g_basins_set_map first");
  // multi-GPU: every rank returns its own z-slab idg(:,:,zlo+1:zhi)
  const long long nnl = (long long)res->n[0] * res->n[1] * (res->zhi - res->zlo);
  if (nnl == 0) return C2G_OK;
  int* d_out = nullptr;
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_out, sizeof(int) * nnl));
  ctx->prof_begin("map_labels");
  k_map_labels<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(nnl, res->d_label, res->d_map, d_out, res->kind == 0 ? 0x7fffffff : -1);
  ctx->prof_end();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(idg, d_out, sizeof(int) * nnl, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  c2g_release(ctx, d_out);
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "c2g_basins_labels: %s", cudaGetErrorString(e));
  ctx->prof_collect();
  return C2G_OK;
}

// Asynchronous variant: the label map runs on the compute stream, the device-to-host copy on a separate stream, so
// that it overlaps later calls (e.g. c2g_integrate, or the upload of the next field).  `idg` must be page-locked for
// a real overlap; it is complete after c2g_synchronize.
int c2g_basins_labels_async(c2g_basins* res, int* idg) {
  if (!res) return C2G_ERR_ARG;
  if (!res->parts.empty()) return c2g_basins_labels(res, idg);
  c2g_context* ctx = res->ctx;
  if (!idg) return ctx->fail(C2G_ERR_ARG, "c2g_basins_labels_async: null output");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_basins_labels_async: call c2g_basins_set_map first");
  int rc = ensure_copy_streams(ctx);
  if (rc != C2G_OK) return rc;
  const long long nnl = (long long)res->n[0] * res->n[1] * (res->zhi - res->zlo);
  if (nnl == 0) return C2G_OK;
  int* d_out = nullptr;
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_out, sizeof(int) * nnl));
  ctx->prof_begin("map_labels");
  k_map_labels<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(nnl, res->d_label, res->d_map, d_out, res->kind == 0 ? 0x7fffffff : -1);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  C2G_CUDA(ctx, cudaEventRecord(ctx->ev_order, ctx->stream));
  C2G_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_order, 0));
  C2G_CUDA(ctx, cudaMemcpyAsync(idg, d_out, sizeof(int) * nnl, cudaMemcpyDeviceToHost, ctx->copy_out));
  ctx->deferred.push_back(d_out);  // handed back to the cache by c2g_synchronize, once the copy is done
  return C2G_OK;
}

// indicator field of one basin: w = 1 where map(label) == idb, else 0 (int_cubew, integration@proc.f90:4455-4458)
__global__ void k_indicator(long long nn, const int* __restrict__ label, const int* __restrict__ map, int mask, int idb,
                            double* __restrict__ w) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int l = label[i] & mask;
    w[i] = (l >= 0 && __ldg(map + l) == idb) ? 1.0 : 0.0;
  }
}

int c2g_basins_weight_grid(c2g_basins* res, int idb, int* handle) {
  if (!res) return C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (!handle) return ctx->fail(C2G_ERR_ARG, "c2g_basins_weight_grid: null handle");
  if (!res->parts.empty() || ctx->parent)
    return ctx->fail(C2G_ERR_STATE, "c2g_basins_weight_grid: not available on a multi-device context (labels are sharded; use a single-device context for WCUBE)");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_basins_weight_grid: call c2g_basins_set_map first");
  if (idb < 1 || idb > res->nattr) return ctx->fail(C2G_ERR_ARG, "c2g_basins_weight_grid: unknown basin %d", idb);
  if (ctx->nranks > 1 && res->kind == 0)
    return ctx->fail(C2G_ERR_STATE, "c2g_basins_weight_grid: Bader labels are sharded on a multi-GPU context");
  int rc = c2g_grid_alloc(ctx, res->n, handle);
  if (rc != C2G_OK) return rc;
  double* w = ctx->grids[*handle].d;
  if (res->kind == 1) {
    rc = c2g_yt_weights_device(res, idb, w);
  } else {
    ctx->prof_begin("basin_indicator");
    k_indicator<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(res->nn, res->d_label, res->d_map, res->kind == 0 ? 0x7fffffff : -1, idb, w);
    ctx->prof_end();
    if (cudaGetLastError() != cudaSuccess) rc = ctx->fail(C2G_ERR_CUDA, "k_indicator launch failed");
  }
  if (rc != C2G_OK) { c2g_grid_free(ctx, *handle); *handle = -1; return rc; }
  return C2G_OK;
}

int c2g_basins_stats(c2g_basins* res, long long stats[8]) {
  if (!res || !stats) return C2G_ERR_ARG;
  for (int i = 0; i < 8; i++) stats[i] = res->stats[i];
  return C2G_OK;
}

void c2g_basins_free(c2g_basins* res) {
  if (!res) return;
  if (!res->parts.empty()) { grp_basins_free(res); return; }
  if (res->d_lbuf) c2g_release(res->ctx, res->d_lbuf);
  else if (res->d_label) c2g_release(res->ctx, res->d_label);
  if (res->d_map) c2g_release(res->ctx, res->d_map);
  if (res->d_vec) cudaFree(res->d_vec);
  if (res->d_area) cudaFree(res->d_area);
  c2g_yt_free_state(res);
  delete res;
}

// ---- profiling ----
int c2g_profile_enable(c2g_context* ctx, int on) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_profile_enable(ctx, on);
  ctx->prof_on = on != 0;
  return C2G_OK;
}
int c2g_profile_count(c2g_context* ctx) { return ctx ? (int)(ctx->group ? c2g_group_sub(ctx, 0) : ctx)->prof.size() : 0; }
int c2g_profile_get(c2g_context* ctx, int i, char name[64], double* ms, int* launches) {
  if (ctx && ctx->group) return grp_profile_get(ctx, i, name, ms, launches);
  if (!ctx || i < 0 || i >= (int)ctx->prof.size()) return C2G_ERR_ARG;
  snprintf(name, 64, "%s", ctx->prof[i].name.c_str());
  if (ms) *ms = ctx->prof[i].ms;
  if (launches) *launches = ctx->prof[i].launches;
  return C2G_OK;
}
int c2g_profile_reset(c2g_context* ctx) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_profile_reset(ctx);
  cudaStreamSynchronize(ctx->stream);
  ctx->prof_collect();
  ctx->prof.clear();
  return C2G_OK;
}
long long c2g_launch_count(c2g_context* ctx) { return ctx ? (ctx->group ? grp_launch_count(ctx) : ctx->launches) : 0; }

int c2g_flush_l2(c2g_context* ctx) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return c2g_group_run(ctx, [](int, c2g_context* s) { return c2g_flush_l2(s); });
  if (!ctx->flushbuf) {
    ctx->flushbytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    C2G_CUDA(ctx, cudaMalloc(&ctx->flushbuf, ctx->flushbytes));
  }
  k_flush<<<ctx->nsm * 4, 256, 0, ctx->stream>>>((float4*)ctx->flushbuf, ctx->flushbytes / 16);
  C2G_KERNEL_CHECK(ctx);
  return C2G_OK;
}

int c2g_timer_start(c2g_context* ctx) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_timer_start(ctx);
  if (!ctx->t0) { C2G_CUDA(ctx, cudaEventCreate(&ctx->t0)); C2G_CUDA(ctx, cudaEventCreate(&ctx->t1)); }
  C2G_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
  return C2G_OK;
}

int c2g_timer_stop(c2g_context* ctx, double* ms) {
  if (ctx && ctx->group) return ms ? grp_timer_stop(ctx, ms) : C2G_ERR_ARG;
  if (!ctx || !ms || !ctx->t0) return C2G_ERR_ARG;
  C2G_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
  C2G_CUDA(ctx, cudaEventSynchronize(ctx->t1));
  float f = 0.f;
  C2G_CUDA(ctx, cudaEventElapsedTime(&f, ctx->t0, ctx->t1));
  *ms = f;
  return C2G_OK;
}

int c2g_synchronize(c2g_context* ctx) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_synchronize(ctx);
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_in) {
    C2G_CUDA(ctx, cudaStreamSynchronize(ctx->copy_in));
    C2G_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
  }
  for (void* p : ctx->deferred) c2g_release(ctx, p);
  ctx->deferred.clear();
  ctx->prof_collect();
  return C2G_OK;
}

}  // extern "C"
