// group.cu -- single-process multi-GPU context (c2g_init(ngpus)).
//
// critic2 is ONE process (src/critic2.F90:24-106): its driver cannot start a rank per GPU or hand an ncclUniqueId
// around.  c2g_init_devices(ngpus) therefore returns one c2g_context that drives `ngpus` devices of the box: it owns
// a per-device context (exactly what c2g_init_multi creates in the one-process-per-GPU model) and one worker
// thread per device, and the entry points of the sharded path -- resident grids, BADER assignment, the basin
// bookkeeping, INTEGRABLE sums / multipoles / attractor images and the NCIPLOT loop -- fan out to the workers, which
// run the same per-rank code (z-slabs, NCCL halo ring, all-reduces) concurrently.  The caller sees whole arrays:
// c2g_grid_upload scatters the z-slabs of ONE host array (every device copies its own slab over PCIe, then the slabs
// are replicated over NVLink), c2g_basins_labels assembles idg(n1,n2,n3) from the slabs, c2g_nci_rdg the (k,j,i)
// arrays from the row ranges.  Paths that do not shard (YT, FFT fields, text codec, HIRSHFELD: "replicas only",
// SURVEY.md 8e) run on the first device, whose grids are complete because every field is replicated.
#include "common.cuh"
#include "group.h"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct c2g_group {
  std::vector<c2g_context*> sub;
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<int(int, c2g_context*)> task;
  unsigned long long epoch = 0;
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
};

namespace {

void worker(c2g_group* G, int r) {
  unsigned long long seen = 0;
  for (;;) {
    std::function<int(int, c2g_context*)> f;
    {
      std::unique_lock<std::mutex> lk(G->mu);
      G->cv_go.wait(lk, [&] { return G->quit || G->epoch != seen; });
      if (G->quit) return;
      seen = G->epoch;
      f = G->task;
    }
    const int rc = f(r, G->sub[r]);
    {
      std::lock_guard<std::mutex> lk(G->mu);
      G->rc[r] = rc;
      if (--G->pending == 0) G->cv_done.notify_all();
    }
  }
}

}  // namespace

// run f(rank, per-device context) on every worker thread concurrently; first non-zero status wins
int c2g_group_run(c2g_context* g, const std::function<int(int, c2g_context*)>& f) {
  c2g_group* G = g->group;
  {
    std::unique_lock<std::mutex> lk(G->mu);
    G->task = f;
    G->pending = (int)G->sub.size();
    G->epoch++;
    G->cv_go.notify_all();
    G->cv_done.wait(lk, [&] { return G->pending == 0; });
  }
  for (size_t r = 0; r < G->sub.size(); r++)
    if (G->rc[r] != C2G_OK) {
      if (G->sub[r]) g->err = "device " + std::to_string(r) + ": " + G->sub[r]->err;
      return G->rc[r];
    }
  return C2G_OK;
}
int c2g_group_size(const c2g_context* g) { return g->group ? (int)g->group->sub.size() : 1; }
c2g_context* c2g_group_sub(c2g_context* g, int r) { return g->group ? g->group->sub[r] : g; }

extern "C" int c2g_init_devices(int ngpus, c2g_context** out) {
  if (!out) return C2G_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return C2G_ERR_CUDA;  // no CPU fallback
  if (ngpus < 1 || ngpus > ndev) return C2G_ERR_ARG;
  if (ngpus == 1) return c2g_init(0, out);
  char uid[128];
  int rc = c2g_nccl_unique_id(uid);
  if (rc != C2G_OK) return rc;
  c2g_context* g = new c2g_context();
  c2g_group* G = new c2g_group();
  g->group = G;
  g->nranks = 1;  // the caller sees one context that owns whole arrays
  G->sub.assign(ngpus, nullptr);
  G->rc.assign(ngpus, C2G_OK);
  for (int r = 0; r < ngpus; r++) G->th.emplace_back(worker, G, r);
  // every worker binds its device and joins the communicator (ncclCommInitRank needs all ranks at once)
  rc = c2g_group_run(g, [&](int r, c2g_context*) -> int {
    c2g_context* s = nullptr;
    const int e = c2g_init_multi(r, r, ngpus, uid, &s);
    if (s) s->parent = g;
    G->sub[r] = s;
    return e;
  });
  if (rc != C2G_OK) {
    c2g_finalize(g);
    return rc;
  }
  g->device = 0;
  g->nsm = G->sub[0]->nsm;
  g->desc = G->sub[0]->desc + " x" + std::to_string(ngpus) + " (one process, z-slabs over NCCL)";
  *out = g;
  return C2G_OK;
}

void c2g_group_finalize(c2g_context* g) {
  c2g_group* G = g->group;
  if (!G) return;
  c2g_group_run(g, [&](int r, c2g_context* s) -> int {
    if (s) c2g_finalize(s);
    G->sub[r] = nullptr;
    return C2G_OK;
  });
  {
    std::lock_guard<std::mutex> lk(G->mu);
    G->quit = true;
    G->cv_go.notify_all();
  }
  for (auto& t : G->th) t.join();
  delete G;
  g->group = nullptr;
  delete g;
}

// ---- resident grids ----
int grp_grid_alloc(c2g_context* g, const int n[3], int* handle) {
  std::vector<int> h(c2g_group_size(g), -1);
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) { return c2g_grid_alloc(s, n, &h[r]); });
  if (rc) return rc;
  for (int v : h)
    if (v != h[0]) return g->fail(C2G_ERR_STATE, "grid handles diverged between the devices");
  *handle = h[0];
  return C2G_OK;
}
int grp_grid_upload(c2g_context* g, const double* f, const int n[3], int* handle) {
  const int G = c2g_group_size(g);
  const size_t plane = (size_t)n[0] * n[1];
  std::vector<int> h(G, -1);
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) {
    int zlo, zhi;
    c2g_slab_bounds(n[2], G, r, &zlo, &zhi);
    return c2g_grid_upload_slab(s, f + plane * zlo, n, &h[r]);  // own slab over PCIe, then NVLink replication
  });
  if (rc) return rc;
  for (int v : h)
    if (v != h[0]) return g->fail(C2G_ERR_STATE, "grid handles diverged between the devices");
  *handle = h[0];
  return C2G_OK;
}
int grp_grid_download(c2g_context* g, int handle, double* f) {
  // every device holds the complete (replicated) field: each one returns its own slab, in parallel
  const int G = c2g_group_size(g);
  return c2g_group_run(g, [&](int r, c2g_context* s) {
    if (handle < 0 || handle >= (int)s->grids.size() || !s->grids[handle].used) return s->fail(C2G_ERR_ARG, "c2g_grid_download: invalid grid handle %d", handle);
    const c2g_grid& gr = s->grids[handle];
    int zlo, zhi;
    c2g_slab_bounds(gr.n[2], G, r, &zlo, &zhi);
    return c2g_grid_download_slab(s, handle, f + (size_t)gr.n[0] * gr.n[1] * zlo);
  });
}
int grp_grid_free(c2g_context* g, int handle) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_grid_free(s, handle); });
}
int grp_grid_promolecular(c2g_context* g, int handle, const double x2c[9], int nat, const double* xat, const double* zat,
                          const double* alpha, int nimg, double rc_) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_grid_promolecular(s, handle, x2c, nat, xat, zat, alpha, nimg, rc_); });
}

// ---- BADER + basin bookkeeping ----
int grp_bader_assign(c2g_context* g, int handle, const double car2lat[9], const double lid[27], int algo, int order, int* nmax,
                     c2g_basins** res_out) {
  const int G = c2g_group_size(g);
  c2g_basins* res = new c2g_basins();
  res->ctx = g;
  res->parts.assign(G, nullptr);
  std::vector<int> nm(G, 0);
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) { return c2g_bader_assign(s, handle, car2lat, lid, algo, order, &nm[r], &res->parts[r]); });
  if (rc == C2G_OK)
    for (int v : nm)
      if (v != nm[0]) rc = g->fail(C2G_ERR_STATE, "the devices disagree on the number of maxima");
  if (rc) {
    c2g_basins_free(res);
    return rc;
  }
  const c2g_basins* p0 = res->parts[0];
  res->kind = p0->kind; res->gridh = handle; res->nn = p0->nn; res->nmax = p0->nmax;
  for (int d = 0; d < 3; d++) res->n[d] = p0->n[d];
  res->zlo = 0; res->zhi = p0->n[2];
  res->max_lin = p0->max_lin;
  for (int i = 0; i < 8; i++) {
    res->stats[i] = 0;
    for (int r = 0; r < G; r++) res->stats[i] = (i == 0 || i == 2 || i == 3 || i == 5) ? res->stats[i] + res->parts[r]->stats[i] : std::max(res->stats[i], res->parts[r]->stats[i]);
  }
  *nmax = nm[0];
  *res_out = res;
  return C2G_OK;
}
void grp_basins_free(c2g_basins* res) {
  c2g_context* g = res->ctx;
  c2g_group_run(g, [&](int r, c2g_context*) {
    if (res->parts[r]) c2g_basins_free(res->parts[r]);
    return C2G_OK;
  });
  delete res;
}
int grp_basins_counts(c2g_basins* res, long long* counts) {
  std::vector<std::vector<long long>> c(res->parts.size(), std::vector<long long>(std::max(res->nmax, 1)));
  int rc = c2g_group_run(res->ctx, [&](int r, c2g_context*) { return c2g_basins_counts(res->parts[r], c[r].data()); });  // all-reduced
  if (rc) return rc;
  for (int i = 0; i < res->nmax; i++) counts[i] = c[0][i];
  return C2G_OK;
}
int grp_basins_set_map(c2g_basins* res, int nattr, const int* map, bool relabel, int nattr0) {
  int rc = c2g_group_run(res->ctx, [&](int r, c2g_context*) {
    return relabel ? c2g_basins_relabel(res->parts[r], nattr0, map, nattr) : c2g_basins_set_map(res->parts[r], nattr, map);
  });
  if (rc) return rc;
  res->nattr = res->parts[0]->nattr;
  res->map = res->parts[0]->map;
  res->has_map = true;
  return C2G_OK;
}
int grp_basins_labels(c2g_basins* res, int* idg) {
  const size_t plane = (size_t)res->n[0] * res->n[1];
  return c2g_group_run(res->ctx, [&](int r, c2g_context*) { return c2g_basins_labels(res->parts[r], idg + plane * res->parts[r]->zlo); });
}

// ---- INTEGRABLE ----
int grp_integrate(c2g_context* g, c2g_basins* res, int nprop, const int* fieldhandles, double omega, double* psum, double* vol) {
  const int G = c2g_group_size(g);
  const size_t np = (size_t)std::max(res->nattr, 1) * std::max(nprop, 1);
  std::vector<std::vector<double>> ps(G, std::vector<double>(np)), vl(G, std::vector<double>(std::max(res->nattr, 1)));
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) { return c2g_integrate(s, res->parts[r], nprop, fieldhandles, omega, ps[r].data(), vol ? vl[r].data() : nullptr); });
  if (rc) return rc;
  // the partial sums are all-reduced on the devices: every rank holds the same numbers
  if (psum) std::copy(ps[0].begin(), ps[0].begin() + (size_t)res->nattr * nprop, psum);
  if (vol) std::copy(vl[0].begin(), vl[0].begin() + res->nattr, vol);
  return C2G_OK;
}
int grp_integrate_multipoles(c2g_context* g, c2g_basins* res, int fieldhandle, int lmax, const double* xattr, const unsigned char* domask,
                             int isortho, int isortho_del, const double x2c[9], const double x2xr[9], const double xr2c[9], int nws,
                             const double* ws, double omega, double* mpole) {
  const int G = c2g_group_size(g);
  const size_t nm = (size_t)(lmax + 1) * (lmax + 1) * std::max(res->nattr, 1);
  std::vector<std::vector<double>> mp(G, std::vector<double>(nm));
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) {
    return c2g_integrate_multipoles(s, res->parts[r], fieldhandle, lmax, xattr, domask, isortho, isortho_del, x2c, x2xr, xr2c, nws, ws, omega, mp[r].data());
  });
  if (rc) return rc;
  std::copy(mp[0].begin(), mp[0].end(), mpole);
  return C2G_OK;
}
int grp_basins_remap(c2g_context* g, c2g_basins* res, const double* xattr, const double c2x[9], int isortho, int isortho_del,
                     const double x2c[9], const double x2xr[9], const double xr2c[9], int nws, const double* ws, int maxattn, int* nattn,
                     int* iatt, int* ilvec, int* idg1) {
  const int G = c2g_group_size(g);
  const size_t plane = (size_t)res->n[0] * res->n[1];
  std::vector<int> na(G, 0);
  std::vector<std::vector<int>> ia(G, std::vector<int>(std::max(maxattn, 1))), il(G, std::vector<int>(3 * (size_t)std::max(maxattn, 1)));
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) {
    return c2g_basins_remap(s, res->parts[r], xattr, c2x, isortho, isortho_del, x2c, x2xr, xr2c, nws, ws, maxattn, &na[r], ia[r].data(), il[r].data(),
                            idg1 ? idg1 + plane * res->parts[r]->zlo : nullptr);
  });
  if (nattn) *nattn = na[0];
  if (rc) return rc;
  std::copy(ia[0].begin(), ia[0].begin() + na[0], iatt);
  std::copy(il[0].begin(), il[0].begin() + 3 * (size_t)na[0], ilvec);
  return C2G_OK;
}

// ---- NCIPLOT: rows [ilo, ihi) of the slowest output index per device, no collective ----
int grp_nci_rdg(c2g_context* g, int handle, const double x0[3], const double xmat[9], const int nstep[3], const double c2x[9],
                const double x2c[9], const double c2xl[9], int nnuc, const double* nuc, double* crho, double* cgrad) {
  const size_t row = (size_t)nstep[2] * nstep[1];
  return c2g_group_run(g, [&](int, c2g_context* s) {
    int ilo, ihi;
    int rc = c2g_nci_range(s, nstep[0], &ilo, &ihi);
    if (rc) return rc;
    return c2g_nci_rdg(s, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc, crho + row * ilo, cgrad + row * ilo);
  });
}

// ---- housekeeping ----
int grp_synchronize(c2g_context* g) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_synchronize(s); });
}
int grp_timer_start(c2g_context* g) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_timer_start(s); });
}
int grp_timer_stop(c2g_context* g, double* ms) {
  std::vector<double> t(c2g_group_size(g), 0.0);
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) { return c2g_timer_stop(s, &t[r]); });
  *ms = *std::max_element(t.begin(), t.end());  // the step is over when the slowest device is done
  return rc;
}
int grp_profile_enable(c2g_context* g, int on) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_profile_enable(s, on); });
}
int grp_profile_reset(c2g_context* g) {
  return c2g_group_run(g, [&](int, c2g_context* s) { return c2g_profile_reset(s); });
}
// per-kernel times: the names of device 0, the maximum over the devices
int grp_profile_get(c2g_context* g, int i, char name[64], double* ms, int* launches) {
  c2g_context* s0 = c2g_group_sub(g, 0);
  if (i < 0 || i >= (int)s0->prof.size()) return C2G_ERR_ARG;
  snprintf(name, 64, "%s", s0->prof[i].name.c_str());
  double m = 0.0;
  for (int r = 0; r < c2g_group_size(g); r++)
    for (const auto& e : c2g_group_sub(g, r)->prof)
      if (e.name == s0->prof[i].name) m = std::max(m, e.ms);
  if (ms) *ms = m;
  if (launches) *launches = s0->prof[i].launches;
  return C2G_OK;
}
long long grp_launch_count(c2g_context* g) {
  long long n = 0;
  for (int r = 0; r < c2g_group_size(g); r++) n += c2g_group_sub(g, r)->launches;
  return n;
}

int grp_nci_rdg_fourier(c2g_context* g, const int h[5], const double x0[3], const double xmat[9], const int nstep[3], const double c2x[9],
                        const double c2xl[9], double* crho, double* cgrad) {
  const size_t row = (size_t)nstep[2] * nstep[1];
  return c2g_group_run(g, [&](int, c2g_context* s) {
    int ilo, ihi;
    int rc = c2g_nci_range(s, nstep[0], &ilo, &ihi);
    if (rc) return rc;
    return c2g_nci_rdg_fourier(s, h, x0, xmat, nstep, c2x, c2xl, crho + row * ilo, cgrad + row * ilo);
  });
}
// replicas: the derived field is computed on every device so that the grid tables stay aligned
int grp_fft_derivative(c2g_context* g, int handle, int iff, const double x2c[9], int* hout) {
  std::vector<int> h(c2g_group_size(g), -1);
  int rc = c2g_group_run(g, [&](int r, c2g_context* s) { return c2g_fft_derivative(s, handle, iff, x2c, &h[r]); });
  if (rc) return rc;
  for (int v : h)
    if (v != h[0]) return g->fail(C2G_ERR_STATE, "grid handles diverged between the devices");
  *hout = h[0];
  return C2G_OK;
}
