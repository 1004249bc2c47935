// fft.cu -- FFT-derived fields (grid3%fft, critic2 src/grid3mod@proc.f90:1757-1872) and the NCIPLOT loop in
// FOURIER mode (src/nci@proc.f90:527-565) on sm_100a.
//
// The reference transforms the real grid as a complex array with cfftnd (forward scaled by 1/ntot,
// cfftnd.f90:33-36), multiplies every coefficient by a function m(G) of the reciprocal vector
//     vgc = i1*bvec(:,1) + i2*bvec(:,2) + i3*bvec(:,3),   i_d in (-n_d/2, n_d/2]          (:1785-1800)
// transforms back and keeps the REAL part (:1860-1866).  Because the input is real its spectrum is
// Hermitian, F(-k) = conj(F(k)), and
//     Re IFFT(m F)  =  IFFT(m_eff F),    m_eff(k) = ( m(k) + conj(m(-k)) ) / 2 ,
// where -k is the index-negated frequency folded back into the reference's range (so the Nyquist
// index n/2 of an even dimension is its own partner and keeps the reference's +n/2 convention).
// m_eff F is Hermitian again, so the whole operation runs on HALF spectra: one cuFFT D2Z, one pointwise
// kernel, one Z2D per output grid -- half the HBM traffic and memory of the complex transform, same result
// up to rounding.  |grad f| (iff = 42) shares one forward transform between its three components.
// cuFFT is resolved with dlopen (like NCCL) so the library loads on machines without it.
#include "common.cuh"
#include "group.h"

#include <cufft.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>

namespace {

struct CufftApi {
  bool ok = false;
  const char* err = "";
  decltype(&cufftPlanMany) PlanMany = nullptr;
  decltype(&cufftPlan3d) Plan3d = nullptr;
  decltype(&cufftSetStream) SetStream = nullptr;
  decltype(&cufftExecD2Z) ExecD2Z = nullptr;
  decltype(&cufftExecZ2D) ExecZ2D = nullptr;
  decltype(&cufftDestroy) Destroy = nullptr;
  decltype(&cufftCreate) Create = nullptr;
  decltype(&cufftSetAutoAllocation) SetAutoAllocation = nullptr;
  decltype(&cufftMakePlan3d) MakePlan3d = nullptr;
  decltype(&cufftSetWorkArea) SetWorkArea = nullptr;
};
// loaded once; thread-safe (a multi-device context calls this from one host thread per device at the same time)
static CufftApi cufft_load() {
  CufftApi api;
  void* h = dlopen("libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libcufft.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("/usr/local/cuda/lib64/libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { api.err = "cannot load libcufft.so.11"; return api; }
#define LOADSYM(field, name)                                  \
  api.field = (decltype(api.field))dlsym(h, name);            \
  if (!api.field) { api.err = "missing symbol " name; return api; }
  LOADSYM(PlanMany, "cufftPlanMany")
  LOADSYM(Plan3d, "cufftPlan3d")
  LOADSYM(SetStream, "cufftSetStream")
  LOADSYM(ExecD2Z, "cufftExecD2Z")
  LOADSYM(ExecZ2D, "cufftExecZ2D")
  LOADSYM(Destroy, "cufftDestroy")
  LOADSYM(Create, "cufftCreate")
  LOADSYM(SetAutoAllocation, "cufftSetAutoAllocation")
  LOADSYM(MakePlan3d, "cufftMakePlan3d")
  LOADSYM(SetWorkArea, "cufftSetWorkArea")
#undef LOADSYM
  api.ok = true;
  return api;
}
CufftApi& cufft_api() {
  static CufftApi api = cufft_load();  // C++11: initialised exactly once, other threads wait
  return api;
}

// Plans are kept per context and grid shape (creating a 1024^3 plan costs tens of ms); their work areas are NOT
// owned by cuFFT: one block from the context's memory cache serves both directions and goes back after the call.
struct FftPlans {
  int n[3];
  cufftHandle pf, pb;
  size_t work;
};
struct FftPlanCache {
  std::vector<FftPlans> plans;
};

struct FftParams {
  int n1, n2, n3, nh;  // nh = n1/2 + 1 stored coefficients along index 1
  double bvec[9];      // column-major, bvec[d + 3*j] = bvec(d+1, j+1)
  double scale;        // 1/ntot (the reference's forward scaling)
  int iff;             // 33..44; for 42 `comp` selects the gradient component
  int comp;
};

__device__ __forceinline__ int ref_freq(int idx, int n) { return idx <= n / 2 ? idx : idx - n; }

// m(k) of the reference for one reciprocal vector v (:1817, :1831-1858); returns (re, im)
__device__ __forceinline__ double2 ref_multiplier(int iff, int comp, const double v[3]) {
  switch (iff) {
    case 33: return make_double2(0.0, -v[0]);
    case 34: return make_double2(0.0, -v[1]);
    case 35: return make_double2(0.0, -v[2]);
    case 36: return make_double2(-v[0] * v[0], 0.0);
    case 37: return make_double2(-v[0] * v[1], 0.0);
    case 38: return make_double2(-v[0] * v[2], 0.0);
    case 39: return make_double2(-v[1] * v[1], 0.0);
    case 40: return make_double2(-v[1] * v[2], 0.0);
    case 41: return make_double2(-v[2] * v[2], 0.0);
    case 42: return make_double2(0.0, v[comp]);  // vgc * cmplx(-aimag(z), dble(z)) = i vgc z
    case 43: return make_double2(-((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), 0.0);
    default: {
      const double v2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
      return make_double2(v2 < 1e-12 ? 0.0 : -1.0 / v2, 0.0);
    }
  }
}

// Y(k) = m_eff(k) * F(k) / ntot over the half spectrum [n3][n2][nh]
__global__ void __launch_bounds__(256) k_fft_multiply(const __grid_constant__ FftParams P, const double2* __restrict__ F,
                                                      double2* __restrict__ Y) {
  const long long total = (long long)P.nh * P.n2 * P.n3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int j1 = (int)(t % P.nh);
    const long long r = t / P.nh;
    const int j2 = (int)(r % P.n2), j3 = (int)(r / P.n2);
    const int i1 = ref_freq(j1, P.n1), i2 = ref_freq(j2, P.n2), i3 = ref_freq(j3, P.n3);
    // partner -k folded into the reference's range
    const int p1 = ref_freq((P.n1 - j1) % P.n1, P.n1), p2 = ref_freq((P.n2 - j2) % P.n2, P.n2),
              p3 = ref_freq((P.n3 - j3) % P.n3, P.n3);
    double v[3], w[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      v[d] = ((double)i1 * P.bvec[d] + (double)i2 * P.bvec[3 + d]) + (double)i3 * P.bvec[6 + d];
      w[d] = ((double)p1 * P.bvec[d] + (double)p2 * P.bvec[3 + d]) + (double)p3 * P.bvec[6 + d];
    }
    const double2 a = ref_multiplier(P.iff, P.comp, v), b = ref_multiplier(P.iff, P.comp, w);
    const double mr = 0.5 * (a.x + b.x) * P.scale, mi = 0.5 * (a.y - b.y) * P.scale;
    const double2 f = F[t];
    Y[t] = make_double2(mr * f.x - mi * f.y, mr * f.y + mi * f.x);
  }
}

// out = y * s (mode 0), out = y^2 (mode 1), out += y^2 (mode 2), out = sqrt(out + y^2) (mode 3)
__global__ void __launch_bounds__(256) k_fft_finish(long long nn, const double* __restrict__ y, double* __restrict__ out,
                                                    int mode, double s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += stride) {
    const double v = y[t];
    double o;
    if (mode == 0) o = v * s;
    else if (mode == 1) o = v * v;
    else if (mode == 2) o = out[t] + v * v;
    else o = sqrt(out[t] + v * v);
    out[t] = o;
  }
}

}  // namespace

extern "C" int c2g_fft_derivative(c2g_context* ctx, int handle, int iff, const double x2c[9], int* hout) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_fft_derivative(ctx, handle, iff, x2c, hout);
  if (!x2c || !hout) return ctx->fail(C2G_ERR_ARG, "c2g_fft_derivative: null argument");
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_fft_derivative: invalid grid handle %d", handle);
  if (iff < C2G_FT_X || iff > C2G_FT_POT) return ctx->fail(C2G_ERR_ARG, "c2g_fft_derivative: unknown derivative code %d", iff);
  // multi-GPU contexts hold the field replicated on every rank: the derived field is computed by every rank on its own
  // copy (replicas only) -- no collective, the result is again a replicated resident grid
  CufftApi& api = cufft_api();
  if (!api.ok) return ctx->fail(C2G_ERR_CUDA, "c2g_fft_derivative: cuFFT unavailable: %s", api.err);
  c2g_grids_ready_all(ctx);
  const int n1 = ctx->grids[handle].n[0], n2 = ctx->grids[handle].n[1], n3 = ctx->grids[handle].n[2];
  const long long nn = ctx->grids[handle].nn;
  cudaStream_t st = ctx->stream;

  FftParams P;
  P.n1 = n1; P.n2 = n2; P.n3 = n3; P.nh = n1 / 2 + 1;
  {  // reciprocal lattice vectors exactly as :1785-1789
    auto cross = [](const double* a, const double* b, double* c) {
      c[0] = a[1] * b[2] - a[2] * b[1];
      c[1] = a[2] * b[0] - a[0] * b[2];
      c[2] = a[0] * b[1] - a[1] * b[0];
    };
    cross(x2c + 6, x2c + 3, P.bvec + 0);
    cross(x2c + 0, x2c + 6, P.bvec + 3);
    cross(x2c + 3, x2c + 0, P.bvec + 6);
    const double det = x2c[0] * (x2c[4] * x2c[8] - x2c[5] * x2c[7]) - x2c[3] * (x2c[1] * x2c[8] - x2c[2] * x2c[7]) +
                       x2c[6] * (x2c[1] * x2c[5] - x2c[2] * x2c[4]);
    if (det == 0.0) return ctx->fail(C2G_ERR_ARG, "c2g_fft_derivative: singular cell matrix");
    const double pi = 3.14159265358979323846264338328;
    for (int q = 0; q < 9; q++) P.bvec[q] = 2.0 * pi / fabs(det) * P.bvec[q];
  }
  P.scale = 1.0 / (double)nn;
  P.iff = iff; P.comp = 0;

  int rc = c2g_grid_alloc(ctx, ctx->grids[handle].n, hout);
  if (rc) return rc;
  const double* src = ctx->grids[handle].d;  // (re-read after the alloc: the grid table may have moved)
  double* out = ctx->grids[*hout].d;

  const size_t nspec = (size_t)P.nh * n2 * n3;
  DevBuf b_F, b_Y, b_y;
  C2G_CUDA(ctx, b_F.alloc(ctx, sizeof(double2) * nspec));
  C2G_CUDA(ctx, b_Y.alloc(ctx, sizeof(double2) * nspec));
  const bool grad = iff == C2G_FT_GRAD;
  if (grad) C2G_CUDA(ctx, b_y.alloc(ctx, sizeof(double) * (size_t)nn));
  // cuFFT is row-major: the slowest dimension first, i.e. (n3, n2, n1) for a Fortran f(n1,n2,n3)
  if (!ctx->fft_cache) ctx->fft_cache = new FftPlanCache();
  FftPlanCache* cache = (FftPlanCache*)ctx->fft_cache;
  FftPlans* pl = nullptr;
  for (auto& q : cache->plans)
    if (q.n[0] == n1 && q.n[1] == n2 && q.n[2] == n3) pl = &q;
  if (!pl) {
    FftPlans q;
    q.n[0] = n1; q.n[1] = n2; q.n[2] = n3;
    size_t wf = 0, wb = 0;
    if (api.Create(&q.pf) != CUFFT_SUCCESS || api.Create(&q.pb) != CUFFT_SUCCESS) return ctx->fail(C2G_ERR_CUDA, "cufftCreate failed");
    api.SetAutoAllocation(q.pf, 0);
    api.SetAutoAllocation(q.pb, 0);
    if (api.MakePlan3d(q.pf, n3, n2, n1, CUFFT_D2Z, &wf) != CUFFT_SUCCESS || api.MakePlan3d(q.pb, n3, n2, n1, CUFFT_Z2D, &wb) != CUFFT_SUCCESS) {
      api.Destroy(q.pf); api.Destroy(q.pb);
      return ctx->fail(C2G_ERR_CUDA, "cufftMakePlan3d failed for %d x %d x %d", n1, n2, n3);
    }
    q.work = std::max(wf, wb);
    api.SetStream(q.pf, st);
    api.SetStream(q.pb, st);
    cache->plans.push_back(q);
    pl = &cache->plans.back();
  }
  const cufftHandle pf = pl->pf, pb = pl->pb;
  DevBuf b_work;
  C2G_CUDA(ctx, b_work.alloc(ctx, std::max<size_t>(pl->work, 16)));
  api.SetWorkArea(pf, b_work.p);
  api.SetWorkArea(pb, b_work.p);

  const int blocks = ctx->nsm * 8;
  ctx->prof_begin("fft_forward_cufft");
  if (api.ExecD2Z(pf, const_cast<double*>(src), (cufftDoubleComplex*)b_F.p) != CUFFT_SUCCESS)
    return ctx->fail(C2G_ERR_CUDA, "cufftExecD2Z failed");
  ctx->prof_end();
  const int ncomp = grad ? 3 : 1;
  for (int c = 0; c < ncomp; c++) {
    P.comp = c;
    ctx->prof_begin("fft_multiply");
    k_fft_multiply<<<blocks, 256, 0, st>>>(P, b_F.as<double2>(), b_Y.as<double2>());
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    double* dst = grad ? b_y.as<double>() : out;
    ctx->prof_begin("fft_backward_cufft");
    if (api.ExecZ2D(pb, (cufftDoubleComplex*)b_Y.p, dst) != CUFFT_SUCCESS) return ctx->fail(C2G_ERR_CUDA, "cufftExecZ2D failed");
    ctx->prof_end();
    if (grad || iff == C2G_FT_POT) {
      const double pi = 3.14159265358979323846264338328;
      ctx->prof_begin("fft_finish");
      k_fft_finish<<<blocks, 256, 0, st>>>(nn, dst, out, grad ? (c == 0 ? 1 : (c == 1 ? 2 : 3)) : 0, -4.0 * pi);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
  }
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->prof_collect();
  return C2G_OK;
}

// called by c2g_finalize
void c2g_fft_free_plans(c2g_context* ctx) {
  if (!ctx->fft_cache) return;
  FftPlanCache* cache = (FftPlanCache*)ctx->fft_cache;
  CufftApi& api = cufft_api();
  if (api.ok)
    for (auto& q : cache->plans) { api.Destroy(q.pf); api.Destroy(q.pb); }
  delete cache;
  ctx->fft_cache = nullptr;
}
