// multipole.cu -- basin multipole moments of an INTEGRABLE field (INTEGRABLE id MULTIPOLES [lmax]).
//
// Replaces the multipole branch of intgrid_fields (critic2 src/integration@proc.f90:1302-1361):
//   Bader / isosurface (:1338-1358): for every grid point p with basin ix = idg(p):
//        dv = p/n - xattr(:,ix); shortest(dv); tosphere; genrlm_real(lmax); mpole(:,ix) += rrlm * fint(p)
//   YT (:1316-1336): for every basin m (docelatom only): w = yt_weights(m); points with |w| < 1e-15 are skipped;
//        mpole(:,m) += rrlm(p/n - xattr(:,m)) * fint(p) * w(p)
//   then mpole = mpole * omega / ntot (:1360).
// The reference runs the Bader loop under an `omp critical` and the YT loop once per basin over the full grid.
//
// Per point this is several hundred fp64 operations (the Masters & Richards-Dinger recursion of genylm,
// src/tools_math@proc.f90:314-377, with its divisions and square roots) against 12 bytes of HBM traffic (label +
// field value): the kernel is FP64-pipe bound, not HBM bound.  Layout: persistent warps walk contiguous chunks of the (slab of the) grid, one
// point per lane; a lane keeps the (lmax+1)^2 partial moments of the warp's current basin in registers (lmax <= 5,
// the reference's default) and the warp reduces them with shuffles only when the basin changes, as k_basin_reduce
// does.  The recursion follows the Fortran evaluation order (the file is built with -fmad=false); the angles of
// tosphere are replaced by their sines and cosines taken directly from the vector (see add_point), so parity is a
// tolerance (1e-10 of the sum of |terms|; measured ~1e-15), not bits.
#include "common.cuh"
#include "group.h"

#include <algorithm>

namespace {

constexpr int MP_LCAP_FAST = 5;    // register-resident accumulators (36 per lane)
constexpr int MP_LCAP_MAX = 10;    // larger lmax: same code, accumulators spill to local memory
constexpr int MP_CHUNK_ITERS = 8;  // 256 consecutive points per warp chunk
constexpr int MP_THREADS = 128;     // ~160 registers per thread: 3 blocks (12 warps) per SM
constexpr int MP_BLOCKS_PER_SM = 3;

struct MpArgs {
  int n1, n2, n3;
  unsigned z0;            // first owned plane
  unsigned nnl;           // owned points
  int lmax, nlm;
  int isortho, isortho_del, nws;
  double x2c[9], x2xr[9], xr2c[9];  // column-major
  const double* ws;       // (3, nws) Cartesian Wigner-Seitz neighbours (ws_ineighc)
  const double* xattr;    // (3, nattr) crystallographic
  const int* label;       // Bader: per-maximum labels of the owned planes
  const int* map;         // maximum -> basin (1-based, 0 = discarded)
  int mask;
  const double* w;        // YT: weights of basin idb on the whole grid
  int idb;
  const double* fint;     // integrand, owned planes
  double* sums;           // (nlm, nattr)
};

__device__ __forceinline__ void matvec3(const double* m, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; i++) y[i] = m[i] * x[0] + m[i + 3] * x[1] + m[i + 6] * x[2];
}

// crystal%shortest, src/crystalmod@proc.f90:1056-1085 (x cryst. in, Cartesian out)
__device__ __forceinline__ void shortest(const MpArgs& a, double x[3]) {
  double t[3];
  if (a.isortho) {
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = x[i] - round(x[i]);  // nint: half away from zero
    matvec3(a.x2c, t, x);
    return;
  }
  matvec3(a.x2xr, x, t);
#pragma unroll
  for (int i = 0; i < 3; i++) t[i] = t[i] - round(t[i]);
  matvec3(a.xr2c, t, x);
  if (!a.isortho_del) {
    double dist = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    for (int i = 0; i < a.nws; i++) {
      const double y0 = x0 + __ldg(a.ws + 3 * i), y1 = x1 + __ldg(a.ws + 3 * i + 1), y2 = x2 + __ldg(a.ws + 3 * i + 2);
      const double d = sqrt(y0 * y0 + y1 * y1 + y2 * y2);
      if (d < dist) {
        x[0] = y0; x[1] = y1; x[2] = y2;
        dist = d;
      }
    }
  }
}

// r**l as gfortran evaluates it (libgcc __powidf2)
__device__ __forceinline__ double powi(double x, int m) {
  unsigned n = (unsigned)m;
  double y = (n & 1u) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n & 1u) y *= x;
  }
  return y;
}

// tosphere + genylm + genrlm_real (src/tools_math@proc.f90:381-406, :314-377, :273-306) fused with the
// accumulation acc(:) += (rrlm * f) * w.  Loops are fully unrolled so that acc, x, zc, zs stay in registers.
template <int LCAP>
__device__ __forceinline__ void add_point(double (&acc)[(LCAP + 1) * (LCAP + 1)], int lmax, const double v[3], double f, double w) {
  constexpr double pi = 3.14159265358979323846, fourpi = 12.566370614359172954, eps = 1e-14;
  const double sh = 1.0 / sqrt(2.0);
  const double r = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  // l = 0: rlm(1) = ylm(1) * sqrt(4 pi / 1) * r**0
  acc[0] += ((0.28209479177387814347 * sqrt(4.0 * pi / 1.0) * 1.0) * f) * w;
  if (lmax == 0) return;
  // The reference goes through the angles: theta = acos(z/r), phi = atan2(y,x), then sin/cos(theta) and
  // cos/sin(m phi).  The same quantities without transcendentals: cos(theta) = z/r, sin(theta) = rho/r with
  // rho = sqrt(x^2+y^2), cos(phi) = x/rho, sin(phi) = y/rho and the angle-addition recurrence for m phi (error
  // <~ m ulp).  This is the better-conditioned form (acos loses half the digits near the poles) and takes the kernel
  // off the libm slow paths; the branches of tosphere are kept (r <= eps, |z/r| >= 1, |x|,|y| <= eps).
  double sn = 0.0, cs = 1.0, c1 = 1.0, s1 = 0.0;
  if (r > eps) {
    const double t1 = v[2] / r;
    const double rho2 = v[0] * v[0] + v[1] * v[1];
    const double rho = sqrt(rho2);
    if (t1 >= 1.0) { cs = 1.0; sn = 0.0; }
    else if (t1 <= -1.0) { cs = -1.0; sn = 1.2246467991473532e-16; }  // cos(pi), sin(pi) in fp64
    else { cs = t1; sn = rho / r; }
    if (fabs(v[0]) > eps || fabs(v[1]) > eps) { c1 = v[0] / rho; s1 = v[1] / rho; }
  }
  double zc[LCAP + 1], zs[LCAP + 1];
  zc[1] = c1; zs[1] = s1;
#pragma unroll
  for (int m = 2; m <= LCAP; m++)
    if (m <= lmax) {
      zc[m] = zc[m - 1] * c1 - zs[m - 1] * s1;
      zs[m] = zs[m - 1] * c1 + zc[m - 1] * s1;
    }
  double x[LCAP + 1];
#pragma unroll
  for (int l = 1; l <= LCAP; l++) {
    if (l > lmax) break;
    x[l] = (l & 1) ? -1.0 : 1.0;
    double dx = 0.0;
#pragma unroll
    for (int m = l; m >= 1; m--) {
      const double t1 = sqrt((double)((l + m) * (l - m + 1)));
      x[m - 1] = -(sn * dx + (double)(2 * m) * cs * x[m]) / t1;
      dx = sn * x[m] * t1;
    }
    double t1 = sn, sum = 0.0;
#pragma unroll
    for (int m = 1; m <= l; m++) {
      x[m] = t1 * x[m];
      sum = sum + x[m] * x[m];
      t1 = t1 * sn;
    }
    sum = 2.0 * sum + x[0] * x[0];
    t1 = sqrt((double)(2 * l + 1) / (fourpi * sum));
    const double s = sqrt(4.0 * pi / (double)(2 * l + 1)), rl = powi(r, l);
    acc[l * l + l] += ((((t1 * x[0]) * s) * rl) * f) * w;
#pragma unroll
    for (int m = 1; m <= l; m++) {
      const double a = t1 * x[m];
      const double re = ((a * zc[m]) * s) * rl, im = ((a * zs[m]) * s) * rl;  // rlm(ip); rlm(im) = iphas * conj
      const double ph_re = (m & 1) ? -re : re, ph_im = (m & 1) ? -im : im;     // iphas * rlm(ip)
      acc[l * l + l - m] += ((sh * (ph_re + ph_re)) * f) * w;                   // cosine harmonic C_lm
      acc[l * l + l + m] += ((sh * (ph_im + ph_im)) * f) * w;                   // sine harmonic S_lm
    }
  }
}

template <int LCAP, bool YT>
__global__ void __launch_bounds__(MP_THREADS) k_multipoles(const __grid_constant__ MpArgs a) {
  constexpr int NLM = (LCAP + 1) * (LCAP + 1);
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned chunk = 32u * MP_CHUNK_ITERS;
  const unsigned plane = (unsigned)a.n1 * (unsigned)a.n2;
  double acc[NLM];
#pragma unroll
  for (int e = 0; e < NLM; e++) acc[e] = 0.0;
  int cur = -1;        // warp-uniform basin (0-based) of the current run
  bool dirty = false;  // this lane added something since the last flush
  auto flush = [&]() {
    if (cur >= 0 && __any_sync(0xffffffffu, dirty)) {
#pragma unroll
      for (int e = 0; e < NLM; e++) {
        if (e < a.nlm) {
          double s = acc[e];
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
          if (lane == 0) atomicAdd(a.sums + (size_t)cur * a.nlm + e, s);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NLM; e++) acc[e] = 0.0;
    dirty = false;
  };
  for (unsigned long long base = (unsigned long long)warp * chunk; base < a.nnl; base += (unsigned long long)nwarps * chunk) {
    for (int k = 0; k < MP_CHUNK_ITERS; k++) {
      const unsigned long long i = base + 32u * k + lane;
      int lab = -1;
      double f = 0.0, w = 1.0;
      if (i < a.nnl) {
        if (YT) {
          w = __ldg(a.w + i);
          lab = fabs(w) < 1e-15 ? -1 : a.idb - 1;
        } else {
          const int l = __ldg(a.label + i) & a.mask;  // isosurface regions: -1 below the contour value
          lab = l >= 0 ? __ldg(a.map + l) - 1 : -1;
        }
        f = __ldg(a.fint + i);
      }
      bool rem = lab >= 0;
      for (;;) {
        if (rem && lab == cur) {
          const unsigned g = (unsigned)i + a.z0 * plane;  // nn < 2^31
          const unsigned iz = g / plane, q = g - iz * plane, iy = q / (unsigned)a.n1, ix = q - iy * (unsigned)a.n1;
          double dv[3] = {(double)ix / (double)a.n1 - __ldg(a.xattr + 3 * lab), (double)iy / (double)a.n2 - __ldg(a.xattr + 3 * lab + 1),
                          (double)iz / (double)a.n3 - __ldg(a.xattr + 3 * lab + 2)};
          shortest(a, dv);
          add_point<LCAP>(acc, a.lmax, dv, f, w);
          dirty = true;
          rem = false;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, rem);
        if (!bal) break;
        flush();
        cur = __shfl_sync(0xffffffffu, lab, __ffs(bal) - 1);
      }
    }
  }
  flush();
}

template <bool YT>
cudaError_t launch(c2g_context* ctx, const MpArgs& a) {
  const int blocks = ctx->nsm * MP_BLOCKS_PER_SM;
  if (a.lmax <= MP_LCAP_FAST) k_multipoles<MP_LCAP_FAST, YT><<<blocks, MP_THREADS, 0, ctx->stream>>>(a);
  else k_multipoles<MP_LCAP_MAX, YT><<<blocks, MP_THREADS, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

extern "C" int c2g_integrate_multipoles(c2g_context* ctx, c2g_basins* res, int fieldhandle, int lmax, const double* xattr,
                                        const unsigned char* domask, int isortho, int isortho_del, const double x2c[9],
                                        const double x2xr[9], const double xr2c[9], int nws, const double* ws_ineighc,
                                        double omega, double* mpole) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group && res && !res->parts.empty())
    return grp_integrate_multipoles(ctx, res, fieldhandle, lmax, xattr, domask, isortho, isortho_del, x2c, x2xr, xr2c, nws, ws_ineighc, omega, mpole);
  if (ctx->group && res) { ctx = res->ctx; cudaSetDevice(ctx->device); }
  if (!res || !xattr || !mpole || !x2c) return ctx->fail(C2G_ERR_ARG, "c2g_integrate_multipoles: bad argument");
  if (lmax < 0 || lmax > MP_LCAP_MAX)
    return ctx->fail(C2G_ERR_ARG, "c2g_integrate_multipoles: lmax = %d out of range (0..%d)", lmax, MP_LCAP_MAX);
  if (!isortho && (!x2xr || !xr2c || nws < 0 || (nws > 0 && !ws_ineighc)))
    return ctx->fail(C2G_ERR_ARG, "c2g_integrate_multipoles: a non-orthogonal cell needs x2xr, xr2c and the WS neighbours");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_integrate_multipoles: call c2g_basins_set_map first");
  const int h = fieldhandle;
  if (h < 0 || h >= (int)ctx->grids.size() || !ctx->grids[h].used) return ctx->fail(C2G_ERR_ARG, "c2g_integrate_multipoles: invalid field handle %d", h);
  c2g_grid_ready(ctx, h);
  const c2g_grid& g = ctx->grids[h];
  if (g.n[0] != res->n[0] || g.n[1] != res->n[1] || g.n[2] != res->n[2])
    return ctx->fail(C2G_ERR_ARG, "c2g_integrate_multipoles: the field has a different grid size");
  const int nattr = res->nattr, nlm = (lmax + 1) * (lmax + 1);
  if (nattr == 0) return C2G_OK;

  MpArgs a;
  memset(&a, 0, sizeof(a));
  a.n1 = res->n[0]; a.n2 = res->n[1]; a.n3 = res->n[2];
  a.lmax = lmax; a.nlm = nlm;
  a.isortho = isortho ? 1 : 0; a.isortho_del = isortho_del ? 1 : 0; a.nws = isortho ? 0 : nws;
  for (int i = 0; i < 9; i++) {
    a.x2c[i] = x2c[i];
    a.x2xr[i] = (!isortho) ? x2xr[i] : 0.0;
    a.xr2c[i] = (!isortho) ? xr2c[i] : 0.0;
  }
  DevBuf b_ws(ctx), b_xattr(ctx), b_sums(ctx), b_w(ctx);
  if (a.nws > 0) {
    C2G_CUDA(ctx, b_ws.alloc(ctx, sizeof(double) * 3 * a.nws));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_ws.p, ws_ineighc, sizeof(double) * 3 * a.nws, cudaMemcpyHostToDevice, ctx->stream));
  }
  C2G_CUDA(ctx, b_xattr.alloc(ctx, sizeof(double) * 3 * nattr));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_xattr.p, xattr, sizeof(double) * 3 * nattr, cudaMemcpyHostToDevice, ctx->stream));
  C2G_CUDA(ctx, b_sums.alloc(ctx, sizeof(double) * (size_t)nlm * nattr));
  C2G_CUDA(ctx, cudaMemsetAsync(b_sums.p, 0, sizeof(double) * (size_t)nlm * nattr, ctx->stream));
  a.ws = b_ws.as<double>();
  a.xattr = b_xattr.as<double>();
  a.sums = b_sums.as<double>();
  a.map = res->d_map;
  const size_t plane = (size_t)res->n[0] * res->n[1];

  if (res->kind != 1) {
    // Bader labels (or isosurface regions): this rank's z-slab, partial moments all-reduced below
    a.z0 = (unsigned)res->zlo;
    a.nnl = (unsigned)(plane * (size_t)(res->zhi - res->zlo));
    a.label = res->d_label;
    a.mask = res->kind == 0 ? 0x7fffffff : -1;
    a.fint = g.d + plane * res->zlo;
    if (a.nnl > 0) {
      ctx->prof_begin("multipoles");
      cudaError_t e = launch<false>(ctx, a);
      ctx->prof_end();
      if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "k_multipoles launch: %s", cudaGetErrorString(e));
    }
    if (ctx->nranks > 1 && res->kind == 0) {  // z-slabs of Bader labels; ISOSURFACE regions are replicated on every rank
      ctx->prof_begin("multipoles_allreduce_nccl");
      ncclResult_t r = ncclAllReduce(a.sums, a.sums, (size_t)nlm * nattr, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream);
      ctx->prof_end();
      if (r != ncclSuccess) return ctx->fail(C2G_ERR_NCCL, "c2g_integrate_multipoles: ncclAllReduce failed");
    }
  } else {
    // YT: one weight field per basin (replicas only on multi-GPU contexts, like the rest of YT)
    a.z0 = 0;
    a.nnl = (unsigned)res->nn;
    a.fint = g.d;
    C2G_CUDA(ctx, b_w.alloc(ctx, sizeof(double) * res->nn));
    a.w = b_w.as<double>();
    for (int m = 1; m <= nattr; m++) {
      if (domask && !domask[m - 1]) continue;
      int rc = c2g_yt_weights_device(res, m, b_w.as<double>());
      if (rc != C2G_OK) return rc;
      a.idb = m;
      ctx->prof_begin("multipoles_yt");
      cudaError_t e = launch<true>(ctx, a);
      ctx->prof_end();
      if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "k_multipoles launch: %s", cudaGetErrorString(e));
    }
  }
  std::vector<double> hs((size_t)nlm * nattr);
  C2G_CUDA(ctx, cudaMemcpyAsync(hs.data(), a.sums, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_collect();
  const double ntot = (double)res->nn;
  for (size_t e = 0; e < hs.size(); e++) mpole[e] = hs[e] * omega / ntot;  // :1360
  return C2G_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Attractor images for DELOC: bader_remap (src/bader@proc.f90:237-296) and yt_remap (src/yt@proc.f90:533-594).
//
// For every point: x = p/n - xattr(:,basin), xs = shortest(x), lattice vector pl = nint(x - c2x(xs)).  Points with
// pl /= 0 belong to the image (basin, pl) of their attractor; the reference numbers the images nattr+1, nattr+2, ...
// in the order in which its scan (index 1 fastest) first meets them and searches a growing list for every such
// point.  Here: one pass takes atomicMin(first[basin][pl], linear index) over a dense (nattr x 7^3) table, the host
// sorts the few occupied entries by that first index (= the reference's numbering; YT: by basin, then first index,
// like its basin-outer loop), a second pass writes idg1 through the resulting table.  Same `shortest` as above.
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int RM_R = 3, RM_W = 2 * RM_R + 1, RM_CODES = RM_W * RM_W * RM_W;

struct RemapArgs {
  MpArgs a;
  double c2x[9];
  int* first;        // (RM_CODES, nattr): smallest linear index of a point of image (basin, code)
  const int* idmap;  // apply pass: (RM_CODES, nattr) -> new id
  int* out;          // apply pass: idg1 of the owned planes
  int* err;
  int apply;
};

template <bool YT>
__global__ void __launch_bounds__(256) k_remap(const __grid_constant__ RemapArgs r) {
  const MpArgs& a = r.a;
  const unsigned plane = (unsigned)a.n1 * (unsigned)a.n2;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nnl; i += stride) {
    int lab;
    if (YT) {
      lab = fabs(__ldg(a.w + i)) < 1e-15 ? -1 : a.idb - 1;
    } else {
      const int l = __ldg(a.label + i) & a.mask;
      lab = l >= 0 ? __ldg(a.map + l) - 1 : -1;
    }
    if (lab < 0) {
      if (r.apply) r.out[i] = 0;
      continue;
    }
    const unsigned g = (unsigned)i + a.z0 * plane;
    const unsigned iz = g / plane, q = g - iz * plane, iy = q / (unsigned)a.n1, ix = q - iy * (unsigned)a.n1;
    const double x[3] = {(double)ix / (double)a.n1 - __ldg(a.xattr + 3 * lab), (double)iy / (double)a.n2 - __ldg(a.xattr + 3 * lab + 1),
                         (double)iz / (double)a.n3 - __ldg(a.xattr + 3 * lab + 2)};
    double xs[3] = {x[0], x[1], x[2]}, xc[3];
    shortest(a, xs);
    matvec3(r.c2x, xs, xc);
    const int p0 = (int)round(x[0] - xc[0]), p1 = (int)round(x[1] - xc[1]), p2 = (int)round(x[2] - xc[2]);
    if ((p0 | p1 | p2) == 0) {
      if (r.apply) r.out[i] = lab + 1;
      continue;
    }
    if (abs(p0) > RM_R || abs(p1) > RM_R || abs(p2) > RM_R) {
      atomicExch(r.err, 1);
      continue;
    }
    const int code = ((p0 + RM_R) * RM_W + (p1 + RM_R)) * RM_W + (p2 + RM_R);
    if (r.apply) r.out[i] = __ldg(r.idmap + (size_t)lab * RM_CODES + code);
    else atomicMin(r.first + (size_t)lab * RM_CODES + code, (int)g);
  }
}

}  // namespace

extern "C" int c2g_basins_remap(c2g_context* ctx, c2g_basins* res, const double* xattr, const double c2x[9], int isortho,
                                int isortho_del, const double x2c[9], const double x2xr[9], const double xr2c[9], int nws,
                                const double* ws_ineighc, int maxattn, int* nattn_out, int* iatt, int* ilvec, int* idg1) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group && res && !res->parts.empty())
    return grp_basins_remap(ctx, res, xattr, c2x, isortho, isortho_del, x2c, x2xr, xr2c, nws, ws_ineighc, maxattn, nattn_out, iatt, ilvec, idg1);
  if (ctx->group && res) { ctx = res->ctx; cudaSetDevice(ctx->device); }
  if (!res || !xattr || !c2x || !x2c || !nattn_out || !iatt || !ilvec) return ctx->fail(C2G_ERR_ARG, "c2g_basins_remap: bad argument");
  if (!isortho && (!x2xr || !xr2c || nws < 0 || (nws > 0 && !ws_ineighc)))
    return ctx->fail(C2G_ERR_ARG, "c2g_basins_remap: a non-orthogonal cell needs x2xr, xr2c and the WS neighbours");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_basins_remap: call c2g_basins_set_map first");
  if (res->kind == 1 && idg1) return ctx->fail(C2G_ERR_ARG, "c2g_basins_remap: yt_remap has no idg1 output");
  const int nattr = res->nattr;
  if (maxattn < nattr) { *nattn_out = nattr; return ctx->fail(C2G_ERR_OVERFLOW, "c2g_basins_remap: maxattn < nattr"); }
  for (int i = 0; i < nattr; i++) {
    iatt[i] = i + 1;
    ilvec[3 * i] = ilvec[3 * i + 1] = ilvec[3 * i + 2] = 0;
  }
  *nattn_out = nattr;
  if (nattr == 0) return C2G_OK;
  cudaStream_t st = ctx->stream;

  RemapArgs r;
  memset(&r, 0, sizeof(r));
  MpArgs& a = r.a;
  a.n1 = res->n[0]; a.n2 = res->n[1]; a.n3 = res->n[2];
  a.isortho = isortho ? 1 : 0; a.isortho_del = isortho_del ? 1 : 0; a.nws = isortho ? 0 : nws;
  for (int i = 0; i < 9; i++) {
    a.x2c[i] = x2c[i];
    a.x2xr[i] = (!isortho) ? x2xr[i] : 0.0;
    a.xr2c[i] = (!isortho) ? xr2c[i] : 0.0;
    r.c2x[i] = c2x[i];
  }
  DevBuf b_ws(ctx), b_xattr(ctx), b_first(ctx), b_idmap(ctx), b_err(ctx), b_w(ctx), b_out(ctx);
  if (a.nws > 0) {
    C2G_CUDA(ctx, b_ws.alloc(ctx, sizeof(double) * 3 * a.nws));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_ws.p, ws_ineighc, sizeof(double) * 3 * a.nws, cudaMemcpyHostToDevice, st));
  }
  C2G_CUDA(ctx, b_xattr.alloc(ctx, sizeof(double) * 3 * nattr));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_xattr.p, xattr, sizeof(double) * 3 * nattr, cudaMemcpyHostToDevice, st));
  const size_t ntab = (size_t)nattr * RM_CODES;
  C2G_CUDA(ctx, b_first.alloc(ctx, sizeof(int) * ntab));
  C2G_CUDA(ctx, cudaMemsetAsync(b_first.p, 0x7f, sizeof(int) * ntab, st));
  C2G_CUDA(ctx, b_err.alloc(ctx, sizeof(int)));
  C2G_CUDA(ctx, cudaMemsetAsync(b_err.p, 0, sizeof(int), st));
  a.ws = b_ws.as<double>();
  a.xattr = b_xattr.as<double>();
  a.map = res->d_map;
  r.first = b_first.as<int>();
  r.err = b_err.as<int>();
  const size_t plane = (size_t)res->n[0] * res->n[1];
  const int blocks = ctx->nsm * 8;

  if (res->kind != 1) {
    a.z0 = (unsigned)res->zlo;
    a.nnl = (unsigned)(plane * (size_t)(res->zhi - res->zlo));
    a.label = res->d_label;
    a.mask = res->kind == 0 ? 0x7fffffff : -1;
    if (a.nnl > 0) {
      ctx->prof_begin("remap_scan");
      k_remap<false><<<blocks, 256, 0, st>>>(r);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    if (ctx->nranks > 1 && res->kind == 0) {  // the first point of an image may lie in another rank's slab
      ncclResult_t rr = ncclAllReduce(r.first, r.first, ntab, ncclInt, ncclMin, (ncclComm_t)ctx->nccl, st);
      if (rr != ncclSuccess) return ctx->fail(C2G_ERR_NCCL, "c2g_basins_remap: ncclAllReduce failed");
    }
  } else {
    a.z0 = 0;
    a.nnl = (unsigned)res->nn;
    C2G_CUDA(ctx, b_w.alloc(ctx, sizeof(double) * res->nn));
    a.w = b_w.as<double>();
    for (int m = 1; m <= nattr; m++) {
      int rc = c2g_yt_weights_device(res, m, b_w.as<double>());
      if (rc != C2G_OK) return rc;
      a.idb = m;
      ctx->prof_begin("remap_scan_yt");
      k_remap<true><<<blocks, 256, 0, st>>>(r);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
  }
  std::vector<int> first(ntab);
  int herr = 0;
  C2G_CUDA(ctx, cudaMemcpyAsync(first.data(), r.first, sizeof(int) * ntab, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(&herr, r.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  if (herr) return ctx->fail(C2G_ERR_OVERFLOW, "c2g_basins_remap: a lattice vector outside -%d..%d (cell not reduced?)", RM_R, RM_R);
  // images in the reference's order of first appearance
  struct Img { int first, basin, code; };
  std::vector<Img> imgs;
  for (int b = 0; b < nattr; b++)
    for (int c = 0; c < RM_CODES; c++)
      if (first[(size_t)b * RM_CODES + c] != 0x7f7f7f7f) imgs.push_back({first[(size_t)b * RM_CODES + c], b, c});
  if (res->kind == 1)
    std::sort(imgs.begin(), imgs.end(), [](const Img& p, const Img& q) { return p.basin != q.basin ? p.basin < q.basin : p.first < q.first; });
  else
    std::sort(imgs.begin(), imgs.end(), [](const Img& p, const Img& q) { return p.first < q.first; });
  const int nattn = nattr + (int)imgs.size();
  *nattn_out = nattn;
  if (nattn > maxattn) return ctx->fail(C2G_ERR_OVERFLOW, "c2g_basins_remap: %d attractor images, maxattn = %d", nattn, maxattn);
  std::vector<int> idmap(ntab, 0);
  for (size_t k = 0; k < imgs.size(); k++) {
    const int id = nattr + (int)k;  // 0-based slot
    iatt[id] = imgs[k].basin + 1;
    const int c = imgs[k].code;
    ilvec[3 * id] = c / (RM_W * RM_W) - RM_R;
    ilvec[3 * id + 1] = (c / RM_W) % RM_W - RM_R;
    ilvec[3 * id + 2] = c % RM_W - RM_R;
    idmap[(size_t)imgs[k].basin * RM_CODES + c] = id + 1;
  }
  if (idg1 && a.nnl > 0) {
    C2G_CUDA(ctx, b_idmap.alloc(ctx, sizeof(int) * ntab));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_idmap.p, idmap.data(), sizeof(int) * ntab, cudaMemcpyHostToDevice, st));
    C2G_CUDA(ctx, b_out.alloc(ctx, sizeof(int) * (size_t)a.nnl));
    r.idmap = b_idmap.as<int>();
    r.out = b_out.as<int>();
    r.apply = 1;
    ctx->prof_begin("remap_apply");
    k_remap<false><<<blocks, 256, 0, st>>>(r);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(idg1, r.out, sizeof(int) * (size_t)a.nnl, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  ctx->prof_collect();
  return C2G_OK;
}
