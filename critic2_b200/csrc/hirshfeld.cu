// hirshfeld.cu -- promolecular density on a grid and Hirshfeld integration (HIRSHFELD keyword on grids).
//
// Replaces promolecular_array3 (critic2 src/crystalmod@complex.f90:436-470; per point promolecular_atom,
// src/crystalmod@env.f90:622-748, with grid1%interp, src/grid1mod@proc.f90:86-137) and the grid loop of
// intgrid_hirshfeld_fields (src/integration@proc.f90:1552-1596):
//     rho_pro(p) = sum over atom images within the species cutoff of max(interp_spc(max(r, r(1), 1e-14)), 0)
//     psum(A,l)  = sum_p  interp_spc(A)(|p - A'|) / max(rho_pro(p), vsmall) * f_l(p)    over the images A' of atom A
// The reference finds the atoms near every point with its block environment (list_near_atoms).  Here the host
// enumerates every periodic image that can reach the cell once; a block owns a tile of 256 consecutive points,
// warp 0 culls the image list against the tile's bounding sphere into shared memory (in list order, so the order of
// the additions is fixed), and every lane then runs over the few survivors.  Basin sums go through a per-block
// shared-memory table (warp shuffles first) and one set of global fp64 atomics per block at the end.
// Values agree with the reference to rounding (different order of the additions), not bit for bit: tolerance 1e-10.
#include "common.cuh"
#include "group.h"

#include <algorithm>
#include <cmath>

namespace {

constexpr int HB_THREADS = 256;
constexpr int HB_LIST = 2048;       // culled images per round
constexpr int HB_TABLE_MAX = 4096;  // doubles of dynamic shared memory for the per-block sums (32 KB, + 8 KB static)
constexpr double VSMALL_H = 1e-80;  // param.F90:27

struct HParams {
  int n1, n2, n3;
  unsigned nn;
  double x2c[9];
  int nimg;
  const int* img_atom;    // atom of every image
  const double* img_x;    // (3, nimg) Cartesian
  const int* atom_spc;    // 0-based species of every atom
  const unsigned char* amask;  // per atom: 0 = skipped (fragment / docelatom); null = all
  const int* ngrid;
  const int* off;
  const double *a, *b, *rmax, *rcut, *rtab, *ftab;
};

// grid1%interp, value only (grid1mod@proc.f90:86-137)
__device__ __forceinline__ double interp(const HParams& P, int is, double r0) {
  const int ng = __ldg(P.ngrid + is);
  if (ng <= 0) return 0.0;
  if (r0 >= __ldg(P.rmax + is)) return 0.0;
  const double* rg = P.rtab + __ldg(P.off + is);
  const double* fg = P.ftab + __ldg(P.off + is);
  int ir;
  double r;
  const double r1 = __ldg(rg);
  if (r0 <= r1) { ir = 1; r = r1; }
  else { ir = 1 + (int)floor(log(r0 / __ldg(P.a + is)) / __ldg(P.b + is)); r = r0; }
  const int i0 = min(max(ir, 2), ng - 2) - 2;
  double rr[4], dr1[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    rr[i] = __ldg(rg + i0 + i);
    dr1[i] = r - rr[i];
  }
  double f = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double prod = 1.0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (i == j) continue;
      // x1dr12(i,j) = 1/(rr(i)-rr(j)) for j < i, -(1/(rr(j)-rr(i))) for j > i (:113-116)
      const double x1 = j < i ? 1.0 / (rr[i] - rr[j]) : -(1.0 / (rr[j] - rr[i]));
      prod = prod * dr1[j] * x1;
    }
    f = f + __ldg(fg + i0 + i) * prod;
  }
  return f;
}

struct Tile {
  double x, y, z;  // Cartesian position of this thread's point
  bool valid;
  unsigned i;
};

__device__ __forceinline__ Tile tile_point(const HParams& P, unsigned tile) {
  Tile t;
  t.i = tile * HB_THREADS + threadIdx.x;
  t.valid = t.i < P.nn;
  const unsigned i = t.valid ? t.i : P.nn - 1;
  const unsigned plane = (unsigned)P.n1 * (unsigned)P.n2;
  const unsigned iz = i / plane, q = i - iz * plane, iy = q / (unsigned)P.n1, ix = q - iy * (unsigned)P.n1;
  const double f0 = (double)ix / (double)P.n1, f1 = (double)iy / (double)P.n2, f2 = (double)iz / (double)P.n3;
  t.x = P.x2c[0] * f0 + P.x2c[3] * f1 + P.x2c[6] * f2;
  t.y = P.x2c[1] * f0 + P.x2c[4] * f1 + P.x2c[7] * f2;
  t.z = P.x2c[2] * f0 + P.x2c[5] * f1 + P.x2c[8] * f2;
  return t;
}

// shared state of the culling
struct Cull {
  double cx, cy, cz;
  unsigned rad_bits;  // float bits of the largest distance of a tile point from (cx,cy,cz)
  int cnt, next;
};

// bounding sphere of the tile: centre = the point of thread 0, radius = max distance (as a float, rounded up)
__device__ __forceinline__ double tile_sphere(Cull& c, const Tile& t) {
  if (threadIdx.x == 0) { c.cx = t.x; c.cy = t.y; c.cz = t.z; c.rad_bits = 0u; c.next = 0; }
  __syncthreads();
  const double dx = t.x - c.cx, dy = t.y - c.cy, dz = t.z - c.cz;
  const float d = __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz));
  atomicMax(&c.rad_bits, __float_as_uint(d));  // non-negative floats order like their bit patterns
  __syncthreads();
  return (double)__uint_as_float(c.rad_bits) * 1.000001 + 1e-9;
}

// warp 0 appends, in list order, the images from c.next on that can reach the tile; stops when the list is full
__device__ __forceinline__ void cull_round(const HParams& P, Cull& c, double radius, int* s_list) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int cnt = 0, base = c.next;
    __syncwarp();  // every lane has read c.next before lane 0 rewrites it below
    for (; base < P.nimg && cnt + 32 <= HB_LIST; base += 32) {
      const int m = base + lane;
      bool keep = false;
      if (m < P.nimg) {
        const int at = __ldg(P.img_atom + m);
        if (!P.amask || P.amask[at]) {
          const int is = __ldg(P.atom_spc + at);
          if (__ldg(P.ngrid + is) > 0) {
            const double dx = __ldg(P.img_x + 3 * m) - c.cx, dy = __ldg(P.img_x + 3 * m + 1) - c.cy, dz = __ldg(P.img_x + 3 * m + 2) - c.cz;
            const double lim = __ldg(P.rcut + is) + radius;
            keep = dx * dx + dy * dy + dz * dz <= lim * lim;
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) s_list[cnt + __popc(bal & ((1u << lane) - 1u))] = m;
      cnt += __popc(bal);
    }
    if (lane == 0) { c.cnt = cnt; c.next = base; }
  }
  __syncthreads();
}

// promolecular_array3: one value per grid point
__global__ void __launch_bounds__(HB_THREADS) k_promolecular(const __grid_constant__ HParams P, unsigned ntiles, double* __restrict__ out) {
  __shared__ Cull c;
  __shared__ int s_list[HB_LIST];
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = tile_point(P, tile);
    const double radius = tile_sphere(c, t);
    double f = 0.0;
    for (;;) {
      cull_round(P, c, radius, s_list);
      const int cnt = c.cnt;
      for (int q = 0; q < cnt; q++) {
        const int m = s_list[q];
        const int is = __ldg(P.atom_spc + __ldg(P.img_atom + m));
        const double dx = t.x - __ldg(P.img_x + 3 * m), dy = t.y - __ldg(P.img_x + 3 * m + 1), dz = t.z - __ldg(P.img_x + 3 * m + 2);
        double r = sqrt(dx * dx + dy * dy + dz * dz);
        if (r > __ldg(P.rcut + is)) continue;                          // list_near_atoms(up2dsp), env.f90:692
        r = fmax(fmax(r, __ldg(P.rtab + __ldg(P.off + is))), 1e-14);   // :724
        f = f + fmax(interp(P, is, r), 0.0);                           // :725-727
      }
      const bool more = c.next < P.nimg;
      __syncthreads();  // everybody is done with s_list and c before the next round / tile
      if (!more) break;
    }
    if (t.valid) out[t.i] = f;
  }
}

// intgrid_hirshfeld_fields: sums[atom*(NP+1) + l], l = NP is the volume
template <int NP>
__global__ void __launch_bounds__(HB_THREADS) k_hirshfeld(const __grid_constant__ HParams P, unsigned ntiles, int nat,
                                                          const double* __restrict__ promol, const double* __restrict__ f0,
                                                          const double* __restrict__ f1, const double* __restrict__ f2,
                                                          const double* __restrict__ f3, double* __restrict__ sums) {
  __shared__ Cull c;
  __shared__ int s_list[HB_LIST];
  extern __shared__ double s_tab[];  // nat * (NP+1)
  const int lane = threadIdx.x & 31;
  const double* fp[4] = {f0, f1, f2, f3};
  for (int e = threadIdx.x; e < nat * (NP + 1); e += blockDim.x) s_tab[e] = 0.0;
  __syncthreads();
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = tile_point(P, tile);
    const double radius = tile_sphere(c, t);
    double fac = 0.0, fv[NP > 0 ? NP : 1];
    if (t.valid) {
      fac = 1.0 / fmax(__ldg(promol + t.i), VSMALL_H);                 // :1561
#pragma unroll
      for (int l = 0; l < NP; l++) fv[l] = __ldg(fp[l] + t.i);
    } else {
#pragma unroll
      for (int l = 0; l < NP; l++) fv[l] = 0.0;
    }
    for (;;) {
      cull_round(P, c, radius, s_list);
      const int cnt = c.cnt;
      for (int q = 0; q < cnt; q++) {
        const int m = s_list[q];
        const int at = __ldg(P.img_atom + m);
        const int is = __ldg(P.atom_spc + at);
        const double dx = t.x - __ldg(P.img_x + 3 * m), dy = t.y - __ldg(P.img_x + 3 * m + 1), dz = t.z - __ldg(P.img_x + 3 * m + 2);
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        double tosum = 0.0;
        if (t.valid && r <= __ldg(P.rcut + is)) tosum = fac * interp(P, is, r);   // :1566-1567
        if (!__any_sync(0xffffffffu, tosum != 0.0)) continue;
        double v[NP + 1];
#pragma unroll
        for (int l = 0; l < NP; l++) v[l] = tosum * fv[l];
        v[NP] = tosum;
#pragma unroll
        for (int l = 0; l <= NP; l++) {
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) v[l] += __shfl_xor_sync(0xffffffffu, v[l], d);
          if (lane == 0) atomicAdd(s_tab + at * (NP + 1) + l, v[l]);
        }
      }
      const bool more = c.next < P.nimg;
      __syncthreads();
      if (!more) break;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nat * (NP + 1); e += blockDim.x)
    if (s_tab[e] != 0.0) atomicAdd(sums + e, s_tab[e]);
}

// host side: images, species tables and atoms on the device
struct HSetup {
  DevBuf img_atom, img_x, atom_spc, amask, ngrid, off, a, b, rmax, rcut, rtab, ftab;
  HParams P;
};

int h_setup(c2g_context* ctx, const char* who, const int n[3], const double x2c[9], int nat, const double* xat, const int* ispc,
            int nspc, const int* spc_ngrid, const int* spc_off, const double* spc_a, const double* spc_b, const double* spc_rmax,
            const double* spc_rcut, const double* rtab, const double* ftab, const unsigned char* amask, HSetup& S) {
  if (!n || !x2c || nat < 1 || !xat || !ispc || nspc < 1 || !spc_ngrid || !spc_off || !spc_a || !spc_b || !spc_rmax || !spc_rcut ||
      !rtab || !ftab)
    return ctx->fail(C2G_ERR_ARG, "%s: bad argument", who);
  const long long nn = (long long)n[0] * n[1] * n[2];
  if (n[0] < 1 || n[1] < 1 || n[2] < 1 || nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "%s: bad grid shape", who);
  int ntab = 0;
  double rcutmax = 0.0;
  for (int i = 0; i < nspc; i++) {
    if (spc_ngrid[i] < 0 || (spc_ngrid[i] > 0 && spc_ngrid[i] < 4)) return ctx->fail(C2G_ERR_ARG, "%s: species %d: a radial grid needs at least 4 nodes", who, i + 1);
    if (spc_ngrid[i] > 0) {
      ntab = std::max(ntab, spc_off[i] + spc_ngrid[i]);
      rcutmax = std::max(rcutmax, spc_rcut[i]);
    }
  }
  for (int a = 0; a < nat; a++)
    if (ispc[a] < 1 || ispc[a] > nspc) return ctx->fail(C2G_ERR_ARG, "%s: atom %d has species %d", who, a + 1, ispc[a]);
  // inverse of x2c (adjugate) for the plane spacings
  auto A = [&](int i, int j) { return x2c[i + 3 * j]; };
  const double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                     A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
  if (det == 0.0) return ctx->fail(C2G_ERR_ARG, "%s: singular cell", who);
  const double d = 1.0 / det;
  double c2x[9];
  c2x[0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d; c2x[3] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d; c2x[6] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
  c2x[1] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d; c2x[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d; c2x[7] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
  c2x[2] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d; c2x[5] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d; c2x[8] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
  int m[3];
  for (int i = 0; i < 3; i++) m[i] = (int)std::ceil(rcutmax * std::sqrt(c2x[i] * c2x[i] + c2x[i + 3] * c2x[i + 3] + c2x[i + 6] * c2x[i + 6])) + 1;
  const long long nimg_ll = (long long)nat * (2 * m[0] + 1) * (2 * m[1] + 1) * (2 * m[2] + 1);
  if (nimg_ll > 50000000ll) return ctx->fail(C2G_ERR_ARG, "%s: %lld atom images (cutoff too large for the cell)", who, nimg_ll);
  std::vector<int> h_atom;
  std::vector<double> h_x;
  h_atom.reserve((size_t)nimg_ll);
  h_x.reserve(3 * (size_t)nimg_ll);
  for (int a = 0; a < nat; a++)
    for (int l1 = -m[0]; l1 <= m[0]; l1++)
      for (int l2 = -m[1]; l2 <= m[1]; l2++)
        for (int l3 = -m[2]; l3 <= m[2]; l3++) {
          const double xf[3] = {xat[3 * a] - std::floor(xat[3 * a]) + l1, xat[3 * a + 1] - std::floor(xat[3 * a + 1]) + l2,
                                xat[3 * a + 2] - std::floor(xat[3 * a + 2]) + l3};
          h_atom.push_back(a);
          for (int i = 0; i < 3; i++) h_x.push_back(x2c[i] * xf[0] + x2c[i + 3] * xf[1] + x2c[i + 6] * xf[2]);
        }
  std::vector<int> h_spc(nat);
  for (int a = 0; a < nat; a++) h_spc[a] = ispc[a] - 1;
  cudaStream_t st = ctx->stream;
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = b.alloc(ctx, std::max<size_t>(bytes, 1));
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st);
    return e;
  };
  C2G_CUDA(ctx, up(S.img_atom, h_atom.data(), sizeof(int) * h_atom.size()));
  C2G_CUDA(ctx, up(S.img_x, h_x.data(), sizeof(double) * h_x.size()));
  C2G_CUDA(ctx, up(S.atom_spc, h_spc.data(), sizeof(int) * nat));
  if (amask) C2G_CUDA(ctx, up(S.amask, amask, (size_t)nat));
  C2G_CUDA(ctx, up(S.ngrid, spc_ngrid, sizeof(int) * nspc));
  C2G_CUDA(ctx, up(S.off, spc_off, sizeof(int) * nspc));
  C2G_CUDA(ctx, up(S.a, spc_a, sizeof(double) * nspc));
  C2G_CUDA(ctx, up(S.b, spc_b, sizeof(double) * nspc));
  C2G_CUDA(ctx, up(S.rmax, spc_rmax, sizeof(double) * nspc));
  C2G_CUDA(ctx, up(S.rcut, spc_rcut, sizeof(double) * nspc));
  C2G_CUDA(ctx, up(S.rtab, rtab, sizeof(double) * ntab));
  C2G_CUDA(ctx, up(S.ftab, ftab, sizeof(double) * ntab));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));  // the host vectors go out of scope
  HParams& P = S.P;
  memset(&P, 0, sizeof(P));
  P.n1 = n[0]; P.n2 = n[1]; P.n3 = n[2]; P.nn = (unsigned)nn;
  for (int i = 0; i < 9; i++) P.x2c[i] = x2c[i];
  P.nimg = (int)h_atom.size();
  P.img_atom = S.img_atom.as<int>(); P.img_x = S.img_x.as<double>(); P.atom_spc = S.atom_spc.as<int>();
  P.amask = amask ? S.amask.as<unsigned char>() : nullptr;
  P.ngrid = S.ngrid.as<int>(); P.off = S.off.as<int>();
  P.a = S.a.as<double>(); P.b = S.b.as<double>(); P.rmax = S.rmax.as<double>(); P.rcut = S.rcut.as<double>();
  P.rtab = S.rtab.as<double>(); P.ftab = S.ftab.as<double>();
  return C2G_OK;
}

}  // namespace

extern "C" int c2g_promolecular_grid(c2g_context* ctx, const int n[3], const double x2c[9], int nat, const double* xat,
                                     const int* ispc, int nspc, const int* spc_ngrid, const int* spc_off, const double* spc_a,
                                     const double* spc_b, const double* spc_rmax, const double* spc_rcut, const double* rtab,
                                     const double* ftab, const unsigned char* infrag, int* handle) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_promolecular_grid");
  if (!handle) return ctx->fail(C2G_ERR_ARG, "c2g_promolecular_grid: null handle");
  HSetup S;
  int rc = h_setup(ctx, "c2g_promolecular_grid", n, x2c, nat, xat, ispc, nspc, spc_ngrid, spc_off, spc_a, spc_b, spc_rmax, spc_rcut, rtab,
                   ftab, infrag, S);
  if (rc != C2G_OK) return rc;
  rc = c2g_grid_alloc(ctx, n, handle);
  if (rc != C2G_OK) return rc;
  const unsigned ntiles = (S.P.nn + HB_THREADS - 1) / HB_THREADS;
  const int blocks = (int)std::min<unsigned>(ntiles, (unsigned)ctx->nsm * 8u);
  ctx->prof_begin("promolecular_atoms");
  k_promolecular<<<blocks, HB_THREADS, 0, ctx->stream>>>(S.P, ntiles, ctx->grids[*handle].d);
  ctx->prof_end();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    c2g_grid_free(ctx, *handle);
    *handle = -1;
    return ctx->fail(C2G_ERR_CUDA, "c2g_promolecular_grid: %s", cudaGetErrorString(e));
  }
  ctx->prof_collect();
  return C2G_OK;
}

extern "C" int c2g_hirshfeld_integrate(c2g_context* ctx, int hpromol, const double x2c[9], int nat, const double* xat,
                                       const int* ispc, int nspc, const int* spc_ngrid, const int* spc_off, const double* spc_a,
                                       const double* spc_b, const double* spc_rmax, const double* spc_rcut, const double* rtab,
                                       const double* ftab, const unsigned char* domask, int nprop, const int* fieldhandles,
                                       double omega, double* psum, double* vol) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_hirshfeld_integrate");
  if (hpromol < 0 || hpromol >= (int)ctx->grids.size() || !ctx->grids[hpromol].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_hirshfeld_integrate: invalid promolecular grid handle %d", hpromol);
  if (nprop < 0 || (nprop > 0 && (!fieldhandles || !psum))) return ctx->fail(C2G_ERR_ARG, "c2g_hirshfeld_integrate: bad argument");
  c2g_grid_ready(ctx, hpromol);
  const c2g_grid& gp = ctx->grids[hpromol];
  for (int k = 0; k < nprop; k++) {
    const int h = fieldhandles[k];
    if (h < 0 || h >= (int)ctx->grids.size() || !ctx->grids[h].used) return ctx->fail(C2G_ERR_ARG, "c2g_hirshfeld_integrate: invalid field handle %d", h);
    c2g_grid_ready(ctx, h);
    const c2g_grid& g = ctx->grids[h];
    if (g.n[0] != gp.n[0] || g.n[1] != gp.n[1] || g.n[2] != gp.n[2])
      return ctx->fail(C2G_ERR_ARG, "c2g_hirshfeld_integrate: field %d has a different grid size", k + 1);
  }
  if ((long long)nat * 5 > HB_TABLE_MAX) return ctx->fail(C2G_ERR_ARG, "c2g_hirshfeld_integrate: more than %d atoms", HB_TABLE_MAX / 5);
  HSetup S;
  int rc = h_setup(ctx, "c2g_hirshfeld_integrate", gp.n, x2c, nat, xat, ispc, nspc, spc_ngrid, spc_off, spc_a, spc_b, spc_rmax, spc_rcut,
                   rtab, ftab, domask, S);
  if (rc != C2G_OK) return rc;
  const unsigned ntiles = (S.P.nn + HB_THREADS - 1) / HB_THREADS;
  const int blocks = (int)std::min<unsigned>(ntiles, (unsigned)ctx->nsm * 4u);
  const double ntot = (double)gp.nn;
  bool vol_done = false;
  for (int k0 = 0; k0 < std::max(nprop, 1); k0 += 4) {
    const int np = std::max(0, std::min(4, nprop - k0));
    const double* fp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int p = 0; p < np; p++) fp[p] = ctx->grids[fieldhandles[k0 + p]].d;
    DevBuf b_sums(ctx);
    const size_t nsum = (size_t)nat * (np + 1);
    C2G_CUDA(ctx, b_sums.alloc(ctx, sizeof(double) * nsum));
    C2G_CUDA(ctx, cudaMemsetAsync(b_sums.p, 0, sizeof(double) * nsum, ctx->stream));
    const size_t smem = sizeof(double) * nsum;
    ctx->prof_begin("hirshfeld_sums");
    switch (np) {
      case 0: k_hirshfeld<0><<<blocks, HB_THREADS, smem, ctx->stream>>>(S.P, ntiles, nat, gp.d, fp[0], fp[1], fp[2], fp[3], b_sums.as<double>()); break;
      case 1: k_hirshfeld<1><<<blocks, HB_THREADS, smem, ctx->stream>>>(S.P, ntiles, nat, gp.d, fp[0], fp[1], fp[2], fp[3], b_sums.as<double>()); break;
      case 2: k_hirshfeld<2><<<blocks, HB_THREADS, smem, ctx->stream>>>(S.P, ntiles, nat, gp.d, fp[0], fp[1], fp[2], fp[3], b_sums.as<double>()); break;
      case 3: k_hirshfeld<3><<<blocks, HB_THREADS, smem, ctx->stream>>>(S.P, ntiles, nat, gp.d, fp[0], fp[1], fp[2], fp[3], b_sums.as<double>()); break;
      default: k_hirshfeld<4><<<blocks, HB_THREADS, smem, ctx->stream>>>(S.P, ntiles, nat, gp.d, fp[0], fp[1], fp[2], fp[3], b_sums.as<double>()); break;
    }
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    std::vector<double> hs(nsum);
    C2G_CUDA(ctx, cudaMemcpyAsync(hs.data(), b_sums.p, sizeof(double) * nsum, cudaMemcpyDeviceToHost, ctx->stream));
    C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int a = 0; a < nat; a++) {
      for (int p = 0; p < np; p++) psum[a + (size_t)nat * (k0 + p)] = hs[(size_t)a * (np + 1) + p] * omega / ntot;  // :1590
      if (!vol_done && vol) vol[a] = hs[(size_t)a * (np + 1) + np] * omega / ntot;
    }
    vol_done = true;
  }
  ctx->prof_collect();
  return C2G_OK;
}

// =================================================================================================
// VORONOI on a grid: voronoi_grid (src/hirshfeld@proc.f90:93-122) = crystal%nearest_atom_grid
// (src/crystalmod@proc.f90:1138-1167): idg(i,j,k) = complete-list id of the atom nearest to the grid node.
// The reference asks its block environment for the nearest atom of every node (nearest_atom ->
// list_near_atoms(up2n = 1), crystalmod@env.f90:516-563).  Here the host enumerates every atom image within half the
// longest body diagonal of the cell (no point of the cell is farther than that from an image of ANY atom); a block
// owns a tile of 256 consecutive points, warp 0 finds the image nearest to the tile's reference point (d0) and keeps
// the images within d0 + 2 * (tile radius) of it -- a superset of the nearest image of every point of the tile -- and
// every lane runs over the survivors.
// Distances are compared as the reference computes them (Cartesian difference, sum of squares in x, y, z order);
// TIES between equidistant atoms go to the lower atom id.  The reference's choice among exactly equidistant atoms
// follows the block traversal and merge sort of list_near_atoms, which is not restated: parity is claimed for nodes
// whose nearest atom is unique, ties are counted by the tests.
// =================================================================================================
namespace {
struct VParams {
  int n1, n2, n3;
  unsigned nn;
  double x2c[9];
  int nimg;
  const int* img_atom;
  const double* img_x;
};
__global__ void __launch_bounds__(HB_THREADS) k_voronoi(const __grid_constant__ VParams V, unsigned ntiles, int* __restrict__ label) {
  __shared__ Cull c;
  __shared__ int s_list[HB_LIST];
  __shared__ double s_d0;
  HParams P;  // tile_point only reads the grid shape and the cell
  P.n1 = V.n1; P.n2 = V.n2; P.n3 = V.n3; P.nn = V.nn;
  for (int i = 0; i < 9; i++) P.x2c[i] = V.x2c[i];
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = tile_point(P, tile);
    const double radius = tile_sphere(c, t);
    // distance of the reference point to its nearest image
    if (threadIdx.x < 32) {
      double best = 1e300;
      for (int m = threadIdx.x; m < V.nimg; m += 32) {
        const double dx = __ldg(V.img_x + 3 * m) - c.cx, dy = __ldg(V.img_x + 3 * m + 1) - c.cy, dz = __ldg(V.img_x + 3 * m + 2) - c.cz;
        best = fmin(best, dx * dx + dy * dy + dz * dz);
      }
      for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (threadIdx.x == 0) s_d0 = sqrt(best);
    }
    __syncthreads();
    const double lim = (s_d0 + 2.0 * radius) * 1.000001 + 1e-9;
    double dbest = 1e300;
    int abest = 0x7fffffff;
    bool more;
    do {  // culling rounds of HB_LIST images, in list order
      if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int cnt = 0, base = c.next;
        __syncwarp();
        for (; base < V.nimg && cnt + 32 <= HB_LIST; base += 32) {
          const int m = base + lane;
          bool keep = false;
          if (m < V.nimg) {
            const double dx = __ldg(V.img_x + 3 * m) - c.cx, dy = __ldg(V.img_x + 3 * m + 1) - c.cy, dz = __ldg(V.img_x + 3 * m + 2) - c.cz;
            keep = dx * dx + dy * dy + dz * dz <= lim * lim;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (keep) s_list[cnt + __popc(bal & ((1u << lane) - 1u))] = m;
          cnt += __popc(bal);
        }
        if (lane == 0) { c.cnt = cnt; c.next = base; }
      }
      __syncthreads();
      const int cnt = c.cnt;
      more = c.next < V.nimg;  // read while nobody writes it: warp 0 rewrites c.next at the top of the next round
      for (int q = 0; q < cnt; q++) {
        const int m = s_list[q];
        const double dx = t.x - __ldg(V.img_x + 3 * m), dy = t.y - __ldg(V.img_x + 3 * m + 1), dz = t.z - __ldg(V.img_x + 3 * m + 2);
        const double d2 = dx * dx + dy * dy + dz * dz;
        const int a = __ldg(V.img_atom + m);
        if (d2 < dbest || (d2 == dbest && a < abest)) { dbest = d2; abest = a; }
      }
      __syncthreads();
    } while (more);
    if (t.valid) label[t.i] = abest;
    __syncthreads();
  }
}
}  // namespace

extern "C" int c2g_voronoi_grid(c2g_context* ctx, const int n[3], const double x2c[9], int nat, const double* xat, c2g_basins** res_out) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_voronoi_grid");
  if (!n || !x2c || nat < 1 || !xat || !res_out) return ctx->fail(C2G_ERR_ARG, "c2g_voronoi_grid: bad argument");
  const long long nn = (long long)n[0] * n[1] * n[2];
  if (n[0] < 1 || n[1] < 1 || n[2] < 1 || nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "c2g_voronoi_grid: bad grid shape");
  auto A = [&](int i, int j) { return x2c[i + 3 * j]; };
  const double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                     A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
  if (det == 0.0) return ctx->fail(C2G_ERR_ARG, "c2g_voronoi_grid: singular cell");
  // reciprocal rows (plane spacings) and the longest body diagonal
  const double d = 1.0 / det;
  double c2x[9];
  c2x[0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d; c2x[3] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d; c2x[6] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
  c2x[1] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d; c2x[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d; c2x[7] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
  c2x[2] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d; c2x[5] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d; c2x[8] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
  double diag = 0.0;
  for (int s1 = -1; s1 <= 1; s1 += 2)
    for (int s2 = -1; s2 <= 1; s2 += 2) {
      double v[3];
      for (int i = 0; i < 3; i++) v[i] = x2c[i] + s1 * x2c[i + 3] + s2 * x2c[i + 6];
      diag = std::max(diag, std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    }
  const double reach = 0.5 * diag * 1.0001;
  int m[3];
  for (int i = 0; i < 3; i++) m[i] = (int)std::ceil(reach * std::sqrt(c2x[i] * c2x[i] + c2x[i + 3] * c2x[i + 3] + c2x[i + 6] * c2x[i + 6])) + 1;
  const long long nimg_ll = (long long)nat * (2 * m[0] + 1) * (2 * m[1] + 1) * (2 * m[2] + 1);
  if (nimg_ll > 50000000ll) return ctx->fail(C2G_ERR_ARG, "c2g_voronoi_grid: %lld atom images", nimg_ll);
  std::vector<int> h_atom;
  std::vector<double> h_x;
  for (int a = 0; a < nat; a++)
    for (int l1 = -m[0]; l1 <= m[0]; l1++)
      for (int l2 = -m[1]; l2 <= m[1]; l2++)
        for (int l3 = -m[2]; l3 <= m[2]; l3++) {
          const double xf[3] = {xat[3 * a] - std::floor(xat[3 * a]) + l1, xat[3 * a + 1] - std::floor(xat[3 * a + 1]) + l2,
                                xat[3 * a + 2] - std::floor(xat[3 * a + 2]) + l3};
          h_atom.push_back(a);
          for (int i = 0; i < 3; i++) h_x.push_back(x2c[i] * xf[0] + x2c[i + 3] * xf[1] + x2c[i + 6] * xf[2]);
        }
  cudaStream_t st = ctx->stream;
  DevBuf b_atom, b_x;
  C2G_CUDA(ctx, b_atom.alloc(ctx, sizeof(int) * h_atom.size()));
  C2G_CUDA(ctx, b_x.alloc(ctx, sizeof(double) * h_x.size()));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_atom.p, h_atom.data(), sizeof(int) * h_atom.size(), cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_x.p, h_x.data(), sizeof(double) * h_x.size(), cudaMemcpyHostToDevice, st));
  c2g_basins* res = new c2g_basins();
  res->ctx = ctx; res->kind = 2; res->gridh = -1;   // plain labels like ISOSURFACE regions: one "maximum" per atom
  res->n[0] = n[0]; res->n[1] = n[1]; res->n[2] = n[2]; res->nn = nn;
  res->zlo = 0; res->zhi = n[2];
  res->nmax = nat;
  res->max_lin.assign(nat, 0);
  for (int a = 0; a < nat; a++) {  // grid node nearest to the atom (for c2g_basins_maxima; the attractors are the atoms themselves)
    long long id = 0, stride = 1;
    for (int i = 0; i < 3; i++) {
      const double f = xat[3 * a + i] - std::floor(xat[3 * a + i]);
      const long long k = ((long long)std::llround(f * n[i])) % n[i];
      id += k * stride; stride *= n[i];
    }
    res->max_lin[a] = (int)id;
  }
  struct Guard { c2g_basins* r; bool ok = false; ~Guard() { if (!ok) c2g_basins_free(r); } } guard{res};
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&res->d_label, sizeof(int) * nn));
  VParams V;
  V.n1 = n[0]; V.n2 = n[1]; V.n3 = n[2]; V.nn = (unsigned)nn;
  for (int i = 0; i < 9; i++) V.x2c[i] = x2c[i];
  V.nimg = (int)h_atom.size(); V.img_atom = b_atom.as<int>(); V.img_x = b_x.as<double>();
  const unsigned ntiles = (V.nn + HB_THREADS - 1) / HB_THREADS;
  ctx->prof_begin("voronoi_nearest_atom");
  k_voronoi<<<std::min<unsigned>(ntiles, (unsigned)ctx->nsm * 8u), HB_THREADS, 0, st>>>(V, ntiles, res->d_label);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  C2G_CUDA(ctx, cudaStreamSynchronize(st));  // the host image vectors go out of scope
  std::vector<int> ident(nat);
  for (int a = 0; a < nat; a++) ident[a] = a + 1;
  int rc = c2g_basins_set_map(res, nat, ident.data());
  if (rc != C2G_OK) return rc;
  ctx->prof_collect();
  guard.ok = true;
  *res_out = res;
  return C2G_OK;
}
