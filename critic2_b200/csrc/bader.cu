// bader.cu -- Bader near-grid basin assignment on sm_100a.
//
// Replaces the scan + refine_edge body of bader_integrate (critic2 src/bader@proc.f90:147-224).
// The reference is a sequential, order-dependent scan; what it converges to (verified against the
// faithful CPU restatement in oracle/) is the labelling in which every grid point carries the
// terminal maximum of ITS OWN near-grid trajectory started with dr = 0:
//     step_neargrid  bader@proc.f90:455-494     rho_grad_dir  :532-567
//     step_ongrid    :500-527                   is_max        :571-597
// This file computes exactly that labelling:
//   C2G_BADER_EXACT  every point walks its complete trajectory (the on-device referee);
//   C2G_BADER_FAST   hierarchical: points of the stride-4 lattice walk; a stride-s cube whose 8
//                    corners agree and that contains no local maximum is filled, the other points
//                    walk; finally every filled point that has a 26-neighbour with a different
//                    label walks too, until none is left -- the same fixed-point condition the
//                    reference's refine_edge enforces (:300-422: every edge point carries the
//                    label of its own trajectory).
// Arithmetic: IEEE fp64, evaluation order of the Fortran source, NO fused multiply-add (this
// translation unit is compiled with -fmad=false; the reference is built -O3 without -march/-ffast-math).
#include "common.cuh"

#include <algorithm>

namespace {

struct BaderParams {
  int n1, n2, n3;
  double c2l[9];   // car2lat, column-major
  double lid[27];  // lat_i_dist, (d1+1)*9+(d2+1)*3+(d3+1)
};

constexpr int PATHCAP = 640;            // per-thread path buffer (local memory) of the fast walker
constexpr unsigned FILLBIT = 0x80000000u;
constexpr int LMASK = 0x7fffffff;

__device__ __forceinline__ int wrapx(int p, int n) {
  if (p < 0) p += n;
  if (p >= n) p -= n;
  if (p < 0 || p >= n) p = ((p % n) + n) % n;  // grids narrower than a tile
  return p;
}
__device__ __forceinline__ int wrapc(int p, int n) {
  while (p < 0) p += n;
  while (p >= n) p -= n;
  return p;
}

// open-addressing hash: linear id of a candidate maximum -> candidate index
struct MaxHash {
  const int* keys;  // -1 = empty
  const int* vals;
  unsigned mask;
};
__device__ __forceinline__ int hash_lookup(const MaxHash& h, int key) {
  unsigned s = ((unsigned)key * 2654435761u) & h.mask;
  for (;;) {
    const int k = __ldg(h.keys + s);
    if (k == key) return __ldg(h.vals + s);
    if (k < 0) return -1;
    s = (s + 1) & h.mask;
  }
}

// wrap of a coordinate that is at most 2 cells outside [0,n) (pbc, bader@proc.f90:601-617)
__device__ __forceinline__ int wrap2(int p, int n) {
  if (p < 0) p += n;
  if (p < 0) p += n;
  if (p >= n) p -= n;
  if (p >= n) p -= n;
  return p;
}
// Fortran nint (half away from zero) for |v| < 1.5, as a double
__device__ __forceinline__ double nint_small(double v) { return v >= 0.5 ? 1.0 : (v <= -0.5 ? -1.0 : 0.0); }

// is_max, bader@proc.f90:571-597 (no neighbour strictly greater)
__device__ __noinline__ bool dev_is_max(const BaderParams& P, const double* __restrict__ rho, int x, int y, int z, double r0) {
  bool ismax = true;
#pragma unroll 1
  for (int d3 = -1; d3 <= 1; d3++) {
    const int zz = wrap2(z + d3, P.n3);
#pragma unroll 1
    for (int d2 = -1; d2 <= 1; d2++) {
      const int yy = wrap2(y + d2, P.n2);
      const int base = P.n1 * (yy + P.n2 * zz);
#pragma unroll
      for (int d1 = -1; d1 <= 1; d1++) {
        const int xx = wrap2(x + d1, P.n1);
        if (__ldg(rho + base + xx) > r0) ismax = false;
      }
    }
  }
  return ismax;
}

// step_ongrid, bader@proc.f90:500-527.  Loop order d1 (outer), d2, d3 (inner), first strictly greater wins.
// Returns the linear id of the chosen point.
__device__ __noinline__ int dev_step_ongrid(const BaderParams& P, const double* __restrict__ rho, int x, int y, int z,
                                            double rho_ctr) {
  double rho_max = rho_ctr;
  int best = x + P.n1 * (y + P.n2 * z);
#pragma unroll 1
  for (int d1 = -1; d1 <= 1; d1++) {
    const int xx = wrap2(x + d1, P.n1);
#pragma unroll 1
    for (int d2 = -1; d2 <= 1; d2++) {
      const int yy = wrap2(y + d2, P.n2);
#pragma unroll
      for (int d3 = -1; d3 <= 1; d3++) {
        const int zz = wrap2(z + d3, P.n3);
        const int q = xx + P.n1 * (yy + P.n2 * zz);
        double rho_tmp = __ldg(rho + q);
        rho_tmp = rho_ctr + (rho_tmp - rho_ctr) * P.lid[(d1 + 1) * 9 + (d2 + 1) * 3 + (d3 + 1)];
        if (rho_tmp > rho_max) {
          rho_max = rho_tmp;
          best = q;
        }
      }
    }
  }
  return best;
}

__device__ __noinline__ bool dev_on_path(const int* path, int len, int nid) {
  for (int j = len - 1; j >= 0; j--)
    if (path[j] == nid) return true;
  return false;
}

// Early-termination map (FAST algorithm only): safe[c] >= 0 means the stride-2^shift cube c and its 26
// neighbouring cubes are uniformly labelled with the terminal maximum safe[c] and hold no maximum of
// their own -- the analogue of the reference's known==2 "interior" points at which max_neargrid stops
// (bader@proc.f90:447), with a wider margin.  Covers the owned cube layers only.
struct SafeMap {
  const int* safe;  // nullptr = disabled
  int shift, c1, c2, zlo, nzl;
};

// One complete near-grid trajectory (max_neargrid, bader@proc.f90:427-450 on a fresh grid).
// Returns the linear id of the terminal maximum, or -1 if the path buffer overflowed.
// The reference's "known(pm)==1" revisit test (:484-488) is answered exactly: a visited point
// can only be hit again when rho(pm) <= max rho along the path, and only then the stored path is
// searched.  ORTHO: car2lat is diagonal (orthogonal cell); the skipped products are exact zeros,
// so the result is bit-identical to the general expression.
template <bool ORTHO>
__device__ __forceinline__ int dev_walk(const BaderParams& P, const double* __restrict__ rho, const MaxHash& h,
                                        const SafeMap& sm, int start, int* path, int cap,
                                        unsigned long long* nsteps_out) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int s2 = n1, s3 = n1 * n2;
  const int wxp = 1 - n1, wxm = n1 - 1, wyp = s2 - s3, wym = s3 - s2, wzp = s3 - s3 * n3, wzm = s3 * n3 - s3;
  int x = start % n1;
  int t = start / n1;
  int y = t % n2;
  int z = t / n2;
  int id = start;
  double dr0 = 0.0, dr1 = 0.0, dr2 = 0.0;
  double rhomax = -1.0e300;
  int len = 0;
  double r0 = __ldg(rho + id);
  for (;;) {
    const double* c = rho + id;
    const double rxp = __ldg(c + ((x + 1 == n1) ? wxp : 1)), rxm = __ldg(c + ((x == 0) ? wxm : -1));
    const double ryp = __ldg(c + ((y + 1 == n2) ? wyp : s2)), rym = __ldg(c + ((y == 0) ? wym : -s2));
    const double rzp = __ldg(c + ((z + 1 == n3) ? wzp : s3)), rzm = __ldg(c + ((z == 0) ? wzm : -s3));
    // rho_grad_dir (:532-567)
    double gl0 = (rxp - rxm) * 0.5, gl1 = (ryp - rym) * 0.5, gl2 = (rzp - rzm) * 0.5;
    if (rxp < r0 && rxm < r0) gl0 = 0.0;
    if (ryp < r0 && rym < r0) gl1 = 0.0;
    if (rzp < r0 && rzm < r0) gl2 = 0.0;
    double g0, g1, g2;
    if (ORTHO) {
      g0 = P.c2l[0] * (gl0 * P.c2l[0]);
      g1 = P.c2l[4] * (gl1 * P.c2l[4]);
      g2 = P.c2l[8] * (gl2 * P.c2l[8]);
    } else {
      const double gc0 = gl0 * P.c2l[0] + gl1 * P.c2l[1] + gl2 * P.c2l[2];
      const double gc1 = gl0 * P.c2l[3] + gl1 * P.c2l[4] + gl2 * P.c2l[5];
      const double gc2 = gl0 * P.c2l[6] + gl1 * P.c2l[7] + gl2 * P.c2l[8];
      g0 = P.c2l[0] * gc0 + P.c2l[3] * gc1 + P.c2l[6] * gc2;
      g1 = P.c2l[1] * gc0 + P.c2l[4] * gc1 + P.c2l[7] * gc2;
      g2 = P.c2l[2] * gc0 + P.c2l[5] * gc1 + P.c2l[8] * gc2;
    }
    const double gmax = fmax(fabs(g0), fmax(fabs(g1), fabs(g2)));
    int nid, nx, ny, nz;
    if (gmax < 1e-30) {  // (:468-476)
      dr0 = dr1 = dr2 = 0.0;
      // is_max (:571-597) == membership in the candidate list built by k_maxima with the same predicate
      if (hash_lookup(h, id) >= 0) break;
      nid = dev_step_ongrid(P, rho, x, y, z, r0);
      nx = nid % n1; t = nid / n1; ny = t % n2; nz = t / n2;
    } else {  // (:477-483)
      const double coeff = 1.0 / gmax;
      g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
      const double a0 = nint_small(g0), a1 = nint_small(g1), a2 = nint_small(g2);
      dr0 = dr0 + g0 - a0; dr1 = dr1 + g1 - a1; dr2 = dr2 + g2 - a2;
      const double b0 = nint_small(dr0), b1 = nint_small(dr1), b2 = nint_small(dr2);
      dr0 = dr0 - b0; dr1 = dr1 - b1; dr2 = dr2 - b2;
      nx = wrap2(x + (int)(a0 + b0), n1);
      ny = wrap2(y + (int)(a1 + b1), n2);
      nz = wrap2(z + (int)(a2 + b2), n3);
      nid = nx + n1 * (ny + n2 * nz);
    }
    // known(p) = 1 (:484)
    if (len >= cap) return -1;
    path[len++] = id;
    rhomax = fmax(rhomax, r0);
    double rn = __ldg(rho + nid);
    if (rn <= rhomax) {  // only then pm can be a point of this path (:487)
      if (dev_on_path(path, len, nid)) {
        nid = dev_step_ongrid(P, rho, x, y, z, r0);
        nx = nid % n1; t = nid / n1; ny = t % n2; nz = t / n2;
        dr0 = dr1 = dr2 = 0.0;
        rn = __ldg(rho + nid);
      }
    }
    if (nid == id) break;  // did not move: maximum (:439)
    id = nid;
    r0 = rn;
    x = nx; y = ny; z = nz;
    if (sm.safe) {  // quit at a known interior point (:447)
      const int cz = nz - sm.zlo;
      if (cz >= 0 && cz < sm.nzl) {
        const int sl = __ldg(sm.safe + (nx >> sm.shift) + sm.c1 * ((ny >> sm.shift) + sm.c2 * (cz >> sm.shift)));
        if (sl >= 0) { id = sl; break; }
      }
    }
  }
  if (nsteps_out) atomicAdd(nsteps_out, (unsigned long long)len);
  return id;
}

// ------------------------------------------------------------------------------------------------
// z-slab bookkeeping.  A rank owns the global planes [zlo, zhi) (boundaries are multiples of 4 so
// that no stride-4 cube straddles two ranks); rho is replicated on every rank, labels are sharded.
// The label buffer holds nzl + 2 planes: local plane 0 = global plane zlo-1 (halo below), planes
// 1..nzl = owned, plane nzl+1 = global plane zhi (halo above); halos are periodic images and are
// filled by Exchange (NCCL send/recv between ranks, a device copy when the rank is its own neighbour).
// ------------------------------------------------------------------------------------------------
struct Slab {
  int zlo, zhi, nzl;
};

// K0: candidate maxima (26-neighbour, is_max) by a separable 3x3x3 box maximum on shared-memory
// tiles with a periodic 1-cell halo; marks the stride-4 / stride-2 cubes that contain a maximum.
constexpr int TX = 32, TY = 8, TZ = 8;
__global__ void __launch_bounds__(256) k_maxima(const __grid_constant__ BaderParams P, const Slab S,
                                                const double* __restrict__ rho, int* __restrict__ cand,
                                                int* __restrict__ ncand, int maxcand, unsigned char* __restrict__ cube4,
                                                unsigned char* __restrict__ cube2) {
  extern __shared__ double sm[];
  double* s0 = sm;                                   // [TZ+2][TY+2][TX+2]
  double* s1 = sm + (TZ + 2) * (TY + 2) * (TX + 2);  // [TZ+2][TY+2][TX] max over x
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int bx0 = blockIdx.x * TX, by0 = blockIdx.y * TY, bz0 = S.zlo + blockIdx.z * TZ;
  const int tid = threadIdx.x;
  constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2;
  {  // row-wise staging: one warp per (y,z) row of 34 values, then the 3-point maximum along x
    const int lane = tid & 31, wid = tid >> 5;
    const int gx0 = wrapx(bx0 - 1 + lane, n1), gx1 = wrapx(bx0 - 1 + 32 + (lane & 1), n1);
    for (int r = wid; r < SY * SZ; r += 8) {
      const int sy = r % SY, sz = r / SY;
      const int gy = wrapx(by0 + sy - 1, n2), gz = wrapx(bz0 + sz - 1, n3);
      const double* row = rho + (size_t)n1 * (gy + (size_t)n2 * gz);
      s0[r * SX + lane] = __ldg(row + gx0);
      if (lane < 2) s0[r * SX + 32 + lane] = __ldg(row + gx1);
    }
    __syncthreads();
    for (int r = wid; r < SY * SZ; r += 8) {
      const double* p = s0 + r * SX + lane;
      s1[r * TX + lane] = fmax(p[0], fmax(p[1], p[2]));
    }
    __syncthreads();
  }
  const int lx = tid % TX, ly = tid / TX;
  const int gx = bx0 + lx, gy = by0 + ly;
  double m[SZ];
#pragma unroll
  for (int sz = 0; sz < SZ; sz++) {
    const double* p = s1 + (sz * SY + ly) * TX + lx;
    m[sz] = fmax(p[0], fmax(p[TX], p[2 * TX]));
  }
  if (gx < n1 && gy < n2) {
#pragma unroll
    for (int lz = 0; lz < TZ; lz++) {
      const int gz = bz0 + lz;
      if (gz >= S.zhi) break;
      const double c = s0[((lz + 1) * SY + (ly + 1)) * SX + lx + 1];
      const double bm = fmax(m[lz], fmax(m[lz + 1], m[lz + 2]));
      if (!(bm > c)) {  // no neighbour strictly greater
        const int slot = atomicAdd(ncand, 1);
        if (slot < maxcand) cand[slot] = gx + n1 * (gy + n2 * gz);
        const int c41 = (n1 + 3) / 4, c42 = (n2 + 3) / 4;
        const int c21 = (n1 + 1) / 2, c22 = (n2 + 1) / 2;
        cube4[(gx >> 2) + c41 * ((gy >> 2) + (size_t)c42 * ((gz - S.zlo) >> 2))] = 1;
        cube2[(gx >> 1) + c21 * ((gy >> 1) + (size_t)c22 * ((gz - S.zlo) >> 1))] = 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// walkers.  `label` is the owned part of the label buffer re-based so that label[global id] is valid
// for every owned point.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void finish_walk(int start, int term, int* __restrict__ label, const MaxHash& h,
                                            unsigned char* __restrict__ reached, int* __restrict__ overflow,
                                            int* __restrict__ noverflow, int* __restrict__ err) {
  if (term < 0) {
    const int slot = atomicAdd(noverflow, 1);
    overflow[slot] = start;
    return;
  }
  label[start] = term;
  const int ci = hash_lookup(h, term);
  if (ci < 0) atomicExch(err, 1);  // terminal is not a candidate maximum: cannot happen
  else if (!reached[ci]) reached[ci] = 1;
}

// every owned point of the stride-s lattice (s = 1: every grid point = the EXACT referee)
template <bool ORTHO>
__global__ void __launch_bounds__(128) k_walk_lattice(const __grid_constant__ BaderParams P, const Slab S,
                                                      const double* __restrict__ rho, int s, int m1, int m2, int m3,
                                                      int* __restrict__ label, MaxHash h, unsigned char* __restrict__ reached,
                                                      int* __restrict__ overflow, int* __restrict__ noverflow,
                                                      int* __restrict__ err, unsigned long long* nsteps) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m1 * m2 * m3) return;
  const int lx = (int)(t % m1), ly = (int)((t / m1) % m2), lz = (int)(t / ((long long)m1 * m2));
  const int start = lx * s + P.n1 * (ly * s + P.n2 * (S.zlo + lz * s));
  int path[PATHCAP];
  const SafeMap nosafe{nullptr, 0, 0, 0, 0, 0};
  const int term = dev_walk<ORTHO>(P, rho, h, nosafe, start, path, PATHCAP, nsteps);
  finish_walk(start, term, label, h, reached, overflow, noverflow, err);
}

template <bool ORTHO>
__global__ void __launch_bounds__(128) k_walk_list(const __grid_constant__ BaderParams P, const double* __restrict__ rho,
                                                   const int* __restrict__ list, int count, int* __restrict__ label,
                                                   MaxHash h, SafeMap sm, unsigned char* __restrict__ reached,
                                                   int* __restrict__ overflow, int* __restrict__ noverflow,
                                                   int* __restrict__ err, unsigned long long* nsteps) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int start = list[t];
  int path[PATHCAP];
  const int term = dev_walk<ORTHO>(P, rho, h, sm, start, path, PATHCAP, nsteps);
  finish_walk(start, term, label, h, reached, overflow, noverflow, err);
}

// rare long trajectories: path buffer in global memory (bigcap entries per walker)
__global__ void __launch_bounds__(64) k_walk_big(const __grid_constant__ BaderParams P, const double* __restrict__ rho,
                                                 const int* __restrict__ list, int count, int* __restrict__ label,
                                                 MaxHash h, unsigned char* __restrict__ reached, int* __restrict__ scratch,
                                                 int bigcap, int* __restrict__ err, unsigned long long* nsteps) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int start = list[t];
  const SafeMap nosafe{nullptr, 0, 0, 0, 0, 0};
  const int term = dev_walk<false>(P, rho, h, nosafe, start, scratch + (size_t)t * bigcap, bigcap, nsteps);
  if (term < 0) { atomicExch(err, 2); return; }
  label[start] = term;
  const int ci = hash_lookup(h, term);
  if (ci < 0) atomicExch(err, 1);
  else reached[ci] = 1;
}

// ------------------------------------------------------------------------------------------------
// classify: one thread per owned stride-s cube.  If its 8 corners (already labelled; the upper ones
// may sit on the halo-above plane) agree and it holds no local maximum, fill its new stride-s/2
// points (label | FILLBIT); otherwise queue them.  lbuf = label buffer including halos.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_classify(int n1, int n2, int n3, const Slab S, int s, int* __restrict__ lbuf,
                                                  const unsigned char* __restrict__ cubemax, int* __restrict__ cubeuni,
                                                  int* __restrict__ list, int* __restrict__ nlist) {
  const int c1 = (n1 + s - 1) / s, c2 = (n2 + s - 1) / s, c3 = (S.nzl + s - 1) / s;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < (long long)c1 * c2 * c3;
  int npush = 0;
  int pts[7];
  if (active) {
    const int cx = (int)(t % c1), cy = (int)((t / c1) % c2), cz = (int)(t / ((long long)c1 * c2));
    const int x0 = cx * s, y0 = cy * s, z0 = S.zlo + cz * s;
    const int x1 = (x0 + s < n1) ? x0 + s : 0, y1 = (y0 + s < n2) ? y0 + s : 0;
    const int p0 = z0 - S.zlo + 1;                              // local plane of z0
    const int p1 = ((z0 + s < n3) ? z0 + s : n3) - S.zlo + 1;   // local plane of the upper corners (may be the halo)
    const int s3 = n1 * n2;
    const int l000 = lbuf[x0 + n1 * y0 + s3 * p0] & LMASK;
    bool uni = !cubemax[t];
    uni = uni && ((lbuf[x1 + n1 * y0 + s3 * p0] & LMASK) == l000);
    uni = uni && ((lbuf[x0 + n1 * y1 + s3 * p0] & LMASK) == l000);
    uni = uni && ((lbuf[x1 + n1 * y1 + s3 * p0] & LMASK) == l000);
    uni = uni && ((lbuf[x0 + n1 * y0 + s3 * p1] & LMASK) == l000);
    uni = uni && ((lbuf[x1 + n1 * y0 + s3 * p1] & LMASK) == l000);
    uni = uni && ((lbuf[x0 + n1 * y1 + s3 * p1] & LMASK) == l000);
    uni = uni && ((lbuf[x1 + n1 * y1 + s3 * p1] & LMASK) == l000);
    cubeuni[t] = uni ? l000 : -1;
    const int hh = s >> 1;
    for (int o = 1; o < 8; o++) {
      const int x = x0 + ((o & 1) ? hh : 0), y = y0 + ((o & 2) ? hh : 0), z = z0 + ((o & 4) ? hh : 0);
      if (x >= n1 || y >= n2 || z >= S.zhi) continue;
      if (uni) lbuf[x + n1 * y + s3 * (z - S.zlo + 1)] = (int)((unsigned)l000 | FILLBIT);
      else pts[npush++] = x + n1 * (y + n2 * z);  // global id
    }
  }
  // block-aggregated append (keeps spatial order inside a block)
  __shared__ int s_base;
  __shared__ int s_warp[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = npush;
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int w = 0; w < 8; w++) { const int v = s_warp[w]; s_warp[w] = acc; acc += v; }
    s_base = acc ? atomicAdd(nlist, acc) : 0;
  }
  __syncthreads();
  if (npush) {
    int off = s_base + s_warp[wid] + incl - npush;
    for (int k = 0; k < npush; k++) list[off + k] = pts[k];
  }
}

// safe[c] = cubeuni[c] if the 26 neighbouring cubes (periodic in x,y; owned layers only in z) carry the
// same uniform label, else -1
__global__ void __launch_bounds__(256) k_safe(int c1, int c2, int c3, const int* __restrict__ cubeuni, int* __restrict__ safe) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)c1 * c2 * c3) return;
  const int v = cubeuni[t];
  int out = v;
  if (v >= 0) {
    const int cx = (int)(t % c1), cy = (int)((t / c1) % c2), cz = (int)(t / ((long long)c1 * c2));
    for (int dz = -1; dz <= 1 && out >= 0; dz++) {
      const int zz = cz + dz;
      if (zz < 0 || zz >= c3) { out = -1; break; }
      for (int dy = -1; dy <= 1 && out >= 0; dy++) {
        const int yy = (cy + dy + c2) % c2;
        for (int dx = -1; dx <= 1; dx++) {
          const int xx = (cx + dx + c1) % c1;
          if (cubeuni[xx + c1 * (yy + (size_t)c2 * zz)] != v) { out = -1; break; }
        }
      }
    }
  }
  safe[t] = out;
}

// ------------------------------------------------------------------------------------------------
// edge fix: every FILLED owned point with a 26-neighbour of a different label is queued for an exact
// walk (the refine_edge criterion, is_vol_edge bader@proc.f90:730-752).  Reads both halo planes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edgefix(int n1, int n2, const Slab S, int* __restrict__ lbuf,
                                                 int* __restrict__ list, int* __restrict__ nlist,
                                                 const unsigned char* __restrict__ dirty_in, unsigned char* __restrict__ dirty_out) {
  const int tile = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (dirty_in && !dirty_in[tile]) return;  // nothing changed near this tile in the previous pass
  __shared__ int sl[(TZ + 2) * (TY + 2) * (TX + 2)];
  constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2;
  const int bx0 = blockIdx.x * TX, by0 = blockIdx.y * TY, lz0 = blockIdx.z * TZ;  // lz0: first owned plane (0-based)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int s3 = n1 * n2;
  // row-wise staging: one warp per (y,z) row of 34 labels
  const int gx0 = wrapx(bx0 - 1 + lane, n1);
  const int gx1 = wrapx(bx0 - 1 + 32 + (lane & 1), n1);
  int first = -1;
  bool uniform = true;
  for (int r = wid; r < SY * SZ; r += 8) {
    const int sy = r % SY, sz = r / SY;
    const int gy = wrapx(by0 + sy - 1, n2);
    int pl = lz0 + sz;  // buffer plane (halo offset included)
    if (pl > S.nzl + 1) pl = S.nzl + 1;
    const int* row = lbuf + (size_t)s3 * pl + (size_t)n1 * gy;
    const int v = row[gx0];
    sl[r * SX + lane] = v;
    if (first < 0) first = v & LMASK;
    uniform = uniform && ((v & LMASK) == first);
    if (lane < 2) {
      const int v2 = row[gx1];
      sl[r * SX + 32 + lane] = v2;
      uniform = uniform && ((v2 & LMASK) == first);
    }
  }
  __syncthreads();
  uniform = uniform && (first == (sl[0] & LMASK));
  if (__syncthreads_and(uniform)) return;
  const int lx = tid % TX, ly = tid / TX;
  const int gx = bx0 + lx, gy = by0 + ly;
  if (gx >= n1 || gy >= n2) return;
  for (int lz = 0; lz < TZ; lz++) {
    if (lz0 + lz >= S.nzl) break;
    const int c = sl[((lz + 1) * SY + (ly + 1)) * SX + lx + 1];
    if (!((unsigned)c & FILLBIT)) continue;
    const int cl = c & LMASK;
    bool edge = false;
#pragma unroll
    for (int dz = 0; dz < 3; dz++)
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++)
          edge = edge || ((sl[((lz + dz) * SY + (ly + dy)) * SX + lx + dx] & LMASK) != cl);
    if (edge) {
      lbuf[gx + n1 * gy + s3 * (lz0 + lz + 1)] = cl;  // clear FILLBIT: walked from now on
      list[atomicAdd(nlist, 1)] = gx + n1 * (gy + n2 * (S.zlo + lz0 + lz));
      // the next pass only has to look at tiles within one cell of a re-walked point
      const int x0 = (lx == 0) ? -1 : 0, x1 = (lx == TX - 1 || gx == n1 - 1) ? 1 : 0;
      const int y0 = (ly == 0) ? -1 : 0, y1 = (ly == TY - 1 || gy == n2 - 1) ? 1 : 0;
      const int z0 = (lz == 0) ? -1 : 0, z1 = (lz == TZ - 1 || lz0 + lz == S.nzl - 1) ? 1 : 0;
      for (int dz = z0; dz <= z1; dz++) {
        const int tz = (int)blockIdx.z + dz;
        if (tz < 0 || tz >= (int)gridDim.z) continue;  // other rank / periodic image: those layers are always rechecked
        for (int dy = y0; dy <= y1; dy++) {
          const int ty = ((int)blockIdx.y + dy + gridDim.y) % gridDim.y;
          for (int dx = x0; dx <= x1; dx++) {
            const int tx = ((int)blockIdx.x + dx + gridDim.x) % gridDim.x;
            dirty_out[tx + gridDim.x * (ty + gridDim.y * tz)] = 1;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compaction of the owned labels: terminal linear id -> index in the ordered maxima list
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact(long long nn, int* __restrict__ label, MaxHash h,
                                                 const int* __restrict__ cand2out, int* __restrict__ err) {
  // 4 consecutive labels per thread (16-byte accesses); the hash lookup is repeated only when the
  // terminal changes, which inside a basin it does not.
  const long long nv = nn >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int last_t = -1, last_o = -1;
  auto conv = [&](int v) -> int {
    const int t = v & LMASK;
    if (t != last_t) {
      const int ci = hash_lookup(h, t);
      if (ci < 0) { atomicExch(err, 1); return v; }
      last_t = t;
      last_o = __ldg(cand2out + ci);
    }
    return last_o;
  };
  int4* l4 = reinterpret_cast<int4*>(label);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    int4 v = l4[i];
    v.x = conv(v.x); v.y = conv(v.y); v.z = conv(v.z); v.w = conv(v.w);
    l4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (nn & 3)) {
    const long long i = (nv << 2) + threadIdx.x;
    label[i] = conv(label[i]);
  }
}

// reference scan-order key of the first point of each basin (C2G_ORDER_SCAN); label = owned planes
__global__ void __launch_bounds__(256) k_firstpoint(int n1, int n2, int n3, const Slab S, const int* __restrict__ label,
                                                    int* __restrict__ first) {
  const long long nn = (long long)n1 * n2 * S.nzl;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int x = (int)(i % n1), y = (int)((i / n1) % n2), z = S.zlo + (int)(i / ((long long)n1 * n2));
    const int key = (x * n2 + y) * n3 + z;
    const int l = label[i];
    if (key < first[l]) atomicMin(first + l, key);
  }
}

__global__ void k_permute_labels(long long nn, int* __restrict__ label, const int* __restrict__ perm) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) label[i] = perm[label[i]];
}

#define C2G_NCCL(ctx, call)                                                                         \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess)                                                                         \
      return (ctx)->fail(C2G_ERR_NCCL, "%s:%d: %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
  } while (0)

// halo planes of the label buffer.  which: 1 = halo above only (needed by classify), 3 = both.
int exchange_halos(c2g_context* ctx, int* lbuf, size_t plane, const Slab& S, int which) {
  cudaStream_t st = ctx->stream;
  const int G = ctx->nranks, r = ctx->rank;
  if (G == 1) {
    if (S.nzl < 1) return C2G_OK;
    // periodic images of the rank's own planes
    C2G_CUDA(ctx, cudaMemcpyAsync(lbuf + plane * (S.nzl + 1), lbuf + plane * 1, plane * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (which & 2)
      C2G_CUDA(ctx, cudaMemcpyAsync(lbuf, lbuf + plane * S.nzl, plane * sizeof(int), cudaMemcpyDeviceToDevice, st));
    ctx->launches += (which & 2) ? 2 : 1;
    return C2G_OK;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  const int up = (r + 1) % G, dn = (r + G - 1) % G;
  ctx->prof_begin("bader_halo_nccl");
  C2G_NCCL(ctx, ncclGroupStart());
  // my first owned plane is the halo-above of the rank below; I receive my halo-above from the rank above
  C2G_NCCL(ctx, ncclSend(lbuf + plane * 1, plane, ncclInt32, dn, comm, st));
  C2G_NCCL(ctx, ncclRecv(lbuf + plane * (S.nzl + 1), plane, ncclInt32, up, comm, st));
  if (which & 2) {
    C2G_NCCL(ctx, ncclSend(lbuf + plane * S.nzl, plane, ncclInt32, up, comm, st));
    C2G_NCCL(ctx, ncclRecv(lbuf, plane, ncclInt32, dn, comm, st));
  }
  C2G_NCCL(ctx, ncclGroupEnd());
  ctx->prof_end();
  return C2G_OK;
}

}  // namespace

// slab boundaries: multiples of 4, as even as possible
void c2g_slab_bounds(int n3, int nranks, int rank, int* zlo, int* zhi) {
  auto bound = [&](int r) -> int {
    if (r >= nranks) return n3;
    long long z = (long long)n3 * r / nranks;
    z = (z + 2) / 4 * 4;
    if (z > n3) z = n3;
    return (int)z;
  };
  *zlo = bound(rank);
  *zhi = bound(rank + 1);
  if (*zhi < *zlo) *zhi = *zlo;
}

// =================================================================================================
extern "C" int c2g_bader_assign(c2g_context* ctx, int handle, const double car2lat[9], const double lat_i_dist[27],
                                int algo, int order, int* nmax_out, c2g_basins** res_out) {
  if (!ctx) return C2G_ERR_ARG;
  if (!nmax_out || !res_out || !car2lat || !lat_i_dist) return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: null argument");
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: invalid grid handle %d", handle);
  const c2g_grid& g = ctx->grids[handle];
  if (g.nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: grid too large for int32 indices");
  cudaStream_t st = ctx->stream;
  const int G = ctx->nranks;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  BaderParams P;
  P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
  memcpy(P.c2l, car2lat, sizeof(P.c2l));
  memcpy(P.lid, lat_i_dist, sizeof(P.lid));
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const size_t plane = (size_t)n1 * n2;
  // diagonal car2lat (orthogonal cell): the off-diagonal products are exact zeros and can be skipped
  const bool ortho = P.c2l[1] == 0.0 && P.c2l[2] == 0.0 && P.c2l[3] == 0.0 && P.c2l[5] == 0.0 && P.c2l[6] == 0.0 && P.c2l[7] == 0.0;
  Slab S;
  c2g_slab_bounds(n3, G, ctx->rank, &S.zlo, &S.zhi);
  S.nzl = S.zhi - S.zlo;
  const long long nnl = (long long)plane * S.nzl;  // owned points

  c2g_basins* res = new c2g_basins();
  res->ctx = ctx; res->kind = 0; res->gridh = handle;
  res->n[0] = n1; res->n[1] = n2; res->n[2] = n3; res->nn = g.nn;
  res->zlo = S.zlo; res->zhi = S.zhi;
  struct Guard { c2g_basins* r; bool ok = false; ~Guard() { if (!ok) c2g_basins_free(r); } } guard{res};

  // the owned planes start one plane into the buffer: pad so that they are 16-byte aligned (int4 passes)
  const size_t pad = (4 - plane % 4) % 4;
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&res->d_lbuf, sizeof(int) * (plane * (S.nzl + 2) + pad)));
  int* lbuf = res->d_lbuf + pad;
  res->d_label = lbuf + plane;                               // owned planes
  int* label_g = lbuf + plane - (long long)plane * S.zlo;    // label_g[global id] for owned points

  // ---- K0: candidate maxima of the slab ----
  const int c41 = (n1 + 3) / 4, c42 = (n2 + 3) / 4, c43 = (S.nzl + 3) / 4;
  const int c21 = (n1 + 1) / 2, c22 = (n2 + 1) / 2, c23 = (S.nzl + 1) / 2;
  const size_t ncube4 = std::max<size_t>(1, (size_t)c41 * c42 * c43), ncube2 = std::max<size_t>(1, (size_t)c21 * c22 * c23);
  DevBuf b_cube4, b_cube2, b_cand, b_cnt;
  C2G_CUDA(ctx, b_cube4.alloc(ctx, ncube4));
  C2G_CUDA(ctx, b_cube2.alloc(ctx, ncube2));
  int maxcand = (int)std::max<long long>(1, std::min<long long>(nnl, std::max<long long>(1 << 16, nnl / 64)));
  C2G_CUDA(ctx, b_cnt.alloc(ctx, 64));
  // counters: [0] ncand, [1] nlist, [2] noverflow, [3] err ; [4..5] nsteps (u64) ; [6] scratch for collectives
  int* cnt = b_cnt.as<int>();
  unsigned long long* nsteps = (unsigned long long*)(cnt + 4);
  int hcnt[8];
  for (int attempt = 0;; attempt++) {
    C2G_CUDA(ctx, b_cand.alloc(ctx, sizeof(int) * (size_t)maxcand));
    C2G_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_cube4.p, 0, ncube4, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_cube2.p, 0, ncube2, st));
    if (S.nzl > 0) {
      dim3 grid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, (S.nzl + TZ - 1) / TZ);
      const size_t smem = sizeof(double) * ((TZ + 2) * (TY + 2) * (TX + 2) + (TZ + 2) * (TY + 2) * TX);
      C2G_CUDA(ctx, cudaFuncSetAttribute(k_maxima, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ctx->prof_begin("bader_maxima");
      k_maxima<<<grid, 256, smem, st>>>(P, S, g.d, b_cand.as<int>(), cnt, maxcand, b_cube4.as<unsigned char>(),
                                        b_cube2.as<unsigned char>());
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hcnt[0] <= maxcand) break;
    if (attempt > 0) return ctx->fail(C2G_ERR_OVERFLOW, "candidate maxima list overflow");
    b_cand.reset();
    maxcand = hcnt[0];
  }
  int ncand_local = hcnt[0];
  std::vector<int> cand;
  if (G == 1) {
    cand.resize(ncand_local);
    if (ncand_local) C2G_CUDA(ctx, cudaMemcpy(cand.data(), b_cand.p, sizeof(int) * ncand_local, cudaMemcpyDeviceToHost));
  } else {
    // all ranks need every candidate (a trajectory may end in any slab): padded all-gather
    DevBuf b_mx, b_all;
    C2G_CUDA(ctx, b_mx.alloc(ctx, sizeof(int)));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_mx.p, &ncand_local, sizeof(int), cudaMemcpyHostToDevice, st));
    C2G_NCCL(ctx, ncclAllReduce(b_mx.p, b_mx.p, 1, ncclInt32, ncclMax, comm, st));
    int mx = 0;
    C2G_CUDA(ctx, cudaMemcpyAsync(&mx, b_mx.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    mx = std::max(mx, 1);
    DevBuf b_pad;
    C2G_CUDA(ctx, b_pad.alloc(ctx, sizeof(int) * (size_t)mx));
    C2G_CUDA(ctx, cudaMemsetAsync(b_pad.p, 0xff, sizeof(int) * (size_t)mx, st));
    if (ncand_local)
      C2G_CUDA(ctx, cudaMemcpyAsync(b_pad.p, b_cand.p, sizeof(int) * ncand_local, cudaMemcpyDeviceToDevice, st));
    C2G_CUDA(ctx, b_all.alloc(ctx, sizeof(int) * (size_t)mx * G));
    ctx->prof_begin("bader_cand_allgather_nccl");
    C2G_NCCL(ctx, ncclAllGather(b_pad.p, b_all.p, mx, ncclInt32, comm, st));
    ctx->prof_end();
    std::vector<int> all((size_t)mx * G);
    C2G_CUDA(ctx, cudaMemcpyAsync(all.data(), b_all.p, sizeof(int) * all.size(), cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    for (int v : all)
      if (v >= 0) cand.push_back(v);
  }
  const int ncand = (int)cand.size();
  if (ncand == 0) return ctx->fail(C2G_ERR_STATE, "c2g_bader_assign: the field has no local maximum (NaN input?)");
  std::sort(cand.begin(), cand.end());
  // hash table on the host, then upload
  unsigned hsize = 1024;
  while (hsize < 4u * (unsigned)ncand) hsize <<= 1;
  std::vector<int> hk(hsize, -1), hv(hsize, -1);
  for (int i = 0; i < ncand; i++) {
    unsigned s = ((unsigned)cand[i] * 2654435761u) & (hsize - 1);
    while (hk[s] >= 0) s = (s + 1) & (hsize - 1);
    hk[s] = cand[i]; hv[s] = i;
  }
  DevBuf b_hk, b_hv, b_reached;
  C2G_CUDA(ctx, b_hk.alloc(ctx, sizeof(int) * hsize));
  C2G_CUDA(ctx, b_hv.alloc(ctx, sizeof(int) * hsize));
  C2G_CUDA(ctx, b_reached.alloc(ctx, ncand));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hk.p, hk.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hv.p, hv.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemsetAsync(b_reached.p, 0, ncand, st));
  MaxHash h{b_hk.as<int>(), b_hv.as<int>(), hsize - 1};
  unsigned char* reached = b_reached.as<unsigned char>();

  // work list / overflow list
  DevBuf b_list, b_over;
  const long long listcap = (algo == C2G_BADER_EXACT) ? 1 : std::max<long long>(1, nnl);
  C2G_CUDA(ctx, b_list.alloc(ctx, sizeof(int) * (size_t)listcap));
  long long overcap = std::max<long long>(1024, nnl / 16);
  C2G_CUDA(ctx, b_over.alloc(ctx, sizeof(int) * (size_t)overcap));
  int* list = b_list.as<int>();
  int* over = b_over.as<int>();
  long long walked = 0, fixpts = 0, fixpasses = 0, noverflow_total = 0;

  // handle walkers whose path buffer overflowed
  auto drain_overflow = [&]() -> int {
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hcnt[3] == 1) return ctx->fail(C2G_ERR_NEWMAX, "bader walk ended on a point that is not a local maximum");
    const int nov = hcnt[2];
    if (nov == 0) return C2G_OK;
    if (nov > overcap) return ctx->fail(C2G_ERR_OVERFLOW, "too many long trajectories (%d)", nov);
    noverflow_total += nov;
    const int bigcap = (int)std::min<long long>(g.nn, 1 << 22);
    const int chunk = (int)std::max<long long>(1, std::min<long long>(nov, (1ll << 31) / bigcap));  // <= 8 GiB scratch
    DevBuf b_scr;
    C2G_CUDA(ctx, b_scr.alloc(ctx, sizeof(int) * (size_t)chunk * bigcap));
    for (int off = 0; off < nov; off += chunk) {
      const int c = std::min(chunk, nov - off);
      ctx->prof_begin("bader_walk_big");
      k_walk_big<<<c2g_blocks_for(c, 64), 64, 0, st>>>(P, g.d, over + off, c, label_g, h, reached, b_scr.as<int>(), bigcap,
                                                       cnt + 3, nsteps);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    C2G_CUDA(ctx, cudaMemsetAsync(cnt + 2, 0, sizeof(int), st));
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hcnt[3] == 2) return ctx->fail(C2G_ERR_OVERFLOW, "trajectory longer than %d steps", bigcap);
    if (hcnt[3] == 1) return ctx->fail(C2G_ERR_NEWMAX, "bader walk ended on a point that is not a local maximum");
    return C2G_OK;
  };
  auto grow_over = [&](long long need) -> int {
    if (need > overcap) {
      b_over.reset();
      overcap = need;
      C2G_CUDA(ctx, b_over.alloc(ctx, sizeof(int) * (size_t)overcap));
      over = b_over.as<int>();
    }
    return C2G_OK;
  };
  SafeMap cursafe{nullptr, 0, 0, 0, 0, 0};
  auto walk_list = [&](int count, const char* name) -> int {
    if (count <= 0) return C2G_OK;
    int rc = grow_over(count);
    if (rc) return rc;
    ctx->prof_begin(name);
    if (ortho)
      k_walk_list<true><<<c2g_blocks_for(count, 128), 128, 0, st>>>(P, g.d, list, count, label_g, h, cursafe, reached, over, cnt + 2,
                                                                    cnt + 3, algo == C2G_BADER_EXACT ? nsteps : nullptr);
    else
      k_walk_list<false><<<c2g_blocks_for(count, 128), 128, 0, st>>>(P, g.d, list, count, label_g, h, cursafe, reached, over, cnt + 2,
                                                                     cnt + 3, algo == C2G_BADER_EXACT ? nsteps : nullptr);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    walked += count;
    return drain_overflow();
  };
  auto walk_lattice = [&](int s, const char* name) -> int {
    const int m1 = (n1 + s - 1) / s, m2 = (n2 + s - 1) / s, m3 = (S.nzl + s - 1) / s;
    const long long m = (long long)m1 * m2 * m3;
    if (m == 0) return C2G_OK;
    int rc = grow_over(m);
    if (rc) return rc;
    ctx->prof_begin(name);
    if (ortho)
      k_walk_lattice<true><<<c2g_blocks_for(m, 128), 128, 0, st>>>(P, S, g.d, s, m1, m2, m3, label_g, h, reached, over, cnt + 2,
                                                                   cnt + 3, algo == C2G_BADER_EXACT ? nsteps : nullptr);
    else
      k_walk_lattice<false><<<c2g_blocks_for(m, 128), 128, 0, st>>>(P, S, g.d, s, m1, m2, m3, label_g, h, reached, over, cnt + 2,
                                                                    cnt + 3, algo == C2G_BADER_EXACT ? nsteps : nullptr);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    walked += m;
    return drain_overflow();
  };

  int rc;
  if (algo == C2G_BADER_EXACT) {
    if ((rc = walk_lattice(1, "bader_walk_all")) != C2G_OK) return rc;
  } else {
    // level 0: stride-4 lattice
    if ((rc = walk_lattice(4, "bader_walk_l4")) != C2G_OK) return rc;
    // levels 4 -> 2 -> 1
    DevBuf b_uni(ctx), b_safe(ctx);
    const bool use_safe = getenv("C2G_NO_EARLY_STOP") == nullptr;
    for (int s = 4; s >= 2; s >>= 1) {
      if ((rc = exchange_halos(ctx, lbuf, plane, S, 1)) != C2G_OK) return rc;
      const int c1 = (n1 + s - 1) / s, c2 = (n2 + s - 1) / s, c3 = (S.nzl + s - 1) / s;
      const long long nc = (long long)c1 * c2 * c3;
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));
      cursafe.safe = nullptr;
      if (nc > 0) {
        C2G_CUDA(ctx, b_uni.alloc(ctx, sizeof(int) * (size_t)nc));
        C2G_CUDA(ctx, b_safe.alloc(ctx, sizeof(int) * (size_t)nc));
        ctx->prof_begin(s == 4 ? "bader_classify4" : "bader_classify2");
        k_classify<<<c2g_blocks_for(nc, 256), 256, 0, st>>>(n1, n2, n3, S, s, lbuf,
                                                            s == 4 ? b_cube4.as<unsigned char>() : b_cube2.as<unsigned char>(),
                                                            b_uni.as<int>(), list, cnt + 1);
        ctx->prof_end();
        C2G_KERNEL_CHECK(ctx);
        if (use_safe) {
          ctx->prof_begin(s == 4 ? "bader_safe4" : "bader_safe2");
          k_safe<<<c2g_blocks_for(nc, 256), 256, 0, st>>>(c1, c2, c3, b_uni.as<int>(), b_safe.as<int>());
          ctx->prof_end();
          C2G_KERNEL_CHECK(ctx);
          cursafe = SafeMap{b_safe.as<int>(), s == 4 ? 2 : 1, c1, c2, S.zlo, S.nzl};
        }
      }
      C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
      C2G_CUDA(ctx, cudaStreamSynchronize(st));
      if ((rc = walk_list(hcnt[1], s == 4 ? "bader_walk_l2" : "bader_walk_l1")) != C2G_OK) return rc;
    }
    // edge fix until no filled point (on any rank) is adjacent to a different label.  After the first
    // pass only the tiles within one cell of a re-walked point (plus the slab's face layers, whose
    // halos may have changed on another rank) are looked at again.
    DevBuf b_dirty[2] = {DevBuf(ctx), DevBuf(ctx)};
    const dim3 egrid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, std::max(1, (S.nzl + TZ - 1) / TZ));
    const size_t ntiles = (size_t)egrid.x * egrid.y * egrid.z;
    C2G_CUDA(ctx, b_dirty[0].alloc(ctx, ntiles));
    C2G_CUDA(ctx, b_dirty[1].alloc(ctx, ntiles));
    for (;;) {
      if ((rc = exchange_halos(ctx, lbuf, plane, S, 3)) != C2G_OK) return rc;
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));
      unsigned char* din = fixpasses == 0 ? nullptr : b_dirty[fixpasses & 1].as<unsigned char>();
      unsigned char* dout = b_dirty[(fixpasses + 1) & 1].as<unsigned char>();
      C2G_CUDA(ctx, cudaMemsetAsync(dout, 0, ntiles, st));
      // face layers are always rechecked
      C2G_CUDA(ctx, cudaMemsetAsync(dout, 1, (size_t)egrid.x * egrid.y, st));
      C2G_CUDA(ctx, cudaMemsetAsync(dout + (size_t)egrid.x * egrid.y * (egrid.z - 1), 1, (size_t)egrid.x * egrid.y, st));
      if (S.nzl > 0) {
        ctx->prof_begin("bader_edgefix");
        k_edgefix<<<egrid, 256, 0, st>>>(n1, n2, S, lbuf, list, cnt + 1, din, dout);
        ctx->prof_end();
        C2G_KERNEL_CHECK(ctx);
      }
      if (G > 1) {
        C2G_CUDA(ctx, cudaMemcpyAsync(cnt + 6, cnt + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
        C2G_NCCL(ctx, ncclAllReduce(cnt + 6, cnt + 6, 1, ncclInt32, ncclMax, comm, st));
      }
      C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
      C2G_CUDA(ctx, cudaStreamSynchronize(st));
      fixpasses++;
      const int any = (G > 1) ? hcnt[6] : hcnt[1];
      if (any == 0) break;
      fixpts += hcnt[1];
      if ((rc = walk_list(hcnt[1], "bader_walk_fix")) != C2G_OK) return rc;
      if (fixpasses > 1000) return ctx->fail(C2G_ERR_STATE, "edge refinement did not converge");
    }
  }

  // ---- maxima actually reached (on any rank), output order, compaction ----
  if (G > 1) C2G_NCCL(ctx, ncclAllReduce(reached, reached, ncand, ncclUint8, ncclMax, comm, st));
  std::vector<unsigned char> hreached(ncand);
  C2G_CUDA(ctx, cudaMemcpyAsync(hreached.data(), reached, ncand, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  std::vector<int> cand2out(ncand, -1);
  int nmax = 0;
  for (int i = 0; i < ncand; i++)
    if (hreached[i]) cand2out[i] = nmax++;
  res->nmax = nmax;
  res->max_lin.resize(nmax);
  for (int i = 0; i < ncand; i++)
    if (cand2out[i] >= 0) res->max_lin[cand2out[i]] = cand[i];
  DevBuf b_c2o;
  C2G_CUDA(ctx, b_c2o.alloc(ctx, sizeof(int) * ncand));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_c2o.p, cand2out.data(), sizeof(int) * ncand, cudaMemcpyHostToDevice, st));
  const int nblk = ctx->nsm * 8;
  if (nnl > 0) {
    ctx->prof_begin("bader_compact");
    k_compact<<<nblk, 256, 0, st>>>(nnl, res->d_label, h, b_c2o.as<int>(), cnt + 3);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
  }
  C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  if (hcnt[3] != 0) return ctx->fail(C2G_ERR_NEWMAX, "label compaction found a terminal that is not a known maximum");

  if (order == C2G_ORDER_SCAN && nmax > 1) {
    DevBuf b_first, b_perm;
    C2G_CUDA(ctx, b_first.alloc(ctx, sizeof(int) * nmax));
    C2G_CUDA(ctx, cudaMemsetAsync(b_first.p, 0x7f, sizeof(int) * nmax, st));
    if (nnl > 0) {
      ctx->prof_begin("bader_firstpoint");
      k_firstpoint<<<nblk, 256, 0, st>>>(n1, n2, n3, S, res->d_label, b_first.as<int>());
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    if (G > 1) C2G_NCCL(ctx, ncclAllReduce(b_first.p, b_first.p, nmax, ncclInt32, ncclMin, comm, st));
    std::vector<int> first(nmax);
    C2G_CUDA(ctx, cudaMemcpyAsync(first.data(), b_first.p, sizeof(int) * nmax, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int> idx(nmax), perm(nmax);
    for (int i = 0; i < nmax; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return first[a] < first[b] || (first[a] == first[b] && a < b); });
    std::vector<int> ml(nmax);
    for (int k = 0; k < nmax; k++) { perm[idx[k]] = k; ml[k] = res->max_lin[idx[k]]; }
    res->max_lin.swap(ml);
    C2G_CUDA(ctx, b_perm.alloc(ctx, sizeof(int) * nmax));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_perm.p, perm.data(), sizeof(int) * nmax, cudaMemcpyHostToDevice, st));
    if (nnl > 0) {
      ctx->prof_begin("bader_permute");
      k_permute_labels<<<nblk, 256, 0, st>>>(nnl, res->d_label, b_perm.as<int>());
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  unsigned long long hsteps = 0;
  C2G_CUDA(ctx, cudaMemcpy(&hsteps, nsteps, sizeof(hsteps), cudaMemcpyDeviceToHost));
  res->stats[0] = walked;
  res->stats[1] = fixpasses;
  res->stats[2] = fixpts;
  res->stats[3] = noverflow_total;
  res->stats[4] = ncand;
  res->stats[5] = (long long)hsteps;
  ctx->prof_collect();
  guard.ok = true;
  *nmax_out = nmax;
  *res_out = res;
  return C2G_OK;
}
