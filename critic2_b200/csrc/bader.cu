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

__device__ __forceinline__ int wrapc(int p, int n) {
  while (p < 0) p += n;
  while (p >= n) p -= n;
  return p;
}

// is_max, bader@proc.f90:571-597 (no neighbour strictly greater)
__device__ bool dev_is_max(const BaderParams& P, const double* __restrict__ rho, int x, int y, int z, double r0) {
  bool ismax = true;
#pragma unroll 1
  for (int d3 = -1; d3 <= 1; d3++) {
    const int zz = wrapc(z + d3, P.n3);
    for (int d2 = -1; d2 <= 1; d2++) {
      const int yy = wrapc(y + d2, P.n2);
      const int base = P.n1 * (yy + P.n2 * zz);
      for (int d1 = -1; d1 <= 1; d1++) {
        const int xx = wrapc(x + d1, P.n1);
        if (__ldg(rho + base + xx) > r0) ismax = false;
      }
    }
  }
  return ismax;
}

// step_ongrid, bader@proc.f90:500-527.  Loop order d1 (outer), d2, d3 (inner), first strictly greater wins.
__device__ void dev_step_ongrid(const BaderParams& P, const double* __restrict__ rho, int x, int y, int z,
                                double rho_ctr, int& ox, int& oy, int& oz) {
  double rho_max = rho_ctr;
  int bx = x, by = y, bz = z;
#pragma unroll 1
  for (int d1 = -1; d1 <= 1; d1++) {
    const int xx = wrapc(x + d1, P.n1);
    for (int d2 = -1; d2 <= 1; d2++) {
      const int yy = wrapc(y + d2, P.n2);
      for (int d3 = -1; d3 <= 1; d3++) {
        const int zz = wrapc(z + d3, P.n3);
        double rho_tmp = __ldg(rho + xx + P.n1 * (yy + P.n2 * zz));
        rho_tmp = rho_ctr + (rho_tmp - rho_ctr) * P.lid[(d1 + 1) * 9 + (d2 + 1) * 3 + (d3 + 1)];
        if (rho_tmp > rho_max) {
          rho_max = rho_tmp;
          bx = xx; by = yy; bz = zz;
        }
      }
    }
  }
  ox = bx; oy = by; oz = bz;
}

// One complete near-grid trajectory (max_neargrid, bader@proc.f90:427-450 on a fresh grid).
// Returns the linear id of the terminal maximum, or -1 if the path buffer overflowed.
// The reference's "known(pm)==1" revisit test (:484-488) is answered exactly: a visited point
// can only be hit again when rho(pm) <= max rho along the path, and only then the stored path is
// searched.
__device__ int dev_walk(const BaderParams& P, const double* __restrict__ rho, int start, int* path, int cap,
                        unsigned long long* nsteps_out) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
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
    const int xp = (x + 1 == n1) ? 0 : x + 1, xm = (x == 0) ? n1 - 1 : x - 1;
    const int yp = (y + 1 == n2) ? 0 : y + 1, ym = (y == 0) ? n2 - 1 : y - 1;
    const int zp = (z + 1 == n3) ? 0 : z + 1, zm = (z == 0) ? n3 - 1 : z - 1;
    const int row = n1 * (y + n2 * z);
    const double rxp = __ldg(rho + row + xp), rxm = __ldg(rho + row + xm);
    const double ryp = __ldg(rho + x + n1 * (yp + n2 * z)), rym = __ldg(rho + x + n1 * (ym + n2 * z));
    const double rzp = __ldg(rho + x + n1 * (y + n2 * zp)), rzm = __ldg(rho + x + n1 * (y + n2 * zm));
    // rho_grad_dir (:532-567)
    double gl0 = (rxp - rxm) * 0.5, gl1 = (ryp - rym) * 0.5, gl2 = (rzp - rzm) * 0.5;
    if (rxp < r0 && rxm < r0) gl0 = 0.0;
    if (ryp < r0 && rym < r0) gl1 = 0.0;
    if (rzp < r0 && rzm < r0) gl2 = 0.0;
    const double gc0 = gl0 * P.c2l[0] + gl1 * P.c2l[1] + gl2 * P.c2l[2];
    const double gc1 = gl0 * P.c2l[3] + gl1 * P.c2l[4] + gl2 * P.c2l[5];
    const double gc2 = gl0 * P.c2l[6] + gl1 * P.c2l[7] + gl2 * P.c2l[8];
    double g0 = P.c2l[0] * gc0 + P.c2l[3] * gc1 + P.c2l[6] * gc2;
    double g1 = P.c2l[1] * gc0 + P.c2l[4] * gc1 + P.c2l[7] * gc2;
    double g2 = P.c2l[2] * gc0 + P.c2l[5] * gc1 + P.c2l[8] * gc2;
    const double gmax = fmax(fabs(g0), fmax(fabs(g1), fabs(g2)));
    int nx, ny, nz;
    if (gmax < 1e-30) {  // (:468-476)
      dr0 = dr1 = dr2 = 0.0;
      if (dev_is_max(P, rho, x, y, z, r0)) break;
      dev_step_ongrid(P, rho, x, y, z, r0, nx, ny, nz);
    } else {  // (:477-483)
      const double coeff = 1.0 / gmax;
      g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
      const double a0 = round(g0), a1 = round(g1), a2 = round(g2);
      dr0 = dr0 + g0 - a0; dr1 = dr1 + g1 - a1; dr2 = dr2 + g2 - a2;
      const double b0 = round(dr0), b1 = round(dr1), b2 = round(dr2);
      dr0 = dr0 - b0; dr1 = dr1 - b1; dr2 = dr2 - b2;
      nx = wrapc(x + (int)a0 + (int)b0, n1);
      ny = wrapc(y + (int)a1 + (int)b1, n2);
      nz = wrapc(z + (int)a2 + (int)b2, n3);
    }
    // known(p) = 1 (:484)
    if (len >= cap) return -1;
    path[len++] = id;
    rhomax = fmax(rhomax, r0);
    int nid = nx + n1 * (ny + n2 * nz);
    double rn = __ldg(rho + nid);
    if (rn <= rhomax) {  // only then pm can be a point of this path (:487)
      bool found = false;
      for (int j = len - 1; j >= 0; j--)
        if (path[j] == nid) { found = true; break; }
      if (found) {
        dev_step_ongrid(P, rho, x, y, z, r0, nx, ny, nz);
        dr0 = dr1 = dr2 = 0.0;
        nid = nx + n1 * (ny + n2 * nz);
        rn = __ldg(rho + nid);
      }
    }
    if (nid == id) break;  // did not move: maximum (:439)
    x = nx; y = ny; z = nz; id = nid; r0 = rn;
  }
  if (nsteps_out) atomicAdd(nsteps_out, (unsigned long long)len);
  return id;
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

// ------------------------------------------------------------------------------------------------
// K0: candidate maxima (26-neighbour, is_max) by a separable 3x3x3 box maximum on shared-memory
// tiles with a periodic 1-cell halo; marks the stride-4 / stride-2 cubes that contain a maximum.
// ------------------------------------------------------------------------------------------------
constexpr int TX = 32, TY = 8, TZ = 8;
__global__ void __launch_bounds__(256) k_maxima(const __grid_constant__ BaderParams P, const double* __restrict__ rho,
                                                int* __restrict__ cand, int* __restrict__ ncand, int maxcand,
                                                unsigned char* __restrict__ cube4, unsigned char* __restrict__ cube2) {
  extern __shared__ double sm[];
  double* s0 = sm;                                   // [TZ+2][TY+2][TX+2]
  double* s1 = sm + (TZ + 2) * (TY + 2) * (TX + 2);  // [TZ+2][TY+2][TX] max over x
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int bx0 = blockIdx.x * TX, by0 = blockIdx.y * TY, bz0 = blockIdx.z * TZ;
  const int tid = threadIdx.x;
  constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2;
  for (int e = tid; e < SX * SY * SZ; e += 256) {
    const int sx = e % SX, sy = (e / SX) % SY, sz = e / (SX * SY);
    const int gx = wrapc(bx0 + sx - 1, n1), gy = wrapc(by0 + sy - 1, n2), gz = wrapc(bz0 + sz - 1, n3);
    s0[e] = __ldg(rho + gx + n1 * (gy + n2 * gz));
  }
  __syncthreads();
  for (int e = tid; e < TX * SY * SZ; e += 256) {
    const int sx = e % TX, r = e / TX;  // r = sy + SY*sz
    const double* p = s0 + r * SX + sx;
    s1[e] = fmax(p[0], fmax(p[1], p[2]));
  }
  __syncthreads();
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
      if (gz >= n3) break;
      const double c = s0[((lz + 1) * SY + (ly + 1)) * SX + lx + 1];
      const double bm = fmax(m[lz], fmax(m[lz + 1], m[lz + 2]));
      if (!(bm > c)) {  // no neighbour strictly greater
        const int slot = atomicAdd(ncand, 1);
        if (slot < maxcand) cand[slot] = gx + n1 * (gy + n2 * gz);
        const int c41 = (n1 + 3) / 4, c42 = (n2 + 3) / 4;
        const int c21 = (n1 + 1) / 2, c22 = (n2 + 1) / 2;
        cube4[(gx >> 2) + c41 * ((gy >> 2) + (size_t)c42 * (gz >> 2))] = 1;
        cube2[(gx >> 1) + c21 * ((gy >> 1) + (size_t)c22 * (gz >> 1))] = 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// walkers
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

// every point of the stride-s lattice (s = 1: every grid point = the EXACT referee)
__global__ void __launch_bounds__(128) k_walk_lattice(const __grid_constant__ BaderParams P, const double* __restrict__ rho,
                                                      int s, int m1, int m2, int m3, int* __restrict__ label,
                                                      MaxHash h, unsigned char* __restrict__ reached,
                                                      int* __restrict__ overflow, int* __restrict__ noverflow,
                                                      int* __restrict__ err, unsigned long long* nsteps) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)m1 * m2 * m3) return;
  const int lx = (int)(t % m1), ly = (int)((t / m1) % m2), lz = (int)(t / ((long long)m1 * m2));
  const int start = lx * s + P.n1 * (ly * s + P.n2 * (lz * s));
  int path[PATHCAP];
  const int term = dev_walk(P, rho, start, path, PATHCAP, nsteps);
  finish_walk(start, term, label, h, reached, overflow, noverflow, err);
}

__global__ void __launch_bounds__(128) k_walk_list(const __grid_constant__ BaderParams P, const double* __restrict__ rho,
                                                   const int* __restrict__ list, int count, int* __restrict__ label,
                                                   MaxHash h, unsigned char* __restrict__ reached,
                                                   int* __restrict__ overflow, int* __restrict__ noverflow,
                                                   int* __restrict__ err, unsigned long long* nsteps) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int start = list[t];
  int path[PATHCAP];
  const int term = dev_walk(P, rho, start, path, PATHCAP, nsteps);
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
  const int term = dev_walk(P, rho, start, scratch + (size_t)t * bigcap, bigcap, nsteps);
  if (term < 0) { atomicExch(err, 2); return; }
  label[start] = term;
  const int ci = hash_lookup(h, term);
  if (ci < 0) atomicExch(err, 1);
  else reached[ci] = 1;
}

// ------------------------------------------------------------------------------------------------
// classify: one thread per stride-s cube.  If its 8 corners (already labelled) agree and it holds
// no local maximum, fill its new stride-s/2 points (label | FILLBIT); otherwise queue them.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_classify(int n1, int n2, int n3, int s, int* __restrict__ label,
                                                  const unsigned char* __restrict__ cubemax,
                                                  int* __restrict__ list, int* __restrict__ nlist) {
  const int c1 = (n1 + s - 1) / s, c2 = (n2 + s - 1) / s, c3 = (n3 + s - 1) / s;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < (long long)c1 * c2 * c3;
  int npush = 0;
  int pts[7];
  if (active) {
    const int cx = (int)(t % c1), cy = (int)((t / c1) % c2), cz = (int)(t / ((long long)c1 * c2));
    const int x0 = cx * s, y0 = cy * s, z0 = cz * s;
    const int x1 = (x0 + s < n1) ? x0 + s : 0, y1 = (y0 + s < n2) ? y0 + s : 0, z1 = (z0 + s < n3) ? z0 + s : 0;
    const int l000 = label[x0 + n1 * (y0 + n2 * z0)] & LMASK;
    bool uni = !cubemax[t];
    uni = uni && ((label[x1 + n1 * (y0 + n2 * z0)] & LMASK) == l000);
    uni = uni && ((label[x0 + n1 * (y1 + n2 * z0)] & LMASK) == l000);
    uni = uni && ((label[x1 + n1 * (y1 + n2 * z0)] & LMASK) == l000);
    uni = uni && ((label[x0 + n1 * (y0 + n2 * z1)] & LMASK) == l000);
    uni = uni && ((label[x1 + n1 * (y0 + n2 * z1)] & LMASK) == l000);
    uni = uni && ((label[x0 + n1 * (y1 + n2 * z1)] & LMASK) == l000);
    uni = uni && ((label[x1 + n1 * (y1 + n2 * z1)] & LMASK) == l000);
    const int h = s >> 1;
    for (int o = 1; o < 8; o++) {
      const int x = x0 + ((o & 1) ? h : 0), y = y0 + ((o & 2) ? h : 0), z = z0 + ((o & 4) ? h : 0);
      if (x >= n1 || y >= n2 || z >= n3) continue;
      const int id = x + n1 * (y + n2 * z);
      if (uni) label[id] = (int)((unsigned)l000 | FILLBIT);
      else pts[npush++] = id;
    }
  }
  // block-aggregated append (keeps spatial order inside a block)
  __shared__ int s_base, s_tot;
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
    s_tot = acc;
    s_base = acc ? atomicAdd(nlist, acc) : 0;
  }
  __syncthreads();
  if (npush) {
    int off = s_base + s_warp[wid] + incl - npush;
    for (int k = 0; k < npush; k++) list[off + k] = pts[k];
  }
}

// ------------------------------------------------------------------------------------------------
// edge fix: every FILLED point with a 26-neighbour of a different label is queued for an exact walk
// (the refine_edge criterion, is_vol_edge bader@proc.f90:730-752).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edgefix(int n1, int n2, int n3, int* __restrict__ label,
                                                 int* __restrict__ list, int* __restrict__ nlist) {
  __shared__ int sl[(TZ + 2) * (TY + 2) * (TX + 2)];
  constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2;
  const int bx0 = blockIdx.x * TX, by0 = blockIdx.y * TY, bz0 = blockIdx.z * TZ;
  const int tid = threadIdx.x;
  int first = 0;
  bool uniform = true;
  for (int e = tid; e < SX * SY * SZ; e += 256) {
    const int sx = e % SX, sy = (e / SX) % SY, sz = e / (SX * SY);
    const int gx = wrapc(bx0 + sx - 1, n1), gy = wrapc(by0 + sy - 1, n2), gz = wrapc(bz0 + sz - 1, n3);
    const int v = label[gx + n1 * (gy + n2 * gz)];
    sl[e] = v;
    if (e == tid) first = v & LMASK;
    else uniform = uniform && ((v & LMASK) == first);
  }
  __syncthreads();
  uniform = uniform && (first == (sl[0] & LMASK));
  if (__syncthreads_and(uniform)) return;
  const int lx = tid % TX, ly = tid / TX;
  const int gx = bx0 + lx, gy = by0 + ly;
  if (gx >= n1 || gy >= n2) return;
  for (int lz = 0; lz < TZ; lz++) {
    const int gz = bz0 + lz;
    if (gz >= n3) break;
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
      const int id = gx + n1 * (gy + n2 * gz);
      label[id] = cl;  // clear FILLBIT: walked from now on
      list[atomicAdd(nlist, 1)] = id;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// compaction of labels: terminal linear id -> index in the ordered maxima list; counts points
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact(long long nn, int* __restrict__ label, MaxHash h,
                                                 const int* __restrict__ cand2out, unsigned long long* __restrict__ counts,
                                                 int* __restrict__ err) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  int last_t = -1, last_o = -1;
  unsigned long long run = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int t = label[i] & LMASK;
    if (t != last_t) {
      if (run) atomicAdd(counts + last_o, run);
      run = 0;
      const int ci = hash_lookup(h, t);
      if (ci < 0) { atomicExch(err, 1); last_t = -1; last_o = -1; continue; }
      last_t = t;
      last_o = cand2out[ci];
    }
    run++;
    label[i] = last_o;
  }
  if (run) atomicAdd(counts + last_o, run);
}

// reference scan-order key of the first point of each basin (C2G_ORDER_SCAN)
__global__ void __launch_bounds__(256) k_firstpoint(int n1, int n2, int n3, const int* __restrict__ label,
                                                    int* __restrict__ first) {
  const long long nn = (long long)n1 * n2 * n3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int x = (int)(i % n1), y = (int)((i / n1) % n2), z = (int)(i / ((long long)n1 * n2));
    const int key = (x * n2 + y) * n3 + z;
    const int l = label[i];
    if (key < first[l]) atomicMin(first + l, key);
  }
}

__global__ void k_permute_labels(long long nn, int* __restrict__ label, const int* __restrict__ perm) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) label[i] = perm[label[i]];
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return (T*)p; }
};

}  // namespace

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
  BaderParams P;
  P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
  memcpy(P.c2l, car2lat, sizeof(P.c2l));
  memcpy(P.lid, lat_i_dist, sizeof(P.lid));
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const long long nn = g.nn;

  c2g_basins* res = new c2g_basins();
  res->ctx = ctx; res->kind = 0; res->gridh = handle;
  res->n[0] = n1; res->n[1] = n2; res->n[2] = n3; res->nn = nn;
  struct Guard { c2g_basins* r; bool ok = false; ~Guard() { if (!ok) c2g_basins_free(r); } } guard{res};

  C2G_CUDA(ctx, cudaMalloc(&res->d_label, sizeof(int) * nn));
  int* label = res->d_label;

  // ---- K0: candidate maxima ----
  const int c41 = (n1 + 3) / 4, c42 = (n2 + 3) / 4, c43 = (n3 + 3) / 4;
  const int c21 = (n1 + 1) / 2, c22 = (n2 + 1) / 2, c23 = (n3 + 1) / 2;
  const size_t ncube4 = (size_t)c41 * c42 * c43, ncube2 = (size_t)c21 * c22 * c23;
  DevBuf b_cube4, b_cube2, b_cand, b_cnt;
  C2G_CUDA(ctx, cudaMalloc(&b_cube4.p, ncube4));
  C2G_CUDA(ctx, cudaMalloc(&b_cube2.p, ncube2));
  int maxcand = (int)std::min<long long>(nn, std::max<long long>(1 << 16, nn / 64));
  C2G_CUDA(ctx, cudaMalloc(&b_cnt.p, 64));
  // counters: [0] ncand, [1] nlist, [2] noverflow, [3] err ; [4..5] nsteps (u64)
  int* cnt = b_cnt.as<int>();
  unsigned long long* nsteps = (unsigned long long*)(cnt + 4);
  int hcnt[8];
  for (int attempt = 0;; attempt++) {
    C2G_CUDA(ctx, cudaMalloc(&b_cand.p, sizeof(int) * (size_t)maxcand));
    C2G_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_cube4.p, 0, ncube4, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_cube2.p, 0, ncube2, st));
    dim3 grid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, (n3 + TZ - 1) / TZ);
    const size_t smem = sizeof(double) * ((TZ + 2) * (TY + 2) * (TX + 2) + (TZ + 2) * (TY + 2) * TX);
    static bool attr_set = false;
    if (!attr_set) {
      C2G_CUDA(ctx, cudaFuncSetAttribute(k_maxima, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    ctx->prof_begin("bader_maxima");
    k_maxima<<<grid, 256, smem, st>>>(P, g.d, b_cand.as<int>(), cnt, maxcand, b_cube4.as<unsigned char>(),
                                      b_cube2.as<unsigned char>());
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hcnt[0] <= maxcand) break;
    if (attempt > 0) return ctx->fail(C2G_ERR_OVERFLOW, "candidate maxima list overflow");
    cudaFree(b_cand.p); b_cand.p = nullptr;
    maxcand = hcnt[0];
  }
  const int ncand = hcnt[0];
  if (ncand == 0) return ctx->fail(C2G_ERR_STATE, "c2g_bader_assign: the field has no local maximum (NaN input?)");
  std::vector<int> cand(ncand);
  C2G_CUDA(ctx, cudaMemcpy(cand.data(), b_cand.p, sizeof(int) * ncand, cudaMemcpyDeviceToHost));
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
  C2G_CUDA(ctx, cudaMalloc(&b_hk.p, sizeof(int) * hsize));
  C2G_CUDA(ctx, cudaMalloc(&b_hv.p, sizeof(int) * hsize));
  C2G_CUDA(ctx, cudaMalloc(&b_reached.p, ncand));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hk.p, hk.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hv.p, hv.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemsetAsync(b_reached.p, 0, ncand, st));
  MaxHash h{b_hk.as<int>(), b_hv.as<int>(), hsize - 1};
  unsigned char* reached = b_reached.as<unsigned char>();

  // work list / overflow list
  DevBuf b_list, b_over;
  const long long listcap = (algo == C2G_BADER_EXACT) ? 1 : nn;
  C2G_CUDA(ctx, cudaMalloc(&b_list.p, sizeof(int) * (size_t)listcap));
  long long overcap = std::max<long long>(1024, nn / 16);
  C2G_CUDA(ctx, cudaMalloc(&b_over.p, sizeof(int) * (size_t)overcap));
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
    const int bigcap = (int)std::min<long long>(nn, 1 << 22);
    const int chunk = (int)std::max<long long>(1, std::min<long long>(nov, (1ll << 31) / bigcap));  // <= 8 GiB scratch
    DevBuf b_scr;
    C2G_CUDA(ctx, cudaMalloc(&b_scr.p, sizeof(int) * (size_t)chunk * bigcap));
    for (int off = 0; off < nov; off += chunk) {
      const int c = std::min(chunk, nov - off);
      ctx->prof_begin("bader_walk_big");
      k_walk_big<<<c2g_blocks_for(c, 64), 64, 0, st>>>(P, g.d, over + off, c, label, h, reached, b_scr.as<int>(), bigcap,
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
  auto walk_list = [&](int count, const char* name) -> int {
    if (count <= 0) return C2G_OK;
    // overflow list must be able to hold every walker of this launch in the worst case
    if (count > overcap) {
      cudaFree(b_over.p); b_over.p = nullptr;
      overcap = count;
      C2G_CUDA(ctx, cudaMalloc(&b_over.p, sizeof(int) * (size_t)overcap));
      over = b_over.as<int>();
    }
    ctx->prof_begin(name);
    k_walk_list<<<c2g_blocks_for(count, 128), 128, 0, st>>>(P, g.d, list, count, label, h, reached, over, cnt + 2, cnt + 3,
                                                            algo == C2G_BADER_EXACT ? nsteps : nullptr);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    walked += count;
    return drain_overflow();
  };
  auto walk_lattice = [&](int s, const char* name) -> int {
    const int m1 = (n1 + s - 1) / s, m2 = (n2 + s - 1) / s, m3 = (n3 + s - 1) / s;
    const long long m = (long long)m1 * m2 * m3;
    if (m > overcap) {
      cudaFree(b_over.p); b_over.p = nullptr;
      overcap = m;
      C2G_CUDA(ctx, cudaMalloc(&b_over.p, sizeof(int) * (size_t)overcap));
      over = b_over.as<int>();
    }
    ctx->prof_begin(name);
    k_walk_lattice<<<c2g_blocks_for(m, 128), 128, 0, st>>>(P, g.d, s, m1, m2, m3, label, h, reached, over, cnt + 2, cnt + 3,
                                                           algo == C2G_BADER_EXACT ? nsteps : nullptr);
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
    for (int s = 4; s >= 2; s >>= 1) {
      const int c1 = (n1 + s - 1) / s, c2 = (n2 + s - 1) / s, c3 = (n3 + s - 1) / s;
      const long long nc = (long long)c1 * c2 * c3;
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));
      ctx->prof_begin(s == 4 ? "bader_classify4" : "bader_classify2");
      k_classify<<<c2g_blocks_for(nc, 256), 256, 0, st>>>(n1, n2, n3, s, label,
                                                          s == 4 ? b_cube4.as<unsigned char>() : b_cube2.as<unsigned char>(),
                                                          list, cnt + 1);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
      C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
      C2G_CUDA(ctx, cudaStreamSynchronize(st));
      if ((rc = walk_list(hcnt[1], s == 4 ? "bader_walk_l2" : "bader_walk_l1")) != C2G_OK) return rc;
    }
    // edge fix until no filled point is adjacent to a different label
    for (;;) {
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));
      dim3 grid((n1 + TX - 1) / TX, (n2 + TY - 1) / TY, (n3 + TZ - 1) / TZ);
      ctx->prof_begin("bader_edgefix");
      k_edgefix<<<grid, 256, 0, st>>>(n1, n2, n3, label, list, cnt + 1);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
      C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
      C2G_CUDA(ctx, cudaStreamSynchronize(st));
      fixpasses++;
      if (hcnt[1] == 0) break;
      fixpts += hcnt[1];
      if ((rc = walk_list(hcnt[1], "bader_walk_fix")) != C2G_OK) return rc;
      if (fixpasses > 1000) return ctx->fail(C2G_ERR_STATE, "edge refinement did not converge");
    }
  }

  // ---- maxima actually reached, output order, compaction ----
  std::vector<unsigned char> hreached(ncand);
  C2G_CUDA(ctx, cudaMemcpy(hreached.data(), reached, ncand, cudaMemcpyDeviceToHost));
  std::vector<int> cand2out(ncand, -1);
  int nmax = 0;
  for (int i = 0; i < ncand; i++)
    if (hreached[i]) cand2out[i] = nmax++;
  res->nmax = nmax;
  res->max_lin.resize(nmax);
  for (int i = 0; i < ncand; i++)
    if (cand2out[i] >= 0) res->max_lin[cand2out[i]] = cand[i];
  DevBuf b_c2o, b_counts;
  C2G_CUDA(ctx, cudaMalloc(&b_c2o.p, sizeof(int) * ncand));
  C2G_CUDA(ctx, cudaMalloc(&b_counts.p, sizeof(unsigned long long) * nmax));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_c2o.p, cand2out.data(), sizeof(int) * ncand, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemsetAsync(b_counts.p, 0, sizeof(unsigned long long) * nmax, st));
  const int nblk = ctx->nsm * 8;
  ctx->prof_begin("bader_compact");
  k_compact<<<nblk, 256, 0, st>>>(nn, label, h, b_c2o.as<int>(), b_counts.as<unsigned long long>(), cnt + 3);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  res->counts.resize(nmax);
  C2G_CUDA(ctx, cudaMemcpyAsync(res->counts.data(), b_counts.p, sizeof(unsigned long long) * nmax, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, 32, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  if (hcnt[3] != 0) return ctx->fail(C2G_ERR_NEWMAX, "label compaction found a terminal that is not a known maximum");

  if (order == C2G_ORDER_SCAN && nmax > 1) {
    DevBuf b_first, b_perm;
    C2G_CUDA(ctx, cudaMalloc(&b_first.p, sizeof(int) * nmax));
    C2G_CUDA(ctx, cudaMemsetAsync(b_first.p, 0x7f, sizeof(int) * nmax, st));
    ctx->prof_begin("bader_firstpoint");
    k_firstpoint<<<nblk, 256, 0, st>>>(n1, n2, n3, label, b_first.as<int>());
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    std::vector<int> first(nmax);
    C2G_CUDA(ctx, cudaMemcpy(first.data(), b_first.p, sizeof(int) * nmax, cudaMemcpyDeviceToHost));
    std::vector<int> idx(nmax), perm(nmax);
    for (int i = 0; i < nmax; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return first[a] < first[b]; });
    std::vector<int> ml(nmax);
    std::vector<long long> cc(nmax);
    for (int k = 0; k < nmax; k++) { perm[idx[k]] = k; ml[k] = res->max_lin[idx[k]]; cc[k] = res->counts[idx[k]]; }
    res->max_lin.swap(ml);
    res->counts.swap(cc);
    C2G_CUDA(ctx, cudaMalloc(&b_perm.p, sizeof(int) * nmax));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_perm.p, perm.data(), sizeof(int) * nmax, cudaMemcpyHostToDevice, st));
    ctx->prof_begin("bader_permute");
    k_permute_labels<<<nblk, 256, 0, st>>>(nn, label, b_perm.as<int>());
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
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
