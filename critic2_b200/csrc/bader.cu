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
//   C2G_BADER_FAST   hierarchical.  The points of a stride-L0 lattice (L0 <= 32) walk.  Then, level by
//                    level (s = L0 .. 2): a stride-s cube whose 8 corners agree and that holds no local
//                    maximum is "uniform" and its stride-s/2 points are filled with the corner label;
//                    the stride-s/2 points of the other cubes walk, and a walk stops as soon as it
//                    reaches a point whose nearest cube-grid vertex is surrounded by 8 uniform cubes of
//                    one label (octet certificate) -- the analogue of the reference's known==2 interior
//                    points at which max_neargrid stops (:447).  Finally every filled point that has a
//                    26-neighbour with a different label walks too, until none is left: the fixed point
//                    refine_edge enforces (:300-422, every edge point carries the label of its own
//                    trajectory).  Early stops are logged and re-walked if a later label change voids
//                    the certificate they relied on (k_requeue).
// Kernels: k_maxima / k_maxima2 (candidate maxima, streaming), k_walk (persistent path-free walkers),
// k_walk_big (the rare walks that need the reference's path search), k_classify + k_vsafe (cube
// classification and certificates per level), k_fill_edge_v (last-level fills + edge detection, driven
// by the certificates), k_requeue, k_items_* (spatially ordered work items).
// Labels are indices into the sorted list of candidate maxima; bit 31 (FILLBIT) marks points that were
// filled, not walked (consumers mask it).
// Arithmetic: IEEE fp64, evaluation order of the Fortran source, NO fused multiply-add (this
// translation unit is compiled with -fmad=false; the reference is built -O3 without -march/-ffast-math).
#include "common.cuh"
#include "group.h"

#include <algorithm>

namespace {

struct BaderParams {
  int n1, n2, n3;
  int in1, in2, in3;  // max(n - 6, 0): points with 3 <= p < n-3 take steps that never touch a periodic seam
  double c2l[9];   // car2lat, column-major
  double lid[27];  // lat_i_dist, (d1+1)*9+(d2+1)*3+(d3+1)
  unsigned mg1, mg12;  // id / n1 and id / (n1*n2) for 0 <= id < 2^31 as (id * mg) >> (31 + sh), see fastdiv_make
  int sh1, sh12;
};

// Division of 0 <= v < 2^31 by a run-time constant d >= 1 (round-up method): with l = ceil(log2 d) and
// m = floor(2^(31+l) / d) + 1 (< 2^32), m*d - 2^(31+l) is in [1, d] <= 2^l, hence floor(v*m / 2^(31+l)) = v / d.
inline void fastdiv_make(unsigned d, unsigned& m, int& sh) {
  int l = 0;
  while ((1ull << l) < d) l++;
  m = (unsigned)((1ull << (31 + l)) / d + 1ull);
  sh = l;
}
__device__ __forceinline__ int fastdiv(int v, unsigned m, int sh) {
  return (int)(((unsigned long long)(unsigned)v * m) >> (31 + sh));
}

constexpr unsigned FILLBIT = 0x80000000u;
constexpr int LMASK = 0x7fffffff;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAXLEV = 5;               // cube strides 2,4,8,16,32

__device__ __forceinline__ int wrapx(int p, int n) {
  if (p < 0) p += n;
  if (p >= n) p -= n;
  if (p < 0 || p >= n) p = ((p % n) + n) % n;  // grids narrower than a tile
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

// is_max, bader@proc.f90:571-597 (no neighbour strictly greater); general periodic wrap
__device__ __noinline__ bool dev_is_max(int n1, int n2, int n3, const double* __restrict__ rho, int x, int y, int z) {
  const double r0 = __ldg(rho + x + (size_t)n1 * (y + (size_t)n2 * z));
  bool ismax = true;
#pragma unroll 1
  for (int d3 = -1; d3 <= 1; d3++) {
    const int zz = wrapx(z + d3, n3);
#pragma unroll 1
    for (int d2 = -1; d2 <= 1; d2++) {
      const int yy = wrapx(y + d2, n2);
      const double* row = rho + (size_t)n1 * (yy + (size_t)n2 * zz);
#pragma unroll
      for (int d1 = -1; d1 <= 1; d1++) {
        const int xx = wrapx(x + d1, n1);
        if (__ldg(row + xx) > r0) ismax = false;
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

// Early-termination map (FAST algorithm only), the analogue of the reference's known==2 interior points.
//   octet = 0: safe[c] >= 0 means the stride-2^shift cube c and its 26 neighbouring cubes are uniformly
//              labelled safe[c] and hold no maximum (every point of c has a uniform 5x5x5 neighbourhood);
//   octet = 1: safe[v] >= 0 means the 8 cubes that meet at cube-grid vertex v are uniformly labelled and hold
//              no maximum; a point q is looked up at its nearest vertex v = (q + 2^(shift-1)) >> shift, whose
//              8 cubes contain the whole 3x3x3 neighbourhood of q;
//   pack  = 1: (octet map of the stride-2 cubes only) an entry carries, above its 23-bit label, one flag for each of
//              the 8 points {2v-1, 2v}^3 whose nearest vertex is v: the octets of ALL the vertices that cover the
//              point's 5x5x5 neighbourhood hold with the same label (k_pack5).  A walk stops at such a point only
//              when its flag is set: the margin at which the reference's refine_edge walks stop (known == 2 is
//              cleared in the 26-neighbourhood of every edge point, bader@proc.f90:756-771, :367-388, so a walk only
//              ends on a point whose 26 neighbours are not edge points either).  Octets of wider cubes cover the
//              5x5x5 neighbourhood of their points by themselves.
// Covers the owned cube layers only.
struct SafeMap {
  int* safe;  // nullptr = disabled; entries are set to -1 when a label inside the certified region changes
  int shift, c1, c2, c3, zlo, nzl, octet, wrapz;
  int hf;  // octet: half a cube, (1 << shift) >> 1
  int pack;
};
constexpr int CERT_LBITS = 23;                       // label bits of a packed entry
constexpr int CERT_LMASK = (1 << CERT_LBITS) - 1;
constexpr int CERT_VOID = (int)0x80000000;           // packed entry without a certificate: negative AND no flag set, so
                                                     // that the walkers test one bit and everyone else the sign
// flag of the point (x, y, cz) in a packed entry: bit CERT_LBITS + (x odd) + 2 (y odd) + 4 (cz odd)
__device__ __forceinline__ int cert_flagbit(int x, int y, int cz) {
  return CERT_LBITS + ((x & 1) | ((y & 1) << 1) | ((cz & 1) << 2));
}
// label certified for the point (x, y, cz) by entry e of a packed map, or -1
__device__ __forceinline__ int cert_unpack(int e, int x, int y, int cz) {
  return ((e >> cert_flagbit(x, y, cz)) & 1) ? (e & CERT_LMASK) : -1;
}
constexpr int STOP_SHIFT = 28;             // stop code = (level index << 28) | map index
constexpr int STOP_MASK = (1 << STOP_SHIFT) - 1;

// ------------------------------------------------------------------------------------------------
// one near-grid step (step_neargrid, bader@proc.f90:455-494) of a trajectory held in WState.
// The reference's "known(pm)==1" revisit test (:484-488) is answered exactly: a visited point can only
// be hit again when rho(pm) <= max rho along the path, and only then the stored path is searched.
// ORTHO: car2lat is diagonal (orthogonal cell); the skipped products are exact zeros, so the result is
// bit-identical to the general expression.
// returns 0 = moved on; 1 = ended on the maximum `out` (linear id); 2 = entered a safe cube labelled `out`;
//         3 = path buffer full.
// ------------------------------------------------------------------------------------------------
struct WState {
  int id, x, y, z, len;
  double dr0, dr1, dr2, rhomax, r0;
};

__device__ __forceinline__ void walk_init(const BaderParams& P, const double* __restrict__ rho, WState& w, int start) {
  w.id = start;
  w.x = start % P.n1;
  const int t = start / P.n1;
  w.y = t % P.n2;
  w.z = t / P.n2;
  w.len = 0;
  w.dr0 = w.dr1 = w.dr2 = 0.0;
  w.rhomax = -1.0e300;
  w.r0 = __ldg(rho + start);
}

__device__ __forceinline__ int safe_lookup(const SafeMap& sm, int nx, int ny, int nz, int& mapidx);
template <bool ORTHO>
__device__ __forceinline__ int walk_step(const BaderParams& P, const double* __restrict__ rho, const MaxHash& h,
                                         const SafeMap& sm, WState& w, int* path, int cap, int& out, int& sli) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int s2 = n1, s3 = n1 * n2;
  const int x = w.x, y = w.y, z = w.z, id = w.id;
  const double r0 = w.r0;
  const double* c = rho + id;
  const double rxp = __ldg(c + ((x + 1 == n1) ? 1 - n1 : 1)), rxm = __ldg(c + ((x == 0) ? n1 - 1 : -1));
  const double ryp = __ldg(c + ((y + 1 == n2) ? s2 - s3 : s2)), rym = __ldg(c + ((y == 0) ? s3 - s2 : -s2));
  const double rzp = __ldg(c + ((z + 1 == n3) ? s3 - s3 * n3 : s3)), rzm = __ldg(c + ((z == 0) ? s3 * n3 - s3 : -s3));
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
    w.dr0 = w.dr1 = w.dr2 = 0.0;
    // is_max (:571-597) == membership in the candidate list built by k_maxima with the same predicate
    if (hash_lookup(h, id) >= 0) { out = id; return 1; }
    nid = dev_step_ongrid(P, rho, x, y, z, r0);
    nx = nid % n1; const int t = nid / n1; ny = t % n2; nz = t / n2;
  } else {  // (:477-483)
    const double coeff = 1.0 / gmax;
    g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
    const double a0 = nint_small(g0), a1 = nint_small(g1), a2 = nint_small(g2);
    double dr0 = w.dr0 + g0 - a0, dr1 = w.dr1 + g1 - a1, dr2 = w.dr2 + g2 - a2;
    const double b0 = nint_small(dr0), b1 = nint_small(dr1), b2 = nint_small(dr2);
    w.dr0 = dr0 - b0; w.dr1 = dr1 - b1; w.dr2 = dr2 - b2;
    nx = wrap2(x + (int)(a0 + b0), n1);
    ny = wrap2(y + (int)(a1 + b1), n2);
    nz = wrap2(z + (int)(a2 + b2), n3);
    nid = nx + n1 * (ny + n2 * nz);
  }
  // known(p) = 1 (:484)
  if (w.len >= cap) return 3;
  path[w.len++] = id;
  w.rhomax = fmax(w.rhomax, r0);
  double rn = __ldg(rho + nid);
  if (rn <= w.rhomax) {  // only then pm can be a point of this path (:487)
    if (dev_on_path(path, w.len, nid)) {
      nid = dev_step_ongrid(P, rho, x, y, z, r0);
      nx = nid % n1; const int t = nid / n1; ny = t % n2; nz = t / n2;
      w.dr0 = w.dr1 = w.dr2 = 0.0;
      rn = __ldg(rho + nid);
    }
  }
  if (nid == id) { out = id; return 1; }  // did not move: maximum (:439)
  w.id = nid; w.r0 = rn;
  w.x = nx; w.y = ny; w.z = nz;
  sli = -1;
  if (sm.safe) {  // quit at a known interior point (:447)
    const int sl = safe_lookup(sm, nx, ny, nz, sli);
    if (sl >= 0) { out = sl; return 2; }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined variant used by the persistent walkers: as soon as the next point is known, ALL the
// loads the next step will need (its density, its 6 neighbours, its early-termination label) are issued
// together, so that a step costs one memory round trip instead of three dependent ones.  Same arithmetic,
// same decisions, same results as walk_step.
// ------------------------------------------------------------------------------------------------
struct Nb {
  double xp, xm, yp, ym, zp, zm;
};
__device__ __forceinline__ void load_nb(const BaderParams& P, const double* __restrict__ rho, int id, int x, int y, int z, Nb& nb) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int s2 = n1, s3 = n1 * n2;
  const double* c = rho + id;
  nb.xp = __ldg(c + ((x + 1 == n1) ? 1 - n1 : 1));
  nb.xm = __ldg(c + ((x == 0) ? n1 - 1 : -1));
  nb.yp = __ldg(c + ((y + 1 == n2) ? s2 - s3 : s2));
  nb.ym = __ldg(c + ((y == 0) ? s3 - s2 : -s2));
  nb.zp = __ldg(c + ((z + 1 == n3) ? s3 - s3 * n3 : s3));
  nb.zm = __ldg(c + ((z == 0) ? s3 * n3 - s3 : -s3));
}
__device__ __forceinline__ int safe_lookup(const SafeMap& sm, int nx, int ny, int nz, int& mapidx) {
  mapidx = -1;
  if (!sm.safe) return -1;
  const int cz = nz - sm.zlo;
  if (sm.octet) {  // one predicated load, no divergent control flow
    const int hf = (1 << sm.shift) >> 1;
    int vx = (nx + hf) >> sm.shift, vy = (ny + hf) >> sm.shift, vz = (cz + hf) >> sm.shift;
    vx = vx == sm.c1 ? 0 : vx;
    vy = vy == sm.c2 ? 0 : vy;
    bool ok = (unsigned)cz < (unsigned)sm.nzl;
    if (vz == sm.c3) { ok = ok && sm.wrapz; vz = 0; }
    if (!ok) return -1;
    mapidx = vx + sm.c1 * (vy + sm.c2 * vz);
    const int e = sm.safe[mapidx];  // plain load: entries may be invalidated while walkers run
    return sm.pack ? cert_unpack(e, nx, ny, cz) : e;
  }
  if (cz < 0 || cz >= sm.nzl) return -1;
  mapidx = (nx >> sm.shift) + sm.c1 * ((ny >> sm.shift) + sm.c2 * (cz >> sm.shift));
  return sm.safe[mapidx];
}
// The same lookup with the branch on the certificate kind hoisted (warp-uniform) and nothing recomputed per call.
__device__ __forceinline__ int safe_lookup2(const SafeMap& sm, int x, int y, int z, int& mapidx) {
  const int cz = z - sm.zlo;
  int vx, vy, vz;
  bool ok = (unsigned)cz < (unsigned)sm.nzl;
  if (sm.octet) {
    const int hf = sm.hf;
    vx = (x + hf) >> sm.shift; vy = (y + hf) >> sm.shift; vz = (cz + hf) >> sm.shift;
    vx = vx == sm.c1 ? 0 : vx;
    vy = vy == sm.c2 ? 0 : vy;
    const bool top = vz == sm.c3;
    ok = ok && (!top || sm.wrapz);
    vz = top ? 0 : vz;
  } else {
    vx = x >> sm.shift; vy = y >> sm.shift; vz = cz >> sm.shift;
  }
  mapidx = vx + sm.c1 * (vy + sm.c2 * vz);
  const int e = ok ? sm.safe[mapidx] : CERT_VOID;  // plain load: entries may be invalidated while walkers run
  return sm.pack ? cert_unpack(e, x, y, cz) : e;     // (CERT_VOID is negative: "none" for the other map kinds too)
}
// w: current point with w.r0, nb = its neighbours, sl = early-termination label of its cube (all loaded by
// the previous call or by the caller for the start point, where sl must be -1)
// Fortran nint of |v| < 1.5 as a double (two fp64 compares, like nint_small) and as an int taken from the high
// word of that double (0x3ff00000 / 0xbff00000 / 0) -- no fp64 -> int conversion
__device__ __forceinline__ double nint_di(double v, int& d) {
  const double a = v >= 0.5 ? 1.0 : (v <= -0.5 ? -1.0 : 0.0);
  const int hi = __double2hiint(a);
  d = (hi >> 31) | ((hi >> 29) & 1);
  return a;
}
template <bool ORTHO>
__device__ __forceinline__ int walk_step_pipe(const BaderParams& P, const double* __restrict__ rho, const MaxHash& h,
                                              const SafeMap& sm, WState& w, Nb& nb, int& sl, int& sli, int& out) {
  if (sl >= 0) { out = sl; return 2; }  // quit at a known interior point (:447); sli = where (for the stop log)
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int x = w.x, y = w.y, z = w.z, id = w.id;
  const double r0 = w.r0;
  // rho_grad_dir (:532-567)
  double gl0 = (nb.xp - nb.xm) * 0.5, gl1 = (nb.yp - nb.ym) * 0.5, gl2 = (nb.zp - nb.zm) * 0.5;
  {
    const bool z0 = (nb.xp < r0) & (nb.xm < r0), z1 = (nb.yp < r0) & (nb.ym < r0), z2 = (nb.zp < r0) & (nb.zm < r0);
    gl0 = z0 ? 0.0 : gl0;
    gl1 = z1 ? 0.0 : gl1;
    gl2 = z2 ? 0.0 : gl2;
  }
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
  // maxval(abs(grad)) with plain compares (= fmax for every non-NaN input; a NaN ends in k_walk_big via rn <= rhomax)
  double gmax = fabs(g0);
  {
    const double t1 = fabs(g1), t2 = fabs(g2);
    gmax = t1 > gmax ? t1 : gmax;
    gmax = t2 > gmax ? t2 : gmax;
  }
  // the point and its next point stay clear of the periodic seams: no wrapping anywhere in this step
  const bool inner = (unsigned)(x - 3) < (unsigned)P.in1 && (unsigned)(y - 3) < (unsigned)P.in2 &&
                     (unsigned)(z - 3) < (unsigned)P.in3;
  int nid, nx, ny, nz;
  if (gmax < 1e-30) {  // (:468-476)
    w.dr0 = w.dr1 = w.dr2 = 0.0;
    if (hash_lookup(h, id) >= 0) { out = id; return 1; }
    nid = dev_step_ongrid(P, rho, x, y, z, r0);
    nx = nid % n1; const int t = nid / n1; ny = t % n2; nz = t / n2;
  } else {  // (:477-483)
    const double coeff = 1.0 / gmax;
    g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
    int d0, d1, d2, e0, e1, e2;
    const double a0 = nint_di(g0, d0), a1 = nint_di(g1, d1), a2 = nint_di(g2, d2);
    const double dr0 = w.dr0 + g0 - a0, dr1 = w.dr1 + g1 - a1, dr2 = w.dr2 + g2 - a2;
    const double b0 = nint_di(dr0, e0), b1 = nint_di(dr1, e1), b2 = nint_di(dr2, e2);
    w.dr0 = dr0 - b0; w.dr1 = dr1 - b1; w.dr2 = dr2 - b2;
    d0 += e0; d1 += e1; d2 += e2;
    if (inner) {
      nx = x + d0; ny = y + d1; nz = z + d2;
      nid = id + d0 + n1 * (d1 + n2 * d2);
    } else {
      nx = wrap2(x + d0, n1);
      ny = wrap2(y + d1, n2);
      nz = wrap2(z + d2, n3);
      nid = nx + n1 * (ny + n2 * nz);
    }
  }
  w.len++;             // known(p) = 1 (:484)
  w.rhomax = r0 > w.rhomax ? r0 : w.rhomax;
  // everything the next step needs, in one batch
  double rn;
  if (inner) {
    const double* c = rho + nid;
    const int s2 = n1, s3 = n1 * n2;
    rn = __ldg(c);
    nb.xp = __ldg(c + 1); nb.xm = __ldg(c - 1);
    nb.yp = __ldg(c + s2); nb.ym = __ldg(c - s2);
    nb.zp = __ldg(c + s3); nb.zm = __ldg(c - s3);
  } else {
    rn = __ldg(rho + nid);
    load_nb(P, rho, nid, nx, ny, nz, nb);
  }
  sl = safe_lookup(sm, nx, ny, nz, sli);
  // The reference now asks whether pm is a point of this path (known(pm) == 1, :487).  That can only be so when
  // rho(pm) <= the largest density seen on the path.  On smooth fields this never happens (0 of 7.6e7 steps at
  // 512^3), so the persistent walkers keep NO path: such a walk is handed to k_walk_big, which repeats it from
  // the start with the full path in global memory and answers the question exactly.
  if (rn <= w.rhomax) return 3;
  if (nid == id) { out = id; return 1; }  // did not move: maximum (:439)
  w.id = nid; w.r0 = rn;
  w.x = nx; w.y = ny; w.z = nz;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// z-slab bookkeeping.  A rank owns the global planes [zlo, zhi) (interior boundaries are multiples of the
// top lattice stride so that no cube straddles two ranks); rho is replicated on every rank, labels are
// sharded.  The label buffer holds nzl + 2 planes: local plane 0 = global plane zlo-1 (halo below),
// planes 1..nzl = owned, plane nzl+1 = global plane zhi (halo above); halos are periodic images and are
// filled by exchange_halos (NCCL send/recv between ranks, a device copy when the rank is its own
// neighbour).  periodic = 1: the rank owns every plane (single GPU) and kernels may wrap z themselves.
// ------------------------------------------------------------------------------------------------
struct Slab {
  int zlo, zhi, nzl, periodic;
};

// "has a local maximum" flags of the stride-(2<<i) cubes, i = 0..nlev-1
struct CubeFlags {
  unsigned char* p[MAXLEV];
  int c1[MAXLEV], c2[MAXLEV];
  int nlev;
};

// 3x3 in-plane combination on a (32 x TY) thread tile with a 1-cell halo in shared memory.
// s: [(TY+2)][34] buffer of T.  Every thread contributes its own value; threads with hslot >= 0 also a
// halo value.  Returns Op(3x3 neighbourhood) of the thread's point.  One __syncthreads per call; the
// caller alternates between two buffers so that no second barrier is needed.
struct OpAgree {  // common value of three labels, or -1
  static __device__ __forceinline__ int c3(int a, int b, int c) { return (a == b && b == c) ? a : -1; }
};
template <class T, class Op>
__device__ __forceinline__ T plane3x3(T* s, T own, int hslot, T hval, int lx, int ly) {
  s[(ly + 1) * 34 + lx + 1] = own;
  if (hslot >= 0) s[hslot] = hval;
  __syncthreads();
  const T* col = s + ly * 34 + lx + 1;
  const T a = Op::c3(col[0], own, col[68]);
  // the two edge lanes take their outer neighbour column from the halo; every lane executes the same code
  const int e = (lx == 0) ? -1 : ((lx == 31) ? 1 : 0);
  const T ae = Op::c3(col[e], col[e + 34], col[e + 68]);
  const T lu = __shfl_up_sync(FULL, a, 1), rd = __shfl_down_sync(FULL, a, 1);
  const T l = (lx == 0) ? ae : lu, r = (lx == 31) ? ae : rd;
  return Op::c3(l, a, r);
}
constexpr int TY = 8;                       // tile rows (256 threads)
constexpr int NHALO = 2 * 34 + 2 * TY;      // halo elements of a 32 x TY tile
// halo element h -> (col, row) in tile coordinates (-1..32, -1..TY)
__device__ __forceinline__ void halo_decode(int h, int& col, int& row) {
  if (h < 34) { row = -1; col = h - 1; }
  else if (h < 68) { row = TY; col = h - 35; }
  else if (h < 68 + TY) { col = -1; row = h - 68; }
  else { col = 32; row = h - 68 - TY; }
}

// K0: candidate maxima (26-neighbour, is_max).  The grid is streamed once: a warp owns a row segment of 32 x
// points and marches along z with the previous / current / next plane values in registers; the x neighbours
// come from shuffles.  A point can only be a maximum if it is >= its two x and two z neighbours (necessary,
// exact in fp64); the rare survivors (points near the line through a nucleus) get the full 26-neighbour test.
// No shared memory, no barriers: one coalesced 8-byte load per point.  Marks the cubes of every level that
// contain a maximum.
constexpr int MZC = 32;  // planes per block
#ifndef C2G_MZP
#define C2G_MZP 4
#endif
#ifndef C2G_MAX2_MINB
#define C2G_MAX2_MINB 1
#endif
constexpr int MZP = C2G_MZP;   // planes per prefetch group of k_maxima
__global__ void __launch_bounds__(256) k_maxima(const __grid_constant__ BaderParams P, const Slab S,
                                                const double* __restrict__ rho, int* __restrict__ cand,
                                                int* __restrict__ ncand, int maxcand, const __grid_constant__ CubeFlags CF) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const size_t s3 = (size_t)n1 * n2;
  // a block = 256 consecutive x points of one row (2 KB contiguous per plane: DRAM-row friendly)
  const int lane = threadIdx.x & 31;
  const int gx = blockIdx.x * 256 + threadIdx.x, gy = blockIdx.y;
  if (gx - lane >= n1) return;  // whole warp past the end of the row
  const int z0 = S.zlo + blockIdx.z * MZC, z1 = min(z0 + MZC, S.zhi);
  const bool valid = gx < n1;
  // lanes past the end of the row hold the periodic images, so that the shuffles give the right neighbours
  const double* cp = rho + wrapx(gx, n1) + (size_t)n1 * gy;
  // lane 0 / lane 31 also fetch the neighbour outside the warp's segment
  const bool edge = lane == 0 || lane == 31;
  const double* ep = rho + wrapx(lane == 0 ? gx - 1 : gx + 1, n1) + (size_t)n1 * gy;
  const size_t wrapback = s3 * (size_t)n3;
  int wz = wrapx(z0 - 1, n3);
  cp += s3 * wz; ep += s3 * wz;
  auto next_plane = [&]() {
    cp += s3; ep += s3;
    if (++wz == n3) { wz = 0; cp -= wrapback; ep -= wrapback; }
  };
  // Planes are fetched MZP at a time, one group ahead of the group being tested, so that every thread keeps
  // MZP (+ MZP for the edge lanes) independent 8-byte loads in flight: ~40 KB per SM, what the HBM latency needs.
  double a[MZP + 2], e[MZP + 2];  // values of planes zb-1 .. zb+MZP of the own point / of the edge neighbour
#pragma unroll
  for (int k = 0; k < MZP + 2; k++) {
    const bool need = z0 - 1 + k <= z1;  // planes z0-1 .. z1
    a[k] = need ? __ldg(cp) : 0.0;
    e[k] = (need && edge) ? __ldg(ep) : 0.0;
    if (need) next_plane();
  }
  for (int zb = z0; zb < z1; zb += MZP) {
    double an[MZP], en[MZP];  // planes zb+MZP+1 .. zb+2*MZP
#pragma unroll
    for (int k = 0; k < MZP; k++) {
      const bool need = zb + MZP + 1 + k <= z1;
      an[k] = need ? __ldg(cp) : 0.0;
      en[k] = (need && edge) ? __ldg(ep) : 0.0;
      if (need) next_plane();
    }
#pragma unroll
    for (int k = 0; k < MZP; k++) {
      const int iz = zb + k;
      if (iz < z1) {  // warp-uniform
        const double vm = a[k], vc = a[k + 1], vp = a[k + 2], ec = e[k + 1];
        double xl = __shfl_up_sync(FULL, vc, 1), xr = __shfl_down_sync(FULL, vc, 1);
        if (lane == 0) xl = ec;
        if (lane == 31) xr = ec;
        if (valid && vc >= xl && vc >= xr && vc >= vm && vc >= vp) {
          if (dev_is_max(n1, n2, n3, rho, gx, gy, iz)) {
            const int slot = atomicAdd(ncand, 1);
            if (slot < maxcand) cand[slot] = gx + n1 * (gy + n2 * iz);
            for (int i = 0; i < CF.nlev; i++)
              CF.p[i][(gx >> (i + 1)) + CF.c1[i] * ((gy >> (i + 1)) + (size_t)CF.c2[i] * ((iz - S.zlo) >> (i + 1)))] = 1;
          }
        }
      }
    }
    a[0] = a[MZP]; a[1] = a[MZP + 1]; e[0] = e[MZP]; e[1] = e[MZP + 1];
#pragma unroll
    for (int k = 0; k < MZP; k++) { a[k + 2] = an[k]; e[k + 2] = en[k]; }
  }
}

// Same pass for even n1: a lane owns TWO consecutive x points (one 16-byte load per plane), which halves the
// shuffles, the edge handling and the address arithmetic per point -- k_maxima is issue-bound (75 % of the issue
// slots at 3.6 TB/s), not load-bound.  A block = 128 threads = 256 consecutive x points of one row.
__global__ void __launch_bounds__(128, C2G_MAX2_MINB) k_maxima2(const __grid_constant__ BaderParams P, const Slab S,
                                                 const double* __restrict__ rho, int* __restrict__ cand,
                                                 int* __restrict__ ncand, int maxcand, const __grid_constant__ CubeFlags CF) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const size_t s3 = (size_t)n1 * n2;
  const int lane = threadIdx.x & 31;
  const int gx = blockIdx.x * 256 + 2 * threadIdx.x, gy = blockIdx.y;  // first of the two points
  if (gx - 2 * lane >= n1) return;  // whole warp past the end of the row
  const int z0 = S.zlo + blockIdx.z * MZC, z1 = min(z0 + MZC, S.zhi);
  const bool valid = gx < n1;       // n1 is even: both points or none
  const double* cp = rho + wrapx(gx, n1) + (size_t)n1 * gy;  // periodic images past the end keep the shuffles right
  const bool edge = lane == 0 || lane == 31;
  const double* ep = rho + wrapx(lane == 0 ? gx - 1 : gx + 2, n1) + (size_t)n1 * gy;
  const size_t wrapback = s3 * (size_t)n3;
  int wz = wrapx(z0 - 1, n3);
  cp += s3 * wz; ep += s3 * wz;
  auto next_plane = [&]() {
    cp += s3; ep += s3;
    if (++wz == n3) { wz = 0; cp -= wrapback; ep -= wrapback; }
  };
  double2 a[MZP + 2];
  double e[MZP + 2];
#pragma unroll
  for (int k = 0; k < MZP + 2; k++) {
    const bool need = z0 - 1 + k <= z1;
    a[k] = need ? __ldg(reinterpret_cast<const double2*>(cp)) : make_double2(0.0, 0.0);
    e[k] = (need && edge) ? __ldg(ep) : 0.0;
    if (need) next_plane();
  }
  for (int zb = z0; zb < z1; zb += MZP) {
    double2 an[MZP];
    double en[MZP];
#pragma unroll
    for (int k = 0; k < MZP; k++) {
      const bool need = zb + MZP + 1 + k <= z1;
      an[k] = need ? __ldg(reinterpret_cast<const double2*>(cp)) : make_double2(0.0, 0.0);
      en[k] = (need && edge) ? __ldg(ep) : 0.0;
      if (need) next_plane();
    }
#pragma unroll
    for (int k = 0; k < MZP; k++) {
      const int iz = zb + k;
      if (iz < z1) {  // warp-uniform
        const double2 vm = a[k], vc = a[k + 1], vp = a[k + 2];
        double xl = __shfl_up_sync(FULL, vc.y, 1), xr = __shfl_down_sync(FULL, vc.x, 1);
        if (lane == 0) xl = e[k + 1];
        if (lane == 31) xr = e[k + 1];
        const bool c0 = vc.x >= xl && vc.x >= vc.y && vc.x >= vm.x && vc.x >= vp.x;
        const bool c1 = vc.y >= vc.x && vc.y >= xr && vc.y >= vm.y && vc.y >= vp.y;
        if (valid && (c0 || c1)) {
#pragma unroll 1
          for (int q = 0; q < 2; q++) {
            if (!(q ? c1 : c0)) continue;
            const int px = gx + q;
            if (dev_is_max(n1, n2, n3, rho, px, gy, iz)) {
              const int slot = atomicAdd(ncand, 1);
              if (slot < maxcand) cand[slot] = px + n1 * (gy + n2 * iz);
              for (int i = 0; i < CF.nlev; i++)
                CF.p[i][(px >> (i + 1)) + CF.c1[i] * ((gy >> (i + 1)) + (size_t)CF.c2[i] * ((iz - S.zlo) >> (i + 1)))] = 1;
            }
          }
        }
      }
    }
    a[0] = a[MZP]; a[1] = a[MZP + 1]; e[0] = e[MZP]; e[1] = e[MZP + 1];
#pragma unroll
    for (int k = 0; k < MZP; k++) { a[k + 2] = an[k]; e[k + 2] = en[k]; }
  }
}

// ------------------------------------------------------------------------------------------------
// walkers: persistent warps with lane refill.  A warp takes batches of work items from a global cursor;
// a lane whose trajectory has ended picks up the next item while the other lanes keep stepping, so the
// long trajectories that run along an interatomic surface do not idle the rest of the warp.
// `label_g` is the owned part of the label buffer re-based so that label_g[global id] is valid for every
// owned point.
// ------------------------------------------------------------------------------------------------
struct WalkArgs {
  const double* rho;
  int* label_g;
  MaxHash h;
  SafeMap sm;
  unsigned char* reached;
  const int* list;          // nullptr: lattice mode; otherwise the dense list of every walker so far
  long long flat_base;      // flat mode: items are list[flat_base .. flat_base + count)
  long long count;
  const int* count_dev;     // k_walk2 / k_walk3: non-null = the number of entries is read from device memory (no host round trip)
  const int* base_dev;      // k_walk3: non-null = flat_base is read from device memory as well
  int* stop;                // stop log parallel to `list`: where a walk was cut short (stop code), or -1
  int sm_level;             // level index of `sm` (goes into the stop code)
  SafeMap maps[MAXLEV];     // FIX: every early-termination map in use, for invalidation
  int nmaps;
  int* ninval;
  const int2* items;        // non-null: work items (first list index, count) cut from a segmented list
  int nitems;               // upper bound; the exact number is *nitems_dev
  const int* nitems_dev;
  int lat_s, lat_m1, lat_m2;  // lattice mode: item t -> (lx, ly, lz) -> start = lx*s + n1*(ly*s + n2*(zlo + lz*s))
  Slab S;
  unsigned long long* cursor;
  int batch;
  int refill_min;           // idle lanes that trigger a refill
  int steps_per_check;      // walker steps between two looks at the work queue
  int2* overflow; int* noverflow; int overcap;  // (start, index in the dense walker list or -1)
  int* err;
  unsigned long long* nsteps;
  int* next; int* nnext; int nextcap;  // FIX: points whose neighbourhood became non-uniform
};

// FIX: `start` was a filled point; if its own trajectory ends elsewhere, its filled 26-neighbours now have
// a neighbour with a different label and are queued for the next pass (claimed by clearing FILLBIT).
__device__ __noinline__ void claim_neighbours(const BaderParams& P, const WalkArgs& A, int start) {
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int x = start % n1, y = (start / n1) % n2, z = start / (n1 * n2);
  // the label of `start` changed: every certificate whose region contains it is void
  for (int m = 0; m < A.nmaps; m++) {
    const SafeMap& sm = A.maps[m];
    if (!sm.safe) continue;
    const int cz = z - sm.zlo;
    if (cz < 0 || cz >= sm.nzl) continue;
    // entries whose certified region contains the point.  Octet: vertices c, c+1 (3x3x3 of the points they serve;
    // for cubes of stride >= 4 that is their 5x5x5 as well); packed octets of the stride-2 cubes: every vertex that
    // serves a point within 2 cells, (p-1)>>1 .. (p+3)>>1; cube certificate: cubes c-1, c, c+1
    int cx = x >> sm.shift, cy = y >> sm.shift, cl = cz >> sm.shift, lo = sm.octet ? 0 : -1, hi = 1;
    if (sm.pack) { cx = (x - 1) >> 1; cy = (y - 1) >> 1; cl = (cz - 1) >> 1; lo = 0; hi = 2; }
    for (int dz = lo; dz <= hi; dz++) {
      int vz = cl + dz;
      if (vz >= sm.c3) { if (sm.octet && sm.wrapz) vz -= sm.c3; else continue; }
      if (vz < 0) { if (sm.octet && sm.wrapz) vz += sm.c3; else continue; }
      for (int dy = lo; dy <= hi; dy++) {
        const int vy = wrapx(cy + dy, sm.c2);
        for (int dx = lo; dx <= hi; dx++) {
          const int vx = wrapx(cx + dx, sm.c1);
          int* e = sm.safe + vx + sm.c1 * (vy + (size_t)sm.c2 * vz);
          if (*e >= 0 && atomicExch(e, sm.pack ? CERT_VOID : -1) >= 0) atomicAdd(A.ninval, 1);
        }
      }
    }
  }
  for (int dz = -1; dz <= 1; dz++) {
    int zz = z + dz;
    if (A.S.periodic) zz = wrapx(zz, n3);
    else if (zz < A.S.zlo || zz >= A.S.zhi) continue;  // the owner sees the change through its halo plane
    for (int dy = -1; dy <= 1; dy++) {
      const int yy = wrapx(y + dy, n2);
      for (int dx = -1; dx <= 1; dx++) {
        const int q = wrapx(x + dx, n1) + n1 * (yy + n2 * zz);
        if (q == start) continue;
        if ((unsigned)A.label_g[q] & FILLBIT) {
          const unsigned old = (unsigned)atomicAnd(A.label_g + q, LMASK);
          if (old & FILLBIT) {
            const int slot = atomicAdd(A.nnext, 1);
            if (slot < A.nextcap) A.next[slot] = q;
            else atomicExch(A.err, 4);
          }
        }
      }
    }
  }
}

// oldlab (FIX only): the label `start` carried when its walk began (read together with the first loads of the
// walk; nobody else writes the label bits of a queued point during a pass)
template <bool FIX>
__device__ __forceinline__ void walk_finish(const BaderParams& P, const WalkArgs& A, int start, int st, int out, int tidx,
                                            int sli, int oldlab) {
  if (A.stop && tidx >= 0) A.stop[tidx] = (st == 2) ? ((A.sm_level << STOP_SHIFT) | sli) : -1;
  if (st == 3) {
    const int slot = atomicAdd(A.noverflow, 1);
    if (slot < A.overcap) A.overflow[slot] = make_int2(start, tidx);
    else atomicExch(A.err, 3);
    return;
  }
  int lab = out;
  if (st == 1) {
    lab = hash_lookup(A.h, out);
    if (lab < 0) { atomicExch(A.err, 1); return; }  // terminal is not a candidate maximum: cannot happen
    if (!A.reached[lab]) A.reached[lab] = 1;
  }
  A.label_g[start] = lab;
  if (FIX && oldlab != lab) claim_neighbours(P, A, start);
}

__device__ __noinline__ int walk_item_lattice(int n1, int n2, int zlo, int lat_s, int lat_m1, int lat_m2, int t) {
  const int lx = t % lat_m1, ly = (t / lat_m1) % lat_m2, lz = t / (lat_m1 * lat_m2);
  return lx * lat_s + n1 * (ly * lat_s + n2 * (zlo + lz * lat_s));
}
__device__ __forceinline__ int walk_item(const BaderParams& P, const WalkArgs& A, int t) {
  if (A.list) return A.list[t];
  const int lx = (int)(t % A.lat_m1), ly = (int)((t / A.lat_m1) % A.lat_m2), lz = (int)(t / ((long long)A.lat_m1 * A.lat_m2));
  return lx * A.lat_s + P.n1 * (ly * A.lat_s + P.n2 * (A.S.zlo + lz * A.lat_s));
}

constexpr int STEPS_PER_CHECK = 3;
constexpr int REFILL_MIN = 20;  // idle lanes that trigger a refill: high, so that the lanes of a warp stay in step and their loads coalesce
// STATS: count the walker steps (diagnostics, C2G_BADER_VERBOSE); off in production, it costs registers
template <bool ORTHO, bool FIX, bool STATS>
__global__ void __launch_bounds__(256, 4) k_walk(const __grid_constant__ BaderParams P, const __grid_constant__ WalkArgs A) {
  const int lane = threadIdx.x & 31;
  int qpos = 0, qend = 0;  // warp-uniform: items [qpos, qend) are this warp's (indices < 2^31: the lists hold at most nn entries)
  bool done = false, active = false;
  int start = 0, sl = -1, sli = -1, oldlab = 0;
  int tidx = -1;
  WState w;
  Nb nb;
  unsigned long long steps = 0;
  int pend_st = 0, pend_out = 0;  // a finished walk waits here until the next look at the queue: one pass for all lanes
  for (;;) {
    if (pend_st) {
      walk_finish<FIX>(P, A, start, pend_st, pend_out, A.list ? tidx : -1, sli, oldlab);
      pend_st = 0;
    }
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle == FULL || (!done && __popc(idle) >= A.refill_min)) {
      if (qpos >= qend && !done) {
        if (A.items) {  // next work item: a spatially compact group of start points
          unsigned long long b = 0;
          if (lane == 0) b = atomicAdd(A.cursor, 1ull);
          b = __shfl_sync(FULL, b, 0);
          if (b >= (unsigned long long)__ldg(A.nitems_dev)) done = true;
          else {
            const int2 it = __ldg(A.items + b);
            qpos = it.x; qend = qpos + it.y;
          }
        } else {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(A.cursor, (unsigned long long)A.batch);
          base = __shfl_sync(FULL, base, 0);
          if ((long long)base >= A.count) done = true;
          else { qpos = (int)(A.flat_base + (long long)base); qend = (int)(A.flat_base + min((long long)base + A.batch, A.count)); }
        }
      }
      const int navail = qend - qpos;
      if (navail > 0) {
        const int rank = __popc(idle & ((1u << lane) - 1u));
        if (!active && rank < navail) {
          tidx = qpos + rank;
          start = walk_item(P, A, tidx);
          walk_init(P, A.rho, w, start);
          load_nb(P, A.rho, start, w.x, w.y, w.z, nb);
          if (FIX) oldlab = A.label_g[start] & LMASK;
          sl = -1;
          active = true;
        }
        qpos += min(navail, __popc(idle));
      } else if (done && idle == FULL) break;
    }
    // a few steps between two looks at the work queue (the test above costs ~45 issue slots)
#pragma unroll 1
    for (int k = 0; k < A.steps_per_check; k++) {
      if (active) {
        int out = 0;
        const int st = walk_step_pipe<ORTHO>(P, A.rho, A.h, A.sm, w, nb, sl, sli, out);
        if (st) {
          if (STATS) steps += (unsigned)w.len;
          pend_st = st; pend_out = out;
          active = false;
        }
      }
    }
  }
  if (STATS && A.nsteps && steps) atomicAdd(A.nsteps, steps);
}

// ------------------------------------------------------------------------------------------------
// k_walk2: the production walkers.  Same trajectories, same decisions, same results as k_walk; what differs is
// the schedule.  A step keeps NO state besides (point, dr, largest density on the path): all loads of a step
// (density of the point, its 6 neighbours, its certificate) are issued together at the top of the step, so that
// starting a new trajectory costs a handful of instructions.  That makes it cheap to refill EVERY idle lane
// before EVERY step: the warp holds the next 32 start points in a register (one coalesced load, a second chunk
// is already in flight), and idle lanes take theirs with one shuffle.  With k_walk's refill thresholds
// (20 / 24 idle lanes) 19 of 32 lanes were active on the last level.
// Work distribution: chunks of 32 consecutive entries of the dense, spatially ordered walker list (or of the
// lattice numbering), handed out through a global cursor.
// The tests of step k+1 that k_walk made at the end of step k are made at the top of step k+1, in the same
// order: possible revisit (density not above the path's maximum: hand over to k_walk_big), then the certificate.
// ------------------------------------------------------------------------------------------------
// nint(v) for a finite |v| < 1.5 as a double (returned) and as an int (d), on the integer ALU: |v| >= 0.5 iff the high
// word of |v| is >= that of 0.5.  NaN and infinities give 0 (such a walk ends "did not move", like with nint_di).
__device__ __forceinline__ double nint_bits(double v, int& d) {
  const int hi = __double2hiint(v);
  const bool big = (unsigned)((hi & 0x7fffffff) - 0x3fe00000) < (unsigned)(0x7ff00000 - 0x3fe00000);
  const int ahi = big ? ((hi & (int)0x80000000) | 0x3ff00000) : 0;
  d = big ? ((hi >> 31) | 1) : 0;
  return __hiloint2double(ahi, 0);
}
// Fortran nint of a finite |v| < 2.5 with |nint| <= 1 expected (|v| < 1.5), as a double: one fp64 compare, the sign of v
// pasted onto 1.0, one select.  NaN gives 0 (such a walk ends "did not move", like with nint_di).
__device__ __forceinline__ double nint_half(double v) {
  const int hi = __double2hiint(v);
  const int ahi = fabs(v) >= 0.5 ? ((hi & (int)0x80000000) | 0x3ff00000) : 0;
  return __hiloint2double(ahi, 0);
}
// g, or 0 when both neighbours are strictly below the centre (rho_grad_dir, :553-558): two chained predicates, one select
__device__ __forceinline__ double zero_if_both_less(double g, double p, double m, double r0) {
  double out;
  asm("{\n\t.reg .pred q;\n\tsetp.lt.f64 q, %2, %4;\n\tsetp.lt.and.f64 q, %3, %4, q;\n\t"
      "selp.f64 %0, 0d0000000000000000, %1, q;\n\t}"
      : "=d"(out) : "d"(g), "d"(p), "d"(m), "d"(r0));
  return out;
}
template <bool ORTHO, bool FIX, bool STATS>
__global__ void __launch_bounds__(256, 4) k_walk2(const __grid_constant__ BaderParams P, const __grid_constant__ WalkArgs A) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int s2 = n1, s3 = n1 * n2;
  const int total = A.count_dev ? __ldg(A.count_dev) : (int)A.count;  // < 2^31: the lists hold at most nn entries
  const int nchunk = (total + 31) >> 5;
  const int* const list = A.list ? A.list + A.flat_base : nullptr;
  const int tbase = (int)A.flat_base;
  unsigned* const cur32 = reinterpret_cast<unsigned*>(A.cursor);  // low word of the (zeroed) 64-bit cursor
  auto grab = [&]() -> int {  // chunk ids beyond nchunk are clamped: at most one surplus grab per warp and rotation
    unsigned b = 0;
    if (lane == 0) b = atomicAdd(cur32, 1u);
    return (int)min(__shfl_sync(FULL, b, 0), (unsigned)nchunk);
  };
  auto load_chunk = [&](int c) -> int {  // start point of entry 32 c + lane
    const int t = (c << 5) + lane;
    if (t >= total) return -1;
    return list ? __ldg(list + t) : walk_item_lattice(P.n1, P.n2, A.S.zlo, A.lat_s, A.lat_m1, A.lat_m2, t);
  };
  int c0 = grab();
  int pre = load_chunk(c0);
  int c1 = grab();
  int pre2 = load_chunk(c1);
  int navail = min(32, total - (c0 << 5));  // entries of `pre` (<= 0: none)
  int consumed = 0;                                                    // ... already handed out
  bool active = false, first = true;
  int id = 0, x = 0, y = 0, z = 0, start = 0, tidx = -1, oldlab = 0;
  double dr0 = 0.0, dr1 = 0.0, dr2 = 0.0, rhomax = 0.0;
  int pend_st = 0, pend_out = 0, pend_sli = -1;
  unsigned long long steps = 0;
  for (;;) {
    if (pend_st) {  // one pass for all lanes whose walk ended in the previous step
      walk_finish<FIX>(P, A, start, pend_st, pend_out, list ? tidx : -1, pend_sli, oldlab);
      pend_st = 0;
    }
    const unsigned idle = __ballot_sync(FULL, !active);
    if (__popc(idle) >= A.refill_min) {
      if (consumed == navail && c0 < nchunk) {  // next chunk (already loaded); fetch the one after it
        c0 = c1; pre = pre2; consumed = 0;
        navail = min(32, total - (c0 << 5));
        c1 = grab();
        pre2 = load_chunk(c1);
      }
      const int take = min(__popc(idle), navail - consumed);
      if (take > 0) {
        const int r = __popc(idle & lt);
        const int s = __shfl_sync(FULL, pre, (consumed + r) & 31);
        if (!active && r < take) {
          start = id = s;
          tidx = tbase + (c0 << 5) + consumed + r;
          z = fastdiv(s, P.mg12, P.sh12);
          const int rem = s - z * s3;
          y = fastdiv(rem, P.mg1, P.sh1);
          x = rem - y * n1;
          dr0 = dr1 = dr2 = 0.0;
          rhomax = __longlong_as_double((long long)0xfff0000000000000ull);  // -inf: the start point is never a revisit
          first = true;
          if (FIX) oldlab = A.label_g[s] & LMASK;
          active = true;
        }
        consumed += take;
      } else if (idle == FULL) {
        break;  // nothing in flight and nothing left to hand out (c0 >= nchunk: the cursor only grows)
      }
    }
    if (active) {
      if (STATS) steps++;
      const double* c = A.rho + id;
      // the point and its next point stay clear of the periodic seams: no wrapping anywhere in this step
      const bool inner = (unsigned)(x - 3) < (unsigned)P.in1 && (unsigned)(y - 3) < (unsigned)P.in2 &&
                         (unsigned)(z - 3) < (unsigned)P.in3;
      double r0, xp, xm, yp, ym, zp, zm;
      r0 = __ldg(c);
      if (inner) {
        xp = __ldg(c + 1); xm = __ldg(c - 1);
        yp = __ldg(c + s2); ym = __ldg(c - s2);
        zp = __ldg(c + s3); zm = __ldg(c - s3);
      } else {
        xp = __ldg(c + ((x + 1 == n1) ? 1 - n1 : 1));
        xm = __ldg(c + ((x == 0) ? n1 - 1 : -1));
        yp = __ldg(c + ((y + 1 == n2) ? s2 - s3 : s2));
        ym = __ldg(c + ((y == 0) ? s3 - s2 : -s2));
        zp = __ldg(c + ((z + 1 == n3) ? s3 - s3 * n3 : s3));
        zm = __ldg(c + ((z == 0) ? s3 * n3 - s3 : -s3));
      }
      int sli = -1, sl = -1;
      if (!first && A.sm.safe) sl = safe_lookup2(A.sm, x, y, z, sli);  // the start point is never looked up
      int st = 0, out = 0;
      if (r0 <= rhomax) {
        st = 3;  // possibly a point of this path (known(pm) == 1, :487): k_walk_big answers exactly
      } else if (sl >= 0) {
        st = 2; out = sl;  // quit at a known interior point (:447)
      } else {
        // rho_grad_dir (:532-567)
        const double gl0 = zero_if_both_less((xp - xm) * 0.5, xp, xm, r0);
        const double gl1 = zero_if_both_less((yp - ym) * 0.5, yp, ym, r0);
        const double gl2 = zero_if_both_less((zp - zm) * 0.5, zp, zm, r0);
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
        double gmax = fabs(g0);
        {
          const double t1 = fabs(g1), t2 = fabs(g2);
          gmax = t1 > gmax ? t1 : gmax;
          gmax = t2 > gmax ? t2 : gmax;
        }
        const int oid = id;
        if (gmax < 1e-30) {  // (:468-476)
          dr0 = dr1 = dr2 = 0.0;
          if (hash_lookup(A.h, id) >= 0) {
            st = 1; out = id;
          } else {
            id = dev_step_ongrid(P, A.rho, x, y, z, r0);
            z = fastdiv(id, P.mg12, P.sh12);
            const int rem = id - z * s3;
            y = fastdiv(rem, P.mg1, P.sh1);
            x = rem - y * n1;
            if (id == oid) st = 3;
          }
        } else {  // (:477-483)
          const double coeff = 1.0 / gmax;
          g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
          int d0, d1, d2, e0, e1, e2;
          const double a0 = nint_bits(g0, d0), a1 = nint_bits(g1, d1), a2 = nint_bits(g2, d2);
          const double t0 = dr0 + g0 - a0, t1 = dr1 + g1 - a1, t2 = dr2 + g2 - a2;
          const double b0 = nint_bits(t0, e0), b1 = nint_bits(t1, e1), b2 = nint_bits(t2, e2);
          dr0 = t0 - b0; dr1 = t1 - b1; dr2 = t2 - b2;
          d0 += e0; d1 += e1; d2 += e2;
          if (inner) {
            x += d0; y += d1; z += d2;
            id += d0 + n1 * (d1 + n2 * d2);
          } else {
            x = wrap2(x + d0, n1);
            y = wrap2(y + d1, n2);
            z = wrap2(z + d2, n3);
            id = x + n1 * (y + n2 * z);
          }
          // did not move although the gradient is not zero: k_walk_big (k_walk hands these over as well)
          if (id == oid) st = 3;
        }
        // known(p) = 1 (:484); harmless when the walk has just ended
        rhomax = r0 > rhomax ? r0 : rhomax;
        first = false;
      }
      if (st) {
        pend_st = st; pend_out = out; pend_sli = sli;
        active = false;
      }
    }
  }
  if (STATS && A.nsteps && steps) atomicAdd(A.nsteps, steps);
}

// ------------------------------------------------------------------------------------------------
// k_walk3: block-cooperative walkers with stable compaction.  Same trajectories, same decisions, same results as
// k_walk; what differs is how lanes are kept busy.  k_walk refills a warp only when 20 (24) of its lanes are idle,
// because refilling single lanes mixes walkers of different age in one warp and their loads stop coalescing: 19 of
// 32 lanes were active on the last level.  Here the 256 threads of a block take `steps_per_check` steps, then --
// once `refill_min` threads are idle -- the surviving walkers are packed, IN THEIR ORDER, into the first threads of
// the block through shared memory, and the free threads at the end take the next consecutive entries of the dense,
// spatially ordered walker list.  A warp therefore always holds walkers that started next to each other at about the
// same time (a cohort, or the packed survivors of neighbouring cohorts), and whole warps -- not scattered lanes --
// run dry at the end of a launch.  The step is the lean one of k_walk2 (all loads of a step issued together at its
// top, no state besides point, dr and the largest density on the path).
// ------------------------------------------------------------------------------------------------
// CERT: 0 = no certificates (top lattice), 1 = any SafeMap, 2 = the packed octets of the stride-2 level
#ifndef C2G_W3_NT
#define C2G_W3_NT 256   // threads per block of k_walk3 (512: measured)
#endif
#ifndef C2G_W3_MINB
#define C2G_W3_MINB (1024 / C2G_W3_NT)   // resident blocks per SM the register allocation aims at
#endif
constexpr int W3_NT = C2G_W3_NT;
template <bool ORTHO, bool FIX, bool STATS, int CERT>
__global__ void __launch_bounds__(W3_NT, C2G_W3_MINB) k_walk3(const __grid_constant__ BaderParams P, const __grid_constant__ WalkArgs A) {
  __shared__ int s_wc[W3_NT / 32];
  __shared__ int s_chunk[2];                       // first entry and number of entries handed to the block
  __shared__ int s_id[W3_NT], s_start[W3_NT], s_tidx[W3_NT], s_old[W3_NT];
  __shared__ double s_dr[3][W3_NT], s_rm[W3_NT];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const int s2 = n1, s3 = n1 * n2;
  const int total = A.count_dev ? __ldg(A.count_dev) : (int)A.count;  // < 2^31: the lists hold at most nn entries
  const int tbase = A.base_dev ? __ldg(A.base_dev) : (int)A.flat_base;
  const int* const list = A.list ? A.list + tbase : nullptr;
  unsigned* const cur32 = reinterpret_cast<unsigned*>(A.cursor);  // low word of the (zeroed) 64-bit cursor
  bool active = false, first = true, exhausted = false;           // exhausted: block-uniform
  int id = 0, x = 0, y = 0, z = 0, start = 0, tidx = -1, oldlab = 0;
  double dr0 = 0.0, dr1 = 0.0, dr2 = 0.0, rhomax = 0.0;
  unsigned long long steps = 0;
  for (;;) {
    // ---- pack the survivors, refill the free threads ----
    const int nact = __syncthreads_count(active);
    if (nact == 0 && exhausted) break;
    if (W3_NT - nact >= A.refill_min || nact == 0) {
      const unsigned m = __ballot_sync(FULL, active);
      if (lane == 0) s_wc[wid] = __popc(m);
      if (tid == 0) {
        int base = 0, n = 0;
        if (!exhausted) {
          const int need = W3_NT - nact;
          const unsigned b = atomicAdd(cur32, (unsigned)need);
          base = (int)min(b, (unsigned)total);
          n = min(need, total - base);
        }
        s_chunk[0] = base; s_chunk[1] = n;
      }
      __syncthreads();
      if (active) {
        int slot = __popc(m & ((1u << lane) - 1u));
        for (int w = 0; w < wid; w++) slot += s_wc[w];
        s_id[slot] = id; s_start[slot] = start; s_tidx[slot] = tidx;
        if (FIX) s_old[slot] = oldlab;
        s_dr[0][slot] = dr0; s_dr[1][slot] = dr1; s_dr[2][slot] = dr2; s_rm[slot] = rhomax;
      }
      __syncthreads();
      const int cbase = s_chunk[0], cn = s_chunk[1];
      if (cn < W3_NT - nact) exhausted = true;
      active = false;
      int s = -1;
      if (tid < nact) {
        s = s_id[tid]; start = s_start[tid]; tidx = s_tidx[tid];
        if (FIX) oldlab = s_old[tid];
        dr0 = s_dr[0][tid]; dr1 = s_dr[1][tid]; dr2 = s_dr[2][tid]; rhomax = s_rm[tid];
        first = false;
        active = true;
      } else if (tid - nact < cn) {
        const int t = cbase + tid - nact;
        s = list ? __ldg(list + t) : walk_item_lattice(P.n1, P.n2, A.S.zlo, A.lat_s, A.lat_m1, A.lat_m2, t);
        start = s; tidx = tbase + t;
        dr0 = dr1 = dr2 = 0.0;
        rhomax = __longlong_as_double((long long)0xfff0000000000000ull);  // -inf: the start point is never a revisit
        first = true;
        if (FIX) oldlab = A.label_g[s] & LMASK;
        active = true;
      }
      if (active) {
        id = s;
        z = fastdiv(s, P.mg12, P.sh12);
        const int rem = s - z * s3;
        y = fastdiv(rem, P.mg1, P.sh1);
        x = rem - y * n1;
      }
      // the shared arrays are written again only after the next __syncthreads_count
    }
    // ---- a few steps ----
    int pend_st = 0, pend_out = 0, pend_sli = -1;
#pragma unroll 1
    for (int k = 0; k < A.steps_per_check; k++) {
      if (!active) break;
      if (STATS) steps++;
      const double* c = A.rho + id;
      // the point and its next point stay clear of the periodic seams: no wrapping anywhere in this step
      const bool inner = (unsigned)(x - 3) < (unsigned)P.in1 && (unsigned)(y - 3) < (unsigned)P.in2 &&
                         (unsigned)(z - 3) < (unsigned)P.in3;
      double r0, xp, xm, yp, ym, zp, zm;
      r0 = __ldg(c);
      if (inner) {
        xp = __ldg(c + 1); xm = __ldg(c - 1);
        yp = __ldg(c + s2); ym = __ldg(c - s2);
        zp = __ldg(c + s3); zm = __ldg(c - s3);
      } else {
        xp = __ldg(c + ((x + 1 == n1) ? 1 - n1 : 1));
        xm = __ldg(c + ((x == 0) ? n1 - 1 : -1));
        yp = __ldg(c + ((y + 1 == n2) ? s2 - s3 : s2));
        ym = __ldg(c + ((y == 0) ? s3 - s2 : -s2));
        zp = __ldg(c + ((z + 1 == n3) ? s3 - s3 * n3 : s3));
        zm = __ldg(c + ((z == 0) ? s3 * n3 - s3 : -s3));
      }
      int sli = -1, sl = -1;
      if (CERT == 2) {
        if (!first) {  // the start point is never looked up
          const SafeMap& sm = A.sm;
          const int cz = z - sm.zlo;
          int vx = (x + 1) >> 1, vy = (y + 1) >> 1, vz = (cz + 1) >> 1;
          bool ok = true;
          if (!(inner && A.S.periodic)) {  // near a seam, or a z-slab: the vertex may wrap or lie outside the owned layers
            vx = vx == sm.c1 ? 0 : vx;
            vy = vy == sm.c2 ? 0 : vy;
            const bool top = vz == sm.c3;
            ok = (unsigned)cz < (unsigned)sm.nzl && (!top || sm.wrapz);
            vz = top ? 0 : vz;
          }
          sli = vx + sm.c1 * (vy + sm.c2 * vz);
          const int e = ok ? sm.safe[sli] : CERT_VOID;  // plain load: entries may be voided while walkers run
          sl = cert_unpack(e, x, y, cz);
        }
      } else if (CERT == 1) {
        if (!first) sl = safe_lookup2(A.sm, x, y, z, sli);
      }
      int st = 0, out = 0;
      if (r0 <= rhomax) {
        st = 3;  // possibly a point of this path (known(pm) == 1, :487): k_walk_big answers exactly
      } else if (sl >= 0) {
        st = 2; out = sl;  // quit at a known interior point (:447)
      } else {
        // rho_grad_dir (:532-567)
        const double gl0 = zero_if_both_less((xp - xm) * 0.5, xp, xm, r0);
        const double gl1 = zero_if_both_less((yp - ym) * 0.5, yp, ym, r0);
        const double gl2 = zero_if_both_less((zp - zm) * 0.5, zp, zm, r0);
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
        // maxval(abs(grad)): the component of largest magnitude is selected as it is, its sign goes at the end
        double gsel = fabs(g1) > fabs(g0) ? g1 : g0;
        gsel = fabs(g2) > fabs(gsel) ? g2 : gsel;
        const double gmax = fabs(gsel);
        const int oid = id;
        if (gmax < 1e-30) {  // (:468-476)
          dr0 = dr1 = dr2 = 0.0;
          if (hash_lookup(A.h, id) >= 0) {
            st = 1; out = id;
          } else {
            id = dev_step_ongrid(P, A.rho, x, y, z, r0);
            z = fastdiv(id, P.mg12, P.sh12);
            const int rem = id - z * s3;
            y = fastdiv(rem, P.mg1, P.sh1);
            x = rem - y * n1;
            if (id == oid) st = 3;
          }
        } else {  // (:477-483)
          const double coeff = 1.0 / gmax;
          g0 = coeff * g0; g1 = coeff * g1; g2 = coeff * g2;
          // pm = p + nint(g) + nint(dr + g - nint(g)); the sum of the two nints is exact (0, +-1, +-2) and converted once
          const double a0 = nint_half(g0), a1 = nint_half(g1), a2 = nint_half(g2);
          const double t0 = dr0 + g0 - a0, t1 = dr1 + g1 - a1, t2 = dr2 + g2 - a2;
          const double b0 = nint_half(t0), b1 = nint_half(t1), b2 = nint_half(t2);
          dr0 = t0 - b0; dr1 = t1 - b1; dr2 = t2 - b2;
          const int d0 = __double2int_rn(a0 + b0), d1 = __double2int_rn(a1 + b1), d2 = __double2int_rn(a2 + b2);
          if (inner) {
            x += d0; y += d1; z += d2;
            id += d0 + n1 * (d1 + n2 * d2);
          } else {
            x = wrap2(x + d0, n1);
            y = wrap2(y + d1, n2);
            z = wrap2(z + d2, n3);
            id = x + n1 * (y + n2 * z);
          }
          // did not move although the gradient is not zero: k_walk_big (k_walk hands these over as well)
          if (id == oid) st = 3;
        }
        rhomax = r0;  // known(p) = 1 (:484): r0 > rhomax here, so the largest density on the path is r0
        first = false;
      }
      if (st) {
        pend_st = st; pend_out = out; pend_sli = sli;
        active = false;
      }
    }
    if (pend_st) walk_finish<FIX>(P, A, start, pend_st, pend_out, list ? tidx : -1, pend_sli, oldlab);
  }
  if (STATS && A.nsteps && steps) atomicAdd(A.nsteps, steps);
}

// cut the non-empty segments of a segmented list into work items of at most `batch` entries, IN SEGMENT
// ORDER (segments are numbered super-block by super-block, and that order is what keeps the walkers that are
// in flight together inside one L2-sized region).  Three small launches: items per 256-segment block,
// exclusive scan of the block sums, write.
__device__ __forceinline__ int block_excl_scan_256(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int base = 0;
  total = 0;
  for (int w = 0; w < 8; w++) {
    if (w < wid) base += s_warp[w];
    total += s_warp[w];
  }
  __syncthreads();
  return base + incl - v;
}
__global__ void __launch_bounds__(256) k_items_count(int nseg, int batch, const int* __restrict__ segcnt, int2* __restrict__ blocksum) {
  __shared__ int s_warp[8];
  const int b = blockIdx.x * 256 + threadIdx.x;
  const int c = b < nseg ? segcnt[b] : 0;
  int ti, te;
  block_excl_scan_256((c + batch - 1) / batch, s_warp, ti);
  block_excl_scan_256(c, s_warp, te);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = make_int2(ti, te);
}
// doff (optional): entries of the dense walker list in use, in device memory; a list that would overflow (capacity dcap)
// raises err = 5 and hands out nothing
__global__ void __launch_bounds__(256) k_items_scan(int nblk, int2* __restrict__ blocksum, int* __restrict__ nitems, int* __restrict__ nentries,
                                                    const int* __restrict__ doff, long long dcap, int* __restrict__ err) {
  __shared__ int s_warp[8];
  int ci = 0, ce = 0;
  for (int i0 = 0; i0 < nblk; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const int2 v = i < nblk ? blocksum[i] : make_int2(0, 0);
    int ti, te;
    const int ei = block_excl_scan_256(v.x, s_warp, ti);
    const int ee = block_excl_scan_256(v.y, s_warp, te);
    if (i < nblk) blocksum[i] = make_int2(ci + ei, ce + ee);
    ci += ti; ce += te;
  }
  if (threadIdx.x == 0) {
    if (doff && (long long)*doff + ce > dcap) { atomicExch(err, 5); ci = 0; ce = 0; }
    *nitems = ci; *nentries = ce;
  }
}
// writes the items and copies the entries of the segments, in segment order, into the dense list at dbase
__global__ void __launch_bounds__(256) k_items_write(int nseg, int segcap, int batch, const int* __restrict__ segcnt,
                                                     const int2* __restrict__ blocksum, const int* __restrict__ seglist,
                                                     int2* __restrict__ items, int* __restrict__ dlist, long long dbase,
                                                     const int* __restrict__ doff, const int* __restrict__ nentries) {
  __shared__ int s_warp[8];
  __shared__ int s_c[256], s_e[256];
  if (doff) {  // offsets in device memory (no host round trip): nothing to do when nothing was queued or the list is full
    if (*nentries == 0) return;
    dbase = *doff;
  }
  const int b = blockIdx.x * 256 + threadIdx.x;
  const int c = b < nseg ? segcnt[b] : 0;
  const int k = (c + batch - 1) / batch;
  int total;
  const int2 bs = blocksum[blockIdx.x];
  const int ibase = bs.x + block_excl_scan_256(k, s_warp, total);
  const int ebase = bs.y + block_excl_scan_256(c, s_warp, total);
  if (items)  // k_walk3 walks the dense list itself and needs no items
    for (int j = 0; j < k; j++) items[ibase + j] = make_int2((int)(dbase + ebase + j * batch), min(batch, c - j * batch));
  s_c[threadIdx.x] = c; s_e[threadIdx.x] = ebase;
  __syncthreads();
  if (total == 0) return;
  // coalesced copy of the segments' entries, one warp per segment (most segments are empty or hold a few dozen entries:
  // with the whole block on every segment the loop was a chain of 256 dependent rounds, 0.13-0.22 ms per level)
  const int lane = threadIdx.x & 31;
  for (int sg = threadIdx.x >> 5; sg < 256; sg += 8) {
    const int cs = s_c[sg];
    if (cs == 0) continue;
    const int* src = seglist + (size_t)(blockIdx.x * 256 + sg) * segcap;
    int* dst = dlist + dbase + s_e[sg];
    for (int e = lane; e < cs; e += 32) dst[e] = src[e];
  }
}

// ---- device-side bookkeeping of the assignment: with these the host enqueues a whole level (classification, list
// compaction, walkers, hand-over walks) without reading a single counter back ----
// counter layout (ints); [0..12] as documented in c2g_bader_assign
constexpr int C_NLIST = 1, C_NOVER = 2, C_ERR = 3, C_NNEXT = 7, C_NITEMS = 10, C_NINVAL = 11, C_NENT = 12, C_DOFF = 13, C_BIGCUR = 14,
              C_WALKED = 15, C_INVSEEN = 16, C_FIXPTS = 17, C_OVERTOT = 18, C_FLATN = 19, C_NPASS = 20, C_NCNT = 64;
// a flat list (edge-fix passes) -> the end of the dense list; its length is read from device memory
__global__ void __launch_bounds__(256) k_items_flat(const int* __restrict__ src, int* __restrict__ cnt, long long dcap, int* __restrict__ dlist) {
  const int count = cnt[C_FLATN], dbase = cnt[C_DOFF];
  const bool full = (long long)dbase + count > dcap;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (!full)
    for (long long e = t0; e < count; e += stride) dlist[dbase + e] = src[e];
  // the counters are read by every block above: they are written by the single-thread kernel that follows
}
__global__ void k_items_flat_end(int* __restrict__ cnt, long long dcap) {
  const int count = cnt[C_FLATN], dbase = cnt[C_DOFF];
  const bool full = (long long)dbase + count > dcap;
  if (full) atomicExch(cnt + C_ERR, 5);
  cnt[C_NENT] = full ? 0 : count;
}
// end of a group of walker launches: their entries now belong to the dense list; the hand-over list is empty again
__global__ void k_walk_end(int* __restrict__ cnt, int fix) {
  const int ne = cnt[C_NENT];
  cnt[C_DOFF] += ne;
  cnt[C_WALKED] += ne;
  if (fix) cnt[C_FIXPTS] += ne;
  cnt[C_NENT] = 0; cnt[C_NITEMS] = 0;
  cnt[C_OVERTOT] += cnt[C_NOVER];
  cnt[C_NOVER] = 0; cnt[C_BIGCUR] = 0;
  cnt[8] = 0; cnt[9] = 0;  // work cursor
}
// start of an edge-fix pass: the list built by the previous pass is consumed, a new one is started
__global__ void k_fix_begin(int* __restrict__ cnt, int first) {
  const int n = cnt[C_NNEXT];
  cnt[C_FLATN] = n;
  cnt[C_NNEXT] = 0;
  if (n + (first ? cnt[C_NLIST] : 0) > 0) cnt[C_NPASS] += 1;
}
__global__ void k_requeue_end(int* __restrict__ cnt) { cnt[C_INVSEEN] = cnt[C_NINVAL]; }

// re-queue the walkers whose early stop rested on a certificate that has been invalidated since
__global__ void __launch_bounds__(256) k_requeue(long long ntotal, const int* __restrict__ dlist, int* __restrict__ stop,
                                                 const __grid_constant__ WalkArgs A, int* __restrict__ out, int* __restrict__ nout,
                                                 int outcap, const int* __restrict__ cnt) {
  if (cnt) {  // device-side gate: nothing to do unless a certificate was voided since the last pass
    if (cnt[C_INVSEEN] == cnt[C_NINVAL]) return;
    ntotal = cnt[C_DOFF];
  }
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t0 = (long long)blockIdx.x * blockDim.x; t0 < ntotal; t0 += stride) {
    const long long t = t0 + threadIdx.x;
    bool push = false;
    if (t < ntotal) {
      const int code = stop[t];
      if (code >= 0) {
        const int lev = code >> STOP_SHIFT;
        if (lev < A.nmaps && A.maps[lev].safe && A.maps[lev].safe[code & STOP_MASK] < 0) { push = true; stop[t] = -2; }
      }
    }
    const unsigned m = __ballot_sync(FULL, push);
    if (m) {
      int base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(nout, __popc(m));
      base = __shfl_sync(FULL, base, __ffs(m) - 1);
      if (push) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < outcap) out[slot] = dlist[t];
        else atomicExch(A.err, 4);
      }
    }
  }
}

// The rare walks the persistent walkers hand over (a possible revisit, :484-488): repeated from the start with the
// whole path in global memory (bigcap entries per walker) and the reference's path search.  They use the level's
// certificates and log their early stops like every other walker of the level.
template <bool FIX>
__global__ void __launch_bounds__(64) k_walk_big(const __grid_constant__ BaderParams P, const __grid_constant__ WalkArgs A,
                                                 int count, int* __restrict__ scratch, int bigcap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int start = A.overflow[t].x, tidx = A.overflow[t].y;
  WState w;
  walk_init(P, A.rho, w, start);
  const SafeMap nosafe{nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const SafeMap& sm = (tidx >= 0 && A.stop) ? A.sm : nosafe;  // no stop log, no early stop
  const int oldlab = FIX ? (A.label_g[start] & LMASK) : 0;
  int out = 0, st, sli = -1;
  do st = walk_step<false>(P, A.rho, A.h, sm, w, scratch + (size_t)t * bigcap, bigcap, out, sli); while (st == 0);
  if (A.nsteps) atomicAdd(A.nsteps, (unsigned long long)w.len);
  if (st == 3) { atomicExch(A.err, 2); return; }
  walk_finish<FIX>(P, A, start, st, out, tidx, sli, oldlab);
}

// The same walks with the number of hand-overs read from device memory: a fixed grid takes them one by one through a
// cursor, every thread owns a path buffer of `bigcap` entries.
template <bool FIX>
__global__ void __launch_bounds__(64) k_walk_big2(const __grid_constant__ BaderParams P, const __grid_constant__ WalkArgs A,
                                                  int* __restrict__ cnt, int* __restrict__ scratch, int bigcap) {
  const int nov = min(cnt[C_NOVER], A.overcap);
  if (nov == 0) return;
  int* path = scratch + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * bigcap;
  const SafeMap nosafe{nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = atomicAdd(cnt + C_BIGCUR, 1); t < nov; t = atomicAdd(cnt + C_BIGCUR, 1)) {
    const int start = A.overflow[t].x, tidx = A.overflow[t].y;
    WState w;
    walk_init(P, A.rho, w, start);
    const SafeMap& sm = (tidx >= 0 && A.stop) ? A.sm : nosafe;  // no stop log, no early stop
    const int oldlab = FIX ? (A.label_g[start] & LMASK) : 0;
    int out = 0, st, sli = -1;
    do st = walk_step<false>(P, A.rho, A.h, sm, w, path, bigcap, out, sli); while (st == 0);
    if (A.nsteps) atomicAdd(A.nsteps, (unsigned long long)w.len);
    if (st == 3) { atomicExch(A.err, 2); continue; }
    walk_finish<FIX>(P, A, start, st, out, tidx, sli, oldlab);
  }
}

// ------------------------------------------------------------------------------------------------
// classify: one thread per owned stride-s cube, one block per tile of 16 x 4 x 4 cubes.
// uni[c] = common label of the cube's 8 corners (already labelled; the upper ones may sit on the
// halo-above plane) if they agree and the cube holds no local maximum, else -1.  The new stride-s/2 points
// of a non-uniform cube are queued for a walk; those of a uniform cube are filled (label | FILLBIT) unless
// FILL is false (last level: k_fill_edge writes them).  The queue is SEGMENTED: block b owns
// list[b*CLS_SEGCAP ..) and writes its count to segcnt[b], so that consecutive work items are spatially
// compact (the walkers of a warp then share cache lines) and no global atomic is needed.
// lbuf = label buffer including halos.
// ------------------------------------------------------------------------------------------------
constexpr int CLS_TX = 16, CLS_TY = 4, CLS_TZ = 4;
constexpr int CLS_SEGCAP = 7 * 256;
// Blocks (tiles) are numbered super-block by super-block (SB_X x SB_Y x SB_Z tiles = 128^3 points at the last
// level) so that the walkers in flight at any time work on a compact 3D region that fits the L2 cache.
constexpr int SB_X = 4, SB_Y = 16, SB_Z = 16;
__host__ __device__ inline int cls_nblocks(int t1, int t2, int t3) {
  return ((t1 + SB_X - 1) / SB_X) * ((t2 + SB_Y - 1) / SB_Y) * ((t3 + SB_Z - 1) / SB_Z) * (SB_X * SB_Y * SB_Z);
}
// Geometry of a classification pass, computed once on the host: the kernel used to derive it per thread with eight
// run-time integer divisions (~150 of its 380 instructions per warp; it is issue-bound).
struct ClsGeom {
  int c1, c2, c3;       // cubes per axis (owned layers in z)
  int t1, t2, t3;       // tiles per axis
  int s1, s12;          // super-blocks along x, and along x times y
  unsigned mg1, mg12;   // fastdiv magics of s1 and s12
  int sh1, sh12;
};
inline ClsGeom cls_geom(int n1, int n2, int nzl, int s) {
  ClsGeom G;
  G.c1 = (n1 + s - 1) / s; G.c2 = (n2 + s - 1) / s; G.c3 = (nzl + s - 1) / s;
  G.t1 = (G.c1 + CLS_TX - 1) / CLS_TX; G.t2 = (G.c2 + CLS_TY - 1) / CLS_TY; G.t3 = (G.c3 + CLS_TZ - 1) / CLS_TZ;
  G.s1 = std::max(1, (G.t1 + SB_X - 1) / SB_X);
  const int s2 = std::max(1, (G.t2 + SB_Y - 1) / SB_Y);
  G.s12 = G.s1 * s2;
  fastdiv_make((unsigned)G.s1, G.mg1, G.sh1);
  fastdiv_make((unsigned)G.s12, G.mg12, G.sh12);
  return G;
}
template <bool FILL>
__global__ void __launch_bounds__(256) k_classify(int n1, int n2, int n3, const Slab S, int s, const __grid_constant__ ClsGeom G,
                                                  int* __restrict__ lbuf,
                                                  const unsigned char* __restrict__ cubemax, int* __restrict__ uni,
                                                  int* __restrict__ list, int* __restrict__ segcnt, int* __restrict__ ntotal) {
  __shared__ unsigned char s_nu[256];   // cube is non-uniform: its new points walk
  __shared__ int s_row[64];             // entries per row (iz, py) of the tile, then their exclusive prefix
  const int c1 = G.c1, c2 = G.c2, c3 = G.c3;
  const int t1 = G.t1, t2 = G.t2, t3 = G.t3;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int tx, ty, tz;
  {
    const int sb = b / (SB_X * SB_Y * SB_Z), w = b % (SB_X * SB_Y * SB_Z);   // compile-time divisors
    const int sz = fastdiv(sb, G.mg12, G.sh12), r = sb - sz * G.s12;
    const int sy = fastdiv(r, G.mg1, G.sh1), sx = r - sy * G.s1;
    tx = sx * SB_X + w % SB_X;
    ty = sy * SB_Y + (w / SB_X) % SB_Y;
    tz = sz * SB_Z + w / (SB_X * SB_Y);
  }
  if (tx >= t1 || ty >= t2 || tz >= t3) {  // padding block of a ragged super-block
    if (tid == 0) segcnt[b] = 0;
    return;
  }
  const int cx = tx * CLS_TX + (tid & 15), cy = ty * CLS_TY + ((tid >> 4) & 3), cz = tz * CLS_TZ + (tid >> 6);
  const bool active = cx < c1 && cy < c2 && cz < c3;
  const size_t s3 = (size_t)n1 * n2;
  const int hh = s >> 1;
  bool nonuni = false;
  if (active) {
    const size_t t = cx + (size_t)c1 * (cy + (size_t)c2 * cz);
    const int x0 = cx * s, y0 = cy * s, z0 = S.zlo + cz * s;
    const int x1 = (x0 + s < n1) ? x0 + s : 0, y1 = (y0 + s < n2) ? y0 + s : 0;
    const size_t p0 = z0 - S.zlo + 1;                              // local plane of z0
    const size_t p1 = ((z0 + s < n3) ? z0 + s : n3) - S.zlo + 1;   // local plane of the upper corners (may be the halo)
    // all 8 corner labels in one batch of independent loads (a short-circuit chain would make them 8 dependent
    // round trips), then one comparison
    const int* r00 = lbuf + (size_t)n1 * y0 + s3 * p0;
    const int* r10 = lbuf + (size_t)n1 * y1 + s3 * p0;
    const int* r01 = lbuf + (size_t)n1 * y0 + s3 * p1;
    const int* r11 = lbuf + (size_t)n1 * y1 + s3 * p1;
    const int c0 = r00[x0], c1_ = r00[x1], c2_ = r10[x0], c3_ = r10[x1];
    const int c4 = r01[x0], c5 = r01[x1], c6 = r11[x0], c7 = r11[x1];
    const unsigned char hasmax = cubemax[t];
    const int l000 = c0 & LMASK;
    const int diff = ((c1_ ^ c0) | (c2_ ^ c0) | (c3_ ^ c0) | (c4 ^ c0) | (c5 ^ c0) | (c6 ^ c0) | (c7 ^ c0)) & LMASK;
    const bool u = !hasmax && diff == 0;
    uni[t] = u ? l000 : -1;
    nonuni = !u;
    if (FILL && u) {
      for (int o = 1; o < 8; o++) {
        const int x = x0 + ((o & 1) ? hh : 0), y = y0 + ((o & 2) ? hh : 0), z = z0 + ((o & 4) ? hh : 0);
        if (x >= n1 || y >= n2 || z >= S.zhi) continue;
        lbuf[x + n1 * y + s3 * (z - S.zlo + 1)] = (int)((unsigned)l000 | FILLBIT);
      }
    }
  }
  s_nu[tid] = nonuni ? 1 : 0;
  if (!__syncthreads_or(nonuni ? 1 : 0)) {  // nothing to queue in this tile
    if (tid == 0) segcnt[b] = 0;
    return;
  }
  // The new points of the non-uniform cubes are queued in memory order (z, y, x): the tile spans 32 x 8 x 8
  // positions of the stride-s/2 lattice; warp w owns the rows py = w, lane = px.
  const int px = lane, py = wid;
  const int gx = (tx * CLS_TX * 2 + px) * hh, gy = (ty * CLS_TY * 2 + py) * hh;
  unsigned mask[8];
#pragma unroll
  for (int iz = 0; iz < 8; iz++) {
    const int gz = S.zlo + (tz * CLS_TZ * 2 + iz) * hh;
    const bool wk = ((px | py | iz) & 1) && s_nu[(px >> 1) + 16 * ((py >> 1) + 4 * (iz >> 1))] && gx < n1 && gy < n2 && gz < S.zhi;
    mask[iz] = __ballot_sync(FULL, wk);
    if (lane == 0) s_row[iz * 8 + py] = __popc(mask[iz]);
  }
  __syncthreads();
  if (wid == 0) {  // exclusive prefix over the 64 rows
    int a0 = s_row[2 * lane], a1 = s_row[2 * lane + 1];
    int incl = a0 + a1;
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += v;
    }
    const int excl = incl - a0 - a1;
    s_row[2 * lane] = excl;
    s_row[2 * lane + 1] = excl + a0;
    if (lane == 31) {
      segcnt[b] = incl;
      if (incl) atomicAdd(ntotal, incl);
    }
  }
  __syncthreads();
  int* out = list + (size_t)b * CLS_SEGCAP;
#pragma unroll
  for (int iz = 0; iz < 8; iz++) {
    if ((mask[iz] >> lane) & 1u) {
      const int gz = S.zlo + (tz * CLS_TZ * 2 + iz) * hh;
      out[s_row[iz * 8 + py] + __popc(mask[iz] & ((1u << lane) - 1u))] = gx + n1 * (gy + n2 * gz);
    }
  }
}

// safe[c] = uni[c] if the 26 neighbouring cubes (periodic in x,y; owned layers only in z) carry the same
// uniform label, else -1.  z-marching separable "agree" stencil, one pass over the cube array.
__global__ void __launch_bounds__(256) k_safe(int c1, int c2, int c3, const int* __restrict__ uni, int* __restrict__ safe) {
  __shared__ int sbuf[2][(TY + 2) * 34];
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
  const int bx0 = blockIdx.x * 32, by0 = blockIdx.y * TY;
  const int z0 = blockIdx.z * MZC, z1 = min(z0 + MZC, c3);
  const int gx = bx0 + lx, gy = by0 + ly;
  const bool valid = gx < c1 && gy < c2;
  const size_t s3 = (size_t)c1 * c2;
  const int* colp = uni + wrapx(gx, c1) + (size_t)c1 * wrapx(gy, c2);
  int hslot = -1;
  const int* hp = uni;
  if (tid < NHALO) {
    int hc, hr;
    halo_decode(tid, hc, hr);
    hslot = (hr + 1) * 34 + hc + 1;
    hp = uni + wrapx(bx0 + hc, c1) + (size_t)c1 * wrapx(by0 + hr, c2);
  }
  int pm0 = -1, pm1 = -1;
  for (int iz = z0 - 1; iz <= z1; iz++) {
    const bool in = iz >= 0 && iz < c3;
    const int v = in ? __ldg(colp + s3 * iz) : -1;
    const int hv = (in && hslot >= 0) ? __ldg(hp + s3 * iz) : -1;
    const int pm2 = plane3x3<int, OpAgree>(sbuf[(iz - z0 + 1) & 1], v, hslot, hv, lx, ly);
    if (iz > z0 && valid) safe[gx + (size_t)c1 * gy + s3 * (iz - 1)] = OpAgree::c3(pm0, pm1, pm2);
    pm0 = pm1; pm1 = pm2;
  }
}

// octet certificate: vsafe[v] = common label of the 8 cubes {v-1, v}^3 that meet at cube-grid vertex v if
// they are all uniform with the same label, else -1.  x, y wrap periodically when the grid dimension is a
// multiple of the cube stride (px, py), otherwise the vertices next to the ragged last cube are unsafe;
// z: owned layers only, wrapping (pz) on a single GPU.  One thread per (vx, vy) column, marching in z.
__global__ void __launch_bounds__(256) k_vsafe(int c1, int c2, int c3, int px, int py, int pz, int zfull,
                                               const int* __restrict__ uni, int* __restrict__ vsafe) {
  const int vx = blockIdx.x * 32 + (threadIdx.x & 31), vy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (vx >= c1 || vy >= c2) return;
  const int z0 = blockIdx.z * MZC, z1 = min(z0 + MZC, c3);
  const int xm = vx ? vx - 1 : c1 - 1, ym = vy ? vy - 1 : c2 - 1;
  const bool okxy = (px || (vx > 0 && vx < c1 - 1)) && (py || (vy > 0 && vy < c2 - 1));
  const size_t s3 = (size_t)c1 * c2;
  auto layer = [&](int cz) -> int {  // agreement of the 2 x 2 cubes of layer cz
    const int* pl = uni + s3 * cz;
    const int a = __ldg(pl + xm + (size_t)c1 * ym), bq = __ldg(pl + vx + (size_t)c1 * ym);
    const int c = __ldg(pl + xm + (size_t)c1 * vy), d = __ldg(pl + vx + (size_t)c1 * vy);
    return (a == bq && c == d && a == c) ? a : -1;
  };
  int prev;
  if (z0 > 0) prev = layer(z0 - 1);
  else prev = pz ? layer(c3 - 1) : -1;
  for (int vz = z0; vz < z1; vz++) {
    const int cur = layer(vz);
    const bool okz = (vz < c3 - 1) || zfull;  // the vertex below a ragged last layer is not certified
    vsafe[vx + (size_t)c1 * vy + s3 * vz] = (okxy && okz && cur == prev) ? cur : -1;
    prev = cur;
  }
}

// packed certificates of the stride-2 level (SafeMap::pack): out[v] = E[v] | flags << CERT_LBITS (CERT_VOID where E[v]
// is none), where E is the octet map of k_vsafe and flag (ox | oy << 1 | oz << 2) belongs to the point
// (2vx - ox, 2vy - oy, 2vz - oz).  The 5x5x5
// neighbourhood of an even coordinate 2v lies inside the octet of v; that of an odd coordinate 2v-1 needs the octets of
// v-1 and v.  x, y, z wrap like in k_vsafe (px, py, pz); the vertices below the first owned layer of a z-slab do not exist.
__global__ void __launch_bounds__(256) k_pack5(int c1, int c2, int c3, int px, int py, int pz, const int* __restrict__ E,
                                               int* __restrict__ out) {
  const int vx = blockIdx.x * 32 + (threadIdx.x & 31), vy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (vx >= c1 || vy >= c2) return;
  const int z0 = blockIdx.z * MZC, z1 = min(z0 + MZC, c3);
  const size_t s3 = (size_t)c1 * c2;
  const int xm = vx ? vx - 1 : (px ? c1 - 1 : -1), ym = vy ? vy - 1 : (py ? c2 - 1 : -1);
  // the four columns (vx, vy), (vx-1, vy), (vx, vy-1), (vx-1, vy-1); a column that does not exist reads as -2
  const int* col[4] = {E + vx + (size_t)c1 * vy, xm >= 0 ? E + xm + (size_t)c1 * vy : nullptr,
                       ym >= 0 ? E + vx + (size_t)c1 * ym : nullptr, (xm >= 0 && ym >= 0) ? E + xm + (size_t)c1 * ym : nullptr};
  int lo[4], cu[4];  // layers vz-1 and vz, marching upwards
  const int zb = z0 > 0 ? z0 - 1 : (pz ? c3 - 1 : -1);
#pragma unroll
  for (int k = 0; k < 4; k++) lo[k] = (col[k] && zb >= 0) ? __ldg(col[k] + s3 * zb) : -2;
  for (int vz = z0; vz < z1; vz++) {
#pragma unroll
    for (int k = 0; k < 4; k++) cu[k] = col[k] ? __ldg(col[k] + s3 * vz) : -2;
    const int e0 = cu[0];
    int res = CERT_VOID;
    if (e0 >= 0) {
      // ok[m]: the vertex shifted down along the axes in m (1 = x, 2 = y, 4 = z) carries e0 as well
      const bool o1 = cu[1] == e0, o2 = cu[2] == e0, o3 = cu[3] == e0;
      const bool o4 = lo[0] == e0, o5 = lo[1] == e0, o6 = lo[2] == e0, o7 = lo[3] == e0;
      // flag index (cert_flagbit): bit set = odd coordinate 2v-1 (needs the lower vertex along that axis)
      int flags = 1;                                        // (even, even, even): the octet of v alone
      flags |= (o1 ? 1 : 0) << 1;                           // x odd
      flags |= (o2 ? 1 : 0) << 2;                           // y odd
      flags |= ((o1 && o2 && o3) ? 1 : 0) << 3;             // x, y odd
      flags |= (o4 ? 1 : 0) << 4;                           // z odd
      flags |= ((o1 && o4 && o5) ? 1 : 0) << 5;             // x, z odd
      flags |= ((o2 && o4 && o6) ? 1 : 0) << 6;             // y, z odd
      flags |= ((o1 && o2 && o3 && o4 && o5 && o6 && o7) ? 1 : 0) << 7;  // x, y, z odd
      res = e0 | (flags << CERT_LBITS);
    }
    out[vx + (size_t)c1 * vy + s3 * vz] = res;
#pragma unroll
    for (int k = 0; k < 4; k++) lo[k] = cu[k];
  }
}

// ------------------------------------------------------------------------------------------------
// fill + edge detection, one streaming pass over the label planes [za, zb) of this rank.
// RULE: the labels of the last level are materialised here -- a point that is not on the stride-2 lattice
// and whose stride-2 cube is uniform gets uni2 | FILLBIT, every other point keeps what the walkers wrote.
// Every FILLED point with a 26-neighbour of a different label (the refine_edge criterion, is_vol_edge
// bader@proc.f90:730-752) loses FILLBIT and is queued for an exact walk.  Without RULE the labels are
// read as they are (halo planes included) and only the edge test is done.
// SEG: the queue is segmented, one segment of FE_SEGCAP entries per block (32 x TY x MZC points), filled in
// memory order (z, y, x) so that the walkers of a warp start next to each other; otherwise entries are
// appended to a flat list.
// skip_lo/skip_hi: do not test the first/last owned plane (multi-GPU: their halos are not valid yet).
// ------------------------------------------------------------------------------------------------
constexpr int FE_SEGCAP = 32 * TY * MZC;
struct FillArgs {
  int n1, n2, n3;
  Slab S;
  int* lbuf;
  const int* uni2;
  int c1, c2;      // stride-2 cube grid
  int za, zb;      // global planes to process
  int skip_lo, skip_hi;
  int* list; int* nlist; int listcap; int* err;
  int* segcnt;     // SEG only
  int g1, g2, g3;  // blocks per axis (32 x TY x MZC points each); SEG: numbered super-block by super-block
};
constexpr int FSB_X = 4, FSB_Y = 16, FSB_Z = 4;  // super-block of 128^3 points
__host__ __device__ inline int fe_nblocks(int g1, int g2, int g3) {
  return ((g1 + FSB_X - 1) / FSB_X) * ((g2 + FSB_Y - 1) / FSB_Y) * ((g3 + FSB_Z - 1) / FSB_Z) * (FSB_X * FSB_Y * FSB_Z);
}
// Label of a point in two dependent loads that the caller issues one plane apart: u = uniform label of the
// point's stride-2 cube if the fill rule applies to the point (else -1), then u | FILLBIT or the label the
// walkers / earlier levels wrote.  PlaneD holds the block-uniform part of the addressing of one z plane.
struct PlaneD {
  long long lboff, uoff;
  int zpar, owned;
  int rz;  // plane index relative to zlo (wrapped into [0, n3) when the slab is periodic)
};
__device__ __forceinline__ void plane_set(const FillArgs& A, PlaneD& d) {
  d.owned = A.S.periodic ? 1 : ((d.rz >= 0 && d.rz < A.S.nzl) ? 1 : 0);
  d.zpar = d.rz & 1;
  d.lboff = (long long)(A.n1 * A.n2) * (d.rz + 1);
  d.uoff = d.owned ? (long long)(A.c1 * A.c2) * (d.rz >> 1) : 0;
}
__device__ __forceinline__ PlaneD plane_desc(const FillArgs& A, int gz) {
  PlaneD d;
  d.rz = gz - A.S.zlo;  // periodic: zlo = 0
  if (A.S.periodic) d.rz = wrapx(gz, A.n3);
  plane_set(A, d);
  return d;
}
__device__ __forceinline__ void plane_advance(const FillArgs& A, PlaneD& d, int k) {  // k = 1 or 2 planes up
  d.rz += k;
  if (A.S.periodic && d.rz >= A.n3) d.rz -= A.n3;
  plane_set(A, d);
}
template <bool RULE>
__device__ __forceinline__ int fill_u(const PlaneD& d, const int* __restrict__ ucol, int xyodd) {
  if (RULE && d.owned && (xyodd | d.zpar)) return __ldg(ucol + d.uoff);
  return -1;
}
__device__ __forceinline__ int fill_lab(const PlaneD& d, const int* lbcol, int u) {
  if (u >= 0) return (int)((unsigned)u | FILLBIT);
  return lbcol[d.lboff];
}
template <bool RULE, bool SEG>
__global__ void __launch_bounds__(256) k_fill_edge(const __grid_constant__ FillArgs A) {
  __shared__ int sbuf[2][(TY + 2) * 34];
  __shared__ __align__(16) int s_cnt[2][TY];  // SEG: edge points per warp of a plane, double-buffered
  const int n1 = A.n1, n2 = A.n2;
  const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5, lane = lx;
  int bx, by, bz;
  if (SEG) {
    const int b = blockIdx.x;
    const int sb = b / (FSB_X * FSB_Y * FSB_Z), w = b % (FSB_X * FSB_Y * FSB_Z);
    const int s1 = (A.g1 + FSB_X - 1) / FSB_X, s2 = (A.g2 + FSB_Y - 1) / FSB_Y;
    bx = (sb % s1) * FSB_X + w % FSB_X;
    by = ((sb / s1) % s2) * FSB_Y + (w / FSB_X) % FSB_Y;
    bz = (sb / (s1 * s2)) * FSB_Z + w / (FSB_X * FSB_Y);
    if (bx >= A.g1 || by >= A.g2 || bz >= A.g3) {  // padding block of a ragged super-block
      if (tid == 0) A.segcnt[b] = 0;
      return;
    }
  } else {
    bx = blockIdx.x; by = blockIdx.y; bz = blockIdx.z;
  }
  const int bx0 = bx * 32, by0 = by * TY;
  const int z0 = A.za + bz * MZC, z1 = min(z0 + MZC, A.zb);
  const int gx = bx0 + lx, gy = by0 + ly;
  const bool valid = gx < n1 && gy < n2;
  const int wx = wrapx(gx, n1), wy = wrapx(gy, n2);
  // per-thread column bases of the own point and (first NHALO threads) of one halo point
  const int* lbc = A.lbuf + wx + (size_t)n1 * wy;
  const int* uc = A.uni2 + (wx >> 1) + (size_t)A.c1 * (wy >> 1);
  const int xyodd = (wx | wy) & 1;
  int hslot = -1, hxyodd = 0;
  const int* hlbc = A.lbuf;
  const int* huc = A.uni2;
  if (tid < NHALO) {
    int hc, hr;
    halo_decode(tid, hc, hr);
    hslot = (hr + 1) * 34 + hc + 1;
    const int hx = wrapx(bx0 + hc, n1), hy = wrapx(by0 + hr, n2);
    hlbc = A.lbuf + hx + (size_t)n1 * hy;
    huc = A.uni2 + (hx >> 1) + (size_t)A.c1 * (hy >> 1);
    hxyodd = (hx | hy) & 1;
  }
  const bool hal = hslot >= 0;
  if (SEG && tid < 2 * TY) (&s_cnt[0][0])[tid] = 0;
  int* wout = A.lbuf + gx + (size_t)n1 * gy - (size_t)n1 * n2 * (A.S.zlo - 1);  // wout[s3 * gz] = label of (gx, gy, gz)
  const size_t s3 = (size_t)n1 * n2;
  int* segout = A.list + (size_t)blockIdx.x * FE_SEGCAP;  // SEG only
  int nseg = 0;          // block-uniform: entries in this block's segment so far
  unsigned m_pend = 0;   // SEG: edge mask of the previous plane, written one iteration later (see below)
  int id_pend = 0;
  int pm0 = -1, pm1 = -1, own1 = 0;
  // software pipeline: cube labels two planes ahead, point labels one plane ahead
  PlaneD d1 = plane_desc(A, z0 - 1);
  int u_a = fill_u<RULE>(d1, uc, xyodd), hu_a = hal ? fill_u<RULE>(d1, huc, hxyodd) : -1;
  int own_n = fill_lab(d1, lbc, u_a), hv_n = hal ? fill_lab(d1, hlbc, hu_a) : 0;
  plane_advance(A, d1, 1);
  u_a = fill_u<RULE>(d1, uc, xyodd); hu_a = hal ? fill_u<RULE>(d1, huc, hxyodd) : -1;
  for (int iz = z0 - 1; iz <= z1; iz++) {
    const int own = own_n, hv = hv_n & LMASK;
    if (iz < z1) {  // d1 describes plane iz+1
      own_n = fill_lab(d1, lbc, u_a);
      if (hal) hv_n = fill_lab(d1, hlbc, hu_a);
      if (iz + 1 < z1) {
        plane_advance(A, d1, 1);
        u_a = fill_u<RULE>(d1, uc, xyodd);
        if (hal) hu_a = fill_u<RULE>(d1, huc, hxyodd);
      }
    }
    const int pm2 = plane3x3<int, OpAgree>(sbuf[(iz - z0 + 1) & 1], own & LMASK, hslot, hv, lx, ly);
    if (SEG) {
      // the barrier inside plane3x3 has published every warp's count of the plane tested one iteration ago:
      // write that plane's entries at block-level positions (z, y, x order)
      const int4* c = reinterpret_cast<const int4*>(s_cnt[(iz - z0) & 1]);
      const int4 c0 = c[0], c1 = c[1];
      const int cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      int base = nseg, total = 0;
#pragma unroll
      for (int w = 0; w < TY; w++) {
        if (w < ly) base += cv[w];
        total += cv[w];
      }
      if ((m_pend >> lane) & 1u) segout[base + __popc(m_pend & ((1u << lane) - 1u))] = id_pend;
      nseg += total;
    }
    unsigned m = 0;
    if (iz > z0) {  // plane iz-1 is complete
      const int gz = iz - 1;
      const bool test = !((A.skip_lo && gz == A.S.zlo) || (A.skip_hi && gz == A.S.zhi - 1));
      const bool filled = ((unsigned)own1 & FILLBIT) != 0;
      const bool edge = valid && filled && test && OpAgree::c3(pm0, pm1, pm2) < 0;
      if (valid && (RULE ? filled : edge)) wout[s3 * gz] = edge ? (own1 & LMASK) : own1;
      m = __ballot_sync(FULL, edge);
      if (m) {
        if (SEG) {
          id_pend = gx + n1 * (gy + n2 * gz);
        } else {
          const int rank = __popc(m & ((1u << lane) - 1u));
          int base = 0;
          if (lane == __ffs(m) - 1) base = atomicAdd(A.nlist, __popc(m));
          base = __shfl_sync(FULL, base, __ffs(m) - 1);
          if (edge) {
            const int slot = base + rank;
            if (slot < A.listcap) A.list[slot] = gx + n1 * (gy + n2 * gz);
            else atomicExch(A.err, 4);
          }
        }
      }
    }
    if (SEG) {
      m_pend = m;
      if (lane == 0) s_cnt[(iz - z0 + 1) & 1][ly] = __popc(m);
    }
    pm0 = pm1; pm1 = pm2; own1 = own;
  }
  if (SEG) {  // entries of the last plane
    __syncthreads();
    const int* c = s_cnt[(z1 + 1 - z0) & 1];
    int base = nseg, total = 0;
#pragma unroll
    for (int w = 0; w < TY; w++) {
      const int v = c[w];
      if (w < ly) base += v;
      total += v;
    }
    if ((m_pend >> lane) & 1u) segout[base + __popc(m_pend & ((1u << lane) - 1u))] = id_pend;
    nseg += total;
    if (tid == 0) {
      A.segcnt[blockIdx.x] = nseg;
      if (nseg) atomicAdd(A.nlist, nseg);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fill + edge detection driven by the octet certificates of the last level (single GPU, periodic slab).
// One thread per cube-grid vertex v = (vx, vy, vz), which owns the points {2v-1, 2v}^3 (the points whose
// nearest vertex is v).  vsafe[v] >= 0 says that the 8 stride-2 cubes around v are uniform with one label:
// every owned point then has a uniformly labelled 3x3x3 neighbourhood, so it is filled (the RULE of
// k_fill_edge) and cannot be an edge point -- 7 stores, nothing else (~90 % of the vertices).  Otherwise the
// labels of the 4x4x4 points around v are evaluated with the same RULE (uniform cube -> its label, else what
// the walkers wrote) and the owned FILLED points are tested against their 26 neighbours exactly like
// is_vol_edge (bader@proc.f90:730-752).  Same labels, same edge set as k_fill_edge<true, true>; the edge
// points of a block (32 x 8 x VZC vertices = 64 x 16 x 2 VZC points) go to the block's segment.
// ------------------------------------------------------------------------------------------------
constexpr int VZC = 16;                          // vertex layers per block
constexpr int FV_SEGCAP = 64 * 16 * 2 * VZC;     // points per block
constexpr int VSB_X = 2, VSB_Y = 8, VSB_Z = 4;   // super-block of 128^3 points
__host__ __device__ inline int fv_nblocks(int g1, int g2, int g3) {
  return ((g1 + VSB_X - 1) / VSB_X) * ((g2 + VSB_Y - 1) / VSB_Y) * ((g3 + VSB_Z - 1) / VSB_Z) * (VSB_X * VSB_Y * VSB_Z);
}
struct FillVArgs {
  int n1, n2, n3;
  int c1, c2, c3;        // stride-2 cube grid of the owned planes = vertex grid
  int nzl, periodic;     // owned planes; periodic = 1: single GPU, z wraps.  Otherwise (z-slab of a multi-GPU run)
                         // the first and last owned planes are only filled, not tested: their neighbours live on
                         // another rank and the caller tests them against fresh halos (edge_faces)
  int nvz;               // vertex layers to visit: c3 (periodic) or c3 + 1 (slab: the layer above the last cube)
  int zlo;               // global plane of the first owned plane (edge ids are global)
  int* label;            // owned planes, label[x + n1*(y + n2*zl)], zl = 0 .. nzl-1
  const int* uni2;
  const int* vsafe;
  int pack;              // vsafe entries are packed (SafeMap::pack): the label is the low CERT_LBITS bits
  int* list; int* nlist; int* segcnt;
  int g1, g2, g3;        // blocks per axis
};
__device__ __forceinline__ int agree3(int a, int b, int c) { return (a == b && b == c) ? a : -1; }
#ifndef FV_MINB
#define FV_MINB 3
#endif
__global__ void __launch_bounds__(256, FV_MINB) k_fill_edge_v(const __grid_constant__ FillVArgs A) {
  __shared__ int s_count, s_nslow;
  __shared__ unsigned short s_slow[32 * 8 * VZC];  // block-local (lx | ly << 5 | lz << 8) of the uncertified vertices
  const int n1 = A.n1, n2 = A.n2, n3 = A.n3;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.x;
  int bx, by, bz;
  {
    const int sb = b / (VSB_X * VSB_Y * VSB_Z), w = b % (VSB_X * VSB_Y * VSB_Z);
    const int s1 = (A.g1 + VSB_X - 1) / VSB_X, s2 = (A.g2 + VSB_Y - 1) / VSB_Y;
    bx = (sb % s1) * VSB_X + w % VSB_X;
    by = ((sb / s1) % s2) * VSB_Y + (w / VSB_X) % VSB_Y;
    bz = (sb / (s1 * s2)) * VSB_Z + w / (VSB_X * VSB_Y);
  }
  if (bx >= A.g1 || by >= A.g2 || bz >= A.g3) {  // padding block of a ragged super-block
    if (tid == 0) A.segcnt[b] = 0;
    return;
  }
  if (tid == 0) { s_count = 0; s_nslow = 0; }
  __syncthreads();
  const size_t s3 = (size_t)n1 * n2, u3 = (size_t)A.c1 * A.c2;
  const int vz0 = bz * VZC, vz1 = min(vz0 + VZC, A.nvz);
  const int per = A.periodic;
  // ---- phase 1: certified vertices store their fills; the others are collected ----
  {
    const int vx = bx * 32 + lane, vy = by * 8 + (tid >> 5);
    const bool vvalid = vx < A.c1 && vy < A.c2;
    // owned coordinates per axis: 2v-1 (absent for v = 0 unless n is even: then it is the point n-1) and 2v
    const int x1 = wrapx(2 * vx - 1, n1), x2 = 2 * vx, y1 = wrapx(2 * vy - 1, n2), y2 = 2 * vy;
    const bool okx1 = vx > 0 || (n1 & 1) == 0, oky1 = vy > 0 || (n2 & 1) == 0;
    const int* vs = A.vsafe + vx + (size_t)A.c1 * vy;
    const int lmask = A.pack ? CERT_LMASK : LMASK;
    int sv_next = (vvalid && vz0 < A.c3) ? __ldg(vs + u3 * vz0) : -1;
    for (int vz = vz0; vz < vz1; vz++) {
      const int sv = sv_next < 0 ? -1 : (sv_next & lmask);
      sv_next = -1;  // the slab's extra layer above the last cube has no certificate
      if (vz + 1 < vz1 && vz + 1 < A.c3 && vvalid) sv_next = __ldg(vs + u3 * (vz + 1));
      const bool slow = vvalid && sv < 0;
      if (vvalid && sv >= 0) {  // the 8 cubes around the vertex are owned and uniform: both z planes exist
        const int fl = (int)((unsigned)sv | FILLBIT);
        const bool okz1 = vz > 0 || (per && (n3 & 1) == 0);
        int* p2 = A.label + s3 * (2 * vz);
        int* p1 = A.label + s3 * (per ? wrapx(2 * vz - 1, n3) : max(2 * vz - 1, 0));
        int* r22 = p2 + (size_t)n1 * y2;
        int* r21 = p2 + (size_t)n1 * y1;
        if (okx1) r22[x1] = fl;              // (x2, y2, z2) is the stride-2 lattice point: it keeps its label
        if (oky1) { r21[x2] = fl; if (okx1) r21[x1] = fl; }
        if (okz1) {
          int* r12 = p1 + (size_t)n1 * y2;
          int* r11 = p1 + (size_t)n1 * y1;
          r12[x2] = fl;
          if (okx1) r12[x1] = fl;
          if (oky1) { r11[x2] = fl; if (okx1) r11[x1] = fl; }
        }
      }
      const unsigned m = __ballot_sync(FULL, slow);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_nslow, __popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (slow) s_slow[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(lane | ((tid >> 5) << 5) | ((vz - vz0) << 8));
      }
    }
  }
  __syncthreads();
  // ---- phase 2: the uncertified vertices, densely ----
  const int nslow = s_nslow;
  int* segout = A.list + (size_t)b * FV_SEGCAP;
  for (int e0 = 0; e0 < nslow; e0 += 256) {
    const int e = e0 + tid;
    int nedge = 0;
    int eid[8];
    if (e < nslow) {
      const int code = s_slow[e];
      const int vx = bx * 32 + (code & 31), vy = by * 8 + ((code >> 5) & 7), vz = vz0 + (code >> 8);
      int X[4], Y[4], Z[4];
      bool zin[4];  // plane is one of the owned planes (always, when z is periodic)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        X[i] = wrapx(2 * vx - 2 + i, n1);
        Y[i] = wrapx(2 * vy - 2 + i, n2);
        const int zl = 2 * vz - 2 + i;
        zin[i] = per || (zl >= 0 && zl < A.nzl);
        Z[i] = per ? wrapx(zl, n3) : min(max(zl, 0), A.nzl - 1);
      }
      const bool okx1 = vx > 0 || (n1 & 1) == 0, oky1 = vy > 0 || (n2 & 1) == 0;
      // existence of the two owned z planes, and whether their points are tested here
      const bool okz[2] = {per ? (vz > 0 || (n3 & 1) == 0) : vz > 0, per ? true : 2 * vz < A.nzl};
      const bool tz[2] = {per != 0 || (2 * vz - 1 >= 1 && 2 * vz - 1 <= A.nzl - 2), per != 0 || (2 * vz >= 1 && 2 * vz <= A.nzl - 2)};
      int Ag[4][2][2];   // in-plane 3x3 agreement around the owned (x, y) positions, per plane
      int own[2][2][2];  // full labels (with FILLBIT) of the owned points [k-1][j-1][i-1]
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!zin[k]) {  // not this rank's plane: the owned points next to it are not tested in this pass
#pragma unroll
          for (int q = 0; q < 4; q++) Ag[k][q >> 1][q & 1] = -1;
          continue;
        }
        int l[4][4];
        const size_t zoff = s3 * Z[k], uoff = u3 * (Z[k] >> 1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const size_t urow = uoff + (size_t)A.c1 * (Y[j] >> 1);
          int ucache = -2, ucube = -1;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int cx = X[i] >> 1;
            if (cx != ucube) { ucube = cx; ucache = __ldg(A.uni2 + urow + cx); }
            const bool owned = k >= 1 && k <= 2 && j >= 1 && j <= 2 && i >= 1 && i <= 2;
            int raw;
            if (ucache >= 0) {
              // uniform cube: its corners carry this label, its other points are filled with it (RULE)
              raw = ucache;  // neighbours: label bits only
              if (owned) {
                const bool lattice = !((X[i] | Y[j] | Z[k]) & 1);
                raw = lattice ? A.label[X[i] + (size_t)n1 * Y[j] + zoff] : (int)((unsigned)ucache | FILLBIT);
              }
            } else {
              raw = A.label[X[i] + (size_t)n1 * Y[j] + zoff];
            }
            l[j][i] = raw & LMASK;
            if (owned) own[k - 1][j - 1][i - 1] = raw;
          }
        }
        int r[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          r[j][0] = agree3(l[j][0], l[j][1], l[j][2]);
          r[j][1] = agree3(l[j][1], l[j][2], l[j][3]);
        }
#pragma unroll
        for (int bb = 0; bb < 2; bb++)
#pragma unroll
          for (int a = 0; a < 2; a++) Ag[k][bb][a] = agree3(r[bb][a], r[bb + 1][a], r[bb + 2][a]);
      }
#pragma unroll
      for (int k = 1; k <= 2; k++)
#pragma unroll
        for (int j = 1; j <= 2; j++)
#pragma unroll
          for (int i = 1; i <= 2; i++) {
            if ((i == 1 && !okx1) || (j == 1 && !oky1) || !okz[k - 1]) continue;
            const int raw = own[k - 1][j - 1][i - 1];
            if (!((unsigned)raw & FILLBIT)) continue;  // walked points keep what the walkers wrote
            const bool edge = tz[k - 1] && agree3(Ag[k - 1][j - 1][i - 1], Ag[k][j - 1][i - 1], Ag[k + 1][j - 1][i - 1]) < 0;
            const size_t off = X[i] + (size_t)n1 * Y[j] + s3 * Z[k];
            const bool lattice = !((X[i] | Y[j] | Z[k]) & 1);
            if (edge) { A.label[off] = raw & LMASK; eid[nedge++] = (int)(off + s3 * A.zlo); }
            else if (!lattice) A.label[off] = raw;
          }
    }
    // queue the edge points of this warp (the order inside a block does not matter: items are cut per segment)
    const unsigned any = __ballot_sync(FULL, nedge > 0);
    if (any) {
      int incl = nedge;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t;
      }
      int base = 0;
      if (lane == 31) base = atomicAdd(&s_count, incl);
      base = __shfl_sync(FULL, base, 31) + incl - nedge;
#pragma unroll
      for (int q = 0; q < 8; q++)
        if (q < nedge) segout[base + q] = eid[q];
    }
  }
  __syncthreads();
  if (tid == 0) {
    A.segcnt[b] = s_count;
    if (s_count) atomicAdd(A.nlist, s_count);
  }
}

// terminal candidate index -> output index (only needed when a candidate maximum was reached by no
// trajectory, e.g. the twin of a two-point plateau); keeps FILLBIT
__global__ void k_permute_labels(long long nn, int* __restrict__ label, const int* __restrict__ perm) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int l = label[i];
    label[i] = (int)((unsigned)perm[l & LMASK] | ((unsigned)l & FILLBIT));
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
    const int l = label[i] & LMASK;
    if (key < first[l]) atomicMin(first + l, key);
  }
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

// alignment of the interior slab boundaries = largest top lattice stride the slabs allow
int c2g_slab_align(int n3, int nranks) {
  int a = 32;
  while (a > 4 && (long long)a * 2 * nranks > n3) a >>= 1;
  return a;
}

// slab boundaries: multiples of c2g_slab_align, as even as possible
void c2g_slab_bounds(int n3, int nranks, int rank, int* zlo, int* zhi) {
  const int a = c2g_slab_align(n3, nranks);
  auto bound = [&](int r) -> int {
    if (r >= nranks) return n3;
    long long z = (long long)n3 * r / nranks;
    z = (z + a / 2) / a * a;
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
  if (ctx->group) return grp_bader_assign(ctx, handle, car2lat, lat_i_dist, algo, order, nmax_out, res_out);
  if (!nmax_out || !res_out || !car2lat || !lat_i_dist) return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: null argument");
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: invalid grid handle %d", handle);
  c2g_grid_ready(ctx, handle);
  const c2g_grid& g = ctx->grids[handle];
  if (g.nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "c2g_bader_assign: grid too large for int32 indices");
  cudaStream_t st = ctx->stream;
  const int G = ctx->nranks;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  BaderParams P;
  P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
  P.in1 = std::max(P.n1 - 6, 0); P.in2 = std::max(P.n2 - 6, 0); P.in3 = std::max(P.n3 - 6, 0);
  memcpy(P.c2l, car2lat, sizeof(P.c2l));
  memcpy(P.lid, lat_i_dist, sizeof(P.lid));
  fastdiv_make((unsigned)P.n1, P.mg1, P.sh1);
  fastdiv_make((unsigned)P.n1 * (unsigned)P.n2, P.mg12, P.sh12);
  const int n1 = P.n1, n2 = P.n2, n3 = P.n3;
  const size_t plane = (size_t)n1 * n2;
  // diagonal car2lat (orthogonal cell): the off-diagonal products are exact zeros and can be skipped
  const bool ortho = P.c2l[1] == 0.0 && P.c2l[2] == 0.0 && P.c2l[3] == 0.0 && P.c2l[5] == 0.0 && P.c2l[6] == 0.0 && P.c2l[7] == 0.0;
  Slab S;
  c2g_slab_bounds(n3, G, ctx->rank, &S.zlo, &S.zhi);
  S.nzl = S.zhi - S.zlo;
  S.periodic = (G == 1) ? 1 : 0;
  const long long nnl = (long long)plane * S.nzl;  // owned points

  // top lattice stride: a power of two <= 32, small against the grid, compatible with the slab boundaries
  int L0 = 4;
  {
    const int minn = std::min(n1, std::min(n2, n3));
    while (L0 < 32 && L0 * 4 <= minn) L0 <<= 1;
    if (G > 1) L0 = std::min(L0, c2g_slab_align(n3, G));
    if (const char* e = getenv("C2G_BADER_L0")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32) L0 = std::min(L0, v); }
  }
  int nlev = 0;
  for (int s = 2; s <= L0; s <<= 1) nlev++;

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

  // ---- per-level cube arrays (level i: stride 2<<i) ----
  struct Level { int s, c1, c2, c3; size_t nc; };
  Level lev[MAXLEV];
  DevBuf b_cubemax[MAXLEV], b_uni[MAXLEV], b_safe[MAXLEV];
  CubeFlags CF;
  memset(&CF, 0, sizeof(CF));
  CF.nlev = (algo == C2G_BADER_EXACT) ? 0 : nlev;
  for (int i = 0; i < nlev && algo != C2G_BADER_EXACT; i++) {
    const int s = 2 << i;
    lev[i].s = s;
    lev[i].c1 = (n1 + s - 1) / s; lev[i].c2 = (n2 + s - 1) / s; lev[i].c3 = (S.nzl + s - 1) / s;
    lev[i].nc = std::max<size_t>(1, (size_t)lev[i].c1 * lev[i].c2 * lev[i].c3);
    C2G_CUDA(ctx, b_cubemax[i].alloc(ctx, lev[i].nc));
    C2G_CUDA(ctx, b_uni[i].alloc(ctx, sizeof(int) * lev[i].nc));
    C2G_CUDA(ctx, b_safe[i].alloc(ctx, sizeof(int) * lev[i].nc));
    CF.p[i] = b_cubemax[i].as<unsigned char>();
    CF.c1[i] = lev[i].c1; CF.c2[i] = lev[i].c2;
  }

  // ---- K0: candidate maxima of the slab ----
  DevBuf b_cand, b_cnt;
  int maxcand = (int)std::max<long long>(1, std::min<long long>(nnl, std::max<long long>(1 << 16, nnl / 64)));
  C2G_CUDA(ctx, b_cnt.alloc(ctx, C_NCNT * sizeof(int)));
  // counters: [0] ncand, [1] nlist, [2] noverflow, [3] err ; [4..5] nsteps (u64) ; [6] scratch for collectives,
  //           [7] nnext ; [8..9] work cursor (u64) ; [10] number of work items
  int* cnt = b_cnt.as<int>();
  unsigned long long* nsteps = (unsigned long long*)(cnt + 4);
  unsigned long long* cursor = (unsigned long long*)(cnt + 8);
  if (!ctx->hpin) C2G_CUDA(ctx, cudaHostAlloc((void**)&ctx->hpin, 64 * sizeof(int), cudaHostAllocDefault));
  int* const hcnt = ctx->hpin;  // page-locked: the counter read-backs between the launches are plain DMA writes
  for (int attempt = 0;; attempt++) {
    C2G_CUDA(ctx, b_cand.alloc(ctx, sizeof(int) * (size_t)maxcand));
    C2G_CUDA(ctx, cudaMemsetAsync(cnt, 0, C_NCNT * sizeof(int), st));
    ctx->prof_begin("bader_clear_flags");
    for (int i = 0; i < CF.nlev; i++) C2G_CUDA(ctx, cudaMemsetAsync(CF.p[i], 0, lev[i].nc, st));
    ctx->prof_end(CF.nlev);
    if (S.nzl > 0) {
      dim3 grid((n1 + 255) / 256, n2, (S.nzl + MZC - 1) / MZC);
      ctx->prof_begin("bader_maxima");
      if (n1 % 2 == 0 && ((uintptr_t)g.d % 16) == 0)
        k_maxima2<<<grid, 128, 0, st>>>(P, S, g.d, b_cand.as<int>(), cnt, maxcand, CF);
      else
        k_maxima<<<grid, 256, 0, st>>>(P, S, g.d, b_cand.as<int>(), cnt, maxcand, CF);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
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

  // top stride actually used: small against the mean basin size (a cube certified by its 8 corners must not
  // be wide enough for another basin to pass between them unseen)
  if (algo != C2G_BADER_EXACT) {
    const double bsize = std::cbrt((double)g.nn / (double)ncand);
    int lmax = 4;
    while (lmax < 32 && lmax * 2 * 6 <= bsize) lmax <<= 1;
    if (getenv("C2G_BADER_L0") == nullptr) L0 = std::min(L0, lmax);
    nlev = 0;
    for (int s = 2; s <= L0; s <<= 1) nlev++;
  }

  // work lists: [0] segmented (classify levels, fill+edge), [1],[2] flat (edge-fix ping-pong); overflow list;
  // dense list of every walker so far + its stop log (where its walk was cut short)
  DevBuf b_list[3], b_over, b_segcnt, b_items, b_blksum, b_dlist, b_stop;
  size_t itemcap = 0, blksumcap = 0;
  size_t segints = 1, nsegmax = 1;
  int fe_gx = (n1 + 31) / 32, fe_gy = (n2 + TY - 1) / TY, fe_gz = std::max(1, (S.nzl + MZC - 1) / MZC);
  if (algo != C2G_BADER_EXACT) {
    for (int i = 0; i < nlev; i++) {
      const size_t nt = (size_t)cls_nblocks((lev[i].c1 + CLS_TX - 1) / CLS_TX, (lev[i].c2 + CLS_TY - 1) / CLS_TY,
                                            std::max(1, (lev[i].c3 + CLS_TZ - 1) / CLS_TZ));
      segints = std::max(segints, nt * CLS_SEGCAP);
      nsegmax = std::max(nsegmax, nt);
    }
    const size_t nfe = (size_t)fe_nblocks(fe_gx, fe_gy, fe_gz);
    segints = std::max(segints, nfe * FE_SEGCAP);
    nsegmax = std::max(nsegmax, nfe);
    if (nlev > 0) {  // vertex-driven fill + edge pass (single GPU)
      const size_t nfv = (size_t)fv_nblocks((lev[0].c1 + 31) / 32, (lev[0].c2 + 7) / 8, std::max(1, (lev[0].c3 + 1 + VZC - 1) / VZC));
      segints = std::max(segints, nfv * FV_SEGCAP);
      nsegmax = std::max(nsegmax, nfv);
    }
  }
  const long long listcap = (algo == C2G_BADER_EXACT) ? 1 : std::max<long long>(1, nnl);
  const long long dcap = (algo == C2G_BADER_EXACT) ? 1 : std::min<long long>(nnl + nnl / 8 + (1 << 20), 0x7fffffffll);
  long long doff = 0;  // entries of the dense list in use
  C2G_CUDA(ctx, b_list[0].alloc(ctx, sizeof(int) * segints));
  C2G_CUDA(ctx, b_segcnt.alloc(ctx, sizeof(int) * nsegmax));
  C2G_CUDA(ctx, b_list[1].alloc(ctx, sizeof(int) * (size_t)listcap));
  C2G_CUDA(ctx, b_list[2].alloc(ctx, sizeof(int) * (size_t)listcap));
  C2G_CUDA(ctx, b_dlist.alloc(ctx, sizeof(int) * (size_t)dcap));
  C2G_CUDA(ctx, b_stop.alloc(ctx, sizeof(int) * (size_t)dcap));
  const long long overcap = std::max<long long>(1024, nnl);  // worst case: every walker of a launch is handed over
  C2G_CUDA(ctx, b_over.alloc(ctx, sizeof(int2) * (size_t)overcap));
  long long walked = 0, fixpts = 0, fixpasses = 0, noverflow_total = 0;

  WalkArgs WA;
  memset(&WA, 0, sizeof(WA));
  WA.rho = g.d; WA.label_g = label_g; WA.h = h; WA.reached = reached; WA.S = S;
  WA.cursor = cursor; WA.overflow = b_over.as<int2>(); WA.noverflow = cnt + 2; WA.overcap = (int)std::min<long long>(overcap, 0x7fffffff);
  WA.err = cnt + 3; WA.nsteps = nsteps; WA.nnext = cnt + 7; WA.ninval = cnt + 11;
  // Refill policy of the persistent walkers.  The dense last level is issue-bound and wants its lanes refilled
  // early; the sparser coarse levels, the top lattice and the fix passes are latency-bound and run faster when a
  // cohort of neighbouring walkers stays in step (their loads coalesce), so they refill late and look at the
  // work queue less often (measured at 1024^3: l2 4.9 -> 3.6 ms, l4 2.4 -> 1.9 ms).
  int refill_fine = REFILL_MIN, spc_fine = STEPS_PER_CHECK, refill_coarse = 24, spc_coarse = 8;
  if (const char* e = getenv("C2G_STEPS_PER_CHECK")) spc_fine = std::max(1, std::min(64, atoi(e)));
  if (const char* e = getenv("C2G_REFILL_MIN")) refill_fine = std::max(1, std::min(32, atoi(e)));
  if (const char* e = getenv("C2G_SPC_COARSE")) spc_coarse = std::max(1, std::min(64, atoi(e)));
  if (const char* e = getenv("C2G_REFILL_COARSE")) refill_coarse = std::max(1, std::min(32, atoi(e)));
  WA.refill_min = refill_coarse;
  WA.steps_per_check = spc_coarse;
  const int walk_occ = std::max(4, C2G_W3_MINB * W3_NT / 256);  // resident 256-thread walker blocks per SM (<= 64 registers)
  const bool walk_stats = getenv("C2G_BADER_VERBOSE") != nullptr || getenv("C2G_BADER_STATS") != nullptr;
  // k_walk2 (every idle lane refilled before every step) is an opt-in experiment: it keeps 30 of 32 lanes busy but
  // its lanes no longer move as cohorts of neighbours, every load then touches up to 32 cache lines, and the last
  // level took 26.4 ms against k_walk's 10.6 ms at 1024^3 (profiles/r02b_bench_1024_kwalk2_rejected.json)
  const bool walk_old = getenv("C2G_WALK2") == nullptr;
  const int wblocks = ctx->nsm * walk_occ;
  // which launches use the block-cooperative walkers (k_walk3): 1 = last level, 2 = edge fix, 4 = coarse levels, 8 = top lattice
  // per class (last level / every other launch): idle threads of a block that trigger a pack + refill, steps between looks
  // (B200, 1024^3: last level 11.95 -> 10.17 ms at 8 steps / 32 idle; edge fix 5.01 -> 2.45, stride-2 level 3.49 -> 3.10,
  // stride-4 level 1.90 -> 1.68 ms at 16 steps / 128 idle; profiles/r02g_sweep_walk3.txt)
  int walk3_mask = 15, w3_idle = 32, w3_k = 8, w3_idle_c = 128, w3_k_c = 16;
  if (const char* e = getenv("C2G_WALK3")) walk3_mask = atoi(e);
  if (const char* e = getenv("C2G_W3_IDLE")) w3_idle = std::max(1, std::min(256, atoi(e)));
  if (const char* e = getenv("C2G_W3_K")) w3_k = std::max(1, std::min(64, atoi(e)));
  if (const char* e = getenv("C2G_W3_IDLEC")) w3_idle_c = std::max(1, std::min(256, atoi(e)));
  if (const char* e = getenv("C2G_W3_KC")) w3_k_c = std::max(1, std::min(64, atoi(e)));

  // Device-side bookkeeping (default): the host enqueues whole levels without reading a counter back -- list lengths,
  // the fill of the dense walker list and the hand-over walks are handled by k_items_*, k_walk_end and k_walk_big2 from
  // counters in device memory; the host synchronises once for the candidate maxima, then only in the edge-fix loop
  // (after the second pass and after every later one) and at the end.  C2G_SYNC_LEVELS=1, the per-level statistics of
  // C2G_BADER_VERBOSE and the older walker kernels keep the host-driven flow (a read-back after every launch group).
  const bool verbose = getenv("C2G_BADER_VERBOSE") != nullptr;
  const bool async = algo != C2G_BADER_EXACT && walk3_mask == 15 && !verbose && getenv("C2G_SYNC_LEVELS") == nullptr;
  const int bigcap2 = (int)std::min<long long>(g.nn, 1 << 16), bigthreads2 = 16 * 64;
  DevBuf b_bigscr;
  if (async) C2G_CUDA(ctx, b_bigscr.alloc(ctx, sizeof(int) * (size_t)bigthreads2 * bigcap2));

  auto check_err = [&]() -> int {
    switch (hcnt[3]) {
      case 0: return C2G_OK;
      case 1: return ctx->fail(C2G_ERR_NEWMAX, "bader walk ended on a point that is not a local maximum");
      case 2: return ctx->fail(C2G_ERR_OVERFLOW, "trajectory longer than the scratch path buffer");
      case 3: return ctx->fail(C2G_ERR_OVERFLOW, "too many long trajectories");
      default: return ctx->fail(C2G_ERR_OVERFLOW, "work list overflow");
    }
  };
  // read the counters; re-walk the trajectories whose path buffer overflowed
  auto drain = [&](bool fix) -> int {
    if (async) {
      ctx->prof_begin("bader_walk_big");
      if (fix) k_walk_big2<true><<<bigthreads2 / 64, 64, 0, st>>>(P, WA, cnt, b_bigscr.as<int>(), bigcap2);
      else k_walk_big2<false><<<bigthreads2 / 64, 64, 0, st>>>(P, WA, cnt, b_bigscr.as<int>(), bigcap2);
      k_walk_end<<<1, 1, 0, st>>>(cnt, fix ? 1 : 0);
      ctx->prof_end(2);
      C2G_KERNEL_CHECK(ctx);
      return C2G_OK;
    }
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    int rc = check_err();
    if (rc) return rc;
    const int nov = hcnt[2];
    if (nov == 0) return C2G_OK;
    noverflow_total += nov;
    const int bigcap = (int)std::min<long long>(g.nn, 1 << 22);
    const int chunk = (int)std::max<long long>(1, std::min<long long>(nov, (1ll << 31) / bigcap));  // <= 8 GiB scratch
    DevBuf b_scr;
    C2G_CUDA(ctx, b_scr.alloc(ctx, sizeof(int) * (size_t)chunk * bigcap));
    WalkArgs W2 = WA;
    for (int off = 0; off < nov; off += chunk) {
      const int c = std::min(chunk, nov - off);
      W2.overflow = b_over.as<int2>() + off;
      ctx->prof_begin("bader_walk_big");
      if (fix) k_walk_big<true><<<c2g_blocks_for(c, 64), 64, 0, st>>>(P, W2, c, b_scr.as<int>(), bigcap);
      else k_walk_big<false><<<c2g_blocks_for(c, 64), 64, 0, st>>>(P, W2, c, b_scr.as<int>(), bigcap);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    C2G_CUDA(ctx, cudaMemsetAsync(cnt + 2, 0, sizeof(int), st));
    C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    hcnt[2] = 0;
    return check_err();
  };
  auto launch_walk = [&](long long count, bool fix, const char* name) -> int {
    C2G_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
    const int blocks = WA.count_dev ? wblocks : (int)std::min<long long>(wblocks, (count + 255) / 256);
    ctx->prof_begin(name);
    const int variant = (ortho ? 4 : 0) | (fix ? 2 : 0) | (walk_stats ? 1 : 0);
    // launch class: 8 = top lattice, 4 = coarse levels, 1 = last level, 2 = edge fix (bits of C2G_WALK3)
    const int cls = fix ? 2 : (WA.list == nullptr ? 8 : (WA.sm_level == 0 ? 1 : 4));
    if (walk3_mask & cls) {
      // k_walk3: refill_min = idle threads of the block that trigger a pack + refill, steps_per_check = steps between two looks
      WA.refill_min = cls == 1 ? w3_idle : w3_idle_c; WA.steps_per_check = cls == 1 ? w3_k : w3_k_c;
      // the sparse lattices (top, strides 8 and 4) run a little faster with fewer looks at the queue (24 steps: stride 4
      // 1.70 -> 1.56 ms at 1024^3), the stride-2 level and the edge fix do not
      if ((cls == 8 || (cls == 4 && WA.sm_level >= 2)) && getenv("C2G_W3_KC") == nullptr) WA.steps_per_check = 24;
      const int certk = WA.sm.safe == nullptr ? 0 : ((WA.sm.pack && WA.sm.octet && WA.sm.shift == 1) ? 2 : 1);
#define C2G_W3(O, F, T)                                                            \
  do {                                                                             \
    const int b3 = std::max(1, blocks * 256 / W3_NT);                                \
    if (certk == 2) k_walk3<O, F, T, 2><<<b3, W3_NT, 0, st>>>(P, WA);              \
    else if (certk == 1) k_walk3<O, F, T, 1><<<b3, W3_NT, 0, st>>>(P, WA);         \
    else k_walk3<O, F, T, 0><<<b3, W3_NT, 0, st>>>(P, WA);                         \
  } while (0)
      switch (variant) {
        case 0: C2G_W3(false, false, false); break;
        case 1: C2G_W3(false, false, true); break;
        case 2: C2G_W3(false, true, false); break;
        case 3: C2G_W3(false, true, true); break;
        case 4: C2G_W3(true, false, false); break;
        case 5: C2G_W3(true, false, true); break;
        case 6: C2G_W3(true, true, false); break;
        default: C2G_W3(true, true, true); break;
      }
#undef C2G_W3
    } else if (!walk_old) {
      // k_walk2 refills idle lanes before every step once `refill_min` lanes are idle (default 1)
      WA.refill_min = 1;
      if (const char* e = getenv("C2G_W2_REFILL")) WA.refill_min = std::max(1, std::min(32, atoi(e)));
      switch (variant) {
        case 0: k_walk2<false, false, false><<<blocks, 256, 0, st>>>(P, WA); break;
        case 1: k_walk2<false, false, true><<<blocks, 256, 0, st>>>(P, WA); break;
        case 2: k_walk2<false, true, false><<<blocks, 256, 0, st>>>(P, WA); break;
        case 3: k_walk2<false, true, true><<<blocks, 256, 0, st>>>(P, WA); break;
        case 4: k_walk2<true, false, false><<<blocks, 256, 0, st>>>(P, WA); break;
        case 5: k_walk2<true, false, true><<<blocks, 256, 0, st>>>(P, WA); break;
        case 6: k_walk2<true, true, false><<<blocks, 256, 0, st>>>(P, WA); break;
        default: k_walk2<true, true, true><<<blocks, 256, 0, st>>>(P, WA); break;
      }
    } else
    switch (variant) {
      case 0: k_walk<false, false, false><<<blocks, 256, 0, st>>>(P, WA); break;
      case 1: k_walk<false, false, true><<<blocks, 256, 0, st>>>(P, WA); break;
      case 2: k_walk<false, true, false><<<blocks, 256, 0, st>>>(P, WA); break;
      case 3: k_walk<false, true, true><<<blocks, 256, 0, st>>>(P, WA); break;
      case 4: k_walk<true, false, false><<<blocks, 256, 0, st>>>(P, WA); break;
      case 5: k_walk<true, false, true><<<blocks, 256, 0, st>>>(P, WA); break;
      case 6: k_walk<true, true, false><<<blocks, 256, 0, st>>>(P, WA); break;
      default: k_walk<true, true, true><<<blocks, 256, 0, st>>>(P, WA); break;
    }
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    if (!WA.count_dev) walked += count;  // otherwise counted on the device (C_WALKED)
    return C2G_OK;
  };
  // walkers over the stride-lat_s lattice of the slab (complete trajectories, nothing logged)
  auto walk_lattice = [&](long long count, int lat_s, const char* name) -> int {
    if (count <= 0) return C2G_OK;
    WA.list = nullptr; WA.stop = nullptr; WA.items = nullptr; WA.nitems = 0; WA.flat_base = 0; WA.count = count;
    WA.count_dev = nullptr; WA.base_dev = nullptr;
    WA.sm = SafeMap{nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; WA.sm_level = 0;
    WA.refill_min = lat_s == 1 ? refill_fine : refill_coarse;
    WA.steps_per_check = lat_s == 1 ? spc_fine : spc_coarse;
    WA.lat_s = lat_s; WA.lat_m1 = (n1 + lat_s - 1) / lat_s; WA.lat_m2 = (n2 + lat_s - 1) / lat_s;
    WA.next = nullptr; WA.nextcap = 0;
    const long long nwarps = (long long)wblocks * 8;
    WA.batch = (int)std::max<long long>(32, std::min<long long>(256, count / (nwarps * 4) / 32 * 32));
    return launch_walk(count, false, name);
  };
  // walkers over a segmented list: its entries are first copied, in segment order, to the end of the dense list
  auto walk_segments = [&](const int* seglist, const int* segcnt, int nseg, int segcap, long long count, bool fix, int* next,
                           const SafeMap& sm, int sm_level, const char* name) -> int {
    if (async) {  // lengths and offsets stay on the device; a full dense list raises err = 5 there
      const int nblk = c2g_blocks_for(nseg, 256);
      if ((size_t)nblk > blksumcap) {
        blksumcap = (size_t)nblk;
        C2G_CUDA(ctx, b_blksum.alloc(ctx, sizeof(int2) * blksumcap));
      }
      ctx->prof_begin("bader_items");
      k_items_count<<<nblk, 256, 0, st>>>(nseg, 64, segcnt, b_blksum.as<int2>());
      k_items_scan<<<1, 256, 0, st>>>(nblk, b_blksum.as<int2>(), cnt + C_NITEMS, cnt + C_NENT, cnt + C_DOFF, dcap, cnt + C_ERR);
      k_items_write<<<nblk, 256, 0, st>>>(nseg, segcap, 64, segcnt, b_blksum.as<int2>(), seglist, nullptr, b_dlist.as<int>(), 0,
                                          cnt + C_DOFF, cnt + C_NENT);
      ctx->prof_end(3);
      C2G_KERNEL_CHECK(ctx);
      WA.list = b_dlist.as<int>(); WA.stop = b_stop.as<int>(); WA.flat_base = 0; WA.count = 0;
      WA.count_dev = cnt + C_NENT; WA.base_dev = cnt + C_DOFF;
      WA.items = nullptr; WA.nitems = 0; WA.nitems_dev = cnt + C_NITEMS;
      WA.sm = sm; WA.sm_level = sm_level; WA.lat_s = 1; WA.lat_m1 = n1; WA.lat_m2 = n2;
      WA.next = next; WA.nextcap = (int)std::min<long long>(listcap, 0x7fffffff);
      WA.batch = 64;
      return launch_walk(0, fix, name);
    }
    WA.count_dev = nullptr; WA.base_dev = nullptr;
    if (count <= 0) return C2G_OK;
    if (doff + count > dcap) return ctx->fail(C2G_ERR_OVERFLOW, "walker list overflow");
    int batch = count < 64ll * wblocks * 8 ? 32 : 64;
    if (const char* e = getenv("C2G_BATCH")) batch = std::max(32, std::min(1024, atoi(e) / 32 * 32));
    const size_t maxitems = (size_t)nseg + (size_t)(count / batch) + 1;
    if (maxitems > itemcap) {
      itemcap = maxitems + maxitems / 4;
      C2G_CUDA(ctx, b_items.alloc(ctx, sizeof(int2) * itemcap));
    }
    const int nblk = c2g_blocks_for(nseg, 256);
    if ((size_t)nblk > blksumcap) {
      blksumcap = (size_t)nblk;
      C2G_CUDA(ctx, b_blksum.alloc(ctx, sizeof(int2) * blksumcap));
    }
    ctx->prof_begin("bader_items");
    k_items_count<<<nblk, 256, 0, st>>>(nseg, batch, segcnt, b_blksum.as<int2>());
    k_items_scan<<<1, 256, 0, st>>>(nblk, b_blksum.as<int2>(), cnt + 10, cnt + 12, nullptr, 0, nullptr);
    k_items_write<<<nblk, 256, 0, st>>>(nseg, segcap, batch, segcnt, b_blksum.as<int2>(), seglist, b_items.as<int2>(),
                                        b_dlist.as<int>(), doff, nullptr, nullptr);
    ctx->prof_end(3);
    C2G_KERNEL_CHECK(ctx);
    WA.list = b_dlist.as<int>(); WA.stop = b_stop.as<int>(); WA.flat_base = doff; WA.count = count;
    WA.items = b_items.as<int2>(); WA.nitems = (int)std::min<size_t>(maxitems, 0x7fffffff); WA.nitems_dev = cnt + 10;
    WA.sm = sm; WA.sm_level = sm_level; WA.lat_s = 1; WA.lat_m1 = n1; WA.lat_m2 = n2;
    WA.next = next; WA.nextcap = (int)std::min<long long>(listcap, 0x7fffffff);
    WA.batch = batch;
    const bool fine = !fix && sm_level == 0;
    WA.refill_min = fine ? refill_fine : refill_coarse;
    WA.steps_per_check = fine ? spc_fine : spc_coarse;
    doff += count;
    return launch_walk(count, fix, name);
  };
  // walkers over a flat list (edge-fix passes): copied to the end of the dense list as well
  auto walk_flat = [&](const int* list, long long count, bool fix, int* next, const SafeMap& sm, int sm_level, const char* name) -> int {
    if (async) {  // the list length is cnt[C_FLATN] (k_fix_begin)
      ctx->prof_begin("bader_items");
      k_items_flat<<<ctx->nsm * 4, 256, 0, st>>>(list, cnt, dcap, b_dlist.as<int>());
      k_items_flat_end<<<1, 1, 0, st>>>(cnt, dcap);
      ctx->prof_end(2);
      C2G_KERNEL_CHECK(ctx);
      WA.list = b_dlist.as<int>(); WA.stop = b_stop.as<int>(); WA.flat_base = 0; WA.count = 0;
      WA.count_dev = cnt + C_NENT; WA.base_dev = cnt + C_DOFF;
      WA.items = nullptr; WA.nitems = 0;
      WA.sm = sm; WA.sm_level = sm_level; WA.lat_s = 1; WA.lat_m1 = n1; WA.lat_m2 = n2;
      WA.next = next; WA.nextcap = (int)std::min<long long>(listcap, 0x7fffffff);
      WA.batch = 32;
      return launch_walk(0, fix, name);
    }
    WA.count_dev = nullptr; WA.base_dev = nullptr;
    if (count <= 0) return C2G_OK;
    if (doff + count > dcap) return ctx->fail(C2G_ERR_OVERFLOW, "walker list overflow");
    C2G_CUDA(ctx, cudaMemcpyAsync(b_dlist.as<int>() + doff, list, sizeof(int) * (size_t)count, cudaMemcpyDeviceToDevice, st));
    WA.list = b_dlist.as<int>(); WA.stop = b_stop.as<int>(); WA.flat_base = doff; WA.count = count;
    WA.items = nullptr; WA.nitems = 0;
    WA.sm = sm; WA.sm_level = sm_level; WA.lat_s = 1; WA.lat_m1 = n1; WA.lat_m2 = n2;
    WA.next = next; WA.nextcap = (int)std::min<long long>(listcap, 0x7fffffff);
    WA.batch = 32;
    WA.refill_min = refill_coarse;
    WA.steps_per_check = spc_coarse;
    doff += count;
    return launch_walk(count, fix, name);
  };
  const SafeMap nosafe{nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0};

  int rc;
  if (algo == C2G_BADER_EXACT) {
    if ((rc = walk_lattice(nnl, 1, "bader_walk_all")) != C2G_OK) return rc;
    if ((rc = drain(false)) != C2G_OK) return rc;
  } else {
    // top level: the stride-L0 lattice walks complete trajectories
    {
      const long long m = (long long)((n1 + L0 - 1) / L0) * ((n2 + L0 - 1) / L0) * ((S.nzl + L0 - 1) / L0);
      if ((rc = walk_lattice(m, L0, "bader_walk_top")) != C2G_OK) return rc;
      if ((rc = drain(false)) != C2G_OK) return rc;
    }
    const long long ntop = walked;
    const bool use_safe = getenv("C2G_NO_EARLY_STOP") == nullptr;
    // Early stops: a walk may end where a certificate says that the neighbourhood is uniformly labelled (the
    // reference's known==2 points).  Certificates rest on labels that may still change (wrong fills), so every
    // stop is logged, every label change voids the certificates around it, and the walkers that relied on a
    // voided certificate walk again (k_requeue).  safe_maxs: coarsest cube stride whose certificates are used.
    int safe_maxs = 16;
    if (const char* e = getenv("C2G_SAFE_MAXS")) safe_maxs = atoi(e);
    // certificate: 1 = octets; at the stride-2 level with the per-point 5x5x5 flags of k_pack5 (the margin at which
    // the reference's refine_edge walks stop), 2 = cube + 26 neighbour cubes (coarser 5x5x5 rule, experiments),
    // 3 = bare octets at every level (3x3x3 at stride 2: leaves a few wrong labels on ridge-running trajectories,
    // 3 of 7.1e6 points in tests/sized_cases.py hetero192; kept for measurements only)
    int cert = 1;
    if (const char* e = getenv("C2G_CERT")) cert = atoi(e);
    unsigned long long steps_prev = 0;
    auto report = [&](const char* what, long long count) {
      if (!verbose) return;
      unsigned long long sn = 0;
      memcpy(&sn, hcnt + 4, sizeof(sn));
      fprintf(stderr, "[c2g bader] %-16s walkers %12lld steps %14llu (%.2f/walker)\n", what, count, sn - steps_prev,
              count ? (double)(sn - steps_prev) / count : 0.0);
      steps_prev = sn;
    };
    report("bader_walk_top", ntop);
    static const char* wname[MAXLEV] = {"bader_walk_l1", "bader_walk_l2", "bader_walk_l4", "bader_walk_l8", "bader_walk_l16"};
    int* seglist = b_list[0].as<int>();
    int* segcnt = b_segcnt.as<int>();
    WA.nmaps = nlev;  // maps[i] = certificate map of level i (stride 2<<i), if any
    for (int i = nlev - 1; i >= 0; i--) {
      const Level& L = lev[i];
      if ((rc = exchange_halos(ctx, lbuf, plane, S, 1)) != C2G_OK) return rc;
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));
      SafeMap sm = nosafe;
      const int t1 = (L.c1 + CLS_TX - 1) / CLS_TX, t2 = (L.c2 + CLS_TY - 1) / CLS_TY, t3 = (L.c3 + CLS_TZ - 1) / CLS_TZ;
      const int ntile = cls_nblocks(t1, t2, std::max(1, t3));
      if (nnl > 0) {
        ctx->prof_begin(i == 0 ? "bader_classify2" : "bader_classify");
        if (i == 0)
          k_classify<false><<<ntile, 256, 0, st>>>(n1, n2, n3, S, L.s, cls_geom(n1, n2, S.nzl, L.s), lbuf, b_cubemax[i].as<unsigned char>(), b_uni[i].as<int>(),
                                                   seglist, segcnt, cnt + 1);
        else
          k_classify<true><<<ntile, 256, 0, st>>>(n1, n2, n3, S, L.s, cls_geom(n1, n2, S.nzl, L.s), lbuf, b_cubemax[i].as<unsigned char>(), b_uni[i].as<int>(),
                                                  seglist, segcnt, cnt + 1);
        ctx->prof_end();
        C2G_KERNEL_CHECK(ctx);
        if (use_safe && L.s <= safe_maxs) {
          const dim3 sg((L.c1 + 31) / 32, (L.c2 + TY - 1) / TY, (L.c3 + MZC - 1) / MZC);
          ctx->prof_begin(i == 0 ? "bader_safe2" : "bader_safe");
          if (cert == 1 || cert == 3) {
            const int px = (n1 % L.s) == 0, py = (n2 % L.s) == 0;
            const int pz = (S.periodic && (n3 % L.s) == 0) ? 1 : 0;
            // stride-2 level: an octet only covers the 3x3x3 neighbourhood of its points; the walkers need 5x5x5 (the
            // reference's margin), so the octets go to a scratch array and k_pack5 adds the per-point flags.  Labels
            // that do not fit the packed entry (> 8.4e6 maxima: noise): no certificates at this level.
            const bool pack = i == 0 && cert == 1;
            if (pack && ncand > CERT_LMASK) {
              ctx->prof_end(0);
            } else {
              int* octets = pack ? b_list[2].as<int>() : b_safe[i].as<int>();  // the flat lists are idle until the edge fix
              k_vsafe<<<sg, 256, 0, st>>>(L.c1, L.c2, L.c3, px, py, pz, (S.nzl % L.s) == 0 ? 1 : 0, b_uni[i].as<int>(), octets);
              if (pack) {
                k_pack5<<<sg, 256, 0, st>>>(L.c1, L.c2, L.c3, px, py, pz, octets, b_safe[i].as<int>());
              }
              sm = SafeMap{b_safe[i].as<int>(), i + 1, L.c1, L.c2, L.c3, S.zlo, S.nzl, 1, pz, (1 << (i + 1)) >> 1, pack ? 1 : 0};
              ctx->prof_end(pack ? 2 : 1);
            }
          } else {
            k_safe<<<sg, 256, 0, st>>>(L.c1, L.c2, L.c3, b_uni[i].as<int>(), b_safe[i].as<int>());
            sm = SafeMap{b_safe[i].as<int>(), i + 1, L.c1, L.c2, L.c3, S.zlo, S.nzl, 0, 0, 0, 0};
            ctx->prof_end();
          }
          C2G_KERNEL_CHECK(ctx);
        }
      }
      WA.maps[i] = sm;
      long long nw = 1;  // device-side bookkeeping: the count stays in cnt[C_NLIST] / segcnt
      if (!async) {
        C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
        C2G_CUDA(ctx, cudaStreamSynchronize(st));
        nw = hcnt[1];
      }
      if ((rc = walk_segments(seglist, segcnt, ntile, CLS_SEGCAP, nw, false, nullptr, sm, i, wname[i])) != C2G_OK) return rc;
      if ((rc = drain(false)) != C2G_OK) return rc;
      report(wname[i], nw);
    }
    const SafeMap fixsafe = WA.maps[0];
    // ---- materialise the last level and find the filled points with a foreign 26-neighbour ----
    FillArgs FA;
    memset(&FA, 0, sizeof(FA));
    FA.n1 = n1; FA.n2 = n2; FA.n3 = n3; FA.S = S; FA.lbuf = lbuf; FA.uni2 = b_uni[0].as<int>();
    FA.c1 = lev[0].c1; FA.c2 = lev[0].c2; FA.err = cnt + 3;
    FA.listcap = (int)std::min<long long>(listcap, 0x7fffffff);
    FA.segcnt = segcnt;
    C2G_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0, sizeof(int), st));   // [1] segmented edge points of the first pass
    C2G_CUDA(ctx, cudaMemsetAsync(cnt + 7, 0, sizeof(int), st));   // [7] flat list being built
    auto edge_faces = [&](int* out, int* nout) -> int {  // multi-GPU: the two face planes against fresh halos
      FillArgs FB = FA;
      FB.list = out; FB.nlist = nout; FB.skip_lo = FB.skip_hi = 0;
      const dim3 eg(fe_gx, fe_gy, 1);
      ctx->prof_begin("bader_edge_faces");
      FB.za = S.zlo; FB.zb = S.zlo + 1;
      k_fill_edge<false, false><<<eg, 256, 0, st>>>(FB);
      if (S.nzl > 1) {
        FB.za = S.zhi - 1; FB.zb = S.zhi;
        k_fill_edge<false, false><<<eg, 256, 0, st>>>(FB);
      }
      ctx->prof_end(S.nzl > 1 ? 2 : 1);
      C2G_KERNEL_CHECK(ctx);
      return C2G_OK;
    };
    int cur = 1;  // flat list being consumed next; the other one receives the claims
    const int nfeblk = fe_nblocks(fe_gx, fe_gy, fe_gz);
    int nfeseg = nfeblk, fesegcap = FE_SEGCAP;
    FA.g1 = fe_gx; FA.g2 = fe_gy; FA.g3 = fe_gz;
    const bool vertex_pass = fixsafe.safe && fixsafe.octet && (S.nzl % 2 == 0 || S.periodic) && getenv("C2G_FILL_OLD") == nullptr;
    if (nnl > 0 && vertex_pass) {
      FillVArgs FV;
      FV.n1 = n1; FV.n2 = n2; FV.n3 = n3;
      FV.c1 = lev[0].c1; FV.c2 = lev[0].c2; FV.c3 = lev[0].c3;
      FV.nzl = S.nzl; FV.periodic = S.periodic; FV.zlo = S.zlo;
      FV.nvz = FV.c3 + (S.periodic ? 0 : 1);
      FV.label = res->d_label; FV.uni2 = b_uni[0].as<int>(); FV.vsafe = fixsafe.safe; FV.pack = fixsafe.pack;
      FV.list = seglist; FV.nlist = cnt + 1; FV.segcnt = segcnt;
      FV.g1 = (FV.c1 + 31) / 32; FV.g2 = (FV.c2 + 7) / 8; FV.g3 = std::max(1, (FV.nvz + VZC - 1) / VZC);
      nfeseg = fv_nblocks(FV.g1, FV.g2, FV.g3); fesegcap = FV_SEGCAP;
      ctx->prof_begin("bader_fill_edge");
      k_fill_edge_v<<<nfeseg, 256, 0, st>>>(FV);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    } else if (nnl > 0) {
      FA.list = seglist; FA.nlist = cnt + 1;
      FA.za = S.zlo; FA.zb = S.zhi; FA.skip_lo = FA.skip_hi = (G > 1) ? 1 : 0;
      ctx->prof_begin("bader_fill_edge");
      k_fill_edge<true, true><<<nfeblk, 256, 0, st>>>(FA);
      ctx->prof_end();
      C2G_KERNEL_CHECK(ctx);
    }
    if (G > 1) {
      if ((rc = exchange_halos(ctx, lbuf, plane, S, 3)) != C2G_OK) return rc;
      if (nnl > 0 && (rc = edge_faces(b_list[cur].as<int>(), cnt + 7)) != C2G_OK) return rc;
    }
    // ---- edge fix until no filled point (on any rank) is adjacent to a different label ----
    // pass 1 consumes the segmented list (+ the face points, multi-GPU); every pass writes into the other flat
    // list the filled neighbours of the points that changed and the walkers whose certificate was voided.
    bool first = true;
    int ninval_seen = 0;
    if (async) {
      // Passes 1 and 2 are enqueued without looking at the counters (an empty pass costs a few empty launches); from then
      // on the host reads the length of the list the last pass built and stops when it is empty on every rank.
      for (int pass = 0;; pass++) {
        if (pass >= 2) {
          C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
          C2G_CUDA(ctx, cudaStreamSynchronize(st));
          // a rank that failed must not leave the others waiting in the next collective: the error flag travels with
          // the "any work left" reduction (2 = some rank failed) and every rank returns
          int any = hcnt[3] != 0 ? 2 : (hcnt[C_NNEXT] > 0 ? 1 : 0);
          if (G > 1) {
            C2G_CUDA(ctx, cudaMemcpyAsync(cnt + 6, &any, sizeof(int), cudaMemcpyHostToDevice, st));
            C2G_NCCL(ctx, ncclAllReduce(cnt + 6, cnt + 6, 1, ncclInt32, ncclMax, comm, st));
            C2G_CUDA(ctx, cudaMemcpyAsync(&any, cnt + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
            C2G_CUDA(ctx, cudaStreamSynchronize(st));
          }
          if ((rc = check_err()) != C2G_OK) return rc;
          if (any == 2) return ctx->fail(C2G_ERR_STATE, "c2g_bader_assign: another rank failed in the edge refinement");
          if (any == 0) break;
        }
        int* in = b_list[cur].as<int>();
        int* out = b_list[cur == 1 ? 2 : 1].as<int>();
        k_fix_begin<<<1, 1, 0, st>>>(cnt, first ? 1 : 0);
        ctx->launches += 1;
        if (first) {
          if ((rc = walk_segments(seglist, segcnt, nfeseg, fesegcap, 1, true, out, fixsafe, 0, "bader_walk_fix")) != C2G_OK) return rc;
          if ((rc = drain(true)) != C2G_OK) return rc;
        }
        if ((rc = walk_flat(in, 1, true, out, fixsafe, 0, "bader_walk_fix")) != C2G_OK) return rc;
        if ((rc = drain(true)) != C2G_OK) return rc;
        // walkers whose certificate has been voided since the last pass go back into the list (gated on the device)
        ctx->prof_begin("bader_requeue");
        k_requeue<<<ctx->nsm * 8, 256, 0, st>>>(0, b_dlist.as<int>(), b_stop.as<int>(), WA, out, cnt + C_NNEXT,
                                                 (int)std::min<long long>(listcap, 0x7fffffff), cnt);
        k_requeue_end<<<1, 1, 0, st>>>(cnt);
        ctx->prof_end(2);
        C2G_KERNEL_CHECK(ctx);
        if (G > 1) {
          if ((rc = exchange_halos(ctx, lbuf, plane, S, 3)) != C2G_OK) return rc;
          if (nnl > 0 && (rc = edge_faces(out, cnt + 7)) != C2G_OK) return rc;
        }
        cur = (cur == 1) ? 2 : 1;
        first = false;
        if (pass > 100000) return ctx->fail(C2G_ERR_STATE, "edge refinement did not converge");
      }
      fixpasses = hcnt[C_NPASS] + 1;
      fixpts = hcnt[C_FIXPTS];
      walked += hcnt[C_WALKED];
      noverflow_total = hcnt[C_OVERTOT];
    } else
    for (;;) {
      C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
      C2G_CUDA(ctx, cudaStreamSynchronize(st));
      const int nseg1 = first ? hcnt[1] : 0, nflat = hcnt[7];
      int any = hcnt[3] != 0 ? 2 : (nseg1 + nflat > 0 ? 1 : 0);  // 2: this rank failed -- agreed on by every rank below
      if (G > 1) {
        C2G_CUDA(ctx, cudaMemcpyAsync(cnt + 6, &any, sizeof(int), cudaMemcpyHostToDevice, st));
        C2G_NCCL(ctx, ncclAllReduce(cnt + 6, cnt + 6, 1, ncclInt32, ncclMax, comm, st));
        C2G_CUDA(ctx, cudaMemcpyAsync(&any, cnt + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
        C2G_CUDA(ctx, cudaStreamSynchronize(st));
      }
      if ((rc = check_err()) != C2G_OK) return rc;
      if (any == 2) return ctx->fail(C2G_ERR_STATE, "c2g_bader_assign: another rank failed in the edge refinement");
      fixpasses++;
      if (any == 0) break;
      fixpts += nseg1 + nflat;
      int* in = b_list[cur].as<int>();
      int* out = b_list[cur == 1 ? 2 : 1].as<int>();
      C2G_CUDA(ctx, cudaMemsetAsync(cnt + 7, 0, sizeof(int), st));
      if (nseg1 && (rc = walk_segments(seglist, segcnt, nfeseg, fesegcap, nseg1, true, out, fixsafe, 0, "bader_walk_fix")) != C2G_OK) return rc;
      if (nflat && (rc = walk_flat(in, nflat, true, out, fixsafe, 0, "bader_walk_fix")) != C2G_OK) return rc;
      if ((rc = drain(true)) != C2G_OK) return rc;
      report("bader_walk_fix", nseg1 + nflat);
      if (hcnt[11] != ninval_seen && doff > 0) {  // certificates were voided: re-queue the walkers that relied on them
        ninval_seen = hcnt[11];
        ctx->prof_begin("bader_requeue");
        k_requeue<<<ctx->nsm * 8, 256, 0, st>>>(doff, b_dlist.as<int>(), b_stop.as<int>(), WA, out, cnt + 7,
                                                 (int)std::min<long long>(listcap, 0x7fffffff), nullptr);
        ctx->prof_end();
        C2G_KERNEL_CHECK(ctx);
      }
      if (G > 1) {
        if ((rc = exchange_halos(ctx, lbuf, plane, S, 3)) != C2G_OK) return rc;
        if (nnl > 0 && (rc = edge_faces(out, cnt + 7)) != C2G_OK) return rc;
      }
      cur = (cur == 1) ? 2 : 1;
      first = false;
      if (fixpasses > 100000) return ctx->fail(C2G_ERR_STATE, "edge refinement did not converge");
    }
    res->stats[7] = hcnt[11];
  }

  // ---- maxima actually reached (on any rank), output order ----
  if (G > 1) C2G_NCCL(ctx, ncclAllReduce(reached, reached, ncand, ncclUint8, ncclMax, comm, st));
  std::vector<unsigned char> hreached(ncand);
  C2G_CUDA(ctx, cudaMemcpyAsync(hreached.data(), reached, ncand, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(hcnt, cnt, C_NCNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  if ((rc = check_err()) != C2G_OK) return rc;
  std::vector<int> cand2out(ncand, -1);
  int nmax = 0;
  for (int i = 0; i < ncand; i++)
    if (hreached[i]) cand2out[i] = nmax++;
  res->nmax = nmax;
  res->max_lin.resize(nmax);
  for (int i = 0; i < ncand; i++)
    if (cand2out[i] >= 0) res->max_lin[cand2out[i]] = cand[i];
  const int nblk = ctx->nsm * 8;
  if (nmax < ncand && nnl > 0) {  // rare: a candidate no trajectory ends on; renumber
    DevBuf b_c2o;
    C2G_CUDA(ctx, b_c2o.alloc(ctx, sizeof(int) * ncand));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_c2o.p, cand2out.data(), sizeof(int) * ncand, cudaMemcpyHostToDevice, st));
    ctx->prof_begin("bader_compact");
    k_permute_labels<<<nblk, 256, 0, st>>>(nnl, res->d_label, b_c2o.as<int>());
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }

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
  res->stats[6] = L0;
  ctx->prof_collect();
  guard.ok = true;
  *nmax_out = nmax;
  *res_out = res;
  return C2G_OK;
}
