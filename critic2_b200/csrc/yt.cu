// yt.cu -- Yu-Trinkle basin weights and weighted integration on sm_100a.
//
// Replaces yt_integrate (critic2 src/yt@proc.f90:77-211), yt_weights (:476-499) and the per-basin
// weighted sums of intgrid_fields (src/integration@proc.f90:1208-1218, :1289-1299).
//
// The reference sorts the grid by density with a serial quicksort and sweeps it once in decreasing
// order, then performs one full downhill sweep PER BASIN to build its weight field.  Here:
//  * "j is higher than i" is the rank comparison of a stable sort, (rho_j, j) > (rho_i, i); it needs
//    no sort at all and equals the reference's permutation whenever the data has no ties
//    (SURVEY.md 7.2-5).
//  * k_scan      : per point, bit mask of higher Voronoi neighbours, csum = sum of fluxes
//                  chi = max(area*(rho_j-rho_i), vsmall) in stencil order (:115-125), steepest
//                  higher neighbour (pointer), local maxima (:129).
//  * k_jump      : pointer jumping to the terminal maximum = candidate basin of every point.
//  * k_seed/bfs  : interatomic-surface (IAS) points = downward closure of the points whose higher
//                  neighbours do not all share one candidate (:170-186); the rest are interior.
//  * k_kahn      : topological levels of the IAS flux graph (persistent cooperative kernel).
//  * integration : ONE uphill (adjoint) sweep per call for all properties instead of one downhill
//                  sweep per basin:  y(j) = f(j) + sum_{IAS i below j} fnear(i->j) y(i);
//                  integral over basin b = sum of y over the interior points of b.  Pull-based and
//                  level-ordered, hence atomics-free and run-to-run deterministic up to the final
//                  per-basin reduction.
//  * weights     : the same levels swept downhill for a single basin (yt_weights).
// Flux fractions fnear = chi/max(csum,vsmall) are recomputed from rho (a stencil) instead of being
// stored as the reference's dense inear/fnear(nvec,nn) arrays.
#include "common.cuh"
#include "group.h"

#include <cub/device/device_radix_sort.cuh>

#include <cooperative_groups.h>

#include <algorithm>

namespace cg = cooperative_groups;

namespace {

constexpr int MAXVEC = 30;
constexpr double VSMALL = 1e-80;  // param.F90:27

struct YtParams {
  int n1, n2, n3;
  int nvec;
  int vec[3 * MAXVEC];
  int opp[MAXVEC];   // index of -vec(k)
  double area[MAXVEC];
};

__device__ __forceinline__ int imod(int a, int n) {
  int r = a % n;
  return r < 0 ? r + n : r;
}
__device__ __forceinline__ bool higher(double rj, int j, double ri, int i) { return (rj > ri) || (rj == ri && j > i); }

struct Pt { int x, y, z; };
__device__ __forceinline__ Pt unlin(const YtParams& P, int i) {
  Pt p;
  p.x = i % P.n1;
  const int t = i / P.n1;
  p.y = t % P.n2;
  p.z = t / P.n2;
  return p;
}
__device__ __forceinline__ int nbr(const YtParams& P, const Pt& p, int k) {
  const int x = imod(p.x + P.vec[3 * k], P.n1), y = imod(p.y + P.vec[3 * k + 1], P.n2), z = imod(p.z + P.vec[3 * k + 2], P.n3);
  return x + P.n1 * (y + P.n2 * z);
}

__global__ void __launch_bounds__(256) k_scan(const __grid_constant__ YtParams P, const double* __restrict__ rho,
                                              unsigned* __restrict__ mask, double* __restrict__ csum, int* __restrict__ up,
                                              int* __restrict__ maxlist, int* __restrict__ nmax, int maxcap) {
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const Pt p = unlin(P, (int)i);
  const double ri = __ldg(rho + i);
  unsigned m = 0;
  double cs = 0.0;
  int best = (int)i;
  double rbest = ri;
  for (int k = 0; k < P.nvec; k++) {
    const int j = nbr(P, p, k);
    const double rj = __ldg(rho + j);
    if (higher(rj, j, ri, (int)i)) {
      m |= 1u << k;
      cs = cs + fmax(P.area[k] * (rj - ri), VSMALL);
      if (higher(rj, j, rbest, best)) { rbest = rj; best = j; }
    }
  }
  mask[i] = m;
  csum[i] = cs;
  up[i] = best;
  if (m == 0) {
    const int s = atomicAdd(nmax, 1);
    if (s < maxcap) maxlist[s] = (int)i;
  }
}

__global__ void __launch_bounds__(256) k_jump(long long nn, int* __restrict__ up, int* __restrict__ changed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  int u = up[i];
  int v = up[u];
  if (v == u) return;
  // a few hops per launch
  for (int it = 0; it < 8; it++) {
    u = v;
    v = up[u];
    if (v == u) break;
  }
  up[i] = u;
  if (up[u] != u) *changed = 1;
}

// terminal maximum (linear id) -> index in the ordered maxima list, in place
__global__ void __launch_bounds__(256) k_cand_index(long long nn, int* __restrict__ up, const int* __restrict__ hk,
                                                    const int* __restrict__ hv, unsigned hmask) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  int last_t = -1, last_o = -1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int t = up[i];
    if (t != last_t) {
      unsigned s = ((unsigned)t * 2654435761u) & hmask;
      while (hk[s] != t) s = (s + 1) & hmask;
      last_t = t;
      last_o = hv[s];
    }
    up[i] = last_o;
  }
}

// Seeds of the IAS closure (yt@proc.f90:170-173): a point below at least one other point is an IAS point when the
// basins of its higher neighbours differ or one of them has none.  cand = index of the maximum reached by steepest
// ascent; map (may be null = every maximum is its own basin) = maximum -> basin id as the reference's sweep writes it
// into ibasin (identify_atom / are_lclose merge maxima into one attractor, a DISCARDed maximum keeps 0, :129-168).
__global__ void __launch_bounds__(256) k_seed(const __grid_constant__ YtParams P, const unsigned* __restrict__ mask,
                                              const int* __restrict__ cand, const int* __restrict__ map,
                                              unsigned char* __restrict__ ias, int* __restrict__ queue, int* __restrict__ qtail) {
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const unsigned m = mask[i];
  if (m == 0) return;
  const Pt p = unlin(P, (int)i);
  const int c0 = map ? __ldg(map + cand[i]) : cand[i] + 1;  // basin of the steepest higher neighbour
  bool seed = c0 == 0;
  for (int k = 0; k < P.nvec; k++)
    if (m & (1u << k)) {
      const int cj = cand[nbr(P, p, k)];
      seed = seed || ((map ? __ldg(map + cj) : cj + 1) != c0);
    }
  if (seed) {
    ias[i] = 1;
    queue[atomicAdd(qtail, 1)] = (int)i;
  }
}

// downward closure: persistent cooperative BFS.  In: ctl[0] = number of seeds (already in queue[]).  Out: ctl[0] =
// queue length, ctl[1] = number of levels.  Like k_kahn: the points found while a level is processed are appended
// behind its end through a per-level counter, so one grid barrier per level is enough.
__global__ void __launch_bounds__(256) k_bfs(const __grid_constant__ YtParams P, const unsigned* __restrict__ mask,
                                             unsigned char* __restrict__ ias, int* __restrict__ queue, int* __restrict__ ctl,
                                             int* __restrict__ cnt, int maxlvl) {
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_next;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const unsigned full = (P.nvec >= 32) ? 0xffffffffu : ((1u << P.nvec) - 1u);
  int lo = 0, hi = ctl[0], levels = 0;
  while (lo < hi) {
    int* next = cnt + min(levels + 1, maxlvl);
    for (int q = lo + tid; q < hi; q += nth) {
      const int i = queue[q];
      const Pt p = unlin(P, i);
      unsigned lower = (~mask[i]) & full;
      while (lower) {
        const int k = __ffs(lower) - 1;
        lower &= lower - 1;
        const int j = nbr(P, p, k);
        if (!ias[j]) {
          // byte flags: claim through a 32-bit atomic on the containing word
          unsigned* w = (unsigned*)(ias + (j & ~3));
          const unsigned bit = 1u << (8 * (j & 3));
          const unsigned old = atomicOr(w, bit);
          if (!(old & bit)) queue[hi + atomicAdd(next, 1)] = j;
        }
      }
    }
    grid.sync();
    if (threadIdx.x == 0) s_next = *((volatile int*)next);
    __syncthreads();
    lo = hi;
    hi += s_next;
    levels++;
    __syncthreads();
  }
  if (tid == 0) { ctl[0] = hi; ctl[1] = levels; }
}

// in-degree of every IAS point = number of IAS points below it among its neighbours
__global__ void __launch_bounds__(256) k_indeg(const __grid_constant__ YtParams P, const unsigned* __restrict__ mask,
                                               const unsigned char* __restrict__ ias, const int* __restrict__ iaslist,
                                               int nias, unsigned char* __restrict__ indeg, int* __restrict__ order,
                                               int* __restrict__ otail) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nias) return;
  const int i = iaslist[q];
  const Pt p = unlin(P, i);
  const unsigned full = (P.nvec >= 32) ? 0xffffffffu : ((1u << P.nvec) - 1u);
  unsigned lower = (~mask[i]) & full;
  int d = 0;
  while (lower) {
    const int k = __ffs(lower) - 1;
    lower &= lower - 1;
    d += ias[nbr(P, p, k)] ? 1 : 0;
  }
  indeg[i] = (unsigned char)d;
  if (d == 0) order[atomicAdd(otail, 1)] = i;
}

// Kahn levels, uphill.  In: ctl[0] = number of in-degree-0 points (level 0, already in order[]).  Out: ctl[0] = points
// ordered, ctl[1] = number of levels; lvl[L] = start of level L in order[].  The points released while level L is
// processed are appended behind the end of level L at positions handed out by a PER-LEVEL counter cnt[L+1]
// (zeroed by the caller), so the end of level L+1 is known after ONE grid barrier per level: nobody touches
// cnt[L+1] once level L is done.
__global__ void __launch_bounds__(256) k_kahn(const __grid_constant__ YtParams P, const unsigned* __restrict__ mask,
                                               const unsigned char* __restrict__ ias, unsigned char* __restrict__ indeg,
                                               int* __restrict__ order, int* __restrict__ ctl, int* __restrict__ lvl,
                                               int* __restrict__ cnt, int maxlvl) {
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_next;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  int lo = 0, hi = ctl[0], levels = 0;
  while (lo < hi) {
    if (tid == 0 && levels < maxlvl) lvl[levels] = lo;
    int* next = cnt + min(levels + 1, maxlvl);
    for (int q = lo + tid; q < hi; q += nth) {
      const int i = order[q];
      const Pt p = unlin(P, i);
      unsigned hm = mask[i];
      while (hm) {
        const int k = __ffs(hm) - 1;
        hm &= hm - 1;
        const int j = nbr(P, p, k);
        if (ias[j]) {
          unsigned* w = (unsigned*)(indeg + (j & ~3));
          const int sh = 8 * (j & 3);
          const unsigned old = atomicSub(w, 1u << sh);
          if (((old >> sh) & 0xffu) == 1u) order[hi + atomicAdd(next, 1)] = j;
        }
      }
    }
    grid.sync();
    // one L2 read per block, not one per thread (1.5e5 threads reading one address serialise in its L2 slice)
    if (threadIdx.x == 0) s_next = *((volatile int*)next);
    __syncthreads();
    lo = hi;
    hi += s_next;
    levels++;
    __syncthreads();
  }
  if (tid == 0) {
    ctl[0] = hi;
    ctl[1] = levels;
    if (levels < maxlvl) lvl[levels] = lo;
  }
}

// Kahn appends the points of a level in the order in which the atomics happen to land, i.e. scattered over the
// whole grid: the sweeps then pull ~1 KB through DRAM per point (ncu: 33.6 GB for 3.4e7 IAS points at 512^3).  The
// order INSIDE a level is free, so every level is re-sorted by linear grid index: key = (level << 32) | point, one
// radix sort of the whole list (CUB, a plain library primitive like cuFFT), and a warp's 32 points become
// neighbours along x that share the sectors of rho, csum, mask, ias and y.
__global__ void __launch_bounds__(256) k_level_keys(int nias, int nlevels, const int* __restrict__ lvl, const int* __restrict__ order,
                                                    unsigned long long* __restrict__ keys) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nias) return;
  int lo = 0, hi = nlevels;  // the level L with lvl[L] <= q < lvl[L+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(lvl + mid) <= q) lo = mid; else hi = mid;
  }
  keys[q] = ((unsigned long long)lo << 32) | (unsigned)order[q];
}
__global__ void __launch_bounds__(256) k_level_unkey(int nias, const unsigned long long* __restrict__ keys, int* __restrict__ order) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nias) order[q] = (int)(keys[q] & 0xffffffffull);
}

// flux fraction pulled by j from the IAS point i = j + vec(k) lying below it: fnear(i->j)
__device__ __forceinline__ double flux_up(const YtParams& P, int k, double rj, double ri, double csum_i) {
  return fmax(P.area[P.opp[k]] * (rj - ri), VSMALL) / fmax(csum_i, VSMALL);
}

// adjoint (uphill) sweep over the IAS levels for NP properties (+ volume in slot NP).
// y arrays: y[p*nn + i].  f pointers may be null only for the volume slot.
template <int NP>
__global__ void __launch_bounds__(256) k_sweep_up(const __grid_constant__ YtParams P, const double* __restrict__ rho,
                                                  const unsigned* __restrict__ mask, const unsigned char* __restrict__ ias,
                                                  const double* __restrict__ csum, const int* __restrict__ order,
                                                  const int* __restrict__ lvl, int nlevels, const double* __restrict__ f0,
                                                  const double* __restrict__ f1, const double* __restrict__ f2,
                                                  const double* __restrict__ f3, double* __restrict__ y) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const unsigned full = (P.nvec >= 32) ? 0xffffffffu : ((1u << P.nvec) - 1u);
  const double* fp[4] = {f0, f1, f2, f3};
  for (int L = 0; L < nlevels; L++) {
    const int lo = lvl[L], hi = lvl[L + 1];
    for (int q = lo + tid; q < hi; q += nth) {
      const int j = order[q];
      const Pt p = unlin(P, j);
      const double rj = __ldg(rho + j);
      double acc[NP + 1];
#pragma unroll
      for (int s = 0; s < NP; s++) acc[s] = __ldg(fp[s] + j);
      acc[NP] = 1.0;
      unsigned lower = (~mask[j]) & full;
      while (lower) {
        const int k = __ffs(lower) - 1;
        lower &= lower - 1;
        const int i = nbr(P, p, k);
        if (ias[i]) {
          const double fr = flux_up(P, k, rj, __ldg(rho + i), csum[i]);
#pragma unroll
          for (int s = 0; s <= NP; s++) acc[s] += fr * __ldcg(y + (size_t)s * nn + i);
        }
      }
#pragma unroll
      for (int s = 0; s <= NP; s++) y[(size_t)s * nn + j] = acc[s];
    }
    grid.sync();
  }
}

// interior points: Y(j) = f(j) + pulls from IAS neighbours below; written to yint[s*nn + j] (IAS: 0)
template <int NP>
__global__ void __launch_bounds__(256) k_interior(const __grid_constant__ YtParams P, const double* __restrict__ rho,
                                                  const unsigned* __restrict__ mask, const unsigned char* __restrict__ ias,
                                                  const double* __restrict__ csum, const double* __restrict__ f0,
                                                  const double* __restrict__ f1, const double* __restrict__ f2,
                                                  const double* __restrict__ f3, double* __restrict__ y) {
  const long long nn = (long long)P.n1 * P.n2 * P.n3;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nn) return;
  if (ias[j]) return;  // handled by the sweep; zeroed afterwards by the reduction's label < 0 test
  const unsigned full = (P.nvec >= 32) ? 0xffffffffu : ((1u << P.nvec) - 1u);
  const double* fp[4] = {f0, f1, f2, f3};
  const Pt p = unlin(P, (int)j);
  const double rj = __ldg(rho + j);
  double acc[NP + 1];
#pragma unroll
  for (int s = 0; s < NP; s++) acc[s] = __ldg(fp[s] + j);
  acc[NP] = 1.0;
  unsigned lower = (~mask[j]) & full;
  while (lower) {
    const int k = __ffs(lower) - 1;
    lower &= lower - 1;
    const int i = nbr(P, p, k);
    if (ias[i]) {
      const double fr = flux_up(P, k, rj, __ldg(rho + i), csum[i]);
#pragma unroll
      for (int s = 0; s <= NP; s++) acc[s] += fr * __ldcg(y + (size_t)s * nn + i);
    }
  }
#pragma unroll
  for (int s = 0; s <= NP; s++) y[(size_t)s * nn + j] = acc[s];
}

// final labels: interior -> index of its maximum in the ordered list, IAS -> -1
__global__ void __launch_bounds__(256) k_yt_labels(long long nn, const int* __restrict__ cand, const unsigned char* __restrict__ ias,
                                                   int* __restrict__ label) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) label[i] = ias[i] ? -1 : cand[i];
}

// downhill sweep for the weight field of one basin (yt_weights): w on IAS points by levels, descending
__global__ void __launch_bounds__(256) k_sweep_down(const __grid_constant__ YtParams P, const double* __restrict__ rho,
                                                    const unsigned* __restrict__ mask, const double* __restrict__ csum,
                                                    const int* __restrict__ order, const int* __restrict__ lvl, int nlevels,
                                                    double* __restrict__ w) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int L = nlevels - 1; L >= 0; L--) {
    const int lo = lvl[L], hi = lvl[L + 1];
    for (int q = lo + tid; q < hi; q += nth) {
      const int i = order[q];
      const Pt p = unlin(P, i);
      const double ri = __ldg(rho + i);
      const double cs = fmax(csum[i], VSMALL);
      double acc = 0.0;
      unsigned hm = mask[i];
      while (hm) {
        const int k = __ffs(hm) - 1;
        hm &= hm - 1;
        const int j = nbr(P, p, k);
        acc += fmax(P.area[k] * (__ldg(rho + j) - ri), VSMALL) / cs * __ldcg(w + j);
      }
      w[i] = acc;
    }
    grid.sync();
  }
}

__global__ void k_gather(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void k_init_w(long long nn, const int* __restrict__ label, const int* __restrict__ map, int idb, double* __restrict__ w) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    const int l = label[i];
    w[i] = (l >= 0 && map[l] == idb) ? 1.0 : 0.0;
  }
}

template <class K>
int coop_grid(c2g_context* ctx, K kernel, int threads, int* blocks) {
  int per_sm = 0;
  C2G_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
  if (per_sm < 1) return ctx->fail(C2G_ERR_CUDA, "cooperative kernel does not fit on an SM");
  *blocks = ctx->nsm * std::min(per_sm, 4);
  return C2G_OK;
}

// YT state kept for integration / weights
struct YtState {
  YtParams P;
  unsigned* mask = nullptr;
  double* csum = nullptr;
  unsigned char* ias = nullptr;
  int* order = nullptr;
  int* lvl = nullptr;
  int* cand = nullptr;            // index of the maximum every point reaches by steepest ascent
  std::vector<int> closure_map;   // maximum -> basin map the IAS closure was built with (empty: one basin per maximum)
  int nlevels = 0;
  int nias = 0;
  c2g_context* ctx = nullptr;
  ~YtState() {
    c2g_release(ctx, cand);
    c2g_release(ctx, mask);
    c2g_release(ctx, csum);
    c2g_release(ctx, ias);
    c2g_release(ctx, order);
    c2g_release(ctx, lvl);
  }
};

}  // namespace

// c2g_basins keeps the YtState behind an opaque pointer
static YtState* yt_state(c2g_basins* res) { return reinterpret_cast<YtState*>(res->yt); }
void c2g_yt_free_state(c2g_basins* res) {
  if (res->yt) {
    delete yt_state(res);
    res->yt = nullptr;
  }
}

// IAS closure, labels and sweep levels for a given maximum -> basin map (device array, null = one basin per maximum).
// Run by c2g_yt_build with the identity and again by c2g_basins_set_map when the caller's attractor identification
// merges maxima (identify_atom / are_lclose) or discards one: the reference classifies every point with the merged
// ids already in ibasin (yt@proc.f90:129-186), so a point between two maxima of ONE attractor is interior there.
static int yt_closure(c2g_context* ctx, c2g_basins* res, const int* d_map) {
  YtState* S = yt_state(res);
  const YtParams& P = S->P;
  cudaStream_t st = ctx->stream;
  const long long nn = res->nn;
  const int nb = c2g_blocks_for(nn, 256);
  DevBuf b_ctl, b_queue, b_indeg;
  C2G_CUDA(ctx, b_ctl.alloc(ctx, 64));
  int* ctl = b_ctl.as<int>();
  int hctl[4];
  // seeds + closure
  C2G_CUDA(ctx, b_queue.alloc(ctx, sizeof(int) * nn));
  C2G_CUDA(ctx, cudaMemsetAsync(S->ias, 0, ((size_t)nn + 3) / 4 * 4, st));
  C2G_CUDA(ctx, cudaMemsetAsync(ctl, 0, 64, st));
  ctx->prof_begin("yt_seed");
  k_seed<<<nb, 256, 0, st>>>(P, S->mask, S->cand, d_map, S->ias, b_queue.as<int>(), ctl);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  {
    int blocks = 0, rc;
    if ((rc = coop_grid(ctx, k_bfs, 256, &blocks)) != C2G_OK) return rc;
    const unsigned* a_mask = S->mask;
    unsigned char* a_ias = S->ias;
    int* a_queue = b_queue.as<int>();
    const int bfs_maxlvl = 1 << 20;  // more levels than any grid dimension allows; the counter index saturates
    DevBuf b_bcnt;
    C2G_CUDA(ctx, b_bcnt.alloc(ctx, sizeof(int) * ((size_t)bfs_maxlvl + 2)));
    C2G_CUDA(ctx, cudaMemsetAsync(b_bcnt.p, 0, sizeof(int) * ((size_t)bfs_maxlvl + 2), st));
    int* a_bcnt = b_bcnt.as<int>();
    int a_bmax = bfs_maxlvl;
    void* args[] = {(void*)&P, (void*)&a_mask, (void*)&a_ias, (void*)&a_queue, (void*)&ctl, (void*)&a_bcnt, (void*)&a_bmax};
    ctx->prof_begin("yt_bfs");
    C2G_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_bfs, dim3(blocks), dim3(256), args, 0, st));
    ctx->prof_end();
  }
  C2G_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, 16, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  const int nias = hctl[0];
  const int bfs_levels = hctl[1];
  S->nias = nias;
  res->n_ias = nias;
  // labels
  ctx->prof_begin("yt_labels");
  k_yt_labels<<<ctx->nsm * 8, 256, 0, st>>>(nn, S->cand, S->ias, res->d_label);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  // Kahn levels of the IAS graph
  const int maxlvl = 1 << 22;
  c2g_release(ctx, S->order); S->order = nullptr;
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&S->order, sizeof(int) * std::max(nias, 1)));
  if (!S->lvl) C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&S->lvl, sizeof(int) * (maxlvl + 1)));
  S->nlevels = 0;
  if (nias > 0) {
    C2G_CUDA(ctx, b_indeg.alloc(ctx, ((size_t)nn + 3) / 4 * 4));
    C2G_CUDA(ctx, cudaMemsetAsync(b_indeg.p, 0, ((size_t)nn + 3) / 4 * 4, st));
    C2G_CUDA(ctx, cudaMemsetAsync(ctl, 0, 64, st));
    ctx->prof_begin("yt_indeg");
    k_indeg<<<c2g_blocks_for(nias, 256), 256, 0, st>>>(P, S->mask, S->ias, b_queue.as<int>(), nias, b_indeg.as<unsigned char>(),
                                                       S->order, ctl);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    int blocks = 0, rc;
    if ((rc = coop_grid(ctx, k_kahn, 256, &blocks)) != C2G_OK) return rc;
    DevBuf b_lcnt;
    C2G_CUDA(ctx, b_lcnt.alloc(ctx, sizeof(int) * ((size_t)maxlvl + 2)));
    C2G_CUDA(ctx, cudaMemsetAsync(b_lcnt.p, 0, sizeof(int) * ((size_t)maxlvl + 2), st));
    int* a_cnt = b_lcnt.as<int>();
    const unsigned* a_mask = S->mask;
    const unsigned char* a_ias = S->ias;
    unsigned char* a_indeg = b_indeg.as<unsigned char>();
    int* a_order = S->order;
    int* a_lvl = S->lvl;
    int a_maxlvl = maxlvl;
    void* args[] = {(void*)&P, (void*)&a_mask, (void*)&a_ias, (void*)&a_indeg, (void*)&a_order, (void*)&ctl, (void*)&a_lvl, (void*)&a_cnt, (void*)&a_maxlvl};
    ctx->prof_begin("yt_kahn");
    C2G_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_kahn, dim3(blocks), dim3(256), args, 0, st));
    ctx->prof_end();
    C2G_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, 16, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hctl[0] != nias) return ctx->fail(C2G_ERR_STATE, "YT: flux graph is not acyclic (%d of %d ordered)", hctl[0], nias);
    if (hctl[1] >= maxlvl) return ctx->fail(C2G_ERR_OVERFLOW, "YT: more than %d sweep levels", maxlvl);
    S->nlevels = hctl[1];
    if (getenv("C2G_YT_NO_LEVEL_SORT") == nullptr && S->nlevels > 0) {  // spatial order inside every level
      DevBuf b_k0, b_k1, b_tmp;
      C2G_CUDA(ctx, b_k0.alloc(ctx, sizeof(unsigned long long) * (size_t)nias));
      C2G_CUDA(ctx, b_k1.alloc(ctx, sizeof(unsigned long long) * (size_t)nias));
      int lbits = 1;
      while ((1ll << lbits) < (long long)S->nlevels + 1) lbits++;
      size_t tmpbytes = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tmpbytes, b_k0.as<unsigned long long>(), b_k1.as<unsigned long long>(), nias, 0, 32 + lbits, st);
      C2G_CUDA(ctx, b_tmp.alloc(ctx, tmpbytes));
      ctx->prof_begin("yt_level_sort");
      k_level_keys<<<c2g_blocks_for(nias, 256), 256, 0, st>>>(nias, S->nlevels, S->lvl, S->order, b_k0.as<unsigned long long>());
      cub::DeviceRadixSort::SortKeys(b_tmp.p, tmpbytes, b_k0.as<unsigned long long>(), b_k1.as<unsigned long long>(), nias, 0, 32 + lbits, st);
      k_level_unkey<<<c2g_blocks_for(nias, 256), 256, 0, st>>>(nias, b_k1.as<unsigned long long>(), S->order);
      ctx->prof_end(3);
      C2G_KERNEL_CHECK(ctx);
    }
  }
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  res->stats[0] = nias;
  res->stats[1] = bfs_levels;
  res->stats[2] = S->nlevels;
  return C2G_OK;
}

// called by c2g_basins_set_map (capi.cu) for YT results: rebuild the closure if the map changes the partition
int c2g_yt_apply_map(c2g_basins* res) {
  c2g_context* ctx = res->ctx;
  YtState* S = yt_state(res);
  if (!S) return ctx->fail(C2G_ERR_STATE, "c2g_basins_set_map: YT state missing");
  // plain map: every maximum is a basin of its own and none is discarded -> the closure of c2g_yt_build
  bool plain = true;
  {
    std::vector<char> seen(res->nattr + 1, 0);
    for (int m = 0; m < res->nmax && plain; m++) {
      const int b = res->map[m];
      if (b <= 0 || seen[b]) plain = false; else seen[b] = 1;
    }
  }
  std::vector<int> key;
  if (!plain) key = res->map;
  if (key == S->closure_map) return C2G_OK;
  int rc = yt_closure(ctx, res, plain ? nullptr : res->d_map);
  if (rc != C2G_OK) return rc;
  S->closure_map.swap(key);
  ctx->prof_collect();
  return C2G_OK;
}

extern "C" int c2g_yt_build(c2g_context* ctx, int handle, int nvec, const int* vec, const double* area, int* nmax_out,
                            c2g_basins** res_out) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_FIRST_DEVICE(ctx);  // YT does not shard (replicas only)
  if (!vec || !area || !nmax_out || !res_out) return ctx->fail(C2G_ERR_ARG, "c2g_yt_build: null argument");
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_yt_build: invalid grid handle %d", handle);
  if (nvec < 1 || nvec > MAXVEC) return ctx->fail(C2G_ERR_ARG, "c2g_yt_build: nvec=%d outside 1..%d", nvec, MAXVEC);
  c2g_grids_ready_all(ctx);
  const c2g_grid& g = ctx->grids[handle];
  if (g.nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "c2g_yt_build: grid too large for int32 indices");
  cudaStream_t st = ctx->stream;
  const long long nn = g.nn;

  c2g_basins* res = new c2g_basins();
  res->ctx = ctx; res->kind = 1; res->gridh = handle;
  res->n[0] = g.n[0]; res->n[1] = g.n[1]; res->n[2] = g.n[2]; res->nn = nn;
  res->nvec = nvec;
  res->zlo = 0; res->zhi = g.n[2];
  struct Guard { c2g_basins* r; bool ok = false; ~Guard() { if (!ok) c2g_basins_free(r); } } guard{res};
  YtState* S = new YtState();
  S->ctx = ctx;
  res->yt = S;
  YtParams& P = S->P;
  P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2]; P.nvec = nvec;
  for (int k = 0; k < nvec; k++) {
    P.vec[3 * k] = vec[3 * k]; P.vec[3 * k + 1] = vec[3 * k + 1]; P.vec[3 * k + 2] = vec[3 * k + 2];
    P.area[k] = area[k];
  }
  for (int k = 0; k < nvec; k++) {
    P.opp[k] = -1;
    for (int q = 0; q < nvec; q++)
      if (vec[3 * q] == -vec[3 * k] && vec[3 * q + 1] == -vec[3 * k + 1] && vec[3 * q + 2] == -vec[3 * k + 2]) P.opp[k] = q;
    if (P.opp[k] < 0) return ctx->fail(C2G_ERR_ARG, "c2g_yt_build: stencil is not centro-symmetric (vec %d has no opposite)", k + 1);
  }

  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&res->d_label, sizeof(int) * nn));
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&S->mask, sizeof(unsigned) * nn));
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&S->csum, sizeof(double) * nn));
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&S->ias, ((size_t)nn + 3) / 4 * 4));
  DevBuf b_up, b_ctl, b_maxl;
  C2G_CUDA(ctx, b_up.alloc(ctx, sizeof(int) * nn));
  C2G_CUDA(ctx, b_ctl.alloc(ctx, 64));
  int maxcap = (int)std::min<long long>(nn, std::max<long long>(1 << 16, nn / 64));
  int* ctl = b_ctl.as<int>();
  int hctl[4];
  const int nb = c2g_blocks_for(nn, 256);
  for (int attempt = 0;; attempt++) {
    C2G_CUDA(ctx, b_maxl.alloc(ctx, sizeof(int) * (size_t)maxcap));
    C2G_CUDA(ctx, cudaMemsetAsync(ctl, 0, 64, st));
    ctx->prof_begin("yt_scan");
    k_scan<<<nb, 256, 0, st>>>(P, g.d, S->mask, S->csum, b_up.as<int>(), b_maxl.as<int>(), ctl, maxcap);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, 16, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (hctl[0] <= maxcap) break;
    if (attempt > 0) return ctx->fail(C2G_ERR_OVERFLOW, "maxima list overflow");
    b_maxl.reset();
    maxcap = hctl[0];
  }
  const int nmax = hctl[0];
  if (nmax == 0) return ctx->fail(C2G_ERR_STATE, "c2g_yt_build: no local maximum (NaN input?)");
  // maxima in decreasing (rho, index): the order in which the reference's sweep meets them (:108-168)
  std::vector<int> ml(nmax);
  C2G_CUDA(ctx, cudaMemcpy(ml.data(), b_maxl.p, sizeof(int) * nmax, cudaMemcpyDeviceToHost));
  std::vector<double> mr(nmax);
  {
    DevBuf b_r;
    C2G_CUDA(ctx, b_r.alloc(ctx, sizeof(double) * nmax));
    k_gather<<<c2g_blocks_for(nmax, 256), 256, 0, st>>>(nmax, b_maxl.as<int>(), g.d, b_r.as<double>());
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(mr.data(), b_r.p, sizeof(double) * nmax, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  std::vector<int> idx(nmax);
  for (int i = 0; i < nmax; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return (mr[a] > mr[b]) || (mr[a] == mr[b] && ml[a] > ml[b]); });
  res->nmax = nmax;
  res->max_lin.resize(nmax);
  for (int k = 0; k < nmax; k++) res->max_lin[k] = ml[idx[k]];
  unsigned hsize = 1024;
  while (hsize < 4u * (unsigned)nmax) hsize <<= 1;
  std::vector<int> hk(hsize, -1), hv(hsize, -1);
  for (int k = 0; k < nmax; k++) {
    unsigned s = ((unsigned)res->max_lin[k] * 2654435761u) & (hsize - 1);
    while (hk[s] >= 0) s = (s + 1) & (hsize - 1);
    hk[s] = res->max_lin[k]; hv[s] = k;
  }
  DevBuf b_hk, b_hv;
  C2G_CUDA(ctx, b_hk.alloc(ctx, sizeof(int) * hsize));
  C2G_CUDA(ctx, b_hv.alloc(ctx, sizeof(int) * hsize));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hk.p, hk.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));
  C2G_CUDA(ctx, cudaMemcpyAsync(b_hv.p, hv.data(), sizeof(int) * hsize, cudaMemcpyHostToDevice, st));

  // pointer jumping
  for (int round = 0; round < 64; round++) {
    C2G_CUDA(ctx, cudaMemsetAsync(ctl + 2, 0, sizeof(int), st));
    ctx->prof_begin("yt_jump");
    k_jump<<<nb, 256, 0, st>>>(nn, b_up.as<int>(), ctl + 2);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, 16, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    if (!hctl[2]) break;
  }
  // terminal maxima -> indices into the ordered list; kept for the closure (which depends on the maximum -> basin map)
  ctx->prof_begin("yt_cand");
  k_cand_index<<<ctx->nsm * 8, 256, 0, st>>>(nn, b_up.as<int>(), b_hk.as<int>(), b_hv.as<int>(), hsize - 1);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  S->cand = b_up.as<int>();
  b_up.p = nullptr;  // ownership moves to the state
  {
    int rc = yt_closure(ctx, res, nullptr);
    if (rc != C2G_OK) return rc;
  }
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  res->stats[4] = nmax;
  ctx->prof_collect();
  guard.ok = true;
  *nmax_out = nmax;
  *res_out = res;
  return C2G_OK;
}

template <int NP>
static int yt_sweep_launch(c2g_context* ctx, YtState* S, const double* rho, const double* const* fp, double* y, long long nn) {
  cudaStream_t st = ctx->stream;
  const double *f0 = fp[0], *f1 = fp[1], *f2 = fp[2], *f3 = fp[3];
  if (S->nias > 0) {
    int blocks = 0, rc;
    if ((rc = coop_grid(ctx, k_sweep_up<NP>, 256, &blocks)) != C2G_OK) return rc;
    const unsigned* a_mask = S->mask;
    const unsigned char* a_ias = S->ias;
    const double* a_csum = S->csum;
    const int* a_order = S->order;
    const int* a_lvl = S->lvl;
    int a_nl = S->nlevels;
    void* args[] = {(void*)&S->P, (void*)&rho, (void*)&a_mask, (void*)&a_ias, (void*)&a_csum, (void*)&a_order, (void*)&a_lvl,
                    (void*)&a_nl, (void*)&f0, (void*)&f1, (void*)&f2, (void*)&f3, (void*)&y};
    ctx->prof_begin("yt_sweep_up");
    C2G_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_sweep_up<NP>, dim3(blocks), dim3(256), args, 0, st));
    ctx->prof_end();
  }
  ctx->prof_begin("yt_interior");
  k_interior<NP><<<c2g_blocks_for(nn, 256), 256, 0, st>>>(S->P, rho, S->mask, S->ias, S->csum, f0, f1, f2, f3, y);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  return C2G_OK;
}

int c2g_yt_integrate_impl(c2g_context* ctx, c2g_basins* res, int nprop, const int* fieldhandles, double omega,
                          double* psum, double* vol) {
  YtState* S = yt_state(res);
  if (!S) return ctx->fail(C2G_ERR_STATE, "c2g_integrate: YT state missing");
  const long long nn = res->nn;
  const int nmax = res->nmax, nattr = res->nattr;
  const double ntot = (double)nn;
  const double* rho = ctx->grids[res->gridh].d;
  if (!ctx->grids[res->gridh].used) return ctx->fail(C2G_ERR_STATE, "c2g_integrate: reference grid was freed");
  cudaStream_t st = ctx->stream;
  std::vector<double> hs((size_t)(nprop + 1) * nmax, 0.0);  // [k*nmax+m], last block = volume
  bool volume_done = false;
  for (int k0 = 0; k0 < std::max(nprop, 1); k0 += 3) {
    // up to 3 properties + the volume slot per sweep (the volume rides along in the first sweep)
    const int np = std::min(3, nprop - k0);
    const int npe = std::max(np, 0);
    const double* fp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int p = 0; p < npe; p++) fp[p] = ctx->grids[fieldhandles[k0 + p]].d;
    DevBuf b_y, b_sums;
    C2G_CUDA(ctx, b_y.alloc(ctx, sizeof(double) * (size_t)(npe + 1) * nn));
    int rc;
    switch (npe) {
      case 0: rc = yt_sweep_launch<0>(ctx, S, rho, fp, b_y.as<double>(), nn); break;
      case 1: rc = yt_sweep_launch<1>(ctx, S, rho, fp, b_y.as<double>(), nn); break;
      case 2: rc = yt_sweep_launch<2>(ctx, S, rho, fp, b_y.as<double>(), nn); break;
      default: rc = yt_sweep_launch<3>(ctx, S, rho, fp, b_y.as<double>(), nn); break;
    }
    if (rc) return rc;
    // per-basin sums over interior points (label < 0 = IAS is skipped by the reduction)
    C2G_CUDA(ctx, b_sums.alloc(ctx, sizeof(double) * (size_t)(npe + 1) * nmax));
    C2G_CUDA(ctx, cudaMemsetAsync(b_sums.p, 0, sizeof(double) * (size_t)(npe + 1) * nmax, st));
    const double* yp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int p = 0; p <= npe; p++) yp[p] = b_y.as<double>() + (size_t)p * nn;
    rc = c2g_launch_basin_reduce(ctx, nn, res->d_label, -1, npe + 1, yp, nmax, b_sums.as<double>(), nullptr);
    if (rc) return rc;
    std::vector<double> part((size_t)(npe + 1) * nmax);
    C2G_CUDA(ctx, cudaMemcpyAsync(part.data(), b_sums.p, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    for (int p = 0; p < npe; p++)
      for (int m = 0; m < nmax; m++) hs[(size_t)(k0 + p) * nmax + m] = part[(size_t)p * nmax + m];
    if (!volume_done) {
      for (int m = 0; m < nmax; m++) hs[(size_t)nprop * nmax + m] = part[(size_t)npe * nmax + m];
      volume_done = true;
    }
  }
  ctx->prof_collect();
  for (int k = 0; k < nprop; k++) {
    for (int i = 0; i < nattr; i++) psum[i + (size_t)nattr * k] = 0.0;
    for (int m = 0; m < nmax; m++) {
      const int b = res->map[m];
      if (b > 0) psum[(b - 1) + (size_t)nattr * k] += hs[(size_t)k * nmax + m];
    }
    for (int i = 0; i < nattr; i++) psum[i + (size_t)nattr * k] = psum[i + (size_t)nattr * k] * omega / ntot;
  }
  if (vol) {
    for (int i = 0; i < nattr; i++) vol[i] = 0.0;
    for (int m = 0; m < nmax; m++) {
      const int b = res->map[m];
      if (b > 0) vol[b - 1] += hs[(size_t)nprop * nmax + m];
    }
    for (int i = 0; i < nattr; i++) vol[i] = vol[i] * omega / ntot;
  }
  return C2G_OK;
}

// weights of basin idb (yt_weights, yt@proc.f90:476-499) into a device array of nn doubles, on ctx->stream
int c2g_yt_weights_device(c2g_basins* res, int idb, double* d_w) {
  c2g_context* ctx = res->ctx;
  YtState* S = yt_state(res);
  const long long nn = res->nn;
  const double* rho = ctx->grids[res->gridh].d;
  cudaStream_t st = ctx->stream;
  ctx->prof_begin("yt_init_w");
  k_init_w<<<ctx->nsm * 8, 256, 0, st>>>(nn, res->d_label, res->d_map, idb, d_w);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  if (S->nias > 0) {
    int blocks = 0, rc;
    if ((rc = coop_grid(ctx, k_sweep_down, 256, &blocks)) != C2G_OK) return rc;
    const unsigned* a_mask = S->mask;
    const double* a_csum = S->csum;
    const int* a_order = S->order;
    const int* a_lvl = S->lvl;
    int a_nl = S->nlevels;
    double* a_w = d_w;
    void* args[] = {(void*)&S->P, (void*)&rho, (void*)&a_mask, (void*)&a_csum, (void*)&a_order, (void*)&a_lvl, (void*)&a_nl, (void*)&a_w};
    ctx->prof_begin("yt_sweep_down");
    C2G_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_sweep_down, dim3(blocks), dim3(256), args, 0, st));
    ctx->prof_end();
  }
  return C2G_OK;
}

extern "C" int c2g_yt_weights(c2g_basins* res, int idb, double* w) {
  if (!res) return C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (res->kind != 1) return ctx->fail(C2G_ERR_STATE, "c2g_yt_weights: not a YT result");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_yt_weights: call c2g_basins_set_map first");
  if (!w || idb < 1 || idb > res->nattr) return ctx->fail(C2G_ERR_ARG, "c2g_yt_weights: unknown basin %d", idb);
  const long long nn = res->nn;
  DevBuf b_w;
  C2G_CUDA(ctx, b_w.alloc(ctx, sizeof(double) * nn));
  int rc = c2g_yt_weights_device(res, idb, b_w.as<double>());
  if (rc != C2G_OK) return rc;
  C2G_CUDA(ctx, cudaMemcpyAsync(w, b_w.p, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_collect();
  return C2G_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ytdata export (yt.f90:36-45; written at yt@proc.f90:191-199, read back by yt_weights :399-530, int_reorder_gridout
// integration@proc.f90:1125-1158 and the BASINS / DI consumers): the record the reference keeps on a scratch file,
// indexed by the POSITION of a point in the density-sorted list (1 = lowest).  The device never sorts for its own
// work; here the permutation is materialised once: a stable radix sort of (rho, index) -- the reference's qcksort
// order whenever no two densities tie -- then one pass per chunk of positions gathers, for every
// point kk, the IAS points ii below it that send flux to it, in the order in which the reference's sweep appends
// them (decreasing ii), with fnear = chi / max(csum, vsmall) of the lower point (:177-186).
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_exp_keys(long long nn, const double* __restrict__ rho, unsigned long long* __restrict__ key,
                                                  int* __restrict__ val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  // order-preserving map of an IEEE double onto an unsigned integer (-0.0 sorts below +0.0; the reference compares
  // them equal, which only matters for tied data)
  const unsigned long long b = (unsigned long long)__double_as_longlong(rho[i]);
  key[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  val[i] = (int)i;
}
__global__ void __launch_bounds__(256) k_exp_iio(long long nn, const int* __restrict__ io, int* __restrict__ iio) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nn) iio[io[p]] = (int)p + 1;
}
// positions [p0, p0 + np): ibasin, nlo and the inear / fnear columns
__global__ void __launch_bounds__(128) k_exp_points(const __grid_constant__ YtParams P, long long p0, int np, const double* __restrict__ rho,
                                                    const unsigned* __restrict__ mask, const double* __restrict__ csum,
                                                    const unsigned char* __restrict__ ias, const int* __restrict__ label,
                                                    const int* __restrict__ map, const int* __restrict__ io, const int* __restrict__ iio,
                                                    int* __restrict__ nlo, int* __restrict__ ibasin, int* __restrict__ inear,
                                                    double* __restrict__ fnear) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= np) return;
  const int j = io[p0 + t];
  const int l = label[j];
  ibasin[t] = l >= 0 ? __ldg(map + l) : 0;
  const Pt pj = unlin(P, j);
  const double rj = rho[j];
  const unsigned mj = mask[j];
  int pos[MAXVEC];
  double fl[MAXVEC];
  int cnt = 0;
  for (int k = 0; k < P.nvec; k++) {
    if ((mj >> k) & 1u) continue;           // that neighbour is higher than j
    const int i = nbr(P, pj, k);
    if (i == j || !ias[i]) continue;        // only IAS points keep their flux records (:177)
    // j = i + vec(opp[k]) is a higher neighbour of i: chi as in k_scan / :121
    const double chi = fmax(P.area[P.opp[k]] * (rj - rho[i]), VSMALL);
    const int pi = iio[i];
    // insertion into the list ordered by decreasing position (the reference appends while ii runs downwards)
    int q = cnt++;
    const double f = chi / fmax(csum[i], VSMALL);
    while (q > 0 && pos[q - 1] < pi) { pos[q] = pos[q - 1]; fl[q] = fl[q - 1]; q--; }
    pos[q] = pi; fl[q] = f;
  }
  nlo[t] = cnt;
  if (inear)
    for (int q = 0; q < P.nvec; q++) inear[(size_t)t * P.nvec + q] = q < cnt ? pos[q] : 0;
  if (fnear)
    for (int q = 0; q < P.nvec; q++) fnear[(size_t)t * P.nvec + q] = q < cnt ? fl[q] : 0.0;
}
}  // namespace

extern "C" int c2g_yt_export(c2g_basins* res, int* nlo, int* ibasin, int* iio, int* inear, double* fnear) {
  if (!res) return C2G_ERR_ARG;
  c2g_context* ctx = res->ctx;
  if (res->kind != 1 || !res->yt) return ctx->fail(C2G_ERR_STATE, "c2g_yt_export: not a YT result");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_yt_export: call c2g_basins_set_map first");
  if (!nlo || !ibasin || !iio) return ctx->fail(C2G_ERR_ARG, "c2g_yt_export: null output");
  cudaSetDevice(ctx->device);
  YtState* S = yt_state(res);
  const YtParams& P = S->P;
  const long long nn = res->nn;
  cudaStream_t st = ctx->stream;
  c2g_grid_ready(ctx, res->gridh);
  const double* rho = ctx->grids[res->gridh].d;
  DevBuf b_key, b_key2, b_val, b_io, b_iio, b_tmp;
  C2G_CUDA(ctx, b_key.alloc(ctx, sizeof(unsigned long long) * nn));
  C2G_CUDA(ctx, b_key2.alloc(ctx, sizeof(unsigned long long) * nn));
  C2G_CUDA(ctx, b_val.alloc(ctx, sizeof(int) * nn));
  C2G_CUDA(ctx, b_io.alloc(ctx, sizeof(int) * nn));
  C2G_CUDA(ctx, b_iio.alloc(ctx, sizeof(int) * nn));
  const int nb = c2g_blocks_for(nn, 256);
  ctx->prof_begin("yt_export_sort");
  k_exp_keys<<<nb, 256, 0, st>>>(nn, rho, b_key.as<unsigned long long>(), b_val.as<int>());
  size_t tmpbytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmpbytes, b_key.as<unsigned long long>(), b_key2.as<unsigned long long>(), b_val.as<int>(),
                                  b_io.as<int>(), (int)nn, 0, 64, st);
  C2G_CUDA(ctx, b_tmp.alloc(ctx, tmpbytes));
  cub::DeviceRadixSort::SortPairs(b_tmp.p, tmpbytes, b_key.as<unsigned long long>(), b_key2.as<unsigned long long>(), b_val.as<int>(),
                                  b_io.as<int>(), (int)nn, 0, 64, st);
  k_exp_iio<<<nb, 256, 0, st>>>(nn, b_io.as<int>(), b_iio.as<int>());
  ctx->prof_end(3);
  C2G_KERNEL_CHECK(ctx);
  C2G_CUDA(ctx, cudaMemcpyAsync(iio, b_iio.p, sizeof(int) * nn, cudaMemcpyDeviceToHost, st));
  b_key.reset(); b_key2.reset(); b_val.reset(); b_tmp.reset();
  // chunks of positions: the inear / fnear columns of 512^3 points with 14 stencil vectors are 22 GB
  const int chunk = (int)std::min<long long>(nn, 1ll << 24);
  DevBuf b_nlo, b_ib, b_in, b_fn;
  C2G_CUDA(ctx, b_nlo.alloc(ctx, sizeof(int) * chunk));
  C2G_CUDA(ctx, b_ib.alloc(ctx, sizeof(int) * chunk));
  if (inear) C2G_CUDA(ctx, b_in.alloc(ctx, sizeof(int) * (size_t)chunk * P.nvec));
  if (fnear) C2G_CUDA(ctx, b_fn.alloc(ctx, sizeof(double) * (size_t)chunk * P.nvec));
  for (long long p0 = 0; p0 < nn; p0 += chunk) {
    const int np = (int)std::min<long long>(chunk, nn - p0);
    ctx->prof_begin("yt_export_points");
    k_exp_points<<<c2g_blocks_for(np, 128), 128, 0, st>>>(P, p0, np, rho, S->mask, S->csum, S->ias, res->d_label, res->d_map,
                                                          b_io.as<int>(), b_iio.as<int>(), b_nlo.as<int>(), b_ib.as<int>(),
                                                          inear ? b_in.as<int>() : nullptr, fnear ? b_fn.as<double>() : nullptr);
    ctx->prof_end();
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(nlo + p0, b_nlo.p, sizeof(int) * np, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaMemcpyAsync(ibasin + p0, b_ib.p, sizeof(int) * np, cudaMemcpyDeviceToHost, st));
    if (inear) C2G_CUDA(ctx, cudaMemcpyAsync(inear + (size_t)p0 * P.nvec, b_in.p, sizeof(int) * (size_t)np * P.nvec, cudaMemcpyDeviceToHost, st));
    if (fnear) C2G_CUDA(ctx, cudaMemcpyAsync(fnear + (size_t)p0 * P.nvec, b_fn.p, sizeof(double) * (size_t)np * P.nvec, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  ctx->prof_collect();
  return C2G_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ISOSURFACE regions: yt_isosurface, src/yt@proc.f90:233-390.
//
// The reference sweeps the sorted grid once more: a point at or above the contour value takes the region of its
// higher neighbours when they agree, else the smallest of their regions, and records imap(other) = smallest for
// every other one -- overwriting what an earlier contact stored (:319-331).  Read as a whole:
//   region(i) = min over the higher neighbours of region(.) = the smallest discovery number among the maxima that
//               i reaches on ascending paths.  Interior points of the YT partition reach one maximum (their YT
//               label); the IAS points take ONE downhill min-sweep over the Kahn levels that c2g_yt_build kept.
//   imap(b)   = the smallest region at the LAST contact (lowest (rho, index)) at which b was not the smallest:
//               three passes of atomicMin over the IAS points (density key, index, then the write).
// The chains of imap are followed on the host exactly like :337-351, surviving regions keep their discovery numbers
// and bas%nattr is the number of survivors (the reference's behaviour, not a renumbering).
// ---------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ unsigned long long rho_key(double r) {  // order-preserving map of a double onto u64
  const unsigned long long b = (unsigned long long)__double_as_longlong(r);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) k_iso_sweep(const __grid_constant__ YtParams P, const unsigned* __restrict__ mask,
                                                   const int* __restrict__ order, const int* __restrict__ lvl, int nlevels,
                                                   int* __restrict__ lab) {
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int L = nlevels - 1; L >= 0; L--) {
    const int lo = lvl[L], hi = lvl[L + 1];
    for (int q = lo + tid; q < hi; q += nth) {
      const int i = order[q];
      const Pt p = unlin(P, i);
      int mn = 0x7fffffff;
      unsigned hm = mask[i];
      while (hm) {
        const int k = __ffs(hm) - 1;
        hm &= hm - 1;
        mn = min(mn, __ldcg(lab + nbr(P, p, k)));
      }
      lab[i] = mn;
    }
    grid.sync();
  }
}

// contacts: IAS points at or above isov whose higher neighbours carry different regions
template <int PASS>
__global__ void __launch_bounds__(256) k_iso_contacts(const __grid_constant__ YtParams P, const double* __restrict__ rho,
                                                      const unsigned* __restrict__ mask, const int* __restrict__ order, int nias,
                                                      double isov, const int* __restrict__ lab,
                                                      unsigned long long* __restrict__ minkey, int* __restrict__ minidx,
                                                      int* __restrict__ imap) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nias) return;
  const int i = order[q];
  const double ri = __ldg(rho + i);
  if (ri < isov) return;
  const Pt p = unlin(P, i);
  const int imin = lab[i];  // = min over the higher neighbours
  const unsigned long long key = rho_key(ri);
  unsigned hm = mask[i];
  while (hm) {
    const int k = __ffs(hm) - 1;
    hm &= hm - 1;
    const int b = lab[nbr(P, p, k)];
    if (b == imin) continue;
    if (PASS == 0) atomicMin(minkey + b, key);
    if (PASS == 1 && minkey[b] == key) atomicMin(minidx + b, i);
    if (PASS == 2 && minkey[b] == key && minidx[b] == i) imap[b] = imin + 1;
  }
}

__global__ void __launch_bounds__(256) k_iso_final(long long nn, const double* __restrict__ rho, double isov, int* __restrict__ lab) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride)
    if (__ldg(rho + i) < isov) lab[i] = -1;
}

}  // namespace

extern "C" int c2g_yt_isosurface(c2g_basins* yt, double isov, int* nraw_out, int* nattr_out, c2g_basins** res_out) {
  if (!yt) return C2G_ERR_ARG;
  c2g_context* ctx = yt->ctx;
  if (yt->kind != 1 || !yt->yt) return ctx->fail(C2G_ERR_STATE, "c2g_yt_isosurface: not a YT result (call c2g_yt_build on the field first)");
  if (!nraw_out || !nattr_out || !res_out) return ctx->fail(C2G_ERR_ARG, "c2g_yt_isosurface: null argument");
  if (!(isov == isov)) return ctx->fail(C2G_ERR_ARG, "c2g_yt_isosurface: the contour value is NaN");
  if (!ctx->grids[yt->gridh].used) return ctx->fail(C2G_ERR_STATE, "c2g_yt_isosurface: the field was freed");
  YtState* S = yt_state(yt);
  const long long nn = yt->nn;
  const double* rho = ctx->grids[yt->gridh].d;
  cudaStream_t st = ctx->stream;

  // regions before merging = maxima at or above the contour value; they lead the (rho, index)-descending list
  const int nmax = yt->nmax;
  std::vector<double> mr(nmax);
  {
    DevBuf b_l, b_r;
    C2G_CUDA(ctx, b_l.alloc(ctx, sizeof(int) * nmax));
    C2G_CUDA(ctx, b_r.alloc(ctx, sizeof(double) * nmax));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_l.p, yt->max_lin.data(), sizeof(int) * nmax, cudaMemcpyHostToDevice, st));
    k_gather<<<c2g_blocks_for(nmax, 256), 256, 0, st>>>(nmax, b_l.as<int>(), rho, b_r.as<double>());
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(mr.data(), b_r.p, sizeof(double) * nmax, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  int nraw = 0;
  while (nraw < nmax && !(mr[nraw] < isov)) nraw++;

  c2g_basins* res = new c2g_basins();
  res->ctx = ctx; res->kind = 2; res->gridh = yt->gridh;
  res->n[0] = yt->n[0]; res->n[1] = yt->n[1]; res->n[2] = yt->n[2]; res->nn = nn;
  res->zlo = 0; res->zhi = yt->n[2];
  res->nmax = nraw;
  res->max_lin.assign(yt->max_lin.begin(), yt->max_lin.begin() + nraw);
  struct Guard { c2g_basins* r; bool ok = false; ~Guard() { if (!ok) c2g_basins_free(r); } } guard{res};
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&res->d_label, sizeof(int) * nn));
  int* lab = res->d_label;
  C2G_CUDA(ctx, cudaMemcpyAsync(lab, yt->d_label, sizeof(int) * nn, cudaMemcpyDeviceToDevice, st));
  std::vector<int> imap(std::max(nraw, 1), 0);
  if (S->nias > 0 && nraw > 0) {
    int blocks = 0, rc;
    if ((rc = coop_grid(ctx, k_iso_sweep, 256, &blocks)) != C2G_OK) return rc;
    const unsigned* a_mask = S->mask;
    const int* a_order = S->order;
    const int* a_lvl = S->lvl;
    int a_nl = S->nlevels;
    void* args[] = {(void*)&S->P, (void*)&a_mask, (void*)&a_order, (void*)&a_lvl, (void*)&a_nl, (void*)&lab};
    ctx->prof_begin("iso_sweep");
    C2G_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_iso_sweep, dim3(blocks), dim3(256), args, 0, st));
    ctx->prof_end();
    DevBuf b_key, b_idx, b_imap;
    C2G_CUDA(ctx, b_key.alloc(ctx, sizeof(unsigned long long) * nraw));
    C2G_CUDA(ctx, b_idx.alloc(ctx, sizeof(int) * nraw));
    C2G_CUDA(ctx, b_imap.alloc(ctx, sizeof(int) * nraw));
    C2G_CUDA(ctx, cudaMemsetAsync(b_key.p, 0xff, sizeof(unsigned long long) * nraw, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_idx.p, 0x7f, sizeof(int) * nraw, st));
    C2G_CUDA(ctx, cudaMemsetAsync(b_imap.p, 0, sizeof(int) * nraw, st));
    const int nb = c2g_blocks_for(S->nias, 256);
    ctx->prof_begin("iso_contacts");
    k_iso_contacts<0><<<nb, 256, 0, st>>>(S->P, rho, S->mask, S->order, S->nias, isov, lab, b_key.as<unsigned long long>(), b_idx.as<int>(), b_imap.as<int>());
    k_iso_contacts<1><<<nb, 256, 0, st>>>(S->P, rho, S->mask, S->order, S->nias, isov, lab, b_key.as<unsigned long long>(), b_idx.as<int>(), b_imap.as<int>());
    k_iso_contacts<2><<<nb, 256, 0, st>>>(S->P, rho, S->mask, S->order, S->nias, isov, lab, b_key.as<unsigned long long>(), b_idx.as<int>(), b_imap.as<int>());
    ctx->prof_end(3);
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaMemcpyAsync(imap.data(), b_imap.p, sizeof(int) * nraw, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  ctx->prof_begin("iso_final");
  k_iso_final<<<ctx->nsm * 8, 256, 0, st>>>(nn, rho, isov, lab);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  // :337-351 -- follow the chains; survivors keep their numbers, nattr = number of survivors
  std::vector<int> map(std::max(nraw, 1), 0);
  int nattr = 0;
  for (int i = 1; i <= nraw; i++) {
    int r = i;
    if (imap[i - 1] == 0) nattr++;
    else
      while (imap[r - 1] != 0) r = imap[r - 1];
    map[i - 1] = r;
  }
  int rc = c2g_basins_set_map(res, nraw, map.data());
  if (rc != C2G_OK) return rc;
  ctx->prof_collect();
  guard.ok = true;
  *nraw_out = nraw;
  *nattr_out = nattr;
  *res_out = res;
  return C2G_OK;
}
