// integrate.cu -- basin integration of INTEGRABLE grid fields on Bader labels.
//
// Replaces the per-attractor masked sums of intgrid_fields (critic2
// src/integration@proc.f90:1208-1218 volume = count(idg==i)*omega/ntot, :1289-1299
// padd = sum(fint, idg==i)*omega/ntot): the reference makes nattr full-grid passes per property;
// this is ONE streaming pass over the labels and up to 4 fields per launch.
//
// Layout: each warp owns contiguous segments of 128*SEG_ITERS points; per iteration a lane reads 4 consecutive
// labels (one 16-byte load) and 4 consecutive values of every field (two 16-byte loads), so a warp covers 128
// points with fully coalesced 512 B / 1 KB requests.  Every lane accumulates privately while the points carry
// the warp's current label; only when the label changes (a basin boundary, about once per basin width) are
// the lanes' partial sums reduced with shuffles and added to the per-maximum accumulators.
// HBM traffic: 4 B (label) + 8 B per field per point.
#include "common.cuh"
#include "group.h"

namespace {

constexpr int SEG_ITERS = 16;      // 128*16 = 2048 points per warp segment
constexpr int RED_BLOCKS_PER_SM = 8;
constexpr int SMEM_TABLE_MAX = 6144;  // doubles per block (48 KB): nmax*(NP+1) must fit to use the shared table
constexpr int LABEL_MASK = 0x7fffffff;  // bit 31 = filled-not-walked flag of bader.cu

// Accumulator target: a per-block shared-memory table when it fits (hot global addresses would
// serialise in the L2 atomic unit: measured ~10 M same-address fp64 atomics/s), else global atomics.
struct Sink {
  double* tab;                    // shared: [(NP+1)][nmax] (slot NP = counts as double) or nullptr
  double* gsums;                  // global sums [p*nmax+m]
  unsigned long long* gcounts;    // global counts (may be null)
  int nmax;
};

template <int NP>
__device__ __forceinline__ void sink_add(const Sink& k, int label, const double* s, unsigned long long c) {
  if (k.tab) {
#pragma unroll
    for (int p = 0; p < NP; p++) atomicAdd(k.tab + (size_t)p * k.nmax + label, s[p]);
    atomicAdd(k.tab + (size_t)NP * k.nmax + label, (double)c);  // exact: counts < 2^53
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) atomicAdd(k.gsums + (size_t)p * k.nmax + label, s[p]);
    if (k.gcounts) atomicAdd(k.gcounts + label, c);
  }
}

template <int NP>
struct Group {  // 4 consecutive points of one lane
  int l[4];
  double v[NP > 0 ? NP : 1][4];
  unsigned valid;  // bit j: point j exists
};

template <int NP>
__device__ __forceinline__ void load_group(Group<NP>& g, long long i, long long nn, const int* __restrict__ label, int mask,
                                           bool vec, const double* const* fp) {
  g.valid = 0;
  if (vec && i + 3 < nn) {  // vec: label and every field are 16-byte aligned
    const int4 l4 = *reinterpret_cast<const int4*>(label + i);
    g.l[0] = l4.x & mask; g.l[1] = l4.y & mask; g.l[2] = l4.z & mask; g.l[3] = l4.w & mask;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(fp[p] + i));
      const double2 b = __ldg(reinterpret_cast<const double2*>(fp[p] + i + 2));
      g.v[p][0] = a.x; g.v[p][1] = a.y; g.v[p][2] = b.x; g.v[p][3] = b.y;
    }
    // negative labels (YT: interatomic-surface points) belong to no basin
    g.valid = (g.l[0] >= 0 ? 1u : 0u) | (g.l[1] >= 0 ? 2u : 0u) | (g.l[2] >= 0 ? 4u : 0u) | (g.l[3] >= 0 ? 8u : 0u);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      g.l[j] = -1;
      if (i + j < nn) {
        g.l[j] = label[i + j] & mask;
#pragma unroll
        for (int p = 0; p < NP; p++) g.v[p][j] = __ldg(fp[p] + i + j);
        if (g.l[j] >= 0) g.valid |= 1u << j;
      }
    }
  }
}

template <int NP>
#ifndef RED_MINB
#define RED_MINB 4
#endif
// three and four integrand grids need 90-118 registers: under the 64-register bound of 4 resident blocks they spilled
// 240-576 bytes (profiles/r01q_ptxas_v.txt); they run 2 blocks per SM instead (no spills)
__global__ void __launch_bounds__(256, (NP <= 2 ? RED_MINB : 2)) k_basin_reduce(long long nn, const int* __restrict__ label,
                                                      const double* __restrict__ f0, const double* __restrict__ f1,
                                                      const double* __restrict__ f2, const double* __restrict__ f3,
                                                      int nmax, int use_table, int mask, int vec, double* __restrict__ sums,
                                                      unsigned long long* __restrict__ counts,
                                                      double* __restrict__ partials) {
  extern __shared__ double s_tab[];
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long seg = 128ll * SEG_ITERS;
  const double* fp[4] = {f0, f1, f2, f3};
  Sink sink{use_table ? s_tab : nullptr, sums, counts, nmax};
  if (use_table) {
    for (int e = threadIdx.x; e < (NP + 1) * nmax; e += blockDim.x) s_tab[e] = 0.0;
    __syncthreads();
  }
  double acc[NP > 0 ? NP : 1];
#pragma unroll
  for (int p = 0; p < (NP > 0 ? NP : 1); p++) acc[p] = 0.0;
  unsigned cnt = 0;
  int cur = -1;  // warp-uniform label of the current run
  auto flush = [&]() {
    if (cur >= 0 && __any_sync(0xffffffffu, cnt != 0)) {
      unsigned c = cnt;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
      double s[NP > 0 ? NP : 1];
#pragma unroll
      for (int p = 0; p < NP; p++) {
        s[p] = acc[p];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s[p] += __shfl_xor_sync(0xffffffffu, s[p], d);
      }
      if (lane == 0) sink_add<NP>(sink, cur, s, (unsigned long long)c);
    }
#pragma unroll
    for (int p = 0; p < (NP > 0 ? NP : 1); p++) acc[p] = 0.0;
    cnt = 0;
  };
  for (long long base = warp * seg; base < nn; base += nwarps * seg) {
    Group<NP> g, gn;
    load_group<NP>(gn, base + 4 * lane, nn, label, mask, vec != 0, fp);
    for (int k = 0; k < SEG_ITERS; k++) {
      if (base + 128ll * k >= nn) break;
      g = gn;
      if (k + 1 < SEG_ITERS) load_group<NP>(gn, base + 128ll * (k + 1) + 4 * lane, nn, label, mask, vec != 0, fp);  // prefetch
      unsigned rem = g.valid;
      for (;;) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (((rem >> j) & 1u) && g.l[j] == cur) {
#pragma unroll
            for (int p = 0; p < NP; p++) acc[p] += g.v[p][j];
            cnt++;
            rem &= ~(1u << j);
          }
        const unsigned bal = __ballot_sync(0xffffffffu, rem != 0);
        if (!bal) break;
        flush();  // the run of `cur` ends inside this group
        const int candl = (rem & 1u) ? g.l[0] : (rem & 2u) ? g.l[1] : (rem & 4u) ? g.l[2] : g.l[3];
        cur = __shfl_sync(0xffffffffu, candl, __ffs(bal) - 1);  // label of the first remaining point in memory order
      }
    }
  }
  flush();
  if (use_table) {
    // block partials, reduced in a fixed order by k_reduce_partials (no hot global atomics)
    __syncthreads();
    double* out = partials + (size_t)blockIdx.x * (NP + 1) * nmax;
    for (int e = threadIdx.x; e < (NP + 1) * nmax; e += blockDim.x) out[e] = s_tab[e];
  }
}

// sums[p*nmax+m] += sum over blocks of partials[b][p][m]; counts from slot np
__global__ void __launch_bounds__(256) k_reduce_partials(int nblocks, int np, int nmax, const double* __restrict__ partials,
                                                         double* __restrict__ sums, unsigned long long* __restrict__ counts) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (np + 1) * nmax) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += partials[(size_t)b * (np + 1) * nmax + e];
  if (e < np * nmax) sums[e] += s;
  else if (counts) counts[e - np * nmax] += (unsigned long long)(s + 0.5);
}

}  // namespace

// sums[(p*nmax)+m], counts[m] per MAXIMUM index (device pointers, accumulated into); fields: up to 4 device arrays
int c2g_launch_basin_reduce(c2g_context* ctx, long long nn, const int* label, int label_mask, int np, const double* const* f,
                            int nmax, double* sums, unsigned long long* counts) {
  const int blocks = ctx->nsm * RED_BLOCKS_PER_SM;
  const double* f0 = np > 0 ? f[0] : nullptr;
  const double* f1 = np > 1 ? f[1] : nullptr;
  const double* f2 = np > 2 ? f[2] : nullptr;
  const double* f3 = np > 3 ? f[3] : nullptr;
  const int use_table = ((long long)(np + 1) * nmax <= SMEM_TABLE_MAX) ? 1 : 0;
  int vec = ((uintptr_t)label % 16 == 0) ? 1 : 0;
  for (int p = 0; p < np && p < 4; p++)
    if ((uintptr_t)f[p] % 16 != 0) vec = 0;
  const size_t smem = use_table ? sizeof(double) * (size_t)(np + 1) * nmax : 0;
  double* partials = nullptr;
  if (use_table) C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&partials, sizeof(double) * (size_t)blocks * (np + 1) * nmax));
  ctx->prof_begin("basin_reduce");
  switch (np) {
    case 0: k_basin_reduce<0><<<blocks, 256, smem, ctx->stream>>>(nn, label, f0, f1, f2, f3, nmax, use_table, label_mask, vec, sums, counts, partials); break;
    case 1: k_basin_reduce<1><<<blocks, 256, smem, ctx->stream>>>(nn, label, f0, f1, f2, f3, nmax, use_table, label_mask, vec, sums, counts, partials); break;
    case 2: k_basin_reduce<2><<<blocks, 256, smem, ctx->stream>>>(nn, label, f0, f1, f2, f3, nmax, use_table, label_mask, vec, sums, counts, partials); break;
    case 3: k_basin_reduce<3><<<blocks, 256, smem, ctx->stream>>>(nn, label, f0, f1, f2, f3, nmax, use_table, label_mask, vec, sums, counts, partials); break;
    default: k_basin_reduce<4><<<blocks, 256, smem, ctx->stream>>>(nn, label, f0, f1, f2, f3, nmax, use_table, label_mask, vec, sums, counts, partials); break;
  }
  int nl = 1;
  if (use_table) {
    k_reduce_partials<<<c2g_blocks_for((long long)(np + 1) * nmax, 256), 256, 0, ctx->stream>>>(blocks, np, nmax, partials, sums, counts);
    nl = 2;
  }
  ctx->prof_end(nl);
  cudaError_t e = cudaGetLastError();
  if (partials) c2g_release(ctx, partials);
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "basin_reduce launch: %s", cudaGetErrorString(e));
  return C2G_OK;
}

int c2g_yt_integrate_impl(c2g_context* ctx, c2g_basins* res, int nprop, const int* fieldhandles, double omega,
                          double* psum, double* vol);

extern "C" int c2g_integrate(c2g_context* ctx, c2g_basins* res, int nprop, const int* fieldhandles, double omega,
                             double* psum, double* vol) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group && res && !res->parts.empty()) return grp_integrate(ctx, res, nprop, fieldhandles, omega, psum, vol);
  if (ctx->group && res) { ctx = res->ctx; cudaSetDevice(ctx->device); }  // a YT result lives on the first device
  if (!res || nprop < 0 || (nprop > 0 && (!fieldhandles || !psum))) return ctx->fail(C2G_ERR_ARG, "c2g_integrate: bad argument");
  if (!res->has_map) return ctx->fail(C2G_ERR_STATE, "c2g_integrate: call c2g_basins_set_map first");
  for (int k = 0; k < nprop; k++) {
    const int h = fieldhandles[k];
    if (h < 0 || h >= (int)ctx->grids.size() || !ctx->grids[h].used) return ctx->fail(C2G_ERR_ARG, "c2g_integrate: invalid field handle %d", h);
    c2g_grid_ready(ctx, h);
    const c2g_grid& g = ctx->grids[h];
    if (g.n[0] != res->n[0] || g.n[1] != res->n[1] || g.n[2] != res->n[2])
      return ctx->fail(C2G_ERR_ARG, "c2g_integrate: field %d has a different grid size", k + 1);
  }
  if (res->kind == 1) return c2g_yt_integrate_impl(ctx, res, nprop, fieldhandles, omega, psum, vol);

  const int nmax = res->nmax, nattr = res->nattr;
  const double ntot = (double)res->nn;
  double* d_sums = nullptr;
  unsigned long long* d_counts = nullptr;
  std::vector<double> hs((size_t)std::max(nprop, 1) * nmax, 0.0);
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_sums, sizeof(double) * hs.size()));
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_counts, sizeof(unsigned long long) * nmax));
  C2G_CUDA(ctx, cudaMemsetAsync(d_sums, 0, sizeof(double) * hs.size(), ctx->stream));
  C2G_CUDA(ctx, cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * nmax, ctx->stream));
  int rc = C2G_OK;
  bool first = true;
  // z-slab owned by this rank (single GPU: the whole grid); partial sums are all-reduced below
  const size_t plane = (size_t)res->n[0] * res->n[1];
  const long long nnl = (long long)plane * (res->zhi - res->zlo);
  for (int k0 = 0; k0 < std::max(nprop, 1) && rc == C2G_OK && nnl > 0; k0 += 4) {
    const int np = std::min(4, nprop - k0);
    const double* fp[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int p = 0; p < np; p++) fp[p] = ctx->grids[fieldhandles[k0 + p]].d + plane * res->zlo;
    rc = c2g_launch_basin_reduce(ctx, nnl, res->d_label, res->kind == 0 ? LABEL_MASK : -1, np > 0 ? np : 0, fp, nmax, d_sums + (size_t)k0 * nmax,
                                 first ? d_counts : nullptr);
    first = false;
  }
  if (rc == C2G_OK && ctx->nranks > 1 && res->kind == 0) {  // Bader labels are z-slabs; ISOSURFACE regions are replicas
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    ctx->prof_begin("basin_allreduce_nccl");
    ncclResult_t r1 = ncclAllReduce(d_sums, d_sums, hs.size(), ncclDouble, ncclSum, comm, ctx->stream);
    ncclResult_t r2 = ncclAllReduce(d_counts, d_counts, nmax, ncclUint64, ncclSum, comm, ctx->stream);
    ctx->prof_end();
    if (r1 != ncclSuccess || r2 != ncclSuccess) rc = ctx->fail(C2G_ERR_NCCL, "c2g_integrate: ncclAllReduce failed");
  }
  std::vector<unsigned long long> hc(nmax);
  cudaError_t e = cudaSuccess;
  if (rc == C2G_OK) {
    e = cudaMemcpyAsync(hs.data(), d_sums, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hc.data(), d_counts, sizeof(unsigned long long) * nmax, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  c2g_release(ctx, d_sums);
  c2g_release(ctx, d_counts);
  if (rc != C2G_OK) return rc;
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "c2g_integrate: %s", cudaGetErrorString(e));
  ctx->prof_collect();
  // maxima -> basins (several maxima may belong to one attractor; map 0 = discarded)
  for (int k = 0; k < nprop; k++) {
    for (int i = 0; i < nattr; i++) psum[i + (size_t)nattr * k] = 0.0;
    for (int m = 0; m < nmax; m++) {
      const int b = res->map[m];
      if (b > 0) psum[(b - 1) + (size_t)nattr * k] += hs[(size_t)k * nmax + m];
    }
    for (int i = 0; i < nattr; i++) psum[i + (size_t)nattr * k] = psum[i + (size_t)nattr * k] * omega / ntot;
  }
  if (vol) {
    std::vector<unsigned long long> cb(nattr, 0);
    for (int m = 0; m < nmax; m++)
      if (res->map[m] > 0) cb[res->map[m] - 1] += hc[m];
    for (int i = 0; i < nattr; i++) vol[i] = (double)cb[i] * omega / ntot;
  }
  return C2G_OK;
}
