// nci.cu -- NCIPLOT reduced-density-gradient loop (grid / tricubic mode) on sm_100a.
//
// Replaces the triple loop of nciplot (critic2 src/nci@proc.f90:543-605) for grid fields evaluated
// with INTERPOLATION GRID (tricubic): per lattice point
//     x  = x0 + i*xmat(:,1) + j*xmat(:,2) + k*xmat(:,3)                      (nci@proc.f90:548)
//     wx = c2x*x, wrapped outside [-1e-4, 1+1e-4], modulo 1                   (fieldmod@proc.f90:921-929, grid3mod@proc.f90:1716)
//     rho, grad, Hessian by the Lekien-Marsden tricubic interpolant           (grid3mod@proc.f90:2649-2821)
//     s  = |grad rho| / (2 (3 pi^2)^(1/3) max(rho,1e-80)^(4/3))              (nci@proc.f90:91,571)
//     crho = sign(rho, lambda_2) * 100 ; cgrad = s                            (nci@proc.f90:598-599)
// The tricubic interpolant with central-difference derivative data is a tensor product of 1-D cubic
// Hermite (Catmull-Rom) kernels, so the 64x64 matrix-vector product of the reference collapses to three
// 4-point contractions; at grid nodes (fractional offset exactly 0) it reduces further to the closed
// forms below, evaluated in the reference's operation order.  Only sign(lambda_2) is consumed, so the
// middle eigenvalue's sign is taken from Descartes' rule on the characteristic polynomial (all roots
// real) instead of a LAPACK dsyev call.
//
// The coordinate chain decides which cell floor() selects and the tricubic Hessian is discontinuous
// across cells, so that chain is evaluated with explicit __dmul_rn/__dadd_rn (never fused) in the
// Fortran left-to-right order.
//
// Mapping: a block owns a 32(i) x 2(j) x 4(k) brick of the output lattice; lanes run along i (the
// fastest index of rho) for the stencil loads, results are staged in shared memory and written with k
// fastest, which is the reference's cgrad(k,j,i) layout.  HBM traffic: 8 B read + 16 B written per point.
#include "common.cuh"
#include "group.h"

#include <cmath>

namespace {

struct NciParams {
  int n1, n2, n3;          // rho grid
  int ns1, ns2, ns3;       // output lattice (ns1 = the i range computed by this rank)
  int i0;                  // first lattice index i of this rank (multi-GPU: the lattice is sharded along i)
  double x0[3], xmat[9];   // Cartesian
  double c2x[9], x2c[9], c2xl[9];
  int nnuc;
  double cst;  // 2 (3 pi^2)^(1/3), nci@proc.f90:91, evaluated once on the host
  int ortho;   // c2xl is diagonal: the skipped products of the Cartesian transforms are exact zeros
};

constexpr int BI = 32, BJ = 2, BK = 4;  // lanes run along i (the contiguous index of rho): coalesced stencil loads; 4 consecutive k = one 32 B output sector

__device__ __forceinline__ int imod(int a, int n) {
  if (a >= -n && a < 2 * n) {  // the usual case: at most one period off (no integer division)
    a += a < 0 ? n : 0;
    return a >= n ? a - n : a;
  }
  const int r = a % n;
  return r < 0 ? r + n : r;
}

__device__ __forceinline__ int positive_roots_middle_sign(double hxx, double hyy, double hzz, double hxy, double hxz,
                                                          double hyz) {
  // characteristic polynomial l^3 - c2 l^2 + c1 l - c0 ; all roots real => Descartes' rule is exact
  const double c2 = hxx + hyy + hzz;
  const double m0 = fma(hyy, hzz, -hyz * hyz), m1 = fma(hxy, hzz, -hyz * hxz), m2 = fma(hxy, hyz, -hyy * hxz);
  const double c1 = (fma(hxx, hyy, -hxy * hxy) + fma(hxx, hzz, -hxz * hxz)) + m0;
  const double c0 = fma(hxz, m2, fma(hxx, m0, -hxy * m1));
  const double seq[4] = {1.0, -c2, c1, -c0};
  int npos = 0, nzero = 0;
  double prev = 1.0;
  for (int q = 1; q < 4; q++) {
    if (seq[q] == 0.0) continue;
    if ((seq[q] > 0.0) != (prev > 0.0)) npos++;
    prev = seq[q];
  }
  if (c0 == 0.0) { nzero = 1; if (c1 == 0.0) { nzero = 2; if (c2 == 0.0) nzero = 3; } }
  return (npos + nzero >= 2) ? 1 : -1;  // lambda_2 >= 0 -> +
}

// general point: tensor-product cubic Hermite with central-difference slopes (value, gradient, Hessian in grid
// units).  Not inlined: the node-aligned lattice never comes here and should not pay for its registers.
__device__ __noinline__ void tricubic_general(const NciParams& P, const double* __restrict__ rho, const int xs_[4], const int ys_[4],
                                              const int zs_[4], const double t[3], double& f, double gl[3], double h[6]) {
  auto G = [&](int a, int b, int c) -> double {  // offsets -1..2
    return __ldg(rho + xs_[a + 1] + (size_t)P.n1 * (ys_[b + 1] + (size_t)P.n2 * zs_[c + 1]));
  };
  {
      // general point: tensor-product cubic Hermite with central-difference slopes
      double w[3][4], w1[3][4], w2[3][4];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double u = t[d], u2 = u * u, u3 = u2 * u;
        w[d][0] = 0.5 * (-u3 + 2.0 * u2 - u);
        w[d][1] = 0.5 * (3.0 * u3 - 5.0 * u2 + 2.0);
        w[d][2] = 0.5 * (-3.0 * u3 + 4.0 * u2 + u);
        w[d][3] = 0.5 * (u3 - u2);
        w1[d][0] = 0.5 * (-3.0 * u2 + 4.0 * u - 1.0);
        w1[d][1] = 0.5 * (9.0 * u2 - 10.0 * u);
        w1[d][2] = 0.5 * (-9.0 * u2 + 8.0 * u + 1.0);
        w1[d][3] = 0.5 * (3.0 * u2 - 2.0 * u);
        w2[d][0] = 0.5 * (-6.0 * u + 4.0);
        w2[d][1] = 0.5 * (18.0 * u - 10.0);
        w2[d][2] = 0.5 * (-18.0 * u + 8.0);
        w2[d][3] = 0.5 * (6.0 * u - 2.0);
      }
      f = 0.0;
      gl[0] = gl[1] = gl[2] = 0.0;
#pragma unroll
      for (int q = 0; q < 6; q++) h[q] = 0.0;
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        double p0 = 0, px = 0, pxx = 0, py = 0, pxy = 0, pyy = 0;  // contracted over x and y
#pragma unroll
        for (int b = 0; b < 4; b++) {
          double r0 = 0, rx = 0, rxx = 0;
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const double v = G(a - 1, b - 1, c - 1);
            r0 += w[0][a] * v;
            rx += w1[0][a] * v;
            rxx += w2[0][a] * v;
          }
          p0 += w[1][b] * r0;
          px += w[1][b] * rx;
          pxx += w[1][b] * rxx;
          py += w1[1][b] * r0;
          pxy += w1[1][b] * rx;
          pyy += w2[1][b] * r0;
        }
        f += w[2][c] * p0;
        gl[0] += w[2][c] * px;
        gl[1] += w[2][c] * py;
        gl[2] += w1[2][c] * p0;
        h[0] += w[2][c] * pxx;
        h[1] += w[2][c] * pyy;
        h[2] += w2[2][c] * p0;
        h[3] += w[2][c] * pxy;
        h[4] += w1[2][c] * px;
        h[5] += w1[2][c] * py;
      }
  }
}

#ifndef NCI_MINB
#define NCI_MINB 4
#endif
__global__ void __launch_bounds__(256, NCI_MINB) k_nci_rdg(const __grid_constant__ NciParams P, const double* __restrict__ rho,
                                                 const double* __restrict__ nuc, double* __restrict__ crho,
                                                 double* __restrict__ cgrad) {
  __shared__ double s_rho[BI * BJ * BK];
  __shared__ double s_grad[BI * BJ * BK];
  const int tid = threadIdx.x;
  const int li = tid % BI, lj = (tid / BI) % BJ, lk = tid / (BI * BJ);
  const int il = blockIdx.x * BI + li, j = blockIdx.y * BJ + lj, k = blockIdx.z * BK + lk;
  const int i = P.i0 + il;
  double out_rho = 0.0, out_grad = 0.0;
  if (il < P.ns1 && j < P.ns2 && k < P.ns3) {
    // ---- coordinate chain (never fused) ----
    double wx[3];
    {
      double x[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double v = __dadd_rn(P.x0[d], __dmul_rn((double)i, P.xmat[d]));
        v = __dadd_rn(v, __dmul_rn((double)j, P.xmat[d + 3]));
        v = __dadd_rn(v, __dmul_rn((double)k, P.xmat[d + 6]));
        x[d] = v;
      }
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double s = __dmul_rn(P.c2x[d], x[0]);
        s = __dadd_rn(s, __dmul_rn(P.c2x[d + 3], x[1]));
        s = __dadd_rn(s, __dmul_rn(P.c2x[d + 6], x[2]));
        if (s < -1e-4 || s > 1.0 + 1e-4) s = s - floor(s);
        wx[d] = s;
      }
    }
    int idx[3];
    double t[3];
    const int nn[3] = {P.n1, P.n2, P.n3};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double xi = wx[d] - floor(wx[d]);  // modulo(xi,1d0)
      if (xi >= 1.0) xi = 0.0;
      const double xs = __dmul_rn(xi, (double)nn[d]);
      const int fl = (int)floor(xs);
      idx[d] = imod(fl, nn[d]);
      t[d] = xs - (double)idx[d];  // (:2772)
    }
    const int xs_[4] = {imod(idx[0] - 1, P.n1), idx[0], imod(idx[0] + 1, P.n1), imod(idx[0] + 2, P.n1)};
    const int ys_[4] = {imod(idx[1] - 1, P.n2), idx[1], imod(idx[1] + 1, P.n2), imod(idx[1] + 2, P.n2)};
    const int zs_[4] = {imod(idx[2] - 1, P.n3), idx[2], imod(idx[2] + 1, P.n3), imod(idx[2] + 2, P.n3)};
    auto G = [&](int a, int b, int c) -> double {  // offsets -1..2
      return __ldg(rho + xs_[a + 1] + (size_t)P.n1 * (ys_[b + 1] + (size_t)P.n2 * zs_[c + 1]));
    };
    double f, gl[3], h[6];  // h: xx, yy, zz, xy, xz, yz in grid (lattice) coordinates, unscaled
    if (t[0] == 0.0 && t[1] == 0.0 && t[2] == 0.0) {
      // node: closed forms of a = matmul(c,b) rows (value, first and second derivatives)
      const double g0 = G(0, 0, 0);
      const double gxp = G(1, 0, 0), gxm = G(-1, 0, 0), gx2 = G(2, 0, 0);
      const double gyp = G(0, 1, 0), gym = G(0, -1, 0), gy2 = G(0, 2, 0);
      const double gzp = G(0, 0, 1), gzm = G(0, 0, -1), gz2 = G(0, 0, 2);
      f = g0;
      gl[0] = 0.5 * (gxp - gxm);
      gl[1] = 0.5 * (gyp - gym);
      gl[2] = 0.5 * (gzp - gzm);
      // a_200 = -3 f0 + 3 f1 - 2 d0 - d1 ; second derivative = 2 a_200
      h[0] = 2.0 * (((-3.0 * g0 + 3.0 * gxp) + (-2.0) * (0.5 * (gxp - gxm))) + (-1.0) * (0.5 * (gx2 - g0)));
      h[1] = 2.0 * (((-3.0 * g0 + 3.0 * gyp) + (-2.0) * (0.5 * (gyp - gym))) + (-1.0) * (0.5 * (gy2 - g0)));
      h[2] = 2.0 * (((-3.0 * g0 + 3.0 * gzp) + (-2.0) * (0.5 * (gzp - gzm))) + (-1.0) * (0.5 * (gz2 - g0)));
      h[3] = 0.25 * (G(1, 1, 0) - G(-1, 1, 0) - G(1, -1, 0) + G(-1, -1, 0));
      h[4] = 0.25 * (G(1, 0, 1) - G(-1, 0, 1) - G(1, 0, -1) + G(-1, 0, -1));
      h[5] = 0.25 * (G(0, 1, 1) - G(0, -1, 1) - G(0, 1, -1) + G(0, -1, -1));
    } else {
      tricubic_general(P, rho, xs_, ys_, zs_, t, f, gl, h);
    }
    // to fractional coordinates (:2811-2817)
    const double dn[3] = {(double)P.n1, (double)P.n2, (double)P.n3};
    double yp[3] = {gl[0] * dn[0], gl[1] * dn[1], gl[2] * dn[2]};
    double H[3][3];
    H[0][0] = h[0] * dn[0] * dn[0];
    H[1][1] = h[1] * dn[1] * dn[1];
    H[2][2] = h[2] * dn[2] * dn[2];
    H[0][1] = H[1][0] = h[3] * dn[0] * dn[1];
    H[0][2] = H[2][0] = h[4] * dn[0] * dn[2];
    H[1][2] = H[2][1] = h[5] * dn[1] * dn[2];
    // to Cartesian (grid3mod@proc.f90:1747-1750): yp = matmul(transpose(c2xl),yp); ypp = c2xl^T ypp c2xl
    double gc[3], HC[3][3];
    if (P.ortho) {  // orthogonal cell: same values, a fraction of the fp64 work (this kernel is FP64-pipe bound)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        gc[a] = __dmul_rn(P.c2xl[4 * a], yp[a]);
#pragma unroll
        for (int b = 0; b < 3; b++) HC[a][b] = (P.c2xl[4 * a] * H[a][b]) * P.c2xl[4 * b];
      }
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        double s = __dmul_rn(P.c2xl[0 + 3 * a], yp[0]);
        s = __dadd_rn(s, __dmul_rn(P.c2xl[1 + 3 * a], yp[1]));
        s = __dadd_rn(s, __dmul_rn(P.c2xl[2 + 3 * a], yp[2]));
        gc[a] = s;
      }
      double T[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
          // (only sign(lambda_2) is taken from the Cartesian Hessian: fused multiply-adds are as good as any order)
          double s = P.c2xl[3 * a] * H[0][b];
          s = fma(P.c2xl[1 + 3 * a], H[1][b], s);
          s = fma(P.c2xl[2 + 3 * a], H[2][b], s);
          T[a][b] = s;
        }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
          double s = T[a][0] * P.c2xl[3 * b];
          s = fma(T[a][1], P.c2xl[1 + 3 * b], s);
          s = fma(T[a][2], P.c2xl[2 + 3 * b], s);
          HC[a][b] = s;
        }
    }
    // nucleus rule (fieldmod@proc.f90:1148-1155): zero gradient within 1e-5 bohr of an atom
    if (P.nnuc > 0) {
      double wc[3];
#pragma unroll
      for (int d = 0; d < 3; d++) wc[d] = P.x2c[d] * wx[0] + P.x2c[d + 3] * wx[1] + P.x2c[d + 6] * wx[2];
      bool isnuc = false;
      for (int a = 0; a < P.nnuc && !isnuc; a++) {
        const double dc0 = wc[0] - nuc[3 * a], dc1 = wc[1] - nuc[3 * a + 1], dc2 = wc[2] - nuc[3 * a + 2];
        double dx[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          dx[d] = P.c2x[d] * dc0 + P.c2x[d + 3] * dc1 + P.c2x[d + 6] * dc2;
          dx[d] -= rint(dx[d]);
        }
        // minimum image among the 27 nearest translations
        for (int ta = -1; ta <= 1 && !isnuc; ta++)
          for (int tb = -1; tb <= 1 && !isnuc; tb++)
            for (int tc = -1; tc <= 1; tc++) {
              const double e0 = dx[0] + ta, e1 = dx[1] + tb, e2 = dx[2] + tc;
              const double q0 = P.x2c[0] * e0 + P.x2c[3] * e1 + P.x2c[6] * e2;
              const double q1 = P.x2c[1] * e0 + P.x2c[4] * e1 + P.x2c[7] * e2;
              const double q2 = P.x2c[2] * e0 + P.x2c[5] * e1 + P.x2c[8] * e2;
              if (sqrt(q0 * q0 + q1 * q1 + q2 * q2) <= 1e-5) { isnuc = true; break; }
            }
      }
      if (isnuc) gc[0] = gc[1] = gc[2] = 0.0;
    }
    const double gfmod = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(gc[0], gc[0]), __dmul_rn(gc[1], gc[1])), __dmul_rn(gc[2], gc[2])));
    // max(rho,vsmall)**(4/3) as rho * cbrt(rho): same value to 1-2 ulp (the 1e-12 contract), a fraction of pow's cost
    const double fm = fmax(f, 1e-80);
    const double dimgrad = gfmod / (P.cst * (fm * cbrt(fm)));
    const int sg = positive_roots_middle_sign(HC[0][0], HC[1][1], HC[2][2], 0.5 * (HC[0][1] + HC[1][0]),
                                              0.5 * (HC[0][2] + HC[2][0]), 0.5 * (HC[1][2] + HC[2][1]));
    out_grad = dimgrad;
    out_rho = (sg > 0 ? fabs(f) : -fabs(f)) * 100.0;
  }
  // stage [li][lj][lk] and write with k fastest
  s_rho[(li * BJ + lj) * BK + lk] = out_rho;
  s_grad[(li * BJ + lj) * BK + lk] = out_grad;
  __syncthreads();
  {
    const int ok = tid % BK, oj = (tid / BK) % BJ, oi = tid / (BK * BJ);
    const int gi = blockIdx.x * BI + oi, gj = blockIdx.y * BJ + oj, gk = blockIdx.z * BK + ok;
    if (gi < P.ns1 && gj < P.ns2 && gk < P.ns3) {
      const size_t o = (size_t)gk + (size_t)P.ns3 * ((size_t)gj + (size_t)P.ns2 * gi);
      crho[o] = s_rho[(oi * BJ + oj) * BK + ok];
      cgrad[o] = s_grad[(oi * BJ + oj) * BK + ok];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FOURIER mode (nci@proc.f90:527-565): every quantity is a field%grd call with nder = 0, i.e. the raw node
// value when the point is within neargrideps = 1e-12 grid units of a node (fieldmod@proc.f90:948-961), else
// grid3%interp: tricubic for rho (the reference field keeps its mode), TRILINEAR for the FFT-derived
// |grad rho|, Hxx, Hyy, Hzz grids (:534-537, grid3mod@proc.f90:2323-2370).
// ------------------------------------------------------------------------------------------------
struct GridPos {
  bool isgrid;
  int node[3];   // nearest node (valid when isgrid)
  int idx[3];    // floor cell
  double t[3];   // tricubic offset xs - idx
  double r[3];   // trilinear weight n*x0 - idx
};
__device__ __forceinline__ void grid_pos(const NciParams& P, const double wx[3], GridPos& g) {
  const int nn[3] = {P.n1, P.n2, P.n3};
  g.isgrid = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double xi = wx[d] - floor(wx[d]);  // modulo(wx,1d0)
    if (xi >= 1.0) xi = 0.0;
    const double xs = __dmul_rn(xi, (double)nn[d]);
    const double nr = floor(xs + 0.5);  // nint of a non-negative number
    if (!(fabs(xs - nr) < 1e-12)) g.isgrid = false;
    g.node[d] = imod((int)nr, nn[d]);
    const int fl = (int)floor(xs);
    g.idx[d] = imod(fl, nn[d]);
    g.t[d] = xs - (double)g.idx[d];
    g.r[d] = __dadd_rn(__dadd_rn(__dmul_rn((double)nn[d], xi), -(double)(g.idx[d] + 1)), 1.0);  // f%n*x0 - idx + 1, idx 1-based
  }
}
__device__ __forceinline__ double tricubic_value(const NciParams& P, const double* __restrict__ rho, const GridPos& g) {
  const int xs_[4] = {imod(g.idx[0] - 1, P.n1), g.idx[0], imod(g.idx[0] + 1, P.n1), imod(g.idx[0] + 2, P.n1)};
  const int ys_[4] = {imod(g.idx[1] - 1, P.n2), g.idx[1], imod(g.idx[1] + 1, P.n2), imod(g.idx[1] + 2, P.n2)};
  const int zs_[4] = {imod(g.idx[2] - 1, P.n3), g.idx[2], imod(g.idx[2] + 1, P.n3), imod(g.idx[2] + 2, P.n3)};
  if (g.t[0] == 0.0 && g.t[1] == 0.0 && g.t[2] == 0.0) return __ldg(rho + xs_[1] + (size_t)P.n1 * (ys_[1] + (size_t)P.n2 * zs_[1]));
  double w[3][4];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double u = g.t[d], u2 = u * u, u3 = u2 * u;
    w[d][0] = 0.5 * (-u3 + 2.0 * u2 - u);
    w[d][1] = 0.5 * (3.0 * u3 - 5.0 * u2 + 2.0);
    w[d][2] = 0.5 * (-3.0 * u3 + 4.0 * u2 + u);
    w[d][3] = 0.5 * (u3 - u2);
  }
  double f = 0.0;
#pragma unroll 1
  for (int c = 0; c < 4; c++) {
    double p0 = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      double r0 = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) r0 += w[0][a] * __ldg(rho + xs_[a] + (size_t)P.n1 * (ys_[b] + (size_t)P.n2 * zs_[c]));
      p0 += w[1][b] * r0;
    }
    f += w[2][c] * p0;
  }
  return f;
}
// grinterp_trilinear value in the reference's operation order (:2349-2362)
__device__ __forceinline__ double trilinear_value(const NciParams& P, const double* __restrict__ f, const GridPos& g) {
  const int x0 = g.idx[0], x1 = imod(g.idx[0] + 1, P.n1);
  const int y0 = g.idx[1], y1 = imod(g.idx[1] + 1, P.n2);
  const int z0 = g.idx[2], z1 = imod(g.idx[2] + 1, P.n3);
  auto F = [&](int x, int y, int z) { return __ldg(f + x + (size_t)P.n1 * (y + (size_t)P.n2 * z)); };
  const double r1 = g.r[0], r2 = g.r[1], r3 = g.r[2];
  const double s1 = 1.0 - r1, s2 = 1.0 - r2, s3 = 1.0 - r3;
  // ff(i,j,2) = ff(i,j,0)*s3 + ff(i,j,1)*r3 ; ff(i,2,2) = ff(i,0,2)*s2 + ff(i,1,2)*r2 ; ff(2,2,2) = ff(0,2,2)*s1 + ff(1,2,2)*r1
  const double a00 = F(x0, y0, z0) * s3 + F(x0, y0, z1) * r3, a01 = F(x0, y1, z0) * s3 + F(x0, y1, z1) * r3;
  const double a10 = F(x1, y0, z0) * s3 + F(x1, y0, z1) * r3, a11 = F(x1, y1, z0) * s3 + F(x1, y1, z1) * r3;
  const double b0 = a00 * s2 + a01 * r2, b1 = a10 * s2 + a11 * r2;
  return b0 * s1 + b1 * r1;
}
__device__ __forceinline__ double grd0_value(const NciParams& P, const double* __restrict__ f, const GridPos& g, bool trilinear) {
  if (g.isgrid) return __ldg(f + g.node[0] + (size_t)P.n1 * (g.node[1] + (size_t)P.n2 * g.node[2]));
  return trilinear ? trilinear_value(P, f, g) : tricubic_value(P, f, g);
}

__global__ void __launch_bounds__(256) k_nci_rdg_fourier(const __grid_constant__ NciParams P, const double* __restrict__ rho,
                                                         const double* __restrict__ fgrad, const double* __restrict__ fxx,
                                                         const double* __restrict__ fyy, const double* __restrict__ fzz,
                                                         double* __restrict__ crho, double* __restrict__ cgrad) {
  __shared__ double s_rho[BI * BJ * BK];
  __shared__ double s_grad[BI * BJ * BK];
  const int tid = threadIdx.x;
  const int li = tid % BI, lj = (tid / BI) % BJ, lk = tid / (BI * BJ);
  const int il = blockIdx.x * BI + li, j = blockIdx.y * BJ + lj, k = blockIdx.z * BK + lk;
  const int i = P.i0 + il;
  double out_rho = 0.0, out_grad = 0.0;
  if (il < P.ns1 && j < P.ns2 && k < P.ns3) {
    double wx[3];
    {
      double x[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double v = __dadd_rn(P.x0[d], __dmul_rn((double)i, P.xmat[d]));
        v = __dadd_rn(v, __dmul_rn((double)j, P.xmat[d + 3]));
        v = __dadd_rn(v, __dmul_rn((double)k, P.xmat[d + 6]));
        x[d] = v;
      }
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double s = __dmul_rn(P.c2x[d], x[0]);
        s = __dadd_rn(s, __dmul_rn(P.c2x[d + 3], x[1]));
        s = __dadd_rn(s, __dmul_rn(P.c2x[d + 6], x[2]));
        if (s < -1e-4 || s > 1.0 + 1e-4) s = s - floor(s);
        wx[d] = s;
      }
    }
    GridPos g;
    grid_pos(P, wx, g);
    const double f = grd0_value(P, rho, g, false);
    const double gm = grd0_value(P, fgrad, g, true);
    const double hx = grd0_value(P, fxx, g, true), hy = grd0_value(P, fyy, g, true), hz = grd0_value(P, fzz, g, true);
    const double fm = fmax(f, 1e-80);
    out_grad = gm / (P.cst * (fm * cbrt(fm)));
    const int npos = (hx > 0.0) + (hy > 0.0) + (hz > 0.0);
    out_rho = (npos >= 2 ? fabs(f) : -fabs(f)) * 100.0;
  }
  s_rho[(li * BJ + lj) * BK + lk] = out_rho;
  s_grad[(li * BJ + lj) * BK + lk] = out_grad;
  __syncthreads();
  {
    const int ok = tid % BK, oj = (tid / BK) % BJ, oi = tid / (BK * BJ);
    const int gi = blockIdx.x * BI + oi, gj = blockIdx.y * BJ + oj, gk = blockIdx.z * BK + ok;
    if (gi < P.ns1 && gj < P.ns2 && gk < P.ns3) {
      const size_t o = (size_t)gk + (size_t)P.ns3 * ((size_t)gj + (size_t)P.ns2 * gi);
      crho[o] = s_rho[(oi * BJ + oj) * BK + ok];
      cgrad[o] = s_grad[(oi * BJ + oj) * BK + ok];
    }
  }
}

// lattice rows i owned by this rank: an even split of nstep(1) (the slowest index of cgrad(k,j,i), so every rank's
// part is one contiguous piece of the reference arrays)
static void nci_range(const c2g_context* ctx, int ns1, int* ilo, int* ihi) {
  *ilo = (int)((long long)ns1 * ctx->rank / ctx->nranks);
  *ihi = (int)((long long)ns1 * (ctx->rank + 1) / ctx->nranks);
}

int nci_launch(c2g_context* ctx, int handle, const double x0[3], const double xmat[9], const int nstep[3],
               const double c2x[9], const double x2c[9], const double c2xl[9], int nnuc, const double* nuc_cart,
               double* d_rho, double* d_grad) {
  c2g_grids_ready_all(ctx);
  const c2g_grid& g = ctx->grids[handle];
  NciParams P;
  P.n1 = g.n[0]; P.n2 = g.n[1]; P.n3 = g.n[2];
  int ilo, ihi;
  nci_range(ctx, nstep[0], &ilo, &ihi);
  P.ns1 = ihi - ilo; P.i0 = ilo; P.ns2 = nstep[1]; P.ns3 = nstep[2];
  memcpy(P.x0, x0, sizeof(P.x0));
  memcpy(P.xmat, xmat, sizeof(P.xmat));
  memcpy(P.c2x, c2x, sizeof(P.c2x));
  memcpy(P.x2c, x2c, sizeof(P.x2c));
  memcpy(P.c2xl, c2xl, sizeof(P.c2xl));
  P.nnuc = nnuc;
  P.ortho = (c2xl[1] == 0.0 && c2xl[2] == 0.0 && c2xl[3] == 0.0 && c2xl[5] == 0.0 && c2xl[6] == 0.0 && c2xl[7] == 0.0) ? 1 : 0;
  P.cst = 2.0 * std::pow(3.0 * 3.14159265358979323846264338328 * 3.14159265358979323846264338328, 1.0 / 3.0);
  double* d_nuc = nullptr;
  if (nnuc > 0) {
    C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_nuc, sizeof(double) * 3 * nnuc));
    C2G_CUDA(ctx, cudaMemcpyAsync(d_nuc, nuc_cart, sizeof(double) * 3 * nnuc, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (P.ns1 < 1) return C2G_OK;  // more ranks than lattice rows
  dim3 grid((P.ns1 + BI - 1) / BI, (nstep[1] + BJ - 1) / BJ, (nstep[2] + BK - 1) / BK);
  ctx->prof_begin("nci_rdg");
  k_nci_rdg<<<grid, 256, 0, ctx->stream>>>(P, g.d, d_nuc, d_rho, d_grad);
  ctx->prof_end();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (d_nuc) c2g_release(ctx, d_nuc);
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "nci_rdg: %s", cudaGetErrorString(e));
  ctx->prof_collect();
  return C2G_OK;
}

int nci_check(c2g_context* ctx, int handle, const double* x0, const double* xmat, const int* nstep, const double* c2x,
              const double* x2c, const double* c2xl, int nnuc, const double* nuc) {
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg: invalid grid handle %d", handle);
  if (!x0 || !xmat || !nstep || !c2x || !x2c || !c2xl || nnuc < 0 || (nnuc > 0 && !nuc))
    return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg: null argument");
  if (nstep[0] < 1 || nstep[1] < 1 || nstep[2] < 1) return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg: bad nstep");
  if (ctx->grids[handle].nn >= (1ll << 31)) return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg: grid too large");
  return C2G_OK;
}

}  // namespace

extern "C" int c2g_nci_rdg(c2g_context* ctx, int handle, const double x0[3], const double xmat[9], const int nstep[3],
                           const double c2x[9], const double x2c[9], const double c2xl[9], int nnuc,
                           const double* nuc_cart, double* crho, double* cgrad) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_nci_rdg(ctx, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc_cart, crho, cgrad);
  int rc = nci_check(ctx, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc_cart);
  if (rc) return rc;
  if (!crho || !cgrad) return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg: null output");
  int ilo, ihi;
  nci_range(ctx, nstep[0], &ilo, &ihi);
  const size_t nout = (size_t)std::max(ihi - ilo, 0) * nstep[1] * nstep[2];
  double *d_rho = nullptr, *d_grad = nullptr;
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_rho, sizeof(double) * nout));
  C2G_CUDA(ctx, c2g_alloc(ctx, (void**)&d_grad, sizeof(double) * nout));
  rc = nci_launch(ctx, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc_cart, d_rho, d_grad);
  cudaError_t e = cudaSuccess;
  if (rc == C2G_OK) {
    e = cudaMemcpyAsync(crho, d_rho, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cgrad, d_grad, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  c2g_release(ctx, d_rho);
  c2g_release(ctx, d_grad);
  if (rc) return rc;
  if (e != cudaSuccess) return ctx->fail(C2G_ERR_CUDA, "c2g_nci_rdg: %s", cudaGetErrorString(e));
  return C2G_OK;
}

extern "C" int c2g_nci_rdg_resident(c2g_context* ctx, int handle, const double x0[3], const double xmat[9],
                                    const int nstep[3], const double c2x[9], const double x2c[9], const double c2xl[9],
                                    int nnuc, const double* nuc_cart, int* hrho, int* hgrad) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_nci_rdg_resident");
  int rc = nci_check(ctx, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc_cart);
  if (rc) return rc;
  if (!hrho || !hgrad) return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg_resident: null output");
  int ilo, ihi;
  nci_range(ctx, nstep[0], &ilo, &ihi);
  const int nout[3] = {nstep[2], nstep[1], std::max(ihi - ilo, 1)};
  if ((rc = c2g_grid_alloc(ctx, nout, hrho)) != C2G_OK) return rc;
  if ((rc = c2g_grid_alloc(ctx, nout, hgrad)) != C2G_OK) return rc;
  return nci_launch(ctx, handle, x0, xmat, nstep, c2x, x2c, c2xl, nnuc, nuc_cart, ctx->grids[*hrho].d, ctx->grids[*hgrad].d);
}

// FOURIER mode: h = {rho, |grad rho|, Hxx, Hyy, Hzz} resident grids of the same shape (the last four from
// c2g_fft_derivative with C2G_FT_GRAD, _XX, _YY, _ZZ, as nci@proc.f90:528-531 loads them)
extern "C" int c2g_nci_rdg_fourier(c2g_context* ctx, const int h[5], const double x0[3], const double xmat[9],
                                   const int nstep[3], const double c2x[9], const double c2xl[9], double* crho,
                                   double* cgrad) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) return grp_nci_rdg_fourier(ctx, h, x0, xmat, nstep, c2x, c2xl, crho, cgrad);
  if (!h || !crho || !cgrad) return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg_fourier: null argument");
  int rc = nci_check(ctx, h[0], x0, xmat, nstep, c2x, c2x, c2xl, 0, nullptr);
  if (rc) return rc;
  c2g_grids_ready_all(ctx);
  const c2g_grid& g0 = ctx->grids[h[0]];
  for (int q = 1; q < 5; q++) {
    if (h[q] < 0 || h[q] >= (int)ctx->grids.size() || !ctx->grids[h[q]].used)
      return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg_fourier: invalid grid handle %d", h[q]);
    const c2g_grid& gq = ctx->grids[h[q]];
    if (gq.n[0] != g0.n[0] || gq.n[1] != g0.n[1] || gq.n[2] != g0.n[2])
      return ctx->fail(C2G_ERR_ARG, "c2g_nci_rdg_fourier: the derived grids must have the shape of rho");
  }
  NciParams P;
  P.n1 = g0.n[0]; P.n2 = g0.n[1]; P.n3 = g0.n[2];
  int ilo, ihi;
  nci_range(ctx, nstep[0], &ilo, &ihi);
  P.ns1 = ihi - ilo; P.i0 = ilo; P.ns2 = nstep[1]; P.ns3 = nstep[2];
  memcpy(P.x0, x0, sizeof(P.x0));
  memcpy(P.xmat, xmat, sizeof(P.xmat));
  memcpy(P.c2x, c2x, sizeof(P.c2x));
  memcpy(P.x2c, c2x, sizeof(P.x2c));  // unused in this mode
  memcpy(P.c2xl, c2xl, sizeof(P.c2xl));
  P.nnuc = 0;
  P.ortho = 0;
  P.cst = 2.0 * std::pow(3.0 * 3.14159265358979323846264338328 * 3.14159265358979323846264338328, 1.0 / 3.0);
  if (P.ns1 < 1) return C2G_OK;
  const size_t nout = (size_t)P.ns1 * nstep[1] * nstep[2];
  DevBuf d_rho, d_grad;
  C2G_CUDA(ctx, d_rho.alloc(ctx, sizeof(double) * nout));
  C2G_CUDA(ctx, d_grad.alloc(ctx, sizeof(double) * nout));
  dim3 grid((P.ns1 + BI - 1) / BI, (nstep[1] + BJ - 1) / BJ, (nstep[2] + BK - 1) / BK);
  ctx->prof_begin("nci_rdg_fourier");
  k_nci_rdg_fourier<<<grid, 256, 0, ctx->stream>>>(P, ctx->grids[h[0]].d, ctx->grids[h[1]].d, ctx->grids[h[2]].d,
                                                   ctx->grids[h[3]].d, ctx->grids[h[4]].d, d_rho.as<double>(), d_grad.as<double>());
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  C2G_CUDA(ctx, cudaMemcpyAsync(crho, d_rho.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaMemcpyAsync(cgrad, d_grad.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_collect();
  return C2G_OK;
}

// Multi-GPU: the lattice rows i (0-based, half open) this rank computes; c2g_nci_rdg* then fill / return only
// cgrad(:,:,ilo+1:ihi), which is one contiguous piece of the reference array.  Single GPU: the whole range.
extern "C" int c2g_nci_range(c2g_context* ctx, int nstep1, int* ilo, int* ihi) {
  if (!ctx) return C2G_ERR_ARG;
  if (ctx->group) { if (!ilo || !ihi) return C2G_ERR_ARG; *ilo = 0; *ihi = nstep1; return C2G_OK; }  // the caller holds whole arrays
  if (!ilo || !ihi || nstep1 < 1) return ctx->fail(C2G_ERR_ARG, "c2g_nci_range: bad argument");
  nci_range(ctx, nstep1, ilo, ihi);
  return C2G_OK;
}
