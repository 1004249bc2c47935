// oracle.cpp -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement (C++17, IEEE fp64, no FMA contraction, no fast-math) of the
// algorithms on critic2's on-grid QTAIM hot path.  It is the checker for the
// CUDA implementation under critic2_b200/csrc and the timed CPU baseline of
// bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.
//
// PARITY STATUS: "parity unpinned".  critic2 is Fortran; no Fortran compiler
// exists in this image and the input grids of the reference's golden .cro files
// (tests/009_intgrid/ref/*.cro) are not shipped, so this restatement cannot be
// pinned against outputs of the reference itself for BADER, YT and the grid-field
// NCI path.  Those are pinned by line-by-line review against the cited source
// ranges, by the published tricubic matrix (checked against
// src/grid3mod@proc.f90:76-340 when /root/reference exists), and by analytic
// known-answer tests (tests/test_oracle_*.py).  PINNED by the reference's own
// outputs (golden files of its nodata tests, tests/golden/cube_golden.json): the
// promolecular density (promolecular_atom + grid1%interp, reproduces the urea
// density cube of 015_grdplot/005_nciplot_basic to its six printed digits), the
// RDG formula (the gradient cube of the same test), the real solid harmonics
// (closed forms in the comment block of genrlm_real); the Fortran edit
// descriptors of the cube writers are restated in oracle.py and pinned there.
//
// All file:line citations are relative to the critic2 source tree.
//
// Layout convention everywhere: Fortran column-major f(n1,n2,n3), index 1
// fastest; 0-based linear id = i1 + n1*(i2 + n2*i3).  3x3 matrices are passed
// column-major as in Fortran: m[i + 3*j] = M(i+1,j+1).
//
// Build: g++ -O3 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile)

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double VSMALL = 1e-80;  // src/param.F90:27

inline int nint_(double x) { return (int)std::lround(x); }  // Fortran nint: half away from zero

// periodic wrap of a 0-based coordinate (bader@proc.f90:601-617 pbc, :619-640 rho_val)
inline int wrap0(int p, int n) {
  while (p < 0) p += n;
  while (p >= n) p -= n;
  return p;
}

// ---------------------------------------------------------------------------
// Bader near-grid method.  src/bader@proc.f90:57-66 (module state)
// ---------------------------------------------------------------------------
struct Bader {
  int n[3];
  const double* f;
  double car2lat[9];      // column-major
  double lat_i_dist[27];  // index (d1+1)*9 + (d2+1)*3 + (d3+1); 0 at centre
  std::vector<int> volnum, known;
  std::vector<int> path;  // linear ids of the current path
  int pnum = 0;
  double dr[3] = {0, 0, 0};  // `save`d correction vector of step_neargrid (:460)
  int nbasin = 0;
  long nsteps = 0;  // statistics: near-grid steps taken

  inline int lin(const int p[3]) const { return p[0] + n[0] * (p[1] + n[1] * p[2]); }
  inline void unlin(int id, int p[3]) const {
    p[0] = id % n[0];
    p[1] = (id / n[0]) % n[1];
    p[2] = id / (n[0] * n[1]);
  }
  inline double rho(int p1, int p2, int p3) const {
    return f[wrap0(p1, n[0]) + n[0] * (wrap0(p2, n[1]) + n[1] * wrap0(p3, n[2]))];
  }
  inline void pbc(int p[3]) const {
    for (int i = 0; i < 3; i++) p[i] = wrap0(p[i], n[i]);
  }

  // rho_grad_dir, bader@proc.f90:532-567
  void rho_grad_dir(const int p[3], double res[3]) const {
    const double rho000 = rho(p[0], p[1], p[2]);
    const double rho001 = rho(p[0], p[1], p[2] + 1);
    const double rho010 = rho(p[0], p[1] + 1, p[2]);
    const double rho100 = rho(p[0] + 1, p[1], p[2]);
    const double rho00_1 = rho(p[0], p[1], p[2] - 1);
    const double rho_100 = rho(p[0] - 1, p[1], p[2]);
    const double rho0_10 = rho(p[0], p[1] - 1, p[2]);
    double gl[3];
    gl[0] = (rho100 - rho_100) / 2.0;
    gl[1] = (rho010 - rho0_10) / 2.0;
    gl[2] = (rho001 - rho00_1) / 2.0;
    if (rho100 < rho000 && rho_100 < rho000) gl[0] = 0.0;
    if (rho010 < rho000 && rho0_10 < rho000) gl[1] = 0.0;
    if (rho001 < rho000 && rho00_1 < rho000) gl[2] = 0.0;
    // rho_grad_car = matmul(rho_grad_lat, car2lat): gc(j) = sum_i gl(i)*C(i,j)   (:562)
    double gc[3];
    for (int j = 0; j < 3; j++) {
      double s = gl[0] * car2lat[0 + 3 * j];
      s = s + gl[1] * car2lat[1 + 3 * j];
      s = s + gl[2] * car2lat[2 + 3 * j];
      gc[j] = s;
    }
    // res = matmul(car2lat, rho_grad_car): res(i) = sum_j C(i,j)*gc(j)           (:565)
    for (int i = 0; i < 3; i++) {
      double s = car2lat[i + 3 * 0] * gc[0];
      s = s + car2lat[i + 3 * 1] * gc[1];
      s = s + car2lat[i + 3 * 2] * gc[2];
      res[i] = s;
    }
  }

  // is_max, bader@proc.f90:571-597
  bool is_max(const int p[3]) const {
    const double r = rho(p[0], p[1], p[2]);
    bool ismax = true;
    for (int d1 = -1; d1 <= 1; d1++)
      for (int d2 = -1; d2 <= 1; d2++)
        for (int d3 = -1; d3 <= 1; d3++)
          if (rho(p[0] + d1, p[1] + d2, p[2] + d3) > r) ismax = false;
    return ismax;
  }

  // step_ongrid, bader@proc.f90:500-527
  void step_ongrid(int p[3]) const {
    int pm[3] = {p[0], p[1], p[2]};
    const double rho_ctr = rho(p[0], p[1], p[2]);
    double rho_max = rho_ctr;
    for (int d1 = -1; d1 <= 1; d1++)
      for (int d2 = -1; d2 <= 1; d2++)
        for (int d3 = -1; d3 <= 1; d3++) {
          double rho_tmp = rho(p[0] + d1, p[1] + d2, p[2] + d3);
          rho_tmp = rho_ctr + (rho_tmp - rho_ctr) * lat_i_dist[(d1 + 1) * 9 + (d2 + 1) * 3 + (d3 + 1)];
          if (rho_tmp > rho_max) {
            rho_max = rho_tmp;
            pm[0] = p[0] + d1;
            pm[1] = p[1] + d2;
            pm[2] = p[2] + d3;
          }
        }
    pbc(pm);
    p[0] = pm[0];
    p[1] = pm[1];
    p[2] = pm[2];
  }

  // step_neargrid, bader@proc.f90:455-494
  void step_neargrid(int p[3]) {
    if (pnum == 1) dr[0] = dr[1] = dr[2] = 0.0;
    double g[3];
    rho_grad_dir(p, g);
    int pm[3];
    const double gmax = std::max(std::fabs(g[0]), std::max(std::fabs(g[1]), std::fabs(g[2])));
    if (gmax < 1e-30) {
      dr[0] = dr[1] = dr[2] = 0.0;
      if (is_max(p)) return;
      pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
      step_ongrid(pm);
    } else {
      const double coeff = 1.0 / gmax;
      for (int i = 0; i < 3; i++) {
        g[i] = coeff * g[i];
        const int ng = nint_(g[i]);
        pm[i] = p[i] + ng;
        dr[i] = dr[i] + g[i] - (double)ng;   // dr = dr + gradrl - nint(gradrl), left to right
        const int nd = nint_(dr[i]);
        pm[i] = pm[i] + nd;
        dr[i] = dr[i] - (double)nd;
      }
    }
    nsteps++;
    known[lin(p)] = 1;
    pbc(pm);
    if (known[lin(pm)] == 1) {
      pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
      step_ongrid(pm);
      dr[0] = dr[1] = dr[2] = 0.0;
    }
    p[0] = pm[0]; p[1] = pm[1]; p[2] = pm[2];
  }

  // max_neargrid, bader@proc.f90:427-450
  void max_neargrid(int p[3]) {
    pnum = 1;
    path.clear();
    path.push_back(lin(p));
    while (true) {
      step_neargrid(p);
      if (lin(p) == path[pnum - 1]) break;  // did not move: maximum
      pnum++;
      path.push_back(lin(p));
      if (known[lin(p)] == 2) break;  // quit at a known point
    }
  }

  inline int volnum_val(int p1, int p2, int p3) const {
    return volnum[wrap0(p1, n[0]) + n[0] * (wrap0(p2, n[1]) + n[1] * wrap0(p3, n[2]))];
  }
  // known_volnum_ongrid, bader@proc.f90:705-726
  void known_volnum_ongrid(const int p[3]) {
    const int v = volnum_val(p[0], p[1], p[2]);
    if (v <= 0) return;
    if (volnum_val(p[0], p[1], p[2] + 1) != v) return;
    if (volnum_val(p[0], p[1], p[2] - 1) != v) return;
    if (volnum_val(p[0], p[1] + 1, p[2]) != v) return;
    if (volnum_val(p[0], p[1] - 1, p[2]) != v) return;
    if (volnum_val(p[0] + 1, p[1], p[2]) != v) return;
    if (volnum_val(p[0] - 1, p[1], p[2]) != v) return;
    known[lin(p)] = 2;
  }
  // assign_surrounding_pts, bader@proc.f90:665-700
  void assign_surrounding_pts(const int p[3]) {
    static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int k = 0; k < 6; k++) {
      int pt[3] = {p[0] + d[k][0], p[1] + d[k][1], p[2] + d[k][2]};
      pbc(pt);
      if (known[lin(pt)] != 2) known_volnum_ongrid(pt);
    }
  }
  // is_vol_edge, bader@proc.f90:730-752
  bool is_vol_edge(const int p[3]) const {
    const int v = volnum[lin(p)];
    for (int d1 = -1; d1 <= 1; d1++)
      for (int d2 = -1; d2 <= 1; d2++)
        for (int d3 = -1; d3 <= 1; d3++) {
          int pt[3] = {p[0] + d1, p[1] + d2, p[2] + d3};
          pbc(pt);
          if (std::abs(volnum[lin(pt)]) != std::abs(v)) return true;
        }
    return false;
  }

  // refine_edge, bader@proc.f90:300-422.  Returns the number of reassignments; -1 on the
  // "should be no new maxima" error.
  long refine_edge(int ref_itrs, long* nedge_out) {
    int p[3], pt[3] = {0, 0, 0};
    long num_edge = 0;
    if (ref_itrs == 1) {
      for (p[0] = 0; p[0] < n[0]; p[0]++)
        for (p[1] = 0; p[1] < n[1]; p[1]++)
          for (p[2] = 0; p[2] < n[2]; p[2]++) {
            const int id = lin(p);
            if (volnum[id] == nbasin + 1) continue;
            if (is_vol_edge(p) && !is_max(p)) {
              num_edge++;
              volnum[id] = -volnum[id];
              known[id] = 0;
              // reassign_volnum_ongrid2, :756-771
              for (int d1 = -1; d1 <= 1; d1++)
                for (int d2 = -1; d2 <= 1; d2++)
                  for (int d3 = -1; d3 <= 1; d3++) {
                    int q[3] = {p[0] + d1, p[1] + d2, p[2] + d3};
                    pbc(q);
                    known[lin(q)] = 0;
                  }
            }
          }
    } else {
      for (p[0] = 0; p[0] < n[0]; p[0]++)
        for (p[1] = 0; p[1] < n[1]; p[1]++)
          for (p[2] = 0; p[2] < n[2]; p[2]++) {
            const int id = lin(p);
            if (volnum[id] == nbasin + 1) continue;
            if (volnum[id] < 0 && known[id] != -1) {
              for (int d1 = -1; d1 <= 1; d1++)
                for (int d2 = -1; d2 <= 1; d2++)
                  for (int d3 = -1; d3 <= 1; d3++) {
                    pt[0] = p[0] + d1; pt[1] = p[1] + d2; pt[2] = p[2] + d3;
                    pbc(pt);
                    const int iq = lin(pt);
                    if (volnum[iq] == nbasin + 1) continue;
                    if (!is_max(pt)) {
                      if (volnum[iq] > 0) {
                        volnum[iq] = -volnum[iq];
                        known[iq] = -1;
                      } else if (volnum[iq] < 0 && known[iq] == 0) {
                        known[iq] = -2;
                      }
                    }
                  }
              // upstream quirk (:358-361): pt is the last visited neighbour p+(1,1,1)
              if (known[lin(pt)] != -2) volnum[id] = std::abs(volnum[id]);
            }
          }
      // make the surrounding points unknown (:367-388)
      for (p[0] = 0; p[0] < n[0]; p[0]++)
        for (p[1] = 0; p[1] < n[1]; p[1]++)
          for (p[2] = 0; p[2] < n[2]; p[2]++) {
            if (volnum[lin(p)] < 0) {
              for (int d1 = -1; d1 <= 1; d1++)
                for (int d2 = -1; d2 <= 1; d2++)
                  for (int d3 = -1; d3 <= 1; d3++) {
                    int q[3] = {p[0] + d1, p[1] + d2, p[2] + d3};
                    pbc(q);
                    if (known[lin(q)] == 2) known[lin(q)] = 0;
                  }
            }
          }
    }
    if (nedge_out) *nedge_out = num_edge;

    long num_reassign = 0;
    for (p[0] = 0; p[0] < n[0]; p[0]++)
      for (p[1] = 0; p[1] < n[1]; p[1]++)
        for (p[2] = 0; p[2] < n[2]; p[2]++) {
          const int id = lin(p);
          const int bvolnum = volnum[id];
          if (bvolnum < 0) {
            int q[3] = {p[0], p[1], p[2]};
            max_neargrid(q);
            const int path_volnum = volnum[lin(q)];
            if (path_volnum < 0 || path_volnum > nbasin) return -1;
            volnum[id] = path_volnum;
            if (std::abs(bvolnum) != path_volnum) {
              num_reassign++;
              volnum[id] = -path_volnum;
            }
            for (int i = 0; i < pnum; i++)
              if (known[path[i]] != 2) known[path[i]] = 0;
          }
        }
    return num_reassign;
  }
};

// ---------------------------------------------------------------------------
// Minimal crystal services used by the attractor identification
// (crystalmod@proc.f90:1054-1085 shortest, :1118-1133 are_lclose,
//  crystalmod@env.f90:593-614 identify_atom -> nearest_atom).
// Restated by brute force over the 125 neighbouring lattice translations;
// adequate for the moderately skewed test cells used here.
// ---------------------------------------------------------------------------
struct Cell {
  double x2c[9];
  void tocart(const double x[3], double c[3]) const {
    for (int i = 0; i < 3; i++) c[i] = x2c[i] * x[0] + x2c[i + 3] * x[1] + x2c[i + 6] * x[2];
  }
  double shortest(const double dx_in[3]) const {
    double dx[3];
    for (int i = 0; i < 3; i++) dx[i] = dx_in[i] - std::round(dx_in[i]);
    double best = 1e300;
    for (int a = -2; a <= 2; a++)
      for (int b = -2; b <= 2; b++)
        for (int c = -2; c <= 2; c++) {
          const double t[3] = {dx[0] + a, dx[1] + b, dx[2] + c};
          double ct[3];
          tocart(t, ct);
          const double d = std::sqrt(ct[0] * ct[0] + ct[1] * ct[1] + ct[2] * ct[2]);
          if (d < best) best = d;
        }
    return best;
  }
  bool are_lclose(const double x0[3], const double x1[3], double eps) const {
    const double dx[3] = {x0[0] - x1[0], x0[1] - x1[1], x0[2] - x1[2]};
    return shortest(dx) < eps;
  }
  // returns 1-based atom id or 0
  int identify_atom(const double x[3], int nat, const double* xat, double distmax) const {
    int best = 0;
    double dbest = 1e300;
    for (int i = 0; i < nat; i++) {
      const double dx[3] = {x[0] - xat[3 * i], x[1] - xat[3 * i + 1], x[2] - xat[3 * i + 2]};
      const double d = shortest(dx);
      if (d < dbest) {
        dbest = d;
        best = i + 1;
      }
    }
    if (best > 0 && dbest <= distmax) return best;
    return 0;
  }
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------
// orc_bader_integrate: bader_integrate, src/bader@proc.f90:80-234 (scan :151-215,
// refine :218-224).  DISCARD expressions are not restated (host expression
// evaluator; out of scope).
//  in : f(n1,n2,n3), car2lat[9], lat_i_dist[27] (host-computed as in :124-145),
//       x2c[9] (crystal cell, for identify_atom/are_lclose), atexist, nat atoms
//       xat(3,nat) (cryst.), ratom.
//  out: idg(n1,n2,n3) (1-based basin ids), *nattr, xattr(3,maxattr) (cryst.),
//       stats[0]=refine iterations, [1]=edge points at iteration 1,
//       [2]=near-grid steps.
//  returns 0, or <0 on error (-1 refine found new maximum, -2 maxattr too small)
// ---------------------------------------------------------------------------
int orc_bader_integrate(const double* f, const int* n, const double* car2lat, const double* lat_i_dist,
                        const double* x2c, int atexist, int nat, const double* xat, double ratom,
                        int* idg, int* nattr, double* xattr, int maxattr, long* stats) {
  Bader b;
  Cell cell;
  std::memcpy(cell.x2c, x2c, sizeof(cell.x2c));
  for (int i = 0; i < 3; i++) b.n[i] = n[i];
  b.f = f;
  std::memcpy(b.car2lat, car2lat, sizeof(b.car2lat));
  std::memcpy(b.lat_i_dist, lat_i_dist, sizeof(b.lat_i_dist));
  const size_t nn = (size_t)n[0] * n[1] * n[2];
  b.volnum.assign(nn, 0);
  b.known.assign(nn, 0);
  b.nbasin = 0;
  if (atexist) {  // :102-111
    if (nat > maxattr) return -2;
    b.nbasin = nat;
    for (int i = 0; i < 3 * nat; i++) xattr[i] = xat[i];
  }
  int p[3];
  // scan: do i; do j; do k  (:151-153) -- k (third index) innermost
  for (int i = 0; i < n[0]; i++)
    for (int j = 0; j < n[1]; j++)
      for (int k = 0; k < n[2]; k++) {
        p[0] = i; p[1] = j; p[2] = k;
        if (b.volnum[b.lin(p)] != 0) continue;
        b.max_neargrid(p);
        int path_volnum = b.volnum[b.lin(p)];
        if (path_volnum == 0) {  // new maximum (:160-199)
          const double dv[3] = {(double)p[0] / n[0], (double)p[1] / n[1], (double)p[2] / n[2]};
          bool isassigned = false;
          if (atexist) {
            const int nid = cell.identify_atom(dv, nat, xat, ratom);
            if (nid > 0) {
              path_volnum = nid;
              isassigned = true;
            }
          }
          if (!isassigned && ratom > VSMALL) {
            for (int l = 0; l < b.nbasin; l++)
              if (cell.are_lclose(dv, &xattr[3 * l], ratom)) {
                path_volnum = l + 1;
                isassigned = true;
                break;
              }
          }
          if (!isassigned) {
            if (b.nbasin + 1 > maxattr) return -2;
            b.nbasin++;
            path_volnum = b.nbasin;
            xattr[3 * (b.nbasin - 1) + 0] = dv[0];
            xattr[3 * (b.nbasin - 1) + 1] = dv[1];
            xattr[3 * (b.nbasin - 1) + 2] = dv[2];
          }
        }
        // assign all points along the trajectory (:202-211)
        for (int l = 0; l < b.pnum; l++) {
          const int id = b.path[l];
          int pt[3];
          b.unlin(id, pt);
          if (b.volnum[id] != -1) b.volnum[id] = path_volnum;
          if (b.known[id] != 2) b.known[id] = 0;
          b.assign_surrounding_pts(pt);
        }
      }
  // refine (:218-224)
  int ref_itrs = 1;
  long nedge1 = 0;
  while (true) {
    long nedge = 0;
    const long nre = b.refine_edge(ref_itrs, &nedge);
    if (ref_itrs == 1) nedge1 = nedge;
    if (nre < 0) return -1;
    if (nre == 0) break;
    ref_itrs++;
  }
  for (size_t i = 0; i < nn; i++) idg[i] = b.volnum[i];
  *nattr = b.nbasin;
  if (stats) {
    stats[0] = ref_itrs;
    stats[1] = nedge1;
    stats[2] = b.nsteps;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// orc_bader_canonical: the order-independent labelling every point would get
// from its OWN complete near-grid trajectory started with dr = 0 on a fresh
// grid (max_neargrid/step_neargrid, bader@proc.f90:427-494, with no known==2
// points).  term[i] = 0-based linear id of the terminal maximum.  This is the
// property the parallel implementation computes; it is compared with
// orc_bader_integrate's labels in tests.  OpenMP over start points (checker
// only; the reference routine is serial).  stats[0] = total steps,
// stats[1] = longest path, stats[2] = #steps that landed on a point with
// rho <= running path maximum, stats[3] = #revisits (on-grid fallbacks).
// ---------------------------------------------------------------------------
int orc_bader_canonical(const double* f, const int* n, const double* car2lat, const double* lat_i_dist,
                        int* term, long* stats) {
  const long nn = (long)n[0] * n[1] * n[2];
  long tot = 0, longest = 0, nle = 0, nrev = 0;
#pragma omp parallel reduction(+ : tot, nle, nrev) reduction(max : longest)
  {
    Bader b;
    for (int i = 0; i < 3; i++) b.n[i] = n[i];
    b.f = f;
    std::memcpy(b.car2lat, car2lat, sizeof(b.car2lat));
    std::memcpy(b.lat_i_dist, lat_i_dist, sizeof(b.lat_i_dist));
    std::vector<int> visited;  // linear ids on the path (the known==1 set)
#pragma omp for schedule(dynamic, 4096)
    for (long s = 0; s < nn; s++) {
      int p[3];
      b.unlin((int)s, p);
      visited.clear();
      double dr[3] = {0, 0, 0};
      double rhomax = -1e300;
      long len = 0;
      while (true) {
        // one step_neargrid with an explicit visited set instead of the `known` array
        double g[3];
        b.rho_grad_dir(p, g);
        int pm[3];
        const double gmax = std::max(std::fabs(g[0]), std::max(std::fabs(g[1]), std::fabs(g[2])));
        const int idp = b.lin(p);
        bool moved = true;
        if (gmax < 1e-30) {
          dr[0] = dr[1] = dr[2] = 0.0;
          if (b.is_max(p)) {
            moved = false;
          } else {
            pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
            b.step_ongrid(pm);
          }
        } else {
          const double coeff = 1.0 / gmax;
          for (int i = 0; i < 3; i++) {
            g[i] = coeff * g[i];
            const int ng = nint_(g[i]);
            pm[i] = p[i] + ng;
            dr[i] = dr[i] + g[i] - (double)ng;
            const int nd = nint_(dr[i]);
            pm[i] = pm[i] + nd;
            dr[i] = dr[i] - (double)nd;
          }
        }
        if (!moved) break;
        len++;
        visited.push_back(idp);
        rhomax = std::max(rhomax, f[idp]);
        b.pbc(pm);
        const int idm = b.lin(pm);
        if (f[idm] <= rhomax) nle++;
        if (std::find(visited.begin(), visited.end(), idm) != visited.end()) {
          pm[0] = p[0]; pm[1] = p[1]; pm[2] = p[2];
          b.step_ongrid(pm);
          dr[0] = dr[1] = dr[2] = 0.0;
          nrev++;
        }
        if (b.lin(pm) == idp) break;  // did not move: maximum
        p[0] = pm[0]; p[1] = pm[1]; p[2] = pm[2];
      }
      term[s] = b.lin(p);
      tot += len;
      longest = std::max(longest, len);
    }
  }
  if (stats) {
    stats[0] = tot;
    stats[1] = longest;
    stats[2] = nle;
    stats[3] = nrev;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// orc_integrate_bader: intgrid_fields, Bader branch,
// src/integration@proc.f90:1208-1218 (volume: count(idg==i)*omega/ntot) and
// :1289-1299 (padd = sum(fint, idg==i)*omega/ntot; naive sum in array element
// order).  OpenMP over attractors exactly where the reference has it.
//  fields: nprop pointers to fint(n1,n2,n3); psum(nattr,nprop) column-major
// ---------------------------------------------------------------------------
void orc_integrate_bader(const int* idg, const int* n, int nattr, int nprop, const double* const* fields,
                         double omega, double* vol, double* psum) {
  const long nn = (long)n[0] * n[1] * n[2];
  const double ntot = (double)nn;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 1; i <= nattr; i++) {
    long cnt = 0;
    for (long q = 0; q < nn; q++)
      if (idg[q] == i) cnt++;
    vol[i - 1] = (double)cnt * omega / ntot;
    for (int k = 0; k < nprop; k++) {
      const double* ff = fields[k];
      double s = 0.0;
      for (long q = 0; q < nn; q++)
        if (idg[q] == i) s += ff[q];
      psum[(i - 1) + (long)nattr * k] = s * omega / ntot;
    }
  }
}

// ---------------------------------------------------------------------------
// orc_qcksort_r8: qcksort_r8, src/tools@proc.f90:83-168, bit-for-bit including
// the LCG pivot (fm=7875, fa=211, fc=1663, fmi=1.2698413e-4), m=7 insertion
// cut-off and the explicit stack (nstack=50).  iord is 1-based (values and
// positions) like the Fortran; arr is indexed arr[iord-1].
// returns 0 or -1 ("Increase nstack").
// ---------------------------------------------------------------------------
int orc_qcksort_r8(const double* arr, int* iord, int first, int last) {
  const int m = 7, nstack = 50;
  const double fm = 7875.0, fa = 211.0, fc = 1663.0, fmi = 1.2698413e-4;
  int istack[nstack + 1];
  int jstack = 0, l = first, ir = last, i, j, na, iq;
  double fx = 0.0, a;
  auto A = [&](int k) { return arr[k - 1]; };   // arr(k)
  auto IO = [&](int k) -> int& { return iord[k - 1]; };  // iord(k)
  while (true) {
    if (ir - l < m) {
      for (j = l + 1; j <= ir; j++) {
        na = IO(j);
        a = A(na);
        for (i = j - 1; i >= first; i--) {
          if (A(IO(i)) <= a) goto L12;
          IO(i + 1) = IO(i);
        }
        i = first - 1;
      L12:
        IO(i + 1) = na;
      }
      if (jstack == 0) return 0;
      ir = istack[jstack];
      l = istack[jstack - 1];
      jstack -= 2;
    } else {
      i = l;
      j = ir;
      fx = std::fmod(fx * fa + fc, fm);
      iq = l + (int)((ir - l + 1) * (fx * fmi));
      na = IO(iq);
      a = A(na);
      IO(iq) = IO(l);
      while (true) {
        while (j >= first && a < A(IO(j))) j--;   // label 21
        if (j <= i) {
          IO(i) = na;
          break;  // goto 30
        }
        IO(i) = IO(j);
        i++;
        while (i <= last && a > A(IO(i))) i++;    // label 22
        if (j <= i) {
          IO(j) = na;
          i = j;
          break;  // goto 30
        }
        IO(j) = IO(i);
        j--;
      }
      jstack += 2;
      if (jstack > nstack) return -1;
      if (ir - i >= i - l) {
        istack[jstack] = ir;
        istack[jstack - 1] = i + 1;
        ir = i - 1;
      } else {
        istack[jstack] = i - 1;
        istack[jstack - 1] = l;
        l = i + 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Yu-Trinkle.  yt_integrate, src/yt@proc.f90:38-224; yt_weights :399-530.
// The ytdata record (yt.f90:36-45) is returned in caller-allocated arrays:
//   nlo(nn), ibasin(nn), iio(nn) (1-based rank of spatial point), inear(nvec,nn)
//   (1-based ranks), fnear(nvec,nn).  All rank-indexed like the reference.
//  stable != 0 replaces qcksort by a stable sort (ties by linear index): the
//  order the GPU path defines; identical whenever the data has no ties.
// ---------------------------------------------------------------------------
int orc_yt_integrate(const double* f, const int* n, int nvec, const int* vec, const double* area,
                     const double* x2c, int atexist, int nat, const double* xat, double ratom, int stable,
                     int* nlo, int* ibasin, int* iio, int* inear, double* fnear, int* nattr, double* xattr,
                     int maxattr) {
  Cell cell;
  std::memcpy(cell.x2c, x2c, sizeof(cell.x2c));
  const int n1 = n[0], n2 = n[1], n3 = n[2];
  const long nn = (long)n1 * n2 * n3;
  const double* g = f;  // g = reshape(bas%f) (:82-83)
  std::vector<int> io(nn);
  for (long i = 0; i < nn; i++) io[i] = (int)i + 1;
  if (stable) {
    std::stable_sort(io.begin(), io.end(), [&](int a, int b) { return g[a - 1] < g[b - 1]; });
  } else {
    if (orc_qcksort_r8(g, io.data(), 1, (int)nn) != 0) return -1;  // (:86-90)
  }
  for (long i = 0; i < nn; i++) iio[io[i] - 1] = (int)i + 1;  // (:91-93)

  int nat_attr = 0;
  if (atexist) {  // (:68-75)
    if (nat > maxattr) return -2;
    nat_attr = nat;
    for (int i = 0; i < 3 * nat; i++) xattr[i] = xat[i];
  }
  std::vector<int> ihi(nvec);
  std::vector<double> chi(nvec);
  std::fill(ibasin, ibasin + nn, 0);
  std::fill(nlo, nlo + nn, 0);
  std::fill(inear, inear + (size_t)nvec * nn, 0);
  std::fill(fnear, fnear + (size_t)nvec * nn, 0.0);
  auto modulo = [](int a, int b) { int r = a % b; return r < 0 ? r + b : r; };
  for (long ii = nn; ii >= 1; ii--) {  // (:108-188)
    int nhi = 0;
    const int i = io[ii - 1];
    const int k0 = i - 1;
    const int ib[3] = {k0 % n1, (k0 / n1) % n2, (k0 / (n1 * n2)) % n3};  // to3 (0-based)
    double csum = 0.0;
    for (int k = 0; k < nvec; k++) {
      const int jb0 = modulo(ib[0] + vec[3 * k + 0], n1);
      const int jb1 = modulo(ib[1] + vec[3 * k + 1], n2);
      const int jb2 = modulo(ib[2] + vec[3 * k + 2], n3);
      const int j = jb0 + n1 * (jb1 + n2 * jb2) + 1;  // to1
      const int jj = iio[j - 1];
      if (jj > ii) {
        ihi[nhi] = jj;
        chi[nhi] = std::max(area[k] * (g[j - 1] - g[i - 1]), VSMALL);
        csum = csum + chi[nhi];
        nhi++;
      }
    }
    nlo[ii - 1] = 0;
    if (nhi == 0) {  // local maximum (:129-168)
      const double dv[3] = {(double)ib[0] / n1, (double)ib[1] / n2, (double)ib[2] / n3};
      bool isassigned = false;
      if (atexist) {
        const int nid = cell.identify_atom(dv, nat, xat, ratom);
        if (nid > 0) {
          ibasin[ii - 1] = nid;
          isassigned = true;
        }
      }
      if (!isassigned && ratom > VSMALL) {
        for (int k = 0; k < nat_attr; k++)
          if (cell.are_lclose(dv, &xattr[3 * k], ratom)) {
            ibasin[ii - 1] = k + 1;
            isassigned = true;
            break;
          }
      }
      if (!isassigned) {
        if (nat_attr + 1 > maxattr) return -2;
        nat_attr++;
        ibasin[ii - 1] = nat_attr;
        xattr[3 * (nat_attr - 1) + 0] = dv[0];
        xattr[3 * (nat_attr - 1) + 1] = dv[1];
        xattr[3 * (nat_attr - 1) + 2] = dv[2];
      }
    } else {
      bool isias = (ibasin[ihi[0] - 1] == 0);  // (:170-173)
      for (int k = 0; k < nhi; k++) isias = isias || (ibasin[ihi[k] - 1] != ibasin[ihi[0] - 1]);
      if (!isias) {
        ibasin[ii - 1] = ibasin[ihi[0] - 1];
      } else {
        ibasin[ii - 1] = 0;
        for (int k = 0; k < nhi; k++) {  // (:179-185)
          const int kk = ihi[k];
          nlo[kk - 1]++;
          inear[(size_t)(nlo[kk - 1] - 1) + (size_t)nvec * (kk - 1)] = (int)ii;
          fnear[(size_t)(nlo[kk - 1] - 1) + (size_t)nvec * (kk - 1)] = chi[k] / std::max(csum, VSMALL);
        }
      }
    }
  }
  *nattr = nat_attr;
  return 0;
}

// ---------------------------------------------------------------------------
// yt_isosurface, src/yt@proc.f90:233-390 (ISOSURFACE keyword): regions of the super-level set f >= isov under the
// Voronoi-stencil connectivity, numbered like the reference does -- including its two quirks:
//  * imap(b) is OVERWRITTEN by every later contact of region b (:327-331), so an earlier merge of b can be lost;
//  * the surviving regions keep their discovery numbers (:337-351: no renumbering), while bas%nattr becomes the
//    NUMBER of survivors and bas%xattr is cut to that many columns (:359).
// The DISCARD expression (:304-311) is not restated: it is evaluated by the host's expression parser.
// Out: idg(n1,n2,n3) spatial region ids (0 below isov), nraw = regions before merging (bas%nattr at :334),
// nattr = survivors, xattr(3,nraw) = coordinates of every regional maximum in discovery order.
// ---------------------------------------------------------------------------
int orc_yt_isosurface(const double* f, const int* n, int nvec, const int* vec, double isov, int stable, int* idg, int* nraw,
                      int* nattr, double* xattr, int maxattr) {
  const int n1 = n[0], n2 = n[1], n3 = n[2];
  const long nn = (long)n1 * n2 * n3;
  const double* g = f;
  std::vector<int> io(nn), iio(nn);
  for (long i = 0; i < nn; i++) io[i] = (int)i + 1;
  if (stable) {
    std::stable_sort(io.begin(), io.end(), [&](int a, int b) { return g[a - 1] < g[b - 1]; });
  } else {
    if (orc_qcksort_r8(g, io.data(), 1, (int)nn) != 0) return -1;  // (:276-280)
  }
  for (long i = 0; i < nn; i++) iio[io[i] - 1] = (int)i + 1;
  std::vector<int> ibasin(nn, 0), ihi(nvec), imap;
  imap.reserve(1024);
  int na = 0;
  auto modulo = [](int a, int b) { int r = a % b; return r < 0 ? r + b : r; };
  for (long ii = nn; ii >= 1; ii--) {  // (:295-333)
    const int i = io[ii - 1];
    if (g[i - 1] < isov) continue;
    const int k0 = i - 1;
    const int ib[3] = {k0 % n1, (k0 / n1) % n2, (k0 / (n1 * n2)) % n3};
    int nhi = 0;
    for (int k = 0; k < nvec; k++) {
      const int j = modulo(ib[0] + vec[3 * k], n1) + n1 * (modulo(ib[1] + vec[3 * k + 1], n2) + n2 * modulo(ib[2] + vec[3 * k + 2], n3)) + 1;
      const int jj = iio[j - 1];
      if (jj > ii) ihi[nhi++] = jj;
    }
    if (nhi == 0) {
      if (na + 1 > maxattr) return -2;
      na++;
      imap.push_back(0);
      ibasin[ii - 1] = na;
      xattr[3 * (na - 1)] = (double)ib[0] / n1;
      xattr[3 * (na - 1) + 1] = (double)ib[1] / n2;
      xattr[3 * (na - 1) + 2] = (double)ib[2] / n3;
    } else {
      bool interior = true;
      for (int k = 1; k < nhi; k++) interior = interior && (ibasin[ihi[k] - 1] == ibasin[ihi[0] - 1]);
      if (interior) {
        ibasin[ii - 1] = ibasin[ihi[0] - 1];
      } else {
        int imin = ibasin[ihi[0] - 1];
        for (int k = 1; k < nhi; k++) imin = std::min(imin, ibasin[ihi[k] - 1]);
        ibasin[ii - 1] = imin;
        for (int k = 0; k < nhi; k++)
          if (ibasin[ihi[k] - 1] != imin) imap[ibasin[ihi[k] - 1] - 1] = imin;
      }
    }
  }
  *nraw = na;
  std::vector<int> root(na + 1, 0);
  int cnt = 0;
  for (int i = 1; i <= na; i++) {  // (:337-351): `where (ibasin == i) ibasin = ii`, as one composed map
    int r = i;
    if (imap[i - 1] == 0) cnt++;
    else
      while (imap[r - 1] != 0) r = imap[r - 1];
    root[i] = r;
  }
  // the `where` passes run for i = 1, 2, ...: a point relabelled to ii < i is not touched again (ii's pass is over),
  // and roots are never relabelled, so the composition is exactly root[]
  *nattr = cnt;
  for (long q = 0; q < nn; q++) idg[q] = root[ibasin[iio[q] - 1]];  // (:356)
  return 0;
}

// yt_weights for one basin (yt@proc.f90:476-499 / :502-524): w is spatial (n1,n2,n3).
void orc_yt_weights(long nn, int nvec, const int* nlo, const int* ibasin, const int* iio, const int* inear,
                    const double* fnear, int idb, double* w) {
  std::vector<double> waux(nn);
  for (long j = 0; j < nn; j++) waux[j] = (ibasin[j] == idb) ? 1.0 : 0.0;
  for (long j = nn; j >= 1; j--) {
    if (std::fabs(waux[j - 1]) > 0.0) {
      for (int k = 0; k < nlo[j - 1]; k++) {
        const size_t q = (size_t)k + (size_t)nvec * (j - 1);
        waux[inear[q] - 1] = waux[inear[q] - 1] + fnear[q] * waux[j - 1];
      }
    }
  }
  for (long s = 0; s < nn; s++) w[s] = waux[iio[s] - 1];
}

// intgrid_fields, YT branch (integration@proc.f90:1208-1218, :1289-1299):
// vol(i) = sum(w)*omega/ntot ; psum(i,k) = sum(w*fint_k)*omega/ntot, per attractor,
// OpenMP over attractors (firstprivate(w)) as in the reference.
void orc_integrate_yt(long nn, int nvec, const int* nlo, const int* ibasin, const int* iio, const int* inear,
                      const double* fnear, int nattr, int nprop, const double* const* fields, double omega,
                      double* vol, double* psum) {
  const double ntot = (double)nn;
#pragma omp parallel
  {
    std::vector<double> w(nn);
#pragma omp for schedule(dynamic, 1)
    for (int i = 1; i <= nattr; i++) {
      orc_yt_weights(nn, nvec, nlo, ibasin, iio, inear, fnear, i, w.data());
      double s = 0.0;
      for (long q = 0; q < nn; q++) s += w[q];
      vol[i - 1] = s * omega / ntot;
      for (int k = 0; k < nprop; k++) {
        const double* ff = fields[k];
        double t = 0.0;
        for (long q = 0; q < nn; q++) t += w[q] * ff[q];
        psum[(i - 1) + (long)nattr * k] = t * omega / ntot;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Tricubic interpolation (Lekien & Marsden), grinterp_tricubic,
// src/grid3mod@proc.f90:2649-2821, with the 64x64 matrix `c` (:76-340).
// The matrix is not copied: it is regenerated as the inverse of the constraint
// matrix B (value, fx, fy, fz, fxy, fxz, fyz, fxyz at the 8 cell corners applied
// to the monomials x^i y^j z^k, a-index l = i + 4j + 16k), which is the
// published definition; tests compare it entry by entry with the reference's
// table when /root/reference is available.
// ---------------------------------------------------------------------------
static double g_tricubic_c[64][64];  // c(l, m): a(l) = sum_m c(l,m) b(m)
static bool g_tricubic_ready = false;

static void build_tricubic_matrix() {
  if (g_tricubic_ready) return;
  // 1-D cubic Hermite: coefficients of x^0..x^3 from (f0, f1, d0, d1)
  static const int H[4][4] = {{1, 0, 0, 0}, {0, 0, 1, 0}, {-3, 3, -2, -1}, {2, -2, 1, 1}};
  // b ordering (:2685-2764): blocks of 8 = [f, fx, fy, fz, fxy, fxz, fyz, fxyz]; inside a
  // block the corner index is cx + 2*cy + 4*cz.
  static const int dflag[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  for (int l = 0; l < 64; l++) {
    const int i = l & 3, j = (l >> 2) & 3, k = (l >> 4) & 3;
    for (int m = 0; m < 64; m++) {
      const int blk = m >> 3, cor = m & 7;
      const int cx = cor & 1, cy = (cor >> 1) & 1, cz = (cor >> 2) & 1;
      // 1-D Hermite input slot: value at corner c -> c ; derivative at corner c -> 2 + c
      const int sx = dflag[blk][0] * 2 + cx, sy = dflag[blk][1] * 2 + cy, sz = dflag[blk][2] * 2 + cz;
      g_tricubic_c[l][m] = (double)(H[i][sx] * H[j][sy] * H[k][sz]);
    }
  }
  g_tricubic_ready = true;
}

void orc_tricubic_matrix(double* c_out /* 64*64, c_out[l + 64*m] = c(l+1,m+1) */) {
  build_tricubic_matrix();
  for (int l = 0; l < 64; l++)
    for (int m = 0; m < 64; m++) c_out[l + 64 * m] = g_tricubic_c[l][m];
}

// grid_floor (:3118-3136) + grinterp_tricubic (:2660-2819) for a periodic grid.
// xi: crystallographic, already modulo 1.  y, yp(3), ypp(3,3 col-major) in grid-cryst. coords.
static void tricubic_eval(const double* f, const int* n, const double xi[3], double* y, double yp[3],
                          double ypp[9]) {
  build_tricubic_matrix();
  auto modulo = [](int a, int b) { int r = a % b; return r < 0 ? r + b : r; };
  int idx[3];  // 0-based floor index
  for (int d = 0; d < 3; d++) idx[d] = modulo((int)std::floor(xi[d] * n[d]), n[d]);
  double g[4][4][4];  // g(i,j,k) with offsets -1..2 stored at +1
  for (int i = -1; i <= 2; i++)
    for (int j = -1; j <= 2; j++)
      for (int k = -1; k <= 2; k++) {
        const int a = modulo(idx[0] + i, n[0]), b = modulo(idx[1] + j, n[1]), c = modulo(idx[2] + k, n[2]);
        g[i + 1][j + 1][k + 1] = f[a + n[0] * (b + (long)n[1] * c)];
      }
  auto G = [&](int i, int j, int k) { return g[i + 1][j + 1][k + 1]; };
  double b[64];
  int m = 0;
  // f
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++) b[m++] = G(cx, cy, cz);
  // fx
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.5 * (G(cx + 1, cy, cz) - G(cx - 1, cy, cz));
  // fy
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.5 * (G(cx, cy + 1, cz) - G(cx, cy - 1, cz));
  // fz
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.5 * (G(cx, cy, cz + 1) - G(cx, cy, cz - 1));
  // fxy (:2724-2731): g(+,+) - g(-,+) - g(+,-) + g(-,-), left to right
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.25 * (G(cx + 1, cy + 1, cz) - G(cx - 1, cy + 1, cz) - G(cx + 1, cy - 1, cz) + G(cx - 1, cy - 1, cz));
  // fxz
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.25 * (G(cx + 1, cy, cz + 1) - G(cx - 1, cy, cz + 1) - G(cx + 1, cy, cz - 1) + G(cx - 1, cy, cz - 1));
  // fyz
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.25 * (G(cx, cy + 1, cz + 1) - G(cx, cy - 1, cz + 1) - G(cx, cy + 1, cz - 1) + G(cx, cy - 1, cz - 1));
  // fxyz (:2757-2764)
  for (int cz = 0; cz < 2; cz++) for (int cy = 0; cy < 2; cy++) for (int cx = 0; cx < 2; cx++)
    b[m++] = 0.125 * (G(cx + 1, cy + 1, cz + 1) - G(cx - 1, cy + 1, cz + 1) - G(cx + 1, cy - 1, cz + 1) +
                      G(cx - 1, cy - 1, cz + 1) - G(cx + 1, cy + 1, cz - 1) + G(cx - 1, cy + 1, cz - 1) +
                      G(cx + 1, cy - 1, cz - 1) - G(cx - 1, cy - 1, cz - 1));
  double a[64];
  for (int l = 0; l < 64; l++) {  // a = matmul(c,b) (:2769)
    double s = 0.0;
    for (int q = 0; q < 64; q++) s += g_tricubic_c[l][q] * b[q];
    a[l] = s;
  }
  double x[3];
  for (int d = 0; d < 3; d++) x[d] = xi[d] * n[d] - (double)idx[d];  // (:2772) with idx-1 -> 0-based idx
  double aa[4], aax[4], aay[4], aaxy[4], aaxx[4], aayy[4], bb[4], bbx[4], bbxx[4];
  int l = 0;
  for (int k = 0; k < 4; k++) {
    for (int j = 0; j < 4; j++) {
      bb[j] = a[l] + x[0] * (a[l + 1] + x[0] * (a[l + 2] + x[0] * a[l + 3]));
      bbx[j] = a[l + 1] + x[0] * (2.0 * a[l + 2] + x[0] * 3.0 * a[l + 3]);
      bbxx[j] = 2.0 * a[l + 2] + 6.0 * x[0] * a[l + 3];
      l += 4;
    }
    aa[k] = bb[0] + x[1] * (bb[1] + x[1] * (bb[2] + x[1] * bb[3]));
    aax[k] = bbx[0] + x[1] * (bbx[1] + x[1] * (bbx[2] + x[1] * bbx[3]));
    aay[k] = bb[1] + x[1] * (2.0 * bb[2] + x[1] * 3.0 * bb[3]);
    aaxy[k] = bbx[1] + x[1] * (2.0 * bbx[2] + x[1] * 3.0 * bbx[3]);
    aaxx[k] = bbxx[0] + x[1] * (bbxx[1] + x[1] * (bbxx[2] + x[1] * bbxx[3]));
    aayy[k] = 2.0 * bb[2] + 6.0 * x[1] * bb[3];
  }
  *y = aa[0] + x[2] * (aa[1] + x[2] * (aa[2] + x[2] * aa[3]));
  yp[0] = aax[0] + x[2] * (aax[1] + x[2] * (aax[2] + x[2] * aax[3]));
  yp[1] = aay[0] + x[2] * (aay[1] + x[2] * (aay[2] + x[2] * aay[3]));
  yp[2] = aa[1] + x[2] * (2.0 * aa[2] + x[2] * 3.0 * aa[3]);
  double h[3][3];
  h[0][0] = aaxx[0] + x[2] * (aaxx[1] + x[2] * (aaxx[2] + x[2] * aaxx[3]));
  h[0][1] = aaxy[0] + x[2] * (aaxy[1] + x[2] * (aaxy[2] + x[2] * aaxy[3]));
  h[0][2] = aax[1] + x[2] * (2.0 * aax[2] + x[2] * 3.0 * aax[3]);
  h[1][1] = aayy[0] + x[2] * (aayy[1] + x[2] * (aayy[2] + x[2] * aayy[3]));
  h[1][2] = aay[1] + x[2] * (2.0 * aay[2] + x[2] * 3.0 * aay[3]);
  h[2][2] = 2.0 * aa[2] + 6.0 * x[2] * aa[3];
  for (int i = 0; i < 3; i++) {  // (:2811-2817)
    yp[i] = yp[i] * n[i];
    for (int j = i; j < 3; j++) {
      h[i][j] = h[i][j] * n[i] * n[j];
      h[j][i] = h[i][j];
    }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) ypp[i + 3 * j] = h[i][j];
}

// grid3%interp for a periodic tricubic grid (grid3mod@proc.f90:1696-1752): modulo, evaluate,
// then yp = matmul(transpose(c2xl),yp), ypp = matmul(matmul(transpose(c2xl),ypp),c2xl).
void orc_grid_interp_tricubic(const double* f, const int* n, const double* c2xl, const double* xi_in,
                              double* y, double* yp, double* ypp) {
  double xi[3];
  for (int d = 0; d < 3; d++) {
    xi[d] = xi_in[d] - std::floor(xi_in[d]);  // modulo(xi,1d0)
    if (xi[d] >= 1.0) xi[d] = 0.0;
  }
  double yp0[3], h0[9];
  tricubic_eval(f, n, xi, y, yp0, h0);
  for (int i = 0; i < 3; i++) {
    double s = 0.0;
    for (int j = 0; j < 3; j++) s += c2xl[j + 3 * i] * yp0[j];  // transpose(c2xl)(i,j) = c2xl(j,i)
    yp[i] = s;
  }
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += c2xl[k + 3 * i] * h0[k + 3 * j];
      t[i + 3 * j] = s;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += t[i + 3 * k] * c2xl[k + 3 * j];
      ypp[i + 3 * j] = s;
    }
}

// eigenvalues of a real symmetric 3x3 matrix, ascending (stands in for eigsym = LAPACK
// dsyev, tools_math@proc.f90:1005-1039; only sign(ehess(2)) is consumed, nci@proc.f90:599).
// Cyclic Jacobi.
static void eig3_sym(const double hin[9], double ev[3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) a[i][j] = 0.5 * (hin[i + 3 * j] + hin[j + 3 * i]);
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
      }
  }
  ev[0] = a[0][0]; ev[1] = a[1][1]; ev[2] = a[2][2];
  std::sort(ev, ev + 3);
}

// ---------------------------------------------------------------------------
// orc_nci_rdg: the NCIPLOT hot loop in "grid" (tricubic) mode,
// src/nci@proc.f90:543-605 with f%grd for type_grid (fieldmod@proc.f90:921-974,
// :1148-1155) and nder = 2.
//  x0(3) Cartesian origin, xmat(3,3) step vectors (Cartesian, column-major),
//  nstep(3); c2x/x2c of the crystal; c2xl of the grid (= c2x for a full-cell grid);
//  nuclei: nnuc Cartesian positions used for the "on a nucleus => zero gradient"
//  rule (distmax = 1e-5 bohr), tested against the minimum image.
//  out: crho, cgrad with index k fastest: [k + nstep3*(j + nstep2*i)]; lam2 (optional).
// OpenMP over i with schedule(dynamic) as in the reference (:543-544).
// ---------------------------------------------------------------------------
void orc_nci_rdg(const double* f, const int* n, const double* x0, const double* xmat, const int* nstep,
                 const double* c2x, const double* x2c, const double* c2xl, int nnuc, const double* nuc_cart,
                 double* crho, double* cgrad, double* lam2) {
  const double pi = 3.14159265358979323846264338328;  // param.F90 pi
  const double cst = 2.0 * std::pow(3.0 * pi * pi, 1.0 / 3.0);  // nci@proc.f90:91
  const double fthirds = 4.0 / 3.0;
  const double flooreps = 1e-4;  // fieldmod@proc.f90:31
  Cell cell;
  std::memcpy(cell.x2c, x2c, sizeof(cell.x2c));
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nstep[0]; i++)
    for (int j = 0; j < nstep[1]; j++)
      for (int k = 0; k < nstep[2]; k++) {
        double x[3], wx[3];
        for (int d = 0; d < 3; d++)  // x = x0 + i*xmat(:,1) + j*xmat(:,2) + k*xmat(:,3) (:548)
          x[d] = ((x0[d] + i * xmat[d + 0]) + j * xmat[d + 3]) + k * xmat[d + 6];
        for (int d = 0; d < 3; d++) {  // wx = matmul(m_c2x, x)
          double s = c2x[d + 0] * x[0];
          s = s + c2x[d + 3] * x[1];
          s = s + c2x[d + 6] * x[2];
          wx[d] = s;
        }
        for (int d = 0; d < 3; d++)  // (:923-929)
          if (wx[d] < -flooreps || wx[d] > 1.0 + flooreps) wx[d] = wx[d] - std::floor(wx[d]);
        double y, yp[3], ypp[9];
        orc_grid_interp_tricubic(f, n, c2xl, wx, &y, yp, ypp);
        // nucleus rule (:1148-1155): wc = x2c(wx); identify_atom(wc, distmax=1d-5)
        bool isnuc = false;
        if (nnuc > 0) {
          double wc[3];
          cell.tocart(wx, wc);
          for (int a = 0; a < nnuc && !isnuc; a++) {
            const double dc[3] = {wc[0] - nuc_cart[3 * a], wc[1] - nuc_cart[3 * a + 1], wc[2] - nuc_cart[3 * a + 2]};
            double dx[3];
            for (int d = 0; d < 3; d++) dx[d] = c2x[d] * dc[0] + c2x[d + 3] * dc[1] + c2x[d + 6] * dc[2];
            if (cell.shortest(dx) <= 1e-5) isnuc = true;
          }
        }
        if (isnuc) yp[0] = yp[1] = yp[2] = 0.0;
        const double gfmod = std::sqrt(yp[0] * yp[0] + yp[1] * yp[1] + yp[2] * yp[2]);  // norm2
        double ev[3];
        eig3_sym(ypp, ev);
        const double dimgrad = gfmod / (cst * std::pow(std::max(y, VSMALL), fthirds));  // (:571)
        const size_t o = (size_t)k + (size_t)nstep[2] * ((size_t)j + (size_t)nstep[1] * i);
        cgrad[o] = dimgrad;
        crho[o] = std::copysign(std::fabs(y), ev[1]) * 100.0;  // sign(res%f, ehess(2))*100 (:599)
        if (lam2) lam2[o] = ev[1];
      }
}

// ---------------------------------------------------------------------------
// Synthetic promolecular-like density used by tests and the bench:
//   rho(x) = sum_atoms sum_images Z * exp(-alpha*r) * cut(r),  cut = (1-(r/rc)^2)^3 (rc>0) or 1.
// Images: lattice translations in [-nimg, nimg]^3.  Stands in for
// promolecular_array3 (crystalmod@complex.f90:436-470), whose atomic tables
// (dat/wfc) are not part of the hot path.
// ---------------------------------------------------------------------------
void orc_promolecular(const int* n, const double* x2c, int nat, const double* xat, const double* zat,
                      const double* alpha, int nimg, double rc, double* f) {
  const int n1 = n[0], n2 = n[1], n3 = n[2];
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < n3; k++)
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n1; i++) {
        const double xf[3] = {(double)i / n1, (double)j / n2, (double)k / n3};
        double s = 0.0;
        for (int a = 0; a < nat; a++)
          for (int ta = -nimg; ta <= nimg; ta++)
            for (int tb = -nimg; tb <= nimg; tb++)
              for (int tc = -nimg; tc <= nimg; tc++) {
                const double d[3] = {xf[0] - xat[3 * a] + ta, xf[1] - xat[3 * a + 1] + tb, xf[2] - xat[3 * a + 2] + tc};
                double c[3];
                for (int q = 0; q < 3; q++) c[q] = x2c[q] * d[0] + x2c[q + 3] * d[1] + x2c[q + 6] * d[2];
                const double r = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
                if (rc > 0.0) {
                  if (r >= rc) continue;
                  const double u = 1.0 - (r / rc) * (r / rc);
                  s += zat[a] * std::exp(-alpha[a] * r) * u * u * u;
                } else {
                  s += zat[a] * std::exp(-alpha[a] * r);
                }
              }
        f[i + (size_t)n1 * (j + (size_t)n2 * k)] = s;
      }
}


// ---------------------------------------------------------------------------
// FFT-derived fields.  Restates grid3%fft (grid3mod@proc.f90:1757-1872) on top of a plain complex
// 3-D DFT with the conventions of cfftnd (cfftnd.f90:30-120, FFTPACK5 cmfm1f/cmfm1b): sgn = -1 is
// the forward transform exp(-2 pi i jk/n) SCALED by 1/ntot, sgn = +1 the unscaled backward one.
// The DFT itself is a mixed-radix decimation-in-time recursion (O(p^2) butterflies for a prime
// factor p); it is not FFTPACK's operation order, so results agree with the reference to
// rounding (a few ulp of max|f| log ntot), which is what the tests allow.
// ---------------------------------------------------------------------------
typedef std::complex<double> cplx;
// 1-D DFT of length n (input with stride), mixed-radix decimation in time; sgn = -1 forward, +1 backward
static void dft1_go(int n, int stride, const cplx* in, cplx* out, int sgn) {
  if (n == 1) { out[0] = in[0]; return; }
  int p = n;
  for (int q = 2; q * q <= n; q++)
    if (n % q == 0) { p = q; break; }
  const int m = n / p;
  const double pi = 3.14159265358979323846264338328;
  if (m == 1) {  // prime length: direct sum
    for (int k = 0; k < n; k++) {
      cplx s = 0.0;
      for (int j = 0; j < n; j++) {
        const double a = sgn * 2.0 * pi * (double)((long long)j * k % n) / n;
        s += in[(size_t)j * stride] * cplx(std::cos(a), std::sin(a));
      }
      out[k] = s;
    }
    return;
  }
  std::vector<cplx> tmp(n);  // p interleaved sub-sequences of length m
  for (int r = 0; r < p; r++) dft1_go(m, stride * p, in + (size_t)r * stride, tmp.data() + (size_t)r * m, sgn);
  for (int k = 0; k < m; k++)
    for (int q = 0; q < p; q++) {
      cplx s = 0.0;
      const int kk = k + q * m;
      for (int r = 0; r < p; r++) {
        const double a = sgn * 2.0 * pi * (double)((long long)r * kk % n) / n;
        s += tmp[(size_t)r * m + k] * cplx(std::cos(a), std::sin(a));
      }
      out[kk] = s;
    }
}
static void dft1(int n, cplx* line, int sgn, std::vector<cplx>& a, std::vector<cplx>& b) {
  a.assign(line, line + n);
  b.resize(n);
  dft1_go(n, 1, a.data(), b.data(), sgn);
  for (int i = 0; i < n; i++) line[i] = b[i];
}
// cfftnd(3,n,sgn,c): in place, Fortran order, forward scaled by 1/ntot (cfftnd.f90:33-36)
static void cfftnd3(const int* n, int sgn, std::vector<cplx>& c) {
  const int n1 = n[0], n2 = n[1], n3 = n[2];
  const size_t ntot = (size_t)n1 * n2 * n3;
#pragma omp parallel
  {
    std::vector<cplx> line, a, b;
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < n3; k++)
      for (int j = 0; j < n2; j++) {
        cplx* p = c.data() + (size_t)n1 * (j + (size_t)n2 * k);
        dft1(n1, p, sgn, a, b);
      }
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < n3; k++)
      for (int i = 0; i < n1; i++) {
        line.resize(n2);
        for (int j = 0; j < n2; j++) line[j] = c[i + (size_t)n1 * (j + (size_t)n2 * k)];
        dft1(n2, line.data(), sgn, a, b);
        for (int j = 0; j < n2; j++) c[i + (size_t)n1 * (j + (size_t)n2 * k)] = line[j];
      }
#pragma omp for collapse(2) schedule(static)
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n1; i++) {
        line.resize(n3);
        for (int k = 0; k < n3; k++) line[k] = c[i + (size_t)n1 * (j + (size_t)n2 * k)];
        dft1(n3, line.data(), sgn, a, b);
        for (int k = 0; k < n3; k++) c[i + (size_t)n1 * (j + (size_t)n2 * k)] = line[k];
      }
  }
  if (sgn < 0) {
    const double sc = 1.0 / (double)ntot;
    for (size_t i = 0; i < ntot; i++) c[i] *= sc;
  }
}

// iff: the reference's ifformat_as_ft_* codes (param.F90:225-236): 33 x, 34 y, 35 z, 36 xx, 37 xy,
// 38 xz, 39 yy, 40 yz, 41 zz, 42 grad (|grad f|), 43 lap, 44 pot.   grid3mod@proc.f90:1757-1872
int orc_fft_derivative(const double* f, const int* n, const double* x2c, int iff, double* out) {
  if (iff < 33 || iff > 44) return 1;
  const int n1 = n[0], n2 = n[1], n3 = n[2];
  const size_t ntot = (size_t)n1 * n2 * n3;
  const double pi = 3.14159265358979323846264338328;
  // reciprocal lattice vectors (:1785-1789): bvec(:,1) = cross(x2c(:,3),x2c(:,2)) ...
  auto cross = [](const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
  };
  double bvec[9];
  cross(x2c + 6, x2c + 3, bvec + 0);
  cross(x2c + 0, x2c + 6, bvec + 3);
  cross(x2c + 3, x2c + 0, bvec + 6);
  const double det = x2c[0] * (x2c[4] * x2c[8] - x2c[5] * x2c[7]) - x2c[3] * (x2c[1] * x2c[8] - x2c[2] * x2c[7]) +
                     x2c[6] * (x2c[1] * x2c[5] - x2c[2] * x2c[4]);
  for (int q = 0; q < 9; q++) bvec[q] = 2.0 * pi / std::fabs(det) * bvec[q];
  // vgc(:,igfft) (:1791-1800)
  std::vector<double> vgc(3 * ntot);
  for (int i1 = n1 / 2 - n1 + 1; i1 <= n1 / 2; i1++)
    for (int i2 = n2 / 2 - n2 + 1; i2 <= n2 / 2; i2++)
      for (int i3 = n3 / 2 - n3 + 1; i3 <= n3 / 2; i3++) {
        const size_t ig = (size_t)(((i3 % n3) + n3) % n3) * n2 * n1 + (size_t)(((i2 % n2) + n2) % n2) * n1 + (((i1 % n1) + n1) % n1);
        for (int d = 0; d < 3; d++) vgc[3 * ig + d] = ((double)i1 * bvec[d] + (double)i2 * bvec[3 + d]) + (double)i3 * bvec[6 + d];
      }
  std::vector<cplx> z(ntot);
  const cplx img(0.0, 1.0);
  if (iff == 42) {  // gradient modulus (:1809-1827)
    for (size_t i = 0; i < ntot; i++) out[i] = 0.0;
    for (int d = 0; d < 3; d++) {
      for (size_t i = 0; i < ntot; i++) z[i] = f[i];
      cfftnd3(n, -1, z);
      for (size_t i = 0; i < ntot; i++) z[i] = vgc[3 * i + d] * cplx(-z[i].imag(), z[i].real());
      cfftnd3(n, +1, z);
      for (size_t i = 0; i < ntot; i++) out[i] = out[i] + z[i].real() * z[i].real();
    }
    for (size_t i = 0; i < ntot; i++) out[i] = std::sqrt(out[i]);
    return 0;
  }
  for (size_t i = 0; i < ntot; i++) z[i] = f[i];
  cfftnd3(n, -1, z);
  for (size_t i = 0; i < ntot; i++) {
    const double* v = &vgc[3 * i];
    switch (iff) {
      case 33: z[i] = -v[0] * img * z[i]; break;
      case 34: z[i] = -v[1] * img * z[i]; break;
      case 35: z[i] = -v[2] * img * z[i]; break;
      case 36: z[i] = -v[0] * v[0] * z[i]; break;
      case 37: z[i] = -v[0] * v[1] * z[i]; break;
      case 38: z[i] = -v[0] * v[2] * z[i]; break;
      case 39: z[i] = -v[1] * v[1] * z[i]; break;
      case 40: z[i] = -v[1] * v[2] * z[i]; break;
      case 41: z[i] = -v[2] * v[2] * z[i]; break;
      case 43: z[i] = -((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) * z[i]; break;
      default: {
        const double vgc2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
        if (vgc2 < 1e-12) z[i] = 0.0;
        else z[i] = -z[i] / vgc2;
      }
    }
  }
  cfftnd3(n, +1, z);
  if (iff == 44)
    for (size_t i = 0; i < ntot; i++) out[i] = -4.0 * pi * z[i].real();
  else
    for (size_t i = 0; i < ntot; i++) out[i] = z[i].real();
  return 0;
}

// grinterp_trilinear value (grid3mod@proc.f90:2323-2370); x0 in [0,1) crystallographic
static double trilinear_value(const double* f, const int* n, const double x0[3]) {
  int idx[3];
  double r[3], s[3];
  for (int d = 0; d < 3; d++) {
    int fl = (int)std::floor(x0[d] * n[d]);     // grid_floor: floor(x*n), modulo n, +1 (:3133-3135)
    fl = ((fl % n[d]) + n[d]) % n[d];
    idx[d] = fl + 1;
    r[d] = n[d] * x0[d] - idx[d] + 1;           // (:2346)
    s[d] = 1.0 - r[d];
  }
  double ff[3][3][3];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        const int a = (idx[0] + i - 1) % n[0], b = (idx[1] + j - 1) % n[1], c = (idx[2] + k - 1) % n[2];
        ff[i][j][k] = f[a + (size_t)n[0] * (b + (size_t)n[1] * c)];
      }
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      ff[i][j][2] = ff[i][j][0] * s[2] + ff[i][j][1] * r[2];
      ff[i][2][j] = ff[i][0][j] * s[1] + ff[i][1][j] * r[1];
      ff[2][i][j] = ff[0][i][j] * s[0] + ff[1][i][j] * r[0];
    }
  for (int i = 0; i < 2; i++) {
    ff[i][2][2] = ff[i][0][2] * s[1] + ff[i][1][2] * r[1];
    ff[2][i][2] = ff[2][i][0] * s[2] + ff[2][i][1] * r[2];
    ff[2][2][i] = ff[0][2][i] * s[0] + ff[1][2][i] * r[0];
  }
  return ff[0][2][2] * s[0] + ff[1][2][2] * r[0];
}

// field%grd with nder = 0 on a grid field (fieldmod@proc.f90:948-961): the raw node value when the
// point is within neargrideps = 1e-12 (in grid units) of a node, else grid3%interp in `mode`
// (0 = tricubic, 1 = trilinear).  wx already wrapped as in :921-929.
static double grd0_value(const double* f, const int* n, const double* c2xl, const double wx[3], int mode) {
  double x[3];
  int idx[3];
  bool isgrid = true;
  for (int d = 0; d < 3; d++) {
    double m = wx[d] - std::floor(wx[d]);  // modulo(wx,1d0)
    if (m >= 1.0) m = 0.0;
    x[d] = m * n[d];
    idx[d] = (int)std::lround(x[d]);       // nint
    if (!(std::fabs(x[d] - idx[d]) < 1e-12)) isgrid = false;
  }
  if (isgrid) {
    for (int d = 0; d < 3; d++) idx[d] = ((idx[d] % n[d]) + n[d]) % n[d];
    return f[idx[0] + (size_t)n[0] * (idx[1] + (size_t)n[1] * idx[2])];
  }
  if (mode == 1) {
    double x0[3];
    for (int d = 0; d < 3; d++) {
      x0[d] = wx[d] - std::floor(wx[d]);
      if (x0[d] >= 1.0) x0[d] = 0.0;
    }
    return trilinear_value(f, n, x0);
  }
  double y, yp[3], ypp[9];
  orc_grid_interp_tricubic(f, n, c2xl, wx, &y, yp, ypp);
  return y;
}

void orc_grid_interp_trilinear(const double* f, const int* n, const double* xi_in, double* y) {
  double x0[3];
  for (int d = 0; d < 3; d++) {
    x0[d] = xi_in[d] - std::floor(xi_in[d]);
    if (x0[d] >= 1.0) x0[d] = 0.0;
  }
  *y = trilinear_value(f, n, x0);
}

// ---------------------------------------------------------------------------
// orc_nci_rdg_fourier: the NCIPLOT hot loop with fourierint = .true. (nci@proc.f90:527-565):
// rho from the reference field (tricubic, nder = 0), |grad rho| and Hxx, Hyy, Hzz from the FFT-derived
// grids (fgrad, fxx, fyy, fzz: orc_fft_derivative 42, 36, 39, 41) read with TRILINEAR interpolation
// (:534-537); sign = + if at least two of Hxx, Hyy, Hzz are positive (:558-562).
// ---------------------------------------------------------------------------
void orc_nci_rdg_fourier(const double* f, const double* fgrad, const double* fxx, const double* fyy, const double* fzz,
                         const int* n, const double* x0, const double* xmat, const int* nstep, const double* c2x,
                         const double* c2xl, double* crho, double* cgrad) {
  const double pi = 3.14159265358979323846264338328;
  const double cst = 2.0 * std::pow(3.0 * pi * pi, 1.0 / 3.0);
  const double fthirds = 4.0 / 3.0;
  const double flooreps = 1e-4;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nstep[0]; i++)
    for (int j = 0; j < nstep[1]; j++)
      for (int k = 0; k < nstep[2]; k++) {
        double x[3], wx[3];
        for (int d = 0; d < 3; d++) x[d] = ((x0[d] + i * xmat[d + 0]) + j * xmat[d + 3]) + k * xmat[d + 6];
        for (int d = 0; d < 3; d++) {
          double s = c2x[d + 0] * x[0];
          s = s + c2x[d + 3] * x[1];
          s = s + c2x[d + 6] * x[2];
          wx[d] = s;
        }
        for (int d = 0; d < 3; d++)
          if (wx[d] < -flooreps || wx[d] > 1.0 + flooreps) wx[d] = wx[d] - std::floor(wx[d]);
        const double rho = grd0_value(f, n, c2xl, wx, 0);
        const double g = grd0_value(fgrad, n, c2xl, wx, 1);
        const double dimgrad = g / (cst * std::pow(std::max(rho, VSMALL), fthirds));  // (:552)
        const double eh[3] = {grd0_value(fxx, n, c2xl, wx, 1), grd0_value(fyy, n, c2xl, wx, 1), grd0_value(fzz, n, c2xl, wx, 1)};
        const int npos = (eh[0] > 0.0) + (eh[1] > 0.0) + (eh[2] > 0.0);
        const double e2 = npos >= 2 ? 1.0 : -1.0;
        const size_t o = (size_t)k + (size_t)nstep[2] * ((size_t)j + (size_t)nstep[1] * i);
        cgrad[o] = dimgrad;
        crho[o] = std::copysign(std::fabs(rho), e2) * 100.0;
      }
}

// ---------------------------------------------------------------------------
// Multipoles of the basins: the `else` branch of intgrid_fields,
// src/integration@proc.f90:1302-1361 (INTEGRABLE id MULTIPOLES [lmax]).
// ---------------------------------------------------------------------------
}  // extern "C"
namespace {

// crystal%shortest, src/crystalmod@proc.f90:1056-1085.  x (cryst.) -> shortest lattice-translated copy (Cartesian).
// matmul(A,x)_i = A(i,1)x1 + A(i,2)x2 + A(i,3)x3, accumulated left to right; norm2 as sqrt of the sum of squares.
struct OrcCell {
  int isortho, isortho_del, nws;
  const double *x2c, *x2xr, *xr2c, *ws;  // 3x3 column-major; ws(3,nws) Cartesian (ws_ineighc)
};
inline void matvec3(const double* m, const double* x, double* y) {
  for (int i = 0; i < 3; i++) y[i] = m[i] * x[0] + m[i + 3] * x[1] + m[i + 6] * x[2];
}
inline void orc_shortest(const OrcCell& c, double x[3]) {
  double t[3];
  if (c.isortho) {
    for (int i = 0; i < 3; i++) t[i] = x[i] - (double)std::lround(x[i]);
    matvec3(c.x2c, t, x);
    return;
  }
  matvec3(c.x2xr, x, t);
  for (int i = 0; i < 3; i++) t[i] = t[i] - (double)std::lround(t[i]);
  matvec3(c.xr2c, t, x);
  double dist = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  if (!c.isortho_del) {
    const double x0[3] = {x[0], x[1], x[2]};
    for (int i = 0; i < c.nws; i++) {
      const double xt[3] = {x0[0] + c.ws[3 * i], x0[1] + c.ws[3 * i + 1], x0[2] + c.ws[3 * i + 2]};
      const double d = std::sqrt(xt[0] * xt[0] + xt[1] * xt[1] + xt[2] * xt[2]);
      if (d < dist) {
        x[0] = xt[0]; x[1] = xt[1]; x[2] = xt[2];
        dist = d;
      }
    }
  }
}

// tosphere, src/tools_math@proc.f90:381-406
inline void orc_tosphere(const double v[3], double& r, double tp[2]) {
  const double eps = 1e-14, pi = 3.14159265358979323846;
  r = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  tp[0] = tp[1] = 0.0;
  if (r > eps) {
    const double t1 = v[2] / r;
    if (t1 >= 1.0) tp[0] = 0.0;
    else if (t1 <= -1.0) tp[0] = pi;
    else tp[0] = std::acos(t1);
    if (std::fabs(v[0]) > eps || std::fabs(v[1]) > eps) tp[1] = std::atan2(v[1], v[0]);
  }
}

// genylm, src/tools_math@proc.f90:314-377 (Masters & Richards-Dinger recursion), ylm(l*(l+1)+m), 0-based
inline void orc_genylm(int lmax, const double tp[2], std::complex<double>* ylm) {
  const double fourpi = 12.566370614359172954;
  ylm[0] = 0.28209479177387814347;
  if (lmax == 0) return;
  std::vector<double> x(lmax + 1);
  std::vector<std::complex<double>> z(lmax + 1);
  const double sn = std::sin(tp[0]), cs = std::cos(tp[0]);
  for (int m = 1; m <= lmax; m++) {
    const double t1 = (double)m * tp[1];
    z[m] = std::complex<double>(std::cos(t1), std::sin(t1));
  }
  for (int l = 1; l <= lmax; l++) {
    x[l] = (l % 2 == 0) ? 1.0 : -1.0;
    double dx = 0.0;
    for (int m = l; m >= 1; m--) {
      const double t1 = std::sqrt((double)((l + m) * (l - m + 1)));
      x[m - 1] = -(sn * dx + (double)(2 * m) * cs * x[m]) / t1;
      dx = sn * x[m] * t1;
    }
    double t1 = sn, sum = 0.0;
    for (int m = 1; m <= l; m++) {
      x[m] = t1 * x[m];
      sum = sum + x[m] * x[m];
      t1 = t1 * sn;
    }
    sum = 2.0 * sum + x[0] * x[0];
    t1 = std::sqrt((double)(2 * l + 1) / (fourpi * sum));
    const int lm0 = l * (l + 1);
    ylm[lm0] = t1 * x[0];
    for (int m = 1; m <= l; m++) {
      const double a = t1 * x[m];  // t1*x(m)*z(m), left to right
      ylm[lm0 + m] = std::complex<double>(a * z[m].real(), a * z[m].imag());
      ylm[lm0 - m] = std::conj(ylm[lm0 + m]);
      if (m % 2 != 0) ylm[lm0 - m] = -ylm[lm0 - m];
    }
  }
}

// r**l with an integer exponent as gfortran evaluates it (libgcc __powidf2: square and multiply)
inline double orc_powi(double x, int m) {
  unsigned n = (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y *= x;
  }
  return y;
}

// genrlm_real, src/tools_math@proc.f90:273-306
inline void orc_genrlm_real(int lmax, double r, const double tp[2], double* rrlm, std::complex<double>* rlm) {
  const double pi = 3.14159265358979323846, sh = 1.0 / std::sqrt(2.0);
  const std::complex<double> img(0.0, 1.0);
  orc_genylm(lmax, tp, rlm);
  for (int l = 0; l <= lmax; l++) {
    const int imin = l * l, imax = (l + 1) * (l + 1);
    const double s = std::sqrt(4.0 * pi / (double)(2 * l + 1)), rl = orc_powi(r, l);
    for (int i = imin; i < imax; i++) rlm[i] = std::complex<double>(rlm[i].real() * s * rl, rlm[i].imag() * s * rl);
    int ip = imin + l;
    rrlm[ip] = rlm[ip].real();
    for (int m = 1; m <= l; m++) {
      ip = imin + l + m;
      const int im = imin + l - m;
      const double iphas = (m % 2 == 0) ? 1.0 : -1.0;
      const std::complex<double> a = iphas * rlm[ip] + rlm[im];
      const std::complex<double> b = -iphas * img * rlm[ip] + img * rlm[im];
      rrlm[im] = sh * a.real();
      rrlm[ip] = sh * b.real();
    }
  }
}

}  // namespace
extern "C" {

// Bader / isosurface branch (integration@proc.f90:1338-1358): mpole(:,ix) += rrlm(p/n - xattr(:,ix)) * fint(p),
// loop order i1, i2, i3 (i3 innermost), then mpole * omega / ntot (:1360).  idg == 0 (points of a discarded
// attractor) are skipped; the reference would read xattr(:,0) there.  mpole((lmax+1)^2, nattr) column-major.
void orc_multipoles_bader(const int* idg, const int* n, int nattr, const double* xattr, int lmax, const double* fint,
                          int isortho, int isortho_del, const double* x2c, const double* x2xr, const double* xr2c, int nws,
                          const double* ws, double omega, double* mpole) {
  const OrcCell cell{isortho, isortho_del, nws, x2c, x2xr, xr2c, ws};
  const int nlm = (lmax + 1) * (lmax + 1);
  std::vector<double> rrlm(nlm);
  std::vector<std::complex<double>> rlm(nlm);
  for (size_t e = 0; e < (size_t)nlm * nattr; e++) mpole[e] = 0.0;
  for (int i1 = 0; i1 < n[0]; i1++)
    for (int i2 = 0; i2 < n[1]; i2++)
      for (int i3 = 0; i3 < n[2]; i3++) {
        const size_t q = (size_t)i1 + (size_t)n[0] * ((size_t)i2 + (size_t)n[1] * i3);
        const int ix = idg[q];
        if (ix < 1 || ix > nattr) continue;
        double dv[3] = {(double)i1 / (double)n[0] - xattr[3 * (ix - 1)], (double)i2 / (double)n[1] - xattr[3 * (ix - 1) + 1],
                        (double)i3 / (double)n[2] - xattr[3 * (ix - 1) + 2]};
        orc_shortest(cell, dv);
        double r, tp[2];
        orc_tosphere(dv, r, tp);
        orc_genrlm_real(lmax, r, tp, rrlm.data(), rlm.data());
        double* out = mpole + (size_t)nlm * (ix - 1);
        for (int e = 0; e < nlm; e++) out[e] = out[e] + rrlm[e] * fint[q];
      }
  const double ntot = (double)n[0] * (double)n[1] * (double)n[2];
  for (size_t e = 0; e < (size_t)nlm * nattr; e++) mpole[e] = mpole[e] * omega / ntot;
}

// YT branch (integration@proc.f90:1316-1336) for ONE basin: w = yt_weights(idb); points with |w| < 1e-15 are
// skipped; mpole_m += rrlm(p/n - xattr_m) * fint * w; scaled by omega/ntot like :1360.  mpole_m((lmax+1)^2).
void orc_multipoles_weighted(const double* w, const int* n, const double* xattr_m, int lmax, const double* fint, int isortho,
                             int isortho_del, const double* x2c, const double* x2xr, const double* xr2c, int nws,
                             const double* ws, double omega, double* mpole_m) {
  const OrcCell cell{isortho, isortho_del, nws, x2c, x2xr, xr2c, ws};
  const int nlm = (lmax + 1) * (lmax + 1);
  std::vector<double> rrlm(nlm);
  std::vector<std::complex<double>> rlm(nlm);
  for (int e = 0; e < nlm; e++) mpole_m[e] = 0.0;
  for (int i1 = 0; i1 < n[0]; i1++)
    for (int i2 = 0; i2 < n[1]; i2++)
      for (int i3 = 0; i3 < n[2]; i3++) {
        const size_t q = (size_t)i1 + (size_t)n[0] * ((size_t)i2 + (size_t)n[1] * i3);
        if (std::fabs(w[q]) < 1e-15) continue;
        double dv[3] = {(double)i1 / (double)n[0] - xattr_m[0], (double)i2 / (double)n[1] - xattr_m[1],
                        (double)i3 / (double)n[2] - xattr_m[2]};
        orc_shortest(cell, dv);
        double r, tp[2];
        orc_tosphere(dv, r, tp);
        orc_genrlm_real(lmax, r, tp, rrlm.data(), rlm.data());
        for (int e = 0; e < nlm; e++) mpole_m[e] = mpole_m[e] + rrlm[e] * fint[q] * w[q];
      }
  const double ntot = (double)n[0] * (double)n[1] * (double)n[2];
  for (int e = 0; e < nlm; e++) mpole_m[e] = mpole_m[e] * omega / ntot;
}

// bader_remap, src/bader@proc.f90:237-296 (DELOC): attractor images.  A point whose shortest vector to its attractor
// differs from the in-cell difference by a lattice vector p /= 0 belongs to the image (attractor, p); images are
// numbered nattr+1, nattr+2, ... in the order in which the scan (m1 fastest, m3 slowest) first meets them.
// Out: idg1(n1,n2,n3), iatt(nattn), ilvec(3,nattn), *nattn.  Returns -2 when maxattn is too small.  idg == 0 is
// skipped (the reference would read xattr(:,0)).
int orc_bader_remap(const int* idg, const int* n, int nattr, const double* xattr, const double* c2x, int isortho,
                    int isortho_del, const double* x2c, const double* x2xr, const double* xr2c, int nws, const double* ws,
                    int maxattn, int* nattn_out, int* iatt, int* ilvec, int* idg1) {
  const OrcCell cell{isortho, isortho_del, nws, x2c, x2xr, xr2c, ws};
  if (nattr > maxattn) return -2;
  int nattn = nattr;
  for (int i = 0; i < nattr; i++) {
    iatt[i] = i + 1;
    ilvec[3 * i] = ilvec[3 * i + 1] = ilvec[3 * i + 2] = 0;
  }
  for (int m3 = 0; m3 < n[2]; m3++)
    for (int m2 = 0; m2 < n[1]; m2++)
      for (int m1 = 0; m1 < n[0]; m1++) {
        const size_t q = (size_t)m1 + (size_t)n[0] * ((size_t)m2 + (size_t)n[1] * m3);
        const int b = idg[q];
        idg1[q] = b;
        if (b < 1 || b > nattr) continue;
        const double x[3] = {(double)m1 / (double)n[0] - xattr[3 * (b - 1)], (double)m2 / (double)n[1] - xattr[3 * (b - 1) + 1],
                             (double)m3 / (double)n[2] - xattr[3 * (b - 1) + 2]};
        double xs[3] = {x[0], x[1], x[2]}, xc[3];
        orc_shortest(cell, xs);
        matvec3(c2x, xs, xc);
        const int p[3] = {(int)std::lround(x[0] - xc[0]), (int)std::lround(x[1] - xc[1]), (int)std::lround(x[2] - xc[2])};
        if (p[0] != 0 || p[1] != 0 || p[2] != 0) {
          bool found = false;
          for (int i = nattr; i < nattn; i++)
            if (iatt[i] == b && ilvec[3 * i] == p[0] && ilvec[3 * i + 1] == p[1] && ilvec[3 * i + 2] == p[2]) {
              found = true;
              idg1[q] = i + 1;
              break;
            }
          if (!found) {
            if (nattn + 1 > maxattn) return -2;
            ilvec[3 * nattn] = p[0]; ilvec[3 * nattn + 1] = p[1]; ilvec[3 * nattn + 2] = p[2];
            iatt[nattn] = b;
            nattn++;
            idg1[q] = nattn;
          }
        }
      }
  *nattn_out = nattn;
  return 0;
}

// One basin of yt_remap, src/yt@proc.f90:557-589: w = yt_weights(idb = ib); the images of basin ib met by the points
// with |w| >= 1e-15 are appended to (iatt, ilvec) in scan order.  *nattn is updated in place (start it at nattr with
// iatt(i) = i, ilvec = 0 like :546-554).
int orc_yt_remap_basin(const double* w, const int* n, int nattr, int ib, const double* xattr_i, const double* c2x, int isortho,
                       int isortho_del, const double* x2c, const double* x2xr, const double* xr2c, int nws, const double* ws,
                       int maxattn, int* nattn_io, int* iatt, int* ilvec) {
  const OrcCell cell{isortho, isortho_del, nws, x2c, x2xr, xr2c, ws};
  int nattn = *nattn_io;
  for (int m3 = 0; m3 < n[2]; m3++)
    for (int m2 = 0; m2 < n[1]; m2++)
      for (int m1 = 0; m1 < n[0]; m1++) {
        const size_t q = (size_t)m1 + (size_t)n[0] * ((size_t)m2 + (size_t)n[1] * m3);
        if (std::fabs(w[q]) < 1e-15) continue;
        const double x[3] = {(double)m1 / (double)n[0] - xattr_i[0], (double)m2 / (double)n[1] - xattr_i[1],
                             (double)m3 / (double)n[2] - xattr_i[2]};
        double xs[3] = {x[0], x[1], x[2]}, xc[3];
        orc_shortest(cell, xs);
        matvec3(c2x, xs, xc);
        const int p[3] = {(int)std::lround(x[0] - xc[0]), (int)std::lround(x[1] - xc[1]), (int)std::lround(x[2] - xc[2])};
        if (p[0] != 0 || p[1] != 0 || p[2] != 0) {
          bool found = false;
          for (int j = nattr; j < nattn; j++)
            if (iatt[j] == ib && ilvec[3 * j] == p[0] && ilvec[3 * j + 1] == p[1] && ilvec[3 * j + 2] == p[2]) {
              found = true;
              break;
            }
          if (!found) {
            if (nattn + 1 > maxattn) return -2;
            ilvec[3 * nattn] = p[0]; ilvec[3 * nattn + 1] = p[1]; ilvec[3 * nattn + 2] = p[2];
            iatt[nattn] = ib;
            nattn++;
          }
        }
      }
  *nattn_io = nattn;
  return 0;
}

// ---------------------------------------------------------------------------
// HIRSHFELD on a grid: promolecular_array3 / promolecular_atom (src/crystalmod@complex.f90:436-470,
// src/crystalmod@env.f90:622-748), grid1%interp (src/grid1mod@proc.f90:86-137) and the loop of
// intgrid_hirshfeld_fields (src/integration@proc.f90:1552-1596).
// The reference finds the atoms near a point with its block environment (list_near_atoms); the sums below run over the
// same set -- every periodic image within the species cutoff -- enumerated by brute force, so values agree to
// rounding (the order of the additions differs), not bit for bit.
// Species tables: spc_ngrid, spc_off (offset into rtab/ftab), spc_rmax (g%rmax), spc_rcut = min(cutrad(z), g%rmax),
// spc_a, spc_b (r(i) = a exp(b (i-1))); ngrid = 0 marks a species without a usable grid (skipped).
// ---------------------------------------------------------------------------
}  // extern "C"
namespace {

struct OrcSpecies {
  int nspc;
  const int *ngrid, *off;
  const double *a, *b, *rmax, *rcut, *rtab, *ftab;
};

// grid1%interp, value only (grid1mod@proc.f90:86-137): 4-node Lagrange on the logarithmic grid
inline double orc_grid1_interp(const OrcSpecies& S, int is, double r0) {
  const int ng = S.ngrid[is];
  if (ng <= 0) return 0.0;
  if (r0 >= S.rmax[is]) return 0.0;
  const double* rg = S.rtab + S.off[is];
  const double* fg = S.ftab + S.off[is];
  int ir;
  double r;
  if (r0 <= rg[0]) { ir = 1; r = rg[0]; }
  else { ir = 1 + (int)std::floor(std::log(r0 / S.a[is]) / S.b[is]); r = r0; }
  const int i0 = std::min(std::max(ir, 2), ng - 2) - 2;  // nodes i0+1 .. i0+4 (1-based) = rg[i0 .. i0+3]
  double rr[4], dr1[4], x1[4][4];
  for (int i = 0; i < 4; i++) {
    rr[i] = rg[i0 + i];
    dr1[i] = r - rr[i];
    for (int j = 0; j < i; j++) {
      x1[i][j] = 1.0 / (rr[i] - rr[j]);
      x1[j][i] = -x1[i][j];
    }
  }
  double f = 0.0;
  for (int i = 0; i < 4; i++) {
    double prod = 1.0;
    for (int j = 0; j < 4; j++) {
      if (i == j) continue;
      prod = prod * dr1[j] * x1[i][j];
    }
    f = f + fg[i0 + i] * prod;
  }
  return f;
}

struct OrcImage { int atom; double x[3]; };

// every periodic image of every atom that can lie within the largest cutoff of some point of the cell
inline std::vector<OrcImage> orc_images(const double* x2c, int nat, const double* xat, double rcutmax) {
  // plane spacings h_i = 1 / |row i of c2x|
  double c2x[9];
  {
    auto A = [&](int i, int j) { return x2c[i + 3 * j]; };
    const double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                       A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    const double d = 1.0 / det;
    c2x[0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d; c2x[3] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d; c2x[6] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
    c2x[1] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d; c2x[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d; c2x[7] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
    c2x[2] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d; c2x[5] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d; c2x[8] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
  }
  int m[3];
  for (int i = 0; i < 3; i++) {
    const double g = std::sqrt(c2x[i] * c2x[i] + c2x[i + 3] * c2x[i + 3] + c2x[i + 6] * c2x[i + 6]);  // 1/h_i
    m[i] = (int)std::ceil(rcutmax * g) + 1;
  }
  std::vector<OrcImage> im;
  for (int a = 0; a < nat; a++)
    for (int l1 = -m[0]; l1 <= m[0]; l1++)
      for (int l2 = -m[1]; l2 <= m[1]; l2++)
        for (int l3 = -m[2]; l3 <= m[2]; l3++) {
          const double xf[3] = {xat[3 * a] - std::floor(xat[3 * a]) + l1, xat[3 * a + 1] - std::floor(xat[3 * a + 1]) + l2,
                                xat[3 * a + 2] - std::floor(xat[3 * a + 2]) + l3};
          OrcImage q;
          q.atom = a;
          matvec3(x2c, xf, q.x);
          im.push_back(q);
        }
  return im;
}

}  // namespace
extern "C" {

double orc_grid1_interp_value(int ngrid, double a, double b, double rmax, const double* rtab, const double* ftab, double r0) {
  const int off = 0;
  const double rcut = rmax;
  const OrcSpecies S{1, &ngrid, &off, &a, &b, &rmax, &rcut, rtab, ftab};
  return orc_grid1_interp(S, 0, r0);
}

// promolecular_array3 (crystalmod@complex.f90:436-470) with the optional fragment (infrag(nat), may be null = all atoms)
void orc_promolecular_grid(const int* n, const double* x2c, int nat, const double* xat, const int* ispc, int nspc,
                           const int* spc_ngrid, const int* spc_off, const double* spc_a, const double* spc_b,
                           const double* spc_rmax, const double* spc_rcut, const double* rtab, const double* ftab,
                           const unsigned char* infrag, double* out) {
  const OrcSpecies S{nspc, spc_ngrid, spc_off, spc_a, spc_b, spc_rmax, spc_rcut, rtab, ftab};
  double rcutmax = 0.0;
  for (int i = 0; i < nspc; i++) rcutmax = std::max(rcutmax, spc_rcut[i]);
  const std::vector<OrcImage> im = orc_images(x2c, nat, xat, rcutmax);
#pragma omp parallel for schedule(dynamic)
  for (int k = 0; k < n[2]; k++)
    for (int j = 0; j < n[1]; j++)
      for (int i = 0; i < n[0]; i++) {
        const double xf[3] = {(double)i / (double)n[0], (double)j / (double)n[1], (double)k / (double)n[2]};
        double xc[3];
        matvec3(x2c, xf, xc);
        double f = 0.0;
        for (const OrcImage& q : im) {
          if (infrag && !infrag[q.atom]) continue;
          const int is = ispc[q.atom] - 1;
          if (spc_ngrid[is] <= 0) continue;
          const double d0 = xc[0] - q.x[0], d1 = xc[1] - q.x[1], d2 = xc[2] - q.x[2];
          double r = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
          if (r > spc_rcut[is]) continue;                                   // list_near_atoms(up2dsp), env.f90:692, :713
          r = std::max(std::max(r, rtab[spc_off[is]]), 1e-14);              // :724
          double rho = orc_grid1_interp(S, is, r);
          rho = std::max(rho, 0.0);                                         // :726
          f = f + rho;
        }
        out[(size_t)i + (size_t)n[0] * ((size_t)j + (size_t)n[1] * k)] = f;
      }
}

// intgrid_hirshfeld_fields, integration@proc.f90:1552-1596: psum(nat,nprop), vol(nat); domask(nat) = docelatom(icp(.))
void orc_hirshfeld_fields(const int* n, const double* x2c, int nat, const double* xat, const int* ispc, int nspc,
                          const int* spc_ngrid, const int* spc_off, const double* spc_a, const double* spc_b,
                          const double* spc_rmax, const double* spc_rcut, const double* rtab, const double* ftab,
                          const double* promol, const unsigned char* domask, int nprop, const double* const* fields, double omega,
                          double* psum, double* vol) {
  const OrcSpecies S{nspc, spc_ngrid, spc_off, spc_a, spc_b, spc_rmax, spc_rcut, rtab, ftab};
  double rcutmax = 0.0;
  for (int i = 0; i < nspc; i++) rcutmax = std::max(rcutmax, spc_rcut[i]);
  const std::vector<OrcImage> im = orc_images(x2c, nat, xat, rcutmax);
  std::vector<double> acc((size_t)nat * (nprop + 1), 0.0);
  for (int k = 0; k < n[2]; k++)
    for (int j = 0; j < n[1]; j++)
      for (int i = 0; i < n[0]; i++) {
        const size_t q0 = (size_t)i + (size_t)n[0] * ((size_t)j + (size_t)n[1] * k);
        const double xf[3] = {(double)i / (double)n[0], (double)j / (double)n[1], (double)k / (double)n[2]};
        double xc[3];
        matvec3(x2c, xf, xc);
        const double fac = 1.0 / std::max(promol[q0], VSMALL);
        for (const OrcImage& q : im) {
          if (domask && !domask[q.atom]) continue;
          const int is = ispc[q.atom] - 1;
          if (spc_ngrid[is] <= 0) continue;
          const double d0 = xc[0] - q.x[0], d1 = xc[1] - q.x[1], d2 = xc[2] - q.x[2];
          const double r = std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
          if (r > spc_rcut[is]) continue;
          const double tosum = fac * orc_grid1_interp(S, is, r);             // :1566-1567
          acc[(size_t)q.atom * (nprop + 1) + nprop] += tosum;               // volume (:1574)
          for (int l = 0; l < nprop; l++) acc[(size_t)q.atom * (nprop + 1) + l] += tosum * fields[l][q0];  // :1576
        }
      }
  const double ntot = (double)n[0] * (double)n[1] * (double)n[2];
  for (int a = 0; a < nat; a++) {
    vol[a] = acc[(size_t)a * (nprop + 1) + nprop] * omega / ntot;           // :1590
    for (int l = 0; l < nprop; l++) psum[a + (size_t)nat * l] = acc[(size_t)a * (nprop + 1) + l] * omega / ntot;
  }
}

// one evaluation of genrlm_real(lmax, tosphere(v)) for the known-answer tests
void orc_rlm_real(const double* v, int lmax, double* rrlm) {
  double r, tp[2];
  orc_tosphere(v, r, tp);
  std::vector<std::complex<double>> rlm((lmax + 1) * (lmax + 1));
  orc_genrlm_real(lmax, r, tp, rrlm, rlm.data());
}

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
