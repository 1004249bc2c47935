// critic2_host.cpp -- see critic2_host.hpp.  Calls the C ABI only; no field arithmetic happens here.
#include "critic2_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace c2h {

namespace {
c2g_context* g_ctx = nullptr;
c2g_basins* g_basins = nullptr;  // result of the last BADER/YT call, consumed by intgrid_fields
int g_hgrid = -1;                // resident copy of bas%f

void check(int ier, const char* routine) {
  if (ier != 0) ferror(routine, std::string("GPU: ") + c2g_last_error(g_ctx));
}

// 3x3 inverse (adjugate); the reference uses matinv = LAPACK dgetrf/dgetri (tools_math@proc.f90:1203-1235)
void matinv3(const double a[9], double inv[9]) {
  auto A = [&](int i, int j) { return a[i + 3 * j]; };
  const double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                     A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
  if (det == 0.0) ferror("matinv", "singular matrix");
  const double d = 1.0 / det;
  inv[0 + 3 * 0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d;
  inv[0 + 3 * 1] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d;
  inv[0 + 3 * 2] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
  inv[1 + 3 * 0] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d;
  inv[1 + 3 * 1] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d;
  inv[1 + 3 * 2] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
  inv[2 + 3 * 0] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d;
  inv[2 + 3 * 1] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d;
  inv[2 + 3 * 2] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
}

// shortest Cartesian length of a crystallographic difference vector over lattice translations
double shortest(const double x2c[9], const double dxin[3]) {
  double dx[3];
  for (int i = 0; i < 3; i++) dx[i] = dxin[i] - std::round(dxin[i]);
  double best = 1e300;
  for (int a = -2; a <= 2; a++)
    for (int b = -2; b <= 2; b++)
      for (int c = -2; c <= 2; c++) {
        const double v[3] = {dx[0] + a, dx[1] + b, dx[2] + c};
        double r2 = 0.0;
        for (int i = 0; i < 3; i++) {
          const double ci = x2c[i] * v[0] + x2c[i + 3] * v[1] + x2c[i + 6] * v[2];
          r2 += ci * ci;
        }
        best = std::min(best, std::sqrt(r2));
      }
  return best;
}

// the per-maximum attractor identification of bader@proc.f90:160-199 / yt@proc.f90:129-168
void identify_attractors(system& s, basindat& bas, int nmax, const std::vector<int>& pmax, std::vector<int>& map) {
  // a maximum that is neither an atom nor a known attractor is kept only if the DISCARD expression vanishes there
  // (:184-190); a discarded one gets map 0 -- the library then treats every point below it as the reference does
  // (Bader: label 0; YT: ibasin = 0, which turns the points below into interatomic-surface points, yt@proc.f90:170)
  if (!bas.expr.empty())
    ferror("identify_attractors", "DISCARD expressions need critic2's arithmetic module (use the Fortran shim, which calls s%eval)");
  map.assign(nmax, 0);
  for (int i = 0; i < nmax; i++) {
    const double dv[3] = {(pmax[3 * i] - 1.0) / bas.n[0], (pmax[3 * i + 1] - 1.0) / bas.n[1], (pmax[3 * i + 2] - 1.0) / bas.n[2]};
    if (bas.atexist) {
      const int nid = s.identify_atom(dv, bas.ratom);
      if (nid > 0) map[i] = nid;
    }
    if (map[i] == 0 && bas.ratom > vsmall) {
      for (int l = 0; l < bas.nattr; l++)
        if (s.are_lclose(dv, &bas.xattr[3 * l], bas.ratom)) { map[i] = l + 1; break; }
    }
    if (map[i] == 0) {  // (a DISCARD expression would be evaluated here, exactly as in the CPU path)
      bas.nattr++;
      bas.xattr.insert(bas.xattr.end(), dv, dv + 3);
      map[i] = bas.nattr;
    }
  }
}

void upload_field(const basindat& bas, const char* routine) {
  if ((long long)bas.f.size() != (long long)bas.n[0] * bas.n[1] * bas.n[2]) ferror(routine, "inconsistent field size");
  if (g_hgrid >= 0) check(c2g_grid_free(g_ctx, g_hgrid), routine);
  g_hgrid = -1;
  check(c2g_grid_upload(g_ctx, bas.f.data(), bas.n, &g_hgrid), routine);
  if (g_basins) { c2g_basins_free(g_basins); g_basins = nullptr; }
}

void finish_assignment(system& s, basindat& bas, int nmax, const char* routine) {
  std::vector<int> pmax(3 * (size_t)nmax), map;
  check(c2g_basins_maxima(g_basins, pmax.data()), routine);
  if (bas.atexist) {  // the cell CP list of the reference field seeds the attractors (bader@proc.f90:113-118)
    const std::vector<double>& seeds = s.cpcel.empty() ? s.xat : s.cpcel;
    bas.nattr = (int)(seeds.size() / 3);
    bas.xattr = seeds;
  } else {
    bas.nattr = 0;
    bas.xattr.clear();
  }
  identify_attractors(s, bas, nmax, pmax, map);
  check(c2g_basins_set_map(g_basins, bas.nattr, map.data()), routine);
  bas.idg.assign(bas.f.size(), 0);
  check(c2g_basins_labels(g_basins, bas.idg.data()), routine);
}
}  // namespace

void ferror(const std::string& routine, const std::string& msg) {
  std::fprintf(stderr, "ERROR : %s: %s\n", routine.c_str(), msg.c_str());
  throw fatal_error(routine + ": " + msg);
}

void system::set_cell(const double x2c[9]) {
  std::memcpy(m_x2c, x2c, sizeof(m_x2c));
  omega = std::fabs(x2c[0] * (x2c[4] * x2c[8] - x2c[7] * x2c[5]) - x2c[3] * (x2c[1] * x2c[8] - x2c[7] * x2c[2]) +
                    x2c[6] * (x2c[1] * x2c[5] - x2c[4] * x2c[2]));
  // crystal%shortest data.  The reference Delaunay-reduces the cell (m_x2xr, m_xr2c) and keeps the Wigner-Seitz
  // neighbours of the reduced cell; this mirror keeps the input cell (m_x2xr = 1) and hands over every lattice
  // vector with coefficients in -2..2, a superset of the WS neighbours that gives the same minimum for the
  // moderately skewed cells of the tests.
  isortho = x2c[1] == 0.0 && x2c[2] == 0.0 && x2c[3] == 0.0 && x2c[5] == 0.0 && x2c[6] == 0.0 && x2c[7] == 0.0;
  isortho_del = false;
  const double one[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(m_x2xr, one, sizeof(one));
  std::memcpy(m_xr2c, x2c, sizeof(m_xr2c));
  ws_ineighc.clear();
  if (!isortho)
    for (int a = -2; a <= 2; a++)
      for (int b = -2; b <= 2; b++)
        for (int c = -2; c <= 2; c++) {
          if (a == 0 && b == 0 && c == 0) continue;
          for (int i = 0; i < 3; i++) ws_ineighc.push_back(x2c[i] * a + x2c[i + 3] * b + x2c[i + 6] * c);
        }
}

int system::identify_atom(const double x[3], double distmax) const {
  int best = 0;
  double dbest = 1e300;
  for (int i = 0; i < nat(); i++) {
    const double dx[3] = {x[0] - xat[3 * i], x[1] - xat[3 * i + 1], x[2] - xat[3 * i + 2]};
    const double d = shortest(m_x2c, dx);
    if (d < dbest) { dbest = d; best = i + 1; }
  }
  return (best > 0 && dbest <= distmax) ? best : 0;
}

bool system::are_lclose(const double x0[3], const double x1[3], double eps) const {
  const double dx[3] = {x0[0] - x1[0], x0[1] - x1[1], x0[2] - x1[2]};
  return shortest(m_x2c, dx) < eps;
}

void gpu_init(int device) {
  if (g_ctx) return;
  const int ier = c2g_init(device, &g_ctx);
  if (ier != 0) ferror("gpu_init", g_ctx ? c2g_last_error(g_ctx) : "no usable CUDA device (there is no CPU fallback)");
}

void gpu_end() {
  if (g_basins) c2g_basins_free(g_basins);
  if (g_ctx) c2g_finalize(g_ctx);
  g_basins = nullptr; g_ctx = nullptr; g_hgrid = -1;
}

bool gpu_enabled() { return g_ctx != nullptr; }

void bader_integrate(system& s, basindat& bas) {
  if (!g_ctx) ferror("bader_integrate", "gpu_init was not called");
  // metrics, bader@proc.f90:124-145
  double lat2car[9], car2lat[9], lid[27];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) lat2car[k + 3 * i] = s.m_x2c[k + 3 * i] / bas.n[i];
  matinv3(lat2car, car2lat);
  for (int i = -1; i <= 1; i++)
    for (int j = -1; j <= 1; j++)
      for (int k = -1; k <= 1; k++) {
        double r2 = 0.0;
        for (int c = 0; c < 3; c++) {
          const double d = lat2car[c] * i + lat2car[c + 3] * j + lat2car[c + 6] * k;
          r2 += d * d;
        }
        lid[(i + 1) * 9 + (j + 1) * 3 + (k + 1)] = (i || j || k) ? 1.0 / std::sqrt(r2) : 0.0;
      }
  upload_field(bas, "bader_integrate");
  int nmax = 0;
  check(c2g_bader_assign(g_ctx, g_hgrid, car2lat, lid, C2G_BADER_FAST, C2G_ORDER_SCAN, &nmax, &g_basins), "bader_integrate");
  bas.is_yt = false;
  finish_assignment(s, bas, nmax, "bader_integrate");
}

void yt_integrate(system& s, basindat& bas) {
  if (!g_ctx) ferror("yt_integrate", "gpu_init was not called");
  if (s.grid.nvec <= 0) ferror("yt_integrate", "the grid has no Voronoi stencil (init_geometry)");
  upload_field(bas, "yt_integrate");
  int nmax = 0;
  check(c2g_yt_build(g_ctx, g_hgrid, s.grid.nvec, s.grid.vec.data(), s.grid.area.data(), &nmax, &g_basins), "yt_integrate");
  bas.is_yt = true;
  finish_assignment(s, bas, nmax, "yt_integrate");
}

void yt_isosurface(system& s, basindat& bas) {
  if (!g_ctx) ferror("yt_isosurface", "gpu_init was not called");
  if (s.grid.nvec <= 0) ferror("yt_isosurface", "the grid has no Voronoi stencil (init_geometry)");
  upload_field(bas, "yt_isosurface");
  int nmax = 0, nraw = 0, nattr = 0;
  c2g_basins* yt = nullptr;
  check(c2g_yt_build(g_ctx, g_hgrid, s.grid.nvec, s.grid.vec.data(), s.grid.area.data(), &nmax, &yt), "yt_isosurface");
  const int rc = c2g_yt_isosurface(yt, bas.isov, &nraw, &nattr, &g_basins);
  c2g_basins_free(yt);
  check(rc, "yt_isosurface");
  bas.is_yt = false;
  std::vector<int> pmax(3 * (size_t)std::max(nraw, 1));
  check(c2g_basins_maxima(g_basins, pmax.data()), "yt_isosurface");
  bas.nattr = nattr;  // survivors; ids in idg run up to nraw (yt@proc.f90:337-352)
  bas.xattr.assign(3 * (size_t)nattr, 0.0);
  for (int i = 0; i < nattr; i++)
    for (int k = 0; k < 3; k++) bas.xattr[3 * i + k] = (pmax[3 * i + k] - 1.0) / bas.n[k];
  bas.idg.assign(bas.f.size(), 0);
  check(c2g_basins_labels(g_basins, bas.idg.data()), "yt_isosurface");
}

void voronoi_grid(system& s, basindat& bas) {
  if (!g_ctx) ferror("voronoi_grid", "gpu_init was not called");
  if (s.nat() < 1) ferror("voronoi_grid", "system does not have crystal");
  if (g_basins) { c2g_basins_free(g_basins); g_basins = nullptr; }
  // atoms are the attractors (hirshfeld@proc.f90:108-117); nodes -> nearest atom (:120)
  const std::vector<double>& seeds = s.cpcel.empty() ? s.xat : s.cpcel;
  bas.nattr = bas.atexist ? (int)(seeds.size() / 3) : 0;
  bas.xattr = bas.atexist ? seeds : std::vector<double>();
  check(c2g_voronoi_grid(g_ctx, bas.n, s.m_x2c, s.nat(), s.xat.data(), &g_basins), "voronoi_grid");
  bas.is_yt = false;
  bas.idg.assign((size_t)bas.n[0] * bas.n[1] * bas.n[2], 0);
  check(c2g_basins_labels(g_basins, bas.idg.data()), "voronoi_grid");
}

void intgrid_fields(const system& s, const basindat& bas, const std::vector<const double*>& fint, std::vector<int_result>& res,
                    std::vector<double>& vol) {
  if (!g_ctx || !g_basins) ferror("intgrid_fields", "no basin assignment on the device");
  const int nprop = (int)fint.size();
  std::vector<int> h(nprop, -1);
  for (int k = 0; k < nprop; k++) check(c2g_grid_upload(g_ctx, fint[k], bas.n, &h[k]), "intgrid_fields");
  // rows written by the library: bas.nattr, except after yt_isosurface, where region ids run beyond the number of
  // surviving regions; the reference's loops stop at bas%nattr (integration@proc.f90:1208, :1290) and so does this copy
  int nrow = 0;
  check(c2g_basins_nattr(g_basins, &nrow), "intgrid_fields");
  if (nrow < bas.nattr) ferror("intgrid_fields", "inconsistent number of attractors");
  std::vector<double> psum((size_t)nrow * std::max(nprop, 1)), v(nrow, 0.0);
  check(c2g_integrate(g_ctx, g_basins, nprop, h.data(), s.omega, psum.data(), v.data()), "intgrid_fields");
  for (int k = 0; k < nprop; k++) check(c2g_grid_free(g_ctx, h[k]), "intgrid_fields");
  vol.assign(v.begin(), v.begin() + bas.nattr);
  res.assign(nprop, int_result());
  for (int k = 0; k < nprop; k++) res[k].psum.assign(psum.begin() + (size_t)k * nrow, psum.begin() + (size_t)k * nrow + bas.nattr);
  // ONLY / ONLY_RANGE: the reference leaves the sums of the attractors it skips at zero (integration@proc.f90:1210, :1292)
  if ((int)bas.docelatom.size() == bas.nattr)
    for (int i = 0; i < bas.nattr; i++)
      if (!bas.docelatom[i]) {
        vol[i] = 0.0;
        for (int k = 0; k < nprop; k++) res[k].psum[i] = 0.0;
      }
}

void intgrid_multipoles(const system& s, const basindat& bas, const double* fint, int lmax,
                        const std::vector<unsigned char>& docelatom, std::vector<double>& mpole) {
  if (!g_ctx || !g_basins) ferror("intgrid_fields", "no basin assignment on the device");
  if (!docelatom.empty() && (int)docelatom.size() != bas.nattr) ferror("intgrid_fields", "docelatom has the wrong size");
  int nrow = 0;
  check(c2g_basins_nattr(g_basins, &nrow), "intgrid_fields");
  if (nrow != bas.nattr) ferror("intgrid_fields", "multipoles need one attractor position per basin id");
  int h = -1;
  check(c2g_grid_upload(g_ctx, fint, bas.n, &h), "intgrid_fields");
  mpole.assign((size_t)(lmax + 1) * (lmax + 1) * bas.nattr, 0.0);
  const int nws = (int)(s.ws_ineighc.size() / 3);
  check(c2g_integrate_multipoles(g_ctx, g_basins, h, lmax, bas.xattr.data(), docelatom.empty() ? nullptr : docelatom.data(),
                                 s.isortho ? 1 : 0, s.isortho_del ? 1 : 0, s.m_x2c, s.m_x2xr, s.m_xr2c, nws,
                                 nws ? s.ws_ineighc.data() : nullptr, s.omega, mpole.data()),
        "intgrid_fields");
  check(c2g_grid_free(g_ctx, h), "intgrid_fields");
}

void basins_remap(const system& s, const basindat& bas, int& nattn, std::vector<int>& iatt, std::vector<int>& ilvec,
                  std::vector<int>* idg1) {
  if (!g_ctx || !g_basins) ferror("bader_remap", "no basin assignment on the device");
  double c2x[9];
  matinv3(s.m_x2c, c2x);
  const int nws = (int)(s.ws_ineighc.size() / 3);
  if (idg1) idg1->assign(bas.f.size(), 0);
  int cap = 27 * std::max(bas.nattr, 1);
  for (;;) {
    iatt.assign(cap, 0);
    ilvec.assign(3 * (size_t)cap, 0);
    const int ier = c2g_basins_remap(g_ctx, g_basins, bas.xattr.data(), c2x, s.isortho ? 1 : 0, s.isortho_del ? 1 : 0, s.m_x2c,
                                     s.m_x2xr, s.m_xr2c, nws, nws ? s.ws_ineighc.data() : nullptr, cap, &nattn, iatt.data(),
                                     ilvec.data(), idg1 ? idg1->data() : nullptr);
    if (ier == C2G_ERR_OVERFLOW && nattn > cap) { cap = nattn; continue; }  // the capacity needed comes back in nattn
    check(ier, "bader_remap");
    break;
  }
  iatt.resize(nattn);
  ilvec.resize(3 * (size_t)nattn);
}

void hirshfeld_fields(const system& s, basindat& bas, const std::vector<int>& ispc, const atomic_grids& g,
                      const std::vector<const double*>& fint, std::vector<int_result>& res, std::vector<double>& vol) {
  if (!g_ctx) ferror("intgrid_hirshfeld_fields", "gpu_init was not called");
  const int nat = s.nat(), nprop = (int)fint.size();
  if ((int)ispc.size() != nat) ferror("intgrid_hirshfeld_fields", "one species per atom is required");
  if (g.nspc() == 0) ferror("intgrid_hirshfeld_fields", "hirshfeld requires atomic grids");  // integration@proc.f90:1544
  int hp = -1;
  check(c2g_promolecular_grid(g_ctx, bas.n, s.m_x2c, nat, s.xat.data(), ispc.data(), g.nspc(), g.ngrid.data(), g.off.data(),
                              g.a.data(), g.b.data(), g.rmax.data(), g.rcut.data(), g.rtab.data(), g.ftab.data(), nullptr, &hp),
        "intgrid_hirshfeld_fields");
  bas.f.assign((size_t)bas.n[0] * bas.n[1] * bas.n[2], 0.0);
  check(c2g_grid_download(g_ctx, hp, bas.f.data()), "intgrid_hirshfeld_fields");   // bas%f = promolecular density (:264-267)
  bas.nattr = nat;                                                               // hirsh_grid: the atoms are the attractors
  bas.xattr = s.xat;
  std::vector<int> h(std::max(nprop, 1), -1);
  for (int k = 0; k < nprop; k++) check(c2g_grid_upload(g_ctx, fint[k], bas.n, &h[k]), "intgrid_hirshfeld_fields");
  std::vector<double> psum((size_t)nat * std::max(nprop, 1), 0.0);
  vol.assign(nat, 0.0);
  const bool masked = (int)bas.docelatom.size() == nat;
  check(c2g_hirshfeld_integrate(g_ctx, hp, s.m_x2c, nat, s.xat.data(), ispc.data(), g.nspc(), g.ngrid.data(), g.off.data(), g.a.data(),
                                g.b.data(), g.rmax.data(), g.rcut.data(), g.rtab.data(), g.ftab.data(),
                                masked ? bas.docelatom.data() : nullptr, nprop, h.data(), s.omega, psum.data(), vol.data()),
        "intgrid_hirshfeld_fields");
  for (int k = 0; k < nprop; k++) check(c2g_grid_free(g_ctx, h[k]), "intgrid_hirshfeld_fields");
  check(c2g_grid_free(g_ctx, hp), "intgrid_hirshfeld_fields");
  res.assign(nprop, int_result());
  for (int k = 0; k < nprop; k++) res[k].psum.assign(psum.begin() + (size_t)k * nat, psum.begin() + (size_t)(k + 1) * nat);
}

void yt_weights(const basindat& bas, int idb, std::vector<double>& w) {
  if (!g_ctx || !g_basins || !bas.is_yt) ferror("yt_weights", "no YT assignment on the device");
  w.assign(bas.f.size(), 0.0);
  check(c2g_yt_weights(g_basins, idb, w.data()), "yt_weights");
}

void nci_rdg(const system& s, std::vector<double>& crho, std::vector<double>& cgrad) {
  if (!g_ctx) ferror("nciplot", "gpu_init was not called");
  const grid3& g = s.grid;
  int h = -1;
  check(c2g_grid_upload(g_ctx, g.f.data(), g.n, &h), "nciplot");
  double c2x[9], xmat[9];
  matinv3(s.m_x2c, c2x);
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) xmat[k + 3 * i] = s.m_x2c[k + 3 * i] / g.n[i];  // nci@proc.f90:426
  const double x0[3] = {0.0, 0.0, 0.0};
  crho.assign(g.f.size(), 0.0);
  cgrad.assign(g.f.size(), 0.0);
  check(c2g_nci_rdg(g_ctx, h, x0, xmat, g.n, c2x, s.m_x2c, c2x, 0, nullptr, crho.data(), cgrad.data()), "nciplot");
  check(c2g_grid_free(g_ctx, h), "nciplot");
}

void grid_fft(const system& s, const grid3& fold, int iff, grid3& fnew) {
  if (!g_ctx) ferror("fft", "gpu_init was not called");
  if (fold.f.empty()) ferror("fft", "no input grid");  // grid3mod@proc.f90:1775
  int h = -1, ho = -1;
  check(c2g_grid_upload(g_ctx, fold.f.data(), fold.n, &h), "fft");
  check(c2g_fft_derivative(g_ctx, h, iff, s.m_x2c, &ho), "fft");
  fnew = grid3();
  for (int i = 0; i < 3; i++) fnew.n[i] = fold.n[i];  // copy_geometry (:1781)
  fnew.nvec = fold.nvec; fnew.vec = fold.vec; fnew.area = fold.area;
  fnew.f.assign(fold.f.size(), 0.0);
  check(c2g_grid_download(g_ctx, ho, fnew.f.data()), "fft");
  check(c2g_grid_free(g_ctx, h), "fft");
  check(c2g_grid_free(g_ctx, ho), "fft");
}

void nci_rdg_fourier(const system& s, std::vector<double>& crho, std::vector<double>& cgrad) {
  if (!g_ctx) ferror("nciplot", "gpu_init was not called");
  const grid3& g = s.grid;
  int h[5] = {-1, -1, -1, -1, -1};
  check(c2g_grid_upload(g_ctx, g.f.data(), g.n, &h[0]), "nciplot");
  const int iffs[4] = {C2G_FT_GRAD, C2G_FT_XX, C2G_FT_YY, C2G_FT_ZZ};  // nci@proc.f90:528-531
  for (int q = 0; q < 4; q++) check(c2g_fft_derivative(g_ctx, h[0], iffs[q], s.m_x2c, &h[q + 1]), "nciplot");
  double c2x[9], xmat[9];
  matinv3(s.m_x2c, c2x);
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) xmat[k + 3 * i] = s.m_x2c[k + 3 * i] / g.n[i];
  const double x0[3] = {0.0, 0.0, 0.0};
  crho.assign(g.f.size(), 0.0);
  cgrad.assign(g.f.size(), 0.0);
  check(c2g_nci_rdg_fourier(g_ctx, h, x0, xmat, g.n, c2x, c2x, crho.data(), cgrad.data()), "nciplot");
  for (int q = 0; q < 5; q++) check(c2g_grid_free(g_ctx, h[q]), "nciplot");
}

size_t grid_read_text(const std::string& text, bool k_fastest, double divisor, grid3& g) {
  if (!g_ctx) ferror("read_cube", "gpu_init was not called");
  int h = -1;
  size_t consumed = 0;
  long long nslow = 0;
  check(c2g_grid_parse_text(g_ctx, text.data(), text.size(), g.n, k_fastest ? C2G_TEXT_ORDER_K_FASTEST : C2G_TEXT_ORDER_I_FASTEST,
                            divisor, &h, &consumed, &nslow), "read_cube");
  g.f.assign((size_t)g.n[0] * g.n[1] * g.n[2], 0.0);
  check(c2g_grid_download(g_ctx, h, g.f.data()), "read_cube");
  check(c2g_grid_free(g_ctx, h), "read_cube");
  return consumed;
}

std::string grid_write_text(const grid3& g, bool cube_order, const int ishift[3], int width, int digits, int scale) {
  if (!g_ctx) ferror("writegrid_cube", "gpu_init was not called");
  int h = -1;
  check(c2g_grid_upload(g_ctx, g.f.data(), g.n, &h), "writegrid_cube");
  const int layout = cube_order ? C2G_TEXT_ROWS_INDEX3 : C2G_TEXT_ROWS_INDEX1;
  size_t nbytes = 0;
  check(c2g_grid_format_text(g_ctx, h, layout, ishift, width, digits, scale, nullptr, 0, &nbytes), "writegrid_cube");
  std::string out(nbytes, ' ');
  check(c2g_grid_format_text(g_ctx, h, layout, ishift, width, digits, scale, &out[0], out.size(), &nbytes), "writegrid_cube");
  check(c2g_grid_free(g_ctx, h), "writegrid_cube");
  return out;
}

}  // namespace c2h
