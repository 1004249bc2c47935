// critic2_host.hpp -- C++ mirror of the reference's driver interface for the on-grid QTAIM path.
//
// critic2's host code is Fortran; this image has no Fortran compiler, so the host side above the C ABI
// (include/critic2_gpu.h) is mirrored here in C++ with the reference's names, argument meaning and error
// behaviour.  Nothing in this directory computes on the CPU what the GPU path computes: every routine
// marshals the reference's objects into C-ABI calls, exactly like fortran/critic2_gpu.f90.
//
//   types.f90:361-383          basindat      ->  c2h::basindat
//   grid3mod.f90:60-125        grid3         ->  c2h::grid3
//   systemmod.f90:44-86        system        ->  c2h::system (crystal cell + atoms + one reference grid field)
//   bader@proc.f90:80-234      bader_integrate(s,bas,iref)
//   yt@proc.f90:38-224         yt_integrate(s,bas)
//   integration@proc.f90:1170  intgrid_fields(bas,res)
//   integration@proc.f90:1302  multipole branch of intgrid_fields -> intgrid_multipoles
//   yt@proc.f90:233-390        yt_isosurface(s,bas)
//   bader@proc.f90:237-296     bader_remap / yt@proc.f90:533-594 yt_remap -> basins_remap
//   integration@proc.f90:1397  intgrid_hirshfeld_fields (+ promolecular_array3 for bas%f) -> hirshfeld_fields
//   nci@proc.f90:543-605       nciplot loop -> nci_rdg, nci_rdg_fourier
//   grid3mod@proc.f90:1757     grid3%fft -> grid_fft
//   grid3mod@proc.f90:559,884  read_cube / read_vasp value blocks -> grid_read_text
//   crystalmod@write.f90:3556  writegrid_cube / write_cube_body value loops -> grid_write_text
//   tools_io@proc.F90:1573     ferror(routine,msg,faterr) -> c2h::ferror (throws c2h::fatal_error)
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/critic2_gpu.h"

namespace c2h {

struct fatal_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// ferror(routine,msg,faterr): the reference prints "ERROR : routine: msg" and stops (tools_io@proc.F90:1573-1643)
[[noreturn]] void ferror(const std::string& routine, const std::string& msg);

constexpr double vsmall = 1e-80;  // param.F90:27

// grid3mod.f90:60-125 (the members the path reads)
struct grid3 {
  int n[3] = {0, 0, 0};
  std::vector<double> f;       // f(n1,n2,n3), index 1 fastest
  int nvec = 0;                // Voronoi-relevant grid steps (init_geometry, grid3mod@proc.f90:3197)
  std::vector<int> vec;        // vec(3,nvec)
  std::vector<double> area;    // area(nvec)
};

// crystal + field: what bader_integrate / yt_integrate read through sy%c and sy%f(iref)
struct system {
  double m_x2c[9] = {0};       // crystallographic -> Cartesian, column-major (crystalmod)
  double omega = 0.0;          // cell volume
  // what crystal%shortest reads (crystalmod.f90:173-209, crystalmod@proc.f90:1056-1085); filled by set_cell
  bool isortho = false, isortho_del = false;
  double m_x2xr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, m_xr2c[9] = {0};
  std::vector<double> ws_ineighc;  // (3,ws_nf) Cartesian
  std::vector<double> xat;     // atoms (nuclear CPs of the reference field), crystallographic, (3,nat)
  std::vector<double> cpcel;   // sy%f(iref)%cpcel(:)%x, (3,ncpcel): the nuclei followed by any critical points found or
                               // loaded before the integration (non-nuclear maxima).  They seed the attractor list
                               // (bader@proc.f90:113-118, yt@proc.f90:66-76).  Empty = no CP beyond the nuclei: xat is used.
  grid3 grid;                  // sy%f(iref)%grid
  int nat() const { return (int)(xat.size() / 3); }
  // crystalmod@env.f90:593-614: id (1-based) of the atom within distmax of x (cryst.), 0 if none
  int identify_atom(const double x[3], double distmax) const;
  // crystalmod@proc.f90:1118-1133: are two crystallographic points closer than eps (any lattice translation)?
  bool are_lclose(const double x0[3], const double x1[3], double eps) const;
  void set_cell(const double x2c[9]);
};

// types.f90:361-383
struct basindat {
  bool atexist = true;         // attractors are assigned to atoms (integration@proc.f90:243-252)
  double ratom = 1.0;          // distance below which a maximum is an atom / a known attractor
  int n[3] = {0, 0, 0};
  std::vector<double> f;       // copy of the reference field (integration@proc.f90:255-271)
  int nattr = 0;
  std::vector<double> xattr;   // (3,nattr) crystallographic
  std::vector<int> idg;        // idg(n1,n2,n3)
  bool is_yt = false;          // weights live on the device (the reference's luw scratch unit)
  std::vector<unsigned char> docelatom;  // per attractor (docelatom(icp(i)), ONLY / ONLY_RANGE); empty = all
  double isov = 0.0;           // ISOSURFACE: contour value (already negated by the driver for LOWER, integration@proc.f90:260)
  std::string expr;            // DISCARD expression (bas%expr).  It is evaluated by critic2's arithmetic module per new
                               // maximum (bader@proc.f90:184-190, yt@proc.f90:152-166): the Fortran shim calls s%eval; this
                               // C++ mirror has no expression evaluator and REFUSES a non-empty expression (ferror)
};

// types.f90:394-410 (the sums only)
struct int_result {
  std::vector<double> psum;    // psum(nattr)
};

// One GPU context per process, like the module variables of fortran/critic2_gpu.f90.
void gpu_init(int device = 0);   // CRITIC2_GPU=1 in the Fortran shim
void gpu_end();
bool gpu_enabled();

// bader@proc.f90:80-234.  Fills bas.idg, bas.nattr, bas.xattr.
void bader_integrate(system& s, basindat& bas);
// yt@proc.f90:38-224.  bas.idg = spatial basin id, 0 on interatomic-surface points.
void yt_integrate(system& s, basindat& bas);
// yt@proc.f90:233-390 (no DISCARD expression): regions of f >= bas.isov.  bas.idg = region ids in the reference's
// numbering, bas.nattr = surviving regions, bas.xattr = the first nattr regional maxima (:359).
void yt_isosurface(system& s, basindat& bas);
// hirshfeld@proc.f90:93-122 (VORONOI): bas.idg = id of the atom nearest to every grid node (crystal%nearest_atom_grid),
// attractors = the atoms; intgrid_fields then integrates over it like over Bader basins.  bas.n must be set.
void voronoi_grid(system& s, basindat& bas);
// integration@proc.f90:1170-1391: res[k].psum(i) = integral of fint[k] over basin i; vol(i) = basin volume.
void intgrid_fields(const system& s, const basindat& bas, const std::vector<const double*>& fint,
                    std::vector<int_result>& res, std::vector<double>& vol);
// integration@proc.f90:1302-1361 (INTEGRABLE ... MULTIPOLES): mpole((lmax+1)^2, nattr), genrlm_real's order; docelatom
// (may be empty) is the per-attractor mask of the YT branch (:1318).
void intgrid_multipoles(const system& s, const basindat& bas, const double* fint, int lmax,
                        const std::vector<unsigned char>& docelatom, std::vector<double>& mpole);
// bader@proc.f90:237-296 / yt@proc.f90:533-594 (DELOC): attractor images of the basins of the last bader_integrate /
// yt_integrate call.  iatt(nattn), ilvec(3,nattn); idg1 (Bader only, pass nullptr otherwise) is resized to the grid.
void basins_remap(const system& s, const basindat& bas, int& nattn, std::vector<int>& iatt, std::vector<int>& ilvec,
                  std::vector<int>* idg1);
// grid1mod.f90: the atomic radial grids of the species, packed like fortran/critic2_gpu.f90's pack_atomic_grids
struct atomic_grids {
  std::vector<int> ngrid, off;                 // per species; ngrid = 0: no usable grid
  std::vector<double> a, b, rmax, rcut;        // r(i) = a exp(b (i-1)); rcut = min(cutrad(z), rmax)
  std::vector<double> rtab, ftab;              // concatenated r(i), f(i)
  int nspc() const { return (int)ngrid.size(); }
};
// HIRSHFELD on a grid (integration@proc.f90:264-267 and :1552-1596): fills bas.f with the promolecular density of the
// atoms s.xat with species ispc (1-based), then res[k].psum(A) / vol(A) per atom; bas.docelatom = the ONLY mask.
void hirshfeld_fields(const system& s, basindat& bas, const std::vector<int>& ispc, const atomic_grids& g,
                      const std::vector<const double*>& fint, std::vector<int_result>& res, std::vector<double>& vol);
// yt@proc.f90:476-499 (yt_weights with idb): dense weight field of one basin.
void yt_weights(const basindat& bas, int idb, std::vector<double>& w);
// nci@proc.f90:543-605, grid interpolation mode on the field's own lattice:
// crho, cgrad(0:n3-1,0:n2-1,0:n1-1), third index fastest.
void nci_rdg(const system& s, std::vector<double>& crho, std::vector<double>& cgrad);
// the same loop with FOURIER interpolation (nci@proc.f90:527-565): the derived grids are built on the device
void nci_rdg_fourier(const system& s, std::vector<double>& crho, std::vector<double>& cgrad);
// grid3%fft (grid3mod@proc.f90:1757-1872): fnew = FFT-derived field of fold; iff = ifformat_as_ft_* (param.F90:225-236)
void grid_fft(const system& s, const grid3& fold, int iff, grid3& fnew);
// the value block of read_cube (k_fastest = true, grid3mod@proc.f90:559) / read_vasp (false, :884; divisor = det3(x2c)
// with vscal): g.n must be set; fills g.f; returns the offset of the first byte after the last value
size_t grid_read_text(const std::string& text, bool k_fastest, double divisor, grid3& g);
// the value loops of writegrid_cube (cube_order = true: width 12, digits 5, scale 1, or precisecube 22, 14, 0;
// crystalmod@write.f90:3556-3565) and write_cube_body (cube_order = false: 13, 5, 1; nci@proc.f90:916-932)
std::string grid_write_text(const grid3& g, bool cube_order, const int ishift[3], int width, int digits, int scale);

}  // namespace c2h
