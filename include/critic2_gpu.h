/*
 * critic2_gpu.h -- C ABI of the B200-native on-grid QTAIM hot path for critic2.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * critic2 has no plugin interface for this path (libcritic2.h:46-185 exports
 * crystal/XRPD calls only), so the entry points below are what a thin
 * ISO_C_BINDING module placed next to src/c_interface_module.f90 would bind
 * (see INTEGRATION.md and fortran/critic2_gpu.f90).  Conventions follow the
 * reference's two existing Fortran<->C idioms:
 *   - opaque context handles + integer status (doqhull.c:19-27,
 *     tools@proc.f90:709-728): every call returns 0 on success;
 *     c2g_last_error() gives the message the Fortran shim forwards to
 *     ferror(routine,msg,faterr) (tools_io@proc.F90:1573-1643);
 *   - "step 1 returns sizes, step 2 fills caller-allocated arrays".
 *
 * All arrays are caller-owned HOST memory in Fortran layout: f(n1,n2,n3) with
 * index 1 fastest; 3x3 matrices column-major (m[i+3*j] = M(i+1,j+1)).  Grid
 * coordinates returned to the caller are 1-based like the Fortran code.
 * Calls must come from one thread at a time per context (all reference call
 * sites are outside OpenMP regions: integration@proc.f90:288,298; nci@proc.f90:525).
 *
 * There is NO CPU fallback: every compute entry point fails with a non-zero
 * status if no CUDA device is usable.
 */
#ifndef CRITIC2_GPU_H
#define CRITIC2_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c2g_context c2g_context; /* one per process / GPU */
typedef struct c2g_basins c2g_basins;   /* result of a BADER or YT assignment (device resident) */

/* ---- status codes ---- */
#define C2G_OK 0
#define C2G_ERR_CUDA 1      /* CUDA runtime error (message has details) */
#define C2G_ERR_ARG 2       /* invalid argument */
#define C2G_ERR_NOMEM 3     /* device allocation failed */
#define C2G_ERR_STATE 4     /* call sequence error (e.g. labels before set_map) */
#define C2G_ERR_NEWMAX 5    /* internal consistency: walk ended on an unknown maximum */
#define C2G_ERR_OVERFLOW 6  /* internal table overflow */
#define C2G_ERR_NCCL 7

/* ---- context ---- */
/* Create a context on CUDA device `device` (single-GPU use). */
int c2g_init(int device, c2g_context** ctx);
/* Multi-GPU from ONE process -- the form critic2 needs (src/critic2.F90:24-106 is a single program; SURVEY.md 8b
 * `c2g_init(ngpus)`): the context drives devices 0..ngpus-1 of the box with one worker thread per device and its own
 * NCCL communicator.  Every entry point below takes WHOLE host arrays exactly as on one GPU: c2g_grid_upload scatters
 * the z-slabs of the array (each device copies its slab, NVLink replicates them), c2g_bader_assign / c2g_integrate /
 * c2g_integrate_multipoles / c2g_basins_remap run sharded over NCCL, c2g_basins_labels assembles idg(n1,n2,n3),
 * c2g_nci_rdg* shard the output rows.  c2g_fft_derivative runs replicated, c2g_yt_* on the first device ("replicas
 * only").  The text codec, HIRSHFELD, WCUBE weight grids and c2g_nci_rdg_resident return C2G_ERR_STATE
 * on such a context.  ngpus = 1 is c2g_init(0). */
int c2g_init_devices(int ngpus, c2g_context** ctx);
/* Multi-GPU, one process per GPU.  `nccl_uid` is the 128-byte ncclUniqueId created by
 * c2g_nccl_unique_id() on rank 0 and distributed by the caller (MPI/torch.distributed/file). */
int c2g_nccl_unique_id(void* uid128);
int c2g_init_multi(int device, int rank, int nranks, const void* nccl_uid128, c2g_context** ctx);
void c2g_finalize(c2g_context* ctx);
const char* c2g_last_error(const c2g_context* ctx);
/* Library / device description, e.g. "critic2_gpu 0.1 sm_100 NVIDIA B200 148 SMs". */
const char* c2g_describe(c2g_context* ctx);

/* ---- grid3 fields resident in HBM (grid3mod.f90:60-125: n(3), f(:,:,:)) ---- */
/* Replaces the bas%f copy of intgrid_driver (integration@proc.f90:255-271). */
int c2g_grid_upload(c2g_context* ctx, const double* f, const int n[3], int* handle);
int c2g_grid_alloc(c2g_context* ctx, const int n[3], int* handle);
/* Asynchronous upload on a separate copy stream: returns at once, every call that reads the grid is ordered after
 * the copy on the device.  `f` should be page-locked (pinned) for a real overlap and must stay valid and unchanged
 * until c2g_synchronize.  Lets the upload of the second INTEGRABLE field overlap c2g_bader_assign on the first. */
int c2g_grid_upload_async(c2g_context* ctx, const double* f, const int n[3], int* handle);
/* Multi-GPU (c2g_init_multi): the grid is sharded as z-slabs f(:,:,zlo+1:zhi) across the ranks
 * (0-based half-open plane range from c2g_slab_range; boundaries are multiples of 4).  Every rank
 * uploads only its own slab; the slabs are replicated over NVLink (NCCL) because the field is
 * read-only and trajectories may cross slabs.  Labels and integrals stay sharded / all-reduced. */
int c2g_slab_range(c2g_context* ctx, int n3, int* zlo, int* zhi);
/* same without a context (pure host arithmetic; usable on a machine without a GPU) */
int c2g_slab_bounds_query(int n3, int nranks, int rank, int* zlo, int* zhi);
int c2g_grid_upload_slab(c2g_context* ctx, const double* fslab, const int n[3], int* handle);
int c2g_grid_download(c2g_context* ctx, int handle, double* f);
/* this rank's z-slab f(:,:,zlo+1:zhi) only */
int c2g_grid_download_slab(c2g_context* ctx, int handle, double* fslab);
int c2g_grid_free(c2g_context* ctx, int handle);
/* Fill a resident grid with the synthetic promolecular-like density
 *   rho(x) = sum_atoms sum_images Z exp(-alpha r) cut(r), cut = (1-(r/rc)^2)^3 for rc>0 else 1
 * (stand-in for LOAD AS PROMOLECULAR / promolecular_array3, crystalmod@complex.f90:436-470;
 * used to create bench inputs of 512^3..1024^3 directly in HBM). xat: (3,nat) cryst. */
int c2g_grid_promolecular(c2g_context* ctx, int handle, const double x2c[9], int nat, const double* xat,
                          const double* zat, const double* alpha, int nimg, double rc);

/* ---- BADER: near-grid basin assignment (bader@proc.f90:147-224) ---- */
/* algo: */
#define C2G_BADER_FAST 0    /* hierarchical exact walks + edge refinement (default) */
#define C2G_BADER_EXACT 1   /* referee: every point follows its own complete trajectory */
/* order of the returned maxima: */
#define C2G_ORDER_INDEX 0   /* by linear grid index of the maximum */
#define C2G_ORDER_SCAN 1    /* by the reference scan-order key ((i1-1)*n2+(i2-1))*n3+(i3-1) of the
                               first point of each basin (approximates discovery order, NOATOMS mode) */
/* Step 1: label every point with its terminal maximum.  car2lat = inverse of lat2car
 * (bader@proc.f90:124-129, computed by the caller with matinv), lat_i_dist(-1:1,-1:1,-1:1)
 * flattened as (d1+1)*9+(d2+1)*3+(d3+1) (:131-145).  Returns the number of distinct maxima. */
int c2g_bader_assign(c2g_context* ctx, int handle, const double car2lat[9], const double lat_i_dist[27],
                     int algo, int order, int* nmax, c2g_basins** res);
/* Step 2: grid coordinates (1-based, 3 x nmax) of the maxima, in the requested order.  The caller
 * keeps the attractor identification of bader@proc.f90:160-199 (identify_atom, are_lclose, DISCARD). */
int c2g_basins_maxima(c2g_basins* res, int* pmax);
/* Number of grid points per maximum (nmax entries). */
int c2g_basins_counts(c2g_basins* res, long long* counts);
/* map(nmax): maximum -> basin id (1..nattr; 0 = discarded attractor, points stay unassigned). */
int c2g_basins_set_map(c2g_basins* res, int nattr, const int* map);
/* Number of basin ids of the current map = rows of psum / vol / mpole written by c2g_integrate and
 * c2g_integrate_multipoles (for an ISOSURFACE result: nraw, the largest region id). */
int c2g_basins_nattr(c2g_basins* res, int* nattr);
/* bas%idg(n1,n2,n3) (move_alloc(volnum,bas%idg), bader@proc.f90:229).  Multi-GPU: every rank
 * receives its own slab idg(:,:,zlo+1:zhi). */
int c2g_basins_labels(c2g_basins* res, int* idg);
/* Asynchronous variant: the device-to-host copy runs on a separate stream and overlaps later calls; idg (page-locked
 * for a real overlap) is complete after c2g_synchronize. */
int c2g_basins_labels_async(c2g_basins* res, int* idg);
/* int_reorder_gridout's nattr0 full-grid `where` passes (integration@proc.f90:1113-1122,
 * :1139-1144) as one composition of maps: new id = assigned(old id), old ids 1..nattr0. */
int c2g_basins_relabel(c2g_basins* res, int nattr0, const int* assigned, int nattr_new);
void c2g_basins_free(c2g_basins* res);
/* statistics of the last assignment: [0] walked points, [1] edge-fix passes, [2] edge-fix points,
 * [3] path overflows, [4] candidate maxima, [5] total walker steps (EXACT algo only, else 0) */
int c2g_basins_stats(c2g_basins* res, long long stats[8]);

/* ---- INTEGRABLE: basin integration (integration@proc.f90:1208-1218, :1289-1299) ---- */
/* psum(nattr,nprop) column-major = sum(fint_k, idg==i)*omega/ntot (Bader) or sum(w_i*fint_k)*omega/ntot
 * (YT); vol(nattr) = count(idg==i)*omega/ntot or sum(w_i)*omega/ntot.  Requires c2g_basins_set_map.
 * fieldhandles: nprop resident grids with the same n. */
int c2g_integrate(c2g_context* ctx, c2g_basins* res, int nprop, const int* fieldhandles, double omega,
                  double* psum, double* vol);

/* ---- INTEGRABLE id MULTIPOLES [lmax]: basin multipole moments (integration@proc.f90:1302-1361) ---- */
/* mpole((lmax+1)^2, nattr) column-major, real regular solid harmonics in genrlm_real's order
 * (tools_math@proc.f90:273-306: C00, C11, C10, S11, C22, C21, C20, S21, S22, ...):
 *   Bader labels (:1338-1358): mpole(:,ix) = sum over the points of basin ix of rrlm(dv) * fint,
 *   YT weights  (:1316-1336): mpole(:,m)  = sum over |w_m| >= 1e-15 of rrlm(dv) * fint * w_m, basins with
 *                                           domask(m) == 0 (the reference's docelatom(icp(m))) are skipped;
 *   dv = shortest(p/n - xattr(:,basin)), then * omega / ntot (:1360).  xattr(3,nattr) crystallographic.
 * The cell arguments are what crystal%shortest reads (crystalmod@proc.f90:1056-1085): isortho, isortho_del, m_x2c,
 * m_x2xr, m_xr2c and ws_ineighc(3,ws_nf) (Cartesian); the last four may be NULL / 0 when isortho != 0.
 * domask may be NULL (all basins).  lmax <= 10 (the reference default is 5, systemmod@proc.f90:1014). */
int c2g_integrate_multipoles(c2g_context* ctx, c2g_basins* res, int fieldhandle, int lmax, const double* xattr,
                             const unsigned char* domask, int isortho, int isortho_del, const double x2c[9],
                             const double x2xr[9], const double xr2c[9], int nws, const double* ws_ineighc,
                             double omega, double* mpole);

/* ---- DELOC attractor images: bader_remap (bader@proc.f90:237-296), yt_remap (yt@proc.f90:533-594) ---- */
/* For every point: x = p/n - xattr(:,basin), xs = shortest(x), lattice vector nint(x - c2x(xs)); a non-zero vector
 * makes the point a member of an IMAGE of its attractor.  iatt(1:nattr) = 1..nattr, ilvec(:,1:nattr) = 0; the images
 * follow, numbered in the order in which the reference's scan (index 1 fastest) first meets them (YT: basin by basin,
 * points with |w| >= 1e-15).  iatt(nattn): attractor of every image, ilvec(3,nattn): its lattice vector,
 * idg1(n1,n2,n3) (Bader labels / ISOSURFACE regions only, may be NULL; must be NULL for a YT result): image id per
 * point.  c2x = crystal%m_c2x; the other cell arguments as in c2g_integrate_multipoles.  Caller-allocated outputs of
 * capacity maxattn; when there are more images the call returns C2G_ERR_OVERFLOW with *nattn set to the number
 * needed, and the caller retries.  Multi-GPU Bader: idg1 is the rank's slab, iatt/ilvec are global. */
int c2g_basins_remap(c2g_context* ctx, c2g_basins* res, const double* xattr, const double c2x[9], int isortho,
                     int isortho_del, const double x2c[9], const double x2xr[9], const double xr2c[9], int nws,
                     const double* ws_ineighc, int maxattn, int* nattn, int* iatt, int* ilvec, int* idg1);

/* ---- HIRSHFELD on a grid (hirshfeld@proc.f90:28-89, integration@proc.f90:264-267, :1397-1597) ---- */
/* Atomic radial grids (grid1mod.f90) per species s = 1..nspc: spc_ngrid(s) nodes r(i) = a exp(b (i-1)) stored at
 * rtab(spc_off(s) + i), densities at ftab(...), spc_rmax = g%rmax, spc_rcut = min(cutrad(z), g%rmax)
 * (crystalmod@env.f90:671-684); spc_ngrid(s) = 0 marks a species without a usable grid (z = 0, z - qat <= 0, not
 * initialised).  Atoms: xat(3,nat) crystallographic, ispc(nat) 1-based species.
 * c2g_promolecular_grid = promolecular_array3 (crystalmod@complex.f90:436-470): the promolecular density on the grid as a
 * resident field (bas%f of the HIRSHFELD driver); infrag(nat) (may be NULL) restricts the sum to a fragment, e.g. one
 * atom for hirsh_weights (hirshfeld@proc.f90:78-86: w = that grid / max(bas%f, vsmall)). */
int c2g_promolecular_grid(c2g_context* ctx, const int n[3], const double x2c[9], int nat, const double* xat, const int* ispc,
                          int nspc, const int* spc_ngrid, const int* spc_off, const double* spc_a, const double* spc_b,
                          const double* spc_rmax, const double* spc_rcut, const double* rtab, const double* ftab,
                          const unsigned char* infrag, int* handle);
/* The grid loop of intgrid_hirshfeld_fields (integration@proc.f90:1552-1596): psum(nat,nprop) column-major =
 * sum_p rho_A(p) / max(rho_pro(p), vsmall) * f_k(p) * omega/ntot over every image of atom A within its cutoff,
 * vol(nat) the same without f_k; atoms with domask(A) == 0 (docelatom(icp(A)), may be NULL) stay at zero.
 * hpromol: the resident promolecular grid (bas%f). */
int c2g_hirshfeld_integrate(c2g_context* ctx, int hpromol, const double x2c[9], int nat, const double* xat, const int* ispc,
                            int nspc, const int* spc_ngrid, const int* spc_off, const double* spc_a, const double* spc_b,
                            const double* spc_rmax, const double* spc_rcut, const double* rtab, const double* ftab,
                            const unsigned char* domask, int nprop, const int* fieldhandles, double omega, double* psum,
                            double* vol);
/* ---- VORONOI on a grid: voronoi_grid (hirshfeld@proc.f90:93-122) = crystal%nearest_atom_grid (crystalmod@proc.f90:1138-1167) ---- */
/* idg(i,j,k) = complete-list id (1-based, the order of xat) of the atom nearest to the grid node, any periodic image.  The
 * result is a c2g_basins with plain labels and the identity map already set (the attractors of VORONOI are the atoms,
 * :108-117): c2g_basins_labels gives bas%idg, c2g_integrate / c2g_integrate_multipoles the atomic properties exactly like the
 * Bader branch of intgrid_fields.  Nodes equidistant from two atoms go to the lower id (the reference's choice follows the
 * traversal order of list_near_atoms; such nodes are a set of measure zero unless the structure is symmetric about them).
 * Single-device contexts. */
int c2g_voronoi_grid(c2g_context* ctx, const int n[3], const double x2c[9], int nat, const double* xat, c2g_basins** res);

/* ---- YT: Yu-Trinkle weights (yt@proc.f90:77-211) ---- */
/* vec(3,nvec), area(nvec): Voronoi-relevant grid steps and facet areas from grid3%init_geometry
 * (grid3mod@proc.f90:3197).  Maxima are returned in decreasing density (= reference discovery order). */
int c2g_yt_build(c2g_context* ctx, int handle, int nvec, const int* vec, const double* area, int* nmax,
                 c2g_basins** res);
/* spatial basin id per point, 0 = interatomic-surface (fractional) point:
 * ibasin(iio(.)) of the ytdata record (yt.f90:36-45). */
/* (use c2g_basins_labels) */
/* dense weight field of one basin, w(n1,n2,n3) (yt_weights, yt@proc.f90:476-499). */
int c2g_yt_weights(c2g_basins* res, int idb, double* w);
/* The ytdata record (yt.f90:36-45) that yt_integrate writes to bas%luw (yt@proc.f90:191-199) and that yt_weights
 * (:399-530), int_reorder_gridout (integration@proc.f90:1125-1158) and the BASINS / DI consumers read back, for hosts
 * that keep using their own yt_weights: nlo(nn), ibasin(nn), iio(nn), inear(nvec,nn), fnear(nvec,nn), all indexed by
 * the position of a point in the density-sorted list (1 = lowest) except iio (grid index -> position).  ibasin
 * carries the ids of the current c2g_basins_set_map / c2g_basins_relabel (what int_reorder_gridout rewrites).  The order
 * is that of a stable sort of (density, index) = the reference's qcksort order on tie-free data.  inear / fnear may be
 * NULL (they are nvec*nn entries each). */
int c2g_yt_export(c2g_basins* res, int* nlo, int* ibasin, int* iio, int* inear, double* fnear);
/* WCUBE (int_cubew, integration@proc.f90:4428-4466): the weight field of basin idb as a RESIDENT grid -- the YT
 * weights (yt_weights with idb, :4451) or, for Bader labels and ISOSURFACE regions, w = 1 where idg == idb
 * (:4455-4458).  Pass the handle to c2g_grid_format_text (layout 1) for the value block of the cube file, then
 * c2g_grid_free it.  Single-GPU contexts only for Bader labels (they are sharded otherwise). */
int c2g_basins_weight_grid(c2g_basins* res, int idb, int* handle);
/* statistics: [0] IAS points, [1] flux records (sum of nhi over IAS points), [2] sweep levels */
/* (use c2g_basins_stats) */

/* ---- ISOSURFACE: regions of f >= isov (yt_isosurface, yt@proc.f90:233-390) ---- */
/* yt: the result of c2g_yt_build on the field (its stencil, maxima order and sweep levels are reused).  Returns a new
 * c2g_basins whose labels are the reference's bas%idg: 0 below the contour value, else the region id -- regions that
 * touch are merged onto the smaller id through imap exactly like :319-351, INCLUDING the reference's behaviour that
 * a later contact overwrites imap(b) and that surviving regions keep their discovery numbers.  nraw = regions before
 * merging (ids run up to nraw), nattr = survivors (the reference's final bas%nattr); c2g_basins_maxima gives the
 * nraw regional maxima in discovery order (bas%xattr = its first nattr columns, :359).  The map is already set:
 * c2g_basins_labels, c2g_integrate (plain sums like the Bader branch, ids 1..nraw) and c2g_integrate_multipoles work on
 * the result.  A DISCARD expression (:304-311) is evaluated by the host parser and is not supported here. */
int c2g_yt_isosurface(c2g_basins* yt, double isov, int* nraw, int* nattr, c2g_basins** res);

/* ---- NCIPLOT: reduced density gradient loop (nci@proc.f90:543-605, grid mode) ---- */
/* x0(3) Cartesian origin, xmat(3,3) Cartesian step vectors, nstep(3); c2x/x2c crystal matrices,
 * c2xl grid matrix (grid3mod.f90 c2xl); nnuc nuclei (Cartesian, 3 x nnuc) for the zero-gradient-at-
 * nucleus rule (fieldmod@proc.f90:1148-1155).  Outputs crho, cgrad(0:nstep3-1,0:nstep2-1,0:nstep1-1),
 * third index fastest, exactly like the reference arrays (nci@proc.f90:598-599). */
int c2g_nci_rdg(c2g_context* ctx, int handle, const double x0[3], const double xmat[9], const int nstep[3],
                const double c2x[9], const double x2c[9], const double c2xl[9], int nnuc,
                const double* nuc_cart, double* crho, double* cgrad);
/* Multi-GPU (c2g_init_multi; the field must be resident on every rank, which c2g_grid_upload_slab ensures): the
 * output lattice is sharded along its slowest index i.  c2g_nci_range gives this rank's rows [ilo, ihi) (0-based);
 * every c2g_nci_rdg* call then computes and returns only cgrad(:,:,ilo+1:ihi) / crho(:,:,ilo+1:ihi), one contiguous
 * piece of the reference arrays.  No collective is involved (NCI shards naturally). */
int c2g_nci_range(c2g_context* ctx, int nstep1, int* ilo, int* ihi);
/* same, results stay in HBM (two new grid handles with n = (nstep3,nstep2,nstep1)) */
int c2g_nci_rdg_resident(c2g_context* ctx, int handle, const double x0[3], const double xmat[9],
                         const int nstep[3], const double c2x[9], const double x2c[9], const double c2xl[9],
                         int nnuc, const double* nuc_cart, int* hrho, int* hgrad);

/* ---- FFT-derived fields: grid3%fft (grid3mod@proc.f90:1757-1872; LOAD AS LAP/GRAD, the `lap`/`gmod`
 * integrables, NCIPLOT FOURIER mode) ---- */
/* iff = the reference's ifformat_as_ft_* codes (param.F90:225-236) */
#define C2G_FT_X 33
#define C2G_FT_Y 34
#define C2G_FT_Z 35
#define C2G_FT_XX 36
#define C2G_FT_XY 37
#define C2G_FT_XZ 38
#define C2G_FT_YY 39
#define C2G_FT_YZ 40
#define C2G_FT_ZZ 41
#define C2G_FT_GRAD 42 /* |grad f| */
#define C2G_FT_LAP 43
#define C2G_FT_POT 44
/* x2c(3,3): crystallographic -> Cartesian matrix of the grid (grid3%x2c).  The result is a new resident
 * grid of the same shape (fnew%f); download it with c2g_grid_download or feed it to c2g_integrate.
 * Single GPU.  Uses cuFFT for the transforms. */
int c2g_fft_derivative(c2g_context* ctx, int handle, int iff, const double x2c[9], int* hout);

/* NCIPLOT loop with FOURIER interpolation (nci@proc.f90:527-565): h = {rho, |grad rho|, Hxx, Hyy, Hzz}
 * resident grids (the derived ones from c2g_fft_derivative with C2G_FT_GRAD, _XX, _YY, _ZZ, :528-531);
 * rho is read with the tricubic interpolant, the derived grids TRILINEARLY (:534-537), each with the
 * node short-cut of field%grd (fieldmod@proc.f90:948-961).  Outputs as c2g_nci_rdg. */
int c2g_nci_rdg_fourier(c2g_context* ctx, const int h[5], const double x0[3], const double xmat[9],
                        const int nstep[3], const double c2x[9], const double c2xl[9], double* crho, double* cgrad);

/* ---- formatted-text grid readers: the numeric block of a cube file (read_cube,
 * grid3mod@proc.f90:512-568) or of a CHGCAR/CHG/ELFCAR file (read_vasp, :842-913) ---- */
#define C2G_TEXT_ORDER_I_FASTEST 0 /* (((f(i,j,k),i=1,n1),j=1,n2),k=1,n3): VASP */
#define C2G_TEXT_ORDER_K_FASTEST 1 /* (((f(i,j,k),k=1,n3),j=1,n2),i=1,n1): Gaussian cube */
/* text: the bytes that follow the header lines (host memory); the first n1*n2*n3 numbers are converted with
 * correct rounding (the value a list-directed READ gives), stored as f(n1,n2,n3) in a new resident grid and
 * divided by `divisor` (det3(x2c) for CHGCAR with vscal, :907-908; 1 otherwise).  consumed (optional) = offset of
 * the first byte after the last value read (the next CHGCAR block starts there); nslow (optional) = number of
 * values that took the exact multi-word path on the device (results within 2^-98 of a rounding boundary, subnormals,
 * more than 19 digits); nothing is converted on the host. */
int c2g_grid_parse_text(c2g_context* ctx, const char* text, size_t nbytes, const int n[3], int order, double divisor,
                        int* handle, size_t* consumed, long long* nslow);

/* Formatted output of a resident grid: the value loops of writegrid_cube (crystalmod@write.f90:3556-3565) and of
 * NCIPLOT's write_cube_body (nci@proc.f90:916-932): every value as " " + Ew.dE3 with scale factor `scale` (0 or 1),
 * 6 per line, a new line after each row; digits correctly rounded (nearest even), fields that do not fit are
 * asterisks like the Fortran run-time prints them.
 *   layout C2G_TEXT_ROWS_INDEX1: rows run along the first (fastest) index as stored -- the crho/cgrad(k,j,i) arrays of
 *     c2g_nci_rdg_resident with (6(" ",1p,e13.5e3)): width 13, digits 5, scale 1;
 *   layout C2G_TEXT_ROWS_INDEX3: cube order of a field f(i,j,k), rows along k for i outer, j inner, with the optional
 *     shift ishift(3) of writegrid_cube: (1p,6(" ",E12.5E3)) = 12, 5, 1 or precisecube (6(" ",E22.14E3)) = 22, 14, 0.
 * out == NULL or cap == 0: only *nbytes is set (step 1); otherwise out(cap) receives *nbytes characters (step 2). */
#define C2G_TEXT_ROWS_INDEX1 0
#define C2G_TEXT_ROWS_INDEX3 1
int c2g_grid_format_text(c2g_context* ctx, int handle, int layout, const int ishift[3], int width, int digits, int scale,
                         char* out, size_t cap, size_t* nbytes);

/* ---- profiling: CUDA-event timings of the kernels launched by the last API call ---- */
int c2g_profile_enable(c2g_context* ctx, int on);
int c2g_profile_count(c2g_context* ctx);
/* name (<=63 chars), milliseconds, number of launches accumulated under that name */
int c2g_profile_get(c2g_context* ctx, int i, char name[64], double* ms, int* launches);
int c2g_profile_reset(c2g_context* ctx);
/* total number of kernels launched by this context so far */
long long c2g_launch_count(c2g_context* ctx);
/* write a buffer larger than L2 (flush between timed iterations) */
int c2g_flush_l2(c2g_context* ctx);
int c2g_synchronize(c2g_context* ctx);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of this library is launched on):
 * elapsed device time between the two calls, including gaps while the host prepares the next launch. */
int c2g_timer_start(c2g_context* ctx);
int c2g_timer_stop(c2g_context* ctx, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* CRITIC2_GPU_H */
