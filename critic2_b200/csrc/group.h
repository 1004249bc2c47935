// group.h -- internal interface of the single-process multi-device context (group.cu)
#pragma once
#include <functional>

struct c2g_context;
struct c2g_basins;

int c2g_group_run(c2g_context* g, const std::function<int(int, c2g_context*)>& f);
int c2g_group_size(const c2g_context* g);
c2g_context* c2g_group_sub(c2g_context* g, int r);
void c2g_group_finalize(c2g_context* g);

int grp_grid_alloc(c2g_context* g, const int n[3], int* handle);
int grp_grid_upload(c2g_context* g, const double* f, const int n[3], int* handle);
int grp_grid_download(c2g_context* g, int handle, double* f);
int grp_grid_free(c2g_context* g, int handle);
int grp_grid_promolecular(c2g_context* g, int handle, const double x2c[9], int nat, const double* xat, const double* zat,
                          const double* alpha, int nimg, double rc);
int grp_bader_assign(c2g_context* g, int handle, const double car2lat[9], const double lid[27], int algo, int order, int* nmax,
                     c2g_basins** res);
void grp_basins_free(c2g_basins* res);
int grp_basins_counts(c2g_basins* res, long long* counts);
int grp_basins_set_map(c2g_basins* res, int nattr, const int* map, bool relabel, int nattr0);
int grp_basins_labels(c2g_basins* res, int* idg);
int grp_integrate(c2g_context* g, c2g_basins* res, int nprop, const int* fieldhandles, double omega, double* psum, double* vol);
int grp_integrate_multipoles(c2g_context* g, c2g_basins* res, int fieldhandle, int lmax, const double* xattr, const unsigned char* domask,
                             int isortho, int isortho_del, const double x2c[9], const double x2xr[9], const double xr2c[9], int nws,
                             const double* ws, double omega, double* mpole);
int grp_basins_remap(c2g_context* g, c2g_basins* res, const double* xattr, const double c2x[9], int isortho, int isortho_del,
                     const double x2c[9], const double x2xr[9], const double xr2c[9], int nws, const double* ws, int maxattn, int* nattn,
                     int* iatt, int* ilvec, int* idg1);
int grp_nci_rdg(c2g_context* g, int handle, const double x0[3], const double xmat[9], const int nstep[3], const double c2x[9],
                const double x2c[9], const double c2xl[9], int nnuc, const double* nuc, double* crho, double* cgrad);
int grp_synchronize(c2g_context* g);
int grp_timer_start(c2g_context* g);
int grp_timer_stop(c2g_context* g, double* ms);
int grp_profile_enable(c2g_context* g, int on);
int grp_profile_reset(c2g_context* g);
int grp_profile_get(c2g_context* g, int i, char name[64], double* ms, int* launches);
long long grp_launch_count(c2g_context* g);
int grp_nci_rdg_fourier(c2g_context* g, const int h[5], const double x0[3], const double xmat[9], const int nstep[3], const double c2x[9],
                        const double c2xl[9], double* crho, double* cgrad);
int grp_fft_derivative(c2g_context* g, int handle, int iff, const double x2c[9], int* hout);
// entry points that do not shard: run on the first device of a multi-device context (its grids are complete)
#define C2G_FIRST_DEVICE(ctx)                          \
  do {                                                 \
    if ((ctx) && (ctx)->group) {                       \
      (ctx) = c2g_group_sub((ctx), 0);                 \
      cudaSetDevice((ctx)->device);                    \
    }                                                  \
  } while (0)
#define C2G_NOT_ON_GROUP(ctx, who)                                                                                    \
  do {                                                                                                                \
    if ((ctx) && (ctx)->group)                                                                                        \
      return (ctx)->fail(C2G_ERR_STATE, who ": not available on a multi-device context; use a single-device context"); \
  } while (0)
