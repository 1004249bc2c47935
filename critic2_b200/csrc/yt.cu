// yt.cu -- placeholder, filled in below
#include "common.cuh"
int c2g_yt_integrate_impl(c2g_context* ctx, c2g_basins*, int, const int*, double, double*, double*) {
  return ctx->fail(C2G_ERR_STATE, "YT not built");
}
