// spcu_render / spcu_render_device: the wavefront loop (host orchestration).
#include "ctx.h"

using namespace spcu;

extern "C" {

int spcu_render_device(spcu_ctx* c, const spcu_partition* part, float* d_rgb_sum, float* d_lum_sumsq, spcu_stats* stats,
                       void* stream)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    return fail(c, SPCU_ERR_INVALID, "render stage not linked into this build");
}

int spcu_render(spcu_ctx* c, const spcu_partition* part, float* rgb_sum, float* lum_sumsq, spcu_stats* stats)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    return fail(c, SPCU_ERR_INVALID, "render stage not linked into this build");
}

int spcu_trace_closest_fast(spcu_ctx* c, const spcu_ray* rays, uint64_t n, spcu_hit* hits)
{
    return spcu_trace_closest(c, rays, n, hits);
}
}
