// Wavefront stages that do not traverse: camera ray generation, BSDF sampling / evaluation, light sampling,
// Russian roulette, accumulation.  (Traversal stages live in trace_kernels.cu.)
#include "kernels.h"
#include "rng.cuh"
#include "shade.cuh"

#include <algorithm>
#include <cfloat>
#include <vector>

namespace spcu {
namespace {

constexpr int kShadeBlock = 128;

inline unsigned grid_for(uint64_t n, int block = kShadeBlock)
{
    return static_cast<unsigned>((n + block - 1) / block);
}

__global__ void __launch_bounds__(kShadeBlock) k_generate_rays(const __grid_constant__ DScene s, const uint32_t* pix,
                                                               const uint32_t* smp, uint64_t n, spcu_ray* rays)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kShadeBlock + threadIdx.x;
    if (i >= n) {
        return;
    }
    float4 o, d;
    camera_ray(s, pix[i], smp[i], o, d);
    reinterpret_cast<float4*>(rays)[2 * i + 0] = o;
    reinterpret_cast<float4*>(rays)[2 * i + 1] = d;
}

// Morton decode of the low 6 bits: TilePixelIterator visits an 8x8 tile in Morton order (base/Tile.h:134-138,
// math/Morton.h:87-92): x = even bits, y = odd bits.
__device__ __forceinline__ void morton8(uint32_t m, uint32_t& x, uint32_t& y)
{
    x = (m & 1u) | ((m >> 1) & 2u) | ((m >> 2) & 4u);
    y = ((m >> 1) & 1u) | ((m >> 2) & 2u) | ((m >> 3) & 4u);
}

// Pixel list of a partition: tiles t = offset, offset+stride, ... in the scheduler's row-major tile order
// (base/TileScheduler.h:66-82), 64 Morton-ordered candidates per tile, those outside the image dropped
// (main.cpp:90-91).  Slot j of the list is computed directly: prefix[k] = pixels in the first k owned tiles is
// evaluated in closed form from the number of full / clipped tiles, so no scan is needed.
__global__ void k_build_pixel_list(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride,
                                   uint32_t n_owned_tiles, const uint32_t* tile_prefix, uint32_t* pix_list)
{
    const uint32_t k = blockIdx.x; // owned tile index
    if (k >= n_owned_tiles) {
        return;
    }
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t t       = tile_offset + k * tile_stride;
    const uint32_t tx = t % tiles_x, ty = t / tiles_x;
    uint32_t       mx, my;
    morton8(threadIdx.x, mx, my);
    const uint32_t x = tx * 8 + mx, y = ty * 8 + my;
    const bool     inside = x < width && y < height;
    // rank of this Morton index among the inside pixels of the tile
    __shared__ uint32_t flags[64];
    flags[threadIdx.x] = inside ? 1u : 0u;
    __syncthreads();
    uint32_t rank = 0;
    for (uint32_t m = 0; m < threadIdx.x; ++m) {
        rank += flags[m];
    }
    if (inside) {
        pix_list[tile_prefix[k] + rank] = y * width + x;
    }
}


// ---- queue helpers ---------------------------------------------------------------------------------------------------
// All wavefront kernels are persistent-style: a fixed grid (a multiple of the SM count) strides over the queue, whole
// warps iterate together so that ballots see all 32 lanes.  Compaction = warp ballot + ONE atomicAdd per warp.
__device__ __forceinline__ void warp_count(unsigned long long* counter, bool pred, unsigned per = 1u)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(counter, static_cast<unsigned long long>(__popc(m)) * per);
    }
}


// records are moved as pairs of 16-byte vectors
__device__ __forceinline__ Rng make_rng(const PathRec& pr, uint64_t seed, uint32_t depth, uint32_t site, uint32_t first_block)
{
    return Rng{ __float_as_uint(pr.tp.w),      __float_as_uint(pr.L.w), static_cast<uint32_t>(seed),
                static_cast<uint32_t>(seed >> 32), rng_stream(depth, site), first_block };
}

#define FOR_EACH_QUEUED(i, active, n)                                                                  \
    for (uint32_t base_ = blockIdx.x * blockDim.x, i = base_ + threadIdx.x, active = (i < (n)) ? 1u : 0u; \
         base_ < (n); base_ += gridDim.x * blockDim.x, i = base_ + threadIdx.x, active = (i < (n)) ? 1u : 0u)

// ---- raygen (main.cpp:90-98): slot = s_local * n_pix + p_local, so neighbouring threads trace neighbouring pixels -------
__global__ void __launch_bounds__(kShadeBlock) k_raygen(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                        const uint32_t* pix_list, uint32_t n_pix, uint32_t sample_begin,
                                                        uint32_t n_samples, uint32_t* queue, uint32_t* n_queue,
                                                        unsigned long long* counters)
{
    const uint32_t n = n_pix * n_samples;
    count_items(counters, kStRaygen, n);
    FOR_EACH_QUEUED(i, active, n)
    {
        if (active) {
            const uint32_t pix = __ldg(pix_list + i % n_pix);
            const uint32_t smp = sample_begin + i / n_pix;
            RayRec         ray;
            camera_ray(s, pix, smp, ray.o, ray.d);
            w.ray[i]      = ray;
            w.path[i]     = PathRec{ make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pix)), make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(smp)) };
            queue[i]      = i;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *n_queue = n;
        atomicAdd(counters + kCntPaths, static_cast<unsigned long long>(n));
    }
}

// ---- shade (Integrator.cpp:558-572, 627-632): miss handling, surface interaction, primary BSDF sample S0 ---------------
#ifndef SPCU_SHADE_MIN_BLOCKS
#define SPCU_SHADE_MIN_BLOCKS 8 // 64 registers instead of 80 (bunny: shade 23.3 -> 22.0 ms)
#endif
template <typename F>
__global__ void __launch_bounds__(kShadeBlock, SPCU_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                       const __grid_constant__ RenderParams p,
                                                       const __grid_constant__ SortedQueue sorted, uint32_t* q_live,
                                                       uint32_t* n_live, uint32_t* q_shadow, uint32_t* n_shadow,
                                                       unsigned long long* counters)
{
  // one material segment after the other: every warp shades a single material (or only misses)
  for (uint32_t seg = 0; seg < sorted.n_segments; ++seg) {
    const uint32_t  n    = sorted.counts[seg];
    const uint32_t* q_in = sorted.slots + static_cast<size_t>(seg) * sorted.capacity;
    count_items(counters, kStShade, n);
    FOR_EACH_QUEUED(i, active, n)
    {
        bool     live    = false;
        bool     sampled = false;
        uint32_t slot    = 0;
        V3       lit_point = v3(0, 0, 0), lit_normal = v3(0, 1, 0);
        if (active) {
            slot               = q_in[i];
            const ExtendRec ex = w.extend[slot];
            const RayRec    ry = w.ray[slot];
            const V3        o = xyz(ry.o), d = xyz(ry.d);
            if (ex.hit.id < 0) {
                if (ex.light >= 0) { // emitter reached, at ANY depth (Integrator.cpp:627-629)
                    const V3 L  = light_hit_L<F>(s, s.lights[ex.light], d);
                    PathRec  pr = w.path[slot];
                    pr.L.x += pr.tp.x * L.x;
                    pr.L.y += pr.tp.y * L.y;
                    pr.L.z += pr.tp.z * L.z;
                    w.path[slot].L = pr.L;
                }
            } else {
                V3       point, normal;
                uint32_t material;
                make_isect<F>(s, ex.hit, o, d, point, normal, material);
                if (p.integrator == SPCU_INTEGRATOR_DIRECT_LIGHTING || p.integrator == SPCU_INTEGRATOR_WHITTED) {
                    live = true; // these sample the lights first; Whitted draws its BSDF sample afterwards (whitted_advance)
                } else {
                    Rng           rng = make_rng(w.path[slot], p.seed, p.depth, kSiteBsdf, 0u);
                    const MSample sr  = material_sample<F>(s, material, -d, normal, rng);
                    sampled           = true;
                    if (!(sr.pdf == 0.0f || is_black(sr.color))) {
                        w.s0[slot] = SampleRec{ f4(sr.dir, sr.pdf), f4(sr.color, 0.0f) };
                        live       = true;
                    }
                }
                w.vertex[slot] = VertexRec{ f4(point, __uint_as_float(material)), f4(normal, 0.0f) };
                lit_point      = point;
                lit_normal     = normal;
            }
        }
        queue_push(q_live, n_live, slot, live);
        warp_count(counters + kCntShadeCalls, sampled);
        // Light::sample for EVERY light of the scene (Integrator.cpp:497-501 / :288-292), while point and normal are in
        // registers: each light has its own random sub-stream, so the lights of a vertex do not depend on each other.
        // Usable samples go to that light's shadow queue.
        if (q_shadow) {
            for (uint32_t li = 0; li < s.n_lights; ++li) {
                bool usable = false;
                if (live) {
                    const spcu_light& light = s.lights[__ldg(s.light_order + li)];
                    Rng               rng   = make_rng(w.path[slot], p.seed, p.depth, kSiteLight0 + li, 0u);
                    float             u0, u1;
                    rng_next2(rng, u0, u1);
                    const LSample ls = light_sample<F>(s, light, lit_point, lit_normal, u0, u1);
                    if (!(ls.pdf == 0.0f || is_black(ls.L))) {
                        w.light[static_cast<size_t>(li) * w.capacity + slot] =
                            LightRec{ f4(ls.wi, ls.t_max), make_float4(ls.t_min, ls.pdf, ls.u, ls.v) };
                        usable = true;
                    }
                }
                queue_push(q_shadow + static_cast<size_t>(li) * w.capacity, n_shadow + li, slot, usable);
            }
        }
    }
  }
}

// radiance of a stored light sample: an image-based light is looked up again from the sample's (u, v); every other
// light's sample carries the light's constant radiance (Lights/Light.h:81-90,155-161,226-249)
template <typename F>
__device__ __forceinline__ V3 light_sample_L(const DScene& s, const spcu_light& light, const LightRec& lr)
{
    return F::ibl && light.kind == SPCU_LIGHT_ENV_IBL ? ibl_lookup(s, light, lr.aux.z, lr.aux.w)
                                            : v3(light.radiance[0], light.radiance[1], light.radiance[2]);
}

// ---- nee_bsdf (Integrator.cpp:503-530): light-strategy term, second BSDF sample, Light::pdf ----------------------------------
#ifndef SPCU_NEE_MIN_BLOCKS
#define SPCU_NEE_MIN_BLOCKS 8 // 64 registers instead of 122: 32 warps per SM instead of 16 (bunny: nee_bsdf 37.2 -> 32.7 ms)
#endif
template <typename F>
__global__ void __launch_bounds__(kShadeBlock, SPCU_NEE_MIN_BLOCKS) k_nee_bsdf(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                          const __grid_constant__ RenderParams p, const uint32_t* q_shadow,
                                                          const uint32_t* n_shadow, uint32_t* q_mis, uint32_t* n_mis,
                                                          unsigned long long* counters)
{
    const uint32_t    n     = *n_shadow;
    count_items(counters, kStNeeBsdf, n);
    const spcu_light& light = s.lights[__ldg(s.light_order + p.light_index)];
    FOR_EACH_QUEUED(i, active, n)
    {
        bool     to_mis = false;
        uint32_t slot   = 0;
        unsigned calls  = 0;
        if (active) {
            slot = q_shadow[i]; // the shadow stage queued unoccluded samples only (Integrator.cpp:503-506)
            const VertexRec vx       = w.vertex[slot];
            const LightRec  lr       = w.light[static_cast<size_t>(p.light_index) * w.capacity + slot];
            const PathRec   pr       = w.path[slot];
            const V3        pt       = xyz(vx.p);
            const uint32_t  material = __float_as_uint(vx.p.w);
            const V3        nn       = xyz(vx.n);
            const V3        wo       = -xyz(w.ray[slot].d);
            const V3        wi       = xyz(lr.wi);
            const float     lpdf_s   = lr.aux.y;
            const V3        lL       = light_sample_L<F>(s, light, lr);
            Rng             rng      = make_rng(pr, p.seed, p.depth, kSiteLight0 + p.light_index, 1u); // draw 0 was the light sample

            // Material::eval / pdf (materials/Material.h:475-490) rebuild the ONB on every call; it is the same basis
            const Onb onb = onb_from_v(nn);
            const V3  wol = to_onb(onb, wo), wil = to_onb(onb, wi);
            V3        A   = v3(0, 0, 0);
            const Coats<F> coats = walk_coats<F>(s, material, wol);
            const V3  f   = material_eval_coats<F>(s, coats, wol, wil, rng);
            ++calls;
            if (!is_black(f)) {
                const float bsdf_pdf = material_pdf_coats<F>(s, coats, wol, wil, rng);
                ++calls;
                if (bsdf_pdf > 0.0f) {
                    const float weight = balance2(lpdf_s, bsdf_pdf);
                    A                  = f * lL * (fabsf(dot(wi, nn)) * weight / lpdf_s);
                }
            }
            MSample ms = material_sample_local<F>(s, material, wol, rng);
            ++calls;
            float lpdf = 0.0f;
            if (!(ms.pdf == 0.0f || is_black(ms.color))) {
                ms.dir = to_world(onb, ms.dir);
                lpdf   = light_pdf<F>(s, light, pt, ms.dir);
            }
            if (lpdf != 0.0f) {
                const float weight = balance2(ms.pdf, lpdf);
                MisRec      mr;
                mr.d        = f4(ms.dir, ray_offset(nn, ms.dir));
                mr.col      = f4(ms.color, ms.pdf);
                mr.cwa      = make_float4(fabsf(dot(ms.dir, nn)), weight, A.x, A.y);
                mr.ab       = A.z;
                mr.light    = -1;
                mr.occluded = 0;
                mr.pad      = 0.0f;
                w.mis[slot] = mr;
                to_mis      = true;
            } else if (!is_black(A)) {
                float4 L = pr.L;
                L.x += pr.tp.x * A.x;
                L.y += pr.tp.y * A.y;
                L.z += pr.tp.z * A.z;
                w.path[slot].L = L;
            }
        }
        queue_push(q_mis, n_mis, slot, to_mis);
        const unsigned total = __reduce_add_sync(0xffffffffu, calls);
        if ((threadIdx.x & 31) == 0 && total) {
            atomicAdd(counters + kCntShadeCalls, static_cast<unsigned long long>(total));
        }
    }
}

// ---- nee_mis_accumulate (Integrator.cpp:531-538): L += throughput * (light term + BSDF-strategy term) ------------------------
template <typename F>
__global__ void __launch_bounds__(kShadeBlock) k_nee_mis_accumulate(const __grid_constant__ DScene s,
                                                                    const __grid_constant__ DWave w, const uint32_t* q_mis,
                                                                    const uint32_t* n_mis, unsigned long long* counters)
{
    const uint32_t n = *n_mis;
    count_items(counters, kStNeeMisAccumulate, n);
    FOR_EACH_QUEUED(i, active, n)
    {
        if (active) {
            const uint32_t slot = q_mis[i];
            const MisRec   mr   = w.mis[slot];
            V3             est  = v3(mr.cwa.z, mr.cwa.w, mr.ab);
            if (mr.light >= 0 && mr.occluded == 0) {
                const V3 Li = light_hit_L<F>(s, s.lights[mr.light], xyz(mr.d));
                est         = est + xyz(mr.col) * Li * mr.cwa.x * mr.cwa.y / mr.col.w;
            }
            if (!is_black(est)) {
                const PathRec pr = w.path[slot];
                float4        L  = pr.L;
                L.x += pr.tp.x * est.x;
                L.y += pr.tp.y * est.y;
                L.z += pr.tp.z * est.z;
                w.path[slot].L = L;
            }
        }
    }
}

// ---- direct_accumulate (Integrator.cpp:296-306): DirectLightingIntegrator's per-light term -----------------------------------
// Order in the reference: eval first, shadow query only when f != black.  Here the shadow query has already run for every
// usable light sample (its result does not change the estimate, only the ray count), then eval consumes its random numbers.
template <typename F>
__global__ void __launch_bounds__(kShadeBlock) k_direct_accumulate(const __grid_constant__ DScene s,
                                                                   const __grid_constant__ DWave w,
                                                                   const __grid_constant__ RenderParams p,
                                                                   const uint32_t* q_shadow, const uint32_t* n_shadow,
                                                                   unsigned long long* counters)
{
    const uint32_t    n     = *n_shadow;
    count_items(counters, kStDirectAccumulate, n);
    const spcu_light& light = s.lights[__ldg(s.light_order + p.light_index)];
    FOR_EACH_QUEUED(i, active, n)
    {
        if (active) {
            const uint32_t  slot     = q_shadow[i];
            const VertexRec vx       = w.vertex[slot];
            const LightRec  lr       = w.light[static_cast<size_t>(p.light_index) * w.capacity + slot];
            const uint32_t  material = __float_as_uint(vx.p.w);
            const V3        nn       = xyz(vx.n);
            const V3        wo       = -xyz(w.ray[slot].d);
            const V3        wi       = xyz(lr.wi);
            Rng             rng      = make_rng(w.path[slot], p.seed, p.depth, kSiteLight0 + p.light_index, 1u);
            const Onb       onb      = onb_from_v(nn);
            const V3        f        = material_eval_local<F>(s, material, to_onb(onb, wo), to_onb(onb, wi), rng);
            if (!is_black(f) && !w.occluded[slot]) {
                const V3 c = f * light_sample_L<F>(s, light, lr) * fabsf(dot(wi, nn)) / lr.aux.y;
                float4   L = w.path[slot].L;
                L.x += c.x;
                L.y += c.y;
                L.z += c.z;
                w.path[slot].L = L;
            }
        }
        warp_count(counters + kCntShadeCalls, active != 0u);
    }
}

// ---- advance (Integrator.cpp:601-626): throughput update, Russian roulette, next segment -----------------------------------
__global__ void __launch_bounds__(kShadeBlock) k_advance(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                         const __grid_constant__ RenderParams p, const uint32_t* q_live,
                                                         const uint32_t* n_live, uint32_t* q_next, uint32_t* n_next,
                                                         unsigned long long* counters)
{
    const uint32_t n = *n_live;
    count_items(counters, kStAdvance, n);
    FOR_EACH_QUEUED(i, active, n)
    {
        bool     alive = false;
        uint32_t slot  = 0;
        if (active) {
            slot                = q_live[i];
            const SampleRec s0  = w.s0[slot];
            const VertexRec vx  = w.vertex[slot];
            const PathRec   pr  = w.path[slot];
            const V3        wi  = xyz(s0.dir);
            const V3        nn  = xyz(vx.n);
            const float     cosine = fabsf(dot(wi, nn));
            V3              tp  = xyz(pr.tp) * (cosine * xyz(s0.col) / s0.dir.w);
            alive               = true;
            if (p.depth >= s.rr_depth) {
                const float lum = luminance(tp);
                if (lum < 0.1f) {
                    const float q   = max_std(0.05f, lum / 0.1f); // probability of continuing
                    Rng         rng = make_rng(pr, p.seed, p.depth, kSiteRoulette, 0u);
                    const float u   = rng_next1(rng);
                    if (u < q) {
                        tp = tp / q;
                    } else {
                        alive = false;
                    }
                }
            }
            if (alive) {
                w.path[slot].tp = f4(tp, pr.tp.w);
                w.ray[slot]     = RayRec{ make_float4(vx.p.x, vx.p.y, vx.p.z, ray_offset_cos(cosine)), f4(wi, kFltMax) };
            }
        }
        queue_push(q_next, n_next, slot, alive);
    }
}

// ---- whitted_advance (Integrator.cpp:357-363): follow the BSDF sample only when it is specular; the reflected ray gets
// default RayLimits and its radiance is added unweighted (the reference's `L += do_integrate(outgoing_ray, ...)`) ------------
template <typename F>
__global__ void __launch_bounds__(kShadeBlock) k_whitted_advance(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                                 const __grid_constant__ RenderParams p, const uint32_t* q_live,
                                                                 const uint32_t* n_live, uint32_t* q_next, uint32_t* n_next,
                                                                 unsigned long long* counters)
{
    const uint32_t n = *n_live;
    count_items(counters, kStAdvance, n);
    FOR_EACH_QUEUED(i, active, n)
    {
        bool     alive = false;
        uint32_t slot  = 0;
        if (active) {
            slot               = q_live[i];
            const VertexRec vx = w.vertex[slot];
            Rng             rng = make_rng(w.path[slot], p.seed, p.depth, kSiteBsdf, 0u);
            const MSample   ms  = material_sample<F>(s, __float_as_uint(vx.p.w), -xyz(w.ray[slot].d), xyz(vx.n), rng);
            if (ms.specular) { // is_specular(properties) alone decides (:359); a black or pdf-0 specular sample is followed too
                w.ray[slot] = RayRec{ make_float4(vx.p.x, vx.p.y, vx.p.z, kEps), f4(ms.dir, kFltMax) };
                alive       = true;
            }
        }
        queue_push(q_next, n_next, slot, alive);
        warp_count(counters + kCntShadeCalls, active != 0u);
    }
}

// ---- resolve (main.cpp:100): per pixel, add this batch's samples in sample order ------------------------------------------
// radiance[(k * n_pix + i) * stride] is the sample of pixel i, sample k: stride 2 / offset 1 reads PathRec::L of the
// wavefront state, stride 1 the persistent path kernel's buffer.
__global__ void __launch_bounds__(kShadeBlock) k_resolve(const float4* radiance, uint32_t stride, const uint32_t* pix_list,
                                                         uint32_t n_pix, uint32_t n_samples, float* rgb_sum, float* lum_sumsq,
                                                         unsigned long long* counters)
{
    count_items(counters, kStResolve, n_pix * n_samples);
    FOR_EACH_QUEUED(i, active, n_pix)
    {
        if (active) {
            const uint32_t pix = __ldg(pix_list + i);
            float r = rgb_sum[3 * pix + 0], g = rgb_sum[3 * pix + 1], b = rgb_sum[3 * pix + 2];
            float sq = lum_sumsq ? lum_sumsq[pix] : 0.0f;
            for (uint32_t k = 0; k < n_samples; ++k) {
                const float4 L = radiance[(static_cast<size_t>(k) * n_pix + i) * stride];
                r += L.x;
                g += L.y;
                b += L.z;
                const float lum = luminance(v3(L.x, L.y, L.z));
                sq += lum * lum;
            }
            rgb_sum[3 * pix + 0] = r;
            rgb_sum[3 * pix + 1] = g;
            rgb_sum[3 * pix + 2] = b;
            if (lum_sumsq) {
                lum_sumsq[pix] = sq;
            }
        }
    }
}

} // namespace

unsigned wavefront_grid(uint32_t max_n, int block, int ctas_per_sm, int sm_count)
{
    const uint64_t need = (static_cast<uint64_t>(max_n) + block - 1) / block;
    const uint64_t fill = static_cast<uint64_t>(std::max(1, ctas_per_sm)) * std::max(1, sm_count);
    return static_cast<unsigned>(std::max<uint64_t>(1, std::min(need, fill)));
}

template <typename K>
static int ctas_per_sm(K kernel, int block)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, 0) != cudaSuccess || n < 1) {
        n = 1;
    }
    return n;
}

#define WAVEFRONT_LAUNCH(kernel, l, max_n, ...)                                                       \
    do {                                                                                              \
        if ((max_n) == 0) return;                                                                     \
        static const int occ_ = ctas_per_sm(kernel, kShadeBlock);                                     \
        kernel<<<wavefront_grid((max_n), kShadeBlock, occ_, (l).sm_count), kShadeBlock, 0, (l).stream>>>(__VA_ARGS__); \
    } while (0)

// shading kernels are instantiated per scene feature set (features.h); the launcher picks the one the scene needs
#define FEATURE_LAUNCH(kernel, l, max_n, ...)                                \
    do {                                                                     \
        if ((l).features == FeatAnalytic::id) {                              \
            WAVEFRONT_LAUNCH(kernel<FeatAnalytic>, l, max_n, __VA_ARGS__);   \
        } else {                                                             \
            WAVEFRONT_LAUNCH(kernel<FeatFull>, l, max_n, __VA_ARGS__);       \
        }                                                                    \
    } while (0)

void launch_raygen(const Launch& l, const DScene& s, const DWave& w, const uint32_t* d_pix_list, uint32_t n_pix,
                   uint32_t sample_begin, uint32_t n_samples, uint32_t* queue, uint32_t* d_n_queue,
                   unsigned long long* d_counters)
{
    WAVEFRONT_LAUNCH(k_raygen, l, n_pix * n_samples, s, w, d_pix_list, n_pix, sample_begin, n_samples, queue, d_n_queue,
                     d_counters);
}

void launch_shade(const Launch& l, const DScene& s, const DWave& w, const RenderParams& p, const SortedQueue& sorted,
                  uint32_t max_n, uint32_t* q_live, uint32_t* d_n_live, uint32_t* q_shadow, uint32_t* d_n_shadow,
                  unsigned long long* d_counters)
{
    FEATURE_LAUNCH(k_shade, l, max_n, s, w, p, sorted, q_live, d_n_live, q_shadow, d_n_shadow, d_counters);
}

void launch_nee_bsdf(const Launch& l, const DScene& s, const DWave& w, const RenderParams& p, const uint32_t* q_shadow,
                     const uint32_t* d_n_shadow, uint32_t max_n, uint32_t* q_mis, uint32_t* d_n_mis,
                     unsigned long long* d_counters)
{
    FEATURE_LAUNCH(k_nee_bsdf, l, max_n, s, w, p, q_shadow, d_n_shadow, q_mis, d_n_mis, d_counters);
}

void launch_nee_mis_accumulate(const Launch& l, const DScene& s, const DWave& w, const uint32_t* q_mis,
                               const uint32_t* d_n_mis, uint32_t max_n, unsigned long long* d_counters)
{
    FEATURE_LAUNCH(k_nee_mis_accumulate, l, max_n, s, w, q_mis, d_n_mis, d_counters);
}

void launch_direct_accumulate(const Launch& l, const DScene& s, const DWave& w, const RenderParams& p,
                              const uint32_t* q_shadow, const uint32_t* d_n_shadow, uint32_t max_n,
                              unsigned long long* d_counters)
{
    FEATURE_LAUNCH(k_direct_accumulate, l, max_n, s, w, p, q_shadow, d_n_shadow, d_counters);
}

void launch_advance(const Launch& l, const DScene& s, const DWave& w, const RenderParams& p, const uint32_t* q_live,
                    const uint32_t* d_n_live, uint32_t max_n, uint32_t* q_next, uint32_t* d_n_next,
                    unsigned long long* d_counters)
{
    WAVEFRONT_LAUNCH(k_advance, l, max_n, s, w, p, q_live, d_n_live, q_next, d_n_next, d_counters);
}

void launch_whitted_advance(const Launch& l, const DScene& s, const DWave& w, const RenderParams& p, const uint32_t* q_live,
                            const uint32_t* d_n_live, uint32_t max_n, uint32_t* q_next, uint32_t* d_n_next,
                            unsigned long long* d_counters)
{
    FEATURE_LAUNCH(k_whitted_advance, l, max_n, s, w, p, q_live, d_n_live, q_next, d_n_next, d_counters);
}

void launch_resolve(const Launch& l, const float4* d_radiance, uint32_t stride, const uint32_t* d_pix_list, uint32_t n_pix,
                    uint32_t n_samples, float* d_rgb_sum, float* d_lum_sumsq, unsigned long long* d_counters)
{
    WAVEFRONT_LAUNCH(k_resolve, l, n_pix, d_radiance, stride, d_pix_list, n_pix, n_samples, d_rgb_sum, d_lum_sumsq, d_counters);
}

void launch_generate_rays(const DScene& s, const uint32_t* d_pix, const uint32_t* d_smp, uint64_t n, spcu_ray* d_rays,
                          cudaStream_t st)
{
    if (n == 0) return;
    k_generate_rays<<<grid_for(n), kShadeBlock, 0, st>>>(s, d_pix, d_smp, n, d_rays);
}

static uint32_t tile_pixels(uint32_t width, uint32_t height, uint32_t t)
{
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t tx = t % tiles_x, ty = t / tiles_x;
    const uint32_t w = std::min(8u, width - tx * 8), h = std::min(8u, height - ty * 8);
    return w * h;
}

uint32_t count_partition_pixels(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride)
{
    const uint32_t n_tiles = ((width + 7) / 8) * ((height + 7) / 8);
    uint64_t       n       = 0;
    for (uint32_t t = tile_offset; t < n_tiles; t += tile_stride) {
        n += tile_pixels(width, height, t);
    }
    return static_cast<uint32_t>(n);
}

cudaError_t build_pixel_list(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride, uint32_t* d_pix_list,
                             uint32_t* d_tile_prefix_scratch, cudaStream_t st)
{
    const uint32_t        n_tiles = ((width + 7) / 8) * ((height + 7) / 8);
    std::vector<uint32_t> prefix;
    uint32_t              run = 0;
    for (uint32_t t = tile_offset; t < n_tiles; t += tile_stride) {
        prefix.push_back(run);
        run += tile_pixels(width, height, t);
    }
    if (prefix.empty()) {
        return cudaSuccess;
    }
    cudaError_t e = cudaMemcpyAsync(d_tile_prefix_scratch, prefix.data(), prefix.size() * sizeof(uint32_t),
                                    cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        return e;
    }
    k_build_pixel_list<<<static_cast<unsigned>(prefix.size()), 64, 0, st>>>(width, height, tile_offset, tile_stride,
                                                                           static_cast<uint32_t>(prefix.size()),
                                                                           d_tile_prefix_scratch, d_pix_list);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        return e;
    }
    return cudaStreamSynchronize(st); // `prefix` must outlive the copy
}

} // namespace spcu
