// The persistent path kernel: ONE launch renders a whole batch of camera samples, each thread carrying a path from the
// camera to its end in registers and fetching a new one the moment it finishes ("path regeneration"), so warps stay
// full although path lengths vary between 1 and max_depth vertices.
//
// Same stages, same order, same random numbers as the wavefront pipeline (spcu_render.cu) — raygen, extend
// (intersect_lights + intersect), BSDF sample, per-light NEE (light sample, shadow ray, eval/pdf, BSDF-strategy ray),
// Russian roulette — but nothing round-trips through HBM between them: per path the kernel reads 4 bytes (its pixel)
// and writes 16 (its radiance sample), against ~1.3 KB per path VERTEX for the queue-based pipeline.
// Traversal is trace.cuh: reference order, exact arithmetic (explicit *_rn intrinsics, immune to FMA contraction), with
// the first levels of the stack in shared memory.
//
// Work distribution: the batch's slots are cut into one contiguous range per CTA (coherent camera rays, balanced to
// ~1 % by the law of large numbers); lanes draw slots from their CTA's shared-memory counter with ONE atomic per warp
// and refill (warp ballot + popc prefix).  A lane whose CTA range is exhausted idles until the CTA's last path ends.
#include "kernels.h"
#include "shade.cuh"
#include "trace.cuh"

namespace spcu {
namespace {

constexpr int kPathBlock = 128; // == kTraceBlock: the traversal stack layout is [level][thread]
static_assert(kPathBlock == kTraceBlock, "stack layout");

struct PathCounters
{
    unsigned paths = 0, rays_closest = 0, rays_any = 0, rays_lights = 0, shade_calls = 0;
};

__device__ __forceinline__ Ray make_ray(V3 o, V3 d, float t_min) { return Ray{ o.x, o.y, o.z, d.x, d.y, d.z, t_min }; }

// estimate_direct_mis (Integrators/Integrator.cpp:486-539) for one light
template <bool kCount, typename F>
__device__ __forceinline__ V3 estimate_direct_mis(const DScene& s, const spcu_light& light, V3 p, V3 n, uint32_t material,
                                                  V3 wo, Rng& rng, int32_t* stack, PathCounters& pc, TraceCounters* tc)
{
    V3    L = v3(0, 0, 0);
    float u0, u1;
    rng_next2(rng, u0, u1); // block 0 of the light's sub-stream; eval / pdf / sample follow from block 1
    const LSample ls = light_sample<F>(s, light, p, n, u0, u1);
    if (ls.pdf == 0.0f || is_black(ls.L)) {
        return L;
    }
    ++pc.rays_any;
    if (scene_any_hit<kCount, F>(s, make_ray(p, ls.wi, ls.t_min), ls.t_max, stack, tc)) {
        return L;
    }
    const Onb onb = onb_from_v(n);
    const V3  wol = to_onb(onb, wo), wil = to_onb(onb, ls.wi);
    const Coats<F> coats = walk_coats<F>(s, material, wol);
    const V3  f   = material_eval_coats<F>(s, coats, wol, wil, rng);
    ++pc.shade_calls;
    if (!is_black(f)) {
        const float bsdf_pdf = material_pdf_coats<F>(s, coats, wol, wil, rng);
        ++pc.shade_calls;
        if (bsdf_pdf > 0.0f) {
            const float weight = balance2(ls.pdf, bsdf_pdf);
            L                  = f * ls.L * (fabsf(dot(ls.wi, n)) * weight / ls.pdf);
        }
    }
    MSample ms = material_sample_local<F>(s, material, wol, rng);
    ++pc.shade_calls;
    if (ms.pdf == 0.0f || is_black(ms.color)) {
        return L;
    }
    ms.dir           = to_world(onb, ms.dir);
    const float lpdf = light_pdf<F>(s, light, p, ms.dir);
    if (lpdf == 0.0f) {
        return L;
    }
    const float weight = balance2(ms.pdf, lpdf);
    const Ray   mr     = make_ray(p, ms.dir, ray_offset(n, ms.dir));
    float       t_max  = kInfinite, beta, gamma;
    ++pc.rays_lights;
    const LightPrimsT<F> lp{ s.lights };
    const int32_t    li = closest_hit<false>(s.lights_accel, lp, mr, t_max, beta, gamma, stack, nullptr);
    if (li >= 0) {
        ++pc.rays_any;
        // limits are NOT shrunk to the light's distance: a sphere light occludes itself, as in the reference (:531-532)
        if (!scene_any_hit<kCount, F>(s, mr, kInfinite, stack, tc)) {
            const V3 Li = light_hit_L<F>(s, s.lights[li], ms.dir);
            L           = L + ms.color * Li * fabsf(dot(ms.dir, n)) * weight / ms.pdf;
        }
    }
    return L;
}

struct PathState
{
    V3       o, d;
    float    t_min;
    V3       throughput, L;
    uint32_t depth;
    uint32_t slot;
    Rng      rng;
};

// One path vertex of IntegratorIterativeRRNEE::do_integrate (Integrator.cpp:556-632), BruteForceIntegratorIterativeRR
// (:219-263, nee == false), DirectLightingIntegrator (:277-312) or WhittedIntegrator (:323-368).  Returns false when the
// path ends at this vertex.
template <bool kCount, typename F>
__device__ __forceinline__ bool path_vertex(const DScene& s, uint32_t integrator, PathState& ps, int32_t* stack,
                                            PathCounters& pc, TraceCounters* tc)
{
    const Ray r     = make_ray(ps.o, ps.d, ps.t_min);
    float     t_max = kInfinite, beta, gamma;

    ++pc.rays_lights;
    const LightPrimsT<F> lp{ s.lights };
    const int32_t    li = closest_hit<false>(s.lights_accel, lp, r, t_max, beta, gamma, stack, nullptr);
    ++pc.rays_closest;
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    const int32_t   gi = closest_hit<kCount>(s.geom, gp, r, t_max, beta, gamma, stack, tc);
    if (gi < 0) {
        if (li >= 0) {
            ps.L = ps.L + ps.throughput * light_hit_L<F>(s, s.lights[li], ps.d);
        }
        return false;
    }
    V3       point, normal;
    uint32_t material;
    make_isect<F>(s, HitRec{ gi, t_max, beta, gamma }, ps.o, ps.d, point, normal, material);
    const V3 wo = -ps.d;

    if (integrator == SPCU_INTEGRATOR_DIRECT_LIGHTING || integrator == SPCU_INTEGRATOR_WHITTED) {
        for (uint32_t k = 0; k < s.n_lights; ++k) {
            const spcu_light& light = s.lights[__ldg(s.light_order + k)];
            float             u0, u1;
            rng_seek(ps.rng, rng_stream(ps.depth, kSiteLight0 + k), 0u);
            rng_next2(ps.rng, u0, u1);
            const LSample ls = light_sample<F>(s, light, point, normal, u0, u1);
            if (ls.pdf == 0.0f || is_black(ls.L)) {
                continue;
            }
            const Onb onb = onb_from_v(normal);
            const V3  f   = material_eval_local<F>(s, material, to_onb(onb, wo), to_onb(onb, ls.wi), ps.rng);
            ++pc.shade_calls;
            if (is_black(f)) {
                continue;
            }
            ++pc.rays_any;
            if (!scene_any_hit<kCount, F>(s, make_ray(point, ls.wi, ls.t_min), ls.t_max, stack, tc)) {
                ps.L = ps.L + f * ls.L * fabsf(dot(ls.wi, normal)) / ls.pdf;
            }
        }
        if (integrator == SPCU_INTEGRATOR_DIRECT_LIGHTING) {
            return false;
        }
        // WhittedIntegrator (Integrator.cpp:357-363): follow the BSDF sample only when it is specular, default limits,
        // radiance of the reflected ray added unweighted
        rng_seek(ps.rng, rng_stream(ps.depth, kSiteBsdf), 0u);
        const MSample ms = material_sample<F>(s, material, wo, normal, ps.rng);
        ++pc.shade_calls;
        if (!ms.specular) { // is_specular(properties) alone decides (:359)
            return false;
        }
        ps.o     = point;
        ps.d     = ms.dir;
        ps.t_min = kRayEpsilon;
        ++ps.depth;
        return ps.depth < s.max_depth;
    }

    rng_seek(ps.rng, rng_stream(ps.depth, kSiteBsdf), 0u);
    const MSample sr = material_sample<F>(s, material, wo, normal, ps.rng);
    ++pc.shade_calls;
    if (sr.pdf == 0.0f || is_black(sr.color)) {
        return false;
    }
    if (integrator == SPCU_INTEGRATOR_ITERATIVE_RRNEE) {
        for (uint32_t k = 0; k < s.n_lights; ++k) {
            const spcu_light& light = s.lights[__ldg(s.light_order + k)];
            rng_seek(ps.rng, rng_stream(ps.depth, kSiteLight0 + k), 0u);
            ps.L = ps.L + ps.throughput * estimate_direct_mis<kCount, F>(s, light, point, normal, material, wo, ps.rng, stack, pc, tc);
        }
    }
    const float cosine = fabsf(dot(sr.dir, normal));
    ps.throughput      = ps.throughput * (cosine * sr.color / sr.pdf);
    if (ps.depth >= s.rr_depth) {
        const float lum = luminance(ps.throughput);
        if (lum < 0.1f) {
            const float q = max_std(0.05f, lum / 0.1f);
            rng_seek(ps.rng, rng_stream(ps.depth, kSiteRoulette), 0u);
            if (rng_next1(ps.rng) < q) {
                ps.throughput = ps.throughput / q;
            } else {
                return false;
            }
        }
    }
    ps.o     = point;
    ps.d     = sr.dir;
    ps.t_min = ray_offset_cos(cosine);
    ++ps.depth;
    return ps.depth < s.max_depth;
}

template <bool kCount, typename F>
__global__ void __launch_bounds__(kPathBlock) k_paths(const __grid_constant__ DScene s, const uint32_t* __restrict__ pix_list,
                                                      uint32_t n_pix, uint32_t sample_begin, uint32_t n_samples, uint64_t seed,
                                                      uint32_t integrator, float4* __restrict__ radiance,
                                                      unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t  stack_smem[kStackShared * kPathBlock];
    __shared__ uint32_t next_slot;

    const uint32_t n       = n_pix * n_samples;
    const uint32_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t begin   = min(n, blockIdx.x * per_cta);
    const uint32_t end     = min(n, begin + per_cta);
    if (threadIdx.x == 0) {
        next_slot = begin;
    }
    __syncthreads();

    int32_t*      stack = stack_smem + threadIdx.x;
    const int     lane  = threadIdx.x & 31;
    PathCounters  pc;
    TraceCounters tc{ 0, 0, 0 };
    PathState     ps;
    bool          have_path = false;
    const bool    no_depth  = s.max_depth == 0; // `depth < max_depth` fails at once: every sample is black

    for (;;) {
        // ---- regeneration: lanes without a path draw the next slots of this CTA's range (one atomic per warp) ----
        const unsigned want = __ballot_sync(0xffffffffu, !have_path);
        bool           drew = false;
        if (want) {
            uint32_t base = 0;
            if (lane == __ffs(want) - 1) {
                base = atomicAdd(&next_slot, static_cast<uint32_t>(__popc(want)));
            }
            base = __shfl_sync(0xffffffffu, base, __ffs(want) - 1);
            if (!have_path) {
                const uint32_t slot = base + __popc(want & ((1u << lane) - 1u));
                if (slot < end) {
                    drew               = true;
                    const uint32_t pix = __ldg(pix_list + slot % n_pix);
                    const uint32_t smp = sample_begin + slot / n_pix;
                    float4         o, d;
                    camera_ray(s, pix, smp, o, d);
                    ps.o          = xyz(o);
                    ps.d          = xyz(d);
                    ps.t_min      = o.w;
                    ps.throughput = v3(1.0f, 1.0f, 1.0f);
                    ps.L          = v3(0.0f, 0.0f, 0.0f);
                    ps.depth      = 0;
                    ps.slot       = slot;
                    ps.rng        = Rng{ pix, smp, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), 0u, 0u };
                    have_path     = true;
                    ++pc.paths;
                    if (no_depth) {
                        radiance[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        have_path      = false;
                    }
                }
            }
        }
        // Warp-uniform exit: no lane carries a path and this round drew nothing, i.e. the CTA's range is used up
        // (slots are handed out in increasing order).
        if (__ballot_sync(0xffffffffu, have_path) == 0u) {
            if (__ballot_sync(0xffffffffu, drew) == 0u) {
                break;
            }
            continue;
        }
        // ---- one vertex for every lane that carries a path --------------------------------------------------------
        if (have_path) {
            if (!path_vertex<kCount, F>(s, integrator, ps, stack, pc, kCount ? &tc : nullptr)) {
                radiance[ps.slot] = make_float4(ps.L.x, ps.L.y, ps.L.z, 0.0f);
                have_path         = false;
            }
        }
    }

    // ---- counters: one atomic per warp and counter ----------------------------------------------------------------------
    unsigned v[5] = { pc.paths, pc.rays_closest, pc.rays_any, pc.rays_lights, pc.shade_calls };
    const int idx[5] = { kCntPaths, kCntRaysClosest, kCntRaysAny, kCntRaysLights, kCntShadeCalls };
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const unsigned total = __reduce_add_sync(0xffffffffu, v[k]);
        if (lane == 0 && total) {
            atomicAdd(counters + idx[k], static_cast<unsigned long long>(total));
        }
    }
    if (kCount) {
        unsigned long long a = tc.nodes, b = tc.tris, c = tc.xf;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
            c += __shfl_down_sync(0xffffffffu, c, off);
        }
        if (lane == 0) {
            if (a) atomicAdd(&cnt->nodes, a);
            if (b) atomicAdd(&cnt->tris, b);
            if (c) atomicAdd(&cnt->xf, c);
        }
    }
}

template <typename K>
int ctas_per_sm(K kernel)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kPathBlock, 0) != cudaSuccess || n < 1) {
        n = 1;
    }
    return n;
}

} // namespace

template <typename F>
static void launch_paths_variant(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                                 uint32_t n_samples, uint64_t seed, uint32_t integrator, float4* d_radiance,
                                 unsigned long long* d_counters, TraceCounters* d_cnt)
{
    const uint32_t n = n_pix * n_samples;
    if (d_cnt) {
        static const int occ = ctas_per_sm(k_paths<true, F>);
        k_paths<true, F><<<wavefront_grid(n, kPathBlock, occ, l.sm_count), kPathBlock, 0, l.stream>>>(
            s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator, d_radiance, d_counters, d_cnt);
    } else {
        static const int occ = ctas_per_sm(k_paths<false, F>);
        k_paths<false, F><<<wavefront_grid(n, kPathBlock, occ, l.sm_count), kPathBlock, 0, l.stream>>>(
            s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator, d_radiance, d_counters, nullptr);
    }
}

void launch_paths(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                  uint32_t n_samples, uint64_t seed, uint32_t integrator, float4* d_radiance, unsigned long long* d_counters,
                  TraceCounters* d_cnt)
{
    const uint32_t n = n_pix * n_samples;
    if (n == 0) {
        return;
    }
    if (l.features == FeatAnalytic::id) {
        launch_paths_variant<FeatAnalytic>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator, d_radiance, d_counters,
                                           d_cnt);
    } else {
        launch_paths_variant<FeatFull>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator, d_radiance, d_counters,
                                       d_cnt);
    }
}

} // namespace spcu
