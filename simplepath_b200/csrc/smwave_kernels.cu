// The SM-local wavefront: a third organisation of the same estimator (same stages, same order, same addressed random
// numbers as spcu_render.cu's queue pipeline and path_kernels.cu's persistent path kernel).
//
// Why: the queue pipeline runs every stage with full warps but moves ~1 KB per path vertex through HBM between its
// kernels; the persistent path kernel moves nothing, but a lane carries one path through all stages, and lanes whose
// path missed, was occluded or ended wait for the slowest lane of the warp: ncu measured 11.7 of 32 lanes active per
// instruction on example_scene (profiles/r01j_*).  Here the wavefront lives in the SM: ONE persistent CTA per SM keeps
// the state of P paths (1.0-1.6 K, 140 B each) in its shared memory (up to 227 KB on sm_100a) together with one work
// queue per stage — extend (both closest-hit queries), shade (surface interaction, primary BSDF sample, roulette), light (light
// sample + shadow ray), mis (eval / pdf / second sample + BSDF-strategy rays) —
// and its warps are workers: a warp picks the stage with the most waiting items, takes up to 32 of them, runs that
// stage for all of them in lockstep and hands each item to its next stage's queue.  Paths that end are regenerated in
// place from the CTA's range of the batch.  Nothing but the pixel index (4 B in) and the radiance sample (16 B out)
// touches HBM, and every stage runs with (nearly) full warps.
//
// Bank-conflict-free by construction: slot i of the state belongs to lane i % 32, for ever.  State is a
// structure of arrays st[field][slot], so a warp's access to one field hits 32 different banks; the queues are 32
// independent rings per stage (one per lane class, [position][lane]), pushed and popped by per-lane shared-memory atomics
// on that lane's own counters — no ballots, no leader election, no warp-wide prefix sums on the hand-over path.
// Publication protocol of a ring entry (producer and consumer are lanes of different warps): the producer reserves a
// position with atomicAdd(tail), writes the item's state, fences, then stores entry = index + 1; the consumer claims a
// position with atomicCAS(head) only while head != tail, spins until the entry is non-zero, clears it and fences before
// reading the state.  A class owns at most P/32 < kRing slots and a slot is in at most one queue, so a ring position is
// never reused while it is still claimed.
#include "kernels.h"
#include "shade.cuh"
#include "trace.cuh"

#include <cstdlib>

namespace spcu {
namespace {

constexpr int kRing    = 64;        // ring entries per (stage, lane class); P / 32 must stay below it
// Work queues.  A stage whose code depends on WHICH light it handles has one queue per parity of the light cursor: with the
// usual one or two lights every warp of `light` / `mis` then runs one light kind (ncu on example_scene, one queue per
// stage: Light::sample ran at 16 of 32 lanes — sphere-light lanes and environment-light lanes took turns).
constexpr int kQExtend = 0, kQShade = 1, kQLight = 2, kQMis = 4, kQueues = 6;

// state fields (32-bit words), st[field * P + slot]
enum : int {
    kFSlot = 0,  // index of the camera sample in the batch (radiance[slot])
    kFPix,       // pixel (RNG key word 0)
    kFSmp,       // global sample index (RNG key word 1)
    kFFlags,     // depth (bits 0-11) | light cursor k (bits 12-27) | kAlive | kRegen
    kFTp,        // throughput (3)
    kFL = kFTp + 3,          // radiance accumulated so far (3)
    kFO = kFL + 3,           // ray origin; after the hit: the surface point (3)
    kFD = kFO + 3,           // ray direction (3)
    kFTmin = kFD + 3,        // ray t_min
    kFN,                     // shading normal (3)
    kFMat = kFN + 3,         // material
    kFNextD,                 // next segment, decided at the hit: direction (3), t_min, throughput (3)
    kFNextTmin = kFNextD + 3,
    kFNextTp,
    kFLwi = kFNextTp + 3,    // light sample handed from `light` to `mis`: direction (3), pdf, radiance (3)
    kFLpdf = kFLwi + 3,
    kFLrad,
    kFields = kFLrad + 3
};
constexpr uint32_t kAlive = 1u << 30, kRegen = 1u << 31;
constexpr uint32_t kDepthMask = 0xfffu, kCursorShift = 12u, kCursorMask = 0xffffu;

struct SmQueues
{
    uint32_t head[kQueues][32];
    uint32_t tail[kQueues][32];
    uint8_t  ring[kQueues][kRing][32];
    uint32_t next_slot; // next camera sample of this CTA's range
    uint32_t live;      // state slots that still carry, or may still draw, a path
    int      preferred; // the stage the CTA ran last
    uint32_t error;     // a bounded wait ran out: every warp leaves, the kernel reports kCntErrors (never a hung GPU)
    uint32_t idle[32];  // per warp: consecutive empty scans (in shared memory: the kernel has no register to spare for it)
};
#ifndef SPCU_SMWAVE_CAPS
#define SPCU_SMWAVE_CAPS 1 // 0: A/B builds without the bounded waits
#endif
constexpr uint32_t kSpinCap = 1u << 22; // polls of a slot whose publication is in flight (normally a handful)
constexpr uint32_t kIdleCap = 1u << 22; // consecutive empty scans (x 64 ns sleep) of a warp while paths are still live

struct PathCounters
{
    unsigned paths = 0, rays_closest = 0, rays_any = 0, rays_lights = 0, shade_calls = 0;
};

__device__ __forceinline__ Ray make_ray(V3 o, V3 d, float t_min) { return Ray{ o.x, o.y, o.z, d.x, d.y, d.z, t_min }; }

// Out-of-line copies of the two queries that several stages issue (instruction footprint, see the header comment).
template <bool kCount, typename F>
static __device__ __noinline__ bool occluded(const DScene& s, V3 o, V3 d, float t_min, float t_max, int32_t* stack, TraceCounters* tc)
{
    return scene_any_hit<kCount, F>(s, make_ray(o, d, t_min), t_max, stack, tc);
}

template <typename F>
static __device__ __noinline__ int32_t nearest_light(const DScene& s, V3 o, V3 d, float t_min, float& t_max, int32_t* stack)
{
    float                beta, gamma;
    const LightPrimsT<F> lp{ s.lights };
    return closest_hit<false>(s.lights_accel, lp, make_ray(o, d, t_min), t_max, beta, gamma, stack, nullptr);
}

// ---- per-lane queue operations -----------------------------------------------------------------------------------------
__device__ __forceinline__ void q_push(SmQueues& q, int stage, int lane, uint32_t idx)
{
    const uint32_t pos = atomicAdd(&q.tail[stage][lane], 1u);
    __threadfence_block(); // the item's state before its publication
    *reinterpret_cast<volatile uint8_t*>(&q.ring[stage][pos & (kRing - 1)][lane]) = static_cast<uint8_t>(idx + 1u);
}

__device__ __forceinline__ int q_pop(SmQueues& q, int stage, int lane)
{
    uint32_t h = *reinterpret_cast<volatile uint32_t*>(&q.head[stage][lane]);
    for (;;) {
        const uint32_t t = *reinterpret_cast<volatile uint32_t*>(&q.tail[stage][lane]);
        if (h == t) {
            return -1;
        }
        const uint32_t old = atomicCAS(&q.head[stage][lane], h, h + 1u);
        if (old == h) {
            break;
        }
        h = old;
    }
    volatile uint8_t* e = reinterpret_cast<volatile uint8_t*>(&q.ring[stage][h & (kRing - 1)][lane]);
    uint32_t          v;
    uint32_t          spins = 0;
    while ((v = *e) == 0u) {
        if (SPCU_SMWAVE_CAPS && ++spins > kSpinCap) { // a lost publication: fail the render, do not hang the device
            *reinterpret_cast<volatile uint32_t*>(&q.error) = 1u;
            return -1;
        }
    }
    *e = 0u;
    __threadfence_block();
    return static_cast<int>(v) - 1;
}

// ---- state access --------------------------------------------------------------------------------------------------------
struct State
{
    float*   st;
    uint32_t P;
    uint32_t slot; // idx * 32 + lane

    __device__ __forceinline__ float&    f(int field) const { return st[field * P + slot]; }
    __device__ __forceinline__ uint32_t& u(int field) const { return reinterpret_cast<uint32_t*>(st)[field * P + slot]; }
    __device__ __forceinline__ V3   get3(int field) const { return v3(f(field), f(field + 1), f(field + 2)); }
    __device__ __forceinline__ void set3(int field, V3 v) const
    {
        f(field)     = v.x;
        f(field + 1) = v.y;
        f(field + 2) = v.z;
    }
};

struct Shared // what every stage needs, by reference
{
    const DScene&   s;
    SmQueues&       q;
    float4*         radiance;
    uint32_t        seed_lo, seed_hi;
    int32_t*        stack; // this thread's column of the traversal stack (F::bvh only)
    PathCounters&   pc;
    TraceCounters*  tc;
};

__device__ __forceinline__ Rng state_rng(const Shared& sh, const State& x, uint32_t depth, uint32_t site, uint32_t ctr)
{
    return Rng{ x.u(kFPix), x.u(kFSmp), sh.seed_lo, sh.seed_hi, rng_stream(depth, site), ctr };
}

// the path ends: its radiance sample is final, its state slot draws a new camera sample at its next extend
__device__ __forceinline__ void terminate(const Shared& sh, const State& x, int lane, V3 L)
{
    sh.radiance[x.u(kFSlot)] = make_float4(L.x, L.y, L.z, 0.0f);
    x.u(kFFlags)             = kRegen;
    q_push(sh.q, kQExtend, lane, x.slot >> 5);
}

// all lights of the vertex are done (or there are none): continue with the segment decided at the hit, or end
template <typename F, uint32_t kI>
__device__ __forceinline__ void finish_vertex(const Shared& sh, const State& x, int lane, uint32_t flags)
{
    const uint32_t depth = flags & kDepthMask;
    if (kI == SPCU_INTEGRATOR_DIRECT_LIGHTING) {
        terminate(sh, x, lane, x.get3(kFL));
        return;
    }
    if (kI == SPCU_INTEGRATOR_WHITTED) {
        // WhittedIntegrator (Integrator.cpp:357-363): follow the BSDF sample only when it is specular, default limits,
        // radiance of the reflected ray added unweighted
        Rng           rng = state_rng(sh, x, depth, kSiteBsdf, 0u);
        const MSample ms  = material_sample<F>(sh.s, x.u(kFMat), -x.get3(kFD), x.get3(kFN), rng);
        ++sh.pc.shade_calls;
        if (!ms.specular || depth + 1u >= sh.s.max_depth) { // is_specular(properties) alone decides (:359)
            terminate(sh, x, lane, x.get3(kFL));
            return;
        }
        x.set3(kFD, ms.dir);
        x.f(kFTmin)  = kRayEpsilon;
        x.u(kFFlags) = depth + 1u;
        q_push(sh.q, kQExtend, lane, x.slot >> 5);
        return;
    }
    if (!(flags & kAlive)) {
        terminate(sh, x, lane, x.get3(kFL));
        return;
    }
    x.set3(kFD, x.get3(kFNextD));
    x.f(kFTmin) = x.f(kFNextTmin);
    x.set3(kFTp, x.get3(kFNextTp));
    x.u(kFFlags) = depth + 1u;
    q_push(sh.q, kQExtend, lane, x.slot >> 5);
}

// the light under the cursor is done: the next light of the vertex, or the end of the vertex
template <typename F, uint32_t kI>
__device__ __forceinline__ void next_light(const Shared& sh, const State& x, int lane, uint32_t flags)
{
    const uint32_t k = ((flags >> kCursorShift) & kCursorMask) + 1u;
    if (k < sh.s.n_lights) {
        x.u(kFFlags) = (flags & ~(kCursorMask << kCursorShift)) | (k << kCursorShift);
        q_push(sh.q, kQLight + static_cast<int>(k & 1u), lane, x.slot >> 5);
    } else {
        finish_vertex<F, kI>(sh, x, lane, flags);
    }
}

// ---- stage: extend (regeneration, Scene::intersect_lights + Scene::intersect, miss / hit handling, primary BSDF sample,
// throughput update and Russian roulette: Integrator.cpp:556-572, 601-632) -----------------------------------------------
template <bool kCount, typename F, uint32_t kI>
__device__ __forceinline__ void stage_extend(const Shared& sh, State x, int lane, bool have, const uint32_t* pix_list,
                                             uint32_t n_pix, uint32_t sample_begin, uint32_t end)
{
    const DScene& s = sh.s;
    // regeneration: lanes whose slot is empty draw the next camera samples of the CTA's range (one atomic per warp)
    const bool     regen = have && (x.u(kFFlags) & kRegen);
    const unsigned want  = __ballot_sync(0xffffffffu, regen);
    if (want) {
        uint32_t base = 0;
        if (lane == __ffs(want) - 1) {
            base = atomicAdd(&sh.q.next_slot, static_cast<uint32_t>(__popc(want)));
        }
        base = __shfl_sync(0xffffffffu, base, __ffs(want) - 1);
        if (regen) {
            const uint32_t slot = base + __popc(want & ((1u << lane) - 1u));
            if (slot >= end) { // the range is used up: this state slot retires
                atomicSub(&sh.q.live, 1u);
                have = false;
            } else {
                const uint32_t pix = __ldg(pix_list + slot % n_pix);
                const uint32_t smp = sample_begin + slot / n_pix;
                float4         o, d;
                camera_ray(s, pix, smp, o, d);
                x.u(kFSlot)  = slot;
                x.u(kFPix)   = pix;
                x.u(kFSmp)   = smp;
                x.u(kFFlags) = 0u;
                x.set3(kFTp, v3(1.0f, 1.0f, 1.0f));
                x.set3(kFL, v3(0.0f, 0.0f, 0.0f));
                x.set3(kFO, xyz(o));
                x.set3(kFD, xyz(d));
                x.f(kFTmin) = o.w;
                ++sh.pc.paths;
                if (s.max_depth == 0u) { // `depth < max_depth` fails at once: every sample is black
                    terminate(sh, x, lane, v3(0.0f, 0.0f, 0.0f));
                    have = false;
                }
            }
        }
    }
    if (!have) {
        return;
    }
    const uint32_t depth = x.u(kFFlags) & kDepthMask;
    const V3       o = x.get3(kFO), d = x.get3(kFD);
    const Ray      r     = make_ray(o, d, x.f(kFTmin));
    float          t_max = kInfinite, beta, gamma;

    ++sh.pc.rays_lights;
    // Scene::intersect_lights inlined at this site (every segment of every path comes through here; the BSDF-strategy ray of
    // the mis stage keeps the out-of-line copy): +2 % paths/s, like the shadow query of the light stage.
    const LightPrimsT<F> light_prims{ s.lights };
    float                light_beta, light_gamma;
    const int32_t        li = closest_hit<false>(s.lights_accel, light_prims, r, t_max, light_beta, light_gamma, sh.stack, nullptr);
    ++sh.pc.rays_closest;
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    const int32_t       gi = closest_hit<kCount>(s.geom, gp, r, t_max, beta, gamma, sh.stack, sh.tc);
    if (gi < 0) {
        V3 L = x.get3(kFL);
        if (li >= 0) { // emitter reached, at ANY depth (Integrator.cpp:627-629)
            L = L + x.get3(kFTp) * light_hit_L<F>(s, s.lights[li], d);
        }
        terminate(sh, x, lane, L);
        return;
    }
    // the hit record travels in the fields the shade stage fills afterwards
    x.u(kFN)     = static_cast<uint32_t>(gi);
    x.f(kFN + 1) = t_max;
    x.f(kFN + 2) = beta;
    x.f(kFMat)   = gamma;
    q_push(sh.q, kQShade, lane, x.slot >> 5);
}

// ---- stage: shade (surface interaction, primary BSDF sample, throughput update and Russian roulette:
// Integrator.cpp:565-572, 601-626) ---------------------------------------------------------------------------------------------
template <bool kCount, typename F, uint32_t kI>
__device__ __forceinline__ void stage_shade(const Shared& sh, State x, int lane)
{
    const DScene&  s     = sh.s;
    const uint32_t depth = x.u(kFFlags) & kDepthMask;
    const V3       o = x.get3(kFO), d = x.get3(kFD);
    V3             point, normal;
    uint32_t       material;
    make_isect<F>(s, HitRec{ static_cast<int32_t>(x.u(kFN)), x.f(kFN + 1), x.f(kFN + 2), x.f(kFMat) }, o, d, point, normal,
                  material);

    x.set3(kFO, point);
    x.set3(kFN, normal);
    x.u(kFMat) = material;

    if (kI == SPCU_INTEGRATOR_DIRECT_LIGHTING || kI == SPCU_INTEGRATOR_WHITTED) {
        // these sample the lights first; Whitted draws its BSDF sample afterwards (finish_vertex)
        if (s.n_lights) {
            x.u(kFFlags) = depth;
            q_push(sh.q, kQLight, lane, x.slot >> 5);
        } else {
            finish_vertex<F, kI>(sh, x, lane, depth);
        }
        return;
    }

    Rng           rng = state_rng(sh, x, depth, kSiteBsdf, 0u);
    const MSample sr  = material_sample<F>(s, material, -d, normal, rng);
    ++sh.pc.shade_calls;
    if (sr.pdf == 0.0f || is_black(sr.color)) {
        terminate(sh, x, lane, x.get3(kFL));
        return;
    }
    // The continuation is decided now (its random numbers are addressed, not streamed, so the order of the stages does
    // not matter); the lights of this vertex still see the throughput the path arrived with.
    const float cosine = fabsf(dot(sr.dir, normal));
    V3          tp     = x.get3(kFTp) * (cosine * sr.color / sr.pdf);
    bool        alive  = true;
    if (depth >= s.rr_depth) {
        const float lum = luminance(tp);
        if (lum < 0.1f) {
            const float q  = max_std(0.05f, lum / 0.1f);
            Rng         rr = state_rng(sh, x, depth, kSiteRoulette, 0u);
            if (rng_next1(rr) < q) {
                tp = tp / q;
            } else {
                alive = false;
            }
        }
    }
    alive = alive && depth + 1u < s.max_depth;
    x.set3(kFNextD, sr.dir);
    x.f(kFNextTmin) = ray_offset_cos(cosine);
    x.set3(kFNextTp, tp);
    const uint32_t flags = depth | (alive ? kAlive : 0u);
    if (kI == SPCU_INTEGRATOR_ITERATIVE_RRNEE && s.n_lights) {
        x.u(kFFlags) = flags;
        q_push(sh.q, kQLight, lane, x.slot >> 5);
    } else {
        finish_vertex<F, kI>(sh, x, lane, flags);
    }
}

// ---- stage: light (Light::sample + the shadow query: Integrator.cpp:497-506; for direct lighting / Whitted the whole
// per-light term, Integrator.cpp:289-306 / :336-355) -------------------------------------------------------------------------
template <bool kCount, typename F, uint32_t kI>
__device__ __forceinline__ void stage_light(const Shared& sh, State x, int lane)
{
    const DScene&     s     = sh.s;
    const uint32_t    flags = x.u(kFFlags);
    const uint32_t    depth = flags & kDepthMask, k = (flags >> kCursorShift) & kCursorMask;
    const spcu_light& light = s.lights[__ldg(s.light_order + k)];
    const V3          point = x.get3(kFO), normal = x.get3(kFN);
    Rng               rng   = state_rng(sh, x, depth, kSiteLight0 + k, 0u);
    float             u0, u1;
    rng_next2(rng, u0, u1);
    const LSample ls = light_sample<F>(s, light, point, normal, u0, u1);
    if (ls.pdf == 0.0f || is_black(ls.L)) {
        next_light<F, kI>(sh, x, lane, flags);
        return;
    }
    if (kI == SPCU_INTEGRATOR_DIRECT_LIGHTING || kI == SPCU_INTEGRATOR_WHITTED) {
        const Onb onb = onb_from_v(normal);
        const V3  f   = material_eval_local<F>(s, x.u(kFMat), to_onb(onb, -x.get3(kFD)), to_onb(onb, ls.wi), rng);
        ++sh.pc.shade_calls;
        if (!is_black(f)) {
            ++sh.pc.rays_any;
            if (!occluded<kCount, F>(s, point, ls.wi, ls.t_min, ls.t_max, sh.stack, sh.tc)) {
                x.set3(kFL, x.get3(kFL) + f * ls.L * fabsf(dot(ls.wi, normal)) / ls.pdf);
            }
        }
        next_light<F, kI>(sh, x, lane, flags);
        return;
    }
    ++sh.pc.rays_any;
    // inlined here, the one call site every light sample of the NEE integrator reaches: the call sequence of the out-of-line
    // copy was 8 % of the kernel's instructions (profiles/hot_lines.py on the r01zc capture); +2 % paths/s.  The rarer sites
    // (BSDF-strategy ray, direct lighting / Whitted) keep the shared copy: inlining them too measured no further gain.
    if (scene_any_hit<kCount, F>(s, make_ray(point, ls.wi, ls.t_min), ls.t_max, sh.stack, sh.tc)) {
        next_light<F, kI>(sh, x, lane, flags);
        return;
    }
    x.set3(kFLwi, ls.wi);
    x.f(kFLpdf) = ls.pdf;
    x.set3(kFLrad, ls.L);
    q_push(sh.q, kQMis + static_cast<int>(k & 1u), lane, x.slot >> 5);
}

// ---- stage: mis (the rest of estimate_direct_mis for an unoccluded light sample: Integrator.cpp:508-538) ----------------
template <bool kCount, typename F, uint32_t kI>
__device__ __forceinline__ void stage_mis(const Shared& sh, State x, int lane)
{
    const DScene&     s     = sh.s;
    const uint32_t    flags = x.u(kFFlags);
    const uint32_t    depth = flags & kDepthMask, k = (flags >> kCursorShift) & kCursorMask;
    const spcu_light& light = s.lights[__ldg(s.light_order + k)];
    const V3          p = x.get3(kFO), n = x.get3(kFN);
    const uint32_t    material = x.u(kFMat);
    const V3          wo = -x.get3(kFD), wi = x.get3(kFLwi);
    const float       lpdf_s = x.f(kFLpdf);
    Rng               rng    = state_rng(sh, x, depth, kSiteLight0 + k, 1u); // draw 0 was the light sample

    V3        L   = v3(0, 0, 0);
    const Onb onb = onb_from_v(n);
    const V3  wol = to_onb(onb, wo), wil = to_onb(onb, wi);
    const Coats<F> coats = walk_coats<F>(s, material, wol);
    const V3  f   = material_eval_coats<F>(s, coats, wol, wil, rng);
    ++sh.pc.shade_calls;
    if (!is_black(f)) {
        const float bsdf_pdf = material_pdf_coats<F>(s, coats, wol, wil, rng);
        ++sh.pc.shade_calls;
        if (bsdf_pdf > 0.0f) {
            const float weight = balance2(lpdf_s, bsdf_pdf);
            L                  = f * x.get3(kFLrad) * (fabsf(dot(wi, n)) * weight / lpdf_s);
        }
    }
    MSample ms = material_sample_local<F>(s, material, wol, rng);
    ++sh.pc.shade_calls;
    float lpdf = 0.0f;
    if (!(ms.pdf == 0.0f || is_black(ms.color))) {
        ms.dir = to_world(onb, ms.dir);
        lpdf   = light_pdf<F>(s, light, p, ms.dir);
    }
    if (lpdf != 0.0f) {
        const float weight = balance2(ms.pdf, lpdf);
        const float t_min  = ray_offset(n, ms.dir);
        float       t_max  = kInfinite;
        ++sh.pc.rays_lights;
        const int32_t li = nearest_light<F>(s, p, ms.dir, t_min, t_max, sh.stack);
        if (li >= 0) {
            ++sh.pc.rays_any;
            // limits are NOT shrunk to the light's distance: a sphere light occludes itself, as in the reference (:531-532)
            if (!occluded<kCount, F>(s, p, ms.dir, t_min, kInfinite, sh.stack, sh.tc)) {
                const V3 Li = light_hit_L<F>(s, s.lights[li], ms.dir);
                L           = L + ms.color * Li * fabsf(dot(ms.dir, n)) * weight / ms.pdf;
            }
        }
    }
    x.set3(kFL, x.get3(kFL) + x.get3(kFTp) * L);
    next_light<F, kI>(sh, x, lane, flags);
}

template <bool kCount, typename F, uint32_t kI, int kSmBlock>
__global__ void __launch_bounds__(kSmBlock, 1)
    k_smwave(const __grid_constant__ DScene s, const uint32_t* __restrict__ pix_list, uint32_t n_pix, uint32_t sample_begin,
             uint32_t n_samples, uint64_t seed, float4* __restrict__ radiance, uint32_t P, int use_affinity,
             unsigned long long* counters, TraceCounters* cnt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmQueues& q     = *reinterpret_cast<SmQueues*>(smem_raw);
    float*    st    = reinterpret_cast<float*>(smem_raw + sizeof(SmQueues));
    int32_t*  stack = nullptr;
    if (F::bvh) { // [group of 128 threads][level][thread]: the layout trace.cuh's Stack expects
        int32_t* base = reinterpret_cast<int32_t*>(st + static_cast<size_t>(kFields) * P);
        stack         = base + (threadIdx.x >> 7) * (kStackShared * kTraceBlock) + (threadIdx.x & (kTraceBlock - 1));
    }

    const uint32_t n       = n_pix * n_samples;
    const uint32_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t begin   = min(n, blockIdx.x * per_cta);
    const uint32_t end     = min(n, begin + per_cta);
    const int      lane    = threadIdx.x & 31;

    // every state slot starts empty and waiting in the extend queue
    for (uint32_t i = threadIdx.x; i < kQueues * 32; i += kSmBlock) {
        (&q.head[0][0])[i] = 0u;
        (&q.tail[0][0])[i] = (i < 32) ? P / 32 : 0u;
    }
    for (uint32_t i = threadIdx.x; i < kQueues * kRing * 32; i += kSmBlock) {
        const uint32_t stage = i / (kRing * 32), pos = (i / 32) % kRing;
        (&q.ring[0][0][0])[i] = (stage == kQExtend && pos < P / 32) ? static_cast<uint8_t>(pos + 1u) : 0u;
    }
    for (uint32_t i = threadIdx.x; i < P; i += kSmBlock) {
        reinterpret_cast<uint32_t*>(st)[kFFlags * P + i] = kRegen;
    }
    if (threadIdx.x == 0) {
        q.next_slot = begin;
        q.live      = P;
        q.preferred = kQExtend;
        q.error     = 0u;
    }
    if (threadIdx.x < 32) {
        q.idle[threadIdx.x] = 0u;
    }
    __syncthreads();
    volatile uint32_t* idle = &q.idle[threadIdx.x >> 5];

    PathCounters  pc;
    TraceCounters tc{ 0, 0, 0 };
    const Shared  sh{ s, q, radiance, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), stack, pc,
                     kCount ? &tc : nullptr };

    for (;;) {
        // The CTA drifts through the stages together: a warp stays with the stage the CTA ran last while that queue can
        // still fill `use_affinity` lanes, and moves the CTA on when it cannot — to the queue with the most lane classes
        // that have an item waiting (ties: the later stage, to drain vertices).  The SM's L1.5 instruction cache is 32 KB
        // and the stages together are ~75 KB of SASS; with every warp on its own stage the issue slots starved on
        // instruction fetch (ncu: stall_no_instruction 2.4-6.3 cycles per issued instruction, 2.2 against 3.3 Gpaths/s).
        const volatile uint32_t* vh = &q.head[0][0];
        const volatile uint32_t* vt = &q.tail[0][0];
        int stage = -1;
        if (use_affinity) {
            const int preferred = *reinterpret_cast<volatile int*>(&q.preferred);
            const int np = __popc(__ballot_sync(0xffffffffu, vh[preferred * 32 + lane] != vt[preferred * 32 + lane]));
            if (np >= use_affinity) {
                stage = preferred;
            }
        }
        if (stage < 0) {
            int best = 0;
#pragma unroll
            for (int k = 0; k < kQueues; ++k) {
                const int nk = __popc(__ballot_sync(0xffffffffu, vh[k * 32 + lane] != vt[k * 32 + lane]));
                if (nk >= best) {
                    best  = nk;
                    stage = k;
                }
            }
            if (best == 0) {
                if (*reinterpret_cast<volatile uint32_t*>(&q.live) == 0u || *reinterpret_cast<volatile uint32_t*>(&q.error) != 0u) {
                    break;
                }
                // live paths but no queue ever fills again: a protocol fault, not a reason to hang (the count is the warp's: lane 0 keeps it)
                if (SPCU_SMWAVE_CAPS && __shfl_sync(0xffffffffu, lane == 0 ? (*idle = *idle + 1u) : 0u, 0) > kIdleCap) {
                    *reinterpret_cast<volatile uint32_t*>(&q.error) = 1u;
                    break;
                }
                __nanosleep(64);
                continue;
            }
            if (SPCU_SMWAVE_CAPS && lane == 0 && *idle) {
                *idle = 0u; // work again
            }
            if (use_affinity && lane == 0) {
                *reinterpret_cast<volatile int*>(&q.preferred) = stage;
            }
        }
        const int   idx = q_pop(q, stage, lane);
        const State x{ st, P, static_cast<uint32_t>(max(idx, 0)) * 32u + static_cast<uint32_t>(lane) };
        if (stage == kQExtend) {
            stage_extend<kCount, F, kI>(sh, x, lane, idx >= 0, pix_list, n_pix, sample_begin, end);
        } else if (idx >= 0) {
            if (stage == kQShade) {
                stage_shade<kCount, F, kI>(sh, x, lane);
            } else if (stage < kQMis) {
                stage_light<kCount, F, kI>(sh, x, lane);
            } else {
                stage_mis<kCount, F, kI>(sh, x, lane);
            }
        }
        __syncwarp();
    }

    if (lane == 0 && *reinterpret_cast<volatile uint32_t*>(&q.error) != 0u) {
        atomicAdd(counters + kCntErrors, 1ull);
    }
    // ---- counters: one atomic per warp and counter ----------------------------------------------------------------------
    unsigned  v[5]   = { pc.paths, pc.rays_closest, pc.rays_any, pc.rays_lights, pc.shade_calls };
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

// P = paths resident per SM: whatever fits beside the queues (and the traversal stacks), a multiple of 32, below 32 * kRing
template <typename F>
uint32_t resident_paths(size_t smem_limit, size_t& smem_bytes, int kSmBlock)
{
    const size_t fixed = sizeof(SmQueues) + (F::bvh ? static_cast<size_t>(kSmBlock) * kStackShared * sizeof(int32_t) : 0);
    size_t       P     = (smem_limit - fixed) / (kFields * sizeof(float));
    P                  = std::min<size_t>(P / 32 * 32, 32 * (kRing - 8));
    smem_bytes         = fixed + P * kFields * sizeof(float);
    return static_cast<uint32_t>(P);
}

template <bool kCount, typename F, uint32_t kI, int kSmBlock>
cudaError_t launch_block(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                           uint32_t n_samples, uint64_t seed, float4* d_radiance, unsigned long long* d_counters,
                           TraceCounters* d_cnt)
{
    static size_t      smem = 0;
    static uint32_t    P    = 0;
    static cudaError_t init = [] {
        int dev = 0, limit = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        P = resident_paths<F>(static_cast<size_t>(limit), smem, kSmBlock);
        return cudaFuncSetAttribute(k_smwave<kCount, F, kI, kSmBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    }();
    if (init != cudaSuccess) {
        return init;
    }
    static const int affinity = [] {
        // lane classes the CTA's current stage must still fill for a warp to stay with it; 0 = every warp picks the
        // fullest queue (measured on example_scene: 2.2 against 3.3 Gpaths/s)
        const char* e = getenv("SPCU_SMWAVE_AFFINITY");
        return e ? atoi(e) : 16;
    }();
    const uint32_t n    = n_pix * n_samples;
    const unsigned grid = static_cast<unsigned>(std::max<uint32_t>(1u, std::min<uint32_t>(l.sm_count, (n + 255u) / 256u)));
    k_smwave<kCount, F, kI, kSmBlock><<<grid, kSmBlock, smem, l.stream>>>(s, d_pix_list, n_pix, sample_begin, n_samples, seed, d_radiance, P, affinity,
                                                                d_counters, d_cnt);
    return cudaSuccess;
}

template <bool kCount, typename F, uint32_t kI>
cudaError_t launch_variant(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                           uint32_t n_samples, uint64_t seed, float4* d_radiance, unsigned long long* d_counters,
                           TraceCounters* d_cnt)
{
    static const int block = [] {
        const char* e = getenv("SPCU_SMWAVE_BLOCK"); // tuning knob; measured best on example_scene: 768 (profiles/)
        return e ? atoi(e) : (F::bvh ? 512 : 768);
    }();
    switch (block) {
    case 768: return launch_block<kCount, F, kI, 768>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, d_radiance, d_counters, d_cnt);
    case 1024: return launch_block<kCount, F, kI, 1024>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, d_radiance, d_counters, d_cnt);
    default: return launch_block<kCount, F, kI, 512>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, d_radiance, d_counters, d_cnt);
    }
}

// the kernel is compiled per (node counting, scene feature set, integrator): what a render does not need is not in its
// instruction stream (the organisation's weak spot is instruction fetch: warps of one SM run different stages)
template <bool kCount, typename F>
cudaError_t launch_integrator(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                              uint32_t n_samples, uint64_t seed, uint32_t integrator, float4* d_radiance,
                              unsigned long long* d_counters, TraceCounters* d_cnt)
{
#define SPCU_SMWAVE_CASE(I)                                                                                                 \
    case I:                                                                                                                 \
        return launch_variant<kCount, F, I>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, d_radiance, d_counters, d_cnt)
    switch (integrator) {
        SPCU_SMWAVE_CASE(SPCU_INTEGRATOR_ITERATIVE_RRNEE);
        SPCU_SMWAVE_CASE(SPCU_INTEGRATOR_BRUTE_FORCE_RR);
        SPCU_SMWAVE_CASE(SPCU_INTEGRATOR_DIRECT_LIGHTING);
        SPCU_SMWAVE_CASE(SPCU_INTEGRATOR_WHITTED);
    default:
        return cudaErrorInvalidValue;
    }
#undef SPCU_SMWAVE_CASE
}

cudaError_t launch_features(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                            uint32_t n_samples, uint64_t seed, uint32_t integrator, float4* d_radiance,
                            unsigned long long* d_counters, TraceCounters* d_cnt)
{
    if (l.features == FeatAnalytic::id) {
        return d_cnt ? launch_integrator<true, FeatAnalytic>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator,
                                                             d_radiance, d_counters, d_cnt)
                     : launch_integrator<false, FeatAnalytic>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator,
                                                              d_radiance, d_counters, d_cnt);
    }
    return d_cnt ? launch_integrator<true, FeatFull>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator,
                                                     d_radiance, d_counters, d_cnt)
                 : launch_integrator<false, FeatFull>(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator,
                                                      d_radiance, d_counters, d_cnt);
}

} // namespace

bool smwave_supports(const DScene& s) { return s.max_depth <= kDepthMask && s.n_lights <= kCursorMask; }

cudaError_t launch_smwave(const Launch& l, const DScene& s, const uint32_t* d_pix_list, uint32_t n_pix, uint32_t sample_begin,
                          uint32_t n_samples, uint64_t seed, uint32_t integrator, float4* d_radiance,
                          unsigned long long* d_counters, TraceCounters* d_cnt)
{
    if (n_pix * n_samples == 0) {
        return cudaSuccess;
    }
    if (!smwave_supports(s)) {
        return cudaErrorInvalidValue;
    }
    return launch_features(l, s, d_pix_list, n_pix, sample_begin, n_samples, seed, integrator, d_radiance, d_counters, d_cnt);
}

} // namespace spcu
