// Traversal kernels of the backend: the batch queries behind spcu_trace_* (parity tests) and the three wavefront
// stages that traverse (extend, shadow, mis).  All of them walk the flattened accelerators in REFERENCE ORDER with
// the reference's exact arithmetic (trace.cuh); this TU is compiled with --fmad=false.
//
// Launch shape: one thread per ray, kTraceBlock threads per CTA, grid = ceil(n / kTraceBlock).  The traversal stack's
// first kStackShared levels live in shared memory ([level][thread], conflict free), the rest spills to local memory.
#include "kernels.h"
#include "rng.cuh"
#include "trace.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace spcu {
namespace {

__device__ __forceinline__ Ray load_ray(const spcu_ray* rays, uint64_t i, float& t_max)
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 0);
    const float4 b = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 1);
    t_max          = b.w;
    return Ray{ a.x, a.y, a.z, b.x, b.y, b.z, a.w };
}

// Warp-aggregated add of per-thread counters: one atomic per warp and counter.
__device__ __forceinline__ void flush_counters(const TraceCounters& local, TraceCounters* global)
{
    unsigned long long n = local.nodes, t = local.tris, x = local.xf;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n += __shfl_down_sync(0xffffffffu, n, off);
        t += __shfl_down_sync(0xffffffffu, t, off);
        x += __shfl_down_sync(0xffffffffu, x, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n) atomicAdd(&global->nodes, n);
        if (t) atomicAdd(&global->tris, t);
        if (x) atomicAdd(&global->xf, x);
    }
}

__device__ __forceinline__ void warp_count(unsigned long long* counter, bool pred)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(counter, static_cast<unsigned long long>(__popc(m)));
    }
}

// ---- batch queries -----------------------------------------------------------------------------------------------
template <bool kCount>
__global__ void __launch_bounds__(kTraceBlock) k_trace_closest(const __grid_constant__ DScene s, const spcu_ray* rays,
                                                               uint64_t n, spcu_hit* hits, TraceCounters* cnt)
{
    __shared__ int32_t stack[kStackShared * kTraceBlock];
    const uint64_t     i = static_cast<uint64_t>(blockIdx.x) * kTraceBlock + threadIdx.x;
    TraceCounters      local{ 0, 0, 0 };
    if (i < n) {
        float           t_max, beta, gamma;
        const Ray       r = load_ray(rays, i, t_max);
        const GeomPrims gp{ s.geom_prims, s.geom_meta };
        const int32_t   id = closest_hit<kCount>(s.geom, gp, r, t_max, beta, gamma, stack + threadIdx.x, &local);
        hits[i]            = spcu_hit{ id, t_max };
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

template <bool kCount>
__global__ void __launch_bounds__(kTraceBlock) k_trace_closest_ordered(const __grid_constant__ DScene s, const spcu_ray* rays,
                                                                       uint64_t n, spcu_hit* hits, TraceCounters* cnt)
{
    __shared__ int32_t stack[kStackShared * kTraceBlock];
    const uint64_t     i = static_cast<uint64_t>(blockIdx.x) * kTraceBlock + threadIdx.x;
    TraceCounters      local{ 0, 0, 0 };
    if (i < n) {
        float           t_max, beta, gamma;
        const Ray       r = load_ray(rays, i, t_max);
        const GeomPrims gp{ s.geom_prims, s.geom_meta };
        const int32_t   id = closest_hit_ordered<kCount>(s.geom, gp, r, t_max, beta, gamma, stack + threadIdx.x, &local);
        hits[i]            = spcu_hit{ id, t_max };
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

__global__ void __launch_bounds__(kTraceBlock) k_trace_any(const __grid_constant__ DScene s, const spcu_ray* rays, uint64_t n,
                                                           uint8_t* out)
{
    __shared__ int32_t stack[kStackShared * kTraceBlock];
    const uint64_t     i = static_cast<uint64_t>(blockIdx.x) * kTraceBlock + threadIdx.x;
    if (i < n) {
        float     t_max;
        const Ray r = load_ray(rays, i, t_max);
        out[i]      = scene_any_hit<false>(s, r, t_max, stack + threadIdx.x, nullptr) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kTraceBlock) k_trace_lights(const __grid_constant__ DScene s, const spcu_ray* rays,
                                                              uint64_t n, spcu_hit* hits)
{
    __shared__ int32_t stack[kStackShared * kTraceBlock];
    const uint64_t     i = static_cast<uint64_t>(blockIdx.x) * kTraceBlock + threadIdx.x;
    if (i < n) {
        float            t_max, beta, gamma;
        const Ray        r = load_ray(rays, i, t_max);
        const LightPrims lp{ s.lights };
        const int32_t    id = closest_hit<false>(s.lights_accel, lp, r, t_max, beta, gamma, stack + threadIdx.x, nullptr);
        hits[i]             = spcu_hit{ id, t_max };
    }
}

// Material-sorted hand-over: lanes of a warp that share a segment reserve their places with ONE atomic per segment.
__device__ __forceinline__ void segment_push(const SortedQueue& q, uint32_t seg, uint32_t slot, bool active)
{
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (!active) {
        return;
    }
    const unsigned peers  = __match_any_sync(act, seg);
    const int      lane   = threadIdx.x & 31;
    const int      leader = __ffs(peers) - 1;
    uint32_t       base   = 0;
    if (lane == leader) {
        base = atomicAdd(q.counts + seg, static_cast<uint32_t>(__popc(peers)));
    }
    base = __shfl_sync(peers, base, leader);
    q.slots[static_cast<size_t>(seg) * q.capacity + base + __popc(peers & ((1u << lane) - 1u))] = slot;
}

// ---- wavefront stages ----------------------------------------------------------------------------------------------
// Work distribution of the persistent traversal stages.  Secondary rays have a violently skewed cost (bunny scene, random
// rays: median 1 node visit, mean 15, p99 121, max 438): a warp that takes 32 rays and waits for its slowest runs at 3-4
// active lanes (ncu, profiles/r01e_*).  Here lanes are refilled individually: a lane whose walk has ended draws the next
// queue entry while the other lanes keep walking, so long rays pile up side by side in a warp instead of each stalling 31
// finished lanes.  Entries are handed out from ONE global cursor in chunks of kChunk per warp (one global atomic per
// chunk), so the whole grid drains the queue evenly and no CTA is left with a private tail.
constexpr int      kLeavesPerRound = 2;  // leaf visits a lane may make before the warp looks for idle lanes again
constexpr uint32_t kChunk          = 64; // queue entries a warp reserves at a time
#ifndef SPCU_REFILL_MIN
#define SPCU_REFILL_MIN 8
#endif
constexpr int      kRefillMin      = SPCU_REFILL_MIN; // idle lanes that make a refill worth its set-up code
#ifndef SPCU_WALK_REFILL_MIN
#define SPCU_WALK_REFILL_MIN 12
#endif
// ... in the begin / walk kernels, where the phase also retires the finished walks.  4 until the walks stopped paying a
// global atomic per finished ray; re-swept with four batches in flight: 2 / 4 / 6 / 8 / 12 / 16 / 20 / 24 / 28 give 451 / 466 /
// 473 / 479 / 482 / 483 / 479 / 475 / 466 Mpaths/s on bunny and 676 / 693 / 697 / 704 / 706-712 / 706 / 700 / 700 / 694 on elf
// (profiles/r04b_*, r04c_*): a phase at 4 of 32 lanes every few steps costs more than lanes waiting for a fuller one.
constexpr int      kWalkRefillMin  = SPCU_WALK_REFILL_MIN;
// phase vote of the walk kernels: a pair leaf step runs when  pairs * NUM > lanes_at_a_node * DEN  (tuned on the GPU: profiles/)
// tuning / A-B switches of the walk kernels (make EXTRA=-D...; defaults are what measured best, profiles/r02*)
#ifndef SPCU_WALK_MIN_BLOCKS
#define SPCU_WALK_MIN_BLOCKS 8 // resident CTAs per SM the walk kernels are compiled for (register cap = 65536 / (128 * this))
#endif
#ifndef SPCU_LEAF_VOTE_NUM
#define SPCU_LEAF_VOTE_NUM 1
#endif
#ifndef SPCU_LEAF_VOTE_DEN
#define SPCU_LEAF_VOTE_DEN 1
#endif

struct LaneFeed
{
    uint32_t* cursor; // global, zero at kernel start
    uint32_t  n;
    uint32_t  next = 0, end = 0; // this warp's current chunk (identical in all lanes)

    // all 32 lanes call; lanes with want == true receive an index (0xffffffff: the queue is exhausted)
    __device__ __forceinline__ uint32_t draw(bool want)
    {
        const unsigned mask = __ballot_sync(0xffffffffu, want);
        if (mask == 0u) {
            return 0xffffffffu;
        }
        const int      lane = threadIdx.x & 31;
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        const uint32_t k    = __popc(mask);
        uint32_t       idx  = 0xffffffffu;
        const uint32_t have = end - next;
        if (want && rank < have) {
            idx = next + rank;
        }
        if (k <= have) {
            next += k;
            return idx;
        }
        // chunk used up: reserve the next one (k - have <= 32 <= kChunk entries are still owed)
        uint32_t base = 0;
        if (lane == 0) {
            base = atomicAdd(cursor, kChunk);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (want && rank >= have) {
            const uint32_t i = base + (rank - have);
            idx              = i < n ? i : 0xffffffffu;
        }
        next = min(n, base + (k - have));
        end  = min(n, base + kChunk);
        return idx;
    }
};

// per-segment push from divergent code: the lanes that arrive together (converged) aggregate among themselves
__device__ __forceinline__ void segment_push_converged(const SortedQueue& q, uint32_t seg, uint32_t slot)
{
    const unsigned act    = __activemask();
    const unsigned peers  = __match_any_sync(act, seg);
    const int      lane   = threadIdx.x & 31;
    const int      leader = __ffs(peers) - 1;
    uint32_t       base   = 0;
    if (lane == leader) {
        base = atomicAdd(q.counts + seg, static_cast<uint32_t>(__popc(peers)));
    }
    base = __shfl_sync(peers, base, leader);
    q.slots[static_cast<size_t>(seg) * q.capacity + base + __popc(peers & ((1u << lane) - 1u))] = slot;
}

// extend: Integrator.cpp:558-563.  intersect_lights first; a light hit shrinks t_max for the geometry query.
template <bool kCount, bool kOrdered, typename F>
__global__ void __launch_bounds__(kTraceBlock) k_extend(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                        const uint32_t* queue, const uint32_t* n_queue, uint32_t* cursor,
                                                        const __grid_constant__ SortedQueue sorted,
                                                        unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t stack_smem[kStackShared * kTraceBlock];
    const uint32_t     n = *n_queue;
    count_items(counters, kStExtend, n);
    LaneFeed feed;
    feed.cursor = cursor;
    feed.n      = n;

    TraceCounters   local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    Stack           stack;
    OrderedStack    ostack;
    stack.sh = stack_smem + threadIdx.x;
    ostack.attach(stack_smem + threadIdx.x);
    Ray         r{};
    RayInv      inv{};
    ClosestWalk walk{};
    walk.link       = kDone;
    uint32_t slot   = 0;
    unsigned traced = 0;
    bool     have   = false;
    int32_t  light_id = -1;
    float    light_t  = 0.0f;

    bool drained = false; // the queue has nothing left to hand out (warp uniform)
    for (;;) {
        // ---- phase vote: what do the lanes of this warp hold? ----------------------------------------------------------
        const int n_node = __popc(__ballot_sync(0xffffffffu, have && at_node(walk)));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && at_leaf(walk)));
        const int n_idle = 32 - __popc(__ballot_sync(0xffffffffu, have));
        // ---- refill: lanes without a ray draw the next queue entries, once enough of them wait (the set-up — ray record,
        // lights accelerator, unbounded primitives — is itself a hundred instructions: better run it for many lanes) ------
        if (!drained && (n_idle >= kRefillMin || n_node + n_leaf == 0)) {
            const uint32_t i = feed.draw(!have);
            drained          = __ballot_sync(0xffffffffu, !have && i == 0xffffffffu) != 0u;
            if (!have && i != 0xffffffffu) {
                slot            = queue[i];
                const RayRec rr = w.ray[slot];
                const float4 o = rr.o, d = rr.d;
                r               = Ray{ o.x, o.y, o.z, d.x, d.y, d.z, o.w };
                inv            = make_inv(r, s.geom.proper_boxes != 0u);
                float t_max = d.w, beta, gamma;
                // Scene::intersect_lights (a handful of lights: walked in one go)
                const LightPrimsT<F> lp{ s.lights };
                const int32_t    li = closest_hit<false>(s.lights_accel, lp, r, t_max, beta, gamma, stack_smem + threadIdx.x, nullptr);
                light_id            = li;
                light_t             = t_max;
                walk.t_max          = t_max;
                closest_begin<kCount>(s.geom, gp, r, walk, &local);
                stack.n = ostack.n = 0;
                have               = true;
                ++traced;
            }
        } else if (n_node + n_leaf == 0) {
            break; // nothing in flight and nothing left to draw
        } else if (kOrdered) {
            // ---- a bounded piece of every lane's walk (lanes without a ray vote along with a finished cursor) -----------
            closest_run_ordered<kCount>(s.geom, gp, r, inv, walk, ostack, kLeavesPerRound, &local, 0xffffffffu);
        } else if (!F::bvh) {
            closest_run<kCount>(s.geom, gp, r, inv, walk, stack, kLeavesPerRound, &local, 0xffffffffu);
        } else if (n_node >= n_leaf) {
            // ---- ONE step of the phase most lanes are in: an internal node (both child boxes) ... ---------------------------
            if (have && at_node(walk)) {
                closest_node_step<kCount>(s.geom, r, inv, walk, stack, &local);
            }
        } else {
            // ---- ... or a leaf (its primitives in list order) --------------------------------------------------------------
            const unsigned leaf_mask = __ballot_sync(0xffffffffu, have && at_leaf(walk));
            if (have && at_leaf(walk)) {
                closest_leaf_step<kCount>(gp, r, walk, stack, &local, leaf_mask);
            }
        }
        if (have && walk.link == kDone) {
            ExtendRec ex;
            ex.hit      = HitRec{ walk.hit_id, walk.t_max, walk.beta, walk.gamma };
            ex.light    = light_id;
            ex.light_t  = light_t;
            ex.pad[0] = ex.pad[1] = 0.0f;
            w.extend[slot] = ex;
            // hand the vertex to the shading stage sorted by material: misses in the last segment
            const uint32_t seg = walk.hit_id < 0 ? sorted.n_segments - 1u
                                                 : min(SPCU_META_MATERIAL(__ldg(s.geom_meta + walk.hit_id)), sorted.n_segments - 2u);
            segment_push_converged(sorted, seg, slot);
            have = false;
        }
    }
    const unsigned total = __reduce_add_sync(0xffffffffu, traced);
    if ((threadIdx.x & 31) == 0 && total) {
        atomicAdd(counters + kCntRaysClosest, static_cast<unsigned long long>(total));
        atomicAdd(counters + kCntRaysLights, static_cast<unsigned long long>(total));
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-kernel form of the traversal stages for scenes with a BVH (exact walk): `begin` + `walk`.
//
// Half of the rays of a path-traced frame never get past the root of the BVH (bunny scene: median 1 node visit), and
// every ray needs the same hundred instructions of set-up before its walk (ray record, lights accelerator, unbounded
// primitives).  In the persistent walk kernel those were run for whichever few lanes happened to be idle.  `begin` runs
// them for ALL rays at full warps — set-up, then the root node's two child boxes — finishes the rays that end there and
// parks the others (cursor, leaf count, "root's right child pending") in the two spare words of their ExtendRec; `walk`
// is the persistent phase-voted kernel, whose refill is now three loads.
//
// Parked cursor: link with bit 30 flipped when the root is on the stack (node and primitive indices stay below 2^30,
// checked at upload).
// ---------------------------------------------------------------------------------------------------------------------
// The `advance` stage of one path (shade_kernels.cu k_advance, Integrator.cpp:601-626) for the fused advance + begin kernel:
// false when Russian roulette ends the path, otherwise the new throughput and the next segment's ray record are written
// (and `rr` returns the latter).
__device__ __forceinline__ bool advance_path(const DScene& s, const DWave& w, const AdvanceArgs& adv, uint32_t slot, RayRec& rr)
{
    const SampleRec s0     = w.s0[slot];
    const VertexRec vx     = w.vertex[slot];
    const PathRec   pr     = w.path[slot];
    const float     cosine = fabsf(s0.dir.x * vx.n.x + s0.dir.y * vx.n.y + s0.dir.z * vx.n.z);
    float           tx = pr.tp.x * (cosine * s0.col.x / s0.dir.w), ty = pr.tp.y * (cosine * s0.col.y / s0.dir.w),
                    tz = pr.tp.z * (cosine * s0.col.z / s0.dir.w);
    if (adv.depth >= s.rr_depth) {
        const float lum = 0.2126f * tx + 0.7152f * ty + 0.0722f * tz; // math/RGB.h:224
        if (lum < 0.1f) {
            const float q = (0.05f < lum / 0.1f) ? lum / 0.1f : 0.05f; // probability of continuing
            Rng rng{ __float_as_uint(pr.tp.w),          __float_as_uint(pr.L.w),           static_cast<uint32_t>(adv.seed),
                     static_cast<uint32_t>(adv.seed >> 32), rng_stream(adv.depth, kSiteRoulette), 0u };
            if (!(rng_next1(rng) < q)) {
                return false;
            }
            tx = tx / q;
            ty = ty / q;
            tz = tz / q;
        }
    }
    w.path[slot].tp = make_float4(tx, ty, tz, pr.tp.w);
    rr              = RayRec{ make_float4(vx.p.x, vx.p.y, vx.p.z, cosine == 0.0f ? kRayEpsilon : kRayEpsilon / cosine),
                              make_float4(s0.dir.x, s0.dir.y, s0.dir.z, kInfinite) };
    w.ray[slot]     = rr;
    return true;
}

constexpr int32_t kPendingBit = 0x40000000;
__device__ __forceinline__ int32_t park_link(int32_t link, bool root_pending) { return root_pending ? (link ^ kPendingBit) : link; }
__device__ __forceinline__ int32_t unpark_link(int32_t parked, bool& root_pending)
{
    root_pending = parked >= 0 ? (parked & kPendingBit) != 0 : (parked & kPendingBit) == 0;
    return root_pending ? (parked ^ kPendingBit) : parked;
}

// kAdvance: the kernel ALSO is the `advance` stage of the previous vertex (Integrator.cpp:601-626: throughput update, Russian
// roulette, next segment).  `queue` is then the previous depth's queue of live vertices; a path that survives gets its
// throughput and its ray record written and goes on into the set-up below with the ray still in registers — one kernel, one
// pass over the records, instead of advance -> queue -> begin (3.7 + 5.8 % of a bunny frame, ncu launch list r02g).
#ifndef SPCU_BEGIN_MIN_BLOCKS
#define SPCU_BEGIN_MIN_BLOCKS 8 // resident CTAs per SM the `begin` kernels are compiled for
#endif
template <bool kCount, bool kOrdered, typename F, bool kAdvance = false>
__global__ void __launch_bounds__(kTraceBlock, SPCU_BEGIN_MIN_BLOCKS) k_extend_begin(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                              const uint32_t* queue, const uint32_t* n_queue,
                                                              uint32_t* q_walk, uint32_t* n_walk,
                                                              const __grid_constant__ SortedQueue sorted,
                                                              unsigned long long* counters, TraceCounters* cnt,
                                                              const __grid_constant__ AdvanceArgs adv)
{
    __shared__ int32_t stack_smem[kStackShared * kTraceBlock];
    const uint32_t     n = *n_queue;
    count_items(counters, kAdvance ? kStAdvance : kStExtend, n);
    TraceCounters       local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    for (uint32_t base = blockIdx.x * kTraceBlock; base < n; base += gridDim.x * kTraceBlock) {
        const uint32_t i      = base + threadIdx.x;
        bool           active = i < n;
        bool           done = false, parked = false;
        uint32_t       slot = 0, seg = 0;
        RayRec         rr{};
        if (active) {
            slot = queue[i];
            if constexpr (kAdvance) {
                active = advance_path(s, w, adv, slot, rr);
            } else {
                rr = w.ray[slot];
            }
        }
        if (kAdvance) {
            warp_count(counters + kNumCounters + kStExtend, active);
        }
        if (active) {
            const float4 o = rr.o, d = rr.d;
            const Ray    r{ o.x, o.y, o.z, d.x, d.y, d.z, o.w };
            float        t_max = d.w, beta, gamma;
            // Scene::intersect_lights (a handful of lights: walked in one go); a light hit shrinks t_max for the geometry
            const LightPrimsT<F> lp{ s.lights };
            const int32_t        li = closest_hit<false>(s.lights_accel, lp, r, t_max, beta, gamma, stack_smem + threadIdx.x, nullptr);
            ExtendRec            ex;
            ex.light   = li;
            ex.light_t = t_max;
            ClosestWalk walk;
            walk.t_max = t_max;
            closest_begin<kCount>(s.geom, gp, r, walk, &local); // unbounded primitives; cursor at the root
            Stack        stack;
            OrderedStack ostack;
            stack.sh = stack_smem + threadIdx.x;
            ostack.attach(stack_smem + threadIdx.x);
            const RayInv inv = kOrdered ? make_inv_wide(r, s.geom) : make_inv(r, s.geom.proper_boxes != 0u);
            if (at_node(walk)) {
                if (!kOrdered) { // the root's two child boxes; its right child may be left pending
                    closest_node_step<kCount>(s.geom, r, inv, walk, stack, &local);
                } else if (!inv.generic) { // the ordered walk starts over at the root: only "does it enter anything" is asked here
                    if (!wide_root_entered(s.geom, r, inv, walk.t_max)) {
                        walk.link = kDone;
                    }
                } else {
                    NodeHalf c0, c1;
                    load_node(s.geom.nodes, walk.link, c0, c1);
                    bool  h0, h1;
                    float e0, e1;
                    slab_pair(s.geom.nodes, walk.link, c0, c1, r, inv, walk.t_max, h0, h1, e0, e1);
                    if (!h0 && !h1) {
                        walk.link = kDone;
                    }
                }
            } else if (at_leaf(walk)) { // the root is a leaf
                if (kOrdered) {
                    closest_run_ordered<kCount>(s.geom, gp, r, inv, walk, ostack, kAllLeaves, &local, __activemask());
                } else {
                    closest_run<kCount>(s.geom, gp, r, inv, walk, stack, kAllLeaves, &local, __activemask());
                }
            }
            ex.hit = HitRec{ walk.hit_id, walk.t_max, walk.beta, walk.gamma };
            done   = walk.link == kDone;
            parked = !done;
            ex.pad[0] = __int_as_float(park_link(walk.link, !kOrdered && stack.n > 0));
            ex.pad[1] = __uint_as_float(walk.count);
            w.extend[slot] = ex;
            // finished vertices go to the shading stage sorted by material: misses in the last segment
            seg = walk.hit_id < 0 ? sorted.n_segments - 1u
                                  : min(SPCU_META_MATERIAL(__ldg(s.geom_meta + walk.hit_id)), sorted.n_segments - 2u);
        }
        segment_push(sorted, seg, slot, done);
        queue_push(q_walk, n_walk, slot, parked);
        warp_count(counters + kCntRaysClosest, active);
        warp_count(counters + kCntRaysLights, active);
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Leaf step with the warp's (ray, triangle) tests spread over all 32 lanes.
//
// In the per-lane leaf step a lane walks through its own leaf's triangles while every lane that is not at a leaf waits, and
// the loop runs to the LARGEST leaf among them: ncu on the bunny scene showed the triangle test at 5-9 of 32 lanes.  Here the
// lanes at a triangle leaf of up to kPairLeafMax primitives publish their (ray, triangle) PAIRS — an exclusive prefix sum of
// the leaf sizes gives every pair a lane —, every lane fetches the ray of its pair's owner by shuffle and runs ONE triangle
// test, and the owners then fold their results in list order.  The fold is the reference's sequence: ListAccelerator tests a
// leaf's primitives in order and a hit shrinks t_max for the next one (shapes/ListAccelerator.h:50-62); the triangle test
// depends on t_max only through its final `t > t_max` rejection (Triangle.h:141), so testing all of a leaf's triangles
// against the t_max of ENTRY and re-applying that one comparison in order gives the same accepted hit, bit for bit.
// Leaves that do not fit this round (more than 32 pairs in the warp) stay where they are and go first in the next one; mixed
// leaves (spheres among the bounded primitives) and leaves of more than kPairLeafMax primitives run the per-lane loop.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kPairLeafMax = 4; // k_max_leaf_elements (shapes/BVHAccelerator.h:211)

struct PairPlan
{
    unsigned part_mask; // lanes whose leaf is tested in this round
    uint32_t offset;    // participant: the lane that runs its first pair
    uint32_t my_n;      // participant: primitives of its leaf; 0 otherwise
    uint32_t total;     // pairs of this round, <= 32
};

__device__ __forceinline__ PairPlan plan_pairs(bool candidate, uint32_t n)
{
#ifndef SPCU_PAIR_SCAN_SHUFFLE // (A/B switch: the five-step shuffle scan)
    // exclusive prefix sum of the leaf sizes (0..kPairLeafMax = 4: three bits) from three ballots — three independent votes
    // instead of a chain of five dependent shuffles in front of every leaf step
    static_assert(kPairLeafMax < 8u, "the ballot scan sums three bits");
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    const uint32_t v  = candidate ? n : 0u;
    const unsigned b0 = __ballot_sync(0xffffffffu, (v & 1u) != 0u), b1 = __ballot_sync(0xffffffffu, (v & 2u) != 0u),
                   b2 = __ballot_sync(0xffffffffu, (v & 4u) != 0u);
    const uint32_t excl = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
    PairPlan       p;
    const bool     part = candidate && excl + v <= 32u; // a prefix of the candidates, in lane order
    p.part_mask         = __ballot_sync(0xffffffffu, part);
    p.my_n              = part ? n : 0u;
    p.offset            = excl;
    p.total             = __popc(b0 & p.part_mask) + 2u * __popc(b1 & p.part_mask) + 4u * __popc(b2 & p.part_mask);
    return p;
#else
    const int lane = threadIdx.x & 31;
    uint32_t  incl = candidate ? n : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) {
            incl += v;
        }
    }
    PairPlan   p;
    const bool part = candidate && incl <= 32u; // a prefix of the candidates, in lane order
    p.part_mask     = __ballot_sync(0xffffffffu, part);
    p.my_n          = part ? n : 0u;
    p.offset        = incl - (candidate ? n : 0u);
    p.total         = p.part_mask ? __shfl_sync(0xffffffffu, incl, 31 - __clz(p.part_mask)) : 0u;
    return p;
#endif
}

// owner lane and index within the owner's leaf of the pair this lane runs (lanes >= total: garbage, not used)
__device__ __forceinline__ void pair_assignment(const PairPlan& p, uint8_t* tbl /* this warp's 32 bytes */, int& owner, uint32_t& k)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (uint32_t j = 0; j < kPairLeafMax; ++j) {
        if (j < p.my_n) {
            tbl[p.offset + j] = static_cast<uint8_t>(lane | (j << 5));
        }
    }
    __syncwarp();
    const uint8_t e = tbl[lane];
    __syncwarp();
    owner = e & 31;
    k     = e >> 5;
}

__device__ __forceinline__ Ray fetch_ray(const Ray& r, int owner)
{
    Ray o;
    o.ox    = __shfl_sync(0xffffffffu, r.ox, owner);
    o.oy    = __shfl_sync(0xffffffffu, r.oy, owner);
    o.oz    = __shfl_sync(0xffffffffu, r.oz, owner);
    o.dx    = __shfl_sync(0xffffffffu, r.dx, owner);
    o.dy    = __shfl_sync(0xffffffffu, r.dy, owner);
    o.dz    = __shfl_sync(0xffffffffu, r.dz, owner);
    o.t_min = __shfl_sync(0xffffffffu, r.t_min, owner);
    return o;
}

// the walk's next cursor after a leaf: the pending right child (exact walk, re-tested there) / the nearest deferred child
template <bool kOrdered, typename StackT>
__device__ __forceinline__ void leaf_pop(const DAccel& acc, StackT& stack, ClosestWalk& w, bool wide_keys)
{
    if constexpr (kOrdered) {
        if (wide_keys) {
            wide_ordered_pop(acc, stack, w);
        } else {
            ordered_pop(acc, stack, w);
        }
    } else {
        if (stack.n > 0) {
            w.link   = stack.pop();
            w.retest = true;
        } else {
            w.link = kDone;
        }
    }
}

// All 32 lanes call.  `mine` = this lane holds a ray whose cursor is at a leaf.
template <bool kCount, bool kOrdered, typename Prims, typename StackT>
__device__ __forceinline__ void closest_leaf_step_pairs(const DAccel& acc, const Prims& prims, const Ray& r, ClosestWalk& w, StackT& stack,
                                                        bool mine, bool wide_keys, uint8_t* tbl, TraceCounters* cnt)
{
    const int      lane      = threadIdx.x & 31;
    const uint32_t n         = w.count & SPCU_LEAF_COUNT_MASK;
    const bool     candidate = mine && !(w.count & SPCU_LEAF_MIXED_FLAG) && n <= kPairLeafMax;
    const PairPlan p         = plan_pairs(candidate, n);
    if (p.part_mask) {
        int      owner;
        uint32_t k;
        pair_assignment(p, tbl, owner, k);
        const Ray      pr    = fetch_ray(r, owner);
        const float    ptmax = __shfl_sync(0xffffffffu, w.t_max, owner);
        const uint32_t first = __shfl_sync(0xffffffffu, static_cast<uint32_t>(~w.link), owner);
        float          t = 0.0f, beta = 0.0f, gamma = 0.0f;
        bool           hit = false;
        if (static_cast<uint32_t>(lane) < p.total) {
            const uint32_t id = first + k;
            const float4   a  = __ldg(prims.prims + 3 * id + 0);
            const float4   b  = __ldg(prims.prims + 3 * id + 1);
            const float4   c  = __ldg(prims.prims + 3 * id + 2);
            if (kCount) ++cnt->tris;
            hit = tri_hit(a, b, c, pr, ptmax, t, beta, gamma);
        }
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        int            win  = -1;
#pragma unroll
        for (uint32_t j = 0; j < kPairLeafMax; ++j) { // the owners fold their leaf's results in list order
            const int   src = min(static_cast<int>(p.offset + j), 31);
            const float tj  = __shfl_sync(0xffffffffu, t, src);
            if (j < p.my_n && ((hits >> src) & 1u) && !(tj > w.t_max)) {
                const int32_t id = static_cast<int32_t>(static_cast<uint32_t>(~w.link) + j);
                if (!kOrdered || tj < w.t_max || id > w.hit_id) {
                    w.t_max  = tj;
                    w.hit_id = id;
                    win      = src;
                }
            }
        }
        const int   from = win >= 0 ? win : lane;
        const float bw = __shfl_sync(0xffffffffu, beta, from), gw = __shfl_sync(0xffffffffu, gamma, from);
        if (win >= 0) {
            w.beta  = bw;
            w.gamma = gw;
        }
        if ((p.part_mask >> lane) & 1u) { // participants (also those with an empty leaf) move on
            leaf_pop<kOrdered>(acc, stack, w, wide_keys);
        }
    }
    // leaves the pair scheme does not take: mixed or oversized (the per-lane loop, reference order)
    const bool     other      = mine && !candidate;
    const unsigned other_mask = __ballot_sync(0xffffffffu, other);
    if (other) {
        float          t, b, g;
        const uint32_t first = static_cast<uint32_t>(~w.link);
        const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
        const uint32_t n_max = __reduce_max_sync(other_mask, n);
#pragma unroll 1
        for (uint32_t i = 0; i < n_max; ++i) {
            const int32_t id = static_cast<int32_t>(first + i);
            if (i < n && prims.template test<kCount>(first + i, mixed, r, w.t_max, t, b, g, cnt) &&
                (!kOrdered || t < w.t_max || id > w.hit_id)) {
                w.t_max  = t;
                w.hit_id = id;
                w.beta   = b;
                w.gamma  = g;
            }
        }
        leaf_pop<kOrdered>(acc, stack, w, wide_keys);
    }
}

// Which phase the warp runs next: the one that puts more lanes to work.  A node step employs the lanes at a node, a pair
// leaf step min(32, triangles of the waiting leaves) lanes.
__device__ __forceinline__ bool vote_leaf_step(int n_node, bool at_leaf_lane, uint32_t leaf_n)
{
    const uint32_t pairs = __reduce_add_sync(0xffffffffu, at_leaf_lane ? min(leaf_n, kPairLeafMax) : 0u);
    return static_cast<int>(min(pairs, 32u)) * SPCU_LEAF_VOTE_NUM > n_node * SPCU_LEAF_VOTE_DEN;
}

// Persistent kernels give up (kCntErrors -> SPCU_ERR_INTERNAL on the host) instead of spinning for ever on a corrupted cursor.
constexpr uint32_t kWalkIterCap = 1u << 27;

template <bool kCount, bool kOrdered, typename F>
__global__ void __launch_bounds__(kTraceBlock, SPCU_WALK_MIN_BLOCKS) k_extend_walk(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                             const uint32_t* q_walk, const uint32_t* n_walk, uint32_t* cursor,
                                                             const __grid_constant__ SortedQueue sorted,
                                                             unsigned long long* counters, TraceCounters* cnt)
{
    // exact walk: 24 four-byte levels per thread; ordered walk: 16 eight-byte entries (wide nodes defer up to three per step)
    __shared__ int32_t stack_smem[(kOrdered ? 2 * kWideSharedOrdered : kStackShared) * kTraceBlock];
    __shared__ uint8_t pair_tbl[kTraceBlock];
    LaneFeed           feed;
    feed.cursor = cursor;
    feed.n      = *n_walk;

    TraceCounters       local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    using StackS = typename std::conditional<kOrdered, WideOrderedStack, Stack>::type;
    StackS stack;
    if constexpr (kOrdered) {
        stack.attach(stack_smem + threadIdx.x);
    } else {
        stack.sh = stack_smem + threadIdx.x;
    }
    uint8_t*    tbl = pair_tbl + (threadIdx.x & ~31);
    Ray         r{};
    RayInv      inv{};
    ClosestWalk walk{};
    walk.link     = kDone;
    uint32_t slot = 0, iters = 0;
    bool     have = false, drained = false;

    for (;;) {
        const int n_node = __popc(__ballot_sync(0xffffffffu, have && at_node(walk)));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && at_leaf(walk)));
        if (++iters > kWalkIterCap) {
            if ((threadIdx.x & 31) == 0) atomicAdd(counters + kCntErrors, 1ull);
            break;
        }
        if (n_node + n_leaf == 0 || (!drained && 32 - n_node - n_leaf >= kWalkRefillMin)) {
            // ---- retire + refill, together.  A finished walk is not written out the moment it ends (one lane, one global atomic
            // whose round trip stalled the whole warp: 15 % of the kernel's stall samples, ncu r02g) but here, with every other
            // lane that has finished since the last refill: one aggregated atomic per material segment, issued next to the loads
            // of the rays that take the freed lanes.
            const bool finished = have && walk.link == kDone;
            uint32_t   seg      = 0;
            if (finished) {
                w.extend[slot].hit = HitRec{ walk.hit_id, walk.t_max, walk.beta, walk.gamma };
                seg = walk.hit_id < 0 ? sorted.n_segments - 1u
                                      : min(SPCU_META_MATERIAL(__ldg(s.geom_meta + walk.hit_id)), sorted.n_segments - 2u);
                have = false;
            }
            segment_push(sorted, seg, slot, finished);
            if (drained) {
                break; // nothing in flight (the other clause needs !drained) and nothing left to draw
            }
            const uint32_t i = feed.draw(!have);
            drained          = __ballot_sync(0xffffffffu, !have && i == 0xffffffffu) != 0u;
            if (!have && i != 0xffffffffu) {
                slot               = q_walk[i];
                const RayRec    rr = w.ray[slot];
                const ExtendRec ex = w.extend[slot];
                r                  = Ray{ rr.o.x, rr.o.y, rr.o.z, rr.d.x, rr.d.y, rr.d.z, rr.o.w };
                inv                = kOrdered ? make_inv_wide(r, s.geom) : make_inv(r, s.geom.proper_boxes != 0u);
                bool root_pending;
                walk.link   = unpark_link(__float_as_int(ex.pad[0]), root_pending);
                walk.count  = __float_as_uint(ex.pad[1]);
                walk.retest = false;
                walk.hit_id = ex.hit.id;
                walk.t_max  = ex.hit.t;
                walk.beta   = ex.hit.beta;
                walk.gamma  = ex.hit.gamma;
                stack.n     = 0;
                if constexpr (!kOrdered) {
                    if (root_pending) {
                        stack.push(s.geom.root);
                    }
                }
                have = true;
            }
        } else if (n_node > 0 && !vote_leaf_step(n_node, have && at_leaf(walk), walk.count & SPCU_LEAF_COUNT_MASK)) {
            if (have && at_node(walk)) {
                if constexpr (!kOrdered) {
                    closest_node_step<kCount>(s.geom, r, inv, walk, stack, &local);
                } else if (!inv.generic) {
                    closest_wide_step_ordered<kCount>(s.geom, r, inv, walk, stack, &local);
                    prefetch_leaf(s.geom_prims, walk.link, walk.count);
                } else {
                    closest_node_step_ordered<kCount>(s.geom, r, inv, walk, stack, &local);
                }
            }
        } else {
            closest_leaf_step_pairs<kCount, kOrdered>(s.geom, gp, r, walk, stack, have && at_leaf(walk), kOrdered && !inv.generic, tbl, &local);
        }
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// shadow_begin: the visibility ray's set-up, the unbounded primitives, the lights accelerator (Scene::intersect_p is an
// OR over both accelerators: base/Scene.h:79-82, so their order is free) and the root's two child boxes.
// kMis: the same two kernels serve the BSDF-strategy ray of the NEE stage (Integrator.cpp:527-532): intersect_lights first,
// and — only when the ray reaches a light — Scene::intersect_p with the UNSHRUNK limits (t_max stays FLT_MAX: a sphere light
// therefore occludes itself, as in the reference); the answer goes to MisRec::light / MisRec::occluded.
template <bool kCount, typename F, bool kMis = false>
__global__ void __launch_bounds__(kTraceBlock, SPCU_BEGIN_MIN_BLOCKS) k_shadow_begin(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                              const uint32_t* queue, const uint32_t* n_queue, uint32_t light_index,
                                                              uint32_t* q_walk, uint32_t* n_walk, uint32_t* q_lit, uint32_t* n_lit,
                                                              unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t stack_smem[kStackShared * kTraceBlock];
    const uint32_t     n = *n_queue;
    count_items(counters, kMis ? kStMisTrace : kStShadow, n);
    TraceCounters       local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    for (uint32_t base = blockIdx.x * kTraceBlock; base < n; base += gridDim.x * kTraceBlock) {
        const uint32_t i      = base + threadIdx.x;
        const bool     active = i < n;
        bool           lit = false, parked = false, traced = !kMis;
        uint32_t       slot = 0;
        if (active) {
            slot            = queue[i];
            const float4 p  = w.vertex[slot].p;
            float4       d;   // direction xyz; w = t_max (shadow) / t_min (mis)
            float        t_min, t_max;
            if (kMis) {
                d     = w.mis[slot].d;
                t_min = d.w;
                t_max = kInfinite;
            } else {
                const LightRec lr = w.light[static_cast<size_t>(light_index) * w.capacity + slot];
                d     = lr.wi;
                t_min = lr.aux.x;
                t_max = lr.wi.w;
            }
            const Ray r{ p.x, p.y, p.z, d.x, d.y, d.z, t_min };
            bool      hit = false;
            if (kMis) { // Scene::intersect_lights; without a light the vertex gets nothing from this strategy and no occlusion query
                float                beta, gamma, t_light = kInfinite;
                const LightPrimsT<F> lp{ s.lights };
                const int32_t li = closest_hit<false>(s.lights_accel, lp, r, t_light, beta, gamma, stack_smem + threadIdx.x, nullptr);
                w.mis[slot].light = li;
                traced            = li >= 0;
            }
            // ListAccelerator::intersect_p_impl: unbounded primitives first
            for (uint32_t k = 0; traced && k < s.geom.n_unbounded && !hit; ++k) {
                float t, b, g;
                hit = gp.template test<kCount>(k, true, r, t_max, t, b, g, &local);
            }
            if (traced && !hit) {
                hit = lights_any_hit<F>(s, r, t_max, stack_smem + threadIdx.x);
            }
            AnyWalk walk{ kDone, 0 };
            Stack   stack;
            stack.sh = stack_smem + threadIdx.x;
            if (traced && !hit) {
                walk = AnyWalk{ s.geom.root, s.geom.root_count };
                const RayInv inv = make_inv_wide(r, s.geom);
                if (at_node(walk)) { // the walk starts over at the root: only "does the ray enter anything" is asked here
                    bool entered;
                    if (!inv.generic) {
                        entered = wide_root_entered(s.geom, r, inv, t_max);
                    } else {
                        NodeHalf c0, c1;
                        load_node(s.geom.nodes, walk.link, c0, c1);
                        bool  h0, h1;
                        float e0, e1;
                        slab_pair(s.geom.nodes, walk.link, c0, c1, r, inv, t_max, h0, h1, e0, e1);
                        entered = h0 || h1;
                    }
                    if (!entered) {
                        walk.link = kDone;
                    }
                } else if (at_leaf(walk)) {
                    auto geom_test = [&](uint32_t id, bool mixed, TraceCounters* c) {
                        float t, b, g;
                        return gp.template test<kCount>(id, mixed, r, t_max, t, b, g, c);
                    };
                    hit = any_run<kCount, true>(s.geom, geom_test, r, inv, t_max, walk, stack, kAllLeaves, &local, __activemask()) == kAnyHit;
                    walk.link = kDone;
                }
            }
            parked = !hit && walk.link != kDone;
            lit    = !hit && !parked;
            if (parked) {
                w.extend[slot].pad[0] = __int_as_float(park_link(walk.link, false));
                w.extend[slot].pad[1] = __uint_as_float(walk.count);
            } else if (kMis) {
                w.mis[slot].occluded = hit ? 1 : 0;
            } else if (!q_lit) {
                w.occluded[slot] = hit ? 1 : 0; // direct lighting reads the flag; the NEE path gets the compacted queue
            }
        }
        if (!kMis && q_lit) { // survivors only go on to the BSDF stages (Integrator.cpp:503-506)
            queue_push(q_lit, n_lit, slot, lit);
        }
        queue_push(q_walk, n_walk, slot, parked);
        if (kMis) {
            warp_count(counters + kCntRaysLights, active);
        }
        warp_count(counters + kCntRaysAny, active && traced);
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// Any-hit leaf step with the pairs spread over the warp (see closest_leaf_step_pairs): the limits of an any-hit query never
// change and its answer is "some primitive of a reachable leaf accepts", so the order of the tests is free.  Returns true
// for the lanes whose leaf holds an accepted primitive (their query is over).
template <bool kCount, typename Prims>
__device__ __forceinline__ bool any_leaf_step_pairs(const DAccel& acc, const Prims& prims, const Ray& r, float t_max, AnyWalk& w,
                                                    WideStack& stack, bool mine, bool wide_keys, uint8_t* tbl, TraceCounters* cnt)
{
    const int      lane      = threadIdx.x & 31;
    const uint32_t n         = w.count & SPCU_LEAF_COUNT_MASK;
    const bool     candidate = mine && !(w.count & SPCU_LEAF_MIXED_FLAG) && n <= kPairLeafMax;
    const PairPlan p         = plan_pairs(candidate, n);
    bool           found     = false;
    if (p.part_mask) {
        int      owner;
        uint32_t k;
        pair_assignment(p, tbl, owner, k);
        const Ray      pr    = fetch_ray(r, owner);
        const float    ptmax = __shfl_sync(0xffffffffu, t_max, owner);
        const uint32_t first = __shfl_sync(0xffffffffu, static_cast<uint32_t>(~w.link), owner);
        bool           hit   = false;
        if (static_cast<uint32_t>(lane) < p.total) {
            const uint32_t id = first + k;
            const float4   a  = __ldg(prims.prims + 3 * id + 0);
            const float4   b  = __ldg(prims.prims + 3 * id + 1);
            const float4   c  = __ldg(prims.prims + 3 * id + 2);
            float          t, beta, gamma;
            if (kCount) ++cnt->tris;
            hit = tri_hit(a, b, c, pr, ptmax, t, beta, gamma);
        }
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        if ((p.part_mask >> lane) & 1u) {
            const unsigned my_lanes = p.my_n ? (((1u << p.my_n) - 1u) << p.offset) : 0u; // my_n <= 4, offset + my_n <= 32
            found                   = (hits & my_lanes) != 0u;
            if (found) {
                w.link = kDone;
            } else if (wide_keys) {
                wide_any_pop(acc, stack, w);
            } else {
                any_pop(acc, stack, w);
            }
        }
    }
    const bool     other      = mine && !candidate;
    const unsigned other_mask = __ballot_sync(0xffffffffu, other);
    if (other) {
        auto geom_test = [&](uint32_t id, bool mixed, TraceCounters* c) {
            float t, b, g;
            return prims.template test<kCount>(id, mixed, r, t_max, t, b, g, c);
        };
        found = any_leaf_step(acc, geom_test, w, stack, cnt, other_mask, wide_keys);
    }
    return found;
}

template <bool kCount, typename F, bool kMis = false>
__global__ void __launch_bounds__(kTraceBlock, SPCU_WALK_MIN_BLOCKS) k_shadow_walk(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                             const uint32_t* q_walk, const uint32_t* n_walk, uint32_t light_index,
                                                             uint32_t* cursor, uint32_t* q_lit, uint32_t* n_lit,
                                                             unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t stack_smem[kWideSharedAny * kTraceBlock];
    __shared__ uint8_t pair_tbl[kTraceBlock];
    LaneFeed           feed;
    feed.cursor = cursor;
    feed.n      = *n_walk;

    TraceCounters       local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    WideStack           stack;
    stack.sh = stack_smem + threadIdx.x;
    uint8_t* tbl = pair_tbl + (threadIdx.x & ~31);
    Ray      r{};
    RayInv   inv{};
    AnyWalk  walk{ kDone, 0 };
    float    t_max = 0.0f;
    uint32_t slot = 0, iters = 0;
    bool     have = false, drained = false, hit = false;

    for (;;) {
        const int n_node = __popc(__ballot_sync(0xffffffffu, have && at_node(walk)));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && at_leaf(walk)));
        if (++iters > kWalkIterCap) {
            if ((threadIdx.x & 31) == 0) atomicAdd(counters + kCntErrors, 1ull);
            break;
        }
        if (n_node + n_leaf == 0 || (!drained && 32 - n_node - n_leaf >= kWalkRefillMin)) {
            // ---- retire + refill, together (see k_extend_walk)
            const bool finished = have && walk.link == kDone;
            if (finished) {
                if (kMis) {
                    w.mis[slot].occluded = hit ? 1 : 0;
                } else if (!q_lit) {
                    w.occluded[slot] = hit ? 1 : 0;
                }
                have = false;
            }
            if (!kMis && q_lit) { // survivors only go on to the BSDF stages (Integrator.cpp:503-506)
                queue_push(q_lit, n_lit, slot, finished && !hit);
            }
            if (drained) {
                break;
            }
            const uint32_t i = feed.draw(!have);
            drained          = __ballot_sync(0xffffffffu, !have && i == 0xffffffffu) != 0u;
            if (!have && i != 0xffffffffu) {
                slot               = q_walk[i];
                const float4    p  = w.vertex[slot].p;
                const ExtendRec ex = w.extend[slot];
                if (kMis) {
                    const float4 d = w.mis[slot].d;
                    r              = Ray{ p.x, p.y, p.z, d.x, d.y, d.z, d.w };
                    t_max          = kInfinite;
                } else {
                    const LightRec lr = w.light[static_cast<size_t>(light_index) * w.capacity + slot];
                    r                 = Ray{ p.x, p.y, p.z, lr.wi.x, lr.wi.y, lr.wi.z, lr.aux.x };
                    t_max             = lr.wi.w;
                }
                inv = make_inv_wide(r, s.geom);
                bool root_pending;
                walk.link  = unpark_link(__float_as_int(ex.pad[0]), root_pending);
                walk.count = __float_as_uint(ex.pad[1]);
                stack.n    = 0;
                have = true;
                hit  = false;
            }
        } else if (n_node > 0 && !vote_leaf_step(n_node, have && at_leaf(walk), walk.count & SPCU_LEAF_COUNT_MASK)) {
            if (have && at_node(walk)) {
                if (!inv.generic) {
                    any_wide_step<kCount>(s.geom, r, inv, t_max, walk, stack, &local);
                    prefetch_leaf(s.geom_prims, walk.link, walk.count);
                } else {
                    any_node_step<kCount>(s.geom, r, inv, t_max, walk, stack, &local);
                }
            }
        } else {
            const bool found = any_leaf_step_pairs<kCount>(s.geom, gp, r, t_max, walk, stack, have && at_leaf(walk), !inv.generic, tbl, &local);
            hit              = hit || found;
        }
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// shadow: Integrator.cpp:503 — Scene::intersect_p of the light sample's visibility ray.
template <bool kCount, typename F>
__global__ void __launch_bounds__(kTraceBlock) k_shadow(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                        const uint32_t* queue, const uint32_t* n_queue, uint32_t light_index,
                                                        uint32_t* cursor, uint32_t* q_lit, uint32_t* n_lit,
                                                        unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t stack_smem[kStackShared * kTraceBlock];
    const uint32_t     n = *n_queue;
    count_items(counters, kStShadow, n);
    LaneFeed feed;
    feed.cursor = cursor;
    feed.n      = n;

    TraceCounters   local{ 0, 0, 0 };
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    Stack           stack;
    stack.sh = stack_smem + threadIdx.x;
    Ray      r{};
    RayInv   inv{};
    AnyWalk  walk{ kDone, 0 };
    float    t_max  = 0.0f;
    uint32_t slot   = 0;
    unsigned traced = 0;
    bool     have   = false;
    int      status = kAnyMiss;

    bool drained = false;
    auto geom_test = [&](uint32_t id, bool mixed, TraceCounters* c) {
        float t, b, g;
        return gp.template test<kCount>(id, mixed, r, t_max, t, b, g, c);
    };
    for (;;) {
        const int n_node = __popc(__ballot_sync(0xffffffffu, have && at_node(walk)));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && at_leaf(walk)));
        const int n_idle = 32 - __popc(__ballot_sync(0xffffffffu, have));
        if (!drained && (n_idle >= kRefillMin || n_node + n_leaf == 0)) {
            const uint32_t i = feed.draw(!have);
            drained          = __ballot_sync(0xffffffffu, !have && i == 0xffffffffu) != 0u;
            if (!have && i != 0xffffffffu) {
                slot              = queue[i];
                const float4   p  = w.vertex[slot].p;
                const LightRec lr = w.light[static_cast<size_t>(light_index) * w.capacity + slot];
                r                 = Ray{ p.x, p.y, p.z, lr.wi.x, lr.wi.y, lr.wi.z, lr.aux.x };
                inv               = make_inv(r, s.geom.proper_boxes != 0u);
                t_max             = lr.wi.w;
                stack.n        = 0;
                walk           = AnyWalk{ s.geom.root, s.geom.root_count };
                have           = true;
                status         = kAnyRunning;
                ++traced;
                // ListAccelerator::intersect_p_impl: unbounded primitives first
                for (uint32_t k = 0; k < s.geom.n_unbounded; ++k) {
                    float t, b, g;
                    if (gp.template test<kCount>(k, true, r, t_max, t, b, g, &local)) {
                        status    = kAnyHit;
                        walk.link = kDone;
                        break;
                    }
                }
            }
        } else if (n_node + n_leaf == 0) {
            break;
        } else if (!F::bvh) {
            const int st = any_run<kCount, false>(s.geom, geom_test, r, inv, t_max, walk, stack, kLeavesPerRound, &local, 0xffffffffu);
            if (have && status == kAnyRunning) {
                status = st;
            }
        } else if (n_node >= n_leaf) {
            if (have && at_node(walk)) {
                any_node_step<kCount>(s.geom, r, inv, t_max, walk, stack, &local);
            }
        } else {
            const unsigned leaf_mask = __ballot_sync(0xffffffffu, have && at_leaf(walk));
            if (have && at_leaf(walk)) {
                if (any_leaf_step(s.geom, geom_test, walk, stack, &local, leaf_mask)) {
                    status = kAnyHit;
                }
            }
        }
        if (have && status == kAnyRunning && walk.link == kDone) {
            status = kAnyMiss;
        }
        if (have && status != kAnyRunning) {
            // Scene::intersect_p: geometry, then the lights accelerator (base/Scene.h:79-82)
            if (status == kAnyMiss && lights_any_hit<F>(s, r, t_max, stack_smem + threadIdx.x)) {
                status = kAnyHit;
            }
            const bool occ = status == kAnyHit;
            if (!q_lit) {
                w.occluded[slot] = occ ? 1 : 0; // direct lighting reads the flag; the NEE path gets the compacted queue
            }
            if (q_lit && !occ) { // survivors only go on to the BSDF stages (Integrator.cpp:503-506)
                const unsigned act    = __activemask();
                const int      lane   = threadIdx.x & 31;
                const int      leader = __ffs(act) - 1;
                uint32_t       base   = 0;
                if (lane == leader) {
                    base = atomicAdd(n_lit, static_cast<uint32_t>(__popc(act)));
                }
                base = __shfl_sync(act, base, leader);
                q_lit[base + __popc(act & ((1u << lane) - 1u))] = slot;
            }
            have      = false;
            walk.link = kDone;
        }
    }
    const unsigned total = __reduce_add_sync(0xffffffffu, traced);
    if ((threadIdx.x & 31) == 0 && total) {
        atomicAdd(counters + kCntRaysAny, static_cast<unsigned long long>(total));
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

// mis: Integrator.cpp:527-532 — intersect_lights of the BSDF-sampled ray and, when it reaches a light, intersect_p
// with the SAME limits (t_max stays FLT_MAX: a sphere light therefore occludes itself, as in the reference).
template <bool kCount, typename F>
__global__ void __launch_bounds__(kTraceBlock) k_mis_trace(const __grid_constant__ DScene s, const __grid_constant__ DWave w,
                                                           const uint32_t* queue, const uint32_t* n_queue,
                                                           unsigned long long* counters, TraceCounters* cnt)
{
    __shared__ int32_t stack[kStackShared * kTraceBlock];
    const uint32_t     n = *n_queue;
    count_items(counters, kStMisTrace, n);
    TraceCounters      local{ 0, 0, 0 };
    for (uint32_t base = blockIdx.x * kTraceBlock; base < n; base += gridDim.x * kTraceBlock) {
    const uint32_t i          = base + threadIdx.x;
    const bool     active     = i < n;
    bool           traced_any = false;
    if (active) {
        const uint32_t slot = queue[i];
        const float4   p    = w.vertex[slot].p;
        const float4   d    = w.mis[slot].d;
        const Ray      r{ p.x, p.y, p.z, d.x, d.y, d.z, d.w };
        float          t_max = kInfinite, beta, gamma;

        const LightPrimsT<F> lp{ s.lights };
        const int32_t    li  = closest_hit<false>(s.lights_accel, lp, r, t_max, beta, gamma, stack + threadIdx.x, nullptr);
        int              occ = 0;
        if (li >= 0) {
            traced_any = true;
            occ        = scene_any_hit<kCount, F>(s, r, kInfinite, stack + threadIdx.x, &local) ? 1 : 0;
        }
        w.mis[slot].light    = li;
        w.mis[slot].occluded = occ;
    }
    warp_count(counters + kCntRaysLights, active);
    warp_count(counters + kCntRaysAny, traced_any);
    }
    if (kCount) {
        flush_counters(local, cnt);
    }
}

inline unsigned grid_for(uint64_t n)
{
    return static_cast<unsigned>((n + kTraceBlock - 1) / kTraceBlock);
}

// ---- ray batches through the wavefront's own stage kernels (spcu_extend_batch / spcu_shadow_batch) -------------------------
// slot i of the wavefront state <- ray i; queue = 0 .. n-1
__global__ void __launch_bounds__(kTraceBlock) k_batch_fill(const __grid_constant__ DWave w, const spcu_ray* rays, uint32_t n,
                                                            uint32_t* queue, uint32_t* n_queue, bool shadow)
{
    const uint32_t i = blockIdx.x * kTraceBlock + threadIdx.x;
    if (i == 0) {
        *n_queue = n;
    }
    if (i < n) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 0); // o, t_min
        const float4 b = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 1); // d, t_max
        queue[i]       = i;
        if (shadow) { // what shade leaves behind for a light sample: VertexRec::p, LightRec (light plane 0)
            w.vertex[i].p = make_float4(a.x, a.y, a.z, 0.0f);
            w.light[i]    = LightRec{ b, make_float4(a.w, 0.0f, 0.0f, 0.0f) };
        } else {
            w.ray[i] = RayRec{ a, b };
        }
    }
}

__global__ void __launch_bounds__(kTraceBlock) k_batch_gather_extend(const __grid_constant__ DWave w, uint32_t n, spcu_hit* hits,
                                                                     spcu_hit* light_hits)
{
    const uint32_t i = blockIdx.x * kTraceBlock + threadIdx.x;
    if (i < n) {
        const ExtendRec ex = w.extend[i];
        hits[i]            = spcu_hit{ ex.hit.id, ex.hit.t };
        if (light_hits) {
            light_hits[i] = spcu_hit{ ex.light, ex.light_t };
        }
    }
}

} // namespace

void launch_batch_fill(const DWave& w, const spcu_ray* d_rays, uint32_t n, uint32_t* queue, uint32_t* d_n_queue, bool shadow,
                       cudaStream_t st)
{
    k_batch_fill<<<std::max(1u, grid_for(n)), kTraceBlock, 0, st>>>(w, d_rays, n, queue, d_n_queue, shadow);
}

void launch_batch_gather_extend(const DWave& w, uint32_t n, spcu_hit* d_hits, spcu_hit* d_light_hits, cudaStream_t st)
{
    if (n == 0) return;
    k_batch_gather_extend<<<grid_for(n), kTraceBlock, 0, st>>>(w, n, d_hits, d_light_hits);
}

void launch_trace_closest(const DScene& s, const spcu_ray* d_rays, uint64_t n, spcu_hit* d_hits, TraceCounters* d_cnt,
                          cudaStream_t st)
{
    if (n == 0) return;
    if (d_cnt) {
        k_trace_closest<true><<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_hits, d_cnt);
    } else {
        k_trace_closest<false><<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_hits, nullptr);
    }
}

void launch_trace_closest_ordered(const DScene& s, const spcu_ray* d_rays, uint64_t n, spcu_hit* d_hits, TraceCounters* d_cnt,
                                  cudaStream_t st)
{
    if (n == 0) return;
    if (d_cnt) {
        k_trace_closest_ordered<true><<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_hits, d_cnt);
    } else {
        k_trace_closest_ordered<false><<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_hits, nullptr);
    }
}

void launch_trace_any(const DScene& s, const spcu_ray* d_rays, uint64_t n, uint8_t* d_out, cudaStream_t st)
{
    if (n == 0) return;
    k_trace_any<<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_out);
}

void launch_trace_lights(const DScene& s, const spcu_ray* d_rays, uint64_t n, spcu_hit* d_hits, cudaStream_t st)
{
    if (n == 0) return;
    k_trace_lights<<<grid_for(n), kTraceBlock, 0, st>>>(s, d_rays, n, d_hits);
}

// SPCU_DEBUG_SYNC=1: synchronise after every traversal launch and name the kernel that failed (stderr)
static void debug_sync(const Launch& l, const char* what)
{
    static const bool on = [] { const char* e = getenv("SPCU_DEBUG_SYNC"); return e && *e == '1'; }();
    if (on) {
        const cudaError_t e = cudaStreamSynchronize(l.stream);
        if (e != cudaSuccess) {
            fprintf(stderr, "SPCU_DEBUG_SYNC: %s failed: %s\n", what, cudaGetErrorString(e));
        }
    }
}

template <typename K>
static int trace_ctas_per_sm(K kernel)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kTraceBlock, 0) != cudaSuccess || n < 1) {
        n = 1;
    }
    return n;
}

// One instantiation per (counting, [ordered,] scene feature set); the launcher picks by the scene's feature set.
template <bool kCount, bool kOrdered, typename F>
static void launch_extend_variant(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue,
                                  const uint32_t* d_n_queue, uint32_t max_n, uint32_t* d_cursor, const SortedQueue& sorted,
                                  unsigned long long* d_counters, TraceCounters* d_cnt)
{
    static const int occ_ = trace_ctas_per_sm(k_extend<kCount, kOrdered, F>);
    k_extend<kCount, kOrdered, F><<<wavefront_grid(max_n, kTraceBlock, occ_, l.sm_count), kTraceBlock, 0, l.stream>>>(
        s, w, queue, d_n_queue, d_cursor, sorted, d_counters, d_cnt);
}

template <typename F>
static void launch_extend_features(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue,
                                   const uint32_t* d_n_queue, uint32_t max_n, uint32_t* d_cursor, const SortedQueue& sorted,
                                   bool ordered, unsigned long long* d_counters, TraceCounters* d_cnt)
{
    if (d_cnt) {
        ordered ? launch_extend_variant<true, true, F>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, d_counters, d_cnt)
                : launch_extend_variant<true, false, F>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, d_counters, d_cnt);
    } else {
        ordered ? launch_extend_variant<false, true, F>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, d_counters, nullptr)
                : launch_extend_variant<false, false, F>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, d_counters, nullptr);
    }
}

// begin + walk (scenes with a BVH)
template <bool kCount, bool kOrdered>
static void launch_extend_split(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue, const uint32_t* d_n_queue,
                                uint32_t max_n, uint32_t* d_cursor, const SortedQueue& sorted, uint32_t* q_walk, uint32_t* d_n_walk,
                                unsigned long long* d_counters, TraceCounters* d_cnt, const AdvanceArgs* adv)
{
    static const int occ_b = trace_ctas_per_sm(k_extend_begin<kCount, kOrdered, FeatFull>);
    static const int occ_w = trace_ctas_per_sm(k_extend_walk<kCount, kOrdered, FeatFull>);
    if constexpr (!kCount) {
        if (adv) { // `queue` = the previous depth's live vertices: advance + begin in one kernel
            static const int occ_a = trace_ctas_per_sm(k_extend_begin<false, kOrdered, FeatFull, true>);
            k_extend_begin<false, kOrdered, FeatFull, true><<<wavefront_grid(max_n, kTraceBlock, occ_a, l.sm_count), kTraceBlock, 0, l.stream>>>(
                s, w, queue, d_n_queue, q_walk, d_n_walk, sorted, d_counters, nullptr, *adv);
        }
    }
    if (kCount || !adv) {
        k_extend_begin<kCount, kOrdered, FeatFull><<<wavefront_grid(max_n, kTraceBlock, occ_b, l.sm_count), kTraceBlock, 0, l.stream>>>(
            s, w, queue, d_n_queue, q_walk, d_n_walk, sorted, d_counters, d_cnt, AdvanceArgs{});
    }
    debug_sync(l, "k_extend_begin");
    k_extend_walk<kCount, kOrdered, FeatFull><<<wavefront_grid(max_n, kTraceBlock, occ_w, l.sm_count), kTraceBlock, 0, l.stream>>>(
        s, w, q_walk, d_n_walk, d_cursor, sorted, d_counters, d_cnt);
    debug_sync(l, "k_extend_walk");
}

bool extend_fuses_advance(const Launch& l, const uint32_t* q_walk, const TraceCounters* d_cnt)
{
#ifdef SPCU_NO_FUSED_ADVANCE // (A/B builds)
    return false;
#else
    return l.features != FeatAnalytic::id && q_walk != nullptr && d_cnt == nullptr;
#endif
}

int launch_extend(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue, const uint32_t* d_n_queue,
                  uint32_t max_n, uint32_t* d_cursor, const SortedQueue& sorted, bool ordered, uint32_t* q_walk,
                  uint32_t* d_n_walk, unsigned long long* d_counters, TraceCounters* d_cnt, const AdvanceArgs* adv)
{
    if (max_n == 0) return 0;
    if (l.features == FeatAnalytic::id) {
        launch_extend_features<FeatAnalytic>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, ordered, d_counters, d_cnt);
        return 1;
    }
    if (!q_walk) {
        launch_extend_features<FeatFull>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, ordered, d_counters, d_cnt);
        return 1;
    }
    if (d_cnt) {
        ordered ? launch_extend_split<true, true>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, q_walk, d_n_walk, d_counters, d_cnt, nullptr)
                : launch_extend_split<true, false>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, q_walk, d_n_walk, d_counters, d_cnt, nullptr);
    } else {
        ordered ? launch_extend_split<false, true>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, q_walk, d_n_walk, d_counters, nullptr, adv)
                : launch_extend_split<false, false>(l, s, w, queue, d_n_queue, max_n, d_cursor, sorted, q_walk, d_n_walk, d_counters, nullptr, adv);
    }
    return 2;
}

template <bool kCount, typename F>
static void launch_shadow_variant(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue,
                                  const uint32_t* d_n_queue, uint32_t max_n, uint32_t light_index, uint32_t* d_cursor,
                                  uint32_t* q_lit, uint32_t* d_n_lit, unsigned long long* d_counters, TraceCounters* d_cnt)
{
    static const int occ_ = trace_ctas_per_sm(k_shadow<kCount, F>);
    k_shadow<kCount, F><<<wavefront_grid(max_n, kTraceBlock, occ_, l.sm_count), kTraceBlock, 0, l.stream>>>(
        s, w, queue, d_n_queue, light_index, d_cursor, q_lit, d_n_lit, d_counters, d_cnt);
}

int launch_shadow(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue, const uint32_t* d_n_queue,
                  uint32_t max_n, uint32_t light_index, uint32_t* d_cursor, uint32_t* q_lit, uint32_t* d_n_lit, uint32_t* q_walk,
                  uint32_t* d_n_walk, unsigned long long* d_counters, TraceCounters* d_cnt)
{
    if (max_n == 0) return 0;
    const bool analytic = l.features == FeatAnalytic::id;
    if (!analytic && !d_cnt && q_walk) {
        // begin + walk.  (Not with node counting: `begin` asks the lights accelerator before the BVH, which skips node
        // visits the reference order would count.)
        static const int occ_b = trace_ctas_per_sm(k_shadow_begin<false, FeatFull>);
        static const int occ_w = trace_ctas_per_sm(k_shadow_walk<false, FeatFull>);
        k_shadow_begin<false, FeatFull><<<wavefront_grid(max_n, kTraceBlock, occ_b, l.sm_count), kTraceBlock, 0, l.stream>>>(
            s, w, queue, d_n_queue, light_index, q_walk, d_n_walk, q_lit, d_n_lit, d_counters, nullptr);
        debug_sync(l, "k_shadow_begin");
        k_shadow_walk<false, FeatFull><<<wavefront_grid(max_n, kTraceBlock, occ_w, l.sm_count), kTraceBlock, 0, l.stream>>>(
            s, w, q_walk, d_n_walk, light_index, d_cursor, q_lit, d_n_lit, d_counters, nullptr);
        debug_sync(l, "k_shadow_walk");
        return 2;
    }
    if (d_cnt) {
        analytic ? launch_shadow_variant<true, FeatAnalytic>(l, s, w, queue, d_n_queue, max_n, light_index, d_cursor, q_lit, d_n_lit, d_counters, d_cnt)
                 : launch_shadow_variant<true, FeatFull>(l, s, w, queue, d_n_queue, max_n, light_index, d_cursor, q_lit, d_n_lit, d_counters, d_cnt);
    } else {
        analytic ? launch_shadow_variant<false, FeatAnalytic>(l, s, w, queue, d_n_queue, max_n, light_index, d_cursor, q_lit, d_n_lit, d_counters, nullptr)
                 : launch_shadow_variant<false, FeatFull>(l, s, w, queue, d_n_queue, max_n, light_index, d_cursor, q_lit, d_n_lit, d_counters, nullptr);
    }
    return 1;
}

template <bool kCount, typename F>
static void launch_mis_variant(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue, const uint32_t* d_n_queue,
                               uint32_t max_n, unsigned long long* d_counters, TraceCounters* d_cnt)
{
    static const int occ_ = trace_ctas_per_sm(k_mis_trace<kCount, F>);
    k_mis_trace<kCount, F><<<wavefront_grid(max_n, kTraceBlock, occ_, l.sm_count), kTraceBlock, 0, l.stream>>>(
        s, w, queue, d_n_queue, d_counters, d_cnt);
}

int launch_mis_trace(const Launch& l, const DScene& s, const DWave& w, const uint32_t* queue, const uint32_t* d_n_queue,
                     uint32_t max_n, uint32_t* d_cursor, uint32_t* q_walk, uint32_t* d_n_walk, unsigned long long* d_counters,
                     TraceCounters* d_cnt)
{
    if (max_n == 0) return 0;
    const bool analytic = l.features == FeatAnalytic::id;
#ifndef SPCU_MIS_SINGLE
#define SPCU_MIS_SINGLE 0 // 1: A/B builds with round 1's one-thread-per-ray kernel
#endif
    if (!analytic && !d_cnt && q_walk && !SPCU_MIS_SINGLE) { // begin + persistent walk, as the shadow stage
        static const int occ_b = trace_ctas_per_sm(k_shadow_begin<false, FeatFull, true>);
        static const int occ_w = trace_ctas_per_sm(k_shadow_walk<false, FeatFull, true>);
        k_shadow_begin<false, FeatFull, true><<<wavefront_grid(max_n, kTraceBlock, occ_b, l.sm_count), kTraceBlock, 0, l.stream>>>(
            s, w, queue, d_n_queue, 0u, q_walk, d_n_walk, nullptr, nullptr, d_counters, nullptr);
        debug_sync(l, "k_mis_begin");
        k_shadow_walk<false, FeatFull, true><<<wavefront_grid(max_n, kTraceBlock, occ_w, l.sm_count), kTraceBlock, 0, l.stream>>>(
            s, w, q_walk, d_n_walk, 0u, d_cursor, nullptr, nullptr, d_counters, nullptr);
        debug_sync(l, "k_mis_walk");
        return 2;
    }
    if (d_cnt) {
        analytic ? launch_mis_variant<true, FeatAnalytic>(l, s, w, queue, d_n_queue, max_n, d_counters, d_cnt)
                 : launch_mis_variant<true, FeatFull>(l, s, w, queue, d_n_queue, max_n, d_counters, d_cnt);
    } else {
        analytic ? launch_mis_variant<false, FeatAnalytic>(l, s, w, queue, d_n_queue, max_n, d_counters, nullptr)
                 : launch_mis_variant<false, FeatFull>(l, s, w, queue, d_n_queue, max_n, d_counters, nullptr);
    }
    return 1;
}

} // namespace spcu
