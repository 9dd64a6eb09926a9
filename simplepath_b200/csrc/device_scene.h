// Device-side view of the flattened scene (include/spcu.h) and of the wavefront state.  Host and device code of
// the backend share this header; nothing here is visible through the C-ABI.
#pragma once

#include "spcu.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace spcu {

struct DAccel
{
    const float4* nodes; // 4 x float4 per node (spcu_bvh_node)
    int32_t       root;
    uint32_t      root_count;
    uint32_t      n_unbounded;
    uint32_t      n_prims;
};

struct DScene
{
    uint32_t width, height, rr_depth, max_depth;
    float    camera[12];

    DAccel          geom;
    const float4*   geom_prims; // 3 x float4 per primitive (spcu_prim_geom)
    const float4*   geom_shade; // 3 x float4 per primitive (spcu_prim_shade)
    const uint32_t* geom_meta;

    DAccel            lights_accel;
    uint32_t          n_lights;
    const spcu_light* lights;
    const uint32_t*   light_order;

    const spcu_material* materials;
    const spcu_bxdf*     bxdfs;
    const float*         pool;

    const float* jitter; // spp x 2
    uint32_t     spp;
};

// Closest-hit record of the extend stage: 16 bytes, one vector store.
struct alignas(16) HitRec
{
    int32_t id;    // reference-order primitive ID or -1
    float   t;     // distance (query t_max on a miss)
    float   beta;  // triangle barycentrics of the accepted hit (Triangle.h:121,130)
    float   gamma;
};

struct TraceCounters
{
    unsigned long long nodes;
    unsigned long long tris;
    unsigned long long xf;
};

// ---- wavefront state (structure of arrays over path slots) -----------------------------------------------
// One slot = one camera sample in flight.  Every array has `capacity` entries; all float4 arrays are 16-byte
// vector loads/stores, coalesced when queues are dense.
struct DWave
{
    uint32_t capacity;

    // camera-sample identity of a slot and its random-number draw counter (RNG contract: rng.cuh)
    uint32_t* pixel;   // global pixel index y*w+x
    uint32_t* sample;  // global sample index
    uint32_t* rng_ctr; // number of draw calls made so far on this path

    // current path segment: sp::Ray + sp::RayLimits (32 B as two float4)
    float4* ray_o; // o.xyz, t_min
    float4* ray_d; // d.xyz, t_max

    float4* throughput; // rgb, (unused)
    float4* radiance;   // L rgb, (unused)

    HitRec* hit;       // Scene::intersect result
    int2*   light_hit; // Scene::intersect_lights result: (light id, float bits of distance)

    // surface interaction kept across the NEE stages of one vertex
    float4* isect_p; // point xyz, material index (as int bits)
    float4* isect_n; // shading normal xyz, (unused)

    // primary BSDF sample S0 of the vertex (Integrator.cpp:569): direction, colour, pdf
    float4* s0_dir; // wi xyz, pdf
    float4* s0_col; // rgb, (unused)

    // NEE light-sampling strategy (Integrator.cpp:497-516); the shadow ray starts at isect_p
    float4*  sh_d;    // shadow ray d, t_max
    float*   sh_tmin; // shadow ray t_min
    float4*  light_L; // light sample radiance rgb, light pdf
    uint8_t* occluded;

    // NEE BSDF-sampling strategy (Integrator.cpp:518-536)
    float4* mis_d;   // material ray d xyz, t_min   (origin = isect point, t_max = FLT_MAX)
    float4* mis_col; // material sample colour rgb, pdf
    float2* mis_cw;  // |cos(n, wi)|, balance-heuristic weight
    float4* nee_acc; // light-strategy term of this light, added together with the BSDF-strategy term (:516 + :534)
    int2*   mis_hit; // (light id or -1, occluded flag)
};

// Output of the extend stage: path vertices grouped by material, so that warps of the shading stages run one material's
// code (SURVEY.md §7 "sort/compact by material before shading").  Segment m < n_segments-1 holds hits on material m (the
// last material segment also takes any higher index), segment n_segments-1 holds misses.
struct SortedQueue
{
    uint32_t* slots;  // [n_segments][capacity]
    uint32_t* counts; // [n_segments], zeroed per depth
    uint32_t  n_segments;
    uint32_t  capacity;
};

// indices into the device counter block (unsigned long long[kNumCounters])
enum Counter : int
{
    kCntPaths = 0,
    kCntRaysClosest,
    kCntRaysAny,
    kCntRaysLights,
    kCntShadeCalls,
    kNumCounters
};

// wavefront stages, in pipeline order; device item counters live at counters[kNumCounters + stage]
enum Stage : int
{
    kStRaygen = 0,
    kStExtend,
    kStShade,
    kStNeeLight,
    kStShadow,
    kStNeeBsdf,
    kStMisTrace,
    kStNeeMisAccumulate,
    kStDirectAccumulate,
    kStAdvance,
    kStResolve,
    kStPaths, // the persistent path kernel (all of the above but resolve, fused)
    kNumStages
};

// Queue compaction: warp ballot + ONE atomicAdd per warp; all 32 lanes of the warp must call it together.
__device__ __forceinline__ void queue_push(uint32_t* queue, uint32_t* n_queue, uint32_t slot, bool pred)
{
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0u) {
        return;
    }
    const int lane   = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t  base   = 0;
    if (lane == leader) {
        base = atomicAdd(n_queue, static_cast<uint32_t>(__popc(mask)));
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        queue[base + __popc(mask & ((1u << lane) - 1u))] = slot;
    }
}

__device__ __forceinline__ void count_items(unsigned long long* counters, Stage stage, uint32_t n)
{
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) {
        atomicAdd(counters + kNumCounters + stage, static_cast<unsigned long long>(n));
    }
}

} // namespace spcu
