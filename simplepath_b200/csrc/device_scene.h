// Device-side view of the flattened scene (include/spcu.h) and of the wavefront state.  Host and device code of
// the backend share this header; nothing here is visible through the C-ABI.
#pragma once

#include "spcu.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace spcu {

// Packed child of a 4-wide node (build_kernels.cu k_build_wide, trace.cuh wide_unpack) — what the wide walks keep on their stacks:
//   bit 31 clear: an internal node, the word is its index;
//   bit 31 set:   a leaf — bit 30 = SPCU_LEAF_MIXED_FLAG, bits 29..27 = its primitive count (0..kWideSmallLeafMax) and bits
//                 26..0 = its first primitive, or count bits = kWideBigLeafTag and bits 26..0 = a slot of DAccel::big.
constexpr uint32_t kWideSmallLeafMax = 6u;
constexpr uint32_t kWideBigLeafTag   = 7u;
constexpr uint32_t kWidePayloadMask  = 0x07ffffffu;

struct DAccel
{
    const float4* nodes; // 4 x float4 per node (spcu_bvh_node)
    const float4* wide;  // 8 x float4 per node: the 4-wide node of binary node i (trace.cuh); NULL without internal nodes
    const int2*   big;   // { link, count } of the leaves a packed child cannot describe (more than kWideSmallLeafMax primitives)
    int32_t       root;
    uint32_t      root_count;
    uint32_t      n_unbounded;
    uint32_t      n_prims;
    uint32_t      proper_boxes; // every child box is finite with lo <= hi: the slab test may take its NaN-free form (trace.cuh)
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

// Closest-hit record as the traversal code produces it (id, distance, triangle barycentrics).
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

// ---- wavefront state ---------------------------------------------------------------------------------------------
// One slot = one camera sample in flight.  Queues are compacted (and sorted by material), so stages reach their slots
// in scattered order: a 4-byte or 16-byte field of a structure-of-arrays layout then costs a whole 32-byte DRAM sector
// (ncu on round 1's SoA layout: 1.5-2.7x the algorithmic bytes, profiles/ncu_traffic_r01_c2.json).  The state is
// therefore an array of 32-byte-aligned records per kind, each holding what one stage reads or writes TOGETHER, every
// access a pair of 16-byte vector loads/stores that use their sector completely.
struct alignas(32) PathRec // identity + accumulators of the path
{
    float4 tp; // throughput rgb, pixel index (bits)
    float4 L;  // radiance rgb,   global sample index (bits)
};
struct alignas(32) RayRec // sp::Ray + sp::RayLimits of the current segment
{
    float4 o; // o.xyz, t_min
    float4 d; // d.xyz, t_max
};
struct alignas(32) VertexRec // surface interaction kept across the NEE stages of one vertex
{
    float4 p; // point xyz, material index (bits)
    float4 n; // shading normal xyz, (unused)
};
struct alignas(32) ExtendRec // result of the extend stage
{
    HitRec hit;      // Scene::intersect
    int32_t light;   // Scene::intersect_lights: light id or -1 ...
    float   light_t; // ... and its distance
    float   pad[2];
};
struct alignas(32) SampleRec // primary BSDF sample S0 of the vertex (Integrator.cpp:569)
{
    float4 dir; // wi xyz, pdf
    float4 col; // rgb, (unused)
};
struct alignas(32) LightRec // light sample of the NEE light strategy (Integrator.cpp:497-516); the ray starts at VertexRec::p
{
    float4 wi;  // direction xyz, t_max
    float4 aux; // t_min, light pdf, (u, v) of an image-based light's sample — its radiance is looked up again from them;
                // the other lights' radiance is a constant of the light
};
struct alignas(64) MisRec // NEE BSDF strategy (Integrator.cpp:518-536)
{
    float4  d;        // material ray d xyz, t_min (origin = VertexRec::p, t_max = FLT_MAX)
    float4  col;      // material sample colour rgb, pdf
    float4  cwa;      // |cos(n, wi)|, balance-heuristic weight, light-strategy term A.r, A.g
    float   ab;       // A.b
    int32_t light;    // mis trace: light reached or -1
    int32_t occluded; // mis trace: Scene::intersect_p of the same ray
    float   pad;
};

struct DWave
{
    uint32_t   capacity;
    PathRec*   path;
    RayRec*    ray;
    VertexRec* vertex;
    ExtendRec* extend;
    SampleRec* s0;
    LightRec*  light;
    MisRec*    mis;
    uint8_t*   occluded; // direct lighting only: shadow result per slot
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
    kCntErrors, // protocol faults a kernel detected instead of hanging (bounded spins / iteration caps): must stay 0
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
