// Extend-ray / any-hit traversal over the flattened copy of the reference's accelerators, in REFERENCE ORDER with
// the reference's exact arithmetic, so that primitive IDs and distances are bit-identical to Scene::intersect,
// Scene::intersect_p and Scene::intersect_lights (base/Scene.h:69-82).
//
// Arithmetic contract (SURVEY.md §0.6, §7 "hard parts"): every operation is spelled with a round-to-nearest
// intrinsic (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn/__fmaf_rn) which nvcc never contracts, reorders or
// replaces by an approximation; fused multiply-adds appear exactly where the reference calls madd/msub
// (math/Math.h:137-162).  This TU is additionally compiled with --fmad=false.
#pragma once

#include "device_scene.h"

#include <cfloat>

namespace spcu {

constexpr int   kTraceBlock   = 128; // threads per CTA of every traversal kernel
constexpr int   kStackShared  = 24;  // stack entries per thread kept in shared memory (bank-conflict free)
constexpr int   kStackLocal   = SPCU_MAX_BVH_DEPTH + 2 - kStackShared; // overflow, thread-local
constexpr float kInfinite     = FLT_MAX; // k_infinite_distance (base/Constants.h:16)
constexpr float kRayEpsilon   = 0.001f;  // k_ray_epsilon (math/Ray.h:11)

struct Ray
{
    float ox, oy, oz, dx, dy, dz;
    float t_min;
};

struct RayInv
{
    float x, y, z; // 1.0f / d, IEEE division: the value BBox.h:130 recomputes at every box
};

__device__ __forceinline__ RayInv make_inv(const Ray& r)
{
    return { __fdiv_rn(1.0f, r.dx), __fdiv_rn(1.0f, r.dy), __fdiv_rn(1.0f, r.dz) };
}

// _mm_dp_ps(a, b, 0x7F) (math/Vector3.h:742-746): (x*x' + y*y') + (z*z' + 0)
__device__ __forceinline__ float dot_dpps(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fadd_rn(__fmul_rn(az, bz), 0.0f));
}

// One axis of sp::intersect_p(BBox, Ray, RayLimits) (math/BBox.h:128-142).  std::max(a,b) = (a<b)?b:a and
// std::min(a,b) = (b<a)?b:a — a NaN t_near/t_far REPLACES the running bound, so fmaxf/fminf must not be used.
__device__ __forceinline__ bool slab_axis(float lo, float hi, float o, float inv, float& t0, float& t1)
{
    float t_near = __fmul_rn(__fsub_rn(lo, o), inv);
    float t_far  = __fmul_rn(__fsub_rn(hi, o), inv);
    if (t_near > t_far) {
        const float tmp = t_near;
        t_near          = t_far;
        t_far           = tmp;
    }
    t0 = (t_near < t0) ? t0 : t_near;
    t1 = (t1 < t_far) ? t1 : t_far;
    return !(t0 > t1);
}

__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r,
                                     const RayInv& inv, float t_max)
{
    float t0 = r.t_min, t1 = t_max;
    return slab_axis(lox, hix, r.ox, inv.x, t0, t1) && slab_axis(loy, hiy, r.oy, inv.y, t0, t1) &&
           slab_axis(loz, hiz, r.oz, inv.z, t0, t1);
}

// Triangle::intersect_impl (shapes/Triangle.h:97-146)
__device__ __forceinline__ bool tri_hit(const float4 p0, const float4 p1, const float4 p2, const Ray& r, float t_max,
                                        float& t_out, float& beta_out, float& gamma_out)
{
    const float A = __fsub_rn(p0.x, p1.x), B = __fsub_rn(p0.y, p1.y), C = __fsub_rn(p0.z, p1.z);
    const float D = __fsub_rn(p0.x, p2.x), E = __fsub_rn(p0.y, p2.y), F = __fsub_rn(p0.z, p2.z);
    const float G = r.dx, H = r.dy, I = r.dz;
    const float J = __fsub_rn(p0.x, r.ox), K = __fsub_rn(p0.y, r.oy), L = __fsub_rn(p0.z, r.oz);

    const float EIHF = __fmaf_rn(E, I, -__fmul_rn(H, F));
    const float GFDI = __fmaf_rn(G, F, -__fmul_rn(D, I));
    const float DHEG = __fmaf_rn(D, H, -__fmul_rn(E, G));

    const float denom = __fmaf_rn(A, EIHF, __fmaf_rn(B, GFDI, __fmul_rn(C, DHEG)));
    if (denom == 0.0f) {
        return false;
    }
    const float beta = __fdiv_rn(__fmaf_rn(J, EIHF, __fmaf_rn(K, GFDI, __fmul_rn(L, DHEG))), denom);
    if (beta <= 0.0f || beta >= 1.0f) {
        return false;
    }
    const float AKJB = __fmaf_rn(A, K, -__fmul_rn(J, B));
    const float JCAL = __fmaf_rn(J, C, -__fmul_rn(A, L));
    const float BLKC = __fmaf_rn(B, L, -__fmul_rn(K, C));

    const float gamma = __fdiv_rn(__fmaf_rn(I, AKJB, __fmaf_rn(H, JCAL, __fmul_rn(G, BLKC))), denom);
    if (gamma <= 0.0f || __fadd_rn(beta, gamma) >= 1.0f) {
        return false;
    }
    const float t = __fdiv_rn(-__fmaf_rn(F, AKJB, __fmaf_rn(E, JCAL, __fmul_rn(D, BLKC))), denom);
    if (t < r.t_min || t > t_max) {
        return false;
    }
    t_out     = t;
    beta_out  = beta;
    gamma_out = gamma;
    return true;
}

// AffineSpace::operator()(Point3) / LinearSpace3x3::operator()(Vector3) (math/AffineSpace.h:79-91,
// math/LinearSpace3x3.h:153-161).  m = c0.x c0.y c0.z c1.x | c1.y c1.z c2.x c2.y | c2.z a.x a.y a.z
struct LocalRay
{
    float ox, oy, oz, dx, dy, dz;
};

__device__ __forceinline__ LocalRay to_local(const float4 m0, const float4 m1, const float4 m2, const Ray& r)
{
    LocalRay l;
    l.ox = __fmaf_rn(r.ox, m0.x, __fmaf_rn(r.oy, m0.w, __fmaf_rn(r.oz, m1.z, m2.y)));
    l.oy = __fmaf_rn(r.ox, m0.y, __fmaf_rn(r.oy, m1.x, __fmaf_rn(r.oz, m1.w, m2.z)));
    l.oz = __fmaf_rn(r.ox, m0.z, __fmaf_rn(r.oy, m1.y, __fmaf_rn(r.oz, m2.x, m2.w)));
    l.dx = __fmaf_rn(r.dx, m0.x, __fmaf_rn(r.dy, m0.w, __fmul_rn(r.dz, m1.z)));
    l.dy = __fmaf_rn(r.dx, m0.y, __fmaf_rn(r.dy, m1.x, __fmul_rn(r.dz, m1.w)));
    l.dz = __fmaf_rn(r.dx, m0.z, __fmaf_rn(r.dy, m1.y, __fmul_rn(r.dz, m2.x)));
    return l;
}

// Sphere::intersect_impl (shapes/Sphere.h:77-97): unit sphere in object space; t is reused in world space.
__device__ __forceinline__ bool sphere_hit(const float4 m0, const float4 m1, const float4 m2, const Ray& r, float t_max,
                                           float& t_out)
{
    const LocalRay l = to_local(m0, m1, m2, r);
    const float    a = dot_dpps(l.dx, l.dy, l.dz, l.dx, l.dy, l.dz);
    const float    b = __fmul_rn(2.0f, dot_dpps(l.dx, l.dy, l.dz, l.ox, l.oy, l.oz));
    const float    c = __fsub_rn(dot_dpps(l.ox, l.oy, l.oz, l.ox, l.oy, l.oz), 1.0f);
    float          disc = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.0f, a), c));
    if (!(disc > 0.0f)) {
        return false;
    }
    disc             = __fsqrt_rn(disc);
    const float two_a = __fmul_rn(2.0f, a);
    float       t     = __fdiv_rn(__fsub_rn(-b, disc), two_a);
    if (t < r.t_min) {
        t = __fdiv_rn(__fadd_rn(-b, disc), two_a);
    }
    if (t < r.t_min || t > t_max) {
        return false;
    }
    t_out = t;
    return true;
}

// Plane::intersect_impl (shapes/Plane.h:49-63): y = 0 in object space.
__device__ __forceinline__ bool plane_hit(const float4 m0, const float4 m1, const float4 m2, const Ray& r, float t_max,
                                          float& t_out)
{
    const float dy = __fmaf_rn(r.dx, m0.y, __fmaf_rn(r.dy, m1.x, __fmul_rn(r.dz, m1.w)));
    if (dy == 0.0f) {
        return false;
    }
    const float oy = __fmaf_rn(r.ox, m0.y, __fmaf_rn(r.oy, m1.x, __fmaf_rn(r.oz, m1.w, m2.z)));
    const float t  = __fdiv_rn(-oy, dy);
    if (t < r.t_min || t > t_max) {
        return false;
    }
    t_out = t;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Primitive policies: what a "primitive" of an accelerator is and how it is tested.
// ---------------------------------------------------------------------------------------------------------------
struct GeomPrims
{
    const float4*   prims;
    const uint32_t* meta;

    // GeometricPrimitive::intersect_impl -> Shape::intersect (shapes/Primitive.h:37-44)
    template <bool kCount>
    __device__ __forceinline__ bool test(uint32_t id, bool mixed, const Ray& r, float t_max, float& t, float& beta,
                                         float& gamma, TraceCounters* cnt) const
    {
        const float4   a    = __ldg(prims + 3 * id + 0);
        const float4   b    = __ldg(prims + 3 * id + 1);
        const float4   c    = __ldg(prims + 3 * id + 2);
        const uint32_t kind = mixed ? SPCU_META_KIND(__ldg(meta + id)) : SPCU_PRIM_TRIANGLE;
        beta = gamma = 0.0f;
        if (kind == SPCU_PRIM_TRIANGLE) {
            if (kCount) ++cnt->tris;
            return tri_hit(a, b, c, r, t_max, t, beta, gamma);
        }
        if (kCount) ++cnt->xf;
        if (kind == SPCU_PRIM_SPHERE) {
            return sphere_hit(a, b, c, r, t_max, t);
        }
        return plane_hit(a, b, c, r, t_max, t);
    }
};

struct LightPrims
{
    const spcu_light* lights;

    __device__ __forceinline__ void load_xf(const spcu_light* l, float4& m0, float4& m1, float4& m2) const
    {
        const float* w = l->world_to_object;
        m0 = make_float4(w[0], w[1], w[2], w[3]);
        m1 = make_float4(w[4], w[5], w[6], w[7]);
        m2 = make_float4(w[8], w[9], w[10], w[11]);
    }

    // Light::intersect_lights_impl: SphereLight (Lights/Light.h:354-361), EnvironmentLight (:135-141),
    // ImageBasedEnvironmentLight (:196-201)
    template <bool kCount>
    __device__ __forceinline__ bool test(uint32_t id, bool, const Ray& r, float t_max, float& t, float& beta,
                                         float& gamma, TraceCounters*) const
    {
        const spcu_light* l = lights + id;
        beta = gamma = 0.0f;
        if (l->kind == SPCU_LIGHT_SPHERE) {
            float4 m0, m1, m2;
            load_xf(l, m0, m1, m2);
            return sphere_hit(m0, m1, m2, r, t_max, t);
        }
        if (t_max < kInfinite) {
            return false;
        }
        t = kInfinite;
        return true;
    }

    // Light::intersect_p_impl: only sphere lights occlude (:363-366 vs :143-146, :211-214)
    __device__ __forceinline__ bool test_any(uint32_t id, const Ray& r, float t_max) const
    {
        const spcu_light* l = lights + id;
        if (l->kind != SPCU_LIGHT_SPHERE) {
            return false;
        }
        float4 m0, m1, m2;
        float  t;
        load_xf(l, m0, m1, m2);
        return sphere_hit(m0, m1, m2, r, t_max, t);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Traversal stack: the first kStackShared levels live in shared memory, laid out [level][thread] so that a warp's
// accesses to one level hit 32 different banks; deeper levels spill to a thread-local array.
// An entry is the index of an internal node whose RIGHT child is still to be tested (the reference tests it only
// after the left subtree returned, with the t_max the left subtree left behind; BVHAccelerator.h:62-77).
// ---------------------------------------------------------------------------------------------------------------
struct Stack
{
    int32_t* sh; // &smem[threadIdx.x]
    int32_t  loc[kStackLocal];
    int      n = 0;

    __device__ __forceinline__ void push(int32_t v)
    {
        if (n < kStackShared) {
            sh[n * kTraceBlock] = v;
        } else {
            loc[n - kStackShared] = v;
        }
        ++n;
    }

    __device__ __forceinline__ int32_t pop()
    {
        --n;
        return (n < kStackShared) ? sh[n * kTraceBlock] : loc[n - kStackShared];
    }
};

struct NodeHalf
{
    float   lox, loy, loz, hix, hiy, hiz;
    int32_t child;
    uint32_t count;
};

__device__ __forceinline__ void load_node(const float4* nodes, int32_t idx, NodeHalf& c0, NodeHalf& c1)
{
    const float4 v0 = __ldg(nodes + 4 * idx + 0);
    const float4 v1 = __ldg(nodes + 4 * idx + 1);
    const float4 v2 = __ldg(nodes + 4 * idx + 2);
    const float4 v3 = __ldg(nodes + 4 * idx + 3);
    c0 = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, __float_as_int(v3.x), __float_as_uint(v3.z) };
    c1 = { v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, __float_as_int(v3.y), __float_as_uint(v3.w) };
}

__device__ __forceinline__ void load_right(const float4* nodes, int32_t idx, NodeHalf& c1)
{
    const float4 v1 = __ldg(nodes + 4 * idx + 1);
    const float4 v2 = __ldg(nodes + 4 * idx + 2);
    const float4 v3 = __ldg(nodes + 4 * idx + 3);
    c1 = { v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, __float_as_int(v3.y), __float_as_uint(v3.w) };
}

// Closest hit over ListAccelerator[unbounded..., BVH] (shapes/ListAccelerator.h:36-62 + BVHAccelerator.h:45-77,110-113).
// `t_max` enters as the query's limit and leaves as the accepted distance; returns the accepted primitive or -1.
template <bool kCount, typename Prims>
__device__ __forceinline__ int32_t closest_hit(const DAccel& acc, const Prims& prims, const Ray& r, float& t_max,
                                               float& beta, float& gamma, int32_t* stack_smem, TraceCounters* cnt)
{
    int32_t hit_id = -1;
    float   t, b, g;
    beta = gamma = 0.0f;

    // top-level list: unbounded primitives first, in list order; a hit shrinks t_max and replaces the result
    for (uint32_t i = 0; i < acc.n_unbounded; ++i) {
        if (prims.template test<kCount>(i, true, r, t_max, t, b, g, cnt)) {
            t_max  = t;
            hit_id = static_cast<int32_t>(i);
            beta   = b;
            gamma  = g;
        }
    }

    const RayInv inv = make_inv(r);
    Stack        stack;
    stack.sh = stack_smem;

    int32_t  link  = acc.root; // the root's own bounds are never tested (BVHAccelerator.h:138-142)
    uint32_t count = acc.root_count;
    for (;;) {
        if (link < 0) {
            // NodeLeaf -> ListAccelerator::intersect_impl over the leaf's primitives
            const uint32_t first = static_cast<uint32_t>(~link);
            const uint32_t n     = count & SPCU_LEAF_COUNT_MASK;
            const bool     mixed = (count & SPCU_LEAF_MIXED_FLAG) != 0u;
            for (uint32_t i = 0; i < n; ++i) {
                if (prims.template test<kCount>(first + i, mixed, r, t_max, t, b, g, cnt)) {
                    t_max  = t;
                    hit_id = static_cast<int32_t>(first + i);
                    beta   = b;
                    gamma  = g;
                }
            }
        } else {
            // NodeInternal: left child now, right child after the left subtree
            NodeHalf c0, c1;
            load_node(acc.nodes, link, c0, c1);
            if (kCount) ++cnt->nodes;
            if (slab(c0.lox, c0.loy, c0.loz, c0.hix, c0.hiy, c0.hiz, r, inv, t_max)) {
                stack.push(link);
                link  = c0.child;
                count = c0.count;
                continue;
            }
            if (slab(c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz, r, inv, t_max)) {
                link  = c1.child;
                count = c1.count;
                continue;
            }
        }
        // return to the innermost node whose right child is pending
        bool descended = false;
        while (stack.n > 0) {
            const int32_t idx = stack.pop();
            NodeHalf      c1;
            load_right(acc.nodes, idx, c1);
            if (slab(c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz, r, inv, t_max)) {
                link      = c1.child;
                count     = c1.count;
                descended = true;
                break;
            }
        }
        if (!descended) {
            break;
        }
    }
    return hit_id;
}

// Any hit over one accelerator (ListAccelerator::intersect_p_impl :64-67, NodeInternal::intersect_p :79-90):
// limits never change, first accepted primitive ends the query.
template <bool kCount, typename AnyTest>
__device__ __forceinline__ bool any_hit(const DAccel& acc, const AnyTest& test, const Ray& r, float t_max,
                                        int32_t* stack_smem, TraceCounters* cnt)
{
    for (uint32_t i = 0; i < acc.n_unbounded; ++i) {
        if (test(i, true, cnt)) {
            return true;
        }
    }
    const RayInv inv = make_inv(r);
    Stack        stack;
    stack.sh = stack_smem;

    int32_t  link  = acc.root;
    uint32_t count = acc.root_count;
    for (;;) {
        if (link < 0) {
            const uint32_t first = static_cast<uint32_t>(~link);
            const uint32_t n     = count & SPCU_LEAF_COUNT_MASK;
            const bool     mixed = (count & SPCU_LEAF_MIXED_FLAG) != 0u;
            for (uint32_t i = 0; i < n; ++i) {
                if (test(first + i, mixed, cnt)) {
                    return true;
                }
            }
        } else {
            NodeHalf c0, c1;
            load_node(acc.nodes, link, c0, c1);
            if (kCount) ++cnt->nodes;
            const bool h0 = slab(c0.lox, c0.loy, c0.loz, c0.hix, c0.hiy, c0.hiz, r, inv, t_max);
            const bool h1 = slab(c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz, r, inv, t_max);
            // limits are constant here, so both boxes can be tested at once; order of descent stays left, right
            if (h0) {
                if (h1) {
                    stack.push(link);
                }
                link  = c0.child;
                count = c0.count;
                continue;
            }
            if (h1) {
                link  = c1.child;
                count = c1.count;
                continue;
            }
        }
        if (stack.n == 0) {
            return false;
        }
        const int32_t idx = stack.pop();
        const float4  v3  = __ldg(acc.nodes + 4 * idx + 3);
        link              = __float_as_int(v3.y);
        count             = __float_as_uint(v3.w);
    }
}

// Scene::intersect_p (base/Scene.h:79-82): geometry accelerator, then lights accelerator.
template <bool kCount>
__device__ __forceinline__ bool scene_any_hit(const DScene& s, const Ray& r, float t_max, int32_t* stack_smem,
                                              TraceCounters* cnt)
{
    const GeomPrims gp{ s.geom_prims, s.geom_meta };
    auto geom_test = [&](uint32_t id, bool mixed, TraceCounters* c) {
        float t, b, g;
        return gp.template test<kCount>(id, mixed, r, t_max, t, b, g, c);
    };
    if (any_hit<kCount>(s.geom, geom_test, r, t_max, stack_smem, cnt)) {
        return true;
    }
    const LightPrims lp{ s.lights };
    auto light_test = [&](uint32_t id, bool, TraceCounters*) { return lp.test_any(id, r, t_max); };
    return any_hit<false>(s.lights_accel, light_test, r, t_max, stack_smem, nullptr);
}

} // namespace spcu
