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
#include "features.h"

#include <cfloat>

namespace spcu {

constexpr int   kTraceBlock   = 128; // threads per CTA of every traversal kernel
constexpr int   kStackShared  = 24;  // stack entries per thread kept in shared memory (bank-conflict free)
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
    bool  generic; // the slab test must run with the reference's literal swap / NaN rules (see slab_fast)
};

__device__ __forceinline__ bool is_finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); }

// `proper_boxes`: every child box of the accelerator is finite with lo <= hi (checked at upload).
__device__ __forceinline__ RayInv make_inv(const Ray& r, bool proper_boxes)
{
    RayInv inv{ __fdiv_rn(1.0f, r.dx), __fdiv_rn(1.0f, r.dy), __fdiv_rn(1.0f, r.dz), false };
    inv.generic = !(proper_boxes && is_finite(inv.x) && is_finite(inv.y) && is_finite(inv.z) && is_finite(r.ox) &&
                    is_finite(r.oy) && is_finite(r.oz));
    return inv;
}

// for the walks that use 4-wide nodes where they can: `generic` also when the accelerator has none
struct DAccel;
__device__ __forceinline__ RayInv make_inv_wide(const Ray& r, const DAccel& acc);

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

__device__ __forceinline__ bool slab_literal(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r,
                                             const RayInv& inv, float t_max, float& t0_out)
{
    float      t0 = r.t_min, t1 = t_max;
    const bool ok = slab_axis(lox, hix, r.ox, inv.x, t0, t1) && slab_axis(loy, hiy, r.oy, inv.y, t0, t1) &&
                    slab_axis(loz, hiz, r.oz, inv.z, t0, t1);
    t0_out = t0;
    return ok;
}

// The same test in 8 instead of 11 instructions per axis, for the rays and boxes that cannot produce a NaN: a finite box
// with lo <= hi, a finite origin and three finite reciprocals (0 * inf is the only NaN source of (b - o) * inv).  Without
// NaNs the swap is min/max of the two products (equal products: nothing to swap), and the running bounds are max / min:
// fmaxf / fminf return the non-NaN operand, which is also what the reference's ternaries do with a NaN t_min / t_max on the
// first axis.  Results are the same VALUES (a zero may carry the other sign; t0 / t1 are only ever compared).  The three
// axes are evaluated without the reference's early exit: each axis only tightens [t0, t1], so the final test decides the same.
__device__ __forceinline__ bool slab_fast(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r,
                                          const RayInv& inv, float t_max, float& t0_out)
{
    const float ax = __fmul_rn(__fsub_rn(lox, r.ox), inv.x), bx = __fmul_rn(__fsub_rn(hix, r.ox), inv.x);
    const float ay = __fmul_rn(__fsub_rn(loy, r.oy), inv.y), by = __fmul_rn(__fsub_rn(hiy, r.oy), inv.y);
    const float az = __fmul_rn(__fsub_rn(loz, r.oz), inv.z), bz = __fmul_rn(__fsub_rn(hiz, r.oz), inv.z);
    const float t0 = fmaxf(fmaxf(fmaxf(r.t_min, fminf(ax, bx)), fminf(ay, by)), fminf(az, bz));
    const float t1 = fminf(fminf(fminf(t_max, fmaxf(ax, bx)), fmaxf(ay, by)), fmaxf(az, bz));
    t0_out         = t0;
    return !(t0 > t1);
}

__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r,
                                     const RayInv& inv, float t_max)
{
    float t0;
    return slab_literal(lox, loy, loz, hix, hiy, hiz, r, inv, t_max, t0);
}

// Same test, also returning the entry distance t0 (used by the ordered traversal to visit the nearer child first).
__device__ __forceinline__ bool slab_t0(float lox, float loy, float loz, float hix, float hiy, float hiz, const Ray& r,
                                        const RayInv& inv, float t_max, float& t0_out)
{
    return slab_literal(lox, loy, loz, hix, hiy, hiz, r, inv, t_max, t0_out);
}

// Early outs of the triangle test that decide `q <= 0` / `q >= 1` for q = fl(num / den) WITHOUT dividing.  They only
// fire when the IEEE quotient is certain to satisfy the comparison, so the decision is the reference's bit for bit:
//   * num == 0, or num and den of strictly opposite signs  =>  q is -0, +0 or negative  =>  q <= 0;
//   * |num| >= |den| with equal signs                      =>  the exact quotient is >= 1, and so is its rounding.
// Everything else (including NaN operands, for which every comparison below is false, and quotients that round up to
// 1 or underflow to 0) falls through to the real division.  Most triangle tests of a leaf end here: the walk is
// issue-bound and an IEEE division costs about ten instructions.
__device__ __forceinline__ bool quotient_not_positive(float num, float den)
{
    return num == 0.0f || (num < 0.0f && den > 0.0f) || (num > 0.0f && den < 0.0f);
}
__device__ __forceinline__ bool quotient_outside_unit(float num, float den)
{
    return quotient_not_positive(num, den) || fabsf(num) >= fabsf(den);
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
    const float beta_num = __fmaf_rn(J, EIHF, __fmaf_rn(K, GFDI, __fmul_rn(L, DHEG)));
    if (quotient_outside_unit(beta_num, denom)) { // beta <= 0 || beta >= 1 decided without the division
        return false;
    }
    const float beta = __fdiv_rn(beta_num, denom);
    if (beta <= 0.0f || beta >= 1.0f) {
        return false;
    }
    const float AKJB = __fmaf_rn(A, K, -__fmul_rn(J, B));
    const float JCAL = __fmaf_rn(J, C, -__fmul_rn(A, L));
    const float BLKC = __fmaf_rn(B, L, -__fmul_rn(K, C));

    const float gamma_num = __fmaf_rn(I, AKJB, __fmaf_rn(H, JCAL, __fmul_rn(G, BLKC)));
    if (quotient_not_positive(gamma_num, denom)) { // gamma <= 0 decided without the division
        return false;
    }
    const float gamma = __fdiv_rn(gamma_num, denom);
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
template <typename F>
struct GeomPrimsT
{
    const float4*   prims;
    const uint32_t* meta;
    static constexpr bool kBvh = F::bvh; // the accelerator has internal nodes (false: its root is a leaf)

    // GeometricPrimitive::intersect_impl -> Shape::intersect (shapes/Primitive.h:37-44)
    template <bool kCount>
    __device__ __forceinline__ bool test(uint32_t id, bool mixed, const Ray& r, float t_max, float& t, float& beta,
                                         float& gamma, TraceCounters* cnt) const
    {
        const float4   a    = __ldg(prims + 3 * id + 0);
        const float4   b    = __ldg(prims + 3 * id + 1);
        const float4   c    = __ldg(prims + 3 * id + 2);
        const uint32_t kind = (mixed || !F::triangles) ? SPCU_META_KIND(__ldg(meta + id)) : SPCU_PRIM_TRIANGLE;
        beta = gamma = 0.0f;
        if (F::triangles && kind == SPCU_PRIM_TRIANGLE) {
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

using GeomPrims = GeomPrimsT<FeatFull>;

template <typename F>
struct LightPrimsT
{
    const spcu_light* lights;
    static constexpr bool kBvh = F::bvh; // feature sets without internal nodes have none in the lights accelerator either

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
using LightPrims = LightPrimsT<FeatFull>;

// ---------------------------------------------------------------------------------------------------------------
// Traversal stack: the first kStackShared levels live in shared memory, laid out [level][thread] so that a warp's
// accesses to one level hit 32 different banks; deeper levels spill to a thread-local array.
// An entry is the index of an internal node whose RIGHT child is still to be tested (the reference tests it only
// after the left subtree returned, with the t_max the left subtree left behind; BVHAccelerator.h:62-77).
// ---------------------------------------------------------------------------------------------------------------
template <int kShared, int kCapacity>
struct StackT
{
    int32_t* sh; // &smem[threadIdx.x]
    int32_t  loc[kCapacity - kShared];
    int      n = 0;

    __device__ __forceinline__ void push(int32_t v)
    {
        if (n < kShared) {
            sh[n * kTraceBlock] = v;
        } else {
            loc[n - kShared] = v;
        }
        ++n;
    }

    __device__ __forceinline__ int32_t pop()
    {
        --n;
        return (n < kShared) ? sh[n * kTraceBlock] : loc[n - kShared];
    }
};
using Stack = StackT<kStackShared, SPCU_MAX_BVH_DEPTH + 2>;

struct NodeHalf
{
    float   lox, loy, loz, hix, hiy, hiz;
    int32_t child;
    uint32_t count;
};

// 256-bit read-only global load (sm_100+: LDG.E.256): 32-byte aligned address, ONE 32-byte sector per lane.  A walk's loads
// are scattered — every lane of a warp reads another node — so the L1's cost is per (lane, sector) tag look-up, not per byte:
// ncu on the bunny scene showed l1tex at 69 % of its peak with four 128-bit loads per 64-byte node, i.e. every sector looked
// up twice.  Two 256-bit loads read the same node with half the look-ups.
struct Float8
{
    float4 lo, hi;
};

__device__ __forceinline__ Float8 ldg256(const void* p)
{
    unsigned long long a, b, c, d; // (four 64-bit registers: cicc 12.9 crashes on eight float outputs in these kernels)
    // volatile: a plain asm is a pure function to the compiler, which then hoists the load out of the branch that guards the
    // pointer (k_shadow_begin read DAccel::wide == NULL for scenes without internal nodes)
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    Float8 r;
    r.lo = make_float4(__uint_as_float(static_cast<unsigned>(a)), __uint_as_float(static_cast<unsigned>(a >> 32)),
                       __uint_as_float(static_cast<unsigned>(b)), __uint_as_float(static_cast<unsigned>(b >> 32)));
    r.hi = make_float4(__uint_as_float(static_cast<unsigned>(c)), __uint_as_float(static_cast<unsigned>(c >> 32)),
                       __uint_as_float(static_cast<unsigned>(d)), __uint_as_float(static_cast<unsigned>(d >> 32)));
    return r;
}

__device__ __forceinline__ void load_node(const float4* nodes, int32_t idx, NodeHalf& c0, NodeHalf& c1)
{
#ifndef SPCU_NODE_LOAD_128 // (A/B switch: four 128-bit loads, as in round 1)
    const Float8 a = ldg256(nodes + 4 * idx), b = ldg256(nodes + 4 * idx + 2);
    c0 = { a.lo.x, a.lo.y, a.lo.z, a.lo.w, a.hi.x, a.hi.y, __float_as_int(b.hi.x), __float_as_uint(b.hi.z) };
    c1 = { a.hi.z, a.hi.w, b.lo.x, b.lo.y, b.lo.z, b.lo.w, __float_as_int(b.hi.y), __float_as_uint(b.hi.w) };
#else
    const float4 v0 = __ldg(nodes + 4 * idx + 0);
    const float4 v1 = __ldg(nodes + 4 * idx + 1);
    const float4 v2 = __ldg(nodes + 4 * idx + 2);
    const float4 v3 = __ldg(nodes + 4 * idx + 3);
    c0 = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, __float_as_int(v3.x), __float_as_uint(v3.z) };
    c1 = { v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, __float_as_int(v3.y), __float_as_uint(v3.w) };
#endif
}

__device__ __forceinline__ void load_right(const float4* nodes, int32_t idx, NodeHalf& c1)
{
    const float4 v1 = __ldg(nodes + 4 * idx + 1);
    const float4 v2 = __ldg(nodes + 4 * idx + 2);
    const float4 v3 = __ldg(nodes + 4 * idx + 3);
    c1 = { v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, __float_as_int(v3.y), __float_as_uint(v3.w) };
}

// Both child boxes of node `idx` against the current limits; e0 / e1 = entry distances (valid where the box is hit).
// The literal form lives out of line and reloads the node: it runs for axis-parallel rays and improper boxes only.
static __device__ __noinline__ unsigned slab_pair_literal(const float4* nodes, int32_t idx, const Ray r, const RayInv inv,
                                                          float t_max, float& e0, float& e1)
{
    NodeHalf c0, c1;
    load_node(nodes, idx, c0, c1);
    const bool h0 = slab_literal(c0.lox, c0.loy, c0.loz, c0.hix, c0.hiy, c0.hiz, r, inv, t_max, e0);
    const bool h1 = slab_literal(c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz, r, inv, t_max, e1);
    return (h0 ? 1u : 0u) | (h1 ? 2u : 0u);
}

__device__ __forceinline__ void slab_pair(const float4* nodes, int32_t idx, const NodeHalf& c0, const NodeHalf& c1, const Ray& r,
                                          const RayInv& inv, float t_max, bool& h0, bool& h1, float& e0, float& e1)
{
#ifndef SPCU_FAST_SLAB
#define SPCU_FAST_SLAB 1 // 0: always the literal form (A/B builds)
#endif
    if (!SPCU_FAST_SLAB || inv.generic) {
        const unsigned m = slab_pair_literal(nodes, idx, r, inv, t_max, e0, e1);
        h0               = (m & 1u) != 0u;
        h1               = (m & 2u) != 0u;
    } else {
        h0 = slab_fast(c0.lox, c0.loy, c0.loz, c0.hix, c0.hiy, c0.hiz, r, inv, t_max, e0);
        h1 = slab_fast(c1.lox, c1.loy, c1.loz, c1.hix, c1.hiy, c1.hiz, r, inv, t_max, e1);
    }
}

constexpr int32_t kDone      = 0x7fffffff; // traversal cursor: no work left
constexpr int     kAllLeaves = 0x7fffffff; // "run to completion"

// ---------------------------------------------------------------------------------------------------------------------
// Closest hit over ListAccelerator[unbounded..., BVH] (shapes/ListAccelerator.h:36-62 + BVHAccelerator.h:45-77,110-113),
// as a RESUMABLE walk: closest_begin scans the unbounded list and sets the cursor, closest_run advances the BVH part by
// at most `max_leaves` leaf visits and can be called again.  The persistent traversal kernels use that to hand new rays
// to lanes whose walk has ended while the other lanes of the warp keep going.
//
// Order and arithmetic are the reference's: children left then right, the right child's box tested against the t_max the
// left subtree left behind, leaf primitives in list order, a hit replaces the result (equal t: the later one wins).
// The loop is shaped for SIMD efficiency ("while-while"): all lanes first walk internal nodes — one iteration = both child
// boxes of one node, branch-free apart from the final select — until each holds a leaf, then all lanes test leaf
// primitives.  The cursor is either a fresh internal node (>= 0), a popped node whose RIGHT child is pending
// (`retest`), a leaf (< 0), or kDone.  The right child is pre-filtered with the current t_max before it is pushed: the
// slab test is monotone in t_max (also through its NaN rule), so a box that fails now fails later too, and nothing that
// the reference would enter is skipped; it is re-tested when popped, exactly where the reference tests it.
// ---------------------------------------------------------------------------------------------------------------------
struct ClosestWalk
{
    int32_t  link;
    uint32_t count;
    bool     retest;
    int32_t  hit_id; // accepted primitive or -1
    float    t_max;  // in: the query's limit; out: the accepted distance
    float    beta, gamma;
};

template <bool kCount, typename Prims>
__device__ __forceinline__ void closest_begin(const DAccel& acc, const Prims& prims, const Ray& r, ClosestWalk& w,
                                              TraceCounters* cnt)
{
    float t, b, g;
    w.hit_id = -1;
    w.beta = w.gamma = 0.0f;
    // top-level list: unbounded primitives first, in list order; a hit shrinks t_max and replaces the result
    for (uint32_t i = 0; i < acc.n_unbounded; ++i) {
        if (prims.template test<kCount>(i, true, r, w.t_max, t, b, g, cnt)) {
            w.t_max  = t;
            w.hit_id = static_cast<int32_t>(i);
            w.beta   = b;
            w.gamma  = g;
        }
    }
    w.link   = acc.root; // the root's own bounds are never tested (BVHAccelerator.h:138-142)
    w.count  = acc.root_count;
    w.retest = false;
}

// `mask` = the lanes that execute this call together (the full warp, or __activemask() of a converged subset).  Loop
// conditions are warp votes over `mask`, so every lane of the group runs the same number of iterations and the
// compiler's reconvergence points sit INSIDE the loops: lanes holding an internal node really do step in lockstep
// (without the votes each lane ends up running its own copy of the loop, serialised — ncu showed 2-4 of 32 lanes active).
template <bool kCount, typename Prims>
__device__ __forceinline__ void closest_run(const DAccel& acc, const Prims& prims, const Ray& r, const RayInv& inv,
                                            ClosestWalk& w, Stack& stack, int max_leaves, TraceCounters* cnt, unsigned mask)
{
    float t, b, g;
    if (!Prims::kBvh) { // feature set without internal nodes: the root is a leaf, scanned in list order
        if (w.link != kDone) {
            const uint32_t first = static_cast<uint32_t>(~w.link), n = w.count & SPCU_LEAF_COUNT_MASK;
            for (uint32_t i = 0; i < n; ++i) {
                if (prims.template test<kCount>(first + i, true, r, w.t_max, t, b, g, cnt)) {
                    w.t_max  = t;
                    w.hit_id = static_cast<int32_t>(first + i);
                    w.beta   = b;
                    w.gamma  = g;
                }
            }
            w.link = kDone;
        }
        return;
    }
#pragma unroll 1
    while (__any_sync(mask, w.link != kDone) && max_leaves > 0) {
        // ---- internal nodes --------------------------------------------------------------------------------------
#pragma unroll 1
        while (__any_sync(mask, w.link >= 0 && w.link != kDone)) {
            if (w.link >= 0 && w.link != kDone) {
                NodeHalf c0, c1;
                load_node(acc.nodes, w.link, c0, c1);
                if (kCount && !w.retest) ++cnt->nodes;
                bool  h0, h1;
                float e0, e1;
                slab_pair(acc.nodes, w.link, c0, c1, r, inv, w.t_max, h0, h1, e0, e1);
                h0 = h0 && !w.retest;
                if (h0) {
                    if (h1) {
                        stack.push(w.link); // right child pending: re-tested against the t_max of that moment
                    }
                    w.link   = c0.child;
                    w.count  = c0.count;
                    w.retest = false;
                } else if (h1) {
                    w.link   = c1.child;
                    w.count  = c1.count;
                    w.retest = false;
                } else if (stack.n > 0) {
                    w.link   = stack.pop();
                    w.retest = true;
                } else {
                    w.link = kDone;
                }
            }
        }
        // ---- leaf: NodeLeaf -> ListAccelerator::intersect_impl over the leaf's primitives ------------------------
        const bool     leaf  = w.link != kDone;
        const uint32_t first = static_cast<uint32_t>(~w.link);
        const uint32_t n     = leaf ? (w.count & SPCU_LEAF_COUNT_MASK) : 0u;
        const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
        const uint32_t n_max = __reduce_max_sync(mask, n);
#pragma unroll 1
        for (uint32_t i = 0; i < n_max; ++i) {
            if (i < n && prims.template test<kCount>(first + i, mixed, r, w.t_max, t, b, g, cnt)) {
                w.t_max  = t;
                w.hit_id = static_cast<int32_t>(first + i);
                w.beta   = b;
                w.gamma  = g;
            }
        }
        if (leaf) {
            if (stack.n > 0) {
                w.link   = stack.pop();
                w.retest = true;
            } else {
                w.link = kDone;
            }
        }
        --max_leaves;
    }
}

// ---- the same walk as single steps, for kernels that vote per step which phase the warp runs --------------------------
// (k_extend / k_shadow: while-while leaves lanes that already hold a leaf waiting through the whole node loop — ncu on the
// bunny scene: the slab tests ran at 5.7 of 32 lanes.)  Per lane the sequence of node and leaf steps is unchanged.
__device__ __forceinline__ bool at_node(const ClosestWalk& w) { return w.link >= 0 && w.link != kDone; }
__device__ __forceinline__ bool at_leaf(const ClosestWalk& w) { return w.link < 0; }

// precondition: at_node(w)
template <bool kCount, typename StackS>
__device__ __forceinline__ void closest_node_step(const DAccel& acc, const Ray& r, const RayInv& inv, ClosestWalk& w,
                                                  StackS& stack, TraceCounters* cnt)
{
    NodeHalf c0, c1;
    load_node(acc.nodes, w.link, c0, c1);
    if (kCount && !w.retest) ++cnt->nodes;
    bool  h0, h1;
    float e0, e1;
    slab_pair(acc.nodes, w.link, c0, c1, r, inv, w.t_max, h0, h1, e0, e1);
    h0 = h0 && !w.retest;
    if (h0) {
        if (h1) {
            stack.push(w.link); // right child pending: re-tested against the t_max of that moment
        }
        w.link   = c0.child;
        w.count  = c0.count;
        w.retest = false;
    } else if (h1) {
        w.link   = c1.child;
        w.count  = c1.count;
        w.retest = false;
    } else if (stack.n > 0) {
        w.link   = stack.pop();
        w.retest = true;
    } else {
        w.link = kDone;
    }
}

// precondition: at_leaf(w) for every lane of `mask`, the lanes that call together
template <bool kCount, typename Prims, typename StackS>
__device__ __forceinline__ void closest_leaf_step(const Prims& prims, const Ray& r, ClosestWalk& w, StackS& stack,
                                                  TraceCounters* cnt, unsigned mask)
{
    float          t, b, g;
    const uint32_t first = static_cast<uint32_t>(~w.link);
    const uint32_t n     = w.count & SPCU_LEAF_COUNT_MASK;
    const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
    const uint32_t n_max = __reduce_max_sync(mask, n);
#pragma unroll 1
    for (uint32_t i = 0; i < n_max; ++i) {
        if (i < n && prims.template test<kCount>(first + i, mixed, r, w.t_max, t, b, g, cnt)) {
            w.t_max  = t;
            w.hit_id = static_cast<int32_t>(first + i);
            w.beta   = b;
            w.gamma  = g;
        }
    }
    if (stack.n > 0) {
        w.link   = stack.pop();
        w.retest = true;
    } else {
        w.link = kDone;
    }
}

// One-shot form: `t_max` enters as the query's limit and leaves as the accepted distance; returns the primitive or -1.
template <bool kCount, typename Prims>
__device__ __forceinline__ int32_t closest_hit(const DAccel& acc, const Prims& prims, const Ray& r, float& t_max,
                                               float& beta, float& gamma, int32_t* stack_smem, TraceCounters* cnt)
{
    ClosestWalk w;
    w.t_max = t_max;
    closest_begin<kCount>(acc, prims, r, w, cnt);
    const RayInv inv = make_inv(r, acc.proper_boxes != 0u);
    Stack        stack;
    stack.sh = stack_smem;
    closest_run<kCount>(acc, prims, r, inv, w, stack, kAllLeaves, cnt, __activemask());
    t_max = w.t_max;
    beta  = w.beta;
    gamma = w.gamma;
    return w.hit_id;
}

// ---------------------------------------------------------------------------------------------------------------------
// Ordered ("fast") closest hit: same boxes, same primitive tests, same arithmetic, but at every internal node the child
// whose box the ray enters first is visited first and the other is deferred with its entry distance; a deferred child
// is dropped when popped if the hit found meanwhile is closer than that entry.  The closest hit is the same as the
// reference-order walk finds, with two provisos that the parity tests count and state:
//  * equal-t candidates: the reference keeps the LAST one it visits, and IDs are assigned in its visiting order, so
//    ties are resolved here towards the higher ID — the same answer whenever the reference visits all tied candidates;
//  * the slab test is not conservative with respect to the primitive tests (different roundings), and it is evaluated
//    against whatever t_max the walk has reached; another visiting order can therefore cull, or fail to cull, a box
//    whose primitive grazes the current hit distance.  These are the "epsilon-tie mismatches".
// Stack entry = (node << 1 | child, entry distance): 8 bytes, first kStackSharedOrdered levels in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStackSharedOrdered = 12; // 12 levels x 128 threads x 8 B = 12 KB, the same footprint as the exact walk

template <int kShared, int kCapacity>
struct OrderedStackT
{
    int2* sh; // &smem[threadIdx.x]
    int2  loc[kCapacity - kShared];
    int   n = 0;

    __device__ __forceinline__ void attach(int32_t* stack_smem)
    {
        sh = reinterpret_cast<int2*>(stack_smem - threadIdx.x) + threadIdx.x; // the same shared block, 8-byte entries
    }
    __device__ __forceinline__ void push(int32_t key, float t0)
    {
        const int2 v = make_int2(key, __float_as_int(t0));
        if (n < kShared) {
            sh[n * kTraceBlock] = v;
        } else {
            loc[n - kShared] = v;
        }
        ++n;
    }
    __device__ __forceinline__ int2 pop()
    {
        --n;
        return (n < kShared) ? sh[n * kTraceBlock] : loc[n - kShared];
    }
};
using OrderedStack = OrderedStackT<kStackSharedOrdered, SPCU_MAX_BVH_DEPTH + 2>;

// next deferred child that is still reachable, or kDone
template <typename StackS>
__device__ __forceinline__ void ordered_pop(const DAccel& acc, StackS& stack, ClosestWalk& w)
{
    w.link = kDone;
    while (stack.n > 0) {
        const int2 e = stack.pop();
        if (!(__int_as_float(e.y) > w.t_max)) { // entry not beyond the current hit
            const float4 v3 = __ldg(acc.nodes + 4 * (e.x >> 1) + 3);
            w.link          = (e.x & 1) ? __float_as_int(v3.y) : __float_as_int(v3.x);
            w.count         = (e.x & 1) ? __float_as_uint(v3.w) : __float_as_uint(v3.z);
            return;
        }
    }
}

// one step of the ordered walk at an internal node (closest_run_ordered's node loop body)
template <bool kCount, typename StackS>
__device__ __forceinline__ void closest_node_step_ordered(const DAccel& acc, const Ray& r, const RayInv& inv, ClosestWalk& w,
                                                          StackS& stack, TraceCounters* cnt)
{
    NodeHalf c0, c1;
    load_node(acc.nodes, w.link, c0, c1);
    if (kCount) ++cnt->nodes;
    bool  h0, h1;
    float e0, e1;
    slab_pair(acc.nodes, w.link, c0, c1, r, inv, w.t_max, h0, h1, e0, e1);
    if (h0 && h1) {
        const bool left_first = !(e1 < e0); // equal entries: the reference's order, left first
        stack.push((w.link << 1) | (left_first ? 1 : 0), left_first ? e1 : e0);
        w.link  = left_first ? c0.child : c1.child;
        w.count = left_first ? c0.count : c1.count;
    } else if (h0) {
        w.link  = c0.child;
        w.count = c0.count;
    } else if (h1) {
        w.link  = c1.child;
        w.count = c1.count;
    } else {
        ordered_pop(acc, stack, w);
    }
}

template <bool kCount, typename Prims, typename StackS>
__device__ __forceinline__ void closest_run_ordered(const DAccel& acc, const Prims& prims, const Ray& r, const RayInv& inv,
                                                    ClosestWalk& w, StackS& stack, int max_leaves, TraceCounters* cnt,
                                                    unsigned mask)
{
    float t, b, g;
    if (!Prims::kBvh) { // no internal nodes: nothing to order, the root leaf in list order (ties: later primitive wins)
        if (w.link != kDone) {
            const uint32_t first = static_cast<uint32_t>(~w.link), n = w.count & SPCU_LEAF_COUNT_MASK;
            for (uint32_t i = 0; i < n; ++i) {
                if (prims.template test<kCount>(first + i, true, r, w.t_max, t, b, g, cnt)) {
                    w.t_max  = t;
                    w.hit_id = static_cast<int32_t>(first + i);
                    w.beta   = b;
                    w.gamma  = g;
                }
            }
            w.link = kDone;
        }
        return;
    }
#pragma unroll 1
    while (__any_sync(mask, w.link != kDone) && max_leaves > 0) {
#pragma unroll 1
        while (__any_sync(mask, w.link >= 0 && w.link != kDone)) {
            if (w.link >= 0 && w.link != kDone) {
                NodeHalf c0, c1;
                load_node(acc.nodes, w.link, c0, c1);
                if (kCount) ++cnt->nodes;
                bool  h0, h1;
                float e0, e1;
                slab_pair(acc.nodes, w.link, c0, c1, r, inv, w.t_max, h0, h1, e0, e1);
                if (h0 && h1) {
                    const bool left_first = !(e1 < e0); // equal entries: the reference's order, left first
                    stack.push((w.link << 1) | (left_first ? 1 : 0), left_first ? e1 : e0);
                    w.link  = left_first ? c0.child : c1.child;
                    w.count = left_first ? c0.count : c1.count;
                } else if (h0) {
                    w.link  = c0.child;
                    w.count = c0.count;
                } else if (h1) {
                    w.link  = c1.child;
                    w.count = c1.count;
                } else {
                    ordered_pop(acc, stack, w);
                }
            }
        }
        const bool     leaf  = w.link != kDone;
        const uint32_t first = static_cast<uint32_t>(~w.link);
        const uint32_t n     = leaf ? (w.count & SPCU_LEAF_COUNT_MASK) : 0u;
        const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
        const uint32_t n_max = __reduce_max_sync(mask, n);
#pragma unroll 1
        for (uint32_t i = 0; i < n_max; ++i) {
            const int32_t id = static_cast<int32_t>(first + i);
            if (i < n && prims.template test<kCount>(first + i, mixed, r, w.t_max, t, b, g, cnt) && (t < w.t_max || id > w.hit_id)) {
                w.t_max  = t;
                w.hit_id = id;
                w.beta   = b;
                w.gamma  = g;
            }
        }
        if (leaf) {
            ordered_pop(acc, stack, w);
        }
        --max_leaves;
    }
}

template <bool kCount, typename Prims>
__device__ __forceinline__ int32_t closest_hit_ordered(const DAccel& acc, const Prims& prims, const Ray& r, float& t_max,
                                                       float& beta, float& gamma, int32_t* stack_smem, TraceCounters* cnt)
{
    ClosestWalk w;
    w.t_max = t_max;
    closest_begin<kCount>(acc, prims, r, w, cnt); // the unbounded list is scanned first, in order, as in the reference
    const RayInv inv = make_inv(r, acc.proper_boxes != 0u);
    OrderedStack stack;
    stack.attach(stack_smem);
    closest_run_ordered<kCount>(acc, prims, r, inv, w, stack, kAllLeaves, cnt, __activemask());
    t_max = w.t_max;
    beta  = w.beta;
    gamma = w.gamma;
    return w.hit_id;
}

// ---------------------------------------------------------------------------------------------------------------------
// Any hit over one accelerator (ListAccelerator::intersect_p_impl :64-67, NodeInternal::intersect_p :79-90):
// limits never change, so a right child that passes its box test when its parent is visited needs no re-test: the
// stack holds nodes whose right child is simply entered when popped.  First accepted primitive ends the query.
// Resumable like the closest walk: any_run returns kAnyRunning / kAnyHit / kAnyMiss.
// ---------------------------------------------------------------------------------------------------------------------
struct AnyWalk
{
    int32_t  link;
    uint32_t count;
};
constexpr int kAnyRunning = 0, kAnyHit = 1, kAnyMiss = 2;

template <typename StackS>
__device__ __forceinline__ void any_pop(const DAccel& acc, StackS& stack, AnyWalk& w)
{
    if (stack.n > 0) {
        const float4 v3 = __ldg(acc.nodes + 4 * stack.pop() + 3);
        w.link          = __float_as_int(v3.y);
        w.count         = __float_as_uint(v3.w);
    } else {
        w.link = kDone;
    }
}

template <bool kCount, bool kBvh, typename AnyTest>
__device__ __forceinline__ int any_run(const DAccel& acc, const AnyTest& test, const Ray& r, const RayInv& inv, float t_max,
                                       AnyWalk& w, Stack& stack, int max_leaves, TraceCounters* cnt, unsigned mask)
{
    bool hit = false;
    if (!kBvh) { // the root is a leaf
        if (w.link != kDone) {
            const uint32_t first = static_cast<uint32_t>(~w.link), n = w.count & SPCU_LEAF_COUNT_MASK;
            for (uint32_t i = 0; i < n && !hit; ++i) {
                hit = test(first + i, true, cnt);
            }
            w.link = kDone;
        }
        return hit ? kAnyHit : kAnyMiss;
    }
#pragma unroll 1
    while (__any_sync(mask, w.link != kDone) && max_leaves > 0) {
#pragma unroll 1
        while (__any_sync(mask, w.link >= 0 && w.link != kDone)) {
            if (w.link >= 0 && w.link != kDone) {
                NodeHalf c0, c1;
                load_node(acc.nodes, w.link, c0, c1);
                if (kCount) ++cnt->nodes;
                bool  h0, h1;
                float e0, e1;
                slab_pair(acc.nodes, w.link, c0, c1, r, inv, t_max, h0, h1, e0, e1);
                if (h0) {
                    if (h1) {
                        stack.push(w.link);
                    }
                    w.link  = c0.child;
                    w.count = c0.count;
                } else if (h1) {
                    w.link  = c1.child;
                    w.count = c1.count;
                } else {
                    any_pop(acc, stack, w);
                }
            }
        }
        const bool     leaf  = w.link != kDone;
        const uint32_t first = static_cast<uint32_t>(~w.link);
        const uint32_t n     = leaf ? (w.count & SPCU_LEAF_COUNT_MASK) : 0u;
        const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
        const uint32_t n_max = __reduce_max_sync(mask, n);
#pragma unroll 1
        for (uint32_t i = 0; i < n_max; ++i) {
            if (i < n && !hit && test(first + i, mixed, cnt)) {
                hit = true; // first accepted primitive ends this lane's query; the others finish theirs
            }
        }
        if (hit) {
            w.link = kDone;
        } else if (leaf) {
            any_pop(acc, stack, w);
        }
        --max_leaves;
    }
    return hit ? kAnyHit : (w.link == kDone ? kAnyMiss : kAnyRunning);
}

template <bool kCount, bool kBvh, typename AnyTest>
__device__ __forceinline__ bool any_hit(const DAccel& acc, const AnyTest& test, const Ray& r, float t_max,
                                        int32_t* stack_smem, TraceCounters* cnt)
{
    for (uint32_t i = 0; i < acc.n_unbounded; ++i) {
        if (test(i, true, cnt)) {
            return true;
        }
    }
    const RayInv inv = make_inv(r, acc.proper_boxes != 0u);
    Stack        stack;
    stack.sh = stack_smem;
    AnyWalk w{ acc.root, acc.root_count };
    return any_run<kCount, kBvh>(acc, test, r, inv, t_max, w, stack, kAllLeaves, cnt, __activemask()) == kAnyHit;
}

// single steps of the any-hit walk (see closest_node_step)
__device__ __forceinline__ bool at_node(const AnyWalk& w) { return w.link >= 0 && w.link != kDone; }
__device__ __forceinline__ bool at_leaf(const AnyWalk& w) { return w.link < 0; }

template <bool kCount, typename StackS>
__device__ __forceinline__ void any_node_step(const DAccel& acc, const Ray& r, const RayInv& inv, float t_max, AnyWalk& w,
                                              StackS& stack, TraceCounters* cnt)
{
    NodeHalf c0, c1;
    load_node(acc.nodes, w.link, c0, c1);
    if (kCount) ++cnt->nodes;
    bool  h0, h1;
    float e0, e1;
    slab_pair(acc.nodes, w.link, c0, c1, r, inv, t_max, h0, h1, e0, e1);
    if (h0) {
        if (h1) {
            stack.push(w.link);
        }
        w.link  = c0.child;
        w.count = c0.count;
    } else if (h1) {
        w.link  = c1.child;
        w.count = c1.count;
    } else {
        any_pop(acc, stack, w);
    }
}

// returns true when a primitive of the leaf is hit (the query is over); otherwise the walk moves on
__device__ __forceinline__ void wide_unpack(const DAccel& acc, int32_t packed, int32_t& link, uint32_t& count);

// (`wide_keys`: the stack holds keys of 4-wide nodes — see "4-wide nodes" below — instead of binary node indices)
template <typename AnyTest, typename StackS>
__device__ __forceinline__ bool any_leaf_step(const DAccel& acc, const AnyTest& test, AnyWalk& w, StackS& stack,
                                              TraceCounters* cnt, unsigned mask, bool wide_keys = false)
{
    const uint32_t first = static_cast<uint32_t>(~w.link);
    const uint32_t n     = w.count & SPCU_LEAF_COUNT_MASK;
    const bool     mixed = (w.count & SPCU_LEAF_MIXED_FLAG) != 0u;
    const uint32_t n_max = __reduce_max_sync(mask, n);
    bool           hit   = false;
#pragma unroll 1
    for (uint32_t i = 0; i < n_max; ++i) {
        if (i < n && !hit && test(first + i, mixed, cnt)) {
            hit = true;
        }
    }
    if (hit) {
        w.link = kDone;
    } else if (!wide_keys) {
        any_pop(acc, stack, w);
    } else if (stack.n > 0) {
        wide_unpack(acc, stack.pop(), w.link, w.count);
    } else {
        w.link = kDone;
    }
    return hit;
}

// ---------------------------------------------------------------------------------------------------------------------
// 4-wide nodes: the reference topology with every other level folded away (device_scene.h DAccel::wide, built at upload
// by k_build_wide, build_kernels.cu).  Wide node i belongs to binary node i and holds the boxes of its GRANDCHILDREN — or of
// a child where that child is a leaf —, in the reference's left-to-right order:
//     child k = the 32-byte sector k:  { lo.x lo.y lo.z hi.x } { hi.y hi.z packed count }
//     (an unused slot is a point box at +3e38: never entered; `packed` = the child in one word, device_scene.h; `count` as in
//     spcu_bvh_node, kept for inspection)
// 128 bytes = four 256-bit loads.  A stack entry is the child's PACKED word: the pop that follows a leaf or a dead end decodes
// it in registers and goes straight for the node or the triangles.  (The first wide walk kept (node, slot) keys and paid one
// dependent 8-byte load per pop — an L1 miss more often than not, the warps of an SM sweep its L1 once per round of steps.)
// The walks of the render's traversal stages are bound by the LATENCY of one dependent node
// fetch per step and by the L1's look-up rate for scattered sectors, not by arithmetic (ncu, profiles/r02b_*: long-scoreboard
// and fixed-latency stalls 7.4 of 11.4 cycles per issued instruction, l1tex at 69 % of peak); a wide step does the work of
// two binary levels behind ONE fetch, and its four slab tests are independent instruction streams.
// Skipping the child's own box is exact wherever the slab test is monotone under box inclusion — the child's box contains
// its children's, so it passes whenever one of them does — i.e. for every ray and box of the NaN-free form (slab_fast);
// rays with an infinite reciprocal take the binary walk over the reference's nodes (the `generic` flag).
// Only the ORDERED closest-hit walk and the any-hit walk use wide nodes: their answers do not depend on the visiting order
// (up to the epsilon ties stated for the ordered walk).  The exact walk stays on the binary nodes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWideStackCapacity = 3 * (SPCU_MAX_BVH_DEPTH / 2 + 1) + 2; // up to three deferred children per wide level
#ifndef SPCU_WIDE_SHARED_ANY
#define SPCU_WIDE_SHARED_ANY 24
#endif
#ifndef SPCU_WIDE_SHARED_ORDERED
#define SPCU_WIDE_SHARED_ORDERED 16
#endif
constexpr int kWideSharedAny     = SPCU_WIDE_SHARED_ANY;     // 4-byte packed children: [24][128] = the 12 KB block
constexpr int kWideSharedOrdered = SPCU_WIDE_SHARED_ORDERED; // 8-byte (packed child, entry distance): needs [32][128] words
using WideStack        = StackT<kWideSharedAny, kWideStackCapacity>;
using WideOrderedStack = OrderedStackT<kWideSharedOrdered, kWideStackCapacity>;

struct WideNode
{
    Float8 q0, q1, q2, q3;
};

__device__ __forceinline__ WideNode load_wide(const float4* wide, int32_t idx)
{
    const float4* p = wide + 8 * static_cast<size_t>(idx);
    return WideNode{ ldg256(p), ldg256(p + 2), ldg256(p + 4), ldg256(p + 6) };
}

// the four slab tests of a wide node (NaN-free form); h = hit mask, e[k] = entry distance of box k (valid where hit)
__device__ __forceinline__ bool wide_slab(const Float8& q, const Ray& r, const RayInv& inv, float t_max, float& e)
{
    return slab_fast(q.lo.x, q.lo.y, q.lo.z, q.lo.w, q.hi.x, q.hi.y, r, inv, t_max, e);
}

__device__ __forceinline__ unsigned wide_slabs(const WideNode& n, const Ray& r, const RayInv& inv, float t_max, float (&e)[4])
{
    const bool h0 = wide_slab(n.q0, r, inv, t_max, e[0]);
    const bool h1 = wide_slab(n.q1, r, inv, t_max, e[1]);
    const bool h2 = wide_slab(n.q2, r, inv, t_max, e[2]);
    const bool h3 = wide_slab(n.q3, r, inv, t_max, e[3]);
    return (h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u);
}

// packed child (device_scene.h) -> the cursor's { link, count } (link / count as in spcu_bvh_node)
__device__ __forceinline__ void wide_unpack(const DAccel& acc, int32_t packed, int32_t& link, uint32_t& count)
{
    link = packed;
    if (packed < 0) {
        const uint32_t p = static_cast<uint32_t>(packed), tag = (p >> 27) & 7u;
        if (tag != kWideBigLeafTag) {
            link  = ~static_cast<int32_t>(p & kWidePayloadMask);
            count = tag | ((p << 1) & SPCU_LEAF_MIXED_FLAG);
        } else { // a leaf of more than kWideSmallLeafMax primitives: the side table (rare)
            const int2 lc = __ldg(acc.big + (p & kWidePayloadMask));
            link          = lc.x;
            count         = static_cast<uint32_t>(lc.y);
        }
    }
}

// The cursor has just become a leaf: start its triangle records on their way to L1 now — the leaf step that tests them runs a few
// votes later (it waits until enough lanes hold a leaf), and would otherwise begin with a full-latency miss.
__device__ __forceinline__ void prefetch_leaf(const float4* prims, int32_t link, uint32_t count)
{
#ifndef SPCU_NO_LEAF_PREFETCH
    if (link < 0 && (count & SPCU_LEAF_COUNT_MASK) != 0u) {
        const float4* p = prims + 3 * static_cast<size_t>(static_cast<uint32_t>(~link));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 3 * ((count & SPCU_LEAF_COUNT_MASK) - 1u) + 2)); // its last 16 bytes
    }
#endif
}

// next deferred child that is still reachable, or kDone (wide keys)
__device__ __forceinline__ void wide_ordered_pop(const DAccel& acc, WideOrderedStack& stack, ClosestWalk& w)
{
    w.link = kDone;
    while (stack.n > 0) {
        const int2 e = stack.pop();
        if (!(__int_as_float(e.y) > w.t_max)) { // entry not beyond the current hit
            wide_unpack(acc, e.x, w.link, w.count);
            return;
        }
    }
}

// One step of the ordered walk at an internal node, over its wide node: the nearest of the (up to four) boxes the ray enters
// becomes the cursor, the others wait on the stack, farthest first (so the nearest of them is popped first).  Equal entry
// distances keep the reference's left-to-right order.
#define SPCU_CSWAP(ka, sa, kb, sb)        \
    {                                     \
        const bool    sw = kb < ka;       \
        const float   tk = sw ? kb : ka;  \
        const int32_t ts = sw ? sb : sa;  \
        kb               = sw ? ka : kb;  \
        sb               = sw ? sa : sb;  \
        ka               = tk;            \
        sa               = ts;            \
    }
template <bool kCount>
__device__ __forceinline__ void closest_wide_step_ordered(const DAccel& acc, const Ray& r, const RayInv& inv, ClosestWalk& w,
                                                          WideOrderedStack& stack, TraceCounters* cnt)
{
    const WideNode n = load_wide(acc.wide, w.link);
    if (kCount) ++cnt->nodes;
    float          e[4];
    const unsigned h = wide_slabs(n, r, inv, w.t_max, e);
    if (h == 0u) {
        wide_ordered_pop(acc, stack, w);
        return;
    }
    // sort (distance, packed child) ascending; a box that is not entered sorts last (+inf)
    const float inf = __int_as_float(0x7f800000);
    float   k0 = (h & 1u) ? e[0] : inf, k1 = (h & 2u) ? e[1] : inf, k2 = (h & 4u) ? e[2] : inf, k3 = (h & 8u) ? e[3] : inf;
    int32_t s0 = __float_as_int(n.q0.hi.z), s1 = __float_as_int(n.q1.hi.z), s2 = __float_as_int(n.q2.hi.z), s3 = __float_as_int(n.q3.hi.z);
    SPCU_CSWAP(k0, s0, k1, s1)
    SPCU_CSWAP(k2, s2, k3, s3)
    SPCU_CSWAP(k0, s0, k2, s2)
    SPCU_CSWAP(k1, s1, k3, s3)
    SPCU_CSWAP(k1, s1, k2, s2)
    if (k3 < inf) stack.push(s3, k3);
    if (k2 < inf) stack.push(s2, k2);
    if (k1 < inf) stack.push(s1, k1);
    wide_unpack(acc, s0, w.link, w.count);
}

__device__ __forceinline__ void wide_any_pop(const DAccel& acc, WideStack& stack, AnyWalk& w)
{
    if (stack.n > 0) {
        wide_unpack(acc, stack.pop(), w.link, w.count);
    } else {
        w.link = kDone;
    }
}

// One step of the any-hit walk over a wide node: the limits never change, so the order is free — the first entered box
// becomes the cursor, the others wait.
template <bool kCount>
__device__ __forceinline__ void any_wide_step(const DAccel& acc, const Ray& r, const RayInv& inv, float t_max, AnyWalk& w,
                                              WideStack& stack, TraceCounters* cnt)
{
    const WideNode n = load_wide(acc.wide, w.link);
    if (kCount) ++cnt->nodes;
    float          e[4];
    const unsigned h = wide_slabs(n, r, inv, t_max, e);
    if (h == 0u) {
        wide_any_pop(acc, stack, w);
        return;
    }
    const int32_t s0 = __float_as_int(n.q0.hi.z), s1 = __float_as_int(n.q1.hi.z), s2 = __float_as_int(n.q2.hi.z), s3 = __float_as_int(n.q3.hi.z);
#ifdef SPCU_ANY_SORTED // (A/B: measured slower — elf shadow 11.9 -> 12.7 ms, profiles/r02m_ab_any_hit_order.jsonl — and left off)
    // Nearest box first, although any order gives the same answer: an occluded ray — half of the shadow rays of a path-traced
    // frame — usually meets its occluder in the nearer boxes, and the walk ends at the first accepted primitive.
    const float inf = __int_as_float(0x7f800000);
    float   k0 = (h & 1u) ? e[0] : inf, k1 = (h & 2u) ? e[1] : inf, k2 = (h & 4u) ? e[2] : inf, k3 = (h & 8u) ? e[3] : inf;
    int32_t c0 = s0, c1 = s1, c2 = s2, c3 = s3;
    SPCU_CSWAP(k0, c0, k1, c1)
    SPCU_CSWAP(k2, c2, k3, c3)
    SPCU_CSWAP(k0, c0, k2, c2)
    SPCU_CSWAP(k1, c1, k3, c3)
    SPCU_CSWAP(k1, c1, k2, c2)
    if (k3 < inf) stack.push(c3);
    if (k2 < inf) stack.push(c2);
    if (k1 < inf) stack.push(c1);
    wide_unpack(acc, c0, w.link, w.count);
#else
    // the first entered box (in slot order) becomes the cursor, the others are pushed in slot order
    const int32_t first = (h & 1u) ? s0 : (h & 2u) ? s1 : (h & 4u) ? s2 : s3;
    const unsigned rest = h & (h - 1u);
    if (rest & 2u) stack.push(s1);
    if (rest & 4u) stack.push(s2);
    if (rest & 8u) stack.push(s3);
    wide_unpack(acc, first, w.link, w.count);
#endif
}
#undef SPCU_CSWAP

__device__ __forceinline__ RayInv make_inv_wide(const Ray& r, const DAccel& acc)
{
    return make_inv(r, acc.proper_boxes != 0u && acc.wide != nullptr);
}

// does the ray enter ANY box of the root's wide node?  (what the `begin` kernels ask before they park a ray for the walk)
__device__ __forceinline__ bool wide_root_entered(const DAccel& acc, const Ray& r, const RayInv& inv, float t_max)
{
    float e[4];
    return wide_slabs(load_wide(acc.wide, acc.root), r, inv, t_max, e) != 0u;
}

// Lights half of Scene::intersect_p: only sphere lights occlude.
template <typename F = FeatFull>
__device__ __forceinline__ bool lights_any_hit(const DScene& s, const Ray& r, float t_max, int32_t* stack_smem)
{
    const LightPrimsT<F> lp{ s.lights };
    auto light_test = [&](uint32_t id, bool, TraceCounters*) { return lp.test_any(id, r, t_max); };
    return any_hit<false, F::bvh>(s.lights_accel, light_test, r, t_max, stack_smem, nullptr);
}

// Scene::intersect_p (base/Scene.h:79-82): geometry accelerator, then lights accelerator.
template <bool kCount, typename F = FeatFull>
__device__ __forceinline__ bool scene_any_hit(const DScene& s, const Ray& r, float t_max, int32_t* stack_smem,
                                              TraceCounters* cnt)
{
    const GeomPrimsT<F> gp{ s.geom_prims, s.geom_meta };
    auto geom_test = [&](uint32_t id, bool mixed, TraceCounters* c) {
        float t, b, g;
        return gp.template test<kCount>(id, mixed, r, t_max, t, b, g, c);
    };
    if (any_hit<kCount, F::bvh>(s.geom, geom_test, r, t_max, stack_smem, cnt)) {
        return true;
    }
    return lights_any_hit<F>(s, r, t_max, stack_smem);
}

} // namespace spcu
