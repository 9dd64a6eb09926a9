/*
 * TEST INFRASTRUCTURE — see sp_oracle.h.  Plain-C restatement of the reference's hot path over the flattened
 * scene.  Compiled with -ffp-contract=off -mfma: the ONLY fused multiply-adds are the explicit fmaf() calls
 * that restate the reference's madd/msub (math/Math.h:137-162 == std::fma under -mavx2), which is the
 * arithmetic contract the canonical reference build (oracle/_ref/libsp_ref.so) follows.
 *
 * Each function cites the reference lines it restates.
 */
#include "sp_oracle.h"

#include <float.h>
#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SP_PI 3.14159265358979323846f /* std::numbers::pi_v<float> */
#define SP_INV_PI 0.318309886183790671538f
#define K_RAY_EPSILON 0.001f /* math/Ray.h:11 */
#define K_INFINITE FLT_MAX    /* base/Constants.h:16 */

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }

/* ===================================================================================================
 * Vector kernel (math/Vector3.h, math/Math.h)
 * =================================================================================================== */

/* dot = _mm_dp_ps(a, b, 0x7F) (math/Vector3.h:742-746): three rounded products, summed (x+y)+(z+0). */
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + 0.0f); }

/* rsqrt (math/Math.h:205-227): hardware estimate + one Newton step, same operation order. */
static inline float sp_rsqrt(float x)
{
    const __m128 a = _mm_set_ss(x);
    __m128       r = _mm_rsqrt_ss(a);
    const __m128 c = _mm_add_ss(_mm_mul_ss(_mm_set_ss(1.5f), r),
                                _mm_mul_ss(_mm_mul_ss(_mm_mul_ss(a, _mm_set_ss(-0.5f)), r), _mm_mul_ss(r, r)));
    return _mm_cvtss_f32(c);
}

/* normalize (math/Vector3.h:797-805) */
static inline v3 normalize3(v3 a)
{
    const float s = sp_rsqrt(dot3(a, a));
    return V(a.x * s, a.y * s, a.z * s);
}

static inline v3 add3(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 scale3(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg3(v3 a) { return V(-a.x, -a.y, -a.z); }

/* AffineSpace::operator()(Point3) (math/AffineSpace.h:79-86): fma chain per component.  m = 12 floats c0 c1 c2 a. */
static inline v3 xf_point(const float* m, v3 p)
{
    return V(fmaf(p.x, m[0], fmaf(p.y, m[3], fmaf(p.z, m[6], m[9]))),
             fmaf(p.x, m[1], fmaf(p.y, m[4], fmaf(p.z, m[7], m[10]))),
             fmaf(p.x, m[2], fmaf(p.y, m[5], fmaf(p.z, m[8], m[11]))));
}

/* LinearSpace3x3::operator()(Vector3) (math/LinearSpace3x3.h:153-161): m = 9 floats c0 c1 c2 */
static inline v3 xf_vector(const float* m, v3 v)
{
    return V(fmaf(v.x, m[0], fmaf(v.y, m[3], v.z * m[6])),
             fmaf(v.x, m[1], fmaf(v.y, m[4], v.z * m[7])),
             fmaf(v.x, m[2], fmaf(v.y, m[5], v.z * m[8])));
}

/* get_ray_offset (math/Ray.h:51-85) */
static inline float ray_offset_cos(float cos_d) { return cos_d == 0.0f ? K_RAY_EPSILON : K_RAY_EPSILON / cos_d; }
static inline float ray_offset(v3 n, v3 d) { return ray_offset_cos(fabsf(dot3(n, d))); }

/* ===================================================================================================
 * Intersection (math/BBox.h, shapes/*.h)
 * =================================================================================================== */
typedef struct {
    v3    o, d;
    float t_min, t_max;
} ray_t;

/* sp::intersect_p(BBox, Ray, RayLimits) (math/BBox.h:122-146).  std::max(a,b) = (a<b)?b:a and
 * std::min(a,b) = (b<a)?b:a: a NaN t_near REPLACES t0.  Do not use fmaxf/fminf here. */
static inline int slab(const float* lo, const float* hi, const ray_t* r, float t_min, float t_max)
{
    const float o[3] = { r->o.x, r->o.y, r->o.z };
    const float d[3] = { r->d.x, r->d.y, r->d.z };
    float       t0 = t_min, t1 = t_max;
    for (int i = 0; i < 3; ++i) {
        const float inv    = 1.0f / d[i];
        float       t_near = (lo[i] - o[i]) * inv;
        float       t_far  = (hi[i] - o[i]) * inv;
        if (t_near > t_far) {
            const float tmp = t_near;
            t_near          = t_far;
            t_far           = tmp;
        }
        t0 = (t_near < t0) ? t0 : t_near; /* std::max(t_near, t0) */
        t1 = (t1 < t_far) ? t1 : t_far;   /* std::min(t_far, t1)  */
        if (t0 > t1) {
            return 0;
        }
    }
    return 1;
}

/* Triangle::intersect_impl (shapes/Triangle.h:97-146).  Returns 1 and *t, *beta, *gamma on a hit. */
static inline int tri_hit(const float* g, const ray_t* r, float t_min, float t_max, float* t_out, float* b_out, float* c_out)
{
    const float A = g[0] - g[4], B = g[1] - g[5], C = g[2] - g[6];
    const float D = g[0] - g[8], E = g[1] - g[9], F = g[2] - g[10];
    const float G = r->d.x, H = r->d.y, I = r->d.z;
    const float J = g[0] - r->o.x, K = g[1] - r->o.y, L = g[2] - r->o.z;

    const float EIHF = fmaf(E, I, -(H * F));
    const float GFDI = fmaf(G, F, -(D * I));
    const float DHEG = fmaf(D, H, -(E * G));

    const float denom = fmaf(A, EIHF, fmaf(B, GFDI, C * DHEG));
    if (denom == 0) {
        return 0;
    }
    const float beta = fmaf(J, EIHF, fmaf(K, GFDI, L * DHEG)) / denom;
    if (beta <= 0.0f || beta >= 1.0f) {
        return 0;
    }
    const float AKJB = fmaf(A, K, -(J * B));
    const float JCAL = fmaf(J, C, -(A * L));
    const float BLKC = fmaf(B, L, -(K * C));

    const float gamma = fmaf(I, AKJB, fmaf(H, JCAL, G * BLKC)) / denom;
    if (gamma <= 0.0f || beta + gamma >= 1.0f) {
        return 0;
    }
    const float t = -fmaf(F, AKJB, fmaf(E, JCAL, D * BLKC)) / denom;
    if (t < t_min || t > t_max) {
        return 0;
    }
    *t_out = t;
    *b_out = beta;
    *c_out = gamma;
    return 1;
}

/* Sphere::intersect_impl (shapes/Sphere.h:77-109): m = world_to_object (12 floats).  Local o, d are returned
 * for the normal computation. */
static inline int sphere_hit(const float* m, const ray_t* r, float t_min, float t_max, float* t_out, v3* lo, v3* ld)
{
    const v3    o = xf_point(m, r->o);
    const v3    d = xf_vector(m, r->d);
    const float a = dot3(d, d);
    const float b = 2.0f * dot3(d, o);
    const float c = dot3(o, o) - 1.0f; /* k_radius * k_radius */
    float       disc = b * b - 4.0f * a * c;
    if (disc > 0.0f) {
        disc    = sqrtf(disc);
        float t = (-b - disc) / (2.0f * a);
        if (t < t_min) {
            t = (-b + disc) / (2.0f * a);
        }
        if (t < t_min || t > t_max) {
            return 0;
        }
        *t_out = t;
        if (lo) {
            *lo = o;
            *ld = d;
        }
        return 1;
    }
    return 0;
}

/* Plane::intersect_impl (shapes/Plane.h:21-71) */
static inline int plane_hit(const float* m, const ray_t* r, float t_min, float t_max, float* t_out)
{
    const v3 d = xf_vector(m, r->d);
    if (d.y == 0.0f) {
        return 0;
    }
    const v3    o = xf_point(m, r->o);
    const float t = -o.y / d.y;
    if (t < t_min || t > t_max) {
        return 0;
    }
    *t_out = t;
    return 1;
}

typedef struct {
    int32_t id;
    float   t;
    float   beta, gamma; /* triangle barycentrics of the accepted hit */
} closest_t;

/* GeometricPrimitive::intersect via ListAccelerator::intersect_impl's accept rule (shapes/ListAccelerator.h:50-62):
 * a hit replaces the result and shrinks t_max, so an equal-t later primitive wins. */
static inline void geom_prim(const spcu_flat_scene* s, uint32_t id, const ray_t* r, float* t_max, closest_t* c, spo_counters* cnt)
{
    const float*   g    = s->geom_prims[id].v;
    const uint32_t kind = SPCU_META_KIND(s->geom_meta[id]);
    float          t, b = 0.0f, gm = 0.0f;
    int            hit;
    if (kind == SPCU_PRIM_TRIANGLE) {
        if (cnt) ++cnt->tris;
        hit = tri_hit(g, r, r->t_min, *t_max, &t, &b, &gm);
    } else if (kind == SPCU_PRIM_SPHERE) {
        if (cnt) ++cnt->xf;
        hit = sphere_hit(g, r, r->t_min, *t_max, &t, NULL, NULL);
    } else {
        if (cnt) ++cnt->xf;
        hit = plane_hit(g, r, r->t_min, *t_max, &t);
    }
    if (hit) {
        *t_max   = t;
        c->id    = (int32_t)id;
        c->t     = t;
        c->beta  = b;
        c->gamma = gm;
    }
}

static inline int geom_prim_any(const spcu_flat_scene* s, uint32_t id, const ray_t* r, spo_counters* cnt)
{
    const float*   g    = s->geom_prims[id].v;
    const uint32_t kind = SPCU_META_KIND(s->geom_meta[id]);
    float          t, b, gm;
    if (kind == SPCU_PRIM_TRIANGLE) {
        if (cnt) ++cnt->tris;
        return tri_hit(g, r, r->t_min, r->t_max, &t, &b, &gm);
    }
    if (cnt) ++cnt->xf;
    if (kind == SPCU_PRIM_SPHERE) return sphere_hit(g, r, r->t_min, r->t_max, &t, NULL, NULL);
    return plane_hit(g, r, r->t_min, r->t_max, &t);
}

/* NodeInternal::intersect / NodeLeaf::intersect (shapes/BVHAccelerator.h:62-77,110-113): children left then right,
 * each child's own bounds tested against the CURRENT limits; the root's bounds are never tested (:138-142). */
static void geom_node(const spcu_flat_scene* s, int32_t link, uint32_t count, const ray_t* r, float* t_max, closest_t* c, spo_counters* cnt)
{
    if (link < 0) {
        const uint32_t first = (uint32_t)~link, n = count & SPCU_LEAF_COUNT_MASK;
        for (uint32_t i = 0; i < n; ++i) {
            geom_prim(s, first + i, r, t_max, c, cnt);
        }
        return;
    }
    const spcu_bvh_node* node = &s->geom.nodes[link];
    if (cnt) ++cnt->nodes;
    for (int k = 0; k < 2; ++k) {
        if (slab(node->box + 6 * k, node->box + 6 * k + 3, r, r->t_min, *t_max)) {
            geom_node(s, node->child[k], node->count[k], r, t_max, c, cnt);
        }
    }
}

static int geom_node_any(const spcu_flat_scene* s, int32_t link, uint32_t count, const ray_t* r, spo_counters* cnt)
{
    if (link < 0) {
        const uint32_t first = (uint32_t)~link, n = count & SPCU_LEAF_COUNT_MASK;
        for (uint32_t i = 0; i < n; ++i) {
            if (geom_prim_any(s, first + i, r, cnt)) return 1;
        }
        return 0;
    }
    const spcu_bvh_node* node = &s->geom.nodes[link];
    if (cnt) ++cnt->nodes;
    for (int k = 0; k < 2; ++k) {
        if (slab(node->box + 6 * k, node->box + 6 * k + 3, r, r->t_min, r->t_max) &&
            geom_node_any(s, node->child[k], node->count[k], r, cnt)) {
            return 1;
        }
    }
    return 0;
}

/* Scene::intersect -> ListAccelerator [unbounded..., BVH] (base/Scene.h:74-77, shapes/ListAccelerator.h:50-62) */
static closest_t scene_intersect(const spcu_flat_scene* s, const ray_t* r, spo_counters* cnt)
{
    closest_t c     = { -1, r->t_max, 0.0f, 0.0f };
    float     t_max = r->t_max;
    for (uint32_t i = 0; i < s->geom.n_unbounded; ++i) {
        geom_prim(s, i, r, &t_max, &c, cnt);
    }
    geom_node(s, s->geom.root, s->geom.root_count, r, &t_max, &c, cnt);
    return c;
}

/* ---- lights accelerator ------------------------------------------------------------------------ */
/* Light::intersect_lights_impl: SphereLight (Lights/Light.h:354-361), EnvironmentLight (:135-141),
 * ImageBasedEnvironmentLight (:196-209) */
static inline int light_hit(const spcu_light* l, const ray_t* r, float t_min, float t_max, float* t_out)
{
    if (l->kind == SPCU_LIGHT_SPHERE) {
        return sphere_hit(l->world_to_object, r, t_min, t_max, t_out, NULL, NULL);
    }
    if (t_max < K_INFINITE) {
        return 0;
    }
    *t_out = K_INFINITE;
    return 1;
}

static inline int light_hit_any(const spcu_light* l, const ray_t* r)
{
    float t;
    if (l->kind == SPCU_LIGHT_SPHERE) { /* SphereLight::intersect_p_impl :363-366 */
        return sphere_hit(l->world_to_object, r, r->t_min, r->t_max, &t, NULL, NULL);
    }
    return 0; /* environment lights never occlude (:143-146, :211-214) */
}

static void lights_node(const spcu_flat_scene* s, int32_t link, uint32_t count, const ray_t* r, float* t_max, closest_t* c)
{
    if (link < 0) {
        const uint32_t first = (uint32_t)~link, n = count & SPCU_LEAF_COUNT_MASK;
        for (uint32_t i = 0; i < n; ++i) {
            float t;
            if (light_hit(&s->lights[first + i], r, r->t_min, *t_max, &t)) {
                *t_max = t;
                c->id  = (int32_t)(first + i);
                c->t   = t;
            }
        }
        return;
    }
    const spcu_bvh_node* node = &s->lights_accel.nodes[link];
    for (int k = 0; k < 2; ++k) {
        if (slab(node->box + 6 * k, node->box + 6 * k + 3, r, r->t_min, *t_max)) {
            lights_node(s, node->child[k], node->count[k], r, t_max, c);
        }
    }
}

static int lights_node_any(const spcu_flat_scene* s, int32_t link, uint32_t count, const ray_t* r)
{
    if (link < 0) {
        const uint32_t first = (uint32_t)~link, n = count & SPCU_LEAF_COUNT_MASK;
        for (uint32_t i = 0; i < n; ++i) {
            if (light_hit_any(&s->lights[first + i], r)) return 1;
        }
        return 0;
    }
    const spcu_bvh_node* node = &s->lights_accel.nodes[link];
    for (int k = 0; k < 2; ++k) {
        if (slab(node->box + 6 * k, node->box + 6 * k + 3, r, r->t_min, r->t_max) &&
            lights_node_any(s, node->child[k], node->count[k], r)) {
            return 1;
        }
    }
    return 0;
}

/* Scene::intersect_lights (base/Scene.h:69-72) */
static closest_t scene_intersect_lights(const spcu_flat_scene* s, const ray_t* r)
{
    closest_t c     = { -1, r->t_max, 0.0f, 0.0f };
    float     t_max = r->t_max;
    for (uint32_t i = 0; i < s->lights_accel.n_unbounded; ++i) {
        float t;
        if (light_hit(&s->lights[i], r, r->t_min, t_max, &t)) {
            t_max = t;
            c.id  = (int32_t)i;
            c.t   = t;
        }
    }
    lights_node(s, s->lights_accel.root, s->lights_accel.root_count, r, &t_max, &c);
    return c;
}

/* Scene::intersect_p (base/Scene.h:79-82): geometry || lights */
static int scene_intersect_p(const spcu_flat_scene* s, const ray_t* r, spo_counters* cnt)
{
    for (uint32_t i = 0; i < s->geom.n_unbounded; ++i) {
        if (geom_prim_any(s, i, r, cnt)) return 1;
    }
    if (geom_node_any(s, s->geom.root, s->geom.root_count, r, cnt)) return 1;
    for (uint32_t i = 0; i < s->lights_accel.n_unbounded; ++i) {
        if (light_hit_any(&s->lights[i], r)) return 1;
    }
    return lights_node_any(s, s->lights_accel.root, s->lights_accel.root_count, r);
}

static inline ray_t load_ray(const spcu_ray* q)
{
    ray_t r = { V(q->ox, q->oy, q->oz), V(q->dx, q->dy, q->dz), q->t_min, q->t_max };
    return r;
}

/* ---- Intersection record (shapes/Intersection.h:17-23) ------------------------------------------ */
typedef struct {
    float    t;
    v3       normal, point;
    uint32_t material;
} isect_t;

/* Triangle.h:148-160, Sphere.h:99-104, Plane.h:65-70 */
static isect_t make_isect(const spcu_flat_scene* s, const ray_t* r, const closest_t* c)
{
    isect_t        is;
    const uint32_t meta = s->geom_meta[c->id];
    const float*   sh   = s->geom_shade[c->id].v;
    is.t                = c->t;
    is.material         = SPCU_META_MATERIAL(meta);
    /* Ray::operator() (math/Ray.h:30-34): origin + direction * d */
    is.point = add3(r->o, scale3(r->d, c->t));
    if (SPCU_META_KIND(meta) == SPCU_PRIM_TRIANGLE) {
        const float alpha = 1.0f - c->beta - c->gamma;
        const v3    n     = V(fmaf(alpha, sh[0], fmaf(c->beta, sh[4], c->gamma * sh[8])),
                              fmaf(alpha, sh[1], fmaf(c->beta, sh[5], c->gamma * sh[9])),
                              fmaf(alpha, sh[2], fmaf(c->beta, sh[6], c->gamma * sh[10])));
        is.normal         = normalize3(n);
    } else if (SPCU_META_KIND(meta) == SPCU_PRIM_SPHERE) {
        const float* m = s->geom_prims[c->id].v;
        const v3     o = xf_point(m, r->o);
        const v3     d = xf_vector(m, r->d);
        /* madd(t, d, o) / k_radius */
        const v3 n = V(fmaf(c->t, d.x, o.x) / 1.0f, fmaf(c->t, d.y, o.y) / 1.0f, fmaf(c->t, d.z, o.z) / 1.0f);
        const v3 w = V(fmaf(n.x, sh[0], fmaf(n.y, sh[4], n.z * sh[8])),
                       fmaf(n.x, sh[1], fmaf(n.y, sh[5], n.z * sh[9])),
                       fmaf(n.x, sh[2], fmaf(n.y, sh[6], n.z * sh[10])));
        is.normal  = normalize3(w);
    } else {
        /* get_object_to_world()(Normal3{0,1,0}) — NOT normalised (Plane.h:66) */
        is.normal = V(fmaf(0.0f, sh[0], fmaf(1.0f, sh[4], 0.0f * sh[8])),
                      fmaf(0.0f, sh[1], fmaf(1.0f, sh[5], 0.0f * sh[9])),
                      fmaf(0.0f, sh[2], fmaf(1.0f, sh[6], 0.0f * sh[10])));
    }
    return is;
}

/* ===================================================================================================
 * Public batch entry points
 * =================================================================================================== */
void spo_trace_closest(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits, spo_counters* cnt)
{
    spo_counters local = { 0, 0, 0 };
    if (!cnt) { /* rays are independent: the 2^20-ray parity batches of tests/test_gpu_scale.py use every host core */
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const ray_t     r = load_ray(&rays[i]);
            const closest_t c = scene_intersect(s, &r, NULL);
            hits[i].id        = c.id;
            hits[i].t         = c.t;
        }
        return;
    }
    for (uint64_t i = 0; i < n; ++i) {
        const ray_t     r = load_ray(&rays[i]);
        const closest_t c = scene_intersect(s, &r, &local);
        hits[i].id        = c.id;
        hits[i].t         = c.t;
    }
    *cnt = local;
}

void spo_trace_any(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, uint8_t* out)
{
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const ray_t r = load_ray(&rays[i]);
        out[i]        = (uint8_t)scene_intersect_p(s, &r, NULL);
    }
}

void spo_trace_lights(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits)
{
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const ray_t     r = load_ray(&rays[i]);
        const closest_t c = scene_intersect_lights(s, &r);
        hits[i].id        = c.id;
        hits[i].t         = c.t;
    }
}

/* PerspectiveCamera::generate_ray_impl (Cameras/Camera.h:119-129): normalize(px*col0 + py*col1 + col2) */
static ray_t camera_ray(const spcu_flat_scene* s, float px, float py)
{
    const float* m = s->camera;
    const v3     d = V((px * m[0] + py * m[3]) + m[6], (px * m[1] + py * m[4]) + m[7], (px * m[2] + py * m[5]) + m[8]);
    ray_t        r = { V(m[9], m[10], m[11]), normalize3(d), K_RAY_EPSILON, K_INFINITE };
    return r;
}

void spo_generate_rays(const spcu_flat_scene* s, const float* jitter, const uint32_t* pix, const uint32_t* smp,
                       uint64_t n, spcu_ray* rays)
{
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t x = pix[i] % s->width, y = pix[i] / s->width;
        /* main.cpp:97: p.x + sample.x with p.x an int */
        const ray_t r = camera_ray(s, (float)(int)x + jitter[2 * smp[i]], (float)(int)y + jitter[2 * smp[i] + 1]);
        const spcu_ray q = { r.o.x, r.o.y, r.o.z, r.t_min, r.d.x, r.d.y, r.d.z, r.t_max };
        rays[i]          = q;
    }
}

void spo_hit_records(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, float* normal_point, int32_t* material)
{
    for (uint64_t i = 0; i < n; ++i) {
        const ray_t     r = load_ray(&rays[i]);
        const closest_t c = scene_intersect(s, &r, NULL);
        float*          o = normal_point + 6 * i;
        memset(o, 0, 6 * sizeof(float));
        if (material) material[i] = -1;
        if (c.id >= 0) {
            const isect_t is = make_isect(s, &r, &c);
            o[0] = is.normal.x; o[1] = is.normal.y; o[2] = is.normal.z;
            o[3] = is.point.x;  o[4] = is.point.y;  o[5] = is.point.z;
            if (material) material[i] = (int32_t)is.material;
        }
    }
}

/* ===================================================================================================
 * BVH construction (shapes/BVHAccelerator.h:175-209)
 * =================================================================================================== */

/* _mm_min_ps(a, b) / _mm_max_ps(a, b) (math/Vector3.h:383-393): the SECOND operand on a tie (matters for +-0). */
static inline float sse_min(float a, float b) { return a < b ? a : b; }
static inline float sse_max(float a, float b) { return a > b ? a : b; }

typedef struct {
    const spcu_bounds* bounds;
    const uint8_t*     non_triangle;
    uint32_t*          order;
    spcu_bvh_node*     nodes;
    uint32_t           capacity, n_nodes, first_id, max_depth;
    int                overflow;
} build_t;

/* bounds = merge(bounds, prim) over [first, last) in range order (BVHAccelerator.h:181-185).  merge returns
 * BBox{min(lo, lo'), max(hi, hi')} (math/BBox.h:60-64), and that constructor sorts its corners once more (:26-30). */
static void build_fold(const build_t* b, uint32_t first, uint32_t last, float lo[3], float hi[3])
{
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; } /* BBox() (math/BBox.h:15-19) */
    for (uint32_t p = first; p < last; ++p) {
        const spcu_bounds* x = &b->bounds[b->order[p]];
        for (int a = 0; a < 3; ++a) {
            const float l = sse_min(lo[a], x->lo[a]);
            const float h = sse_max(hi[a], x->hi[a]);
            lo[a]         = sse_min(l, h);
            hi[a]         = sse_max(l, h);
        }
    }
}

/* max_dim (math/Vector3.h:653-670) */
static int build_max_dim(const float v[3])
{
    const float x = fabsf(v[0]), y = fabsf(v[1]), z = fabsf(v[2]);
    if (x > y) return x > z ? 0 : 2;
    return y > z ? 1 : 2;
}

/* Returns the link of the subtree over [first, last); writes its bounds and its leaf count word. */
static int32_t build_construct(build_t* b, uint32_t first, uint32_t last, uint32_t depth, float lo[3], float hi[3],
                               uint32_t* count_word)
{
    build_fold(b, first, last, lo, hi);
    uint32_t split = first;
    if (last - first > 4u) { /* k_max_leaf_elements (BVHAccelerator.h:211) */
        const float size[3] = { hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] };
        const int   d       = build_max_dim(size);
        const float at      = (lo[d] + hi[d]) / 2.0f; /* center (math/BBox.h:114-118) */
        /* std::partition, libstdc++ bidirectional version (bits/stl_algo.h __partition): Hoare scheme */
        uint32_t f = first, l = last;
        for (;;) {
            for (;;) {
                if (f == l) goto done;
                const spcu_bounds* x = &b->bounds[b->order[f]];
                if ((x->lo[d] + x->hi[d]) / 2.0f < at) ++f; else break;
            }
            --l;
            for (;;) {
                if (f == l) goto done;
                const spcu_bounds* x = &b->bounds[b->order[l]];
                if (!((x->lo[d] + x->hi[d]) / 2.0f < at)) --l; else break;
            }
            { const uint32_t t = b->order[f]; b->order[f] = b->order[l]; b->order[l] = t; }
            ++f;
        }
    done:
        split = f;
    }
    if (split == first || split == last) { /* NodeLeaf (also for the small ranges, where split stayed == first) */
        uint32_t mixed = 0;
        if (b->non_triangle)
            for (uint32_t p = first; p < last; ++p) mixed |= b->non_triangle[b->order[p]];
        *count_word = (last - first) | (mixed ? SPCU_LEAF_MIXED_FLAG : 0u);
        return ~(int32_t)(b->first_id + first);
    }
    if (depth + 1 > b->max_depth) b->max_depth = depth + 1;
    const uint32_t idx = b->n_nodes++;
    spcu_bvh_node  n;
    memset(&n, 0, sizeof n);
    if (idx >= b->capacity) b->overflow = 1;
    const uint32_t range[3] = { first, split, last };
    for (int k = 0; k < 2; ++k) {
        float clo[3], chi[3];
        n.child[k] = build_construct(b, range[k], range[k + 1], depth + 1, clo, chi, &n.count[k]);
        memcpy(n.box + 6 * k, clo, sizeof clo);
        memcpy(n.box + 6 * k + 3, chi, sizeof chi);
    }
    if (idx < b->capacity) b->nodes[idx] = n;
    *count_word = 0;
    return (int32_t)idx;
}

int spo_build_bvh(const spcu_bounds* bounds, uint32_t n, const uint8_t* non_triangle, uint32_t first_id, uint32_t* order,
                  spcu_bvh_node* nodes, uint32_t capacity, spcu_accel* accel, float* root_bounds)
{
    build_t b = { bounds, non_triangle, order, nodes, capacity, 0, first_id, 0, 0 };
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    float      lo[3], hi[3];
    spcu_accel a;
    memset(&a, 0, sizeof a);
    a.root        = build_construct(&b, 0, n, 0, lo, hi, &a.root_count);
    a.n_unbounded = first_id;
    a.n_prims     = first_id + n;
    a.n_nodes     = b.n_nodes;
    a.max_depth   = b.max_depth;
    a.nodes       = nodes;
    *accel        = a;
    if (root_bounds) { memcpy(root_bounds, lo, sizeof lo); memcpy(root_bounds + 3, hi, sizeof hi); }
    return b.overflow ? -1 : 0;
}

/* Triangle::get_world_bounds_impl (shapes/Triangle.h:228-237): BBox::extend(p) = { min(p, m_min), max(p, m_max) }
 * (math/BBox.h:42-46) over the three vertices — here the running value is the SECOND operand. */
void spo_triangle_bounds(const spcu_prim_geom* tris, uint32_t n, spcu_bounds* out)
{
    for (uint32_t i = 0; i < n; ++i) {
        spcu_bounds r = { { INFINITY, INFINITY, INFINITY }, { -INFINITY, -INFINITY, -INFINITY } };
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) {
                const float p = tris[i].v[4 * k + a];
                r.lo[a]       = sse_min(p, r.lo[a]);
                r.hi[a]       = sse_max(p, r.hi[a]);
            }
        out[i] = r;
    }
}

/* ===================================================================================================
 * Output side (main.cpp:100-102, Image/Image.cpp:14-55, Image/Image.h:38-50)
 * =================================================================================================== */
static float pack_to_srgb(float u) /* rgb_to_srgb (Image/Image.h:38-45); std::pow(float, float) = powf */
{
    if (u <= 0.0031308f) return 12.92f * u;
    return 1.055f * powf(u, 1.0f / 2.4f) - 0.055f;
}

void spo_pack_image(const float* rgb_sum, uint32_t width, uint32_t height, uint32_t spp, uint32_t format, void* out)
{
    const float n = (float)spp;
    size_t      o = 0;
    for (int j = (int)height - 1; j >= 0; --j) {
        for (uint32_t i = 0; i < width; ++i, ++o) {
            const float* c = rgb_sum + ((size_t)j * width + i) * 3;
            for (int k = 0; k < 3; ++k) {
                const float m = c[k] / n; /* RGB::operator/=(float) (math/RGB.h:126-132) */
                if (format == SPCU_IMAGE_PFM) {
                    ((float*)out)[3 * o + k] = m;
                } else {
                    int v = (int)(255.99f * pack_to_srgb(m)); /* write_ppm (Image/Image.cpp:22-24) */
                    v     = v < 0 ? 0 : (v > 65535 ? 65535 : v);
                    ((uint16_t*)out)[3 * o + k] = (uint16_t)v;
                }
            }
        }
    }
}

/* ===================================================================================================
 * Mesh ingest (base/PlyReader.cpp:487-531, shapes/Triangle.h:25-51)
 * =================================================================================================== */
/* difference_of_products (math/Vector3.h:489-498): cd = c*d; err = nmadd(c, d, cd); dop = msub(a, b, cd); dop + err */
static inline float mesh_dop(float a, float b, float c, float d)
{
    const float cd  = c * d;
    const float err = fmaf(-c, d, cd);
    const float dop = fmaf(a, b, -cd);
    return dop + err;
}
/* cross (math/Vector3.h:769-775) */
static inline v3 mesh_cross(v3 a, v3 b)
{
    return V(mesh_dop(a.y, b.z, a.z, b.y), mesh_dop(a.z, b.x, a.x, b.z), mesh_dop(a.x, b.y, a.y, b.x));
}

/* is_zero (math/Vector3.h:644-647) = float_compare(c, 0) per component (math/Math.h:265-272): |c| <= 1e-5 */
static int mesh_is_zero(v3 a) { return fabsf(a.x) <= 1.0e-05f && fabsf(a.y) <= 1.0e-05f && fabsf(a.z) <= 1.0e-05f; }

/* file_normals == NULL: read_ply; else read_binary_stl (base/STLReader.cpp:107-118: stored normal unless is_zero, then the
 * cross product; still is_zero: no contribution to the vertex normals, but the triangle stays — its indices were pushed at :96) */
static void mesh_ingest(const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float* file_normals,
                        const float object_to_world[12], const float normal_xf[9], uint32_t material, spcu_prim_geom* prims,
                        spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept, float* world_vertices, float* world_normals)
{
    v3*       vn   = (v3*)calloc(nv ? nv : 1, sizeof(v3)); /* vertex_normals(num_vertices, Normal3{0,0,0}) (PlyReader.cpp:510) */
    v3*       wv   = (v3*)malloc((nv ? nv : 1) * sizeof(v3));
    uint32_t* kept = (uint32_t*)malloc((nf ? nf : 1) * sizeof(uint32_t));
    uint32_t  nk   = 0;
#define VERT(i) V(vertices[3 * (size_t)(i)], vertices[3 * (size_t)(i) + 1], vertices[3 * (size_t)(i) + 2])
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t* ix = faces + 3 * (size_t)f;
        const v3        e0 = sub3(VERT(ix[1]), VERT(ix[0])), e1 = sub3(VERT(ix[2]), VERT(ix[0]));
        v3              n;
        if (file_normals) {
            n = V(file_normals[3 * (size_t)f], file_normals[3 * (size_t)f + 1], file_normals[3 * (size_t)f + 2]);
            if (mesh_is_zero(n)) n = mesh_cross(e0, e1);
            kept[nk++] = f; /* the indices are already in the mesh (STLReader.cpp:95-96) */
            if (mesh_is_zero(n)) continue;
        } else {
            n = mesh_cross(e0, e1);
            if (dot3(n, n) == 0.0f) continue; /* zero-area face: skipped (PlyReader.cpp:497-500) */
            kept[nk++] = f;
        }
        n = normalize3(n);
        for (int k = 0; k < 3; ++k) vn[ix[k]] = add3(vn[ix[k]], n); /* in face order (PlyReader.cpp:511-515) */
    }
    for (uint32_t v = 0; v < nv; ++v) {
        v3 n = vn[v];
        n    = (n.x != 0.0f || n.y != 0.0f || n.z != 0.0f) ? normalize3(n) : V(0.0f, 1.0f, 0.0f); /* PlyReader.cpp:517-528 */
        wv[v] = xf_point(object_to_world, VERT(v));                                                /* Triangle.h:37-41 */
        vn[v] = xf_vector(normal_xf, n);                                                           /* Triangle.h:43-47 */
        if (world_vertices) { world_vertices[3 * (size_t)v] = wv[v].x; world_vertices[3 * (size_t)v + 1] = wv[v].y; world_vertices[3 * (size_t)v + 2] = wv[v].z; }
        if (world_normals)  { world_normals[3 * (size_t)v] = vn[v].x;  world_normals[3 * (size_t)v + 1] = vn[v].y;  world_normals[3 * (size_t)v + 2] = vn[v].z; }
    }
#undef VERT
    for (uint32_t t = 0; t < nk; ++t) {
        const uint32_t* ix = faces + 3 * (size_t)kept[t];
        memset(&prims[t], 0, sizeof prims[t]);
        memset(&shade[t], 0, sizeof shade[t]);
        for (int k = 0; k < 3; ++k) {
            prims[t].v[4 * k] = wv[ix[k]].x, prims[t].v[4 * k + 1] = wv[ix[k]].y, prims[t].v[4 * k + 2] = wv[ix[k]].z;
            shade[t].v[4 * k] = vn[ix[k]].x, shade[t].v[4 * k + 1] = vn[ix[k]].y, shade[t].v[4 * k + 2] = vn[ix[k]].z;
        }
        meta[t] = SPCU_MAKE_META(SPCU_PRIM_TRIANGLE, material);
    }
    *n_kept = nk;
    free(vn);
    free(wv);
    free(kept);
}

void spo_ingest_mesh(const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float object_to_world[12],
                     const float normal_xf[9], uint32_t material, spcu_prim_geom* prims, spcu_prim_shade* shade, uint32_t* meta,
                     uint32_t* n_kept, float* world_vertices, float* world_normals)
{
    mesh_ingest(vertices, nv, faces, nf, NULL, object_to_world, normal_xf, material, prims, shade, meta, n_kept, world_vertices,
                world_normals);
}

void spo_ingest_mesh_stl(const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float* face_normals,
                         const float object_to_world[12], const float normal_xf[9], uint32_t material, spcu_prim_geom* prims,
                         spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept, float* world_vertices, float* world_normals)
{
    mesh_ingest(vertices, nv, faces, nf, face_normals, object_to_world, normal_xf, material, prims, shade, meta, n_kept,
                world_vertices, world_normals);
}

#include "sp_oracle_shade.inc"
