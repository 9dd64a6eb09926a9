// Device-side shading model of the backend: the reference's materials (materials/Material.h, Material.cpp), lights
// (Lights/Light.h, math/Distribution1D.h, math/Distribution2D.h) and sampling maps (math/Sampling.h), written for one
// thread = one path.  Every quirk of the reference's estimator that changes the expected image is kept on purpose
// (SURVEY.md §0.8); each function cites the lines it follows.  Arithmetic here is ordinary fp32 (FMA contraction
// allowed, CUDA libm transcendentals): parity of this stage with the reference is statistical, parity with the
// oracle — which consumes the same counter-based random numbers (rng.cuh) — is per pixel up to rounding.
#pragma once

#include "device_scene.h"
#include "features.h"
#include "rng.cuh"

#include <cfloat>

namespace spcu {

constexpr float kPi      = 3.14159265358979323846f;
constexpr float kInvPi   = 0.318309886183790671538f;
constexpr float kInv2Pi  = 1.0f / (2.0f * kPi);
constexpr float kEps     = 0.001f;   // k_ray_epsilon (math/Ray.h:11)
constexpr float kFltMax  = FLT_MAX;  // k_infinite_distance (base/Constants.h:16)
constexpr unsigned kRhoEvals = 16u;  // OneSampleMaterial::get_selection_weights (materials/Material.h:548)

struct V3
{
    float x, y, z;
};

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{ x, y, z }; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalize(V3 a) { return a * rsqrtf(dot(a, a)); }
__device__ __forceinline__ bool is_black(V3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }
__device__ __forceinline__ float luminance(V3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; } // math/RGB.h:224
__device__ __forceinline__ float sqr(float x) { return x * x; }
// std::max / std::clamp keep the first argument on NaN comparisons; fmaxf would not.
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float clamp_std(float x, float lo, float hi) { return x < lo ? lo : (hi < x ? hi : x); }

// get_ray_offset (math/Ray.h:51-85)
__device__ __forceinline__ float ray_offset_cos(float c) { return c == 0.0f ? kEps : kEps / c; }
__device__ __forceinline__ float ray_offset(V3 n, V3 d) { return ray_offset_cos(fabsf(dot(n, d))); }

// column-major 3x3 (+ affine) transforms as the flattener stores them
__device__ __forceinline__ V3 xf_vector(const float* m, V3 v)
{
    return v3(fmaf(v.x, m[0], fmaf(v.y, m[3], v.z * m[6])), fmaf(v.x, m[1], fmaf(v.y, m[4], v.z * m[7])),
              fmaf(v.x, m[2], fmaf(v.y, m[5], v.z * m[8])));
}
__device__ __forceinline__ V3 xf_point(const float* m, V3 p)
{
    return v3(fmaf(p.x, m[0], fmaf(p.y, m[3], fmaf(p.z, m[6], m[9]))), fmaf(p.x, m[1], fmaf(p.y, m[4], fmaf(p.z, m[7], m[10]))),
              fmaf(p.x, m[2], fmaf(p.y, m[5], fmaf(p.z, m[8], m[11]))));
}

// ---- ONB::from_v (math/ONB.h:12-32,57-62): u = b2, v = normalize(n), w = b1 ------------------------------------------
struct Onb
{
    V3 u, v, w;
};

__device__ __forceinline__ Onb onb_from_v(V3 n)
{
    Onb         o;
    const V3    v    = normalize(n);
    const float sign = copysignf(1.0f, v.z);
    const float a    = -1.0f / (sign + v.z);
    const float b    = v.x * v.y * a;
    o.w = v3(1.0f + sign * v.x * v.x * a, sign * b, -sign * v.x);
    o.u = v3(b, sign + v.y * v.y * a, -v.y);
    o.v = v;
    return o;
}
__device__ __forceinline__ V3 to_onb(const Onb& o, V3 a) { return v3(dot(a, o.u), dot(a, o.v), dot(a, o.w)); }
__device__ __forceinline__ V3 to_world(const Onb& o, V3 a) { return a.x * o.u + a.y * o.v + a.z * o.w; }

// ---- local-frame trigonometry (materials/Material.h:56-111), y up ------------------------------------------------------
__device__ __forceinline__ float cos_theta(V3 w) { return w.y; }
__device__ __forceinline__ float cos2_theta(V3 w) { return w.y * w.y; }
__device__ __forceinline__ float abs_cos_theta(V3 w) { return fabsf(w.y); }
__device__ __forceinline__ float sin2_theta(V3 w) { return max_std(0.0f, 1.0f - cos2_theta(w)); }
__device__ __forceinline__ float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
__device__ __forceinline__ bool  same_hemisphere(V3 a, V3 b) { return a.y * b.y > 0.0f; }

// cos(phi), sin(phi) with the reference's defaults when sin(theta) == 0 (both 1)
__device__ __forceinline__ void cos_sin_phi(V3 w, float& c, float& s)
{
    const float st = sin_theta(w);
    c = (st == 0.0f) ? 1.0f : clamp_std(w.x / st, -1.0f, 1.0f);
    s = (st == 0.0f) ? 1.0f : clamp_std(w.z / st, -1.0f, 1.0f);
}

// fresnel_dielectric (materials/Material.h:114-143)
__device__ __forceinline__ float fresnel_dielectric(float cos_i, float eta_i, float eta_t)
{
    cos_i = clamp_std(cos_i, -1.0f, 1.0f);
    if (!(cos_i > 0.0f)) {
        const float tmp = eta_i;
        eta_i           = eta_t;
        eta_t           = tmp;
        cos_i           = fabsf(cos_i);
    }
    const float sin_i = sqrtf(max_std(0.0f, 1.0f - cos_i * cos_i));
    const float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1.0f) {
        return 1.0f;
    }
    const float cos_t = sqrtf(max_std(0.0f, 1.0f - sin_t * sin_t));
    const float parl  = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    const float perp  = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (parl * parl + perp * perp) / 2.0f;
}

// erfinv (math/Math.h:230-261)
__device__ __forceinline__ float erfinv_sp(float a)
{
    float       p;
    const float t = logf(fmaf(a, 0.0f - a, 1.0f));
    if (fabsf(t) > 6.125f) {
        p = 3.03697567e-10f;
        p = fmaf(p, t, 2.93243101e-8f);
        p = fmaf(p, t, 1.22150334e-6f);
        p = fmaf(p, t, 2.84108955e-5f);
        p = fmaf(p, t, 3.93552968e-4f);
        p = fmaf(p, t, 3.02698812e-3f);
        p = fmaf(p, t, 4.83185798e-3f);
        p = fmaf(p, t, -2.64646143e-1f);
        p = fmaf(p, t, 8.40016484e-1f);
    } else {
        p = 5.43877832e-9f;
        p = fmaf(p, t, 1.43285448e-7f);
        p = fmaf(p, t, 1.22774793e-6f);
        p = fmaf(p, t, 1.12963626e-7f);
        p = fmaf(p, t, -5.61530760e-5f);
        p = fmaf(p, t, -1.47697632e-4f);
        p = fmaf(p, t, 2.31468678e-3f);
        p = fmaf(p, t, 1.15392581e-2f);
        p = fmaf(p, t, -2.32015476e-1f);
        p = fmaf(p, t, 8.86226892e-1f);
    }
    return a * p;
}

// ---- Beckmann distribution (materials/Material.h:213-267, materials/Material.cpp:14-157) --------------------------------
// beckmann_sample11 (Material.cpp:14-92): Newton iteration with bisection safeguard, at most 9 rounds.  What depends on the
// incident direction alone — a third of the function, with erf, acos and exp in it — is split off (Sample11Setup): the
// 16-sample albedo estimates of OneSampleMaterial call this 16 times with the SAME direction (bxdf_rho_microfacet).
struct Sample11Setup
{
    bool  normal_incidence; // cos_i > .9999: the closed-form branch
    float tan_i, c0, fit, normalization;
};

__device__ __forceinline__ Sample11Setup beckmann_sample11_setup(float cos_i)
{
    Sample11Setup s{};
    s.normal_incidence = cos_i > .9999f;
    if (s.normal_incidence) {
        return s;
    }
    const float sin_i = sqrtf(max_std(0.0f, 1.0f - sqr(cos_i)));
    s.tan_i           = sin_i / cos_i;
    const float cot_i = 1.0f / s.tan_i;
    s.c0              = erff(cot_i);
    const float theta_i     = acosf(cos_i);
    s.fit                   = 1.0f + theta_i * (-0.876f + theta_i * (0.4265f - 0.0594f * theta_i));
    const float sqrt_pi_inv = 1.0f / sqrtf(kPi);
    s.normalization         = 1.0f / (1.0f + s.c0 + sqrt_pi_inv * s.tan_i * expf(-cot_i * cot_i));
    return s;
}

__device__ __forceinline__ void beckmann_sample11_draw(const Sample11Setup& st, float U1, float U2, float& slope_x, float& slope_y)
{
    if (st.normal_incidence) {
        const float r = sqrtf(-logf(1.0f - U1));
        float       s, c;
        sincosf(2.0f * kPi * U2, &s, &c);
        slope_x = r * c;
        slope_y = r * s;
        return;
    }
    const float tan_i = st.tan_i;

    float       a        = -1.0f;
    float       c        = st.c0;
    const float sample_x = max_std(U1, 1e-6f);

    float b = c - (1.0f + c) * powf(1.0f - sample_x, st.fit);

    const float sqrt_pi_inv   = 1.0f / sqrtf(kPi);
    const float normalization = st.normalization;

    for (int it = 0; it < 9; ++it) {
        if (!(b >= a && b <= c)) {
            b = 0.5f * (a + c);
        }
        const float inv_erf    = erfinv_sp(b);
        const float value      = normalization * (1.0f + b + sqrt_pi_inv * tan_i * expf(-inv_erf * inv_erf)) - sample_x;
        const float derivative = normalization * (1.0f - inv_erf * tan_i);
        if (fabsf(value) < 1e-5f) {
            break;
        }
        if (value > 0.0f) {
            c = b;
        } else {
            a = b;
        }
        b -= value / derivative;
    }
    slope_x = erfinv_sp(b);
    slope_y = erfinv_sp(2.0f * max_std(U2, 1e-6f) - 1.0f);
}

__device__ __forceinline__ void beckmann_sample11(float cos_i, float U1, float U2, float& slope_x, float& slope_y)
{
    beckmann_sample11_draw(beckmann_sample11_setup(cos_i), U1, U2, slope_x, slope_y);
}

// beckmann_sample (Material.cpp:94-114)
__device__ __forceinline__ V3 beckmann_sample(V3 wi, float ax, float ay, float U1, float U2)
{
    const V3 ws = normalize(v3(ax * wi.x, wi.y, ay * wi.z));
    float    sx, sy, cp, sp;
    beckmann_sample11(cos_theta(ws), U1, U2, sx, sy);
    cos_sin_phi(ws, cp, sp);
    const float tmp = cp * sx - sp * sy;
    sy              = sp * sx + cp * sy;
    sx              = tmp;
    sx              = ax * sx;
    sy              = ay * sy;
    return normalize(v3(-sx, 1.0f, -sy));
}

// BeckmannDistribution::sample_wh_impl (Material.cpp:117-157).  In the visible-area branch the two get_next_1D()
// arguments are ONE draw call (rng.cuh).  The other branch keeps the reference's 2*phi (golden ratio) factor.
static __device__ __noinline__ V3 beckmann_sample_wh(const spcu_bxdf& bx, V3 wo, Rng& rng)
{
    if (!bx.sample_visible) {
        float tan2, phi;
        if (bx.alpha_x == bx.alpha_y) {
            const float log_sample = logf(1.0f - rng_next1(rng));
            tan2                   = -bx.alpha_x * bx.alpha_x * log_sample;
            phi                    = rng_next1(rng) * 2.0f * 1.6180339887498948482f;
        } else {
            const float log_sample = logf(1.0f - rng_next1(rng));
            const float u1         = rng_next1(rng);
            phi                    = atanf(bx.alpha_y / bx.alpha_x * tanf(2.0f * kPi * u1 + 0.5f * kPi));
            if (u1 > 0.5f) {
                phi += kPi;
            }
            float sp, cp;
            sincosf(phi, &sp, &cp);
            tan2 = -log_sample / (sqr(cp) / sqr(bx.alpha_x) + sqr(sp) / sqr(bx.alpha_y));
        }
        const float ct = 1.0f / sqrtf(1.0f + tan2);
        const float st = sqrtf(max_std(0.0f, 1.0f - sqr(ct)));
        float       sp, cp;
        sincosf(phi, &sp, &cp);
        V3 wh = v3(st * cp, ct, st * sp);
        if (!same_hemisphere(wo, wh)) {
            wh = -wh;
        }
        return wh;
    }
    const bool flip = wo.y < 0.0f;
    float      U1, U2;
    rng_next2(rng, U1, U2);
    V3 wh = beckmann_sample(flip ? -wo : wo, bx.alpha_x, bx.alpha_y, U1, U2);
    if (flip) {
        wh = -wh;
    }
    return wh;
}

// BeckmannDistribution::D_impl (Material.h:237-247)
__device__ __forceinline__ float beckmann_D(V3 wh, float ax, float ay)
{
    const float t2 = sin2_theta(wh) / cos2_theta(wh);
    if (isinf(t2)) {
        return 0.0f;
    }
    const float cos4 = sqr(cos2_theta(wh));
    if (ax == ay) { // isotropic (every material the .sp parser creates): cos^2 phi + sin^2 phi = 1 up to rounding
        return expf(-t2 / sqr(ax)) / (kPi * ax * ay * cos4);
    }
    float       cp, sp;
    cos_sin_phi(wh, cp, sp);
    return expf(-t2 * (sqr(cp) / sqr(ax) + sqr(sp) / sqr(ay))) / (kPi * ax * ay * cos4);
}

// BeckmannDistribution::lambda (Material.h:249-261)
__device__ __forceinline__ float beckmann_lambda(V3 w, float ax, float ay)
{
    const float abs_tan = fabsf(sin_theta(w) / cos_theta(w));
    if (isinf(abs_tan)) {
        return 0.0f;
    }
    float alpha = ax;
    if (ax != ay) { // (isotropic: sqrt(cos^2 phi + sin^2 phi) * ax = ax up to rounding; the azimuth is not needed)
        float cp, sp;
        cos_sin_phi(w, cp, sp);
        alpha = sqrtf(sqr(cp) * sqr(ax) + sqr(sp) * sqr(ay));
    }
    const float a     = 1.0f / (alpha * abs_tan);
    if (a >= 1.6f) {
        return 0.0f;
    }
    return (1.0f - 1.259f * a + 0.396f * sqr(a)) / (3.535f * a + 2.181f * sqr(a));
}

// MicrofacetDistribution::pdf (Material.h:185-192)
__device__ __forceinline__ float distribution_pdf(const spcu_bxdf& bx, V3 wo, V3 wh)
{
    if (bx.sample_visible) {
        const float g1 = 1.0f / (1.0f + beckmann_lambda(wo, bx.alpha_x, bx.alpha_y));
        return beckmann_D(wh, bx.alpha_x, bx.alpha_y) * g1 * fabsf(dot(wo, wh)) / abs_cos_theta(wo);
    }
    return beckmann_D(wh, bx.alpha_x, bx.alpha_y) * abs_cos_theta(wh);
}

// ---- BxDFs (materials/Material.h:269-454) -------------------------------------------------------------------------------
struct MSample
{
    V3    color;
    V3    dir;
    float pdf;
    bool  specular; // is_specular(properties) (materials/BSDFProperties.h:48-50); only the Whitted integrator reads it
};

__device__ __forceinline__ MSample degenerate_sample() { return MSample{ v3(0, 0, 0), v3(0, 0, 0), 0.0f, false }; }
__device__ __forceinline__ V3 bxdf_r(const spcu_bxdf& bx) { return v3(bx.r[0], bx.r[1], bx.r[2]); }

// BRDF::eval: LambertianBRDF :334-337, SpecularReflectionBRDF :371-374, MicrofacetReflection :424-440
template <typename F>
__device__ __forceinline__ V3 bxdf_eval(const spcu_bxdf& bx, V3 wo, V3 wi)
{
    if (!(F::microfacet || F::specular_bxdf) || bx.kind == SPCU_BXDF_LAMBERT) {
        return bxdf_r(bx);
    }
    if (!F::microfacet || bx.kind == SPCU_BXDF_SPECULAR) {
        return v3(0, 0, 0);
    }
    const float co = abs_cos_theta(wo), ci = abs_cos_theta(wi);
    if (ci == 0.0f || co == 0.0f) {
        return v3(0, 0, 0);
    }
    V3 wh = wi + wo;
    if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) {
        return v3(0, 0, 0);
    }
    wh            = normalize(wh);
    const float f = fresnel_dielectric(dot(wi, wh), 1.0f, bx.ior);
    const float D = beckmann_D(wh, bx.alpha_x, bx.alpha_y);
    const float G = 1.0f / (1.0f + beckmann_lambda(wo, bx.alpha_x, bx.alpha_y) + beckmann_lambda(wi, bx.alpha_x, bx.alpha_y));
    return bxdf_r(bx) * D * G * f / (4.0f * ci * co);
}

// BRDF::pdf: Lambert :339-342 (uniform hemisphere), specular :376-381, microfacet :442-449
template <typename F>
__device__ __forceinline__ float bxdf_pdf(const spcu_bxdf& bx, V3 wo, V3 wi)
{
    if (!(F::microfacet || F::specular_bxdf) || bx.kind == SPCU_BXDF_LAMBERT) {
        return kInv2Pi;
    }
    if (!F::microfacet || bx.kind == SPCU_BXDF_SPECULAR) {
        return 0.0f;
    }
    if (!same_hemisphere(wo, wi)) {
        return 0.0f;
    }
    const V3 wh = normalize(wo + wi);
    return distribution_pdf(bx, wo, wh) / (4.0f * dot(wo, wh));
}

// BRDF::sample: Lambert :322-332 (uniform hemisphere, not cosine weighted), specular :362-369, microfacet :398-422
template <typename F>
__device__ __forceinline__ MSample bxdf_sample(const spcu_bxdf& bx, V3 wo, Rng& rng)
{
    MSample s;
    if (!(F::microfacet || F::specular_bxdf) || bx.kind == SPCU_BXDF_LAMBERT) {
        float u0, u1;
        rng_next2(rng, u0, u1);
        const float y = u0;
        const float r = sqrtf(max_std(0.0f, 1.0f - y * y));
        float       sp, cp;
        sincosf(2.0f * kPi * u1, &sp, &cp);
        s.dir      = v3(r * cp, y, r * sp);
        s.color    = bxdf_r(bx);
        s.pdf      = kInv2Pi;
        s.specular = false;
        return s;
    }
    if (!F::microfacet || bx.kind == SPCU_BXDF_SPECULAR) {
        s.dir   = v3(-wo.x, wo.y, -wo.z);
        s.color    = fresnel_dielectric(cos_theta(s.dir), 1.0f, 1.5f) * bxdf_r(bx) / abs_cos_theta(s.dir);
        s.pdf      = 1.0f;
        s.specular = true;
        return s;
    }
    if (wo.y == 0.0f) {
        return degenerate_sample();
    }
    const V3    wh = beckmann_sample_wh(bx, wo, rng);
    const float dp = dot(wo, wh);
    if (dp < 0.0f) {
        return degenerate_sample();
    }
    const V3 wi = -wo + 2.0f * dot(wo, wh) * wh; // specular_reflection (Material.h:45-48)
    if (!same_hemisphere(wo, wi)) {
        return degenerate_sample();
    }
    s.pdf      = distribution_pdf(bx, wo, wh) / (4.0f * dp);
    s.color    = bxdf_eval<F>(bx, wo, wi);
    s.dir      = wi;
    s.specular = false;
    return s;
}

// BRDF::rho_impl (:299-310) for a MicrofacetReflection over a Beckmann distribution that samples visible normals — the
// 16-sample albedo estimate behind every OneSampleMaterial::sample / eval / pdf of a glossy material, i.e. where the NEE stage
// spends its time (ncu, profiles/r02b_*: 58 K thread instructions per vertex and light, nearly all of it in these loops).
// The same samples from the same random numbers, with two algebraic facts used:
//  * everything that depends on wo alone is computed once, not 16 times: the stretched direction and its azimuth
//    (beckmann_sample, Material.cpp:94-114), the incidence-dependent third of beckmann_sample11, Lambda(wo);
//  * a sample's term is  f * |cos wi| / pdf  with  f = r D G F / (4 ci co)  (MicrofacetReflection::eval, :424-440) and
//    pdf = D G1(wo) |wo.wh| / co / (4 wo.wh)  (:442-449 over MicrofacetDistribution::pdf :185-192): D, ci and co cancel and
//    the term is  r F (1 + Lambda(wo)) / (1 + Lambda(wo) + Lambda(wi)),  with F at wo.wh (= wi.wh for a mirror direction).
// The rejections of bxdf_sample (wo.wh < 0, wi below the horizon) are kept; what the cancellation drops is the case D == 0
// (the exponential underflowing for a SAMPLED normal: probability below e^-87) and rounding in the last digits — parity of this
// stage with the oracle is per pixel up to rounding (tests/test_gpu_render.py), parity with the reference statistical.
static __device__ __noinline__ V3 bxdf_rho_microfacet(const spcu_bxdf& bx, V3 wo, Rng& rng)
{
    if (wo.y == 0.0f) {
        return v3(0, 0, 0); // bxdf_sample returns before it draws (:400): no random numbers consumed
    }
    const bool  flip = wo.y < 0.0f;
    const V3    wv   = flip ? -wo : wo;
    const float ax = bx.alpha_x, ay = bx.alpha_y;
    const V3    ws = normalize(v3(ax * wv.x, wv.y, ay * wv.z));
    float       cp, sp;
    cos_sin_phi(ws, cp, sp);
    const Sample11Setup st        = beckmann_sample11_setup(cos_theta(ws));
    const float         one_lo    = 1.0f + beckmann_lambda(wo, ax, ay);
    float               sum       = 0.0f;
#pragma unroll 1
    for (unsigned i = 0; i < kRhoEvals; ++i) {
        float U1, U2, sx, sy;
        rng_next2(rng, U1, U2);
        beckmann_sample11_draw(st, U1, U2, sx, sy);
        const float tmp = cp * sx - sp * sy;
        sy              = sp * sx + cp * sy;
        sx              = ax * tmp;
        sy              = ay * sy;
        V3 wh = normalize(v3(-sx, 1.0f, -sy));
        if (flip) {
            wh = -wh;
        }
        const float dp = dot(wo, wh);
        if (!(dp > 0.0f) || wh.y == 0.0f) { // dp < 0: rejected; dp == 0: pdf = 0 / 0, rejected by `pdf > 0`; wh.y == 0: D = 0
            continue;
        }
        const V3 wi = -wo + 2.0f * dp * wh;
        if (!same_hemisphere(wo, wi)) {
            continue;
        }
        const float f = fresnel_dielectric(dp, 1.0f, bx.ior);
        sum += f * one_lo / (one_lo + beckmann_lambda(wi, ax, ay));
    }
    return bxdf_r(bx) * (sum / static_cast<float>(kRhoEvals));
}

// BRDF::rho: LambertianBRDF overrides it (:344-347, no random numbers); the others run BRDF::rho_impl (:299-310)
template <typename F>
__device__ __forceinline__ V3 bxdf_rho(const spcu_bxdf& bx, V3 wo, Rng& rng)
{
    if (!(F::microfacet || F::specular_bxdf) || bx.kind == SPCU_BXDF_LAMBERT) {
        return bxdf_r(bx) * kPi;
    }
#ifndef SPCU_RHO_LITERAL // (A/B switch: the literal loop below for every BxDF)
    if (F::microfacet && bx.kind == SPCU_BXDF_MICROFACET && bx.sample_visible) {
        return bxdf_rho_microfacet(bx, wo, rng);
    }
#endif
    V3 r = v3(0, 0, 0);
#pragma unroll 1
    for (unsigned i = 0; i < kRhoEvals; ++i) {
        const MSample s = bxdf_sample<F>(bx, wo, rng);
        if (s.pdf > 0.0f) {
            r = r + s.color * abs_cos_theta(s.dir) / s.pdf;
        }
    }
    return r / static_cast<float>(kRhoEvals);
}

// ---- materials (materials/Material.h:456-806) ----------------------------------------------------------------------------
// OneSampleMaterial::get_selection_weights :545-572 — a fresh 16-sample albedo estimate per BxDF on EVERY call
template <typename F>
static __device__ __noinline__ void selection_weights(const spcu_bxdf* bx, uint32_t n, V3 wo, Rng& rng, float* w)
{
    float sum = 0.0f;
#pragma unroll 1
    for (uint32_t i = 0; i < n; ++i) {
        w[i] = luminance(bxdf_rho<F>(bx[i], wo, rng));
        sum += w[i];
    }
    for (uint32_t i = 0; i < n; ++i) {
        w[i] = w[i] / sum;
    }
}

__device__ __forceinline__ float balance1(float p, float inner) { return inner == 0.0f ? 0.0f : p / inner; } // math/Math.h:82-89

// OneSampleMaterial::sample_impl :577-667
template <typename F>
__device__ __forceinline__ MSample one_sample_sample(const DScene& s, const spcu_material& m, V3 wo, Rng& rng)
{
    const spcu_bxdf* bx = s.bxdfs + m.first_bxdf;
    const uint32_t   n  = m.n_bxdfs;
    if (!F::multi_bxdf || n == 1) {
        return bxdf_sample<F>(bx[0], wo, rng);
    }
    float w[SPCU_MAX_BXDFS];
    selection_weights<F>(bx, n, wo, rng, w);

    const float u        = rng_next1(rng);
    float       running  = 0.0f;
    uint32_t    selected = n - 1u; // left uninitialised by the reference when no term exceeds u, then clamped (:604-610)
    for (uint32_t i = 0; i < n; ++i) {
        if (w[i] + running > u) {
            selected = i;
            break;
        }
        running += w[i];
    }
    const MSample r = bxdf_sample<F>(bx[selected], wo, rng);
    if (r.pdf == 0.0f || is_black(r.color)) {
        return degenerate_sample();
    }
    V3    values[SPCU_MAX_BXDFS];
    float pdfs[SPCU_MAX_BXDFS];
    float inner = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        if (i == selected) {
            values[i] = r.color;
            pdfs[i]   = r.pdf * w[i];
        } else {
            values[i] = bxdf_eval<F>(bx[i], wo, r.dir);
            pdfs[i]   = bxdf_pdf<F>(bx[i], wo, r.dir) * w[i];
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        inner += pdfs[i];
    }
    MSample out{ v3(0, 0, 0), r.dir, 0.0f, r.specular }; // result.properties of the selected BxDF (:666)
    for (uint32_t i = 0; i < n; ++i) {
        if (pdfs[i] > 0.0f) {
            out.color = out.color + balance1(pdfs[i], inner) * values[i];
            out.pdf += pdfs[i];
        }
    }
    return out;
}

// OneSampleMaterial::pdf_impl :669-683
template <typename F>
__device__ __forceinline__ float one_sample_pdf(const DScene& s, const spcu_material& m, V3 wo, V3 wi, Rng& rng)
{
    const spcu_bxdf* bx = s.bxdfs + m.first_bxdf;
    if (!F::multi_bxdf) { // one BxDF, weight x / x == 1 (features.h)
        return bxdf_pdf<F>(bx[0], wo, wi);
    }
    float w[SPCU_MAX_BXDFS];
    selection_weights<F>(bx, m.n_bxdfs, wo, rng, w);
    float pdf = 0.0f;
    for (uint32_t i = 0; i < m.n_bxdfs; ++i) {
        pdf += w[i] * bxdf_pdf<F>(bx[i], wo, wi);
    }
    return pdf;
}

// OneSampleMaterial::eval_impl :685-715
template <typename F>
__device__ __forceinline__ V3 one_sample_eval(const DScene& s, const spcu_material& m, V3 wo, V3 wi, Rng& rng)
{
    const spcu_bxdf* bx = s.bxdfs + m.first_bxdf;
    if (!F::multi_bxdf) { // one BxDF: pdf * 1 > 0 for every BxDF this feature set admits, balance p / p == 1
        return bxdf_eval<F>(bx[0], wo, wi);
    }
    float w[SPCU_MAX_BXDFS], pdfs[SPCU_MAX_BXDFS];
    selection_weights<F>(bx, m.n_bxdfs, wo, rng, w);
    float inner = 0.0f;
    for (uint32_t i = 0; i < m.n_bxdfs; ++i) {
        pdfs[i] = bxdf_pdf<F>(bx[i], wo, wi) * w[i];
    }
    for (uint32_t i = 0; i < m.n_bxdfs; ++i) {
        inner += pdfs[i];
    }
    V3 result = v3(0, 0, 0);
    for (uint32_t i = 0; i < m.n_bxdfs; ++i) {
        if (pdfs[i] > 0.0f) {
            result = result + balance1(pdfs[i], inner) * bxdf_eval<F>(bx[i], wo, wi);
        }
    }
    return result;
}

// ClearcoatMaterial (:723-806) wraps a base material; the reference recurses through m_base, here the chain of coats
// is walked iteratively (the flattener bounds its length by SPCU_MAX_COAT_DEPTH; feature sets without nested coats walk
// at most one, with no arrays and no loops) and unwound in the same order.  The walk does not depend on the incident
// direction, so eval and pdf of the same (material, wo) share one (material_*_coats).
template <typename F>
struct Coats
{
    static constexpr int kMax = F::nested_coats ? static_cast<int>(SPCU_MAX_COAT_DEPTH) : 1;
    float    f[kMax];   // Fresnel term of coat i, outermost first
    uint32_t mat[kMax]; // the coat's material (its specular colour scales the base's sample)
    int      depth;
    uint32_t base;      // the OneSampleMaterial under the coats
};

template <typename F>
__device__ __forceinline__ Coats<F> walk_coats(const DScene& s, uint32_t mat, V3 wo)
{
    Coats<F> c;
    c.depth = 0;
#pragma unroll
    for (int i = 0; i < Coats<F>::kMax; ++i) {
        if (s.materials[mat].kind != SPCU_MAT_CLEARCOAT) {
            break;
        }
        c.f[i]   = fresnel_dielectric(cos_theta(wo), 1.0f, s.materials[mat].ior);
        c.mat[i] = mat;
        c.depth  = i + 1;
        mat      = s.materials[mat].base;
    }
    c.base = mat;
    return c;
}

// ClearcoatMaterial::sample_impl :734-765 over OneSampleMaterial::sample_impl.  Coats are entered until one reflects
// specularly (probability f, one draw per coat) or the base is reached; the base is sampled at ONE place, after the walk,
// so lanes under different numbers of coats run it together.
template <typename F>
__device__ __forceinline__ MSample material_sample_local(const DScene& s, uint32_t mat, V3 wo, Rng& rng)
{
    float    cf[Coats<F>::kMax];
    uint32_t cm[Coats<F>::kMax];
    int      depth    = 0;
    bool     specular = false;
    float    f        = 0.0f;
#pragma unroll
    for (int i = 0; i < Coats<F>::kMax; ++i) {
        if (specular || s.materials[mat].kind != SPCU_MAT_CLEARCOAT) {
            break;
        }
        f = fresnel_dielectric(cos_theta(wo), 1.0f, s.materials[mat].ior);
        if (rng_next1(rng) < f) {
            specular = true;
        } else {
            cf[i] = f;
            cm[i] = mat;
            depth = i + 1;
            mat   = s.materials[mat].base;
        }
    }
    MSample r;
    if (specular) {
        const spcu_material& m = s.materials[mat];
        r.dir      = v3(-wo.x, wo.y, -wo.z);
        r.color    = f * v3(m.specular[0], m.specular[1], m.specular[2]) / abs_cos_theta(r.dir);
        r.pdf      = f;
        r.specular = true;
    } else {
        r = one_sample_sample<F>(s, s.materials[mat], wo, rng);
    }
#pragma unroll
    for (int i = Coats<F>::kMax - 1; i >= 0; --i) {
        if (i < depth && r.pdf != 0.0f) { // pdf == 0: `return base_result` at every enclosing level
            const spcu_material& m = s.materials[cm[i]];
            r.pdf                  = (1.0f - cf[i]) * r.pdf;
            r.color = (v3(1.0f, 1.0f, 1.0f) - cf[i] * v3(m.specular[0], m.specular[1], m.specular[2])) * r.color;
        }
    }
    return r;
}

// product of (1 - f) over the coats, applied innermost first as the recursion unwinds (:767-801)
template <typename F>
__device__ __forceinline__ float material_pdf_coats(const DScene& s, const Coats<F>& c, V3 wo, V3 wi, Rng& rng)
{
    float pdf = one_sample_pdf<F>(s, s.materials[c.base], wo, wi, rng);
#pragma unroll
    for (int i = Coats<F>::kMax - 1; i >= 0; --i) {
        if (i < c.depth) {
            pdf = (1.0f - c.f[i]) * pdf;
        }
    }
    return pdf;
}

template <typename F>
__device__ __forceinline__ V3 material_eval_coats(const DScene& s, const Coats<F>& c, V3 wo, V3 wi, Rng& rng)
{
    V3 f = one_sample_eval<F>(s, s.materials[c.base], wo, wi, rng);
#pragma unroll
    for (int i = Coats<F>::kMax - 1; i >= 0; --i) {
        if (i < c.depth) {
            f = (1.0f - c.f[i]) * f;
        }
    }
    return f;
}

template <typename F>
__device__ __forceinline__ float material_pdf_local(const DScene& s, uint32_t mat, V3 wo, V3 wi, Rng& rng)
{
    return material_pdf_coats<F>(s, walk_coats<F>(s, mat, wo), wo, wi, rng);
}

template <typename F>
__device__ __forceinline__ V3 material_eval_local(const DScene& s, uint32_t mat, V3 wo, V3 wi, Rng& rng)
{
    return material_eval_coats<F>(s, walk_coats<F>(s, mat, wo), wo, wi, rng);
}

// Material::sample (materials/Material.h:461-473): result direction returned in world space
template <typename F>
__device__ __forceinline__ MSample material_sample(const DScene& s, uint32_t mat, V3 wo, V3 n, Rng& rng)
{
    const Onb onb = onb_from_v(n);
    MSample   r   = material_sample_local<F>(s, mat, to_onb(onb, wo), rng);
    if (r.pdf == 0.0f || is_black(r.color)) {
        return r;
    }
    r.dir = to_world(onb, r.dir);
    return r;
}

// ---- lights (Lights/Light.h) -----------------------------------------------------------------------------------------------
// std::ranges::upper_bound exactly as libstdc++ bisects (bits/stl_algo.h): the reference's CDF tables are not sorted
// at their last entry (math/Distribution1D.h:42-43 shifts the normalised values left by one), so the probe sequence
// itself is part of the contract.
__device__ __forceinline__ uint32_t upper_bound_f(const float* a, uint32_t n, float val)
{
    uint32_t first = 0, len = n;
    while (len > 0) {
        const uint32_t half = len >> 1, middle = first + half;
        if (val < __ldg(a + middle)) {
            len = half;
        } else {
            first = middle + 1;
            len   = len - half - 1;
        }
    }
    return first;
}

// Distribution1D::sample_continuous (math/Distribution1D.h:77-98, get_offset :135-143), m_min = 0, m_max = 1
__device__ __forceinline__ float dist_sample_continuous(const float* func, const float* cdf, float integral, uint32_t n,
                                                        float u, float& pdf, uint32_t& offset)
{
    const uint32_t it = upper_bound_f(cdf, n + 1, u);
    offset            = (it == n + 1 || it == n) ? n - 1 : it;
    const float c0 = __ldg(cdf + offset), c1 = __ldg(cdf + offset + 1);
    float       du = u - c0;
    if ((c1 - c0) > 0.0f) {
        du /= (c1 - c0);
    }
    pdf           = (integral > 0.0f) ? __ldg(func + offset) / integral : 0.0f;
    const float x = (static_cast<float>(offset) + du) / static_cast<float>(n);
    return (1.0f - x) * 0.0f + x * 1.0f; // sp::lerp(x, m_min, m_max)
}

// sample_nearest_neighbor(img, s, t, RemapWrap, RemapClamp) (Image/Image.h:74-115)
__device__ __forceinline__ V3 ibl_lookup(const DScene& s, const spcu_light& l, float u, float v)
{
    u = fmodf(1.0f + fmodf(u, 1.0f), 1.0f);
    v = clamp_std(v, 0.0f, 0x1.fffffep-1f);
    const float fx = roundf(u * static_cast<float>(l.img_w));
    const float fy = roundf(v * static_cast<float>(l.img_h));
    uint32_t    x  = static_cast<uint32_t>(fx), y = static_cast<uint32_t>(fy);
    x              = min(x, l.img_w - 1u);
    y              = min(y, l.img_h - 1u);
    const float* p = s.pool + l.img_off + (static_cast<uint64_t>(y) * l.img_w + x) * 3u;
    return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

__device__ __forceinline__ float spherical_theta(V3 v) { return acosf(clamp_std(v.y, -1.0f, 1.0f)); } // math/Sampling.h:84
__device__ __forceinline__ float spherical_phi(V3 v)                                                   // math/Sampling.h:89
{
    const float p = atan2f(v.z, v.x);
    return (p < 0.0f) ? (p + 2.0f * kPi) : p;
}

// sample_to_uniform_sphere (math/Sampling.h:20-26); the reference's `2.0 *` promotes the product to double
__device__ __forceinline__ V3 sample_uniform_sphere(float u0, float u1)
{
    const float z   = 1.0f - 2.0f * u0;
    const float r   = sqrtf(max_std(0.0f, 1.0f - z * z));
    const float phi = static_cast<float>(2.0 * static_cast<double>(kPi) * static_cast<double>(u1));
    float       sp, cp;
    sincosf(phi, &sp, &cp);
    return v3(r * cp, r * sp, z);
}

// sample_to_concentric_disk (math/Sampling.cpp:10-34) + sample_to_cosine_hemisphere (math/Sampling.h:45-50)
__device__ __forceinline__ V3 sample_cosine_hemisphere(float u0, float u1)
{
    const float ox = 2.0f * u0 - 1.0f, oy = 2.0f * u1 - 1.0f;
    float       dx = 0.0f, dy = 0.0f;
    if (!(ox == 0.0f && oy == 0.0f)) {
        float theta, r;
        if (fabsf(ox) > fabsf(oy)) {
            r     = ox;
            theta = (kPi / 4.0f) * (oy / ox);
        } else {
            r     = oy;
            theta = (kPi / 2.0f) - (kPi / 4.0f) * (ox / oy);
        }
        float st, ct;
        sincosf(theta, &st, &ct);
        dx = r * ct;
        dy = r * st;
    }
    const float y = sqrtf(max_std(0.0f, 1.0f - dx * dx - dy * dy));
    return v3(dx, y, dy);
}

// Sphere::pdf (shapes/Sphere.h:53-74): the uniform-cone density, whatever the direction
__device__ __forceinline__ float sphere_pdf(const spcu_light& l, V3 p)
{
    const V3    o  = xf_point(l.world_to_object, p);
    const float d2 = dot(o, o);
    if (d2 <= 1.0f) {
        return 1.0f / (4.0f * kPi);
    }
    const float sin2_max = 1.0f / d2;
    const float cos_max  = sqrtf(max_std(0.0f, 1.0f - sin2_max));
    const float omc      = (sin2_max < 0.00068523f) ? sin2_max / 2.0f : 1.0f - cos_max;
    return 1.0f / (2.0f * kPi * omc);
}

struct LSample
{
    V3    L;
    float pdf;
    V3    wi;
    float t_min, t_max;
    float u, v; // image-based light: the sampled map coordinates (its L is ibl_lookup(u, v)); 0 otherwise
};

// Light::sample (Lights/Light.h:38-49) over SphereLight (ObjectLight::sample_impl :81-90 + Sphere::sample
// shapes/Sphere.h:20-51), EnvironmentLight (:155-161), ImageBasedEnvironmentLight (:226-249)
template <typename F>
__device__ __forceinline__ LSample light_sample(const DScene& s, const spcu_light& l, V3 p, V3 n, float u0, float u1)
{
    LSample out;
    out.t_max = kFltMax;
    out.u = out.v = 0.0f;
    if (l.kind == SPCU_LIGHT_SPHERE) {
        const V3 obs = xf_point(l.world_to_object, p);
        V3       local;
        if (dot(obs, obs) <= 1.0f) {
            local = sample_uniform_sphere(u0, u1);
        } else {
            const V3  c   = sample_cosine_hemisphere(u0, u1);
            const Onb onb = onb_from_v(obs);
            local         = to_world(onb, c);
        }
        const V3 sp_point  = xf_point(l.object_to_world, local);
        const V3 sp_normal = xf_vector(l.normal_xf, local);
        const V3 to_sample = sp_point - p;
        out.wi             = normalize(to_sample);
        out.pdf            = sphere_pdf(l, p);
        out.t_max          = sqrtf(dot(to_sample, to_sample)) - ray_offset(sp_normal, -out.wi);
        out.L              = v3(l.radiance[0], l.radiance[1], l.radiance[2]);
    } else if (!F::ibl || l.kind == SPCU_LIGHT_ENV_CONST) {
        out.wi  = sample_uniform_sphere(u0, u1);
        out.pdf = 1.0f / (4.0f * kPi);
        out.L   = v3(l.radiance[0], l.radiance[1], l.radiance[2]);
    } else {
        const float* pool = s.pool;
        float        pdf0, pdf1;
        uint32_t     v, dummy;
        const float  d1 = dist_sample_continuous(pool + l.marg_func_off, pool + l.marg_cdf_off, l.marg_integral, l.nv, u1, pdf1, v);
        const float  d0 = dist_sample_continuous(pool + l.cond_func_off + static_cast<uint64_t>(v) * l.nu,
                                                 pool + l.cond_cdf_off + static_cast<uint64_t>(v) * (l.nu + 1u),
                                                 __ldg(pool + l.cond_int_off + v), l.nu, u0, pdf0, dummy);
        const float  map_pdf = pdf0 * pdf1;
        if (map_pdf == 0.0f) {
            out.pdf = 0.0f;
            out.L   = v3(0, 0, 0);
            out.wi  = v3(0.0f, 1.0f, 0.0f);
        } else {
            float st, ct, sp, cp;
            sincosf(d1 * kPi, &st, &ct);
            sincosf(d0 * 2.0f * kPi, &sp, &cp);
            out.wi  = xf_vector(l.light_to_world, v3(st * cp, ct, st * sp));
            out.pdf = (st == 0.0f) ? 0.0f : map_pdf / (2.0f * sqr(kPi) * st);
            out.L   = ibl_lookup(s, l, d0, d1);
            out.u   = d0;
            out.v   = d1;
        }
    }
    out.t_min = ray_offset(n, out.wi);
    return out;
}

// Light::pdf: SphereLight -> Sphere::pdf; EnvironmentLight :163-166; ImageBasedEnvironmentLight :251-266 (theta * pi
// as written there) over Distribution2D::pdf (math/Distribution2D.h:31-38)
template <typename F>
__device__ __forceinline__ float light_pdf(const DScene& s, const spcu_light& l, V3 p, V3 wi)
{
    if (l.kind == SPCU_LIGHT_SPHERE) {
        return sphere_pdf(l, p);
    }
    if (!F::ibl || l.kind == SPCU_LIGHT_ENV_CONST) {
        return 1.0f / (4.0f * kPi);
    }
    const V3    w     = xf_vector(l.world_to_light, wi);
    const float theta = spherical_theta(w);
    const float phi   = spherical_phi(w);
    const float st    = sinf(theta);
    if (st == 0.0f) {
        return 0.0f;
    }
    const float pu = phi * kInv2Pi, pv = theta * kPi;
    // static_cast<size_t> of a non-negative float, then clamp to the last cell
    const uint32_t iu = static_cast<uint32_t>(fminf(pu * static_cast<float>(l.nu), static_cast<float>(l.nu - 1u)));
    const uint32_t iv = static_cast<uint32_t>(fminf(pv * static_cast<float>(l.nv), static_cast<float>(l.nv - 1u)));
    const float    d2 = __ldg(s.pool + l.cond_func_off + static_cast<uint64_t>(iv) * l.nu + iu) / l.marg_integral;
    return d2 / (2.0f * sqr(kPi) * st);
}

// LightIntersection::L (Light::intersect_lights_impl: SphereLight :354-361, EnvironmentLight :135-141,
// ImageBasedEnvironmentLight :196-209)
template <typename F>
__device__ __forceinline__ V3 light_hit_L(const DScene& s, const spcu_light& l, V3 dir)
{
    if (!F::ibl || l.kind != SPCU_LIGHT_ENV_IBL) {
        return v3(l.radiance[0], l.radiance[1], l.radiance[2]);
    }
    const V3 w = normalize(xf_vector(l.world_to_light, dir));
    return ibl_lookup(s, l, spherical_phi(w) * kInv2Pi, spherical_theta(w) * kInvPi);
}

__device__ __forceinline__ V3 xyz(const float4 v) { return v3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 f4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

// PerspectiveCamera::generate_ray_impl (Cameras/Camera.h:119-129): direction = normalize(px*col0 + py*col1 + col2),
// products and sums rounded separately as in the canonical reference build; normalize() there is the SSE rsqrt
// estimate + one Newton step (math/Math.h:205-227), which no GPU instruction reproduces — here it is the correctly
// rounded reciprocal square root, so directions agree to a few ulp, not bitwise.
__device__ __forceinline__ void camera_ray(const DScene& s, uint32_t pix, uint32_t smp, float4& o, float4& d)
{
    const uint32_t x  = pix % s.width;
    const uint32_t y  = pix / s.width;
    const float    px = __fadd_rn(static_cast<float>(static_cast<int>(x)), __ldg(s.jitter + 2 * smp + 0)); // main.cpp:97
    const float    py = __fadd_rn(static_cast<float>(static_cast<int>(y)), __ldg(s.jitter + 2 * smp + 1));
    const float*   m  = s.camera;
    const float    dx = __fadd_rn(__fadd_rn(__fmul_rn(px, m[0]), __fmul_rn(py, m[3])), m[6]);
    const float    dy = __fadd_rn(__fadd_rn(__fmul_rn(px, m[1]), __fmul_rn(py, m[4])), m[7]);
    const float    dz = __fadd_rn(__fadd_rn(__fmul_rn(px, m[2]), __fmul_rn(py, m[5])), m[8]);
    const float    len2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fadd_rn(__fmul_rn(dz, dz), 0.0f));
    const float    inv  = __frsqrt_rn(len2);
    o = make_float4(m[9], m[10], m[11], 0.001f);                                             // RayLimits default t_min
    d = make_float4(__fmul_rn(dx, inv), __fmul_rn(dy, inv), __fmul_rn(dz, inv), FLT_MAX);     // ... and t_max
}

// Intersection record of the accepted hit: Triangle.h:148-160, Sphere.h:99-104, Plane.h:65-70
template <typename F>
__device__ __forceinline__ void make_isect(const DScene& s, const HitRec& h, V3 o, V3 d, V3& point, V3& normal,
                                           uint32_t& material)
{
    const uint32_t meta = __ldg(s.geom_meta + h.id);
    const float4   a    = __ldg(s.geom_shade + 3 * h.id + 0);
    const float4   b    = __ldg(s.geom_shade + 3 * h.id + 1);
    const float4   c    = __ldg(s.geom_shade + 3 * h.id + 2);
    material            = SPCU_META_MATERIAL(meta);
    point               = o + d * h.t; // Ray::operator() (math/Ray.h:30-34)
    const uint32_t kind = SPCU_META_KIND(meta);
    if (F::triangles && kind == SPCU_PRIM_TRIANGLE) {
        const float alpha = 1.0f - h.beta - h.gamma;
        normal = normalize(v3(fmaf(alpha, a.x, fmaf(h.beta, b.x, h.gamma * c.x)), fmaf(alpha, a.y, fmaf(h.beta, b.y, h.gamma * c.y)),
                              fmaf(alpha, a.z, fmaf(h.beta, b.z, h.gamma * c.z))));
    } else if (kind == SPCU_PRIM_SPHERE) {
        const float4 m0 = __ldg(s.geom_prims + 3 * h.id + 0);
        const float4 m1 = __ldg(s.geom_prims + 3 * h.id + 1);
        const float4 m2 = __ldg(s.geom_prims + 3 * h.id + 2);
        const float  m[12] = { m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w };
        const V3     lo = xf_point(m, o);
        const V3     ld = xf_vector(m, d);
        const V3     n  = v3(fmaf(h.t, ld.x, lo.x), fmaf(h.t, ld.y, lo.y), fmaf(h.t, ld.z, lo.z)); // madd(t, d, o) / k_radius
        normal = normalize(v3(fmaf(n.x, a.x, fmaf(n.y, b.x, n.z * c.x)), fmaf(n.x, a.y, fmaf(n.y, b.y, n.z * c.y)),
                              fmaf(n.x, a.z, fmaf(n.y, b.z, n.z * c.z))));
    } else {
        normal = v3(b.x, b.y, b.z); // object_to_world(Normal3{0,1,0}) = second column of the normal matrix, NOT normalised
    }
}

// balance_heuristic(1, f_pdf, 1, g_pdf) (math/Math.h:91-94)
__device__ __forceinline__ float balance2(float f_pdf, float g_pdf)
{
    const float inner = f_pdf + g_pdf;
    return inner == 0.0f ? 0.0f : f_pdf / inner;
}

} // namespace spcu
