// RNG contract of the backend (the reference's per-pixel mt19937_64 stream, main.cpp:73-76, cannot be replayed on a
// GPU: SURVEY.md §8a-3).  Random numbers are addressed instead of streamed:
//
//     block k of sub-stream s of path (pixel, sample) = Philox4x32-10( counter = (k, s, seed_lo, seed_hi), key = (pixel, sample) )
//
// A sub-stream belongs to one call site of the integrator at one path depth, s = depth << 16 | site:
//     site 0      the vertex's primary BSDF sample            (Integrator.cpp:569 / :228)
//     site 1      Russian roulette                            (Integrator.cpp:615 / :246)
//     site 2 + l  next-event estimation for light l: draw 0 = the light sample (Integrator.cpp:497 / :289), draws 1.. =
//                 Material::eval, pdf and the second sample, in the reference's order (:508-518 / :296)
// Within a sub-stream the *draw calls* of the reference (Sampler::get_next_1D or get_next_2D, math/Sampler.h:76-94) are
// numbered in call order, and draw d is served by HALF a block: block d >> 1, words (0, 1) when d is even and (2, 3) when it
// is odd; a 1D draw uses the first word of its pair, a 2D draw both; the two get_next_1D() arguments of beckmann_sample
// (materials/Material.cpp:150) are one draw (U1, U2 = the pair).  (Round 1 spent a whole block per draw: the ten rounds
// are ~70 instructions and a microfacet vertex draws ~50 times per light — 40 % of the NEE stage's instructions.)  A word w
// maps to the float (w >> 8) * 2^-24 in [0, 1).  The mapping is a pure function of (path, sub-stream, d): no stage has to
// carry generator state to the next one — a stage that starts at an odd draw recomputes that block —, and the lights of a
// vertex are independent of each other.  oracle/sp_oracle_shade.inc restates the same contract, so both sides draw
// identical numbers.
#pragma once

#include <stdint.h>

namespace spcu {

struct Rng
{
    uint32_t pixel, sample;
    uint32_t seed_lo, seed_hi;
    uint32_t stream; // depth << 16 | site
    uint32_t ctr;    // next draw of the sub-stream
    // words 2 and 3 of block ctr >> 1, kept from the even draw for the odd one that follows (`fresh`); not part of the contract
    uint32_t w2 = 0u, w3 = 0u;
    bool     fresh = false;
};

// continue at draw `ctr` of another sub-stream
__host__ __device__ __forceinline__ void rng_seek(Rng& r, uint32_t stream, uint32_t ctr)
{
    r.stream = stream;
    r.ctr    = ctr;
    r.fresh  = false;
}

constexpr uint32_t kSiteBsdf = 0u, kSiteRoulette = 1u, kSiteLight0 = 2u;
__host__ __device__ __forceinline__ uint32_t rng_stream(uint32_t depth, uint32_t site) { return (depth << 16) | site; }

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__host__ __device__ __forceinline__ float word_to_unit(uint32_t w)
{
    return static_cast<float>(w >> 8) * 5.9604644775390625e-8f; // 2^-24
}

#ifdef __CUDACC__
// Words 0 and 1 of a block as ONE out-of-line copy per kernel: the ten rounds are ~70 instructions and a kernel draws
// at some twenty-five call sites; inlined they were 11 % of the path kernels' SASS, whose limiter is instruction fetch.
static __device__ __noinline__ uint4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
#endif

// One draw call: returns its pair of words as floats in [0,1) and advances the path's draw counter.
#ifdef SPCU_RNG_NOINLINE
__host__ __device__ __noinline__ void rng_next2(Rng& r, float& u0, float& u1)
#else
__host__ __device__ __forceinline__ void rng_next2(Rng& r, float& u0, float& u1)
#endif
{
#if defined(SPCU_RNG_WHOLE_BLOCK) && defined(__CUDA_ARCH__) // A/B only (profiles/): round 1's contract, a whole block per draw — NOT what the oracle draws
    {
        const uint4 w = philox_block(r.ctr, r.stream, r.seed_lo, r.seed_hi, r.pixel, r.sample);
        ++r.ctr;
        u0 = word_to_unit(w.x);
        u1 = word_to_unit(w.y);
        return;
    }
#endif
    const bool odd = (r.ctr & 1u) != 0u;
    uint32_t   a, b;
    if (odd && r.fresh) {
        a       = r.w2;
        b       = r.w3;
        r.fresh = false;
    } else {
#ifdef __CUDA_ARCH__
        const uint4 w = philox_block(r.ctr >> 1, r.stream, r.seed_lo, r.seed_hi, r.pixel, r.sample);
#else
        uint32_t o[4];
        philox4x32_10(r.ctr >> 1, r.stream, r.seed_lo, r.seed_hi, r.pixel, r.sample, o);
        const struct { uint32_t x, y, z, w; } w = { o[0], o[1], o[2], o[3] };
#endif
        a       = odd ? w.z : w.x;
        b       = odd ? w.w : w.y;
        r.w2    = w.z;
        r.w3    = w.w;
        r.fresh = !odd;
    }
    ++r.ctr;
    u0 = word_to_unit(a);
    u1 = word_to_unit(b);
}

__host__ __device__ __forceinline__ float rng_next1(Rng& r)
{
    float u0, u1;
    rng_next2(r, u0, u1);
    return u0;
}

} // namespace spcu
