// RNG contract of the backend (the reference's per-pixel mt19937_64 stream, main.cpp:73-76, cannot be replayed on a
// GPU: SURVEY.md §8a-3).  Random numbers are addressed instead of streamed:
//
//     block k of sub-stream s of path (pixel, sample) = Philox4x32-10( counter = (k, s, seed_lo, seed_hi), key = (pixel, sample) )
//
// A sub-stream belongs to one call site of the integrator at one path depth, s = depth << 16 | site:
//     site 0      the vertex's primary BSDF sample            (Integrator.cpp:569 / :228)
//     site 1      Russian roulette                            (Integrator.cpp:615 / :246)
//     site 2 + l  next-event estimation for light l: block 0 = the light sample (Integrator.cpp:497 / :289), blocks 1.. =
//                 Material::eval, pdf and the second sample, in the reference's order (:508-518 / :296)
// Within a sub-stream every *draw call* of the reference (Sampler::get_next_1D or get_next_2D, math/Sampler.h:76-94)
// consumes ONE block, in call order: 1D uses word 0, 2D uses words 0 and 1; the two get_next_1D() arguments of
// beckmann_sample (materials/Material.cpp:150) are served by one block (U1 = word 0, U2 = word 1).  A word w maps to the
// float (w >> 8) * 2^-24 in [0, 1).  No stage has to carry a draw counter to the next one, and the lights of a vertex
// are independent of each other.  oracle/sp_oracle_shade.inc restates the same contract, so both sides draw identical
// numbers.
#pragma once

#include <stdint.h>

namespace spcu {

struct Rng
{
    uint32_t pixel, sample;
    uint32_t seed_lo, seed_hi;
    uint32_t stream; // depth << 16 | site
    uint32_t ctr;    // next block of the sub-stream
};

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
static __device__ __noinline__ uint2 philox_block01(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    uint32_t o[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, o);
    return make_uint2(o[0], o[1]);
}
#endif

// One draw call: returns words 0 and 1 as floats in [0,1) and advances the path's counter.
__host__ __device__ __forceinline__ void rng_next2(Rng& r, float& u0, float& u1)
{
#ifdef __CUDA_ARCH__
    const uint2 w = philox_block01(r.ctr, r.stream, r.seed_lo, r.seed_hi, r.pixel, r.sample);
    ++r.ctr;
    u0 = word_to_unit(w.x);
    u1 = word_to_unit(w.y);
#else
    uint32_t o[4];
    philox4x32_10(r.ctr, r.stream, r.seed_lo, r.seed_hi, r.pixel, r.sample, o);
    ++r.ctr;
    u0 = word_to_unit(o[0]);
    u1 = word_to_unit(o[1]);
#endif
}

__host__ __device__ __forceinline__ float rng_next1(Rng& r)
{
    float u0, u1;
    rng_next2(r, u0, u1);
    return u0;
}

} // namespace spcu
