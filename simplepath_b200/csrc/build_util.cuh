// Shared by the construction-side translation units (build_kernels.cu, mesh_kernels.cu): scratch allocations that live for one
// call, grid sizing, and an exclusive prefix sum over n (+1) elements in three passes — per-tile sums, one block over the tile
// sums, per-tile scan — for byte and 32-bit inputs.  Inputs must be zero-padded to a whole number of tiles.
#pragma once

#include "ctx.h"

namespace spcu {
namespace bu { // a NAMED namespace: nvcc's host stubs cannot tell two anonymous namespaces of one translation unit apart

constexpr int      kBlock    = 256;
constexpr uint32_t kScanTile = 2048; // elements per block of the prefix sum: 256 threads x 8

struct Scratch
{
    std::vector<DevBuf> bufs;
    ~Scratch()
    {
        for (auto& b : bufs) {
            b.release();
        }
    }
    template <typename T>
    cudaError_t get(T** out, size_t count)
    {
        bufs.emplace_back();
        const cudaError_t e = bufs.back().reserve(std::max<size_t>(count * sizeof(T), 16));
        *out                = bufs.back().as<T>();
        return e;
    }
};

inline unsigned grid_for(uint64_t n, int sm_count)
{
    const uint64_t want = (n + kBlock - 1) / kBlock;
    return static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(want, static_cast<uint64_t>(sm_count) * 8)));
}

inline uint32_t scan_tiles(uint32_t n) { return n / kScanTile + 1; } // covers index n

// eight consecutive elements of a thread
struct Eight
{
    uint32_t v[8];
};

__device__ __forceinline__ Eight load_eight(const uint8_t* in, size_t t)
{
    const uint2 p = reinterpret_cast<const uint2*>(in)[t];
    Eight       e;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        e.v[j] = ((j < 4 ? p.x : p.y) >> (8 * (j & 3))) & 0xFFu;
    }
    return e;
}

__device__ __forceinline__ Eight load_eight(const uint32_t* in, size_t t)
{
    const uint4 a = reinterpret_cast<const uint4*>(in)[2 * t], b = reinterpret_cast<const uint4*>(in)[2 * t + 1];
    return Eight{ { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w } };
}

__device__ __forceinline__ uint32_t sum_eight(const Eight& e)
{
    return ((e.v[0] + e.v[1]) + (e.v[2] + e.v[3])) + ((e.v[4] + e.v[5]) + (e.v[6] + e.v[7]));
}

__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[kBlock / 32];
    const uint32_t      lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t            inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, off);
        if (lane >= static_cast<uint32_t>(off)) {
            inc += o;
        }
    }
    __syncthreads(); // warp_sums may still be read by a previous call
    if (lane == 31u) {
        warp_sums[w] = inc;
    }
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (int i = 0; i < kBlock / 32; ++i) {
        const uint32_t s = warp_sums[i];
        base += static_cast<uint32_t>(i) < w ? s : 0u;
        all += s;
    }
    *total = all;
    return base + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_scan_reduce(const T* in, uint32_t* partials)
{
    const Eight e = load_eight(in, static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x);
    uint32_t    total;
    block_exclusive(sum_eight(e), &total);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
    }
}

static __global__ void __launch_bounds__(kBlock) k_scan_partials(uint32_t* partials, uint32_t count)
{
    uint32_t running = 0;
    for (uint32_t base = 0; base < count; base += kBlock) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < count ? partials[i] : 0u;
        uint32_t       total;
        const uint32_t e = block_exclusive(v, &total);
        if (i < count) {
            partials[i] = running + e;
        }
        running += total;
    }
}

// out[i] = sum of in[0 .. i) for i in [0, n]  (in[] is zero beyond n)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_scan_apply(const T* in, const uint32_t* partials, uint32_t* out, uint32_t n)
{
    const size_t t = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x;
    const Eight  e = load_eight(in, t);
    uint32_t     total;
    uint32_t     run = partials[blockIdx.x] + block_exclusive(sum_eight(e), &total);
    const size_t i0  = t * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (i0 + j <= n) {
            out[i0 + j] = run;
        }
        run += e.v[j];
    }
}

// in: scan_tiles(n) * kScanTile elements (zero beyond n); out: n + 1; partials: scan_tiles(n)
template <typename T>
inline void exclusive_scan(const T* in, uint32_t* out, uint32_t n, uint32_t* partials, cudaStream_t st)
{
    const uint32_t n_tiles = scan_tiles(n);
    k_scan_reduce<T><<<n_tiles, kBlock, 0, st>>>(in, partials);
    k_scan_partials<<<1, kBlock, 0, st>>>(partials, n_tiles);
    k_scan_apply<T><<<n_tiles, kBlock, 0, st>>>(in, partials, out, n);
}

} // namespace bu
using namespace bu;
} // namespace spcu
