// Wavefront stages that do not traverse: camera ray generation, BSDF sampling / evaluation, light sampling,
// Russian roulette, accumulation.  (Traversal stages live in trace_kernels.cu.)
#include "kernels.h"
#include "rng.cuh"
#include "shade.cuh"

#include <algorithm>
#include <cfloat>
#include <vector>

namespace spcu {
namespace {

constexpr int kShadeBlock = 128;

inline unsigned grid_for(uint64_t n, int block = kShadeBlock)
{
    return static_cast<unsigned>((n + block - 1) / block);
}

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

__global__ void __launch_bounds__(kShadeBlock) k_generate_rays(const __grid_constant__ DScene s, const uint32_t* pix,
                                                               const uint32_t* smp, uint64_t n, spcu_ray* rays)
{
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * kShadeBlock + threadIdx.x;
    if (i >= n) {
        return;
    }
    float4 o, d;
    camera_ray(s, pix[i], smp[i], o, d);
    reinterpret_cast<float4*>(rays)[2 * i + 0] = o;
    reinterpret_cast<float4*>(rays)[2 * i + 1] = d;
}

// Morton decode of the low 6 bits: TilePixelIterator visits an 8x8 tile in Morton order (base/Tile.h:134-138,
// math/Morton.h:87-92): x = even bits, y = odd bits.
__device__ __forceinline__ void morton8(uint32_t m, uint32_t& x, uint32_t& y)
{
    x = (m & 1u) | ((m >> 1) & 2u) | ((m >> 2) & 4u);
    y = ((m >> 1) & 1u) | ((m >> 2) & 2u) | ((m >> 3) & 4u);
}

// Pixel list of a partition: tiles t = offset, offset+stride, ... in the scheduler's row-major tile order
// (base/TileScheduler.h:66-82), 64 Morton-ordered candidates per tile, those outside the image dropped
// (main.cpp:90-91).  Slot j of the list is computed directly: prefix[k] = pixels in the first k owned tiles is
// evaluated in closed form from the number of full / clipped tiles, so no scan is needed.
__global__ void k_build_pixel_list(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride,
                                   uint32_t n_owned_tiles, const uint32_t* tile_prefix, uint32_t* pix_list)
{
    const uint32_t k = blockIdx.x; // owned tile index
    if (k >= n_owned_tiles) {
        return;
    }
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t t       = tile_offset + k * tile_stride;
    const uint32_t tx = t % tiles_x, ty = t / tiles_x;
    uint32_t       mx, my;
    morton8(threadIdx.x, mx, my);
    const uint32_t x = tx * 8 + mx, y = ty * 8 + my;
    const bool     inside = x < width && y < height;
    // rank of this Morton index among the inside pixels of the tile
    __shared__ uint32_t flags[64];
    flags[threadIdx.x] = inside ? 1u : 0u;
    __syncthreads();
    uint32_t rank = 0;
    for (uint32_t m = 0; m < threadIdx.x; ++m) {
        rank += flags[m];
    }
    if (inside) {
        pix_list[tile_prefix[k] + rank] = y * width + x;
    }
}

} // namespace

void launch_generate_rays(const DScene& s, const uint32_t* d_pix, const uint32_t* d_smp, uint64_t n, spcu_ray* d_rays,
                          cudaStream_t st)
{
    if (n == 0) return;
    k_generate_rays<<<grid_for(n), kShadeBlock, 0, st>>>(s, d_pix, d_smp, n, d_rays);
}

static uint32_t tile_pixels(uint32_t width, uint32_t height, uint32_t t)
{
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t tx = t % tiles_x, ty = t / tiles_x;
    const uint32_t w = std::min(8u, width - tx * 8), h = std::min(8u, height - ty * 8);
    return w * h;
}

uint32_t count_partition_pixels(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride)
{
    const uint32_t n_tiles = ((width + 7) / 8) * ((height + 7) / 8);
    uint64_t       n       = 0;
    for (uint32_t t = tile_offset; t < n_tiles; t += tile_stride) {
        n += tile_pixels(width, height, t);
    }
    return static_cast<uint32_t>(n);
}

cudaError_t build_pixel_list(uint32_t width, uint32_t height, uint32_t tile_offset, uint32_t tile_stride, uint32_t* d_pix_list,
                             uint32_t* d_tile_prefix_scratch, cudaStream_t st)
{
    const uint32_t        n_tiles = ((width + 7) / 8) * ((height + 7) / 8);
    std::vector<uint32_t> prefix;
    uint32_t              run = 0;
    for (uint32_t t = tile_offset; t < n_tiles; t += tile_stride) {
        prefix.push_back(run);
        run += tile_pixels(width, height, t);
    }
    if (prefix.empty()) {
        return cudaSuccess;
    }
    cudaError_t e = cudaMemcpyAsync(d_tile_prefix_scratch, prefix.data(), prefix.size() * sizeof(uint32_t),
                                    cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        return e;
    }
    k_build_pixel_list<<<static_cast<unsigned>(prefix.size()), 64, 0, st>>>(width, height, tile_offset, tile_stride,
                                                                           static_cast<uint32_t>(prefix.size()),
                                                                           d_tile_prefix_scratch, d_pix_list);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        return e;
    }
    return cudaStreamSynchronize(st); // `prefix` must outlive the copy
}

} // namespace spcu
