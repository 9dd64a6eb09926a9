// Microbenchmark behind the quad-cooperative node step: how fast can the SMs fetch scattered 128-byte
// nodes from an L2-resident array when (A) every lane loads its own node with four 256-bit loads (32 distinct lines per
// load instruction) or (B) the four lanes of a quad load one 32-byte sector each of ONE node per round, four rounds
// (8 distinct lines per load instruction)?  Both variants follow a dependent chain (the next index comes out of the
// data), as a traversal does.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a node_fetch.cu -o node_fetch
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const void* p)
{
    F8 r; unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    r.lo = make_float4(__uint_as_float((unsigned)a), __uint_as_float((unsigned)(a >> 32)), __uint_as_float((unsigned)b), __uint_as_float((unsigned)(b >> 32)));
    r.hi = make_float4(__uint_as_float((unsigned)c), __uint_as_float((unsigned)(c >> 32)), __uint_as_float((unsigned)d), __uint_as_float((unsigned)(d >> 32)));
    return r;
}
__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ unsigned fold(const F8& v)
{
    return __float_as_uint(v.lo.x) ^ __float_as_uint(v.lo.y) ^ __float_as_uint(v.lo.z) ^ __float_as_uint(v.lo.w) ^
           __float_as_uint(v.hi.x) ^ __float_as_uint(v.hi.y) ^ __float_as_uint(v.hi.z) ^ __float_as_uint(v.hi.w);
}

template <int kAlu>
__global__ void __launch_bounds__(128, 8) k_lane(const float4* nodes, unsigned n_nodes, int steps, unsigned* out)
{
    unsigned st = mix(blockIdx.x * 128 + threadIdx.x), idx = st % n_nodes, acc = 0;
    for (int s = 0; s < steps; ++s) {
        const float4* p = nodes + 8 * (size_t)idx;
        const F8 a = ldg256(p), b = ldg256(p + 2), c = ldg256(p + 4), d = ldg256(p + 6);
        unsigned h = fold(a) ^ fold(b) ^ fold(c) ^ fold(d);
#pragma unroll
        for (int i = 0; i < kAlu; ++i) h = h * 1664525u + 1013904223u;
        acc += h;
        st = mix(st + h + s); idx = st % n_nodes;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int kAlu>
__global__ void __launch_bounds__(128, 8) k_quad(const float4* nodes, unsigned n_nodes, int steps, unsigned* out)
{
    const int lane = threadIdx.x & 31, j = lane & 3, qb = lane & ~3;
    unsigned st = mix(blockIdx.x * 128 + threadIdx.x), idx = st % n_nodes, acc = 0;
    for (int s = 0; s < steps; ++s) {
        unsigned h = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned ik = __shfl_sync(0xffffffffu, idx, qb + k);
            const F8 v = ldg256(nodes + 8 * (size_t)ik + 2 * j);
            unsigned f = fold(v);
#pragma unroll
            for (int i = 0; i < kAlu / 4; ++i) f = f * 1664525u + 1013904223u;
            f ^= __shfl_xor_sync(0xffffffffu, f, 1);
            f ^= __shfl_xor_sync(0xffffffffu, f, 2);
            if (k == j) h = f;
        }
        acc += h;
        st = mix(st + h + s); idx = st % n_nodes;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <typename K>
static void run(const char* name, K kernel, const float4* d_nodes, unsigned n_nodes, unsigned* d_out, int grid, int steps)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<grid, 128>>>(d_nodes, n_nodes, steps / 8, d_out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kernel<<<grid, 128>>>(d_nodes, n_nodes, steps, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fetches = (double)grid * 128 * steps;
    printf("{\"variant\": \"%s\", \"nodes\": %u, \"ms\": %.3f, \"gnode_fetches_per_s\": %.2f, \"tb_per_s\": %.2f, \"err\": \"%s\"}\n", name, n_nodes, ms,
           fetches / ms * 1e-6, fetches * 128 / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, steps = 4096;
    unsigned* d_out; cudaMalloc(&d_out, (size_t)grid * 128 * 4);
    for (unsigned n_nodes : { 1u << 10, 55000u, 1u << 20 }) {  // 128 KB (L1-resident), 7 MB (bunny's wide nodes), 128 MB (beyond L2)
        float4* d_nodes; cudaMalloc(&d_nodes, (size_t)n_nodes * 128);
        cudaMemset(d_nodes, 0x3c, (size_t)n_nodes * 128);
        run("lane_4x256_alu0", k_lane<0>, d_nodes, n_nodes, d_out, grid, steps);
        run("quad_4rounds_alu0", k_quad<0>, d_nodes, n_nodes, d_out, grid, steps);
        run("lane_4x256_alu128", k_lane<128>, d_nodes, n_nodes, d_out, grid, steps);
        run("quad_4rounds_alu128", k_quad<128>, d_nodes, n_nodes, d_out, grid, steps);
        cudaFree(d_nodes);
    }
    return 0;
}
