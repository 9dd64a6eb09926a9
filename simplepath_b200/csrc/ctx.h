// Private to the backend's host code: the context behind the opaque spcu_ctx handle and error plumbing.
#pragma once

#include "kernels.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace spcu {

struct DevBuf
{
    void*  p     = nullptr;
    size_t bytes = 0;

    cudaError_t reserve(size_t n)
    {
        if (n <= bytes) {
            return cudaSuccess;
        }
        if (p) {
            cudaFree(p);
            p     = nullptr;
            bytes = 0;
        }
        const cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) {
            bytes = n;
        }
        return e;
    }

    void release()
    {
        if (p) {
            cudaFree(p);
        }
        p     = nullptr;
        bytes = 0;
    }

    template <typename T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

constexpr int      kNumQueues      = 7;        // cur, next, live, shadow, lit, mis, walk
constexpr int      kMaxQueueCounts = 4096;     // device queue-length words zeroed once per batch
constexpr uint64_t kBatchRays      = 1u << 22; // rays per chunk of the batch query entry points
constexpr uint32_t kMaxMaterialSegments = 15;  // materials beyond share the last segment

constexpr int      kMaxBatchLanes  = 8;        // SPCU_OPT_BATCH_LANES
constexpr int      kDefaultBatchLanes = 4;

// Wavefront state of one batch in flight beyond the first (whose state lives in the context itself): SPCU_OPT_BATCH_LANES
struct WaveLane
{
    DWave               wave{};
    std::vector<DevBuf> wave_bufs;
    DevBuf              queues[kNumQueues];
    DevBuf              queue_counts;
    DevBuf              sorted_queue;
    uint32_t            wave_lights = 0;
    cudaStream_t        stream   = nullptr;
    cudaEvent_t         resolved = nullptr; // recorded after each of the lane's resolve kernels

    void release()
    {
        for (auto& b : wave_bufs) b.release();
        wave_bufs.clear();
        for (auto& q : queues) q.release();
        queue_counts.release();
        sorted_queue.release();
        wave.capacity = 0;
    }
};

} // namespace spcu

struct spcu_ctx
{
    int          device = 0;
    std::string  err;
    cudaStream_t stream = nullptr;
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    int          sm_count = 148;
    void*        nccl_comm   = nullptr; // ncclComm_t of this context's rank (comm.cu); NULL: single GPU
    int          comm_rank   = 0;
    int          comm_nranks = 1;
    bool         scratch_in_use = false;   // a render has been enqueued: its last event is ev1 on scratch_stream
    cudaStream_t scratch_stream = nullptr;

    // scene
    bool         have_scene = false;
    spcu::DScene ds{};
    spcu::DevBuf geom_wide; // 4-wide copy of geom_nodes (built at upload)
    spcu::DevBuf geom_big;  // its side table of oversized leaves + the slot counter (device_scene.h)
    spcu::DevBuf geom_nodes, geom_prims, geom_shade, geom_meta, light_nodes, lights, light_order, materials, bxdfs, pool,
        jitter;
    uint64_t scene_bytes = 0;
    spcu_accel built_geom{}; // geometry accelerator header of the resident scene (nodes = NULL: they live on the device)

    // batch-query scratch
    spcu::DevBuf q_rays, q_out, q_aux, q_cnt;

    // wavefront
    uint64_t                  wavefront_size = 0; // 0 = default
    spcu::DWave               wave{};
    std::vector<spcu::DevBuf> wave_bufs;
    spcu::DevBuf              queues[spcu::kNumQueues];
    spcu::DevBuf              queue_counts; // uint32[kMaxQueueCounts]
    spcu::DevBuf              counters;     // unsigned long long[kNumCounters] followed by TraceCounters
    spcu::DevBuf              pix_list;
    uint32_t                  pix_list_offset = ~0u, pix_list_stride = 0, pix_list_n = 0;
    spcu::DevBuf              host_rgb, host_sq; // device accumulators of the host-buffer entry point
    spcu::DevBuf              packed;            // packed output image (spcu_render_image / spcu_pack_image)
    void*                     stage[2]    = { nullptr, nullptr }; // page-locked staging of large host->device copies
    cudaEvent_t               stage_ev[2] = { nullptr, nullptr };
    bool                      stage_busy[2] = { false, false };
    spcu::DevBuf              path_radiance;     // float4 per slot of a batch (SPCU_PIPELINE_PATHS)
    spcu::DevBuf              sorted_queue;      // material-sorted hand-over between extend and shade
    uint32_t                  n_materials = 0;
    int                       features    = 0; // feature set of the uploaded scene (features.h)
    uint32_t                  wave_lights = 0; // light planes the wavefront buffers were sized for
    cudaEvent_t               resolved = nullptr;       // lane 0's "resolve done" event
    std::vector<spcu::WaveLane> extra_lanes;            // batches in flight beyond the first (SPCU_OPT_BATCH_LANES)

    uint32_t                 options[SPCU_OPT_COUNT_] = {};
    std::vector<cudaEvent_t> stage_events; // pairs, when SPCU_OPT_STAGE_TIMING is on
    std::vector<int>         stage_kinds;  // spcu::Stage of each pair
    spcu_stage_time          stage_report[spcu::kNumStages] = {};
};

namespace spcu {

int fail(spcu_ctx* c, int code, const char* fmt, ...);
int need_scene(spcu_ctx* c);
// spcu_api.cu: host->device copy on the context's stream.  Large copies from pageable memory go through two page-locked
// staging buffers (CPU memcpy of chunk k+1 overlaps the DMA of chunk k); the source is fully consumed when the call returns.
int copy_to_device(spcu_ctx* c, void* dst, const void* src, size_t bytes);
// image_kernels.cu: mean, row order and sRGB quantisation of device-resident sums, result copied to the host buffer `out`
int pack_device_image(spcu_ctx* c, const float* d_rgb_sum, uint32_t w, uint32_t h, uint32_t spp, uint32_t format, void* out);
// build_kernels.cu: geometry of an UNBUILT scene -> bounds, BVH and leaf-order gather on the device (spcu_upload_scene_build)
int build_scene_geometry(spcu_ctx* c, const spcu_flat_scene* s, const spcu_bounds* extra_bounds, uint32_t* order_out,
                         spcu_accel* built, bool* proper_boxes);

#define CK(ctx, call)                                                                                                     \
    do {                                                                                                                  \
        const cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) {                                                                                          \
            return spcu::fail((ctx), SPCU_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                                 \
    } while (0)

} // namespace spcu
