// spcu_render / spcu_render_device: host orchestration of the wavefront loop.
//
// One call renders a partition (a set of 8x8 tiles x a sample range) in batches of at most `wavefront_size` paths:
//
//   raygen -> for depth in [0, max_depth):  extend -> shade (+ all light samples) -> { per light: shadow -> nee_bsdf -> mis ->
//             nee_mis_accumulate } -> advance        -> resolve (per-pixel sums, in sample order)
//
// Every stage is a persistent-style kernel reading its queue length from device memory, so a batch is enqueued without
// any host synchronisation; queues are compacted on the device (warp ballot + one atomic per warp).
#include "ctx.h"

#include <chrono>

using namespace spcu;

namespace {

enum Queue { kQCur = 0, kQNext, kQLive, kQShadow, kQLit, kQMis, kQWalk };
constexpr int kCounterBlock = kNumCounters + kNumStages; // + per-stage item counters

const char* const kStageNames[kNumStages] = { "raygen",   "extend",    "shade",     "nee_light",         "shadow",           "nee_bsdf",
                                              "mis_trace", "nee_mis_accumulate", "direct_accumulate", "advance", "resolve", "paths" };
inline bool stage_traverses(int st) { return st == kStExtend || st == kStShadow || st == kStMisTrace; }

template <typename T>
int wave_array(spcu_ctx* c, std::vector<DevBuf>& bufs, T*& ptr, size_t n)
{
    bufs.emplace_back();
    CK(c, bufs.back().reserve(std::max<size_t>(n, 1) * sizeof(T)));
    ptr = bufs.back().as<T>();
    return SPCU_OK;
}

// the wavefront state of one batch in flight: records, queues, queue counters
int ensure_wave_buffers(spcu_ctx* c, DWave& w, std::vector<DevBuf>& bufs, DevBuf* queues, DevBuf& queue_counts, uint32_t& wave_lights,
                        uint32_t capacity)
{
    if (w.capacity == capacity && wave_lights == c->ds.n_lights) {
        return SPCU_OK; // (capacity is also the stride of the per-light planes, so it must match exactly)
    }
    wave_lights = c->ds.n_lights;
    for (auto& b : bufs) {
        b.release();
    }
    bufs.clear();
    bufs.reserve(32);
    w.capacity = 0;
    int rc;
#define WAVE(field) \
    if ((rc = wave_array(c, bufs, w.field, capacity)) != SPCU_OK) return rc
    WAVE(path);
    WAVE(ray);
    WAVE(vertex);
    WAVE(extend);
    WAVE(s0);
    if ((rc = wave_array(c, bufs, w.light, static_cast<size_t>(capacity) * std::max(1u, c->ds.n_lights))) != SPCU_OK) return rc;
    WAVE(mis);
    WAVE(occluded);
#undef WAVE
    for (int i = 0; i < kNumQueues; ++i) { // the shadow queue has one plane per light
        const size_t planes = (i == kQShadow) ? std::max(1u, c->ds.n_lights) : 1u;
        CK(c, queues[i].reserve(static_cast<size_t>(capacity) * planes * sizeof(uint32_t)));
    }
    CK(c, queue_counts.reserve(kMaxQueueCounts * sizeof(uint32_t)));
    w.capacity = capacity;
    return SPCU_OK;
}

int ensure_wave(spcu_ctx* c, uint32_t capacity)
{
    if (int rc = ensure_wave_buffers(c, c->wave, c->wave_bufs, c->queues, c->queue_counts, c->wave_lights, capacity); rc != SPCU_OK) {
        return rc;
    }
    CK(c, c->counters.reserve(kCounterBlock * sizeof(unsigned long long) + sizeof(TraceCounters)));
    return SPCU_OK;
}

// Lane k >= 1 of SPCU_OPT_BATCH_LANES: its own stream, event and wavefront state.  A lane that cannot be allocated is not
// an error: the render goes on with the lanes it has (the caller shrinks its lane count to what this returns OK for).
int ensure_extra_lane(spcu_ctx* c, size_t k, uint32_t capacity, size_t sorted_bytes)
{
    if (c->extra_lanes.size() <= k) {
        c->extra_lanes.resize(k + 1);
    }
    WaveLane& lane = c->extra_lanes[k];
    if (!lane.stream) {
        CK(c, cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
        CK(c, cudaEventCreateWithFlags(&lane.resolved, cudaEventDisableTiming));
    }
    int rc = ensure_wave_buffers(c, lane.wave, lane.wave_bufs, lane.queues, lane.queue_counts, lane.wave_lights, capacity);
    if (rc == SPCU_OK && lane.sorted_queue.reserve(sorted_bytes) != cudaSuccess) {
        rc = SPCU_ERR_CUDA;
    }
    if (rc != SPCU_OK) {
        cudaGetLastError(); // (out of memory is sticky-free; clear it)
        lane.release();
    }
    return rc;
}

int ensure_pixel_list(spcu_ctx* c, const spcu_partition& part)
{
    if (c->pix_list_stride == part.tile_stride && c->pix_list_offset == part.tile_offset) {
        return SPCU_OK;
    }
    const uint32_t w = c->ds.width, h = c->ds.height;
    const uint32_t n_pix   = count_partition_pixels(w, h, part.tile_offset, part.tile_stride);
    const uint32_t n_tiles = ((w + 7) / 8) * ((h + 7) / 8);
    const uint32_t owned   = part.tile_offset < n_tiles ? (n_tiles - part.tile_offset + part.tile_stride - 1) / part.tile_stride : 0;
    // pixel list followed by the per-tile prefix scratch
    CK(c, c->pix_list.reserve((static_cast<size_t>(n_pix) + owned + 1) * sizeof(uint32_t)));
    CK(c, build_pixel_list(w, h, part.tile_offset, part.tile_stride, c->pix_list.as<uint32_t>(),
                           c->pix_list.as<uint32_t>() + n_pix, c->stream));
    c->pix_list_offset = part.tile_offset;
    c->pix_list_stride = part.tile_stride;
    c->pix_list_n      = n_pix;
    return SPCU_OK;
}

// Optional per-launch timing: an event pair around every kernel launch, summed per stage kind afterwards.
struct StageTimer
{
    spcu_ctx*    c;
    bool         on;
    cudaStream_t st;
    size_t       used = 0;

    void begin(int kind)
    {
        ++launches[kind];
        if (!on) return;
        if (used + 2 > c->stage_events.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            c->stage_events.push_back(a);
            c->stage_events.push_back(b);
        }
        c->stage_kinds.resize(c->stage_events.size() / 2);
        c->stage_kinds[used / 2] = kind;
        cudaEventRecord(c->stage_events[used], st);
    }
    void end()
    {
        if (!on) return;
        cudaEventRecord(c->stage_events[used + 1], st);
        used += 2;
    }
    uint64_t launches[kNumStages] = {};

    void collect(float& trace_ms, float& shade_ms)
    {
        trace_ms = shade_ms = 0.0f;
        for (size_t i = 0; i < used; i += 2) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, c->stage_events[i], c->stage_events[i + 1]);
            const int st = c->stage_kinds[i / 2];
            c->stage_report[st].ms += ms;
            (stage_traverses(st) ? trace_ms : shade_ms) += ms;
        }
    }
};

// Device counters -> spcu_stats (+ the per-stage report).  Synchronises `st`.
int read_back_stats(spcu_ctx* c, cudaStream_t st, spcu_stats* stats, uint64_t launches, StageTimer& timer)
{
    auto* d_counters = c->counters.as<unsigned long long>();
    CK(c, cudaEventSynchronize(c->ev1));
    unsigned long long h[kCounterBlock];
    TraceCounters      tc{};
    CK(c, cudaMemcpyAsync(h, d_counters, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(&tc, d_counters + kCounterBlock, sizeof tc, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    stats->paths           = h[kCntPaths];
    stats->rays_closest    = h[kCntRaysClosest];
    stats->rays_any        = h[kCntRaysAny];
    stats->rays_lights     = h[kCntRaysLights];
    stats->shade_calls     = h[kCntShadeCalls];
    stats->nodes_visited   = tc.nodes;
    stats->prims_tested    = tc.tris;
    stats->xf_prims_tested = tc.xf;
    stats->kernel_launches = launches;
    if (h[kCntErrors]) {
        return fail(c, SPCU_ERR_INTERNAL, "%llu kernel(s) gave up on a bounded wait (queue protocol fault); the image is void",
                    h[kCntErrors]);
    }
    CK(c, cudaEventElapsedTime(&stats->device_ms, c->ev0, c->ev1));
    for (int i = 0; i < kNumStages; ++i) {
        spcu_stage_time& r = c->stage_report[i];
        std::memset(&r, 0, sizeof r);
        std::snprintf(r.name, sizeof r.name, "%s", kStageNames[i]);
        r.launches  = timer.launches[i];
        r.items     = h[kNumCounters + i];
        r.traverses = stage_traverses(i) ? 1u : 0u;
    }
    timer.collect(stats->trace_ms, stats->shade_ms);
    return SPCU_OK;
}

// SPCU_PIPELINE_PATHS: one persistent kernel per batch (path_kernels.cu) + resolve.
int render_paths(spcu_ctx* c, const spcu_partition* part, float* d_rgb_sum, float* d_lum_sumsq, spcu_stats* stats,
                 cudaStream_t st, uint32_t n_pix, uint32_t n_samples, bool sm_local)
{
    const DScene& s = c->ds;
    // batch = whole pixel list x as many samples as fit in the radiance buffer (16 B per path)
    const uint64_t target = c->wavefront_size ? c->wavefront_size : (1ull << 27);
    uint32_t       pix_per_batch, smp_per_batch;
    if (n_pix <= target) {
        pix_per_batch = n_pix;
        smp_per_batch = static_cast<uint32_t>(std::min<uint64_t>(n_samples, std::max<uint64_t>(1, target / n_pix)));
    } else {
        pix_per_batch = static_cast<uint32_t>(target);
        smp_per_batch = 1;
    }
    const size_t capacity = static_cast<size_t>(pix_per_batch) * smp_per_batch;
    CK(c, c->path_radiance.reserve(capacity * sizeof(float4)));
    CK(c, c->counters.reserve(kCounterBlock * sizeof(unsigned long long) + sizeof(TraceCounters)));
    auto*          d_counters = c->counters.as<unsigned long long>();
    TraceCounters* d_cnt = c->options[SPCU_OPT_COUNT_NODES] ? reinterpret_cast<TraceCounters*>(d_counters + kCounterBlock) : nullptr;
    const uint32_t* d_pix_list = c->pix_list.as<uint32_t>();
    const Launch    L{ c->sm_count, st, c->features };
    StageTimer      timer{ c, c->options[SPCU_OPT_STAGE_TIMING] != 0, st };
    uint64_t        launches = 0;
    float4*         d_radiance = c->path_radiance.as<float4>();

    CK(c, cudaMemsetAsync(d_counters, 0, kCounterBlock * sizeof(unsigned long long) + sizeof(TraceCounters), st));
    CK(c, cudaEventRecord(c->ev0, st));
    for (uint32_t pb = 0; pb < n_pix; pb += pix_per_batch) {
        const uint32_t np = std::min(pix_per_batch, n_pix - pb);
        for (uint32_t sb = 0; sb < n_samples; sb += smp_per_batch) {
            const uint32_t ns = std::min(smp_per_batch, n_samples - sb);
            timer.begin(kStPaths);
            if (sm_local) {
                CK(c, launch_smwave(L, s, d_pix_list + pb, np, part->sample_begin + sb, ns, part->seed, part->integrator,
                                    d_radiance, d_counters, d_cnt));
            } else {
                launch_paths(L, s, d_pix_list + pb, np, part->sample_begin + sb, ns, part->seed, part->integrator, d_radiance,
                             d_counters, d_cnt);
            }
            timer.end();
            timer.begin(kStResolve);
            launch_resolve(L, d_radiance, 1, d_pix_list + pb, np, ns, d_rgb_sum, d_lum_sumsq, d_counters);
            timer.end();
            launches += 2;
            CK(c, cudaGetLastError());
        }
    }
    CK(c, cudaEventRecord(c->ev1, st));
    if (stats) {
        if (int rc = read_back_stats(c, st, stats, launches, timer); rc != SPCU_OK) return rc;
        c->stage_report[kStPaths].items = stats->paths;
    }
    return SPCU_OK;
}

// SPCU_PIPELINE_AUTO: the organisation that measures faster for the scene's feature set
uint32_t resolve_pipeline(const spcu_ctx* c)
{
    const uint32_t pipeline = c->options[SPCU_OPT_PIPELINE];
    if (pipeline != SPCU_PIPELINE_AUTO) {
        return pipeline;
    }
    return (c->features == FeatAnalytic::id && smwave_supports(c->ds)) ? SPCU_PIPELINE_SMWAVE : SPCU_PIPELINE_WAVEFRONT;
}

int render_impl(spcu_ctx* c, const spcu_partition* part, float* d_rgb_sum, float* d_lum_sumsq, spcu_stats* stats,
                cudaStream_t caller_stream, bool have_caller_stream)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    if (!part || !d_rgb_sum) {
        return fail(c, SPCU_ERR_INVALID, "partition or accumulator is NULL");
    }
    const DScene& s = c->ds;
    if (part->tile_stride == 0 || part->tile_offset >= part->tile_stride) {
        return fail(c, SPCU_ERR_INVALID, "bad tile partition %u/%u", part->tile_offset, part->tile_stride);
    }
    if (part->sample_begin > part->sample_end || part->sample_end > part->spp_total || part->spp_total > s.spp) {
        return fail(c, SPCU_ERR_INVALID, "bad sample range [%u,%u) of %u (jitter table holds %u)", part->sample_begin,
                    part->sample_end, part->spp_total, s.spp);
    }
    if (part->integrator > SPCU_INTEGRATOR_WHITTED) {
        return fail(c, SPCU_ERR_INVALID, "unknown integrator %u", part->integrator);
    }
    // The caller's stream (spcu_render_device) orders this call after the caller's earlier work on that stream.
    const cudaStream_t st = have_caller_stream ? caller_stream : c->stream;
    // The wavefront state, queues and counters are context-owned scratch: a render on ANOTHER stream than the previous one
    // must not start before that one has finished with them (ev1 is recorded at the end of every render).
    if (c->scratch_in_use && c->scratch_stream != st) {
        CK(c, cudaStreamWaitEvent(st, c->ev1, 0));
    }
    c->scratch_in_use = true;
    c->scratch_stream = st;

    if (int rc = ensure_pixel_list(c, *part); rc != SPCU_OK) return rc;
    const uint32_t n_pix     = c->pix_list_n;
    const uint32_t n_samples = part->sample_end - part->sample_begin;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
    }
    if (n_pix == 0 || n_samples == 0) {
        return SPCU_OK;
    }

    const uint32_t pipeline = resolve_pipeline(c);
    if (pipeline == SPCU_PIPELINE_SMWAVE && !smwave_supports(s)) {
        return fail(c, SPCU_ERR_INVALID, "SPCU_PIPELINE_SMWAVE: max_depth %u / %u lights exceed its packed path flags",
                    s.max_depth, s.n_lights);
    }
    if (pipeline == SPCU_PIPELINE_PATHS || pipeline == SPCU_PIPELINE_SMWAVE) {
        return render_paths(c, part, d_rgb_sum, d_lum_sumsq, stats, st, n_pix, n_samples, pipeline == SPCU_PIPELINE_SMWAVE);
    }

    // batch shape: whole pixel list x as many samples as fit, or a slice of the pixel list x one sample
    const uint64_t target = c->wavefront_size ? c->wavefront_size : (1ull << 24);
    uint32_t       pix_per_batch, smp_per_batch;
    if (n_pix <= target) {
        pix_per_batch = n_pix;
        smp_per_batch = static_cast<uint32_t>(std::min<uint64_t>(n_samples, std::max<uint64_t>(1, target / n_pix)));
    } else {
        pix_per_batch = static_cast<uint32_t>(target);
        smp_per_batch = 1;
    }
    const uint32_t capacity = pix_per_batch * smp_per_batch;
    if (int rc = ensure_wave(c, capacity); rc != SPCU_OK) return rc;

    const uint32_t n_lights    = s.n_lights;
    const uint32_t max_depth   = part->integrator == SPCU_INTEGRATOR_DIRECT_LIGHTING ? std::min(1u, s.max_depth) : s.max_depth;
    const bool     nee         = part->integrator == SPCU_INTEGRATOR_ITERATIVE_RRNEE;
    const bool     whitted     = part->integrator == SPCU_INTEGRATOR_WHITTED;
    const bool     direct      = part->integrator == SPCU_INTEGRATOR_DIRECT_LIGHTING || whitted; // per-light direct term
    const uint32_t n_segments  = std::min<uint32_t>(c->n_materials, kMaxMaterialSegments) + 1u; // + the miss segment
    const uint32_t counts_need = 1 + max_depth * (4 + 7 * n_lights + n_segments);
    CK(c, c->sorted_queue.reserve(static_cast<size_t>(n_segments) * capacity * sizeof(uint32_t)));
    if (counts_need > static_cast<uint32_t>(kMaxQueueCounts)) {
        return fail(c, SPCU_ERR_LIMIT, "max_depth x lights needs %u queue counters (limit %d)", counts_need, kMaxQueueCounts);
    }

    auto*          d_counters = c->counters.as<unsigned long long>();
    TraceCounters* d_cnt      = c->options[SPCU_OPT_COUNT_NODES] ? reinterpret_cast<TraceCounters*>(d_counters + kCounterBlock) : nullptr;
    const uint32_t* d_pix_list = c->pix_list.as<uint32_t>();
    StageTimer      timer{ c, c->options[SPCU_OPT_STAGE_TIMING] != 0, st };
    uint64_t        launches = 0;

    // ---- batches in flight (SPCU_OPT_BATCH_LANES): lane 0 = the context's own wavefront state on the call's stream ----------
    const uint64_t n_batches = static_cast<uint64_t>((n_pix + pix_per_batch - 1) / pix_per_batch) * ((n_samples + smp_per_batch - 1) / smp_per_batch);
    uint32_t       want_lanes = c->options[SPCU_OPT_BATCH_LANES] ? c->options[SPCU_OPT_BATCH_LANES] : kDefaultBatchLanes;
    if (timer.on || d_cnt) {
        want_lanes = 1; // per-launch events and the node counters describe one batch at a time
    }
    want_lanes = static_cast<uint32_t>(std::min<uint64_t>({ want_lanes, kMaxBatchLanes, n_batches }));
    struct LaneView
    {
        DWave        wave;
        uint32_t*    q[kNumQueues];
        uint32_t*    d_counts;
        uint32_t*    sorted;
        cudaStream_t st;
        cudaEvent_t  resolved;
    };
    std::vector<LaneView> lanes;
    if (!c->resolved) {
        CK(c, cudaEventCreateWithFlags(&c->resolved, cudaEventDisableTiming));
    }
    {
        LaneView v{ c->wave, {}, c->queue_counts.as<uint32_t>(), c->sorted_queue.as<uint32_t>(), st, c->resolved };
        for (int i = 0; i < kNumQueues; ++i) v.q[i] = c->queues[i].as<uint32_t>();
        lanes.push_back(v);
    }
    for (uint32_t k = 1; k < want_lanes; ++k) {
        if (ensure_extra_lane(c, k - 1, capacity, static_cast<size_t>(n_segments) * capacity * sizeof(uint32_t)) != SPCU_OK) {
            break; // not enough memory for another lane: render with the ones we have
        }
        WaveLane& l = c->extra_lanes[k - 1];
        LaneView  v{ l.wave, {}, l.queue_counts.as<uint32_t>(), l.sorted_queue.as<uint32_t>(), l.stream, l.resolved };
        for (int i = 0; i < kNumQueues; ++i) v.q[i] = l.queues[i].as<uint32_t>();
        lanes.push_back(v);
    }

    CK(c, cudaMemsetAsync(d_counters, 0, kCounterBlock * sizeof(unsigned long long) + sizeof(TraceCounters), st));
    CK(c, cudaEventRecord(c->ev0, st));
    for (size_t k = 1; k < lanes.size(); ++k) { // the other lanes start after whatever the call's stream held before
        CK(c, cudaStreamWaitEvent(lanes[k].st, c->ev0, 0));
    }

    uint64_t batch = 0;
    for (uint32_t pb = 0; pb < n_pix; pb += pix_per_batch) {
        const uint32_t np = std::min(pix_per_batch, n_pix - pb);
        for (uint32_t sb = 0; sb < n_samples; sb += smp_per_batch) {
            const uint32_t ns    = std::min(smp_per_batch, n_samples - sb);
            const uint32_t max_n = np * ns;
            LaneView&      lane     = lanes[batch % lanes.size()];
            uint32_t**     q        = lane.q;
            uint32_t*      d_counts = lane.d_counts;
            const Launch   L{ c->sm_count, lane.st, c->features };
            CK(c, cudaMemsetAsync(d_counts, 0, counts_need * sizeof(uint32_t), lane.st));
            uint32_t next_count = 0;
            auto     new_count  = [&]() { return d_counts + next_count++; };

            uint32_t* n_cur = new_count();
            timer.begin(kStRaygen);
            launch_raygen(L, s, lane.wave, d_pix_list + pb, np, part->sample_begin + sb, ns, q[kQCur], n_cur, d_counters);
            timer.end();
            ++launches;

            uint32_t* q_cur  = q[kQCur];
            uint32_t* q_next = q[kQNext];
            // advance (of depth d) + the set-up of extend (of depth d + 1) as ONE kernel where the extend stage has a `begin`
            // kernel to put it in; the live queue of depth d then goes to extend directly
            const bool fuse_advance = !whitted && extend_fuses_advance(L, q[kQWalk], d_cnt);
            bool       pending_advance = false;
            for (uint32_t depth = 0; depth < max_depth; ++depth) {
                RenderParams p{ part->seed, part->integrator, depth, 0 };
                SortedQueue sorted{ lane.sorted, d_counts + next_count, n_segments, capacity };
                next_count += n_segments;
                timer.begin(kStExtend);
                uint32_t*         cursor = new_count();
                const AdvanceArgs adv{ part->seed, depth - 1u };
                launches += launch_extend(L, s, lane.wave, q_cur, n_cur, max_n, cursor, sorted,
                                          // (node counting is DEFINED on the reference-order walk: DESIGN.md byte model)
                                          c->options[SPCU_OPT_TRAVERSAL] == SPCU_TRAVERSAL_ORDERED && !d_cnt, q[kQWalk], new_count(),
                                          d_counters, d_cnt, pending_advance ? &adv : nullptr);
                pending_advance = false;
                timer.end();
                uint32_t* n_live   = new_count();
                uint32_t* n_shadow = d_counts + next_count; // one shadow queue (and counter) per light
                next_count += n_lights;
                uint32_t* q_shadow = (nee || direct) ? q[kQShadow] : nullptr;
                timer.begin(kStShade);
                launch_shade(L, s, lane.wave, p, sorted, max_n, q[kQLive], n_live, q_shadow, n_shadow, d_counters);
                timer.end();
                launches += 1;
                if (nee || direct) {
                    for (uint32_t li = 0; li < n_lights; ++li) {
                        p.light_index           = li;
                        uint32_t* q_shadow_l    = q_shadow + static_cast<size_t>(li) * capacity;
                        uint32_t* n_lit         = direct ? nullptr : new_count();
                        timer.begin(kStShadow);
                        uint32_t* cursor_l = new_count();
                        launches += launch_shadow(L, s, lane.wave, q_shadow_l, n_shadow + li, max_n, li, cursor_l,
                                                  direct ? nullptr : q[kQLit], n_lit, q[kQWalk], new_count(), d_counters, d_cnt);
                        timer.end();
                        if (direct) {
                            timer.begin(kStDirectAccumulate);
                            launch_direct_accumulate(L, s, lane.wave, p, q_shadow_l, n_shadow + li, max_n, d_counters);
                            timer.end();
                            ++launches;
                            continue;
                        }
                        uint32_t* n_mis = new_count();
                        timer.begin(kStNeeBsdf);
                        launch_nee_bsdf(L, s, lane.wave, p, q[kQLit], n_lit, max_n, q[kQMis], n_mis, d_counters);
                        timer.end();
                        timer.begin(kStMisTrace);
                        uint32_t* cursor_m = new_count();
                        launches += launch_mis_trace(L, s, lane.wave, q[kQMis], n_mis, max_n, cursor_m, q[kQWalk], new_count(), d_counters, d_cnt);
                        timer.end();
                        timer.begin(kStNeeMisAccumulate);
                        launch_nee_mis_accumulate(L, s, lane.wave, q[kQMis], n_mis, max_n, d_counters);
                        timer.end();
                        launches += 2;
                    }
                }
                if (direct && !whitted) {
                    break;
                }
                if (fuse_advance) {
                    // the live queue is read by the next depth's extend (and rewritten only by the shade stage after it):
                    // two buffers in turn, so that this depth's q_live is intact while the next one's is being filled
                    q_cur           = q[kQLive];
                    n_cur           = n_live;
                    pending_advance = true;
                    std::swap(q[kQLive], q[kQNext]);
                    continue;
                }
                uint32_t* n_next = new_count();
                timer.begin(kStAdvance);
                if (whitted) {
                    launch_whitted_advance(L, s, lane.wave, p, q[kQLive], n_live, max_n, q_next, n_next, d_counters);
                } else {
                    launch_advance(L, s, lane.wave, p, q[kQLive], n_live, max_n, q_next, n_next, d_counters);
                }
                timer.end();
                ++launches;
                std::swap(q_cur, q_next);
                n_cur = n_next;
            }
            // batches add their samples to the accumulators in batch order, whichever lane ran them: this resolve waits for the
            // previous batch's (the float sums, and so the image, do not depend on the number of lanes)
            if (lanes.size() > 1 && batch > 0) {
                CK(c, cudaStreamWaitEvent(lane.st, lanes[(batch - 1) % lanes.size()].resolved, 0));
            }
            timer.begin(kStResolve);
            launch_resolve(L, &lane.wave.path->L, 2, d_pix_list + pb, np, ns, d_rgb_sum, d_lum_sumsq, d_counters);
            timer.end();
            ++launches;
            if (lanes.size() > 1) {
                CK(c, cudaEventRecord(lane.resolved, lane.st));
            }
            ++batch;
            CK(c, cudaGetLastError());
        }
    }
    for (size_t k = 1; k < lanes.size(); ++k) { // join: the call's stream ends after every lane's last resolve
        CK(c, cudaStreamWaitEvent(st, lanes[k].resolved, 0));
    }
    CK(c, cudaEventRecord(c->ev1, st));

    if (stats) {
        if (int rc = read_back_stats(c, st, stats, launches, timer); rc != SPCU_OK) return rc;
    }
    return SPCU_OK;
}

// The host-buffer entry points synchronise anyway: they also read the fault counter (kCntErrors) when no stats were asked
// for, so a kernel that gave up on a bounded wait fails the call instead of returning a void image.
int sync_and_check_faults(spcu_ctx* c, bool already_checked)
{
    unsigned long long faults = 0;
    if (!already_checked && c->counters.p) {
        CK(c, cudaMemcpyAsync(&faults, c->counters.as<unsigned long long>() + kCntErrors, sizeof faults, cudaMemcpyDeviceToHost,
                              c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    if (faults) {
        return fail(c, SPCU_ERR_INTERNAL, "%llu kernel(s) gave up on a bounded wait (queue protocol fault); the image is void", faults);
    }
    return SPCU_OK;
}

} // namespace

extern "C" {

int spcu_render_device(spcu_ctx* c, const spcu_partition* part, float* d_rgb_sum, float* d_lum_sumsq, spcu_stats* stats,
                       void* stream)
{
    // stream == NULL means the legacy default stream, as in the CUDA runtime
    return render_impl(c, part, d_rgb_sum, d_lum_sumsq, stats, static_cast<cudaStream_t>(stream), true);
}

int spcu_render(spcu_ctx* c, const spcu_partition* part, float* rgb_sum, float* lum_sumsq, spcu_stats* stats)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    if (!rgb_sum) {
        return fail(c, SPCU_ERR_INVALID, "rgb_sum is NULL");
    }
    const size_t n_pixels = static_cast<size_t>(c->ds.width) * c->ds.height;
    CK(c, c->host_rgb.reserve(n_pixels * 3 * sizeof(float)));
    CK(c, cudaMemcpyAsync(c->host_rgb.p, rgb_sum, n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    float* d_sq = nullptr;
    if (lum_sumsq) {
        CK(c, c->host_sq.reserve(n_pixels * sizeof(float)));
        CK(c, cudaMemcpyAsync(c->host_sq.p, lum_sumsq, n_pixels * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        d_sq = c->host_sq.as<float>();
    }
    if (int rc = render_impl(c, part, c->host_rgb.as<float>(), d_sq, stats, nullptr, false); rc != SPCU_OK) return rc;
    CK(c, cudaMemcpyAsync(rgb_sum, c->host_rgb.p, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (lum_sumsq) {
        CK(c, cudaMemcpyAsync(lum_sumsq, c->host_sq.p, n_pixels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    return sync_and_check_faults(c, stats != nullptr);
}

int spcu_resolved_pipeline(const spcu_ctx* c)
{
    return (c && c->have_scene) ? static_cast<int>(resolve_pipeline(c)) : -1;
}

int spcu_render_frame(spcu_ctx* c, const spcu_partition* part, float* rgb_sum, float* lum_sumsq, spcu_stats* stats)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    if (!rgb_sum) {
        return fail(c, SPCU_ERR_INVALID, "rgb_sum is NULL");
    }
    const size_t n_pixels = static_cast<size_t>(c->ds.width) * c->ds.height;
    CK(c, c->host_rgb.reserve(n_pixels * 3 * sizeof(float)));
    CK(c, cudaMemsetAsync(c->host_rgb.p, 0, n_pixels * 3 * sizeof(float), c->stream));
    float* d_sq = nullptr;
    if (lum_sumsq) {
        CK(c, c->host_sq.reserve(n_pixels * sizeof(float)));
        CK(c, cudaMemsetAsync(c->host_sq.p, 0, n_pixels * sizeof(float), c->stream));
        d_sq = c->host_sq.as<float>();
    }
    if (int rc = render_impl(c, part, c->host_rgb.as<float>(), d_sq, stats, nullptr, false); rc != SPCU_OK) return rc;
    CK(c, cudaMemcpyAsync(rgb_sum, c->host_rgb.p, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (lum_sumsq) {
        CK(c, cudaMemcpyAsync(lum_sumsq, c->host_sq.p, n_pixels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    return sync_and_check_faults(c, stats != nullptr);
}

int spcu_render_frame_reduced(spcu_ctx* c, const spcu_partition* part, int with_sumsq, float* rgb_sum, float* lum_sumsq,
                              spcu_stats* stats)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    const bool root = c->comm_rank == 0;
    if (root && (!rgb_sum || (with_sumsq && !lum_sumsq))) {
        return fail(c, SPCU_ERR_INVALID, "rank 0 needs the host buffers");
    }
    const size_t n_pixels = static_cast<size_t>(c->ds.width) * c->ds.height;
    CK(c, c->host_rgb.reserve(n_pixels * 3 * sizeof(float)));
    CK(c, cudaMemsetAsync(c->host_rgb.p, 0, n_pixels * 3 * sizeof(float), c->stream));
    float* d_sq = nullptr;
    if (with_sumsq) {
        CK(c, c->host_sq.reserve(n_pixels * sizeof(float)));
        CK(c, cudaMemsetAsync(c->host_sq.p, 0, n_pixels * sizeof(float), c->stream));
        d_sq = c->host_sq.as<float>();
    }
    if (int rc = render_impl(c, part, c->host_rgb.as<float>(), d_sq, stats, nullptr, false); rc != SPCU_OK) return rc;
    if (int rc = spcu_reduce_to_root(c, c->host_rgb.as<float>(), d_sq, c->stream); rc != SPCU_OK) return rc;
    if (root) {
        CK(c, cudaMemcpyAsync(rgb_sum, c->host_rgb.p, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (with_sumsq) {
            CK(c, cudaMemcpyAsync(lum_sumsq, c->host_sq.p, n_pixels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    return sync_and_check_faults(c, stats != nullptr);
}

int spcu_render_image(spcu_ctx* c, const spcu_partition* part, uint32_t format, void* out, spcu_stats* stats)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    if (!part || part->sample_end <= part->sample_begin) {
        return fail(c, SPCU_ERR_INVALID, "empty sample range");
    }
    const size_t n_pixels = static_cast<size_t>(c->ds.width) * c->ds.height;
    CK(c, c->host_rgb.reserve(n_pixels * 3 * sizeof(float)));
    CK(c, cudaMemsetAsync(c->host_rgb.p, 0, n_pixels * 3 * sizeof(float), c->stream));
    if (int rc = render_impl(c, part, c->host_rgb.as<float>(), nullptr, stats, nullptr, false); rc != SPCU_OK) return rc;
    if (int rc = sync_and_check_faults(c, stats != nullptr); rc != SPCU_OK) return rc;
    return pack_device_image(c, c->host_rgb.as<float>(), c->ds.width, c->ds.height, part->sample_end - part->sample_begin, format, out);
}

// Ray batches through the renderer's own traversal stages: the kernels a frame runs (begin + persistent walk, lane refill,
// pair leaf steps), not the one-thread-per-ray kernels behind spcu_trace_*.
static int stage_batch(spcu_ctx* c, const spcu_ray* rays, uint64_t n, uint32_t traversal, spcu_hit* hits, spcu_hit* light_hits,
                       uint8_t* occluded)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    const bool shadow = occluded != nullptr;
    if (n && (!rays || (!shadow && !hits))) {
        return fail(c, SPCU_ERR_INVALID, "NULL ray or result buffer");
    }
    if (traversal > SPCU_TRAVERSAL_ORDERED) {
        return fail(c, SPCU_ERR_INVALID, "unknown traversal %u", traversal);
    }
    const cudaStream_t st = c->stream;
    if (c->scratch_in_use && c->scratch_stream != st) {
        CK(c, cudaStreamWaitEvent(st, c->ev1, 0));
    }
    const uint32_t chunk = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(n, 1), kBatchRays));
    // the batch borrows the render's wavefront state: keep its capacity when it is already large enough (capacity is also the
    // stride of the per-light planes, and shadow rays use plane 0)
    const uint32_t capacity = (c->wave.capacity >= chunk && c->wave_lights == c->ds.n_lights) ? c->wave.capacity : chunk;
    if (int rc = ensure_wave(c, capacity); rc != SPCU_OK) return rc;
    const uint32_t n_segments = std::min<uint32_t>(c->n_materials, kMaxMaterialSegments) + 1u;
    CK(c, c->sorted_queue.reserve(static_cast<size_t>(n_segments) * capacity * sizeof(uint32_t)));
    CK(c, c->q_rays.reserve(static_cast<size_t>(chunk) * sizeof(spcu_ray)));
    CK(c, c->q_out.reserve(static_cast<size_t>(chunk) * sizeof(spcu_hit)));
    CK(c, c->q_aux.reserve(static_cast<size_t>(chunk) * sizeof(spcu_hit)));
    auto*        d_counters = c->counters.as<unsigned long long>();
    uint32_t*    d_counts   = c->queue_counts.as<uint32_t>();
    const Launch L{ c->sm_count, st, c->features };
    for (uint64_t done = 0; done < n; done += chunk) {
        const uint32_t m = static_cast<uint32_t>(std::min<uint64_t>(chunk, n - done));
        CK(c, cudaMemcpyAsync(c->q_rays.p, rays + done, static_cast<size_t>(m) * sizeof(spcu_ray), cudaMemcpyHostToDevice, st));
        CK(c, cudaMemsetAsync(d_counts, 0, (8 + n_segments) * sizeof(uint32_t), st));
        CK(c, cudaMemsetAsync(d_counters, 0, kCounterBlock * sizeof(unsigned long long) + sizeof(TraceCounters), st));
        uint32_t* n_cur = d_counts + 0;
        launch_batch_fill(c->wave, c->q_rays.as<spcu_ray>(), m, c->queues[kQCur].as<uint32_t>(), n_cur, shadow, st);
        if (shadow) {
            launch_shadow(L, c->ds, c->wave, c->queues[kQCur].as<uint32_t>(), n_cur, m, 0, d_counts + 1, nullptr, nullptr,
                          c->queues[kQWalk].as<uint32_t>(), d_counts + 2, d_counters, nullptr);
            CK(c, cudaGetLastError());
            CK(c, cudaMemcpyAsync(occluded + done, c->wave.occluded, m, cudaMemcpyDeviceToHost, st));
        } else {
            SortedQueue sorted{ c->sorted_queue.as<uint32_t>(), d_counts + 8, n_segments, capacity };
            launch_extend(L, c->ds, c->wave, c->queues[kQCur].as<uint32_t>(), n_cur, m, d_counts + 1, sorted,
                          traversal == SPCU_TRAVERSAL_ORDERED, c->queues[kQWalk].as<uint32_t>(), d_counts + 2, d_counters, nullptr);
            launch_batch_gather_extend(c->wave, m, c->q_out.as<spcu_hit>(), light_hits ? c->q_aux.as<spcu_hit>() : nullptr, st);
            CK(c, cudaGetLastError());
            CK(c, cudaMemcpyAsync(hits + done, c->q_out.p, static_cast<size_t>(m) * sizeof(spcu_hit), cudaMemcpyDeviceToHost, st));
            if (light_hits) {
                CK(c, cudaMemcpyAsync(light_hits + done, c->q_aux.p, static_cast<size_t>(m) * sizeof(spcu_hit), cudaMemcpyDeviceToHost, st));
            }
        }
        if (int rc = sync_and_check_faults(c, false); rc != SPCU_OK) return rc;
    }
    return SPCU_OK;
}

int spcu_extend_batch(spcu_ctx* c, const spcu_ray* rays, uint64_t n, uint32_t traversal, spcu_hit* hits, spcu_hit* light_hits)
{
    return stage_batch(c, rays, n, traversal, hits, light_hits, nullptr);
}

int spcu_shadow_batch(spcu_ctx* c, const spcu_ray* rays, uint64_t n, uint8_t* occluded)
{
    if (n && !occluded) {
        return fail(c, SPCU_ERR_INVALID, "NULL result buffer");
    }
    uint8_t dummy = 0;
    return stage_batch(c, rays, n, SPCU_TRAVERSAL_EXACT, nullptr, nullptr, occluded ? occluded : &dummy);
}

int spcu_stage_times(spcu_ctx* c, spcu_stage_time* out, uint32_t capacity, uint32_t* n_out)
{
    if (!c || !out || !n_out) {
        return fail(c, SPCU_ERR_INVALID, "NULL argument");
    }
    const uint32_t n = std::min<uint32_t>(capacity, kNumStages);
    std::memcpy(out, c->stage_report, n * sizeof(spcu_stage_time));
    *n_out = n;
    return SPCU_OK;
}

} // extern "C"
