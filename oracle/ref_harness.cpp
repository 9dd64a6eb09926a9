// TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
// baseline legs may load the library built from this file (oracle/_ref/libsp_ref*.so).
//
// Drives the REAL reference (compiled from /root/reference by oracle/Makefile) to produce ground truth:
//   * spref_trace_*    : Scene::intersect / intersect_p / intersect_lights on ray batches, with primitive IDs.
//                        The reference's Intersection carries no ID (shapes/Intersection.h:17-23), so the walk
//                        below re-enacts ListAccelerator::intersect_impl (shapes/ListAccelerator.h:36-67) and
//                        BVHAccelerator::NodeInternal::intersect (shapes/BVHAccelerator.h:45-90) by calling the
//                        reference's own sp::intersect_p(bounds, ...) and Hitable::intersect(...) on the
//                        reference's own objects, recording which primitive last updated the result, and
//                        cross-checks the distance bit-for-bit against Scene::intersect().
//   * spref_render     : main.cpp:77-142's render loop with the reference's integrators and samplers, plus a
//                        per-pixel RunningStats (base/RunningStats.h) of sample luminance.
//   * spref_flat       : the product's flattener applied to the very same in-memory Scene.
// Built with -fno-access-control (the reference keeps all of this private).

#include "flat_scene.h"

#include "base/FileParser.h"
#include "base/PlyReader.h"
#include "base/STLReader.h"
#include "base/Logger.h"
#include "base/MemoryArena.h"
#include "base/RunningStats.h"
#include "base/Scene.h"
#include "base/Tile.h"
#include "base/TileScheduler.h"
#include "Cameras/Camera.h"
#include "Image/Image.h"
#include "Integrators/Integrator.h"
#include "math/Sampler.h"
#include "shapes/BVHAccelerator.h"
#include "shapes/ListAccelerator.h"
#include "shapes/Primitive.h"
#include "shapes/Triangle.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <functional>
#include <fstream>
#include <memory>
#include <ranges>
#include <thread>
#include <unistd.h>
#include <unordered_map>

namespace sp {
int k_pretty_print_key = -1; // defined in the reference's main.cpp:33, which is not linked here
}

namespace {

using sp::BVHAccelerator;
using sp::Hitable;

// The reference's AccumulatedLogger singleton joins a worker thread from a static destructor and never
// returns.  Once a scene has been parsed (which creates the singleton) we arrange for the
// process to leave through _exit with its real status, after flushing stdio.
void leave_now(int status, void*)
{
    std::fflush(nullptr);
    _exit(status);
}

void arm_exit_guard()
{
    static std::once_flag once;
    std::call_once(once, [] { on_exit(&leave_now, nullptr); });
}

struct IdMap
{
    std::unordered_map<const Hitable*, int32_t> id;
    int32_t                                     next = 0;

    void add_leaf_dfs(const BVHAccelerator::NodeBase* n)
    {
        if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(n)) {
            for (const auto& p : leaf->m_primitives.m_primitives) {
                id.emplace(p.get(), next++);
            }
            return;
        }
        const auto* inner = static_cast<const BVHAccelerator::NodeInternal*>(n);
        add_leaf_dfs(inner->m_children[0].get());
        add_leaf_dfs(inner->m_children[1].get());
    }

    void build(const sp::ListAccelerator& top)
    {
        for (const auto& p : top.m_primitives) {
            if (const auto* bvh = dynamic_cast<const BVHAccelerator*>(p.get())) {
                add_leaf_dfs(bvh->m_root.get());
            } else {
                id.emplace(p.get(), next++);
            }
        }
    }
};

struct Counters
{
    uint64_t nodes = 0; // NodeInternal visits
    uint64_t tris  = 0; // triangle tests
    uint64_t xf    = 0; // sphere / plane tests
};

struct Closest
{
    int32_t id = -1;
    float   t  = 0.0f;
};

// ---- geometry: closest -------------------------------------------------------------------------
struct GeomTracer
{
    const IdMap& ids;
    Counters*    cnt;

    bool leaf_prim(const Hitable* p, const sp::Ray& ray, sp::RayLimits& limits, Closest& out) const
    {
        if (cnt) {
            const auto* gp = static_cast<const sp::GeometricPrimitive*>(p);
            if (dynamic_cast<const sp::Triangle*>(gp->m_shape.get().get())) {
                ++cnt->tris;
            } else {
                ++cnt->xf;
            }
        }
        if (const auto hit = p->intersect(ray, limits); hit) {
            limits.m_t_max = hit->m_distance;
            out.id         = ids.id.at(p);
            out.t          = hit->m_distance;
            return true;
        }
        return false;
    }

    // Mirrors NodeInternal::intersect / NodeLeaf::intersect; `limits` is the caller's copy semantics:
    // every level works on its own copy and only the hit distance propagates upward.
    bool node(const BVHAccelerator::NodeBase* n, const sp::Ray& ray, const sp::RayLimits& limits_in, Closest& out) const
    {
        bool any = false;
        auto limits{ limits_in };
        if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(n)) {
            for (const auto& p : leaf->m_primitives.m_primitives) {
                any |= leaf_prim(p.get(), ray, limits, out);
            }
            return any;
        }
        const auto* inner = static_cast<const BVHAccelerator::NodeInternal*>(n);
        if (cnt) {
            ++cnt->nodes;
        }
        for (const auto& child : inner->m_children) {
            if (sp::intersect_p(child->m_bounds, ray, limits)) {
                Closest sub;
                if (node(child.get(), ray, limits, sub)) {
                    limits.m_t_max = sub.t;
                    out            = sub;
                    any            = true;
                }
            }
        }
        return any;
    }

    Closest top(const sp::ListAccelerator& list, const sp::Ray& ray, const sp::RayLimits& limits_in) const
    {
        Closest out;
        out.t = limits_in.m_t_max;
        auto limits{ limits_in };
        for (const auto& p : list.m_primitives) {
            if (const auto* bvh = dynamic_cast<const BVHAccelerator*>(p.get())) {
                Closest sub;
                if (node(bvh->m_root.get(), ray, limits, sub)) {
                    limits.m_t_max = sub.t;
                    out            = sub;
                }
            } else {
                leaf_prim(p.get(), ray, limits, out);
            }
        }
        return out;
    }
};

// ---- lights: closest ---------------------------------------------------------------------------
struct LightTracer
{
    const IdMap& ids;

    bool prim(const Hitable* p, const sp::Ray& ray, sp::RayLimits& limits, Closest& out) const
    {
        if (const auto hit = p->intersect_lights(ray, limits); hit) {
            limits.m_t_max = hit->m_distance;
            out.id         = ids.id.at(p);
            out.t          = hit->m_distance;
            return true;
        }
        return false;
    }

    bool node(const BVHAccelerator::NodeBase* n, const sp::Ray& ray, const sp::RayLimits& limits_in, Closest& out) const
    {
        bool any = false;
        auto limits{ limits_in };
        if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(n)) {
            for (const auto& p : leaf->m_primitives.m_primitives) {
                any |= prim(p.get(), ray, limits, out);
            }
            return any;
        }
        const auto* inner = static_cast<const BVHAccelerator::NodeInternal*>(n);
        for (const auto& child : inner->m_children) {
            if (sp::intersect_p(child->m_bounds, ray, limits)) {
                Closest sub;
                if (node(child.get(), ray, limits, sub)) {
                    limits.m_t_max = sub.t;
                    out            = sub;
                    any            = true;
                }
            }
        }
        return any;
    }

    Closest top(const sp::ListAccelerator& list, const sp::Ray& ray, const sp::RayLimits& limits_in) const
    {
        Closest out;
        out.t = limits_in.m_t_max;
        auto limits{ limits_in };
        for (const auto& p : list.m_primitives) {
            if (const auto* bvh = dynamic_cast<const BVHAccelerator*>(p.get())) {
                Closest sub;
                if (node(bvh->m_root.get(), ray, limits, sub)) {
                    limits.m_t_max = sub.t;
                    out            = sub;
                }
            } else {
                prim(p.get(), ray, limits, out);
            }
        }
        return out;
    }
};

sp::Ray to_ray(const spcu_ray& r)
{
    return sp::Ray{ sp::Point3{ r.ox, r.oy, r.oz }, sp::Vector3{ r.dx, r.dy, r.dz } };
}

sp::RayLimits to_limits(const spcu_ray& r)
{
    return sp::RayLimits{ .m_t_min = r.t_min, .m_t_max = r.t_max };
}

uint32_t float_bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

std::unique_ptr<sp::Integrator> make_integrator(const std::string& name)
{
    if (name == "iterative_rrnee") return std::make_unique<sp::IntegratorIterativeRRNEE>();
    if (name == "brute_force_iterative_rr") return std::make_unique<sp::BruteForceIntegratorIterativeRR>();
    if (name == "brute_force_iterative") return std::make_unique<sp::BruteForceIntegratorIterative>();
    if (name == "direct_lighting") return std::make_unique<sp::DirectLightingIntegrator>();
    if (name == "whitted") return std::make_unique<sp::WhittedIntegrator>();
    return nullptr;
}

} // namespace

struct spref_scene
{
    std::unique_ptr<sp::Scene> scene;
    IdMap                      geom_ids;
    IdMap                      light_ids;
    spb200::FlatScene          flat;
    bool                       have_flat = false;
};

template <typename F>
static int guarded(char* err, size_t errlen, F&& f)
{
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen) {
            std::snprintf(err, errlen, "%s", e.what());
        }
        return -1;
    } catch (...) {
        if (err && errlen) {
            std::snprintf(err, errlen, "unknown exception");
        }
        return -1;
    }
}


void put3f(float* dst, const auto& v)
{
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
}

// ---- BVH construction by the reference's own BVHAccelerator (shapes/BVHAccelerator.h:123-209) ---------------------------
// A bounded Hitable that is nothing but its world bounds: lets the reference build a tree over arbitrary boxes.
struct BoxHitable final : Hitable
{
    sp::BBox3 box;
    uint32_t  index;
    bool      non_triangle;

    std::optional<sp::LightIntersection> intersect_lights_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return {}; }
    std::optional<sp::Intersection>      intersect_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return {}; }
    bool                                 intersect_p_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return false; }
    sp::BBox3                            get_world_bounds_impl() const noexcept override { return box; }
    bool                                 is_bounded_impl() const noexcept override { return true; }
};

// An unbounded primitive that is nothing but that (a plane's stand-in for the partition of base/Scene.h:33).
struct UnboundedStub final : Hitable
{
    std::optional<sp::LightIntersection> intersect_lights_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return {}; }
    std::optional<sp::Intersection>      intersect_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return {}; }
    bool                                 intersect_p_impl(const sp::Ray&, const sp::RayLimits&) const noexcept override { return false; }
    sp::BBox3                            get_world_bounds_impl() const noexcept override { return {}; }
    bool                                 is_bounded_impl() const noexcept override { return false; }
};

struct BuildWalker
{
    spcu_bvh_node* nodes;
    uint32_t       capacity;
    uint32_t       n_nodes   = 0;
    uint32_t       n_prims   = 0;
    uint32_t       max_depth = 0;
    uint32_t       first_id;
    uint32_t*      order;

    int32_t walk(const BVHAccelerator::NodeBase* node, uint32_t& count_word, uint32_t depth)
    {
        if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(node)) {
            const uint32_t first = n_prims;
            bool           mixed = false;
            for (const auto& p : leaf->m_primitives.m_primitives) {
                const auto* b    = static_cast<const BoxHitable*>(p.get());
                order[n_prims++] = b->index;
                mixed |= b->non_triangle;
            }
            count_word = (n_prims - first) | (mixed ? SPCU_LEAF_MIXED_FLAG : 0u);
            return ~static_cast<int32_t>(first_id + first);
        }
        const auto* inner = static_cast<const BVHAccelerator::NodeInternal*>(node);
        max_depth         = std::max(max_depth, depth + 1);
        const uint32_t idx = n_nodes++;
        if (idx >= capacity) {
            throw std::runtime_error("spref_build_bvh: node capacity exceeded");
        }
        spcu_bvh_node n{};
        for (int k = 0; k < 2; ++k) {
            const auto& b = inner->m_children[k]->m_bounds;
            const float v[6] = { b.get_lower().x, b.get_lower().y, b.get_lower().z, b.get_upper().x, b.get_upper().y, b.get_upper().z };
            std::memcpy(n.box + 6 * k, v, sizeof v);
            n.child[k] = walk(inner->m_children[k].get(), n.count[k], depth + 1);
        }
        nodes[idx] = n;
        count_word = 0;
        return static_cast<int32_t>(idx);
    }
};

extern "C" {

// Parse a .sp file with the reference's FileParser (base/FileParser.cpp:928-932).  Relative asset paths in
// the file are resolved against the file's directory, like running the reference from there.
spref_scene* spref_load(const char* sp_path, char* err, size_t errlen)
{
    spref_scene* out = nullptr;
    guarded(err, errlen, [&] {
        namespace fs = std::filesystem;
        sp::Logger::set_level(sp::Logger::LoggingLevel::error);
        const fs::path path = fs::absolute(sp_path);
        std::ifstream  ins(path);
        if (!ins) {
            throw std::runtime_error("cannot open " + path.string());
        }
        const fs::path old = fs::current_path();
        fs::current_path(path.parent_path());
        auto s = std::make_unique<spref_scene>();
        try {
            s->scene = std::make_unique<sp::Scene>(sp::parse_file(ins));
        } catch (...) {
            fs::current_path(old);
            arm_exit_guard();
            throw;
        }
        fs::current_path(old);
        arm_exit_guard();
        s->geom_ids.build(s->scene->m_accelerator_geometry);
        s->light_ids.build(s->scene->m_accelerator_lights);
        out = s.release();
    });
    return out;
}

void spref_free(spref_scene* s)
{
    delete s;
}

const spcu_flat_scene* spref_flat(spref_scene* s, char* err, size_t errlen)
{
    if (!s->have_flat) {
        if (guarded(err, errlen, [&] { s->flat = spb200::flatten_scene(*s->scene); }) != 0) {
            return nullptr;
        }
        s->have_flat = true;
    }
    return &s->flat.view;
}

int spref_save_flat(spref_scene* s, const char* path, char* err, size_t errlen)
{
    if (!spref_flat(s, err, errlen)) {
        return -1;
    }
    return guarded(err, errlen, [&] { spb200::save_flat_scene(s->flat, path); });
}

void spref_jitter(unsigned spp, float* out)
{
    const auto t = spb200::jitter_table(spp);
    std::memcpy(out, t.data(), t.size() * sizeof(float));
}

// Scene::intersect with IDs.  counters (may be NULL) = {internal nodes visited, triangle tests, sphere/plane
// tests} summed over the batch.  Returns the number of rays whose walked distance differs (bitwise) from
// Scene::intersect's — must be 0.
int64_t spref_trace_closest(spref_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits, uint64_t* counters)
{
    Counters         cnt;
    const GeomTracer tracer{ s->geom_ids, counters ? &cnt : nullptr };
    int64_t          bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const auto ray    = to_ray(rays[i]);
        const auto limits = to_limits(rays[i]);
        const auto c      = tracer.top(s->scene->m_accelerator_geometry, ray, limits);
        hits[i].id        = c.id;
        hits[i].t         = c.t;
        const auto ref    = s->scene->intersect(ray, limits);
        if (static_cast<bool>(ref) != (c.id >= 0) || (ref && float_bits(ref->m_distance) != float_bits(c.t))) {
            ++bad;
        }
    }
    if (counters) {
        counters[0] = cnt.nodes;
        counters[1] = cnt.tris;
        counters[2] = cnt.xf;
    }
    return bad;
}

// Scene::intersect_p (geometry || lights) straight from the reference.
void spref_trace_any(spref_scene* s, const spcu_ray* rays, uint64_t n, uint8_t* out)
{
    for (uint64_t i = 0; i < n; ++i) {
        out[i] = s->scene->intersect_p(to_ray(rays[i]), to_limits(rays[i])) ? 1 : 0;
    }
}

// Scene::intersect_lights with light IDs (accelerator order).  Returns the cross-check mismatch count.
int64_t spref_trace_lights(spref_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits)
{
    const LightTracer tracer{ s->light_ids };
    int64_t           bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const auto ray    = to_ray(rays[i]);
        const auto limits = to_limits(rays[i]);
        const auto c      = tracer.top(s->scene->m_accelerator_lights, ray, limits);
        hits[i].id        = c.id;
        hits[i].t         = c.t;
        const auto ref    = s->scene->intersect_lights(ray, limits);
        if (static_cast<bool>(ref) != (c.id >= 0) || (ref && float_bits(ref->m_distance) != float_bits(c.t))) {
            ++bad;
        }
    }
    return bad;
}

// Camera rays as render_thread makes them (main.cpp:94-98): pixel index = y*width + x, sample index into the
// R-sequence.  t_min/t_max get RayLimits' defaults (math/Ray.h:13-19).
void spref_generate_rays(spref_scene* s, const uint32_t* pix, const uint32_t* smp, uint64_t n, unsigned spp, spcu_ray* out)
{
    const auto   jitter = spb200::jitter_table(spp);
    const auto   w      = static_cast<uint32_t>(s->scene->image_width);
    sp::RayLimits lim;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t x = pix[i] % w;
        const uint32_t y = pix[i] / w;
        // p.x + sample.x: int + float (main.cpp:97)
        const float    px  = static_cast<float>(static_cast<int>(x)) + jitter[2 * smp[i] + 0];
        const float    py  = static_cast<float>(static_cast<int>(y)) + jitter[2 * smp[i] + 1];
        const sp::Ray  ray = s->scene->m_camera->generate_ray(px, py);
        out[i] = spcu_ray{ ray.get_origin().x,    ray.get_origin().y,    ray.get_origin().z,    lim.m_t_min,
                           ray.get_direction().x, ray.get_direction().y, ray.get_direction().z, lim.m_t_max };
    }
}

// For each ray: the reference's Intersection record (normal xyz, point xyz) or zeros on a miss.
void spref_hit_records(spref_scene* s, const spcu_ray* rays, uint64_t n, float* normal_point)
{
    for (uint64_t i = 0; i < n; ++i) {
        float* o = normal_point + 6 * i;
        std::memset(o, 0, 6 * sizeof(float));
        if (const auto hit = s->scene->intersect(to_ray(rays[i]), to_limits(rays[i])); hit) {
            o[0] = hit->m_normal.x;
            o[1] = hit->m_normal.y;
            o[2] = hit->m_normal.z;
            o[3] = hit->m_point.x;
            o[4] = hit->m_point.y;
            o[5] = hit->m_point.z;
        }
    }
}

// The reference render (main.cpp:77-142) with its own integrators, samplers, tile scheduler and Morton pixel
// order.  rgb_mean = W*H*3 (row major, y down as in Image(x,y)); lum_mean / lum_var (may be NULL) = per-pixel
// RunningStats<float> of relative_luminance(sample).  Returns wall seconds of the render loop, < 0 on error.
double spref_render(spref_scene* s, const char* integrator_name, unsigned spp, unsigned n_threads, float* rgb_mean,
                    float* lum_mean, float* lum_var)
{
    const auto integrator = make_integrator(integrator_name);
    if (!integrator) {
        return -1.0;
    }
    const sp::Scene& scene = *s->scene;
    const int        w     = scene.image_width;
    const int        h     = scene.image_height;

    sp::ColumnMajorTileScheduler scheduler{ w, h, 1 };
    const auto                   t0 = std::chrono::steady_clock::now();
    {
        std::vector<std::jthread> threads;
        for (unsigned t = 0; t < std::max(1u, n_threads); ++t) {
            threads.emplace_back([&] {
                sp::MemoryArena arena;
                while (auto scheduled = scheduler.get_next_tile()) {
                    const auto& tile    = scheduled->tile;
                    auto        in_tile = [&tile](const sp::Point2i& p) noexcept { return contains(tile, p); };
                    for (auto p : std::views::all(tile) | std::views::filter(in_tile)) {
                        const auto ux = static_cast<std::uint32_t>(p.x);
                        const auto uy = static_cast<std::uint32_t>(p.y);
                        auto pixel_sampler      = sp::RSequenceSampler::create_new_sequence(sp::Seed{ ux << 16u | uy });
                        auto integrator_sampler = sp::IncoherentSampler::create_new_sequence(sp::Seed{ (ux << 16u | uy) ^ 0xb0ae9d99 });
                        sp::RGB                 sum = sp::RGB::black();
                        sp::RunningStats<float> stats;
                        for (unsigned i = 0; i < spp; ++i) {
                            arena.release_all();
                            const auto       sample = pixel_sampler.get_next_2D();
                            const sp::Point2 pixel_coords{ p.x + sample.x, p.y + sample.y };
                            const sp::Ray    ray = scene.m_camera->generate_ray(pixel_coords.x, pixel_coords.y);
                            const sp::RGB    L   = integrator->integrate(ray, scene, arena, integrator_sampler, pixel_coords);
                            sum += L;
                            stats.push(sp::relative_luminance(L));
                        }
                        sum /= static_cast<float>(spp);
                        const size_t idx     = static_cast<size_t>(p.y) * w + p.x;
                        rgb_mean[3 * idx + 0] = sum.r;
                        rgb_mean[3 * idx + 1] = sum.g;
                        rgb_mean[3 * idx + 2] = sum.b;
                        if (lum_mean) lum_mean[idx] = stats.mean();
                        if (lum_var) lum_var[idx] = stats.variance();
                    }
                }
            });
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// BVHAccelerator(first, last) (shapes/BVHAccelerator.h:123-129,175-209) over n boxes given in INITIAL order
// (bounds = n x {lo.xyz, hi.xyz}), flattened exactly like the product's flattener numbers a tree: internal nodes in
// depth-first pre-order, leaf link = ~(first_id + leaf position), order[k] = initial index of the primitive at leaf
// position k.  root_bounds (6 floats, may be NULL) = BVHAccelerator::get_world_bounds().  Returns 0, -1 on error.
int spref_build_bvh(const float* bounds, uint32_t n, const uint8_t* non_triangle, uint32_t first_id, uint32_t* order,
                    spcu_bvh_node* nodes, uint32_t capacity, spcu_accel* accel, float* root_bounds, char* err, size_t errlen)
{
    return guarded(err, errlen, [&] {
        std::vector<std::shared_ptr<const Hitable>> prims;
        prims.reserve(n);
        for (uint32_t i = 0; i < n; ++i) {
            auto         b = std::make_shared<BoxHitable>();
            const float* f = bounds + 6 * i;
            // set the members directly: BBox(a, b) would re-sort the corners (math/BBox.h:26-30)
            b->box.m_min    = sp::Point3{ f[0], f[1], f[2] };
            b->box.m_max    = sp::Point3{ f[3], f[4], f[5] };
            b->index        = i;
            b->non_triangle = non_triangle && non_triangle[i];
            prims.push_back(std::move(b));
        }
        const BVHAccelerator bvh(prims.begin(), prims.end());
        BuildWalker          w{ nodes, capacity, 0, 0, 0, first_id, order };
        spcu_accel           a{};
        a.n_unbounded = first_id;
        a.root        = w.walk(bvh.m_root.get(), a.root_count, 0);
        a.n_prims     = first_id + w.n_prims;
        a.n_nodes     = w.n_nodes;
        a.max_depth   = w.max_depth;
        a.nodes       = nodes;
        *accel        = a;
        if (root_bounds) {
            const auto b = bvh.get_world_bounds();
            const float v[6] = { b.get_lower().x, b.get_lower().y, b.get_lower().z, b.get_upper().x, b.get_upper().y, b.get_upper().z };
            std::memcpy(root_bounds, v, sizeof v);
        }
    });
}

// Hitable::get_world_bounds() (shapes/Hitable.h:32-35) of every geometry primitive, in ID order: out = n_prims x
// {lo.xyz, hi.xyz}; unbounded primitives (planes) get zeros.  Returns the number of primitives.
uint32_t spref_geom_bounds(spref_scene* s, float* out)
{
    for (const auto& [prim, id] : s->geom_ids.id) {
        float* o = out + 6 * static_cast<size_t>(id);
        std::memset(o, 0, 6 * sizeof(float));
        if (prim->is_bounded()) {
            const auto b = prim->get_world_bounds();
            const float v[6] = { b.get_lower().x, b.get_lower().y, b.get_lower().z, b.get_upper().x, b.get_upper().y, b.get_upper().z };
            std::memcpy(o, v, sizeof v);
        }
    }
    return static_cast<uint32_t>(s->geom_ids.id.size());
}

// sp::write (Image/Image.cpp:66-76: write_pfm / write_ppm by extension) of an image whose pixels are per-pixel sums divided
// the way render_thread divides them: image(p.x, p.y) /= num_pixel_samples with an unsigned divisor (main.cpp:100-102).
int spref_write_image(const float* rgb_sum, uint32_t w, uint32_t h, unsigned spp, const char* path, char* err, size_t errlen)
{
    return guarded(err, errlen, [&] {
        sp::Image image(w, h);
        for (uint32_t y = 0; y < h; ++y) {
            for (uint32_t x = 0; x < w; ++x) {
                const float* c = rgb_sum + (static_cast<size_t>(y) * w + x) * 3;
                image(x, y)    = sp::RGB{ c[0], c[1], c[2] };
                image(x, y) /= spp;
            }
        }
        sp::write(path, image);
    });
}

// The reference's own read_ply (base/PlyReader.cpp:326-531) — or read_stl (base/STLReader.cpp:139-156) for a ".stl" path — with object_to_world = the given AffineSpace (12 floats:
// c0.xyz c1.xyz c2.xyz affine.xyz): Mesh::m_vertices / m_normals (world space) and m_indices.  Also returns the normal
// matrix the reference applies, inverse(linear).transposed() (math/LinearSpace3x3.h:163-167), column major.
int spref_read_ply(const char* path, const float xf[12], uint32_t cap_vertices, uint32_t cap_triangles, uint32_t* n_vertices,
                   uint32_t* n_triangles, float* vertices, float* normals, uint32_t* indices, float normal_xf[9], char* err,
                   size_t errlen)
{
    return guarded(err, errlen, [&] {
        sp::Logger::set_level(sp::Logger::LoggingLevel::error);
        const sp::LinearSpace3x3 lin{ sp::Vector3{ xf[0], xf[1], xf[2] }, sp::Vector3{ xf[3], xf[4], xf[5] }, sp::Vector3{ xf[6], xf[7], xf[8] } };
        const sp::AffineSpace    aff{ lin, sp::Vector3{ xf[9], xf[10], xf[11] } };
        const auto               xform = sp::AffineTransformation::compute_inverse(aff);
        const bool               stl   = std::filesystem::path(path).extension() == ".stl"; // FileParser.cpp:577-581
        const sp::Mesh           mesh  = stl ? sp::read_stl(path, xform) : sp::read_ply(path, xform);
        arm_exit_guard();
        const auto nm = aff.get_linear().inverse().transposed();
        const float m9[9] = { nm.col0().x, nm.col0().y, nm.col0().z, nm.col1().x, nm.col1().y, nm.col1().z, nm.col2().x, nm.col2().y, nm.col2().z };
        std::memcpy(normal_xf, m9, sizeof m9);
        *n_vertices  = static_cast<uint32_t>(mesh.m_vertices.size());
        *n_triangles = static_cast<uint32_t>(mesh.get_num_triangles());
        if (*n_vertices > cap_vertices || *n_triangles > cap_triangles) {
            throw std::runtime_error("spref_read_ply: output capacity too small");
        }
        for (uint32_t i = 0; i < *n_vertices; ++i) {
            put3f(vertices + 3 * i, mesh.m_vertices[i]);
            put3f(normals + 3 * i, mesh.m_normals[i]);
        }
        for (size_t i = 0; i < mesh.m_indices.size(); ++i) {
            indices[i] = static_cast<uint32_t>(mesh.m_indices[i]);
        }
    });
}

// internal::create_acceleration_structure (base/Scene.h:27-45) — what Scene's constructor runs on the parser's primitive
// list — over a list made of the REAL triangles of a mesh the reference itself reads (read_ply + Mesh + Triangle, in face
// order, as FileParser appends them: base/FileParser.cpp:589-598) and unbounded stand-ins: unbounded[i] != 0 puts a stand-in
// at list position i, every other position takes the next triangle.  order[k] = list position of the primitive with ID k
// (unbounded list first, then the BVH's leaves), nodes / accel as the flattener numbers them.
int spref_accel_from_mesh(const char* ply, const float xf[12], const uint8_t* unbounded, uint32_t n_total, uint32_t* order,
                          spcu_bvh_node* nodes, uint32_t capacity, spcu_accel* accel, char* err, size_t errlen)
{
    return guarded(err, errlen, [&] {
        sp::Logger::set_level(sp::Logger::LoggingLevel::error);
        const sp::LinearSpace3x3 lin{ sp::Vector3{ xf[0], xf[1], xf[2] }, sp::Vector3{ xf[3], xf[4], xf[5] }, sp::Vector3{ xf[6], xf[7], xf[8] } };
        const sp::AffineSpace    aff{ lin, sp::Vector3{ xf[9], xf[10], xf[11] } };
        auto mesh = std::make_shared<sp::Mesh>(sp::read_ply(ply, sp::AffineTransformation::compute_inverse(aff)));
        arm_exit_guard();
        std::vector<std::shared_ptr<const Hitable>>   list;
        std::unordered_map<const Hitable*, uint32_t> position;
        size_t                                        next_triangle = 0;
        for (uint32_t i = 0; i < n_total; ++i) {
            std::shared_ptr<const Hitable> p;
            if (unbounded[i]) {
                p = std::make_shared<UnboundedStub>();
            } else {
                if (next_triangle >= mesh->get_num_triangles()) {
                    throw std::runtime_error("spref_accel_from_mesh: the list asks for more triangles than the mesh has");
                }
                p = std::make_shared<sp::Triangle>(mesh, next_triangle++);
            }
            position.emplace(p.get(), i);
            list.push_back(std::move(p));
        }
        const sp::ListAccelerator top = sp::internal::create_acceleration_structure(list.begin(), list.end());
        uint32_t                  n_prims = 0, n_nodes = 0, max_depth = 0;
        spcu_accel                a{};
        std::function<int32_t(const BVHAccelerator::NodeBase*, uint32_t&, uint32_t)> walk =
            [&](const BVHAccelerator::NodeBase* node, uint32_t& count_word, uint32_t depth) -> int32_t {
            if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(node)) {
                const uint32_t first = n_prims;
                for (const auto& p : leaf->m_primitives.m_primitives) {
                    order[n_prims++] = position.at(p.get());
                }
                count_word = n_prims - first;
                return ~static_cast<int32_t>(first);
            }
            const auto* inner = static_cast<const BVHAccelerator::NodeInternal*>(node);
            max_depth         = std::max(max_depth, depth + 1);
            const uint32_t idx = n_nodes++;
            if (idx >= capacity) {
                throw std::runtime_error("spref_accel_from_mesh: node capacity exceeded");
            }
            spcu_bvh_node n{};
            for (int k = 0; k < 2; ++k) {
                put3f(n.box + 6 * k, inner->m_children[k]->m_bounds.get_lower());
                put3f(n.box + 6 * k + 3, inner->m_children[k]->m_bounds.get_upper());
                n.child[k] = walk(inner->m_children[k].get(), n.count[k], depth + 1);
            }
            nodes[idx] = n;
            count_word = 0;
            return static_cast<int32_t>(idx);
        };
        for (const auto& p : top.m_primitives) {
            if (const auto* bvh = dynamic_cast<const BVHAccelerator*>(p.get())) {
                a.n_unbounded = n_prims;
                a.root        = walk(bvh->m_root.get(), a.root_count, 0);
            } else {
                order[n_prims++] = position.at(p.get());
            }
        }
        a.n_prims   = n_prims;
        a.n_nodes   = n_nodes;
        a.max_depth = max_depth;
        a.nodes     = nodes;
        *accel      = a;
    });
}

int spref_width(spref_scene* s) { return s->scene->image_width; }
int spref_height(spref_scene* s) { return s->scene->image_height; }
} // extern "C"
