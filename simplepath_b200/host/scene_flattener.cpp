// SceneFlattener: sp::Scene (reference object graph) -> spcu_flat_scene (POD, include/spcu.h).
//
// The reference hides everything the device needs behind `private` (base/Scene.h:101-105,
// shapes/BVHAccelerator.h:21-121,216).  This translation unit is compiled with -fno-access-control and
// walks the objects the reference itself built, so that the device traverses *the same* topology in
// *the same* order (std::partition in base/Scene.h:33 and BVHAccelerator.h:196 is not stable; rebuilding
// would not reproduce it).  IDs are assigned in that order: unbounded list first, then BVH leaves
// left-to-right depth first.

#include "flat_scene.h"

#include "base/Scene.h"
#include "Cameras/Camera.h"
#include "Lights/Light.h"
#include "materials/Material.h"
#include "math/Sampler.h"
#include "shapes/BVHAccelerator.h"
#include "shapes/ListAccelerator.h"
#include "shapes/Plane.h"
#include "shapes/Primitive.h"
#include "shapes/Sphere.h"
#include "shapes/Triangle.h"

#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <unordered_map>

namespace spb200 {
namespace {

using sp::BVHAccelerator;
using sp::Hitable;

void put3(float* dst, const auto& v)
{
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
}

// 12 floats: c0.xyz c1.xyz c2.xyz affine.xyz
void put_affine(float* dst, const sp::AffineSpace& a)
{
    put3(dst + 0, a.get_linear().col0());
    put3(dst + 3, a.get_linear().col1());
    put3(dst + 6, a.get_linear().col2());
    put3(dst + 9, a.get_affine());
}

void put_linear(float* dst, const sp::LinearSpace3x3& l)
{
    put3(dst + 0, l.col0());
    put3(dst + 3, l.col1());
    put3(dst + 6, l.col2());
}

// The matrix LinearSpace3x3::operator()(const Normal3&) builds on every call (math/LinearSpace3x3.h:163-167).
sp::LinearSpace3x3 normal_matrix(const sp::AffineSpace& object_to_world)
{
    return object_to_world.get_linear().inverse().transposed();
}

// ---------------------------------------------------------------------------------------------------
// Generic walk over ListAccelerator[unbounded..., BVHAccelerator] (base/Scene.h:27-45).
// emit(prim) appends the primitive's records and returns true when it is NOT a triangle.
// ---------------------------------------------------------------------------------------------------
struct AccelWalker
{
    std::vector<spcu_bvh_node>&          nodes;
    std::function<bool(const Hitable*)>  emit;
    uint32_t                             n_prims   = 0;
    uint32_t                             max_depth = 0;
    std::vector<spcu_bounds>             unbuilt_bounds; // world bounds of the bounded primitives of an UNBUILT list

    // Returns the link for this subtree and writes the leaf count word.
    int32_t walk(const BVHAccelerator::NodeBase* node, uint32_t& count_word, uint32_t depth)
    {
        if (const auto* leaf = dynamic_cast<const BVHAccelerator::NodeLeaf*>(node)) {
            const uint32_t first = n_prims;
            bool           mixed = false;
            for (const auto& p : leaf->m_primitives.m_primitives) {
                mixed |= emit(p.get());
                ++n_prims;
            }
            const uint32_t count = n_prims - first;
            if (count > SPCU_LEAF_COUNT_MASK) {
                throw std::runtime_error("flatten: leaf too large");
            }
            count_word = count | (mixed ? SPCU_LEAF_MIXED_FLAG : 0u);
            return ~static_cast<int32_t>(first);
        }
        const auto* inner = dynamic_cast<const BVHAccelerator::NodeInternal*>(node);
        if (!inner) {
            throw std::runtime_error("flatten: unknown BVH node type");
        }
        max_depth = std::max(max_depth, depth + 1);
        const auto idx = static_cast<int32_t>(nodes.size());
        nodes.emplace_back();
        spcu_bvh_node n{};
        for (int k = 0; k < 2; ++k) {
            const auto& b = inner->m_children[k]->m_bounds;
            put3(n.box + 6 * k + 0, b.get_lower());
            put3(n.box + 6 * k + 3, b.get_upper());
            n.child[k] = walk(inner->m_children[k].get(), n.count[k], depth + 1);
        }
        nodes[idx] = n;
        count_word = 0;
        return idx;
    }

    spcu_accel run(const sp::ListAccelerator& top)
    {
        spcu_accel            a{};
        const BVHAccelerator* bvh = nullptr;
        for (const auto& p : top.m_primitives) {
            if (const auto* b = dynamic_cast<const BVHAccelerator*>(p.get())) {
                if (bvh) {
                    throw std::runtime_error("flatten: more than one BVH in a top-level list");
                }
                bvh = b;
                continue;
            }
            if (bvh) {
                throw std::runtime_error("flatten: primitive after the BVH in a top-level list");
            }
            if (p->is_bounded()) { // only an unbuilt list holds bounded primitives directly: [unbounded..., bounded...]
                const auto b = p->get_world_bounds();
                spcu_bounds out{};
                put3(out.lo, b.get_lower());
                put3(out.hi, b.get_upper());
                unbuilt_bounds.push_back(out);
            } else if (!unbuilt_bounds.empty()) {
                throw std::runtime_error("flatten: unbounded primitive after a bounded one in an unbuilt list");
            }
            emit(p.get());
            ++n_prims;
        }
        a.n_unbounded = n_prims - static_cast<uint32_t>(unbuilt_bounds.size());
        if (bvh && !unbuilt_bounds.empty()) {
            throw std::runtime_error("flatten: bounded primitives beside a BVH in a top-level list");
        }
        if (bvh) {
            a.root = walk(bvh->m_root.get(), a.root_count, 0);
        } else {
            a.root       = ~static_cast<int32_t>(a.n_unbounded);
            a.root_count = 0;
        }
        a.n_prims   = n_prims;
        a.n_nodes   = static_cast<uint32_t>(nodes.size());
        a.max_depth = max_depth;
        if (a.max_depth > SPCU_MAX_BVH_DEPTH) {
            throw std::runtime_error("flatten: BVH deeper than SPCU_MAX_BVH_DEPTH");
        }
        return a;
    }
};

// ---------------------------------------------------------------------------------------------------
struct MaterialTable
{
    FlatScene&                                     fs;
    std::unordered_map<const sp::Material*, uint32_t> index;

    uint32_t get(const sp::Material* m, unsigned depth = 0)
    {
        if (const auto it = index.find(m); it != index.end()) {
            return it->second;
        }
        spcu_material out{};
        if (const auto* one = dynamic_cast<const sp::OneSampleMaterial*>(m)) {
            out.kind       = SPCU_MAT_ONE_SAMPLE;
            out.n_bxdfs    = static_cast<uint32_t>(one->m_bxdfs.size());
            out.first_bxdf = static_cast<uint32_t>(fs.bxdfs.size());
            if (out.n_bxdfs == 0 || out.n_bxdfs > SPCU_MAX_BXDFS) {
                throw std::runtime_error("flatten: unsupported BxDF count");
            }
            for (const auto& b : one->m_bxdfs) {
                fs.bxdfs.push_back(make_bxdf(b.get()));
            }
        } else if (const auto* coat = dynamic_cast<const sp::ClearcoatMaterial*>(m)) {
            if (depth + 1 >= SPCU_MAX_COAT_DEPTH) {
                throw std::runtime_error("flatten: clearcoat nesting too deep");
            }
            out.kind        = SPCU_MAT_CLEARCOAT;
            out.base        = get(coat->m_base.get(), depth + 1);
            out.ior         = coat->m_ior;
            out.specular[0] = coat->m_specular_color.r;
            out.specular[1] = coat->m_specular_color.g;
            out.specular[2] = coat->m_specular_color.b;
        } else {
            throw std::runtime_error("flatten: unknown Material subclass");
        }
        const auto id = static_cast<uint32_t>(fs.materials.size());
        fs.materials.push_back(out);
        index.emplace(m, id);
        return id;
    }

    static spcu_bxdf make_bxdf(const sp::BRDF* b)
    {
        spcu_bxdf out{};
        if (const auto* l = dynamic_cast<const sp::LambertianBRDF*>(b)) {
            out.kind = SPCU_BXDF_LAMBERT;
            out.r[0] = l->m_albedo.r; // already albedo / pi (materials/Material.h:317)
            out.r[1] = l->m_albedo.g;
            out.r[2] = l->m_albedo.b;
        } else if (const auto* mf = dynamic_cast<const sp::MicrofacetReflection*>(b)) {
            const auto* d = dynamic_cast<const sp::BeckmannDistribution*>(mf->m_distribution.get());
            if (!d) {
                throw std::runtime_error("flatten: unknown MicrofacetDistribution subclass");
            }
            out.kind           = SPCU_BXDF_MICROFACET;
            out.r[0]           = mf->m_r.r;
            out.r[1]           = mf->m_r.g;
            out.r[2]           = mf->m_r.b;
            out.alpha_x        = d->m_alpha_x;
            out.alpha_y        = d->m_alpha_y;
            out.ior            = mf->m_ior;
            out.sample_visible = d->m_sample_visible_area ? 1u : 0u;
        } else if (const auto* s = dynamic_cast<const sp::SpecularReflectionBRDF*>(b)) {
            out.kind = SPCU_BXDF_SPECULAR;
            out.r[0] = s->m_r.r;
            out.r[1] = s->m_r.g;
            out.r[2] = s->m_r.b;
        } else {
            throw std::runtime_error("flatten: unknown BRDF subclass");
        }
        return out;
    }
};

uint64_t pool_append(FlatScene& fs, const std::vector<float>& v)
{
    const uint64_t off = fs.float_pool.size();
    fs.float_pool.insert(fs.float_pool.end(), v.begin(), v.end());
    return off;
}

spcu_light make_light(FlatScene& fs, const sp::Light* light)
{
    spcu_light out{};
    if (const auto* sl = dynamic_cast<const sp::SphereLight*>(light)) {
        out.kind        = SPCU_LIGHT_SPHERE;
        out.radiance[0] = sl->m_radiance.r;
        out.radiance[1] = sl->m_radiance.g;
        out.radiance[2] = sl->m_radiance.b;
        const auto& xf  = sl->m_sphere.m_object_to_world;
        put_affine(out.world_to_object, xf.get_inverse());
        put_affine(out.object_to_world, xf.get_transform());
        put_linear(out.normal_xf, normal_matrix(xf.get_transform()));
    } else if (const auto* el = dynamic_cast<const sp::EnvironmentLight*>(light)) {
        out.kind        = SPCU_LIGHT_ENV_CONST;
        out.radiance[0] = el->m_radiance.r;
        out.radiance[1] = el->m_radiance.g;
        out.radiance[2] = el->m_radiance.b;
    } else if (const auto* il = dynamic_cast<const sp::ImageBasedEnvironmentLight*>(light)) {
        out.kind = SPCU_LIGHT_ENV_IBL;
        put_linear(out.light_to_world, il->m_light_to_world.get_transform());
        put_linear(out.world_to_light, il->m_light_to_world.get_inverse());
        const auto& img = il->m_radiance;
        out.img_w       = static_cast<uint32_t>(img.width());
        out.img_h       = static_cast<uint32_t>(img.height());
        std::vector<float> pix(static_cast<size_t>(out.img_w) * out.img_h * 3);
        for (uint32_t y = 0; y < out.img_h; ++y) {
            for (uint32_t x = 0; x < out.img_w; ++x) {
                const auto& c = img(x, y);
                float*      d = &pix[(static_cast<size_t>(y) * out.img_w + x) * 3];
                d[0]          = c.r;
                d[1]          = c.g;
                d[2]          = c.b;
            }
        }
        out.img_off = pool_append(fs, pix);

        const auto& d2 = il->m_distribution_2d;
        out.nv         = static_cast<uint32_t>(d2.p_conditional.size());
        out.nu         = out.nv ? static_cast<uint32_t>(d2.p_conditional[0].m_function.size()) : 0u;
        std::vector<float> func, cdf, integ;
        func.reserve(static_cast<size_t>(out.nu) * out.nv);
        cdf.reserve(static_cast<size_t>(out.nu + 1) * out.nv);
        for (const auto& c : d2.p_conditional) {
            if (c.m_function.size() != out.nu || c.m_cdf.size() != out.nu + 1u || c.m_min != 0.0f || c.m_max != 1.0f) {
                throw std::runtime_error("flatten: ragged Distribution2D");
            }
            func.insert(func.end(), c.m_function.begin(), c.m_function.end());
            cdf.insert(cdf.end(), c.m_cdf.begin(), c.m_cdf.end());
            integ.push_back(c.m_function_integral);
        }
        out.cond_func_off = pool_append(fs, func);
        out.cond_cdf_off  = pool_append(fs, cdf);
        out.cond_int_off  = pool_append(fs, integ);
        out.marg_func_off = pool_append(fs, d2.p_marginal.m_function);
        out.marg_cdf_off  = pool_append(fs, d2.p_marginal.m_cdf);
        out.marg_integral = d2.p_marginal.m_function_integral;
    } else {
        throw std::runtime_error("flatten: unknown Light subclass");
    }
    return out;
}

} // namespace

void FlatScene::finalize()
{
    view.abi_version        = SPCU_ABI_VERSION;
    view.geom.nodes         = geom_nodes.data();
    view.geom_prims         = geom_prims.data();
    view.geom_shade         = geom_shade.data();
    view.geom_meta          = geom_meta.data();
    view.lights_accel.nodes = light_nodes.data();
    view.n_lights           = static_cast<uint32_t>(lights.size());
    view.lights             = lights.data();
    view.light_order        = light_order.data();
    view.n_materials        = static_cast<uint32_t>(materials.size());
    view.n_bxdfs            = static_cast<uint32_t>(bxdfs.size());
    view.materials          = materials.data();
    view.bxdfs              = bxdfs.data();
    view.n_pool             = float_pool.size();
    view.float_pool         = float_pool.data();
}

FlatScene flatten_scene(const sp::Scene& scene)
{
    FlatScene fs;
    fs.view.width     = static_cast<uint32_t>(scene.image_width);
    fs.view.height    = static_cast<uint32_t>(scene.image_height);
    fs.view.rr_depth  = static_cast<uint32_t>(std::max(scene.russian_roulette_depth, 0));
    fs.view.max_depth = static_cast<uint32_t>(std::max(scene.max_depth, 0));

    const auto* cam = dynamic_cast<const sp::ProjectiveCamera*>(scene.m_camera.get());
    if (!cam) {
        throw std::runtime_error("flatten: scene has no projective camera");
    }
    put_affine(fs.view.camera, cam->m_transform);

    // ---- geometry ----
    MaterialTable materials{ fs, {} };
    AccelWalker   geom{ fs.geom_nodes, [&](const Hitable* h) -> bool {
                         const auto* prim = dynamic_cast<const sp::GeometricPrimitive*>(h);
                         if (!prim) {
                             throw std::runtime_error("flatten: geometry accelerator holds a non-primitive");
                         }
                         const sp::Shape* shape    = prim->m_shape.get().get();
                         const uint32_t   material = materials.get(prim->m_material.get().get());
                         spcu_prim_geom   g{};
                         spcu_prim_shade  s{};
                         uint32_t         kind;
                         if (const auto* tri = dynamic_cast<const sp::Triangle*>(shape)) {
                             kind = SPCU_PRIM_TRIANGLE;
                             for (int k = 0; k < 3; ++k) {
                                 put3(g.v + 4 * k, tri->m_mesh->m_vertices[tri->m_indices[k]]);
                                 put3(s.v + 4 * k, tri->m_mesh->m_normals[tri->m_indices[k]]);
                             }
                         } else if (const auto* xs = dynamic_cast<const sp::TransformableShape*>(shape)) {
                             if (dynamic_cast<const sp::Sphere*>(shape)) {
                                 kind = SPCU_PRIM_SPHERE;
                             } else if (dynamic_cast<const sp::Plane*>(shape)) {
                                 kind = SPCU_PRIM_PLANE;
                             } else {
                                 throw std::runtime_error("flatten: unknown TransformableShape subclass");
                             }
                             put_affine(g.v, xs->m_object_to_world.get_inverse());
                             const auto nm = normal_matrix(xs->m_object_to_world.get_transform());
                             put3(s.v + 0, nm.col0());
                             put3(s.v + 4, nm.col1());
                             put3(s.v + 8, nm.col2());
                         } else {
                             throw std::runtime_error("flatten: unknown Shape subclass");
                         }
                         fs.geom_prims.push_back(g);
                         fs.geom_shade.push_back(s);
                         fs.geom_meta.push_back(SPCU_MAKE_META(kind, material));
                         return kind != SPCU_PRIM_TRIANGLE;
                     } };
    fs.view.geom    = geom.run(scene.m_accelerator_geometry);
    fs.geom_unbuilt = !geom.unbuilt_bounds.empty();
    fs.geom_bounds  = std::move(geom.unbuilt_bounds);

    // ---- lights ----
    std::unordered_map<const sp::Light*, uint32_t> light_id;
    AccelWalker lights{ fs.light_nodes, [&](const Hitable* h) -> bool {
                           const auto* light = dynamic_cast<const sp::Light*>(h);
                           if (!light) {
                               throw std::runtime_error("flatten: lights accelerator holds a non-light");
                           }
                           light_id.emplace(light, static_cast<uint32_t>(fs.lights.size()));
                           fs.lights.push_back(make_light(fs, light));
                           return true;
                       } };
    fs.view.lights_accel = lights.run(scene.m_accelerator_lights);
    for (const auto& l : scene.m_lights) {
        const auto it = light_id.find(l.get());
        if (it == light_id.end()) {
            throw std::runtime_error("flatten: light missing from the lights accelerator");
        }
        fs.light_order.push_back(it->second);
    }

    fs.finalize();
    return fs;
}

std::vector<float> jitter_table(unsigned spp)
{
    // Seed of pixel (0,0); float(seed)/FLT_MAX < 2^-96 is absorbed by the first addition in
    // RSequence::r_sequence (math/Sampler.h:35-44), so the table is the same for every pixel.
    auto               sampler = sp::RSequenceSampler::create_new_sequence(sp::Seed{ 0u << 16u | 0u });
    std::vector<float> out(static_cast<size_t>(spp) * 2);
    for (unsigned i = 0; i < spp; ++i) {
        const auto s   = sampler.get_next_2D();
        out[2 * i + 0] = s.x;
        out[2 * i + 1] = s.y;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------
// .spflat: little-endian dump of the vectors, for boxes where the reference sources do not exist.
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr char k_magic[8] = { 'S', 'P', 'F', 'L', 'A', 'T', '0', '1' };

template <typename T>
void write_vec(std::FILE* f, const std::vector<T>& v)
{
    const uint64_t n = v.size();
    std::fwrite(&n, sizeof n, 1, f);
    if (n) {
        std::fwrite(v.data(), sizeof(T), n, f);
    }
}

template <typename T>
void read_vec(std::FILE* f, std::vector<T>& v)
{
    uint64_t n = 0;
    if (std::fread(&n, sizeof n, 1, f) != 1) {
        throw std::runtime_error("spflat: truncated");
    }
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) {
        throw std::runtime_error("spflat: truncated");
    }
}
} // namespace

void save_flat_scene(const FlatScene& fs, const std::string& path)
{
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        throw std::runtime_error("spflat: cannot open " + path);
    }
    std::fwrite(k_magic, 1, 8, f);
    spcu_flat_scene head = fs.view; // pointers are meaningless on disk and are rewired on load
    std::fwrite(&head, sizeof head, 1, f);
    write_vec(f, fs.geom_nodes);
    write_vec(f, fs.geom_prims);
    write_vec(f, fs.geom_shade);
    write_vec(f, fs.geom_meta);
    write_vec(f, fs.light_nodes);
    write_vec(f, fs.lights);
    write_vec(f, fs.light_order);
    write_vec(f, fs.materials);
    write_vec(f, fs.bxdfs);
    write_vec(f, fs.float_pool);
    std::fclose(f);
}

FlatScene load_flat_scene(const std::string& path)
{
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        throw std::runtime_error("spflat: cannot open " + path);
    }
    FlatScene fs;
    char      magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, k_magic, 8) != 0 ||
        std::fread(&fs.view, sizeof fs.view, 1, f) != 1) {
        std::fclose(f);
        throw std::runtime_error("spflat: bad header in " + path);
    }
    try {
        read_vec(f, fs.geom_nodes);
        read_vec(f, fs.geom_prims);
        read_vec(f, fs.geom_shade);
        read_vec(f, fs.geom_meta);
        read_vec(f, fs.light_nodes);
        read_vec(f, fs.lights);
        read_vec(f, fs.light_order);
        read_vec(f, fs.materials);
        read_vec(f, fs.bxdfs);
        read_vec(f, fs.float_pool);
    } catch (...) {
        std::fclose(f);
        throw;
    }
    std::fclose(f);
    fs.finalize();
    return fs;
}

} // namespace spb200
