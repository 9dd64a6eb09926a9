/*
 * TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's hot path in plain C over the flattened
 * scene of include/spcu.h.  It is the checker for the CUDA backend, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load libsp_oracle.so.
 *
 * PARITY PIN: traversal / intersection functions are checked bit-for-bit against the real reference compiled
 * from /root/reference (oracle/_ref/libsp_ref.so, see tests/test_oracle_vs_reference.py and the committed
 * vectors under tests/golden/ produced by tests/golden/make_golden.py).  Shading functions are checked against
 * the reference's own classes through the same harness (statistically where the reference consumes its
 * mt19937_64 stream).  The reference's own unit tests hold no vectors for this path.
 */
#ifndef SP_ORACLE_H
#define SP_ORACLE_H

#include "spcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spo_counters {
    uint64_t nodes; /* internal nodes visited   (NodeInternal::intersect calls) */
    uint64_t tris;  /* triangle tests                                            */
    uint64_t xf;    /* sphere / plane tests                                      */
} spo_counters;

/* Scene::intersect (base/Scene.h:74-77) */
void spo_trace_closest(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits, spo_counters* cnt);
/* Scene::intersect_p (base/Scene.h:79-82) */
void spo_trace_any(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, uint8_t* out);
/* Scene::intersect_lights (base/Scene.h:69-72) */
void spo_trace_lights(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, spcu_hit* hits);
/* Camera::generate_ray + pixel jitter (Cameras/Camera.h:119-129, main.cpp:96-98) */
void spo_generate_rays(const spcu_flat_scene* s, const float* jitter, const uint32_t* pix, const uint32_t* smp,
                       uint64_t n, spcu_ray* rays);
/* Intersection record of the closest hit: normal xyz, point xyz (zeros on a miss), and material index or -1 */
void spo_hit_records(const spcu_flat_scene* s, const spcu_ray* rays, uint64_t n, float* normal_point, int32_t* material);

/* render_thread + Integrator::integrate (main.cpp:77-107; Integrators/Integrator.cpp) with the SAME counter-based
 * random numbers as the CUDA backend (oracle/sp_oracle.c: "RNG contract").  Accumulates like spcu_render:
 * rgb_sum[(y*w+x)*3+c] and lum_sumsq[y*w+x]; stats (may be NULL) receives the ray counters.  threads = OpenMP. */
void spo_render(const spcu_flat_scene* s, const float* jitter, const spcu_partition* part, float* rgb_sum,
                float* lum_sumsq, spcu_stats* stats, int threads);

/* The RNG contract itself, exposed for known-answer tests: the 4 words of block `ctr` of sub-stream `stream`
 * (depth << 16 | site) of path (pixel, sample) as floats in [0,1). */
void spo_rng4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t ctr, float out[4]);

/* Pieces of the shading model, exposed so tests can pin them against the reference's classes.
 * All vectors are 3 floats.  u = uniform numbers consumed in the order documented at each function. */
void  spo_onb_from_v(const float n[3], float u_out[3], float v_out[3], float w_out[3]);
float spo_fresnel_dielectric(float cos_theta_i, float eta_i, float eta_t);
float spo_erfinv(float a);
void  spo_beckmann_sample_wh(const float wo[3], float alpha_x, float alpha_y, float u1, float u2, float wh[3]);
float spo_beckmann_D(const float wh[3], float alpha_x, float alpha_y);
float spo_beckmann_lambda(const float w[3], float alpha_x, float alpha_y);
void  spo_sphere_light_sample(const spcu_light* l, const float p[3], const float n[3], float u0, float u1,
                              float wi[3], float* pdf, float* t_min, float* t_max);
float spo_sphere_light_pdf(const spcu_light* l, const float p[3]);

/* BVHAccelerator::construct (shapes/BVHAccelerator.h:175-209) restated sequentially; same arguments and outputs as
 * spcu_build_bvh (include/spcu.h).  root_bounds (6 floats, may be NULL) = bounds of the root node.  Returns 0, or -1
 * when `capacity` nodes do not suffice.  Pinned bit for bit against the reference's own BVHAccelerator
 * (oracle/ref_harness.cpp spref_build_bvh; tests/golden/bvh_build.npz). */
int spo_build_bvh(const spcu_bounds* bounds, uint32_t n, const uint8_t* non_triangle, uint32_t first_id, uint32_t* order,
                  spcu_bvh_node* nodes, uint32_t capacity, spcu_accel* accel, float* root_bounds);
/* Triangle::get_world_bounds_impl (shapes/Triangle.h:228-237) */
void spo_triangle_bounds(const spcu_prim_geom* tris, uint32_t n, spcu_bounds* out);

/* image(x, y) /= spp (main.cpp:100-102), then write_pfm's payload (format SPCU_IMAGE_PFM: float[h][w][3]) or the numbers
 * write_ppm prints (SPCU_IMAGE_PPM: uint16_t[h][w][3], clamped to [0, 65535]), rows bottom-up (Image/Image.cpp:14-55).
 * Pinned against files the reference's own sp::write produced (tests/golden/image_pack.npz). */
void spo_pack_image(const float* rgb_sum, uint32_t width, uint32_t height, uint32_t spp, uint32_t format, void* out);

/* read_ply's face / vertex-normal passes + Mesh's constructor (base/PlyReader.cpp:487-531, shapes/Triangle.h:25-51); same
 * arguments and outputs as spcu_ingest_mesh.  Pinned bit for bit (x86: same rsqrtss) against the reference's own read_ply
 * (oracle/ref_harness.cpp spref_read_ply; tests/golden/mesh_ingest.npz). */
void spo_ingest_mesh(const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float object_to_world[12],
                     const float normal_xf[9], uint32_t material, spcu_prim_geom* prims, spcu_prim_shade* shade, uint32_t* meta,
                     uint32_t* n_kept, float* world_vertices, float* world_normals);
/* read_binary_stl's face / vertex-normal passes (base/STLReader.cpp:107-137) + Mesh's constructor; as spcu_ingest_mesh_stl.
 * Pinned against the reference's own read_stl (spref_read_stl; tests/golden/mesh_ingest_stl.npz). */
void spo_ingest_mesh_stl(const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float* face_normals,
                         const float object_to_world[12], const float normal_xf[9], uint32_t material, spcu_prim_geom* prims,
                         spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept, float* world_vertices, float* world_normals);

#ifdef __cplusplus
}
#endif
#endif
