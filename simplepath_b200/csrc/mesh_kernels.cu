// Mesh ingest on the device: what read_ply does after it has read the vertex and face lists
// (reference base/PlyReader.cpp:487-531) and what Mesh's constructor does with the result (shapes/Triangle.h:25-51):
//
//   per face      edge0 = v1 - v0, edge1 = v2 - v0, n = cross(edge0, edge1)  (compensated products, math/Vector3.h:489-498,769-775);
//                 a face with sqr_length(n) == 0 is dropped; n = normalize(n)
//   per vertex    normal = sum of the normals of the kept faces that use it, IN FACE ORDER (float addition does not commute
//                 with reordering); normalize, or (0, 1, 0) when the sum is zero
//   Mesh          vertex -> object_to_world(vertex) (fma chain, math/AffineSpace.h:79-86); normal -> object_to_world(normal)
//                 = inverse(linear).transposed() * normal (math/LinearSpace3x3.h:163-167), NOT renormalised
//   per triangle  the pre-gathered records of include/spcu.h (three world-space vertices / three normals), in face order =
//                 the order the parser appends the mesh's triangles to the scene = the pre-construction order
//                 spcu_upload_scene_build expects.
//
// Everything that decides geometry (which faces survive, vertex positions, hence bounds, tree and hits) is bit-exact: explicit
// round-to-nearest intrinsics, fma exactly where the reference calls madd / msub / nmadd, --fmad=false.  normalize() is the
// one operation that cannot be: the reference multiplies by an SSE rsqrtss estimate refined by one Newton step
// (math/Math.h:205-227), a value no other hardware reproduces; the device uses the correctly rounded reciprocal square root,
// so shading normals agree to a few ulp (stated and tested), as for camera directions.
#include "build_util.cuh"

using namespace spcu;

namespace {

struct F3
{
    float x, y, z;
};

__device__ __forceinline__ F3 ld3(const float* p, size_t i) { return F3{ __ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2) }; }
__device__ __forceinline__ F3 sub(F3 a, F3 b) { return F3{ __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z) }; }

// difference_of_products (math/Vector3.h:489-498): a*b - c*d with the rounding error of c*d recovered
__device__ __forceinline__ float dop(float a, float b, float c, float d)
{
    const float cd  = __fmul_rn(c, d);
    const float err = __fmaf_rn(-c, d, cd);
    const float r   = __fmaf_rn(a, b, -cd);
    return __fadd_rn(r, err);
}
// cross (math/Vector3.h:769-775)
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return F3{ dop(a.y, b.z, a.z, b.y), dop(a.z, b.x, a.x, b.z), dop(a.x, b.y, a.y, b.x) }; }
// dot = _mm_dp_ps: (x + y) + (z + 0) (math/Vector3.h:742-746)
__device__ __forceinline__ float dot(F3 a, F3 b)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.z, b.z), 0.0f));
}
__device__ __forceinline__ F3 normalize(F3 a)
{
    const float s = __frsqrt_rn(dot(a, a));
    return F3{ __fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s) };
}

// is_zero (math/Vector3.h:644-647) through float_compare(a, 0) (math/Math.h:265-272): |a| <= 1e-5 (the relative branch can
// never hold against zero)
__device__ __forceinline__ bool is_zero_eps(F3 a) { return fabsf(a.x) <= 1.0e-05f && fabsf(a.y) <= 1.0e-05f && fabsf(a.z) <= 1.0e-05f; }

// face pass.  PLY (file_normals == NULL; base/PlyReader.cpp:493-503): the normal is the cross product, a face whose normal
// has sqr_length == 0 exactly is dropped altogether.  STL (base/STLReader.cpp:107-118): the normal comes from the file unless
// it is_zero (then the cross product); a face whose normal still is_zero adds nothing to the vertex normals — but its indices
// were pushed before the test (:95-96), so the triangle STAYS in the mesh.  contributes: adds to vertex normals; keep: emitted.
__global__ void __launch_bounds__(kBlock) k_face_normals(const float* verts, const uint32_t* faces, uint32_t nf, const float* file_normals,
                                                         float* face_n, uint8_t* keep, uint8_t* contributes)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += stride) {
        const uint32_t i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
        const F3       v0 = ld3(verts, i0);
        F3             n;
        bool           ok;
        if (file_normals) {
            n = ld3(file_normals, f);
            if (is_zero_eps(n)) {
                n = cross(sub(ld3(verts, i1), v0), sub(ld3(verts, i2), v0));
            }
            ok      = !is_zero_eps(n);
            keep[f] = 1;
        } else {
            n       = cross(sub(ld3(verts, i1), v0), sub(ld3(verts, i2), v0));
            ok      = dot(n, n) != 0.0f;
            keep[f] = ok ? 1 : 0;
        }
        contributes[f] = ok ? 1 : 0;
        if (ok) {
            n = normalize(n);
        }
        face_n[3 * f] = n.x, face_n[3 * f + 1] = n.y, face_n[3 * f + 2] = n.z;
    }
}

// kept faces, compacted in order: kept_faces[rank] = face; and the vertex degrees over kept faces
__global__ void __launch_bounds__(kBlock) k_compact_faces(const uint32_t* faces, const uint8_t* keep, const uint8_t* contributes,
                                                          const uint32_t* rank, uint32_t nf, uint32_t* kept, uint32_t* degree)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += stride) {
        if (keep[f]) {
            kept[rank[f]] = f;
        }
        if (!contributes[f]) {
            continue;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicAdd(&degree[__ldg(faces + 3 * f + k)], 1u);
        }
    }
}

// adjacency lists (vertex -> kept faces that use it, one entry per corner), filled in arbitrary order
__global__ void __launch_bounds__(kBlock) k_fill_adjacency(const uint32_t* faces, const uint8_t* contributes, uint32_t nf,
                                                           const uint32_t* offset, uint32_t* cursor, uint32_t* adjacency)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += stride) {
        if (!contributes[f]) {
            continue;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t v                               = __ldg(faces + 3 * f + k);
            adjacency[offset[v] + atomicAdd(&cursor[v], 1u)] = f;
        }
    }
}

// In-place ascending sort of a vertex's face list: insertion sort for the usual handful of entries, heap sort beyond (a
// vertex shared by a huge fan must not turn into a quadratic loop on one thread).
__device__ void sort_ascending(uint32_t* list, uint32_t n)
{
    if (n <= 32u) {
        for (uint32_t i = 1; i < n; ++i) {
            const uint32_t key = list[i];
            uint32_t       j   = i;
            for (; j > 0 && list[j - 1] > key; --j) {
                list[j] = list[j - 1];
            }
            list[j] = key;
        }
        return;
    }
    auto sift = [list](uint32_t root, uint32_t end) { // max-heap on [0, end)
        const uint32_t key = list[root];
        for (;;) {
            uint32_t child = 2u * root + 1u;
            if (child >= end) {
                break;
            }
            if (child + 1u < end && list[child + 1u] > list[child]) {
                ++child;
            }
            if (list[child] <= key) {
                break;
            }
            list[root] = list[child];
            root       = child;
        }
        list[root] = key;
    };
    for (uint32_t i = n / 2u; i-- > 0u;) {
        sift(i, n);
    }
    for (uint32_t end = n - 1u; end > 0u; --end) {
        const uint32_t t = list[0];
        list[0]          = list[end];
        list[end]        = t;
        sift(0u, end);
    }
}

// vertex normals (base/PlyReader.cpp:509-528) and Mesh's transforms (shapes/Triangle.h:37-47).  The sequential loop adds
// face normals face by face, so a vertex sees its faces in ascending face index: sort the (short) list, then add in order.
__global__ void __launch_bounds__(kBlock) k_vertices(const float* verts, uint32_t nv, const uint32_t* offset, uint32_t* adjacency,
                                                     const float* face_n, const float* xf, const float* nxf, float* world_v,
                                                     float* world_n)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += stride) {
        uint32_t*      list = adjacency + offset[v];
        const uint32_t deg  = offset[v + 1] - offset[v];
        sort_ascending(list, deg);
        F3 n{ 0.0f, 0.0f, 0.0f };
        for (uint32_t i = 0; i < deg; ++i) {
            const F3 f = ld3(face_n, list[i]);
            n          = F3{ __fadd_rn(n.x, f.x), __fadd_rn(n.y, f.y), __fadd_rn(n.z, f.z) };
        }
        n = (n.x != 0.0f || n.y != 0.0f || n.z != 0.0f) ? normalize(n) : F3{ 0.0f, 1.0f, 0.0f };
        // AffineSpace::operator()(Point3): madd(x, c0, madd(y, c1, madd(z, c2, affine)))
        const F3 p = ld3(verts, v);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            world_v[3 * v + a] = __fmaf_rn(p.x, xf[a], __fmaf_rn(p.y, xf[3 + a], __fmaf_rn(p.z, xf[6 + a], xf[9 + a])));
            // LinearSpace3x3::operator()(Normal3): madd(x, c0, madd(y, c1, z * c2)) with the inverse-transposed matrix
            world_n[3 * v + a] = __fmaf_rn(n.x, nxf[a], __fmaf_rn(n.y, nxf[3 + a], __fmul_rn(n.z, nxf[6 + a])));
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_triangle_records(const uint32_t* faces, const uint32_t* kept, uint32_t n_kept, const float* world_v,
                                                             const float* world_n, uint32_t meta_word, float4* prims, float4* shade,
                                                             uint32_t* meta)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_kept; t += stride) {
        const uint32_t f = kept[t];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = __ldg(faces + 3 * f + k);
            const F3       p = ld3(world_v, v), n = ld3(world_n, v);
            prims[3 * static_cast<size_t>(t) + k] = make_float4(p.x, p.y, p.z, 0.0f);
            shade[3 * static_cast<size_t>(t) + k] = make_float4(n.x, n.y, n.z, 0.0f);
        }
        meta[t] = meta_word;
    }
}

} // namespace

static int ingest(spcu_ctx* c, const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const float* file_normals,
                  const float object_to_world[12], const float normal_xf[9], uint32_t material, spcu_prim_geom* prims,
                  spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept_out, float* world_vertices, float* world_normals,
                  float* device_ms)
{
    if (!c) {
        return SPCU_ERR_INVALID;
    }
    if (!n_kept_out || !object_to_world || !normal_xf || (nv && !vertices) || (nf && (!faces || !prims || !shade || !meta))) {
        return fail(c, SPCU_ERR_INVALID, "spcu_ingest_mesh: NULL argument");
    }
    if (nf >= (1u << 30) || nv >= (1u << 30)) {
        return fail(c, SPCU_ERR_LIMIT, "spcu_ingest_mesh: more than 2^30 vertices or faces");
    }
    for (uint32_t i = 0; i < 3 * nf; ++i) { // vertices.at(...) throws in the reference (base/PlyReader.cpp:493-494)
        if (faces[i] >= nv) {
            return fail(c, SPCU_ERR_INVALID, "spcu_ingest_mesh: face %u refers to vertex %u of %u", i / 3, faces[i], nv);
        }
    }
    CK(c, cudaSetDevice(c->device));
    *n_kept_out = 0;
    if (device_ms) {
        *device_ms = 0.0f;
    }
    const cudaStream_t st = c->stream;
    Scratch            mem;
    float *            d_v = nullptr, *d_face_n = nullptr, *d_xf = nullptr, *d_world_v = nullptr, *d_world_n = nullptr;
    uint32_t *         d_f = nullptr, *d_rank = nullptr, *d_kept = nullptr, *d_degree = nullptr, *d_offset = nullptr, *d_cursor = nullptr,
             *d_adj = nullptr, *d_partials = nullptr;
    uint8_t *d_keep = nullptr, *d_contributes = nullptr;
    float*   d_file_n = nullptr;
    const size_t f_pad = static_cast<size_t>(scan_tiles(nf)) * kScanTile, v_pad = static_cast<size_t>(scan_tiles(nv)) * kScanTile;
    CK(c, mem.get(&d_v, 3 * static_cast<size_t>(nv)));
    CK(c, mem.get(&d_f, 3 * static_cast<size_t>(nf)));
    CK(c, mem.get(&d_face_n, 3 * static_cast<size_t>(nf)));
    CK(c, mem.get(&d_keep, f_pad));
    CK(c, mem.get(&d_contributes, std::max<size_t>(nf, 1)));
    if (file_normals && nf) {
        CK(c, mem.get(&d_file_n, 3 * static_cast<size_t>(nf)));
        CK(c, cudaMemcpyAsync(d_file_n, file_normals, 3 * static_cast<size_t>(nf) * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    CK(c, mem.get(&d_rank, static_cast<size_t>(nf) + 1));
    CK(c, mem.get(&d_kept, nf));
    CK(c, mem.get(&d_degree, v_pad));
    CK(c, mem.get(&d_offset, static_cast<size_t>(nv) + 1));
    CK(c, mem.get(&d_cursor, nv));
    CK(c, mem.get(&d_adj, 3 * static_cast<size_t>(nf)));
    CK(c, mem.get(&d_partials, std::max(scan_tiles(nf), scan_tiles(nv))));
    CK(c, mem.get(&d_xf, 24));
    CK(c, mem.get(&d_world_v, 3 * static_cast<size_t>(nv)));
    CK(c, mem.get(&d_world_n, 3 * static_cast<size_t>(nv)));
    float4 *  d_prims = nullptr, *d_shade = nullptr; // sized for every face: no allocation between the timed kernels
    uint32_t* d_meta = nullptr;
    CK(c, mem.get(&d_prims, 3 * static_cast<size_t>(nf)));
    CK(c, mem.get(&d_shade, 3 * static_cast<size_t>(nf)));
    CK(c, mem.get(&d_meta, nf));
    if (nv) {
        if (const int rc = copy_to_device(c, d_v, vertices, 3 * static_cast<size_t>(nv) * sizeof(float)); rc != SPCU_OK) return rc;
    }
    if (nf) {
        if (const int rc = copy_to_device(c, d_f, faces, 3 * static_cast<size_t>(nf) * sizeof(uint32_t)); rc != SPCU_OK) return rc;
    }
    CK(c, cudaMemcpyAsync(d_xf, object_to_world, 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(d_xf + 12, normal_xf, 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(c, cudaMemsetAsync(d_keep, 0, f_pad, st));
    CK(c, cudaMemsetAsync(d_degree, 0, v_pad * sizeof(uint32_t), st));
    CK(c, cudaMemsetAsync(d_cursor, 0, std::max<size_t>(nv, 1) * sizeof(uint32_t), st));

    CK(c, cudaEventRecord(c->ev0, st));
    uint32_t n_kept = 0;
    if (nf) {
        k_face_normals<<<grid_for(nf, c->sm_count), kBlock, 0, st>>>(d_v, d_f, nf, d_file_n, d_face_n, d_keep, d_contributes);
        exclusive_scan(d_keep, d_rank, nf, d_partials, st);
        k_compact_faces<<<grid_for(nf, c->sm_count), kBlock, 0, st>>>(d_f, d_keep, d_contributes, d_rank, nf, d_kept, d_degree);
    }
    exclusive_scan(d_degree, d_offset, nv, d_partials, st);
    if (nf) {
        k_fill_adjacency<<<grid_for(nf, c->sm_count), kBlock, 0, st>>>(d_f, d_contributes, nf, d_offset, d_cursor, d_adj);
        CK(c, cudaMemcpyAsync(&n_kept, d_rank + nf, sizeof n_kept, cudaMemcpyDeviceToHost, st));
    }
    if (nv) {
        k_vertices<<<grid_for(nv, c->sm_count), kBlock, 0, st>>>(d_v, nv, d_offset, d_adj, d_face_n, d_xf, d_xf + 12, d_world_v, d_world_n);
    }
    CK(c, cudaStreamSynchronize(st)); // n_kept
    if (n_kept) {
        k_triangle_records<<<grid_for(n_kept, c->sm_count), kBlock, 0, st>>>(d_f, d_kept, n_kept, d_world_v, d_world_n,
                                                                             SPCU_MAKE_META(SPCU_PRIM_TRIANGLE, material), d_prims, d_shade,
                                                                             d_meta);
    }
    CK(c, cudaEventRecord(c->ev1, st));
    CK(c, cudaGetLastError());
    if (n_kept) {
        CK(c, cudaMemcpyAsync(prims, d_prims, static_cast<size_t>(n_kept) * sizeof(spcu_prim_geom), cudaMemcpyDeviceToHost, st));
        CK(c, cudaMemcpyAsync(shade, d_shade, static_cast<size_t>(n_kept) * sizeof(spcu_prim_shade), cudaMemcpyDeviceToHost, st));
        CK(c, cudaMemcpyAsync(meta, d_meta, static_cast<size_t>(n_kept) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (world_vertices && nv) {
        CK(c, cudaMemcpyAsync(world_vertices, d_world_v, 3 * static_cast<size_t>(nv) * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (world_normals && nv) {
        CK(c, cudaMemcpyAsync(world_normals, d_world_n, 3 * static_cast<size_t>(nv) * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    CK(c, cudaStreamSynchronize(st));
    if (device_ms) {
        CK(c, cudaEventElapsedTime(device_ms, c->ev0, c->ev1));
    }
    *n_kept_out = n_kept;
    return SPCU_OK;
}

extern "C" int spcu_ingest_mesh(spcu_ctx* c, const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf,
                                const float object_to_world[12], const float normal_xf[9], uint32_t material, spcu_prim_geom* prims,
                                spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept, float* world_vertices, float* world_normals,
                                float* device_ms)
{
    return ingest(c, vertices, nv, faces, nf, nullptr, object_to_world, normal_xf, material, prims, shade, meta, n_kept, world_vertices,
                  world_normals, device_ms);
}

extern "C" int spcu_ingest_mesh_stl(spcu_ctx* c, const float* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf,
                                    const float* face_normals, const float object_to_world[12], const float normal_xf[9],
                                    uint32_t material, spcu_prim_geom* prims, spcu_prim_shade* shade, uint32_t* meta, uint32_t* n_kept,
                                    float* world_vertices, float* world_normals, float* device_ms)
{
    if (c && nf && !face_normals) {
        return fail(c, SPCU_ERR_INVALID, "spcu_ingest_mesh_stl: face_normals is NULL");
    }
    return ingest(c, vertices, nv, faces, nf, face_normals, object_to_world, normal_xf, material, prims, shade, meta, n_kept,
                  world_vertices, world_normals, device_ms);
}
