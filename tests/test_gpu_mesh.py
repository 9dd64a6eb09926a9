"""spcu_ingest_mesh (include/spcu.h) through the C-ABI.  Bit-exact: which faces survive, their order, every vertex position
(hence bounds, tree and hits downstream).  Tolerance, stated: shading normals, because normalize() in the reference multiplies
by an SSE rsqrtss estimate (math/Math.h:205-227) that only x86 reproduces — every component within 4 * 2^-23 of the normal's
length of the reference's value, times the conditioning of that vertex's sum (meshcases.normal_condition; 1 for ordinary
vertices, hundreds where opposing face normals nearly cancel)."""
import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_mesh import golden
import meshcases

pytestmark = pytest.mark.gpu
NORMAL_TOLERANCE = 4.0 * 2.0 ** -23


def normals_close(got, want, condition=None):
    """condition (meshcases.normal_condition): faces summed / |sum of their unit normals| per normal.  Where a vertex's face
    normals nearly cancel (the hub of a 100 K-face fan sums 10^5 unit vectors to a length of a few hundred) the last-bit
    differences of the summands are amplified by that ratio, and so is the tolerance."""
    scale = np.linalg.norm(want.astype(np.float64), axis=-1, keepdims=True)
    err = (np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale).max(axis=-1)
    tol = NORMAL_TOLERANCE * (np.ones(err.shape) if condition is None else condition)
    assert (err <= tol).all(), float((err / tol).max())


def test_ingest_matches_reference_mesh(ctx):
    z = golden()
    r = ctx.ingest_mesh(z["in_vertices"], z["in_faces"], z["object_to_world"], z["normal_xf"], material=3)
    idx = z["indices"]
    assert r["world_vertices"].tobytes() == z["vertices"].tobytes()
    assert len(r["prims"]) == len(idx)
    assert np.array_equal(r["prims"].reshape(-1, 3, 4)[:, :, :3], z["vertices"][idx])
    assert (r["prims"].reshape(-1, 3, 4)[:, :, 3] == 0).all() and (r["meta"] == (3 << 2)).all()
    cond_vertex, cond_corner = meshcases.normal_condition(z["in_vertices"], z["in_faces"])
    normals_close(r["world_normals"], z["normals"], cond_vertex)
    normals_close(r["shade"].reshape(-1, 3, 4)[:, :, :3], z["normals"][idx], cond_corner)


@pytest.mark.parametrize("n_tris", [12, 2048, 300_000])
def test_ingest_matches_oracle(ctx, oracle_port, n_tris):
    v, f = meshcases.mesh(n_tris=n_tris, seed=n_tris, fan=200 if n_tris < 300_000 else 100_000)  # one vertex in 100 K faces
    xf = meshcases.transform(seed=n_tris)
    m = xf[:9].reshape(3, 3).T.astype(np.float64)               # columns c0 c1 c2
    nxf = np.linalg.inv(m).T.T.reshape(9).astype(np.float32)    # inverse-transposed, column major (= inverse, row major)
    got, want = ctx.ingest_mesh(v, f, xf, nxf, 1), oracle_port.ingest_mesh(v, f, xf, nxf, 1)
    assert got["world_vertices"].tobytes() == want["world_vertices"].tobytes()
    assert got["prims"].tobytes() == want["prims"].tobytes()
    assert np.array_equal(got["meta"], want["meta"])
    cond_vertex, cond_corner = meshcases.normal_condition(v, f)
    assert cond_corner.shape[0] == len(want["prims"])
    normals_close(got["world_normals"], want["world_normals"], cond_vertex)
    normals_close(got["shade"].reshape(-1, 3, 4)[:, :, :3], want["shade"].reshape(-1, 3, 4)[:, :, :3], cond_corner)


def test_ingest_edge_cases(ctx, oracle_port):
    from simplepath_b200 import capi
    xf = meshcases.transform()
    nxf = np.eye(3, dtype=np.float32).reshape(9)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], dtype=np.float32)
    r = ctx.ingest_mesh(v, np.zeros((0, 3), np.uint32), xf, nxf)                      # no faces: every normal is (0, 1, 0)
    assert len(r["prims"]) == 0 and np.array_equal(r["world_normals"], np.tile(np.float32([0, 1, 0]), (4, 1)))
    r = ctx.ingest_mesh(v, np.array([[0, 0, 1], [0, 1, 2]], np.uint32), xf, nxf)      # first face has no area
    w = oracle_port.ingest_mesh(v, np.array([[0, 0, 1], [0, 1, 2]], np.uint32), xf, nxf)
    assert len(r["prims"]) == 1 and r["prims"].tobytes() == w["prims"].tobytes()
    with pytest.raises(capi.SpcuError):
        ctx.ingest_mesh(v, np.array([[0, 1, 4]], np.uint32), xf, nxf)                 # vertices.at(4) throws in the reference


def test_file_to_hits_on_the_device(ctx, oracle_port):
    """The whole construction side: vertex / face lists -> spcu_ingest_mesh -> spcu_upload_scene_build -> batch queries,
    against the all-CPU pipeline (oracle ingest -> oracle construction -> oracle traversal): IDs and distances bit-equal."""
    from simplepath_b200 import capi
    from simplepath_b200.flat import FlatSceneData
    from test_gpu_build import _clone
    shell = FlatSceneData.load(GOLDEN / "g_bunny.flat.npz")
    vec = np.load(GOLDEN / "g_bunny.vectors.npz")
    nu = shell.head["geom"]["n_unbounded"]
    material = int(shell.arrays["geom_meta"].view(np.uint32).reshape(-1)[nu] >> 2)
    z = golden()

    def scene_with(ing):
        s = _clone(shell)
        k = len(ing["prims"])
        for name, rec, width in (("geom_prims", ing["prims"], 48), ("geom_shade", ing["shade"], 48), ("geom_meta", ing["meta"], 4)):
            s.arrays[name] = np.concatenate([shell.arrays[name][:nu], np.ascontiguousarray(rec).view(np.uint8).reshape(k, width)])
        s.head["geom"] = dict(shell.head["geom"], n_prims=nu + k, n_nodes=0, root=~nu, root_count=0, max_depth=0)
        s.arrays["geom_nodes"] = np.zeros((0, 64), dtype=np.uint8)
        return s

    dev = scene_with(ctx.ingest_mesh(z["in_vertices"], z["in_faces"], z["object_to_world"], z["normal_xf"], material))
    order, head = ctx.upload_scene_build(dev.pointer(), vec["jitter"], keepalive=dev)

    cpu_in = oracle_port.ingest_mesh(z["in_vertices"], z["in_faces"], z["object_to_world"], z["normal_xf"], material)
    built = oracle_port.build_bvh(oracle_port.triangle_bounds(cpu_in["prims"]), None, nu)
    cpu = scene_with({k: cpu_in[k][built["order"]] for k in ("prims", "shade", "meta")})
    cpu.arrays["geom_nodes"] = built["nodes"].view(np.uint8).reshape(-1, 64).copy()
    cpu.head["geom"] = built["head"]
    assert head == built["head"] and np.array_equal(order, built["order"])

    rng = np.random.default_rng(4)
    n = 1 << 15
    tri = z["vertices"][z["indices"][rng.integers(0, len(z["indices"]), n)]]           # aim at random triangles, from nearby
    w = rng.random((n, 3, 1)).astype(np.float32)
    target = (tri * (w / w.sum(axis=1, keepdims=True))).sum(axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["o"] = target - d * np.float32(0.5)
    rays["d"] = d
    rays["t_min"] = 0.0
    rays["t_max"] = 4.0
    got, want = ctx.trace_closest(rays), oracle_port.trace_closest(cpu.pointer(), rays)
    assert (want["id"] >= nu).mean() > 0.2
    assert np.array_equal(got["id"], want["id"]) and got["t"].tobytes() == want["t"].tobytes()


@pytest.mark.parametrize("name", ["plane_first", "planes_around", "planes_inside", "no_plane"])
def test_device_pipeline_reproduces_the_reference_scene_accelerator(ctx, name):
    """vertex / face lists -> spcu_ingest_mesh -> std::partition(is_bounded) order -> spcu_upload_scene_build, against what the
    reference's own create_acceleration_structure built over the triangles of the mesh it read itself (scene_accel.npz): the
    same primitive behind every ID and the same header; and spcu_build_bvh on the same bounds gives the same nodes."""
    from simplepath_b200.flat import FlatSceneData
    from test_gpu_build import _clone
    from test_oracle_mesh import golden_accel, list_order
    z, g = golden(), golden_accel(name)
    shell = FlatSceneData.load(GOLDEN / "g_bunny.flat.npz")
    vec = np.load(GOLDEN / "g_bunny.vectors.npz")
    material = int(shell.arrays["geom_meta"].view(np.uint32).reshape(-1)[1] >> 2)
    ing = ctx.ingest_mesh(z["in_vertices"], z["in_faces"], z["object_to_world"], z["normal_xf"], material)
    bounded, unb, rec = list_order(ing, g["unbounded"])
    s = _clone(shell)
    plane = {k: shell.arrays[k][:1] for k in ("geom_prims", "geom_shade", "geom_meta")}   # g_bunny's plane, once per stand-in
    for key, width in (("geom_prims", 48), ("geom_shade", 48), ("geom_meta", 4)):
        short = key.split("_")[1]
        s.arrays[key] = np.concatenate([np.repeat(plane[key], len(unb), axis=0),
                                        np.ascontiguousarray(rec[short]).view(np.uint8).reshape(len(bounded), width)])
    s.head["geom"] = dict(shell.head["geom"], n_prims=len(unb) + len(bounded), n_unbounded=len(unb), n_nodes=0, root=~len(unb),
                          root_count=0, max_depth=0)
    s.arrays["geom_nodes"] = np.zeros((0, 64), dtype=np.uint8)
    order, head = ctx.upload_scene_build(s.pointer(), vec["jitter"], keepalive=s)
    assert head == g["head"]
    assert np.array_equal(np.concatenate([unb, bounded[order]]), g["order"])
    built = ctx.build_bvh(ctx.triangle_bounds(rec["prims"]), None, len(unb))
    assert built["nodes"].tobytes() == g["nodes"].tobytes()


def test_stl_ingest_matches_reference_mesh(ctx, oracle_port):
    """spcu_ingest_mesh_stl against the reference's own read_stl (tests/golden/mesh_ingest_stl.npz) and the oracle."""
    from test_oracle_mesh import golden_stl
    z, v, f = golden_stl()
    r = ctx.ingest_mesh_stl(v, f, z["stored_normals"], z["object_to_world"], z["normal_xf"], material=2)
    w = oracle_port.ingest_mesh_stl(v, f, z["stored_normals"], z["object_to_world"], z["normal_xf"], material=2)
    assert len(r["prims"]) == len(f)
    assert r["world_vertices"].tobytes() == z["vertices"].tobytes()
    assert r["prims"].tobytes() == w["prims"].tobytes() and np.array_equal(r["meta"], w["meta"])
    # conditioning of each vertex sum, from the normals that actually contribute
    tri = z["corners"].astype(np.float64)
    n = z["stored_normals"].astype(np.float64)
    cross = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    zero = (np.abs(n) <= 1e-5).all(axis=1)
    n[zero] = cross[zero]
    used = ~(np.abs(n) <= 1e-5).all(axis=1)
    unit = n[used] / np.linalg.norm(n[used], axis=1, keepdims=True)
    total, count = np.zeros((len(v), 3)), np.zeros(len(v))
    for k in range(3):
        np.add.at(total, f[used][:, k].astype(np.int64), unit)
        np.add.at(count, f[used][:, k].astype(np.int64), 1.0)
    length = np.linalg.norm(total, axis=1)
    cond = np.maximum(np.where(length > 0, count / np.maximum(length, 1e-300), 1.0), 1.0)
    normals_close(r["world_normals"], z["normals"], cond)
    normals_close(r["shade"].reshape(-1, 3, 4)[:, :, :3], z["normals"][z["indices"]], cond[f.astype(np.int64)])
