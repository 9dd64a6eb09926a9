"""TEST INFRASTRUCTURE — ctypes access to oracle/_ref/libsp_ref*.so (the real reference, compiled by
oracle/Makefile from /root/reference).  Used only by tests/, smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from simplepath_b200.capi import Accel, FlatScene, HIT_DTYPE, RAY_DTYPE, run_build

HERE = Path(__file__).resolve().parent
STRICT = HERE / "_ref" / "libsp_ref.so"        # -ffp-contract=off: the canonical parity build
FAST = HERE / "_ref" / "libsp_ref_fast.so"     # GCC default contraction: noise-floor report
STOCK_BINARY = HERE / "_ref" / "SimplePath"    # unmodified main.cpp, CMake-equivalent flags


def available(path: Path = STRICT) -> bool:
    return path.exists()


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class RefScene:
    """A scene parsed and held by the reference's own code."""

    def __init__(self, sp_path, lib_path: Path = STRICT):
        self.lib = lib = C.CDLL(str(lib_path))
        vp = C.c_void_p
        lib.spref_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        lib.spref_load.restype = vp
        lib.spref_free.argtypes = [vp]
        lib.spref_flat.argtypes = [vp, C.c_char_p, C.c_size_t]
        lib.spref_flat.restype = C.POINTER(FlatScene)
        lib.spref_save_flat.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.spref_jitter.argtypes = [C.c_uint, vp]
        lib.spref_trace_closest.argtypes = [vp, vp, C.c_uint64, vp, vp]
        lib.spref_trace_closest.restype = C.c_int64
        lib.spref_trace_any.argtypes = [vp, vp, C.c_uint64, vp]
        lib.spref_trace_lights.argtypes = [vp, vp, C.c_uint64, vp]
        lib.spref_trace_lights.restype = C.c_int64
        lib.spref_generate_rays.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint, vp]
        lib.spref_hit_records.argtypes = [vp, vp, C.c_uint64, vp]
        lib.spref_render.argtypes = [vp, C.c_char_p, C.c_uint, C.c_uint, vp, vp, vp]
        lib.spref_render.restype = C.c_double
        lib.spref_width.argtypes = [vp]
        lib.spref_height.argtypes = [vp]
        err = C.create_string_buffer(512)
        self.h = lib.spref_load(str(sp_path).encode(), err, 512)
        if not self.h:
            raise RuntimeError(f"reference failed to parse {sp_path}: {err.value.decode()}")
        self.width = lib.spref_width(self.h)
        self.height = lib.spref_height(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.spref_free(self.h)
            self.h = None

    __del__ = close

    def flat(self) -> "C.POINTER(FlatScene)":
        err = C.create_string_buffer(512)
        p = self.lib.spref_flat(self.h, err, 512)
        if not p:
            raise RuntimeError(f"flatten failed: {err.value.decode()}")
        return p

    def save_flat(self, path) -> None:
        err = C.create_string_buffer(512)
        if self.lib.spref_save_flat(self.h, str(path).encode(), err, 512) != 0:
            raise RuntimeError(f"save_flat failed: {err.value.decode()}")

    def jitter(self, spp: int) -> np.ndarray:
        out = np.empty((spp, 2), dtype=np.float32)
        self.lib.spref_jitter(spp, _p(out))
        return out

    def trace_closest(self, rays, counters: bool = False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        cnt = np.zeros(3, dtype=np.uint64)
        bad = self.lib.spref_trace_closest(self.h, _p(rays), rays.shape[0], _p(hits), _p(cnt) if counters else None)
        if bad != 0:
            raise AssertionError(f"harness walk disagrees with Scene::intersect on {bad} rays")
        return (hits, cnt) if counters else hits

    def trace_any(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.empty(rays.shape[0], dtype=np.uint8)
        self.lib.spref_trace_any(self.h, _p(rays), rays.shape[0], _p(out))
        return out

    def trace_lights(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        bad = self.lib.spref_trace_lights(self.h, _p(rays), rays.shape[0], _p(hits))
        if bad != 0:
            raise AssertionError(f"harness walk disagrees with Scene::intersect_lights on {bad} rays")
        return hits

    def generate_rays(self, pix, smp, spp: int) -> np.ndarray:
        pix = np.ascontiguousarray(pix, dtype=np.uint32)
        smp = np.ascontiguousarray(smp, dtype=np.uint32)
        rays = np.empty(pix.shape[0], dtype=RAY_DTYPE)
        self.lib.spref_generate_rays(self.h, _p(pix), _p(smp), pix.shape[0], spp, _p(rays))
        return rays

    def hit_records(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.empty((rays.shape[0], 6), dtype=np.float32)
        self.lib.spref_hit_records(self.h, _p(rays), rays.shape[0], _p(out))
        return out

    def geom_bounds(self, n_prims: int) -> np.ndarray:
        """Hitable::get_world_bounds of every geometry primitive in ID order [n_prims, 6] (zeros for planes)."""
        self.lib.spref_geom_bounds.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.spref_geom_bounds.restype = C.c_uint32
        out = np.zeros((n_prims, 6), dtype=np.float32)
        n = self.lib.spref_geom_bounds(self.h, _p(out))
        assert n == n_prims, (n, n_prims)
        return out

    def render(self, integrator: str, spp: int, threads: int):
        rgb = np.zeros((self.height, self.width, 3), dtype=np.float32)
        mean = np.zeros((self.height, self.width), dtype=np.float32)
        var = np.zeros((self.height, self.width), dtype=np.float32)
        secs = self.lib.spref_render(self.h, integrator.encode(), spp, threads, _p(rgb), _p(mean), _p(var))
        if secs < 0:
            raise ValueError(f"unknown integrator {integrator}")
        return rgb, mean, var, secs


def build_bvh(bounds, non_triangle=None, first_id: int = 0, capacity: int | None = None, lib_path: Path = STRICT) -> dict:
    """The reference's own BVHAccelerator(first, last) over boxes (spref_build_bvh): {nodes, order, head, root_bounds}."""
    lib = C.CDLL(str(lib_path))
    lib.spref_build_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                    C.POINTER(Accel), C.c_void_p, C.c_char_p, C.c_size_t]
    lib.spref_build_bvh.restype = C.c_int
    bounds = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 6)
    err = C.create_string_buffer(512)

    def call(*a):
        if lib.spref_build_bvh(*a) != 0:
            raise RuntimeError(f"spref_build_bvh: {err.value.decode()}")
    return run_build(call, bounds, non_triangle, first_id, capacity, with_ms=False, extra=(err, 512))


def write_image(rgb_sum, spp: int, path, lib_path: Path = STRICT) -> None:
    """The reference's own sp::write (Image/Image.cpp) of sums divided as main.cpp:100-102 divides them; format by extension."""
    lib = C.CDLL(str(lib_path))
    lib.spref_write_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.spref_write_image.restype = C.c_int
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float32)
    h, w = rgb_sum.shape[:2]
    err = C.create_string_buffer(512)
    if lib.spref_write_image(_p(rgb_sum), w, h, spp, str(path).encode(), err, 512) != 0:
        raise RuntimeError(f"spref_write_image: {err.value.decode()}")


def read_ply(path, object_to_world, cap_vertices: int, cap_triangles: int, lib_path: Path = STRICT) -> dict:
    """The reference's own read_ply + Mesh constructor: {vertices, normals (world space), indices [t,3], normal_xf [9]}."""
    lib = C.CDLL(str(lib_path))
    u32p = C.POINTER(C.c_uint32)
    lib.spref_read_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, u32p, u32p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_char_p, C.c_size_t]
    lib.spref_read_ply.restype = C.c_int
    xf = np.ascontiguousarray(object_to_world, dtype=np.float32).reshape(12)
    v = np.zeros((max(cap_vertices, 1), 3), dtype=np.float32)
    n = np.zeros((max(cap_vertices, 1), 3), dtype=np.float32)
    idx = np.zeros((max(cap_triangles, 1), 3), dtype=np.uint32)
    nxf = np.zeros(9, dtype=np.float32)
    nv, nt = C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(512)
    if lib.spref_read_ply(str(path).encode(), _p(xf), cap_vertices, cap_triangles, C.byref(nv), C.byref(nt), _p(v), _p(n), _p(idx),
                          _p(nxf), err, 512) != 0:
        raise RuntimeError(f"spref_read_ply: {err.value.decode()}")
    return {"vertices": v[:nv.value].copy(), "normals": n[:nv.value].copy(), "indices": idx[:nt.value].copy(), "normal_xf": nxf}


def accel_from_mesh(ply, object_to_world, unbounded, lib_path: Path = STRICT) -> dict:
    """internal::create_acceleration_structure (base/Scene.h:27-45) by the reference over [triangles of the mesh it reads
    itself | unbounded stand-ins where unbounded[i] != 0]: {order (ID -> list position), nodes, head}."""
    from simplepath_b200.capi import NODE_DTYPE
    lib = C.CDLL(str(lib_path))
    lib.spref_accel_from_mesh.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                          C.POINTER(Accel), C.c_char_p, C.c_size_t]
    lib.spref_accel_from_mesh.restype = C.c_int
    xf = np.ascontiguousarray(object_to_world, dtype=np.float32).reshape(12)
    ub = np.ascontiguousarray(unbounded, dtype=np.uint8)
    n = ub.shape[0]
    order = np.zeros(max(n, 1), dtype=np.uint32)
    nodes = np.zeros(max(n, 1), dtype=NODE_DTYPE)
    accel = Accel()
    err = C.create_string_buffer(512)
    if lib.spref_accel_from_mesh(str(ply).encode(), _p(xf), _p(ub), n, _p(order), _p(nodes), n, C.byref(accel), err, 512) != 0:
        raise RuntimeError(f"spref_accel_from_mesh: {err.value.decode()}")
    head = {k: int(getattr(accel, k)) for k in ("n_prims", "n_unbounded", "n_nodes", "root", "root_count", "max_depth")}
    return {"order": order[:n].copy(), "nodes": nodes[:head["n_nodes"]].copy(), "head": head}
