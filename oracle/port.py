"""TEST INFRASTRUCTURE — ctypes access to oracle/libsp_oracle.so, the plain-C restatement of the hot path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from simplepath_b200.capi import Accel, FlatScene, HIT_DTYPE, RAY_DTYPE, Light, Partition, Stats, run_build, run_ingest

HERE = Path(__file__).resolve().parent
LIB = HERE / "libsp_oracle.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(str(LIB))
        vp, fsp = C.c_void_p, C.POINTER(FlatScene)
        l.spo_trace_closest.argtypes = [fsp, vp, C.c_uint64, vp, vp]
        l.spo_trace_any.argtypes = [fsp, vp, C.c_uint64, vp]
        l.spo_trace_lights.argtypes = [fsp, vp, C.c_uint64, vp]
        l.spo_generate_rays.argtypes = [fsp, vp, vp, vp, C.c_uint64, vp]
        l.spo_hit_records.argtypes = [fsp, vp, C.c_uint64, vp, vp]
        for fn in (l.spo_trace_closest, l.spo_trace_any, l.spo_trace_lights, l.spo_generate_rays, l.spo_hit_records):
            fn.restype = None
        _lib = l
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def trace_closest(flat, rays, counters=False):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
    cnt = np.zeros(3, dtype=np.uint64)
    lib().spo_trace_closest(flat, _p(rays), rays.shape[0], _p(hits), _p(cnt) if counters else None)
    return (hits, cnt) if counters else hits


def trace_any(flat, rays):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.empty(rays.shape[0], dtype=np.uint8)
    lib().spo_trace_any(flat, _p(rays), rays.shape[0], _p(out))
    return out


def trace_lights(flat, rays):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
    lib().spo_trace_lights(flat, _p(rays), rays.shape[0], _p(hits))
    return hits


def generate_rays(flat, jitter, pix, smp):
    jitter = np.ascontiguousarray(jitter, dtype=np.float32)
    pix = np.ascontiguousarray(pix, dtype=np.uint32)
    smp = np.ascontiguousarray(smp, dtype=np.uint32)
    rays = np.empty(pix.shape[0], dtype=RAY_DTYPE)
    lib().spo_generate_rays(flat, _p(jitter), _p(pix), _p(smp), pix.shape[0], _p(rays))
    return rays


def hit_records(flat, rays):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.empty((rays.shape[0], 6), dtype=np.float32)
    mat = np.empty(rays.shape[0], dtype=np.int32)
    lib().spo_hit_records(flat, _p(rays), rays.shape[0], _p(out), _p(mat))
    return out, mat


def render(flat, jitter, part: Partition, threads: int = 0, want_sumsq: bool = True):
    """spo_render over a partition: (rgb_sum [H,W,3], lum_sumsq [H,W] | None, stats dict)."""
    import os
    l = lib()
    l.spo_render.argtypes = [C.POINTER(FlatScene), C.c_void_p, C.POINTER(Partition), C.c_void_p, C.c_void_p,
                             C.POINTER(Stats), C.c_int]
    l.spo_render.restype = None
    fs = flat.contents if isinstance(flat, C.POINTER(FlatScene)) else flat
    jitter = np.ascontiguousarray(jitter, dtype=np.float32)
    rgb = np.zeros((fs.height, fs.width, 3), dtype=np.float32)
    sq = np.zeros((fs.height, fs.width), dtype=np.float32) if want_sumsq else None
    st = Stats()
    l.spo_render(flat, _p(jitter), C.byref(part), _p(rgb), _p(sq) if want_sumsq else None, C.byref(st),
                 threads or (os.cpu_count() or 1))
    return rgb, sq, st.as_dict()


def rng4(seed: int, pixel: int, sample: int, stream: int, ctr: int) -> np.ndarray:
    l = lib()
    l.spo_rng4.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    l.spo_rng4.restype = None
    out = np.zeros(4, dtype=np.float32)
    l.spo_rng4(seed, pixel, sample, stream, ctr, _p(out))
    return out


def build_bvh(bounds, non_triangle=None, first_id: int = 0, capacity: int | None = None) -> dict:
    """spo_build_bvh: {nodes, order, head, root_bounds}."""
    l = lib()
    l.spo_build_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                C.POINTER(Accel), C.c_void_p]
    l.spo_build_bvh.restype = C.c_int
    bounds = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 6)

    def call(*a):
        if l.spo_build_bvh(*a) != 0:
            raise RuntimeError("spo_build_bvh: node capacity exceeded")
    return run_build(call, bounds, non_triangle, first_id, capacity, with_ms=False)


def triangle_bounds(tris) -> np.ndarray:
    l = lib()
    l.spo_triangle_bounds.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    l.spo_triangle_bounds.restype = None
    tris = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 12)
    out = np.empty((tris.shape[0], 6), dtype=np.float32)
    l.spo_triangle_bounds(_p(tris), tris.shape[0], _p(out))
    return out


def pack_image(rgb_sum, spp: int, fmt: int) -> np.ndarray:
    """spo_pack_image: write_pfm payload (fmt 0, float32) / write_ppm numbers (fmt 1, uint16), rows bottom-up."""
    l = lib()
    l.spo_pack_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    l.spo_pack_image.restype = None
    rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float32)
    h, w = rgb_sum.shape[:2]
    out = np.empty((h, w, 3), dtype=np.float32 if fmt == 0 else np.uint16)
    l.spo_pack_image(_p(rgb_sum), w, h, spp, fmt, _p(out))
    return out


def ingest_mesh(vertices, faces, object_to_world, normal_xf, material: int = 0) -> dict:
    """spo_ingest_mesh (read_ply's normal passes + Mesh's constructor): same dict as capi.Context.ingest_mesh."""
    l = lib()
    vp = C.c_void_p
    l.spo_ingest_mesh.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32), vp, vp]
    l.spo_ingest_mesh.restype = None
    return run_ingest(l.spo_ingest_mesh, vertices, faces, object_to_world, normal_xf, material, with_ms=False)


def ingest_mesh_stl(vertices, faces, face_normals, object_to_world, normal_xf, material: int = 0) -> dict:
    """spo_ingest_mesh_stl (read_binary_stl's normal passes + Mesh's constructor)."""
    l = lib()
    vp = C.c_void_p
    l.spo_ingest_mesh_stl.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32), vp, vp]
    l.spo_ingest_mesh_stl.restype = None
    return run_ingest(l.spo_ingest_mesh_stl, vertices, faces, object_to_world, normal_xf, material, with_ms=False,
                      face_normals=face_normals)
