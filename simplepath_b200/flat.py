"""Flattened scenes (include/spcu.h `spcu_flat_scene`) as numpy arrays: copy out of a C struct, rebuild the C struct,
save / load as .npz.  Used by the host mirror, the tests' golden fixtures and bench.py."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .capi import (ABI_VERSION, Accel, Bxdf, BvhNode, FlatScene, Light, Material, PrimGeom, PrimShade)

_LIGHT_DTYPE = np.dtype((np.void, C.sizeof(Light)))


def _copy(ptr, n: int, ctype) -> np.ndarray:
    """n elements of `ctype` behind a ctypes pointer -> owned uint8 array [n, sizeof]."""
    size = C.sizeof(ctype)
    if n == 0:
        return np.zeros((0, size), dtype=np.uint8)
    buf = (C.c_uint8 * (n * size)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=np.uint8).reshape(n, size).copy()


class FlatSceneData:
    """Owns the arrays of one flattened scene and exposes a ctypes `FlatScene` view of them."""

    ARRAYS = ("geom_nodes", "geom_prims", "geom_shade", "geom_meta", "light_nodes", "lights", "light_order",
              "materials", "bxdfs", "float_pool")

    def __init__(self, head: dict, arrays: dict):
        self.head = dict(head)
        self.arrays = {k: np.ascontiguousarray(arrays[k]) for k in self.ARRAYS}
        self._struct = None

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_struct(cls, fs: "FlatScene | C.POINTER(FlatScene)") -> "FlatSceneData":
        s = fs.contents if isinstance(fs, C.POINTER(FlatScene)) else fs
        head = {
            "width": s.width, "height": s.height, "rr_depth": s.rr_depth, "max_depth": s.max_depth,
            "camera": [float(x) for x in s.camera],
            "geom": _accel_head(s.geom), "lights_accel": _accel_head(s.lights_accel),
        }
        arrays = {
            "geom_nodes": _copy(s.geom.nodes, s.geom.n_nodes, BvhNode),
            "geom_prims": _copy(s.geom_prims, s.geom.n_prims, PrimGeom),
            "geom_shade": _copy(s.geom_shade, s.geom.n_prims, PrimShade),
            "geom_meta": _copy(s.geom_meta, s.geom.n_prims, C.c_uint32),
            "light_nodes": _copy(s.lights_accel.nodes, s.lights_accel.n_nodes, BvhNode),
            "lights": _copy(s.lights, s.n_lights, Light),
            "light_order": _copy(s.light_order, s.n_lights, C.c_uint32),
            "materials": _copy(s.materials, s.n_materials, Material),
            "bxdfs": _copy(s.bxdfs, s.n_bxdfs, Bxdf),
            "float_pool": _copy(s.float_pool, s.n_pool, C.c_float),
        }
        return cls(head, arrays)

    @classmethod
    def load(cls, path: Path | str) -> "FlatSceneData":
        with np.load(str(path), allow_pickle=False) as z:
            head = {
                "width": int(z["dims"][0]), "height": int(z["dims"][1]), "rr_depth": int(z["dims"][2]),
                "max_depth": int(z["dims"][3]), "camera": [float(x) for x in z["camera"]],
                "geom": dict(zip(_ACCEL_KEYS, (int(v) for v in z["geom_head"]))),
                "lights_accel": dict(zip(_ACCEL_KEYS, (int(v) for v in z["lights_head"]))),
            }
            arrays = {k: z[k] for k in cls.ARRAYS}
        return cls(head, arrays)

    def save(self, path: Path | str) -> None:
        h = self.head
        np.savez_compressed(
            str(path),
            dims=np.array([h["width"], h["height"], h["rr_depth"], h["max_depth"]], dtype=np.int64),
            camera=np.array(h["camera"], dtype=np.float32),
            geom_head=np.array([h["geom"][k] for k in _ACCEL_KEYS], dtype=np.int64),
            lights_head=np.array([h["lights_accel"][k] for k in _ACCEL_KEYS], dtype=np.int64),
            **self.arrays)

    # ---- views ---------------------------------------------------------------------------------------------------
    @property
    def width(self) -> int:
        return self.head["width"]

    @property
    def height(self) -> int:
        return self.head["height"]

    @property
    def n_prims(self) -> int:
        return self.head["geom"]["n_prims"]

    @property
    def n_nodes(self) -> int:
        return self.head["geom"]["n_nodes"]

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in self.arrays.values()))

    def struct(self) -> FlatScene:
        """ctypes view; valid while this object is alive."""
        if self._struct is None:
            a, h = self.arrays, self.head
            s = FlatScene()
            s.abi_version = ABI_VERSION
            s.width, s.height, s.rr_depth, s.max_depth = h["width"], h["height"], h["rr_depth"], h["max_depth"]
            s.camera = (C.c_float * 12)(*h["camera"])
            s.geom = _accel_struct(h["geom"], a["geom_nodes"])
            s.geom_prims = a["geom_prims"].ctypes.data_as(C.POINTER(PrimGeom))
            s.geom_shade = a["geom_shade"].ctypes.data_as(C.POINTER(PrimShade))
            s.geom_meta = a["geom_meta"].ctypes.data_as(C.POINTER(C.c_uint32))
            s.lights_accel = _accel_struct(h["lights_accel"], a["light_nodes"])
            s.n_lights = a["lights"].shape[0]
            s.lights = a["lights"].ctypes.data_as(C.POINTER(Light))
            s.light_order = a["light_order"].ctypes.data_as(C.POINTER(C.c_uint32))
            s.n_materials = a["materials"].shape[0]
            s.n_bxdfs = a["bxdfs"].shape[0]
            s.materials = a["materials"].ctypes.data_as(C.POINTER(Material))
            s.bxdfs = a["bxdfs"].ctypes.data_as(C.POINTER(Bxdf))
            s.n_pool = a["float_pool"].shape[0]
            s.float_pool = a["float_pool"].ctypes.data_as(C.POINTER(C.c_float))
            self._struct = s
        return self._struct

    def pointer(self) -> "C.POINTER(FlatScene)":
        return C.pointer(self.struct())


_ACCEL_KEYS = ("n_prims", "n_unbounded", "n_nodes", "root", "root_count", "max_depth")


def _accel_head(a: Accel) -> dict:
    return {k: int(getattr(a, k)) for k in _ACCEL_KEYS}


def _accel_struct(h: dict, nodes: np.ndarray) -> Accel:
    a = Accel()
    for k in _ACCEL_KEYS:
        setattr(a, k, h[k])
    a.nodes = nodes.ctypes.data_as(C.POINTER(BvhNode))
    return a


def pre_construction_order(is_bounded) -> tuple[np.ndarray, np.ndarray]:
    """What `std::partition(first, last, is_bounded)` of internal::create_acceleration_structure (reference base/Scene.h:33)
    leaves behind for a primitive list in parser order: (positions of the bounded primitives in the order
    BVHAccelerator(first, part_it) receives them, positions of the unbounded ones in the order the top-level list keeps
    them).  libstdc++'s bidirectional partition is Hoare's scheme — the first misplaced element from the left trades places
    with the first from the right — so the result is NOT the parser order once a plane precedes a mesh.  This is the order
    `spcu_upload_scene_build` expects its primitive records in: unbounded first, then bounded."""
    flags = np.asarray(is_bounded, dtype=bool)
    pos = np.arange(flags.shape[0], dtype=np.uint32)
    k = int(flags.sum())
    left_false = np.flatnonzero(~flags[:k])          # misplaced in [0, k), ascending
    right_true = np.flatnonzero(flags[k:])[::-1] + k  # misplaced in [k, n), from the right
    assert left_false.shape == right_true.shape
    pos[left_false], pos[right_true] = right_true.astype(np.uint32), left_false.astype(np.uint32)
    return pos[:k].copy(), pos[k:].copy()


def unbuilt_scene_with_mesh(ctx, standin: FlatSceneData, vertices, faces, object_to_world, normal_xf) -> FlatSceneData:
    """A scene whose geometry arrives UNBUILT (spcu_upload_scene_build): camera, materials, lights and the top-level
    unbounded primitives of `standin` (a flattened scene holding ONE mesh), with that mesh replaced by (vertices, faces)
    under the given transform.  The mesh is ingested on the device (spcu_ingest_mesh: read_ply's normal passes + Mesh's
    constructor); the triangle records stay in face order, the pre-construction order of a scene whose only bounded
    primitives are this mesh's (reference base/Scene.h:29-45)."""
    nu = standin.head["geom"]["n_unbounded"]
    meta = standin.arrays["geom_meta"].view(np.uint32).reshape(-1)
    material = int(meta[nu] >> 2)                      # the stand-in mesh's material
    mesh = ctx.ingest_mesh(np.ascontiguousarray(vertices, dtype=np.float32), np.ascontiguousarray(faces, dtype=np.uint32),
                           np.asarray(object_to_world, dtype=np.float32), np.asarray(normal_xf, dtype=np.float32), material)
    k = mesh["prims"].shape[0]

    def cat(key, new, width):
        top = standin.arrays[key].view(np.uint8).reshape(-1, width)[:nu]
        return np.concatenate([top, np.ascontiguousarray(new).view(np.uint8).reshape(-1, width)])
    arrays = dict(standin.arrays)
    arrays["geom_prims"] = cat("geom_prims", mesh["prims"], 48)
    arrays["geom_shade"] = cat("geom_shade", mesh["shade"], 48)
    arrays["geom_meta"] = cat("geom_meta", mesh["meta"], 4)
    arrays["geom_nodes"] = np.zeros((0, C.sizeof(BvhNode)), dtype=np.uint8)
    head = dict(standin.head)
    head["geom"] = {"n_prims": nu + k, "n_unbounded": nu, "n_nodes": 0, "root": ~nu, "root_count": 0, "max_depth": 0}
    return FlatSceneData(head, arrays)
