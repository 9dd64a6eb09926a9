"""Python mirror of the host side of the drop-in: loads `.sp` scenes through the reference's own parser
(libsphost.so = reference objects + the scene flattener, built by simplepath_b200/host/Makefile where the reference
sources exist) and hands out flattened scenes for the C-ABI.  Mirrors what sp::CudaIntegrator does in C++."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import rsequence, scenes
from .capi import FlatScene
from .flat import FlatSceneData

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "host" / "libsphost.so"
DRIVER = PKG / "host" / "SimplePathCuda"
FLAT_DIR = PKG.parent / "scenes" / "flat"   # committed flattened scenes of the small workloads

_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        from . import capi
        capi.load()  # libsphost.so links against libspcu.so
        l = C.CDLL(str(LIB_PATH))
        l.sphost_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        l.sphost_load.restype = C.c_void_p
        l.sphost_free.argtypes = [C.c_void_p]
        l.sphost_free.restype = None
        l.sphost_flat.argtypes = [C.c_void_p]
        l.sphost_flat.restype = C.POINTER(FlatScene)
        l.sphost_output_file_name.argtypes = [C.c_void_p]
        l.sphost_output_file_name.restype = C.c_char_p
        l.sphost_jitter.argtypes = [C.c_uint, C.c_void_p]
        l.sphost_jitter.restype = None
        _lib = l
    return _lib


def parse(sp_path: Path | str) -> FlatSceneData:
    """FileParser -> Scene -> flattener, then copied into numpy arrays (the C++ objects are freed)."""
    l = lib()
    err = C.create_string_buffer(512)
    h = l.sphost_load(str(sp_path).encode(), err, 512)
    if not h:
        raise RuntimeError(f"{sp_path}: {err.value.decode()}")
    try:
        return FlatSceneData.from_struct(l.sphost_flat(h))
    finally:
        l.sphost_free(h)


def jitter(spp: int) -> np.ndarray:
    """Pixel jitter table; the pure-Python restatement is bitwise equal to the reference's (tests/test_host_logic.py)."""
    return rsequence.jitter_table(spp)


def workload(name: str) -> FlatSceneData:
    """A named scene of simplepath_b200.scenes, flattened.  Uses the parser library when it is present (development
    container, or a GPU box that received the prebuilt library), else the committed flattened copy."""
    if available():
        return parse(scenes.ensure(name))
    path = FLAT_DIR / f"{name}.flat.npz"
    if path.exists():
        return FlatSceneData.load(path)
    raise RuntimeError(f"cannot load workload {name}: {LIB_PATH} is not built and {path} does not exist")
