"""The C-ABI library loads and exports every entry point include/spcu.h declares (no compute calls: runs without a GPU)."""
import ctypes
import re

from conftest import ROOT


def declared_functions():
    text = (ROOT / "include" / "spcu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spcu_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("spcu_create", "spcu_upload_scene", "spcu_trace_closest", "spcu_trace_any", "spcu_trace_lights",
                 "spcu_generate_rays", "spcu_render", "spcu_render_device"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    from simplepath_b200 import capi
    lib = ctypes.CDLL(str(capi.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"libspcu.so lacks {missing}"
    assert lib.spcu_abi_version() == capi.ABI_VERSION


def test_python_mirror_matches_header_sizes(built):
    """ctypes structs must have the C layout: compile a tiny C probe against the header and compare sizeof()."""
    import subprocess
    import tempfile
    from simplepath_b200 import capi
    probe = r'''
    #include "spcu.h"
    #include <stdio.h>
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(spcu_ray), sizeof(spcu_hit),
        sizeof(spcu_bvh_node), sizeof(spcu_prim_geom), sizeof(spcu_accel), sizeof(spcu_bxdf), sizeof(spcu_material),
        sizeof(spcu_light), sizeof(spcu_flat_scene), sizeof(spcu_partition), sizeof(spcu_stats), sizeof(spcu_bounds),
        sizeof(spcu_prim_shade)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        src = f"{d}/p.c"
        open(src, "w").write(probe)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), src, "-o", f"{d}/p"])
        sizes = [int(x) for x in subprocess.check_output([f"{d}/p"]).split()]
    mine = [ctypes.sizeof(t) for t in (capi.Ray, capi.Hit, capi.BvhNode, capi.PrimGeom, capi.Accel, capi.Bxdf,
                                       capi.Material, capi.Light, capi.FlatScene, capi.Partition, capi.Stats)]
    mine += [capi.BOUNDS_DTYPE.itemsize, ctypes.sizeof(capi.PrimShade)]
    assert sizes == mine
    assert capi.NODE_DTYPE.itemsize == ctypes.sizeof(capi.BvhNode) and capi.RAY_DTYPE.itemsize == ctypes.sizeof(capi.Ray)


def test_no_device_is_an_error_not_a_fallback(built):
    """Without a GPU spcu_create must fail loudly."""
    import torch
    from simplepath_b200 import capi
    if torch.cuda.is_available():
        return
    try:
        capi.Context(0)
    except capi.SpcuError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("Context() succeeded without a CUDA device")
