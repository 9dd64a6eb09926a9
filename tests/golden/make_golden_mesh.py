"""Regenerates tests/golden/mesh_ingest.npz from the REAL reference: its own read_ply (base/PlyReader.cpp) + Mesh
constructor (shapes/Triangle.h:25-51) applied to the PLY of tests/meshcases.py — world-space vertices and normals, the index
list of the kept faces, and the normal matrix it applies.  Development container only:
    python tests/golden/make_golden_mesh.py"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref  # noqa: E402
from simplepath_b200 import scenes  # noqa: E402
import meshcases  # noqa: E402


def main() -> None:
    if not ref.available():
        raise SystemExit("oracle/_ref/libsp_ref.so missing")
    v, f = meshcases.mesh()
    xf = meshcases.transform()
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "m.ply"
        scenes.write_ply(path, v, f)
        r = ref.read_ply(path, xf, len(v), len(f))
    np.savez_compressed(Path(__file__).resolve().parent / "mesh_ingest.npz", in_vertices=v, in_faces=f, object_to_world=xf,
                        vertices=r["vertices"], normals=r["normals"], indices=r["indices"], normal_xf=r["normal_xf"])
    print(len(v), "vertices", len(f), "faces ->", len(r["indices"]), "triangles kept")


if __name__ == "__main__":
    main()


def accel_cases(n_triangles: int) -> dict:
    """Positions of the unbounded stand-ins (planes) in the parser's primitive list, per case."""
    return {"plane_first": [0], "planes_around": [0, n_triangles + 1], "planes_inside": [5, 100, n_triangles + 2], "no_plane": []}


def main_accel() -> None:
    """tests/golden/scene_accel.npz: internal::create_acceleration_structure (base/Scene.h:27-45) run by the reference over
    the real triangles of the mesh above (read by its own read_ply) and unbounded stand-ins at the positions of accel_cases."""
    v, f = meshcases.mesh()
    xf = meshcases.transform()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "m.ply"
        scenes.write_ply(path, v, f)
        nt = len(ref.read_ply(path, xf, len(v), len(f))["indices"])
        for name, where in accel_cases(nt).items():
            ub = np.zeros(nt + len(where), dtype=np.uint8)
            ub[where] = 1
            r = ref.accel_from_mesh(path, xf, ub)
            out[f"{name}.unbounded"] = ub
            out[f"{name}.order"] = r["order"]
            out[f"{name}.nodes"] = r["nodes"].view(np.uint8).reshape(-1, 64)
            out[f"{name}.head"] = np.array([r["head"][k] for k in ("n_prims", "n_unbounded", "n_nodes", "root", "root_count", "max_depth")],
                                           dtype=np.int64)
            print(name, r["head"])
    np.savez_compressed(Path(__file__).resolve().parent / "scene_accel.npz", **out)


def main_stl() -> None:
    """tests/golden/mesh_ingest_stl.npz: the reference's own read_stl (base/STLReader.cpp) + Mesh on the binary STL of
    meshcases.stl_soup()."""
    corners, stored = meshcases.stl_soup()
    xf = meshcases.transform(seed=4)
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "m.stl"
        meshcases.write_stl(path, corners, stored)
        r = ref.read_ply(path, xf, 3 * len(corners), len(corners))
    np.savez_compressed(Path(__file__).resolve().parent / "mesh_ingest_stl.npz", corners=corners, stored_normals=stored,
                        object_to_world=xf, vertices=r["vertices"], normals=r["normals"], indices=r["indices"], normal_xf=r["normal_xf"])
    print("stl:", len(corners), "triangles ->", len(r["vertices"]), "vertices,", len(r["indices"]), "triangles in the mesh")


if __name__ == "__main__":
    main_accel()
    main_stl()
