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
