"""Host-side pieces that need no GPU: the jitter table, flattened-scene containers, scene generators."""
import numpy as np

from conftest import GOLDEN
from simplepath_b200 import rsequence, scenes
from simplepath_b200.flat import FlatSceneData


def test_jitter_table_is_the_reference_sequence_bitwise():
    want = np.load(GOLDEN / "jitter.npz")["jitter4096"]
    for spp in (1, 4, 16, 64, 256, 4096):
        got = rsequence.jitter_table(spp)
        assert got.dtype == np.float32 and got.shape == (spp, 2)
        assert got.tobytes() == want[:spp].tobytes()
    assert ((want >= 0) & (want < 1)).all()


def test_flat_scene_round_trip(tmp_path):
    flat = FlatSceneData.load(GOLDEN / "g_bunny.flat.npz")
    flat.save(tmp_path / "copy.npz")
    again = FlatSceneData.load(tmp_path / "copy.npz")
    assert again.head == flat.head
    for k in FlatSceneData.ARRAYS:
        assert again.arrays[k].tobytes() == flat.arrays[k].tobytes()
    s = again.struct()
    assert s.geom.n_prims == flat.n_prims and s.n_lights == 1


def test_flat_scene_invariants():
    """Reference-order IDs: unbounded prims first, leaves cover [n_unbounded, n_prims) exactly once, left to right."""
    for name in ("g_bunny", "g_elf"):
        flat = FlatSceneData.load(GOLDEN / f"{name}.flat.npz")
        g = flat.head["geom"]
        nodes = flat.arrays["geom_nodes"].view(np.int32).reshape(-1, 16)
        child = nodes[:, 12:14]
        count = nodes[:, 14:16].view(np.uint32) & 0x7FFFFFFF
        leaves = []

        def walk(link, cnt):
            if link < 0:
                leaves.append((~link, cnt))
                return
            for k in range(2):
                walk(int(child[link, k]), int(count[link, k]))

        walk(g["root"], g["root_count"] & 0x7FFFFFFF)
        pos = g["n_unbounded"]
        for first, cnt in leaves:
            assert first == pos
            pos += cnt
        assert pos == g["n_prims"]


def test_generated_scenes_are_deterministic(tmp_path):
    a = scenes.ensure("g_bunny", tmp_path / "a")
    b = scenes.ensure("g_bunny", tmp_path / "b")
    assert a.read_text() == b.read_text()
    pa = next((tmp_path / "a").glob("*.ply")).read_bytes()
    pb = next((tmp_path / "b").glob("*.ply")).read_bytes()
    assert pa == pb
    v, f = scenes.bumpy_sphere(scenes.BUNNY_TRIS, scenes.BUNNY_LO, scenes.BUNNY_HI)
    assert f.shape == (scenes.BUNNY_TRIS, 3)


def test_scene_assets_can_be_generated_by_concurrent_ranks(tmp_path):
    """Every rank of a torchrun job calls scenes.ensure() on a fresh box: writers must not trip over each other's files."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(6) as pool:
        texts = pool.starmap(_ensure_text, [("g_bunny", str(tmp_path))] * 6)
    assert len(set(texts)) == 1
    assert not [p for p in tmp_path.iterdir() if p.name.endswith(".tmp")]


def _ensure_text(name, out):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from simplepath_b200 import scenes
    return scenes.ensure(name, out).read_text()
