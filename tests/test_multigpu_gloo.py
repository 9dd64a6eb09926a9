"""The N > 1 path on CPU: two processes over gloo, each rendering its partition (the oracle's restatement stands in for
the device, it takes the same `spcu_partition`), accumulators reduced to rank 0 — against the single-process render of
the union.  Covers both partitionings of simplepath_b200.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, mode: str, out_path: str):
    import sys
    sys.path.insert(0, str(ROOT))
    from oracle import port as oracle_port
    from simplepath_b200 import distributed, rsequence
    from simplepath_b200.flat import FlatSceneData
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = FlatSceneData.load(GOLDEN / "g_example.flat.npz")
    spp = 2
    if mode == "samples":
        part = distributed.sample_partition(rank, world, spp, seed=31)
        jitter = rsequence.jitter_table(spp * world)
    else:
        part = distributed.tile_partition(rank, world, spp * world, seed=31)
        jitter = rsequence.jitter_table(spp * world)
    rgb, sq, st = oracle_port.render(flat.pointer(), jitter, part, threads=2)
    t_rgb, t_sq = torch.from_numpy(rgb), torch.from_numpy(sq)
    paths = torch.tensor([st["paths"]], dtype=torch.int64)
    distributed.reduce_to_root(t_rgb, t_sq, paths)
    if rank == 0:
        np.savez(out_path, rgb=t_rgb.numpy(), sq=t_sq.numpy(), paths=paths.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_ranks_reproduce_the_single_rank_render(tmp_path, oracle_port, mode):
    from simplepath_b200 import rsequence
    from simplepath_b200.capi import INTEGRATORS, Partition
    from simplepath_b200.flat import FlatSceneData
    world, spp = 2, 2
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    got = np.load(out)
    flat = FlatSceneData.load(GOLDEN / "g_example.flat.npz")
    total = spp * world
    whole = Partition(0, 1, 0, total, total, INTEGRATORS["iterative_rrnee"], 31)
    want, want_sq, st = oracle_port.render(flat.pointer(), rsequence.jitter_table(total), whole, threads=2)
    assert int(got["paths"][0]) == st["paths"] == flat.width * flat.height * total
    if mode == "tiles":   # disjoint pixels: exact
        assert got["rgb"].tobytes() == want.tobytes() and got["sq"].tobytes() == want_sq.tobytes()
    else:                 # same samples, added in a different order
        np.testing.assert_allclose(got["rgb"], want, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got["sq"], want_sq, rtol=1e-5, atol=1e-6)


def test_partition_arithmetic():
    from simplepath_b200 import distributed
    parts = [distributed.sample_partition(r, 4, 16) for r in range(4)]
    assert [(p.sample_begin, p.sample_end) for p in parts] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert all(p.spp_total == 64 and p.tile_stride == 1 for p in parts)
    tiles = [distributed.tile_partition(r, 4, 16) for r in range(4)]
    assert [(p.tile_offset, p.tile_stride) for p in tiles] == [(0, 4), (1, 4), (2, 4), (3, 4)]
    with pytest.raises(ValueError):
        distributed.sample_partition(4, 4, 16)
