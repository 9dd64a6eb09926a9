"""Multi-GPU plumbing of the hot path: one process per GPU, scene replicated, work partitioned, per-pixel accumulators
summed to rank 0.  The partition is pure arithmetic and the reduction is a single collective, so both
are testable on CPU with the gloo backend (tests/test_multigpu_gloo.py)."""
from __future__ import annotations

from .capi import INTEGRATORS, Partition


def sample_partition(rank: int, world: int, spp_per_rank: int, integrator: str = "iterative_rrnee",
                     seed: int = 0) -> Partition:
    """Rank r renders global samples [r * spp, (r + 1) * spp) of world * spp for EVERY pixel.  The random numbers are
    keyed by (pixel, global sample index), so the union over ranks is exactly the world*spp-sample render."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return Partition(0, 1, rank * spp_per_rank, (rank + 1) * spp_per_rank, world * spp_per_rank, INTEGRATORS[integrator], seed)


def tile_partition(rank: int, world: int, spp: int, integrator: str = "iterative_rrnee", seed: int = 0) -> Partition:
    """Rank r renders all samples of the 8x8 tiles t with t % world == r (TileScheduler order, base/TileScheduler.h:66-82).
    Disjoint pixels: the reduction adds zeros and the result is bit-identical to one rank's."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return Partition(rank, world, 0, spp, spp, INTEGRATORS[integrator], seed)


def init_product_comm(ctx) -> None:
    """Form the product library's own NCCL communicator over the ranks of the torch.distributed job: rank 0 draws the
    ncclUniqueId (spcu_comm_unique_id), torch.distributed only CARRIES its 128 bytes to the other ranks (what mpirun or a
    shared file would do), every rank joins with spcu_comm_init_rank.  The frame's reduction then runs inside libspcu.so
    (spcu_reduce_to_root / spcu_render_frame_reduced), not in Python."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    box = [ctx.comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init_rank(dist.get_world_size(), dist.get_rank(), box[0])


def reduce_to_root(*tensors) -> None:
    """Sum the accumulators of all ranks into rank 0 (NCCL over NVLink for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
