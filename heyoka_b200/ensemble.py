"""Multi-GPU sharding of a batch: one process per GPU, lanes split in contiguous blocks, no collective inside
the integration, one all_gather of the final state at the end.

The reference's counterpart is ensemble_propagate_*_batch() (src/ensemble_propagate.cpp:192-311): n_iter
independent integrators under a TBB parallel_for, results collected in a std::vector. Lanes never interact, so
the only exchange here is the final gather (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def lane_shard(n_lanes, rank, world):
    """Contiguous block of lanes owned by `rank`: [begin, end). The first n_lanes % world ranks get one more."""
    base, rem = divmod(int(n_lanes), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_shards(n_lanes, world):
    return [lane_shard(n_lanes, r, world) for r in range(world)]


def gather_lanes(local, n_lanes, group=None):
    """all_gather of a lane-sharded torch tensor [..., n_local] into [..., n_lanes] (same result on every rank).

    Shards may differ by one lane: every rank pads to the largest shard, the padding is dropped after the gather.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    shards = all_shards(n_lanes, world)
    width = max(e - b for b, e in shards)
    pad = torch.zeros(local.shape[:-1] + (width,), dtype=local.dtype, device=local.device)
    pad[..., :local.shape[-1]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[..., :e - b] for o, (b, e) in zip(out, shards)], dim=-1)
