"""Multi-GPU plumbing of the kinetic path: particles are partitioned by index over the ranks, meshes
are replicated, and the only exchange is the additive deposit (NCCL allreduce inside libstarfish_gpu.so).
torch.distributed is used for the rendezvous only (any backend: nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous, balanced slice [first, first+count) of particle indices owned by `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_total), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def exchange_unique_id(make_id):
    """Rank 0 creates the 128-byte NCCL unique id (sfgpu_comm_unique_id), everybody receives it."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def attach_communicator(km):
    """Give a KineticMaterial its NCCL communicator: one rank per GPU of the default process group."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    uid = exchange_unique_id(type(km).commUniqueId)
    km.commInit(dist.get_world_size(), dist.get_rank(), uid)
