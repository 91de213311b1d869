"""Data parallelism over robots: every robot's MPC+WBC is independent, so rank g owns the contiguous
robot range [g*B/G, (g+1)*B/G) and the ONLY collective is one all-gather of the torque buffer per tick
(torch.distributed: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` robots for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend=None):
    """Read RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment (torchrun); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allgather_torque(local, total, rank, world):
    """All-gather the per-rank torque rows [B_local, C] into [total, C] in original robot order (one collective).
    Shards may differ by one robot, so ranks pad to the largest shard."""
    if world == 1:
        return local
    sizes = [shard_range(total, r, world) for r in range(world)]; mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def max_over_ranks(value, device):
    """Max of a python float over ranks (bench timing rule)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def init_comm(solver, rank, world):
    """Bootstrap of the library's own NCCL communicator (include/qmb200.h, qmb200_comm_*): rank 0 creates the ncclUniqueId, torch.distributed only carries the
    128 bytes to the other ranks; every collective of the data path is then issued by the C++ host (qmb200_allgather_torque).  No-op for one rank."""
    if world == 1:
        return
    ids = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    solver.comm_init(world, rank, ids[0])
