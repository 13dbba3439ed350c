"""Clip-level data parallelism: one process per GPU, clips sharded contiguously, ONE gather of finished
motions to rank 0 (SURVEY.md sections 2a, 8(e)).  The reference has no counterpart (its dist_util is dead code,
reference main/utils/dist_util.py:18-67); inside the sampling loop there is nothing to exchange, so no
collective is fused into any kernel.  NCCL on GPUs, gloo in the CPU tests."""
import os

import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous split, first ``total % world`` ranks take one extra clip: returns [lo, hi)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend=None):
    """Join the process group torchrun described (RANK / WORLD_SIZE / MASTER_*); no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1, 0
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    return rank, world, local


def gather_motions(local, total, dst=0):
    """local: [B_local, n, J] on this rank's device -> [total, n, J] on rank ``dst`` (None elsewhere): ONE ``gather`` (NCCL:
    ncclSend/ncclRecv pairs towards ``dst``; gloo in the CPU tests) straight from the device buffer the sampler wrote, before
    any device->host copy.  Ranks may hold different clip counts (``shard_bounds``): shards are padded to the largest and
    trimmed on ``dst``; only ``dst`` allocates the [world * bmax, n, J] receive buffer."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_bounds(total, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in counts)
    if local.shape[0] == bmax:
        send = local.contiguous()
    else:
        send = torch.zeros((bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    if rank != dst:
        dist.gather(send, None, dst=dst)
        return None
    out = torch.empty((world, bmax) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.gather(send, [out[r] for r in range(world)], dst=dst)
    if all(hi - lo == bmax for lo, hi in counts):
        return out.view((world * bmax,) + tuple(local.shape[1:]))
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(counts)], dim=0)


def barrier_max_ms(ms, device=None):
    """Max over ranks of a per-rank elapsed time (device-timed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
