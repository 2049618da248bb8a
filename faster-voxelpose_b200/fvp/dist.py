"""Frame-sharded multi-GPU driver: one process per GPU, frames are independent units
(project_whole.py:71-84 and joint_localization_net.py:72 loop per frame, BatchNorm
is in eval mode), weights and calibrations are replicated, and the only collective is ONE all_gather of
the final ``fused_poses`` rows ([B/R,P,J,5] fp32, 3 KB per frame) - the multi-GPU form of
``torch.cat(all_fused_poses)`` in run/validate.py:114.  Backend: NCCL over NVLink on GPUs, gloo on CPU
(tests)."""
from __future__ import annotations

import os
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group if needed."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_bounds(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(shard: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """THE collective of the path: one all-gather of equally sized per-rank row blocks [n,...] into one contiguous
    [world*n,...] tensor (rank-major = frame order for contiguous frame blocks) - ``ncclAllGather`` on GPUs.  Under gloo
    (CPU tests, or several ranks sharing one GPU) the rows are staged through host memory: gloo has no CUDA all-gather."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * shard.shape[0],) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    if dist.get_backend() == "gloo" and shard.is_cuda:
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(host, shard.cpu().contiguous())
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, shard.contiguous())
    return out


def gather_frames(local_rows: torch.Tensor, num_frames: int, rank: int, world: int) -> torch.Tensor:
    """all_gather of per-rank result rows ([n_local, ...]) into frame order [num_frames, ...].
    Blocks may be ragged (num_frames % world != 0): rows are padded to the largest block."""
    if world == 1:
        return local_rows
    sizes = [shard_bounds(num_frames, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    out = gather_shards(pad).view((world, mx) + tuple(local_rows.shape[1:]))
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def pin_rank_to_cores(local_rank: int, local_world: int) -> list:
    """Give every rank of a node its own block of host cores (8 ranks x 7 pinned-copy pipelines otherwise migrate over
    all cores and contend).  Returns the cores this process may now run on ([] when affinity is not available)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(1, local_world)
        if local_world > 1 and per >= 2:
            mine = cores[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            return mine
        return cores
    except (AttributeError, OSError):
        return []


def sharded_forward(run_frames: Callable[[int, int], torch.Tensor], num_frames: int, rank: int, world: int) -> torch.Tensor:
    """``run_frames(lo, hi)`` computes fused_poses rows of frames [lo,hi) on this rank's GPU; returns all frames' rows
    on every rank."""
    lo, hi = shard_bounds(num_frames, rank, world)
    return gather_frames(run_frames(lo, hi), num_frames, rank, world)
