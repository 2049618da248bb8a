"""Frame-sharded multi-GPU driver: one process per GPU, frames are independent units
(project_whole.py:71-84 and joint_localization_net.py:72 loop per frame, BatchNorm
is in eval mode), weights and calibrations are replicated, and the only collective is ONE all_gather of
the final ``fused_poses`` rows ([B/R,P,J,5] fp32, 3 KB per frame) - the multi-GPU form of
``torch.cat(all_fused_poses)`` in run/validate.py:114.  Backend: NCCL over NVLink on GPUs, gloo on CPU
(tests)."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

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


def _parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read()
    except OSError:
        return None


def gpu_numa_node(pci_bus_id: str, sysfs: str = "/sys") -> Optional[int]:
    """NUMA node the GPU at ``pci_bus_id`` ('0000:1b:00.0') hangs off, None when the kernel does not say (-1, no file)."""
    txt = _read(os.path.join(sysfs, "bus/pci/devices", pci_bus_id.lower(), "numa_node"))
    try:
        node = int(txt) if txt is not None else -1
    except ValueError:
        node = -1
    return node if node >= 0 else None


def local_gpu_numa_nodes(local_world: int, sysfs: str = "/sys") -> List[Optional[int]]:
    """NUMA node of cuda:0 .. cuda:local_world-1 (local rank i drives cuda:i), None where unknown."""
    nodes: List[Optional[int]] = []
    for i in range(local_world):
        try:
            p = torch.cuda.get_device_properties(i)
            nodes.append(gpu_numa_node("%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id), sysfs))
        except Exception:       # no such device / attributes missing: unknown
            nodes.append(None)
    return nodes


def _physical_cores(cpus: Sequence[int], sysfs: str) -> List[List[int]]:
    """Group logical CPUs into physical cores (hyper-thread siblings stay together), ordered by lowest CPU id."""
    allowed, seen, cores = set(cpus), set(), []
    for c in sorted(allowed):
        if c in seen:
            continue
        txt = _read(os.path.join(sysfs, "devices/system/cpu/cpu%d/topology/thread_siblings_list" % c))
        sib = [x for x in (_parse_cpulist(txt) if txt else [c]) if x in allowed] or [c]
        if c not in sib:
            sib.append(c)
        seen.update(sib)
        cores.append(sorted(sib))
    return cores


def plan_rank_cores(allowed: Sequence[int], local_world: int, rank_nodes: Sequence[Optional[int]],
                    sysfs: str = "/sys") -> List[List[int]]:
    """Disjoint host-core sets for the ``local_world`` ranks of a node, whole physical cores each.

    NUMA-aware when every rank's GPU node is known and owns at least 3/4 of an even share of ``allowed`` per rank: the ranks of
    a node share that node's cores, so a rank's pinned staging buffers (first touch) and its copy threads sit next to
    the PCIe root of its GPU - H2D traffic does not cross the socket interconnect.  Otherwise an even split of all
    allowed cores in rank order.  Returns [] per rank when there are fewer physical cores than ranks."""
    def split(cores: List[List[int]], n: int) -> List[List[int]]:
        per = len(cores) // n
        return [sorted(c for core in cores[i * per:(i + 1) * per] for c in core) for i in range(n)]

    phys = _physical_cores(allowed, sysfs)
    if local_world < 1 or len(phys) < local_world:
        return [[] for _ in range(max(0, local_world))]
    even = len(phys) // local_world                         # locality must not cost a rank more than a quarter of this
    if len(rank_nodes) == local_world and all(n is not None for n in rank_nodes):
        plan: List[Optional[List[int]]] = [None] * local_world
        ok = True
        for node in sorted(set(rank_nodes)):
            ranks = [r for r in range(local_world) if rank_nodes[r] == node]
            txt = _read(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node))
            node_cpus = set(_parse_cpulist(txt)) if txt else set()
            cores = [core for core in phys if core[0] in node_cpus]
            share = len(cores) // len(ranks)
            if share < 1 or 4 * share < 3 * even:           # e.g. a cpuset that leaves one socket a handful of cores
                ok = False
                break
            for r, mine in zip(ranks, split(cores, len(ranks))):
                plan[r] = mine
        if ok:
            return [p or [] for p in plan]
    return split(phys, local_world)


def pin_rank_to_cores(local_rank: int, local_world: int, rank_nodes: Optional[Sequence[Optional[int]]] = None,
                      sysfs: str = "/sys") -> list:
    """Give every rank of a node its own block of host cores, next to its GPU when the topology is known
    (:func:`plan_rank_cores`; 8 ranks x 7 pinned-copy pipelines otherwise migrate over all cores, contend, and their
    pinned buffers land on whichever socket the process started on).  Call it BEFORE allocating pinned memory.
    ``rank_nodes``: NUMA node of every local rank's GPU (default: asked from the driver / sysfs when CUDA is available).
    Returns the cores this process may now run on ([] when affinity is not available)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world <= 1:
            return cores
        if rank_nodes is None:
            rank_nodes = local_gpu_numa_nodes(local_world, sysfs) if torch.cuda.is_available() else [None] * local_world
        mine = plan_rank_cores(cores, local_world, rank_nodes, sysfs)[local_rank]
        if len(mine) >= 1 and len(cores) // local_world >= 2:
            os.sched_setaffinity(0, mine)
            return mine
        return cores
    except Exception:       # affinity unsupported, odd sysfs contents, ...: pinning is an optimisation, never an error
        return []


def sharded_forward(run_frames: Callable[[int, int], torch.Tensor], num_frames: int, rank: int, world: int) -> torch.Tensor:
    """``run_frames(lo, hi)`` computes fused_poses rows of frames [lo,hi) on this rank's GPU; returns all frames' rows
    on every rank."""
    lo, hi = shard_bounds(num_frames, rank, world)
    return gather_frames(run_frames(lo, hi), num_frames, rank, world)
