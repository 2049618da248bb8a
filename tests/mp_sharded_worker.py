"""Worker of tests/test_gpu_multirank.py: one rank of a frame-sharded run with the REAL engine (BASELINE configs[2] in small).

Launched once per rank with torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT).  Every rank
runs its contiguous block of frames through fvp.engine.EngineLanes (rows written straight into the rank's shard buffer),
then the single collective of the path (fvp.dist.gather_frames -> one all_gather_into_tensor).  Rank 0 also runs ALL frames
on one context and requires the gathered tensor to be bit-identical (run/validate.py:114: torch.cat of per-batch results).
Backend: NCCL with one GPU per rank when the box has enough GPUs, otherwise all ranks share cuda:0 and the rows travel
through gloo (NCCL refuses two ranks on one device)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (os.path.join(ROOT, "tests"), os.path.join(PKG, "lib"), PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def frame(base: np.ndarray, index: int) -> np.ndarray:
    """Deterministic function of the GLOBAL frame index (what every rank must agree on)."""
    rng = np.random.default_rng(1000 + index)
    noise = (rng.random(base.shape, dtype=np.float32) < 0.02).astype(np.float32) * 0.25
    return np.clip(base + noise, 0.0, 1.0).astype(np.float32)


def main() -> int:
    num_frames = int(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    from golden_util import Golden
    from fvp import dist as fdist
    from fvp.engine import Engine, EngineLanes
    nccl = torch.cuda.device_count() >= world
    dev = torch.device("cuda", local if nccl else 0)
    torch.cuda.set_device(dev)
    fdist.init_from_env("nccl" if nccl else "gloo")
    g = Golden("panoptic_256x192")
    base = g.heatmaps[0]
    lanes = EngineLanes(g.cfg, dev, lanes=2, max_batch=1, max_sequences=1, axes=g.axes)
    lanes.load_state_dict(g.weights)
    slot = lanes.sequence_slot(g.cams, g.resize)
    lanes.use_cuda_graph(True)
    lo, hi = fdist.shard_bounds(num_frames, rank, world)
    shard = torch.zeros((hi - lo, g.P, g.J, 5), device=dev)
    for i in range(lo, hi):
        lanes.submit(torch.from_numpy(frame(base, i)[None]).to(dev), [slot], out_fused=shard[i - lo:i - lo + 1])
        if lanes.outstanding() == len(lanes):
            lanes.collect()
    while lanes.outstanding():
        lanes.collect()
    torch.cuda.current_stream(dev).synchronize()
    lanes.engines[0].check_range()
    everything = fdist.gather_frames(shard, num_frames, rank, world)       # THE collective
    assert tuple(everything.shape) == (num_frames, g.P, g.J, 5)
    ok = True
    if rank == 0:
        eng = Engine(g.cfg, dev, max_batch=1, max_sequences=1, axes=g.axes)
        eng.load_state_dict(g.weights)
        s1 = eng.sequence_slot(g.cams, g.resize)
        single = torch.cat([eng.forward(torch.from_numpy(frame(base, i)[None]).to(dev), [s1])[0] for i in range(num_frames)])
        ok = bool(torch.equal(single, everything))
        valid = int((single[:, :, 0, 3] >= 0).sum())
        print("rank 0: %d frames over %d ranks (%s), %d valid people, gathered == single-rank: %s" % (
            num_frames, world, "nccl" if nccl else "gloo on one GPU", valid, ok), flush=True)
        ok = ok and valid > 0
        eng.close()
    flag = torch.tensor([1 if ok else 0])
    if nccl:
        flag = flag.to(dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    lanes.close()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
