#!/usr/bin/env python
"""Benchmark of the Faster-VoxelPose hot path on B200 (contract: see the task brief / DESIGN.md "measurement").

    python bench.py --gpus N --steps K --warmup W          # ours, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference ...                   # the reference's CPU path (oracle port) on host cores

One *step* = one forward of the whole hot path (K0..finalize) over one batch of synthetic frames per rank (default
batch 1 = BASELINE.json configs[1]: Panoptic 5-view jln64, 80x80x20 voxel, batch=1).  With N > 1 ranks the frames of a
rank form shards of `--shard` (32) frames - BASELINE.json configs[2]: 256 frames over 8 GPUs = 32 frames per rank - and
every shard ends in THE collective of the path: one `all_gather_into_tensor` of the rank's [32,P,J,5] pose rows into one
contiguous [N*32,P,J,5] tensor (run/validate.py:114 `torch.cat`), issued on a side stream so the next shard's kernels
run under it.  Per-GPU work is the same at every N (weak scaling).
`value` = frames/s of the whole job with the inputs already in HBM; `e2e` = the same through the host-buffer C-ABI
entry (fvp_submit_host / fvp_wait: H2D of every step's heat maps from pinned memory + forward + D2H of the step's
result tensors inside the timed region).  Prints exactly ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (PKG, os.path.join(PKG, "lib"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "3D-pose FPS (5-view->80x80x20 voxel)"
UNIT = "frames/s"
POOL = 12                # distinct input frames rotated through: 12 x 14.7 MB = 177 MB > 126 MB L2


def workload(preset: str):
    from fvp import config as fcfg, synth
    cfg = fcfg.preset(preset)
    cfg.CAPTURE_SPEC.MIN_SCORE = -1.0e30          # every proposal slot valid: worst case N = P persons
    if preset.startswith("ring8"):
        cams = synth.ring_cameras(8, cfg.CAPTURE_SPEC.SPACE_CENTER)
    elif preset in ("campus", "shelf"):           # BASELINE configs[0] / [3]: dataset geometry, synthetic ring calibration
        V = int(cfg.DATASET.CAMERA_NUM)
        ow, oh = [float(v) for v in cfg.DATASET.ORI_IMAGE_SIZE]
        far = preset == "campus"
        cams = synth.ring_cameras(V, cfg.CAPTURE_SPEC.SPACE_CENTER, radius=9000.0 if far else 5000.0,
                                  height=3000.0 if far else 2600.0, f=1.22 * ow if far else 1.07 * ow, cx=ow / 2, cy=oh / 2,
                                  k=(0.0, 0.0, 0.0) if far else (-0.1, 0.02, 0.0))
    else:
        cal = os.path.join(ROOT, "tests", "golden", "panoptic_256x192.npz")
        cams = synth.cameras_from_array(np.load(cal)["cameras"])       # Panoptic HD cameras (demo/calibration.json)
    resize = synth.resize_transform(cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE)
    return cfg, cams, resize


def make_frames(cfg, cams, count: int, seed0: int) -> np.ndarray:
    """[count,V,J,H,W] Gaussian-blob heat maps of P synthetic people per frame (seed = seed0 + frame)."""
    from fvp import synth
    P = int(cfg.CAPTURE_SPEC.MAX_PEOPLE)
    out = []
    for i in range(count):
        sk = synth.make_skeletons(cfg, P, seed=seed0 + i)
        out.append(synth.render_heatmaps(cfg, cams, sk, sigma=float(cfg.NETWORK.SIGMA)))
    return np.stack(out)


def workload_config(cfg, preset: str, batch: int, world: int, shard: int) -> dict:
    """The `config` object of the JSON line: identical keys (and, for one command line, values) in both arms."""
    J, P, V = int(cfg.DATASET.NUM_JOINTS), int(cfg.CAPTURE_SPEC.MAX_PEOPLE), int(cfg.DATASET.CAMERA_NUM)
    W, H = [int(v) for v in cfg.DATASET.HEATMAP_SIZE]
    X, Y, Z = [int(v) for v in cfg.CAPTURE_SPEC.VOXELS_PER_AXIS]
    name = {"panoptic_256x192": "Panoptic 5-view jln64", "panoptic": "Panoptic 5-view jln64", "campus": "Campus 3-view jln64",
            "shelf": "Shelf 5-view jln64 (10-person crowd)", "ring8_160": "8-view synthetic ring"}.get(preset, preset)
    which = "BASELINE configs[1]: batch=1 inference on 1 GPU" if world == 1 else \
        "BASELINE configs[2]: frames sharded over %d GPUs, %d frames per rank per gather" % (world, shard * batch)
    return {"workload": "%s geometry, %dx%d synthetic heat maps, %dx%dx%d coarse grid, batch=%d per forward call, P=%d people "
                        "all valid (worst case), random-conditioned weights; %s" % (name, W, H, X, Y, Z, batch, P, which),
            "preset": preset, "batch_per_gpu": batch, "people": P, "views": V, "joints": J,
            "l2": "inputs rotate through %d distinct frames (%.0f MB) > 126 MB L2" % (POOL, POOL * 4.0 * V * J * H * W / 1e6),
            "parallelism": "frame-sharded x%d" % world, "frames_per_gather": shard * batch if world > 1 else None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_fps(cfg, cams, resize, frames: np.ndarray, sd_np, warm: int, steps: int):
    """The reference's CPU path (oracle port: the same PyTorch-CPU ops in the same order as the reference,
    pinned bit-exact to it by oracle/gen_golden.py) on this host's cores, one frame per step."""
    from oracle import fvp_oracle as O
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    rz = torch.as_tensor(resize, dtype=torch.float)
    cameras = {"s": cams}
    # the tiny per-person convolutions of this path scale badly past a few dozen threads (128 threads are ~20x
    # slower than 16 on the B200 host): give the baseline its best thread count among {8,16,32,all}
    ncpu = os.cpu_count() or 1
    best = (None, 1e30)
    with torch.no_grad():
        for nt in sorted({min(8, ncpu), min(16, ncpu), min(32, ncpu), ncpu}):
            torch.set_num_threads(nt)
            hm = torch.from_numpy(frames[0][None])
            O.forward(cfg, sd, hm, ["s"], cameras, rz, taps=False)
            t0 = time.perf_counter()
            O.forward(cfg, sd, hm, ["s"], cameras, rz, taps=False)
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (nt, dt)
            if dt > 4 * best[1]:
                break
    torch.set_num_threads(best[0])
    ts = []
    with torch.no_grad():
        for i in range(warm + steps):
            hm = torch.from_numpy(frames[i % frames.shape[0]][None])
            t0 = time.perf_counter()
            O.forward(cfg, sd, hm, ["s"], cameras, rz, taps=False)
            ts.append(time.perf_counter() - t0)
    ts = ts[warm:]
    return len(ts) / sum(ts), torch.get_num_threads(), float(np.median(ts)) * 1e3


def gpu_reference_port_fps(cfg, cams, resize, frames: np.ndarray, sd_np, dev, warm: int = 3, steps: int = 10) -> dict:
    """Baseline leg (like cpu_baseline, never the product): the reference's own op sequence as PyTorch CUDA ops on this
    GPU - cached sample grids, cudnn.benchmark (tools/ref_gpu_port.py) - i.e. what `DEVICE='cuda:0'` of the reference
    does (run/validate.py:61-63).  Timed twice: with PyTorch's STOCK precision switches (cuDNN convolutions may use TF32,
    `torch.backends.cudnn.allow_tf32 = True` is the default the reference never touches) = `value`, the figure the
    north-star target "whole pipeline >= 10x the reference's 1-GPU PyTorch path" is measured against, and with TF32 off
    (`fp32`), the arithmetic our kernels are held to."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_gpu_port
    dev = torch.device(dev)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    saved = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    def timed(cudnn_tf32: bool) -> float:
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = False        # PyTorch's default since 1.12; the reference has no matmul on this path
        torch.backends.cudnn.allow_tf32 = cudnn_tf32
        ref = ref_gpu_port.CachedReference(cfg, sd, cams, torch.as_tensor(resize, dtype=torch.float), dev)
        pool = [torch.from_numpy(frames[i][None]).to(dev) for i in range(min(4, frames.shape[0]))]
        for i in range(warm):
            ref.forward(pool[i % len(pool)])
        sync()
        t0 = time.perf_counter()
        for i in range(steps):
            ref.forward(pool[i % len(pool)])
        sync()
        dt = time.perf_counter() - t0
        del ref
        return dt

    try:
        dt_stock = timed(True)
        dt_fp32 = timed(False)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return {"value": steps / dt_stock, "unit": UNIT, "ms_per_step": dt_stock / steps * 1e3, "steps": steps, "warmup": warm,
            "batch": 1, "fp32": {"value": steps / dt_fp32, "ms_per_step": dt_fp32 / steps * 1e3},
            "what": "reference op sequence (oracle port) as PyTorch %s CUDA ops on this GPU: cached sample grids, "
                    "cudnn.benchmark, per-person host syncs as in project_individual.py; value = PyTorch's stock switches "
                    "(cudnn.allow_tf32 on, as the reference runs), fp32 = TF32 off" % torch.__version__}


def conv_tensor_rooflines(stage_ms: dict, batch: int, people_per_frame: int, J: int, grid_xy, peak_tflops: float) -> dict:
    """Tensor-pipe figures of the two conv trunks from their CUDA-event stage times.  useful = 2 x MACs of the reference
    layers (fvp.netspec); executed = 3 x useful: the fp16 hi/lo split issues A_hi*B_hi, A_hi*B_lo and
    A_lo*B_hi (csrc/fvp_conv_tc.cu).  peak = dense bf16/fp16 TFLOP/s measured on this pool (MEASURED_PEAKS.json)."""
    from fvp import netspec
    out = {}
    work = {"p2p_net": netspec.macs_per_image(netspec.p2p_net(J), (64, 64)) * 3 * people_per_frame * batch,
            "center_net": netspec.macs_per_image(netspec.center_net(J), tuple(grid_xy)) * batch}
    for name, macs in work.items():
        ms = float(stage_ms[name])
        useful = 2.0 * macs / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        out[name] = {"bound": "tensor", "ms": ms, "gmac": macs / 1e9, "useful_tflops": useful, "executed_tflops": 3.0 * useful,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac_useful": useful / peak_tflops,
                     "frac_executed": 3.0 * useful / peak_tflops}
    return out


def _emit(line: dict) -> None:
    """Exactly one JSON line on the real stdout (library chatter such as NCCL's version banner goes to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)          # anything a library prints to fd 1 from here on lands on stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--batch", type=int, default=1, help="frames per forward call (1 = BASELINE configs[1])")
    ap.add_argument("--shard", type=int, default=32, help="forward calls per all-gather when --gpus > 1 (configs[2]: 32 frames per rank)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="panoptic_256x192")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--conv-mode", type=int, default=-1, help="-1 library default, 0 fp32 CUDA cores, 1 tcgen05 3xTF32")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU / GPU reference-port baseline legs")
    ap.add_argument("--min-time", type=float, default=0.5, help="repeat the timed K-step region until it covers this many seconds")
    ap.add_argument("--lanes", type=int, default=6, help="frames in flight on one GPU (lane contexts sharing one weight set and "
                    "one grid cache, fvp.engine.EngineLanes); 1 = strictly serial forwards")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cfg, cams, resize = workload(args.preset)
    from fvp import synth
    J, P, V = int(cfg.DATASET.NUM_JOINTS), int(cfg.CAPTURE_SPEC.MAX_PEOPLE), int(cfg.DATASET.CAMERA_NUM)
    W, H = [int(v) for v in cfg.DATASET.HEATMAP_SIZE]
    sd_np = synth.make_weights(J, seed=2024)
    conf = workload_config(cfg, args.preset, args.batch, world, args.shard)

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        frames = make_frames(cfg, cams, 4, seed0=5000)
        fps, cores, med_ms = cpu_reference_fps(cfg, cams, resize, frames, sd_np, args.warmup, args.steps)
        sample = "%d forwards of one frame each (after %d warm-up), best of {8,16,32,all} = %d host threads" % (args.steps, args.warmup, cores)
        _emit(({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": conf,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is pure Python and cannot travel to the GPU box: timed is the oracle port (oracle/fvp_oracle.py)"}))
        return 0

    # ------------------------------------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    import torch.distributed as dist
    from fvp import dist as fdist
    from fvp.engine import EngineLanes
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # ranks pin themselves to physical cores of their GPU's NUMA node BEFORE any pinned buffer is allocated (first touch)
    gpu_nodes = fdist.local_gpu_numa_nodes(local_world) if world > 1 else []
    my_cores = fdist.pin_rank_to_cores(local, local_world, gpu_nodes) if world > 1 else []
    if world > 1:
        fdist.init_from_env("nccl")
    B = args.batch
    L = max(1, args.lanes)
    S = max(1, args.shard)
    lanes = EngineLanes(cfg, dev, lanes=L, max_batch=B, max_sequences=1)
    lanes.load_state_dict(sd_np)
    if args.conv_mode >= 0:
        lanes.set_conv_mode(args.conv_mode)
    slot = lanes.sequence_slot(cams, resize)
    slots = [slot] * B
    eng = lanes.engines[0]                       # lane 0 alone: serial latency and per-stage times
    pipelining = ("%d independent batch-%d forwards in flight on %d streams (lane contexts: own workspaces + CUDA graph, shared "
                  "weights and sample grids); per-frame latency is reported as serial_ms_per_step" % (L, B, L)) if L > 1 \
        else "none (serial forwards)"
    frames = make_frames(cfg, cams, POOL, seed0=1000 + 100 * rank)           # distinct frames per rank
    pool_dev = [torch.from_numpy(np.stack([frames[(i + b) % POOL] for b in range(B)])).to(dev) for i in range(POOL)]
    pool_host = [torch.from_numpy(np.stack([frames[(i + b) % POOL] for b in range(B)])).pin_memory() for i in range(POOL)]
    if not args.no_graph:
        lanes.use_cuda_graph(True)

    # ---- the sharded loop: every frame's rows land in the rank's shard buffer, ONE all-gather per shard ------------
    side = torch.cuda.Stream(device=dev)
    shard_dev = [torch.zeros((S * B, P, J, 5), device=dev) for _ in range(2)]
    gathered = [torch.zeros((world * S * B, P, J, 5), device=dev) for _ in range(2)] if world > 1 else None
    gather_done = [torch.cuda.Event() for _ in range(2)]
    coll_ev = []                                 # (start, end) events of every collective on the side stream

    def launch_gather(p, timed):
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)                        # every frame of the shard was collected on `cur`
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
            fdist.gather_shards(shard_dev[p], gathered[p])      # THE collective (run/validate.py:114)
            if timed:
                e1.record(side)
                coll_ev.append((e0, e1))
            gather_done[p].record(side)

    def run_steps(i0, nsteps, timed=False):
        """nsteps forwards, L in flight, gather per shard; all work is ordered before the current stream's tail."""
        cur = torch.cuda.current_stream(dev)
        submitted = collected = 0
        out = None

        def finish_one():
            nonlocal collected, out
            out = lanes.collect()[0]             # the current stream now waits for that frame
            collected += 1
            if world > 1 and collected % S == 0:
                launch_gather((collected // S - 1) % 2, timed)

        for i in range(i0, i0 + nsteps):
            p, row = (submitted // S) % 2, submitted % S
            if world > 1 and row == 0 and submitted >= 2 * S:
                cur.wait_event(gather_done[p])   # the collective still reading this buffer must be through
            lanes.submit(pool_dev[i % POOL], slots, out_fused=shard_dev[p][row * B:(row + 1) * B])
            submitted += 1
            if lanes.outstanding() == L:
                finish_one()
        while lanes.outstanding():
            finish_one()
        if world > 1:
            if collected % S:                    # the last, partial shard
                launch_gather((collected // S) % 2, timed)
            cur.wait_stream(side)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run_steps(0, max(args.warmup, 2 * L))
    sync_all()
    launches_per_step = eng.last_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # time EXACTLY K steps per repetition, bracketed by barrier + synchronize; repeat until >= --min-time seconds are covered
    reps, rep_ms = 0, []
    while True:
        sync_all()
        ev0.record()
        out = run_steps(reps * args.steps, args.steps, timed=True)
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        rep_ms.append(ms)
        reps += 1
        if sum(rep_ms) >= args.min_time * 1e3 or reps >= 50:
            break
    ms = float(np.mean(rep_ms))                  # per K steps (max over ranks per repetition)
    clocks = sampler.stop() if sampler else None
    n_valid = int((out[..., 0, 3] >= 0).sum().item())
    coll_ms = [a.elapsed_time(b) for a, b in coll_ev]
    eng.check_range()                            # synchronises; raises if an activation left the fp16 range

    # ---- serial forwards on one lane (one frame in flight): the per-frame latency ------------------------------
    # From here on lane 0 runs alone, so it uses the latency kernels (fvp_set_latency_mode; an engine without lanes picks them
    # by itself): C2CNet on one 8-CTA cluster per column.  The pipelined `value` above ran the throughput kernels.
    eng.set_latency_mode(1)
    ser_steps = max(10, args.steps // 4)
    for i in range(3):
        eng.forward(pool_dev[i % POOL], slots)
    sync_all()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(ser_steps + 1)]
    marks[0].record()
    for i in range(ser_steps):
        eng.forward(pool_dev[i % POOL], slots)
        marks[i + 1].record()
    sync_all()
    serial_ms = marks[0].elapsed_time(marks[-1]) / ser_steps
    per_step = np.array([marks[i].elapsed_time(marks[i + 1]) for i in range(ser_steps)])     # per-frame latency distribution
    serial_p50, serial_p95 = float(np.percentile(per_step, 50)), float(np.percentile(per_step, 95))
    eng.set_latency_mode(-1)                     # back to automatic (throughput kernels while the lanes exist) for the e2e pipeline

    # ---- the reference-facing plugin: models.faster_voxelpose.get(cfg)(...) as run/validate.py:102-105 calls it ----
    plugin = None
    if rank == 0:
        try:
            import models
            pcfg = workload(args.preset)[0]
            pcfg.DEVICE = str(dev)
            pcfg.TEST.BATCH_SIZE = B
            model = models.faster_voxelpose.get(pcfg)
            model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()})
            model = model.to(dev).eval()
            meta, camd = {"seq": ["bench"] * B}, {"bench": cams}
            rz = torch.as_tensor(resize, dtype=torch.float, device=dev)
            with torch.no_grad():
                for i in range(5):
                    model(meta=meta, input_heatmaps=pool_dev[i % POOL], cameras=camd, resize_transform=rz)
                model.engine().use_cuda_graph(not args.no_graph)
                for i in range(5):
                    model(meta=meta, input_heatmaps=pool_dev[i % POOL], cameras=camd, resize_transform=rz)
                torch.cuda.synchronize(dev)
                pl_steps = max(20, args.steps // 2)
                # (a) exactly the reference's loop (run/validate.py:94-114): results are appended, one torch.cat at the end -
                #     no host synchronisation per call, the Python side of call i+1 runs under the kernels of call i
                t0 = time.perf_counter()
                keep = []
                for i in range(pl_steps):
                    keep.append(model(meta=meta, input_heatmaps=pool_dev[i % POOL], cameras=camd, resize_transform=rz)[0])
                allp = torch.cat(keep, dim=0)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                # (b) a caller that reads every batch's poses back before the next call
                t0 = time.perf_counter()
                for i in range(pl_steps):
                    fp = model(meta=meta, input_heatmaps=pool_dev[i % POOL], cameras=camd, resize_transform=rz)[0]
                    fp_host = fp.cpu()
                dt_sync = time.perf_counter() - t0
            plugin = {"value": B * pl_steps / dt, "unit": UNIT, "ms_per_call": dt / pl_steps * 1e3, "steps": pl_steps,
                      "what": "nn.Module.forward of models.faster_voxelpose.get(cfg) driven like run/validate.py:94-114 (append the "
                              "poses, one torch.cat + synchronise at the end), one stream, wall clock",
                      "with_cpu_readback_per_call": {"value": B * pl_steps / dt_sync, "ms_per_call": dt_sync / pl_steps * 1e3}}
            del allp, keep
            del fp_host
            model.engine().close()
            del model
        except Exception as e:      # pragma: no cover - never take the bench line down
            plugin = {"error": repr(e)}

    # ---- e2e through the host-buffer entry point -----------------------------------------------------
    # (a) blocking call (fvp_forward_host): per-call latency; (b) the pipeline (fvp_submit_host / fvp_wait): every step
    # copies its own inputs H2D from pinned memory and reads its own results back D2H, copies overlapping the kernels of
    # the other lanes.  (b) is the throughput figure reported as e2e.value.  With N > 1 the fused rows of a shard land in
    # one pinned block which goes to the GPU once per shard for the same all-gather.
    host_out = eng.new_host_outputs(B)
    depth = L + 1 if L > 1 else 2                # frames in flight through the host entry (<= 2 tickets per lane)
    host_shard = [torch.zeros((S * B, P, J, 5)).pin_memory() for _ in range(2)]
    host_rest = [eng.new_host_outputs(B)[1:] for _ in range(depth + 1)]
    e2e_steps = max(10, args.steps // 2)
    for i in range(3):
        eng.forward_host(pool_host[i % POOL], slots, host_out)
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.forward_host(pool_host[i % POOL], slots, host_out)
    sync_all()
    sync_call_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    def pipelined(nsteps):
        submitted = done = 0

        def host_gather(p):
            gather_done[p].synchronize()         # earlier collective reading shard_dev[p] / host_shard[p]
            with torch.cuda.stream(side):
                shard_dev[p].copy_(host_shard[p], non_blocking=True)
                fdist.gather_shards(shard_dev[p], gathered[p])
                gather_done[p].record(side)

        def finish():
            nonlocal done
            lanes.wait_oldest()                  # results of step `done` are now in its pinned host rows
            done += 1
            if world > 1 and done % S == 0:
                host_gather((done // S - 1) % 2)

        for i in range(nsteps):
            p, row = (submitted // S) % 2, submitted % S
            if world > 1 and row == 0 and submitted >= 2 * S:
                gather_done[p].synchronize()
            rest = host_rest[submitted % len(host_rest)]
            lanes.submit_host(pool_host[i % POOL], slots, (host_shard[p][row * B:(row + 1) * B], rest[0], rest[1]))
            submitted += 1
            if lanes.host_outstanding() == depth:
                finish()
        while lanes.host_outstanding():
            finish()
        if world > 1:
            if done % S:
                host_gather((done // S) % 2)
            side.synchronize()

    pipelined(2 * depth)
    sync_all()
    e2e_reps, e2e_total = 0, 0.0
    while True:
        sync_all()
        t0 = time.perf_counter()
        pipelined(e2e_steps)
        sync_all()
        dt = (time.perf_counter() - t0) * 1e3    # wall clock: the host side is part of the figure
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_total += dt
        e2e_reps += 1
        if e2e_total >= args.min_time * 1e3 or e2e_reps >= 50:
            break
    e2e_ms = e2e_total / e2e_reps
    h2d = int(B * V * J * H * W * 4)
    d2h = int(sum(t.numel() for t in host_out) * 4)

    # ---- per-stage CUDA-event times (same stream, graph off) -> roofline of the back-projection ----
    eng.use_cuda_graph(False)
    eng.set_latency_mode(1)                      # lane 0 alone again: the stage times belong to the serial figure
    eng.set_profiling(True)
    acc = np.zeros(9)
    prof_steps = min(args.steps, 50)
    for i in range(3 + prof_steps):
        eng.forward(pool_dev[i % POOL], slots)
        if i >= 3:
            acc += np.array(eng.stage_times_ms())
    eng.set_profiling(False)
    eng.set_latency_mode(-1)
    stage = acc / prof_steps
    k1_bytes, k3_bytes = eng.algorithmic_bytes(max(1, n_valid // B))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    k3_ms = stage[5] / 1.0
    k3_gbs = B * k3_bytes / (k3_ms * 1e-3) / 1e9
    k1_ms = stage[0] + stage[1]
    k1_gbs = B * k1_bytes / (k1_ms * 1e-3) / 1e9
    samples = B * max(1, n_valid // B) * V * J * 64 ** 3
    roofline = {"kernel": "k3_jln_patch: fused per-person back-projection + 3-plane max (incl. the memset of the planes)", "bound": "hbm",
                "achieved": k3_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k3_gbs / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": B * k3_bytes, "ms_per_launch": k3_ms,
                "traffic": (traffic or {}).get("k3_dram_bytes_per_launch"),
                "traffic_source": (traffic or {}).get("source", "none") + " (static: read from profiles/traffic.json, not measured in this run)",
                "bilinear_samples_per_s": samples / (k3_ms * 1e-3),
                "l1_wavefront_pct_of_peak_ncu": (traffic or {}).get("k3_l1_wavefront_pct_of_peak_batch8"),
                "note": "K3 reads each heat map once from HBM but gathers 4 x 64 B per voxel-view on chip: ncu shows the L1 data "
                        "pipe at 86 % of its peak wavefront rate at batch 8 (70 % at batch 1) - that, not HBM, is its roofline "
                        "(DESIGN.md 4.2); frac is reported against HBM as the contract asks; see also roofline_k1"}
    # secondary (honest) bound of K3: every voxel-view gathers 4 taps x 64 B (16 joints fp32) from L1; the L1 data pipe
    # delivers 128 B/clk/SM.  This is the roofline K3 actually runs against.
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    l1_bytes = B * max(1, n_valid // B) * V * 64 ** 3 * 4 * 16 * ((J + 3) // 4)
    l1_tbs = l1_bytes / (k3_ms * 1e-3) / 1e12
    roofline["on_chip"] = {"bound": "l1", "achieved": l1_tbs, "peak": l1_peak, "unit": "TB/s", "frac": l1_tbs / l1_peak,
                           "bytes_per_launch": l1_bytes, "peak_source": "148 SMs x 128 B/clk x sm_max_mhz"}
    roofline_k1 = {"kernel": "k0_stage_heatmaps + k1_hdn_project_zmax: whole-space back-projection + z-max (two launches)",
                   "bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                   "peak_source": peak_src, "algorithmic_bytes_per_launch": B * k1_bytes, "ms_per_launch": k1_ms,
                   "traffic": ((traffic or {}).get("k0_dram_bytes_per_launch", 0) + (traffic or {}).get("k1_dram_bytes_per_launch", 0)) or None,
                   "traffic_source": "static: profiles/traffic.json"}
    extra_kernels = {
        "k0+k1_hdn_project": {"ms": k1_ms, "algorithmic_bytes": B * k1_bytes, "achieved_gbs": k1_gbs, "frac_hbm": k1_gbs / hbm_peak},
        "stage_ms": {n: float(v) for n, v in zip(["k0_stage", "k1_hdn_project", "center_net", "nms_topk", "proposals_c2c",
                                                  "k3_jln_project", "p2p_net", "pose_head", "total"], stage)},
    }

    try:        # secondary rooflines (tensor pipe) of the conv trunks; never allowed to take the bench line down
        extra_kernels["conv_tensor_pipe"] = conv_tensor_rooflines(
            extra_kernels["stage_ms"], B, max(1, n_valid // B), J, [int(v) for v in cfg.CAPTURE_SPEC.VOXELS_PER_AXIS[:2]],
            float(peaks.get("bf16_tflops", 2250.0)))
    except Exception as e:      # pragma: no cover
        extra_kernels["conv_tensor_pipe"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_base = gpu_port = None
    if world == 1 and not args.no_cpu_baseline:
        lanes.close()                            # free the workspaces before the PyTorch baseline allocates its own
        torch.cuda.empty_cache()
        try:
            gpu_port = gpu_reference_port_fps(cfg, cams, resize, frames[:4], sd_np, dev)
        except Exception as e:      # pragma: no cover
            gpu_port = {"error": repr(e)}
        fps_c, cores, med = cpu_reference_fps(cfg, cams, resize, frames[:4], sd_np, 2, 12)
        cpu_base = {"value": fps_c, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "12 single-frame forwards of the oracle port after 2 warm-ups (median %.0f ms)" % med}
    fps = world * B * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": conf, "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": launches_per_step, "timed_reps": reps, "timed_total_ms": float(sum(rep_ms)),
        "frames_in_flight": L, "pipelining": pipelining,
        "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "reps": e2e_reps, "ms_per_step": e2e_ms / e2e_steps,
                "api": "fvp_submit_host/fvp_wait, %d frames in flight over %d lanes" % (depth, L),
                "blocking_call_ms": sync_call_ms},
        "roofline": roofline, "roofline_k1": roofline_k1, "cpu_baseline": cpu_base, "gpu_reference_port": gpu_port,
        "plugin_forward": plugin, "kernels": extra_kernels, "valid_people_last_step": n_valid,
        "kernel_modes": {"value_and_e2e": "throughput" if L > 1 else "latency", "serial_and_stage_ms": "latency",
                         "what": "fvp_set_latency_mode: proposal stage on one CTA per column (throughput) or one 8-CTA cluster per column (latency)"},
        "cuda_graph": not args.no_graph, "serial_ms_per_step": serial_ms, "serial_fps": world * B / (serial_ms * 1e-3),
        "serial_ms_p50": serial_p50, "serial_ms_p95": serial_p95,
        "collective": None if world == 1 else {
            "op": "all_gather_into_tensor [%d,%d,%d,5] fp32 per rank -> [%d,...] (NCCL, side stream)" % (S * B, P, J, world * S * B),
            "per_shard_frames": S * B, "count_in_timed_region": len(coll_ms),
            "ms_mean": float(np.mean(coll_ms)) if coll_ms else None, "ms_max": float(np.max(coll_ms)) if coll_ms else None,
            "ms_per_step_amortised": float(np.sum(coll_ms) / max(1, reps * args.steps)) if coll_ms else None,
            "cores_of_rank0": len(my_cores), "gpu_numa_nodes": gpu_nodes},
    }
    if gpu_port and "value" in gpu_port:
        line["vs_gpu_reference_port"] = {"pipelined": fps / gpu_port["value"], "serial": (B / (serial_ms * 1e-3)) / gpu_port["value"]}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
