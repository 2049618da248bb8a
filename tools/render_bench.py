#!/usr/bin/env python
"""N1 measurement: the GPU heat-map renderer (fvp_render_heatmaps) on the benchmark geometry (Panoptic 5 views, 256x192 maps,
10 people per view), beside the reference's NumPy loop (oracle port of generate_input_heatmap) on this host.
Bound: HBM writes - the kernel stores every map value exactly once: algorithmic bytes = 4*V*J*H*W per frame."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (PKG, ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fvp import config as fcfg  # noqa: E402
from fvp.engine import Engine  # noqa: E402
from fvp.render import HeatmapRenderer  # noqa: E402

cfg = fcfg.preset("panoptic_256x192")
peak = 6554.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"workload": "Panoptic 5-view, 256x192 maps, 10 people x 15 joints per view", "hbm_peak_gbs": peak}
for B in (1, 32):
    eng = Engine(cfg, torch.device("cuda:0"), max_batch=B, max_sequences=1)
    R = HeatmapRenderer(eng)
    V, J, W, H = eng.V, eng.J, eng.W, eng.H
    rng = np.random.default_rng(0)
    N = 10
    centre = rng.uniform([150.0, 150.0], [870.0, 620.0], (B, V, N, 1, 2))
    joints = centre + rng.uniform(-0.5, 0.5, (B, V, N, J, 2)) * np.array([120.0, 260.0])
    num = np.full((B, V), N, np.int32)
    for _ in range(3):
        hm = R.render(joints, num)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        hm = R.render(joints, num)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ms = e0.elapsed_time(e1) / reps
    nbytes = 4.0 * B * V * J * H * W
    out["batch_%d" % B] = {"device_ms_per_call": ms, "wall_ms_per_call": wall * 1e3, "frames_per_s": B / (ms * 1e-3),
                           "write_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
                           "nonzero_fraction": float((hm > 0).float().mean())}
    if B == 1:
        from oracle import heatmap_oracle as HO          # CPU baseline: the reference's NumPy loop (oracle port)
        t0 = time.perf_counter()
        for v in range(V):
            HO.render_input_heatmap([joints[0, v, n] for n in range(N)], None, cfg.DATASET.HEATMAP_SIZE, cfg.DATASET.IMAGE_SIZE,
                                    cfg.NETWORK.SIGMA)
        out["cpu_port_ms_per_frame"] = (time.perf_counter() - t0) * 1e3
    eng.close()
print(json.dumps(out))
