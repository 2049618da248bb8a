#!/usr/bin/env python
"""BASELINE configs[4]: back-projection bandwidth sweep of K0 + K1 (stage heat maps, HDN back-projection + z-max) on the
synthetic ring calibration: grid in {80x80x20, 120x120x32, 160x160x40} x views in {4, 5, 8}, 256x192 heat maps, uniform
random inputs (timing is value-independent).  Prints achieved GB/s on the algorithmic bytes of SURVEY.md 8(d)
(Hm + 4*J*X*Y + 84*V per frame) and bilinear samples/s; CUDA events, inputs rotated through a pool larger than L2."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (PKG, ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fvp import config as fcfg, synth  # noqa: E402
from fvp.engine import Engine  # noqa: E402

B, REPS = 8, 20
peak = 6554.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print("# K0+K1 sweep, batch %d, HBM peak %.0f GB/s (measured copy)" % (B, peak))
print("%-12s %2s | %9s %9s | %9s %8s | %12s" % ("grid", "V", "k0 us/fr", "k1 us/fr", "GB/s(alg)", "frac", "samples/s"))
for vox in ((80, 80, 20), (120, 120, 32), (160, 160, 40)):   # Z % 4 == 0 (C2CNet pools z twice)
    for V in (4, 5, 8):
        cfg = fcfg.preset("ring8_160")
        cfg.DATASET.CAMERA_NUM = V
        cfg.CAPTURE_SPEC.VOXELS_PER_AXIS = list(vox)
        J = int(cfg.DATASET.NUM_JOINTS)
        W, H = [int(v) for v in cfg.DATASET.HEATMAP_SIZE]
        cams = synth.ring_cameras(V, cfg.CAPTURE_SPEC.SPACE_CENTER)
        resize = synth.resize_transform(cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE)
        eng = Engine(cfg, torch.device("cuda:0"), max_batch=B, max_sequences=1)
        slot = eng.sequence_slot(cams, resize)
        g = torch.Generator(device="cuda").manual_seed(0)
        npool = max(2, int(np.ceil(160e6 / (B * V * J * H * W * 4))))          # pool > 126 MB L2
        pool = [torch.rand((B, V, J, H, W), device="cuda", generator=g) for _ in range(npool)]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = t1 = 0.0
        for r in range(3 + REPS):
            hm = pool[r % npool]
            e[0].record()
            eng.stage_heatmaps(hm)
            e[1].record()
            eng.hdn_project(B, [slot] * B)
            e[2].record()
            torch.cuda.synchronize()
            if r >= 3:
                t0 += e[0].elapsed_time(e[1])
                t1 += e[1].elapsed_time(e[2])
        t0, t1 = t0 / REPS / B, t1 / REPS / B                                    # ms per frame
        alg = 4.0 * V * J * H * W + 4.0 * J * vox[0] * vox[1] + 84.0 * V
        gbs = alg / ((t0 + t1) * 1e-3) / 1e9
        samples = V * J * vox[0] * vox[1] * vox[2] / ((t0 + t1) * 1e-3)
        print("%-12s %2d | %9.1f %9.1f | %9.1f %8.4f | %12.3e" % ("%dx%dx%d" % vox, V, t0 * 1e3, t1 * 1e3, gbs, gbs / peak, samples))
        sys.stdout.flush()
        eng.close()
        del pool
        torch.cuda.empty_cache()
