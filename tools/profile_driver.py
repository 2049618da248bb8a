#!/usr/bin/env python
"""Short driver for ncu: a few eager forwards of the bench workload (batch 1, P=10 valid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from fvp import synth
from fvp.engine import Engine
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
PRESET = os.environ.get("FVP_PROFILE_PRESET", "panoptic_256x192")
cfg, cams, resize = bench.workload(PRESET)
eng = Engine(cfg, torch.device("cuda:0"), max_batch=batch, max_sequences=1)
eng.load_state_dict(synth.make_weights(int(cfg.DATASET.NUM_JOINTS), seed=2024))
if 'FVP_CONV_MODE' in os.environ:
    eng.set_conv_mode(int(os.environ['FVP_CONV_MODE']))
slot = eng.sequence_slot(cams, resize)
frames = bench.make_frames(cfg, cams, batch, seed0=1000)
hm = torch.from_numpy(frames).cuda()
for i in range(n_iter):
    out = eng.forward(hm, [slot] * batch)
torch.cuda.synchronize()
print("valid", int((out[0][..., 0, 3] >= 0).sum()), "launches", eng.last_launch_count())
