#!/usr/bin/env python
"""PoseResNet backbone (N2) on the GPU: time per 5-view frame of 960x512 images and achieved useful TMAC/s.
    python tools/backbone_bench.py [num_layers=50] [views=5] [h=512] [w=960] [reps=20]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from fvp import backbone_spec as BS, config as fcfg, synth
from fvp.backbone import Backbone
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 50
V = int(sys.argv[2]) if len(sys.argv) > 2 else 5
h = int(sys.argv[3]) if len(sys.argv) > 3 else 512
w = int(sys.argv[4]) if len(sys.argv) > 4 else 960
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
cfg = fcfg.preset("panoptic"); cfg.RESNET.NUM_LAYERS = nl
layers = BS.from_cfg(cfg)
macs = sum(r["macs"] for r in BS.shapes_and_macs(layers, h, w))
bb = Backbone(nl, int(cfg.DATASET.NUM_JOINTS), "cuda:0", max_images=V, max_h=h, max_w=w)
bb.load_state_dict(synth.make_backbone_weights(layers, 7))
x = torch.from_numpy(np.random.default_rng(0).standard_normal((V, 3, h, w)).astype(np.float32)).cuda()
for _ in range(3):
    y = bb.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = bb.forward(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"backbone": "PoseResNet-%d" % nl, "views": V, "image": [h, w], "ms_per_frame": ms, "frames_per_s": 1e3 / ms,
                  "reference_gmac_per_frame": V * macs / 1e9, "useful_tmac_per_s": V * macs / 1e12 / (ms * 1e-3),
                  "finite": bool(torch.isfinite(y).all())}))
