#!/usr/bin/env python
"""Steady-state time of the trunk's conv layer shapes (n images through fvp_debug_conv): conv_layers.py [n ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden
from fvp.engine import Engine
g = Golden("panoptic_none_valid")
eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=1, max_sequences=1, axes=g.axes)
rng = np.random.default_rng(0)
ns = [int(v) for v in sys.argv[1:]] or [30, 960]
MODE = int(os.environ.get("FVP_CONV_DEBUG_MODE", "2"))     # 2 = engine 2 with fp32 activations (legacy loaders), 4 = split activations (TMA-fed)
LAYERS = [("7x7 16->16 64x64", 64, 16, 16, 7), ("3x3 16->32 64x64", 64, 16, 32, 3), ("3x3 32->32 64x64", 64, 32, 32, 3),
          ("1x1 32->16 64x64", 64, 32, 16, 1), ("3x3 32->64 32x32", 32, 32, 64, 3), ("3x3 64->64 32x32", 32, 64, 64, 3),
          ("3x3 64->128 16x16", 16, 64, 128, 3), ("3x3 128->128 16x16", 16, 128, 128, 3), ("1x1 128->256 16x16", 16, 128, 256, 1)]
for n in ns:
    tot = 0.0
    for tag, S, cin, cout, k in LAYERS:
        x = torch.from_numpy(rng.standard_normal((n, S, S, cin)).astype(np.float32)).cuda()
        w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
        y, ms = eng.debug_conv(x, w, np.zeros(cout, np.float32), True, MODE if (k != 7 or MODE != 4) else 2, repeat=10, want_ms=True)
        tot += ms
        m = min(n, 64)                                   # value check against an fp64 convolution of the first images
        ref = torch.nn.functional.conv2d(x[:m].permute(0, 3, 1, 2).double(), torch.from_numpy(w).double().cuda(), padding=k // 2)
        err = float((y[:m] - ref.permute(0, 2, 3, 1).clamp_min(0).float()).abs().max())
        err2 = float((y[-1] - torch.nn.functional.conv2d(x[-1:].permute(0, 3, 1, 2).double(), torch.from_numpy(w).double().cuda(),
                                                         padding=k // 2).permute(0, 2, 3, 1).clamp_min(0).float()[0]).abs().max())
        print("%-22s n=%-4d %8.1f us %7.1f TMAC/s  max|err| %.1e / %.1e" % (tag, n, ms * 1e3, n * S * S * cin * cout * k * k / 1e9 / ms, err, err2)); sys.stdout.flush()
    print("sum n=%d: %.1f us" % (n, tot * 1e3))
