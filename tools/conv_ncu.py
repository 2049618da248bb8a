#!/usr/bin/env python
"""One conv layer through fvp_debug_conv, for `ncu -k regex:k_conv_tc` captures: conv_ncu.py n H W cin cout k [mode] [repeat]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden
from fvp.engine import Engine
n, H, W, cin, cout, k = [int(v) for v in sys.argv[1:7]]
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 2
repeat = int(sys.argv[8]) if len(sys.argv) > 8 else 3
g = Golden("panoptic_none_valid")
eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=1, max_sequences=1, axes=g.axes)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32)).cuda()
w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
b = np.zeros(cout, np.float32)
_, ms = eng.debug_conv(x, w, b, True, mode, repeat=repeat, want_ms=True)
print("%dx%d %d->%d %dx%d n=%d mode %d: %.1f us, %.1f TMAC/s" % (k, k, cin, cout, H, W, n, mode, ms * 1e3, n * H * W * cin * cout * k * k / 1e9 / ms))
