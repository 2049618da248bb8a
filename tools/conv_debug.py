#!/usr/bin/env python
"""Single-layer TC-vs-FFMA conv comparison (per tap), plus timing of both engines on a P2PNet-sized layer."""
import os, sys, json
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden
from fvp.engine import Engine
g = Golden("panoptic_none_valid")
eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=1, max_sequences=1, axes=g.axes)
rng = np.random.default_rng(0)
def ref(x, w, b, relu):
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), torch.from_numpy(w).double().cuda(), torch.from_numpy(b).double().cuda(), padding=w.shape[2] // 2)
    y = y.permute(0, 2, 3, 1)
    return (y.clamp_min(0) if relu else y).float()
def run(tag, n, H, W, cin, cout, k, wmask=None):
    x = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32)).cuda()
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    if wmask is not None:
        w = w * wmask
    b = rng.standard_normal(cout).astype(np.float32) * 0.1
    r = ref(x, w, b, False)
    e = [float((eng.debug_conv(x, w, b, False, m) - r).abs().max()) for m in (0, 1, 2)]
    print("%-28s ffma %.2e  tf32x3 %.2e  fp16x3 %.2e" % (tag, e[0], e[1], e[2])); sys.stdout.flush()
run("7x7 16->16 16x8", 1, 16, 8, 16, 16, 7)
run("3x3 16->32 16x8", 1, 16, 8, 16, 32, 3)
run("1x1 32->32 16x8", 1, 16, 8, 32, 32, 1)
run("1x1 64->128 32x32", 2, 32, 32, 64, 128, 1)
for dy in range(3):
    for dx in range(3):
        m = np.zeros((1, 1, 3, 3), np.float32); m[0, 0, dy, dx] = 1
        run("3x3 32->32 tap(%d,%d) 16x8" % (dy, dx), 1, 16, 8, 32, 32, 3, m)
run("3x3 32->32 full 64x64", 2, 64, 64, 32, 32, 3)
run("3x3 64->64 32x32", 2, 32, 32, 64, 64, 3)
run("3x3 128->128 16x16", 2, 16, 16, 128, 128, 3)
run("3x3 16->32 64x64", 1, 64, 64, 16, 32, 3)
run("7x7 16->16 64x64", 1, 64, 64, 16, 16, 7)
run("3x3 32->32 80x80", 1, 80, 80, 32, 32, 3)
run("3x3 32->64 20x20", 1, 20, 20, 32, 64, 3)
# steady-state timing (CUDA events around repeated launches, weights already uploaded)
def timing(tag, n, H, W, cin, cout, k):
    x = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32)).cuda()
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32); b = np.zeros(cout, np.float32)
    r = []
    for mode in (0, 1, 2):
        _, ms = eng.debug_conv(x, w, b, True, mode, repeat=20, want_ms=True)
        r.append(ms * 1000)
    gmac = n * H * W * cin * cout * k * k / 1e9
    print("%-30s ffma %7.1f us (%5.1f)  tf32x3 %7.1f us (%5.1f)  fp16x3 %7.1f us (%5.1f TMAC/s)" % (
        tag, r[0], gmac / r[0] * 1e3, r[1], gmac / r[1] * 1e3, r[2], gmac / r[2] * 1e3)); sys.stdout.flush()
timing("3x3 32->32 64x64 n=30", 30, 64, 64, 32, 32, 3)
timing("3x3 32->32 64x64 n=120", 120, 64, 64, 32, 32, 3)
timing("3x3 32->32 64x64 n=480", 480, 64, 64, 32, 32, 3)
timing("3x3 64->64 32x32 n=480", 480, 32, 32, 64, 64, 3)
timing("3x3 128->128 16x16 n=480", 480, 16, 16, 128, 128, 3)
timing("7x7 16->16 64x64 n=120", 120, 64, 64, 16, 16, 7)
timing("1x1 64->128 32x32 n=480", 480, 32, 32, 64, 128, 1)
timing("3x3 16->32 64x64 n=120", 120, 64, 64, 16, 32, 3)
