#!/usr/bin/env python
"""Where does the tcgen05 conv spend its time?  Per-role wait/busy cycles (debug counters of k_conv_tc, mode | 0x100).

Counters are summed over CTAs; printed per CTA per item in cycles:
  ld_wait   loader thread 0 waiting for a free A stage      ld_tot   loader loop total
  mma_a     MMA thread waiting for a staged A               mma_acc  MMA thread waiting for a drained accumulator
  mma_b     MMA thread waiting for weights                  mma_tot  MMA loop total
  epi_wait  epilogue waiting for a finished accumulator     epi_tot  epilogue loop total
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden
from fvp.engine import Engine
g = Golden("panoptic_none_valid")
eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=1, max_sequences=1, axes=g.axes)
rng = np.random.default_rng(0)
print("%-28s %8s %6s | %8s %8s | %8s %8s %8s %8s | %8s %8s %8s  (cycles per item)" % (
    "layer", "us", "TMAC/s", "ld_wait", "ld_tot", "mma_a", "mma_acc", "mma_b", "mma_tot", "epi_wait", "epi_tot", "epi_tmem"))
def roles(tag, n, H, W, cin, cout, k, mode=2):
    x = torch.from_numpy(rng.standard_normal((n, H, W, cin)).astype(np.float32)).cuda()
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32); b = np.zeros(cout, np.float32)
    _, ms, c = eng.debug_conv(x, w, b, True, mode | 0x100, repeat=10)
    items = max(c[8], 1e-9) * 1000.0           # counters are in kilo-units
    per = [v * 1000.0 / items for v in c[:8]]
    gmac = n * H * W * cin * cout * k * k / 1e9
    print("%-28s %8.1f %6.1f | %8.0f %8.0f | %8.0f %8.0f %8.0f %8.0f | %8.0f %8.0f %8.0f   items %d" % (
        tag, ms * 1000, gmac / ms, per[0], per[1], per[2], per[3], per[4], per[5], per[6], per[7], c[9] * 1000.0 / items, round(items)))
    sys.stdout.flush()
for n in (30, 960):
    roles("3x3 32->32 64x64 n=%d" % n, n, 64, 64, 32, 32, 3)
    roles("3x3 64->64 32x32 n=%d" % n, n, 32, 32, 64, 64, 3)
    roles("3x3 128->128 16x16 n=%d" % n, n, 16, 16, 128, 128, 3)
    roles("7x7 16->16 64x64 n=%d" % n, n, 64, 64, 16, 16, 7)
    roles("3x3 16->32 64x64 n=%d" % n, n, 64, 64, 16, 32, 3)
    roles("1x1 64->128 32x32 n=%d" % n, n, 32, 32, 64, 128, 1)
    roles("1x1 32->16 64x64 n=%d" % n, n, 64, 64, 32, 16, 1)
roles("3x3 32->32 80x80 n=1", 1, 80, 80, 32, 32, 3)
roles("3x3 64->64 40x40 n=1", 1, 40, 40, 64, 64, 3)
roles("3x3 128->128 20x20 n=1", 1, 20, 20, 128, 128, 3)
