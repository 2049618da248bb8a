"""BASELINE configs[0] on real data, on the GPU: frame 400 of the reference's shipped Campus / Shelf detection files
(tests/golden/{campus,shelf}_frame400.npz, written by oracle/gen_golden.py from the unmodified reference).
detections -> GPU renderer (N1) -> the reference's maps; the reference's maps -> plugin forward -> the reference's outputs."""
import numpy as np
import pytest
import torch

from golden_util import REAL_CASES
from test_gpu_parity import _check_plugin_forward

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", REAL_CASES)
def test_real_detections_render_then_forward(built_library, golden, name):
    from fvp.engine import Engine
    from fvp.render import HeatmapRenderer
    g = golden(name)
    # 1. the renderer on the real detections against the maps the reference rendered (float64 arithmetic on both sides,
    #    exp is the only operation not bit-specified: 1 float32 ulp, >= 99.99 % of the values identical, support exact)
    eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=2, max_sequences=1)
    got = HeatmapRenderer(eng).from_pred([g.real_preds()], g.resize).cpu().numpy()[0]
    eng.close()
    ref = g.rendered()
    assert got.shape == ref.shape and np.array_equal(got == 0, ref == 0)
    d = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4, (int(d.max()), float((d != 0).mean()))
    # 2. the whole path on the reference's maps: cells / flags bit-exact, joints inside the reference's own noise floor
    _check_plugin_forward(g)
