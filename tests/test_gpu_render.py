"""N1: the GPU heat-map renderer through the C ABI (fvp_render_heatmaps) against the maps rendered by the
unmodified reference (tests/golden/heatmaps_*.npz) and against size-independent properties."""
import numpy as np
import pytest
import torch

from golden_util import HEATMAP_CASES, HeatmapGolden

pytestmark = pytest.mark.gpu


def _renderer(cfg, max_batch=2):
    from fvp.engine import Engine
    from fvp.render import HeatmapRenderer
    eng = Engine(cfg, torch.device("cuda:0"), max_batch=max_batch, max_sequences=1)
    return eng, HeatmapRenderer(eng)


def _ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("name", HEATMAP_CASES)
def test_renderer_matches_reference_maps(built_library, name):
    """'pred' and 'gt' sources.  The arithmetic is float64 on both sides; the only operation not bit-specified is exp
    (<= 1 ulp of float64 in CUDA and in NumPy), so after the single rounding to float32 the maps may differ by one
    float32 ulp in a few values per million: tolerance 1 ulp, and at least 99.99 % of the values bit-identical."""
    g = HeatmapGolden(name)
    eng, R = _renderer(g.cfg)
    for which, got in (("pred", R.from_pred([g.preds], g.resize)),
                       ("gt", R.from_gt([g.joints_3d], [g.joints_3d_vis], [g.cams], g.resize))):
        got = got.cpu().numpy()[0]
        ref = g.dense(which)
        assert got.shape == ref.shape
        assert np.array_equal(got == 0, ref == 0), which            # the support (patch windows, int() truncation) is exact
        d = _ulp_diff(got, ref)
        assert d.max() <= 1, (which, int(d.max()))
        assert (d != 0).mean() <= 1e-4, (which, float((d != 0).mean()))
    eng.close()


def test_renderer_batch_and_padding(built_library):
    """Frames of a batch are independent, padded person slots are ignored, an empty view renders zeros."""
    g = HeatmapGolden("heatmaps_panoptic_256x192_crowd")
    eng, R = _renderer(g.cfg, max_batch=3)
    one = R.from_pred([g.preds], g.resize)
    empty_view = [g.preds[0], [], g.preds[2], g.preds[3][:1], g.preds[4]]
    both = R.from_pred([empty_view, g.preds, g.preds], g.resize)
    assert torch.equal(both[1], one[0]) and torch.equal(both[2], one[0])
    assert float(both[0, 1].abs().max()) == 0.0
    assert torch.equal(both[0, 0], one[0, 0]) and torch.equal(both[0, 4], one[0, 4])
    assert float(both[0, 3].max()) == 1.0
    eng.close()


def test_renderer_properties_full_size(built_library):
    """At the benchmark size: values in [0,1]; every drawn joint peaks at exactly 1.0; the map is the per-pixel maximum
    over persons (rendering persons separately and taking the maximum gives the same bits); a joint's patch is
    symmetric about its centre."""
    from fvp import config as fcfg
    cfg = fcfg.preset("panoptic_256x192")
    eng, R = _renderer(cfg, max_batch=1)
    rng = np.random.default_rng(9)
    V, J = eng.V, eng.J
    N = 10
    joints = rng.uniform([60.0, 60.0], [960.0, 700.0], (1, V, N, J, 2))
    num = np.full((1, V), N, np.int32)
    full = R.render(joints, num)
    assert float(full.min()) >= 0.0 and float(full.max()) == 1.0
    acc = torch.zeros_like(full)
    for n in range(N):
        acc = torch.maximum(acc, R.render(joints[:, :, n:n + 1], np.ones((1, V), np.int32)))
    assert torch.equal(acc, full)
    single = R.render(joints[:, :, :1], np.ones((1, V), np.int32))[0, 0, 0].cpu().numpy()
    ys, xs = np.nonzero(single == 1.0)
    assert len(ys) == 1
    cy, cx = int(ys[0]), int(xs[0])
    r = 5
    patch = single[cy - r:cy + r + 1, cx - r:cx + r + 1]
    assert np.array_equal(patch, patch[::-1, :]) and np.array_equal(patch, patch[:, ::-1]) and np.array_equal(patch, patch.T)
    eng.close()


def test_rendered_maps_feed_the_hot_path(built_library, golden):
    """2-D poses -> GPU renderer -> fvp_forward without leaving the device gives exactly what the forward gives on the
    same maps uploaded from the host (the renderer writes the layout the hot path consumes)."""
    from fvp.engine import Engine
    from fvp.render import HeatmapRenderer
    g = golden("panoptic_256x192")
    hg = HeatmapGolden("heatmaps_panoptic_256x192_crowd")
    eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=1, max_sequences=1, axes=g.axes)
    eng.load_state_dict(g.weights)
    slot = eng.sequence_slot(g.cams, g.resize)
    maps = HeatmapRenderer(eng).from_pred([hg.preds], hg.resize)
    a = eng.forward(maps, [slot])
    b = eng.forward(torch.from_numpy(maps.cpu().numpy()).cuda(), [slot])
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(ValueError):
        HeatmapRenderer(eng).render(np.zeros((1, eng.V, 17, eng.J, 2)), np.zeros((1, eng.V), np.int32))
    eng.close()
