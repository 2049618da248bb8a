"""PoseResNet backbone groundwork: layer table, weight generator and CPU oracle against goldens
produced by the unmodified reference (oracle/gen_golden_backbone.py).  CPU only - the CUDA kernels of this row do not
exist yet."""
import os

import numpy as np
import pytest
import torch

from fvp import backbone_spec as BS, config as fcfg, synth
from golden_util import weights_sha
from oracle import backbone_oracle as BO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = fcfg.preset("panoptic")
    cfg.RESNET.NUM_LAYERS = int(z["num_layers"])
    layers = BS.from_cfg(cfg)
    sd_np = synth.make_backbone_weights(layers, int(z["weight_seed"]))
    assert weights_sha(sd_np) == str(z["weights_sha256"]), "backbone weight generator drifted from the golden"
    x = torch.from_numpy(z["image_u8"]).permute(0, 3, 1, 2).float().div(255)
    x = (x - torch.from_numpy(z["mean"]).view(1, 3, 1, 1)) / torch.from_numpy(z["std"]).view(1, 3, 1, 1)
    return z, layers, {k: torch.from_numpy(v) for k, v in sd_np.items()}, x


@pytest.mark.parametrize("name", ["backbone_resnet50", "backbone_resnet18"])
def test_backbone_oracle_is_bit_identical_to_reference_golden(name):
    z, layers, sd, x = _load(name)
    taps = {}
    with torch.no_grad():
        y = BO.forward(layers, sd, x, taps)
    assert sorted(taps) == list(z["tap_names"])
    # The pin itself (oracle == live reference, every tap, bit for bit) is asserted by oracle/gen_golden_backbone.py when the
    # golden is made.  Here the stored reference output is expected bit-exact too on the machine class the golden came
    # from; another CPU's convolution kernels (different vector width -> summation order) may differ in the last bits.
    if not np.array_equal(y.numpy(), z["output"]):
        assert np.allclose(y.numpy(), z["output"], rtol=2e-5, atol=2e-6)
    for k, amax, total in zip(z["tap_names"], z["tap_absmax"], z["tap_sum"]):
        t = taps[str(k)]
        assert abs(float(t.abs().max()) - amax) <= 1e-5 * max(1.0, amax), k
        assert abs(float(t.double().sum()) - total) <= 1e-6 * float(t.abs().double().sum()) + 1e-6, k


@pytest.mark.parametrize("name", ["backbone_resnet50", "backbone_resnet18"])
def test_param_table_has_the_reference_keys(name):
    z, layers, sd, _ = _load(name)
    table = BS.param_table(layers)
    assert [k for k, _, _ in table] == list(z["keys"]) == list(sd.keys())
    for k, shape, dtype in table:
        assert tuple(sd[k].shape) == tuple(shape) and str(sd[k].dtype).endswith(dtype)


def test_layer_table_geometry_and_macs():
    L = BS.pose_resnet(50, 15)
    assert len(L) == 1 + (3 + 4 + 6 + 3) * 3 + 4 + 3 + 1            # stem, bottleneck convs, 4 downsamples, 3 deconvs, final
    rows = BS.shapes_and_macs(L, 512, 960)
    assert rows[0]["out"] == (256, 480) and rows[1]["in"] == (128, 240)          # stem stride 2, max-pool stride 2
    assert rows[-1]["out"] == (128, 240)                                          # heat maps at 1/4 resolution (jln64.yaml:28-30)
    assert [r for r in rows if r["key"] == "layer4.2.conv3"][0]["out"] == (16, 30)
    assert abs(sum(r["macs"] for r in rows) / 1e9 - 54.2) < 0.1                    # per view; x5 views = 271 GMAC per frame
    # deconv geometry for the other kernel sizes the reference accepts (resnet.py:148-161)
    for k in (2, 3, 4):
        c = BS.pose_resnet(18, 15, (64,), (k,))[-2]
        assert c.transposed and BS.out_hw(c, 10, 7) == (20, 14)
    with pytest.raises(KeyError):
        BS.pose_resnet(42)
