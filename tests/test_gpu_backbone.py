"""N2 first slice: stem + max-pool + layer1 of the PoseResNet backbone on the GPU against the CPU oracle
(oracle/backbone_oracle.py, pinned bit-identical to the unmodified reference by oracle/gen_golden_backbone.py) and against
the reference-produced tap statistics stored in tests/golden/backbone_*.npz."""
import os

import numpy as np
import pytest
import torch

from fvp import backbone_spec as BS, config as fcfg, synth
from golden_util import weights_sha

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["backbone_resnet50", "backbone_resnet18"])
def test_backbone_stem_maxpool_layer1_match_reference_taps(built_library, name):
    from fvp.backbone import BackboneSlice
    from oracle import backbone_oracle as BO
    z = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = fcfg.preset("panoptic")
    cfg.RESNET.NUM_LAYERS = int(z["num_layers"])
    layers = BS.from_cfg(cfg)
    sd_np = synth.make_backbone_weights(layers, int(z["weight_seed"]))
    assert weights_sha(sd_np) == str(z["weights_sha256"])
    x = torch.from_numpy(z["image_u8"]).permute(0, 3, 1, 2).float().div(255)
    x = (x - torch.from_numpy(z["mean"]).view(1, 3, 1, 1)) / torch.from_numpy(z["std"]).view(1, 3, 1, 1)
    taps = {}
    with torch.no_grad():
        BO.forward(layers, {k: torch.from_numpy(v) for k, v in sd_np.items()}, x, taps)
    golden = {str(k): (float(a), float(s)) for k, a, s in zip(z["tap_names"], z["tap_absmax"], z["tap_sum"])}
    n, _, h, w = x.shape
    bb = BackboneSlice(int(z["num_layers"]), "cuda:0", max_images=n, max_h=h, max_w=w)
    bb.load_state_dict(sd_np)
    names = ["stem"] + ["layer1.%d" % b for b in range(bb.blocks)]
    for stage, tap in enumerate(names):
        got = bb.forward_slice(x, stage).cpu()
        want = taps[tap]
        assert got.shape == want.shape, tap
        scale = float(want.abs().max())
        err = float((got - want).abs().max())
        print("\n[backbone %s] %-9s max|ours - oracle| = %.3g (values up to %.3g)" % (name, tap, err, scale))
        assert err <= 2e-6 * max(1.0, scale) * (1 + stage), tap          # fp32-grade: summation order and the 22-bit operand split
        amax, total = golden[tap]                                         # statistics written by the unmodified reference
        assert abs(float(got.abs().max()) - amax) <= 1e-5 * max(1.0, amax), tap
        assert abs(float(got.double().sum()) - total) <= 2e-6 * float(got.abs().double().sum()) + 1e-6, tap
    # the slice refuses what it does not have yet instead of computing something else
    from fvp.capi import FvpError
    with pytest.raises(FvpError):
        bb.forward_slice(x, bb.blocks + 1)
    bb.close()
