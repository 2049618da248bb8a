"""N2: the PoseResNet backbone on the GPU against the CPU oracle (oracle/backbone_oracle.py, pinned bit-identical
to the unmodified reference by oracle/gen_golden_backbone.py), against the reference-produced heat maps and tap statistics in
tests/golden/backbone_*.npz, and behind the reference's plugin interface ``models.resnet.get(cfg)``."""
import os

import numpy as np
import pytest
import torch

from fvp import backbone_spec as BS, config as fcfg, synth
from golden_util import weights_sha

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(name):
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
        y = BO.forward(layers, {k: torch.from_numpy(v) for k, v in sd_np.items()}, x, taps)
    return z, cfg, layers, sd_np, x, taps, y


@pytest.mark.parametrize("name", ["backbone_resnet50", "backbone_resnet18"])
def test_backbone_every_stage_matches_the_reference(built_library, name):
    """Every tap of the network - after the max-pool, after each residual block (stride-2 stages included), after each
    transposed convolution - and the final heat maps: fp32-grade agreement with the oracle, the reference's own tap
    statistics and the reference-written heat maps of the golden."""
    from fvp.backbone import Backbone
    z, cfg, layers, sd_np, x, taps, y_ref = _case(name)
    golden = {str(k): (float(a), float(s)) for k, a, s in zip(z["tap_names"], z["tap_absmax"], z["tap_sum"])}
    n, _, h, w = x.shape
    bb = Backbone(int(z["num_layers"]), int(cfg.DATASET.NUM_JOINTS), "cuda:0", max_images=n, max_h=h, max_w=w)
    bb.load_state_dict(sd_np)
    kind, blocks = BS.RESNET_SPEC[int(z["num_layers"])]
    names = ["stem"] + ["layer%d.%d" % (li + 1, b) for li, nb in enumerate(blocks) for b in range(nb)] + \
            ["deconv_layers.%d" % (3 * d) for d in range(3)]
    assert len(names) + 1 == bb.num_stages
    worst = 0.0
    for stage, tap in enumerate(names):
        got = bb.forward_slice(x, stage).cpu()
        want = taps[tap]
        assert got.shape == want.shape, tap
        scale = max(1.0, float(want.abs().max()))
        err = float((got - want).abs().max()) / scale
        worst = max(worst, err)
        # fp32-grade: summation order + the 22-bit operand split; the transposed convolutions sum 4 x 2048 terms per output
        assert err <= (3e-5 if tap.startswith("deconv") else 1e-5), (tap, err)
        amax, total = golden[tap]                                          # statistics written by the unmodified reference
        assert abs(float(got.abs().max()) - amax) <= 4e-5 * max(1.0, amax), tap
        assert abs(float(got.double().sum()) - total) <= 1e-5 * float(got.abs().double().sum()) + 1e-5, tap
    hm = bb.forward(x).cpu()
    assert hm.shape == y_ref.shape == tuple(z["output"].shape)
    e_or = float((hm - y_ref).abs().max())
    e_gold = float(np.abs(hm.numpy() - z["output"]).max())
    print("\n[backbone %s] %d stages: worst relative tap error %.3g; heat maps: max|ours - oracle| = %.3g, max|ours - reference golden| = %.3g "
          "(values up to %.3g)" % (name, len(names), worst, e_or, e_gold, float(np.abs(z["output"]).max())))
    assert e_or <= 5e-5 * max(1.0, float(y_ref.abs().max())) and e_gold <= 5e-5 * max(1.0, float(np.abs(z["output"]).max()))
    assert torch.equal(bb.forward_slice(x, bb.num_stages - 1).cpu(), hm)
    from fvp.capi import FvpError
    with pytest.raises((FvpError, ValueError)):
        bb.forward_slice(x, bb.num_stages)
    bb.close()


def test_backbone_plugin_module_and_image_source_forward(built_library, golden):
    """models.resnet.get(cfg) like run/validate.py:69-74 (strict load of the reference's 338 keys, .to(device), .eval()) and
    the TEST_HEATMAP_SRC = 'image' branch of FasterVoxelPoseNet.forward (faster_voxelpose.py:36-38): views -> backbone per
    camera -> heat maps -> poses, equal to feeding the stacked backbone outputs as input_heatmaps."""
    import models
    z, cfg, layers, sd_np, x, taps, y_ref = _case("backbone_resnet50")
    cfg.DEVICE = "cuda:0"
    backbone = models.resnet.get(cfg)
    assert [k for k, _ in backbone.state_dict().items()] == list(z["keys"])
    backbone.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()})       # strict
    backbone = backbone.to("cuda:0").eval()
    with torch.no_grad():
        hm = backbone(x.cuda())
    assert float((hm.cpu() - y_ref).abs().max()) <= 5e-5 * max(1.0, float(y_ref.abs().max()))
    with pytest.raises(NotImplementedError):
        backbone.train()(x.cuda())
    backbone.eval()
    # the image branch of the pose model: V views of h x w images, heat maps of h/4 x w/4
    g = golden("panoptic_none_valid")
    pcfg = g.cfg
    V, J = int(pcfg.DATASET.CAMERA_NUM), int(pcfg.DATASET.NUM_JOINTS)
    W, H = [int(v) for v in pcfg.DATASET.HEATMAP_SIZE]
    pcfg.DEVICE = "cuda:0"
    model = models.faster_voxelpose.get(pcfg)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in g.weights.items()})
    model = model.to("cuda:0").eval()
    bcfg = fcfg.preset("panoptic")
    bcfg.DEVICE = "cuda:0"
    bmod = models.resnet.get(bcfg)
    bmod.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()})
    bmod = bmod.to("cuda:0").eval()
    rng = np.random.default_rng(3)
    views = torch.from_numpy(rng.standard_normal((1, V, 3, 4 * H, 4 * W)).astype(np.float32)).cuda()
    rz = torch.as_tensor(g.resize, dtype=torch.float).cuda()
    with torch.no_grad():
        a = model(backbone=bmod, views=views, meta={"seq": [g.seq]}, cameras=g.cameras, resize_transform=rz)
        hms = torch.stack([bmod(views[:, c]) for c in range(V)], dim=1)
        b = model(backbone=None, meta={"seq": [g.seq]}, input_heatmaps=hms, cameras=g.cameras, resize_transform=rz)
    assert tuple(a[3].shape) == (1, V, J, H, W)
    assert torch.equal(a[3], hms) and all(torch.equal(p, q) for p, q in zip(a[:3], b[:3]))


def test_backbone_error_behaviour(built_library):
    from fvp.backbone import Backbone
    from fvp.capi import FvpError
    with pytest.raises(FvpError):
        Backbone(50, 15, "cuda:0", max_images=1, max_h=100, max_w=96)          # sides must be multiples of 32
    with pytest.raises(FvpError):
        Backbone(42, 15, "cuda:0", max_images=1, max_h=96, max_w=96)           # not a ResNet depth of resnet.py:204-208
    bb = Backbone(18, 15, "cuda:0", max_images=1, max_h=96, max_w=96)
    x = torch.zeros(1, 3, 96, 96)
    with pytest.raises(FvpError):
        bb.forward(x)                                                          # nothing loaded
    cfg = fcfg.preset("panoptic")
    cfg.RESNET.NUM_LAYERS = 18
    sd = synth.make_backbone_weights(BS.from_cfg(cfg), 3)
    part = {k: v for k, v in sd.items() if not k.startswith("layer3.1.")}
    with pytest.raises(FvpError):
        bb.load_state_dict(part)                                               # a missing layer is named, never skipped
    bb.load_state_dict(sd)
    assert tuple(bb.forward(x).shape) == (1, 15, 24, 24)
    with pytest.raises(FvpError):
        bb.forward(torch.zeros(1, 3, 128, 96))                                 # larger than the object was created for
    with pytest.raises(FvpError):
        bb.forward(torch.zeros(2, 3, 96, 96)[:, :, :80])                       # not a multiple of 32
    bb.close()

