#!/usr/bin/env python
"""Golden vectors of the PoseResNet backbone, produced by the UNMODIFIED reference.

Run in the build container only:    python oracle/gen_golden_backbone.py      # writes tests/golden/backbone_<case>.npz

Builds ``models.resnet.get(cfg)`` from /root/reference/lib (ResNet-50 with the reference's default deconvolution head and
a ResNet-18 variant for the BasicBlock path), loads deterministic conditioned weights, runs it on a small synthetic image
batch with forward hooks on the stem, every residual block and every transposed convolution, and
  1. asserts the state_dict keys / shapes equal ``fvp.backbone_spec.param_table`` (the boundary),
  2. asserts ``oracle.backbone_oracle.forward`` bit-identical on every tap and on the output (the oracle pin),
  3. stores the input (uint8 image, the normalisation constants), the output, per-tap checksums and the weight seed.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
sys.path.insert(0, ROOT)

from fvp import backbone_spec as BS, config as fcfg, synth     # noqa: E402
from oracle import backbone_oracle as BO                       # noqa: E402
from oracle import gen_golden as GG                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def weights_sha(sd) -> str:          # same digest as tests/golden_util.py
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


MEAN, STD = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)   # run/validate.py:45-46


def make_image(seed: int, b: int, h: int, w: int) -> np.ndarray:
    """uint8 [b,h,w,3]: smooth gradients + blobs + noise (structure at every scale the network sees)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((b, h, w, 3), np.float32)
    for i in range(b):
        for c in range(3):
            img[i, :, :, c] = 90 + 60 * np.sin(xx / rng.uniform(9, 40) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(9, 40))
        for _ in range(6):
            cy, cx, s = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(3, 14)
            img[i] += rng.uniform(-80, 80, 3) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))[..., None]
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def normalise(img_u8: np.ndarray) -> torch.Tensor:
    """ToTensor + Normalize as run/validate.py:44-50 builds them: float32 /255, (x - mean) / std, NCHW."""
    x = torch.from_numpy(img_u8).permute(0, 3, 1, 2).float().div(255)
    return (x - torch.from_numpy(MEAN).view(1, 3, 1, 1)) / torch.from_numpy(STD).view(1, 3, 1, 1)


def case(name: str, num_layers: int, seed: int, b: int, h: int, w: int):
    GG._install_reference()
    from models import resnet as ref_resnet
    cfg = fcfg.preset("panoptic")
    cfg.RESNET.NUM_LAYERS = num_layers
    layers = BS.from_cfg(cfg)
    model = ref_resnet.get(cfg).eval()
    ref_sd = model.state_dict()
    table = BS.param_table(layers)
    assert [k for k, _, _ in table] == list(ref_sd.keys()), "key order differs from the reference"
    for k, shape, dtype in table:
        assert tuple(ref_sd[k].shape) == tuple(shape) and str(ref_sd[k].dtype).endswith(dtype), k
    sd_np = synth.make_backbone_weights(layers, seed)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    model.load_state_dict(sd, strict=True)
    taps_ref = {}
    def hook(key):
        return lambda m, i, o: taps_ref.__setitem__(key, o.detach().clone())
    model.maxpool.register_forward_hook(hook("stem"))
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(model, "layer%d" % li)):
            blk.register_forward_hook(hook("layer%d.%d" % (li, bi)))
    for c in layers:
        if c.role == "deconv":                      # tap after BN + ReLU: the ReLU module 2 places after the deconv
            idx = int(c.key.split(".")[1])
            model.deconv_layers[idx + 2].register_forward_hook(hook(c.key))
    img = make_image(seed, b, h, w)
    x = normalise(img)
    with torch.no_grad():
        y_ref = model(x)
        taps_or = {}
        y_or = BO.forward(layers, sd, x, taps_or)
    assert set(taps_ref) == set(taps_or), (sorted(taps_ref), sorted(taps_or))
    for k in taps_ref:
        assert torch.equal(taps_ref[k], taps_or[k]), "oracle differs from the reference at " + k
    assert torch.equal(y_ref, y_or), "oracle output differs from the reference"
    assert torch.isfinite(y_ref).all()
    macs = sum(r["macs"] for r in BS.shapes_and_macs(layers, h, w))
    tap_stats = {k: [float(v.abs().max()), float(v.double().sum())] for k, v in taps_ref.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), num_layers=num_layers, weight_seed=seed, image_u8=img,
                        mean=MEAN, std=STD, output=y_ref.numpy(), keys=np.array(list(ref_sd.keys())),
                        weights_sha256=weights_sha(sd_np),
                        tap_names=np.array(sorted(tap_stats)), tap_absmax=np.array([tap_stats[k][0] for k in sorted(tap_stats)]),
                        tap_sum=np.array([tap_stats[k][1] for k in sorted(tap_stats)]))
    print(name, "out", tuple(y_ref.shape), "absmax %.3f" % float(y_ref.abs().max()), "MACs/image %.2f G" % (macs / 1e9),
          "params %.1f M" % (sum(v.size for v in sd_np.values()) / 1e6))
    return {"num_layers": num_layers, "input": [b, 3, h, w], "output": list(y_ref.shape), "gmacs_per_image": macs / 1e9,
            "out_absmax": float(y_ref.abs().max())}


def main():
    manifest = {"backbone_resnet50": case("backbone_resnet50", 50, 31, 2, 128, 96),
                "backbone_resnet18": case("backbone_resnet18", 18, 32, 1, 96, 160)}
    full = BS.shapes_and_macs(BS.pose_resnet(50, 15), 512, 960)
    manifest["resnet50_960x512_gmacs_per_view"] = sum(r["macs"] for r in full) / 1e9
    with open(os.path.join(OUT, "MANIFEST_backbone.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("ResNet-50 @ 960x512: %.1f GMAC per view" % manifest["resnet50_960x512_gmacs_per_view"])


if __name__ == "__main__":
    main()
