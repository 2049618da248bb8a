"""Drop-in for the reference's ``lib/models/resnet.py``: same ``get(cfg)`` (resnet.py:211-215), same ``state_dict`` keys
(338 tensors for ResNet-50: conv1 / bn1 / layer1-4 / deconv_layers / final_layer), same ``forward(x)`` - normalised images
``[B,3,h,w]`` -> heat maps ``[B,J,h/4,w/4]`` (resnet.py:188-201) - but ``forward`` is one call into libfvp_b200.so
(``fvp_backbone_forward``: stem and max-pool on CUDA cores, every other layer on the tcgen05 engine).

``run/validate.py:69-74`` builds it with ``eval('models.' + config.BACKBONE + '.get')(config)``, loads
``config.NETWORK.PRETRAINED_BACKBONE`` into it, moves it to the device and puts it in eval mode; ``FasterVoxelPoseNet.forward``
then calls ``backbone(views[:, c])`` once per camera (faster_voxelpose.py:36-38).  Inference only, CUDA only.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # faster-voxelpose_b200/
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from fvp import backbone_spec as BS  # noqa: E402
from fvp.backbone import Backbone  # noqa: E402


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of the B200 backbone hold parameters only; call the top-level model")


def _register(root: nn.Module, dotted: str, shape, dtype: str) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Holder())
        mod = getattr(mod, p)
    leaf = parts[-1]
    if dtype == "int64":
        mod.register_buffer(leaf, torch.zeros(shape, dtype=torch.long))
    elif leaf in ("running_mean", "running_var"):
        mod.register_buffer(leaf, torch.ones(shape) if leaf == "running_var" else torch.zeros(shape))
    else:
        init = torch.ones(shape) if (leaf == "weight" and len(shape) == 1) else torch.zeros(shape)
        mod.register_parameter(leaf, nn.Parameter(init, requires_grad=False))


class ResNet(nn.Module):
    def __init__(self, cfg, max_images: int = None):
        super().__init__()
        self.cfg = cfg
        self.layers = BS.from_cfg(cfg)
        for c in self.layers:
            if c.transposed and c.k != 4:
                raise NotImplementedError("the B200 backbone supports 4x4 transposed convolutions (NUM_DECONV_KERNELS) only")
            if c.role == "final" and c.k != 1:
                raise NotImplementedError("the B200 backbone supports FINAL_CONV_KERNEL = 1 only")
        for key, shape, dtype in BS.param_table(self.layers):
            _register(self, key, shape, dtype)
        self._max_images = int(max_images if max_images is not None else max(int(cfg.TEST.BATCH_SIZE), 1))
        self._engine = None
        self._engine_key = None
        self._weights_tag = None

    def _tag(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def engine(self, n: int, h: int, w: int) -> Backbone:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("the B200 backbone runs on a CUDA device only (backbone.to('cuda:0')); there is no CPU path")
        key = (dev, max(n, self._max_images), h, w)
        if self._engine is None or self._engine_key[0] != dev or self._engine_key[1] < n or self._engine_key[2:] != (h, w):
            if self._engine is not None:
                self._engine.close()
            self._engine = Backbone(int(self.cfg.RESNET.NUM_LAYERS), int(self.cfg.DATASET.NUM_JOINTS), dev, key[1], h, w)
            self._engine_key = key
            self._weights_tag = None
        tag = self._tag()
        if tag != self._weights_tag:
            self._engine.load_state_dict(self.state_dict())
            self._weights_tag = tag
        return self._engine

    def forward(self, x):
        if self.training:
            raise NotImplementedError("the B200 backbone is inference only (run/validate.py:74 freezes it with .eval())")
        n, _, h, w = x.shape
        return self.engine(int(n), int(h), int(w)).forward(x)


def get(cfg):
    return ResNet(cfg)
