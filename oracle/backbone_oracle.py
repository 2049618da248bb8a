"""CPU oracle of the PoseResNet backbone - TEST INFRASTRUCTURE ONLY: tests and golden generators may
import it, the product (faster-voxelpose_b200/) never does.

Functional, state_dict-driven restatement of ``ResNet.forward`` (lib/models/resnet.py:188-201) with the same
PyTorch-CPU primitives in the same order: conv -> BatchNorm(eval) -> ReLU, ``MaxPool2d(3, 2, 1)`` after the stem
(resnet.py:108,192), blocks with ``out += residual; relu`` (resnet.py:40-54, 74-95), transposed convolutions + BN + ReLU
(resnet.py:163-186), final convolution.  Pinned bit-identical to the unmodified reference by oracle/gen_golden_backbone.py
(every block output and the result), re-checked against tests/golden/backbone_*.npz by tests/test_backbone_oracle.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from fvp.backbone_spec import Layer


def _conv_bn(sd: Dict[str, Tensor], c: Layer, x: Tensor) -> Tensor:
    bias = sd.get(c.key + ".bias") if c.bias else None
    if c.transposed:
        y = F.conv_transpose2d(x, sd[c.key + ".weight"], bias, stride=c.stride, padding=c.pad, output_padding=c.out_pad)
    else:
        y = F.conv2d(x, sd[c.key + ".weight"], bias, stride=c.stride, padding=c.pad)
    if c.bn:
        y = F.batch_norm(y, sd[c.bn + ".running_mean"], sd[c.bn + ".running_var"], sd[c.bn + ".weight"], sd[c.bn + ".bias"],
                         training=False, momentum=0.1, eps=1e-5)
    return F.relu(y) if c.relu else y


def forward(layers: List[Layer], sd: Dict[str, Tensor], x: Tensor, taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """[B,3,h,w] -> [B,J,h/4,w/4]; ``taps`` (if given) receives the output of the stem, of every residual block and of
    every transposed convolution, keyed like the reference's sub-modules."""
    i = 0
    block_in = residual = None
    while i < len(layers):
        c = layers[i]
        if c.role == "stem":
            x = F.max_pool2d(_conv_bn(sd, c, x), kernel_size=3, stride=2, padding=1)
            if taps is not None:
                taps["stem"] = x
        elif c.role == "conv":
            if c.key.endswith(".conv1"):
                block_in = x
            x = _conv_bn(sd, c, x)
        elif c.role == "last":
            y = _conv_bn(sd, c, x)
            residual = block_in
            if i + 1 < len(layers) and layers[i + 1].role == "downsample":
                residual = _conv_bn(sd, layers[i + 1], block_in)
                i += 1
            y = y + residual            # the reference's in-place `out += residual` computes the same sum
            x = F.relu(y)
            if taps is not None:
                taps[c.key.rsplit(".", 1)[0]] = x
        elif c.role in ("deconv", "final"):
            x = _conv_bn(sd, c, x)
            if taps is not None and c.role == "deconv":
                taps[c.key] = x
        else:
            raise ValueError("unexpected layer role %r at %s" % (c.role, c.key))
        i += 1
    return x
