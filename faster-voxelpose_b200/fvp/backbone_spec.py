"""Declarative description of the PoseResNet backbone (``TEST_HEATMAP_SRC = 'image'``).

The reference builds it as nested ``nn.Module`` classes (lib/models/resnet.py:22-215); ``faster_voxelpose.py:36-38`` calls
``backbone(views[:, c])`` once per camera: ``[B,3,h,w]`` normalised images -> ``[B,J,h/4,w/4]`` heat maps.  Like
``fvp.netspec`` for the voxel networks, this module turns it into a *table*: one :class:`Layer` per convolution in
execution order, keyed by the reference ``state_dict`` prefix, so that the parameter holder, the weight generator, the
CPU oracle (oracle/backbone_oracle.py) and - next - the CUDA engine enumerate one list.

Status: the table, the oracle and its pin against the unmodified reference are in, and the whole network runs on the GPU
(csrc/fvp_backbone.cu, fvp/backbone.py, lib/models/resnet.py, tests/test_gpu_backbone.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

RESNET_SPEC = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]), 50: ("bottleneck", [3, 4, 6, 3]),
               101: ("bottleneck", [3, 4, 23, 3]), 152: ("bottleneck", [3, 8, 36, 3])}       # resnet.py:204-208


@dataclass(frozen=True)
class Layer:
    key: str            # state_dict prefix of the convolution
    cin: int
    cout: int
    k: int
    stride: int
    pad: int
    bn: str             # state_dict prefix of the BatchNorm that follows ("" = none)
    bias: bool          # the convolution has a bias
    relu: bool          # ReLU directly after conv+BN (before any residual add)
    transposed: bool = False
    out_pad: int = 0
    role: str = ""      # "stem" | "conv" | "downsample" | "last" (block's last conv: residual add + ReLU follow) | "deconv" | "final"


def pose_resnet(num_layers: int = 50, num_joints: int = 15, deconv_filters: Sequence[int] = (256, 256, 256),
                deconv_kernels: Sequence[int] = (4, 4, 4), final_kernel: int = 1, deconv_bias: bool = False) -> List[Layer]:
    """Layers of ``ResNet(block, layers, cfg)`` in execution order (resnet.py:98-200); within a block the downsample
    branch is listed after the block's own convolutions, which is also the reference's registration order."""
    kind, blocks = RESNET_SPEC[num_layers]
    exp = 4 if kind == "bottleneck" else 1
    L: List[Layer] = [Layer("conv1", 3, 64, 7, 2, 3, "bn1", False, True, role="stem")]
    inplanes = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), blocks), start=1):
        for b in range(n):
            stride = 2 if (b == 0 and li > 1) else 1
            p = "layer%d.%d" % (li, b)
            if kind == "bottleneck":                                          # resnet.py:57-95 (stride on the 3x3)
                L.append(Layer(p + ".conv1", inplanes, planes, 1, 1, 0, p + ".bn1", False, True, role="conv"))
                L.append(Layer(p + ".conv2", planes, planes, 3, stride, 1, p + ".bn2", False, True, role="conv"))
                L.append(Layer(p + ".conv3", planes, planes * 4, 1, 1, 0, p + ".bn3", False, False, role="last"))
            else:                                                             # resnet.py:22-54
                L.append(Layer(p + ".conv1", inplanes, planes, 3, stride, 1, p + ".bn1", False, True, role="conv"))
                L.append(Layer(p + ".conv2", planes, planes, 3, 1, 1, p + ".bn2", False, False, role="last"))
            if b == 0 and (stride != 1 or inplanes != planes * exp):          # resnet.py:131-137
                L.append(Layer(p + ".downsample.0", inplanes, planes * exp, 1, stride, 0, p + ".downsample.1", False, False,
                               role="downsample"))
            inplanes = planes * exp
    for i, (f, k) in enumerate(zip(deconv_filters, deconv_kernels)):         # resnet.py:148-186
        pad, out_pad = {4: (1, 0), 3: (1, 1), 2: (0, 0)}[int(k)]
        L.append(Layer("deconv_layers.%d" % (3 * i), inplanes, int(f), int(k), 2, pad, "deconv_layers.%d" % (3 * i + 1),
                       bool(deconv_bias), True, transposed=True, out_pad=out_pad, role="deconv"))
        inplanes = int(f)
    L.append(Layer("final_layer", inplanes, num_joints, int(final_kernel), 1, 1 if final_kernel == 3 else 0, "", True, False,
                   role="final"))
    return L


def from_cfg(cfg) -> List[Layer]:
    r = cfg.RESNET
    return pose_resnet(int(r.NUM_LAYERS), int(cfg.DATASET.NUM_JOINTS), list(r.NUM_DECONV_FILTERS), list(r.NUM_DECONV_KERNELS),
                       int(r.FINAL_CONV_KERNEL), bool(r.DECONV_WITH_BIAS))


def weight_shape(c: Layer) -> Tuple[int, ...]:
    return (c.cin, c.cout, c.k, c.k) if c.transposed else (c.cout, c.cin, c.k, c.k)


def param_table(layers: List[Layer]) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(state_dict key, shape, dtype) in the reference's registration order (conv, then its BatchNorm)."""
    rows: List[Tuple[str, Tuple[int, ...], str]] = []
    for c in layers:
        rows.append((c.key + ".weight", weight_shape(c), "float32"))
        if c.bias:
            rows.append((c.key + ".bias", (c.cout,), "float32"))
        if c.bn:
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                rows.append((c.bn + "." + leaf, (c.cout,), "float32"))
            rows.append((c.bn + ".num_batches_tracked", (), "int64"))
    return rows


def out_hw(c: Layer, h: int, w: int) -> Tuple[int, int]:
    if c.transposed:
        f = lambda n: (n - 1) * c.stride - 2 * c.pad + c.k + c.out_pad
    else:
        f = lambda n: (n + 2 * c.pad - c.k) // c.stride + 1
    return f(h), f(w)


def shapes_and_macs(layers: List[Layer], h: int, w: int) -> List[Dict]:
    """Per layer: input / output extent and multiply-accumulates for one ``[3,h,w]`` image (the 3x3 stride-2 max-pool
    sits between the stem and layer1; downsample branches see their block's input)."""
    rows, cur, block_in = [], (h, w), None
    for c in layers:
        if c.key.endswith(".conv1") and c.key.startswith("layer"):
            block_in = cur
        src = block_in if c.role == "downsample" else cur
        dst = out_hw(c, *src)
        macs = (src[0] * src[1] if c.transposed else dst[0] * dst[1]) * c.cin * c.cout * c.k * c.k
        rows.append({"key": c.key, "in": src, "out": dst, "macs": macs})
        if c.role == "stem":
            cur = ((dst[0] + 2 - 3) // 2 + 1, (dst[1] + 2 - 3) // 2 + 1)      # MaxPool2d(3, 2, 1), resnet.py:108
        elif c.role != "downsample":
            cur = dst
    return rows
