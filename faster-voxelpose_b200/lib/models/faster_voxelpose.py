"""Drop-in for the reference's ``lib/models/faster_voxelpose.py``: same ``get(cfg)``, same
``forward(backbone, views, meta, targets, input_heatmaps, cameras, resize_transform)`` signature and
5-tuple result (faster_voxelpose.py:34,105), same ``state_dict`` keys (485 tensors, strict loading of a
reference ``model_best.pth.tar``), same ``model.pose_net`` / ``model.joint_net`` attribute tree - but
``forward`` is one call into libfvp_b200.so (hand-written sm_100a kernels), not a PyTorch graph.

Inference only: the training branch (faster_voxelpose.py:51-98) is out of scope and raises.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # faster-voxelpose_b200/
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from fvp import netspec  # noqa: E402
from fvp.engine import Engine  # noqa: E402


class _Holder(nn.Module):
    """Parameter container; nesting reproduces the reference's state_dict key paths."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of the B200 model hold parameters only; call the top-level model")


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


class FasterVoxelPoseNet(nn.Module):
    def __init__(self, cfg, max_batch: int = None, max_sequences: int = 8):
        super().__init__()
        self.cfg = cfg
        self.max_people = cfg.CAPTURE_SPEC.MAX_PEOPLE
        self.num_joints = cfg.DATASET.NUM_JOINTS
        self.device = torch.device(cfg.DEVICE)
        self._max_batch = int(max_batch if max_batch is not None else max(int(cfg.TEST.BATCH_SIZE), 1))
        self._max_sequences = int(max_sequences)
        for key, shape, dtype in netspec.param_table(int(cfg.DATASET.NUM_JOINTS), int(cfg.NETWORK.NUM_CHANNEL_JOINT_FEAT),
                                                     int(cfg.NETWORK.NUM_CHANNEL_JOINT_HIDDEN)):
            _register(self, key, shape, dtype)
        self._engine = None
        self._weights_tag = None
        self._tensor_cache = None

    # ---- engine management ---------------------------------------------------------------------
    # The packed device weights must follow the module's parameters (load_state_dict, .to(), in-place edits).  Every
    # forward compares a tag of (storage address, version counter) per tensor; walking state_dict() for it cost 1.2 ms per
    # call - twice the GPU time of a frame - so the 485 tensor objects are cached and the cache is dropped whenever the
    # module machinery may have replaced them (_apply: .to()/.cuda()/.float(); load_state_dict(assign=True)).
    def _tensors(self):
        if self._tensor_cache is None:
            self._tensor_cache = list(self.state_dict(keep_vars=True).values())
        return self._tensor_cache

    def _apply(self, fn, *args, **kwargs):
        self._tensor_cache = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._tensor_cache = None
        return out

    def _tag(self):
        return tuple((t.data_ptr(), t._version) for t in self._tensors())

    def engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("the B200 model runs on a CUDA device only (model.to('cuda:0')); there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.cfg, dev, self._max_batch, self._max_sequences)
            self._weights_tag = None
        tag = self._tag()
        if tag != self._weights_tag:
            self._engine.load_state_dict(self.state_dict())
            self._weights_tag = tag
        return self._engine

    # ---- reference signature -------------------------------------------------------------------
    def forward(self, backbone=None, views=None, meta=None, targets=None, input_heatmaps=None, cameras=None,
                resize_transform=None):
        if self.training:
            raise NotImplementedError("training branch (faster_voxelpose.py:51-98) is out of scope of the B200 path; "
                                      "call model.eval()")
        if views is not None:      # faster_voxelpose.py:36-38 (backbone: any callable; models.resnet.get(cfg) is the B200 one)
            num_views = views.shape[1]
            input_heatmaps = torch.stack([backbone(views[:, c]) for c in range(num_views)], dim=1)
        eng = self.engine()
        B = input_heatmaps.shape[0]
        if len(set(meta["seq"][:B])) > eng.max_sequences:
            # every calibration of a batch needs its own device slot (camera block + sample-grid caches); recycling one
            # while the batch is assembled would silently project some frames with another sequence's cameras
            raise ValueError("batch mixes %d sequences but the model was built for max_sequences=%d"
                             % (len(set(meta["seq"][:B])), eng.max_sequences))
        slots, slot_of = [], {}
        for i in range(B):
            seq = meta["seq"][i]
            if seq not in slot_of:     # one calibration lookup (hash of the camera values) per distinct sequence of the batch
                assert seq in cameras.keys(), "missing camera parameters for the current sequence"
                assert len(cameras[seq]) == input_heatmaps.shape[1], "inconsistent number of cameras"
                slot_of[seq] = eng.sequence_slot(cameras[seq], resize_transform)
            slots.append(slot_of[seq])
        if B > eng.max_batch:      # chunk oversized batches
            outs = [eng.forward(input_heatmaps[i:i + eng.max_batch], slots[i:i + eng.max_batch])
                    for i in range(0, B, eng.max_batch)]
            fused = torch.cat([o[0] for o in outs], 0)
            plane = torch.cat([o[1] for o in outs], 1)
            centers = torch.cat([o[2] for o in outs], 0)
        else:
            fused, plane, centers = eng.forward(input_heatmaps, slots)
        return fused, plane, centers, input_heatmaps, None


def get(cfg):
    return FasterVoxelPoseNet(cfg)
