"""Host mirror of the PoseResNet backbone in libfvp_b200.so (lib/models/resnet.py:98-215).

``Backbone`` wraps the C object (``fvp_backbone_*``): it takes the reference's backbone ``state_dict`` unchanged and turns
normalised images ``[n,3,h,w]`` into heat maps ``[n,J,h/4,w/4]`` (``forward``) or into the output of any stage of the
network (``forward_slice``, used by the parity tests).  The reference-facing ``nn.Module`` with the reference's 338
``state_dict`` keys is ``lib/models/resnet.py`` (``models.resnet.get(cfg)``, run/validate.py:69-74)."""
from __future__ import annotations

import ctypes as C
from typing import List, Mapping, Tuple

import numpy as np
import torch

from . import backbone_spec as BS, capi


class Backbone:
    def __init__(self, num_layers: int, num_joints: int, device, max_images: int, max_h: int, max_w: int):
        self.lib = capi.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the backbone runs on a CUDA (sm_100a) device only; there is no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.num_layers, self.num_joints = int(num_layers), int(num_joints)
        self.max_images, self.max_h, self.max_w = int(max_images), int(max_h), int(max_w)
        ctx = C.c_void_p()
        rc = self.lib.fvp_backbone_create(self.num_layers, self.num_joints, self.max_images, self.max_h, self.max_w, idx, C.byref(ctx))
        if rc != capi.FVP_OK:
            raise capi.FvpError(rc, (self.lib.fvp_backbone_last_error(None) or b"?").decode())
        kind, blocks = BS.RESNET_SPEC[self.num_layers]
        exp = 4 if kind == "bottleneck" else 1
        # (channels, log2 of the down-sampling factor) of every stage output: max-pool, residual blocks, deconvs, heat maps
        self.stage_shapes: List[Tuple[int, int]] = [(64, 2)]
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), blocks)):
            self.stage_shapes += [(planes * exp, 2 + li)] * n
        self.stage_shapes += [(256, 4), (256, 3), (256, 2), (self.num_joints, 2)]
        self.ctx = ctx

    def _ck(self, rc: int) -> None:
        if rc != capi.FVP_OK:
            raise capi.FvpError(rc, (self.lib.fvp_backbone_last_error(self.ctx) or b"?").decode())

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.fvp_backbone_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd: Mapping[str, object]) -> None:
        for k, v in sd.items():
            if k.endswith("num_batches_tracked"):
                continue
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            a = np.ascontiguousarray(np.asarray(v), np.float32)
            self._ck(self.lib.fvp_backbone_set_param(self.ctx, k.encode(), a.ctypes.data, a.size))
        self._ck(self.lib.fvp_backbone_finalize(self.ctx))
        assert int(self.lib.fvp_backbone_num_stages(self.ctx)) == len(self.stage_shapes)

    @property
    def num_stages(self) -> int:
        return len(self.stage_shapes)

    def _images(self, images: torch.Tensor) -> torch.Tensor:
        x = images.to(self.device, torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("images must be [n,3,h,w], got %s" % (tuple(x.shape),))
        return x

    def forward_slice(self, images: torch.Tensor, stage: int) -> torch.Tensor:
        """Output (NCHW) of one stage: 0 = after the max-pool, 1..B = after residual block stage-1, B+1..B+3 = after a
        transposed convolution, B+4 = the heat maps."""
        x = self._images(images)
        n, _, h, w = x.shape
        if not 0 <= int(stage) < len(self.stage_shapes):
            raise ValueError("stage %d outside [0, %d)" % (stage, len(self.stage_shapes)))
        ch, sh = self.stage_shapes[int(stage)]
        out = torch.empty((n, ch, h >> sh, w >> sh), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._ck(self.lib.fvp_backbone_forward_slice(self.ctx, x.data_ptr(), n, h, w, int(stage), out.data_ptr(),
                                                         int(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """[n,3,h,w] normalised images -> heat maps [n,J,h/4,w/4] (ResNet.forward, resnet.py:188-201); larger batches than
        ``max_images`` are processed in chunks."""
        x = self._images(images)
        n, _, h, w = x.shape
        out = torch.empty((n, self.num_joints, h // 4, w // 4), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            for i in range(0, n, self.max_images):
                m = min(self.max_images, n - i)
                self._ck(self.lib.fvp_backbone_forward(self.ctx, x[i:i + m].data_ptr(), m, h, w, out[i:i + m].data_ptr(),
                                                       int(torch.cuda.current_stream(self.device).cuda_stream)))
        return out
