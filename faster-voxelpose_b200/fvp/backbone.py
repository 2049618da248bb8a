"""Host mirror of the PoseResNet backbone slice in libfvp_b200.so (lib/models/resnet.py:98-201).

Built so far: stem + max-pool + layer1 (``fvp_backbone_forward_slice``).  The wrapper takes the reference's backbone
``state_dict`` unchanged; keys of layers that are not built yet are accepted and ignored by the library."""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import capi


class BackboneSlice:
    def __init__(self, num_layers: int, device, max_images: int, max_h: int, max_w: int):
        self.lib = capi.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the backbone runs on a CUDA (sm_100a) device only; there is no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.num_layers = int(num_layers)
        self.channels = 256 if self.num_layers >= 50 else 64
        self.blocks = 2 if self.num_layers == 18 else 3
        ctx = C.c_void_p()
        rc = self.lib.fvp_backbone_create(self.num_layers, int(max_images), int(max_h), int(max_w), idx, C.byref(ctx))
        if rc != capi.FVP_OK:
            raise capi.FvpError(rc, (self.lib.fvp_backbone_last_error(None) or b"?").decode())
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

    def forward_slice(self, images: torch.Tensor, stage: int) -> torch.Tensor:
        """[n,3,h,w] fp32 normalised images -> NCHW tap: stage 0 = after the max-pool, k = after layer1 block k-1."""
        x = images.to(self.device, torch.float32).contiguous()
        n, c, h, w = x.shape
        assert c == 3
        ch = 64 if stage == 0 else self.channels
        out = torch.empty((n, ch, h // 4, w // 4), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._ck(self.lib.fvp_backbone_forward_slice(self.ctx, x.data_ptr(), n, h, w, int(stage), out.data_ptr(),
                                                         int(torch.cuda.current_stream(self.device).cuda_stream)))
        return out
