"""Host-side engine over the C ABI: owns one ``fvp_ctx`` on one GPU.

PyTorch is used for what the task allows it for - device memory, streams, ``torch.distributed`` -
never for arithmetic on the hot path.  Every compute call goes through ``libfvp_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import capi


def _f32_ptr(a: np.ndarray) -> int:
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def camera_rows(cams: Sequence[Mapping]) -> np.ndarray:
    """[V,21] fp32 = R(9) T(3) fx fy cx cy k(3) p(2): float64 calibration rounded to fp32 exactly as
    ``unfold_camera_param`` does (lib/utils/cameras.py:11-18)."""
    out = np.zeros((len(cams), 21), np.float32)
    for i, c in enumerate(cams):
        out[i, 0:9] = np.asarray(c["R"], np.float64).reshape(9).astype(np.float32)
        out[i, 9:12] = np.asarray(c["T"], np.float64).reshape(3).astype(np.float32)
        out[i, 12:16] = np.array([c["fx"], c["fy"], c["cx"], c["cy"]], np.float64).astype(np.float32)
        out[i, 16:19] = np.asarray(c["k"], np.float64).reshape(3).astype(np.float32)
        out[i, 19:21] = np.asarray(c["p"], np.float64).reshape(2).astype(np.float32)
    return out


def reference_axes(cfg) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Voxel-centre tables with the reference's own expression ``torch.linspace(-S/2, S/2, N) + centre``
    (project_whole.py:34-40, project_individual.py:25,37-41), evaluated by this host's PyTorch.  Passing
    them to the library makes the in-kernel projection bit-identical to a CPU run of the reference on
    the same host (torch's vectorised CPU linspace is not the scalar start+i*step formula)."""
    whole = [float(v) for v in cfg.CAPTURE_SPEC.SPACE_SIZE]
    ctr = [float(v) for v in cfg.CAPTURE_SPEC.SPACE_CENTER]
    ind = [float(v) for v in cfg.INDIVIDUAL_SPEC.SPACE_SIZE]
    nb = [int(v) for v in cfg.CAPTURE_SPEC.VOXELS_PER_AXIS]
    iv = [int(v) for v in cfg.INDIVIDUAL_SPEC.VOXELS_PER_AXIS]
    whole_t, ind_t = torch.tensor(whole), torch.tensor(ind)
    fine = ((whole_t / ind_t * (torch.tensor(iv, dtype=torch.int32) - 1)).int() + 1).tolist()

    def ax(size, c, n):
        return (torch.linspace(-size / 2, size / 2, int(n)) + c).numpy().astype(np.float32)

    coarse = np.concatenate([ax(whole[d], ctr[d], nb[d]) for d in range(3)])
    fine_a = np.concatenate([ax(whole[d], ctr[d], fine[d]) for d in range(3)])
    indiv = np.concatenate([ax(ind[d], ctr[d], iv[d]) for d in range(3)])
    return coarse, fine_a, indiv


class Engine:
    """One ``fvp_ctx``.  With ``parent`` given the context is a *lane* of that engine (``fvp_create_lane``): it owns
    workspaces, streams and a CUDA graph for one more frame in flight and shares the parent's weights, axis tables,
    calibrations and sample-grid caches; parameters and calibrations are then managed through the parent only."""

    def __init__(self, cfg, device: Optional[torch.device] = None, max_batch: int = 8, max_sequences: int = 8,
                 axes: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None, parent: Optional["Engine"] = None):
        self.lib = capi.load()
        self.parent = parent
        if parent is not None:
            cfg, device, max_sequences = parent.cfg, parent.device, parent.max_sequences
        if not torch.cuda.is_available():
            raise RuntimeError("faster-voxelpose_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device if device is not None else getattr(cfg, "DEVICE", "cuda:0"))
        if self.device.type != "cuda":
            raise RuntimeError("DEVICE must be a cuda device, got %s" % self.device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.cfg = cfg
        c = capi.FvpConfig()
        ds, cs, isp = cfg.DATASET, cfg.CAPTURE_SPEC, cfg.INDIVIDUAL_SPEC
        c.num_views, c.num_joints = int(ds.CAMERA_NUM), int(ds.NUM_JOINTS)
        c.hm_w, c.hm_h = int(ds.HEATMAP_SIZE[0]), int(ds.HEATMAP_SIZE[1])
        c.image_w, c.image_h = float(ds.IMAGE_SIZE[0]), float(ds.IMAGE_SIZE[1])
        c.ori_w, c.ori_h = float(ds.ORI_IMAGE_SIZE[0]), float(ds.ORI_IMAGE_SIZE[1])
        for d in range(3):
            c.space_size[d] = float(cs.SPACE_SIZE[d])
            c.space_center[d] = float(cs.SPACE_CENTER[d])
            c.voxels[d] = int(cs.VOXELS_PER_AXIS[d])
            c.ind_space_size[d] = float(isp.SPACE_SIZE[d])
            c.ind_voxels[d] = int(isp.VOXELS_PER_AXIS[d])
        c.max_people, c.min_score = int(cs.MAX_PEOPLE), float(cs.MIN_SCORE)
        c.beta = float(cfg.NETWORK.BETA)
        c.feat_channels = int(cfg.NETWORK.NUM_CHANNEL_JOINT_FEAT)
        c.hidden_channels = int(cfg.NETWORK.NUM_CHANNEL_JOINT_HIDDEN)
        c.max_batch, c.max_sequences = int(max_batch), int(max_sequences)
        self.c = c
        self.V, self.J, self.P = c.num_views, c.num_joints, c.max_people
        self.H, self.W = c.hm_h, c.hm_w
        self.X, self.Y, self.Z = c.voxels[0], c.voxels[1], c.voxels[2]
        self.max_batch, self.max_sequences = int(max_batch), int(max_sequences)
        ctx = C.c_void_p()
        self._seq_slots: Dict[str, int] = {}     # calibration fingerprint -> slot
        self._seq_lru: List[str] = []
        self._params_loaded = False
        self._lanes: List["Engine"] = []
        if parent is not None:
            capi.check(self.lib, parent.ctx, self.lib.fvp_create_lane(parent.ctx, int(max_batch), C.byref(ctx)))
            self.ctx = ctx
            self.fine = list(parent.fine)
            parent._lanes.append(self)
            return
        rc = self.lib.fvp_create(C.byref(c), idx, C.byref(ctx))
        if rc != capi.FVP_OK:
            raise capi.FvpError(rc, (self.lib.fvp_last_error(None) or b"?").decode())
        self.ctx = ctx
        coarse, fine, indiv = axes if axes is not None else reference_axes(cfg)
        fv = (C.c_int32 * 3)()
        self._ck(self.lib.fvp_fine_voxels(self.ctx, C.byref(fv)))
        self.fine = [int(v) for v in fv]
        assert coarse.size == self.X + self.Y + self.Z and fine.size == sum(self.fine) and indiv.size == 192
        self._ck(self.lib.fvp_set_axes(self.ctx, _f32_ptr(np.ascontiguousarray(coarse, np.float32)),
                                       _f32_ptr(np.ascontiguousarray(fine, np.float32)),
                                       _f32_ptr(np.ascontiguousarray(indiv, np.float32))))

    # ------------------------------------------------------------------------------------------
    def _ck(self, rc: int) -> None:
        capi.check(self.lib, self.ctx, rc)

    def new_lane(self, max_batch: Optional[int] = None) -> "Engine":
        """A lane of this engine: own workspaces / stream / graph, shared weights, calibrations and sample grids."""
        assert self.parent is None, "lanes hang off a root engine"
        return Engine(None, parent=self, max_batch=int(max_batch or self.max_batch))

    def close(self) -> None:
        if getattr(self, "ctx", None):
            for lane in list(getattr(self, "_lanes", [])):      # lanes point into this context: they go first
                lane.close()
            self.lib.fvp_destroy(self.ctx)
            self.ctx = None
            if self.parent is not None and self in self.parent._lanes:
                self.parent._lanes.remove(self)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- parameters --------------------------------------------------------------------------
    def param_names(self) -> List[str]:
        n = self.lib.fvp_param_count(self.ctx)
        return [self.lib.fvp_param_name(self.ctx, i).decode() for i in range(n)]

    def load_state_dict(self, sd: Mapping[str, object]) -> None:
        """Reference ``state_dict`` (torch tensors or numpy arrays, any device) -> folded, packed, uploaded."""
        if self.parent is not None:
            raise RuntimeError("load_state_dict: a lane shares its parent's weights; load them into the parent engine")
        names = self.param_names()
        missing = [k for k in names if k not in sd]
        unexpected = [k for k in sd if k not in set(names)]
        if missing or unexpected:
            raise RuntimeError("Error(s) in loading state_dict: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
        for k in names:
            v = sd[k]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            v = np.asarray(v)
            if k.endswith("num_batches_tracked"):
                continue
            a = np.ascontiguousarray(v, np.float32)
            self._ck(self.lib.fvp_set_param(self.ctx, k.encode(), _f32_ptr(a), a.size))
        self._ck(self.lib.fvp_finalize_params(self.ctx))
        self._params_loaded = True

    # ---- calibration -------------------------------------------------------------------------
    def sequence_slot(self, cams: Sequence[Mapping], resize) -> int:
        """Slot of a calibration; uploads it on first use.  Keyed on the calibration *values* (the
        reference keys its grid cache on the sequence name only, SURVEY.md 3.3 - a stale-cache hazard)."""
        if self.parent is not None:
            return self.parent.sequence_slot(cams, resize)
        rows = camera_rows(cams)
        rz = np.ascontiguousarray(np.asarray(resize.detach().cpu() if isinstance(resize, torch.Tensor) else resize,
                                             np.float64).astype(np.float32).reshape(6))
        key = hashlib.sha1(rows.tobytes() + rz.tobytes()).hexdigest()
        if key in self._seq_slots:
            self._seq_lru.remove(key)
            self._seq_lru.append(key)
            return self._seq_slots[key]
        if len(self._seq_slots) < self.max_sequences:
            slot = len(self._seq_slots)
        else:
            old = self._seq_lru.pop(0)
            slot = self._seq_slots.pop(old)
        self._ck(self.lib.fvp_set_sequence(self.ctx, slot, _f32_ptr(rows), rows.shape[0], _f32_ptr(rz)))
        self._seq_slots[key] = slot
        self._seq_lru.append(key)
        return slot

    # ---- forward -----------------------------------------------------------------------------
    def _slots_arr(self, slots: Sequence[int]):
        return (C.c_int32 * len(slots))(*[int(s) for s in slots])

    def _check_hm(self, hm: torch.Tensor) -> int:
        if hm.dim() != 5 or tuple(hm.shape[1:]) != (self.V, self.J, self.H, self.W):
            raise ValueError("input_heatmaps must be [B,%d,%d,%d,%d], got %s" % (self.V, self.J, self.H, self.W, tuple(hm.shape)))
        if hm.shape[0] > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (hm.shape[0], self.max_batch))
        return int(hm.shape[0])

    def forward(self, heatmaps: torch.Tensor, slots: Sequence[int], out_fused: Optional[torch.Tensor] = None):
        """[B,V,J,H,W] fp32 on this device -> (fused_poses [B,P,J,5], plane_poses [3,B,P,J,2], proposal_centers [B,P,7]).
        ``out_fused``: optional preallocated contiguous [B,P,J,5] destination (e.g. this frame's rows of a rank's shard
        buffer, so the multi-GPU gather needs no extra copy)."""
        B = self._check_hm(heatmaps)
        hm = heatmaps.to(device=self.device, dtype=torch.float32).contiguous()
        if out_fused is None:
            fused = torch.empty((B, self.P, self.J, 5), device=self.device, dtype=torch.float32)
        else:
            fused = out_fused
            assert fused.is_contiguous() and fused.dtype == torch.float32 and fused.device == self.device
            assert tuple(fused.shape) == (B, self.P, self.J, 5)
        plane = torch.empty((3, B, self.P, self.J, 2), device=self.device, dtype=torch.float32)
        centers = torch.empty((B, self.P, 7), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            self._ck(self.lib.fvp_forward(self.ctx, hm.data_ptr(), B, self._slots_arr(slots), fused.data_ptr(),
                                          plane.data_ptr(), centers.data_ptr(), self._stream()))
        return fused, plane, centers

    def forward_host(self, heatmaps: torch.Tensor, slots: Sequence[int], out: Optional[Tuple[torch.Tensor, ...]] = None):
        """Same through host buffers (H2D + forward + D2H inside the call, synchronous)."""
        B = self._check_hm(heatmaps)
        assert heatmaps.device.type == "cpu" and heatmaps.dtype == torch.float32 and heatmaps.is_contiguous()
        if out is None:
            out = (torch.empty((B, self.P, self.J, 5)).pin_memory(), torch.empty((3, B, self.P, self.J, 2)).pin_memory(),
                   torch.empty((B, self.P, 7)).pin_memory())
        with torch.cuda.device(self.device):
            self._ck(self.lib.fvp_forward_host(self.ctx, heatmaps.data_ptr(), B, self._slots_arr(slots),
                                               out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), self._stream()))
        return out

    def new_host_outputs(self, B: int) -> Tuple[torch.Tensor, ...]:
        """Pinned (fused_poses, plane_poses, proposal_centers) host tensors for the host entry points."""
        return (torch.empty((B, self.P, self.J, 5)).pin_memory(), torch.empty((3, B, self.P, self.J, 2)).pin_memory(),
                torch.empty((B, self.P, 7)).pin_memory())

    def submit_host(self, heatmaps: torch.Tensor, slots: Sequence[int], out: Tuple[torch.Tensor, ...]) -> int:
        """Pipelined host entry (fvp_submit_host): enqueue H2D -> forward -> D2H and return a ticket without waiting.
        At most two tickets may be outstanding; `out` must not be reused until wait(ticket) returned."""
        B = self._check_hm(heatmaps)
        assert heatmaps.device.type == "cpu" and heatmaps.dtype == torch.float32 and heatmaps.is_contiguous()
        assert all(t.device.type == "cpu" and t.is_contiguous() for t in out)
        ticket = C.c_longlong(-1)
        self._ck(self.lib.fvp_submit_host(self.ctx, heatmaps.data_ptr(), B, self._slots_arr(slots), out[0].data_ptr(),
                                          out[1].data_ptr(), out[2].data_ptr(), C.byref(ticket)))
        return int(ticket.value)

    def wait(self, ticket: int) -> None:
        self._ck(self.lib.fvp_wait(self.ctx, int(ticket)))

    def stream_host(self, frames, slots_of=None):
        """Run an iterable of host heat-map batches through the two-deep pipeline; yields (index, outputs) in order.
        Mirrors the reference's validation loop (lib/core/function.py: one model call per loader batch) with the
        copy of batch i+1 overlapping the kernels of batch i.  The yielded tensors are reused two batches later."""
        bufs, pending = {}, []
        for i, hm in enumerate(frames):
            B = hm.shape[0]
            slots = slots_of(i) if slots_of else [0] * B
            key = (B, i & 1)
            if key not in bufs:
                bufs[key] = self.new_host_outputs(B)
            pending.append((i, self.submit_host(hm, slots, bufs[key]), bufs[key]))
            if len(pending) == 2:
                j, t, o = pending.pop(0)
                self.wait(t)
                yield j, o
        for j, t, o in pending:
            self.wait(t)
            yield j, o

    def use_cuda_graph(self, on: bool = True) -> None:
        self._ck(self.lib.fvp_use_cuda_graph(self.ctx, 1 if on else 0))

    def check_range(self, synchronize: bool = True) -> None:
        """Raise ``FvpError(FVP_E_RANGE)`` if a convolution stored an activation outside the fp16 range of the default
        hi/lo engine since the last check (the host entry points check by themselves; the stream-ordered ``forward`` cannot
        without a synchronise)."""
        if synchronize:
            torch.cuda.synchronize(self.device)
        self._ck(self.lib.fvp_check_range(self.ctx))

    def fp16_fallback_layers(self) -> int:
        """Conv layers whose BN-folded weights left the fp16 range and therefore run on the 3xTF32 engine."""
        return int(self.lib.fvp_fp16_fallback_layers(self.ctx))

    def set_conv_mode(self, mode: int) -> None:
        """0 = fp32 CUDA-core convolutions, 1 = tcgen05 3xTF32, 2 = tcgen05 fp16 hi/lo split (default)."""
        self._ck(self.lib.fvp_set_conv_mode(self.ctx, int(mode)))

    def set_latency_mode(self, mode: int) -> None:
        """1 = latency kernels (C2CNet on one 8-CTA cluster per column), 0 = throughput kernels (one CTA per column),
        -1 = automatic: latency while this engine runs alone, throughput once lanes share its device."""
        self._ck(self.lib.fvp_set_latency_mode(self.ctx, int(mode)))

    def set_profiling(self, on: bool) -> None:
        self._ck(self.lib.fvp_set_profiling(self.ctx, 1 if on else 0))

    def stage_times_ms(self) -> List[float]:
        t = (C.c_float * 9)()
        self._ck(self.lib.fvp_stage_times_ms(self.ctx, C.byref(t)))
        return [float(v) for v in t]

    def last_launch_count(self) -> int:
        return int(self.lib.fvp_last_launch_count(self.ctx))

    def algorithmic_bytes(self, num_valid: int) -> Tuple[float, float]:
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.fvp_algorithmic_bytes(self.ctx, int(num_valid), C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- stage entry points (tests / profiling) -------------------------------------------------
    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(shape, device=self.device, dtype=dtype)

    def debug_conv(self, x_nhwc: torch.Tensor, weight: np.ndarray, bias: np.ndarray, relu: bool, mode: int, repeat: int = 1,
                   want_ms: bool = False):
        x = x_nhwc.to(self.device, torch.float32).contiguous()
        n, H, W, cin = x.shape
        w = np.ascontiguousarray(weight, np.float32)
        b = np.ascontiguousarray(bias, np.float32)
        cout, k = w.shape[0], w.shape[2]
        out = torch.zeros((n, H, W, (cout + 3) // 4 * 4), device=self.device)
        ms = (C.c_float * 13)()                  # [0] = ms per launch; [1..12] = role counters when mode has bit 0x100
        self._ck(self.lib.fvp_debug_conv(self.ctx, x.data_ptr(), n, H, W, cin, w.ctypes.data, b.ctypes.data, cout, k,
                                         1 if relu else 0, int(mode), out.data_ptr(), int(repeat),
                                         C.cast(ms, C.POINTER(C.c_float)), self._stream()))
        if int(mode) & 0x100:
            return out[..., :cout], float(ms[0]), [float(v) for v in ms[1:]]
        return (out[..., :cout], float(ms[0])) if want_ms else out[..., :cout]

    def debug_project(self, slot: int, points: torch.Tensor):
        p = points.to(self.device, torch.float32).contiguous()
        n = p.shape[0]
        ix, iy = self._new(self.V, n), self._new(self.V, n)
        self._ck(self.lib.fvp_debug_project(self.ctx, int(slot), p.data_ptr(), n, ix.data_ptr(), iy.data_ptr(), self._stream()))
        return ix, iy

    def stage_heatmaps(self, hm: torch.Tensor) -> None:
        B = self._check_hm(hm)
        self._keep = hm.to(self.device, torch.float32).contiguous()
        self._ck(self.lib.fvp_stage_heatmaps(self.ctx, self._keep.data_ptr(), B, self._stream()))

    def hdn_project(self, B: int, slots) -> torch.Tensor:
        out = self._new(B, self.J, self.X, self.Y)
        self._ck(self.lib.fvp_hdn_project(self.ctx, B, self._slots_arr(slots), out.data_ptr(), self._stream()))
        return out

    def center_net(self, plane: Optional[torch.Tensor], B: int):
        hm, size = self._new(B, self.X, self.Y), self._new(B, 2, self.X, self.Y)
        p = plane.to(self.device, torch.float32).contiguous() if plane is not None else None
        self._ck(self.lib.fvp_center_net(self.ctx, p.data_ptr() if p is not None else None, B, hm.data_ptr(),
                                         size.data_ptr(), self._stream()))
        return hm, size

    def nms_topk(self, hm: torch.Tensor):
        B = hm.shape[0]
        h = hm.to(self.device, torch.float32).contiguous()
        conf, flat = self._new(B, self.P), self._new(B, self.P, dtype=torch.int32)
        self._ck(self.lib.fvp_nms_topk(self.ctx, h.data_ptr(), B, conf.data_ptr(), flat.data_ptr(), self._stream()))
        return conf, flat

    def proposals(self, B: int, slots, conf2d: torch.Tensor, flat: torch.Tensor, size: torch.Tensor):
        c = conf2d.to(self.device, torch.float32).contiguous()
        f = flat.to(self.device, torch.int32).contiguous()
        s = size.to(self.device, torch.float32).contiguous()
        cols, hm1d = self._new(B * self.P, self.J, self.Z), self._new(B * self.P, self.Z)
        centers = self._new(B, self.P, 7)
        self._ck(self.lib.fvp_proposals(self.ctx, B, self._slots_arr(slots), c.data_ptr(), f.data_ptr(), s.data_ptr(),
                                        cols.data_ptr(), hm1d.data_ptr(), centers.data_ptr(), self._stream()))
        return cols, hm1d, centers

    def c2c_net(self, cols: torch.Tensor) -> torch.Tensor:
        x = cols.to(self.device, torch.float32).contiguous()
        out = self._new(x.shape[0], self.Z)
        self._ck(self.lib.fvp_c2c_net(self.ctx, x.data_ptr(), x.shape[0], out.data_ptr(), self._stream()))
        return out

    def jln_project(self, B: int, slots, centers: torch.Tensor):
        c = centers.to(self.device, torch.float32).contiguous()
        n = B * self.P
        planes, off = self._new(3, n, self.J, 64, 64), self._new(n, 3)
        self._ck(self.lib.fvp_jln_project(self.ctx, B, self._slots_arr(slots), c.data_ptr(), planes.data_ptr(),
                                          off.data_ptr(), self._stream()))
        return planes, off

    def p2p_net(self, planes: torch.Tensor, valid: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = planes.to(self.device, torch.float32).contiguous().view(-1, self.J, 64, 64)
        v = valid.to(self.device, torch.int32).contiguous() if valid is not None else None
        out = torch.zeros_like(x)
        self._ck(self.lib.fvp_p2p_net(self.ctx, x.data_ptr(), x.shape[0], v.data_ptr() if v is not None else None,
                                      out.data_ptr(), self._stream()))
        return out

    def pose_head(self, feat: torch.Tensor, offset: torch.Tensor):
        f = feat.to(self.device, torch.float32).contiguous()          # [3,n,J,64,64]
        n = f.shape[1]
        o = offset.to(self.device, torch.float32).contiguous()
        pose, conf = self._new(3, n, self.J, 2), self._new(n)
        w, fused = self._new(3, n, self.J), self._new(n, self.J, 3)
        self._ck(self.lib.fvp_pose_head(self.ctx, f.data_ptr(), o.data_ptr(), n, pose.data_ptr(), conf.data_ptr(),
                                        w.data_ptr(), fused.data_ptr(), self._stream()))
        return pose, conf, w, fused


class EngineLanes:
    """Several frames in flight on ONE GPU: a root context plus ``lanes - 1`` lane contexts (``fvp_create_lane``) that
    share its weights, calibrations and sample-grid caches - L workspaces, one weight set, one 164 MB fine grid.

    A batch-1 forward is a chain of ~55 kernels of which many cannot fill 148 SMs (CenterNet on one 80x80 plane is
    <= 50 CTAs, the proposal kernel is one CTA per slot, ...).  Consecutive frames are independent
    (project_whole.py:71-84, joint_localization_net.py:72 loop per frame), so frame i+1 may run its latency-bound
    stages under the throughput-bound stages of frame i.  Every lane owns its workspaces, CUDA graph and stream;
    frames are dispatched round-robin and complete in order per lane.  The arithmetic of a frame does not depend on
    the lane it ran on (tested: bit-identical to the single-lane result).
    """

    def __init__(self, cfg, device=None, lanes: int = 2, max_batch: int = 1, max_sequences: int = 8, axes=None):
        assert lanes >= 1
        root = Engine(cfg, device, max_batch=max_batch, max_sequences=max_sequences, axes=axes)
        self.engines = [root] + [root.new_lane(max_batch) for _ in range(lanes - 1)]
        e0 = root
        self.device, self.P, self.J, self.V = e0.device, e0.P, e0.J, e0.V
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self._next = 0
        self._pending: List[Tuple[int, torch.cuda.Event, Tuple[torch.Tensor, ...]]] = []
        self._host_pending: List[Tuple[int, int]] = []     # (lane, lane ticket) in submission order

    def __len__(self) -> int:
        return len(self.engines)

    def close(self) -> None:
        self.engines[0].close()                             # closes its lanes first

    # ---- weights and calibrations live in the root context, every lane reads them ------------------
    def load_state_dict(self, sd) -> None:
        torch.cuda.synchronize(self.device)                 # no forward of any lane may be in flight
        self.engines[0].load_state_dict(sd)

    def sequence_slot(self, cams, resize) -> int:
        return self.engines[0].sequence_slot(cams, resize)

    def use_cuda_graph(self, on: bool = True) -> None:
        for e in self.engines:
            e.use_cuda_graph(on)

    def set_conv_mode(self, mode: int) -> None:
        for e in self.engines:
            e.set_conv_mode(mode)

    def last_launch_count(self) -> int:
        return self.engines[0].last_launch_count()

    # ---- device-resident frames -------------------------------------------------------------------
    def submit(self, heatmaps: torch.Tensor, slots: Sequence[int], out_fused: Optional[torch.Tensor] = None) -> int:
        """Enqueue one forward on the next lane (ordered after the work already queued on the caller's current
        stream); returns a ticket for collect().  Nothing blocks the host.  ``out_fused``: see Engine.forward."""
        lane = self._next
        self._next = (lane + 1) % len(self.engines)
        cur = torch.cuda.current_stream(self.device)
        st = self.streams[lane]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            out = self.engines[lane].forward(heatmaps, slots, out_fused)
            ev = torch.cuda.Event()
            ev.record(st)
        heatmaps.record_stream(st)
        self._pending.append((lane, ev, out))
        return len(self._pending) - 1

    def collect(self):
        """Oldest outstanding submit(): makes the caller's current stream wait for it (stream-ordered, the host is
        not blocked) and returns (fused_poses, plane_poses, proposal_centers)."""
        lane, ev, out = self._pending.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in out:
            t.record_stream(cur)
        return out

    def outstanding(self) -> int:
        return len(self._pending)

    # ---- host-resident frames (fvp_submit_host / fvp_wait of each lane) ----------------------------
    def new_host_outputs(self, B: int):
        return self.engines[0].new_host_outputs(B)

    def submit_host(self, heatmaps: torch.Tensor, slots: Sequence[int], out) -> int:
        lane = self._next
        self._next = (lane + 1) % len(self.engines)
        t = self.engines[lane].submit_host(heatmaps, slots, out)
        self._host_pending.append((lane, t))
        return len(self._host_pending) - 1

    def wait_oldest(self) -> None:
        lane, t = self._host_pending.pop(0)
        self.engines[lane].wait(t)

    def host_outstanding(self) -> int:
        return len(self._host_pending)

    def stream_host(self, frames, slots_of=None, depth: Optional[int] = None):
        """Run an iterable of pinned host heat-map batches through all lanes, `depth` frames in flight (default one
        per lane plus one, at most two per lane); yields (index, outputs) in submission order."""
        depth = min(2 * len(self.engines), depth or len(self.engines) + 1)
        bufs, pending = {}, []
        for i, hm in enumerate(frames):
            B = hm.shape[0]
            key = (B, i % (depth + 1))
            if key not in bufs:
                bufs[key] = self.new_host_outputs(B)
            self.submit_host(hm, slots_of(i) if slots_of else [0] * B, bufs[key])
            pending.append((i, bufs[key]))
            if len(pending) == depth:
                j, o = pending.pop(0)
                self.wait_oldest()
                yield j, o
        for j, o in pending:
            self.wait_oldest()
            yield j, o
