"""Host side of the heat-map renderer.

Mirrors the two branches of the reference's ``JointsDataset.__getitem__`` that build ``input_heatmaps`` from 2-D / 3-D
poses (lib/dataset/JointsDataset.py:144-190): the per-joint bookkeeping (resize affine, camera projection, visibility
tests - a few hundred float64 points per frame) stays on the host in NumPy exactly as the reference evaluates it; the
rendering itself (``generate_input_heatmap``, :271-337 - the slow per-joint NumPy loop of the data-loader workers) runs
on the GPU through ``fvp_render_heatmaps`` and leaves the maps in device memory, in the layout the hot path consumes.
No CPU fallback: without the library / a CUDA device this module raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Sequence

import numpy as np
import torch

from .engine import Engine

MAX_PEOPLE = 16          # FVP_MAX_PEOPLE of the library


def affine_points(pts: np.ndarray, t: np.ndarray) -> np.ndarray:
    """``affine_transform`` (lib/utils/transforms.py:53-56) on [...,2] float64 points: t @ [x, y, 1]."""
    t = np.asarray(t, np.float64)
    x, y = np.asarray(pts[..., 0], np.float64), np.asarray(pts[..., 1], np.float64)
    return np.stack([t[0, 0] * x + t[0, 1] * y + t[0, 2] * 1.0, t[1, 0] * x + t[1, 1] * y + t[1, 2] * 1.0], axis=-1)


def project_pose_cpu(x: np.ndarray, cam: Mapping) -> np.ndarray:
    """``project_pose_cpu`` (lib/utils/cameras.py:20-27,58-93): [n,3] world mm -> [n,2] original-image pixels, float64."""
    R = np.asarray(cam["R"], np.float64)
    T = np.asarray(cam["T"], np.float64).reshape(3, 1)
    f = np.array([[cam["fx"]], [cam["fy"]]], np.float64)
    c = np.array([[cam["cx"]], [cam["cy"]]], np.float64)
    k = np.asarray(cam["k"], np.float64).reshape(3, 1)
    p = np.asarray(cam["p"], np.float64).reshape(2, 1)
    xcam = np.matmul(R, np.asarray(x, np.float64).T - T)
    y = xcam[:2] / (xcam[2] + 1e-5)
    r = np.sum(y ** 2, axis=0)
    d = 1 + k[0] * r + k[1] * r * r + k[2] * r * r * r
    u = y[0, :] * d + 2 * p[0] * y[0, :] * y[1, :] + p[1] * (r + 2 * y[0, :] * y[0, :])
    v = y[1, :] * d + 2 * p[1] * y[0, :] * y[1, :] + p[0] * (r + 2 * y[1, :] * y[1, :])
    return (f * np.stack([u, v]) + c).T


class HeatmapRenderer:
    """2-D / 3-D poses -> ``input_heatmaps`` [B,V,J,H,W] on the engine's GPU."""

    def __init__(self, engine: Engine, sigma: Optional[float] = None):
        self.eng = engine
        self.sigma = float(sigma if sigma is not None else engine.cfg.NETWORK.SIGMA)
        ds = engine.cfg.DATASET
        self.ori_size = [float(v) for v in ds.ORI_IMAGE_SIZE]
        self.image_size = [float(v) for v in ds.IMAGE_SIZE]

    # ---- the device call -----------------------------------------------------------------------------------------
    def render(self, joints_img: np.ndarray, num_people: np.ndarray, vis: Optional[np.ndarray] = None) -> torch.Tensor:
        """joints_img [B,V,N,J,2] float64 in IMAGE_SIZE pixels, num_people [B,V], vis [B,V,N,J] (0/1) or None."""
        e = self.eng
        j = np.ascontiguousarray(joints_img, np.float64)
        if j.ndim != 5 or j.shape[1] != e.V or j.shape[3] != e.J or j.shape[4] != 2:
            raise ValueError("joints must be [B,%d,N,%d,2], got %s" % (e.V, e.J, j.shape))
        B, N = j.shape[0], j.shape[2]
        if B > e.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (B, e.max_batch))
        if not 1 <= N <= MAX_PEOPLE:
            raise ValueError("people per view must be in [1,%d], got %d" % (MAX_PEOPLE, N))
        n = np.ascontiguousarray(num_people, np.int32).reshape(B, e.V)
        v = None
        if vis is not None:
            v = np.ascontiguousarray(np.asarray(vis) != 0, np.uint8).reshape(B, e.V, N, e.J)
        out = torch.empty((B, e.V, e.J, e.H, e.W), device=e.device, dtype=torch.float32)
        with torch.cuda.device(e.device):
            e._ck(e.lib.fvp_render_heatmaps(e.ctx, j.ctypes.data, n.ctypes.data, v.ctypes.data if v is not None else None,
                                            B, N, C.c_double(self.sigma), out.data_ptr(), e._stream()))
        return out

    # ---- TEST_HEATMAP_SRC == 'pred' (JointsDataset.py:144-154) -----------------------------------------------------
    def from_pred(self, batch_preds: Sequence[Sequence[Sequence[np.ndarray]]], resize_transform) -> torch.Tensor:
        """batch_preds[frame][view][person] = [J,>=2] poses in ORIGINAL image pixels (db_rec['pred_pose2d'])."""
        e = self.eng
        B = len(batch_preds)
        N = max(1, max(len(view) for frame in batch_preds for view in frame))
        joints = np.zeros((B, e.V, N, e.J, 2), np.float64)
        num = np.zeros((B, e.V), np.int32)
        for b, frame in enumerate(batch_preds):
            assert len(frame) == e.V, "one pose list per camera view"
            for v, people in enumerate(frame):
                num[b, v] = len(people)
                for n, pose in enumerate(people):
                    joints[b, v, n] = affine_points(np.asarray(pose, np.float64)[:, :2], resize_transform)
        return self.render(joints, num)

    # ---- TEST_HEATMAP_SRC == 'gt' (JointsDataset.py:156-190) -------------------------------------------------------
    def from_gt(self, batch_joints_3d: Sequence[Sequence[np.ndarray]], batch_vis: Sequence[Sequence[np.ndarray]],
                batch_cameras: Sequence[Sequence[Mapping]], resize_transform) -> torch.Tensor:
        """batch_joints_3d[frame][person] = [J,3] world mm, batch_vis[frame][person] = [J], batch_cameras[frame] = V cameras."""
        e = self.eng
        B = len(batch_joints_3d)
        N = max(1, max(len(f) for f in batch_joints_3d))
        joints = np.zeros((B, e.V, N, e.J, 2), np.float64)
        vis = np.zeros((B, e.V, N, e.J), np.uint8)
        num = np.zeros((B, e.V), np.int32)
        ow, oh = self.ori_size
        iw, ih = self.image_size
        for b in range(B):
            cams = batch_cameras[b]
            assert len(cams) == e.V, "inconsistent number of cameras"
            for v, cam in enumerate(cams):
                num[b, v] = len(batch_joints_3d[b])
                for n, (j3, jv) in enumerate(zip(batch_joints_3d[b], batch_vis[b])):
                    pose = project_pose_cpu(np.asarray(j3, np.float64), cam)
                    ok = (pose[:, 0] >= 0) & (pose[:, 0] <= ow - 1) & (pose[:, 1] >= 0) & (pose[:, 1] <= oh - 1)
                    pose = affine_points(pose, resize_transform)
                    bad = (np.min(pose, axis=1) < 0) | (pose[:, 0] >= iw) | (pose[:, 1] >= ih)
                    joints[b, v, n] = pose
                    vis[b, v, n] = (np.asarray(jv) > 0) & ok & ~bad
        return self.render(joints, num, vis)
