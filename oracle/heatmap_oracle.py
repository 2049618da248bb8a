"""CPU oracle of the heat-map renderer - TEST INFRASTRUCTURE ONLY.

Restates what the reference does between a list of 2-D poses and ``input_heatmaps`` when
``TEST_HEATMAP_SRC`` is 'pred' or 'gt' (the step immediately before the hot path):

* ``render_input_heatmap``  <- ``JointsDataset.generate_input_heatmap``  lib/dataset/JointsDataset.py:271-337
                               (evaluation branch: ``data_augmentation`` off) and ``compute_human_scale`` :197-203
* ``pred_heatmaps``         <- the 'pred' branch of ``__getitem__``       lib/dataset/JointsDataset.py:144-154
                               (``affine_transform`` lib/utils/transforms.py:53-56 on every joint, one map per view)
* ``gt_heatmaps``           <- the 'gt' branch of ``__getitem__``         lib/dataset/JointsDataset.py:156-190
                               (``project_pose_cpu`` lib/utils/cameras.py:59-93, visibility tests, affine, render)

Arithmetic follows the reference under the NumPy of this image (2.x, NEP 50 promotion): person scale, sigma, patch
bounds and the Gaussian argument are float64, ``exp`` is float64, the result is rounded to float32 when it is written
into the float32 map (``np.maximum`` + assignment).  Under the reference's pinned NumPy 1.x the Gaussian would be
evaluated in float32 (value-based casting); the parity pin below is "the reference run in THIS container".

Pin: ``oracle/gen_golden_heatmaps.py`` imports the unmodified ``JointsDataset`` from /root/reference and asserts this
restatement bit-identical on every generated case before writing ``tests/golden/heatmaps_*.npz``.

Nothing under faster-voxelpose_b200/ imports this module.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def affine_points(pts: np.ndarray, t: np.ndarray) -> np.ndarray:
    """[...,2] float64 -> t[:, :2] @ p + t[:, 2] with the reference's evaluation order
    (np.dot of a 2x3 matrix with [x, y, 1.]: (t0*x + t1*y) + t2*1, transforms.py:53-56)."""
    t = np.asarray(t, np.float64)
    x, y = np.asarray(pts[..., 0], np.float64), np.asarray(pts[..., 1], np.float64)
    return np.stack([t[0, 0] * x + t[0, 1] * y + t[0, 2] * 1.0, t[1, 0] * x + t[1, 1] * y + t[1, 2] * 1.0], axis=-1)


def human_scale(pose_hm: np.ndarray) -> float:
    """compute_human_scale with all joints 'visible' (JointsDataset.py:197-203, called with np.ones at :278-279)."""
    minx, maxx = np.min(pose_hm[:, 0]), np.max(pose_hm[:, 0])
    miny, maxy = np.min(pose_hm[:, 1]), np.max(pose_hm[:, 1])
    return np.clip(np.maximum(maxy - miny, maxx - minx) ** 2, 1.0 / 4 * 96 ** 2, 4 * 96 ** 2)


def render_input_heatmap(joints: Sequence[np.ndarray], joints_vis: Optional[Sequence[np.ndarray]], heatmap_size,
                         image_size, sigma) -> np.ndarray:
    """[J,H,W] float32 from a list of [J,>=2] poses in IMAGE_SIZE pixels (JointsDataset.py:271-337)."""
    heatmap_size = np.asarray(heatmap_size)
    image_size = np.asarray(image_size)
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    J = joints[0].shape[0]
    target = np.zeros((J, H, W), np.float32)
    stride = image_size / heatmap_size                                   # float64 [2]
    xs = np.arange(W)
    ys = np.arange(H)
    for n in range(len(joints)):
        hs = 2 * human_scale(joints[n][:, :2] / stride)
        if hs == 0:
            continue
        cur_sigma = sigma * np.sqrt(hs / (96.0 * 96.0))
        tmp = cur_sigma * 3
        size = 2 * tmp + 1
        c0 = size // 2                                                   # float64: centre index inside the patch
        for j in range(J):
            if joints_vis is not None and joints_vis[n][j] == 0:
                continue
            mu_x = int(joints[n][j][0] / stride[0])                      # int(): truncation toward zero
            mu_y = int(joints[n][j][1] / stride[1])
            ul = (int(mu_x - tmp), int(mu_y - tmp))
            br = (int(mu_x + tmp + 1), int(mu_y + tmp + 1))
            if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
                continue
            x0, x1 = max(0, ul[0]), min(br[0], W)
            y0, y1 = max(0, ul[1]), min(br[1], H)
            if x0 >= x1 or y0 >= y1:
                continue
            # patch sample index of an image pixel: px - ul; the patch itself is np.arange(0, size) as float32
            gx = (xs[x0:x1] - ul[0]).astype(np.float32)
            gy = (ys[y0:y1] - ul[1]).astype(np.float32)[:, None]
            g = np.exp(-((gx - c0) ** 2 + (gy - c0) ** 2) / (2 * cur_sigma ** 2))    # float64
            target[j, y0:y1, x0:x1] = np.maximum(target[j, y0:y1, x0:x1], g)
        target = np.clip(target, 0, 1)
    return target


def pred_heatmaps(all_preds: Sequence[Sequence[np.ndarray]], resize_transform: np.ndarray, heatmap_size, image_size,
                  sigma) -> np.ndarray:
    """'pred' source: all_preds[view][person] = [J,>=2] poses in ORIGINAL image pixels -> [V,J,H,W]
    (JointsDataset.py:144-154).  A view without detections cannot be rendered by the reference (IndexError on
    ``joints[0]``); here it gives an all-zero map (J from the first non-empty view)."""
    J = next(p[0].shape[0] for p in all_preds if len(p))
    out = []
    for preds in all_preds:
        if not len(preds):
            out.append(np.zeros((J, int(heatmap_size[1]), int(heatmap_size[0])), np.float32))
            continue
        moved = []
        for p in preds:
            q = np.array(p, np.float64, copy=True)
            q[:, :2] = affine_points(q[:, :2], resize_transform)
            moved.append(q)
        out.append(render_input_heatmap(moved, None, heatmap_size, image_size, sigma))
    return np.stack(out)


def project_pose_f64(x: np.ndarray, cam: dict) -> np.ndarray:
    """project_pose_cpu / project_point_cpu (cameras.py:20-27,58-93): the float64 numpy twin of the torch chain."""
    R = np.asarray(cam["R"], np.float64)
    T = np.asarray(cam["T"], np.float64).reshape(3, 1)
    f = np.array([[cam["fx"]], [cam["fy"]]], np.float64)
    c = np.array([[cam["cx"]], [cam["cy"]]], np.float64)
    k = np.asarray(cam["k"], np.float64).reshape(3, 1)
    p = np.asarray(cam["p"], np.float64).reshape(2, 1)
    xcam = np.matmul(R, x.T - T)
    y = xcam[:2] / (xcam[2] + 1e-5)
    r = np.sum(y ** 2, axis=0)
    d = 1 + k[0] * r + k[1] * r * r + k[2] * r * r * r
    u = y[0, :] * d + 2 * p[0] * y[0, :] * y[1, :] + p[1] * (r + 2 * y[0, :] * y[0, :])
    v = y[1, :] * d + 2 * p[1] * y[0, :] * y[1, :] + p[0] * (r + 2 * y[1, :] * y[1, :])
    return (f * np.stack([u, v]) + c).T


def gt_heatmaps(joints_3d: Sequence[np.ndarray], joints_3d_vis: Sequence[np.ndarray], cams: Sequence[dict],
                resize_transform: np.ndarray, ori_image_size, image_size, heatmap_size, sigma) -> np.ndarray:
    """'gt' source: 3-D joints [J,3] + visibility [J] per person -> [V,J,H,W] (JointsDataset.py:156-190)."""
    out = []
    for cam in cams:
        j2, jv = [], []
        for n in range(len(joints_3d)):
            pose = project_pose_f64(np.asarray(joints_3d[n], np.float64), cam)
            ok = (pose[:, 0] >= 0) & (pose[:, 0] <= ori_image_size[0] - 1) & (pose[:, 1] >= 0) & (pose[:, 1] <= ori_image_size[1] - 1)
            vis = np.asarray(joints_3d_vis[n]) > 0
            vis = vis & ok
            pose = affine_points(pose, resize_transform)
            bad = (np.min(pose, axis=1) < 0) | (pose[:, 0] >= image_size[0]) | (pose[:, 1] >= image_size[1])
            vis = vis & ~bad
            j2.append(pose)
            jv.append(vis)
        out.append(render_input_heatmap(j2, jv, heatmap_size, image_size, sigma))
    return np.stack(out)
