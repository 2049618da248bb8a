"""The caller of the hot path, ``run/validate.py:92-118``, without the reference's dataset classes and DataLoader.

``validate_pred``      Campus / Shelf, ``TEST_HEATMAP_SRC = 'pred'`` (``configs/{campus,shelf}/jln64.yaml``):
    detections of a frame block (fvp.datasets)  ->  heat maps on the GPU (fvp.render, N1)  ->  model(...) (the hot path)
    ->  torch.cat(all_fused_poses)  ->  PCP (fvp.evaluate, N3)
``validate_panoptic``  Panoptic, ``TEST_HEATMAP_SRC = 'image'`` (``configs/panoptic/jln64.yaml``: images -> backbone (N2)
    inside ``model(views=...)``) or ``'gt'`` (3-D ground truth -> heat maps on the GPU):  ->  AP / recall / MPJPE

The loop is the reference's: batches of ``cfg.TEST.BATCH_SIZE`` consecutive frames in ``frame_range`` order, no shuffling,
the poses of every call appended and concatenated once at the end (run/validate.py:95-114).  Nothing here computes: the
renderer, the model and the metrics are the pieces named above; host logic only.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import datasets, evaluate, synth


def validate_pred(cfg, model: Callable, renderer, cameras: Sequence[dict], pred2d: dict, frames: Sequence[int],
                  seq: str, actors=None, batch_size: Optional[int] = None, progress: Optional[Callable] = None) -> dict:
    """Run ``model`` over ``frames`` of a detection file and evaluate.

    model     ``models.faster_voxelpose.get(cfg)`` in eval mode on a CUDA device (any callable with the reference's
              ``forward`` signature); renderer  ``fvp.render.HeatmapRenderer`` (or anything with ``from_pred``)
    cameras   the sequence's cameras (``fvp.datasets.load_calibration``); pred2d  ``fvp.datasets.load_pred_pose2d``
    frames    e.g. ``fvp.datasets.CAMPUS_FRAMES``; seq  ``'campus'`` / ``'shelf'`` (``meta['seq']``, and the PCP head model)
    actors    ``fvp.evaluate.load_actors('actorsGT.mat')`` or None (no metric, poses only)
    Returns ``{'fused_poses': [len(frames), P, J, 5] tensor, 'metric', 'msg', 'detail'}`` (the last three None without actors).
    """
    V = int(cfg.DATASET.CAMERA_NUM)
    if len(cameras) != V:
        raise AssertionError("inconsistent number of cameras")          # project_whole.py:74
    B = int(batch_size if batch_size is not None else cfg.TEST.BATCH_SIZE)
    if B < 1:
        raise ValueError("batch size must be >= 1")
    resize = synth.resize_transform(cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE)     # JointsDataset.py:51-56
    resize_t = None
    all_fused = []
    with torch.no_grad():
        for lo in range(0, len(frames), B):
            block = list(frames[lo:lo + B])
            hm = renderer.from_pred(datasets.batch_preds(pred2d, block, V), resize)
            if resize_t is None:
                resize_t = torch.as_tensor(resize, dtype=torch.float, device=hm.device)
            fused, _, _, _, _ = model(backbone=None, meta={"seq": [seq] * len(block)}, input_heatmaps=hm,
                                      cameras={seq: cameras}, resize_transform=resize_t)
            all_fused.append(fused)
            if progress is not None:
                progress(lo + len(block), len(frames))
        fused_all = torch.cat(all_fused, dim=0) if all_fused else torch.zeros((0,))
    out = {"fused_poses": fused_all, "metric": None, "msg": None, "detail": None}
    if actors is not None:
        preds = fused_all.detach().cpu().numpy()
        out["metric"], out["msg"], out["detail"] = evaluate.evaluate_pcp(preds, actors, list(frames), seq)
    return out


def validate_panoptic(cfg, model: Callable, cameras: dict, records: Sequence[dict], source: Optional[str] = None, renderer=None,
                      backbone=None, batch_size: Optional[int] = None, load_views: Callable = datasets.load_views,
                      progress: Optional[Callable] = None) -> dict:
    """Run ``model`` over Panoptic ``records`` (``fvp.datasets.panoptic_records``) and evaluate.

    cameras   ``{sequence: fvp.datasets.panoptic_cameras(calibration_<sequence>.json)}``; batches may mix sequences
    source    ``'image'`` (``model(backbone=backbone, views=...)``, run/validate.py:97-101; backbone =
              ``models.resnet.get(cfg)``) or ``'gt'`` (``renderer.from_gt`` -> ``input_heatmaps``); default
              ``cfg.DATASET.TEST_HEATMAP_SRC``
    Returns ``{'fused_poses', 'metric', 'msg', 'detail'}`` with the metric of ``Panoptic.evaluate`` (panoptic.py:214-266).
    """
    source = source or str(cfg.DATASET.TEST_HEATMAP_SRC)
    if source not in ("image", "gt"):
        raise ValueError("source must be 'image' or 'gt' (use validate_pred for 'pred'), got %r" % source)
    if source == "image" and backbone is None:
        raise ValueError("the 'image' source needs the backbone (models.resnet.get(cfg))")
    if source == "gt" and renderer is None:
        raise ValueError("the 'gt' source needs a HeatmapRenderer")
    B = int(batch_size if batch_size is not None else cfg.TEST.BATCH_SIZE)
    if B < 1:
        raise ValueError("batch size must be >= 1")
    resize = synth.resize_transform(cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE)
    device = torch.device(cfg.DEVICE)
    resize_t = torch.as_tensor(resize, dtype=torch.float, device=device)
    color_rgb = bool(cfg.DATASET.COLOR_RGB)
    all_fused = []
    with torch.no_grad():
        for lo in range(0, len(records), B):
            block = records[lo:lo + B]
            meta = {"seq": [r["seq"] for r in block]}
            if source == "image":
                views = torch.from_numpy(np.stack([load_views(r["all_image_path"], color_rgb) for r in block])).to(device)
                fused, _, _, _, _ = model(backbone=backbone, views=views, meta=meta, cameras=cameras, resize_transform=resize_t)
            else:
                hm = renderer.from_gt([r["joints_3d"] for r in block], [r["joints_3d_vis"] for r in block],
                                      [cameras[r["seq"]] for r in block], resize)
                fused, _, _, _, _ = model(backbone=None, meta=meta, input_heatmaps=hm, cameras=cameras,
                                          resize_transform=resize_t.to(hm.device))
            all_fused.append(fused)
            if progress is not None:
                progress(lo + len(block), len(records))
        fused_all = torch.cat(all_fused, dim=0) if all_fused else torch.zeros((0,))
    preds = fused_all.detach().cpu().numpy()
    metric, msg, detail = evaluate.evaluate_panoptic(preds, [np.asarray(r["joints_3d"]) for r in records],
                                                     [np.asarray(r["joints_3d_vis"]) for r in records])
    return {"fused_poses": fused_all, "metric": metric, "msg": msg, "detail": detail}
