"""Readers for the data formats on either side of the hot path: the calibration files and the
2-D detection files the reference's dataset classes load, without the dataset classes themselves (they need the image
folders, ``json_tricks`` and ``actorsGT.mat`` even when only heat maps from 2-D poses are wanted).

* Campus / Shelf calibration  ``calibration_{campus,shelf}.json``  - ``{"0": {R, T, fx, fy, cx, cy, k, p}, ...}``
  (``lib/dataset/campus.py:114-129``, ``shelf.py`` likewise) and the demo's ``{"<sequence>": [cam, ...]}`` form
  (``demo/calibration.json``, read by the notebook) -> lists of camera dicts with NumPy values, the ``cameras[seq]``
  the model's ``forward`` receives;
* Panoptic ``calibration_<seq>.json`` (``{"cameras": [{panel, node, K, distCoef, R, t}, ...]}``) -> the five HD cameras in
  the reference's convention ``x_cam = R (x - T)``, millimetres, y-up to z-up (``lib/dataset/panoptic.py:171-205``);
* Panoptic ``hdPose3d_stage1_coco19/body3DScene_*.json`` -> the ground-truth poses and visibilities of a frame as
  ``Panoptic._get_db`` stores them (``panoptic.py:109-163``), i.e. what ``fvp.evaluate.evaluate_panoptic`` and
  ``HeatmapRenderer.from_gt`` take;
* ``pred_{campus,shelf}_maskrcnn_hrnet_coco.pkl``: ``{"<view>_<frame>": [{"pred": [17][x, y, score]}, ...]}`` -> the
  per-view, per-person arrays of one frame, i.e. ``db_rec['pred_pose2d']`` (``campus.py:92-97``), which
  ``fvp.render.HeatmapRenderer.from_pred`` turns into ``input_heatmaps`` on the GPU.

Host logic only (NumPy); pinned against the reference's own loaders by ``oracle/gen_golden_datasets.py`` /
``tests/test_datasets.py``.
"""
from __future__ import annotations

import json
import pickle
from typing import Dict, Iterable, Iterator, List, Sequence, Tuple, Union

import numpy as np

CAMPUS_FRAMES: List[int] = list(range(350, 471)) + list(range(650, 751))      # campus.py:55
SHELF_FRAMES: List[int] = list(range(300, 601))                               # shelf.py:81
PANOPTIC_CAM_LIST: List[Tuple[int, int]] = [(0, 3), (0, 6), (0, 12), (0, 13), (0, 23)]   # panoptic.py:79
PANOPTIC_VAL_LIST: List[str] = ["160906_pizza1", "160422_haggling1", "160906_ian5", "160906_band4"]   # panoptic.py:35-40

_CAM_KEYS = ("R", "T", "fx", "fy", "cx", "cy", "k", "p")


def _as_camera(cam: dict) -> dict:
    """Every value as a NumPy array, exactly what ``np.array(v)`` makes of the JSON value (campus.py:119-121)."""
    missing = [k for k in _CAM_KEYS if k not in cam]
    if missing:
        raise KeyError("camera entry lacks %s" % missing)
    return {k: np.array(v) for k, v in cam.items()}


def load_calibration(path: str) -> Union[List[dict], Dict[str, List[dict]]]:
    """Campus / Shelf file (string keys '0'..'V-1') -> list of cameras in view order; demo-style file (sequence name ->
    list of cameras) -> ``{sequence: [camera, ...]}``."""
    with open(path) as f:
        doc = json.load(f)
    if not isinstance(doc, dict) or not doc:
        raise ValueError("%s: not a calibration file" % path)
    if all(isinstance(v, dict) for v in doc.values()):
        try:
            order = sorted(doc, key=int)
        except ValueError:
            raise ValueError("%s: camera keys must be integers" % path)
        if [int(k) for k in order] != list(range(len(order))):
            raise ValueError("%s: camera keys must be 0..V-1, got %s" % (path, order))
        return [_as_camera(doc[k]) for k in order]
    if all(isinstance(v, list) for v in doc.values()):
        return {seq: [_as_camera(c) for c in cams] for seq, cams in doc.items()}
    raise ValueError("%s: neither {view: camera} nor {sequence: [camera, ...]}" % path)


def panoptic_cameras(calib: Union[str, dict], num_views: int = 5) -> List[dict]:
    """The reference's camera selection and conversion for one Panoptic sequence (panoptic.py:171-205): HD cameras
    ``PANOPTIC_CAM_LIST[:num_views]`` in file order, ``R' = R M`` (y-up -> z-up), ``T = -R'^T t * 10`` (cm -> mm, camera
    centre instead of translation), intrinsics from ``K``, ``k = distCoef[[0, 1, 4]]``, ``p = distCoef[[2, 3]]``."""
    if isinstance(calib, str):
        with open(calib) as f:
            calib = json.load(f)
    M = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
    want = PANOPTIC_CAM_LIST[:num_views]
    out = []
    for cam in calib["cameras"]:
        if (cam["panel"], cam["node"]) in want:
            K, dist = np.array(cam["K"]), np.array(cam["distCoef"])
            R = np.array(cam["R"]).dot(M)
            t = np.array(cam["t"]).reshape((3, 1))
            out.append({"R": R, "T": -np.dot(R.T, t) * 10.0, "fx": K[0, 0], "fy": K[1, 1], "cx": K[0, 2], "cy": K[1, 2],
                        "k": dist[[0, 1, 4]].reshape(3, 1), "p": dist[[2, 3]].reshape(2, 1)})
    return out


def panoptic_annotation_files(seq_dir: str, interval: int = 12) -> List[str]:
    """The ``hdPose3d_stage1_coco19/*.json`` files of a sequence the reference visits (panoptic.py:111-116): sorted, every
    ``interval``-th (12 for validation, 3 for training, panoptic.py:81-88)."""
    import glob
    import os
    files = sorted(glob.iglob("{:s}/*.json".format(os.path.join(seq_dir, "hdPose3d_stage1_coco19"))))
    return [f for i, f in enumerate(files) if i % interval == 0]


def panoptic_image_paths(dataset_dir: str, seq: str, anno_file: str, num_views: int = 5) -> List[str]:
    """The HD images that belong to an annotation file (panoptic.py:122-133), existing or not."""
    import os
    out = []
    for k in range(num_views):
        suffix = os.path.basename(anno_file).replace("body3DScene", "")
        prefix = "{:02d}_{:02d}".format(PANOPTIC_CAM_LIST[k][0], PANOPTIC_CAM_LIST[k][1])
        out.append(os.path.join(dataset_dir, seq, "hdImgs", prefix, prefix + suffix).replace("json", "jpg"))
    return out


def panoptic_frame_gt(anno: Union[str, dict], num_joints: int = 15, root_id: int = 2) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Ground truth of one ``body3DScene_*.json`` as the reference's ``_get_db`` stores it (panoptic.py:138-157):
    per body the first ``num_joints`` of ``joints19`` -> ``[J,3]`` world millimetres (y-up centimetres times ``M``, times
    10) and ``[J]`` visibilities clipped at 0; bodies whose root joint has visibility <= 0.1 are dropped."""
    if isinstance(anno, str):
        with open(anno, "r") as f:
            anno = json.load(f)
    M = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
    joints, vis = [], []
    for body in anno["bodies"]:
        pose3d = np.array(body["joints19"]).reshape((-1, 4))
        pose3d = pose3d[:num_joints]
        joints_vis = np.maximum(pose3d[:, -1], 0.0)
        if joints_vis[root_id] <= 0.1:
            continue
        pose3d[:, 0:3] = pose3d[:, 0:3].dot(M)
        joints.append(pose3d[:, 0:3] * 10.0)
        vis.append(joints_vis)
    return joints, vis


def panoptic_records(dataset_dir: str, sequences: Sequence[str] = tuple(PANOPTIC_VAL_LIST), interval: int = 12,
                     num_views: int = 5, num_joints: int = 15, root_id: int = 2) -> List[dict]:
    """The database ``Panoptic._get_db`` builds (panoptic.py:109-163), one record per visited annotation file that has at
    least one kept body and all its HD images on disk: ``{'seq', 'all_image_path', 'joints_3d', 'joints_3d_vis'}``."""
    import os
    db = []
    for seq in sequences:
        for anno_file in panoptic_annotation_files(os.path.join(dataset_dir, seq), interval):
            with open(anno_file, "r") as f:
                anno = json.load(f)
            if len(anno["bodies"]) == 0:
                continue
            paths = panoptic_image_paths(dataset_dir, seq, anno_file, num_views)
            if not all(os.path.exists(p) for p in paths):
                continue
            joints, vis = panoptic_frame_gt(anno, num_joints, root_id)
            if len(joints) > 0:
                db.append({"seq": seq, "all_image_path": paths, "joints_3d": joints, "joints_3d_vis": vis})
    return db


IMAGENET_MEAN = (0.485, 0.456, 0.406)      # run/validate.py:45-46
IMAGENET_STD = (0.229, 0.224, 0.225)


def load_views(paths: Sequence[str], color_rgb: bool = True) -> np.ndarray:
    """The ``views`` of one frame as the reference's loader produces them (JointsDataset.py:124-134 with the transform
    of run/validate.py:44-52): ``cv2.imread`` (colour, EXIF orientation ignored) -> BGR to RGB when ``COLOR_RGB`` ->
    ``ToTensor`` (uint8 HWC -> float32 CHW / 255) -> ``Normalize(mean, std)``; float32 arithmetic like torchvision's.
    Returns ``[V, 3, h, w]`` float32.  The reference does not resize here: Panoptic frames are resized once by its
    ``preprocess.py``."""
    import cv2
    mean = np.array(IMAGENET_MEAN, np.float32).reshape(3, 1, 1)
    std = np.array(IMAGENET_STD, np.float32).reshape(3, 1, 1)
    out = []
    for p in paths:
        img = cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        if img is None:
            raise FileNotFoundError(p)
        if color_rgb:
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        x = np.ascontiguousarray(img.transpose(2, 0, 1)).astype(np.float32) / np.float32(255)
        out.append((x - mean) / std)
    return np.stack(out)


def load_pred_pose2d(path: str) -> dict:
    """The detection file of Campus / Shelf (campus.py:61-67): ``{"<view>_<frame>": [{"pred": ...}, ...]}``.
    A pickle: only open files you trust (the reference does the same)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def frame_preds(pred2d: dict, frame: int, num_views: int) -> List[List[np.ndarray]]:
    """``db_rec['pred_pose2d']`` of one frame (campus.py:92-97): per view the detected people as ``[J, 3]`` arrays
    (x, y in original-image pixels, score).  A view without an entry behaves like the container does, as in the
    reference: the shipped files unpickle to ``defaultdict(list)`` and answer with no people, a plain dict raises."""
    return [[np.array(p["pred"]) for p in pred2d["%d_%d" % (k, frame)]] for k in range(num_views)]


def pred_frames(pred2d: dict, frames: Iterable[int], num_views: int) -> Iterator[Tuple[int, List[List[np.ndarray]]]]:
    """(frame, per-view detections) over a frame range, e.g. ``CAMPUS_FRAMES`` - the order the reference's ``_get_db``
    builds its records in."""
    for i in frames:
        yield i, frame_preds(pred2d, i, num_views)


def batch_preds(pred2d: dict, frames: Sequence[int], num_views: int) -> List[List[List[np.ndarray]]]:
    """``[frame][view][person]`` for ``HeatmapRenderer.from_pred``."""
    return [frame_preds(pred2d, i, num_views) for i in frames]
