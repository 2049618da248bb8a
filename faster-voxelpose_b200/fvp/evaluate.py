"""Task metrics on the hot path's output: the consumers of ``fused_poses``.

Host-side mirror of the reference's ``dataset.evaluate(preds)`` (called by run/validate.py:116-118 on the
``torch.cat`` of all ``fused_poses [B,P,J,5]``):

* Panoptic  - AP@25..150 mm, recall@500 mm, MPJPE@500 mm   (lib/dataset/panoptic.py:214-311)
* Campus    - PCP per actor / per bone group, recall@500   (lib/dataset/campus.py:138-230)
* Shelf     - same with the Shelf head model                (lib/dataset/shelf.py:162-263)

Same inputs, same return value ``(metric, msg)`` (plus a dict of the individual numbers), same arithmetic: every
float is produced by the same NumPy expression on the same dtypes as in the reference (predictions float32, ground
truth float64), so results are equal to the last bit - tests/test_evaluate.py pins them against goldens produced by the
unmodified reference (oracle/gen_golden_eval.py).  Ground truth is passed in explicitly (the reference reads it from its
dataset object / ``actorsGT.mat``); ``load_actors`` parses that file the way the reference does.

Quirks kept on purpose (they change the numbers):
  * Panoptic matches every predicted pose to its nearest ground-truth person of the frame (no one-to-one assignment at
    this stage; duplicates are resolved later, by score order, inside AP / MPJPE) and skips frames without people.
  * AP: true positive = MPJPE below the threshold AND that ground-truth id not claimed by a higher-scoring pose; the
    precision envelope is taken right-to-left; recall uses ``total_gt + 1e-5``.
  * PCP: every actor present in a frame is matched to the predicted pose with the smallest mean joint distance, even
    when that distance is huge; a frame without any valid prediction is skipped entirely (its actors count neither as
    misses nor as parts) in Campus - and would raise in Shelf (np.stack of an empty list), which we reproduce as a skip
    only for Campus.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np

AP_THRESHOLDS = np.arange(25, 155, 25)          # mm (panoptic.py:249)
LIMBS = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8], [9, 10], [10, 11], [12, 13]]   # campus.py:146-147 / shelf.py:170
BONE_GROUPS = OrderedDict([("Head", [8]), ("Torso", [9]), ("Upper arms", [5, 6]), ("Lower arms", [4, 7]),
                           ("Upper legs", [1, 2]), ("Lower legs", [0, 3])])                 # campus.py:196-198
COCO_TO_14 = np.array([16, 14, 12, 11, 13, 15, 10, 8, 6, 5, 7, 9])                         # campus.py:222 / shelf.py:246


def _np(x) -> np.ndarray:
    """torch tensor (any device) or array -> numpy, dtype preserved (the reference calls .detach().cpu().numpy())."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


# ---------------------------------------------------------------------------------------------------
# Panoptic
# ---------------------------------------------------------------------------------------------------
def match_poses(preds: Sequence, gt_joints: Sequence, gt_vis: Sequence) -> Tuple[List[dict], int]:
    """panoptic.py:219-247: one record {mpjpe, score, gt_id} per valid predicted pose; total number of GT people."""
    assert len(preds) == len(gt_joints) == len(gt_vis), "number mismatch"
    records: List[dict] = []
    total_gt = 0
    for pred, joints, vis in zip(preds, gt_joints, gt_vis):
        if len(joints) == 0:
            continue
        pred = _np(pred)
        pred = pred[pred[:, 0, 3] >= 0]
        for pose in pred:
            errs = []
            for gt, gv in zip(joints, vis):
                seen = np.asarray(gv) > 0.1
                errs.append(np.mean(np.sqrt(np.sum((pose[seen, 0:3] - np.asarray(gt)[seen]) ** 2, axis=-1))))
            best = int(np.argmin(errs))
            records.append({"mpjpe": float(np.min(errs)), "score": float(pose[0, 4]), "gt_id": int(total_gt + best)})
        total_gt += len(joints)
    return records, total_gt


def _by_score(records: List[dict]) -> List[dict]:
    return sorted(records, key=lambda r: r["score"], reverse=True)      # stable, like list.sort in the reference


def average_precision(records: List[dict], total_gt: int, threshold: float) -> Tuple[float, float]:
    """panoptic.py:268-292 (returns AP and the final recall)."""
    ordered = _by_score(records)
    n = len(ordered)
    hit = np.zeros(n)
    claimed = set()
    for i, r in enumerate(ordered):
        if r["mpjpe"] < threshold and r["gt_id"] not in claimed:
            hit[i] = 1
            claimed.add(r["gt_id"])
    tp = np.cumsum(hit)
    fp = np.cumsum(1 - hit)
    recall = tp / (total_gt + 1e-5)
    precision = tp / (tp + fp + 1e-5)
    for k in range(n - 2, -1, -1):
        precision[k] = max(precision[k], precision[k + 1])
    precision = np.concatenate(([0], precision, [0]))
    recall = np.concatenate(([0], recall, [1]))
    step = np.where(recall[1:] != recall[:-1])[0]
    return float(np.sum((recall[step + 1] - recall[step]) * precision[step + 1])), float(recall[-2])


def matched_mpjpe(records: List[dict], threshold: float = 500) -> float:
    """panoptic.py:294-305: mean error of the first (highest-score) match of every GT person below the threshold."""
    claimed, errs = set(), []
    for r in _by_score(records):
        if r["mpjpe"] < threshold and r["gt_id"] not in claimed:
            errs.append(r["mpjpe"])
            claimed.add(r["gt_id"])
    return float(np.mean(errs)) if errs else float("inf")


def matched_recall(records: List[dict], total_gt: int, threshold: float = 500) -> float:
    """panoptic.py:307-311."""
    return len(np.unique([r["gt_id"] for r in records if r["mpjpe"] < threshold])) / total_gt


def evaluate_panoptic(preds: Sequence, gt_joints: Sequence, gt_vis: Sequence):
    """``Panoptic.evaluate`` (panoptic.py:214-266).  preds[i]: [P,J,5] (x,y,z mm, flag, score); gt_joints[i]: [n_i,J,3]
    mm; gt_vis[i]: [n_i,J].  Returns (mean AP, message, details)."""
    records, total_gt = match_poses(preds, gt_joints, gt_vis)
    aps, recs = [], []
    for t in AP_THRESHOLDS:
        ap, rec = average_precision(records, total_gt, t)
        aps.append(ap)
        recs.append(rec)
    mpjpe = matched_mpjpe(records)
    recall = matched_recall(records, total_gt)
    msg = ("Evaluation results on Panoptic dataset:\nap@25: {:.4f}\tap@50: {:.4f}\tap@75: {:.4f}\tap@100: {:.4f}\t"
           "ap@125: {:.4f}\tap@150: {:.4f}\trecall@500mm: {:.4f}\tmpjpe@500mm: {:.3f}").format(*aps, recall, mpjpe)
    metric = float(np.mean(aps))
    return metric, msg, {"aps": aps, "recalls": recs, "mpjpe": mpjpe, "recall": recall, "total_gt": total_gt,
                         "poses": len(records)}


# ---------------------------------------------------------------------------------------------------
# Campus / Shelf
# ---------------------------------------------------------------------------------------------------
def _head_points(coco_pose: np.ndarray):
    mid_shoulder = (coco_pose[5] + coco_pose[6]) / 2
    ear_centre = (coco_pose[3] + coco_pose[4]) / 2
    head_bottom = (mid_shoulder + ear_centre) / 2
    head_top = head_bottom + (ear_centre - head_bottom) * 2
    return head_bottom, head_top


def coco2campus3D(coco_pose: np.ndarray) -> np.ndarray:
    """[17,3] COCO-order pose -> [14,3] Campus order, head from shoulders and ears (campus.py:213-232)."""
    out = np.zeros((14, 3))
    out[0:12] += coco_pose[COCO_TO_14]
    head_bottom, head_top = _head_points(coco_pose)
    out[12] += head_bottom
    out[13] += head_top
    return out


def coco2shelf3D(coco_pose: np.ndarray) -> np.ndarray:
    """[17,3] COCO-order pose -> [14,3] Shelf order; head = 0.75 x (shoulder/nose model) + 0.25 x (ear model)
    (shelf.py:237-263)."""
    out = np.zeros((14, 3))
    out[0:12] += coco_pose[COCO_TO_14]
    head_bottom, head_top = _head_points(coco_pose)
    out[12] = (out[8] + out[9]) / 2
    out[13] = coco_pose[0]
    out[13] = out[12] + (out[13] - out[12]) * np.array([0.75, 0.75, 1.5])
    out[12] = out[12] + (coco_pose[0] - out[12]) * np.array([0.5, 0.5, 0.5])
    alpha = 0.75
    out[13] = out[13] * alpha + head_top * (1 - alpha)
    out[12] = out[12] * alpha + head_bottom * (1 - alpha)
    return out


def load_actors(mat_path: str):
    """``actorsGT.mat`` -> object array [person][frame] of [14,3] poses in metres or empty entries (campus.py:139-141)."""
    import scipy.io as scio
    actor_3d = scio.loadmat(mat_path)["actor3D"]
    return np.array(np.array(actor_3d.tolist()).tolist(), dtype=object).squeeze()


def evaluate_pcp(preds: Sequence, actor_3d, frame_range: Sequence[int], dataset: str, recall_threshold: float = 500):
    """``Campus.evaluate`` / ``Shelf.evaluate`` (campus.py:138-210, shelf.py:162-234).  preds[i]: [P,17,>=4] for frame
    frame_range[i]; actor_3d[person][frame]: [14,3] in metres, or an empty entry when the actor is absent.
    Returns (average PCP of the first three actors, message, details)."""
    assert dataset in ("campus", "shelf")
    convert = coco2campus3D if dataset == "campus" else coco2shelf3D
    num_person = len(actor_3d)
    total_gt = match_gt = 0
    alpha = 0.5
    correct_parts = np.zeros(num_person)
    total_parts = np.zeros(num_person)
    bone_correct = np.zeros((num_person, 10))
    for i, fi in enumerate(frame_range):
        p = _np(preds[i])
        p = p[p[:, 0, 3] >= 0, :, :3]
        if len(p) == 0:
            if dataset == "campus":
                continue                                     # campus.py:156-157
            raise ValueError("need at least one array to stack")   # what shelf.py:179 does on an empty frame
        pred = np.stack([convert(q) for q in p.copy()])
        for person in range(num_person):
            gt = actor_3d[person][fi] * 1000.0
            if len(gt[0]) == 0:
                continue
            dist = np.mean(np.sqrt(np.sum((gt[np.newaxis] - pred) ** 2, axis=-1)), axis=-1)
            m = int(np.argmin(dist))
            if np.min(dist) < recall_threshold:
                match_gt += 1
            total_gt += 1
            for j, (s, e) in enumerate(LIMBS):
                total_parts[person] += 1
                err_s = np.linalg.norm(pred[m, s, 0:3] - gt[s])
                err_e = np.linalg.norm(pred[m, e, 0:3] - gt[e])
                if (err_s + err_e) / 2.0 <= alpha * np.linalg.norm(gt[s] - gt[e]):
                    correct_parts[person] += 1
                    bone_correct[person, j] += 1
            pred_hip = (pred[m, 2, 0:3] + pred[m, 3, 0:3]) / 2.0             # torso: hip centre -> head bottom
            gt_hip = (gt[2] + gt[3]) / 2.0
            total_parts[person] += 1
            err_s = np.linalg.norm(pred_hip - gt_hip)
            err_e = np.linalg.norm(pred[m, 12, 0:3] - gt[12])
            if (err_s + err_e) / 2.0 <= alpha * np.linalg.norm(gt_hip - gt[12]):
                correct_parts[person] += 1
                bone_correct[person, 9] += 1
    actor_pcp = correct_parts / (total_parts + 1e-8)
    avg_pcp = np.mean(actor_pcp[:3])
    bone_pcp: Dict[str, np.ndarray] = OrderedDict()
    for name, idx in BONE_GROUPS.items():
        bone_pcp[name] = np.sum(bone_correct[:, idx], axis=-1) / (total_parts / 10 * len(idx) + 1e-8)
    recall = match_gt / (total_gt + 1e-8)
    msg = ("     | Actor 1 | Actor 2 | Actor 3 | Average | \n"
           " PCP |  {:.2f}  |  {:.2f}  |  {:.2f}  |  {:.2f}  |\t Recall@500mm: {:.4f}").format(
               actor_pcp[0] * 100, actor_pcp[1] * 100, actor_pcp[2] * 100, avg_pcp * 100, recall)
    return float(np.mean(avg_pcp)), msg, {"actor_pcp": actor_pcp, "bone_pcp": bone_pcp, "recall": recall,
                                          "total_gt": total_gt, "match_gt": match_gt}
