#!/usr/bin/env python
"""Golden vectors of the task metrics, produced by the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden_eval.py          # writes tests/golden/eval_<dataset>.npz

The reference's ``evaluate`` methods live on its dataset classes (lib/dataset/panoptic.py:214, campus.py:138,
shelf.py:162), whose constructors need the datasets.  The methods themselves only touch ``self.db`` / ``self.db_size``
(Panoptic) or ``self.frame_range`` / ``self.dataset_dir`` + ``actorsGT.mat`` (Campus, Shelf), so the script creates the
objects with ``object.__new__``, fills those attributes with a synthetic scene and calls the reference methods unchanged;
``scipy.io.loadmat`` is pointed at an in-memory ``actor3D`` cell array of the real file's shape (1 x actors, each
frames x 1 of 14x3 doubles or empty).  Scenes contain what the metrics are sensitive to: exact and noisy detections,
duplicates of one person, false positives, misses, invisible joints, frames without people, frames without predictions.
The script asserts ``fvp.evaluate`` equal to the reference (metric, message, every intermediate number) and stores
inputs + reference outputs.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
sys.path.insert(0, ROOT)

from fvp import evaluate as E                  # noqa: E402
from oracle import gen_golden as GG            # noqa: E402  (easydict shim + reference path)

OUT = os.path.join(ROOT, "tests", "golden")
COCO_FROM_14 = {16: 0, 14: 1, 12: 2, 11: 3, 13: 4, 15: 5, 10: 6, 8: 7, 6: 8, 5: 9, 7: 10, 9: 11}   # coco joint -> 14-order joint


def reference_datasets():
    GG._install_reference()
    sys.modules.setdefault("json_tricks", types.ModuleType("json_tricks"))     # only used by the loaders we never call
    import importlib                      # dataset/__init__.py rebinds the sub-module names to the classes
    return tuple(importlib.import_module("dataset." + m) for m in ("panoptic", "campus", "shelf"))


def random_person(rng, J):
    root = np.array([rng.uniform(-2500, 2500), rng.uniform(-3000, 2000), 900.0])
    return root + rng.uniform(-0.5, 0.5, (J, 3)) * np.array([500.0, 500.0, 1600.0])


def predictions_for(rng, gts, P, J, flags_all_invalid=False):
    """[P,J,5] float32 slots: noisy copies, duplicates, false positives, invalid slots (flag -1, zeros)."""
    out = np.zeros((P, J, 5), np.float32)
    out[:, :, 3] = -1.0
    poses = []
    for g in gts:
        r = rng.random()
        if r < 0.12:
            continue                                                    # missed person
        sigma = [3.0, 15.0, 40.0, 90.0, 300.0][int(rng.integers(0, 5))]
        poses.append(g + rng.normal(0, sigma, g.shape))
        if r > 0.85:
            poses.append(g + rng.normal(0, 25.0, g.shape))              # duplicate detection of the same person
    for _ in range(int(rng.integers(0, 3))):
        poses.append(random_person(rng, J))                             # false positive
    order = rng.permutation(len(poses))[:P]
    if not flags_all_invalid:
        for slot, k in enumerate(order):
            out[slot, :, :3] = poses[k]
            out[slot, :, 3] = 0.0
            out[slot, :, 4] = rng.uniform(0.05, 1.0)
    return out


def panoptic_case(panoptic):
    rng = np.random.default_rng(7)
    J, P, F, MAXP = 15, 10, 40, 10
    db, preds, n_people = [], [], []
    gt_j = np.zeros((F, MAXP, J, 3))
    gt_v = np.zeros((F, MAXP, J))
    for f in range(F):
        n = int(rng.integers(0, 6)) if f % 7 else 0                       # some frames without people
        gts = [random_person(rng, J) for _ in range(n)]
        vis = [np.where(rng.random(J) > 0.2, rng.uniform(0.2, 1.0, J), rng.uniform(0.0, 0.1, J)) for _ in range(n)]
        for v in vis:
            v[2] = 0.9                                                    # root always visible (panoptic.py:145)
        for k in range(n):
            gt_j[f, k], gt_v[f, k] = gts[k], vis[k]
        db.append({"meta": {"num_person": n, "joints_3d": gt_j[f].copy(), "joints_3d_vis": gt_v[f].copy()}})
        preds.append(predictions_for(rng, gts, P, J, flags_all_invalid=(f % 11 == 5)))
        n_people.append(n)
    ds = object.__new__(panoptic.Panoptic)
    ds.db, ds.db_size = db, F
    metric, msg = ds.evaluate([torch.from_numpy(p) for p in preds])
    ours = E.evaluate_panoptic(preds, [gt_j[f, :n_people[f]] for f in range(F)], [gt_v[f, :n_people[f]] for f in range(F)])
    assert ours[0] == metric and ours[1] == msg, (ours[:2], metric, msg)
    # the static helpers, on our records
    rec, total = E.match_poses(preds, [gt_j[f, :n_people[f]] for f in range(F)], [gt_v[f, :n_people[f]] for f in range(F)])
    ref_aps = [panoptic.Panoptic._eval_list_to_ap(list(rec), total, t) for t in E.AP_THRESHOLDS]
    assert [(a, r) for a, r in zip(ours[2]["aps"], ours[2]["recalls"])] == [(float(a), float(r)) for a, r in ref_aps]
    assert ours[2]["mpjpe"] == float(panoptic.Panoptic._eval_list_to_mpjpe(list(rec)))
    assert ours[2]["recall"] == panoptic.Panoptic._eval_list_to_recall(list(rec), total)
    np.savez_compressed(os.path.join(OUT, "eval_panoptic.npz"), preds=np.stack(preds), gt_joints=gt_j, gt_vis=gt_v,
                        num_person=np.array(n_people), metric=metric, msg=msg, aps=np.array(ours[2]["aps"]),
                        recalls=np.array(ours[2]["recalls"]), mpjpe=ours[2]["mpjpe"], recall=ours[2]["recall"])
    print("panoptic:", msg.replace("\n", " | "))
    return {"metric": metric, "poses": ours[2]["poses"], "total_gt": total}


def coco_from_14(rng, g14):
    """A COCO-order [17,3] pose whose limbs coincide with the 14-joint ground truth (head joints invented around it)."""
    c = np.zeros((17, 3))
    for cj, k in COCO_FROM_14.items():
        c[cj] = g14[k]
    head = g14[12] + (g14[13] - g14[12]) * 0.5
    c[0] = head + rng.normal(0, 20, 3)             # nose
    c[1], c[2] = head + rng.normal(0, 20, 3), head + rng.normal(0, 20, 3)
    c[3], c[4] = head + np.array([-70.0, 0, 0]) + rng.normal(0, 10, 3), head + np.array([70.0, 0, 0]) + rng.normal(0, 10, 3)
    return c


def pcp_case(mod, cls_name, dataset, frame_range, seed):
    import scipy.io as scio
    rng = np.random.default_rng(seed)
    A, P, J = 4, 10, 17
    nframes = max(frame_range) + 1
    cells = np.empty((1, A), dtype=object)
    present = np.zeros((A, nframes), bool)
    gt_m = np.zeros((A, nframes, 14, 3))
    for a in range(A):
        col = np.empty((nframes, 1), dtype=object)
        for f in range(nframes):
            if rng.random() < 0.7:
                g = random_person(rng, 14) / 1000.0                       # metres, like actorsGT.mat
                g[13] = g[12] + np.array([0.0, 0.0, 0.2]) + rng.normal(0, 0.01, 3)
                col[f, 0] = g
                present[a, f], gt_m[a, f] = True, g
            else:
                col[f, 0] = np.zeros((1, 0))                      # an absent actor: `len(gt[0]) == 0` (campus.py:164)
        cells[0, a] = col
    preds = []
    for i, fi in enumerate(frame_range):
        gts = [coco_from_14(rng, gt_m[a, fi] * 1000.0) for a in range(A) if present[a, fi]]
        empty = dataset == "campus" and i % 9 == 4                        # Shelf.evaluate cannot digest an empty frame
        p = predictions_for(rng, gts, P, J, flags_all_invalid=empty)
        if dataset == "shelf" and not (p[:, 0, 3] >= 0).any():
            p[0, :, :3], p[0, :, 3], p[0, :, 4] = random_person(rng, J), 0.0, 0.5
        preds.append(p)
    ds = object.__new__(getattr(mod, cls_name))
    ds.frame_range, ds.dataset_dir = list(frame_range), "/nonexistent"
    real_loadmat = scio.loadmat
    mod.scio.loadmat = lambda path: {"actor3D": cells}
    try:
        metric, msg = ds.evaluate([torch.from_numpy(p) for p in preds])
    finally:
        mod.scio.loadmat = real_loadmat
    actors = np.array(np.array(cells.tolist()).tolist(), dtype=object).squeeze()      # campus.py:141
    ours = E.evaluate_pcp(preds, actors, frame_range, dataset)
    assert ours[0] == metric and ours[1] == msg, (ours[:2], metric, msg)
    conv_ref = getattr(getattr(mod, cls_name), "coco2campus3D" if dataset == "campus" else "coco2shelf3D")
    conv = E.coco2campus3D if dataset == "campus" else E.coco2shelf3D
    for p in preds[:5]:
        for q in p[p[:, 0, 3] >= 0, :, :3]:
            assert np.array_equal(conv_ref(q.copy()), conv(q.copy()))
    fr = np.array(list(frame_range))                                     # only the evaluated frames are stored
    np.savez_compressed(os.path.join(OUT, "eval_%s.npz" % dataset), preds=np.stack(preds), gt_metres=gt_m[:, fr], present=present[:, fr],
                        frame_range=fr, metric=metric, msg=msg, actor_pcp=ours[2]["actor_pcp"],
                        bone_pcp=np.stack(list(ours[2]["bone_pcp"].values())), recall=ours[2]["recall"])
    print(dataset + ":", msg.replace("\n", " | "))
    return {"metric": metric, "total_gt": ours[2]["total_gt"], "match_gt": ours[2]["match_gt"]}


def main():
    panoptic, campus, shelf = reference_datasets()
    manifest = {"panoptic": panoptic_case(panoptic),
                "campus": pcp_case(campus, "Campus", "campus", list(range(350, 380)) + list(range(650, 670)), 11),
                "shelf": pcp_case(shelf, "Shelf", "shelf", list(range(300, 345)), 12)}
    with open(os.path.join(OUT, "MANIFEST_eval.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
