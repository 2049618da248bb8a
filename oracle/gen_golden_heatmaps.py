#!/usr/bin/env python
"""Golden vectors of the heat-map renderer, produced by the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden_heatmaps.py          # writes tests/golden/heatmaps_<case>.npz

For every case the script builds the reference ``JointsDataset`` (lib/dataset/JointsDataset.py, evaluation mode), feeds a
synthetic database record through the reference's own ``__getitem__`` ('pred' and 'gt' heat-map sources) and
  1. asserts ``oracle.heatmap_oracle`` bit-identical to the reference output (the oracle pin),
  2. stores the poses (float64), the reference's ``resize_transform``, and the rendered maps (float32) in an npz.
Cases cover: crowds with overlapping people, joints outside the image / negative coordinates, tiny and huge people
(both clip bounds of compute_human_scale), patches cut by every border, joints_vis masks, Panoptic / Campus / Shelf sizes.
"""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
sys.path.insert(0, ROOT)

from fvp import config as fcfg, synth          # noqa: E402
from oracle import heatmap_oracle as HO        # noqa: E402
from oracle import gen_golden as GG            # noqa: E402  (easydict shim + reference path)

OUT = os.path.join(ROOT, "tests", "golden")


def make_poses(rng, ori_size, n_people, J, mode):
    """[n][J,3] poses in original-image pixels (x, y, score)."""
    W, H = ori_size
    poses = []
    for n in range(n_people):
        if mode == "crowd":
            c = rng.uniform([0.1 * W, 0.1 * H], [0.9 * W, 0.9 * H])
            ext = rng.uniform(60, 500)
        elif mode == "tiny":
            c = rng.uniform([0.2 * W, 0.2 * H], [0.8 * W, 0.8 * H])
            ext = rng.uniform(2, 40)                       # below the lower clip of compute_human_scale
        elif mode == "huge":
            c = rng.uniform([0.3 * W, 0.3 * H], [0.7 * W, 0.7 * H])
            ext = rng.uniform(900, 2500)                   # above the upper clip
        else:  # "border": people hanging over every image edge, some completely outside
            edge = n % 5
            c = [np.array([-30.0, 0.5 * H]), np.array([W + 20.0, 0.4 * H]), np.array([0.5 * W, -25.0]),
                 np.array([0.6 * W, H + 15.0]), np.array([-900.0, -900.0])][edge] + rng.uniform(-20, 20, 2)
            ext = rng.uniform(80, 400)
        p = np.zeros((J, 3))
        p[:, :2] = c + rng.uniform(-0.5, 0.5, (J, 2)) * ext * np.array([0.5, 1.0])
        p[:, 2] = rng.uniform(0.2, 1.0, J)
        poses.append(p)
    return poses


def main():
    GG._install_reference()
    # the reference class, unmodified; loaded from its file because lib/dataset/__init__.py pulls the concrete
    # datasets, which need json_tricks (absent here)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_JointsDataset", os.path.join(GG.REF, "lib/dataset/JointsDataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    JointsDataset = mod.JointsDataset
    from utils.transforms import affine_transform

    cases = [("panoptic_256x192", "crowd", 10, 101), ("panoptic", "border", 10, 102), ("campus", "tiny", 4, 103),
             ("shelf", "huge", 3, 104), ("shelf", "crowd", 10, 105)]
    manifest = {}
    for preset, mode, n_people, seed in cases:
        cfg = fcfg.preset(preset)
        cfg.DATASET.TEST_HEATMAP_SRC = "pred"
        V, J = int(cfg.DATASET.CAMERA_NUM), int(cfg.DATASET.NUM_JOINTS)
        rng = np.random.default_rng(seed)
        ds = JointsDataset(cfg, is_train=False)
        # ---- 'pred' source through the reference's own __getitem__ ---------------------------------------
        all_preds = [make_poses(rng, cfg.DATASET.ORI_IMAGE_SIZE, n_people - (v % 2), J, mode) for v in range(V)]
        rec = {"pred_pose2d": copy.deepcopy(all_preds), "target": 0, "meta": {"seq": "s"}, "image": ""}
        ds.db = [rec]
        _, _, _, hm_ref = ds[0]
        hm_ref = hm_ref.numpy()
        hm_or = HO.pred_heatmaps(all_preds, ds.resize_transform, cfg.DATASET.HEATMAP_SIZE, cfg.DATASET.IMAGE_SIZE, cfg.NETWORK.SIGMA)
        assert hm_ref.dtype == np.float32 and hm_ref.shape == hm_or.shape
        nd = int((hm_ref.view(np.int32) != hm_or.view(np.int32)).sum())
        assert nd == 0, "%s/%s: oracle differs from the reference renderer in %d values" % (preset, mode, nd)
        # the affine as the reference applies it (np.dot) vs the oracle's explicit order
        moved_ref = np.array([[[affine_transform(p[j, :2], ds.resize_transform) for j in range(J)] for p in view] for view in all_preds[:1]])
        moved_or = np.array([HO.affine_points(np.array(p)[:, :2], ds.resize_transform) for p in all_preds[0]])[None]
        assert np.array_equal(moved_ref, moved_or), "affine order differs"
        # ---- 'gt' source (3-D joints + calibration) -------------------------------------------------------
        cal = {"panoptic_256x192": "panoptic", "panoptic": "panoptic", "campus": "campus", "shelf": "shelf"}[preset]
        cams = GG._load_calibration(cal)
        sk = synth.make_skeletons(cfg, min(n_people, 6), seed=seed)
        j3 = [np.asarray(s, np.float64) for s in sk]
        vis = [(rng.random(J) > 0.15).astype(np.float64) for _ in j3]
        cfg.DATASET.TEST_HEATMAP_SRC = "gt"
        ds2 = JointsDataset(cfg, is_train=False)
        ds2.cameras = {"s": cams}
        ds2.db = [{"target": 0, "meta": {"seq": "s", "joints_3d": copy.deepcopy(j3), "joints_3d_vis": copy.deepcopy(vis)}, "image": ""}]
        _, _, _, gt_ref = ds2[0]
        gt_ref = gt_ref.numpy()
        gt_or = HO.gt_heatmaps(j3, vis, cams, ds2.resize_transform, cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE,
                               cfg.DATASET.HEATMAP_SIZE, cfg.NETWORK.SIGMA)
        nd = int((gt_ref.view(np.int32) != gt_or.view(np.int32)).sum())
        assert nd == 0, "%s gt: oracle differs from the reference in %d values" % (preset, nd)
        name = "heatmaps_%s_%s" % (preset, mode)
        npeople = np.array([len(v) for v in all_preds], np.int32)
        preds = np.zeros((V, n_people, J, 3))
        for v in range(V):
            preds[v, : npeople[v]] = np.array(all_preds[v])
        # the maps are sparse: store the non-zero values only
        nz = np.flatnonzero(hm_ref)
        nzg = np.flatnonzero(gt_ref)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), preset=preset, mode=mode, preds=preds, num_people=npeople,
                            resize=ds.resize_transform, pred_nz_index=nz.astype(np.int64), pred_nz_value=hm_ref.ravel()[nz],
                            joints_3d=np.array(j3), joints_3d_vis=np.array(vis), cameras=synth.cameras_to_array(cams),
                            gt_nz_index=nzg.astype(np.int64), gt_nz_value=gt_ref.ravel()[nzg], shape=np.array(hm_ref.shape))
        manifest[name] = {"preset": preset, "mode": mode, "views": V, "people": int(n_people), "nonzero_pred": int(nz.size),
                          "nonzero_gt": int(nzg.size), "max": float(hm_ref.max())}
        print(name, manifest[name])
    with open(os.path.join(OUT, "MANIFEST_heatmaps.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
