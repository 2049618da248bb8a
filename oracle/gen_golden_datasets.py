#!/usr/bin/env python
"""Fixtures that pin ``fvp.datasets`` (the data-format readers either side of the hot path) to the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden_datasets.py        # writes tests/golden/datasets_fixture.npz + pred_pose2d_frame400.pkl

What it does
  1. calls the reference's own ``Campus._get_cam`` / ``Shelf._get_cam`` (lib/dataset/campus.py:114-129, shelf.py) on the
     calibration files shipped in /root/reference/data and asserts ``fvp.datasets.load_calibration`` returns the same
     values (``==`` on every array, same dtypes); stores the calibration JSON text and the reference's arrays;
  2. writes a synthetic Panoptic ``calibration_<seq>.json`` (eight HD / VGA cameras, five of them the reference's
     ``cam_list``), runs the reference's ``Panoptic._get_cam`` (panoptic.py:171-205) on it and asserts
     ``fvp.datasets.panoptic_cameras`` equal; stores the JSON text and the expected cameras;
  2b. builds a synthetic Panoptic sequence folder (annotation files, empty image files), runs the reference's
     ``Panoptic._get_db`` (panoptic.py:109-167, with the unmodified ``JointsDataset`` constructor and ``_rebuild_db``) and
     asserts ``fvp.datasets.panoptic_annotation_files`` / ``panoptic_image_paths`` / ``panoptic_frame_gt`` equal;
  3. cuts frame 400 (BASELINE configs[0], SURVEY.md 8d "Config 1") out of the shipped detection files
     ``pred_{campus,shelf}_maskrcnn_hrnet_coco.pkl`` into a small pickle, builds ``db_rec['pred_pose2d']`` with the very
     expression of ``Campus._get_db`` (campus.py:92-97) and asserts ``fvp.datasets.frame_preds`` equal.
"""
from __future__ import annotations

import importlib
import json
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
sys.path.insert(0, ROOT)

from fvp import datasets as D                   # noqa: E402
from oracle import gen_golden as GG             # noqa: E402  (easydict shim + reference path)

OUT = os.path.join(ROOT, "tests", "golden")
FRAME = 400


def reference_dataset_modules():
    GG._install_reference()
    sys.modules["json_tricks"] = json           # the loaders only call json.load on plain JSON files
    return tuple(importlib.import_module("dataset." + m) for m in ("panoptic", "campus", "shelf"))


def cams_equal(ours, ref) -> bool:
    if len(ours) != len(ref):
        return False
    for a, b in zip(ours, ref):
        if sorted(a) != sorted(b):
            return False
        for k in a:
            x, y = np.asarray(a[k]), np.asarray(b[k])
            if x.shape != y.shape or x.dtype != y.dtype or not np.array_equal(x, y):
                return False
    return True


def cams_to_arrays(cams) -> dict:
    """{R [V,3,3], T [V,3,1], f [V,2], c [V,2], k [V,3,1], p [V,2,1]} float64."""
    return {"R": np.stack([np.asarray(c["R"], np.float64) for c in cams]),
            "T": np.stack([np.asarray(c["T"], np.float64).reshape(3, 1) for c in cams]),
            "f": np.array([[float(c["fx"]), float(c["fy"])] for c in cams]),
            "c": np.array([[float(c["cx"]), float(c["cy"])] for c in cams]),
            "k": np.stack([np.asarray(c["k"], np.float64).reshape(3, 1) for c in cams]),
            "p": np.stack([np.asarray(c["p"], np.float64).reshape(2, 1) for c in cams])}


def synthetic_panoptic_calibration(rng) -> dict:
    cams = []
    nodes = [(0, 1), (0, 3), (5, 7), (0, 6), (0, 12), (0, 13), (3, 3), (0, 23)]       # HD panel 0 + two VGA panels
    for panel, node in nodes:
        a = rng.uniform(0, 2 * np.pi)
        q = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        q *= np.sign(np.linalg.det(q))
        K = np.array([[rng.uniform(1300, 1700), 0.0, rng.uniform(900, 1000)], [0.0, rng.uniform(1300, 1700), rng.uniform(500, 580)],
                      [0.0, 0.0, 1.0]])
        cams.append({"name": "%02d_%02d" % (panel, node), "type": "hd" if panel == 0 else "vga", "panel": panel, "node": node,
                     "resolution": [1920, 1080], "K": K.tolist(), "distCoef": rng.uniform(-0.3, 0.3, 5).tolist(),
                     "R": q.tolist(), "t": (rng.uniform(-300, 300, (3, 1)) + np.array([[np.cos(a)], [0.0], [np.sin(a)]]) * 200).tolist()})
    return {"calibDataSource": "synthetic", "cameras": cams}


def main():
    panoptic_m, campus_m, shelf_m = reference_dataset_modules()
    store = {}

    # ---- 1. Campus / Shelf calibration files through the reference's _get_cam ----------------------------------
    for name, mod, cls, sub, fn in (("campus", campus_m, "Campus", "Campus", "calibration_campus.json"),
                                    ("shelf", shelf_m, "Shelf", "Shelf", "calibration_shelf.json")):
        ds = object.__new__(getattr(mod, cls))
        ds.dataset_dir = os.path.join(GG.REF, "data", sub)
        ref = ds._get_cam()[name]                                  # {int view: camera}
        ref_list = [ref[i] for i in range(len(ref))]
        path = os.path.join(ds.dataset_dir, fn)
        ours = D.load_calibration(path)
        assert cams_equal(ours, ref_list), "%s: load_calibration differs from the reference's _get_cam" % name
        store[name + "_json"] = np.array(open(path).read())
        for k, v in cams_to_arrays(ref_list).items():
            store["%s_%s" % (name, k)] = v
        print("%-8s calibration: %d cameras, load_calibration == reference" % (name, len(ours)))
    demo = D.load_calibration(os.path.join(GG.REF, "demo", "calibration.json"))
    assert list(demo) == ["customized_sequence"] and len(demo["customized_sequence"]) == 5
    store["demo_json"] = np.array(open(os.path.join(GG.REF, "demo", "calibration.json")).read())

    # ---- 2. Panoptic conversion on a synthetic calibration_<seq>.json ------------------------------------------
    rng = np.random.default_rng(2024)
    calib = synthetic_panoptic_calibration(rng)
    with tempfile.TemporaryDirectory() as tmp:
        seq = "synthetic_seq"
        os.makedirs(os.path.join(tmp, seq))
        with open(os.path.join(tmp, seq, "calibration_%s.json" % seq), "w") as f:
            json.dump(calib, f)
        for nv in (5, 3):
            ds = object.__new__(panoptic_m.Panoptic)
            ds.dataset_dir, ds.sequence_list, ds.num_views = tmp, [seq], nv
            ds.cam_list = [(0, 3), (0, 6), (0, 12), (0, 13), (0, 23)][:nv]
            ref = ds._get_cam()[seq]
            ours = D.panoptic_cameras(os.path.join(tmp, seq, "calibration_%s.json" % seq), nv)
            assert cams_equal(ours, ref), "panoptic_cameras differs from the reference's _get_cam (%d views)" % nv
            if nv == 5:
                for k, v in cams_to_arrays(ref).items():
                    store["panoptic_%s" % k] = v
    assert D.PANOPTIC_VAL_LIST == panoptic_m.VAL_LIST
    store["panoptic_json"] = np.array(json.dumps(calib))
    print("panoptic calibration: 5 / 3 of 8 cameras selected and converted, panoptic_cameras == reference")

    # ---- 2b. Panoptic annotations through the reference's _get_db on a synthetic sequence folder ---------------------
    from fvp import config as fcfg
    cfg = fcfg.preset("panoptic")
    seq = "synthetic_seq"
    with tempfile.TemporaryDirectory() as tmp:
        adir = os.path.join(tmp, seq, "hdPose3d_stage1_coco19")
        os.makedirs(adir)
        anno_texts = {}
        for i in range(26):                                                    # validation interval 12 -> files 0, 12, 24
            bodies = []
            nb = {0: 4, 12: 0, 24: 3}.get(i, 1)
            for b in range(nb):
                # centimetres, y up; the root must fall inside the capture space (JointsDataset.generate_target asserts it):
                # world = (x, z, -y) * 10 mm with z in [-200, 1800] mm
                j19 = np.concatenate([rng.uniform(-200, 200, (19, 1)), rng.uniform(-170, 10, (19, 1)), rng.uniform(-200, 200, (19, 1)),
                                      rng.uniform(-0.3, 1.0, (19, 1))], axis=1)
                j19[2, 3] = max(j19[2, 3], 0.5)
                if i == 0 and b == 1:
                    j19[2, 3] = 0.05                                           # root visibility <= 0.1: body dropped
                if i == 0 and b == 2:
                    j19[5, 3] = -1.0                                           # negative visibility clipped to 0
                bodies.append({"id": b, "joints19": j19.reshape(-1).tolist()})
            name = "body3DScene_%08d.json" % (100 + i)
            text = json.dumps({"version": 0.7, "univTime": 0.0, "bodies": bodies})
            with open(os.path.join(adir, name), "w") as f:
                f.write(text)
            if i in (0, 12, 24):
                anno_texts[name] = text
            for (panel, node) in D.PANOPTIC_CAM_LIST:
                idir = os.path.join(tmp, seq, "hdImgs", "%02d_%02d" % (panel, node))
                os.makedirs(idir, exist_ok=True)
                open(os.path.join(idir, "%02d_%02d_%08d.jpg" % (panel, node, 100 + i)), "w").close()
        cfg.DATASET.DATADIR = tmp
        ds = object.__new__(panoptic_m.Panoptic)
        panoptic_m.JointsDataset.__init__(ds, cfg, False, None)                # the parent constructor, unmodified
        ds.num_joints, ds.num_views, ds.root_id = 15, 5, cfg.DATASET.ROOT_JOINT_ID
        ds.cam_list, ds.sequence_list, ds._interval = [(0, 3), (0, 6), (0, 12), (0, 13), (0, 23)], [seq], 12
        ds._get_db()                                                            # panoptic.py:109-167 + JointsDataset._rebuild_db
        files = D.panoptic_annotation_files(os.path.join(tmp, seq), 12)
        assert [os.path.basename(f) for f in files] == sorted(anno_texts)
        recs = []
        for f in files:
            j, v = D.panoptic_frame_gt(f, 15, cfg.DATASET.ROOT_JOINT_ID)
            if len(j) > 0:                                                      # panoptic.py:118-119,159: empty frames are skipped
                recs.append((f, j, v))
        assert len(recs) == len(ds.db) == 2
        for (f, j, v), rec in zip(recs, ds.db):
            m = rec["meta"]
            n = m["num_person"]
            assert n == len(j) and m["all_image_path"] == D.panoptic_image_paths(tmp, seq, f, 5)
            assert np.array_equal(m["joints_3d"][:n], np.array(j)) and np.array_equal(m["joints_3d_vis"][:n], np.array(v))
            assert not m["joints_3d"][n:].any()
        # the whole record list, and the missing-image rule (panoptic.py:128-136): drop one HD image of the last frame
        def reference_db():
            d2 = object.__new__(panoptic_m.Panoptic)
            panoptic_m.JointsDataset.__init__(d2, cfg, False, None)
            d2.num_joints, d2.num_views, d2.root_id = 15, 5, cfg.DATASET.ROOT_JOINT_ID
            d2.cam_list, d2.sequence_list, d2._interval = [(0, 3), (0, 6), (0, 12), (0, 13), (0, 23)], [seq], 12
            d2._get_db()
            return d2.db
        for round_ in range(2):
            ref_db, ours_db = reference_db(), D.panoptic_records(tmp, [seq])
            assert len(ref_db) == len(ours_db) == 2 - round_
            for a, b in zip(ours_db, ref_db):
                n = b["meta"]["num_person"]
                assert a["seq"] == b["meta"]["seq"] and a["all_image_path"] == b["meta"]["all_image_path"]
                assert np.array_equal(np.array(a["joints_3d"]), b["meta"]["joints_3d"][:n])
                assert np.array_equal(np.array(a["joints_3d_vis"]), b["meta"]["joints_3d_vis"][:n])
            os.remove(ours_db[-1]["all_image_path"][2])
        store["panoptic_anno_names"] = np.array(sorted(anno_texts))
        store["panoptic_anno_texts"] = np.array([anno_texts[k] for k in sorted(anno_texts)])
        store["panoptic_gt_counts"] = np.array([len(D.panoptic_frame_gt(json.loads(anno_texts[k]))[0]) for k in sorted(anno_texts)])
        store["panoptic_gt_joints_first"] = np.array(recs[0][1])
        store["panoptic_gt_vis_first"] = np.array(recs[0][2])
    print("panoptic annotations: 3 visited files (1 empty), %s people kept, panoptic_frame_gt == reference _get_db"
          % list(store["panoptic_gt_counts"]))

    # ---- 3. frame 400 of the shipped detection files -----------------------------------------------------------
    small = {}
    for name, sub, fn, V in (("campus", "Campus", "pred_campus_maskrcnn_hrnet_coco.pkl", 3),
                             ("shelf", "Shelf", "pred_shelf_maskrcnn_hrnet_coco.pkl", 5)):
        full = D.load_pred_pose2d(os.path.join(GG.REF, "data", sub, fn))
        cut = {"%d_%d" % (k, FRAME): full["%d_%d" % (k, FRAME)] for k in range(V)}
        small[name] = cut
        all_preds = []
        for k in range(V):                                           # campus.py:92-97
            preds = full["{}_{}".format(k, FRAME)]
            preds = [np.array(p["pred"]) for p in preds]
            all_preds.append(preds)
        ours = D.frame_preds(cut, FRAME, V)
        assert len(ours) == V and all(len(a) == len(b) and all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b))
                                      for a, b in zip(ours, all_preds))
        store["%s_people_per_view" % name] = np.array([len(v) for v in all_preds])
        print("%-8s frame %d: %s people per view, frame_preds == reference" % (name, FRAME, [len(v) for v in all_preds]))
    frames = {"campus": (campus_m.Campus, "Campus"), "shelf": (shelf_m.Shelf, "Shelf")}
    # frame ranges as the reference's constructors set them (read from the source: the constructors need the image folders)
    import inspect
    assert "list(range(350, 471)) + list(range(650, 751))" in inspect.getsource(frames["campus"][0].__init__)
    assert "list(range(300, 601))" in inspect.getsource(frames["shelf"][0].__init__)
    with open(os.path.join(OUT, "pred_pose2d_frame400.pkl"), "wb") as f:
        pickle.dump(small, f, protocol=4)
    np.savez_compressed(os.path.join(OUT, "datasets_fixture.npz"), **store)
    print("wrote", os.path.join(OUT, "datasets_fixture.npz"), os.path.getsize(os.path.join(OUT, "datasets_fixture.npz")), "bytes;",
          "pred_pose2d_frame400.pkl", os.path.getsize(os.path.join(OUT, "pred_pose2d_frame400.pkl")), "bytes")


if __name__ == "__main__":
    main()
