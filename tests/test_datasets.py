"""fvp.datasets (readers of the reference's calibration / detection files) against fixtures written by
oracle/gen_golden_datasets.py, which asserted equality with the reference's own dataset loaders (Campus/Shelf/Panoptic
``_get_cam``, ``_get_db``) when it produced them.  CPU only, no reference needed at run time."""
import json
import os
import pickle

import numpy as np
import pytest

from golden_util import GOLDEN_DIR

from fvp import datasets as D


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN_DIR, "datasets_fixture.npz"))


def _check(cams, fx, prefix):
    assert len(cams) == fx[prefix + "_R"].shape[0]
    for v, c in enumerate(cams):
        assert np.array_equal(np.asarray(c["R"], np.float64), fx[prefix + "_R"][v])
        assert np.array_equal(np.asarray(c["T"], np.float64).reshape(3, 1), fx[prefix + "_T"][v])
        assert [float(c["fx"]), float(c["fy"])] == list(fx[prefix + "_f"][v])
        assert [float(c["cx"]), float(c["cy"])] == list(fx[prefix + "_c"][v])
        assert np.array_equal(np.asarray(c["k"], np.float64).reshape(3, 1), fx[prefix + "_k"][v])
        assert np.array_equal(np.asarray(c["p"], np.float64).reshape(2, 1), fx[prefix + "_p"][v])


@pytest.mark.parametrize("name,views", [("campus", 3), ("shelf", 5)])
def test_campus_shelf_calibration_files(fx, tmp_path, name, views):
    """calibration_{campus,shelf}.json -> cameras[seq] exactly as Campus/Shelf._get_cam builds it (campus.py:114-129)."""
    path = tmp_path / ("calibration_%s.json" % name)
    path.write_text(str(fx[name + "_json"]))
    cams = D.load_calibration(str(path))
    assert isinstance(cams, list) and len(cams) == views and all(isinstance(c["R"], np.ndarray) for c in cams)
    _check(cams, fx, name)


def test_demo_calibration_and_rejections(fx, tmp_path):
    path = tmp_path / "calibration.json"
    path.write_text(str(fx["demo_json"]))
    d = D.load_calibration(str(path))
    assert list(d) == ["customized_sequence"] and len(d["customized_sequence"]) == 5
    assert d["customized_sequence"][0]["R"].shape == (3, 3)
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"0": {"R": [[1, 0, 0]] * 3}}))
    with pytest.raises(KeyError):
        D.load_calibration(str(bad))
    doc = json.loads(str(fx["campus_json"]))
    bad.write_text(json.dumps({"0": doc["0"], "2": doc["2"]}))               # view 1 missing
    with pytest.raises(ValueError):
        D.load_calibration(str(bad))
    bad.write_text(json.dumps([1, 2, 3]))
    with pytest.raises(ValueError):
        D.load_calibration(str(bad))


def test_panoptic_camera_selection_and_conversion(fx, tmp_path):
    """calibration_<seq>.json -> the five HD cameras, R M, T = -R'^T t * 10, k / p from distCoef (panoptic.py:171-205)."""
    calib = json.loads(str(fx["panoptic_json"]))
    assert len(calib["cameras"]) == 8
    cams = D.panoptic_cameras(calib)
    _check(cams, fx, "panoptic")
    path = tmp_path / "calibration_seq.json"
    path.write_text(str(fx["panoptic_json"]))
    three = D.panoptic_cameras(str(path), 3)
    assert len(three) == 3 and all(np.array_equal(a["T"], b["T"]) for a, b in zip(three, cams))
    R = cams[0]["R"]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and cams[0]["T"].shape == (3, 1)


def test_detection_file_frames(fx):
    """pred_*_maskrcnn_hrnet_coco.pkl -> db_rec['pred_pose2d'] of a frame (campus.py:92-97); frame 400 of both files."""
    pred = D.load_pred_pose2d(os.path.join(GOLDEN_DIR, "pred_pose2d_frame400.pkl"))
    for name, views in (("campus", 3), ("shelf", 5)):
        frame = D.frame_preds(pred[name], 400, views)
        assert [len(v) for v in frame] == list(fx[name + "_people_per_view"])
        assert all(p.shape == (17, 3) and p.dtype == np.float64 for v in frame for p in v)
        assert D.batch_preds(pred[name], [400, 400], views)[1][0][0] is not frame[0][0]
        assert [i for i, _ in D.pred_frames(pred[name], [400], views)] == [400]
    with pytest.raises(KeyError):
        D.frame_preds(pred["campus"], 401, 3)                                    # a plain dict raises on a missing frame
    assert len(D.CAMPUS_FRAMES) == 222 and D.CAMPUS_FRAMES[0] == 350 and D.CAMPUS_FRAMES[-1] == 750 and 500 not in D.CAMPUS_FRAMES
    assert D.SHELF_FRAMES == list(range(300, 601))


def test_panoptic_annotation_files(fx, tmp_path):
    """body3DScene_*.json -> per-frame ground truth as Panoptic._get_db stores it (panoptic.py:109-163): y-up centimetres to
    z-up millimetres, visibilities clipped at 0, bodies with an invisible root dropped, every 12th file visited."""
    names, texts = [str(n) for n in fx["panoptic_anno_names"]], [str(t) for t in fx["panoptic_anno_texts"]]
    adir = tmp_path / "seq" / "hdPose3d_stage1_coco19"
    adir.mkdir(parents=True)
    for i in range(26):
        name = "body3DScene_%08d.json" % (100 + i)
        (adir / name).write_text(texts[names.index(name)] if name in names else json.dumps({"bodies": []}))
    files = D.panoptic_annotation_files(str(tmp_path / "seq"), 12)
    assert [os.path.basename(f) for f in files] == names
    assert len(D.panoptic_annotation_files(str(tmp_path / "seq"), 3)) == 9
    counts = [len(D.panoptic_frame_gt(f)[0]) for f in files]
    assert counts == list(fx["panoptic_gt_counts"]) == [3, 0, 3]
    j, v = D.panoptic_frame_gt(json.loads(texts[0]))
    assert np.array_equal(np.array(j), fx["panoptic_gt_joints_first"]) and np.array_equal(np.array(v), fx["panoptic_gt_vis_first"])
    assert all(a.shape == (15, 3) and b.shape == (15,) and b.min() >= 0.0 for a, b in zip(j, v))
    raw = np.array(json.loads(texts[0])["bodies"][0]["joints19"]).reshape(-1, 4)
    assert np.allclose(j[0][4], [raw[4, 0] * 10, raw[4, 2] * 10, -raw[4, 1] * 10])          # (x, y, z) cm y-up -> (x, z, -y) mm
    paths = D.panoptic_image_paths("/data/panoptic", "160906_pizza1", files[0], 5)
    assert paths[0] == "/data/panoptic/160906_pizza1/hdImgs/00_03/00_03_00000100.jpg" and paths[4].endswith("00_23/00_23_00000100.jpg")


def test_load_views_equals_torchvision_transform(tmp_path):
    """imread -> BGR2RGB -> ToTensor -> Normalize (JointsDataset.py:124-134 + run/validate.py:44-52), bit for bit against
    torchvision, the library the reference calls."""
    cv2 = pytest.importorskip("cv2")
    tv = pytest.importorskip("torchvision.transforms")
    import torch
    rng = np.random.default_rng(5)
    paths = []
    for v in range(3):
        img = rng.integers(0, 256, (24, 40, 3), dtype=np.uint8)
        p = str(tmp_path / ("v%d.png" % v))
        assert cv2.imwrite(p, img)
        paths.append(p)
    tf = tv.Compose([tv.ToTensor(), tv.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ref = []
    for p in paths:
        inp = cv2.imread(p, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        inp = cv2.cvtColor(inp, cv2.COLOR_BGR2RGB)
        ref.append(tf(inp))
    ref = torch.stack(ref, dim=0).numpy()
    got = D.load_views(paths)
    assert got.dtype == np.float32 and got.shape == (3, 3, 24, 40)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))
    bgr = D.load_views(paths, color_rgb=False)
    assert np.array_equal(bgr[:, 1], got[:, 1]) and not np.array_equal(bgr[:, 0], got[:, 0])     # green stays, red / blue swap
    with pytest.raises(FileNotFoundError):
        D.load_views([str(tmp_path / "missing.png")])


def test_panoptic_records_skip_rules(fx, tmp_path):
    """panoptic.py:116-163: every 12th annotation file; files without bodies, with a missing HD image or without a kept body
    give no record."""
    names, texts = [str(n) for n in fx["panoptic_anno_names"]], [str(t) for t in fx["panoptic_anno_texts"]]
    seq = "160906_pizza1"
    adir = tmp_path / seq / "hdPose3d_stage1_coco19"
    adir.mkdir(parents=True)
    for i in range(26):
        name = "body3DScene_%08d.json" % (100 + i)
        (adir / name).write_text(texts[names.index(name)] if name in names else json.dumps({"bodies": []}))
        for panel, node in D.PANOPTIC_CAM_LIST:
            idir = tmp_path / seq / "hdImgs" / ("%02d_%02d" % (panel, node))
            idir.mkdir(parents=True, exist_ok=True)
            (idir / ("%02d_%02d_%08d.jpg" % (panel, node, 100 + i))).write_bytes(b"")
    db = D.panoptic_records(str(tmp_path), [seq])
    assert [os.path.basename(r["all_image_path"][0]) for r in db] == ["00_03_00000100.jpg", "00_03_00000124.jpg"]
    assert [len(r["joints_3d"]) for r in db] == [3, 3] and db[0]["seq"] == seq and len(db[0]["all_image_path"]) == 5
    os.remove(db[1]["all_image_path"][2])                                         # one HD image of the last frame missing
    assert len(D.panoptic_records(str(tmp_path), [seq])) == 1
    assert len(D.panoptic_records(str(tmp_path), [seq], num_views=2)) == 2        # ... which a 2-view run does not need
