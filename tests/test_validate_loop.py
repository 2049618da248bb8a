"""fvp.validate.validate_pred - the run/validate.py:92-118 loop for the 'pred' heat-map source - with a stand-in renderer
and model (CPU): batching, frame order, what the model is called with, the metric hand-off.  The real pieces are tested
where they live (tests/test_gpu_render.py, test_gpu_parity.py, test_gpu_real_detections.py, test_evaluate.py)."""
import copy
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR

from fvp import config as fcfg, datasets as D, synth
from fvp.validate import validate_pred


class _Renderer:
    def __init__(self, cfg):
        self.cfg, self.calls = cfg, []

    def from_pred(self, batch, resize):
        ds = self.cfg.DATASET
        V, J, (W, H) = int(ds.CAMERA_NUM), int(ds.NUM_JOINTS), ds.HEATMAP_SIZE
        assert np.array_equal(resize, synth.resize_transform(ds.ORI_IMAGE_SIZE, ds.IMAGE_SIZE))
        self.calls.append(len(batch))
        hm = torch.zeros((len(batch), V, J, int(H), int(W)))
        for b, frame in enumerate(batch):
            assert len(frame) == V
            hm[b, 0, 0, 0, 0] = float(frame[0][0][0, 0])           # x of joint 0 of the first person in view 0 marks the frame
        return hm


class _Model:
    def __init__(self, cfg):
        self.P, self.J, self.calls = int(cfg.CAPTURE_SPEC.MAX_PEOPLE), int(cfg.DATASET.NUM_JOINTS), []

    def __call__(self, backbone=None, views=None, meta=None, targets=None, input_heatmaps=None, cameras=None,
                 resize_transform=None):
        B = input_heatmaps.shape[0]
        self.calls.append((list(meta["seq"]), sorted(cameras), tuple(resize_transform.shape), backbone, views))
        fused = torch.zeros((B, self.P, self.J, 5))
        fused[:, :, :, 3] = -1.0
        fused[:, 0, :, 3] = 0.0
        fused[:, 0, :, 0] = input_heatmaps[:, 0, 0, 0, 0][:, None]
        return fused, None, None, input_heatmaps, None


def _pred2d(name, views, frames):
    base = D.load_pred_pose2d(os.path.join(GOLDEN_DIR, "pred_pose2d_frame400.pkl"))[name]
    out = {}
    for i, f in enumerate(frames):
        for k in range(views):
            people = copy.deepcopy(base["%d_400" % k])
            for p in people:
                p["pred"] = (np.array(p["pred"]) + np.array([float(i), 0.0, 0.0])).tolist()
            out["%d_%d" % (k, f)] = people
    return out


def test_validation_loop_batches_frames_in_order_and_hands_over_to_the_metric():
    cfg = fcfg.preset("shelf")
    cams = [dict(R=np.eye(3), T=np.zeros((3, 1)), fx=1.0, fy=1.0, cx=0.0, cy=0.0, k=np.zeros((3, 1)), p=np.zeros((2, 1)))] * 5
    frames = [300, 301, 302, 303, 304]
    pred = _pred2d("shelf", 5, frames)
    R, M = _Renderer(cfg), _Model(cfg)
    seen = []
    out = validate_pred(cfg, M, R, cams, pred, frames, "shelf", batch_size=2, progress=lambda a, b: seen.append((a, b)))
    assert R.calls == [2, 2, 1] and seen == [(2, 5), (4, 5), (5, 5)]
    assert [c[0] for c in M.calls] == [["shelf"] * 2, ["shelf"] * 2, ["shelf"]]
    assert all(c[1] == ["shelf"] and c[2] == (2, 3) and c[3] is None and c[4] is None for c in M.calls)
    f = out["fused_poses"]
    assert tuple(f.shape) == (5, 10, 17, 5) and out["metric"] is None
    x0 = float(np.array(pred["0_300"][0]["pred"])[0, 0])
    assert np.allclose(f[:, 0, 0, 0].numpy(), x0 + np.arange(5), atol=1e-4)       # torch.cat keeps the frame order
    # the default batch size is the config's (run/validate.py:56)
    R2 = _Renderer(cfg)
    validate_pred(cfg, _Model(cfg), R2, cams, pred, frames, "shelf")
    assert R2.calls == [int(cfg.TEST.BATCH_SIZE)] * (5 // int(cfg.TEST.BATCH_SIZE)) + ([5 % int(cfg.TEST.BATCH_SIZE)] if 5 % int(cfg.TEST.BATCH_SIZE) else [])
    # metric hand-off: actors as load_actors returns them ([person][frame] -> [14,3] in metres)
    rng = np.random.default_rng(0)
    actors = np.empty((4, 400), dtype=object)
    for a in range(4):
        for fr in range(400):
            actors[a][fr] = rng.uniform(-1, 1, (14, 3)) if fr in frames and a < 3 else np.zeros((1, 0))
    out = validate_pred(cfg, _Model(cfg), _Renderer(cfg), cams, pred, frames, "shelf", actors=actors, batch_size=4)
    assert isinstance(out["metric"], float) and 0.0 <= out["metric"] <= 1.0 and "PCP" in out["msg"]
    assert out["detail"]["total_gt"] == 15
    with pytest.raises(AssertionError):
        validate_pred(cfg, _Model(cfg), _Renderer(cfg), cams[:3], pred, frames, "shelf")


class _GtRenderer:
    def __init__(self, cfg):
        self.cfg, self.calls = cfg, []

    def from_gt(self, joints, vis, cams, resize):
        ds = self.cfg.DATASET
        assert len(joints) == len(vis) == len(cams) and all(len(c) == int(ds.CAMERA_NUM) for c in cams)
        self.calls.append(len(joints))
        hm = torch.zeros((len(joints), int(ds.CAMERA_NUM), int(ds.NUM_JOINTS), 4, 4))
        for b, people in enumerate(joints):
            hm[b, 0, 0, 0, 0] = float(people[0][0, 0])
        return hm


class _PanopticModel:
    """Returns, per frame, the frame's ground truth itself (passed through the stand-in inputs) as P slots."""

    def __init__(self, cfg, gt_by_marker):
        self.P, self.J, self.gt, self.calls = int(cfg.CAPTURE_SPEC.MAX_PEOPLE), int(cfg.DATASET.NUM_JOINTS), gt_by_marker, []

    def __call__(self, backbone=None, views=None, meta=None, targets=None, input_heatmaps=None, cameras=None, resize_transform=None):
        x = views if views is not None else input_heatmaps
        self.calls.append((list(meta["seq"]), backbone, views is not None))
        fused = torch.zeros((x.shape[0], self.P, self.J, 5))
        fused[..., 3] = -1.0
        for b in range(x.shape[0]):
            marker = round(float(x[b].reshape(-1)[0]), 3)
            for n, pose in enumerate(self.gt[marker]):
                fused[b, n, :, :3] = torch.from_numpy(pose).float()
                fused[b, n, :, 3] = 0.0
                fused[b, n, :, 4] = 0.9
        return fused, None, None, x, None


def test_panoptic_validation_loop_gt_and_image_sources():
    from fvp.validate import validate_panoptic
    cfg = fcfg.preset("panoptic")
    cfg.DEVICE = "cpu"
    rng = np.random.default_rng(1)
    cam = dict(R=np.eye(3), T=np.zeros((3, 1)), fx=1.0, fy=1.0, cx=0.0, cy=0.0, k=np.zeros((3, 1)), p=np.zeros((2, 1)))
    cameras = {"seqA": [cam] * 5, "seqB": [cam] * 5}
    records, gt_by_marker = [], {}
    for i in range(5):
        people = [rng.uniform(-1000, 1000, (15, 3)) for _ in range(1 + i % 3)]
        people[0][0, 0] = 10.0 + i                                                   # marks the frame
        gt_by_marker[round(10.0 + i, 3)] = people
        records.append({"seq": "seqA" if i < 3 else "seqB", "all_image_path": ["f%d_v%d" % (i, v) for v in range(5)],
                        "joints_3d": people, "joints_3d_vis": [np.ones(15) for _ in people]})
    # 'gt' source: heat maps from the renderer
    R, M = _GtRenderer(cfg), _PanopticModel(cfg, gt_by_marker)
    out = validate_panoptic(cfg, M, cameras, records, source="gt", renderer=R, batch_size=2)
    assert R.calls == [2, 2, 1] and [c[0] for c in M.calls] == [["seqA", "seqA"], ["seqA", "seqB"], ["seqB"]]
    assert all(c[1] is None and not c[2] for c in M.calls)
    assert tuple(out["fused_poses"].shape) == (5, 10, 15, 5)
    assert out["detail"]["total_gt"] == sum(len(r["joints_3d"]) for r in records) and out["detail"]["mpjpe"] < 1e-3
    assert out["metric"] == pytest.approx(1.0, abs=1e-4) and "ap@25" in out["msg"]              # predictions == ground truth
    # 'image' source: views loaded per frame and handed to model(backbone=..., views=...)
    loaded = []

    def fake_views(paths, color_rgb):
        loaded.append((tuple(paths), color_rgb))
        i = int(paths[0][1])
        v = np.zeros((5, 3, 8, 8), np.float32)
        v[0, 0, 0, 0] = 10.0 + i
        return v
    backbone = object()
    M2 = _PanopticModel(cfg, gt_by_marker)
    out2 = validate_panoptic(cfg, M2, cameras, records, source="image", backbone=backbone, batch_size=4, load_views=fake_views)
    assert [len(c[0]) for c in M2.calls] == [4, 1] and all(c[1] is backbone and c[2] for c in M2.calls)
    assert [p[0][0] for p in loaded] == ["f%d_v0" % i for i in range(5)] and all(p[1] == bool(cfg.DATASET.COLOR_RGB) for p in loaded)
    assert torch.equal(out2["fused_poses"], out["fused_poses"])
    with pytest.raises(ValueError):
        validate_panoptic(cfg, M, cameras, records, source="pred", renderer=R)
    with pytest.raises(ValueError):
        validate_panoptic(cfg, M, cameras, records, source="image")
