"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py asserted bit-equality on the generating host; another host's oneDNN may round
convolutions differently, hence small tolerances instead of bit-equality here)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_util import CASES
from oracle import fvp_oracle as O


def _sd(g):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in g.weights.items()}


@pytest.mark.parametrize("name", ["campus_b1", "panoptic_mixed", "panoptic_none_valid", "shelf_crowd", "campus_frame400"])
def test_oracle_forward_matches_reference_golden(golden, name):
    g = golden(name)
    with torch.no_grad():
        out = O.forward(g.cfg, _sd(g), torch.from_numpy(g.heatmaps), [g.seq] * g.B, g.cameras,
                        torch.as_tensor(g.resize, dtype=torch.float))
    assert np.array_equal(out["hdn"]["flat"].numpy(), g["flat"])                       # proposal cells: exact
    assert np.array_equal(out["hdn"]["idx_z"].numpy(), g["idx_z"])
    np.testing.assert_allclose(out["hdn"]["plane"].numpy(), g["hdn_plane"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(out["proposal_centers"].numpy(), g["proposal_centers"], atol=1e-5, rtol=0)
    assert np.array_equal(out["fused_poses"].numpy()[..., 3], g["fused_poses"][..., 3])  # validity flags
    np.testing.assert_allclose(out["fused_poses"].numpy(), g["fused_poses"], atol=2e-2, rtol=0)   # mm
    np.testing.assert_allclose(out["plane_poses"].numpy(), g["plane_poses"], atol=2e-2, rtol=0)


@pytest.mark.parametrize("name", ["panoptic_b2", "campus_b1"])
def test_explicit_projection_chain_is_bit_identical_to_torch_chain(golden, name):
    """project_chain_np (the CUDA kernels' specification) == the reference expression chain, bit for bit."""
    g = golden(name)
    cfg = g.cfg
    vox = O.voxel_grid(cfg.CAPTURE_SPEC.SPACE_SIZE, cfg.CAPTURE_SPEC.SPACE_CENTER, cfg.CAPTURE_SPEC.VOXELS_PER_AXIS)[::5]
    W, H = [int(v) for v in cfg.DATASET.HEATMAP_SIZE]
    resize = torch.as_tensor(g.resize, dtype=torch.float)
    ori, img = cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE
    for cam in g.cams:
        grid = O.sample_grid(vox, cam, resize, ori, img, (W, H)).numpy()
        rx = (((grid[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(W - 1)).astype(np.float32)
        ry = (((grid[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(H - 1)).astype(np.float32)
        v = vox.numpy()
        with np.errstate(all="ignore"):
            ix, iy = O.project_chain_np(v[:, 0], v[:, 1], v[:, 2], O.cam21_f32(cam), resize.numpy().reshape(6),
                                        float(max(ori)), (W, H), (float(img[0]), float(img[1])))
        assert np.array_equal(ix.view(np.int32), rx.view(np.int32))
        assert np.array_equal(iy.view(np.int32), ry.view(np.int32))


def test_explicit_bilinear_matches_grid_sample():
    rng = np.random.default_rng(0)
    hm = rng.random((3, 17, 23), dtype=np.float32)
    ix = rng.uniform(-2.5, 24.5, 500).astype(np.float32)
    iy = rng.uniform(-2.5, 18.5, 500).astype(np.float32)
    gx = ix / np.float32(22) * 2 - 1
    gy = iy / np.float32(16) * 2 - 1
    grid = torch.from_numpy(np.stack([gx, gy], 1)).view(1, 1, -1, 2)
    ref = F.grid_sample(torch.from_numpy(hm)[None], grid, align_corners=True)[0, :, 0].numpy()
    # use the coordinates grid_sample itself derives so only the tap arithmetic is compared
    ix2 = (((gx + np.float32(1)) / np.float32(2)) * np.float32(22)).astype(np.float32)
    iy2 = (((gy + np.float32(1)) / np.float32(2)) * np.float32(16)).astype(np.float32)
    got = O.bilinear_zeros_np(hm, ix2, iy2)
    np.testing.assert_allclose(got, ref, atol=2e-7, rtol=0)


def test_nms_topk_quirks():
    """flat -> (x, y) uses the X extent for both div and mod (core/proposal.py:16-17); suppressed cells are zero."""
    hm = torch.zeros(1, 1, 8, 8)
    hm[0, 0, 2, 3] = 0.9
    hm[0, 0, 2, 4] = 0.8       # neighbour of a larger value: suppressed
    hm[0, 0, 6, 1] = 0.7
    vals, idx, flat = O.nms_topk(hm, 3)
    assert flat[0, :2].tolist() == [2 * 8 + 3, 6 * 8 + 1]
    assert idx[0, 0].tolist() == [2, 3] and idx[0, 1].tolist() == [6, 1]
    assert vals[0, 2].item() == 0.0


def test_crop_params_edge_cases(golden):
    """border crops clip to the fine grid; a negative bbox gives an empty crop (project_individual.py:110-125)."""
    g = golden("panoptic_mixed")
    K = O.JlnConstants(g.cfg)
    centers = torch.tensor([[-3990.0, -4400.0, -150.0, 0, 1, 0.85, 0.85],      # corner of the space
                            [0.0, -500.0, 800.0, 0, 1, -0.5, 0.9]])             # bbox < 0 -> mask 47 > 32 -> empty
    crop = O.jln_crop_params(K, centers)
    assert (crop["start"][0] >= 0).all() and (crop["start"][0] == 0).any()
    assert (crop["start"][1] >= crop["end"][1]).any()
    assert (crop["end"] <= K.fine).all()


# ---- N1: heat-map renderer oracle vs the reference's JointsDataset output ----------------------------------------
from golden_util import HEATMAP_CASES, HeatmapGolden  # noqa: E402
from oracle import heatmap_oracle as HO               # noqa: E402


@pytest.mark.parametrize("name", HEATMAP_CASES)
def test_heatmap_oracle_is_bit_identical_to_reference_golden(name):
    """generate_input_heatmap restated (oracle/heatmap_oracle.py) vs the maps the unmodified reference rendered:
    bit-exact for the 'pred' and the 'gt' source (float64 arithmetic, one rounding to float32)."""
    g = HeatmapGolden(name)
    ds = g.cfg.DATASET
    pred = HO.pred_heatmaps(g.preds, g.resize, ds.HEATMAP_SIZE, ds.IMAGE_SIZE, g.cfg.NETWORK.SIGMA)
    assert np.array_equal(pred.view(np.int32), g.dense("pred").view(np.int32))
    gt = HO.gt_heatmaps(g.joints_3d, g.joints_3d_vis, g.cams, g.resize, ds.ORI_IMAGE_SIZE, ds.IMAGE_SIZE, ds.HEATMAP_SIZE,
                        g.cfg.NETWORK.SIGMA)
    assert np.array_equal(gt.view(np.int32), g.dense("gt").view(np.int32))


def test_heatmap_oracle_edge_cases():
    """int() truncation toward zero at negative coordinates, the 'completely outside' skip, a masked joint."""
    W, H = 40, 30
    pose = np.array([[-9.0, 12.0], [170.0, 60.0], [80.0, 500.0]])         # image px, stride 4: mu = (-2,3), (42,15), (20,125)
    hm = HO.render_input_heatmap([pose], None, (W, H), (160, 120), 3)
    assert hm.shape == (3, H, W) and hm[0].max() > 0 and hm[2].max() == 0   # joint 2 lies below the map: skipped
    # person extent 122 px -> sigma 5.39 -> half width 16.2: joint 1 (mu_x = 42) starts at column int(42 - 16.2) = 25
    assert hm[1, :, 25:].max() > 0 and hm[1, :, :25].max() == 0             # patch cut by the right border
    masked = HO.render_input_heatmap([pose], [np.array([0, 1, 1])], (W, H), (160, 120), 3)
    assert masked[0].max() == 0 and np.array_equal(masked[1], hm[1])


@pytest.mark.parametrize("name,views", [("campus_frame400", 3), ("shelf_frame400", 5)])
def test_real_detections_frame400_chain(golden, name, views):
    """BASELINE configs[0] on real data: the detection file -> fvp.datasets.frame_preds -> the heat-map oracle reproduces,
    bit for bit, the maps the reference's JointsDataset rendered from frame 400 of its shipped detection file; the
    golden's model input is those maps on the 1/4096 lattice (off by at most half a step)."""
    import os
    from golden_util import GOLDEN_DIR
    from fvp import datasets as D
    g = golden(name)
    pred = D.load_pred_pose2d(os.path.join(GOLDEN_DIR, "pred_pose2d_frame400.pkl"))[name.split("_")[0]]
    frame = D.frame_preds(pred, int(g["frame"]), views)
    stored = g.real_preds()
    assert [len(v) for v in frame] == [len(v) for v in stored]
    assert all(np.array_equal(a, b) for va, vb in zip(frame, stored) for a, b in zip(va, vb))
    ds = g.cfg.DATASET
    hm = HO.pred_heatmaps(frame, g.resize, ds.HEATMAP_SIZE, ds.IMAGE_SIZE, g.cfg.NETWORK.SIGMA)
    ref = g.rendered()
    assert np.array_equal(hm.view(np.int32), ref.view(np.int32))
    assert hm.max() == 1.0 and float(np.abs(g.heatmaps[0] - ref).max()) <= 0.5 / 4096 + 1e-9
