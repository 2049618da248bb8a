#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference/lib) on deterministic synthetic inputs, and pin the oracle against it.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/<case>.npz + MANIFEST.json

For every case the script
  1. builds the reference model (``models.faster_voxelpose.get(cfg)``, DEVICE='cpu', eval) and loads
     the synthetic state_dict of ``fvp.synth.make_weights`` (strict),
  2. runs the reference forward with forward-hooks on its sub-modules to tap every stage,
  3. runs ``oracle.fvp_oracle.forward`` on the same inputs and asserts BIT-EXACT equality of every
     tap and output (the oracle pin), and asserts the explicit-rounding projection chain
     (``project_chain_np``) reproduces the reference's cached HDN sample grid bit-for-bit,
  4. stores inputs (heatmaps as uint16 on the 1/4096 lattice, calibration, resize transform, weight
     seed + SHA-256) and reference outputs / taps in a compressed npz.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
sys.path.insert(0, ROOT)

from fvp import config as fcfg, synth  # noqa: E402
from oracle import fvp_oracle as O     # noqa: E402


def _install_reference():
    """easydict shim + reference lib on sys.path (the reference needs nothing else for lib/models)."""
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        __setattr__ = __setitem__

    m = types.ModuleType("easydict")
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m
    # the reference's lib/ must shadow nothing of ours: we only import `models`, `utils`, `core` from it
    sys.path.insert(0, os.path.join(REF, "lib"))


def _load_calibration(name: str):
    if name == "panoptic":
        with open(os.path.join(REF, "demo/calibration.json")) as f:
            cams = json.load(f)["customized_sequence"]
    else:
        fn = {"campus": "data/Campus/calibration_campus.json", "shelf": "data/Shelf/calibration_shelf.json"}[name]
        with open(os.path.join(REF, fn)) as f:
            d = json.load(f)
        cams = [d[str(i)] for i in range(len(d))]
    out = []
    for c in cams:
        out.append({k: (np.array(v, dtype=np.float64) if isinstance(v, list) else float(v)) for k, v in c.items()})
    return out


CASES = {
    # name: preset, calibration, people per frame, min_score override (None = keep), seed
    "panoptic_b2": dict(preset="panoptic", calib="panoptic", people=[10, 4], min_score=None, seed=11),
    "panoptic_mixed": dict(preset="panoptic", calib="panoptic", people=[6], min_score="median", seed=12),
    "panoptic_none_valid": dict(preset="panoptic", calib="panoptic", people=[3], min_score=1e9, seed=13),
    "panoptic_256x192": dict(preset="panoptic_256x192", calib="panoptic", people=[10], min_score=None, seed=14),
    "campus_b1": dict(preset="campus", calib="campus", people=[3], min_score=None, seed=15),
    "shelf_crowd": dict(preset="shelf", calib="shelf", people=[10], min_score=None, seed=16),
    # BASELINE configs[0] / SURVEY.md 8d "Config 1" on REAL detections: frame 400 of the shipped
    # pred_{campus,shelf}_maskrcnn_hrnet_coco.pkl (3/2/2 and 4/5/2/2/2 detected people per view), heat maps rendered by
    # the reference's own JointsDataset.__getitem__ ('pred' source, sigma from the config), real calibration files
    "campus_frame400": dict(preset="campus", calib="campus", people=[0], min_score=None, seed=17, frame=400),
    "shelf_frame400": dict(preset="shelf", calib="shelf", people=[0], min_score=None, seed=18, frame=400),
}
_PRED_FILES = {"campus": "data/Campus/pred_campus_maskrcnn_hrnet_coco.pkl", "shelf": "data/Shelf/pred_shelf_maskrcnn_hrnet_coco.pkl"}


def real_frame_heatmaps(cfg, calib: str, frame: int):
    """[V,J,H,W] float32: the reference's JointsDataset.__getitem__ ('pred' heat-map source, JointsDataset.py:144-154,
    271-337) on the detections of one frame of the shipped file, built into a record with the expression of
    Campus._get_db (campus.py:92-97).  Returns the maps rounded to the 1/4096 lattice the goldens store, and the record's
    detections (before the in-place resize the reference applies to them)."""
    import copy
    import importlib.util
    import pickle
    spec = importlib.util.spec_from_file_location("ref_JointsDataset", os.path.join(REF, "lib/dataset/JointsDataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(REF, _PRED_FILES[calib]), "rb") as f:
        pred2d = pickle.load(f)
    V = int(cfg.DATASET.CAMERA_NUM)
    all_preds = []
    for k in range(V):
        preds = pred2d["{}_{}".format(k, frame)]
        all_preds.append([np.array(p["pred"]) for p in preds])
    src = cfg.DATASET.TEST_HEATMAP_SRC
    cfg.DATASET.TEST_HEATMAP_SRC = "pred"
    try:
        ds = mod.JointsDataset(cfg, is_train=False)
        ds.db = [{"pred_pose2d": copy.deepcopy(all_preds), "target": 0, "meta": {"seq": calib}, "image": ""}]
        _, _, _, hm = ds[0]
    finally:
        cfg.DATASET.TEST_HEATMAP_SRC = src
    hm = hm.numpy()
    assert hm.dtype == np.float32 and hm.shape[0] == V
    q = np.floor(hm.astype(np.float64) * 4096 + 0.5)
    return (q / 4096).astype(np.float32), all_preds, hm


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def weights_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


def bit_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    a, b = a.detach().contiguous(), b.detach().contiguous()
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return bool(torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                            b.view(torch.int32) if b.dtype == torch.float32 else b))


def run_case(name: str, spec: dict, outdir: str, ref_models, ref_tf) -> dict:
    cfg = fcfg.preset(spec["preset"])
    cfg.DEVICE = "cpu"
    J, V = int(cfg.DATASET.NUM_JOINTS), int(cfg.DATASET.CAMERA_NUM)
    cams = _load_calibration(spec["calib"])
    assert len(cams) == V
    # resize transform exactly as JointsDataset._get_resize_transform (JointsDataset.py:51-56)
    ori, img = cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE
    c = np.array([ori[0] / 2.0, ori[1] / 2.0])
    s = ref_tf.get_scale((ori[0], ori[1]), img)
    A_ref = ref_tf.get_affine_transform(c, s, 0, img)
    A_closed = synth.resize_transform(ori, img)
    resize = torch.as_tensor(A_ref, dtype=torch.float)
    resize_closed_equal = bool(np.array_equal(A_ref.astype(np.float32), A_closed.astype(np.float32)))

    wseed = 1000 + spec["seed"]
    sd_np = synth.make_weights(J, seed=wseed)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}

    B = len(spec["people"])
    sigma = float(cfg.NETWORK.SIGMA)
    hms, skels = [], []
    real = {}
    for b, npeople in enumerate(spec["people"]):
        if spec.get("frame") is not None:            # real detections rendered by the reference itself
            hm_q, preds, hm_exact = real_frame_heatmaps(cfg, spec["calib"], spec["frame"])
            skels.append(np.zeros((0, J, 3)))
            hms.append(hm_q)
            nmax = max(len(v) for v in preds)
            pp = np.zeros((V, nmax, J, 3))
            for v in range(V):
                pp[v, :len(preds[v])] = np.array(preds[v])
            nz = np.flatnonzero(hm_exact)
            real = {"frame": np.array(spec["frame"]), "preds": pp, "preds_per_view": np.array([len(v) for v in preds], np.int32),
                    "rendered_nz_index": nz.astype(np.int64), "rendered_nz_value": hm_exact.ravel()[nz],
                    "rendered_max_quant_error": np.array(float(np.abs(hm_q - hm_exact).max()))}
            continue
        sk = synth.make_skeletons(cfg, npeople, seed=spec["seed"] * 100 + b)
        skels.append(sk)
        hms.append(synth.render_heatmaps(cfg, cams, sk, sigma=sigma))
    heatmaps = torch.from_numpy(np.stack(hms))                      # [B,V,J,H,W]
    seqs = ["seq_%s" % spec["calib"]] * B
    cameras = {seqs[0]: cams}

    # min-score override
    ms = spec["min_score"]
    if ms == "median":
        probe = O.forward(cfg, sd, heatmaps, seqs, cameras, resize, taps=True)
        conf = probe["hdn_centers"][:, :, 4].reshape(-1)
        ms = float(conf.sort()[0][conf.numel() // 2 - 1] + conf.sort()[0][conf.numel() // 2]) / 2.0
    if ms is None:
        ms = -1.0e30   # all P proposals valid
    cfg.CAPTURE_SPEC.MIN_SCORE = float(ms)

    # ---- reference ----------------------------------------------------------------------------
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref_models.faster_voxelpose.get(cfg).eval()
    missing = model.load_state_dict(sd, strict=True)
    taps = {}

    def hook(tag):
        def fn(mod, inp, out):
            taps.setdefault(tag, []).append(out)
        return fn

    model.pose_net.project_layer.register_forward_hook(hook("hdn_cubes"))
    model.pose_net.center_net.register_forward_hook(hook("center_net"))
    model.pose_net.c2c_net.register_forward_hook(hook("c2c"))
    model.joint_net.project_layer.register_forward_hook(hook("jln_project"))
    model.joint_net.conv_net.register_forward_hook(hook("p2p"))
    model.joint_net.soft_argmax_layer.register_forward_hook(hook("softargmax"))
    model.joint_net.weight_net.register_forward_hook(hook("weightnet"))
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        fused_r, plane_r, centers_r, _, loss = model(meta={"seq": seqs}, input_heatmaps=heatmaps,
                                                     cameras=cameras, resize_transform=resize)
    assert loss is None

    # ---- oracle -------------------------------------------------------------------------------
    with torch.no_grad():
        o = O.forward(cfg, sd, heatmaps, seqs, cameras, resize, taps=True)

    checks = {}
    checks["fused_poses"] = bit_equal(o["fused_poses"], fused_r)
    checks["plane_poses"] = bit_equal(o["plane_poses"], plane_r)
    checks["proposal_centers"] = bit_equal(o["proposal_centers"], centers_r)
    cubes_r = taps["hdn_cubes"][0]
    checks["hdn_plane"] = bit_equal(o["hdn"]["plane"], torch.max(cubes_r, dim=4)[0])
    checks["hm2d"] = bit_equal(o["hdn"]["hm2d"], taps["center_net"][0][0])
    checks["size"] = bit_equal(o["hdn"]["size"], taps["center_net"][0][1])
    checks["hm1d"] = bit_equal(o["hdn"]["hm1d"].reshape(-1), taps["c2c"][0].reshape(-1))
    ji = 0
    for b in range(B):
        jb = o["jln"][b]
        if jb is None:
            continue
        cubes_j, offset_j = taps["jln_project"][ji]
        checks["planes_b%d" % b] = bit_equal(jb["planes"], O.three_planes(cubes_j))
        checks["offset_b%d" % b] = bit_equal(jb["crop_offset"], offset_j)
        checks["feat_b%d" % b] = bit_equal(jb["feat"], torch.stack(torch.chunk(taps["p2p"][ji], 3), dim=0))
        checks["weights_b%d" % b] = bit_equal(jb["weights"], taps["weightnet"][ji])
        checks["confs_b%d" % b] = bit_equal(jb["confs"], taps["softargmax"][ji][1])
        ji += 1

    # explicit-rounding projection chain == reference's cached HDN sample grid -> pixel coordinates
    W, H = int(cfg.DATASET.HEATMAP_SIZE[0]), int(cfg.DATASET.HEATMAP_SIZE[1])
    grid_ref = model.pose_net.project_layer.sample_grid[seqs[0]]          # [V,1,nbins,2]
    vox = O.voxel_grid(cfg.CAPTURE_SPEC.SPACE_SIZE, cfg.CAPTURE_SPEC.SPACE_CENTER,
                       cfg.CAPTURE_SPEC.VOXELS_PER_AXIS).numpy()
    chain_ok = True
    chain_bad = 0
    for v in range(V):
        ix, iy = O.project_chain_np(vox[:, 0], vox[:, 1], vox[:, 2], O.cam21_f32(cams[v]),
                                    resize.numpy().reshape(6), float(max(ori[0], ori[1])), (W, H),
                                    (float(img[0]), float(img[1])))
        g = grid_ref[v, 0].numpy()
        rx = (((g[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(W - 1)).astype(np.float32)
        ry = (((g[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(H - 1)).astype(np.float32)
        bad = int((ix.view(np.int32) != rx.view(np.int32)).sum() + (iy.view(np.int32) != ry.view(np.int32)).sum())
        chain_bad += bad
        chain_ok &= bad == 0
    checks["project_chain_np"] = chain_ok

    failed = [k for k, v in checks.items() if not v]
    print("%-22s oracle==reference: %s%s  (resize closed-form equal: %s, chain mismatches: %d)" % (
        name, "ALL BIT-EXACT" if not failed else "MISMATCH " + str(failed), "", resize_closed_equal, chain_bad))
    if failed:
        raise SystemExit("oracle is not bit-identical to the reference on case %s: %s" % (name, failed))

    # ---- store --------------------------------------------------------------------------------
    h = o["hdn"]
    store = {
        "preset": np.array(spec["preset"]), "min_score": np.array(ms, np.float64),
        "weight_seed": np.array(wseed), "weights_sha256": np.array(weights_sha(sd_np)),
        "heatmaps_u16": synth.quantise_u16(heatmaps.numpy()),
        "cameras": synth.cameras_to_array(cams), "resize": A_ref.astype(np.float64),
        "skeletons": np.concatenate([np.pad(s_, ((0, 10 - s_.shape[0]), (0, 0), (0, 0))) for s_ in skels]),
        "fused_poses": fused_r.numpy(), "plane_poses": plane_r.numpy(), "proposal_centers": centers_r.numpy(),
        "hdn_centers": h["centers"].numpy(), "hdn_plane": h["plane"].numpy(), "hm2d": h["hm2d"].numpy(),
        "size": h["size"].numpy(), "conf2d": h["conf2d"].numpy(), "flat": h["flat"].numpy(),
        "idx_z": h["idx_z"].numpy(), "conf1d": h["conf1d"].numpy(), "cols": h["cols"].numpy(),
        "hm1d": h["hm1d"].numpy(),
    }
    store.update(real)
    Kc = O.JlnConstants(cfg)
    store["axes_coarse"] = np.concatenate([O.axis_coords(cfg.CAPTURE_SPEC.SPACE_SIZE[d], cfg.CAPTURE_SPEC.SPACE_CENTER[d],
                                                         cfg.CAPTURE_SPEC.VOXELS_PER_AXIS[d]).numpy() for d in range(3)])
    store["axes_fine"] = np.concatenate([a.numpy() for a in Kc.fine_axes])
    store["axes_ind"] = np.concatenate([O.axis_coords(cfg.INDIVIDUAL_SPEC.SPACE_SIZE[d], cfg.CAPTURE_SPEC.SPACE_CENTER[d],
                                                      cfg.INDIVIDUAL_SPEC.VOXELS_PER_AXIS[d]).numpy() for d in range(3)])
    KEEP = 2  # persons per frame whose full planes / features are stored
    for b in range(B):
        jb = o["jln"][b]
        if jb is None:
            continue
        n = jb["crop_tl"].shape[0]
        keep = list(range(min(KEEP, n)))
        planes = jb["planes"].view(3, n, *jb["planes"].shape[1:])
        store.update({
            "b%d_crop_tl" % b: jb["crop_tl"].numpy(), "b%d_crop_offset" % b: jb["crop_offset"].numpy(),
            "b%d_crop_start" % b: jb["crop_start"].numpy(), "b%d_crop_end" % b: jb["crop_end"].numpy(),
            "b%d_planes_keep" % b: planes[:, keep].numpy(), "b%d_feat_keep" % b: jb["feat"][:, keep].numpy(),
            "b%d_planes_sum" % b: planes.double().sum(dim=(3, 4)).numpy(),
            "b%d_planes_max" % b: planes.amax(dim=(3, 4)).numpy(),
            "b%d_feat_sum" % b: jb["feat"].double().sum(dim=(3, 4)).numpy(),
            "b%d_pose" % b: jb["pose"].numpy(), "b%d_confs" % b: jb["confs"].numpy(),
            "b%d_weights" % b: jb["weights"].numpy(), "b%d_fused" % b: jb["fused"].numpy(),
        })
    path = os.path.join(outdir, name + ".npz")
    np.savez_compressed(path, **store)
    nvalid = (centers_r[:, :, 3] >= 0).sum(dim=1).tolist()
    return {"file": os.path.basename(path), "bytes": os.path.getsize(path), "valid_per_frame": nvalid,
            "sha256_fused": sha(fused_r.numpy()), "checks": sorted(checks), "resize_closed_form_equal": resize_closed_equal,
            "torch": torch.__version__, "threads": torch.get_num_threads()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    _install_reference()
    import models as ref_models           # reference lib/models
    import models.faster_voxelpose        # noqa: F401
    from utils import transforms as ref_tf
    os.makedirs(args.out, exist_ok=True)
    manifest = {}
    mpath = os.path.join(args.out, "MANIFEST.json")
    if os.path.exists(mpath) and args.only:
        manifest = json.load(open(mpath))
    for name, spec in CASES.items():
        if args.only and name not in args.only.split(","):
            continue
        manifest[name] = run_case(name, spec, args.out, ref_models, ref_tf)
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps({k: (v["bytes"], v["valid_per_frame"]) for k, v in manifest.items()}))


if __name__ == "__main__":
    main()
