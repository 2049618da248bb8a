"""Load tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-voxelpose_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from fvp import config as fcfg, synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CASES = ["panoptic_b2", "panoptic_mixed", "panoptic_none_valid", "panoptic_256x192", "campus_b1", "shelf_crowd"]
# BASELINE configs[0] on real detections: frame 400 of the reference's shipped Campus / Shelf detection files, heat maps
# rendered by the reference's own JointsDataset (oracle/gen_golden.py::real_frame_heatmaps)
REAL_CASES = ["campus_frame400", "shelf_frame400"]


def weights_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


class Golden:
    def __init__(self, name: str):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        z = self.z
        self.cfg = fcfg.preset(str(z["preset"]))
        self.cfg.CAPTURE_SPEC.MIN_SCORE = float(z["min_score"])
        self.J = int(self.cfg.DATASET.NUM_JOINTS)
        self.P = int(self.cfg.CAPTURE_SPEC.MAX_PEOPLE)
        self.heatmaps = synth.dequantise_u16(z["heatmaps_u16"])          # [B,V,J,H,W]
        self.B = self.heatmaps.shape[0]
        self.cams = synth.cameras_from_array(z["cameras"])
        self.resize = z["resize"]
        self.weights = synth.make_weights(self.J, seed=int(z["weight_seed"]))
        assert weights_sha(self.weights) == str(z["weights_sha256"]), "weight generator drifted from the golden"
        X, Y, Zc = [int(v) for v in self.cfg.CAPTURE_SPEC.VOXELS_PER_AXIS]
        self.axes = (z["axes_coarse"], z["axes_fine"], z["axes_ind"])
        self.seq = "seq0"
        self.cameras = {self.seq: self.cams}

    def __getitem__(self, k):
        return self.z[k]

    def has(self, k):
        return k in self.z.files

    # ---- real-detection cases only ------------------------------------------------------------------------------
    def real_preds(self):
        """[view][person] -> [J,3] detections of the frame (original-image pixels), as db_rec['pred_pose2d']."""
        n = self.z["preds_per_view"]
        return [[self.z["preds"][v, i] for i in range(int(n[v]))] for v in range(len(n))]

    def rendered(self) -> np.ndarray:
        """The maps the reference rendered from real_preds(), exact float32 ([V,J,H,W]; heatmaps holds them rounded to
        the 1/4096 lattice)."""
        out = np.zeros(int(np.prod(self.heatmaps.shape[1:])), np.float32)
        out[self.z["rendered_nz_index"]] = self.z["rendered_nz_value"]
        return out.reshape(self.heatmaps.shape[1:])


HEATMAP_CASES = ["heatmaps_panoptic_256x192_crowd", "heatmaps_panoptic_border", "heatmaps_campus_tiny",
                 "heatmaps_shelf_huge", "heatmaps_shelf_crowd"]


class HeatmapGolden:
    """tests/golden/heatmaps_*.npz (oracle/gen_golden_heatmaps.py: the reference's JointsDataset.__getitem__ on synthetic
    'pred' / 'gt' records).  The rendered maps are stored sparsely (non-zero values)."""

    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.cfg = fcfg.preset(str(z["preset"]))
        self.shape = tuple(int(v) for v in z["shape"])                  # [V,J,H,W]
        self.resize = z["resize"]
        self.num_people = z["num_people"]
        self.preds = [[z["preds"][v, n] for n in range(int(self.num_people[v]))] for v in range(self.shape[0])]
        self.joints_3d = [a for a in z["joints_3d"]]
        self.joints_3d_vis = [a for a in z["joints_3d_vis"]]
        self.cams = synth.cameras_from_array(z["cameras"])

    def dense(self, which: str) -> np.ndarray:
        out = np.zeros(int(np.prod(self.shape)), np.float32)
        out[self.z[which + "_nz_index"]] = self.z[which + "_nz_value"]
        return out.reshape(self.shape)
