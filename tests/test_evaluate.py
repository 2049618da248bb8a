"""Task metrics: fvp.evaluate against goldens produced by the unmodified reference's
Panoptic / Campus / Shelf ``evaluate`` (oracle/gen_golden_eval.py), plus edge cases and invariances.  CPU only."""
import os

import numpy as np
import pytest
import torch

from fvp import evaluate as E

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _panoptic():
    z = np.load(os.path.join(GOLD, "eval_panoptic.npz"))
    n = z["num_person"]
    return z, list(z["preds"]), [z["gt_joints"][f, :n[f]] for f in range(len(n))], [z["gt_vis"][f, :n[f]] for f in range(len(n))]


def _actors(z):
    """object array [actor][frame] like load_actors(): [14,3] metres, or an empty (1,0) entry for an absent actor."""
    fr = z["frame_range"]
    A = z["gt_metres"].shape[0]
    actors = np.empty((A, int(fr.max()) + 1), dtype=object)
    for a in range(A):
        for f in range(actors.shape[1]):
            actors[a, f] = np.zeros((1, 0))
        for i, f in enumerate(fr):
            if z["present"][a, i]:
                actors[a, f] = z["gt_metres"][a, i]
    return actors


def test_panoptic_metrics_equal_the_reference():
    z, preds, gj, gv = _panoptic()
    metric, msg, d = E.evaluate_panoptic(preds, gj, gv)
    assert metric == float(z["metric"]) and msg == str(z["msg"])
    assert np.array_equal(d["aps"], z["aps"]) and np.array_equal(d["recalls"], z["recalls"])
    assert d["mpjpe"] == float(z["mpjpe"]) and d["recall"] == float(z["recall"])
    # torch tensors in, as run/validate.py passes them
    assert E.evaluate_panoptic([torch.from_numpy(p) for p in preds], gj, gv)[0] == metric


@pytest.mark.parametrize("dataset", ["campus", "shelf"])
def test_pcp_metrics_equal_the_reference(dataset):
    z = np.load(os.path.join(GOLD, "eval_%s.npz" % dataset))
    metric, msg, d = E.evaluate_pcp(list(z["preds"]), _actors(z), list(z["frame_range"]), dataset)
    assert metric == float(z["metric"]) and msg == str(z["msg"])
    assert np.array_equal(d["actor_pcp"], z["actor_pcp"]) and d["recall"] == float(z["recall"])
    assert np.array_equal(np.stack(list(d["bone_pcp"].values())), z["bone_pcp"])


def test_panoptic_is_invariant_to_slot_order_and_ignores_invalid_slots():
    z, preds, gj, gv = _panoptic()
    base = E.evaluate_panoptic(preds, gj, gv)
    rng = np.random.default_rng(0)
    shuffled = [p[rng.permutation(len(p))] for p in preds]
    assert E.evaluate_panoptic(shuffled, gj, gv)[:2] == base[:2]
    junk = [p.copy() for p in preds]
    for p in junk:                                     # coordinates of invalid slots must not matter
        p[p[:, 0, 3] < 0, :, :3] = 1.0e6
    assert E.evaluate_panoptic(junk, gj, gv)[:2] == base[:2]


def test_panoptic_edge_cases():
    z, preds, gj, gv = _panoptic()
    none = [np.full_like(p, -1.0) for p in preds]
    metric, msg, d = E.evaluate_panoptic(none, gj, gv)
    assert metric == 0.0 and d["mpjpe"] == float("inf") and d["recall"] == 0.0 and d["poses"] == 0
    # perfect detections: AP = recall = 1 up to the reference's 1e-5 regularisers, MPJPE = 0
    perfect = []
    for j3, p in zip(gj, preds):
        q = np.full_like(p, -1.0)
        q[: len(j3), :, :3], q[: len(j3), :, 3], q[: len(j3), :, 4] = j3, 0.0, 0.9
        perfect.append(q)
    metric, _, d = E.evaluate_panoptic(perfect, gj, gv)
    assert abs(metric - 1.0) < 1e-4 and d["recall"] == 1.0 and d["mpjpe"] < 1e-3      # float32 rounding of the joints
    with pytest.raises(AssertionError):
        E.evaluate_panoptic(preds[:-1], gj, gv)


def test_pcp_edge_cases():
    z = np.load(os.path.join(GOLD, "eval_campus.npz"))
    actors, fr = _actors(z), list(z["frame_range"])
    none = [np.full_like(p, -1.0) for p in z["preds"]]
    metric, _, d = E.evaluate_pcp(none, actors, fr, "campus")          # Campus skips frames without predictions
    assert metric == 0.0 and d["total_gt"] == 0
    with pytest.raises(ValueError):                                    # Shelf stacks an empty list (shelf.py:179)
        E.evaluate_pcp(none, actors, fr, "shelf")
    with pytest.raises(AssertionError):
        E.evaluate_pcp(none, actors, fr, "panoptic")


def test_head_models():
    rng = np.random.default_rng(1)
    c = rng.normal(0, 300, (17, 3))
    a, b = E.coco2campus3D(c), E.coco2shelf3D(c)
    assert a.shape == b.shape == (14, 3) and np.array_equal(a[:12], b[:12]) and np.array_equal(a[:12], c[E.COCO_TO_14])
    ears, sho = (c[3] + c[4]) / 2, (c[5] + c[6]) / 2
    assert np.allclose(a[12], (ears + sho) / 2) and np.allclose(a[13], a[12] + (ears - a[12]) * 2)
    assert not np.allclose(a[12:], b[12:])                             # Shelf blends in the nose/shoulder model
