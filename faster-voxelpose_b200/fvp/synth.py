"""Deterministic synthetic inputs: calibrations, skeletons, heatmaps and network weights.

Everything is derived from ``numpy.random.PCG64`` *uniform* draws and correctly-rounded IEEE
arithmetic only, so a (seed, config) pair regenerates bit-identical tensors on any host (no
``exp``/``log`` of libm enters a *weight*; heatmaps go through ``exp`` but are quantised to
multiples of 1/4096, which are exact in fp32).

Recipes follow SURVEY.md §8(d) "Synthetic input distributions" and §7.3 H3 ("oracle conditioning"):
the reference's own ``_initialize_weights`` (N(0, 0.001), ``cnns_2d.py:137-144``) gives degenerate
outputs, so weights use fan-in uniform bounds, randomised BatchNorm statistics and a bbox head
biased to ~0.85 (-> bbox masks of 4-5 fine voxels, ``project_individual.py:114``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

from . import netspec


class Rng:
    """Uniform-only generator (bit-stable across platforms)."""

    def __init__(self, seed: int):
        self._g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, shape, lo: float, hi: float) -> np.ndarray:
        u = self._g.random(size=shape)  # float64 in [0,1), integer-derived
        return (lo + (hi - lo) * u).astype(np.float32)

    def uniform64(self, shape, lo: float, hi: float) -> np.ndarray:
        return lo + (hi - lo) * self._g.random(size=shape)


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
def make_backbone_weights(layers, seed: int = 0) -> Dict[str, np.ndarray]:
    """state_dict (numpy) of the PoseResNet backbone described by ``fvp.backbone_spec`` layers, keyed like the reference's.

    conv: U(-b, b) with b = sqrt(3 / fan_in) (unit gain, so activations neither die nor explode over 50 layers);
    BatchNorm: gamma U(0.7,1.3) - U(0.15,0.35) on the last BN of a residual block so the residual sum stays O(1) -
    beta / running_mean U(-0.1,0.1), running_var U(0.6,1.4); ``num_batches_tracked`` = 0."""
    rng = Rng(seed)
    sd: Dict[str, np.ndarray] = {}
    for c in layers:
        fan_in = (c.cout if c.transposed else c.cin) * c.k * c.k
        b = math.sqrt(3.0 / fan_in)
        shape = (c.cin, c.cout, c.k, c.k) if c.transposed else (c.cout, c.cin, c.k, c.k)
        sd[c.key + ".weight"] = rng.uniform(shape, -b, b)
        if c.bias:
            sd[c.key + ".bias"] = rng.uniform((c.cout,), -0.1, 0.1)
        if c.bn:
            lo, hi = (0.15, 0.35) if c.role == "last" else (0.7, 1.3)
            sd[c.bn + ".weight"] = rng.uniform((c.cout,), lo, hi)
            sd[c.bn + ".bias"] = rng.uniform((c.cout,), -0.1, 0.1)
            sd[c.bn + ".running_mean"] = rng.uniform((c.cout,), -0.1, 0.1)
            sd[c.bn + ".running_var"] = rng.uniform((c.cout,), 0.6, 1.4)
            sd[c.bn + ".num_batches_tracked"] = np.zeros((), np.int64)
    return sd


def make_weights(J: int, seed: int = 0, feat: int = 32, hidden: int = 64,
                 p2p_out_gain: float = 0.25, hm_head_bias: float = 0.5,
                 c2c_head_bias: float = 0.5, c2c_in_gain: float = 4.0, wn_in_gain: float = 20.0,
                 wn_out_gain: float = 10.0) -> Dict[str, np.ndarray]:
    """state_dict (numpy) for the whole FasterVoxelPoseNet, keyed like the reference's.

    * conv / linear: U(-b, b), b = 1/sqrt(fan_in) (PyTorch's default ``reset_parameters`` bound)
    * BatchNorm: gamma U(0.5,1.5), beta U(-0.17,0.17), running_mean U(-0.17,0.17), running_var U(0.5,1.5)
    * ``output_size.2`` (bbox head): weight x0.05, bias 0.85  
    * heat-map heads get a positive bias so confidences are positive; ``output_layer`` of P2PNet is
      scaled by ``p2p_out_gain`` so that softmax(100*x) keeps a few dozen pixels of support.
    """
    rng = Rng(seed)
    sd: Dict[str, np.ndarray] = {}
    for c in netspec.all_layers(J, feat, hidden):
        fan_in = c.cin * (c.k ** c.ndim if c.ndim else 1)
        if c.transposed:  # torch computes fan_in of ConvTranspose from dim 1 (= cout) * k^nd
            fan_in = c.cout * (c.k ** c.ndim)
        b = 1.0 / math.sqrt(fan_in)
        w = rng.uniform(netspec.weight_shape(c), -b, b)
        bias = rng.uniform((c.cout,), -b, b)
        if c.key.endswith("output_size.2"):
            w = (w * np.float32(0.05)).astype(np.float32)
            bias = np.full((c.cout,), 0.85, np.float32)
        if c.key.endswith("output_hm.2"):
            bias = np.full((c.cout,), hm_head_bias, np.float32)
        if c.key.endswith("c2c_net.output_hm"):
            bias = np.full((c.cout,), c2c_head_bias, np.float32)
        if c.key.endswith("c2c_net.front_layers.0.block.0"):
            w = (w * np.float32(c2c_in_gain)).astype(np.float32)       # z-argmax depends on the column
        if c.key.endswith("weight_net.heatmap_feature_net.0"):
            w = (w * np.float32(wn_in_gain)).astype(np.float32)        # fusion weights spread over ~0.2-0.45
        if c.key.endswith("weight_net.output.2"):
            w = (w * np.float32(wn_out_gain)).astype(np.float32)
        if c.key.endswith("output_layer"):
            w = (w * np.float32(p2p_out_gain)).astype(np.float32)
            bias = (bias * np.float32(p2p_out_gain)).astype(np.float32)
        sd[c.key + ".weight"] = w
        sd[c.key + ".bias"] = bias
        if c.bn:
            sd[c.bn + ".weight"] = rng.uniform((c.cout,), 0.5, 1.5)
            sd[c.bn + ".bias"] = rng.uniform((c.cout,), -0.17, 0.17)
            sd[c.bn + ".running_mean"] = rng.uniform((c.cout,), -0.17, 0.17)
            sd[c.bn + ".running_var"] = rng.uniform((c.cout,), 0.5, 1.5)
            sd[c.bn + ".num_batches_tracked"] = np.array(1, np.int64)
    return sd


# ----------------------------------------------------------------------------------------------
# calibration
# ----------------------------------------------------------------------------------------------
def resize_transform(ori_size: Sequence[float], image_size: Sequence[float]) -> np.ndarray:
    """2x3 affine original-pixel -> network-input-pixel, aspect-preserving with centred padding.

    Closed form of what the reference obtains via ``get_scale`` + ``get_affine_transform(rot=0)``
    (``lib/utils/transforms.py:15-50,81-92``, used by ``JointsDataset._get_resize_transform``
    ``lib/dataset/JointsDataset.py:51-56``): uniform scale s = dst_w / padded_src_w about the centres.
    """
    ow, oh = float(ori_size[0]), float(ori_size[1])
    iw, ih = float(image_size[0]), float(image_size[1])
    if ow / iw < oh / ih:
        w_pad = oh / ih * iw
    else:
        w_pad = ow
    s = iw / w_pad
    return np.array([[s, 0.0, iw * 0.5 - s * ow * 0.5],
                     [0.0, s, ih * 0.5 - s * oh * 0.5]], dtype=np.float64)


def _look_at(cam_pos: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Rotation world->camera with +z forward, +y down (image convention)."""
    fwd = target - cam_pos
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd])


def ring_cameras(num_views: int, space_center: Sequence[float], radius: float = 5500.0,
                 height: float = 2600.0, look_z: float = 900.0, f: float = 1400.0,
                 cx: float = 960.0, cy: float = 540.0,
                 k=(-0.28, 0.18, -0.04), p=(0.0, 0.0), phase: float = 0.1) -> List[dict]:
    """Synthetic ring calibration; ``T`` is the camera position,
    reference convention xcam = R (x - T) (``lib/utils/cameras.py:43``)."""
    cams = []
    c = np.asarray(space_center, np.float64)
    for i in range(num_views):
        a = 2.0 * math.pi * i / num_views + phase
        pos = np.array([c[0] + radius * math.cos(a), c[1] + radius * math.sin(a), height])
        R = _look_at(pos, np.array([c[0], c[1], look_z]))
        cams.append({
            "R": R, "T": pos.reshape(3, 1), "fx": f, "fy": f, "cx": cx, "cy": cy,
            "k": np.asarray(k, np.float64).reshape(3, 1), "p": np.asarray(p, np.float64).reshape(2, 1),
        })
    return cams


def cameras_to_array(cams: Sequence[dict]) -> np.ndarray:
    """[V,21] float64: R(9, row-major) T(3) fx fy cx cy k(3) p(2)."""
    out = np.zeros((len(cams), 21), np.float64)
    for i, c in enumerate(cams):
        out[i, 0:9] = np.asarray(c["R"], np.float64).reshape(9)
        out[i, 9:12] = np.asarray(c["T"], np.float64).reshape(3)
        out[i, 12:16] = [c["fx"], c["fy"], c["cx"], c["cy"]]
        out[i, 16:19] = np.asarray(c["k"], np.float64).reshape(3)
        out[i, 19:21] = np.asarray(c["p"], np.float64).reshape(2)
    return out


def cameras_from_array(arr: np.ndarray) -> List[dict]:
    cams = []
    for row in np.asarray(arr, np.float64):
        cams.append({"R": row[0:9].reshape(3, 3).copy(), "T": row[9:12].reshape(3, 1).copy(),
                     "fx": float(row[12]), "fy": float(row[13]), "cx": float(row[14]), "cy": float(row[15]),
                     "k": row[16:19].reshape(3, 1).copy(), "p": row[19:21].reshape(2, 1).copy()})
    return cams


def project_f64(x: np.ndarray, cam: dict) -> np.ndarray:
    """Pinhole + radial/tangential projection in float64 ([n,3] -> [n,2] original pixels);
    same formulas as ``lib/utils/cameras.py:58-84`` (numpy twin used by the datasets)."""
    R = np.asarray(cam["R"], np.float64)
    T = np.asarray(cam["T"], np.float64).reshape(3, 1)
    q = R @ (x.T - T)
    y = q[:2] / (q[2] + 1e-5)
    r = (y * y).sum(0)
    k = np.asarray(cam["k"], np.float64).reshape(3)
    p = np.asarray(cam["p"], np.float64).reshape(2)
    d = 1 + k[0] * r + k[1] * r * r + k[2] * r * r * r
    u = y[0] * d + 2 * p[0] * y[0] * y[1] + p[1] * (r + 2 * y[0] * y[0])
    v = y[1] * d + 2 * p[1] * y[0] * y[1] + p[0] * (r + 2 * y[1] * y[1])
    return np.stack([cam["fx"] * u + cam["cx"], cam["fy"] * v + cam["cy"]], 1)


# ----------------------------------------------------------------------------------------------
# people and heatmaps
# ----------------------------------------------------------------------------------------------
def make_skeletons(cfg, num_people: int, seed: int, min_sep: float = 800.0) -> np.ndarray:
    """[P,J,3] world mm. Roots uniform in the central 50 % of the capture space."""
    rng = Rng(seed)
    J = int(cfg.DATASET.NUM_JOINTS)
    size = np.asarray(cfg.CAPTURE_SPEC.SPACE_SIZE, np.float64)
    ctr = np.asarray(cfg.CAPTURE_SPEC.SPACE_CENTER, np.float64)
    roots: List[np.ndarray] = []
    guard = 0
    while len(roots) < num_people:
        guard += 1
        if guard > 100000:
            raise RuntimeError("cannot place %d people %g mm apart" % (num_people, min_sep))
        xy = rng.uniform64((2,), -0.25, 0.25) * size[:2] + ctr[:2]
        if all(np.hypot(*(xy - r)) >= min_sep for r in roots):
            roots.append(xy)
    out = np.zeros((num_people, J, 3), np.float64)
    for pi, xy in enumerate(roots):
        off = rng.uniform64((J, 3), -0.5, 0.5) * np.array([500.0, 500.0, 1600.0])
        out[pi] = np.array([xy[0], xy[1], 900.0]) + off
    return out


def render_heatmaps(cfg, cams: Sequence[dict], skeletons: np.ndarray, sigma: float = 3.0,
                    quant: int = 4096) -> np.ndarray:
    """[V,J,H,W] fp32 Gaussian-blob heatmaps, per-joint max over persons (the recipe of
    ``JointsDataset.generate_input_heatmap`` ``lib/dataset/JointsDataset.py:271-337`` without its
    noise model), quantised to multiples of 1/``quant``."""
    W, H = int(cfg.DATASET.HEATMAP_SIZE[0]), int(cfg.DATASET.HEATMAP_SIZE[1])
    iw = float(cfg.DATASET.IMAGE_SIZE[0])
    A = resize_transform(cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE)
    stride = iw / W
    P, J, _ = skeletons.shape
    xs = np.arange(W, dtype=np.float64)[None, :]
    ys = np.arange(H, dtype=np.float64)[:, None]
    out = np.zeros((len(cams), J, H, W), np.float64)
    for v, cam in enumerate(cams):
        px = project_f64(skeletons.reshape(-1, 3), cam)          # original pixels
        px = (A[:, :2] @ px.T + A[:, 2:3]).T / stride            # heatmap pixels
        px = px.reshape(P, J, 2)
        for p in range(P):
            for j in range(J):
                mx, my = px[p, j]
                if not (-3 * sigma <= mx < W + 3 * sigma and -3 * sigma <= my < H + 3 * sigma):
                    continue
                g = np.exp(-((xs - mx) ** 2 + (ys - my) ** 2) / (2.0 * sigma * sigma))
                np.maximum(out[v, j], g, out=out[v, j])
    q = np.floor(out * quant + 0.5)
    return (q / quant).astype(np.float32)


def quantise_u16(hm: np.ndarray, quant: int = 4096) -> np.ndarray:
    q = np.rint(hm.astype(np.float64) * quant)
    assert np.array_equal((q / quant).astype(np.float32), hm), "heatmap is not on the 1/%d lattice" % quant
    return q.astype(np.uint16)


def dequantise_u16(q: np.ndarray, quant: int = 4096) -> np.ndarray:
    return (q.astype(np.float64) / quant).astype(np.float32)


def random_heatmaps(shape, seed: int) -> np.ndarray:
    """Bandwidth inputs: uniform [0,1) fp32 (timing is value independent)."""
    return Rng(seed).uniform(shape, 0.0, 1.0)
