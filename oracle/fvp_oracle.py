"""CPU ORACLE for the Faster-VoxelPose inference hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product path (``faster-voxelpose_b200/``) never does and fails
loudly when its CUDA library is missing.

What it is: a functional, ``state_dict``-driven restatement in CPU PyTorch fp32 of
``FasterVoxelPoseNet.forward`` in eval mode on precomputed heatmaps.
The arithmetic primitives (``F.conv*``, ``F.batch_norm``, ``F.grid_sample``, ``topk``, ``softmax``)
are the same third-party PyTorch (2.11.0 in this image; the reference pins 1.12.0, README.md:31)
ops the reference itself calls, applied in the same order, so on one host the oracle is
bit-identical to the reference; every function cites the reference lines it follows.

Parity pin: the reference ships no tests or golden vectors, so the pin is the
reference itself, imported from /root/reference by ``oracle/gen_golden.py`` in the build container:
that script asserts oracle == reference bit-for-bit on every stage tap of every fixture and writes
``tests/golden/*.npz``, which ``tests/test_oracle_golden.py`` re-checks wherever the suite runs.

``project_chain_np`` is a second, op-by-op numpy restatement of the voxel -> sample-coordinate chain
with every fp32 rounding explicit; it is the *specification of the CUDA kernels' in-kernel
projection* and is asserted bit-identical to the torch chain (``torch.mm`` with k=3 evaluates as an
in-order FMA chain on CPU - measured).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
f32 = np.float32


# =============================================================================================
# a1  voxel grids                                       project_whole.py:28-47 / project_individual.py:44-63
# =============================================================================================
def axis_coords(size: float, center: float, n: int) -> Tensor:
    """linspace(-size/2, size/2, n) + center, fp32 (project_whole.py:34-40)."""
    return torch.linspace(-size / 2, size / 2, int(n)) + center


def voxel_grid(space_size: Sequence[float], space_center: Sequence[float], nbins: Sequence[int]) -> Tensor:
    """[X*Y*Z, 3] world coordinates, flat index (ix*Y+iy)*Z+iz (project_whole.py:37-47)."""
    ax = [axis_coords(space_size[d], space_center[d], nbins[d]) for d in range(3)]
    gx, gy, gz = torch.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    return torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1).contiguous()


# =============================================================================================
# a2-a4  camera projection chain                         cameras.py:11-56, transforms.py:59-63,
#                                                        project_whole.py:49-60
# =============================================================================================
def camera_tensors(cam: dict) -> Tuple[Tensor, ...]:
    """float64 calibration -> fp32 tensors (cameras.py:11-18)."""
    R = torch.as_tensor(np.asarray(cam["R"]), dtype=torch.float)
    T = torch.as_tensor(np.asarray(cam["T"]), dtype=torch.float).reshape(3, 1)
    f = torch.tensor(np.array([cam["fx"], cam["fy"]]), dtype=torch.float).reshape(2, 1)
    c = torch.tensor(np.array([cam["cx"], cam["cy"]]), dtype=torch.float).reshape(2, 1)
    k = torch.as_tensor(np.asarray(cam["k"]), dtype=torch.float).reshape(3, 1)
    p = torch.as_tensor(np.asarray(cam["p"]), dtype=torch.float).reshape(2, 1)
    return R, T, f, c, k, p


def project_points(x: Tensor, cam: dict) -> Tensor:
    """[n,3] world -> [n,2] original-image pixels (cameras.py:30-56). No cheirality test."""
    R, T, f, c, k, p = camera_tensors(cam)
    q = torch.mm(R, x.T - T)
    y = q[:2] / (q[2] + 1e-5)
    r = torch.sum(y ** 2, dim=0)
    d = 1 + k[0] * r + k[1] * r * r + k[2] * r * r * r
    u = y[0, :] * d + 2 * p[0] * y[0, :] * y[1, :] + p[1] * (r + 2 * y[0, :] * y[0, :])
    v = y[1, :] * d + 2 * p[1] * y[0, :] * y[1, :] + p[0] * (r + 2 * y[1, :] * y[1, :])
    y = torch.stack([u, v])
    return (f * y + c).T


def sample_grid(points: Tensor, cam: dict, resize: Tensor, ori_size, image_size, hm_size) -> Tensor:
    """[n,2] normalised grid_sample coordinates of world points (project_whole.py:49-60)."""
    w, h = int(hm_size[0]), int(hm_size[1])
    xy = project_points(points, cam)
    xy = torch.clamp(xy, -1.0, float(max(ori_size[0], ori_size[1])))      # same bound for x and y
    homo = torch.cat([xy, torch.ones(xy.shape[0], 1)], dim=1)
    xy = torch.mm(resize, homo.T)[:2].T                                    # transforms.py:59-63
    xy = xy * torch.tensor([w, h], dtype=torch.float) / torch.tensor(
        [float(image_size[0]), float(image_size[1])], dtype=torch.float)
    g = xy / torch.tensor([w - 1, h - 1], dtype=torch.float) * 2.0 - 1.0
    return torch.clamp(g, -1.1, 1.1)


def _fma(a, b, c):
    """fp32 fused multiply-add, emulated in x87 extended precision: the product of two fp32 values is
    exact (48 bits) and the 64-bit-mantissa sum is rounded once more to fp32 - a float64 emulation
    double-rounds about once per 10^7 operations (observed: 1 of 2.56 M coordinates)."""
    wide = np.longdouble
    return (np.asarray(a, wide) * np.asarray(b, wide) + np.asarray(c, wide)).astype(f32)


def project_chain_np(px: np.ndarray, py: np.ndarray, pz: np.ndarray, cam21: np.ndarray,
                     resize6: np.ndarray, ori_max: float, hm_wh, img_wh) -> Tuple[np.ndarray, np.ndarray]:
    """Explicitly rounded fp32 restatement of ``sample_grid`` + grid_sample's un-normalisation.

    Inputs are fp32 world coordinates; ``cam21`` = fp32 [R9 T3 fx fy cx cy k3 p2]; ``resize6`` = fp32
    2x3 row-major.  Returns heat-map pixel coordinates (ix, iy) exactly as ``F.grid_sample(...,
    align_corners=True)`` computes them: ((g+1)/2)*(size-1).  This is the kernels' contract
    (csrc/fvp_project.cuh implements the same sequence with __fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn).
    """
    c = np.asarray(cam21, f32)
    A = np.asarray(resize6, f32)
    px, py, pz = (np.asarray(a, f32) for a in (px, py, pz))
    dx, dy, dz = (px - c[9]).astype(f32), (py - c[10]).astype(f32), (pz - c[11]).astype(f32)
    q = []
    for r in range(3):  # torch.mm, k=3: acc = R0*dx; acc = fma(R1,dy,acc); acc = fma(R2,dz,acc)
        acc = (c[3 * r] * dx).astype(f32)
        acc = _fma(c[3 * r + 1], dy, acc)
        acc = _fma(c[3 * r + 2], dz, acc)
        q.append(acc)
    den = (q[2] + f32(1e-5)).astype(f32)
    y0 = (q[0] / den).astype(f32)
    y1 = (q[1] / den).astype(f32)
    r2 = ((y0 * y0).astype(f32) + (y1 * y1).astype(f32)).astype(f32)
    k0, k1, k2, p0, p1 = c[16], c[17], c[18], c[19], c[20]
    d = (f32(1) + (k0 * r2).astype(f32)).astype(f32)
    d = (d + ((k1 * r2).astype(f32) * r2).astype(f32)).astype(f32)
    d = (d + (((k2 * r2).astype(f32) * r2).astype(f32) * r2).astype(f32)).astype(f32)
    two = f32(2)
    u = ((y0 * d).astype(f32) + (((two * p0).astype(f32) * y0).astype(f32) * y1).astype(f32)).astype(f32)
    u = (u + (p1 * (r2 + ((two * y0).astype(f32) * y0).astype(f32)).astype(f32)).astype(f32)).astype(f32)
    v = ((y1 * d).astype(f32) + (((two * p1).astype(f32) * y0).astype(f32) * y1).astype(f32)).astype(f32)
    v = (v + (p0 * (r2 + ((two * y1).astype(f32) * y1).astype(f32)).astype(f32)).astype(f32)).astype(f32)
    X = ((c[12] * u).astype(f32) + c[14]).astype(f32)
    Y = ((c[13] * v).astype(f32) + c[15]).astype(f32)
    hi = f32(ori_max)
    X = np.minimum(np.maximum(X, f32(-1)), hi)
    Y = np.minimum(np.maximum(Y, f32(-1)), hi)
    # torch.mm(A[2x3], [X;Y;1]): acc = A0*X; fma(A1,Y,acc); fma(A2,1,acc)
    ax = _fma(A[2], f32(1), _fma(A[1], Y, (A[0] * X).astype(f32)))
    ay = _fma(A[5], f32(1), _fma(A[4], Y, (A[3] * X).astype(f32)))
    w, h = f32(hm_wh[0]), f32(hm_wh[1])
    sx = ((ax * w).astype(f32) / f32(img_wh[0])).astype(f32)
    sy = ((ay * h).astype(f32) / f32(img_wh[1])).astype(f32)
    gx = (((sx / f32(hm_wh[0] - 1)).astype(f32) * two).astype(f32) - f32(1)).astype(f32)
    gy = (((sy / f32(hm_wh[1] - 1)).astype(f32) * two).astype(f32) - f32(1)).astype(f32)
    gx = np.minimum(np.maximum(gx, f32(-1.1)), f32(1.1))
    gy = np.minimum(np.maximum(gy, f32(-1.1)), f32(1.1))
    ix = (((gx + f32(1)).astype(f32) / two).astype(f32) * f32(hm_wh[0] - 1)).astype(f32)
    iy = (((gy + f32(1)).astype(f32) / two).astype(f32) * f32(hm_wh[1] - 1)).astype(f32)
    return ix, iy


def cam21_f32(cam: dict) -> np.ndarray:
    c = np.zeros(21, f32)
    c[0:9] = np.asarray(cam["R"], np.float64).reshape(9).astype(f32)
    c[9:12] = np.asarray(cam["T"], np.float64).reshape(3).astype(f32)
    c[12:16] = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float64).astype(f32)
    c[16:19] = np.asarray(cam["k"], np.float64).reshape(3).astype(f32)
    c[19:21] = np.asarray(cam["p"], np.float64).reshape(2).astype(f32)
    return c


def bilinear_zeros_np(hm: np.ndarray, ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    """Explicit 4-tap bilinear with zero padding ([C,H,W], [n], [n] -> [C,n]): the CUDA kernels'
    tap arithmetic (weights (x1-x)(y1-y)...; accumulate NW,NE,SW,SE)."""
    C, H, W = hm.shape
    x0 = np.floor(ix).astype(f32)
    y0 = np.floor(iy).astype(f32)
    x1, y1 = (x0 + f32(1)).astype(f32), (y0 + f32(1)).astype(f32)
    wx1, wy1 = (ix - x0).astype(f32), (iy - y0).astype(f32)
    wx0, wy0 = (x1 - ix).astype(f32), (y1 - iy).astype(f32)
    out = np.zeros((C, ix.shape[0]), f32)
    for (xx, yy, ww) in ((x0, y0, (wx0 * wy0).astype(f32)), (x1, y0, (wx1 * wy0).astype(f32)),
                         (x0, y1, (wx0 * wy1).astype(f32)), (x1, y1, (wx1 * wy1).astype(f32))):
        xi, yi = xx.astype(np.int64), yy.astype(np.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        val = np.where(ok[None, :], hm[:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)], f32(0))
        out = (out + (val * ww[None, :]).astype(f32)).astype(f32)
    return out


# =============================================================================================
# a5-a6  whole-space back-projection + z-max              project_whole.py:62-88, cnns_2d.py:174
# =============================================================================================
def hdn_sample_grids(cfg, cams: Sequence[dict], resize: Tensor) -> Tensor:
    """[V,1,nbins,2] cached per sequence in the reference (project_whole.py:75-80)."""
    grid = voxel_grid(cfg.CAPTURE_SPEC.SPACE_SIZE, cfg.CAPTURE_SPEC.SPACE_CENTER, cfg.CAPTURE_SPEC.VOXELS_PER_AXIS)
    gs = [sample_grid(grid, cam, resize, cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE,
                      cfg.DATASET.HEATMAP_SIZE).view(1, -1, 2) for cam in cams]
    return torch.stack(gs, dim=0)


def hdn_cubes(cfg, heatmaps: Tensor, grids: Sequence[Tensor]) -> Tensor:
    """[B,J,X,Y,Z] = clamp(mean_v grid_sample(hm[b,v], g[v]), 0, 1) (project_whole.py:83-87).
    ``grids[b]`` is the [V,1,nbins,2] grid of frame b's sequence."""
    B, V, J = heatmaps.shape[:3]
    X, Y, Z = [int(v) for v in cfg.CAPTURE_SPEC.VOXELS_PER_AXIS]
    cubes = torch.zeros(B, J, 1, X * Y * Z)
    for b in range(B):
        cubes[b] = torch.mean(F.grid_sample(heatmaps[b], grids[b], align_corners=True), dim=0).squeeze(0)
    return cubes.clamp(0.0, 1.0).view(B, J, X, Y, Z)


# =============================================================================================
# a7 / a10 / a16  CNN trunks (functional, state_dict driven)   cnns_2d.py:12-178, cnns_1d.py:10-132
# =============================================================================================
def _conv(sd, key: str, x: Tensor, nd: int, pad: int) -> Tensor:
    fn = F.conv2d if nd == 2 else F.conv1d
    return fn(x, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=pad)


def _bn(sd, key: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], training=False, momentum=0.1, eps=1e-5)


def _res_block(sd, p: str, x: Tensor, nd: int) -> Tensor:
    """Res2DBlock / Res1DBlock (cnns_2d.py:25-47, cnns_1d.py:23-45)."""
    y = F.relu(_bn(sd, p + ".res_branch.1", _conv(sd, p + ".res_branch.0", x, nd, 1)))
    y = _bn(sd, p + ".res_branch.4", _conv(sd, p + ".res_branch.3", y, nd, 1))
    if (p + ".skip_con.0.weight") in sd:
        s = _bn(sd, p + ".skip_con.1", _conv(sd, p + ".skip_con.0", x, nd, 0))
    else:
        s = x
    return F.relu(y + s)


def _upsample(sd, p: str, x: Tensor, nd: int) -> Tensor:
    """ConvTranspose(k=2,s=2)+BN+ReLU (cnns_2d.py:58-71)."""
    fn = F.conv_transpose2d if nd == 2 else F.conv_transpose1d
    y = fn(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], stride=2)
    return F.relu(_bn(sd, p + ".block.1", y))


def trunk(sd, prefix: str, x: Tensor, nd: int) -> Tensor:
    """front_layers + EncoderDecorder (cnns_2d.py:94-112,131-133)."""
    pool = (lambda t: F.max_pool2d(t, 2, 2)) if nd == 2 else (lambda t: F.max_pool1d(t, 2, 2))
    fl, ed = prefix + ".front_layers", prefix + ".encoder_decoder"
    x = F.relu(_bn(sd, fl + ".0.block.1", _conv(sd, fl + ".0.block.0", x, nd, 3)))
    x = _res_block(sd, fl + ".1", x, nd)
    skip1 = _res_block(sd, ed + ".skip_res1", x, nd)
    x = _res_block(sd, ed + ".encoder_res1", pool(x), nd)
    skip2 = _res_block(sd, ed + ".skip_res2", x, nd)
    x = _res_block(sd, ed + ".encoder_res2", pool(x), nd)
    x = _res_block(sd, ed + ".mid_res", x, nd)
    x = _res_block(sd, ed + ".decoder_res2", x, nd)
    x = _upsample(sd, ed + ".decoder_upsample2", x, nd) + skip2
    x = _res_block(sd, ed + ".decoder_res1", x, nd)
    x = _upsample(sd, ed + ".decoder_upsample1", x, nd) + skip1
    return x


def center_net(sd, plane: Tensor, prefix: str = "pose_net.center_net") -> Tuple[Tensor, Tensor]:
    """[B,J,X,Y] (already z-maxed) -> hm [B,1,X,Y], size [B,2,X,Y] (cnns_2d.py:173-178)."""
    x = trunk(sd, prefix, plane, 2)
    hm = _conv(sd, prefix + ".output_hm.2", F.relu(_conv(sd, prefix + ".output_hm.0", x, 2, 1)), 2, 0)
    size = _conv(sd, prefix + ".output_size.2", F.relu(_conv(sd, prefix + ".output_size.0", x, 2, 1)), 2, 0)
    return hm, size


def c2c_net(sd, cols: Tensor, prefix: str = "pose_net.c2c_net") -> Tensor:
    """[n,J,Z] -> [n,1,Z] (cnns_1d.py:128-132)."""
    return _conv(sd, prefix + ".output_hm", trunk(sd, prefix, cols, 1), 1, 0)


def p2p_net(sd, planes: Tensor, prefix: str = "joint_net.conv_net") -> Tensor:
    """[n,J,64,64] -> [n,J,64,64] (cnns_2d.py:131-135)."""
    return _conv(sd, prefix + ".output_layer", trunk(sd, prefix, planes, 2), 2, 0)


# =============================================================================================
# a8-a11  NMS / top-k / gathers / proposals            core/proposal.py:13-33, human_detection_net.py:44-104
# =============================================================================================
def nms_topk(hm: Tensor, max_num: int) -> Tuple[Tensor, Tensor, Tensor]:
    """keep=(hm==maxpool3x3(hm)); topk over flat X*Y; (x,y)=(flat//X', flat%X') with X'=shape[1]
    of the [1,X,Y] map, i.e. the X extent (core/proposal.py:13-33; harmless quirk when X==Y)."""
    B = hm.shape[0]
    mx = F.max_pool2d(hm, kernel_size=3, stride=1, padding=1)
    nms = ((hm == mx).float() * hm).reshape(B, -1)
    vals, flat = nms.topk(max_num)
    div = hm.shape[2]
    idx = torch.stack([torch.div(flat, div, rounding_mode="trunc"), flat % div], dim=2)
    return vals, idx, flat


def hdn_head(cfg, sd, cubes: Tensor) -> Dict[str, Tensor]:
    """CenterNet -> NMS/top-k -> bbox + z-column gather -> C2CNet -> proposals
    (human_detection_net.py:76-104, ProposalLayer.forward :44-65 eval branch)."""
    B, J = cubes.shape[:2]
    P = int(cfg.CAPTURE_SPEC.MAX_PEOPLE)
    plane = torch.max(cubes, dim=4)[0]
    hm2d, size = center_net(sd, plane)
    conf2d, idx_xy, flat = nms_topk(hm2d, P)
    bbox = torch.gather(torch.flatten(size, 2, 3).permute(0, 2, 1), 1, flat.unsqueeze(2).repeat(1, 1, 2))
    cols = torch.gather(torch.flatten(cubes, 2, 3).permute(0, 2, 1, 3), 1,
                        flat.view(B, -1, 1, 1).repeat(1, 1, J, cubes.shape[4]))
    hm1d = c2c_net(sd, torch.flatten(cols, 0, 1)).view(B, P, -1)
    conf1d, idx_z = hm1d.topk(1)
    index = torch.cat([idx_xy, idx_z], dim=2)
    conf = conf2d * conf1d.squeeze(2)
    size_t = torch.tensor(cfg.CAPTURE_SPEC.SPACE_SIZE)
    scale = size_t / (torch.tensor(cfg.CAPTURE_SPEC.VOXELS_PER_AXIS) - 1)
    bias = torch.tensor(cfg.CAPTURE_SPEC.SPACE_CENTER) - size_t / 2.0
    centers = torch.zeros(B, P, 7)
    centers[:, :, 0:3] = index.float() * scale + bias
    centers[:, :, 4] = conf
    centers[:, :, 3] = (conf > float(cfg.CAPTURE_SPEC.MIN_SCORE)).float() - 1.0
    centers[:, :, 5:7] = bbox
    return {"plane": plane, "hm2d": hm2d, "size": size, "conf2d": conf2d, "idx_xy": idx_xy, "flat": flat,
            "bbox": bbox, "cols": cols, "hm1d": hm1d, "conf1d": conf1d.squeeze(2), "idx_z": idx_z.squeeze(2),
            "centers": centers}


# =============================================================================================
# a13-a15  per-person crops + three orthographic planes   project_individual.py:14-136,
#                                                         joint_localization_net.py:80-81
# =============================================================================================
class JlnConstants:
    """Constants of the fine whole-space grid (project_individual.py:21-42)."""

    def __init__(self, cfg):
        self.center = torch.tensor(cfg.CAPTURE_SPEC.SPACE_CENTER)
        self.whole = torch.tensor(cfg.CAPTURE_SPEC.SPACE_SIZE)
        self.ind = torch.tensor(cfg.INDIVIDUAL_SPEC.SPACE_SIZE)
        self.vox = torch.tensor(cfg.INDIVIDUAL_SPEC.VOXELS_PER_AXIS, dtype=torch.int32)
        self.fine = (self.whole / self.ind * (self.vox - 1)).int() + 1
        self.scale = (self.fine.float() - 1) / self.whole
        self.bias = -self.ind / 2.0 / self.whole * (self.fine - 1) - self.scale * (self.center - self.whole / 2.0)
        g = voxel_grid(self.ind.tolist(), self.center.tolist(), self.vox.tolist()).view(
            int(self.vox[0]), int(self.vox[1]), int(self.vox[2]), 3)
        # plane coordinates for soft-argmax: xy, xz, yz (project_individual.py:38-40)
        self.center_grid = torch.stack([g[:, :, 0, :2].reshape(-1, 2), g[:, 0, :, ::2].reshape(-1, 2),
                                        g[0, :, :, 1:].reshape(-1, 2)])
        self.fine_axes = [axis_coords(float(self.whole[d]), float(self.center[d]), int(self.fine[d])) for d in range(3)]


def jln_crop_params(K: JlnConstants, centers: Tensor) -> Dict[str, Tensor]:
    """tl / offset / bbox mask / start / end for [n,7] proposals (project_individual.py:110-121)."""
    n = centers.shape[0]
    tl = torch.round(centers[:, :3].float() * K.scale + K.bias).int()
    offset = tl.float() / (K.fine - 1) * K.whole - K.whole / 2.0 + K.ind / 2.0
    m = ((1 - centers[:, 5:7]) / 2 * (K.vox[0:2] - 1)).int()
    m[m < 0] = 0
    m = torch.cat([m, torch.zeros((n, 1), dtype=torch.int32)], dim=1)
    start = torch.where(tl + m >= 0, tl + m, torch.zeros_like(tl))
    end = torch.where(tl + K.vox - m <= K.fine, tl + K.vox - m, K.fine)
    return {"tl": tl, "offset": offset, "mask": m, "start": start, "end": end}


def jln_cubes(cfg, K: JlnConstants, heatmaps_b: Tensor, cams: Sequence[dict], resize: Tensor,
              crop: Dict[str, Tensor]) -> Tensor:
    """[n,J,64,64,64] person cubes of one frame (project_individual.py:96-136).

    The reference caches the projected *fine* grid of the whole space (164 MB) and slices it; the
    sample coordinate of fine voxel (gx,gy,gz) depends only on its world coordinate, so this
    restatement projects just the cropped sub-grid with the same elementwise expressions."""
    V, J = heatmaps_b.shape[:2]
    n = crop["tl"].shape[0]
    vx = [int(v) for v in K.vox]
    cubes = torch.zeros(n, J, vx[0], vx[1], vx[2])
    for i in range(n):
        s, e, tl = crop["start"][i].tolist(), crop["end"][i].tolist(), crop["tl"][i].tolist()
        if any(s[d] >= e[d] for d in range(3)):
            continue
        ax = [K.fine_axes[d][s[d]:e[d]] for d in range(3)]
        gx, gy, gz = torch.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
        pts = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1).contiguous()
        grid = torch.stack([sample_grid(pts, cam, resize, cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE,
                                        cfg.DATASET.HEATMAP_SIZE).view(1, -1, 2) for cam in cams], dim=0)
        acc = torch.mean(F.grid_sample(heatmaps_b, grid, align_corners=True), dim=0)
        cubes[i, :, s[0] - tl[0]:e[0] - tl[0], s[1] - tl[1]:e[1] - tl[1], s[2] - tl[2]:e[2] - tl[2]] = \
            acc.view(J, e[0] - s[0], e[1] - s[1], e[2] - s[2])
    return cubes.clamp(0.0, 1.0)


def three_planes(cubes: Tensor) -> Tensor:
    """[n,J,a,b,c] -> [3n,J,64,64] plane-major: xy=max_c, xz=max_b, yz=max_a
    (joint_localization_net.py:80-81)."""
    return torch.cat([torch.max(cubes, dim=4)[0], torch.max(cubes, dim=3)[0], torch.max(cubes, dim=2)[0]])


# =============================================================================================
# a17-a19  soft-argmax, WeightNet, fusion        joint_localization_net.py:15-62, weight_net.py:48-80
# =============================================================================================
def soft_argmax(feat: Tensor, center_grid: Tensor, beta: float) -> Tuple[Tensor, Tensor]:
    """[3,n,J,64,64] -> pose [3,n,J,2], conf [n] (joint_localization_net.py:20-33)."""
    n, J = feat.shape[1], feat.shape[2]
    x = F.softmax(beta * feat.reshape(3, n, J, -1, 1), dim=3)
    confs = torch.mean(torch.max(x, dim=3)[0].squeeze(3), dim=(0, 2))
    pos = torch.sum(torch.mul(x, center_grid.reshape(3, 1, 1, -1, 2)), dim=3)
    return pos, confs


def weight_net(sd, feat: Tensor, prefix: str = "joint_net.weight_net") -> Tensor:
    """[3,n,J,64,64] -> [3n,J,1] (weight_net.py:69-80)."""
    x = torch.flatten(feat, 0, 1)
    n3, J, H, W = x.shape
    x = x.reshape(n3 * J, 1, H, W)
    x = _bn(sd, prefix + ".heatmap_feature_net.1", _conv(sd, prefix + ".heatmap_feature_net.0", x, 2, 1))
    x = F.relu(F.max_pool2d(x, 2))
    x = F.adaptive_avg_pool2d(x, 1).view(n3 * J, -1)
    x = F.relu(F.linear(x, sd[prefix + ".output.0.weight"], sd[prefix + ".output.0.bias"]))
    x = torch.sigmoid(F.linear(x, sd[prefix + ".output.2.weight"], sd[prefix + ".output.2.bias"]))
    return x.view(n3, J, 1)


def fuse(pose: Tensor, weights: Tensor) -> Tensor:
    """[3,n,J,2] + [3n,J,1] -> [n,J,3] (joint_localization_net.py:44-62): normalise, then blend."""
    w_xy, w_xz, w_yz = torch.chunk(weights, 3)
    xy, xz, yz = pose[0], pose[1], pose[2]
    wx = torch.cat([w_xy, w_xz], dim=2)
    wy = torch.cat([w_xy, w_yz], dim=2)
    wz = torch.cat([w_xz, w_yz], dim=2)
    wx = wx / torch.sum(wx, dim=2).unsqueeze(2)
    wy = wy / torch.sum(wy, dim=2).unsqueeze(2)
    wz = wz / torch.sum(wz, dim=2).unsqueeze(2)
    x = wx[:, :, :1] * xy[:, :, :1] + wx[:, :, 1:] * xz[:, :, :1]
    y = wy[:, :, :1] * xy[:, :, 1:] + wy[:, :, 1:] * yz[:, :, :1]
    z = wz[:, :, :1] * xz[:, :, 1:] + wz[:, :, 1:] * yz[:, :, 1:]
    return torch.cat([x, y, z], dim=2)


# =============================================================================================
# float64 yardstick of a14-a19 ("measure the reference's own fp32-vs-fp64 noise floor in the same run")
# =============================================================================================
def jln_fp64(cfg, sd: Dict[str, Tensor], heatmaps_b: Tensor, cams: Sequence[dict], resize: Tensor,
             centers_valid: Tensor) -> Dict[str, Tensor]:
    """JointLocalizationNet for the valid proposals [n,7] of one frame, evaluated in float64 downstream of everything
    that is bit-exact on both sides of the parity tests: proposal centres, crop parameters and the fp32 sample positions
    (ix, iy) of every fine voxel are the reference's own fp32 values; the bilinear taps, view mean, clamp, three-plane
    max, P2PNet, soft-argmax, WeightNet and fusion run in float64 on float64 copies of the weights.  |reference_fp32 -
    this| is the reference's own rounding noise; an implementation is "as good as the reference" when it is at most
    that far from this result (tests/test_gpu_parity.py prints and asserts both)."""
    K = JlnConstants(cfg)
    crop = jln_crop_params(K, centers_valid)
    V, J, H, W = heatmaps_b.shape
    n = crop["tl"].shape[0]
    hm64 = heatmaps_b.double().numpy()
    vx = [int(v) for v in K.vox]
    cubes = np.zeros((n, J, vx[0], vx[1], vx[2]), np.float64)
    for i in range(n):
        s, e, tl = crop["start"][i].tolist(), crop["end"][i].tolist(), crop["tl"][i].tolist()
        if any(s[d] >= e[d] for d in range(3)):
            continue
        ax = [K.fine_axes[d][s[d]:e[d]] for d in range(3)]
        gx, gy, gz = torch.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
        pts = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1).contiguous()
        acc = np.zeros((J, pts.shape[0]), np.float64)
        for v, cam in enumerate(cams):
            g = sample_grid(pts, cam, resize, cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE, cfg.DATASET.HEATMAP_SIZE)
            # ATen grid_sampler_unnormalize, align_corners: ((g + 1) / 2) * (size - 1), in fp32 like the reference
            ix = (((g[:, 0] + 1.0) / 2.0) * float(W - 1)).numpy().astype(np.float64)
            iy = (((g[:, 1] + 1.0) / 2.0) * float(H - 1)).numpy().astype(np.float64)
            x0, y0 = np.floor(ix), np.floor(iy)
            fx, fy = ix - x0, iy - y0
            for xx, yy, ww in ((x0, y0, (1 - fx) * (1 - fy)), (x0 + 1, y0, fx * (1 - fy)), (x0, y0 + 1, (1 - fx) * fy),
                               (x0 + 1, y0 + 1, fx * fy)):
                xi, yi = xx.astype(np.int64), yy.astype(np.int64)
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                acc += np.where(ok[None, :], hm64[v][:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)], 0.0) * ww[None, :]
        cubes[i, :, s[0] - tl[0]:e[0] - tl[0], s[1] - tl[1]:e[1] - tl[1], s[2] - tl[2]:e[2] - tl[2]] = \
            (acc / V).reshape(J, e[0] - s[0], e[1] - s[1], e[2] - s[2])
    cubes_t = torch.from_numpy(cubes).clamp(0.0, 1.0)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    planes = three_planes(cubes_t)
    feat = torch.stack(torch.chunk(p2p_net(sd64, planes), 3), dim=0)
    pose, confs = soft_argmax(feat, K.center_grid.double(), float(cfg.NETWORK.BETA))
    off = crop["offset"].double().reshape(-1, 1, 3)
    pose[0] += off[:, :, :2]
    pose[1] += off[:, :, ::2]
    pose[2] += off[:, :, 1:]
    w = weight_net(sd64, feat)
    return {"planes": planes, "feat": feat, "pose": pose, "confs": confs, "weights": w, "fused": fuse(pose, w)}


# =============================================================================================
# a12  whole forward (eval branch)                      faster_voxelpose.py:34-48,99-105
# =============================================================================================
def forward(cfg, sd: Dict[str, Tensor], heatmaps: Tensor, seqs: Sequence[str], cameras: Dict[str, Sequence[dict]],
            resize: Tensor, taps: bool = True) -> Dict[str, object]:
    """Eval-mode forward on precomputed heatmaps [B,V,J,H,W].  Returns the reference's outputs
    (``fused_poses`` [B,P,J,5], ``plane_poses`` [3,B,P,J,2], ``proposal_centers`` [B,P,7]) plus every
    intermediate named in SURVEY.md §8a when ``taps``."""
    B, V, J = heatmaps.shape[:3]
    P = int(cfg.CAPTURE_SPEC.MAX_PEOPLE)
    beta = float(cfg.NETWORK.BETA)
    out: Dict[str, object] = {}
    grid_cache: Dict[str, Tensor] = {}
    for s in seqs:
        assert s in cameras, "missing camera parameters for the current sequence"
        assert len(cameras[s]) == V, "inconsistent number of cameras"
        if s not in grid_cache:
            grid_cache[s] = hdn_sample_grids(cfg, cameras[s], resize)
    cubes = hdn_cubes(cfg, heatmaps, [grid_cache[s] for s in seqs])
    hdn = hdn_head(cfg, sd, cubes)
    centers = hdn["centers"].clone()
    valid = centers[:, :, 3] >= 0

    K = JlnConstants(cfg)
    fused = torch.zeros(B, P, J, 3)
    plane_poses = torch.zeros(3, B, P, J, 2)
    jl: List[Optional[Dict[str, Tensor]]] = []
    for b in range(B):
        if int(valid[b].sum()) == 0:
            jl.append(None)
            continue
        crop = jln_crop_params(K, centers[b, valid[b]])
        pc = jln_cubes(cfg, K, heatmaps[b], cameras[seqs[b]], resize, crop)
        planes = three_planes(pc)
        feat = torch.stack(torch.chunk(p2p_net(sd, planes), 3), dim=0)
        pose, confs = soft_argmax(feat, K.center_grid, beta)
        off = crop["offset"].reshape(-1, 1, 3)
        pose[0] += off[:, :, :2]
        pose[1] += off[:, :, ::2]
        pose[2] += off[:, :, 1:]
        w = weight_net(sd, feat)
        fp = fuse(pose, w)
        fused[b, valid[b]] = fp
        plane_poses[:, b, valid[b]] = pose
        centers[b, valid[b], 4] = confs                     # joint_localization_net.py:98 (alias write)
        jl.append({"crop_tl": crop["tl"], "crop_offset": crop["offset"], "crop_mask": crop["mask"],
                   "crop_start": crop["start"], "crop_end": crop["end"], "planes": planes, "feat": feat,
                   "pose": pose, "confs": confs, "weights": w, "fused": fp})
    out["fused_poses"] = torch.cat([fused, centers[:, :, 3:5].reshape(B, -1, 1, 2).repeat(1, 1, J, 1)], dim=3)
    out["plane_poses"] = plane_poses
    out["proposal_centers"] = centers
    if taps:
        out["hdn"] = hdn
        out["hdn_centers"] = hdn["centers"]
        out["jln"] = jl
    return out
