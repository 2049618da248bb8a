#!/usr/bin/env python
"""The reference's 1-GPU PyTorch path, timed on the same B200 (north-star target: ours >= 10x this).

The reference itself cannot travel to the GPU box (pure Python under /root/reference), so this tool runs the oracle
port (oracle/fvp_oracle.py: the reference's PyTorch ops in the reference's order) with every tensor on cuda:0, and -
like the reference (project_whole.py:75-80, project_individual.py:104-106) - with the whole-space and the 164 MB fine
sample grids cached per sequence, so the timed forward does what FasterVoxelPoseNet.forward does on a GPU: grid_sample
from cached grids, cuDNN convolutions (cudnn.benchmark as run/validate.py:61-63), top-k, soft-argmax, WeightNet.
This is measurement tooling (a reported baseline), never part of the product path.

    python tools/ref_gpu_port.py [--device cuda:0] [--steps 30] [--warmup 5] [--check]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import fvp_oracle as O  # noqa: E402


class CachedReference:
    """O.forward with the reference's per-sequence grid caches (one sequence)."""

    def __init__(self, cfg, sd, cams, resize, device):
        self.cfg, self.dev = cfg, torch.device(device)
        self.P = int(cfg.CAPTURE_SPEC.MAX_PEOPLE)
        self.beta = float(cfg.NETWORK.BETA)
        with torch.device(self.dev):
            self.sd = {k: v.to(self.dev) for k, v in sd.items()}
            rz = resize.to(self.dev)
            self.grid = O.hdn_sample_grids(cfg, cams, rz)                       # [V,1,nbins,2]
            self.K = O.JlnConstants(cfg)
            fine = [int(v) for v in self.K.fine]
            pts = O.voxel_grid(cfg.CAPTURE_SPEC.SPACE_SIZE, cfg.CAPTURE_SPEC.SPACE_CENTER, fine)
            self.fine_grid = torch.stack([O.sample_grid(pts, cam, rz, cfg.DATASET.ORI_IMAGE_SIZE, cfg.DATASET.IMAGE_SIZE,
                                                        cfg.DATASET.HEATMAP_SIZE).view(fine[0], fine[1], fine[2], 2)
                                          for cam in cams], dim=0)              # [V,fx,fy,fz,2] (project_individual.py:79-94)

    def _jln_cubes(self, hm_b, crop):
        V, J = hm_b.shape[:2]
        n = crop["tl"].shape[0]
        cubes = torch.zeros(n, J, 64, 64, 64)
        for i in range(n):
            s, e, tl = crop["start"][i].tolist(), crop["end"][i].tolist(), crop["tl"][i].tolist()   # host syncs, as the reference
            if any(s[d] >= e[d] for d in range(3)):
                continue
            g = self.fine_grid[:, s[0]:e[0], s[1]:e[1], s[2]:e[2]].reshape(V, 1, -1, 2)
            acc = torch.mean(F.grid_sample(hm_b, g, align_corners=True), dim=0)
            cubes[i, :, s[0] - tl[0]:e[0] - tl[0], s[1] - tl[1]:e[1] - tl[1], s[2] - tl[2]:e[2] - tl[2]] = \
                acc.view(J, e[0] - s[0], e[1] - s[1], e[2] - s[2])
        return cubes.clamp(0.0, 1.0)

    @torch.no_grad()
    def forward(self, heatmaps):
        cfg, sd, K, P = self.cfg, self.sd, self.K, self.P
        with torch.device(self.dev):
            B, V, J = heatmaps.shape[:3]
            cubes = O.hdn_cubes(cfg, heatmaps, [self.grid] * B)
            hdn = O.hdn_head(cfg, sd, cubes)
            centers = hdn["centers"].clone()
            valid = centers[:, :, 3] >= 0
            fused = torch.zeros(B, P, J, 3)
            for b in range(B):
                if int(valid[b].sum()) == 0:
                    continue
                crop = O.jln_crop_params(K, centers[b, valid[b]])
                planes = O.three_planes(self._jln_cubes(heatmaps[b], crop))
                feat = torch.stack(torch.chunk(O.p2p_net(sd, planes), 3), dim=0)
                pose, confs = O.soft_argmax(feat, K.center_grid, self.beta)
                off = crop["offset"].reshape(-1, 1, 3)
                pose[0] += off[:, :, :2]
                pose[1] += off[:, :, ::2]
                pose[2] += off[:, :, 1:]
                fused[b, valid[b]] = O.fuse(pose, O.weight_net(sd, feat))
                centers[b, valid[b], 4] = confs
            return torch.cat([fused, centers[:, :, 3:5].reshape(B, -1, 1, 2).repeat(1, 1, J, 1)], dim=3), centers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--check", action="store_true", help="compare with O.forward on the CPU (first frame)")
    args = ap.parse_args()
    import bench                                              # (lazily: bench.py imports CachedReference from this module)
    cfg, cams, resize = bench.workload("panoptic_256x192")
    from fvp import synth
    J = int(cfg.DATASET.NUM_JOINTS)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synth.make_weights(J, seed=2024).items()}
    rz = torch.as_tensor(resize, dtype=torch.float)
    frames = bench.make_frames(cfg, cams, 4, seed0=1000)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = False            # keep the reference's fp32 arithmetic
    torch.backends.cudnn.allow_tf32 = False
    ref = CachedReference(cfg, sd, cams, rz, args.device)
    dev = torch.device(args.device)
    pool = [torch.from_numpy(np.stack([frames[(i + b) % 4] for b in range(args.batch)])).to(dev) for i in range(4)]
    if args.check:
        want = O.forward(cfg, sd, torch.from_numpy(frames[:1]), ["s"], {"s": cams}, rz, taps=False)
        got, ctr = ref.forward(pool[0][:1])
        print("check vs CPU oracle: centers xyz/flag equal = %s, fused max|d| = %.3g mm" % (
            bool(torch.equal(ctr.cpu()[..., :4], want["proposal_centers"][..., :4])),
            float((got.cpu()[..., :3] - want["fused_poses"][..., :3]).abs().max())))

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        ref.forward(pool[i % 4])
    sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out, _ = ref.forward(pool[i % 4])
    sync()
    dt = time.perf_counter() - t0
    n_valid = int((out[..., 0, 3] >= 0).sum())
    print(json.dumps({"impl": "reference-port on %s (PyTorch %s ops, cached sample grids, fp32, cudnn.benchmark)" % (args.device, torch.__version__),
                      "metric": bench.METRIC, "value": args.batch * args.steps / dt, "unit": bench.UNIT,
                      "ms_per_step": dt / args.steps * 1e3, "batch": args.batch, "steps": args.steps, "valid_people_last_step": n_valid,
                      "gpu": torch.cuda.get_device_name(dev) if dev.type == "cuda" else "cpu"}))


if __name__ == "__main__":
    main()
