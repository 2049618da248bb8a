#!/usr/bin/env python
"""First-light parity + timing report on a B200 (prints measured errors for every stage)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden, CASES
from fvp.engine import Engine

def err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0

def run_case(name):
    g = Golden(name)
    eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=max(2, g.B), max_sequences=2, axes=g.axes)
    eng.load_state_dict(g.weights)
    slot = eng.sequence_slot(g.cams, g.resize)
    slots = [slot] * g.B
    hm = torch.from_numpy(g.heatmaps).cuda()
    rep = {}
    # stages
    eng.stage_heatmaps(hm)
    plane = eng.hdn_project(g.B, slots); torch.cuda.synchronize()
    rep["plane"] = err(plane.cpu(), g["hdn_plane"])
    hm2d, size = eng.center_net(torch.from_numpy(g["hdn_plane"]), g.B)
    rep["hm2d|plane"] = err(hm2d.cpu(), g["hm2d"][:, 0]); rep["size|plane"] = err(size.cpu(), g["size"])
    conf, flat = eng.nms_topk(torch.from_numpy(g["hm2d"][:, 0]))
    rep["flat==golden"] = bool(np.array_equal(flat.cpu().numpy(), g["flat"])) ; rep["conf2d"] = err(conf.cpu(), g["conf2d"])
    cols, hm1d, centers = eng.proposals(g.B, slots, torch.from_numpy(g["conf2d"]), torch.from_numpy(g["flat"]).int(), torch.from_numpy(g["size"]))
    rep["cols"] = err(cols.cpu().view(g.B, g.P, g.J, -1), g["cols"]); rep["hm1d"] = err(hm1d.cpu().view(g.B, g.P, -1), g["hm1d"])
    rep["centers"] = err(centers.cpu(), g["hdn_centers"])
    hm1d2 = eng.c2c_net(torch.from_numpy(g["cols"]).view(-1, g.J, g["cols"].shape[-1]))
    rep["hm1d|cols"] = err(hm1d2.cpu().view(g.B, g.P, -1), g["hm1d"])
    planes, off = eng.jln_project(g.B, slots, torch.from_numpy(g["hdn_centers"]))
    torch.cuda.synchronize()
    planes = planes.cpu().numpy().reshape(3, g.B, g.P, g.J, 64, 64); off = off.cpu().numpy().reshape(g.B, g.P, 3)
    valid = g["hdn_centers"][:, :, 3] >= 0
    for b in range(g.B):
        if not g.has("b%d_pose" % b): continue
        idx = np.nonzero(valid[b])[0]
        keep = g["b%d_planes_keep" % b]
        rep["planes_b%d" % b] = err(planes[:, b, idx[:keep.shape[1]]], keep)
        rep["planes_sum_b%d" % b] = err(planes[:, b, idx].astype(np.float64).sum(axis=(3, 4)), g["b%d_planes_sum" % b])
        rep["offset_b%d" % b] = err(off[b, idx], g["b%d_crop_offset" % b])
        fk = g["b%d_feat_keep" % b]
        feat = eng.p2p_net(torch.from_numpy(keep.reshape(-1, g.J, 64, 64)))
        rep["feat|planes_b%d" % b] = err(feat.cpu().numpy().reshape(fk.shape), fk)
        pose, confs, w, fused = eng.pose_head(torch.from_numpy(fk), torch.from_numpy(g["b%d_crop_offset" % b][:fk.shape[1]]))
        n = fk.shape[1]
        rep["pose|feat_b%d" % b] = err(pose.cpu(), g["b%d_pose" % b][:, :n]); rep["confs|feat_b%d" % b] = err(confs.cpu(), g["b%d_confs" % b][:n])
        rep["weights|feat_b%d" % b] = err(w.cpu().numpy().reshape(3 * n, g.J), g["b%d_weights" % b].reshape(3, -1, g.J)[:, :n].reshape(3 * n, g.J))
        rep["fused|feat_b%d" % b] = err(fused.cpu(), g["b%d_fused" % b][:n])
    # end to end
    fused, plane_poses, centers = eng.forward(hm, slots); torch.cuda.synchronize()
    rep["e2e centers"] = err(centers.cpu(), g["proposal_centers"]); rep["e2e fused"] = err(fused.cpu(), g["fused_poses"])
    rep["e2e plane_poses"] = err(plane_poses.cpu(), g["plane_poses"])
    rep["e2e flag equal"] = bool(np.array_equal(fused.cpu().numpy()[..., 3], g["fused_poses"][..., 3]))
    rep["launches"] = eng.last_launch_count()
    print(name, "PARTIAL", json.dumps(rep)); sys.stdout.flush()
    # graph replay equals eager
    eng.use_cuda_graph(True)
    for _ in range(3):
        f2, p2, c2 = eng.forward(hm, slots)
    torch.cuda.synchronize()
    rep["graph==eager"] = bool(torch.equal(f2, fused) and torch.equal(c2, centers)); rep["graph launches"] = eng.last_launch_count()
    eng.use_cuda_graph(False)
    eng.set_profiling(True)
    for _ in range(3): eng.forward(hm, slots)
    rep["stage_ms"] = [round(t, 4) for t in eng.stage_times_ms()]
    eng.set_profiling(False)
    eng.close()
    print(name, json.dumps(rep, indent=None)); sys.stdout.flush()
    return rep

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    names = sys.argv[1:] or CASES
    for n in names:
        try:
            run_case(n)
        except Exception as e:
            import traceback; traceback.print_exc()
