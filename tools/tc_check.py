#!/usr/bin/env python
"""tcgen05 conv path vs the exact-fp32 CUDA-core path and vs the golden, stage by stage."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from golden_util import Golden
from fvp.engine import Engine
def err(a, b): return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
for name in (sys.argv[1:] or ["panoptic_256x192", "campus_b1"]):
    g = Golden(name)
    eng = Engine(g.cfg, torch.device("cuda:0"), max_batch=max(2, g.B), max_sequences=2, axes=g.axes)
    eng.load_state_dict(g.weights)
    slot = eng.sequence_slot(g.cams, g.resize); slots = [slot] * g.B
    rep = {}
    plane = torch.from_numpy(g["hdn_plane"])
    keep, fk = g["b0_planes_keep"], g["b0_feat_keep"]
    res = {}
    for mode in (0, 1):
        eng.set_conv_mode(mode)
        hm, size = eng.center_net(plane, g.B); torch.cuda.synchronize()
        feat = eng.p2p_net(torch.from_numpy(keep.reshape(-1, g.J, 64, 64))); torch.cuda.synchronize()
        res[mode] = (hm.cpu().numpy(), size.cpu().numpy(), feat.cpu().numpy().reshape(fk.shape))
        rep["hm2d_vs_golden_m%d" % mode] = err(res[mode][0], g["hm2d"][:, 0])
        rep["size_vs_golden_m%d" % mode] = err(res[mode][1], g["size"])
        rep["feat_vs_golden_m%d" % mode] = err(res[mode][2], fk)
    rep["hm2d_tc_vs_ffma"] = err(res[1][0], res[0][0]); rep["feat_tc_vs_ffma"] = err(res[1][2], res[0][2])
    hm = torch.from_numpy(g.heatmaps).cuda()
    out = {}
    for mode in (0, 1):
        eng.set_conv_mode(mode)
        f, p, c = eng.forward(hm, slots); torch.cuda.synchronize()
        out[mode] = (f.cpu().numpy(), c.cpu().numpy())
        rep["e2e_fused_m%d" % mode] = err(out[mode][0][..., :3], g["fused_poses"][..., :3])
        rep["e2e_cells_equal_m%d" % mode] = bool(np.array_equal(out[mode][1][..., :4], g["proposal_centers"][..., :4]))
        eng.set_profiling(True)
        for _ in range(3): eng.forward(hm, slots)
        rep["stage_ms_m%d" % mode] = [round(t, 4) for t in eng.stage_times_ms()]
        eng.set_profiling(False)
    print(name, json.dumps(rep)); sys.stdout.flush()
    eng.close()
