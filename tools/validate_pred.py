#!/usr/bin/env python
"""Campus / Shelf validation on this library without the reference's dataset classes (run/validate.py for
TEST_HEATMAP_SRC = 'pred'): detection file -> GPU heat maps -> hot path -> PCP.  Needs a B200 and the dataset folder:

    python tools/validate_pred.py --dataset shelf --data-dir data/Shelf --weights output/.../model_best.pth.tar
                                  [--cfg configs/shelf/jln64.yaml] [--batch 8] [--max-frames N]

data-dir holds calibration_<dataset>.json, pred_<dataset>_maskrcnn_hrnet_coco.pkl and (for the metric) actorsGT.mat.
Without --weights the model runs on the deterministic synthetic weights of fvp.synth (plumbing / timing only)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200", "lib"))
sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dataset", required=True, choices=["campus", "shelf"])
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--cfg", default="", help="reference-format YAML; default: the built-in preset of the dataset")
    ap.add_argument("--weights", default="", help="model_best.pth.tar of the reference (a bare state_dict)")
    ap.add_argument("--batch", type=int, default=0, help="frames per forward call (default: TEST.BATCH_SIZE of the config)")
    ap.add_argument("--max-frames", type=int, default=0)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args()

    import numpy as np
    import torch
    import models
    from fvp import config as fcfg, datasets as D, evaluate as E, synth
    from fvp.render import HeatmapRenderer
    from fvp.validate import validate_pred

    cfg = fcfg.load_yaml(args.cfg) if args.cfg else fcfg.preset(args.dataset)
    cfg.DEVICE = args.device
    name = args.dataset
    cams = D.load_calibration(os.path.join(args.data_dir, "calibration_%s.json" % name))
    pred = D.load_pred_pose2d(os.path.join(args.data_dir, "pred_%s_maskrcnn_hrnet_coco.pkl" % name))
    frames = D.CAMPUS_FRAMES if name == "campus" else D.SHELF_FRAMES
    if args.max_frames:
        frames = frames[:args.max_frames]
    mat = os.path.join(args.data_dir, "actorsGT.mat")
    actors = E.load_actors(mat) if os.path.isfile(mat) else None
    if actors is None:
        print("no actorsGT.mat in %s: poses only, no PCP" % args.data_dir)

    batch = args.batch or int(cfg.TEST.BATCH_SIZE)
    model = models.faster_voxelpose.FasterVoxelPoseNet(cfg, max_batch=batch)
    if args.weights:
        model.load_state_dict(torch.load(args.weights, map_location="cpu"))
    else:
        print("no --weights: deterministic synthetic weights (the poses mean nothing)")
        sd = synth.make_weights(int(cfg.DATASET.NUM_JOINTS), seed=7)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model = model.to(args.device).eval()
    renderer = HeatmapRenderer(model.engine())

    t0 = time.perf_counter()
    out = validate_pred(cfg, model, renderer, cams, pred, frames, name, actors=actors, batch_size=batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d frames in %.2f s (%.1f frames/s incl. rendering and metric)" % (len(frames), dt, len(frames) / dt))
    if out["msg"]:
        print(out["msg"])


if __name__ == "__main__":
    main()
