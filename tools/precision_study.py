#!/usr/bin/env python
"""CPU study for the next conv design step: what does it cost numerically to keep trunk
activations ONLY as the fp16 hi/lo pair the tensor cores consume (22 significant bits) instead of fp32?

Emulates P2PNet on golden JLN planes in float64 with the roundings of each engine made explicit:
  fp64      : everything in float64 (the yardstick)
  oracle    : the fp32 PyTorch-CPU oracle (= the reference)
  engine A  : today's engine 2 - every conv rounds BOTH operands to hi + lo*2^-11 (fp16 pair), accumulates (emulated in
              float64), epilogue in fp32, activations stored as fp32 (residuals and max-pool see fp32)
  engine B  : proposed - same, but the epilogue stores only the hi/lo pair, so residuals, max-pool and the next layer see
              the 22-bit value
Reports max |feature error| and the soft-argmax joint displacement in mm against the float64 yardstick.
"""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "faster-voxelpose_b200"))
from golden_util import Golden
from oracle import fvp_oracle as O


def r22(x: torch.Tensor) -> torch.Tensor:
    """hi = fp16(x), lo = fp16((x - hi) * 2^11); returns hi + lo * 2^-11 (float64 in, float64 out)."""
    hi = x.to(torch.float16).to(torch.float64)
    lo = ((x - hi) * 2048.0).to(torch.float16).to(torch.float64)
    return hi + lo / 2048.0


def f32(x):
    return x.to(torch.float32).to(torch.float64)


class Emu:
    def __init__(self, sd, mode):
        self.sd, self.mode = {k: v.double() for k, v in sd.items()}, mode      # mode: 'fp64' | 'A' | 'B'

    def store(self, x):                       # what the epilogue leaves in memory
        if self.mode == "fp64":
            return x
        x = f32(x)
        return r22(x) if self.mode == "B" else x

    def conv_bn(self, key_c, key_bn, x, pad, relu, res=None, res_after_relu=False, transpose=False):
        sd = self.sd
        w, b = sd[key_c + ".weight"], sd[key_c + ".bias"]
        if key_bn is not None:                # BN folded into the weights on the host, as the engine does (float64 fold -> fp32)
            g = sd[key_bn + ".weight"] / torch.sqrt(sd[key_bn + ".running_var"] + 1e-5)
            shape = (1, -1, 1, 1) if transpose else (-1, 1, 1, 1)
            w = w * g.reshape(shape)
            b = (b - sd[key_bn + ".running_mean"]) * g + sd[key_bn + ".bias"]
        if self.mode != "fp64":
            w, b, x = r22(f32(w)), f32(b), r22(x)
        y = (F.conv_transpose2d(x, w, b, stride=2) if transpose else F.conv2d(x, w, b, padding=pad))
        if res is not None and not res_after_relu:
            y = y + res
        if relu:
            y = F.relu(y)
        if res is not None and res_after_relu:
            y = y + res
        return self.store(y)

    def res_block(self, p, x):
        y = self.conv_bn(p + ".res_branch.0", p + ".res_branch.1", x, 1, True)
        if (p + ".skip_con.0.weight") in self.sd:      # the engine fuses the 1x1 skip conv as extra K-blocks of the second conv
            s = self.conv_bn(p + ".skip_con.0", p + ".skip_con.1", x, 0, False) if self.mode == "fp64" else None
            if s is None:
                xs = r22(x)
                sd = self.sd
                g = sd[p + ".skip_con.1.weight"] / torch.sqrt(sd[p + ".skip_con.1.running_var"] + 1e-5)
                w = r22(f32(sd[p + ".skip_con.0.weight"] * g.reshape(-1, 1, 1, 1)))
                b = f32((sd[p + ".skip_con.0.bias"] - sd[p + ".skip_con.1.running_mean"]) * g + sd[p + ".skip_con.1.bias"])
                s = F.conv2d(xs, w, b)                   # stays in the accumulator: not stored, not rounded
        else:
            s = x
        return self.conv_bn(p + ".res_branch.3", p + ".res_branch.4", y, 1, True, res=s)

    def trunk(self, prefix, x):
        fl, ed = prefix + ".front_layers", prefix + ".encoder_decoder"
        x = self.store(x)
        x = self.conv_bn(fl + ".0.block.0", fl + ".0.block.1", x, 3, True)
        x = self.res_block(fl + ".1", x)
        skip1 = self.res_block(ed + ".skip_res1", x)
        x = self.res_block(ed + ".encoder_res1", F.max_pool2d(x, 2, 2))
        skip2 = self.res_block(ed + ".skip_res2", x)
        x = self.res_block(ed + ".encoder_res2", F.max_pool2d(x, 2, 2))
        x = self.res_block(ed + ".mid_res", x)
        x = self.res_block(ed + ".decoder_res2", x)
        x = self.conv_bn(ed + ".decoder_upsample2.block.0", ed + ".decoder_upsample2.block.1", x, 0, True, res=skip2,
                         res_after_relu=True, transpose=True)
        x = self.res_block(ed + ".decoder_res1", x)
        x = self.conv_bn(ed + ".decoder_upsample1.block.0", ed + ".decoder_upsample1.block.1", x, 0, True, res=skip1,
                         res_after_relu=True, transpose=True)
        return x

    def p2p(self, planes):
        x = self.trunk("joint_net.conv_net", planes)
        old, self.mode = self.mode, ("fp64" if self.mode == "fp64" else "A")       # the last layer always stores fp32
        y = self.conv_bn("joint_net.conv_net.output_layer", None, x, 0, False)
        self.mode = old
        return y


def main():
    for case in ("panoptic_b2", "shelf_crowd"):
        g = Golden(case)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in g.weights.items()}
        planes = torch.from_numpy(np.concatenate([g["b0_planes_keep"].reshape(-1, g.J, 64, 64)])).double()
        ref = Emu(sd, "fp64").p2p(planes)
        with torch.no_grad():
            outs = {"oracle fp32": O.p2p_net({k: v.float() for k, v in sd.items()}, planes.float()).double(),
                    "engine A (fp32 activations)": Emu(sd, "A").p2p(planes), "engine B (hi/lo activations)": Emu(sd, "B").p2p(planes)}
        K = O.JlnConstants(g.cfg)
        beta = float(g.cfg.NETWORK.BETA)

        def joints(feat):                       # soft-argmax positions of every (plane image, joint) in mm, float64
            n = feat.shape[0]
            w = torch.softmax(beta * feat.reshape(n, g.J, -1), dim=2)
            cg = K.center_grid.double()[0]      # [4096, 2] (plane type does not matter for a displacement)
            return torch.einsum("njp,pc->njc", w, cg)
        jr = joints(ref)
        print("== %s: %d plane images, |feat| max %.3f" % (case, planes.shape[0], float(ref.abs().max())))
        for name, o in outs.items():
            print("  %-30s max|feat - fp64| = %.2e   max joint displacement = %.2e mm" % (
                name, float((o - ref).abs().max()), float((joints(o) - jr).abs().max())))


if __name__ == "__main__":
    main()
