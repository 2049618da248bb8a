"""Declarative description of the four small CNNs on the hot path.

The reference builds these as nested ``nn.Module`` classes (``lib/models/cnns_2d.py:12-187``,
``lib/models/cnns_1d.py:10-143``, ``lib/models/weight_net.py:48-67``).  Here they are *tables*:
every convolution is one :class:`Conv` record whose ``key`` is the reference ``state_dict`` prefix
(SURVEY.md §5 "checkpoint": 485 tensors), so that

* ``models/`` can register parameters under exactly the reference's names (strict
  ``load_state_dict`` of a reference ``model_best.pth.tar`` works),
* the weight generator, the CPU oracle and the CUDA engine's parameter table all enumerate the
  same list.

The C library repeats this enumeration in ``csrc/fvp_params.cu``; ``tests/test_netspec.py`` checks
both against each other through ``fvp_param_count`` / ``fvp_param_name``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class Conv:
    key: str          # state_dict prefix of the conv / linear ("....block.0")
    cin: int
    cout: int
    k: int            # kernel extent per spatial dim (1, 2, 3 or 7)
    ndim: int         # 2, 1 (conv) or 0 (linear)
    bn: str = ""      # state_dict prefix of the BatchNorm that follows ("" = none)
    transposed: bool = False


def _res(prefix: str, cin: int, cout: int, nd: int) -> List[Conv]:
    out = [
        Conv(prefix + ".res_branch.0", cin, cout, 3, nd, prefix + ".res_branch.1"),
        Conv(prefix + ".res_branch.3", cout, cout, 3, nd, prefix + ".res_branch.4"),
    ]
    if cin != cout:
        out.append(Conv(prefix + ".skip_con.0", cin, cout, 1, nd, prefix + ".skip_con.1"))
    return out


def trunk(prefix: str, cin: int, nd: int) -> List[Conv]:
    """front_layers + encoder_decoder shared by CenterNet / P2PNet / C2CNet
    (cnns_2d.py:74-129,150-155; cnns_1d.py:71-125)."""
    ed = prefix + ".encoder_decoder"
    L: List[Conv] = [Conv(prefix + ".front_layers.0.block.0", cin, 16, 7, nd, prefix + ".front_layers.0.block.1")]
    L += _res(prefix + ".front_layers.1", 16, 32, nd)
    # registration order of the reference's EncoderDecorder.__init__ (cnns_2d.py:78-92)
    L += _res(ed + ".encoder_res1", 32, 64, nd)
    L += _res(ed + ".encoder_res2", 64, 128, nd)
    L += _res(ed + ".mid_res", 128, 128, nd)
    L += _res(ed + ".decoder_res2", 128, 128, nd)
    L.append(Conv(ed + ".decoder_upsample2.block.0", 128, 64, 2, nd, ed + ".decoder_upsample2.block.1", True))
    L += _res(ed + ".decoder_res1", 64, 64, nd)
    L.append(Conv(ed + ".decoder_upsample1.block.0", 64, 32, 2, nd, ed + ".decoder_upsample1.block.1", True))
    L += _res(ed + ".skip_res1", 32, 32, nd)
    L += _res(ed + ".skip_res2", 64, 64, nd)
    return L


def center_net(J: int, prefix: str = "pose_net.center_net") -> List[Conv]:
    L = trunk(prefix, J, 2)
    L.append(Conv(prefix + ".output_hm.0", 32, 32, 3, 2))
    L.append(Conv(prefix + ".output_hm.2", 32, 1, 1, 2))
    L.append(Conv(prefix + ".output_size.0", 32, 32, 3, 2))
    L.append(Conv(prefix + ".output_size.2", 32, 2, 1, 2))
    return L


def c2c_net(J: int, prefix: str = "pose_net.c2c_net") -> List[Conv]:
    L = trunk(prefix, J, 1)
    L.append(Conv(prefix + ".output_hm", 32, 1, 1, 1))
    return L


def p2p_net(J: int, prefix: str = "joint_net.conv_net") -> List[Conv]:
    L = trunk(prefix, J, 2)
    L.append(Conv(prefix + ".output_layer", 32, J, 1, 2))
    return L


def weight_net(feat: int = 32, hidden: int = 64, prefix: str = "joint_net.weight_net") -> List[Conv]:
    return [
        Conv(prefix + ".heatmap_feature_net.0", 1, feat, 3, 2, prefix + ".heatmap_feature_net.1"),
        Conv(prefix + ".output.0", feat, hidden, 1, 0),
        Conv(prefix + ".output.2", hidden, 1, 1, 0),
    ]


def all_layers(J: int, feat: int = 32, hidden: int = 64) -> List[Conv]:
    return center_net(J) + c2c_net(J) + p2p_net(J) + weight_net(feat, hidden)


def weight_shape(c: Conv) -> Tuple[int, ...]:
    if c.ndim == 0:
        return (c.cout, c.cin)
    if c.transposed:  # ConvTranspose: [in, out, k(,k)]
        return (c.cin, c.cout) + (c.k,) * c.ndim
    return (c.cout, c.cin) + (c.k,) * c.ndim


def param_table(J: int, feat: int = 32, hidden: int = 64) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(state_dict key, shape, dtype)] in the reference's state_dict order per layer."""
    rows: List[Tuple[str, Tuple[int, ...], str]] = []
    for c in all_layers(J, feat, hidden):
        rows.append((c.key + ".weight", weight_shape(c), "float32"))
        rows.append((c.key + ".bias", (c.cout,), "float32"))
        if c.bn:
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                rows.append((c.bn + "." + leaf, (c.cout,), "float32"))
            rows.append((c.bn + ".num_batches_tracked", (), "int64"))
    return rows


def macs_per_image(layers: List[Conv], hw: Tuple[int, ...]) -> int:
    """Multiply-accumulates of one trunk+head at input resolution ``hw``."""
    import math

    full = math.prod(hw)
    half = math.prod(max(1, s // 2) for s in hw)
    quarter = math.prod(max(1, s // 4) for s in hw)
    total = 0
    for c in layers:
        name = c.key
        if "encoder_res2" in name or "mid_res" in name or "decoder_res2" in name:
            px = quarter
        elif "decoder_upsample2" in name:
            px = quarter  # each input pixel feeds k^nd outputs
        elif "encoder_res1" in name or "skip_res2" in name or "decoder_res1" in name:
            px = half
        elif "decoder_upsample1" in name:
            px = half
        else:
            px = full
        total += px * c.cin * c.cout * (c.k ** max(c.ndim, 1) if c.ndim else 1)
    return total


def by_key(layers: List[Conv]) -> Dict[str, Conv]:
    return {c.key: c for c in layers}
