"""Attribute-dict configuration accepted by ``models.faster_voxelpose.get(cfg)``.

The reference hands its model an ``easydict`` built by ``lib/core/config.py:15-144`` and
overridden from YAML by ``update_config`` (``lib/core/config.py:174-188``).  The hot path only
*reads attributes* (SURVEY.md §5 "config / flags"), so any object with the same attribute tree
works.  This module provides

* :class:`AttrDict` - a minimal recursive attribute dict (no ``easydict`` dependency),
* :func:`preset` - the geometry of the three shipped configs (``configs/{panoptic,campus,shelf}/jln64.yaml``)
  plus the two synthetic benchmark variants named in BASELINE.json,
* :func:`load_yaml` - merge a reference-format YAML file onto the defaults.  Like the reference
  (``lib/core/config.py:168-171,187-188``) unknown keys raise ``ValueError``.
"""
from __future__ import annotations

import copy
from typing import Any, Dict


class AttrDict(dict):
    """dict whose items are also attributes, recursively."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - attribute protocol
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_DEFAULTS: Dict[str, Any] = {
    "CUDNN": {"BENCHMARK": True, "DETERMINISTIC": False, "ENABLED": True},
    "BACKBONE": "resnet",
    "MODEL": "faster_voxelpose",
    "DEVICE": "cuda:0",
    "WORKERS": 8,
    "PRINT_FREQ": 100,
    "OUTPUT_DIR": "output",
    "LOG_DIR": "log",
    "DATASET": {
        "DATADIR": "",
        "COLOR_RGB": False,
        "DATA_AUGMENTATION": False,
        "TRAIN_DATASET": "panoptic",
        "TRAIN_HEATMAP_SRC": "image",
        "TEST_DATASET": "panoptic",
        "TEST_HEATMAP_SRC": "image",
        "CAMERA_NUM": 5,
        "ORI_IMAGE_SIZE": [1920, 1080],
        "IMAGE_SIZE": [960, 512],
        "HEATMAP_SIZE": [240, 128],
        "NUM_JOINTS": 15,
        "ROOT_JOINT_ID": 2,
    },
    "SYNTHETIC": {
        "CAMERA_FILE": "",
        "POSE_FILE": "",
        "MAX_PEOPLE": 10,
        "NUM_DATA": 10000,
        "DATA_AUGMENTATION": True,
    },
    "NETWORK": {
        "PRETRAINED_BACKBONE": "",
        "NUM_CHANNEL_JOINT_FEAT": 32,
        "NUM_CHANNEL_JOINT_HIDDEN": 64,
        "SIGMA": 3,
        "BETA": 100.0,
    },
    "RESNET": {                       # lib/core/config.py:101-108 (pose_resnet backbone)
        "NUM_LAYERS": 50,
        "DECONV_WITH_BIAS": False,
        "NUM_DECONV_LAYERS": 3,
        "NUM_DECONV_FILTERS": [256, 256, 256],
        "NUM_DECONV_KERNELS": [4, 4, 4],
        "FINAL_CONV_KERNEL": 1,
    },
    "TRAIN": {
        "BATCH_SIZE": 8,
        "SHUFFLE": True,
        "BEGIN_EPOCH": 0,
        "END_EPOCH": 20,
        "RESUME": False,
        "OPTIMIZER": "adam",
        "LR": 1e-4,
        "LAMBDA_LOSS_2D": 1.0,
        "LAMBDA_LOSS_1D": 1.0,
        "LAMBDA_LOSS_BBOX": 0.1,
        "LAMBDA_LOSS_FUSED": 5.0,
        "VISUALIZATION": False,
        "VIS_TYPE": ["2d_planes"],
    },
    "TEST": {
        "MODEL_FILE": "model_best.pth.tar",
        "BATCH_SIZE": 8,
        "VISUALIZATION": False,
        "VIS_TYPE": ["2d_planes"],
    },
    "CAPTURE_SPEC": {
        "SPACE_SIZE": [8000.0, 8000.0, 2000.0],
        "SPACE_CENTER": [0.0, -500.0, 800.0],
        "VOXELS_PER_AXIS": [80, 80, 20],
        "MAX_PEOPLE": 10,
        "MIN_SCORE": 0.3,
    },
    "INDIVIDUAL_SPEC": {
        "SPACE_SIZE": [2000.0, 2000.0, 2000.0],
        "VOXELS_PER_AXIS": [64, 64, 64],
    },
}

# Geometry-only overrides of the shipped YAMLs (values are data, restated from
# configs/panoptic/jln64.yaml:20-33,61-86, configs/campus/jln64.yaml, configs/shelf/jln64.yaml)
_PRESETS: Dict[str, Dict[str, Any]] = {
    "panoptic": {},
    # BASELINE.json north_star: "synthetic 5-view 256x192 heatmaps into an 80x80x20 grid"
    # (W=256, H=192, stride 4 -> IMAGE_SIZE 1024x768)
    "panoptic_256x192": {
        "DATASET": {"IMAGE_SIZE": [1024, 768], "HEATMAP_SIZE": [256, 192]},
    },
    "campus": {
        "DATASET": {
            "CAMERA_NUM": 3,
            "ORI_IMAGE_SIZE": [360, 288],
            "IMAGE_SIZE": [800, 640],
            "HEATMAP_SIZE": [200, 160],
            "NUM_JOINTS": 17,
            "ROOT_JOINT_ID": [11, 12],
            "TEST_DATASET": "campus",
            "TEST_HEATMAP_SRC": "pred",
        },
        "NETWORK": {"SIGMA": 4},
        "TEST": {"BATCH_SIZE": 16},
        "CAPTURE_SPEC": {
            "SPACE_SIZE": [12000.0, 12000.0, 2000.0],
            "SPACE_CENTER": [3000.0, 4500.0, 1000.0],
            "MAX_PEOPLE": 5,
            "MIN_SCORE": 0.1,
        },
    },
    "shelf": {
        "DATASET": {
            "CAMERA_NUM": 5,
            "ORI_IMAGE_SIZE": [1032, 776],
            "IMAGE_SIZE": [800, 608],
            "HEATMAP_SIZE": [200, 152],
            "NUM_JOINTS": 17,
            "ROOT_JOINT_ID": [11, 12],
            "TEST_DATASET": "shelf",
            "TEST_HEATMAP_SRC": "pred",
        },
        "TEST": {"BATCH_SIZE": 16},
        "CAPTURE_SPEC": {
            "SPACE_SIZE": [8000.0, 8000.0, 2000.0],
            "SPACE_CENTER": [450.0, -320.0, 800.0],
            "MAX_PEOPLE": 10,
            "MIN_SCORE": 0.1,
        },
    },
    # BASELINE.json configs[4]: 8-view synthetic ring, 160x160x40 grid, 256x192 heatmaps
    "ring8_160": {
        "DATASET": {"CAMERA_NUM": 8, "IMAGE_SIZE": [1024, 768], "HEATMAP_SIZE": [256, 192]},
        "CAPTURE_SPEC": {"VOXELS_PER_AXIS": [160, 160, 40]},
    },
}


def _merge(dst: AttrDict, src: Dict[str, Any], path: str = "") -> None:
    for k, v in src.items():
        if k not in dst:
            raise ValueError("{}{} not exist in config".format(path, k))
        if isinstance(v, dict):
            if not isinstance(dst[k], dict):
                raise ValueError("{}{} is not a section".format(path, k))
            _merge(dst[k], v, path + k + ".")
        else:
            dst[k] = v


def defaults() -> AttrDict:
    return AttrDict(copy.deepcopy(_DEFAULTS))


def preset(name: str, **overrides) -> AttrDict:
    """Config for one of: panoptic, panoptic_256x192, campus, shelf, ring8_160."""
    if name not in _PRESETS:
        raise ValueError("unknown preset %r (have %s)" % (name, sorted(_PRESETS)))
    cfg = defaults()
    _merge(cfg, _PRESETS[name])
    _merge(cfg, overrides)
    return cfg


def load_yaml(path: str) -> AttrDict:
    """Merge a reference-format YAML (e.g. configs/panoptic/jln64.yaml) onto the defaults."""
    import yaml

    with open(path) as f:
        doc = yaml.safe_load(f)
    cfg = defaults()
    _merge(cfg, doc or {})
    return cfg
