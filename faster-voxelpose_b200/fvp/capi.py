"""ctypes binding of libfvp_b200.so (include/fvp_b200.h).

There is deliberately no fallback: if the shared library is missing or does not export every
symbol the header declares, importing the engine fails loudly (``FvpLibraryError``).  The product
path never routes through PyTorch ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # .../faster-voxelpose_b200
LIB_NAME = "libfvp_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)

FVP_OK, FVP_E_INVALID, FVP_E_CUDA, FVP_E_STATE, FVP_E_NOTFOUND, FVP_E_CALIB, FVP_E_RANGE = 0, -1, -2, -3, -4, -5, -6
ABI_VERSION = 1


class FvpLibraryError(ImportError):
    pass


class FvpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libfvp_b200 error %d: %s" % (code, msg))
        self.code = code


class FvpConfig(C.Structure):
    _fields_ = [
        ("num_views", C.c_int32), ("num_joints", C.c_int32), ("hm_w", C.c_int32), ("hm_h", C.c_int32),
        ("image_w", C.c_float), ("image_h", C.c_float), ("ori_w", C.c_float), ("ori_h", C.c_float),
        ("space_size", C.c_float * 3), ("space_center", C.c_float * 3), ("voxels", C.c_int32 * 3),
        ("ind_space_size", C.c_float * 3), ("ind_voxels", C.c_int32 * 3),
        ("max_people", C.c_int32), ("min_score", C.c_float), ("beta", C.c_float),
        ("feat_channels", C.c_int32), ("hidden_channels", C.c_int32),
        ("max_batch", C.c_int32), ("max_sequences", C.c_int32),
    ]


_P = C.c_void_p      # device / host pointers travel as integers
_CTX = C.c_void_p

# name -> (restype, argtypes): every symbol include/fvp_b200.h declares
SYMBOLS = {
    "fvp_create": (C.c_int, [C.POINTER(FvpConfig), C.c_int, C.POINTER(_CTX)]),
    "fvp_create_lane": (C.c_int, [_CTX, C.c_int, C.POINTER(_CTX)]),
    "fvp_destroy": (None, [_CTX]),
    "fvp_last_error": (C.c_char_p, [_CTX]),
    "fvp_abi_version": (C.c_int, []),
    "fvp_param_count": (C.c_int, [_CTX]),
    "fvp_param_name": (C.c_char_p, [_CTX, C.c_int]),
    "fvp_param_numel": (C.c_int64, [_CTX, C.c_int]),
    "fvp_set_param": (C.c_int, [_CTX, C.c_char_p, _P, C.c_int64]),
    "fvp_finalize_params": (C.c_int, [_CTX]),
    "fvp_set_axes": (C.c_int, [_CTX, _P, _P, _P]),
    "fvp_fine_voxels": (C.c_int, [_CTX, C.POINTER(C.c_int32 * 3)]),
    "fvp_set_sequence": (C.c_int, [_CTX, C.c_int, _P, C.c_int, _P]),
    "fvp_forward": (C.c_int, [_CTX, _P, C.c_int, _P, _P, _P, _P, C.c_size_t]),
    "fvp_forward_host": (C.c_int, [_CTX, _P, C.c_int, _P, _P, _P, _P, C.c_size_t]),
    "fvp_submit_host": (C.c_int, [_CTX, _P, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_longlong)]),
    "fvp_wait": (C.c_int, [_CTX, C.c_longlong]),
    "fvp_use_cuda_graph": (C.c_int, [_CTX, C.c_int]),
    "fvp_debug_conv": (C.c_int, [_CTX, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_float), C.c_size_t]),
    "fvp_debug_project": (C.c_int, [_CTX, C.c_int, _P, C.c_int, _P, _P, C.c_size_t]),
    "fvp_debug_conv_plan": (C.c_int, [C.c_int] * 9 + [C.POINTER(C.c_int)]),
    "fvp_debug_pack_tc16": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_longlong, C.POINTER(C.c_longlong)]),
    "fvp_stage_heatmaps": (C.c_int, [_CTX, _P, C.c_int, C.c_size_t]),
    "fvp_hdn_project": (C.c_int, [_CTX, C.c_int, _P, _P, C.c_size_t]),
    "fvp_center_net": (C.c_int, [_CTX, _P, C.c_int, _P, _P, C.c_size_t]),
    "fvp_nms_topk": (C.c_int, [_CTX, _P, C.c_int, _P, _P, C.c_size_t]),
    "fvp_proposals": (C.c_int, [_CTX, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_size_t]),
    "fvp_jln_project": (C.c_int, [_CTX, C.c_int, _P, _P, _P, _P, C.c_size_t]),
    "fvp_p2p_net": (C.c_int, [_CTX, _P, C.c_int, _P, _P, C.c_size_t]),
    "fvp_pose_head": (C.c_int, [_CTX, _P, _P, C.c_int, _P, _P, _P, _P, C.c_size_t]),
    "fvp_c2c_net": (C.c_int, [_CTX, _P, C.c_int, _P, C.c_size_t]),
    "fvp_render_heatmaps": (C.c_int, [_CTX, _P, _P, _P, C.c_int, C.c_int, C.c_double, _P, C.c_size_t]),
    "fvp_backbone_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_CTX)]),
    "fvp_backbone_forward": (C.c_int, [_CTX, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    "fvp_backbone_num_stages": (C.c_int, [_CTX]),
    "fvp_backbone_destroy": (None, [_CTX]),
    "fvp_backbone_last_error": (C.c_char_p, [_CTX]),
    "fvp_backbone_set_param": (C.c_int, [_CTX, C.c_char_p, _P, C.c_int64]),
    "fvp_backbone_finalize": (C.c_int, [_CTX]),
    "fvp_backbone_forward_slice": (C.c_int, [_CTX, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t]),
    "fvp_set_conv_mode": (C.c_int, [_CTX, C.c_int]),
    "fvp_set_latency_mode": (C.c_int, [_CTX, C.c_int]),
    "fvp_check_range": (C.c_int, [_CTX]),
    "fvp_fp16_fallback_layers": (C.c_int, [_CTX]),
    "fvp_last_launch_count": (C.c_int, [_CTX]),
    "fvp_set_profiling": (C.c_int, [_CTX, C.c_int]),
    "fvp_stage_times_ms": (C.c_int, [_CTX, C.POINTER(C.c_float * 9)]),
    "fvp_algorithmic_bytes": (C.c_int, [_CTX, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the library and bind every declared symbol (no compute, no GPU needed)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FVP_B200_LIB", LIB_PATH)
    if not os.path.isfile(p):
        raise FvpLibraryError(
            "%s not found at %s - build it with `python __graft_entry__.py` (or "
            "faster-voxelpose_b200/csrc/build.sh); there is no CPU / PyTorch fallback" % (LIB_NAME, p))
    try:
        lib = C.CDLL(p)
    except OSError as e:  # pragma: no cover
        raise FvpLibraryError("cannot load %s: %s" % (p, e)) from e
    lax = path is not None or "FVP_B200_LIB" in os.environ          # tooling only (A/B of older builds): tolerate missing symbols
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            if lax and os.environ.get("FVP_B200_LAX_SYMBOLS") == "1":
                continue
            raise FvpLibraryError("%s does not export %s" % (p, name)) from e
        fn.restype = res
        fn.argtypes = args
    if lib.fvp_abi_version() != ABI_VERSION:
        raise FvpLibraryError("ABI mismatch: library %d, binding %d" % (lib.fvp_abi_version(), ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, ctx, rc: int) -> None:
    if rc == FVP_OK:
        return
    msg = lib.fvp_last_error(ctx)
    msg = msg.decode() if msg else "?"
    if rc == FVP_E_CALIB:
        # the reference asserts on calibration problems (lib/models/project_whole.py:73-74)
        raise AssertionError(msg)
    raise FvpError(rc, msg)
