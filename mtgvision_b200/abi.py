"""ctypes binding of libmtgv.so (include/mtgv.h).

The CUDA library is the product: if it is missing or cannot be loaded this module raises,
there is no CPU fallback anywhere in the package.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmtgv.so")

MTGV_FIELD_PHILOX = -1
MTGV_TAPE_MAX_OPS = 18
MTGV_X_MAX_OPS = 16

OUT_F16, OUT_U8, OUT_F32 = 0, 1, 2
KIND_VIRTUAL, KIND_CROPPED, KIND_BG_ONLY = 0, 1, 2

# tape opcodes (mtgv_tape_opcode)
(OP_NONE, OP_DOWNUP, OP_WARP, OP_AFFINE, OP_PERSPECTIVE, OP_TINT, OP_FADE_BLACK, OP_FADE_WHITE, OP_BC, OP_FLIP,
 OP_ROTATE, OP_WARP_INV, OP_BLUR, OP_SHARPEN, OP_NOISE, OP_GAUSS_NOISE, OP_SALT_PEPPER, OP_ERASE, OP_CUTOUT) = range(19)
# expanded opcodes (mtgv_x_opcode)
(X_NONE, X_ELEM, X_DOWNUP, X_WARP_PERSP, X_WARP_AFFINE, X_BLUR3, X_SHARPEN, X_NOISE, X_GAUSS_NOISE, X_SALT_PEPPER,
 X_ERASE, X_CUTOUT) = range(12)


class TapeOp(C.Structure):
    _fields_ = [
        ("code", C.c_int32), ("n_field", C.c_int32), ("n_field2", C.c_int32), ("_pad", C.c_int32),
        ("i", C.c_int32 * 16), ("d", C.c_double * 8), ("field", C.c_int64), ("field2", C.c_int64),
    ]


class EncTape(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("card", C.c_int32), ("swap_choice", C.c_int32), ("bg", C.c_int32),
        ("upsidedown", C.c_int32), ("n_fg", C.c_int32), ("n_bg", C.c_int32), ("n_vrtl", C.c_int32),
        ("seed", C.c_uint64), ("ops", TapeOp * MTGV_TAPE_MAX_OPS),
    ]


class XOp(C.Structure):
    _fields_ = [
        ("code", C.c_int32), ("n_field", C.c_int32), ("n_field2", C.c_int32), ("_pad", C.c_int32),
        ("i", C.c_int32 * 16), ("f", C.c_float * 8), ("d", C.c_double * 9), ("field", C.c_int64),
        ("field2", C.c_int64),
    ]


class EncParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("card", C.c_int32), ("bg", C.c_int32), ("upsidedown", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("card_h", C.c_int32), ("card_w", C.c_int32),
        ("src_y0", C.c_int32), ("src_x0", C.c_int32), ("src_h", C.c_int32), ("src_w", C.c_int32),
        ("fg_rh", C.c_int32), ("fg_rw", C.c_int32), ("fg_y0", C.c_int32), ("fg_x0", C.c_int32),
        ("bg_h", C.c_int32), ("bg_w", C.c_int32), ("flip_h", C.c_int32), ("flip_v", C.c_int32),
        ("rot_nh", C.c_int32), ("rot_nw", C.c_int32),
        ("bg_rh", C.c_int32), ("bg_rw", C.c_int32), ("bg_y0", C.c_int32), ("bg_x0", C.c_int32),
        ("n_fg", C.c_int32), ("n_pre", C.c_int32), ("n_post", C.c_int32), ("n_vrtl", C.c_int32),
        ("rot_inv", C.c_double * 6), ("winv", C.c_double * 9), ("seed", C.c_uint64),
        ("status", C.c_int32), ("_pad", C.c_int32), ("ops", XOp * MTGV_X_MAX_OPS),
    ]


class EncConfig(C.Structure):
    _fields_ = [
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("y_h", C.c_int32), ("y_w", C.c_int32),
        ("target_is_input_prob", C.c_double), ("similar_neg_prob", C.c_double),
        ("half_upsidedown", C.c_int32), ("paired", C.c_int32), ("targets", C.c_int32), ("_pad", C.c_int32),
    ]


# ---- detection path (mtgv_det_* in include/mtgv.h) ----
DET_MAX_CARDS, DET_MAX_ATTEMPTS, DET_MAX_KP, DET_MAX_KPOLY = 32, 10, 8, 3
DET_MAX_PRE, DET_MAX_POST, DET_MAX_CARD_OPS = 4, 8, 3
(PH_NONE, PH_RBC, PH_HSV, PH_GAUSS_NOISE, PH_GAUSS_BLUR, PH_ERASE, PH_ISO_NOISE, PH_SHOT_NOISE, PH_MEDIAN_BLUR,
 PH_MOTION_BLUR, PH_GLASS_BLUR) = range(11)


class PhotoOp(C.Structure):
    _fields_ = [("code", C.c_int32), ("i", C.c_int32 * 5), ("d", C.c_double * 3), ("field", C.c_int64)]


class DetAttempt(C.Structure):
    _fields_ = [("cx", C.c_int32), ("cy", C.c_int32), ("dst_given", C.c_int32), ("_pad", C.c_int32),
                ("deg", C.c_double), ("area", C.c_double), ("jitter", C.c_double * 4), ("dst", C.c_float * 8)]


class DetCard(C.Structure):
    _fields_ = [("card", C.c_int32), ("n_attempts", C.c_int32), ("n_photo", C.c_int32), ("_pad", C.c_int32),
                ("photo", PhotoOp * DET_MAX_CARD_OPS), ("att", DetAttempt * DET_MAX_ATTEMPTS)]


class DetTape(C.Structure):
    _fields_ = [("bg_only", C.c_int32), ("bg", C.c_int32), ("bg_deg", C.c_int32), ("n_cards", C.c_int32),
                ("n_pre", C.c_int32), ("n_post", C.c_int32), ("seed", C.c_uint64),
                ("bg_ab_given", C.c_int32), ("_pad", C.c_int32), ("bg_ab", C.c_double * 2),
                ("pre", PhotoOp * DET_MAX_PRE), ("post", PhotoOp * DET_MAX_POST), ("cards", DetCard * DET_MAX_CARDS)]


class DetConfig(C.Structure):
    _fields_ = [("size_h", C.c_int32), ("size_w", C.c_int32), ("num_cards_min", C.c_int32), ("num_cards_max", C.c_int32),
                ("min_visible", C.c_double), ("min_visible_edges", C.c_double), ("jitter_ratio", C.c_double),
                ("min_area_ratio", C.c_double), ("max_area_ratio", C.c_double), ("ratio_bg", C.c_double),
                ("no_contains", C.c_int32), ("max_attempts", C.c_int32), ("kind", C.c_int32), ("photometrics", C.c_int32),
                ("size_sample_mode", C.c_int32), ("n_bgs_first", C.c_int32), ("bg_first_prob", C.c_double)]


DET_TAPE_DTYPE = np.dtype(DetTape)

assert C.sizeof(TapeOp) == 160, C.sizeof(TapeOp)
assert C.sizeof(XOp) == 200, C.sizeof(XOp)

TAPE_DTYPE = np.dtype(EncTape)
PARAMS_DTYPE = np.dtype(EncParams)
XOP_DTYPE = np.dtype(XOp)


LAYOUT_NHWC, LAYOUT_NCHW = 0, 1


class MtgvError(RuntimeError):
    pass


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libmtgv.so or raise: the package has no other execution path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MTGV_LIB") or LIB_PATH  # MTGV_LIB: kernel-variant experiments (tests/build_variant.sh)
    if not os.path.isfile(p):
        raise MtgvError(
            f"libmtgv.so not found at {p}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  mtgvision_b200 has no CPU fallback."
        )
    lib = C.CDLL(p)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    protos = {
        "mtgv_abi_version": (i32, []),
        "mtgv_create": (vp, [i32]),
        "mtgv_destroy": (None, [vp]),
        "mtgv_last_error": (C.c_char_p, [vp]),
        "mtgv_set_card_pool": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i32]),
        "mtgv_set_bg_pool": (i32, [vp, vp, vp, vp, i32]),
        "mtgv_set_encoder_config": (i32, [vp, vp]),
        "mtgv_sample_encoder_tape": (i32, [vp, u64, i64, i32, vp, vp]),
        "mtgv_sample_encoder_tape_ex": (i32, [vp, u64, i64, i32, vp, vp, C.c_double, C.c_double, vp, vp]),
        "mtgv_update_card_images": (i32, [vp, vp, i32, i32, vp]),
        "mtgv_update_bg_images": (i32, [vp, vp, i32, i32, vp]),
        "mtgv_get_mask": (i32, [vp, i32, vp, vp]),
        "mtgv_expand_params": (i32, [vp, vp, i32, vp, vp, vp]),
        "mtgv_encoder_batch": (i32, [vp, vp, i32, vp, i32, vp, vp]),
        "mtgv_encoder_targets": (i32, [vp, vp, i32, vp, i32, vp]),
        "mtgv_warp_perspective": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp]),
        "mtgv_run_plane_ops": (i32, [vp, vp, i32, i32, i32, i32, vp, i32, vp, u64, vp]),
        "mtgv_extract_dewarped": (i32, [vp, vp, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]),
        "mtgv_jpeg_info": (i32, [vp, vp, i64, vp]),
        "mtgv_jpeg_info_batch": (i32, [vp, vp, vp, i32, vp]),
        "mtgv_gather_files": (i32, [vp, vp, vp, i32, vp, i64, vp]),
        "mtgv_jpeg_last_kernel_ms": (i32, [vp, vp]),
        "mtgv_encode_jpeg_batch": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, i64, vp, vp]),
        "mtgv_jpeg_encode_last_kernel_ms": (i32, [vp, vp]),
        "mtgv_compact_jpeg_files": (i32, [vp, vp, i64, vp, i32, vp, vp, vp]),
        "mtgv_decode_jpeg_batch": (i32, [vp, vp, vp, i32, vp, vp, vp, vp]),
        "mtgv_decode_jpeg_to_pools": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "mtgv_launch_count": (i64, [vp]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _DET_PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mtgv_abi_version() != 1:
        raise MtgvError("libmtgv.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


_vp, _i32 = C.c_void_p, C.c_int
_DET_PROTOS: dict = {
    "mtgv_set_det_config": (_i32, [_vp, _vp]),
    "mtgv_sample_det_tape": (_i32, [_vp, C.c_uint64, C.c_int64, _i32, _vp, _vp]),
    "mtgv_det_place": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mtgv_det_params_size": (_i32, []),
    "mtgv_det_batch": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, _vp]),
}


def declared_symbols() -> list[str]:
    """Every `mtgv_*` function declared in include/mtgv.h (used by the symbol-export test)."""
    import re

    hdr = os.path.join(os.path.dirname(_HERE), "include", "mtgv.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mtgv_[a-z0-9_]+)\s*\(", txt)))
