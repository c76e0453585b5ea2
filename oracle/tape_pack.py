"""TEST INFRASTRUCTURE ONLY - pack the oracle's recorded tapes (plain dicts, see
oracle/encoder_oracle.py) into the C ABI structs of include/mtgv.h so the CUDA path can
be driven with exactly the parameters the reference/oracle run sampled.
"""

from __future__ import annotations

import numpy as np

from mtgvision_b200 import abi


class FieldBuffer:
    """Concatenation of injected random fields as 4-byte words."""

    def __init__(self):
        self.chunks: list[np.ndarray] = []
        self.words = 0

    def add(self, arr: np.ndarray) -> int:
        a = np.ascontiguousarray(arr)
        assert a.dtype in (np.float32, np.int32), a.dtype
        off = self.words
        self.chunks.append(a.reshape(-1).view(np.uint32))
        self.words += a.size
        return off

    def to_array(self) -> np.ndarray:
        if not self.chunks:
            return np.zeros(1, dtype=np.uint32)
        return np.concatenate(self.chunks)


def _pack_op(rec: dict, op, fields: FieldBuffer, host_transcendentals: bool):
    name = rec["op"]
    op["field"] = op["field2"] = abi.MTGV_FIELD_PHILOX
    i, d = op["i"], op["d"]
    if name == "downup":
        op["code"] = abi.OP_DOWNUP
        i[0], i[1], i[2] = rec["n"], rec["down"], rec["up"]
    elif name in ("warp", "warp_inv", "perspective"):
        op["code"] = {"warp": abi.OP_WARP, "warp_inv": abi.OP_WARP_INV, "perspective": abi.OP_PERSPECTIVE}[name]
        d[:8] = np.asarray(rec["u"], dtype=np.float64).reshape(-1)
    elif name == "affine":
        op["code"] = abi.OP_AFFINE
        d[0], d[1], d[2], d[3], d[4] = rec["angle"], rec["tx"], rec["ty"], rec["scale"], rec["shear"]
        if host_transcendentals:
            i[0] = 1
            d[5], d[6] = rec["alpha"], rec["beta"]
    elif name == "tint":
        op["code"] = abi.OP_TINT
        d[:3] = rec["u"]
    elif name == "fade_black":
        op["code"] = abi.OP_FADE_BLACK
        d[0] = rec["u"]
    elif name == "fade_white":
        op["code"] = abi.OP_FADE_WHITE
        d[0] = rec["u"]
    elif name == "bc":
        op["code"] = abi.OP_BC
        d[0], d[1] = rec["uc"], rec["ub"]
    elif name == "flip":
        op["code"] = abi.OP_FLIP
        i[0], i[1] = int(rec["horr"]), int(rec["vert"])
    elif name == "rotate":
        op["code"] = abi.OP_ROTATE
        d[0] = rec["u"]
        if host_transcendentals:
            i[0] = 1
            d[1], d[2] = rec["alpha"], rec["beta"]
    elif name == "blur":
        op["code"] = abi.OP_BLUR
        i[0] = rec["n"]
    elif name == "sharpen":
        op["code"] = abi.OP_SHARPEN
    elif name == "noise":
        op["code"] = abi.OP_NOISE
        i[0] = rec["kind"]
        d[0] = rec["u"]
        if rec["kind"] == 2:
            op["field"] = fields.add(rec["salt"].astype(np.int32))
            op["field2"] = fields.add(rec["pepper"].astype(np.int32))
            op["n_field"], op["n_field2"] = len(rec["salt"]), len(rec["pepper"])
        else:
            op["field"] = fields.add(rec["field"].astype(np.float32))
    elif name == "gaussian_noise":
        op["code"] = abi.OP_GAUSS_NOISE
        op["field"] = fields.add(rec["field"].astype(np.float32))
    elif name == "salt_pepper":
        op["code"] = abi.OP_SALT_PEPPER
        op["field"] = fields.add(rec["salt"].astype(np.int32))
        op["field2"] = fields.add(rec["pepper"].astype(np.int32))
        op["n_field"], op["n_field2"] = len(rec["salt"]), len(rec["pepper"])
    elif name == "erase":
        op["code"] = abi.OP_ERASE
        d[0], d[1], d[2] = rec["u_scale"], rec["u_aspect"], rec["u_flip"]
        i[0] = 2 if rec["active"] else (1 if "rect" in rec else 0)
        i[1], i[2] = rec.get("cx", 0), rec.get("cy", 0)
        i[3] = rec.get("color", 0)
        if "c" in rec:
            d[3:6] = rec["c"]
        if "field" in rec:
            op["field"] = fields.add(rec["field"].astype(np.float32))
        i[4], i[5], i[6] = 1, rec["bw"], rec["bh"]
    elif name == "cutout":
        op["code"] = abi.OP_CUTOUT
        i[:16] = np.asarray(rec["holes"], dtype=np.int32).reshape(-1)
    else:
        raise KeyError(name)


def pack_tapes(tapes: list[dict], fields: FieldBuffer | None = None, host_transcendentals: bool = True,
               seed: int = 0):
    """tapes: list of per-sample dicts from the oracle.  Each needs `card`, `bg` (pool
    indices; for x2 tapes `base_card`/`swapped` may resolve the hard negative on device).
    Returns (structured ndarray of mtgv_enc_tape, FieldBuffer)."""
    fields = fields or FieldBuffer()
    arr = np.zeros(len(tapes), dtype=abi.TAPE_DTYPE)
    for s, t in enumerate(tapes):
        e = arr[s]
        e["kind"] = abi.KIND_VIRTUAL if t["kind"] == "virtual" else abi.KIND_CROPPED
        e["card"] = t.get("device_card", t["card"])
        e["swap_choice"] = t.get("swap_choice", -1)
        e["bg"] = t.get("bg", 0)
        e["upsidedown"] = int(t.get("upsidedown", False))
        e["seed"] = seed + s
        ops = []
        if t["kind"] == "virtual":
            ops = list(t["fg_ops"]) + list(t["bg_ops"]) + list(t["vrtl_ops"])
            e["n_fg"], e["n_bg"], e["n_vrtl"] = len(t["fg_ops"]), len(t["bg_ops"]), len(t["vrtl_ops"])
        assert len(ops) <= abi.MTGV_TAPE_MAX_OPS
        for k, rec in enumerate(ops):
            _pack_op(rec, e["ops"][k], fields, host_transcendentals)
    return arr, fields


# --------------------------------------------------------------------------- #
# detection tapes (oracle/det_oracle.py records -> mtgv_det_tape)              #
# --------------------------------------------------------------------------- #


def _pack_photo(rec: dict, op, fields: FieldBuffer):
    from oracle import det_oracle as DO

    op["field"] = abi.MTGV_FIELD_PHILOX
    ph = rec["ph"]
    if ph == DO.PH_RBC:
        op["code"] = abi.PH_RBC
        op["d"][0], op["d"][1] = rec["alpha"], rec["beta"]
    elif ph == DO.PH_HSV:
        op["code"] = abi.PH_HSV
        op["d"][0], op["d"][1], op["d"][2] = rec["hue"], rec["sat"], rec["val"]
    elif ph == DO.PH_GAUSS_NOISE:
        op["code"] = abi.PH_GAUSS_NOISE
        op["d"][0] = rec["sigma"]
        op["field"] = fields.add(rec["field"].astype(np.float32))
    elif ph == DO.PH_GAUSS_BLUR:
        op["code"] = abi.PH_GAUSS_BLUR
        op["d"][0] = rec["sigma"]
    elif ph == DO.PH_ERASE:
        op["code"] = abi.PH_ERASE
        if rec["active"]:
            op["i"][0], op["i"][1], op["i"][2], op["i"][3] = rec["top"], rec["left"], rec["eh"], rec["ew"]
        op["i"][4] = rec["fill"]
        if "color" in rec:
            op["d"][:3] = rec["color"]
        if "field" in rec:
            op["field"] = fields.add(rec["field"].astype(np.float32))
    elif ph == DO.PH_ISO_NOISE:
        op["code"] = abi.PH_ISO_NOISE
        op["d"][0], op["d"][1] = rec["color_shift"], rec["intensity"]
        if "lum" in rec:  # drawn when the oracle applied the op: [H,W,2] = (Poisson luminance counts, unit normals)
            op["field"] = fields.add(np.stack([rec["lum"], rec["col"]], axis=-1).astype(np.float32))
    elif ph == DO.PH_SHOT_NOISE:
        op["code"] = abi.PH_SHOT_NOISE
        op["d"][0] = rec["scale"]
        if "field" in rec:
            op["field"] = fields.add(rec["field"].astype(np.float32))
    elif ph == DO.PH_MEDIAN_BLUR:
        op["code"] = abi.PH_MEDIAN_BLUR
        op["i"][0] = rec["ksize"]
    elif ph == DO.PH_MOTION_BLUR:
        op["code"] = abi.PH_MOTION_BLUR
        op["i"][0] = rec["ksize"]
        bits = np.flatnonzero(rec["mask"].reshape(-1))
        words = [0, 0, 0, 0]
        for b in bits:
            words[b >> 5] |= 1 << (b & 31)
        for k in range(4):
            op["i"][1 + k] = words[k] - (1 << 32) if words[k] >= (1 << 31) else words[k]
    elif ph == DO.PH_GLASS_BLUR:
        op["code"] = abi.PH_GLASS_BLUR
        op["d"][0] = rec["sigma"]
        op["i"][0], op["i"][1] = rec["max_delta"], rec["iterations"]
        op["field"] = fields.add(np.ascontiguousarray(rec["dxy"], dtype=np.int32))
    else:
        raise KeyError(ph)


def pack_det_tapes(tapes: list[dict], fields: FieldBuffer | None = None, host_transcendentals: bool = True, seed: int = 0):
    """Scene tapes recorded by oracle.det_oracle.DetOracle -> (ndarray of mtgv_det_tape, FieldBuffer)."""
    import math

    fields = fields or FieldBuffer()
    arr = np.zeros(len(tapes), dtype=abi.DET_TAPE_DTYPE)
    for s, t in enumerate(tapes):
        e = arr[s]
        e["bg_only"] = int(t["bg_only"])
        e["bg"], e["bg_deg"], e["n_cards"] = t["bg"], t["bg_deg"], t["n_cards"]
        e["seed"] = seed + s
        if host_transcendentals:
            # getRotationMatrix2D's alpha/beta as cv2 computes them with the host libm (od_datasets.py:106)
            h, w = t["bg_hw"]
            S = t["size_hw"]
            scale = math.hypot(S[0] / max(S), S[1] / max(S)) * max(S) / min(h, w)
            a = t["bg_deg"] * (math.pi / 180.0)
            e["bg_ab_given"] = 1
            e["bg_ab"][0], e["bg_ab"][1] = math.cos(a) * scale, math.sin(a) * scale
        for name, cap in (("pre", abi.DET_MAX_PRE), ("post", abi.DET_MAX_POST)):
            assert len(t[name]) <= cap
            e["n_" + name] = len(t[name])
            for k, rec in enumerate(t[name]):
                _pack_photo(rec, e[name][k], fields)
        for ci, c in enumerate(t["cards"]):
            cc = e["cards"][ci]
            cc["card"] = c["card"]
            cc["n_attempts"] = len(c["attempts"])
            for ai, a in enumerate(c["attempts"]):
                at = cc["att"][ai]
                at["cx"], at["cy"], at["deg"], at["area"] = a["cx"], a["cy"], a["deg"], a["area"]
                at["jitter"][:] = a["jitter"]
                if host_transcendentals:
                    at["dst_given"] = 1
                    at["dst"][:] = np.asarray(a["dst"], dtype=np.float32).reshape(-1)
            ops = c.get("photo", [])
            cc["n_photo"] = len(ops)
            for k, rec in enumerate(ops):
                _pack_photo(rec, cc["photo"][k], fields)
    return arr, fields
