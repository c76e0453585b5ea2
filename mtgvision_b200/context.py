"""Thin object wrapper over the C ABI: one `Context` per GPU.

PyTorch is plumbing here (device memory, streams); every pixel and label is produced by
libmtgv.so.  All tensors returned live on the context's device.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi

_TORCH_OUT = {abi.OUT_F16: torch.float16, abi.OUT_U8: torch.uint8, abi.OUT_F32: torch.float32}


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


class Context:
    def __init__(self, device: int | torch.device | None = None):
        if not torch.cuda.is_available():
            raise abi.MtgvError("mtgvision_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = abi.load_library()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self._h = self.lib.mtgv_create(self.device.index)
        if not self._h:
            raise abi.MtgvError(f"mtgv_create({self.device.index}) failed")
        self.cfg: abi.EncConfig | None = None
        self.n_cards = 0
        self.n_bgs = 0
        self.card_hw = (0, 0)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.mtgv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.mtgv_last_error(self._h)
            raise abi.MtgvError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.mtgv_launch_count(self._h))

    # ------------------------------------------------------------------ pools
    def set_card_pool(self, images, labels3, grp_off, grp_mem):
        """images: (N,H,W,3) uint8 numpy array or CUDA tensor."""
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images))
        if images.dtype != torch.uint8 or images.ndim != 4 or images.shape[-1] != 3:
            raise ValueError("card images must be (N,H,W,3) uint8")
        n, h, w, _ = images.shape
        # ingest in chunks so the HWC staging copy stays small next to the planar pool
        dev = images.to(self.device, non_blocking=False).contiguous()
        lab = torch.as_tensor(np.ascontiguousarray(labels3, dtype=np.int32)).to(self.device)
        off = torch.as_tensor(np.ascontiguousarray(grp_off, dtype=np.int32)).to(self.device)
        mem = torch.as_tensor(np.ascontiguousarray(grp_mem, dtype=np.int32)).to(self.device)
        if lab.shape != (n, 3) or off.numel() != n + 1:
            raise ValueError("labels3 must be (N,3) and grp_off (N+1,)")
        rc = self.lib.mtgv_set_card_pool(self._h, _ptr(dev), n, h, w, _ptr(lab), _ptr(off), _ptr(mem), mem.numel())
        self._check(rc, "mtgv_set_card_pool")
        self.n_cards, self.card_hw = n, (h, w)

    def set_bg_pool(self, images):
        """images: list of (h,w,3) uint8 numpy arrays (sizes may differ)."""
        hw = np.asarray([im.shape[:2] for im in images], dtype=np.int32)
        sizes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        flat = np.concatenate([np.ascontiguousarray(im, dtype=np.uint8).reshape(-1) for im in images])
        dev = torch.from_numpy(flat).to(self.device)
        rc = self.lib.mtgv_set_bg_pool(self._h, _ptr(dev), offsets.ctypes.data_as(C.c_void_p),
                                       hw.ctypes.data_as(C.c_void_p), len(images))
        self._check(rc, "mtgv_set_bg_pool")
        self.n_bgs = len(images)
        self.bg_hw = hw

    def set_encoder_config(self, *, x_size_hw=(192, 128), y_size_hw=(192, 128), target_is_input_prob=0.05,
                           similar_neg_prob=0.2, half_upsidedown=False, paired=True, targets=False):
        cfg = abi.EncConfig(int(x_size_hw[0]), int(x_size_hw[1]), int(y_size_hw[0]), int(y_size_hw[1]),
                            float(target_is_input_prob), float(similar_neg_prob), int(half_upsidedown), int(paired),
                            int(targets), 0)
        self._check(self.lib.mtgv_set_encoder_config(self._h, C.byref(cfg)), "mtgv_set_encoder_config")
        self.cfg = cfg

    # ------------------------------------------------------------------ encoder path
    def update_card_images(self, images: torch.Tensor, first: int = 0):
        """Overwrite pool cards [first, first+n) with (n,H,W,3) uint8 images already on the device."""
        assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous()
        assert tuple(images.shape[1:]) == (*self.card_hw, 3)
        rc = self.lib.mtgv_update_card_images(self._h, _ptr(images), first, images.shape[0], self._stream())
        self._check(rc, "mtgv_update_card_images")

    def update_bg_images(self, images: torch.Tensor, first: int = 0):
        assert images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous() and images.ndim == 4
        rc = self.lib.mtgv_update_bg_images(self._h, _ptr(images), first, images.shape[0], self._stream())
        self._check(rc, "mtgv_update_bg_images")

    def sample_encoder_tape(self, seed: int, first_index: int, n_pairs: int, cards: torch.Tensor | None = None,
                            bgs: torch.Tensor | None = None, target_is_input_prob: float | None = None, similar_neg_prob: float | None = None,
                            out: torch.Tensor | None = None) -> torch.Tensor:
        """Tapes of n_pairs x-samples followed (when paired) by n_pairs x2-samples."""
        n = n_pairs * (2 if self.cfg.paired else 1)
        tape = out if out is not None else torch.zeros((n, abi.TAPE_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        if cards is not None:
            cards = cards.to(self.device, dtype=torch.int32).contiguous()
            assert cards.numel() == n_pairs
        if bgs is not None:
            bgs = bgs.to(self.device, dtype=torch.int32).contiguous()
            assert bgs.numel() == n_pairs
        rc = self.lib.mtgv_sample_encoder_tape_ex(
            self._h, C.c_uint64(seed & (2**64 - 1)), C.c_int64(first_index), n_pairs, _ptr(cards), _ptr(bgs),
            -1.0 if target_is_input_prob is None else float(target_is_input_prob),
            -1.0 if similar_neg_prob is None else float(similar_neg_prob), _ptr(tape), self._stream())
        self._check(rc, "mtgv_sample_encoder_tape_ex")
        return tape

    def get_mask(self, which: str = "encoder") -> torch.Tensor:
        out = torch.empty(self.card_hw, dtype=torch.float32, device=self.device)
        self._check(self.lib.mtgv_get_mask(self._h, 0 if which == "encoder" else 1, _ptr(out), self._stream()), "mtgv_get_mask")
        return out

    def upload_tape(self, tape_np: np.ndarray) -> torch.Tensor:
        assert tape_np.dtype == abi.TAPE_DTYPE
        return torch.from_numpy(tape_np.view(np.uint8).reshape(len(tape_np), -1).copy()).to(self.device)

    def expand_params(self, tape: torch.Tensor, want_labels: bool = True, params: torch.Tensor | None = None,
                      labels: torch.Tensor | None = None):
        n = tape.shape[0]
        if params is None:
            params = torch.empty((n, abi.PARAMS_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        if labels is None and want_labels:
            labels = torch.empty((n, 3), dtype=torch.int64, device=self.device)
        rc = self.lib.mtgv_expand_params(self._h, _ptr(tape), n, _ptr(params), _ptr(labels), self._stream())
        self._check(rc, "mtgv_expand_params")
        return params, labels

    def encoder_batch(self, params: torch.Tensor, out_dtype: int = abi.OUT_F16, fields: torch.Tensor | None = None,
                      out: torch.Tensor | None = None) -> torch.Tensor:
        n = params.shape[0]
        shape = (n, 3, self.cfg.out_h, self.cfg.out_w)
        if out is None:
            out = torch.empty(shape, dtype=_TORCH_OUT[out_dtype], device=self.device)
        else:
            assert out.shape == shape and out.dtype == _TORCH_OUT[out_dtype] and out.is_contiguous()
        rc = self.lib.mtgv_encoder_batch(self._h, _ptr(params), n, _ptr(out), out_dtype, _ptr(fields), self._stream())
        self._check(rc, "mtgv_encoder_batch")
        return out

    def encoder_targets(self, cards: torch.Tensor, out_dtype: int = abi.OUT_F16) -> torch.Tensor:
        cards = cards.to(self.device, dtype=torch.int32).contiguous()
        n = cards.numel()
        out = torch.empty((n, 3, self.cfg.y_h, self.cfg.y_w), dtype=_TORCH_OUT[out_dtype], device=self.device)
        rc = self.lib.mtgv_encoder_targets(self._h, _ptr(cards), n, _ptr(out), out_dtype, self._stream())
        self._check(rc, "mtgv_encoder_targets")
        return out

    # ------------------------------------------------------------------ detection path
    def set_det_config(self, *, bg_size_hw=640, num_cards_min=1, num_cards_max=10, card_min_visible_ratio=0.5,
                       card_min_visible_ratio_edges=1.0, card_jitter_ratio=0.3, card_min_area_ratio=0.02,
                       card_max_area_ratio=0.9, card_no_contains=True, card_max_place_attempts=10, ratio_bg=None,
                       kind="obb", photometrics=True, card_size_sample_mode="log_uniform", n_bgs_first=0, bg_first_prob=1.0):
        hw = (bg_size_hw, bg_size_hw) if isinstance(bg_size_hw, int) else tuple(bg_size_hw)
        cfg = abi.DetConfig(int(hw[0]), int(hw[1]), int(num_cards_min), int(num_cards_max), float(card_min_visible_ratio),
                            -1.0 if card_min_visible_ratio_edges is None else float(card_min_visible_ratio_edges),
                            float(card_jitter_ratio), float(card_min_area_ratio), float(card_max_area_ratio),
                            float(ratio_bg or 0.0), int(card_no_contains), int(card_max_place_attempts),
                            {"obb": 0, "seg": 1}[kind], int(photometrics),
                            {"log_uniform": 0, "uniform": 1}[card_size_sample_mode], int(n_bgs_first), float(bg_first_prob))
        self._check(self.lib.mtgv_set_det_config(self._h, C.byref(cfg)), "mtgv_set_det_config")
        self.det_cfg = cfg

    def sample_det_tape(self, seed: int, first_index: int, n: int) -> torch.Tensor:
        tape = torch.zeros((n, abi.DET_TAPE_DTYPE.itemsize), dtype=torch.uint8, device=self.device)
        rc = self.lib.mtgv_sample_det_tape(self._h, C.c_uint64(seed & (2**64 - 1)), C.c_int64(first_index), n, _ptr(tape),
                                           self._stream())
        self._check(rc, "mtgv_sample_det_tape")
        return tape

    def upload_det_tape(self, tape_np: np.ndarray) -> torch.Tensor:
        assert tape_np.dtype == abi.DET_TAPE_DTYPE
        return torch.from_numpy(tape_np.view(np.uint8).reshape(len(tape_np), -1).copy()).to(self.device)

    def det_place(self, tape: torch.Tensor):
        """-> (params, accepted [n,32], keypoints [n,96,8,2] f64, labels [n,96], counts [n])"""
        n = tape.shape[0]
        nk = abi.DET_MAX_CARDS * abi.DET_MAX_KPOLY
        params = torch.empty((n, self.lib.mtgv_det_params_size()), dtype=torch.uint8, device=self.device)
        accepted = torch.empty((n, abi.DET_MAX_CARDS), dtype=torch.int32, device=self.device)
        keypoints = torch.zeros((n, nk, abi.DET_MAX_KP, 2), dtype=torch.float64, device=self.device)
        labels = torch.empty((n, nk), dtype=torch.int32, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        rc = self.lib.mtgv_det_place(self._h, _ptr(tape), n, _ptr(params), _ptr(accepted), _ptr(keypoints), _ptr(labels),
                                     _ptr(counts), self._stream())
        self._check(rc, "mtgv_det_place")
        return params, accepted, keypoints, labels, counts

    def det_batch(self, params: torch.Tensor, out_dtype: int = abi.OUT_U8, fields: torch.Tensor | None = None,
                  out: torch.Tensor | None = None) -> torch.Tensor:
        n = params.shape[0]
        shape = (n, 3, self.det_cfg.size_h, self.det_cfg.size_w)
        if out is None:
            out = torch.empty(shape, dtype=_TORCH_OUT[out_dtype], device=self.device)
        rc = self.lib.mtgv_det_batch(self._h, _ptr(params), n, _ptr(out), out_dtype, _ptr(fields), self._stream())
        self._check(rc, "mtgv_det_batch")
        return out

    # ------------------------------------------------------------------ serving-side dewarp
    def extract_dewarped(self, frame: torch.Tensor, quads: torch.Tensor, out_size_hw=(192, 128), expand_ratio: float = 0.05) -> torch.Tensor:
        """InstanceSeg.extract_dewarped (mtgvision/od_export.py:95-111) for every detected card of a frame.
        frame (H,W,C) uint8; quads (n,4,2) corner points ordered like `xyxyxyxy`; returns (n,h,w,C) uint8 on the device."""
        frame = frame.to(self.device, dtype=torch.uint8).contiguous()
        assert frame.ndim == 3 and 1 <= frame.shape[2] <= 4
        quads = torch.as_tensor(quads).to(torch.float32).reshape(-1, 8).to(self.device).contiguous()  # .astype(np.float32) (:106)
        h, w = int(out_size_hw[0]), int(out_size_hw[1])
        # dst_pts = (1 + r) * [[0,0],[w,0],[w,h],[0,h]] - 0.5 * r * [w,h] in float64, then float32 (:101-107)
        dst = (1.0 + expand_ratio) * np.asarray([[0, 0], [w, 0], [w, h], [0, h]]) - (0.5 * expand_ratio) * np.asarray([w, h])
        dst_t = torch.from_numpy(dst.astype(np.float32).reshape(8)).to(self.device)
        n = quads.shape[0]
        out = torch.empty((n, h, w, frame.shape[2]), dtype=torch.uint8, device=self.device)
        rc = self.lib.mtgv_extract_dewarped(self._h, _ptr(frame), frame.shape[0], frame.shape[1], frame.shape[2], _ptr(quads), n,
                                            _ptr(dst_t), _ptr(out), h, w, self._stream())
        self._check(rc, "mtgv_extract_dewarped")
        return out

    # ------------------------------------------------------------------ image decode into the pools
    MAX_BG_AREA_SCALE = 24.0  # kBgMaxAreaScale (mtgv_geom.cuh): largest INTER_AREA reduction of the background chain

    def oversized_backgrounds(self, x_size_hw) -> np.ndarray:
        """Pool indices of backgrounds so large that some rotation needs an INTER_AREA reduction beyond the kernels' limit
        (factor 24): `rotate_bounded` grows the canvas up to the image diagonal (util/image.py:380-398) and `crop_to_size` then
        shrinks it by min(canvas_h / out_h, canvas_w / out_w) (:349-377).  At 192x128 that is a diagonal above ~4600 px.
        The dataset classes REFUSE such pools when they are built, so no sample can fail for its size at run time."""
        hw = np.asarray(getattr(self, "bg_hw", np.zeros((0, 2))), dtype=np.float64).reshape(-1, 2)
        diag = np.hypot(hw[:, 0], hw[:, 1])
        return np.nonzero(np.minimum(diag / x_size_hw[0], diag / x_size_hw[1]) > self.MAX_BG_AREA_SCALE)[0]

    def jpeg_info(self, data: bytes) -> tuple[int, int]:
        """(h, w) of a baseline JPEG file; raises MtgvError for files the device decoder does not support."""
        buf = np.frombuffer(data, dtype=np.uint8)
        hw = np.zeros(2, dtype=np.int32)
        rc = self.lib.mtgv_jpeg_info(self._h, buf.ctypes.data_as(C.c_void_p), len(buf), hw.ctypes.data_as(C.c_void_p))
        self._check(rc, "mtgv_jpeg_info")
        return int(hw[0]), int(hw[1])

    def jpeg_info_batch(self, blob, file_off: np.ndarray) -> np.ndarray:
        """(h, w) of every file of a concatenated host buffer in one C call (threaded marker walks): int32 [n, 2]."""
        file_off = np.ascontiguousarray(file_off, dtype=np.int64)
        n = len(file_off) - 1
        hw = np.zeros((max(n, 0), 2), dtype=np.int32)
        if n > 0:
            base = blob.data_ptr() if isinstance(blob, torch.Tensor) else np.frombuffer(blob, dtype=np.uint8).ctypes.data
            rc = self.lib.mtgv_jpeg_info_batch(self._h, C.c_void_p(base), file_off.ctypes.data_as(C.c_void_p), n,
                                               hw.ctypes.data_as(C.c_void_p))
            self._check(rc, "mtgv_jpeg_info_batch")
        return hw

    def prepare_jpeg_blob(self, blob: torch.Tensor, file_off, hw: np.ndarray | None = None) -> dict:
        """Zero-copy batch for `decode_prepared`: `blob` is a uint8 CPU tensor (pinned for an asynchronous upload - e.g.
        the arena a loader `readinto`s its files) that already holds the files back to back, file i =
        blob[file_off[i]:file_off[i+1]].  `hw` (int32 [n,2]) skips the header pass when the caller knows the frame sizes;
        mtgv_decode_jpeg_batch checks them against every file either way."""
        assert blob.dtype == torch.uint8 and not blob.is_cuda and blob.is_contiguous()
        file_off = np.ascontiguousarray(file_off, dtype=np.int64)
        n = len(file_off) - 1
        if hw is None:
            hw = self.jpeg_info_batch(blob, file_off)
        hw = np.ascontiguousarray(hw, dtype=np.int32).reshape(n, 2)
        out_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(hw[:, 0].astype(np.int64) * hw[:, 1] * 3, out=out_off[1:])
        return {"n": n, "blob": blob, "file_off": file_off, "out_off": out_off, "hw": hw}

    def prepare_jpegs(self, files: list[bytes], hw: np.ndarray | None = None) -> dict:
        """Host-side batch of JPEG files for `decode_prepared`: the files are copied back to back into one of three rotating
        pinned staging buffers the context keeps (grown on demand; three, so that a buffer is not rewritten while the two
        batches a streaming pipeline has in flight still upload from theirs) and their headers parsed in one C call
        (skipped when the caller passes the frame sizes `hw`: the decoder checks every file against them anyway)."""
        n = len(files)
        files = [f if isinstance(f, bytes) else bytes(f) for f in files]
        lens = np.fromiter((len(f) for f in files), dtype=np.int64, count=n)
        file_off = np.zeros(n + 1, dtype=np.int64)
        total = int(lens.sum())
        ring = self.__dict__.setdefault("_jpeg_stage", [None, None, None])
        k = self.__dict__["_jpeg_stage_next"] = (self.__dict__.get("_jpeg_stage_next", -1) + 1) % 3
        if ring[k] is None or ring[k].numel() < max(total, 1):
            ring[k] = torch.empty(max(total + total // 8, 1 << 16), dtype=torch.uint8).pin_memory()
        blob = ring[k][: max(total, 1)]
        # mtgv_gather_files: the n buffers copied back to back by a few host threads (the GIL is released for the call)
        srcs = (C.c_char_p * max(n, 1))(*files)
        rc = self.lib.mtgv_gather_files(self._h, C.cast(srcs, C.c_void_p), lens.ctypes.data_as(C.c_void_p), n, C.c_void_p(blob.data_ptr()),
                                        int(blob.numel()), file_off.ctypes.data_as(C.c_void_p))
        self._check(rc, "mtgv_gather_files")
        return self.prepare_jpeg_blob(blob, file_off, hw=hw)

    def decode_prepared(self, batch: dict, out: torch.Tensor | None = None) -> torch.Tensor:
        """mtgv_decode_jpeg_batch on a prepared batch; returns the flat uint8 device tensor of all images."""
        total = int(batch["out_off"][-1])
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.numel() >= total
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.mtgv_decode_jpeg_batch(self._h, C.c_void_p(batch["blob"].data_ptr()), vp(batch["file_off"]), batch["n"], _ptr(out),
                                             vp(batch["out_off"]), vp(batch["hw"]), self._stream())
        self._check(rc, "mtgv_decode_jpeg_batch")
        # the pinned file bytes are read by an asynchronous copy: keep them alive until the next call (which starts by
        # waiting for the stream) instead of letting the caller's dict be the only reference
        self._jpeg_inflight = batch
        return out

    def decode_into_pools(self, batch: dict, n_cards: int, first_card: int, n_bgs: int, first_bg: int) -> None:
        """mtgv_decode_jpeg_to_pools: files [0, n_cards) of a prepared batch into card slots [first_card, ...), the next n_bgs
        files into background slots [first_bg, ...), decoded straight into the pools' own layouts (no HWC intermediate)."""
        assert n_cards + n_bgs == batch["n"]
        rc = self.lib.mtgv_decode_jpeg_to_pools(self._h, C.c_void_p(batch["blob"].data_ptr()), batch["file_off"].ctypes.data_as(C.c_void_p),
                                                int(n_cards), int(first_card), int(n_bgs), int(first_bg), self._stream())
        self._check(rc, "mtgv_decode_jpeg_to_pools")
        self._jpeg_inflight = batch  # the pinned bytes are read asynchronously: keep them alive until the next call

    def jpeg_last_kernel_ms(self) -> tuple[float, float, float]:
        ms = (C.c_float * 3)()
        self._check(self.lib.mtgv_jpeg_last_kernel_ms(self._h, ms), "mtgv_jpeg_last_kernel_ms")
        return ms[0], ms[1], ms[2]

    def decode_jpegs(self, files: list[bytes]):
        """cv2.imread(path, IMREAD_COLOR_RGB) (mtgvision/util/image.py:107-114) for a list of JPEG files, on the device.
        Returns (flat uint8 device tensor, byte offsets int64 [n], hw int32 [n,2]); image i is
        flat[off[i] : off[i] + 3*h*w].view(h, w, 3) - the arguments of mtgv_set_bg_pool."""
        batch = self.prepare_jpegs(files)
        return self.decode_prepared(batch), batch["out_off"][:-1].copy(), batch["hw"]

    def set_card_pool_from_jpegs(self, files: list[bytes], labels3, grp_off, grp_mem):
        """Card pool straight from JPEG file bytes (all of one size): decoded on the device, never materialised on the host."""
        flat, _, hw = self.decode_jpegs(files)
        if len(files) == 0 or not (hw == hw[0]).all():
            raise ValueError("card files must all have the same size")
        torch.cuda.current_stream(self.device).synchronize()
        self.set_card_pool(flat.view(len(files), int(hw[0, 0]), int(hw[0, 1]), 3), labels3, grp_off, grp_mem)

    def set_bg_pool_from_jpegs(self, files: list[bytes]):
        """Background pool straight from JPEG file bytes: decoded on the device, never materialised on the host."""
        flat, off, hw = self.decode_jpegs(files)
        torch.cuda.current_stream(self.device).synchronize()  # mtgv_set_bg_pool ingests on its own stream
        rc = self.lib.mtgv_set_bg_pool(self._h, _ptr(flat), off.ctypes.data_as(C.c_void_p), hw.ctypes.data_as(C.c_void_p), len(files))
        self._check(rc, "mtgv_set_bg_pool")
        self.n_bgs = len(files)
        self.bg_hw = hw

    # ------------------------------------------------------------------ image encode for the dataset writer
    def encode_jpegs(self, images: torch.Tensor, quality: int = 95, layout: str | None = None, cap: int | None = None) -> list[bytes]:
        """cv2.imwrite's JPEG bytes (save_sample -> imwrite, od_datasets.py:829-831) for a batch of uint8 RGB images on the
        device: (n,3,H,W) or (n,H,W,3), any size.  Returns one `bytes` per image."""
        return [v.tobytes() for v in self.encode_jpegs_host(images, quality, layout, cap)]

    def encode_jpegs_host(self, images: torch.Tensor, quality: int = 95, layout: str | None = None, cap: int | None = None) -> list[np.ndarray]:
        """Like `encode_jpegs`, without the per-file copy: the files are compacted on the device, brought over in ONE
        transfer into a pinned buffer the context keeps, and returned as uint8 views of it (valid until the next call)."""
        out, lens = self.encode_jpegs_device(images, quality, layout, cap)
        n, capb = out.shape
        if n == 0:
            return []
        if getattr(self, "_jpeg_compact", None) is None or self._jpeg_compact.numel() < n * capb:
            self._jpeg_compact = torch.empty(n * capb, dtype=torch.uint8, device=self.device)
        offsets = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        rc = self.lib.mtgv_compact_jpeg_files(self._h, _ptr(out), C.c_int64(capb), _ptr(lens), n, _ptr(self._jpeg_compact), _ptr(offsets),
                                              self._stream())
        self._check(rc, "mtgv_compact_jpeg_files")
        meta = torch.cat([offsets, lens.to(torch.int64)]).cpu().numpy()  # one small transfer, synchronises
        off_h, lens_h = meta[: n + 1], meta[n + 1:]
        if (lens_h < 0).any():
            raise abi.MtgvError(f"mtgv_encode_jpeg_batch: {int((lens_h < 0).sum())} image(s) do not fit in cap={capb} bytes")
        total = int(off_h[n])
        if getattr(self, "_jpeg_pinned", None) is None or self._jpeg_pinned.numel() < total:
            self._jpeg_pinned = torch.empty(max(total + total // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        self._jpeg_pinned[:total].copy_(self._jpeg_compact[:total])
        host = self._jpeg_pinned.numpy()
        return [host[off_h[i]: off_h[i] + lens_h[i]] for i in range(n)]

    def encode_jpegs_device(self, images: torch.Tensor, quality: int = 95, layout: str | None = None, cap: int | None = None):
        assert images.is_cuda and images.dtype == torch.uint8 and images.ndim == 4 and images.is_contiguous()
        if layout is None:
            layout = "nchw" if images.shape[1] == 3 and images.shape[3] != 3 else "nhwc"
        n = images.shape[0]
        h, w = (images.shape[2], images.shape[3]) if layout == "nchw" else (images.shape[1], images.shape[2])
        cap = int(cap) if cap is not None else (((h + 15) // 16) * ((w + 15) // 16) * 384 + 4096) // 4 * 4
        out = torch.empty((n, cap), dtype=torch.uint8, device=self.device)
        lens = torch.empty(n, dtype=torch.int32, device=self.device)
        rc = self.lib.mtgv_encode_jpeg_batch(self._h, _ptr(images), n, h, w, abi.LAYOUT_NCHW if layout == "nchw" else abi.LAYOUT_NHWC,
                                             int(quality), _ptr(out), C.c_int64(cap), _ptr(lens), self._stream())
        self._check(rc, "mtgv_encode_jpeg_batch")
        return out, lens

    def jpeg_encode_last_kernel_ms(self) -> tuple[float, float]:
        ms = (C.c_float * 2)()
        self._check(self.lib.mtgv_jpeg_encode_last_kernel_ms(self._h, ms), "mtgv_jpeg_encode_last_kernel_ms")
        return ms[0], ms[1]

    # ------------------------------------------------------------------ parity / debug entries
    def warp_perspective(self, src: torch.Tensor, M: torch.Tensor, dsize_hw) -> torch.Tensor:
        """src (n,h,w,c) float32, M (n,3,3) float64 -> (n,dh,dw,c) float32, cv2.warpPerspective semantics."""
        src = src.to(self.device, dtype=torch.float32).contiguous()
        M = M.to(self.device, dtype=torch.float64).contiguous()
        n, sh, sw, c = src.shape
        dh, dw = dsize_hw
        dst = torch.empty((n, dh, dw, c), dtype=torch.float32, device=self.device)
        rc = self.lib.mtgv_warp_perspective(self._h, _ptr(src), n, sh, sw, c, _ptr(M), _ptr(dst), dh, dw, self._stream())
        self._check(rc, "mtgv_warp_perspective")
        return dst

    def run_plane_ops(self, img: torch.Tensor, ops_np: np.ndarray, fields: torch.Tensor | None = None, seed: int = 0):
        """img (n,h,w,c) float32 (modified in place and returned); ops_np: array of abi.XOP_DTYPE."""
        assert ops_np.dtype == abi.XOP_DTYPE
        img = img.to(self.device, dtype=torch.float32).contiguous()
        n, h, w, c = img.shape
        ops = torch.from_numpy(ops_np.view(np.uint8).reshape(len(ops_np), -1).copy()).to(self.device)
        rc = self.lib.mtgv_run_plane_ops(self._h, _ptr(img), n, h, w, c, _ptr(ops), len(ops_np), _ptr(fields),
                                         C.c_uint64(seed), self._stream())
        self._check(rc, "mtgv_run_plane_ops")
        return img
