"""Serving-side dewarp of detected cards - the one piece of `mtgvision/od_export.py` that is image
arithmetic (`InstanceSeg.extract_dewarped`, reference :95-111); the CoreML / ultralytics / shapely parts
of that file are out of scope.

    dewarper = Dewarper(device=0)
    crops = dewarper.extract_dewarped(frame_u8_hwc, xyxyxyxy_points, out_size_hw=(192, 128))   # (n,192,128,3) uint8 cuda

Results equal `cv2.warpPerspective(frame, cv2.getPerspectiveTransform(pts.astype(float32), dst.astype(float32)),
(w, h))` bit for bit (tests/test_gpu_dewarp.py).
"""

from __future__ import annotations

import numpy as np
import torch

from .context import Context


class Dewarper:
    def __init__(self, device: int | None = None, ctx: Context | None = None):
        self.ctx = ctx if ctx is not None else Context(device)

    def extract_dewarped(self, frame, xyxyxyxy, out_size_hw=(192, 128), expand_ratio: float = 0.05) -> torch.Tensor:
        """frame: (H,W,C) uint8 numpy array or tensor; xyxyxyxy: (n,4,2) or (4,2) corner points in the
        reference's order (top-left first, od_export.py:88-92)."""
        if isinstance(frame, np.ndarray):
            frame = torch.from_numpy(np.ascontiguousarray(frame))
        q = torch.as_tensor(np.asarray(xyxyxyxy)) if not torch.is_tensor(xyxyxyxy) else xyxyxyxy
        single = q.ndim == 2
        out = self.ctx.extract_dewarped(frame, q.reshape(-1, 4, 2), out_size_hw, expand_ratio)
        return out[0] if single else out
