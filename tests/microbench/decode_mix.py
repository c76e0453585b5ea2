"""Decode timing of the e2e batch mix (512 synthetic cards 680x488 + 512 backgrounds 375x500, quality 90):
python tests/microbench/decode_mix.py [repeats] [pools]   - `pools` decodes into the pool layouts (the e2e path: planar
cards, RGBX backgrounds) instead of packed RGB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2, numpy as np, torch
from mtgvision_b200 import synth
from mtgvision_b200.context import Context

q = [cv2.IMWRITE_JPEG_QUALITY, 90]
cards = [cv2.imencode(".jpg", synth.synth_card(k)[:, :, ::-1], q)[1].tobytes() for k in range(128)]
bgs = [cv2.imencode(".jpg", synth.synth_bg(k)[:, :, ::-1], q)[1].tobytes() for k in range(128)]
files = [cards[k % 128] for k in range(512)] + [bgs[k % 128] for k in range(512)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
pools = len(sys.argv) > 2 and sys.argv[2] == "pools"
if pools:
    from tests import parity_util as PU
    pool = synth.CardPool(np.zeros((512, 680, 488, 3), np.uint8), synth.synth_faces(512))
    ctx = PU.make_context(pool, [np.zeros((375, 500, 3), np.uint8) for _ in range(512)])
    batch = ctx.prepare_jpegs(files)
    run = lambda: ctx.decode_into_pools(batch, 512, 0, 512, 0)  # noqa: E731
else:
    ctx = Context(0)
    batch = ctx.prepare_jpegs(files)
    out = torch.empty(int(batch["out_off"][-1]), dtype=torch.uint8, device="cuda")
    run = lambda: ctx.decode_prepared(batch, out)  # noqa: E731
for _ in range(3):
    run()
torch.cuda.synchronize()
kms = np.zeros(3)
for _ in range(n):
    run()
    kms += ctx.jpeg_last_kernel_ms()
print({"layout": "pools" if pools else "rgb", "files": len(files), "file_mbytes": round(sum(len(f) for f in files) / 1e6, 1), "entropy_ms": round(kms[0] / n, 4),
       "idct_ms": round(kms[1] / n, 4), "color_ms": round(kms[2] / n, 4)})
