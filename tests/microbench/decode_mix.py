"""Decode timing of the e2e batch mix (512 synthetic cards 680x488 + 512 backgrounds 375x500, quality 90): python tests/microbench/decode_mix.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cv2, numpy as np, torch
from mtgvision_b200 import synth
from mtgvision_b200.context import Context

q = [cv2.IMWRITE_JPEG_QUALITY, 90]
cards = [cv2.imencode(".jpg", synth.synth_card(k)[:, :, ::-1], q)[1].tobytes() for k in range(128)]
bgs = [cv2.imencode(".jpg", synth.synth_bg(k)[:, :, ::-1], q)[1].tobytes() for k in range(128)]
files = [cards[k % 128] for k in range(512)] + [bgs[k % 128] for k in range(512)]
ctx = Context(0)
batch = ctx.prepare_jpegs(files)
out = torch.empty(int(batch["out_off"][-1]), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.decode_prepared(batch, out)
torch.cuda.synchronize()
kms = np.zeros(3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for _ in range(n):
    ctx.decode_prepared(batch, out)
    kms += ctx.jpeg_last_kernel_ms()
print({"files": len(files), "file_mbytes": sum(len(f) for f in files) / 1e6, "entropy_ms": kms[0] / n, "idct_ms": kms[1] / n, "color_ms": kms[2] / n,
       "mean_card_kb": np.mean([len(f) for f in cards]) / 1e3, "mean_bg_kb": np.mean([len(f) for f in bgs]) / 1e3})
