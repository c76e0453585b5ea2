"""Shared helpers for the parity tests: run the oracle with seeds, feed its recorded
tape to the CUDA path through the C ABI, compare."""

from __future__ import annotations

import random

import numpy as np

from mtgvision_b200 import abi, synth
from oracle import encoder_oracle as EO
from oracle import tape_pack

_POOLS = {}


def small_pools(n_cards=8, n_bgs=8, card_hw=synth.CARD_HW, bg_hw=synth.BG_HW):
    key = (n_cards, n_bgs, card_hw, bg_hw)
    if key not in _POOLS:
        _POOLS[key] = (synth.make_card_pool(n_cards, card_hw), synth.make_bg_pool(n_bgs, bg_hw))
    return _POOLS[key]


def make_context(pool, bgs, **cfg):
    from mtgvision_b200.context import Context

    ctx = Context(0)
    ctx.set_encoder_config(**cfg)
    ctx.set_card_pool(pool.images, pool.labels3, pool.grp_off, pool.grp_mem)
    ctx.set_bg_pool(bgs)
    return ctx


def oracle_virtual(pool, bgs, seed, card, bg, size_hw=(192, 128), half_upsidedown=True, fresh_shuffle=True):
    """One oracle make_virtual run with explicit seeds; returns (image f32 HWC, tape)."""
    random.seed(seed)
    np.random.seed(seed)
    if fresh_shuffle:
        EO.reset_shuffle_state()
    t = {"card": int(card), "bg": int(bg)}
    img = EO.make_virtual(EO.u8_to_f32(pool.images[card]), EO.u8_to_f32(bgs[bg]), size_hw, half_upsidedown, tape=t)
    return img, t


def gpu_run_tapes(ctx, tapes, out_dtype=abi.OUT_F32, host_transcendentals=True):
    """tapes (oracle dicts) -> (images [n,H,W,3] numpy in the out dtype, labels [n,3], params ndarray)."""
    import torch

    arr, fields = tape_pack.pack_tapes(tapes, host_transcendentals=host_transcendentals)
    tape_dev = ctx.upload_tape(arr)
    f = torch.from_numpy(fields.to_array().view(np.int32)).to(ctx.device)
    params, labels = ctx.expand_params(tape_dev)
    out = ctx.encoder_batch(params, out_dtype, fields=f)
    torch.cuda.synchronize()
    p = params.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)
    return out.permute(0, 2, 3, 1).contiguous().cpu().numpy(), labels.cpu().numpy(), p


def lsb_diff(gpu, ref_f32):
    """Per-pixel difference in uint8 LSB between the CUDA output and the float32 reference.
    Returns (max |round(255 g) - round(255 r)|, max |g - r| * 255)."""
    if gpu.dtype == np.uint8:
        g8 = gpu.astype(np.int32)
        gf = gpu.astype(np.float64) / 255.0
    else:
        gf = gpu.astype(np.float64)
        g8 = np.rint(np.clip(gf, 0, 1) * 255.0).astype(np.int32)
    r8 = np.rint(np.clip(ref_f32.astype(np.float64), 0, 1) * 255.0).astype(np.int32)
    return int(np.abs(g8 - r8).max()), float(np.abs(gf - ref_f32).max() * 255.0)


def smoke_check():
    """__graft_entry__.smoke(): 6 oracle samples (+ one cropped) vs the CUDA path on cuda:0."""
    pool, bgs = small_pools(4, 4)
    ctx = make_context(pool, bgs, half_upsidedown=True)
    refs, tapes = [], []
    for seed in range(6):
        img, t = oracle_virtual(pool, bgs, seed, seed % 4, (seed // 2) % 4)
        refs.append(img)
        tapes.append(t)
    t = {"card": 1, "bg": 0}
    refs.append(EO.make_cropped(EO.u8_to_f32(pool.images[1]), (192, 128), tape=t))
    tapes.append(t)
    out, labels, _ = gpu_run_tapes(ctx, tapes, abi.OUT_F16)
    worst = max(lsb_diff(o, r)[0] for o, r in zip(out, refs))
    assert worst <= 1, f"smoke: max uint8 LSB error {worst} > 1"
    want = np.asarray([pool.labels3[t["card"]] for t in tapes], dtype=np.int64)
    assert np.array_equal(labels, want), "smoke: labels differ"
    ctx.close()
