"""First-contact GPU diagnostic: per-op and per-sample parity numbers with details, written
to gpurun_out/diag.txt.  Not a pytest file; run as `python tests/gpu_diag.py`."""
import os, sys, time, traceback
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2, torch
from mtgvision_b200 import abi
from tests import parity_util as PU
from oracle import encoder_oracle as EO, cv2_restate as R

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n"); LOG.flush()


def xops(n):
    a = np.zeros(n, dtype=abi.XOP_DTYPE)
    a["field"] = -1; a["field2"] = -1
    return a


def main():
    log(torch.cuda.get_device_name(0), cv2.__version__)
    pool, bgs = PU.small_pools(8, 8)
    t0 = time.time()
    ctx = PU.make_context(pool, bgs, half_upsidedown=True)
    log("context + pools", round(time.time() - t0, 2), "s")
    rng = np.random.default_rng(0)
    # 1. warp_perspective parity entry
    for (sh, sw, c, dh, dw) in [(192, 128, 4, 192, 128), (680, 488, 3, 640, 640), (375, 500, 1, 200, 300)]:
        src = rng.random((2, sh, sw, c), dtype=np.float32)
        Ms = []
        for k in range(2):
            s = np.float32([[0, 0], [sw, 0], [0, sh], [sw, sh]])
            d = (s * (dw / sw, dh / sh) + rng.uniform(-0.15, 0.15, (4, 2)) * [dw, dh]).astype(np.float32)
            Ms.append(cv2.getPerspectiveTransform(s, d))
        out = ctx.warp_perspective(torch.from_numpy(src), torch.from_numpy(np.stack(Ms)), (dh, dw)).cpu().numpy()
        for k in range(2):
            ref = cv2.warpPerspective(src[k], Ms[k], (dw, dh)).reshape(dh, dw, c)
            log("warp_perspective", (sh, sw, c, dh, dw), "max abs", float(np.abs(out[k] - ref).max()), "exact", np.array_equal(out[k], ref))
    # 2. plane ops one by one
    img = rng.random((2, 192, 128, 4), dtype=np.float32)
    def run(ops, im=img, fields=None):
        return ctx.run_plane_ops(torch.from_numpy(im.copy()), ops, fields).cpu().numpy()
    o = xops(1); o[0]["code"] = abi.X_WARP_PERSP
    s = np.float32([[0, 0], [128, 0], [0, 192], [128, 192]]); d = (s + rng.uniform(-12, 12, (4, 2))).astype(np.float32)
    M = cv2.getPerspectiveTransform(s, d); o[0]["d"][:] = cv2.invert(M)[1].reshape(-1)
    got = run(o)
    for k in range(2):
        ref = cv2.warpPerspective(img[k], M, (128, 192)); log("op persp exact", np.array_equal(got[k], ref), float(np.abs(got[k] - ref).max()))
    o = xops(1); o[0]["code"] = abi.X_WARP_AFFINE
    A = cv2.getRotationMatrix2D((64, 96), 4.0, 1.05); A[:, 2] += (3.3, -7.1)
    o[0]["d"][:6] = R.invert_affine(A).reshape(-1)
    got = run(o); ref = cv2.warpAffine(img[0], A, (128, 192)); log("op affine exact", np.array_equal(got[0], ref), float(np.abs(got[0] - ref).max()))
    for n in (1, 2):
        for dn in (0, 1, 2):
            for up in (0, 1, 2):
                o = xops(1); o[0]["code"] = abi.X_DOWNUP; o[0]["i"][:3] = (n, dn, up)
                got = run(o)
                ref = cv2.resize(cv2.resize(img[0], (128 >> n, 192 >> n), interpolation=dn), (128, 192), interpolation=up)
                log("op downup", n, dn, up, "max abs", float(np.abs(got[0] - ref).max()))
    o = xops(1); o[0]["code"] = abi.X_BLUR3
    got = run(o); log("op blur max abs", float(np.abs(got[0] - cv2.GaussianBlur(img[0], (3, 3), 0)).max()))
    o = xops(1); o[0]["code"] = abi.X_SHARPEN
    got = run(o); k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]]); log("op sharpen max abs", float(np.abs(got[0] - np.clip(cv2.filter2D(img[0], -1, k), 0, 1)).max()))
    # 3. expansion: device == host harness
    import ctypes as C
    refs, tapes = [], []
    for seed in range(64):
        im, t = PU.oracle_virtual(pool, bgs, seed, seed % 8, (seed // 8) % 8)
        refs.append(im); tapes.append(t)
    out, labels, params = PU.gpu_run_tapes(ctx, tapes, abi.OUT_F32)
    hh = C.CDLL(os.path.join(ROOT, "tests", "host_harness", "libmtgv_hostharness.so"))
    from oracle import tape_pack
    arr, _ = tape_pack.pack_tapes(tapes)
    hp = np.zeros(len(tapes), dtype=abi.PARAMS_DTYPE)
    bg_hw = np.asarray([b.shape[:2] for b in bgs], dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    cfg = ctx.cfg
    hh.hh_expand_encoder(vp(arr), len(tapes), C.byref(cfg), 680, 488, len(pool), vp(pool.labels3), vp(pool.grp_off), vp(pool.grp_mem), len(bgs), vp(bg_hw), vp(hp))
    same = params.tobytes() == hp.tobytes()
    log("expand device==host bitwise", same)
    if not same:
        for name in abi.PARAMS_DTYPE.names:
            if name == "ops": continue
            if not np.array_equal(params[name], hp[name]): log("  differs:", name)
        for name in abi.XOP_DTYPE.names:
            if not np.array_equal(params["ops"][name], hp["ops"][name]): log("  ops differ:", name)
    log("status", np.unique(params["status"]))
    # 4. full samples
    worst = []
    for k, (o_, r_, t) in enumerate(zip(out, refs, tapes)):
        d8, df = PU.lsb_diff(o_, r_)
        worst.append(d8)
        if d8 > 1 or k < 8:
            frac = float((np.abs(o_ - r_) * 255 > 1).mean())
            log(f"sample {k} lsb {d8} maxabs*255 {df:.3f} frac>1 {frac:.5f} fg {[x['op'] for x in t['fg_ops']]} bg {[x['op'] for x in t['bg_ops']]} v {[(x['op'], x.get('kind', x.get('n',''))) for x in t['vrtl_ops']]}")
    log("virtual parity: worst LSB", max(worst), "hist", np.bincount(np.asarray(worst)).tolist())
    want = np.asarray([pool.labels3[t["card"]] for t in tapes], dtype=np.int64)
    log("labels exact", np.array_equal(labels, want))
    # 5. production sampler: run a batch
    torch.cuda.synchronize(); t0 = time.time()
    tape = ctx.sample_encoder_tape(1234, 0, 256)
    params, labels = ctx.expand_params(tape)
    x = ctx.encoder_batch(params, abi.OUT_F16)
    torch.cuda.synchronize(); t1 = time.time()
    p = params.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)
    log("sampled batch 512 x-samples", round(t1 - t0, 4), "s; status", np.unique(p["status"], return_counts=True), "kinds", np.bincount(p["kind"]))
    xf = x.float()
    log("x finite", bool(torch.isfinite(xf).all()), "min", float(xf.min()), "max", float(xf.max()), "mean", float(xf.mean()))
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        tape = ctx.sample_encoder_tape(99 + rep, 0, 512)
        params, labels = ctx.expand_params(tape)
        x = ctx.encoder_batch(params, abi.OUT_F16)
        torch.cuda.synchronize(); t1 = time.time()
        log("batch 1024 x-samples", round((t1 - t0) * 1e3, 2), "ms ->", round(1024 / (t1 - t0)), "x-samples/s")
    ctx.close()


try:
    main()
except Exception:
    log(traceback.format_exc())
    sys.exit(1)
