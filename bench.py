#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 synthetic-sample generator.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU generator (oracle port) on host cores

Workload (BASELINE.json configs[1]): encoder training batches of 512 positive pairs with
hard-negative same-name swaps, synthetic 488x680 cards and 500x375 backgrounds,
x/x2 [512,3,192,128] fp16 NCHW + [512,3] int64 labels.  One step = one batch
(1024 augmented x-samples) per GPU; N GPUs run N independent shards (weak scaling,
no collective on the data path).  Metric = augmented x-samples/s, whole job.
One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "augmented samples/sec (encoder pairs)"
UNIT = "x-samples/s"
PAIRS = 512
X_HW = (192, 128)
CARD_HW = (680, 488)
BG_HW = (375, 500)
BYTES_CARD = CARD_HW[0] * CARD_HW[1] * 3
BYTES_BG = BG_HW[0] * BG_HW[1] * 3
BYTES_OUT_F16 = 3 * X_HW[0] * X_HW[1] * 2


def workload_config(args):
    """The `config` of the JSON line - identical for the CUDA arm and the reference arm (run-dependent counts go to `stats`)."""
    return {
        "workload": "encoder training batches: 512 positive pairs (x, x2) with hard-negative same-name swaps, "
                    "synthetic 680x488 cards + 375x500 backgrounds -> [512,3,192,128] fp16 NCHW + int64 labels",
        "pairs_per_step": PAIRS,
        "x_samples_per_step_per_gpu": 2 * PAIRS,
        "card_pool": args.pool_cards,
        "bg_pool": args.pool_bgs,
        "target_is_input_prob": 0.05,
        "similar_neg_prob": 0.2,
        "inputs_e2e": "JPEG files (quality 90) of the batch's cards and backgrounds in host memory, as the reference's loaders read them",
        "parallelism": f"{args.gpus} independent shard(s), one process per GPU, no collective",
        "l2": "inputs larger than L2: each step reads ~1000 distinct cards (~1 GB) of a multi-GB resident pool; no flush",
    }


# --------------------------------------------------------------------------------------- #
# reference arm / cpu_baseline: the oracle port on host cores                              #
# --------------------------------------------------------------------------------------- #

_CPU_STATE = {}
JPEG_Q = 90


def _encode_chunk(args):
    import cv2

    kind, a, b = args
    src = _CPU_STATE["cards"].images if kind == "c" else _CPU_STATE["bgs"]
    cv2.setNumThreads(1)
    return [cv2.imencode(".jpg", src[k][:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, JPEG_Q])[1].tobytes() for k in range(a, b)]


def encode_pools_as_jpeg(cards, bgs, workers):
    """The pools as the files a loader would find on disk (cv2.imwrite quality 90): (card_files, bg_files)."""
    import multiprocessing as mp

    _CPU_STATE["cards"], _CPU_STATE["bgs"] = cards, bgs
    nc, nb = len(cards.images), len(bgs)
    jobs = [("c", a, min(a + 32, nc)) for a in range(0, nc, 32)] + [("b", a, min(a + 32, nb)) for a in range(0, nb, 32)]
    with mp.get_context("fork").Pool(max(1, workers)) as pool:
        parts = pool.map(_encode_chunk, jobs)
    files = [f for part in parts for f in part]
    return files[:nc], files[nc:]


class _DecodingImages:
    """images[k] = cv2.imread of file k (imread_float, util/image.py:107-114: IMREAD_COLOR_RGB) - every access decodes,
    like `ran_card` -> dl_and_open_im_resized and IlsvrcImages.ran -> _load_image do per drawn sample."""

    def __init__(self, files):
        self.files = files

    def __len__(self):
        return len(self.files)

    def __getitem__(self, k):
        import cv2
        import numpy as np

        return cv2.imdecode(np.frombuffer(self.files[k], np.uint8), cv2.IMREAD_COLOR_RGB)


class _FileCards:
    def __init__(self, pool, files):
        self.images, self.labels3, self.group_of = _DecodingImages(files), pool.labels3, pool.group_of


def _cpu_worker(args):
    wid, step, n_pairs, from_files = args
    import random

    import cv2
    import numpy as np

    from oracle import encoder_oracle as EO

    cv2.setNumThreads(1)
    random.seed(1000 + 7919 * (step + 1) + wid)
    np.random.seed(1000 + 7919 * (step + 1) + wid)
    EO.reset_shuffle_state()
    cards, bgs = _CPU_STATE["cards"], _CPU_STATE["bgs"]
    if from_files:
        cards, bgs = _FileCards(cards, _CPU_STATE["card_files"]), _DecodingImages(_CPU_STATE["bg_files"])
    bo = EO.BatchOracle(cards, bgs, paired=True, targets=False, x_size_hw=X_HW, target_is_input_prob=0.05, similar_neg_prob=0.2)
    t0 = time.perf_counter()
    imgs, _, _ = bo.random_image_batch(n_pairs)  # one batch former call per worker, like a DataLoader worker
    assert imgs["x"].shape == (n_pairs, X_HW[0], X_HW[1], 3) and imgs["x2"].shape == imgs["x"].shape
    return 2 * n_pairs, time.perf_counter() - t0


class CpuGenerator:
    """The reference generator (oracle port: the same cv2/numpy calls in the same order as mtgvision/encoder_datasets.py +
    encoder_train.py:149-230) on all host cores: ONE persistent pool of worker processes, one per core, like the
    reference's DataLoader workers (encoder_train.py:517-523), over pools of the size the config states."""

    def __init__(self, pool_cards, pool_bgs, workers=None, cards=None, bgs=None, files=None):
        import multiprocessing as mp

        from mtgvision_b200 import synth

        self.workers = workers or os.cpu_count() or 1
        if cards is None:
            cards = synth.make_card_pool(pool_cards, workers=self.workers)
            bgs = synth.make_bg_pool(pool_bgs, workers=self.workers)
        _CPU_STATE["cards"], _CPU_STATE["bgs"] = cards, bgs
        if files is None:
            files = encode_pools_as_jpeg(cards, bgs, self.workers)
        _CPU_STATE["card_files"], _CPU_STATE["bg_files"] = files
        self.pool = mp.get_context("fork").Pool(self.workers)  # forked after the pools exist: workers share them copy-on-write
        self.pool.map(_cpu_worker, [(w, -1, 2, True) for w in range(self.workers)])  # warm-up: imports, page-in
        self._step = 0

    def step(self, pairs_per_worker, from_files=True):
        """-> (x_samples, wall seconds) of one map over all workers."""
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(w, self._step, pairs_per_worker, from_files) for w in range(self.workers)])
        self._step += 1
        return sum(r[0] for r in res), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = max(1, args.ref_pairs_per_worker)
    gen = CpuGenerator(args.pool_cards, args.pool_bgs)
    workers = gen.workers
    vals = []
    for step in range(args.warmup + args.steps):
        n, wall = gen.step(pairs, from_files=True)
        if step >= args.warmup:
            vals.append((n, wall))
    n_arr, w_arr = gen.step(pairs, from_files=False)  # context: the same step on decoded arrays (no cv2.imdecode)
    gen.close()
    n_sum = sum(n for n, _ in vals)
    w_sum = sum(w for _, w in vals)
    value = n_sum / w_sum
    sample = (f"each step: {workers} worker processes (one per host core, cv2 threads=1, one persistent pool) x one batch of {pairs} pairs = "
              f"{2 * pairs * workers} x-samples, drawn from pools of {args.pool_cards} cards / {args.pool_bgs} backgrounds held as quality-{JPEG_Q} JPEG "
              f"files in host memory; per drawn card/background cv2.imdecode (imread_float) + the generator; {n_sum} x-samples in {w_sum:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * w_sum / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "from_decoded_arrays": n_arr / w_arr},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU generator (oracle port of the reference's cv2/numpy code; the reference itself is Python and needs "
                "mtgdata/kornia/lightning, absent here) on all host cores, files in host memory -> batches in host memory",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- #
# clocks                                                                                   #
# --------------------------------------------------------------------------------------- #


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        # NVML in-process (a sample every 5 ms) when the bindings load; else nvidia-smi's own 20 ms loop
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                import torch

                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml, self.handle, self.stop_flag = pynvml, h, False
            self.samples = []
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.perf_counter(), sm, mx, r))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def _stop_nvml(self, t0, t1):
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        sm, mx, reasons = [], [], set()
        for t, c, m, r in self.samples:
            if t < t0 or t > t1:
                continue
            sm.append(float(c))
            mx.append(float(m))
            reasons.update(k for k, b in bits.items() if r & b)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}

    def stop(self, t0, t1):
        if getattr(self, "nvml", None) is not None:
            return self._stop_nvml(t0, t1)
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, row in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- #
# CUDA arm                                                                                 #
# --------------------------------------------------------------------------------------- #


def run_cuda(args):
    import numpy as np
    import torch

    from mtgvision_b200 import abi, synth
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.encoder_train import RanMtgEncDecDataset

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # ---- host-side setup first (untimed): everything that forks worker processes runs while this process is still
    # single-threaded and holds no CUDA context / NCCL communicator ----
    host_workers = max(1, (os.cpu_count() or 1) // max(1, world))
    t_setup = time.perf_counter()
    cards = synth.make_card_pool(args.pool_cards, workers=host_workers)
    bgs = synth.make_bg_pool(args.pool_bgs, workers=host_workers)
    card_files = bg_files = None
    if not args.no_e2e or not args.no_cpu_baseline:
        card_files, bg_files = encode_pools_as_jpeg(cards, bgs, host_workers)
    cpu_gen = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU baseline's worker processes are forked NOW and idle until the end of the run
        cpu_gen = CpuGenerator(args.pool_cards, args.pool_bgs, cards=cards, bgs=bgs, files=(card_files, bg_files))

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        # rank 0's stdout must hold ONE JSON line: NCCL prints its version banner to stdout when the first communicator comes up
        sys.stdout.flush()
        _stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    # ---- resident pools (setup, untimed) ----
    ds = RanMtgEncDecDataset(PAIRS, paired=True, targets=False, x_size_hw=X_HW, y_size_hw=X_HW, half_upsidedown=False,
                             target_is_input_prob=0.05, similar_neg_prob=0.2,
                             mtg=SyntheticBgFgMtgImages(pool=cards), ilsvrc=IlsvrcImages(images=bgs),
                             device=local_rank, out_dtype="float16", seed=20261018, rank=rank, world_size=world)
    ctx = ds.ctx
    t_setup = time.perf_counter() - t_setup
    n_x = 2 * PAIRS

    # preallocated per-step buffers (the public API allocates through torch's caching allocator;
    # here the same three C-ABI calls are issued with explicit events around the plane kernel)
    total_steps = args.warmup + args.steps
    tape = torch.empty((n_x, abi.TAPE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    n_keep = min(args.steps, 64)  # parameter blocks kept for the sample statistics; longer runs reuse them cyclically
    params = [torch.empty((n_x, abi.PARAMS_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(n_keep)]
    labels = torch.empty((n_x, 3), dtype=torch.int64, device=dev)
    out = torch.empty((n_x, 3, X_HW[0], X_HW[1]), dtype=torch.float16, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(i, timed_idx=None):
        first = (i * world + rank) * PAIRS
        ctx.sample_encoder_tape(ds.seed, first, PAIRS, out=tape)
        p = params[timed_idx % n_keep] if timed_idx is not None else params[0]
        ctx.expand_params(tape, params=p, labels=labels)
        if timed_idx is not None:
            ev[timed_idx][0].record()
        ctx.encoder_batch(p, abi.OUT_F16, out=out)
        if timed_idx is not None:
            ev[timed_idx][1].record()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # setup, untimed: bring a fresh box to steady state (clock ramp, lazy module load, allocator, L2/TLB) before the
    # W warm-up steps the contract asks for - measured 7 % slower steps in the first ~0.3 s of a cold process
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.6:
        for i in range(8):
            step(i)
        torch.cuda.synchronize()
    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    launches0 = ctx.launch_count()
    free0 = torch.cuda.mem_get_info(dev)[0]
    t0 = time.perf_counter()
    e_start.record()
    for k in range(args.steps):
        step(args.warmup + k, k)
    e_end.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    launches = ctx.launch_count() - launches0
    free1 = torch.cuda.mem_get_info(dev)[0]
    ms_total = e_start.elapsed_time(e_end)
    clk = clocks.stop(t0, t1)
    barrier()
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * n_x * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (k_encoder), live CUDA events on its stream ----
    k_ms = [a.elapsed_time(b) for a, b in ev]
    kinds = np.concatenate([p.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)["kind"] for p in params])
    status = np.concatenate([p.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)["status"] for p in params])
    n_virtual = int((kinds == abi.KIND_VIRTUAL).sum())
    n_cropped = int((kinds == abi.KIND_CROPPED).sum())
    alg_bytes = n_virtual * (BYTES_CARD + BYTES_BG + BYTES_OUT_F16) + n_cropped * (BYTES_CARD + BYTES_OUT_F16)
    alg_per_launch = alg_bytes / n_keep  # sample kinds of the last n_keep steps
    k_avg_ms = sum(k_ms) / len(k_ms)
    tenth = max(1, len(k_ms) // 10)
    drift = {"kernel_ms_first_tenth": sum(k_ms[:tenth]) / tenth, "kernel_ms_last_tenth": sum(k_ms[-tenth:]) / tenth,
             "kernel_ms_min": min(k_ms), "kernel_ms_median": statistics.median(k_ms), "kernel_ms_max": max(k_ms),
             "device_memory_growth_bytes": int(free0 - free1)}
    achieved = alg_per_launch / (k_avg_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp_["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["pixel_pipeline_dram_bytes_per_step"]
    except Exception:
        pass
    roofline = {"bound": "hbm",
                "kernel": "encoder pixel pipeline of one mtgv_encoder_batch call (k_background + k_foreground + k_encoder, "
                          "timed together; k_background dominates, see profiles/)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "kernel_ms": k_avg_ms,
                "kernel_share_of_step": k_avg_ms * args.steps / ms_total}

    # ---- e2e: host buffers in -> host buffers out through the public dataset API ----
    # Headline pair (like for like with the reference arm): the batch's 512 cards and 512 backgrounds arrive as the JPEG
    # FILES the reference's loaders read (encoder_train.py:149-156 -> dl_and_open_im_resized / imread_float,
    # util/image.py:107-114), in pinned host memory; every step parses / validates its batch (prepare_jpeg_batch_pinned),
    # uploads the compressed bytes, decodes on the device, ingests into the pools, generates, and downloads x, x2, labels.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, args.e2e_max_steps))  # pipeline fill and drain are inside the timed region
        n_rot = 4  # distinct pinned batches rotated through, so consecutive steps upload and decode different files
        rot = []
        for r_ in range(n_rot):
            sel_c = [card_files[(r_ * PAIRS + k) % len(card_files)] for k in range(PAIRS)]
            sel_b = [bg_files[(r_ * PAIRS + k) % len(bg_files)] for k in range(PAIRS)]
            lens = [len(f) for f in sel_c + sel_b]
            off = np.zeros(2 * PAIRS + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            blob = torch.empty(int(off[-1]), dtype=torch.uint8).pin_memory()
            blob.numpy()[:] = np.frombuffer(b"".join(sel_c + sel_b), dtype=np.uint8)
            rot.append((blob, off))
        h2d_files = sum(int(o[-1]) for _, o in rot) / n_rot

        def feed_jpeg(k):
            for i in range(k):
                blob, off = rot[i % n_rot]
                yield ds.prepare_jpeg_batch_pinned(blob, off, bg_hw=BG_HW)  # inside the timed loop

        def timed(feed_fn, k):
            barrier()
            es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_w0 = time.perf_counter()
            es.record()
            chk, res = 0.0, None
            for res in ds.host_tensor_batches(feed_fn(k)):
                chk += float(res["x_labels"][0, 0])  # the consumer touches every batch on the host
            ee.record()
            torch.cuda.synchronize()
            t_w1 = time.perf_counter()
            ms = torch.tensor([max(es.elapsed_time(ee), 1e3 * (t_w1 - t_w0))], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), res

        for _ in ds.host_tensor_batches(feed_jpeg(4)):  # warm-up: staging buffers, pinned outputs, decoder scratch
            pass
        j_ms, res = timed(feed_jpeg, e2e_steps)
        d2h = sum(v.numel() * v.element_size() for v in res.values())
        e2e = {"value": world * n_x * e2e_steps / (j_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_files),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": j_ms / e2e_steps,
               "api": "RanMtgEncDecDataset.host_tensor_batches fed with prepare_jpeg_batch_pinned items (called inside the timed loop): "
                      "the batch's 512 cards + 512 backgrounds as quality-90 JPEG files in pinned host memory -> device decode "
                      "(bit-exact with cv2.imread) -> pool ingest -> generation -> pinned fp16 x/x2 + int64 labels; upload+decode of "
                      "batch i+1 / kernels of batch i / download of batch i-1 overlap on four streams; "
                      "timed = max(CUDA events, host wall clock) over all steps including pipeline fill and drain"}

        # ---- the same call fed with decoded uint8 arrays (798 MB per batch over PCIe: the link is the bound) ----
        if not args.no_e2e_raw:
            hc = torch.from_numpy(cards.images[np.arange(PAIRS) % len(cards.images)]).pin_memory()
            hb = torch.from_numpy(np.stack([bgs[j % len(bgs)] for j in range(PAIRS)])).pin_memory()

            def feed(k):
                for _ in range(k):
                    yield hc, hb

            for _ in ds.host_tensor_batches(feed(3)):
                pass
            raw_steps = min(e2e_steps, 30)
            r_ms, _ = timed(feed, raw_steps)
            h2d = int(hc.numel() + hb.numel())
            e2e["from_raw_arrays"] = {
                "value": world * n_x * raw_steps / (r_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                "steps": raw_steps, "ms_per_step": r_ms / raw_steps, "h2d_gbs_needed_at_this_rate": h2d * raw_steps / (r_ms * 1e-3) / 1e9,
                "api": "the same call fed with pinned uint8 card/background ARRAYS (decoded images): PCIe-bound, 798 MB per batch"}
            del hc, hb

        # ---- the same call fed with lists of `bytes` objects (prepare_jpeg_batch: one extra host copy into pinned staging) ----
        def feed_bytes(k):
            def item(i):
                return ds.prepare_jpeg_batch([card_files[(i * PAIRS + q) % len(card_files)] for q in range(PAIRS)],
                                             [bg_files[(i * PAIRS + q) % len(bg_files)] for q in range(PAIRS)], bg_hw=BG_HW)
            return ds.lookahead((lambda i=i: item(i)) for i in range(k))

        b_steps = e2e_steps
        for _ in ds.host_tensor_batches(feed_bytes(3)):  # warm-up: the three pinned staging buffers of prepare_jpegs
            pass
        b_ms, _ = timed(feed_bytes, b_steps)
        e2e["from_bytes_objects"] = {"value": world * n_x * b_steps / (b_ms * 1e-3), "unit": UNIT, "steps": b_steps,
                                     "api": "prepare_jpeg_batch(list of bytes, list of bytes) inside the loop, one batch ahead on a helper thread (RanMtgEncDecDataset.lookahead): the 1024 bytes objects are copied into pinned staging by mtgv_gather_files"}

    # ---- context only: the training-loop call (pools resident, nothing uploaded), batches delivered in pinned host memory ----
    if e2e is not None:
        r_steps = min(e2e["steps"], 30)
        feed_resident = lambda k: iter([PAIRS] * k)  # noqa: E731
        for _ in ds.host_tensor_batches(feed_resident(3)):
            pass
        r_ms, res = timed(feed_resident, r_steps)
        e2e["resident_pool"] = {"value": world * n_x * r_steps / (r_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                                "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in res.values())), "steps": r_steps,
                                "api": "RanMtgEncDecDataset.host_tensor_batches fed with batch sizes (cards and backgrounds stay in HBM, the "
                                       "kernels of batch i overlap the download of batch i-1); not the e2e number: nothing is uploaded"}

    cpu_baseline = None
    if cpu_gen is not None:
        gen = cpu_gen
        n, wall = gen.step(args.cpu_pairs_per_worker, from_files=True)
        n2, wall2 = gen.step(max(1, args.cpu_pairs_per_worker // 4), from_files=False)
        gen.close()
        cpu_baseline = {"value": n / wall, "unit": UNIT, "cores": gen.workers, "kind": "port",
                        "sample": f"one batch of {args.cpu_pairs_per_worker} pairs per worker x {gen.workers} worker processes (cv2 threads=1), pools of "
                                  f"{args.pool_cards} cards / {args.pool_bgs} backgrounds as quality-{JPEG_Q} JPEG files in host memory, cv2.imdecode per drawn image "
                                  f"+ generator: {n} x-samples in {wall:.1f} s",
                        "from_decoded_arrays": n2 / wall2}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if args.epoch_samples > 0 else "weak",
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "stats": {"out_dtype": "fp16 NCHW", "setup_s": round(t_setup, 1), "virtual": n_virtual, "cropped": n_cropped,
                      "failed_samples": int((status != 0).sum()), "sampled_steps": n_keep, **drift},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if args.epoch_samples > 0:
            total = world * n_x * args.steps
            line["epoch"] = {"x_samples": total, "seconds_device_resident": ms_max * 1e-3,
                             "seconds_e2e_from_files": (total / e2e["value"]) if e2e else None,
                             "seconds_cpu_extrapolated": (total / cpu_baseline["value"]) if cpu_baseline else None,
                             "note": "BASELINE configs[4]: one epoch of x-samples split into contiguous global-index shards, one per GPU; the CPU "
                                     "figure extrapolates the bounded cpu_baseline sample"}
        if dist is not None:
            sys.stdout.flush()
            os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------- #
# detection workloads (BASELINE configs[2], configs[3]) - own measurements for profiles/      #
# --------------------------------------------------------------------------------------- #

_DET_STATE = {}


def _det_cpu_worker(args):
    wid, n_scenes, S, max_cards = args
    import random

    import cv2
    import numpy as np

    from oracle import det_oracle as DO

    cv2.setNumThreads(1)
    random.seed(77 + wid)
    np.random.seed(77 + wid)
    o = DO.DetOracle(_DET_STATE["cards"], _DET_STATE["bgs"], bg_size_hw=S, num_cards_min=1, num_cards_max=max_cards,
                     card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg")
    for _ in range(n_scenes):
        o.random()
    return n_scenes


def _clip_area(poly, S):
    """Area of a convex polygon (k,2) clipped to the frame [0,S]x[0,S] (Sutherland-Hodgman) - the A_k of SURVEY 8d."""
    import numpy as np

    pts = [tuple(p) for p in poly]
    for axis, bound, keep_less in ((0, 0.0, False), (0, float(S), True), (1, 0.0, False), (1, float(S), True)):
        out = []
        for i in range(len(pts)):
            p, q = pts[i], pts[(i + 1) % len(pts)]
            pin = p[axis] <= bound if keep_less else p[axis] >= bound
            qin = q[axis] <= bound if keep_less else q[axis] >= bound
            if pin:
                out.append(p)
            if pin != qin:
                t = (bound - p[axis]) / (q[axis] - p[axis])
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        pts = out
        if len(pts) < 3:
            return 0.0
    a = np.asarray(pts)
    return 0.5 * abs(float(np.dot(a[:, 0], np.roll(a[:, 1], -1)) - np.dot(a[:, 1], np.roll(a[:, 0], -1))))


def run_det(args):
    """BASELINE configs[2] / configs[3]: detection scenes (640^2 up to 8 cards, 1280^2 up to 32 cards, batch 256), the author's
    generator settings (od_datasets.py:861-868), kind='seg', photometrics on.  Same JSON contract as the encoder line."""
    import multiprocessing as mp

    import numpy as np
    import torch

    from mtgvision_b200 import abi, synth
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.od_datasets import Gen

    S, max_cards, batch = (640, 9, 256) if args.workload == "det640" else (1280, 33, 256)
    workers = os.cpu_count() or 1
    cards = synth.make_card_pool(args.pool_cards, workers=workers)
    bgs = synth.make_bg_pool(args.pool_bgs, workers=workers)
    cpu_pool = None
    if not args.no_cpu_baseline:  # forked before CUDA exists in this process
        _DET_STATE["cards"] = [cards.images[k] for k in range(32)]
        _DET_STATE["bgs"] = bgs[:32]
        cpu_pool = mp.get_context("fork").Pool(workers)
        cpu_pool.map(_det_cpu_worker, [(i, 1, S, max_cards) for i in range(workers)])
    gen = Gen(bg_size_hw=S, num_cards_min=1, num_cards_max=max_cards, card_min_visible_ratio=0.5,
              card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",
              mtg_ds=SyntheticBgFgMtgImages(pool=cards), bg_ds=IlsvrcImages(images=bgs), seed=7)
    ctx = gen.ctx
    dev = torch.device("cuda", 0)
    for _ in range(max(args.warmup, 3)):
        gen.random_batch(batch)
    torch.cuda.synchronize()
    # live timing of the pixel kernel(s) of one mtgv_det_batch call, CUDA events on the launch stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(0)
    clocks.start()
    time.sleep(0.25)
    l0 = ctx.launch_count()
    alg = 0.0
    keep = []
    t0 = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        first = gen._shards.next_first(batch)
        tape = ctx.sample_det_tape(gen.seed, first, batch)
        params, accepted, keypoints, labels, counts = ctx.det_place(tape)
        ev[k][0].record()
        image = ctx.det_batch(params, abi.OUT_U8)
        ev[k][1].record()
        if k >= args.steps - 2:
            keep.append((keypoints, counts))
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    launches = int(ctx.launch_count() - l0)
    clk = clocks.stop(t0, t1)
    ms = e0.elapsed_time(e1)
    k_ms = [a.elapsed_time(b) for a, b in ev]
    k_avg = sum(k_ms) / len(k_ms)
    # algorithmic bytes: bg source + output + per placed card min(card bytes, 12 * visible quad area)
    n_placed = 0
    for keypoints, counts in keep:
        kp, cnt = keypoints.cpu().numpy(), counts.cpu().numpy()
        for s_ in range(batch):
            alg += BYTES_BG + 3 * S * S
            for q in range(int(cnt[s_])):
                quad = kp[s_, q][[0, 1, 2, 7]]  # the card box of the 8-vertex seg polygon
                alg += min(BYTES_CARD, 12.0 * _clip_area(quad, S))
                n_placed += 1
    alg_per_launch = alg / len(keep)
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    achieved = alg_per_launch / (k_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_det_pixels (all passes of one mtgv_det_batch call; k_det_lists included)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch, "kernel_ms": k_avg, "kernel_share_of_step": k_avg * args.steps / ms,
                "placed_cards_per_scene": n_placed / (len(keep) * batch)}
    # e2e: the public call, host tensors out (uint8 scenes + labels to pinned memory every step; nothing is uploaded: the
    # generator's inputs are the resident pools); kernels of batch i overlap the download of batch i-1 (Gen.host_batches)
    for _ in gen.host_batches([batch] * 3):  # untimed: pinned targets
        pass
    torch.cuda.synchronize()
    tw0 = time.perf_counter()
    chk, res = 0, None
    for res in gen.host_batches([batch] * args.steps):
        chk += int(res["counts"][0])  # the consumer touches every batch on the host
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - tw0) * 1e3 / args.steps
    d2h = sum(v.numel() * v.element_size() for v in res.values())
    # the dataset writer's path (create_yolo_obb_dataset): scenes generated AND JPEG-encoded on the device, files to pinned host memory
    ctx.encode_jpegs_host(gen.random_batch(batch)["image"])  # untimed: allocates the encoder's buffers
    torch.cuda.synchronize()
    tw0 = time.perf_counter()
    file_bytes = 0
    for _ in range(args.steps):
        files = ctx.encode_jpegs_host(gen.random_batch(batch)["image"])
        file_bytes += sum(len(f) for f in files)
    writer_ms = (time.perf_counter() - tw0) * 1e3 / args.steps
    cpu = None
    if cpu_pool is not None:
        per = 6 if S == 640 else 2
        tw0 = time.perf_counter()
        n = sum(cpu_pool.map(_det_cpu_worker, [(i, per, S, max_cards) for i in range(workers)]))
        wall = time.perf_counter() - tw0
        cpu_pool.close()
        cpu = {"value": n / wall, "unit": "scenes/s", "cores": workers, "kind": "port",
               "sample": f"{n} scenes in {wall:.1f} s: oracle port of generate_synthetic_image (cv2 warps + composite + restated placement test and "
                         "photometric subset), one process per core, pools of 32 cards / 32 backgrounds"}
    print(json.dumps({"metric": f"detection scenes/sec {S}x{S}", "value": batch * args.steps / (ms * 1e-3), "unit": "scenes/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"detection scenes {S}x{S}, up to {max_cards - 1} cards, batch {batch}, overlap rejection sampling, seg polygon "
                                             "labels, photometric augments (BASELINE configs[%d])" % (2 if S == 640 else 3),
                                 "batch": batch, "max_cards": max_cards - 1, "kind": "seg", "out": "uint8 NCHW",
                                 "card_pool": args.pool_cards, "bg_pool": args.pool_bgs,
                                 "l2": "inputs larger than L2: 256 scenes read ~1000 card images of a multi-GB resident pool; no flush"},
                      "clocks": clk, "roofline": roofline,
                      "e2e": {"value": batch / (e2e_ms * 1e-3), "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h),
                              "ms_per_step": e2e_ms,
                              "api": "Gen.host_batches: random_batch(256) with image / keypoints / labels / counts / accepted delivered in pinned host "
                                     "memory, kernels of batch i overlapping the download of batch i-1; host wall clock over all steps incl. fill and "
                                     "drain (the generator has no per-step host input: cards and backgrounds are the resident pools)"},
                      "writer": {"value": batch / (writer_ms * 1e-3), "unit": "scenes/s", "ms_per_batch": writer_ms,
                                 "d2h_bytes_per_step": file_bytes // args.steps,
                                 "api": "Gen.random_batch + Context.encode_jpegs_host: scene generation and JPEG encode on the device, "
                                        "files (cv2.imwrite's bytes) in pinned host memory; host wall clock"},
                      "gpu_launches": launches, "cpu_baseline": cpu}), flush=True)


def run_dewarp(args):
    """Serving-side dewarp (od_export.py:95-111): 32 detected cards of one 1080p uint8 frame per step."""
    import cv2
    import numpy as np
    import torch

    from mtgvision_b200.od_export import Dewarper

    rng = np.random.default_rng(3)
    H, W, n = 1080, 1920, 32
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    quads = []
    for _ in range(n):
        cx, cy, s, a = rng.uniform(0.2, 0.8) * W, rng.uniform(0.2, 0.8) * H, rng.uniform(60, 200), rng.uniform(0, 6.283)
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        quads.append((np.array([[-0.7, -1.0], [0.7, -1.0], [0.7, 1.0], [-0.7, 1.0]]) * s) @ rot.T + (cx, cy))
    quads = np.stack(quads)
    dw = Dewarper(device=0)
    fd = torch.from_numpy(frame).cuda()
    qd = torch.from_numpy(quads).cuda()
    for _ in range(args.warmup):
        dw.extract_dewarped(fd, qd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = dw.extract_dewarped(fd, qd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: pinned frame in, pinned crops out
    hf = torch.from_numpy(frame).pin_memory()
    ho = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fd.copy_(hf, non_blocking=True)
        ho.copy_(dw.extract_dewarped(fd, qd), non_blocking=True)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    cv2.setNumThreads(1)
    dst = ((1 + 0.05) * np.asarray([[0, 0], [128, 0], [128, 192], [0, 192]]) - 0.025 * np.asarray([128, 192])).astype(np.float32)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        for q in quads:
            cv2.warpPerspective(frame, cv2.getPerspectiveTransform(q.astype(np.float32), dst), (128, 192))
    cpu = reps * n / (time.perf_counter() - t0)
    out_bytes = n * 192 * 128 * 3
    print(json.dumps({"metric": "dewarped cards/sec (1080p uint8 frame, 32 cards, 192x128 out)", "value": n / (ms * 1e-3), "unit": "cards/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "dtype": "u8",
                      "config": {"workload": "dewarp", "frame": [H, W, 3], "cards": n},
                      "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "cards/s", "h2d_bytes_per_step": int(frame.nbytes), "d2h_bytes_per_step": out_bytes},
                      "cpu_baseline": {"value": cpu, "unit": "cards/s", "cores": 1, "kind": "reference",
                                       "sample": f"cv2.getPerspectiveTransform + cv2.warpPerspective, {reps * n} cards, 1 thread"}}), flush=True)


def run_jpeg(args):
    """Pool ingest from JPEG files (imread_float util/image.py:107-114): --jpeg-files background-sized files per step."""
    import cv2
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mtgvision_b200 import synth
    from mtgvision_b200.context import Context

    n, rst = args.jpeg_files, args.jpeg_rst
    files = []
    for j in range(min(n, 256)):
        src = synth.synth_bg(j)
        ok, buf = cv2.imencode(".jpg", src[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        files.append(buf.tobytes())
    files = [files[j % len(files)] for j in range(n)]
    ctx = Context(0)
    batch = ctx.prepare_jpegs(files)
    out = torch.empty(int(batch["out_off"][-1]), dtype=torch.uint8, device="cuda")
    for _ in range(args.warmup):
        ctx.decode_prepared(batch, out)
    torch.cuda.synchronize()
    kms = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.decode_prepared(batch, out)
        kms += ctx.jpeg_last_kernel_ms()  # waits for the batch
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    kms /= args.steps
    ms = float(kms.sum())
    flat, off, hw = out, batch["out_off"][:-1], batch["hw"]
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    for f in files[:256]:
        cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR_RGB)
    cpu = min(n, 256) / (time.perf_counter() - t0)
    ref = cv2.imdecode(np.frombuffer(files[-1], np.uint8), cv2.IMREAD_COLOR_RGB)
    h, w = hw[-1]
    same = bool(np.array_equal(flat[off[-1]:].cpu().numpy().reshape(h, w, 3), ref))
    file_bytes = sum(len(f) for f in files)
    print(json.dumps({"metric": "JPEG files decoded into the pool format per second (375x500, q90, 4:2:0)", "value": n / (ms * 1e-3),
                      "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "dtype": "i32",
                      "config": {"workload": "jpeg", "files": n, "restart_interval": rst, "mean_file_bytes": file_bytes // n,
                                 "kernel_ms": {"k_jpeg_entropy": kms[0], "k_jpeg_idct": kms[1], "k_jpeg_color": kms[2]},
                                 "timing": "value: CUDA events around the three kernels inside the library; e2e: host wall clock of "
                                           "mtgv_decode_jpeg_batch from pinned file bytes incl. host marker parse and file upload"},
                      "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": file_bytes, "d2h_bytes_per_step": 0},
                      "bit_exact_vs_cv2": same,
                      "cpu_baseline": {"value": cpu, "unit": "images/s", "cores": 1, "kind": "reference",
                                       "sample": f"cv2.imdecode of {min(n, 256)} of the same files, 1 thread"}}), flush=True)


def run_jpegenc(args):
    """Dataset-writer image encode (save_sample -> imwrite, od_datasets.py:829-831): 256 scenes of 640x640 per step."""
    import cv2
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mtgvision_b200 import synth
    from mtgvision_b200.context import Context

    n, S = 256, 640
    rng = np.random.default_rng(0)
    scenes = np.empty((n, S, S, 3), np.uint8)
    for k in range(n):  # low-pass background with sharp-edged patches: scene-like statistics
        bg = cv2.resize(synth.synth_bg(k % 64), (S, S), interpolation=cv2.INTER_LINEAR)
        for _ in range(4):
            y, x, h, w = rng.integers(0, S - 200), rng.integers(0, S - 200), rng.integers(60, 200), rng.integers(60, 200)
            bg[y:y + h, x:x + w] = synth.synth_card(int(rng.integers(0, 16)))[:h, :w]
        scenes[k] = bg
    ctx = Context(0)
    dev = torch.from_numpy(scenes).cuda().permute(0, 3, 1, 2).contiguous()  # mtgv_det_batch's layout
    for _ in range(args.warmup):
        out, lens = ctx.encode_jpegs_device(dev)
    torch.cuda.synchronize()
    kms = np.zeros(2)
    for _ in range(args.steps):
        out, lens = ctx.encode_jpegs_device(dev)
        kms += ctx.jpeg_encode_last_kernel_ms()
    kms /= args.steps
    ms = float(kms.sum())
    ctx.encode_jpegs_host(dev)  # untimed: allocates the compact and pinned buffers the context keeps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        files = ctx.encode_jpegs_host(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    ref = [cv2.imencode(".jpg", cv2.cvtColor(scenes[k], cv2.COLOR_RGB2BGR))[1].tobytes() for k in range(64)]
    cpu = 64 / (time.perf_counter() - t0)
    same = all(files[k].tobytes() == ref[k] for k in range(64))
    out_bytes = sum(len(f) for f in files)
    print(json.dumps({"metric": "scenes JPEG-encoded per second (640x640, quality 95, 4:2:0)", "value": n / (ms * 1e-3), "unit": "images/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "dtype": "i32",
                      "config": {"workload": "jpegenc", "scenes": n, "size": [S, S], "mean_file_bytes": out_bytes // n,
                                 "kernel_ms": {"k_jpegenc_dct": kms[0], "huffman stage (bit buffer clear, k_jpegenc_size, _scan, _emit, _stuff)": kms[1]},
                                 "timing": "value: CUDA events around the two kernels inside the library; e2e: host wall clock of "
                                           "Context.encode_jpegs_host (device uint8 NCHW scenes in, the files in pinned host memory out)"},
                      "e2e": {"value": n / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": out_bytes},
                      "bytes_equal_cv2": same,
                      "cpu_baseline": {"value": cpu, "unit": "images/s", "cores": 1, "kind": "reference",
                                       "sample": "cv2.imencode of 64 of the same scenes, 1 thread"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jpeg-rst", type=int, default=0)
    ap.add_argument("--jpeg-files", type=int, default=2048)
    ap.add_argument("--workload", default="encoder", choices=["encoder", "det640", "det1280", "dewarp", "jpeg", "jpegenc"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--pool-cards", type=int, default=2048)
    ap.add_argument("--pool-bgs", type=int, default=1024)
    ap.add_argument("--cpu-pairs-per-worker", type=int, default=384)
    ap.add_argument("--ref-pairs-per-worker", type=int, default=128)
    ap.add_argument("--e2e-max-steps", type=int, default=30)
    ap.add_argument("--epoch-samples", type=int, default=0,
                    help="BASELINE configs[4]: generate this many x-samples in total, split into contiguous shards over the ranks "
                         "(strong scaling): overrides --steps with ceil(samples / (1024 * world)) and times the e2e path over as many steps")
    ap.add_argument("--no-e2e-raw", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload == "dewarp":
        run_dewarp(args)
    elif args.workload == "jpeg":
        run_jpeg(args)
    elif args.workload == "jpegenc":
        run_jpegenc(args)
    elif args.workload != "encoder":
        run_det(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.epoch_samples > 0:
            world = int(os.environ.get("WORLD_SIZE", "1"))
            args.steps = -(-args.epoch_samples // (2 * PAIRS * world))
            args.e2e_max_steps = args.steps
        run_cuda(args)


if __name__ == "__main__":
    main()
