#!/usr/bin/env python
"""Benchmark of the picha hot path on B200 (contract: see the build brief; summary in DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's own C++ on the host cores

A step is one pass of the hot path over one batch of synthetic images, one rank per GPU, each
rank owning its own block of images (weak scaling, no data-path collective).  Rank 0 prints
ONE JSON line.

  value     whole-job output Mpix/s, batch resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the C-ABI host entry point with pinned HOST buffers:
            H2D + kernel + D2H inside the timed region
  roofline  algorithmic bytes (payload read + written) of the dominant kernel per launch /
            its launch time, against the measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own C++ (oracle/_ref; else the C port) on the box's host
            cores over a bounded sample of the same workload, and its outputs compared with
            the GPU's for the same images (the parity figures)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (src w, h, dst w, h, pixel, filter name or None (= default cubic), filterScale, images/GPU, seed)
RESIZE_WORKLOADS = {
    "cfg3": dict(sw=3840, sh=2160, dw=960, dh=540, pixel="rgba", filter="lanczos", width=1.0, batch=256, seed=1237,
                 name="cfg3: 4K (3840x2160) rgba -> 960x540 lanczos, batch 256 per GPU"),
    "cfg4": dict(sw=2048, sh=2048, dw=4096, dh=4096, pixel="r16g16b16a16", filter="mitchel", width=1.0, batch=16,
                 seed=1238, name="cfg4: 2048x2048 r16g16b16a16 -> 4096x4096 mitchel, batch 16 per GPU"),
    "cfg5": dict(sw=1920, sh=1080, dw=256, dh=256, pixel="rgb", filter=None, width=0.70, batch=1024, seed=1239,
                 name="cfg5: 1080p rgb -> 256x256 cubic@0.70 thumbnails, 1024 per GPU (8192 on 8)"),
}
# resize, then convert, in one kernel (picha_b200_resize_convert): the thumbnail pipeline straight to grey
RESIZE_WORKLOADS["cfg5-grey"] = dict(RESIZE_WORKLOADS["cfg5"], to="grey", seed=1239,
                                     name="cfg5 fused: 1080p rgb -> 256x256 cubic@0.70 -> grey in one kernel, 1024 per GPU")
CONVERT_WORKLOADS = {
    "cfg2": dict(w=1920, h=1080, src="rgba", dst="rgb", batch=256, seed=1236,
                 name="cfg2: 1080p rgba -> rgb colorConvert, batch 256 per GPU"),
    "cfg2-grey": dict(w=1920, h=1080, src="rgba", dst="grey", batch=256, seed=1236,
                      name="cfg2: 1080p rgba -> grey colorConvert, batch 256 per GPU"),
    "cfg2-greya": dict(w=1920, h=1080, src="rgba", dst="greya", batch=256, seed=1236,
                       name="cfg2: 1080p rgba -> greya colorConvert, batch 256 per GPU"),
}
PIXEL_BYTES = {"rgb": 3, "rgba": 4, "grey": 1, "greya": 2, "r16": 2, "r16g16": 4, "r16g16b16": 6, "r16g16b16a16": 8}


# ---- distributed plumbing (timing barrier and max over ranks only; images never move) ----------

def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _reduce(value, op_name):
    import torch
    dist = _dist()
    if dist is None:
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(ms):
    return _reduce(ms, "MAX")


def sum_over_ranks(v):
    return type(v)(_reduce(v, "SUM"))


def barrier():
    dist = _dist()
    if dist is not None:
        dist.barrier()


# ---- helpers ---------------------------------------------------------------------------------------

def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def profiled_traffic(workload):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi takes 0.1 - 1 s to print its first row and the timed region lasts tens of milliseconds:
            # do not start timing before the sampler is actually sampling
            deadline = time.perf_counter() + 5.0
            while not self.rows and self.proc.poll() is None and time.perf_counter() < deadline:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.1)]
        if not rows:
            rows = [r for (_, r) in self.rows]
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[6]))
            except Exception:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def resolve_filter(P, w):
    from picha_b200 import _native as N
    tag, width = ctypes.c_int(0), ctypes.c_float(0)
    has = w["filter"] is not None
    N.check(N.lib.picha_b200_resolve_resize_options(int(has), N.FILTERS.index(w["filter"]) if has else 0, 1,
                                                    float(w["width"]), ctypes.byref(tag), ctypes.byref(width)))
    return tag.value, width.value


# ---- the GPU arm -------------------------------------------------------------------------------------

def time_device_steps(fn, steps, warmup):
    """K steps, each timed with CUDA events on torch's current stream (the stream the library
    launches on), after W warm-up steps; barrier + synchronize on both sides.  Returns the list of
    per-step milliseconds and the whole-region milliseconds."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return per, evs[0].elapsed_time(evs[-1])


def run_resize_workload(key, args, rank, world, local_rank, with_e2e=True, with_cpu=True, batch=None, extras=False,
                        with_latency=False):
    import numpy as np
    import torch
    import picha_b200 as P
    from picha_b200 import _native as N
    from picha_b200 import device as D

    w = RESIZE_WORKLOADS[key]
    n = batch or w["batch"]
    bpp = PIXEL_BYTES[w["pixel"]]
    src = D.DeviceBatch(n, w["sw"], w["sh"], w["pixel"])
    dst = D.DeviceBatch(n, w["dw"], w["dh"], w.get("to", w["pixel"]))
    src.fill_synthetic(w["seed"], first_image=rank * n)     # every rank owns its own block of images
    torch.cuda.synchronize()
    tag, fwidth = resolve_filter(P, w)
    s0, d0 = src.cimage(), dst.cimage()
    stream = torch.cuda.current_stream().cuda_stream

    cs = (ctypes.c_float * 3)()
    N.lib.picha_b200_resolve_color_settings(float("nan"), float("nan"), float("nan"), cs)

    def step():
        if "to" in w:
            N.check(N.lib.picha_b200_resize_convert_device(n, ctypes.byref(s0), src.step, ctypes.byref(d0), dst.step, tag, fwidth,
                                                           cs[0], cs[1], cs[2], 0, stream))
        else:
            N.check(N.lib.picha_b200_resize_device(n, ctypes.byref(s0), src.step, ctypes.byref(d0), dst.step, tag, fwidth,
                                                   0, stream))

    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = P.launch_count()
    t0 = time.perf_counter()
    per, total_ms = time_device_steps(step, args.steps, args.warmup)
    t1 = time.perf_counter()
    warm_launches = (P.launch_count() - launches0)
    launches = warm_launches * args.steps // (args.steps + args.warmup)
    # clocks under load: from the first warm-up step to the end of the timed region (the timed part alone
    # lasts tens of milliseconds, shorter than nvidia-smi's sampling period)
    clocks = sampler.stop(t0, t1) if sampler else None

    ms_per_step = max_over_ranks(total_ms / args.steps)
    out_mpix_rank = n * w["dw"] * w["dh"] / 1e6
    total_mpix = sum_over_ranks(out_mpix_rank)
    value = total_mpix / (ms_per_step / 1e3)

    algo_bytes = n * (w["sw"] * w["sh"] * bpp + w["dw"] * w["dh"] * PIXEL_BYTES[w.get("to", w["pixel"])])   # payload read + written per step
    launch_ms = sum(per) / len(per)                                       # the step's resize launches, back to back
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (launch_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": profiled_traffic(key), "peak_source": peak_src,
                "kernel": f"resize kernel, {max(1, launches // max(1, args.steps))} launch(es) per step (one per group of row "
                          "bands); bytes and time are per step",
                "algorithmic_bytes_per_launch": algo_bytes,
                "launch_ms": round(launch_ms, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}

    res = {"value": value, "ms_per_step": ms_per_step, "roofline": roofline, "gpu_launches": launches,
           "clocks": clocks, "images_per_gpu": n}

    if "to" in w:
        with_e2e = with_cpu = with_latency = False      # (device-resident line only; parity of the fused op is in tests/)
    if with_e2e:
        big = key == "cfg5"            # thumbnails: enough images per step for the library to cut chunks of one launch each
        res["e2e"] = e2e_resize(w, args, rank, local_rank, src, tag, fwidth, batch=min(256, n) if big else None)
        if extras:
            ceiling = h2d_ceiling(local_rank)
            res["e2e"]["h2d_ceiling_gbs_per_gpu"] = ceiling
            res["e2e"]["frac_of_h2d_ceiling"] = round(res["e2e"]["h2d_gbs_per_gpu"] / ceiling, 3) if ceiling else None
            if world == 1:
                pg = e2e_resize(w, args, rank, local_rank, src, tag, fwidth, pinned=False, batch=min(64, n) if big else 16,
                                steps=max(2, args.steps // 4))
                res["e2e_pageable"] = {k: pg.get(k) for k in ("value", "unit", "ms_per_step", "images_per_step_per_gpu", "h2d_gbs_per_gpu", "api")}
                res["call_latency"] = [call_latency(w, src, tag, fwidth, t) for t in (4, 16)]
            else:
                barrier()
                if rank == 0:
                    res["e2e_sharder"] = sharder_e2e(w, args, src, tag, fwidth, world, per_gpu=128 if big else 16)
                barrier()
    if with_latency and rank == 0 and world == 1 and "call_latency" not in res:
        res["call_latency"] = [call_latency(w, src, tag, fwidth, t, calls=4) for t in (4, 16)]
    if with_cpu and rank == 0 and world == 1:
        res["cpu_baseline"] = cpu_baseline_resize(w, src, dst, args)
    del src, dst
    torch.cuda.empty_cache()
    return res


class HostImages:
    """n copies-by-value of the workload's synthetic images in HOST memory (pinned through the library's allocator,
    or ordinary pageable numpy memory -- what a Node Buffer is), plus destination buffers, as C-ABI image arrays."""

    def __init__(self, w, src, n, pinned=True):
        import numpy as np
        from picha_b200 import _native as N
        self.N, self.n, self.pinned = N, n, pinned
        bpp = PIXEL_BYTES[w["pixel"]]
        sstride, dstride = w["sw"] * bpp, (w["dw"] * bpp + 3) & ~3
        sbytes, dbytes = sstride * w["sh"], dstride * w["dh"]
        self.ptrs = []
        if pinned:
            hp_src, hp_dst = N.lib.picha_b200_host_alloc(n * sbytes), N.lib.picha_b200_host_alloc(n * dbytes)
            self.ptrs = [hp_src, hp_dst]
            if not hp_src or not hp_dst:
                raise MemoryError("pinned allocation failed")
            host_src = np.ctypeslib.as_array(ctypes.cast(hp_src, ctypes.POINTER(ctypes.c_ubyte)), shape=(n * sbytes,))
        else:
            self.keep = (np.empty(n * sbytes, np.uint8), np.empty(n * dbytes, np.uint8))
            host_src = self.keep[0]
            hp_src, hp_dst = self.keep[0].ctypes.data, self.keep[1].ctypes.data
        distinct = min(n, src.n, 8)
        for i in range(n):   # the same synthetic images the device batch holds, now in host memory
            if i < distinct:
                host_src[i * sbytes:(i + 1) * sbytes].reshape(w["sh"], sstride)[:] = src.image(i).rows()
            else:
                host_src[i * sbytes:(i + 1) * sbytes] = host_src[(i % distinct) * sbytes:(i % distinct + 1) * sbytes]
        pix = N.PIXELS.index(w["pixel"])
        self.srcs = (N.CImage * n)(*[N.CImage(hp_src + i * sbytes, sstride, w["sw"], w["sh"], pix) for i in range(n)])
        self.dsts = (N.CImage * n)(*[N.CImage(hp_dst + i * dbytes, dstride, w["dw"], w["dh"], pix) for i in range(n)])
        self.h2d, self.d2h = n * w["sw"] * w["sh"] * bpp, n * w["dw"] * w["dh"] * bpp

    def free(self):
        for p in self.ptrs:
            if p:
                self.N.lib.picha_b200_host_free(p)
        self.ptrs = []


def e2e_resize(w, args, rank, local_rank, src, tag, fwidth, pinned=True, batch=None, steps=None):
    """The same metric through picha_b200_resize_batch with HOST buffers: every step copies its inputs
    host->device and its results device->host inside the timed region."""
    import torch
    import picha_b200 as P
    from picha_b200 import _native as N

    n = batch or min(args.e2e_batch, src.n)
    steps = steps or args.steps
    try:
        host = HostImages(w, src, n, pinned)
    except MemoryError as e:
        return {"value": None, "unit": "Mpix/s", "error": str(e)}
    try:
        def step():
            N.check(N.lib.picha_b200_resize_batch(n, host.srcs, host.dsts, tag, fwidth, 0, local_rank))

        for _ in range(max(1, min(args.warmup, 3))):
            step()
        torch.cuda.synchronize()
        barrier()
        l0 = P.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()                      # blocking: returns when the results are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches = (P.launch_count() - l0) // steps
        barrier()
        ms = max_over_ranks(dt * 1e3 / steps)
        mpix = sum_over_ranks(n * w["dw"] * w["dh"] / 1e6)
        return {"value": round(mpix / (ms / 1e3), 1), "unit": "Mpix/s",
                "h2d_bytes_per_step": host.h2d, "d2h_bytes_per_step": host.d2h,
                "ms_per_step": round(ms, 3), "images_per_step_per_gpu": n, "launches_per_step": launches,
                "h2d_gbs_per_gpu": round(host.h2d / (ms / 1e3) / 1e9, 1),
                "api": "picha_b200_resize_batch (C-ABI, " + ("pinned host buffers from picha_b200_host_alloc)" if pinned
                                                              else "pageable host buffers, staged by the library)")}
    finally:
        host.free()


def h2d_ceiling(local_rank, mib=256, copies=8):
    """The platform's ceiling for the e2e numbers: raw pinned host -> device copies on every rank at once
    (cudaMemcpyAsync through torch), GB/s per GPU as the max time over ranks sees it."""
    import torch
    nbytes = mib << 20
    hsrc = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    ddst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
    ddst.copy_(hsrc, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(copies):
        ddst.copy_(hsrc, non_blocking=True)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    return round(nbytes * copies / dt / 1e9, 1)


def sharder_e2e(w, args, src, tag, fwidth, gpus, per_gpu=16):
    """One process, every GPU of the box: picha_b200_resize_batch(..., device=-1), the library's own sharder
    (one host thread per GPU inside the call), pinned host buffers.  Rank 0 only; the other ranks wait."""
    import torch
    from picha_b200 import _native as N
    n = per_gpu * gpus
    try:
        host = HostImages(w, src, n, True)
    except MemoryError as e:
        return {"value": None, "error": str(e)}
    try:
        def step():
            N.check(N.lib.picha_b200_resize_batch(n, host.srcs, host.dsts, tag, fwidth, 0, -1))
        step()
        steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return {"value": round(n * w["dw"] * w["dh"] / 1e6 / (ms / 1e3), 1), "unit": "Mpix/s", "ms_per_step": round(ms, 3),
                "images_per_step": n, "gpus": gpus, "h2d_gbs_total": round(host.h2d / (ms / 1e3) / 1e9, 1),
                "api": "picha_b200_resize_batch(device=-1): one process, one host thread per GPU inside the library"}
    finally:
        host.free()


def call_latency(w, src, tag, fwidth, threads, calls=8):
    """What picha.resize does (src/resize.cc:362-364): one image per call, `threads` calls in flight from as many
    host threads, ordinary (pageable) buffers.  Median latency of a call and the aggregate rate."""
    import numpy as np
    from picha_b200 import _native as N
    host = HostImages(w, src, threads, pinned=False)
    lat, errs = [[] for _ in range(threads)], []
    gate = threading.Barrier(threads + 1)

    def work(i):
        try:
            for c in range(calls + 2):
                if c == 2:
                    gate.wait()          # every thread has its lane (stream, staging memory) and the plan: start the clock
                t0 = time.perf_counter()
                N.check(N.lib.picha_b200_resize(ctypes.byref(host.srcs[i]), ctypes.byref(host.dsts[i]), tag, fwidth))
                if c >= 2:
                    lat[i].append(time.perf_counter() - t0)
        except Exception as e:   # pragma: no cover
            errs.append(repr(e))
            gate.abort()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    try:
        gate.wait()
    except threading.BrokenBarrierError:
        pass
    t0 = time.perf_counter()
    [t.join() for t in ts]
    wall = time.perf_counter() - t0
    if errs:
        return {"error": errs[0]}
    flat = sorted(x for l in lat for x in l)
    return {"threads": threads, "median_ms": round(flat[len(flat) // 2] * 1e3, 3), "p90_ms": round(flat[int(len(flat) * 0.9)] * 1e3, 3),
            "images_per_s": round(threads * calls / wall, 1), "out_mpix_per_s": round(threads * calls * w["dw"] * w["dh"] / 1e6 / wall, 1),
            "calls": len(flat), "buffers": "pageable"}


def _cpu_pool_rate(fn, items, threads, passes=2):
    """Best-of-`passes` wall time of fn over `items` on a thread pool (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    best = None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(passes):
            t0 = time.perf_counter()
            list(ex.map(fn, items))
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
    return best


def cpu_baseline_resize(w, src, dst, args, per_thread=2, compare=True):
    """The reference's resizeImage on the box's host cores, one image per thread (what libuv's pool
    does for picha.resize), on a bounded sample of the same synthetic images; the first few outputs
    are also compared with the GPU's (parity)."""
    import numpy as np
    import oracle as O          # the one place bench.py touches oracle/: the CPU baseline leg

    kind = "reference" if O.have_ref() else "port"
    impl = "ref" if kind == "reference" else "port"
    threads = max(1, min(os.cpu_count() or 1, args.cpu_threads))
    distinct = min(4, src.n)
    imgs = [src.image(i) for i in range(distinct)]
    bufs = [np.ascontiguousarray(im.data) for im in imgs]
    filt = w["filter"] or "cubic"
    outs = {}

    def one(i):
        k = i % distinct
        d, ds = O.resize(bufs[k], imgs[k].stride, w["sw"], w["sh"], w["pixel"], w["dw"], w["dh"], filt, w["width"], impl)
        if i < distinct:
            outs[i] = (d, ds)

    t1 = _cpu_pool_rate(one, list(range(distinct)), 1, passes=1) / distinct      # single thread, s per image
    n_items = threads * per_thread
    dt = _cpu_pool_rate(one, list(range(n_items)), threads)
    mpix = n_items * w["dw"] * w["dh"] / 1e6
    res = {"value": round(mpix / dt, 2), "unit": "Mpix/s", "cores": threads, "kind": kind,
           "sample": f"{n_items} images ({per_thread} per thread, {distinct} distinct synthetic images of the workload), "
                     f"best of 2 passes",
           "one_thread_ms_per_image": round(t1 * 1e3, 2), "one_thread_mpix_s": round(w["dw"] * w["dh"] / 1e6 / t1, 2)}
    if compare:
        mx, mean = 0, 0.0
        for i in range(distinct):
            got = dst.image(i)
            want = O.channels(outs[i][0], outs[i][1], w["dw"], w["dh"], w["pixel"]).astype(np.int64)
            g = np.ascontiguousarray(got.rows())
            if w["pixel"].startswith("r16"):
                g = g.view(np.uint16)
            d = np.abs(g.astype(np.int64) - want)
            mx = max(mx, int(d.max()))
            mean = max(mean, float(d.mean()))
        res["parity_vs_gpu"] = {"images": distinct, "max_abs_diff": mx, "worst_mean_abs_diff": round(mean, 5),
                                "bound": "max <= 1, mean <= 0.05"}
    return res


def run_convert_workload(key, args, rank, world, local_rank, with_e2e=True, with_cpu=True):
    import numpy as np
    import torch
    import picha_b200 as P
    from picha_b200 import _native as N
    from picha_b200 import device as D

    w = CONVERT_WORKLOADS[key]
    n = w["batch"]
    src = D.DeviceBatch(n, w["w"], w["h"], w["src"])
    dst = D.DeviceBatch(n, w["w"], w["h"], w["dst"])
    src.fill_synthetic(w["seed"], first_image=rank * n)
    torch.cuda.synchronize()
    cs = (ctypes.c_float * 3)()
    nan = float("nan")
    N.lib.picha_b200_resolve_color_settings(nan, nan, nan, cs)
    s0, d0 = src.cimage(), dst.cimage()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        N.check(N.lib.picha_b200_color_convert_device(n, ctypes.byref(s0), src.step, ctypes.byref(d0), dst.step,
                                                      cs[0], cs[1], cs[2], stream))

    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = P.launch_count()
    t0 = time.perf_counter()
    per, total_ms = time_device_steps(step, args.steps, args.warmup)
    t1 = time.perf_counter()
    launches = (P.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_per_step = max_over_ranks(total_ms / args.steps)
    value = sum_over_ranks(n * w["w"] * w["h"] / 1e6) / (ms_per_step / 1e3)
    algo = n * w["w"] * w["h"] * (PIXEL_BYTES[w["src"]] + PIXEL_BYTES[w["dst"]])
    launch_ms = sum(per) / len(per)
    peak, peak_src = measured_peak()
    achieved = algo / (launch_ms / 1e3) / 1e9
    res = {"value": value, "ms_per_step": ms_per_step, "gpu_launches": launches, "clocks": clocks, "images_per_gpu": n,
           "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": profiled_traffic(key), "peak_source": peak_src,
                        "kernel": "color convert (one launch per step)", "algorithmic_bytes_per_launch": algo,
                        "launch_ms": round(launch_ms, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}}
    if with_cpu and rank == 0 and world == 1:
        import oracle as O
        kind = "reference" if O.have_ref() else "port"
        impl = "ref" if kind == "reference" else "port"
        threads = max(1, min(os.cpu_count() or 1, args.cpu_threads))
        img = src.image(0)
        buf = np.ascontiguousarray(img.data)
        out = {}

        def one(i):
            d, ds = O.color_convert(buf, img.stride, w["w"], w["h"], w["src"], w["dst"], None, impl)
            if i == 0:
                out[0] = (d, ds)

        n_items = threads * 8
        dt = _cpu_pool_rate(one, list(range(n_items)), threads)
        same = bool(np.array_equal(dst.image(0).rows(), O.payload(out[0][0], out[0][1], w["w"], w["h"], w["dst"])))
        res["cpu_baseline"] = {"value": round(n_items * w["w"] * w["h"] / 1e6 / dt, 1), "unit": "Mpix/s", "cores": threads,
                               "kind": kind, "sample": f"{n_items} conversions of one synthetic 1080p image, best of 2 passes",
                               "parity_vs_gpu": {"bit_exact": same}}
    del src, dst
    torch.cuda.empty_cache()
    return res


def cfg1_latency(args):
    """BASELINE cfg1 (README example): rgb 50x50 -> 100x100, default options, one blocking call through the
    C-ABI with ordinary (pageable) host buffers.  Launch-latency bound: reported as latency, no roofline."""
    import numpy as np
    from picha_b200 import _native as N
    from picha_b200.synthetic import fill_host
    src = fill_host(50, 50, 3, 152, 1235, 0)
    dst = np.zeros(300 * 100, np.uint8)
    s = N.CImage(src.ctypes.data, 152, 50, 50, 0)
    d = N.CImage(dst.ctypes.data, 300, 100, 100, 0)
    tag, width = ctypes.c_int(0), ctypes.c_float(0)
    N.check(N.lib.picha_b200_resolve_resize_options(0, 0, 0, 0.0, ctypes.byref(tag), ctypes.byref(width)))
    times = []
    for i in range(220):
        t0 = time.perf_counter()
        N.check(N.lib.picha_b200_resize(ctypes.byref(s), ctypes.byref(d), tag.value, width.value))
        times.append(time.perf_counter() - t0)
    times = sorted(times[20:])
    out = {"median_us": round(times[len(times) // 2] * 1e6, 1), "p90_us": round(times[int(len(times) * 0.9)] * 1e6, 1),
           "calls": len(times), "api": "picha_b200_resize (host buffers, blocking)"}
    if not args.no_cpu:
        import oracle as O
        impl = "ref" if O.have_ref() else "port"
        ct = []
        for i in range(60):
            t0 = time.perf_counter()
            want, ws = O.resize(src, 152, 50, 50, "rgb", 100, 100, "cubic", width.value, impl)
            ct.append(time.perf_counter() - t0)
        ct.sort()
        out["cpu_reference_median_us"] = round(ct[len(ct) // 2] * 1e6, 1)
        out["bit_exact_vs_reference"] = bool(np.array_equal(dst.reshape(100, 300), O.payload(want, ws, 100, 100, "rgb")))
    return out


# ---- the reference arm: the reference's own CPU implementation on the host cores ----------------------

def _fill_host_fn():
    """picha_b200/synthetic.py loaded by path: the reference arm must not import the picha_b200 package
    (that would map libpicha_b200.so into the reference's process)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_picha_synthetic", os.path.join(ROOT, "picha_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.fill_host


def shared_config(key, world):
    """The `config` object both arms print (the driver compares them)."""
    w = RESIZE_WORKLOADS.get(key) or CONVERT_WORKLOADS[key]
    return {"workload": w["name"], "images_per_gpu": w["batch"],
            "parallelism": f"dp{world} (sharded by image, no collective)",
            "l2": "GPU arm: inputs larger than L2 (per-step input >= 2 GB vs 126 MB L2), no explicit flush"}


def run_reference_arm(args):
    import numpy as np
    import oracle as O
    fill_host = _fill_host_fn()

    key = args.workload
    kind = "reference" if O.have_ref() else "port"
    impl = "ref" if kind == "reference" else "port"
    threads = max(1, min(os.cpu_count() or 1, args.cpu_threads))
    if key in RESIZE_WORKLOADS:
        w = RESIZE_WORKLOADS[key]
        bpp = PIXEL_BYTES[w["pixel"]]
        distinct = 4
        bufs = [fill_host(w["sw"], w["sh"], bpp, w["sw"] * bpp, w["seed"], i) for i in range(distinct)]
        filt = w["filter"] or "cubic"
        n_items = threads * args.ref_per_thread

        def one(i):
            O.resize(bufs[i % distinct], w["sw"] * bpp, w["sw"], w["sh"], w["pixel"], w["dw"], w["dh"], filt, w["width"], impl)

        mpix_item = w["dw"] * w["dh"] / 1e6
        name = w["name"]
    else:
        w = CONVERT_WORKLOADS[key]
        buf = fill_host(w["w"], w["h"], PIXEL_BYTES[w["src"]], w["w"] * PIXEL_BYTES[w["src"]], w["seed"], 0)
        n_items = threads * args.ref_per_thread * 4

        def one(i):
            O.color_convert(buf, w["w"] * PIXEL_BYTES[w["src"]], w["w"], w["h"], w["src"], w["dst"], None, impl)

        mpix_item = w["w"] * w["h"] / 1e6
        name = w["name"]
    from concurrent.futures import ThreadPoolExecutor
    items = list(range(n_items))
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(args.warmup):
            list(ex.map(one, items[:threads]))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(one, items))
        dt = time.perf_counter() - t0
    ms = dt * 1e3 / args.steps
    value = n_items * mpix_item / (ms / 1e3)
    sample = f"each step = {n_items} images of the workload on {threads} host threads, one image per thread at a time"
    line = {"impl": "reference", "metric": "output_mpix_per_s", "value": round(value, 2), "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(key, max(1, args.gpus)),
            "notes": {"filter": w.get("filter") or ("cubic" if key in RESIZE_WORKLOADS else None), "host_threads": threads,
                      "timing": "wall clock around each step on the host", "step": sample},
            "cpu_baseline": {"value": round(value, 2), "unit": "Mpix/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 2), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- main --------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(RESIZE_WORKLOADS) + list(CONVERT_WORKLOADS))
    ap.add_argument("--also", default="auto", help="comma list of extra workloads to report under 'also' "
                                                   "('auto' = the other BASELINE configs at N=1, 'none')")
    ap.add_argument("--e2e-batch", type=int, default=32, help="images per end-to-end step per GPU")
    ap.add_argument("--cpu-threads", type=int, default=64)
    ap.add_argument("--ref-per-thread", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args)
        return

    if world == 1 and args.gpus > 1:
        # not under torchrun: relaunch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
        sys.exit(subprocess.call(cmd + sys.argv[1:]))

    import torch
    import picha_b200 as P

    if not torch.cuda.is_available() or P.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: picha_b200 has no CPU fallback "
                         "(use --impl reference for the host-core baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    key = args.workload
    runner = run_resize_workload if key in RESIZE_WORKLOADS else run_convert_workload
    kw = {"extras": not args.no_e2e} if key in RESIZE_WORKLOADS else {}
    res = runner(key, args, rank, world, local_rank, with_e2e=not args.no_e2e, with_cpu=not args.no_cpu, **kw)

    also = {}
    extra = []
    if args.also == "auto":
        extra = [k for k in ("cfg2", "cfg2-grey", "cfg2-greya", "cfg4", "cfg5", "cfg5-grey") if k != key] if world == 1 else []
    elif args.also != "none":
        extra = [k for k in args.also.split(",") if k]
    for k in extra:
        e2e_too = k == "cfg5" and not args.no_e2e        # the thumbnail pipeline is a host-buffer workload by definition
        kw2 = {"with_latency": k == "cfg4" and not args.no_e2e} if k in RESIZE_WORKLOADS else {}
        r = (run_resize_workload if k in RESIZE_WORKLOADS else run_convert_workload)(
            k, args, rank, world, local_rank, with_e2e=e2e_too, with_cpu=not args.no_cpu, **kw2)
        also[k] = {"value": round(r["value"], 1), "unit": "Mpix/s", "ms_per_step": round(r["ms_per_step"], 4),
                   "roofline_frac": r["roofline"]["frac"], "achieved_GBs": r["roofline"]["achieved"],
                   "images_per_gpu": r["images_per_gpu"], "cpu_baseline": r.get("cpu_baseline")}
        for extra_key in ("e2e", "call_latency"):
            if r.get(extra_key):
                also[k][extra_key] = r[extra_key]

    if rank == 0 and world == 1 and args.also == "auto":
        also["cfg1"] = cfg1_latency(args)

    if rank == 0:
        w = RESIZE_WORKLOADS.get(key) or CONVERT_WORKLOADS[key]
        line = {
            "metric": "output_mpix_per_s", "value": round(res["value"], 1), "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(res["ms_per_step"], 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(key, world),
            "notes": {"timing": "CUDA events on the launching stream, max over ranks"},
            "e2e": res.get("e2e"), "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
            "roofline": res["roofline"], "cpu_baseline": res.get("cpu_baseline"),
        }
        for extra_key in ("e2e_pageable", "call_latency", "e2e_sharder"):
            if res.get(extra_key) is not None:
                line[extra_key] = res[extra_key]
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
