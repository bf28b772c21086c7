#!/usr/bin/env python
"""Throughput of the kernels either side of the path on the 12 MP configuration, no torch:
    gpurun --timeout 120 -- 'python tools/bench_aux.py > gpurun_out/r1_bench_aux.jsonl 2> gpurun_out/bench_aux.err'
Wall clock around K back-to-back launches between two stream synchronisations (each launch is 0.1-1 ms of GPU
work, so launch overhead hides behind the queue).  Bytes are algorithmic: every input once, every output once."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
s3 = importlib.import_module("3dscan_b200")
from helpers import load_calib_c1, scaled_calib   # noqa: E402  (calibration fixture only; nothing under oracle/)

W, H, NF, K = 4096, 3000, 56, 20
peak = 6456.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
c = scaled_calib(load_calib_c1(), W / 1600.0, W / 1280.0)
cal = s3.make_calib(*[c[k] for k in ("Kc", "dc", "Kp", "dp", "rc", "tc", "rp", "tp")])
ctx = s3.Scan3D(s3.make_config(W, H, W, H, 8, 10, 10, 4, 4, 2), 0, cal)
plane = W * H
src, _ = s3.peer_alloc(0, NF * plane)
dst, _ = s3.peer_alloc(0, NF * plane)
ctx.generate_patterns_dev(0, src)                    # 28 planes of real pattern bytes ...
ctx.generate_patterns_dev(1, src + 28 * plane)       # ... and 28 more
ctx.sync()


def timed(fn, k=K):
    fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / k


def line(name, seconds, nbytes, **kw):
    print(json.dumps(dict(kernel=name, us=round(seconds * 1e6, 1), algorithmic_bytes=nbytes,
                          achieved_gbs=round(nbytes / seconds / 1e9, 1), peak_gbs=peak,
                          frac=round(nbytes / seconds / 1e9 / peak, 3), **kw)), flush=True)


t0 = time.perf_counter()
ctx.undistort_frames_dev(src, 1, dst, 0)             # first call builds the map
ctx.sync()
t_first = time.perf_counter() - t0
t1 = timed(lambda: ctx.undistort_frames_dev(src, 1, dst, 0))
line("k_undistort_map (once per calibration, incl. first 1-frame remap)", t_first, plane * 6, workload="4096x3000")
line("k_remap_frames, 1 frame", t1, plane * (2 + 6), workload="4096x3000 x 1 frame")
tN = timed(lambda: ctx.undistort_frames_dev(src, NF, dst, 0))
line("k_remap_frames, 56-frame stack", tN, plane * (2 * NF + 6), workload="4096x3000 x 56 frames",
     mpix_per_s=round(plane / tN / 1e6, 1))
tr = timed(lambda: ctx._ck(ctx.L.scan3d_roi_fill_dev(ctx.h, src, dst, dst + plane)))
line("k_roi_fill", tr, plane * 3, workload="4096x3000 outline -> roi + filled")
n = plane
tp = timed(lambda: ctx.register_points_dev(src, dst, n, 36.0, 1.5, -2.0, 880.0))
line("k_register_points", tp, n * 24, workload="12.3 M points")
ctx.close()

# ---- raw captures -> points: scan3d_reconstruct_raw_dev (remap + worklist + single-pass kernel chained on one stream),
#      one context, then four contexts on four streams with one CTA slot per SM each for the persistent kernel: the
#      remap of one scan (shared-memory / LSU bound) then runs beside the decode of another (FP64 / issue bound)
import numpy as np   # noqa: E402
import torch         # noqa: E402

cfg = s3.make_config(W, H, W, H, 8, 10, 10, 4, 4, 2, flags=s3.FLAG_FAST_TRIANGULATION)
stack, roi = s3.synth_stack(cfg, cal)
d_stack = torch.from_numpy(stack).cuda()
d_roi = torch.from_numpy(roi).cuda()
del stack
for n_ctx, limit in ((1, 0), (4, 1), (4, 0)):
    streams = [torch.cuda.Stream() for _ in range(n_ctx)]
    ctxs = [s3.Scan3D(cfg, 0, cal, stream=st.cuda_stream) for st in streams]
    for c in ctxs:
        if limit:
            c.set_cta_limit(limit)
        c.reconstruct_raw_dev(d_stack.data_ptr(), d_roi.data_ptr())     # builds the map, allocates
    torch.cuda.synchronize()
    n_scans = 48
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for st in streams:
        st.wait_event(ev0)
    for k in range(n_scans):
        ctxs[k % n_ctx].reconstruct_raw_dev(d_stack.data_ptr(), d_roi.data_ptr())
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    ev1.record()
    torch.cuda.synchronize()
    sec = ev0.elapsed_time(ev1) * 1e-3 / n_scans
    line("raw captures -> points (remap + worklist + k_fused7), %d context(s)%s" % (n_ctx, ", 1 CTA slot per SM each" if limit else ""),
         sec, plane * (2 * NF + 6) + plane * 90, workload="4096x3000 x 56 raw frames per scan", mpix_per_s=round(plane / sec / 1e6, 1),
         points=int(ctxs[0].point_count()))
    for c in ctxs:
        c.close()
