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
