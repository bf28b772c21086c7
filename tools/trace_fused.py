#!/usr/bin/env python
"""Pipeline timeline of the fused kernel (diagnostics).

Needs a library built with the trace hooks:
    export SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_trace
    SCAN3D_BUILD_TRACE=1 python 3dscan_b200/build.py && SCAN3D_TRACE=1 python tools/trace_fused.py
"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("SCAN3D_TRACE", "1")
from helpers import load_calib_c1, scaled_calib
s3 = importlib.import_module("3dscan_b200")
W, H, PW, PH, N, Mv, Mh, fwv, fwh, dirs = 4096, 3000, 4096, 3000, 8, 10, 10, 4, 4, 2
_c = scaled_calib(load_calib_c1(), W / 1600.0, PW / 1280.0)
cal = s3.make_calib(*[_c[k] for k in ("Kc", "dc", "Kp", "dp", "rc", "tc", "rp", "tp")])
cfg = s3.make_config(W, H, PW, PH, N, Mv, Mh, fwv, fwh, dirs, flags=s3.FLAG_FAST_TRIANGULATION)
stack, roi = s3.synth_stack(cfg, cal)
ctx = s3.Scan3D(cfg, 0, cal)
for _ in range(3):
    ctx.reconstruct(stack, roi)
buf = np.zeros(1024 * 64 * 8, np.uint64)
ctx.L.scan3d_debug_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
ctx._ck(ctx.L.scan3d_debug_get_trace(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size))
t = buf.reshape(1024, 64, 8).astype(np.int64)
names = ["slot_free", "loads_issued", "data_ready", "int_done", "staged", "count_out", "prefix_out", "drained"]
# events: 0 input slot handed back, 1 next tile's bulk loads issued, 2 tile data landed (consumer saw it),
#         3 integer phase done, 4 points staged in shared memory (triangulation done), 5 tile's count published
#         (before its triangulation), 6 look-back resolved + inclusive prefix published, 7 points streamed out
for cta in (0, 5, 147, 200, 295):
    tt = t[cta]
    t0 = tt[0, 0]
    print("CTA", cta)
    for it in range(3, 10):
        if tt[it, 2] == 0 or tt[it, 7] == 0:
            break
        r = tt[it] - t0
        print("  it %2d" % it, " ".join("%s=%7d" % (n, v) for n, v in zip(names, r)))
# aggregate over CTAs / tiles (tiles 4.. of every CTA; the trace buffer is cleared before each launch)
cur, prev = t[:, 4:40], t[:, 3:39]
ok = (cur[:, :, 2] > 0) & (cur[:, :, 7] > 0) & (prev[:, :, 4] > 0)
def stat(x): x = x[ok]; return "mean %6.0f p50 %6.0f p90 %6.0f" % (x.mean(), np.median(x), np.percentile(x, 90))
print("tiles traced:", int(ok.sum()), " tiles per CTA: min %d max %d" % ((t[:, :, 2] > 0).sum(1)[(t[:, :, 2] > 0).sum(1) > 0].min(), (t[:, :, 2] > 0).sum(1).max()))
print("cycles: integer phase               ", stat(cur[:, :, 3] - cur[:, :, 2]))
print("cycles: FP64 decode + wait + triang.", stat(cur[:, :, 4] - cur[:, :, 3]))
print("cycles: consumer idle before data   ", stat(cur[:, :, 2] - prev[:, :, 4]))
print("cycles: load latency (issue->ready) ", stat(cur[:, :, 2] - cur[:, :, 1]))
print("cycles: count published -> prefix   ", stat(cur[:, :, 6] - cur[:, :, 5]))
print("cycles: staged -> drained           ", stat(cur[:, :, 7] - cur[:, :, 4]))
print("cycles: tile period                 ", stat(cur[:, :, 4] - prev[:, :, 4]))
