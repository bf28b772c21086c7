#!/usr/bin/env python
"""Pipeline timeline of the fused kernel (diagnostics).

Needs a library built with the trace hooks:
    SCAN3D_BUILD_TRACE=1 python 3dscan_b200/build.py --force && SCAN3D_TRACE=1 python tools/trace_fused.py
"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("SCAN3D_TRACE", "1")
from gpu_common import calibs, s3
W, H, PW, PH, N, Mv, Mh, fwv, fwh, dirs = 4096, 3000, 4096, 3000, 8, 10, 10, 4, 4, 2
cal, ocal, _ = calibs(W / 1600.0, PW / 1280.0)
cfg = s3.make_config(W, H, PW, PH, N, Mv, Mh, fwv, fwh, dirs, flags=s3.FLAG_FAST_TRIANGULATION)
stack, roi = s3.synth_stack(cfg, cal)
ctx = s3.Scan3D(cfg, 0, cal)
for _ in range(3):
    ctx.reconstruct(stack, roi)
buf = np.zeros(1024 * 64 * 8, np.uint64)
ctx.L.scan3d_debug_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
ctx._ck(ctx.L.scan3d_debug_get_trace(ctx.h, buf.ctypes.data_as(C.c_void_p), buf.size))
t = buf.reshape(1024, 64, 8).astype(np.int64)
names = ["slot_free", "loads_issued", "data_ready", "int_done", "dp_done", "epi_start", "lookback_done", "slot_released"]
for cta in (0, 5, 147, 200, 295):
    tt = t[cta]
    t0 = tt[0, 0]
    print("CTA", cta)
    for it in range(3, 12):
        if tt[it, 2] == 0:
            break
        r = tt[it] - t0
        print("  it %2d" % it, " ".join("%s=%7d" % (n, v) for n, v in zip(names, r)),
              "| wait_data=%d int=%d dp=%d epi_wait=%d lookback=%d scatter+drain=%d" % (
                  tt[it, 2] - max(tt[it - 1, 4], tt[it, 1]) if it else 0, tt[it, 3] - tt[it, 2], tt[it, 4] - tt[it, 3],
                  tt[it, 5] - tt[it, 4], tt[it, 6] - tt[it, 5], tt[it, 7] - tt[it, 6]))
# aggregate over CTAs / tiles
ok = t[:, 4:40, 2] > 0
def stat(x): x = x[ok]; return "mean %.0f p50 %.0f p90 %.0f" % (x.mean(), np.median(x), np.percentile(x, 90))
print("cycles: int phase", stat(t[:, 4:40, 3] - t[:, 4:40, 2]))
print("cycles: dp phase ", stat(t[:, 4:40, 4] - t[:, 4:40, 3]))
print("cycles: consumer idle before data", stat(t[:, 4:40, 2] - t[:, 3:39, 4]))
print("cycles: load latency (issue->ready)", stat(t[:, 4:40, 2] - t[:, 4:40, 1]))
print("cycles: epilogue lookback wait", stat(t[:, 4:40, 6] - t[:, 4:40, 5]))
print("cycles: epilogue scatter+drain", stat(t[:, 4:40, 7] - t[:, 4:40, 6]))
print("cycles: tile period", stat(t[:, 4:40, 4] - t[:, 3:39, 4]))
