#!/usr/bin/env python
"""Three 56-frame remaps of a 12 MP stack, for an ncu capture of k_remap_tiled (tools/gpu/r2_ncu_remap.sh)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
s3 = importlib.import_module("3dscan_b200")
from helpers import load_calib_c1, scaled_calib   # noqa: E402

W, H, NF = 4096, 3000, 56
c = scaled_calib(load_calib_c1(), W / 1600.0, W / 1280.0)
cal = s3.make_calib(*[c[k] for k in ("Kc", "dc", "Kp", "dp", "rc", "tc", "rp", "tp")])
ctx = s3.Scan3D(s3.make_config(W, H, W, H, 8, 10, 10, 4, 4, 2), 0, cal)
plane = W * H
src, _ = s3.peer_alloc(0, NF * plane)
dst, _ = s3.peer_alloc(0, NF * plane)
ctx.generate_patterns_dev(0, src)
ctx.generate_patterns_dev(1, src + 28 * plane)
for _ in range(3):
    ctx.undistort_frames_dev(src, NF, dst, 0)
ctx.sync()
print("ok")
