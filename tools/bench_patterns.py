#!/usr/bin/env python
"""Device pattern generator (scan3d_generate_patterns_dev): write rate against the HBM roofline.
One JSON line per configuration.  python tools/bench_patterns.py"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
s3 = importlib.import_module("3dscan_b200")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6456.2
for name, (PW, PH, N, M, fw) in {"c3 4096x3000 8-step 10-bit": (4096, 3000, 8, 10, 4), "c5 8192x6144 8-step 10-bit": (8192, 6144, 8, 10, 8),
                                 "c1 1280x720 3-step 6-bit": (1280, 720, 3, 6, 32)}.items():
    cfg = s3.make_config(64, 16, PW, PH, N, M, M, fw, fw, 2)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx = s3.Scan3D(cfg, 0, None, stream=st.cuda_stream)
    nbytes = (N + 2 * M) * PW * PH
    out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for d in (0, 1):
        for _ in range(3):
            ctx.generate_patterns_dev(d, out.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        reps = 10
        for _ in range(reps):
            ctx.generate_patterns_dev(d, out.data_ptr())
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"workload": name, "direction": "vertical" if d == 0 else "horizontal", "patterns": N + 2 * M,
                          "bytes": nbytes, "ms": ms, "GB/s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                          "note": "profiles cached in the context after the first call; pure write stream"}))
    ctx.close()
