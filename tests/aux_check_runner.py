#!/usr/bin/env python
"""Runs tests/test_gpu_zaux.py's checks without pytest/torch start-up (a few seconds on the GPU box):
    gpurun --timeout 120 -- 'python tests/aux_check_runner.py > gpurun_out/aux_check.log 2>&1'
"""
import os
import pathlib
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import test_gpu_zaux as t   # noqa: E402

CASES = [
    ("register", lambda: t.test_register_points_match_oracle_and_cv2()),
    ("roi 320x240", lambda: t.test_roi_fill_matches_reference_loop(320, 240)),
    ("roi 97x40", lambda: t.test_roi_fill_matches_reference_loop(97, 40)),
    ("roi 4096x16", lambda: t.test_roi_fill_matches_reference_loop(4096, 16)),
    ("undistort cv2 golden", lambda: t.test_undistort_frames_match_cv2_golden()),
    ("undistort 640x480", lambda: t.test_undistort_frames_match_oracle(640, 480, 1.0)),
    ("undistort 333x250", lambda: t.test_undistort_frames_match_oracle(333, 250, 3.0)),
    ("undistort 1000x37", lambda: t.test_undistort_frames_match_oracle(1000, 37, 6.0)),
    ("undistort 4096x64", lambda: t.test_undistort_frames_match_oracle(4096, 64, 1.0)),
    ("undistort args", lambda: t.test_undistort_argument_checks()),
    ("compat register/scissor", lambda: t.test_register_point_clouds_call(pathlib.Path(tempfile.mkdtemp()))),
    ("undistort 1600x1200", lambda: t.test_undistort_frames_match_oracle(1600, 1200, 1.0)),
    ("roi 1600x1200", lambda: t.test_roi_fill_matches_reference_loop(1600, 1200)),
]

bad = 0
for name, fn in CASES:
    t0 = time.time()
    try:
        fn()
        print(f"PASS {name} ({time.time() - t0:.2f} s)", flush=True)
    except Exception:
        bad += 1
        print(f"FAIL {name}", flush=True)
        traceback.print_exc()
        sys.stdout.flush()
print("aux check:", "all passed" if bad == 0 else f"{bad} failed", flush=True)
sys.exit(1 if bad else 0)
