#!/usr/bin/env python
"""Runs tests/test_gpu_compat.py::test_reference_main_program without pytest/torch start-up:
    gpurun --timeout 60 -- 'python tests/main_program_runner.py > gpurun_out/main_program.log 2>&1'
"""
import os
import pathlib
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import test_gpu_compat as t   # noqa: E402

t0 = time.time()
t.test_reference_main_program(pathlib.Path(tempfile.mkdtemp()))
print(f"PASS reference main program ({time.time() - t0:.1f} s)")
