# round-2 GPU call 8: v7 + warp-level skip of the FP64 phase, two contexts
D=gpurun_out/c8; mkdir -p $D
B="python bench.py --no-e2e --no-cpu-baseline --steps 20"
timeout 120 $B > $D/v7_skip.json 2>/dev/null
timeout 120 $B --contexts 2 > $D/v7_skip_ctx2.json 2>/dev/null
timeout 120 $B --contexts 3 > $D/v7_skip_ctx3.json 2>/dev/null
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_noskip timeout 120 $B > $D/v7_noskip.json 2>/dev/null
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_noskip timeout 120 $B --contexts 2 > $D/v7_noskip_ctx2.json 2>/dev/null
timeout 120 $B --contexts 2 --exact-triangulation > $D/v7_skip_ctx2_exact.json 2>/dev/null
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c8/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} launches/scan {d['roofline']['launches_per_scan']} pts {d['points_last_scan']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
