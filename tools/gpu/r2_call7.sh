# round-2 GPU call 7: v7 + LUT prefetch, two contexts; v8 shapes
D=gpurun_out/c7; mkdir -p $D
B="python bench.py --no-e2e --no-cpu-baseline --steps 20"
SCAN3D_FUSED_IMPL=7 timeout 120 $B > $D/v7_lut.json 2>/dev/null
SCAN3D_FUSED_IMPL=7 timeout 120 $B --contexts 2 > $D/v7_lut_ctx2.json 2>/dev/null
timeout 120 $B > $D/v8.json 2>/dev/null
timeout 120 $B --contexts 2 > $D/v8_ctx2.json 2>/dev/null
for v in v8w5 v8cold; do
  SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_$v timeout 120 $B > $D/$v.json 2>/dev/null
done
SCAN3D_FUSED_IMPL=7 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or c1_" 2>&1 | tail -2
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c7/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} launches/scan {d['roofline']['launches_per_scan']} pts {d['points_last_scan']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
