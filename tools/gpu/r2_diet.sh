D=gpurun_out/diet; mkdir -p $D
timeout 900 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $D/pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 20 --batch 16"
timeout 200 $B --contexts 1 > $D/short_1ctx.json 2>/dev/null
timeout 200 $B --contexts 3 > $D/short_3ctx.json 2>/dev/null
timeout 200 $B --contexts 3 --exact-triangulation > $D/short_3ctx_exact.json 2>/dev/null
timeout 300 python bench.py --no-e2e --no-cpu-baseline > $D/default.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/diet/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
