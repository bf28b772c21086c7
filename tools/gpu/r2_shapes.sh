D=gpurun_out/shapes; mkdir -p $D
B="python bench.py --no-e2e --no-cpu-baseline --steps 20 --batch 16"
for cfg in 7,3 5,4 4,4 9,2 6,2 7,2; do
SCAN3D_FUSED_CFG=$cfg timeout 200 $B --contexts 1 > $D/s_${cfg/,/x}_1ctx.json 2>/dev/null
SCAN3D_FUSED_CFG=$cfg timeout 200 $B --contexts 3 > $D/s_${cfg/,/x}_3ctx.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/shapes/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
