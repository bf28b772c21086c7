D=gpurun_out/ab; mkdir -p $D
B="python bench.py --no-e2e --no-cpu-baseline --steps 20 --batch 16 --contexts 1"
for i in 1 2 3; do
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_noreg timeout 200 $B > $D/noreg_$i.json 2>/dev/null
timeout 200 $B > $D/cur_$i.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
