D=gpurun_out/ctal; mkdir -p $D
B="python bench.py --no-e2e --no-cpu-baseline --steps 10"
for w in c3_12mp_8step_10bit_vh c1_1600x1200_3step_6bit_vh c2_1080p_3step_8bit_v; do
 bt=16; [ $w = c3_12mp_8step_10bit_vh ] || bt=256
 for cfg in "3 0" "3 1" "6 1" "4 1" "2 1" "6 2"; do set -- $cfg
  timeout 200 $B --workload $w --batch $bt --contexts $1 --cta-limit $2 > $D/${w%%_*}_c$1_l$2.json 2>/dev/null
 done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ctal/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:40s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
