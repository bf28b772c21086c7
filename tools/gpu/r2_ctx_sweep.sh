# contexts x CTA-limit sweep on the 12 MP workload, short region (no power cap) and the driver-style long region
D=gpurun_out/ctxs; rm -rf $D; mkdir -p $D
for cfg in "1 0" "3 1" "4 1" "5 1" "6 1" "3 0"; do set -- $cfg
  timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-rowshard --steps 10 --batch 24 --contexts $1 --cta-limit $2 > $D/short_c$1_l$2.json 2>/dev/null
done
for cfg in "3 1" "4 1" "6 1"; do set -- $cfg
  timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-rowshard --contexts $1 --cta-limit $2 > $D/long_c$1_l$2.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ctxs/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:40s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
