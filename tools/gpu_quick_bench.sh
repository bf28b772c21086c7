# quick parity + bench loop used while optimising the fused kernel
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PYFMT='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["config"]["workload"][:3], d["config"]["fused_cfg"], d["config"]["triangulation"][:9], "us/scan %.1f"%r["avg_launch_us"], "frac %.3f"%r["frac"], "scans/s %.0f"%d["scans_per_s"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
timeout 90 python bench.py --steps 5 --ring 2 --batch 8 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$PYFMT"
timeout 90 python bench.py --steps 5 --ring 2 --batch 8 --no-e2e --no-cpu-baseline --exact-triangulation 2>/dev/null | tail -1 | python -c "$PYFMT"
for wl in c2_1080p_3step_8bit_v c1_1600x1200_3step_6bit_vh; do
  timeout 90 python bench.py --workload $wl --steps 5 --ring 8 --batch 32 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$PYFMT"
done
