timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
PYFMT='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["config"]["fused_cfg"], d["config"]["triangulation"][:9], "us/scan %.1f"%r["avg_launch_us"], "frac %.3f"%r["frac"], "scans/s %.0f"%d["scans_per_s"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
for mode in "" "--exact-triangulation"; do
for cfg in 4,3,2 6,2,2 4,2,3 8,1,3; do
  SCAN3D_FUSED_CFG=$cfg timeout 120 python bench.py --steps 5 --ring 2 --batch 8 --no-e2e --no-cpu-baseline $mode 2>&1 | tail -1 | python -c "$PYFMT"
done
done
SCAN3D_FUSED_CFG=6,2,2 timeout 300 python tools/trace_fused.py 2>&1 | tail -7
