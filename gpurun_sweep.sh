timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PYFMT='import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]; print(d["config"]["workload"][:3], d["config"]["fused_cfg"], d["config"]["triangulation"][:9], "us/scan %.1f"%r["avg_launch_us"], "frac %.3f"%r["frac"], "scans/s %.0f"%d["scans_per_s"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
timeout 90 python bench.py --steps 5 --ring 2 --batch 8 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$PYFMT"
timeout 90 python bench.py --steps 5 --ring 2 --batch 8 --no-e2e --no-cpu-baseline --exact-triangulation 2>/dev/null | tail -1 | python -c "$PYFMT"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 1 --warmup 3 --batch 2 --ring 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_v7_c2.csv python bench.py --workload c2_1080p_3step_8bit_v --steps 1 --warmup 3 --batch 2 --ring 2 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/launches_v7.csv","gpurun_out/launches_v7_c2.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    tot={}
    for r in rows: tot.setdefault(r[4].split('(')[0],[]).append(float(r[-1]))
    print(f, {k:(len(v), round(sum(v)/len(v)/1e3,1)) for k,v in tot.items()})
PY
