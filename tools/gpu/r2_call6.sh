# round-2 GPU call 6: warp-autonomous v8 after the ring-overflow fix -- parity + stress, bench, ncu (tight timeouts)
D=gpurun_out/c6; mkdir -p $D
timeout 500 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $D/pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 20"
timeout 120 $B > $D/v8.json 2> $D/v8.err; echo "bench rc=$?"
timeout 120 $B --exact-triangulation > $D/v8_exact.json 2>/dev/null
timeout 120 $B --workload c1_1600x1200_3step_6bit_vh > $D/v8_c1.json 2>/dev/null
timeout 120 $B --workload c2_1080p_3step_8bit_v > $D/v8_c2.json 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused8 -s 8 -c 1 -o $D/fused8 python bench.py --no-e2e --no-cpu-baseline --steps 1 --batch 4 > $D/ncu.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c6/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} launches/scan {d['roofline']['launches_per_scan']} pts {d['points_last_scan']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 $D/v8.err
