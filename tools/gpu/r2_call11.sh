# round-2 GPU call 11: modulation in the single-pass kernel, folded registration + raw entry, default bench
D=gpurun_out/c11; mkdir -p $D
timeout 900 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $D/pytest.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline > $D/c3_default.json 2>$D/c3_default.err
timeout 300 python bench.py --no-e2e --no-cpu-baseline --exact-triangulation > $D/c3_exact.json 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_remap_tiled -c 1 -o $D/remap python tools/bench_aux.py > $D/ncu_remap.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c11/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} ms/step {d['ms_per_step']:.1f}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
