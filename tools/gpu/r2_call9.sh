# round-2 GPU call 9: full parity suite (new tests), remap bench, small frames, sanitizer
D=gpurun_out/c9; mkdir -p $D
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -14 $D/pytest.log
timeout 120 python tools/bench_aux.py > $D/bench_aux.jsonl 2> $D/bench_aux.err; cat $D/bench_aux.jsonl | cut -c1-200
B="python bench.py --no-e2e --no-cpu-baseline --steps 5"
for w in c1_1600x1200_3step_6bit_vh c2_1080p_3step_8bit_v; do for c in 1 2 3; do
  timeout 120 $B --workload $w --contexts $c > $D/${w%%_*}_ctx$c.json 2>/dev/null
done; done
timeout 200 python bench.py --no-e2e --no-cpu-baseline > $D/c3_default.json 2>$D/c3_default.err
CS=/usr/local/cuda/bin/compute-sanitizer
T="tests/test_gpu_parity.py::test_fused_matches_oracle[case2] tests/test_gpu_parity.py::test_fused_matches_oracle[case5] tests/test_gpu_parity.py::test_fused_matches_oracle[case6] tests/test_gpu_parity.py::test_fused_roi_edge_cases"
for tool in memcheck racecheck synccheck; do
  timeout 400 $CS --tool $tool --print-limit 20 python -m pytest $T -x -q > $D/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 $D/sanitizer_$tool.log
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c9/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']} ms/step {d['ms_per_step']:.1f}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
