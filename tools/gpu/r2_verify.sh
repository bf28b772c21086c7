# round-2 verification: what the driver runs at round end (tests, smoke, bench both arms)
D=gpurun_out/verify; mkdir -p $D
timeout 900 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $D/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > $D/bench_ref.json 2> $D/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/verify/bench.json"))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","vs_baseline","dtype","scaling")}, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["clocks"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["cpu_baseline"].get("single_thread_value"))
r=json.load(open("gpurun_out/verify/bench_ref.json")); print("ref", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
