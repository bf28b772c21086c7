# round-2 GPU call 10 (1 GPU): parity suite incl. single-rank row-shard schedule, remap, C5 on one GPU
D=gpurun_out/c10; mkdir -p $D
timeout 900 python -m pytest tests -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $D/pytest.log
timeout 120 python tools/bench_aux.py > $D/bench_aux.jsonl 2> $D/bench_aux.err; cut -c1-160 $D/bench_aux.jsonl
timeout 300 python bench.py --workload c5_50mp_rowshard_8step_10bit_vh --no-e2e --no-cpu-baseline --steps 20 > $D/c5_n1.json 2> $D/c5_n1.err; tail -2 $D/c5_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/c10/c5_n1.json")); print("c5 n1", d["ms_per_step"], d["roofline"])
PY
