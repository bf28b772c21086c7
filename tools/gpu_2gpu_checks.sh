# 2-GPU validation + measurements (run on the GPU box from the repo root: gpurun --gpus 2 -- 'bash tools/gpu_2gpu_checks.sh')
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py 2>&1 | grep identical
timeout 400 $TR --master-port 29512 bench.py --gpus 2 2>/dev/null | tail -1 > gpurun_out/bench_r1_n2.json
timeout 400 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1_reference_n2.json
for ex in peer nccl; do
  timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex --workload c5_50mp_rowshard_8step_10bit_vh 2>/dev/null | tail -1 > gpurun_out/bench_r1_c5_n2_$ex.json
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --workload c5_50mp_rowshard_8step_10bit_vh 2>/dev/null | tail -1 > gpurun_out/bench_r1_c5_n1.json
for f in n2 reference_n2 c5_n2_peer c5_n2_nccl c5_n1; do python - $f <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_r1_%s.json" % sys.argv[1]).read())
print(sys.argv[1], "value %.0f %s, ms/step %.3f, frac %s, clocks %s" % (d["value"], d["unit"], d["ms_per_step"], (d.get("roofline") or {}).get("frac"), d.get("clocks")))
PY
done
