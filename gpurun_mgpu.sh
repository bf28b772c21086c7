export MASTER_ADDR=127.0.0.1
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1_n2.json; tail -c 600 gpurun_out/bench_r1_n2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload c5_50mp_rowshard_8step_10bit_vh 2>/dev/null | tail -1 > gpurun_out/bench_r1_c5_n2.json; tail -c 700 gpurun_out/bench_r1_c5_n2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c5_50mp_rowshard_8step_10bit_vh 2>/dev/null | tail -1 > gpurun_out/bench_r1_c5_n1.json; tail -c 500 gpurun_out/bench_r1_c5_n1.json
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_r1_default.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_default.json').read()); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
