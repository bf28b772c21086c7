export MASTER_ADDR=127.0.0.1
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --ring 2 --batch 8 --no-cpu-baseline 2>&1 | tail -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --workload c5_50mp_rowshard_8step_10bit_vh 2>&1 | tail -1
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload c5_50mp_rowshard_8step_10bit_vh 2>&1 | tail -1
