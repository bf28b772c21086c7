export MASTER_ADDR=127.0.0.1
for ex in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex --workload c5_50mp_rowshard_8step_10bit_vh 2>gpurun_out/c5n2.err | tail -1 > gpurun_out/bench_r1_c5_n2_$ex.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_c5_n2_$ex.json').read()); print('$ex', 'scans/s', d['scans_per_s'], 'ms/scan', d['ms_per_step'], 'frac', d['roofline']['frac'], d['points_last_scan'], d['clocks'])" || tail -5 gpurun_out/c5n2.err
done
