export MASTER_ADDR=127.0.0.1
for r in 16 8 32; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((r%10)) bench.py --gpus 2 --steps 20 --warmup 3 --reserve-sms $r --workload c5_50mp_rowshard_8step_10bit_vh 2>gpurun_out/c5n2.err | tail -1 > gpurun_out/bench_r1_c5_n2_r$r.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_c5_n2_r$r.json').read()); print('reserve', $r, 'scans/s', d['scans_per_s'], 'ms/scan', d['ms_per_step'], 'frac', d['roofline']['frac'], d['points_last_scan'])"
done
tail -3 gpurun_out/c5n2.err
