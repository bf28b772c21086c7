# round-2 multi-GPU call: N = number of visible GPUs
N=$(python -c "import torch; print(torch.cuda.device_count())")
D=gpurun_out/mg$N; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 tests/mgpu_check.py > $D/mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; grep -E "identical|MISMATCH|Error|error" $D/mgpu_check.log | head
timeout 400 $TR --master-port 29612 bench.py --gpus $N --workload c5_50mp_rowshard_8step_10bit_vh --no-e2e --no-cpu-baseline --steps 20 > $D/c5_peer.json 2> $D/c5_peer.err; echo "c5 peer rc=$?"; tail -2 $D/c5_peer.err
[ -n "$SKIP_NCCL" ] || timeout 400 $TR --master-port 29613 bench.py --gpus $N --workload c5_50mp_rowshard_8step_10bit_vh --exchange nccl --no-e2e --no-cpu-baseline --steps 20 > $D/c5_nccl.json 2> $D/c5_nccl.err; echo "c5 nccl rc=$?"; tail -2 $D/c5_nccl.err
[ -n "$SKIP_DEFAULT" ] || timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 5 > $D/default.json 2> $D/default.err; echo "default rc=$?"; tail -2 $D/default.err
python - <<PY
import json, os
def first_json(path):
    if not os.path.exists(path): return None
    for l in open(path):
        if l.startswith("{"): return json.loads(l)
for f in ("c5_peer", "c5_nccl"):
    d = first_json("$D/%s.json" % f)
    if d:
        r = d["roofline"]
        print(f, "ms/scan %.3f" % d["ms_per_step"], "Mpix/s %.0f" % d["value"], "frac n*hbm %.3f" % r["frac"], "floors hbm %.3f link %.3f" % (r["floor_ms_hbm"], r["floor_ms_nvlink_ingest"]), "of binding %.3f" % r["frac_of_binding_floor"], "ingest GB/s %.0f" % r["nvlink_ingest_gbs"])
d = first_json("$D/default.json")
if d: print("default value %.0f Mpix/s" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], "rowshard", d.get("rowshard"))
PY
