# round-2 GPU call 1: time the prepared variants, per-line ncu profile of the default fused kernel
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1/smi.txt
B="python bench.py --no-e2e --no-cpu-baseline --steps 20"
timeout 300 $B > gpurun_out/c1/default.json 2> gpurun_out/c1/default.err
for v in all pass4 io500 io1000; do
  SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_$v timeout 300 $B > gpurun_out/c1/$v.json 2>/dev/null
done
for v in remapu2 remapwin remaptile remaptile2; do
  SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_$v timeout 100 python tests/aux_check_runner.py 2>&1 | tail -2 > gpurun_out/c1/$v.check
  SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_$v timeout 100 python tools/bench_aux.py 2>/dev/null | grep remap_frames > gpurun_out/c1/$v.jsonl
done
timeout 100 python tools/bench_aux.py 2>/dev/null > gpurun_out/c1/aux_default.jsonl
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_all timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "wrapped_phase or atan2 or fused or c1_crop" 2>&1 | tail -3 > gpurun_out/c1/all.parity
# ncu: full set with source for the fused kernel (default build)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 6 -c 1 -o gpurun_out/c1/fused_default python bench.py --no-e2e --no-cpu-baseline --steps 1 --batch 2 > gpurun_out/c1/ncu.log 2>&1
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_var_all timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 6 -c 1 -o gpurun_out/c1/fused_all python bench.py --no-e2e --no-cpu-baseline --steps 1 --batch 2 > gpurun_out/c1/ncu_all.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c1/*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
cat gpurun_out/c1/*.jsonl gpurun_out/c1/*.check gpurun_out/c1/all.parity
