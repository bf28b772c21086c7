# Measures the experimental kernel variants built by tools/build_variants.py (3dscan_b200/lib_var_*/) on one B200:
#   python tools/build_variants.py && gpurun --timeout 1500 -- bash tools/gpu_variants.sh
# For every variant: the GPU parity tests of the fused path against the oracle (a variant that is not bit-exact is
# out, whatever its speed), then the bench line (kernel time only) next to the default build's.
mkdir -p gpurun_out
timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/variant_default.json 2>/dev/null
for d in 3dscan_b200/lib_var_*/; do
  v=$(basename $d); v=${v#lib_var_}
  echo "== $v"
  case $v in remap*)
    SCAN3D_LIBDIR=$PWD/$d timeout 100 python tests/aux_check_runner.py 2>&1 | tail -1
    SCAN3D_LIBDIR=$PWD/$d timeout 100 python tools/bench_aux.py 2>/dev/null | grep remap_frames | tee gpurun_out/variant_$v.jsonl
    continue;;
  esac
  SCAN3D_LIBDIR=$PWD/$d timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "wrapped_phase or atan2 or fused or c1_crop" 2>&1 | tail -2
  SCAN3D_LIBDIR=$PWD/$d timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/variant_$v.json 2>/dev/null
  SCAN3D_LIBDIR=$PWD/$d timeout 200 python bench.py --exact-triangulation --no-e2e --no-cpu-baseline > gpurun_out/variant_${v}_exact.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/variant_*.json")):
    try:
        d = json.load(open(f))
        print(f"{f:45s} {d['roofline']['avg_launch_us']:8.1f} us/scan  frac {d['roofline']['frac']:.3f}  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
