# remap kernel: parity (GPU aux tests, also under compute-sanitizer memcheck), throughput, launch time
D=gpurun_out/remap; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_zaux.py -x -q -m gpu > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $D/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zaux.py -x -q -m gpu -k "undistort" > $D/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $D/memcheck.log | tail -3
timeout 300 python tools/bench_aux.py > $D/bench_aux.jsonl 2> $D/bench_aux.err; echo "bench_aux rc=$?"; grep remap $D/bench_aux.jsonl
