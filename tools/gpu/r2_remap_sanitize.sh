# compute-sanitizer racecheck / synccheck over the remap kernel's GPU tests
D=gpurun_out/remap; mkdir -p $D
for tool in synccheck racecheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_zaux.py -x -q -m gpu -k "undistort" > $D/$tool.log 2>&1; echo "$tool rc=$?"; grep -E "SUMMARY|passed|failed" $D/$tool.log | tail -3
done
