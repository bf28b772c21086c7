# Round-end artefacts on one B200 (from the repo root: gpurun -- bash tools/gpu_round_artifacts.sh); needs a trace-enabled build in 3dscan_b200/lib_trace for the last line
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; tail -c 400 gpurun_out/bench_r1_default.json; tail -3 gpurun_out/bench_r1_default.err
timeout 300 python bench.py --exact-triangulation --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_exact.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null
timeout 200 python bench.py --workload c2_1080p_3step_8bit_v --ring 8 --batch 32 --no-cpu-baseline > gpurun_out/bench_r1_c2.json 2>/dev/null
timeout 200 python bench.py --workload c1_1600x1200_3step_6bit_vh --ring 8 --batch 32 --no-cpu-baseline > gpurun_out/bench_r1_c1.json 2>/dev/null
SCAN3D_FUSED_IMPL=6 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_v6kernel.json 2>/dev/null
SCAN3D_FUSED_DYN=0 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_static_schedule.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --batch 2 --ring 2 --no-e2e --no-cpu-baseline > gpurun_out/launches_r1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 3 -c 1 -o gpurun_out/fused_r1_final python bench.py --steps 1 --warmup 3 --batch 2 --ring 2 --no-e2e --no-cpu-baseline > gpurun_out/fused_r1_final.log 2>&1
SCAN3D_LIBDIR=$PWD/3dscan_b200/lib_trace SCAN3D_TRACE=1 timeout 200 python tools/trace_fused.py > gpurun_out/trace_r1.txt 2>&1; tail -9 gpurun_out/trace_r1.txt
timeout 60 python tests/aux_check_runner.py > gpurun_out/aux_check.log 2>&1; tail -1 gpurun_out/aux_check.log
timeout 60 python tools/bench_aux.py > gpurun_out/bench_aux.jsonl 2> gpurun_out/bench_aux.err; cat gpurun_out/bench_aux.jsonl
ls -la gpurun_out | tail -5
