# round-2 artefacts for profiles/: launch list, one full ncu capture of the fused kernel, bench lines
D=gpurun_out/prof; mkdir -p $D
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --batch 2 --ring 2 --contexts 1 --cta-limit 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/r2_launches.csv $CMD > $D/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 4 -c 1 -o $D/r2_fused $CMD > $D/ncu.log 2>&1
timeout 300 python bench.py > $D/r2_bench_default.json 2> $D/r2_bench_default.err
timeout 300 python bench.py --exact-triangulation --no-cpu-baseline --no-e2e > $D/r2_bench_exact.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 > $D/r2_bench_reference.json 2>/dev/null
timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 --workload c1_1600x1200_3step_6bit_vh > $D/r2_bench_c1.json 2>/dev/null
timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 5 --workload c2_1080p_3step_8bit_v > $D/r2_bench_c2.json 2>/dev/null
timeout 120 python tools/bench_aux.py > $D/r2_bench_aux.jsonl 2>/dev/null
ls -la $D
