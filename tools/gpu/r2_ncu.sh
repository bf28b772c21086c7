D=gpurun_out/prof; mkdir -p $D
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --batch 2 --ring 2 --contexts 1 --cta-limit 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/r2_launches.csv $CMD > $D/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 4 -c 1 -o $D/r2_fused $CMD > $D/ncu.log 2>&1
