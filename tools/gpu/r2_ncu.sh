D=gpurun_out/ncu2; mkdir -p $D
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 4 -c 1 -o $D/fused python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --batch 2 --ring 2 --contexts 1 > $D/ncu.log 2>&1
ls -la $D
