# one ncu --set full capture of the fused kernel
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused7 -s 3 -c 1 -o gpurun_out/fused7_c python bench.py --steps 1 --warmup 3 --batch 2 --ring 2 --no-e2e --no-cpu-baseline > gpurun_out/fused7_c.log 2>&1
tail -c 150 gpurun_out/fused7_c.log
