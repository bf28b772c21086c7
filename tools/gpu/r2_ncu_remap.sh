D=gpurun_out/prof; mkdir -p $D
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_remap_tiled -s 1 -c 1 -o $D/r2_remap python tools/gpu/remap_once.py > $D/ncu_remap.log 2>&1; tail -2 $D/ncu_remap.log
