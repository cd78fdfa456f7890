#!/bin/bash
# ncu --set full of the warp-per-particle proximity-graph kernel at c2-default-100k (the reference's spawn cube)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_warp_kernel -s 3 -c 1 -f \
    -o gpurun_out/r02_graphw_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-graphs --workload c2-default-100k --graph-kernel 2 > gpurun_out/r02_graphw_c2.log 2>&1
ncu -i gpurun_out/r02_graphw_c2.ncu-rep --page raw --csv > gpurun_out/r02_graphw_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_graphw_c2.ncu-rep --page source --csv > gpurun_out/r02_graphw_c2_source.csv 2>/dev/null
ls -la gpurun_out | grep r02_graphw
