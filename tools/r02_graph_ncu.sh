#!/bin/bash
# ncu --set full of the proximity-graph kernel (thread per particle, default for sparse lists) at c5-settings-2M
mkdir -p gpurun_out
for gk in ${GRAPH_KERNELS:-1}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_kernel -s 3 -c 1 -f \
      -o gpurun_out/r02_graph_gk${gk} \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-graphs --workload c5-settings-2M --graph-kernel $gk > gpurun_out/r02_graph_gk${gk}.log 2>&1
  ncu -i gpurun_out/r02_graph_gk${gk}.ncu-rep --page raw --csv > gpurun_out/r02_graph_gk${gk}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_graph_gk${gk}.ncu-rep --page source --csv > gpurun_out/r02_graph_gk${gk}_source.csv 2>/dev/null
done
ls -la gpurun_out | grep r02_graph
