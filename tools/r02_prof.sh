#!/bin/bash
# ncu --set full with source of the pair-force kernel, for the workloads in $WORKLOADS (one GPU).
mkdir -p gpurun_out
TAG=${TAG:-r02}
for wl in ${WORKLOADS:-c3-eater-1M c3-pulser-1M}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_ -s 3 -c 1 -f \
      -o gpurun_out/force_${TAG}_${wl} \
      python bench.py --steps 2 --warmup 3 --no-cpu --workload $wl ${BENCH_ARGS} > gpurun_out/ncu_${TAG}_${wl}.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_${wl}.log
done
ls -la gpurun_out
