#!/bin/bash
# ncu capture of the pair-force kernel (one GPU): full set with source, plus the launch list.
mkdir -p gpurun_out
WL=${WORKLOAD:-c3-eater-1M}
ncu --set full --clock-control none --import-source on -k regex:force_ -s 3 -c 1 -f -o gpurun_out/force_${TAG:-prof} \
    python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL ${BENCH_ARGS} > gpurun_out/ncu_force.log 2>&1
tail -3 gpurun_out/ncu_force.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG:-prof}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --workload $WL ${BENCH_ARGS} > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
