#!/bin/bash
# Round-2 evidence: ncu --set full of the pair-force kernel (with source), and a dram-bytes / duration pass over
# every kernel of a step, for $WORKLOADS.  Raw exports are written next to the reports.
mkdir -p gpurun_out
for wl in ${WORKLOADS:-c3-eater-1M}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_ -s 3 -c 1 -f \
      -o gpurun_out/r02_force_${wl} \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --workload $wl > gpurun_out/r02_force_${wl}.log 2>&1
  ncu -i gpurun_out/r02_force_${wl}.ncu-rep --page raw --csv > gpurun_out/r02_force_${wl}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_force_${wl}.ncu-rep --page source --csv > gpurun_out/r02_force_${wl}_source.csv 2>/dev/null
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 160 --csv \
      --log-file gpurun_out/r02_dram_${wl}.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-graphs --workload $wl > gpurun_out/r02_dram_${wl}.log 2>&1
  tail -1 gpurun_out/r02_dram_${wl}.log | cut -c1-200
done
ls -la gpurun_out | grep r02_
