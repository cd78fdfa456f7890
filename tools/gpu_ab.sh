#!/bin/bash
# A/B on the GPU box: parity tests, then short benches of several workloads with several force kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-c3-eater-1M c3-pulser-1M c5-settings-2M}; do
  for fk in ${FORCE_KERNELS:-2 3}; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl --force-kernel $fk ${BENCH_ARGS} > gpurun_out/ab_${wl}_fk$fk.json 2> gpurun_out/ab_${wl}_fk$fk.err
    tail -1 gpurun_out/ab_${wl}_fk$fk.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl fk$fk', {k:d[k] for k in ('value','ms_per_step','phases_ms')}, d['roofline']['frac'], d['e2e']['value'], d['config']['mean_neighbours'])" || tail -5 gpurun_out/ab_${wl}_fk$fk.err
  done
done
