#!/bin/bash
# Iteration loop on the GPU box: parity tests then a short bench of the default workload.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
for fk in ${FORCE_KERNELS:-0}; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --force-kernel $fk ${BENCH_ARGS} > gpurun_out/bench_fk$fk.log 2>&1
  tail -1 gpurun_out/bench_fk$fk.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}, d['roofline']['frac'], d['roofline']['pair_tests_per_s'], d['e2e']['value'], d['config']['mean_neighbours'])" || tail -5 gpurun_out/bench_fk$fk.log
done
