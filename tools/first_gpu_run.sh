#!/bin/bash
# First GPU round trip: golden vectors from the reference, pipe microbenchmarks, smoke, parity
# tests, a short bench.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
python tools/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
./tools/microbench/pipes > gpurun_out/pipes.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pipes.txt gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/golden.log
