#!/bin/bash
# Two-GPU check: full GPU test suite (multi-rank tests included), then 1- and 2-GPU benches.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1
tail -8 gpurun_out/pytest_gpu2.log
for wl in ${WORKLOADS:-c3-eater-1M c5-settings-2M}; do
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --workload $wl > gpurun_out/two_${wl}_1.json 2> gpurun_out/two_${wl}_1.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 10 --warmup 3 --workload $wl > gpurun_out/two_${wl}_2.json 2> gpurun_out/two_${wl}_2.err
  for n in 1 2; do
    tail -1 gpurun_out/two_${wl}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['n_gpus'], d['value'], d['ms_per_step'], d.get('phases_ms'), d['e2e']['value'])" || tail -5 gpurun_out/two_${wl}_$n.err
  done
done
