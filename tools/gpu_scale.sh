#!/bin/bash
# Multi-GPU validation on one box: multi-rank parity (dist_check) and the weak-scaling sweep.
mkdir -p gpurun_out
NG=${NG:-8}
python -m pytest tests/test_slab_gpu.py -m gpu -q -k multi_rank > gpurun_out/pytest_multi.log 2>&1
tail -3 gpurun_out/pytest_multi.log
for wl in ${WORKLOADS:-c3-eater-1M c5-settings-2M}; do
  for n in 1 2 4 8; do
    [ $n -gt $NG ] && continue
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --workload $wl > gpurun_out/scale_${wl}_$n.json 2> gpurun_out/scale_${wl}_$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) \
        bench.py --gpus $n --steps 10 --warmup 3 --workload $wl > gpurun_out/scale_${wl}_$n.json 2> gpurun_out/scale_${wl}_$n.err
    fi
    tail -1 gpurun_out/scale_${wl}_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['n_gpus'], d['value'], d['ms_per_step'], d.get('phases_ms'), d['e2e']['value'])" || tail -5 gpurun_out/scale_${wl}_$n.err
  done
done
