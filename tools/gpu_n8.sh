#!/bin/bash
# One 8-GPU data point per workload (the 1- and 2-GPU points come from tools/gpu_two.sh).
mkdir -p gpurun_out
NG=${NG:-8}
for wl in ${WORKLOADS:-c3-eater-1M c5-settings-2M}; do
  CF_SLAB_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus $NG --steps 10 --warmup 3 --workload $wl > gpurun_out/n${NG}_${wl}.json 2> gpurun_out/n${NG}_${wl}.err
  tail -1 gpurun_out/n${NG}_${wl}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['n_gpus'], d['value'], d['ms_per_step'], d.get('phases_ms'), d['e2e']['value'])" || tail -5 gpurun_out/n${NG}_${wl}.err
done
grep -h "slab build" gpurun_out/n${NG}_*.err | head -4
