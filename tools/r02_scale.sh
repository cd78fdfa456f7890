#!/bin/bash
# bench.py at N = $NGS GPUs of one box (the driver's own launch line), headline + extras; prints the summary.
mkdir -p gpurun_out
TAG=${TAG:-scale}
for n in ${NGS:-1 2}; do
  out=gpurun_out/${TAG}_n$n.json
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > $out 2> ${out%.json}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) \
      bench.py --gpus $n --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > $out 2> ${out%.json}.err
  fi
  tail -1 $out | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d %s value %.1f M ms/step %.4f e2e %.1f M' % (d['n_gpus'], d['config']['workload'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6), d['phases_ms'], d.get('parity'))
for k,v in d.get('extra',{}).items(): print('   extra', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','phases_ms','error','owned_max_over_mean','cpu_reference_single_thread')})
" || tail -8 ${out%.json}.err
done
