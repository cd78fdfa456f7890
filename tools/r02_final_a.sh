#!/bin/bash
# Round-end evidence on one GPU: the GPU test suite, smoke(), the reference arm and the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final2_ref.json 2> gpurun_out/final2_ref.err; tail -1 gpurun_out/final2_ref.json | cut -c1-600
python bench.py --steps 20 --warmup 5 > gpurun_out/final2_n1.json 2> gpurun_out/final2_n1.err
tail -1 gpurun_out/final2_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d %s value %.1f M ms/step %.4f e2e %.1f M frac %.4f' % (d['n_gpus'], d['config']['workload'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac']), d['phases_ms'], d.get('parity'), d.get('clocks'))
print('   cpu', d.get('cpu_baseline'))
for k,v in d.get('extra',{}).items(): print('   extra', k, {a:b for a,b in v.items() if a in ('value','ms_per_step','phases_ms','error','cpu_reference_single_thread')})
" || tail -8 gpurun_out/final2_n1.err
