#!/bin/bash
# ncu launch list (device time of every launch; cold-cache, serialised: shares only) for $WORKLOADS.
mkdir -p gpurun_out
TAG=${TAG:-r02}
for wl in ${WORKLOADS:-c3-eater-1M}; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-400} --csv \
      --log-file gpurun_out/launches_${TAG}_${wl}.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-graphs --workload $wl ${BENCH_ARGS} > gpurun_out/launches_${TAG}_${wl}.log 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_${TAG}_${wl}.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
names=[]; agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[iv].replace(',','')); u=r[iu]
    v = v/1e3 if u in ('ns','nsecond') else (v*1e3 if u in ('ms','msecond') else v)
    k=r[ik].split('(')[0][:60]
    agg.setdefault(k,[]).append(v)
print('$wl')
for k,v in agg.items(): print(f'  {k:60s} n={len(v):3d} median {sorted(v)[len(v)//2]:9.1f} us  last {v[-1]:9.1f}')
PY
done
