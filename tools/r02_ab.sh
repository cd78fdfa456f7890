#!/bin/bash
# A/B on one GPU box: (optionally) the GPU parity tests, then short benches of $WORKLOADS for every
# variant in $VARIANTS (each a quoted string of extra bench.py arguments; "-" = none).
mkdir -p gpurun_out
TAG=${TAG:-ab}
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_${TAG}.log 2>&1
  tail -6 gpurun_out/pytest_${TAG}.log
fi
IFS=';' read -ra VARS <<< "${VARIANTS:--}"
for wl in ${WORKLOADS:-c3-eater-1M c3-pulser-1M c5-settings-2M}; do
  vi=0
  for v in "${VARS[@]}"; do
    [ "$v" = "-" ] && v=""
    out=gpurun_out/${TAG}_${wl}_v${vi}.json
    lib=""
    case "$v" in LIB=*) lib="${v%% *}"; lib="${lib#LIB=}"; v="${v#LIB=* }"; [ "$v" = "LIB=$lib" ] && v="";; esac
    [ -n "$lib" ] && export CELLFLOW_B200_LIB=$PWD/cellflow_b200/lib/$lib || unset CELLFLOW_B200_LIB
    timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-extra --workload $wl $v > $out 2> ${out%.json}.err
    tail -1 $out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl [$v]', d['ms_per_step'], d['phases_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'nbrs', d['details']['mean_neighbours'], d.get('parity'), d['roofline'].get('block_counters'))" || tail -5 ${out%.json}.err
    vi=$((vi+1))
  done
done
