#!/bin/bash
# Multi-rank parity (tests/dist_check.py) at world size $NG; the log is what profiles/r02_dist_check_N.log holds.
mkdir -p gpurun_out
NG=${NG:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600+NG)) \
    tests/dist_check.py > gpurun_out/dist_check_$NG.log 2>&1
echo "rc=$?" >> gpurun_out/dist_check_$NG.log
grep -E "DIST_|rc=|Error|error|assert" gpurun_out/dist_check_$NG.log | tail -20
