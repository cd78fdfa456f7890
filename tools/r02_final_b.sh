#!/bin/bash
# Round-end ncu evidence: --set full of the force kernel (eater, pulser) + per-kernel dram pass + launch lists
WORKLOADS="c3-eater-1M c3-pulser-1M" bash tools/r02_profile_final.sh 2>&1 | tail -12
TAG=fin WORKLOADS="c3-eater-1M c5-settings-2M c2-default-100k" COUNT=300 bash tools/r02_launches.sh 2>&1 | tail -60
