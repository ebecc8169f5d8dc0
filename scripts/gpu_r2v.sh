#!/bin/bash
# compute-sanitizer over the round-2 kernels: ghost-plan stage (two-phase ghost copy, zone-list con2prim), pgen, gather/scatter
set -u
TAG=${TAG:-r02v}
OUT=gpurun_out; mkdir -p $OUT
K='(stages_match and (torus3d or torus2d_fail or reflect2d_prim) and (flagged or plan_unflagged) and not exhaustive) or gather_scatter or (torus_pgen and torus2d and not polar)'
for tool in memcheck initcheck racecheck; do
  echo "### $tool" | tee -a $OUT/${TAG}_sanitizer.txt
  timeout 1200 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" 2>&1 | grep -v "Host Frame\|^=========     at\|^=========$" | tail -8 | tee -a $OUT/${TAG}_sanitizer.txt
done
