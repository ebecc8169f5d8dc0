#!/bin/bash
# quick parity + per-task timing of the current library (and optional variants given as args)
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-q}
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.txt
python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1 | tee $OUT/${TAG}_task128.txt
python scripts/task_bench.py 256 256 256 3 2>&1 | tail -1 | tee $OUT/${TAG}_task256.txt
for lib in "$@"; do
  PHOEBUS_B200_LIB=$PWD/$lib python scripts/task_bench.py 256 256 256 3 2>&1 | tail -1 | tee -a $OUT/${TAG}_variants.txt
done
