#!/bin/bash
set -u
TAG=${TAG:-r02t}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather or stages_match" 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_subset.txt
