#!/bin/bash
set -u
TAG=${TAG:-r02x}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "benchmark_size" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_subset.txt
