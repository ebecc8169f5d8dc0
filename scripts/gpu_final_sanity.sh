#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/final_sanity.txt
timeout 150 python -m pytest tests/test_host_cpp.py tests/test_gpu_parity.py -m gpu -q -x -k "host or (stages_match and torus3d)" 2>&1 | tail -2 | tee -a $OUT/final_sanity.txt
