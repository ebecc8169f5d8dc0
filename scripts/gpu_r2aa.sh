#!/bin/bash
# nghost = 2 + linear reconstruction (the reference's defaults): parity of every parametrised test on the two ng2 fixtures,
# then the same stage / flux tests under compute-sanitizer memcheck (the round-1 review's out-of-bounds stencil loads)
set -u
TAG=${TAG:-r02aa}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ng2" 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_ng2.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ng2 and (stages_match or fused or tasks_match)" 2>&1 | grep -v "Host Frame\|^=========     at\|^=========$" | tail -8 | tee $OUT/${TAG}_sanitizer_ng2.txt
