#!/bin/bash
# validation + source-level profile of the two hottest kernels (128^3 zones, one stage)
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/r1b_pytest.txt
python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1 | tee $OUT/r1b_task128.txt
python scripts/task_bench.py 256 256 256 3 2>&1 | tail -1 | tee $OUT/r1b_task256.txt
ncu --set full --clock-control none --import-source on -k 'regex:flux_kernel|c2p_kernel|source_kernel|update_stream_kernel|halo_local_kernel' -s 22 -c 7 -f -o $OUT/prof_r1b \
    python scripts/task_bench.py 128 128 128 1 > $OUT/prof_r1b.log 2>&1
ls -la $OUT
