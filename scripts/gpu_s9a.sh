#!/bin/bash
# session-9 validation: GPU tests (incl. the C++ host layer), default bench line, per-task timing,
# source-level ncu capture of the flux and c2p kernels (128^3, one stage)
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/s9a_pytest.txt
python bench.py > $OUT/s9a_bench.json 2> $OUT/s9a_bench.err; tail -c 2500 $OUT/s9a_bench.json
python scripts/task_bench.py 256 256 256 3 2>&1 | tail -1 | tee $OUT/s9a_task256.txt
ncu --set full --clock-control none --import-source on -k 'regex:flux_kernel|c2p_kernel' -s 3 -c 4 -f -o $OUT/prof_s9a \
    python scripts/task_bench.py 128 128 128 1 > $OUT/prof_s9a.log 2>&1
ls -la $OUT
