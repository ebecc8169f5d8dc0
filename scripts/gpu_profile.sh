#!/bin/bash
# Run on the B200 box (under gpurun): bench line, ncu launch list, one full ncu capture of the
# main kernels of one RK stage. Outputs land in gpurun_out/ (scratch); summaries are copied to
# profiles/ by scripts/summarize_profile.py.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
KERNELS='regex:flux_kernel|update_fused_kernel|c2p_kernel|floors_kernel'
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 3000 $OUT/bench_$TAG.json
# every launch of our kernels with its device time (torch's init kernels filtered out)
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:_kernel' --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
# one stage of the main kernels, full counters + source
ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 44 -c 7 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/prof_$TAG.log 2>&1
ls -la $OUT
