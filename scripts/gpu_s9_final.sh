#!/bin/bash
# round-1 evidence run on the B200 box: GPU tests, default bench line, ncu launch list of the same
# command, one ncu --set full capture of the main kernels of one RK stage at 256^3 zones.
set -u
TAG=${TAG:-s9z}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 3500 $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 1200 $OUT/${TAG}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:_kernel' --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k 'regex:recon_kernel|riemann_kernel|c2p_kernel|update_stream_kernel|source_kernel|halo_local_kernel|emf_kernel|ct_flux_kernel' \
    -s 60 -c 14 -f -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/prof_$TAG.log 2>&1
ls -la $OUT | tail -12
