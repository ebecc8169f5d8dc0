#!/bin/bash
# ncu evidence for the bench command (kept under the 64 MiB gpurun_out limit): launch list of OUR kernels
# and one --set full capture of each main kernel of one RK stage at 256^3 zones.
set -u
TAG=${TAG:-s9p}
OUT=gpurun_out; mkdir -p $OUT
K='regex:recon_kernel|riemann_kernel|flux_kernel|c2p_kernel|c2p_fixup|floors_|flags_scan|update_|source_kernel|divergence_kernel|emf_kernel|ct_flux|fix_fluxes|halo_|bc_kernel|p2c_kernel|dt_kernel|weighted_sum'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
ncu --set full --clock-control none -k 'regex:recon_kernel|riemann_kernel|c2p_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|bc_kernel' \
    -s 66 -c 13 -f -o $OUT/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/prof_$TAG.log 2>&1
ls -la $OUT | tail -6; du -sh $OUT
