#!/bin/bash
# ncu evidence for the bench command, summarised ON the box (the .ncu-rep files exceed the 64 MiB
# gpurun_out limit): launch list of OUR kernels and one --set full capture of each main kernel of one
# RK stage at 256^3 zones per GPU.
set -u
TAG=${TAG:-s9p}
OUT=gpurun_out; mkdir -p $OUT
K='regex:recon_kernel|riemann_kernel|flux_kernel|c2p_kernel|c2p_fixup|floors_|flags_scan|update_|source_kernel|divergence_kernel|emf_kernel|ct_flux|fix_fluxes|halo_|bc_kernel|p2c_kernel|dt_kernel|weighted_sum'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
python scripts/summarize_profile.py launches $OUT/launches_$TAG.csv $OUT/${TAG}_launches.md > /dev/null
ncu --set full --clock-control none --import-source on -k 'regex:recon_kernel|riemann_kernel|c2p_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|bc_kernel' \
    -s 66 -c 13 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/prof_$TAG.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.md > /dev/null
for k in recon_kernel riemann_kernel c2p_kernel; do echo "### $k"; python scripts/sass_stalls.py /tmp/prof_$TAG.ncu-rep $k 0; echo; done > $OUT/${TAG}_sass_stalls.txt 2>&1
ls -la $OUT | tail -8; du -sh $OUT
