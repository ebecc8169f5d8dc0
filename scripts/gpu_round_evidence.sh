#!/bin/bash
# Round evidence on the B200 box, everything summarised ON the box (the .ncu-rep files exceed the 64 MiB
# gpurun_out limit): GPU tests, bench line (ours + reference arm), ncu launch list of the bench command,
# one ncu --set full capture of each main kernel of one RK stage at 256^3 zones per GPU.
set -u
TAG=${TAG:-r01}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest.txt
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 400 $OUT/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
K='regex:recon_kernel|riemann_|flux_kernel|c2p_|floors_|flags_scan|update_|source_kernel|divergence_kernel|emf_kernel|ct_flux|fix_fluxes|halo_|bc_kernel|p2c_kernel|dt_kernel|divb_kernel|weighted_sum'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file /tmp/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
python scripts/summarize_profile.py launches /tmp/launches_$TAG.csv $OUT/${TAG}_launches.md > /dev/null
ncu --set full --clock-control none --import-source on -k 'regex:recon_kernel|riemann_kernel|c2p_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|bc_kernel' \
    -s 70 -c 14 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.md > /dev/null
for k in recon_kernel riemann_kernel c2p_kernel; do echo "### $k"; python scripts/sass_stalls.py /tmp/prof_$TAG.ncu-rep $k 0; echo; done > $OUT/${TAG}_sass_stalls.txt 2>&1
ls -la $OUT | tail -12; du -sh $OUT
