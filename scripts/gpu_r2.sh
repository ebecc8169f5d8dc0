#!/bin/bash
# Round-2 evidence on the B200 box: GPU tests, bench line, ncu launch list of the bench command. TAG names the outputs.
set -u
TAG=${TAG:-r02a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
fi
timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
K='regex:recon_kernel|riemann_|flux_kernel|c2p_|floors_|flags_scan|flag_|update_|source_kernel|divergence_kernel|emf_kernel|ct_flux|fix_fluxes|halo_|bc_kernel|p2c_kernel|dt_kernel|divb_kernel|weighted_sum|pgen_|scale_bfield'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file /tmp/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
python scripts/summarize_profile.py launches /tmp/launches_$TAG.csv $OUT/${TAG}_launches.md > /dev/null
fi
ls -la $OUT | tail -12
