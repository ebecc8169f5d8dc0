#!/bin/bash
set -u
TAG=${TAG:-r02d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
Q="--steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); b=d["breakdown_ms_per_step"]; z=d["zone_stats"]
    print("%-28s ms/cycle %.2f c2p %.2f flux %.2f (recon %.2f riemann %.2f) src %.2f upd %.2f halo %.2f parked %s mean_it %.2f redo %d" % (sys.argv[1].split("/")[-1], d["ms_per_step"], b.get("c2p",0), b.get("flux",0), b.get("recon",0), b.get("riemann",0), b.get("source",0), b.get("update",0), b.get("halo",0), z.get("c2p_parked_by_warp_vote"), z.get("c2p_mean_iterations",0), z.get("arith_redo_items",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $OUT/${TAG}_v_$name.json 2> $OUT/${TAG}_v_$name.err; summ $OUT/${TAG}_v_$name.json | tee -a $OUT/${TAG}_variants.txt; }
run default A=1
run split PB_C2P_ARITH=split
run d0 PHOEBUS_B200_LIB=$PWD/variants/libpb_d0.so
run minb4 PHOEBUS_B200_LIB=$PWD/variants/libpb_u5.so
run minb6 PHOEBUS_B200_LIB=$PWD/variants/libpb_u6.so
run fused_tma PB_FLUX_MODE=fused PB_FLUX_STAGE=tma
run fused_cpasync PB_FLUX_MODE=fused PB_FLUX_STAGE=cpasync
run fma PHOEBUS_B200_LIB=$PWD/variants/libpb_fma.so
PHOEBUS_B200_LIB=$PWD/variants/libpb_fma.so timeout 600 python scripts/fma_price.py $OUT/${TAG}_fma_price.md > $OUT/${TAG}_fma_price.log 2>&1; tail -25 $OUT/${TAG}_fma_price.md
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1800 $OUT/${TAG}_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:recon_kernel|riemann_kernel|c2p_park_kernel|c2p_finish_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|emf_kernel|bc_kernel' \
    -s 70 -c 14 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.md > /dev/null
python scripts/summarize_profile.py json /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_kernel_metrics.json > /dev/null
PB_FLUX_MODE=fused timeout 600 ncu --set full --clock-control none -k 'regex:flux_kernel' -s 12 -c 3 -f -o /tmp/prof_fused_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof_fused.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_fused_$TAG.ncu-rep $OUT/${TAG}_ncu_fused_tma.md > /dev/null
for k in riemann_kernel c2p_park_kernel; do echo "### $k"; python scripts/sass_stalls.py /tmp/prof_$TAG.ncu-rep $k 0; echo; done > $OUT/${TAG}_sass_stalls.txt 2>&1
ls -la $OUT | tail -8
