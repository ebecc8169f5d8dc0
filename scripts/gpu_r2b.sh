#!/bin/bash
# Round-2 run B (1 GPU): GPU tests with the split con2prim, c2p tuning variants, bench, ncu --set full of the main kernels.
set -u
TAG=${TAG:-r02b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
Q="--steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); b=d["breakdown_ms_per_step"]; z=d["zone_stats"]
    print("%-28s ms/cycle %.2f c2p %.2f flux %.2f (recon %.2f riemann %.2f) parked %s mean_it %.2f redo %d" % (sys.argv[1].split("/")[-1], d["ms_per_step"], b.get("c2p",0), b.get("flux",0), b.get("recon",0), b.get("riemann",0), z.get("c2p_parked_by_warp_vote"), z.get("c2p_mean_iterations",0), z.get("arith_redo_items",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 600 python bench.py $Q > $OUT/${TAG}_v_default.json 2> $OUT/${TAG}_v_default.err; summ $OUT/${TAG}_v_default.json | tee -a $OUT/${TAG}_variants.txt
PB_C2P_ARITH=queue timeout 600 python bench.py $Q > $OUT/${TAG}_v_mono.json 2> $OUT/${TAG}_v_mono.err; summ $OUT/${TAG}_v_mono.json | tee -a $OUT/${TAG}_variants.txt
for v in d0 d4 d16 m5 m8; do
  PHOEBUS_B200_LIB=$PWD/variants/libpb_$v.so timeout 600 python bench.py $Q > $OUT/${TAG}_v_$v.json 2> $OUT/${TAG}_v_$v.err; summ $OUT/${TAG}_v_$v.json | tee -a $OUT/${TAG}_variants.txt
done
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:recon_kernel|riemann_kernel|c2p_solve_kernel|c2p_slow_kernel|c2p_unwrap_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|emf_kernel|bc_kernel' \
    -s 80 -c 16 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.md > /dev/null
python scripts/summarize_profile.py json /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_kernel_metrics.json > /dev/null
for k in riemann_kernel c2p_solve_kernel c2p_unwrap_kernel; do echo "### $k"; python scripts/sass_stalls.py /tmp/prof_$TAG.ncu-rep $k 0; echo; done > $OUT/${TAG}_sass_stalls.txt 2>&1
ls -la $OUT | tail -20
