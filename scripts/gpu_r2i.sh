#!/bin/bash
set -u
TAG=${TAG:-r02i}
OUT=gpurun_out; mkdir -p $OUT
for v in pair zone hoist; do
  PB_HALO_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages_match or ghost_plan or con2prim_variants or graph" 2>&1 | tail -3 | tee -a $OUT/${TAG}_pytest_variants.txt
done
Q="--steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); b=d["breakdown_ms_per_step"]; z=d["zone_stats"]
    print("%-28s ms/cycle %.2f c2p %.2f flux %.2f (recon %.2f riemann %.2f) src %.2f upd %.2f halo %.2f bcs %.2f ff %.2f" % (sys.argv[1].split("/")[-1], d["ms_per_step"], b.get("c2p",0), b.get("flux",0), b.get("recon",0), b.get("riemann",0), b.get("source",0), b.get("update",0), b.get("halo",0), b.get("bcs",0), b.get("fixup_floors",0)+b.get("stale_fixup_floors",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $OUT/${TAG}_v_$name.json 2> $OUT/${TAG}_v_$name.err; summ $OUT/${TAG}_v_$name.json | tee -a $OUT/${TAG}_variants.txt; }
run pair PB_HALO_VARIANT=pair
run zone PB_HALO_VARIANT=zone
run hoist PB_HALO_VARIANT=hoist
run noplan_pair PB_GHOST_PLAN=0 PB_HALO_VARIANT=pair
run noplan_hoist PB_GHOST_PLAN=0 PB_HALO_VARIANT=hoist
PB_HALO_VARIANT=pair timeout 900 ncu --set full --clock-control none -k 'regex:halo_local_kernel' -s 2 -c 2 -f -o $OUT/${TAG}_halo_pair python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1
PB_HALO_VARIANT=zone timeout 900 ncu --set full --clock-control none -k 'regex:halo_local_kernel' -s 2 -c 2 -f -o $OUT/${TAG}_halo_zone python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline >> $OUT/${TAG}_prof.log 2>&1
ls -la $OUT | tail -8
