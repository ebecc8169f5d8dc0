#!/bin/bash
set -u
TAG=${TAG:-r02k}
OUT=gpurun_out; mkdir -p $OUT
Q="--steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); b=d["breakdown_ms_per_step"]; z=d["zone_stats"]
    print("%-28s ms/cycle %.2f c2p %.2f flux %.2f (recon %.2f riemann %.2f) src %.2f upd %.2f ct %.2f halo %.2f bcs %.2f ff %.2f" % (sys.argv[1].split("/")[-1], d["ms_per_step"], b.get("c2p",0), b.get("flux",0), b.get("recon",0), b.get("riemann",0), b.get("source",0), b.get("update",0), b.get("ct_fix",0), b.get("halo",0), b.get("bcs",0), b.get("fixup_floors",0)+b.get("stale_fixup_floors",0)))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $OUT/${TAG}_v_$name.json 2> $OUT/${TAG}_v_$name.err; grep "L2 fetch" $OUT/${TAG}_v_$name.err | tee -a $OUT/${TAG}_variants.txt; summ $OUT/${TAG}_v_$name.json | tee -a $OUT/${TAG}_variants.txt; }
run plan_l2_32 PB_L2_FETCH=32
run plan_l2_64 PB_L2_FETCH=64
run plan_l2_128 PB_L2_FETCH=128
run noplan_l2_32 PB_L2_FETCH=32 PB_GHOST_PLAN=0
run noplan_l2_128 PB_L2_FETCH=128 PB_GHOST_PLAN=0
ls -la $OUT | tail -3
