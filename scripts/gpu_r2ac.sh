#!/bin/bash
set -u
TAG=${TAG:-r02ac}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -q -x -k "stages_match or counts_match or con2prim_variants or ghost_plan or hashed or graph or split_post or benchmark_size" 2>&1 | tail -6 | tee $OUT/${TAG}_pytest_subset.txt
Q="--steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); b=d["breakdown_ms_per_step"]; z=d["zone_stats"]
    print("%-28s ms/cycle %.2f c2p %.2f flux %.2f (recon %.2f riemann %.2f) src %.2f upd %.2f ct %.2f halo %.2f bcs %.2f ff %.2f plan %s" % (sys.argv[1].split("/")[-1], d["ms_per_step"], b.get("c2p",0), b.get("flux",0), b.get("recon",0), b.get("riemann",0), b.get("source",0), b.get("update",0), b.get("ct_fix",0), b.get("halo",0), b.get("bcs",0), b.get("fixup_floors",0)+b.get("stale_fixup_floors",0), d["config"].get("ghost_plan")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $OUT/${TAG}_v_$name.json 2> $OUT/${TAG}_v_$name.err; summ $OUT/${TAG}_v_$name.json | tee -a $OUT/${TAG}_variants.txt; }
run twins A=1
run twins_two_kernels PHOEBUS_B200_LIB=$PWD/variants/libpb_twin_colmajor.so
