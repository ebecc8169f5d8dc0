#!/bin/bash
set -u
TAG=${TAG:-r02c}
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
for v in d0 m5 u5; do
  PHOEBUS_B200_LIB=$PWD/variants/libpb_$v.so timeout 600 python bench.py $Q > $OUT/${TAG}_v_$v.json 2> $OUT/${TAG}_v_$v.err; summ $OUT/${TAG}_v_$v.json | tee -a $OUT/${TAG}_variants.txt
done
timeout 600 python bench.py $Q --graphs > $OUT/${TAG}_v_graphs.json 2> $OUT/${TAG}_v_graphs.err; python -c "
import json; d=json.load(open('$OUT/${TAG}_v_graphs.json')); print('graphs ms/cycle', d['ms_per_step'])" | tee -a $OUT/${TAG}_variants.txt
timeout 600 python bench.py --strong $Q > $OUT/${TAG}_strong1.json 2> $OUT/${TAG}_strong1.err; timeout 600 python bench.py --strong $Q --graphs > $OUT/${TAG}_strong1g.json 2> $OUT/${TAG}_strong1g.err
python -c "
import json
for f in ('strong1','strong1g'):
    d=json.load(open('$OUT/${TAG}_'+f+'.json')); print(f, 'ms/cycle', d['ms_per_step'], 'zc/s %.3e' % d['value'])" | tee -a $OUT/${TAG}_variants.txt
ls -la $OUT | tail -5
