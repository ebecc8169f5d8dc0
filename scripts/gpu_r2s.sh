#!/bin/bash
# final single-GPU evidence of round 2: tests, the default bench line (+ reference arm), strong-scaling base lines,
# launch list, ncu --set full of one stage
set -u
TAG=${TAG:-r02s}
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 600 python bench.py --strong --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_strong_1gpu.json 2> $OUT/${TAG}_strong.err
timeout 600 python bench.py --strong --graphs --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_strong_graphs_1gpu.json 2> $OUT/${TAG}_strong_graphs.err
timeout 600 python bench.py --graphs --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_weak_graphs_1gpu.json 2> $OUT/${TAG}_weak_graphs.err
for f in strong strong_graphs weak_graphs; do python - $OUT/${TAG}_bench_${f}_1gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print("%-44s ms/cycle %.3f zc/s %.4e" % (sys.argv[1].split("/")[-1], d["ms_per_step"], d["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done | tee $OUT/${TAG}_summary_1gpu.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
python scripts/summarize_profile.py launches $OUT/${TAG}_launches.csv $OUT/${TAG}_launches_summary.md > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:recon_kernel|riemann_kernel|c2p_park_kernel|c2p_finish_kernel|update_stream_kernel|source_kernel|halo_local_kernel|ct_flux_kernel|emf_kernel|bc_kernel' \
    -s 80 -c 16 -f -o /tmp/prof_$TAG python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_prof.log 2>&1
python scripts/summarize_profile.py full /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.md > /dev/null
python scripts/summarize_profile.py json /tmp/prof_$TAG.ncu-rep $OUT/${TAG}_ncu_full.json > /dev/null
ls -la $OUT | tail -8
