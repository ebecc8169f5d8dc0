#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): NCCL brick parity test, weak and strong scaling lines at N GPUs.
set -u
N=${N:-2}
TAG=${TAG:-r02m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_gpus.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests/test_multirank.py tests/test_host_cpp.py -m gpu -q -x -rs -k "nccl or two_gpus" 2>&1 | tail -12 | tee $OUT/${TAG}_pytest_multirank_${N}gpu.txt
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711"
Q="--gpus $N --steps 8 --warmup 3 --no-cpu-baseline"
IFS=';' read -ra MODELIST <<< "${MODES:-weak:;weak_torch:--halo torch;strong:--strong;strong_graphs:--strong --graphs;weak_graphs:--graphs}"
for mode in "${MODELIST[@]}"; do
  name=${mode%%:*}; args=${mode#*:}
  timeout 900 $TR bench.py $Q $args > $OUT/${TAG}_${name}_${N}gpu.json 2> $OUT/${TAG}_${name}_${N}gpu.err
  python - "$OUT/${TAG}_${name}_${N}gpu.json" <<'PY' | tee -a $OUT/${TAG}_summary_${N}gpu.txt
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]); print("%-40s n=%d ms/cycle %.3f zc/s %.4e e2e %s launches %s" % (sys.argv[1].split("/")[-1], d["n_gpus"], d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"), d.get("gpu_launches")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
ls -la $OUT | tail -8
