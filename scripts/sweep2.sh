#!/bin/bash
for r in 64 72 80 96 112 128; do
  for ft in 136 204 272 340 408; do
    PHOEBUS_B200_LIB=$PWD/phoebus_b200/libpb_r$r.so PB_FLUX_THREADS=$ft python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1
  done
done
