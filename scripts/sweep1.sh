#!/bin/bash
# tuning sweep: register caps, FMA contraction, flux tile sizes (128^3 zones)
for lib in libphoebus_b200.so libpb_r128.so libpb_r96.so libpb_fma.so; do
  for ft in 136 204 272 340; do
    PHOEBUS_B200_LIB=$PWD/phoebus_b200/$lib PB_FLUX_THREADS=$ft python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1
  done
done
for zt in 64 256; do
  PB_ZONE_THREADS=$zt python scripts/task_bench.py 128 128 128 3 2>&1 | tail -1
done
