#!/usr/bin/env python3
"""Per-task device timing of the hot path on the synthetic torus (tuning tool, not the bench).
usage: python scripts/task_bench.py [N1 N2 N3] [reps]   (env: PHOEBUS_B200_LIB, PB_FLUX_THREADS, PB_ZONE_THREADS)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from phoebus_b200.driver import FluidDriver  # noqa: E402


def main():
    nx = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [128, 128, 128]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3

    class A:
        mesh, block = nx, 32
    dev = torch.device("cuda:0")
    mesh, fp, _ = bench.build_mesh(A, 0, 1, dev)
    drv = FluidDriver(mesh)
    bench.init_state(mesh, drv)
    dt = 0.1 * drv.estimate_dt()
    for _ in range(2):
        drv.step(dt)
    torch.cuda.synchronize()
    drv.enable_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        drv.step(dt)
    e1.record()
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1) / reps
    zones = nx[0] * nx[1] * nx[2]
    tag = "%s ft=%s zt=%s" % (os.path.basename(os.environ.get("PHOEBUS_B200_LIB", "default")), os.environ.get("PB_FLUX_THREADS", "-"),
                              os.environ.get("PB_ZONE_THREADS", "-"))
    parts = " ".join("%s=%.2f" % (k, v[1] / reps) for k, v in sorted(drv.timer_summary().items()))
    print("%-40s cycle %.2f ms  %.3e zc/s | %s" % (tag, tot, zones / tot * 1e3, parts), flush=True)


if __name__ == "__main__":
    main()
