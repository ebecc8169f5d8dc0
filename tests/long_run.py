"""Shared logic of the long-run parity tests: evolve a golden fixture's init state for one full wave
period with the reference's time-step control, compare with the reference's sparse dumps bit for bit,
and compare the L1 error against the analytic solution (after one period the exact solution is the
initial state) with the reference's own L1 (BASELINE.json north_star: within 1 %).

Time-step control: the init dump carries the first (dt_init_fact-scaled)
dt; afterwards dt = min(2 dt_used, cfl * EstimateTimestep), clipped to tlim
(ref: external/parthenon/src/driver/driver.cpp:216-240, src/phoebus_driver.cpp:121-132)."""
import numpy as np

from tests.helpers import stage_tags

LONG_CASES = ["linmodes2d_long", "linmodes3d_long"]


def evolve(g, stage_fn, estimate_fn, check_fn):
    """stage_fn(stage, dt); estimate_fn() -> cfl-scaled estimate; check_fn(tag) called after a dumped stage"""
    tags = stage_tags(g)
    dumped = {(g.manifest(t)["cycle"], g.manifest(t)["stage"]): t for t in tags}
    last = g.manifest(tags[-1])
    tlim = last["time"] + last["dt"]
    ncycles = last["cycle"] + 1
    dt = g.manifest("init")["dt"]
    time = g.manifest("init")["time"]
    for cycle in range(ncycles):
        if (cycle, 1) in dumped:  # the reference's dt of this cycle must be what our control produced
            assert g.manifest(dumped[(cycle, 1)])["dt"] == dt, (cycle, dt)
        for stage in (1, 2):
            stage_fn(stage, dt)
            if (cycle, stage) in dumped:
                check_fn(dumped[(cycle, stage)])
        time += dt
        new_dt = min(2.0 * dt, estimate_fn())
        if time < tlim and tlim - time < new_dt:
            new_dt = tlim - time
        dt = new_dt
    return ncycles


def l1_vs_initial(final, initial, interior):
    """mean |q(T) - q(0)| over interior zones, per primitive component"""
    d = np.abs(final - initial)[(slice(None),) + interior]
    return d.reshape(d.shape[0], -1).mean(axis=1)
