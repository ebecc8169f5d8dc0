"""Input producers next to the hot path: the torus problem modifier and a
synthetic torus-like initial state generated directly in HBM for the benchmark configurations.

The synthetic state is NOT the Fishbone-Moncrief solution of the reference's torus pgen
(ref: src/pgen/torus.cpp:130-474); it is a smooth, magnetised, rotating density torus embedded in
a floor-dominated atmosphere on the same FMKS grid, so that the hot path sees the same mix of
work (smooth torus zones, floored funnel/atmosphere zones, horizon-crossing inner boundary).
"""
import math

import numpy as np


def torus_problem_modifier(pin):
    """phoebus::ProblemModifier for the torus: x1 range from r_outer and the horizon radius
    (ref: src/pgen/torus.cpp:476-498)."""
    router = pin.real("coordinates", "r_outer", 40.0)
    x1max = math.log(router)
    a = pin.real("geometry", "a", 0.0)
    Rh = 1.0 + math.sqrt(1.0 - a * a)
    xh = math.log(Rh)
    ninside = pin.integer("torus", "n_inside_horizon", 5)
    nx1 = pin.integer("parthenon/mesh", "nx1", 1)
    dx = (x1max - xh) / (nx1 - ninside)
    x1min = xh - (ninside + 0.5) * dx
    mesh = pin.blocks.setdefault("parthenon/mesh", {})
    mesh["x1min"] = repr(x1min)
    mesh["x1max"] = repr(x1max)
    return x1min, x1max


def synthetic_torus(mesh, seed_amp=0.02):
    """Fill mesh.prim (all zones, ghosts included) with a torus-like state, device-side."""
    t = mesh.torch
    p = mesh.p
    dev = mesh.device
    gam = p.gamma
    h = p.geom_par["h"]
    r0, w = p_get(p, "rmax", 12.0), 4.0
    kappa = p_get(p, "kappa", 0.01)
    for n in range(mesh.nblocks):
        x1, x2, x3 = (t.from_numpy(np.ascontiguousarray(a)).to(dev) for a in mesh.cell_centers(n))
        X1 = x1.view(1, 1, -1)
        X2 = x2.view(1, -1, 1)
        X3 = x3.view(-1, 1, 1)
        r = t.exp(X1)
        th = math.pi * X2 + ((1.0 - h) / 2.0) * t.sin(2.0 * math.pi * X2)
        s = t.sin(th).abs()
        tor = t.exp(-((r - r0) / w) ** 2) * s ** 8
        # compact support, as in the reference pgen: outside the torus surface the atmosphere is at rest and
        # unmagnetised (v = 0, B = 0 exactly; ref: src/pgen/torus.cpp:246-262 atmosphere branch, and
        # A_phi = max(rho/rho_max - 0.2, 0), torus.cpp:330-352) instead of Gaussian tails down to denormals
        tor_v = t.where(tor > 1.0e-3, tor, t.zeros_like(tor))
        tor_b = t.where(tor > 0.2, tor - 0.2, t.zeros_like(tor))
        wig = 1.0 + seed_amp * t.sin(3.0 * X3 + 5.0 * X1) * t.cos(7.0 * th)
        rho = 1.0e-5 * r ** -1.5 + tor * wig
        u = kappa * rho ** gam / (gam - 1.0) + 1.0e-7 * r ** -2.5
        prim = mesh.prim[n]
        shape = prim.shape[1:]
        prim[0] = rho.expand(shape)
        prim[1] = 0.0
        prim[2] = 0.0
        prim[3] = (0.5 * tor_v * r ** -1.5).expand(shape)      # rotation about the axis (x3 = phi)
        prim[4] = u.expand(shape)
        if p.mhd:
            b0 = 2.0e-3
            prim[5] = (b0 * tor_b * t.cos(th) / r).expand(shape)
            prim[6] = (b0 * tor_b * s / (r * r)).expand(shape)
            prim[7] = 0.0
        prim[mesh.nprim - 2] = ((gam - 1.0) * u).expand(shape)
        prim[mesh.nprim - 1] = gam


def p_get(p, key, default):
    return float(getattr(p, "torus", {}).get(key, default)) if hasattr(p, "torus") else default
