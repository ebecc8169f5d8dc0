"""TEST INFRASTRUCTURE ONLY — build an OracleMesh (oracle/pyoracle.py) from a reference dump
and compare states. Used by tests/ and by oracle/make_golden.py's self-check."""
import numpy as np

from .pyoracle import BC, NGDER, NGLOC, Block, OracleMesh


def geom_from_dump(get, gid, shape):
    """Assemble the oracle geometry table [4][21][gk][gj][gi] + [33][gk][gj][gi] from the
    reference's cached geometry fields (cached_system.hpp:42-52). Face tables carry one
    extra face in their own direction, which is never read on the hot path: dropped."""
    c_alpha = get(gid, "base", "g.c.alpha")
    gk, gj, gi = c_alpha.shape[-3:]
    tab = np.zeros((4 * NGLOC + NGDER, gk, gj, gi))
    for loc, nm in enumerate(("c", "f1", "f2", "f3")):
        o = loc * NGLOC

        def f(name):
            a = get(gid, "base", "g.%s.%s" % (nm, name))
            a = a.reshape((-1,) + a.shape[-3:])
            return a[:, :gk, :gj, :gi]
        tab[o + 0] = f("alpha")[0]
        tab[o + 1:o + 4] = f("bcon")
        tab[o + 4:o + 14] = f("gcov")
        tab[o + 14:o + 20] = f("gamcon")
        tab[o + 20] = f("detgam")[0]
    o = 4 * NGLOC
    tab[o:o + 3] = get(gid, "base", "g.c.dalpha").reshape(3, gk, gj, gi)
    tab[o + 3:o + 33] = get(gid, "base", "g.c.dg").reshape(30, gk, gj, gi)
    return tab, (gk, gj, gi)


def block_bc(p, man_block):
    """per-face oracle bc codes from parthenon boundary flags (block=-1, periodic=1, user=2)."""
    out = []
    for f in range(6):
        fl = man_block["bflag"][f]
        if fl == -1:
            out.append(0)
        elif fl == 1:
            out.append(-1)
        else:
            out.append(BC[p["mesh_bc"][f]])
    return out


PRIM_FIELDS = ["p.density", "p.velocity", "p.energy", "p.bfield", "pressure", "gamma1"]
CONS_FIELDS = ["c.density", "c.momentum", "c.energy", "c.bfield"]


def stack(get, gid, cont, names, shape):
    parts = []
    for n in names:
        a = get(gid, cont, n)
        if a is None:
            continue
        parts.append(a.reshape((-1,) + shape))
    return np.concatenate(parts, axis=0)


def mesh_from_dump(p, man, get, geom_override=None):
    blocks = []
    nblk = [max(1, p["mesh_nx"][d] // p["nx"][d]) for d in range(3)]
    for mb in man["blocks"]:
        gid = mb["gid"]
        shape = tuple(get(gid, "base", "p.density").shape[-3:])
        if geom_override is not None:
            geom, gshape = geom_override(mb)
        else:
            geom, gshape = geom_from_dump(get, gid, shape)
        # logical position from coordinates (loc.lx* is per-tree in Parthenon's forest)
        mesh_min = [man["mesh"]["x%dmin" % (d + 1)] for d in range(3)]
        lx = []
        for d in range(3):
            g = p["ng"] if d < p["ndim"] else 0
            x0 = mb["xmin"][d] + g * mb["dx"][d]
            lx.append(int(round((x0 - mesh_min[d]) / (p["nx"][d] * mb["dx"][d]))))
        b = Block(p, mb["xmin"], mb["dx"], block_bc(p, mb), lx, geom, gshape)
        load_state(b, get, gid, shape)
        blocks.append(b)
    periodic = [p["mesh_bc"][2 * d] == "periodic" for d in range(3)]
    return OracleMesh(p, blocks, nblk, periodic)


def load_state(b, get, gid, shape):
    b.prim[:] = stack(get, gid, "base", PRIM_FIELDS, shape)
    b.temp[:] = get(gid, "base", "temperature").reshape((1,) + shape)
    b.c2p_mu[:] = get(gid, "base", "c2p_mu").reshape((1,) + shape)
    b.fail[:] = get(gid, "base", "fail").reshape((1,) + shape)
    b.csig[:] = get(gid, "base", "cell_signal_speed").reshape((-1,) + shape)
    b.fsig[:] = get(gid, "base", "face_signal_speed").reshape((-1,) + shape)
    for cont in ("base", "1"):
        if get(gid, cont, "c.density") is not None:
            b.cons[cont][:] = stack(get, gid, cont, CONS_FIELDS, shape)


def ref_state(get, gid, shape, sc1):
    """reference fields after a stage, in oracle layout"""
    out = dict(
        prim=stack(get, gid, "base", PRIM_FIELDS, shape),
        temp=get(gid, "base", "temperature").reshape((1,) + shape),
        c2p_mu=get(gid, "base", "c2p_mu").reshape((1,) + shape),
        fail=get(gid, "base", "fail").reshape((1,) + shape),
        csig=get(gid, "base", "cell_signal_speed").reshape((-1,) + shape),
        fsig=get(gid, "base", "face_signal_speed").reshape((-1,) + shape),
        cons=stack(get, gid, sc1, CONS_FIELDS, shape),
        dudt=stack(get, gid, "dUdt", CONS_FIELDS, shape),
    )
    fl = stack(get, gid, "base", ["bnd_flux::" + n for n in CONS_FIELDS], shape)
    out["flux_raw"] = fl
    return out


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(a), np.abs(b))
    scale = np.where(scale > 0, scale, 1.0)
    d = np.abs(a - b) / scale
    d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    return float(np.nanmax(d)) if d.size else 0.0
