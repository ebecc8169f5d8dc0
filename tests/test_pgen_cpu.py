"""CPU tests of the torus problem generator: the oracle's restatement of the reference's start-up sequence against the
reference's own "init" dumps (bit for bit), the product's host table against the oracle, the perturbation hook against
the harness's hash, and the oracle against the hashed 32^3-MeshBlock fixture."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import deck, from_dump as fd, pyoracle as po, refdump, torus_init as ti
from tests.helpers import GOLDEN, getter, load_case, oracle_mesh

FIELDS = ("prim", "temp", "c2p_mu", "fail", "csig", "cons")


def _init_ref(g, gid, shape):
    get = getter(g, "init")
    return dict(prim=fd.stack(get, gid, "base", fd.PRIM_FIELDS, shape), temp=get(gid, "base", "temperature").reshape((1,) + shape),
                c2p_mu=get(gid, "base", "c2p_mu").reshape((1,) + shape), fail=get(gid, "base", "fail").reshape((1,) + shape),
                csig=get(gid, "base", "cell_signal_speed").reshape((-1,) + shape),
                cons=fd.stack(get, gid, "base", fd.CONS_FIELDS, shape))


def _oracle_init(g, p):
    om = oracle_mesh(g, p)  # block coordinates, boundary codes and geometry tables as the reference had them
    man = g.manifest("init")
    par = po.geom_params(1, man["geom_params"])
    tpar = ti.torus_params(g.meta["deck_blocks"])
    amp = float(g.meta.get("env", {}).get("REF_DUMP_PERTURB", 0) or 0)
    mesh_xmin = [man["mesh"]["x%dmin" % (d + 1)] for d in range(3)]
    ti.oracle_torus_init(om, par, tpar, mesh_xmin, p["mesh_nx"][2], amp or None)
    return om, man


@pytest.mark.parametrize("name", ["torus2d", "torus3d", "torus2d_polar"])
def test_oracle_torus_init_matches_reference_init(name):
    """pgen -> PrimitiveToConserved -> ApplyFloors -> ghost exchange + fill -> FillDerived [-> harness perturbation] ->
    B normalisation, restated with the oracle, reproduces every field of the reference's state before its first cycle"""
    g, p = load_case(name)
    om, man = _oracle_init(g, p)
    for b, mb in zip(om.blocks, man["blocks"]):
        ref = _init_ref(g, mb["gid"], b.shape)
        for k in FIELDS:
            mine = b.cons["base"] if k == "cons" else getattr(b, k)
            assert np.array_equal(mine, ref[k], equal_nan=True), (name, mb["gid"], k, fd.relerr(mine, ref[k]))


def test_product_torus_table_matches_oracle_pgen():
    """pb_pgen_torus_table_host (phoebus_b200/csrc/geometry_host.cpp, written independently of oracle_geometry.cpp)
    against the oracle's zone-by-zone pgen: velocities everywhere, density / energy wherever the floors do not act"""
    from phoebus_b200 import pgen
    from tests.helpers import product_mesh
    g, p = load_case("torus2d")
    mesh, fp = product_mesh(g, alloc_fields=False)
    tables = pgen.torus_tables(mesh)
    om = oracle_mesh(g, p)
    man = g.manifest("init")
    par = po.geom_params(1, man["geom_params"])
    tpar = ti.torus_params(g.meta["deck_blocks"])
    tp = np.array([tpar["rin"], tpar["rmax"], tpar["kappa"], p["gamma"], tpar["nsub"], 0.0])
    from tests.helpers import block_lx
    nin = 0
    for b, mb in zip(om.blocks, man["blocks"]):
        prim, temp = np.zeros_like(b.prim), np.zeros_like(b.temp)
        assert po.lib().orc_pgen_torus(C.byref(b.cfg), par, tp, b.geom, prim, temp) == 0
        n = mesh.local_index[block_lx(mesh, man, mb, mesh.ndim, mesh.ng, mesh.nx)]
        tab = tables[mesh.geom_slot[n]]  # [5][1][nj][ni]
        for q, comp in ((2, 1), (3, 2), (4, 3)):
            assert np.array_equal(tab[q], prim[comp])
        inside = tab[0] > 0
        nin += int(inside.sum())
        unfloored = inside & (tab[0] == prim[0])
        assert unfloored.sum() > 0.9 * inside.sum()
        assert np.array_equal(tab[1][unfloored & (tab[1] == prim[4])], prim[4][unfloored & (tab[1] == prim[4])])
        assert (tab[0][~inside] == 0).all() and (tab[2:][:, ~inside] == 0).all()
    assert nin > 50


def test_hash_perturbation_matches_harness_hash():
    """phoebus_b200.pgen.hash_perturbation (torch, int64 wrap-around) == HashUnit of the reference harness"""
    import torch
    from phoebus_b200 import pgen

    class M:
        pass
    m = M()
    m.torch, m.device, m.ng, m.ndim, m.nblocks = torch, torch.device("cpu"), 4, 3, 1
    m.ni, m.nj, m.nk = 16, 12, 14
    m.bxmin = np.array([[0.5 - 4 * 0.125, -4 * 0.25, 2.0 - 4 * 0.5]])
    m.bdx = np.array([[0.125, 0.25, 0.5]])
    m.mesh_xmin = (-0.5, 0.0, 0.0)

    class P:
        mesh_nx = (64, 32, 12)
    m.p = P()
    m.prim = torch.ones((1, 10, m.nk, m.nj, m.ni), dtype=torch.float64)
    pgen.hash_perturbation(m, 0.05)(m)
    org = [8, 0, 4]
    f = ti.perturbation_factors((m.nk, m.nj, m.ni), 4, 3, org, 12, 0.05)
    assert np.array_equal(m.prim[0, 0].numpy(), f[0]) and np.array_equal(m.prim[0, 4].numpy(), f[1])
    assert np.array_equal(m.prim[0, 3].numpy(), f[2]) and (m.prim[0, 1] == 1).all()
    assert len(np.unique(f[0])) > 100


def test_oracle_matches_hashed_fixture_with_32cube_meshblocks():
    """torus3d_b32: 2x2x2 MeshBlocks of 32^3 (the production block shape), 2 cycles; the committed fixture holds one hash
    per array of the reference's dumps. Initial state from the oracle's pgen + the harness perturbation."""
    h = refdump.Hashed(os.path.join(GOLDEN, "torus3d_b32.json"))
    p = deck.params_from_deck(h.meta["deck_blocks"])
    man = h.manifest("init")
    par = po.geom_params(1, man["geom_params"])
    blocks = []
    nblk = [max(1, p["mesh_nx"][d] // p["nx"][d]) for d in range(3)]
    mesh_min = [man["mesh"]["x%dmin" % (d + 1)] for d in range(3)]
    for mb in man["blocks"]:
        lx = [int(round((mb["xmin"][d] + p["ng"] * mb["dx"][d] - mesh_min[d]) / (p["nx"][d] * mb["dx"][d]))) for d in range(3)]
        gshape = (1, p["nx"][1] + 2 * p["ng"], p["nx"][0] + 2 * p["ng"])  # geometry/axisymmetric = true
        cfg = po.make_cfg(p, mb["xmin"], mb["dx"], fd.block_bc(p, mb), gshape)
        geom = np.zeros((4 * po.NGLOC + po.NGDER,) + gshape)
        po.lib().orc_build_geometry(1, par, C.byref(cfg), geom)
        blocks.append(po.Block(p, mb["xmin"], mb["dx"], fd.block_bc(p, mb), lx, geom, gshape))
    om = po.OracleMesh(p, blocks, nblk, [p["mesh_bc"][2 * d] == "periodic" for d in range(3)])
    tpar = ti.torus_params(h.meta["deck_blocks"])
    ti.oracle_torus_init(om, par, tpar, mesh_min, p["mesh_nx"][2], float(h.meta["env"]["REF_DUMP_PERTURB"]))
    om.estimate_dt()  # the reference's EstimateTimestep before the first cycle (sets has_face_speeds)

    def check(tag, sc1):
        for b, mb in zip(om.blocks, h.manifest(tag)["blocks"]):
            gid = mb["gid"]
            names = {"p.density": b.prim[0:1], "p.velocity": b.prim[1:4], "p.energy": b.prim[4:5], "p.bfield": b.prim[5:8],
                     "pressure": b.prim[8:9], "gamma1": b.prim[9:10], "temperature": b.temp, "c2p_mu": b.c2p_mu, "fail": b.fail,
                     "cell_signal_speed": b.csig}
            for nm, arr in names.items():
                assert refdump.array_hash(arr) == h.hash(tag, gid, "base", nm), (tag, gid, nm)
            u = b.cons[sc1]
            for nm, arr in {"c.density": u[0:1], "c.momentum": u[1:4], "c.energy": u[4:5], "c.bfield": u[5:8]}.items():
                assert refdump.array_hash(arr) == h.hash(tag, gid, sc1, nm), (tag, gid, sc1, nm)
    check("init", "base")
    for tag in h.meta["tags"][1:]:
        m = h.manifest(tag)
        om.stage(m["stage"], m["dt"])
        check(tag, m["sc1"])
