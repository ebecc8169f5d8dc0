"""CPU-side checks of the product's host logic (no GPU needed): the C-ABI library loads and exports
every symbol include/phoebus_b200.h declares; deck parameters, block coordinates, boundary codes and
the host-built geometry tables agree bit-for-bit with the reference's own values recorded in the
golden fixtures; inter-GPU halo plans pair up."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import deck as odeck, from_dump as fd
from tests.helpers import CASES, ROOT, block_lx, getter, load_case, product_mesh, product_params


def test_library_exports_every_declared_symbol():
    from phoebus_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "phoebus_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().pb_version()


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors vs the C header, as the C compiler lays the structs out"""
    import subprocess
    from phoebus_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "phoebus_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(pb_block_t), sizeof(pb_cfg_t), sizeof(pb_pack_t),'
                   ' offsetof(pb_pack_t, recon_lr), offsetof(pb_cfg_t, ox1_inflow_check), offsetof(pb_block_t, nbr)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.PbBlock), C.sizeof(_lib.PbCfg), C.sizeof(_lib.PbPack), _lib.PbPack.recon_lr.offset,
            _lib.PbCfg.ox1_inflow_check.offset, _lib.PbBlock.nbr.offset]
    assert got == want, (got, want)


@pytest.mark.parametrize("name", CASES)
def test_deck_parameters_match(name):
    """FluidParams (product) vs the oracle's independent deck mapping"""
    g, p = load_case(name)
    fp = product_params(g)
    c = fp.make_cfg((1, 1, 1))
    for k in ("ndim", "ng", "mhd", "c2p_tol", "c2p_max_iter", "c2p_floor_scale_fac", "c2p_fail_on_floors",
              "c2p_fail_on_ceilings", "rho0", "sie0", "rexp", "sexp", "rho_min_eos", "sie_min_eos", "gam_max", "sie_max",
              "bsqorho_max", "bsqou_max", "enable_floors", "enable_mhd_floors", "enable_c2p_fixup", "enable_flux_fixup",
              "c2p_failure_strategy", "Cv", "bc_prim", "ox1_inflow_check"):
        assert getattr(c, k) == p[k], k
    assert c.gm1 == p["gamma"] - 1.0
    from oracle import pyoracle as po
    assert c.recon == po.RECON[p["recon"]] and c.riemann == po.RIEMANN[p["riemann"]]
    assert c.floor_type == (po.FLOOR[p["floor_type"]] if isinstance(p["floor_type"], str) else p["floor_type"])
    assert tuple(c.nx) == tuple(p["nx"])
    assert fp.mesh_bc == p["mesh_bc"] and fp.cfl == p["cfl"] and fp.dt_init_fact == p["dt_init_fact"]


@pytest.mark.parametrize("name", CASES)
def test_block_coordinates_and_bcs_match_reference(name):
    g, p = load_case(name)
    mesh, fp = product_mesh(g, alloc_fields=False)
    man = g.manifest("init")
    assert mesh.nblocks == len(man["blocks"])
    for mb in man["blocks"]:
        n = mesh.local_index[block_lx(mesh, man, mb, mesh.ndim, mesh.ng, mesh.nx)]
        assert np.array_equal(mesh.bxmin[n], np.array(mb["xmin"])), (mb, mesh.bxmin[n])
        assert np.array_equal(mesh.bdx[n], np.array(mb["dx"]))
        want = fd.block_bc(p, mb)
        got = list(mesh.blocks_host[n].bc)
        assert got[:2 * mesh.ndim] == want[:2 * mesh.ndim]


@pytest.mark.parametrize("name", CASES)
def test_geometry_tables_match_reference_cache(name):
    """pb_build_geometry_host vs the reference's Geometry::Cached fields: bit-exact, FMKS metric
    derivatives by finite differences included (Minkowski tables are 1x1x1 broadcasts)."""
    g, p = load_case(name)
    mesh, fp = product_mesh(g, alloc_fields=False)
    man = g.manifest("init")
    if g.meta["geometry"] == 1:
        gp = man["geom_params"]
        for k, v in gp.items():
            assert fp.geom_par[k] == v, k
    get = getter(g, "init")
    for mb in man["blocks"]:
        n = mesh.local_index[block_lx(mesh, man, mb, mesh.ndim, mesh.ng, mesh.nx)]
        ref, gshape = fd.geom_from_dump(get, mb["gid"], mesh.shape)
        mine = mesh.geom_host[mesh.geom_slot[n]].reshape((-1,) + mesh.gshape)
        if mesh.gshape == (1, 1, 1):
            assert np.array_equal(np.broadcast_to(mine, ref.shape), ref)
        else:
            assert mine.shape == ref.shape
            assert np.array_equal(mine, ref)


def test_floor_table():
    g, p = load_case("torus2d")
    mesh, fp = product_mesh(g, alloc_fields=False)
    assert fp.floor_type == "ExpX1RhoSie"
    for n in range(mesh.nblocks):
        t = mesh.floor_host[mesh.floor_slot[n]].reshape(2, 2, mesh.ni)
        x1c = mesh.bxmin[n, 0] + (np.arange(mesh.ni) + 0.5) * mesh.bdx[n, 0]
        x1f = mesh.bxmin[n, 0] + np.arange(mesh.ni) * mesh.bdx[n, 0]
        import math
        assert np.array_equal(t[0, 0], np.array([fp.rho0 * math.exp(fp.rexp * x) for x in x1c]))
        assert np.array_equal(t[1, 1], np.array([fp.sie0 * math.exp(fp.sexp * x) for x in x1f]))


def _mk_params(nx, bnx, periodic=True, ndim=3):
    from phoebus_b200.params import FluidParams, ParameterInput
    bc = "periodic" if periodic else "gr_outflow"
    txt = "<parthenon/mesh>\nnghost = 4\n"
    for d in range(3):
        txt += "nx%d = %d\nx%dmin = 0\nx%dmax = 1\nix%d_bc = %s\nox%d_bc = %s\n" % (d + 1, nx[d], d + 1, d + 1, d + 1, bc, d + 1, bc)
    txt += "<parthenon/meshblock>\n" + "".join("nx%d = %d\n" % (d + 1, bnx[d]) for d in range(3))
    txt += "<fluid>\nmhd = true\nrecon = weno5\n"
    return FluidParams(ParameterInput(txt))


@pytest.mark.parametrize("world,periodic", [(2, True), (4, True), (8, True), (8, False), (2, False)])
def test_halo_plans_pair_up(world, periodic):
    """every rank's send plan to a peer has the same segment sizes, in the same order, as the
    peer's receive plan from it, and the receive plans cover exactly the ghost zones whose owner
    lives on another rank"""
    from phoebus_b200.mesh import MeshPack
    fp = _mk_params((32, 16, 16), (8, 8, 8), periodic)
    meshes = [MeshPack(fp, rank=r, world=world, alloc_fields=False) for r in range(world)]
    assert sum(m.nblocks for m in meshes) == 4 * 2 * 2
    def sizes(plan):
        return [int((r[2] - r[1] + 1) * (r[4] - r[3] + 1) * (r[6] - r[5] + 1)) for r in plan]
    for r, m in enumerate(meshes):
        for peer, ((splan, sn), (rplan, rn)) in m.halo_plans().items():
            (psplan, psn), (prplan, prn) = meshes[peer].halo_plans()[r]
            assert sn == prn and rn == psn
            assert sizes(splan) == sizes(prplan)
            assert sizes(rplan) == sizes(psplan)
        # coverage: ghost zones with a remote owner
        want = 0
        for n, lx in enumerate(m.lx):
            for code in range(27):
                if m.blocks_host[n].nbr[code] == -2:
                    o = (code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1)
                    want += int(np.prod([m.nx[d] if o[d] == 0 else m.ng for d in range(3)]))
        got = sum(sum(sizes(rp)) for (_, _), (rp, _) in m.halo_plans().values())
        assert got == want


def test_rank_grid_bricks():
    from phoebus_b200.mesh import rank_grid
    assert rank_grid(8, [16, 16, 16], 3) == [2, 2, 2]
    assert sorted(rank_grid(4, [16, 8, 8], 3)) == [1, 2, 2]
    assert rank_grid(2, [8, 4, 4], 3) == [2, 1, 1]
    assert rank_grid(1, [8, 4, 4], 3) == [1, 1, 1]


def test_abi_argument_errors_without_a_gpu():
    """error behaviour of the C ABI that needs no device: bad arguments return PB_EINVAL (-1) with a message,
    never throw, never touch CUDA; options round-trip; on a box without a GPU a launch reports PB_ECUDA"""
    from phoebus_b200 import _lib
    L = _lib.lib()
    g, p = load_case("torus2d")
    mesh, fp = product_mesh(g, alloc_fields=False)
    cfg = mesh.cfg
    pack = _lib.PbPack()
    assert L.pb_calculate_fluxes(None, C.byref(cfg), None) == -1 and b"NULL" in L.pb_last_error()
    assert L.pb_calculate_fluxes(C.byref(pack), C.byref(cfg), None) == -1 and b"no blocks" in L.pb_last_error()
    pack.nblocks = 1
    assert L.pb_cons_to_prim(C.byref(pack), C.byref(cfg), None, None, None, None, None) == -1
    assert b"NULL pack field" in L.pb_last_error()
    bad = type(cfg).from_buffer_copy(cfg)
    bad.ndim = 4
    assert L.pb_flux_ct(C.byref(pack), C.byref(bad), None) == -1 and b"ndim" in L.pb_last_error()
    assert L.pb_set_option(b"no_such_option", 1) == -1
    assert L.pb_set_option(b"c2p_arith", 7) == -1
    old = L.pb_get_option(b"flux_split")
    assert L.pb_set_option(b"flux_split", 0) == 0 and L.pb_get_option(b"flux_split") == 0
    assert L.pb_set_option(b"flux_split", old) == 0
    assert L.pb_nprim(C.byref(cfg)) == 10 and L.pb_ncons(C.byref(cfg)) == 8
    assert L.pb_cells_per_block(C.byref(cfg)) == mesh.ncell
    assert L.pb_selftest_arith(0, 1, 8, None, None) == -1
