"""The ghost plan (phoebus_b200/csrc/ghost_plan_host.cpp): ghost zones whose con2prim result the stage sequencer copies
from the same-GPU neighbour instead of solving again. CPU tests of the host-side plan builder, including the claim the
shortcut rests on, checked against the UNMODIFIED reference's own per-stage dumps: in every deferred ghost zone the
reference holds exactly what it holds in the neighbour's interior zone."""
import numpy as np
import pytest

from oracle import from_dump as fd
from tests.helpers import CASES, block_lx, getter, load_case, product_mesh, stage_tags

PRIM_BC = ["torus2d", "torus3d", "torus2d_fail", "torus2d_polar", "reflect2d_prim"]


def _plan(name):
    g, p = load_case(name)
    mesh, fp = product_mesh(g, alloc_fields=False)
    mask, zones, twins = mesh.ghost_plan_host()
    mesh.plan_twins = twins
    return g, mesh, fp, mask.reshape((mesh.nblocks,) + mesh.shape), zones


def _offsets(mesh):
    """per zone of a block: neighbour-slot code and source cell index"""
    ng, nd = mesh.ng, mesh.ndim
    nk, nj, ni = mesh.shape
    k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
    def off(c, n, active):
        if not active:
            return np.zeros_like(c)
        return np.where(c < ng, -1, np.where(c >= ng + n, 1, 0))
    o1, o2, o3 = off(i, mesh.nx[0], True), off(j, mesh.nx[1], nd >= 2), off(k, mesh.nx[2], nd >= 3)
    code = (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)
    src = ((k - o3 * mesh.nx[2]) * nj + (j - o2 * mesh.nx[1])) * ni + (i - o1 * mesh.nx[0])
    return (o1, o2, o3), code, src


@pytest.mark.parametrize("name", CASES)
def test_plan_is_a_partition_and_respects_the_conditions(name):
    g, mesh, fp, mask, zones = _plan(name)
    ncell = mesh.ncell
    tot = mesh.nblocks * ncell
    (o1, o2, o3), code, src = _offsets(mesh)
    ghost = (o1 != 0) | (o2 != 0) | (o3 != 0)
    # interior zones + listed ghost zones + deferred ghost zones = every zone exactly once
    seen = np.tile((~ghost).reshape(-1).astype(np.int32), mesh.nblocks)
    np.add.at(seen, zones.astype(np.int64), 1)
    assert np.array_equal(seen + (mask.reshape(-1) != 0), np.ones(tot, dtype=np.int32))
    assert np.all(np.diff(zones.astype(np.int64)) > 0)
    # twins: exactly the zones with mask value 2, all of them x1 ghost zones; without the twin part they are ordinary listed zones
    tw = mesh.plan_twins.astype(np.int64)
    assert sorted(tw.tolist()) == np.nonzero(mask.reshape(-1) == 2)[0].tolist()
    assert (o1.reshape(-1)[tw % ncell] != 0).all()
    # list order (block, x1 ghost column, k, j) = increasing slot of the side buffer; slots are unique and inside the buffer
    if len(tw):
        from phoebus_b200 import _lib
        import ctypes as C
        S = int(_lib.lib().pb_twin_entries(C.byref(fp.make_cfg(mesh.gshape)), mesh.nblocks))
        tb, tc = tw // ncell, tw % ncell
        ti, tj, tk = tc % mesh.shape[2], (tc // mesh.shape[2]) % mesh.shape[1], tc // (mesh.shape[2] * mesh.shape[1])
        col = np.where(ti < mesh.ng, ti, ti - mesh.nx[0])
        slot = ((tb * 2 * mesh.ng + col) * mesh.shape[0] + tk) * mesh.shape[1] + tj
        assert np.all(np.diff(slot) > 0) and slot[-1] < S and (col >= 0).all() and (col < 2 * mesh.ng).all()
    mask0, zones0, tw0 = mesh.ghost_plan_host(twins=False)
    assert len(tw0) == 0 and np.array_equal(mask0 == 1, mask.reshape(-1) == 1) and not (mask0 == 2).any()
    assert sorted(zones0.tolist()) == sorted(zones.tolist() + tw.tolist())
    if not fp.bc_prim:
        assert mask.sum() == 0
        return
    for n in range(mesh.nblocks):
        blk = mesh.blocks_host[n]
        m = mask[n] != 0
        assert not (m & ~ghost).any()  # interior zones always solve
        for d, o in enumerate((o1, o2, o3)):
            if d >= mesh.ndim:
                continue
            # no deferred zone in a face that is not towards another MeshBlock (physical, periodic)
            if blk.bc[2 * d] != 0:
                assert not (m & (o == -1)).any()
            if blk.bc[2 * d + 1] != 0:
                assert not (m & (o == 1)).any()
        nb = np.array([blk.nbr[c] for c in range(27)])
        assert (nb[code[m]] >= 0).all()  # the neighbour lives in this pack


@pytest.mark.parametrize("name", PRIM_BC)
def test_reference_dumps_hold_the_neighbours_result_in_every_deferred_zone(name):
    g, mesh, fp, mask, zones = _plan(name)
    if mesh.nblocks == 1 and mask.sum() == 0:
        pytest.skip("single MeshBlock without a same-pack neighbour")
    assert (mask == 1).sum() > 0, "fixture has neighbouring MeshBlocks but no deferred zone"
    _, code, src = _offsets(mesh)
    checked = 0
    for tag in stage_tags(g):
        man = g.manifest(tag)
        get = getter(g, tag)
        state = {}
        for mb in man["blocks"]:
            n = mesh.local_index[block_lx(mesh, man, mb, mesh.ndim, mesh.ng, mesh.nx)]
            ref = fd.ref_state(get, mb["gid"], mesh.shape, man["sc1"])
            state[n] = {k: np.asarray(v).reshape((-1, mesh.ncell)) for k, v in ref.items()}
        for n in range(mesh.nblocks):
            m = mask[n].reshape(-1) == 1  # (twin zones hold their OWN result: not the neighbour's bits)
            if not m.any():
                continue
            nb = np.array([mesh.blocks_host[n].nbr[c] for c in range(27)])[code.reshape(-1)[m]]
            so = src.reshape(-1)[m]
            mine = np.nonzero(m)[0]
            for q in np.unique(nb):
                sel = nb == q
                # zones the neighbour's fixup rewrote afterwards (interior only) are excluded
                good = state[int(q)]["fail"][0, so[sel]] == 1.0
                for k in ("prim", "temp", "c2p_mu", "fail", "csig", "cons"):
                    a = state[n][k][:, mine[sel][good]]
                    b = state[int(q)][k][:, so[sel][good]]
                    assert np.array_equal(a.view(np.int64), b.view(np.int64)), (name, tag, n, int(q), k)
                checked += int(good.sum())
    assert checked > 0


@pytest.mark.parametrize("world", [2, 8])
def test_remote_faces_are_never_deferred_in_a_brick_decomposition(world):
    """one rank's brick of the 3-D torus (8 x 4 x 4 MeshBlocks of 16 x 8 x 8 over `world` GPUs): ghost zones whose source
    block lives on another GPU arrive by NCCL after con2prim could have used them — neither deferred nor twins; the
    interior faces of the brick still are"""
    from phoebus_b200 import _lib
    from phoebus_b200.mesh import MeshPack
    g, p = load_case("torus3d")
    from phoebus_b200.params import FluidParams, ParameterInput
    pin = ParameterInput.from_blocks(g.meta["deck_blocks"])
    for k, v in g.manifest("init")["mesh"].items():
        pin.blocks.setdefault("parthenon/mesh", {})[k] = repr(float(v))
    pin.blocks.setdefault("geometry", {})["kind"] = "fmks"
    pin.blocks["parthenon/mesh"].update(nx1="128", nx2="32", nx3="32")
    fp = FluidParams(pin)
    seen_remote = seen_deferred = 0
    for rank in range(world):
        mesh = MeshPack(fp, device=None, rank=rank, world=world, alloc_fields=False)
        mask, zones, twins = mesh.ghost_plan_host()
        mask = mask.reshape((mesh.nblocks,) + mesh.shape)
        (o1, o2, o3), code, src = _offsets(mesh)
        for n in range(mesh.nblocks):
            nb = np.array([mesh.blocks_host[n].nbr[c] for c in range(27)])[code]
            remote = nb == _lib.PB_NBR_REMOTE
            seen_remote += int(remote.sum())
            assert not mask[n][remote].any()
            assert (nb[mask[n] != 0] >= 0).all()
        seen_deferred += int((mask == 1).sum())
    assert seen_remote > 0 and seen_deferred > 0
