"""Shared helpers for parity tests: load golden fixtures into oracle meshes."""
import os

from oracle import deck, from_dump as fd, refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["shocktube", "linmodes2d", "blast3d", "torus2d",
         # round 2: 3-D FMKS (the benchmarked configuration), real c2p failures, polar / reflect / conserved-FMKS boundaries
         "torus3d", "torus2d_fail", "torus2d_polar", "torus2d_cons", "reflect3d", "reflect2d_prim",
         # the reference's default ghost depth for linear reconstruction (nghost = 2): stencil loads stay inside the slab
         "linmodes2d_ng2", "shocktube_ng2"]


def load_case(name):
    g = refdump.Golden(os.path.join(GOLDEN, name + ".npz"))
    p = deck.params_from_deck(g.meta["deck_blocks"])
    return g, p


def getter(g, tag):
    return lambda gid, cont, nm: g.get(tag, gid, cont, nm)


def oracle_mesh(g, p):
    m = fd.mesh_from_dump(p, g.manifest("init"), getter(g, "init"))
    m.has_face_speeds = True  # the reference already called EstimateTimestep at init
    return m


def stage_tags(g):
    return [t for t in g.meta["tags"] if t != "init"]


# ---------------------------------------------------------------------------------------------
# product-side helpers: build a phoebus_b200 MeshPack from a golden fixture
def product_params(g):
    """FluidParams from the fixture's deck; mesh bounds and FMKS parameters as the reference had
    them AFTER phoebus::ProblemModifier (the torus pgen rewrites x1min/x1max)."""
    from phoebus_b200.params import FluidParams, ParameterInput
    pin = ParameterInput.from_blocks(g.meta["deck_blocks"])
    man = g.manifest("init")
    mesh = pin.blocks.setdefault("parthenon/mesh", {})
    for k, v in man["mesh"].items():
        mesh[k] = repr(float(v))
    pin.blocks.setdefault("geometry", {})["kind"] = "fmks" if g.meta["geometry"] == 1 else "minkowski"
    return FluidParams(pin)


def block_lx(p_or_mesh, man, mb, ndim, ng, nx):
    mesh_min = [man["mesh"]["x%dmin" % (d + 1)] for d in range(3)]
    lx = []
    for d in range(3):
        gg = ng if d < ndim else 0
        x0 = mb["xmin"][d] + gg * mb["dx"][d]
        lx.append(int(round((x0 - mesh_min[d]) / (nx[d] * mb["dx"][d]))))
    return tuple(lx)


def product_mesh(g, device=None, alloc_fields=True, **kw):
    from phoebus_b200.mesh import MeshPack
    fp = product_params(g)
    mesh = MeshPack(fp, device=device, alloc_fields=alloc_fields, **kw)
    return mesh, fp


def load_product_state(mesh, g, tag="init"):
    """copy a dump's fields into the pack (host -> device)"""
    import numpy as np
    import torch
    man = g.manifest(tag)
    get = getter(g, tag)
    order = {}
    for mb in man["blocks"]:
        lx = block_lx(mesh, man, mb, mesh.ndim, mesh.ng, mesh.nx)
        n = mesh.local_index[lx]
        order[mb["gid"]] = n
        gid = mb["gid"]
        shape = mesh.shape
        st = dict(prim=fd.stack(get, gid, "base", fd.PRIM_FIELDS, shape),
                  temp=get(gid, "base", "temperature").reshape(shape),
                  c2p_mu=get(gid, "base", "c2p_mu").reshape(shape),
                  fail=get(gid, "base", "fail").reshape(shape),
                  csig=get(gid, "base", "cell_signal_speed").reshape((-1,) + shape),
                  fsig=get(gid, "base", "face_signal_speed").reshape((-1,) + shape))
        for k, v in st.items():
            getattr(mesh, k)[n].copy_(torch.from_numpy(np.ascontiguousarray(v)))
        for cont in ("base", "1"):
            if get(gid, cont, "c.density") is not None:
                mesh.cons[cont][n].copy_(torch.from_numpy(fd.stack(get, gid, cont, fd.CONS_FIELDS, shape)))
    mesh.invalidate_flags()
    return order
