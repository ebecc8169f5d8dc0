"""Problem generators next to the hot path.

init_torus: the reference's start-up sequence for the Fishbone-Moncrief torus (`problem = torus`), through the C ABI:

  phoebus::ProblemModifier            torus_problem_modifier (x1 range from r_outer and the horizon, ref: src/pgen/torus.cpp:476-498)
  Mesh::Initialize                    (ref: external/parthenon/src/mesh/mesh.cpp:745-856)
    ProblemGenerator                  pb_pgen_torus_table_host (host libm, one table per MeshBlock column) -> pb_pgen_torus_fill ->
                                      pb_pgen_torus_bfield -> pb_prim_to_cons -> pb_apply_floors   (ref: torus.cpp:130-474)
    CommunicateBoundaries             ghost exchange + pb_apply_bcs_ex(fill only)   (ref: mesh.cpp:640-707)
    FillDerived                       pb_cons_to_prim over the entire blocks         (ref: mesh.cpp:725-739)
  PostInitializationModifier          pb_pgen_betas -> pb_scale_bfield -> pb_prim_to_cons   (ref: torus.cpp:525-737)

The result is bit-identical to the reference's state at the start of its first cycle (tests/test_gpu_parity.py,
torus2d / torus3d "init" dumps). With torus/u_jitter != 0 the energy jitter comes from a counter-based hash of the
zone coordinates (the reference's Kokkos RNG stream depends on the thread count and is not reproduced).
"""
import ctypes as C
import math

import numpy as np

from . import _lib


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


def torus_tables(mesh):
    """host tables of the torus solution, one per geometry slot (= MeshBlock column): [nslot][5][tk][nj][ni]"""
    L = _lib.lib()
    p = mesh.p
    if mesh.geom_kind != 1:
        raise ValueError('Problem "torus" requires FMKS geometry!')  # ref: torus.cpp:132-133
    t = p.torus
    tpar = np.array([t["rin"], t["rmax"], t["kappa"], p.gamma, float(t["nsub"])], dtype=np.float64)
    gpar = np.ascontiguousarray(mesh.geom_par, dtype=np.float64)
    tk = mesh.gshape[0] if mesh.gshape[0] > 1 else 1
    nslot = mesh.geom_host.shape[0]
    out = np.zeros((nslot, 5, tk, mesh.nj, mesh.ni))
    done = set()
    D = C.POINTER(C.c_double)
    for n in range(mesh.nblocks):
        s = mesh.geom_slot[n]
        if s in done:
            continue
        done.add(s)
        xm = (C.c_double * 3)(*mesh.bxmin[n])
        dx = (C.c_double * 3)(*mesh.bdx[n])
        _lib.check(L.pb_pgen_torus_table_host(gpar.ctypes.data_as(D), tpar.ctypes.data_as(D), C.byref(mesh.cfg), xm, dx,
                                              mesh.geom_host[s].ctypes.data, out[s].ctypes.data), "pb_pgen_torus_table_host")
    return out


def init_torus(mesh, drv, after_pgen=None):
    """Fill the pack with the reference's torus initial state ("base" container + all OneCopy fields).
    after_pgen(mesh): optional hook run between Mesh::Initialize and PostInitializationModifier (the test harness's
    REF_DUMP_PERTURB acts there); it may modify mesh.prim and must leave the conserved variables to this function."""
    L = _lib.lib()
    t, p = mesh.torch, mesh.p
    pk, cfg = C.byref(mesh.pack), C.byref(mesh.cfg)
    s = drv._stream()
    base = mesh.cons["base"]
    I2 = C.c_int * 2
    ent = (I2(0, mesh.ni - 1), I2(0, mesh.nj - 1), I2(0, mesh.nk - 1))
    for f in (mesh.prim, mesh.temp, mesh.c2p_mu, mesh.fail, mesh.csig, mesh.fsig, base, mesh.cons["1"]):
        f.zero_()  # Kokkos views start from zero
    # ---- ProblemGenerator
    table = t.from_numpy(torus_tables(mesh)).to(mesh.device)
    _lib.check(L.pb_pgen_torus_fill(pk, cfg, table.data_ptr(), p.torus["u_jitter"], p.torus["seed"], s), "pb_pgen_torus_fill")
    _lib.check(L.pb_pgen_torus_bfield(pk, cfg, s), "pb_pgen_torus_bfield")
    _lib.check(L.pb_prim_to_cons(pk, cfg, base.data_ptr(), *ent, s), "pb_prim_to_cons")
    _lib.check(L.pb_apply_floors(pk, cfg, base.data_ptr(), None, s), "pb_apply_floors")
    # ---- CommunicateBoundaries (fill, no conversion) and FillDerived
    drv.exchange(base)
    _lib.check(L.pb_apply_bcs_ex(pk, cfg, base.data_ptr(), 1, s), "pb_apply_bcs_ex")
    _lib.check(L.pb_cons_to_prim(pk, cfg, base.data_ptr(), *ent, s), "pb_cons_to_prim")
    if after_pgen is not None:
        after_pgen(mesh)
        _lib.check(L.pb_prim_to_cons(pk, cfg, base.data_ptr(), *ent, s), "pb_prim_to_cons")
    # ---- PostInitializationModifier: B normalisation (ref: torus.cpp:525-545,721-731)
    fac = None
    if p.mhd:
        partial = t.zeros(3 * mesh.nblocks, dtype=t.float64, device=mesh.device)
        out = (C.c_double * 3)()
        _lib.check(L.pb_pgen_betas(pk, cfg, p.torus["rho_min_bnorm"], partial.data_ptr(), out, s), "pb_pgen_betas")
        ext = list(out)
        if drv.dist is not None:
            v = t.tensor([-ext[0], ext[1], ext[2]], dtype=t.float64, device=mesh.device)
            drv.dist.all_reduce(v, op=drv.dist.ReduceOp.MAX)
            ext = [-float(v[0]), float(v[1]), float(v[2])]
        small = 10.0 * np.finfo(np.float64).tiny
        den = 0.5 * ext[1]
        beta_pmax = ext[2] / (den + (small if den >= 0 else -small))  # robust::ratio
        beta_norm = beta_pmax if p.torus["harm_beta_normalization"] else ext[0]
        fac = math.sqrt(beta_norm / p.torus["target_beta"]) if p.torus["magnetized"] else 0.0
        _lib.check(L.pb_scale_bfield(pk, cfg, fac, s), "pb_scale_bfield")
        _lib.check(L.pb_prim_to_cons(pk, cfg, base.data_ptr(), *ent, s), "pb_prim_to_cons")
    mesh.invalidate_flags()
    drv.has_face_speeds = False  # ref: fluid.cpp:911,940 — the first EstimateTimestep uses cell speeds only
    return fac


def hash_perturbation(mesh, amp):
    """after_pgen hook reproducing oracle/ref_harness/ref_dump.cpp's REF_DUMP_PERTURB on the device: rho, u, v^3 of every
    zone times 1 + amp (h - 0.5), h an integer hash of the zone's global index (periodic in x3)"""
    t = mesh.torch
    dev = mesh.device

    def hash_unit(gi, gj, gk, salt):
        def mul(a, c):  # 64-bit wrap-around product; torch has no uint64 arithmetic, int64 wraps identically
            return a * t.tensor(c if c < (1 << 63) else c - (1 << 64), dtype=t.int64, device=dev)
        x = mul(gi + 4096, 73856093) ^ mul(gj + 4096, 19349663) ^ mul(gk + 4096, 83492791) ^ int(salt * 2654435761)
        x = x ^ ((x >> 15) & ((1 << 49) - 1))   # logical shift of the two's-complement pattern
        x = mul(x, 0x2c1b3c6d)
        x = x ^ ((x >> 12) & ((1 << 52) - 1))
        return (x & 1023).to(t.float64) / 1024.0

    def hook(mesh):
        ng, ndim = mesh.ng, mesh.ndim
        for n in range(mesh.nblocks):
            org = []
            for d in range(3):
                if d >= ndim:
                    org.append(0)
                    continue
                xf = mesh.bxmin[n, d] + ng * mesh.bdx[n, d]
                org.append(int(round((xf - mesh.mesh_xmin[d]) / mesh.bdx[n, d])))
            gi = (t.arange(mesh.ni, dtype=t.int64, device=dev) + (org[0] - ng)).view(1, 1, -1)
            gj = ((t.arange(mesh.nj, dtype=t.int64, device=dev) + (org[1] - ng)) if ndim >= 2 else t.zeros(1, dtype=t.int64, device=dev)).view(1, -1, 1)
            if ndim >= 3:
                nx3 = mesh.p.mesh_nx[2]
                gk = t.remainder(t.arange(mesh.nk, dtype=t.int64, device=dev) + (org[2] - ng), nx3).view(-1, 1, 1)
            else:
                gk = t.zeros(1, dtype=t.int64, device=dev).view(1, 1, 1)
            gi, gj, gk = t.broadcast_tensors(gi, gj, gk)
            prim = mesh.prim[n]
            prim[0] *= 1.0 + amp * (hash_unit(gi, gj, gk, 1) - 0.5)
            prim[4] *= 1.0 + amp * (hash_unit(gi, gj, gk, 2) - 0.5)
            prim[3] *= 1.0 + amp * (hash_unit(gi, gj, gk, 3) - 0.5)
    return hook
