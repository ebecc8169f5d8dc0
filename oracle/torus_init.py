"""TEST INFRASTRUCTURE ONLY — the reference's start-up sequence for the torus problem, restated with the oracle's
task functions, so that the problem generator can be pinned against the reference's "init" dumps:

  Mesh::Initialize (external/parthenon/src/mesh/mesh.cpp:745-856)
      ProblemGenerator (src/pgen/torus.cpp:130-474): prims -> vector potential -> B -> PrimitiveToConserved -> ApplyFloors
      CommunicateBoundaries (mesh.cpp:640-707): ghost exchange + physical ghost fill (NO ConvertBoundaryConditions)
      FillDerived (mesh.cpp:725-739): ConservedToPrimitive over the entire block
  [oracle/ref_harness/ref_dump.cpp REF_DUMP_PERTURB: prims x (1 + amp (hash - 0.5)), PrimitiveToConserved]
  PostInitializationModifier (torus.cpp:525-737): ComputeBetas, B *= sqrt(beta / beta_target) on the interior,
      PrimitiveToConserved over the entire block
  (PostInitializationCommunication is off by default: phoebus_package.cpp:36-37.)
"""
import ctypes as C
import math

import numpy as np

from .pyoracle import lib


def hash_unit(gi, gj, gk, salt):
    """HashUnit of oracle/ref_harness/ref_dump.cpp on integer arrays (broadcast)"""
    M = np.uint64
    with np.errstate(over="ignore"):
        x = ((gi + 4096).astype(M) * M(73856093)) ^ ((gj + 4096).astype(M) * M(19349663)) ^ \
            ((gk + 4096).astype(M) * M(83492791)) ^ (M(salt) * M(2654435761))
        x = x ^ (x >> M(15))
        x = x * M(0x2c1b3c6d)
        x = x ^ (x >> M(12))
    return (x & M(1023)).astype(np.float64) / 1024.0


def perturbation_factors(shape, ng, ndim, origin, mesh_nx3, amp):
    """the three factor arrays [k][j][i] (rho, u, v^3) of PerturbState for a block whose first interior zone has the
    global index `origin`"""
    nk, nj, ni = shape
    i = np.arange(ni, dtype=np.int64)[None, None, :] + (origin[0] - ng)
    j = (np.arange(nj, dtype=np.int64)[None, :, None] + (origin[1] - ng)) if ndim >= 2 else np.zeros((1, 1, 1), dtype=np.int64)
    if ndim >= 3:
        k = np.arange(nk, dtype=np.int64)[:, None, None] + (origin[2] - ng)
        k = ((k % mesh_nx3) + mesh_nx3) % mesh_nx3
    else:
        k = np.zeros((1, 1, 1), dtype=np.int64)
    i, j, k = np.broadcast_arrays(i, j, k)
    return [1.0 + amp * (hash_unit(i, j, k, s) - 0.5) for s in (1, 2, 3)]


def block_origin(b, mesh_xmin, ndim, ng):
    """global index of the block's first interior zone, as PerturbState computes it (lround of a coordinate ratio)"""
    o = []
    for d in range(3):
        if d >= ndim:
            o.append(0)
            continue
        xf = b.cfg.xmin[d] + ng * b.cfg.dx[d]
        o.append(int(round((xf - mesh_xmin[d]) / b.cfg.dx[d])))
    return o


def torus_params(deck_blocks):
    t = deck_blocks.get("torus", {})
    return dict(rin=float(t.get("rin", 6.0)), rmax=float(t.get("rmax", 12.0)), kappa=float(t.get("kappa", 1e-2)),
                nsub=int(t.get("nsub", 1)), u_jitter=float(t.get("u_jitter", 1e-2)),
                target_beta=float(t.get("target_beta", 100.0)), rho_min_bnorm=float(t.get("rho_min_bnorm", 1e-4)),
                harm=str(t.get("harm_beta_normalization", "true")).lower() in ("true", "1"),
                magnetized=str(t.get("magnetized", "true")).lower() in ("true", "1"))


def oracle_torus_init(om, geom_par, tpar, mesh_xmin, mesh_nx3, perturb_amp=None):
    """run the sequence above on an OracleMesh whose blocks carry coordinates, boundary codes and geometry tables;
    fields start from zero, as Kokkos views do"""
    L = lib()
    p = om.p
    cp = C.byref
    tp = np.array([tpar["rin"], tpar["rmax"], tpar["kappa"], p["gamma"], tpar["nsub"], tpar["u_jitter"]])
    par = np.ascontiguousarray(geom_par, dtype=np.float64)
    for b in om.blocks:
        for a in (b.prim, b.temp, b.c2p_mu, b.fail, b.csig, b.fsig, b.cons["base"], b.cons["1"]):
            a[...] = 0.0
        rc = L.orc_pgen_torus(cp(b.cfg), par, tp, b.geom, b.prim, b.temp)
        assert rc == 0
        ib, jb, kb = b.entire()
        L.orc_prim_to_cons(cp(b.cfg), b.geom, b.prim, b.cons["base"], b.csig, ib, jb, kb)
        L.orc_apply_floors(cp(b.cfg), b.geom, b.cons["base"], b.prim, b.temp, b.c2p_mu, b.csig, None)
    om.exchange("base")
    for b in om.blocks:
        L.orc_apply_bcs_ex(cp(b.cfg), b.geom, b.cons["base"], b.prim, b.temp, b.c2p_mu, b.fail, b.csig, 1)
    for b in om.blocks:
        ib, jb, kb = b.entire()
        L.orc_cons_to_prim(cp(b.cfg), b.geom, b.cons["base"], b.prim, b.temp, b.c2p_mu, b.fail, b.csig, ib, jb, kb)
    if perturb_amp:
        for b in om.blocks:
            f = perturbation_factors(b.shape, p["ng"], p["ndim"], block_origin(b, mesh_xmin, p["ndim"], p["ng"]), mesh_nx3, perturb_amp)
            b.prim[0] *= f[0]
            b.prim[4] *= f[1]
            b.prim[3] *= f[2]
            ib, jb, kb = b.entire()
            L.orc_prim_to_cons(cp(b.cfg), b.geom, b.prim, b.cons["base"], b.csig, ib, jb, kb)
    if p["mhd"]:
        ext = np.array([np.inf, -np.inf, -np.inf])
        for b in om.blocks:
            o = np.zeros(3)
            L.orc_pgen_betas(cp(b.cfg), b.geom, b.prim, tpar["rho_min_bnorm"], o)
            ext = np.array([min(ext[0], o[0]), max(ext[1], o[1]), max(ext[2], o[2])])
        SMALL = 10 * np.finfo(np.float64).tiny
        den = 0.5 * ext[1]
        beta_pmax = ext[2] / (den + (SMALL if den >= 0 else -SMALL))
        beta_norm = beta_pmax if tpar["harm"] else ext[0]
        fac = math.sqrt(beta_norm / tpar["target_beta"]) if tpar["magnetized"] else 0.0
        ng, ndim = p["ng"], p["ndim"]
        inter = tuple(slice(ng, -ng) if d < ndim else slice(None) for d in (2, 1, 0))
        for b in om.blocks:
            for n in (5, 6, 7):
                b.prim[(n,) + inter] *= fac
            ib, jb, kb = b.entire()
            L.orc_prim_to_cons(cp(b.cfg), b.geom, b.prim, b.cons["base"], b.csig, ib, jb, kb)
        return fac
    return None
