"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle.so (oracle/oracle.h) and a
block-list driver that sequences one Runge-Kutta stage the way the reference does
(src/phoebus_driver.cpp:204-738, fluid-only sub-graph).

Nothing under phoebus_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NGLOC, NGDER = 21, 33


class OrcCfg(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("nx", C.c_int * 3), ("ng", C.c_int), ("mhd", C.c_int),
        ("recon", C.c_int), ("riemann", C.c_int),
        ("c2p_tol", C.c_double), ("c2p_max_iter", C.c_int), ("c2p_floor_scale_fac", C.c_double),
        ("c2p_fail_on_floors", C.c_int), ("c2p_fail_on_ceilings", C.c_int),
        ("floor_type", C.c_int),
        ("rho0", C.c_double), ("sie0", C.c_double), ("rexp", C.c_double), ("sexp", C.c_double),
        ("rho_min_eos", C.c_double), ("sie_min_eos", C.c_double),
        ("gam_max", C.c_double), ("sie_max", C.c_double),
        ("bsqorho_max", C.c_double), ("bsqou_max", C.c_double),
        ("enable_floors", C.c_int), ("enable_mhd_floors", C.c_int), ("enable_c2p_fixup", C.c_int),
        ("enable_flux_fixup", C.c_int), ("c2p_failure_strategy", C.c_int),
        ("gm1", C.c_double), ("Cv", C.c_double),
        ("gk", C.c_int), ("gj", C.c_int), ("gi", C.c_int),
        ("xmin", C.c_double * 3), ("dx", C.c_double * 3),
        ("bc", C.c_int * 6), ("bc_prim", C.c_int), ("ox1_inflow_check", C.c_int),
    ]


DBL_MAX = np.finfo(np.float64).max
RECON = {"constant": 0, "linear": 1, "weno5": 2, "weno5z": 2, "mp5": 3}
RIEMANN = {"llf": 0, "hll": 1}
FLOOR = {"ConstantRhoSie": 0, "ExpX1RhoSie": 1, "ExpX1RhoU": 2, "X1RhoSie": 3, "RRhoSie": 4}
BC = {"gr_outflow": 1, "reflect": 2, "polar": 3, "periodic": -1, "block": 0}


def default_params():
    """Defaults of the reference's input keys (fluid.cpp:53-137, fixup.cpp:53-277,
    eos_phoebus.cpp:86-95); floors disabled => -DBL_MAX floors, DBL_MAX ceilings
    (fixup.hpp:74-76,182-185)."""
    return dict(
        ndim=3, nx=(8, 8, 8), ng=4, mhd=1, recon="linear", riemann="hll",
        c2p_tol=1e-8, c2p_max_iter=20, c2p_floor_scale_fac=1.0, c2p_fail_on_floors=0,
        c2p_fail_on_ceilings=0, floor_type="ConstantRhoSie", rho0=-DBL_MAX, sie0=-DBL_MAX,
        rexp=0.0, sexp=0.0, rho_min_eos=0.0, sie_min_eos=0.0, gam_max=DBL_MAX, sie_max=DBL_MAX,
        bsqorho_max=DBL_MAX, bsqou_max=DBL_MAX, enable_floors=0, enable_mhd_floors=0,
        enable_c2p_fixup=0, enable_flux_fixup=0, c2p_failure_strategy=0, gamma=5.0 / 3.0, Cv=1.0,
        bc_prim=0, ox1_inflow_check=0, cfl=0.8,
    )


def make_cfg(p, xmin, dx, bc, gshape):
    c = OrcCfg()
    c.ndim = p["ndim"]
    for d in range(3):
        c.nx[d] = p["nx"][d]
        c.xmin[d] = xmin[d]
        c.dx[d] = dx[d]
    c.ng = p["ng"]
    c.mhd = p["mhd"]
    c.recon = RECON[p["recon"]] if isinstance(p["recon"], str) else p["recon"]
    c.riemann = RIEMANN[p["riemann"]] if isinstance(p["riemann"], str) else p["riemann"]
    for k in ("c2p_tol", "c2p_max_iter", "c2p_floor_scale_fac", "c2p_fail_on_floors", "c2p_fail_on_ceilings",
              "rho0", "sie0", "rexp", "sexp", "rho_min_eos", "sie_min_eos", "gam_max", "sie_max", "bsqorho_max",
              "bsqou_max", "enable_floors", "enable_mhd_floors", "enable_c2p_fixup", "enable_flux_fixup",
              "c2p_failure_strategy", "Cv", "bc_prim", "ox1_inflow_check"):
        setattr(c, k, p[k])
    c.floor_type = FLOOR[p["floor_type"]] if isinstance(p["floor_type"], str) else p["floor_type"]
    c.gm1 = p["gamma"] - 1.0
    c.gk, c.gj, c.gi = gshape
    for f in range(6):
        c.bc[f] = bc[f]
    return c


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle.cpp", "oracle_geometry.cpp", "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip2 = C.c_int * 2
    cp = C.POINTER(OrcCfg)
    L.orc_calculate_fluxes.argtypes = [cp, dp, dp, dp, dp]
    L.orc_flux_ct.argtypes = [cp, dp, dp]
    L.orc_fix_fluxes.argtypes = [cp, dp]
    L.orc_flux_divergence.argtypes = [cp, dp, dp]
    L.orc_source_terms.argtypes = [cp, dp, dp, dp]
    L.orc_divb.argtypes = [cp, dp, dp]
    L.orc_history_fluid.argtypes = [cp, dp, dp, dp]
    L.orc_history_torus.argtypes = [cp, dp, dp, dp, C.c_double, C.c_double, dp]
    L.orc_rk_update.argtypes = [cp, dp, C.c_void_p, dp, dp, C.c_double, C.c_double, dp]
    L.orc_cons_to_prim.argtypes = [cp, dp, dp, dp, dp, dp, dp, dp, ip2, ip2, ip2]
    L.orc_c2p_fixup.argtypes = [cp, dp, dp, dp, dp, dp, dp]
    L.orc_apply_floors.argtypes = [cp, dp, dp, dp, dp, dp, dp, C.c_void_p]
    L.orc_prim_to_cons.argtypes = [cp, dp, dp, dp, dp, ip2, ip2, ip2]
    L.orc_apply_bcs.argtypes = [cp, dp, dp, dp, dp, dp, dp, dp]
    L.orc_apply_bcs_ex.argtypes = [cp, dp, dp, dp, dp, dp, dp, dp, C.c_int]
    L.orc_pgen_torus.argtypes = [cp, dp, dp, dp, dp, dp]
    L.orc_pgen_torus.restype = C.c_int
    L.orc_pgen_betas.argtypes = [cp, dp, dp, C.c_double, dp]
    L.orc_estimate_dt.argtypes = [cp, dp, dp, C.c_int]
    L.orc_estimate_dt.restype = C.c_double
    L.orc_recon5.argtypes = [C.c_int, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_build_geometry.argtypes = [C.c_int, dp, cp, dp]
    L.orc_geometry_point.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double, dp]
    L.orc_error_count.restype = C.c_longlong
    _lib = L
    return L


def geom_params(kind, gp=None):
    """par = {a, h, xt, alpha, x0, smooth, hexp_br, hexp_nsq, hexp_csq, derefine_poles}"""
    if kind == 0:
        return np.zeros(10)
    d = dict(a=0.0, h=0.3, xt=0.82, alpha=14.0, x0=0.0, smooth=0.5, hexp_br=1e30, hexp_nsq=4.0, hexp_csq=4.0,
             derefine_poles=1)
    d.update(gp or {})
    return np.array([d[k] for k in ("a", "h", "xt", "alpha", "x0", "smooth", "hexp_br", "hexp_nsq", "hexp_csq",
                                    "derefine_poles")], dtype=np.float64)


class Block:
    """One MeshBlock's fields (numpy, [comp][k][j][i]) + its oracle cfg."""

    def __init__(self, p, xmin, dx, bc, lx, geom, gshape):
        self.p = p
        self.cfg = make_cfg(p, xmin, dx, bc, gshape)
        self.lx = tuple(lx)
        ndim, ng = p["ndim"], p["ng"]
        self.ni = p["nx"][0] + 2 * ng
        self.nj = p["nx"][1] + 2 * ng if ndim >= 2 else 1
        self.nk = p["nx"][2] + 2 * ng if ndim >= 3 else 1
        self.shape = (self.nk, self.nj, self.ni)
        self.nprim = 7 + 3 * p["mhd"]
        self.ncons = 5 + 3 * p["mhd"]
        z = lambda n: np.zeros((n,) + self.shape)
        self.prim = z(self.nprim)
        self.temp = z(1)
        self.c2p_mu = z(1)
        self.fail = z(1)
        self.csig = z(ndim)
        self.fsig = z(ndim)
        self.emf = z(3)
        self.flux = np.zeros((ndim, self.ncons) + self.shape)
        self.cons = {"base": z(self.ncons), "1": z(self.ncons)}
        self.dudt = z(self.ncons)
        self.gsrc = z(4)
        self.geom = np.ascontiguousarray(geom, dtype=np.float64)

    def entire(self):
        I2 = C.c_int * 2
        return I2(0, self.ni - 1), I2(0, self.nj - 1), I2(0, self.nk - 1)


def halo_exchange(blocks, nblk, periodic, fields, ng, ndim):
    """Ghost fill from neighbouring blocks (Parthenon SendBoundBufs/SetBounds,
    external/parthenon/src/bvals/comms/boundary_communication.cpp:49-346): every ghost zone
    whose source lies inside the mesh (after periodic wrap) gets the owner's interior value."""
    by_lx = {b.lx: b for b in blocks}
    offs = [(a, bb, cc) for cc in ((-1, 0, 1) if ndim >= 3 else (0,)) for bb in ((-1, 0, 1) if ndim >= 2 else (0,))
            for a in (-1, 0, 1) if (a, bb, cc) != (0, 0, 0)]
    staged = []
    for b in blocks:
        n = (b.p["nx"][0], b.p["nx"][1] if ndim >= 2 else 1, b.p["nx"][2] if ndim >= 3 else 1)
        for off in offs:
            src_lx, ok = [], True
            for d in range(3):
                l = b.lx[d] + off[d]
                if l < 0 or l >= nblk[d]:
                    if periodic[d]:
                        l %= nblk[d]
                    else:
                        ok = False
                src_lx.append(l)
            if not ok:
                continue
            nb = by_lx[tuple(src_lx)]
            dst, src = [], []
            for d in range(3):
                active = d < ndim
                g = ng if active else 0
                if off[d] == 0:
                    dst.append(slice(g, g + n[d]))
                    src.append(slice(g, g + n[d]))
                elif off[d] < 0:
                    dst.append(slice(0, g))
                    src.append(slice(n[d], n[d] + g))
                else:
                    dst.append(slice(g + n[d], n[d] + 2 * g))
                    src.append(slice(g, 2 * g))
            for f in fields:
                staged.append((f(b), (slice(None), dst[2], dst[1], dst[0]), f(nb)[(slice(None), src[2], src[1], src[0])].copy()))
    for arr, sl, val in staged:
        arr[sl] = val


class OracleMesh:
    """Block list + reference stage sequence, everything computed by liboracle.so."""

    def __init__(self, p, blocks, nblk, periodic):
        self.p, self.blocks, self.nblk, self.periodic = p, blocks, nblk, periodic
        self.L = lib()
        self.has_face_speeds = False
        self.counters = np.zeros(3, dtype=np.int64)

    def stage(self, stage, dt, nstages=2):
        """RK2 (low_storage_integrator.cpp:67-86): stage names base,1,base; beta = 1, 0.5."""
        L, p = self.L, self.p
        names = ["base", "1", "base"]
        beta = [1.0, 0.5][stage - 1]
        sc0, sc1 = names[stage - 1], names[stage]
        cp = C.byref
        for b in self.blocks:
            c = b.cfg
            L.orc_source_terms(cp(c), b.geom, b.prim, b.gsrc)
            L.orc_calculate_fluxes(cp(c), b.geom, b.prim, b.flux, b.fsig)
            L.orc_flux_ct(cp(c), b.flux, b.emf)
            L.orc_fix_fluxes(cp(c), b.flux)
            L.orc_flux_divergence(cp(c), b.flux, b.dudt)
            u0 = b.cons[sc0]
            out = b.cons[sc1]
            L.orc_rk_update(cp(c), b.dudt, b.gsrc.ctypes.data, u0, b.cons["base"] if sc0 != "base" else u0.copy(),
                            beta, beta * dt, out)
            # sync_region_fixup (phoebus_driver.cpp:580-620): stale prims / previous fail flags
            L.orc_c2p_fixup(cp(c), b.geom, out, b.prim, b.temp, b.fail, b.csig)
            L.orc_apply_floors(cp(c), b.geom, out, b.prim, b.temp, b.c2p_mu, b.csig, self.counters.ctypes.data)
        self.exchange(sc1)
        for b in self.blocks:
            c = b.cfg
            out = b.cons[sc1]
            # the reference fills the physical ghosts twice (inside AddBoundaryExchangeTasks and again as set_bc,
            # phoebus_driver.cpp:646-661; the fill is idempotent) and converts them ONCE (convert_bc): with
            # bc_vars = conserved the conversion is not idempotent (its c2p advances the ghost zones' c2p_mu history)
            L.orc_apply_bcs(cp(c), b.geom, out, b.prim, b.temp, b.c2p_mu, b.fail, b.csig)
            ib, jb, kb = b.entire()
            L.orc_cons_to_prim(cp(c), b.geom, out, b.prim, b.temp, b.c2p_mu, b.fail, b.csig, ib, jb, kb)
            L.orc_c2p_fixup(cp(c), b.geom, out, b.prim, b.temp, b.fail, b.csig)
            L.orc_apply_floors(cp(c), b.geom, out, b.prim, b.temp, b.c2p_mu, b.csig, self.counters.ctypes.data)

    def exchange(self, sc):
        fields = [lambda b: b.cons[sc]]
        if self.p["bc_prim"]:
            fields += [lambda b: b.prim, lambda b: b.temp, lambda b: b.c2p_mu, lambda b: b.fail]
        halo_exchange(self.blocks, self.nblk, self.periodic, fields, self.p["ng"], self.p["ndim"])

    def estimate_dt(self):
        """min over blocks of fluid::EstimateTimestepBlock (fluid.cpp:899-944)."""
        dts = [self.L.orc_estimate_dt(C.byref(b.cfg), b.csig, b.fsig, int(self.has_face_speeds)) for b in self.blocks]
        self.has_face_speeds = True
        return self.p["cfl"] * min(dts)

    def step(self, dt):
        self.stage(1, dt)
        self.stage(2, dt)
