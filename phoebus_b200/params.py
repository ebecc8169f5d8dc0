"""Input-deck reader and the flattening of the reference's package Params into one pb_cfg_t.

Mirrors parthenon::ParameterInput (`<block>` / `key = value` / `# comment`, command-line overrides
`block/key=value`) and the GetOrAdd defaults of the packages on the hot path:
fluid (ref: src/fluid/fluid.cpp:53-137), fixup (ref: src/fixup/fixup.cpp:53-277),
eos (ref: src/microphysics/eos_phoebus/eos_phoebus.cpp:73-95), geometry/FMKS
(ref: src/geometry/fmks.cpp:38-91, src/geometry/cached_system.hpp:474-515).
"""
import re
import sys

from ._lib import PbCfg

DBL_MAX = sys.float_info.max
RECON = {"constant": 0, "linear": 1, "weno5": 2, "weno5z": 2, "mp5": 3}
RIEMANN = {"llf": 0, "hll": 1}
FLOOR = {"ConstantRhoSie": 0, "ExpX1RhoSie": 1, "ExpX1RhoU": 2, "X1RhoSie": 3, "RRhoSie": 4}
BC = {"gr_outflow": 1, "reflect": 2, "polar": 3, "periodic": -1}


class ParameterInput:
    """<block> key = value decks, as parthenon::ParameterInput reads them."""

    def __init__(self, text="", overrides=()):
        self.blocks = {}
        cur = None
        for line in text.splitlines():
            line = line.split("#")[0].strip()
            if not line:
                continue
            m = re.match(r"<([^>]+)>", line)
            if m:
                cur = self.blocks.setdefault(m.group(1), {})
            elif "=" in line and cur is not None:
                k, v = line.split("=", 1)
                cur[k.strip()] = v.strip()
        for ov in overrides:
            path, v = ov.split("=", 1)
            blk, key = path.rsplit("/", 1)
            self.blocks.setdefault(blk, {})[key] = v

    @classmethod
    def from_blocks(cls, blocks):
        p = cls()
        p.blocks = {b: dict(kv) for b, kv in blocks.items()}
        return p

    def get(self, block, key, default=None):
        return self.blocks.get(block, {}).get(key, default)

    def real(self, block, key, default):
        return float(self.get(block, key, default))

    def integer(self, block, key, default):
        return int(self.get(block, key, default))

    def boolean(self, block, key, default):
        return str(self.get(block, key, default)).lower() in ("true", "1")

    def string(self, block, key, default):
        return str(self.get(block, key, default))


class FluidParams:
    """Everything the hot path reads from the deck, plus mesh shape and boundary names."""

    def __init__(self, pin):
        g = pin
        self.mesh_nx = tuple(g.integer("parthenon/mesh", "nx%d" % d, 1) for d in (1, 2, 3))
        self.ndim = 3 if self.mesh_nx[2] > 1 else (2 if self.mesh_nx[1] > 1 else 1)
        self.block_nx = tuple(g.integer("parthenon/meshblock", "nx%d" % d, self.mesh_nx[d - 1]) for d in (1, 2, 3))
        self.ng = g.integer("parthenon/mesh", "nghost", 2)
        self.xmin = tuple(g.real("parthenon/mesh", "x%dmin" % d, 0.0) for d in (1, 2, 3))
        self.xmax = tuple(g.real("parthenon/mesh", "x%dmax" % d, 1.0) for d in (1, 2, 3))
        self.mhd = g.boolean("fluid", "mhd", False)
        self.recon = g.string("fluid", "recon", "linear")
        self.riemann = g.string("fluid", "riemann", "hll")
        if self.recon not in RECON:
            raise ValueError("Reconstruction method invalid")  # ref: fluid.cpp:113-115
        if self.riemann not in RIEMANN:
            raise ValueError("Riemann solver invalid")
        if self.recon in ("weno5", "weno5z", "mp5") and self.ng < 4:
            raise ValueError("weno5 or mp5 requires 4+ ghost cells")  # ref: fluid.cpp:98-104
        self.cfl = g.real("fluid", "cfl", 0.8)
        self.c2p_tol = g.real("fluid", "c2p_tol", 1e-8)
        self.c2p_max_iter = g.integer("fluid", "c2p_max_iter", 20)
        self.c2p_floor_scale_fac = g.real("fluid", "c2p_floor_scale_fac", 1.0)
        self.c2p_fail_on_floors = g.boolean("fluid", "c2p_fail_on_floors", False)
        self.c2p_fail_on_ceilings = g.boolean("fluid", "c2p_fail_on_ceilings", False)
        if g.string("fluid", "c2p_method", "robust") != "robust":
            raise ValueError("only c2p_method = robust is on the accelerated path")
        self.gamma = g.real("eos", "Gamma", 5.0 / 3.0)
        self.Cv = g.real("eos", "Cv", 1.0)
        self.rho_min_eos = g.real("eos", "rho_min", 0.0)
        self.sie_min_eos = g.real("eos", "sie_min", 0.0)
        # ---- fixup (ref: fixup.cpp:53-277)
        en_floors = g.boolean("fixup", "enable_floors", False)
        en_mhd = g.boolean("fixup", "enable_mhd_floors", False) and self.mhd
        en_c2p = g.boolean("fixup", "enable_c2p_fixup", False)
        if en_c2p or en_mhd:
            en_floors = True  # ref: fixup.cpp:117-131
        self.enable_floors, self.enable_mhd_floors, self.enable_c2p_fixup = en_floors, en_mhd, en_c2p
        self.enable_flux_fixup = g.boolean("fixup", "enable_flux_fixup", False)
        self.ox1_inflow_check = g.boolean("fixup", "enable_ox1_fmks_inflow_check", False)
        # disabled floors are -DBL_MAX, disabled ceilings DBL_MAX (ref: fixup.hpp:74-76,182-185)
        self.floor_type, self.rho0, self.sie0, self.rexp, self.sexp = "ConstantRhoSie", -DBL_MAX, -DBL_MAX, 0.0, 0.0
        if en_floors:
            self.floor_type = g.string("fixup", "floor_type", "ConstantRhoSie")
            if self.floor_type not in FLOOR:
                raise ValueError("invalid floor_type")
            self.rho0 = g.real("fixup", "rho0_floor", 0.0)
            if self.floor_type in ("ExpX1RhoU", "X1RhoSie", "RRhoSie"):  # ref: fixup.cpp:148-165
                self.sie0 = g.real("fixup", "u0_floor", 0.0)
                self.rexp = g.real("fixup", "rho_exp_floor", -2.0)
                self.sexp = g.real("fixup", "u_exp_floor", -3.0)
            else:
                self.sie0 = g.real("fixup", "sie0_floor", 0.0)
                self.rexp = g.real("fixup", "rho_exp_floor", -2.0)
                self.sexp = g.real("fixup", "sie_exp_floor", -1.0)
        self.gam_max, self.sie_max = DBL_MAX, DBL_MAX
        if g.boolean("fixup", "enable_ceilings", False):
            self.gam_max = g.real("fixup", "gam0_ceiling", 1000.0)
            self.sie_max = g.real("fixup", "sie0_ceiling", 100.0)
        self.bsqorho_max, self.bsqou_max = DBL_MAX, DBL_MAX
        if en_mhd:
            self.bsqorho_max = g.real("fixup", "bsqorho0_ceiling", 50.0)
            self.bsqou_max = g.real("fixup", "bsqou0_ceiling", 2500.0)
        strat = g.string("fixup", "fluid_c2p_failure_strategy", "interpolate")
        if strat not in ("interpolate", "floors"):
            raise ValueError("fixup/fluid_c2p_failure_strategy invalid")
        self.c2p_failure_strategy = 0 if strat == "interpolate" else 1
        self.bc_prim = g.string("phoebus/mesh", "bc_vars", "conserved") == "primitive"
        # boundary names: parthenon flag "user" defers to <phoebus> ix1_bc etc. (ref: phoebus_boundaries.cpp:450-543)
        self.mesh_bc = []
        for d in (1, 2, 3):
            for io in ("i", "o"):
                key = "%sx%d_bc" % (io, d)
                parth = g.string("parthenon/mesh", key, "periodic")
                self.mesh_bc.append("periodic" if parth == "periodic" else g.string("phoebus", key, parth))
        self.dt_init_fact = g.real("parthenon/time", "dt_init_fact", 1.0)
        self.integrator = g.string("parthenon/time", "integrator", "rk2")
        # ---- geometry: a compile-time choice in the reference (-DPHOEBUS_GEOMETRY, ref:
        # cmake/Geometry.cmake:13-47), a run-time one here: <geometry> kind = minkowski | fmks.
        # FMKS keys as read by ref: src/geometry/fmks.cpp:38-91
        self.geometry = g.string("geometry", "kind", "minkowski").lower()
        x1min, x1max = self.xmin[0], self.xmax[0]
        dx1 = (x1max - x1min) / self.mesh_nx[0]
        x0 = min(x1min, max(0.0, x1min - self.ng * dx1))
        self.geom_par = dict(
            a=g.real("geometry", "a", 0.0), h=g.real("coordinates", "hslope", 0.3),
            xt=g.real("coordinates", "poly_xt", 0.82), alpha=g.real("coordinates", "poly_alpha", 14.0), x0=x0,
            smooth=g.real("coordinates", "smooth", 0.5), hexp_br=g.real("coordinates", "hexp_br", 1e30),
            hexp_nsq=g.real("coordinates", "hexp_nsq", 4.0), hexp_csq=g.real("coordinates", "hexp_csq", 4.0),
            derefine_poles=float(g.boolean("coordinates", "derefine_poles", True)))
        self.axisymmetric = g.boolean("geometry", "axisymmetric", False)
        # ---- <torus> (ref: src/pgen/torus.cpp:176-185,526-531)
        self.torus = dict(rin=g.real("torus", "rin", 6.0), rmax=g.real("torus", "rmax", 12.0), kappa=g.real("torus", "kappa", 1.e-2),
                          u_jitter=g.real("torus", "u_jitter", 1.e-2), seed=g.integer("torus", "seed", 0),
                          nsub=g.integer("torus", "nsub", 1), magnetized=g.boolean("torus", "magnetized", True),
                          target_beta=g.real("torus", "target_beta", 100.0),
                          harm_beta_normalization=g.boolean("torus", "harm_beta_normalization", True),
                          rho_min_bnorm=g.real("torus", "rho_min_bnorm", 1.e-4))
        # the reference picks the ghost-fill function from <parthenon/mesh> and the FixFluxes branch from <phoebus>
        # (phoebus_boundaries.cpp:497-505 vs fluid.cpp:178-189); one code per face here, so the two must agree
        for d in (1, 2, 3):
            for io in ("i", "o"):
                key = "%sx%d_bc" % (io, d)
                parth = g.string("parthenon/mesh", key, "periodic")
                ph = g.get("phoebus", key)
                if parth != "periodic" and ph is not None and ph != parth:
                    raise ValueError("<phoebus>/%s = %s disagrees with <parthenon/mesh>/%s = %s" % (key, ph, key, parth))
        if any(b == "polar" for b in self.mesh_bc) and not self.bc_prim:
            raise ValueError("Polar X2 reflecting BCs not supported")  # ref: phoebus_boundaries.cpp:167,199 (bc_vars == primitive)

    def geom_par_array(self):
        import numpy as np
        return np.array([self.geom_par[k] for k in ("a", "h", "xt", "alpha", "x0", "smooth", "hexp_br", "hexp_nsq",
                                                    "hexp_csq", "derefine_poles")], dtype=np.float64)

    def make_cfg(self, gshape):
        """pb_cfg_t for blocks of this mesh; gshape = geometry table extents (gk, gj, gi)."""
        c = PbCfg()
        c.ndim = self.ndim
        for d in range(3):
            c.nx[d] = self.block_nx[d]
        c.ng = self.ng
        c.mhd = int(self.mhd)
        c.recon = RECON[self.recon]
        c.riemann = RIEMANN[self.riemann]
        c.c2p_tol, c.c2p_max_iter, c.c2p_floor_scale_fac = self.c2p_tol, self.c2p_max_iter, self.c2p_floor_scale_fac
        c.c2p_fail_on_floors, c.c2p_fail_on_ceilings = int(self.c2p_fail_on_floors), int(self.c2p_fail_on_ceilings)
        c.floor_type = FLOOR[self.floor_type]
        c.rho0, c.sie0, c.rexp, c.sexp = self.rho0, self.sie0, self.rexp, self.sexp
        c.rho_min_eos, c.sie_min_eos = self.rho_min_eos, self.sie_min_eos
        c.gam_max, c.sie_max, c.bsqorho_max, c.bsqou_max = self.gam_max, self.sie_max, self.bsqorho_max, self.bsqou_max
        c.enable_floors, c.enable_mhd_floors = int(self.enable_floors), int(self.enable_mhd_floors)
        c.enable_c2p_fixup, c.enable_flux_fixup = int(self.enable_c2p_fixup), int(self.enable_flux_fixup)
        c.c2p_failure_strategy = self.c2p_failure_strategy
        c.gm1, c.Cv = self.gamma - 1.0, self.Cv
        c.gk, c.gj, c.gi = gshape
        c.bc_prim, c.ox1_inflow_check = int(self.bc_prim), int(self.ox1_inflow_check)
        return c

    @property
    def nprim(self):
        return 7 + 3 * int(self.mhd)

    @property
    def ncons(self):
        return 5 + 3 * int(self.mhd)

    def rk_betas(self):
        """ref: external/parthenon/src/time_integration/low_storage_integrator.cpp:67-86"""
        if self.integrator == "rk1":
            return [1.0]
        if self.integrator == "rk2":
            return [1.0, 0.5]
        raise ValueError("integrator %s is not on the accelerated path" % self.integrator)
