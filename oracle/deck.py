"""TEST INFRASTRUCTURE ONLY — minimal Athena/Parthenon input-deck reader and the mapping
from the reference's input keys to the oracle parameter dict (fluid.cpp:53-137,
fixup.cpp:53-277, eos_phoebus.cpp:73-95)."""
import re

from .pyoracle import DBL_MAX, default_params


def parse_deck(text, overrides=()):
    blocks, cur = {}, None
    for line in text.splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        m = re.match(r"<([^>]+)>", line)
        if m:
            cur = blocks.setdefault(m.group(1), {})
            continue
        if "=" in line and cur is not None:
            k, v = line.split("=", 1)
            cur[k.strip()] = v.strip()
    for ov in overrides:
        path, v = ov.split("=", 1)
        blk, key = path.rsplit("/", 1)
        blocks.setdefault(blk, {})[key] = v
    return blocks


def _b(v):
    return str(v).lower() in ("true", "1")


def params_from_deck(blocks):
    p = default_params()
    mesh = blocks.get("parthenon/mesh", {})
    mb = blocks.get("parthenon/meshblock", {})
    nx = [int(mesh.get("nx%d" % d, 1)) for d in (1, 2, 3)]
    ndim = 3 if nx[2] > 1 else (2 if nx[1] > 1 else 1)
    bnx = [int(mb.get("nx%d" % d, nx[d - 1])) for d in (1, 2, 3)]
    p.update(ndim=ndim, nx=tuple(bnx), mesh_nx=tuple(nx), ng=int(mesh.get("nghost", 2)))
    fl = blocks.get("fluid", {})
    p["mhd"] = int(_b(fl.get("mhd", "false")))
    p["recon"] = fl.get("recon", "linear")
    p["riemann"] = fl.get("riemann", "hll")
    p["cfl"] = float(fl.get("cfl", 0.8))
    p["c2p_tol"] = float(fl.get("c2p_tol", 1e-8))
    p["c2p_max_iter"] = int(fl.get("c2p_max_iter", 20))
    p["c2p_floor_scale_fac"] = float(fl.get("c2p_floor_scale_fac", 1.0))
    p["c2p_fail_on_floors"] = int(_b(fl.get("c2p_fail_on_floors", "false")))
    p["c2p_fail_on_ceilings"] = int(_b(fl.get("c2p_fail_on_ceilings", "false")))
    eos = blocks.get("eos", {})
    p["gamma"] = float(eos.get("Gamma", 5.0 / 3.0))
    p["Cv"] = float(eos.get("Cv", 1.0))
    p["rho_min_eos"] = float(eos.get("rho_min", 0.0))
    p["sie_min_eos"] = float(eos.get("sie_min", 0.0))
    fx = blocks.get("fixup", {})
    en_floors = _b(fx.get("enable_floors", "false"))
    en_mhd = _b(fx.get("enable_mhd_floors", "false")) and p["mhd"]
    en_c2p = _b(fx.get("enable_c2p_fixup", "false"))
    if en_c2p or en_mhd:
        en_floors = True  # fixup.cpp:117-131
    p["enable_floors"], p["enable_mhd_floors"], p["enable_c2p_fixup"] = int(en_floors), int(en_mhd), int(en_c2p)
    p["enable_flux_fixup"] = int(_b(fx.get("enable_flux_fixup", "false")))
    p["ox1_inflow_check"] = int(_b(fx.get("enable_ox1_fmks_inflow_check", "false")))
    if en_floors:
        ft = fx["floor_type"]
        p["floor_type"] = ft
        p["rho0"] = float(fx.get("rho0_floor", 0.0))
        if ft in ("ConstantRhoSie", "ExpX1RhoSie"):
            p["sie0"] = float(fx.get("sie0_floor", 0.0))
            p["rexp"] = float(fx.get("rho_exp_floor", -2.0))
            p["sexp"] = float(fx.get("sie_exp_floor", -1.0))
        else:
            p["sie0"] = float(fx.get("u0_floor", 0.0))
            p["rexp"] = float(fx.get("rho_exp_floor", -2.0))
            p["sexp"] = float(fx.get("u_exp_floor", -3.0))
    if _b(fx.get("enable_ceilings", "false")):
        p["gam_max"] = float(fx.get("gam0_ceiling", 1000.0))
        p["sie_max"] = float(fx.get("sie0_ceiling", 100.0))
    if en_mhd:
        p["bsqorho_max"] = float(fx.get("bsqorho0_ceiling", 50.0))
        p["bsqou_max"] = float(fx.get("bsqou0_ceiling", 2500.0))
    p["c2p_failure_strategy"] = 0 if fx.get("fluid_c2p_failure_strategy", "interpolate") == "interpolate" else 1
    p["bc_prim"] = int(blocks.get("phoebus/mesh", {}).get("bc_vars", "conserved") == "primitive")
    ph = blocks.get("phoebus", {})
    bcs = []
    for d in (1, 2, 3):
        for io in ("i", "o"):
            key = "%sx%d_bc" % (io, d)
            parth = mesh.get(key, "periodic")
            bcs.append("periodic" if parth == "periodic" else ph.get(key, parth))
    p["mesh_bc"] = bcs
    p["dt_init_fact"] = float(blocks.get("parthenon/time", {}).get("dt_init_fact", 1.0))
    return p
