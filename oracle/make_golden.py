#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — regenerates tests/golden/*.npz by running the UNMODIFIED
reference (oracle/_ref/ref_dump_{mink,fmks}, built by oracle/build_ref.sh from
/root/reference) on small versions of BASELINE.json's configs and packing the per-stage
field dumps. Every fixture is self-checked against the oracle restatement before it is
written. Run from the repo root: python oracle/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import deck as deckmod  # noqa: E402
from oracle import refdump  # noqa: E402

B3 = lambda n1, n2, n3, b1, b2, b3: [
    "parthenon/mesh/nx1=%d" % n1, "parthenon/mesh/nx2=%d" % n2, "parthenon/mesh/nx3=%d" % n3,
    "parthenon/meshblock/nx1=%d" % b1, "parthenon/meshblock/nx2=%d" % b2, "parthenon/meshblock/nx3=%d" % b3]

# name -> (binary, deck, overrides, nlim, geometry kind)
CASES = {
    # configs[0]: 1-D SR MHD shocktube, weno5 + hll, gr_outflow BCs
    "shocktube": ("mink", "mhd_shocktube", ["parthenon/mesh/nx1=64", "parthenon/time/dt_init_fact=1.0"], 3, 0),
    # configs[1] (2-D slice, 4 blocks, periodic halos): MHD fast mode, linear + hll, c2p_tol 1e-12
    "linmodes2d": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast"]
                   + B3(16, 16, 1, 8, 8, 1), 2, 0),
    # configs[2]: 3-D blast wave, mp5 + llf, MHD on, ceilings active (one periodic 8^3 block)
    "blast3d": ("mink", "blast_wave_3d", B3(8, 8, 8, 8, 8, 8) + [
        "parthenon/mesh/refinement=none", "fluid/mhd=true", "fixup/enable_floors=true",
        "fixup/floor_type=ConstantRhoSie", "fixup/rho0_floor=1e-2", "fixup/sie0_floor=0.5",
        "fixup/enable_c2p_fixup=true", "fixup/enable_ceilings=true", "fixup/sie0_ceiling=50",
        "fixup/gam0_ceiling=1.0001", "fixup/enable_mhd_floors=true", "sedov/rinner=0.4",
        "sedov/explosion_energy=1", "parthenon/time/dt_init_fact=1.0"], 4, 0),
    # configs[3] (2-D slice): FM torus in FMKS, weno5 + llf, primitive BCs, floors/ceilings/fixups;
    # c2p_floor_scale_fac < 1 makes ApplyFloors fire (thousands of zones per stage)
    "torus2d": ("fmks", "torus", B3(32, 16, 1, 16, 8, 1) + [
        "torus/seed=7", "torus/u_jitter=0", "fluid/c2p_floor_scale_fac=0.1",
        "parthenon/time/dt_init_fact=1e-2"], 3, 1),
}
# ---- round 2: the benchmarked configuration and the branches no round-1 fixture reached
TORUS_BASE = ["torus/seed=7", "torus/u_jitter=0", "parthenon/time/dt_init_fact=1e-2"]
CASES.update({
    # configs[3]/[4] in small: 3-D FMKS torus, 4 MeshBlocks of 16x8x8, deck defaults (c2p_floor_scale_fac = 1); the harness
    # perturbs rho, u, v^3 by +-2.5 % with a hash of the global zone index, so the state is genuinely 3-D
    "torus3d": ("fmks", "torus", B3(32, 16, 8, 16, 8, 8) + TORUS_BASE, 1, 1, {"REF_DUMP_PERTURB": "0.05"}),
    # ConservedToPrimitiveFixup on real failures: c2p_fail_on_floors marks every zone whose solve touched a floor
    "torus2d_fail": ("fmks", "torus", B3(32, 16, 1, 16, 8, 1) + TORUS_BASE + ["fluid/c2p_fail_on_floors=true"], 2, 1,
                     {"REF_DUMP_PERTURB": "0.05"}),
    # polar X2 boundaries (FixFluxes B2 mirror, sign flips in the ghost fill), primitive and conserved bc_vars
    "torus2d_polar": ("fmks", "torus", B3(32, 16, 1, 16, 8, 1) + TORUS_BASE + [
        "parthenon/mesh/ix2_bc=polar", "parthenon/mesh/ox2_bc=polar", "phoebus/ix2_bc=polar", "phoebus/ox2_bc=polar"], 2, 1,
        {"REF_DUMP_PERTURB": "0.05"}),
    # (polar requires bc_vars = primitive in the reference, phoebus_boundaries.cpp:167.) gr_outflow with
    # bc_vars = conserved in curved space: the sqrt(gamma) ratio of the conserved ghost fill is != 1
    "torus2d_cons": ("fmks", "torus", B3(32, 16, 1, 16, 8, 1) + TORUS_BASE + ["phoebus/mesh/bc_vars=conserved"], 2, 1,
                     {"REF_DUMP_PERTURB": "0.05"}),
    # reflecting walls on all six faces (FixFluxes reflect branches, ghost sign flips, edge / corner composition)
    "reflect3d": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast", "fixup/enable_flux_fixup=true"]
                  + B3(16, 8, 8, 8, 8, 8) + ["parthenon/mesh/%sx%d_bc=reflect" % (io, d) for d in (1, 2, 3) for io in "io"]
                  + ["phoebus/%sx%d_bc=reflect" % (io, d) for d in (1, 2, 3) for io in "io"], 2, 0,
                  {"REF_DUMP_PERTURB": "0.2"}),
    "reflect2d_prim": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast", "fixup/enable_flux_fixup=true",
                                                   "phoebus/mesh/bc_vars=primitive"]
                       + B3(16, 16, 1, 8, 8, 1) + ["parthenon/mesh/%sx%d_bc=reflect" % (io, d) for d in (1, 2) for io in "io"]
                       + ["phoebus/%sx%d_bc=reflect" % (io, d) for d in (1, 2) for io in "io"], 2, 0,
                       {"REF_DUMP_PERTURB": "0.2"}),
})
# the reference's DEFAULT ghost depth for linear reconstruction (nghost = 2): the stencil loads of the flux kernels must
# stay inside the slab (round-1 review), periodic halos and a reflecting shock tube at two ghost zones
CASES.update({
    "linmodes2d_ng2": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast", "parthenon/mesh/nghost=2"]
                       + B3(16, 16, 1, 8, 8, 1), 2, 0, {"REF_DUMP_PERTURB": "0.2"}),
    "shocktube_ng2": ("mink", "mhd_shocktube", ["parthenon/mesh/nx1=64", "parthenon/mesh/nghost=2", "fluid/recon=linear",
                                                "parthenon/time/dt_init_fact=1.0"], 2, 0),
})
# the production block shape: 3-D FMKS torus with 32^3 MeshBlocks (2x2x2 of them), stored as per-array hashes
# (REF_DUMP_FIELDS=hash); the test rebuilds the initial state with the product's own torus pgen + the same perturbation
HASHED = {
    "torus3d_b32": ("fmks", "torus", B3(64, 64, 64, 32, 32, 32) + TORUS_BASE, 2, 1,
                    {"REF_DUMP_PERTURB": "0.05", "REF_DUMP_FIELDS": "hash"}),
}
# larger cases for ad-hoc runs against oracle/_ref (not part of the committed fixtures)
LIVE = {
    "linmodes3d": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast", "fluid/recon=weno5"]
                   + B3(16, 16, 16, 8, 8, 8), 3, 0),
}


# long runs to the deck's own tlim (one wave period, set by the pgen): stages dumped every `every` cycles
# and for the last cycle; name -> (binary, deck, overrides, every, geometry kind)
LONG = {
    # 2-D MHD fast mode, linear + hll (deck defaults), 4 blocks of 16^2: 123 cycles
    "linmodes2d_long": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast"]
                        + B3(32, 32, 1, 16, 16, 1), 40, 0),
    # 3-D, weno5 + llf, 8 blocks of 8^3
    "linmodes3d_long": ("mink", "linear_modes_3d", ["hydro_modes/physics=mhd", "hydro_modes/mode=fast", "fluid/recon=weno5",
                                                     "fluid/riemann=llf"] + B3(16, 16, 16, 8, 8, 8), 30, 0),
}


def run_ref_long(case, outdir):
    binary, deck, ov, every, _ = case
    env = dict(os.environ, REF_DUMP_DIR=outdir, OMP_NUM_THREADS="1", REF_DUMP_EVERY=str(every))
    exe = os.path.join(HERE, "_ref", "ref_dump_" + binary)
    deckf = os.path.join(HERE, "_ref", "decks", deck + ".pin")
    with open(os.path.join(outdir, "log.txt"), "w") as log:
        subprocess.check_call([exe, "-i", deckf] + ov, cwd=outdir, env=env, stdout=log, stderr=subprocess.STDOUT)
    tags = sorted(f[:-5] for f in os.listdir(outdir) if f.endswith(".json") and f != "init.json")
    return ["init"] + tags


def run_ref_hst(case, out_hst, extra=()):
    """the reference's own executable with a history output block: one .hst row per cycle
    (ref: external/parthenon/src/outputs/history.cpp; columns from src/fluid/fluid.cpp:325-347)"""
    binary, deck, ov, _, _ = case
    exe = os.path.join(HERE, "_ref", "phoebus_" + binary)
    deckf = os.path.join(HERE, "_ref", "decks", deck + ".pin")
    tmp = tempfile.mkdtemp(prefix="hst_")
    with open(os.path.join(tmp, "log.txt"), "w") as log:
        subprocess.check_call([exe, "-i", deckf] + ov + list(extra) + ["parthenon/output0/file_type=hst", "parthenon/output0/dt=1e-12"], cwd=tmp,
                              env=dict(os.environ, OMP_NUM_THREADS="1"), stdout=log, stderr=subprocess.STDOUT)
    hst = [f for f in os.listdir(tmp) if f.endswith(".hst")]
    shutil.copy(os.path.join(tmp, hst[0]), out_hst)
    shutil.rmtree(tmp)


def run_ref(case, outdir, qlqr=False):
    binary, deck, ov, nlim = case[:4]
    env = dict(os.environ, REF_DUMP_DIR=outdir, OMP_NUM_THREADS="1")
    env.update(case[5] if len(case) > 5 else {})
    if qlqr:
        env["REF_DUMP_QLQR"] = "1"
    exe = os.path.join(HERE, "_ref", "ref_dump_" + binary)
    deckf = os.path.join(HERE, "_ref", "decks", deck + ".pin")
    with open(os.path.join(outdir, "log.txt"), "w") as log:
        subprocess.check_call([exe, "-i", deckf] + ov + ["parthenon/time/nlim=%d" % nlim], cwd=outdir, env=env,
                              stdout=log, stderr=subprocess.STDOUT)
    return ["init"] + ["c%04ds%d" % (c, s) for c in range(nlim) for s in (1, 2)]


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        tmp = tempfile.mkdtemp(prefix="gold_" + name)
        tags = run_ref(case, tmp)
        out = os.path.join(gold, name + ".npz")
        with open(os.path.join(HERE, "_ref", "decks", case[1] + ".pin")) as f:
            blocks = deckmod.parse_deck(f.read(), case[2])
        refdump.pack_golden(tmp, tags, out, extra_meta={"deck_blocks": blocks, "deck": case[1], "overrides": case[2], "nlim": case[3],
                                                         "geometry": case[4], "binary": case[0],
                                                         "env": case[5] if len(case) > 5 else {}})
        shutil.rmtree(tmp)
        print(name, "->", out, "%.2f MB" % (os.path.getsize(out) / 1e6))
        if name == "torus2d":  # the horizon-flux history columns (mdot, jet fluxes, Phi) of the same run
            run_ref_hst(case, os.path.join(gold, name + ".hst"), ["parthenon/time/nlim=%d" % case[3]])
    for name, case in HASHED.items():
        if only and name not in only:
            continue
        tmp = tempfile.mkdtemp(prefix="gold_" + name)
        tags = run_ref(case, tmp)
        out = os.path.join(gold, name + ".json")
        with open(os.path.join(HERE, "_ref", "decks", case[1] + ".pin")) as f:
            blocks = deckmod.parse_deck(f.read(), case[2])
        refdump.pack_hashed(tmp, tags, out, extra_meta={"deck_blocks": blocks, "deck": case[1], "overrides": case[2], "nlim": case[3],
                                                        "geometry": case[4], "binary": case[0], "env": case[5]})
        shutil.rmtree(tmp)
        print(name, "->", out, "%.2f MB" % (os.path.getsize(out) / 1e6))
    for name, case in LONG.items():
        if only and name not in only:
            continue
        tmp = tempfile.mkdtemp(prefix="gold_" + name)
        tags = run_ref_long(case, tmp)
        out = os.path.join(gold, name + ".npz")
        with open(os.path.join(HERE, "_ref", "decks", case[1] + ".pin")) as f:
            blocks = deckmod.parse_deck(f.read(), case[2])
        refdump.pack_golden(tmp, tags, out, extra_meta={"deck_blocks": blocks, "deck": case[1], "overrides": case[2], "every": case[3],
                                                         "geometry": case[4], "binary": case[0]})
        shutil.rmtree(tmp)
        print(name, "->", out, "%.2f MB" % (os.path.getsize(out) / 1e6))
        run_ref_hst(case, os.path.join(gold, name + ".hst"))


if __name__ == "__main__":
    main()
