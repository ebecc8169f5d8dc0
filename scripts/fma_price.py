#!/usr/bin/env python3
"""The price of bit parity: run every golden fixture through whatever library PHOEBUS_B200_LIB points at
(e.g. a --fmad=true build) and report, per fixture, how many zone-values leave bit equality with the reference's dumps and
how many leave 1e-12 relative, after the first stage and after the last.
usage (GPU box): PHOEBUS_B200_LIB=$PWD/variants/libpb_fma.so python scripts/fma_price.py out.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import from_dump as fd  # noqa: E402
from tests.helpers import CASES, getter, load_case, load_product_state, product_mesh, stage_tags  # noqa: E402


def main(out):
    import torch
    from phoebus_b200.driver import FluidDriver
    dev = torch.device("cuda:0")
    rows = []
    for name in CASES:
        g, p = load_case(name)
        mesh, fp = product_mesh(g, device=dev)
        order = load_product_state(mesh, g, "init")
        drv = FluidDriver(mesh)
        drv.has_face_speeds = True
        tags = stage_tags(g)
        for n, tag in enumerate(tags):
            man = g.manifest(tag)
            drv.stage(man["stage"], man["dt"])
            if n not in (0, len(tags) - 1):
                continue
            torch.cuda.synchronize()
            nval = nbit = ntol = 0
            worst = 0.0
            for mb in man["blocks"]:
                k = order[mb["gid"]]
                ref = fd.ref_state(getter(g, tag), mb["gid"], mesh.shape, man["sc1"])
                for f in ("prim", "cons"):
                    mine = (mesh.cons[man["sc1"]][k] if f == "cons" else mesh.prim[k]).cpu().numpy().reshape(ref[f].shape)
                    r = ref[f]
                    both_nan = np.isnan(mine) & np.isnan(r)
                    scale = np.maximum(np.abs(mine), np.abs(r))
                    rel = np.where(scale > 0, np.abs(mine - r) / np.where(scale > 0, scale, 1.0), 0.0)
                    rel = np.where(both_nan, 0.0, rel)
                    nval += mine.size
                    nbit += int(((mine != r) & ~both_nan).sum())
                    ntol += int((rel > 1e-12).sum())
                    worst = max(worst, float(np.nanmax(rel)))
            rows.append((name, tag, nval, nbit, ntol, worst))
    lib = os.environ.get("PHOEBUS_B200_LIB", "phoebus_b200/libphoebus_b200.so")
    with open(out, "w") as f:
        f.write("# Deviation from the reference's per-stage dumps, library `%s`\n\n" % os.path.relpath(lib, ROOT))
        f.write("prim + cons values of every MeshBlock (ghosts included) after the first and after the last dumped stage.\n\n")
        f.write("| fixture | stage | values | not bit-equal | > 1e-12 relative | worst relative |\n|---|---|---:|---:|---:|---:|\n")
        for r in rows:
            f.write("| %s | %s | %d | %d | %d | %.2e |\n" % r)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1])
