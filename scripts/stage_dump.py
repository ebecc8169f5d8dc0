#!/usr/bin/env python3
"""debug: run the first RK stage of a golden case, save every field (npz)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from phoebus_b200.driver import FluidDriver
from tests.helpers import load_case, load_product_state, product_mesh, stage_tags
name, out = sys.argv[1], sys.argv[2]
g, p = load_case(name)
mesh, fp = product_mesh(g, device=torch.device("cuda:0"))
load_product_state(mesh, g, "init")
drv = FluidDriver(mesh); drv.has_face_speeds = True
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for tag in stage_tags(g)[:nst]:
    man = g.manifest(tag)
    prim_in = mesh.prim.cpu().numpy().copy()
    drv.stage(man["stage"], man["dt"])
torch.cuda.synchronize()
print(stage_tags(g))
np.savez_compressed(out, flux=mesh.flux.cpu().numpy(), fsig=mesh.fsig.cpu().numpy(), prim=mesh.prim.cpu().numpy(),
                    cons=mesh.cons[man["sc1"]].cpu().numpy(), emf=mesh.emf.cpu().numpy(), prim_in=prim_in)
