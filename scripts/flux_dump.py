#!/usr/bin/env python3
"""debug: run pb_calculate_fluxes on a golden case's init state, save flux + fsig (npz)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from phoebus_b200 import _lib
from tests.helpers import load_case, load_product_state, product_mesh
name, out = sys.argv[1], sys.argv[2]
g, p = load_case(name)
mesh, fp = product_mesh(g, device=torch.device("cuda:0"))
load_product_state(mesh, g, "init")
L = _lib.lib()
_lib.check(L.pb_calculate_fluxes(C.byref(mesh.pack), C.byref(mesh.cfg), None), "flux")
torch.cuda.synchronize()
np.savez_compressed(out, flux=mesh.flux.cpu().numpy(), fsig=mesh.fsig.cpu().numpy(), prim=mesh.prim.cpu().numpy())
