"""phoebus_b200 — B200-native GRMHD fluid update behind the Phoebus/Parthenon task API.

Only what the hot path needs: the C-ABI CUDA library (csrc/, include/phoebus_b200.h) and the
host-side mirror of the reference interface (deck parameters, MeshBlock packs, the RK-stage
driver). See DESIGN.md.
"""
from ._lib import PhoebusB200Error, lib  # noqa: F401
from .params import FluidParams, ParameterInput  # noqa: F401


def __getattr__(name):
    if name == "MeshPack":
        from .mesh import MeshPack
        return MeshPack
    if name == "FluidDriver":
        from .driver import FluidDriver
        return FluidDriver
    raise AttributeError(name)
