"""Shared helpers for parity tests: load golden fixtures into oracle meshes."""
import os

from oracle import deck, from_dump as fd, refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["shocktube", "linmodes2d", "blast3d", "torus2d"]


def load_case(name):
    g = refdump.Golden(os.path.join(GOLDEN, name + ".npz"))
    p = deck.params_from_deck(g.meta["deck_blocks"])
    return g, p


def getter(g, tag):
    return lambda gid, cont, nm: g.get(tag, gid, cont, nm)


def oracle_mesh(g, p):
    m = fd.mesh_from_dump(p, g.manifest("init"), getter(g, "init"))
    m.has_face_speeds = True  # the reference already called EstimateTimestep at init
    return m


def stage_tags(g):
    return [t for t in g.meta["tags"] if t != "init"]
