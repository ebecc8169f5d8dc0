#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.

Derives runnable input decks for the reference binaries in oracle/_ref/ from the
reference's own decks (/root/reference/inputs/*.pin): strips every <parthenon/output*>
block (HDF5 is not available here, SURVEY.md §8c caveat 3) and writes the result to
oracle/_ref/decks/ (git-ignored; travels to the GPU box with the built binaries).
Problem-specific overrides are NOT baked in: they are
passed on the command line as block/key=value by the callers (make_golden.py, bench.py).
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INPUTS = os.environ.get("REFERENCE_DIR", "/root/reference") + "/inputs"
DECKS = ["mhd_shocktube", "linear_modes", "linear_modes_3d", "blast_wave_3d", "torus", "bondi", "p2c2p"]


def strip_outputs(text):
    out, skip = [], False
    for line in text.splitlines():
        m = re.match(r"\s*<([^>]+)>", line)
        if m:
            skip = m.group(1).startswith("parthenon/output")
        if not skip:
            out.append(line)
    return "\n".join(out) + "\n"


def main():
    dst = os.path.join(HERE, "_ref", "decks")
    os.makedirs(dst, exist_ok=True)
    for d in DECKS:
        src = os.path.join(REF_INPUTS, d + ".pin")
        if not os.path.exists(src):
            print("missing", src, file=sys.stderr)
            continue
        with open(src) as f:
            txt = strip_outputs(f.read())
        with open(os.path.join(dst, d + ".pin"), "w") as f:
            f.write(txt)
    print("decks written to", dst)


if __name__ == "__main__":
    main()
