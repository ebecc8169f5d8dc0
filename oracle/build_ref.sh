#!/bin/bash
# TEST INFRASTRUCTURE ONLY — builds oracle/_ref/ (git-ignored, travels to the GPU box).
#
# oracle/_ref/ holds binaries made from the UNMODIFIED reference sources:
#   phoebus_mink, phoebus_fmks      the reference's own executable (src/main.cpp), Minkowski / FMKS
#   ref_dump_mink, ref_dump_fmks    oracle/ref_harness/ref_dump.cpp linked against the reference's
#                                   libphoebuslib.a (per-stage raw field dumps = golden vectors)
#
# The reference is a CMake project (Kokkos + Parthenon + singularity-eos); its hot path is not
# compilable from "a few source files". SURVEY.md §8(c) documents the recipe that builds it in
# this container (5 min, 8 cores); this script re-uses that build tree when it exists
# ($REF_BUILD_ROOT, default /tmp/oracle) and re-creates it with the same recipe otherwise.
# Reference sources are never copied into the repository; the copy made under $REF_BUILD_ROOT
# exists only because /root/reference is read-only and two un-checked-out submodules
# (ports-of-call, spiner) must be grafted in from singularity-opac's copies.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="${REF_BUILD_ROOT:-/tmp/oracle}"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"

build_tree() {  # $1 = build dir name, $2 = geometry, $3 = openmp ON/OFF
  local bdir="$ROOT/$1"
  if [ -f "$bdir/src/libphoebuslib.a" ] && [ -f "$bdir/src/phoebus" ]; then return 0; fi
  if [ ! -d "$REF" ]; then echo "no reference at $REF and no build at $bdir" >&2; return 1; fi
  if [ ! -d "$ROOT/src_tree" ]; then
    mkdir -p "$ROOT"
    cp -r "$REF" "$ROOT/src_tree" && chmod -R u+w "$ROOT/src_tree"
    ( cd "$ROOT/src_tree/external/singularity-eos/utils" && rmdir ports-of-call spiner 2>/dev/null || true
      [ -d ports-of-call ] || cp -r ../../singularity-opac/utils/ports-of-call .
      [ -d spiner ] || cp -r ../../singularity-opac/utils/spiner . )
  fi
  mkdir -p "$bdir" && cd "$bdir"
  cmake -G Ninja -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ -DCMAKE_BUILD_TYPE=Release \
    -DPHOEBUS_ENABLE_MPI=OFF -DPHOEBUS_ENABLE_HDF5=OFF -DPHOEBUS_ENABLE_OPENMP=$3 -DKokkos_ENABLE_OPENMP=$3 \
    -DPHOEBUS_GEOMETRY=$2 -DPHOEBUS_CACHE_GEOMETRY=ON -DMAX_NUMBER_CONSERVED_VARS=10 \
    -DPHOEBUS_DO_NU_ELECTRON_ANTI=Off -DPHOEBUS_DO_NU_HEAVY=Off \
    -DPARTHENON_ENABLE_PYTHON_MODULE_CHECK=OFF "$ROOT/src_tree" > "$bdir/configure.log" 2>&1
  ninja phoebus > "$bdir/build.log" 2>&1
}

link_harness() {  # $1 = build dir name, $2 = suffix
  local bdir="$ROOT/$1"
  python3 - "$bdir" "$HERE/ref_harness/ref_dump.cpp" "$OUT/ref_dump_$2" <<'PY'
import json, shlex, subprocess, sys, re
bdir, src, out = sys.argv[1:4]
cc = json.load(open(bdir + '/compile_commands.json'))
cmd = next(c for c in cc if c['file'].endswith('/src/main.cpp'))
args = shlex.split(cmd['command'])
cxx = args[0]
keep = [a for a in args[1:] if a.startswith(('-D', '-I', '-O', '-std', '-W', '-f'))]
ninja = open(bdir + '/build.ninja').read()
m = re.search(r'build src/phoebus:.*?LINK_LIBRARIES = ([^\n]*)', ninja, re.S)
libs = shlex.split(m.group(1))
libs = [(bdir + '/' + l) if (l.endswith('.a') and not l.startswith('/')) else l for l in libs]
full = [cxx] + keep + ['-o', out, src] + libs + libs
print(' '.join(full[:3]), '...', out)
subprocess.check_call(full, cwd=cmd['directory'])
PY
  cp -f "$bdir/src/phoebus" "$OUT/phoebus_$2"
}

MINK=build_mink; FMKS=build_fmks_omp
build_tree $MINK Minkowski OFF
build_tree $FMKS FMKS ON
link_harness $MINK mink
link_harness $FMKS fmks
ls -la "$OUT"
python3 "$HERE/make_decks.py"
