#!/usr/bin/env python3
"""Summarise ptxas -v logs: registers, spills, shared memory per kernel (build/*.ptxas.log)."""
import glob, re, subprocess, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "build/*.ptxas.log")):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", txt):
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        print("%-60s regs %3s stack %5s spill st/ld %5s/%5s %s" % (name[:60], m.group(5), m.group(2), m.group(3), m.group(4), m.group(6).strip()[:40]))
