#!/usr/bin/env python3
"""Per-phase opcode / stall summary of one kernel from `ncu --page source --csv --print-source sass`.
usage: sass_stalls.py report.ncu-rep kernel-regex [instance]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
starts = [n for n, r in enumerate(rows) if r and r[0] == "Kernel Name"]
start = starts[inst]
print(rows[start][1])
hdr = rows[start + 1]
body = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(dict(zip(hdr, r)))
def I(d, k):
    try:
        return int(d.get(k, "0") or 0)
    except ValueError:
        return 0
def opclass(s):
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]
tot_s = sum(I(d, "# Samples") for d in body)
tot_i = sum(I(d, "Instructions Executed") for d in body)
print("sass", len(body), "samples", tot_s, "warp instr", tot_i)
bars = [n for n, d in enumerate(body) if "BAR" in d["Source"] and I(d, "Instructions Executed") > 0]
cuts = [0] + bars[:1] + [len(body)]
for a, b in zip(cuts[:-1], cuts[1:]):
    seg = body[a:b]
    s = sum(I(d, "# Samples") for d in seg)
    i = sum(I(d, "Instructions Executed") for d in seg)
    c, cs, st = collections.Counter(), collections.Counter(), collections.Counter()
    for d in seg:
        c[opclass(d["Source"])] += I(d, "Instructions Executed")
        cs[opclass(d["Source"])] += I(d, "# Samples")
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                st[k] += I(d, k)
    print("phase [%d:%d) samples %d (%.1f%%) instr %d (%.1f%%)" % (a, b, s, 100.0 * s / max(tot_s, 1), i, 100.0 * i / max(tot_i, 1)))
    print("  ops:", [(k, v, cs[k]) for k, v in c.most_common(16)])
    print("  stalls:", st.most_common(9))
