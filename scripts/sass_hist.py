#!/usr/bin/env python3
"""opcode histogram of a cuobjdump -sass listing (static counts)"""
import re, collections, sys
c = collections.Counter()
for l in open(sys.argv[1]):
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m:
        c[m.group(2).split('.')[0]] += 1
print(sum(c.values()), c.most_common(30))
