"""TEST INFRASTRUCTURE ONLY — reader for the raw dumps written by oracle/_ref/ref_dump_*
(oracle/ref_harness/ref_dump.cpp) and for the compressed golden fixtures in tests/golden/."""
import json
import os

import numpy as np


def read_dump(dirname, tag):
    with open(os.path.join(dirname, tag + ".json")) as f:
        man = json.load(f)
    raw = np.fromfile(os.path.join(dirname, tag + ".bin"), dtype=np.float64)
    arrays = {}
    for a in man["arrays"]:
        dims = a["dims"]  # [n6,n5,n4,k,j,i]
        arr = raw[a["offset"]:a["offset"] + a["count"]].reshape(dims)
        arrays[(a["gid"], a["container"], a["name"])] = arr
    return man, arrays


def pack_golden(dirname, tags, out_npz, extra_meta=None):
    """Pack dumps into one compressed npz: key '<tag>|<gid>|<container>|<name>'."""
    data, meta = {}, {"tags": tags, "manifests": {}}
    for t in tags:
        man, arrays = read_dump(dirname, t)
        meta["manifests"][t] = {k: v for k, v in man.items() if k != "arrays"}
        for (gid, cont, name), arr in arrays.items():
            data["%s|%d|%s|%s" % (t, gid, cont, name)] = arr
    meta.update(extra_meta or {})
    data["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_npz, **data)


class Golden:
    def __init__(self, path):
        self.z = np.load(path)
        self.meta = json.loads(bytes(self.z["__meta__"]).decode())

    def get(self, tag, gid, cont, name):
        k = "%s|%d|%s|%s" % (tag, gid, cont, name)
        return self.z[k] if k in self.z.files else None

    def manifest(self, tag):
        return self.meta["manifests"][tag]
