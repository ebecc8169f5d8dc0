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


def array_hash(a):
    """the harness's ArrayHash (oracle/ref_harness/ref_dump.cpp): position-weighted sum of the canonicalised bit
    patterns (-0 -> +0, NaN -> one pattern) modulo 2^64"""
    y = np.ascontiguousarray(a, dtype=np.float64).ravel() + 0.0
    b = y.view(np.uint64).copy()
    b[np.isnan(y)] = np.uint64(0x7ff8000000000000)
    i = np.arange(b.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        w = (i * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD1B54A32D192ED03)) | np.uint64(1)
        return int((b * w).sum(dtype=np.uint64))


def pack_hashed(dirname, tags, out_json, extra_meta=None):
    """REF_DUMP_FIELDS=hash dumps: the manifests (with one hash per array) of all tags in one JSON file"""
    meta = {"tags": tags, "manifests": {}}
    for t in tags:
        with open(os.path.join(dirname, t + ".json")) as f:
            meta["manifests"][t] = json.load(f)
    meta.update(extra_meta or {})
    with open(out_json, "w") as f:
        json.dump(meta, f, separators=(",", ":"))


class Hashed:
    def __init__(self, path):
        with open(path) as f:
            self.meta = json.load(f)
        self.h = {}
        for t, man in self.meta["manifests"].items():
            for a in man["arrays"]:
                self.h[(t, a["gid"], a["container"], a["name"])] = (int(a["hash"], 16), a["dims"])

    def manifest(self, tag):
        return self.meta["manifests"][tag]

    def hash(self, tag, gid, cont, name):
        return self.h.get((tag, gid, cont, name), (None, None))[0]


class Golden:
    def __init__(self, path):
        self.z = np.load(path)
        self.meta = json.loads(bytes(self.z["__meta__"]).decode())

    def get(self, tag, gid, cont, name):
        k = "%s|%d|%s|%s" % (tag, gid, cont, name)
        return self.z[k] if k in self.z.files else None

    def manifest(self, tag):
        return self.meta["manifests"][tag]
