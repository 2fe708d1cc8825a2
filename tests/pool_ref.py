"""Driver for oracle/_ref/ref_pool: the UNMODIFIED reference GLVolumePool (compiled from /root/reference by
oracle/Makefile over a recording null-GL) replayed on a scenario.  Returns what the shader would see (the
R32UI metadata texture = page table, the pool atlas) plus the CPU-side slot table and visibility counts."""
import os
import subprocess

import numpy as np

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_pool")


def have_ref_pool():
    if not os.path.exists(BIN) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(BIN)


class RefPoolResult:
    def __init__(self):
        self.create = None      # dict(total, lods, capacity, offsets, metadim)
        self.events = []        # ("counts", (a,b,c,d)) | ("paged", n) | ("first",) | ("dump", dict)

    def dumps(self):
        return [e[1] for e in self.events if e[0] == "dump"]


def run(tmp_path, octree, vol_size, brick, overlap, dtype, pool_size, ops, max3d=16384, with_voxels=True):
    """ops: list of ("first",) | ("vis1d", a, b) | ("vis2d", a, b, c, d) | ("visiso", v) |
    ("upload", [(x,y,z,lod), ...]) | ("dump",).  The atlas of the LAST dump is returned as a numpy array."""
    o = octree
    bits = {orc.U8: 8, orc.U16: 16, orc.F32: 32}[dtype]
    tmp_path = str(tmp_path)
    keys = list(o.iter_bricks())
    sizes = np.array([o.brick_size(*k) for k in keys], np.uint32)
    sizes.tofile(os.path.join(tmp_path, "sizes.bin"))
    np.ascontiguousarray(o.minmax, np.float64).tofile(os.path.join(tmp_path, "minmax.bin"))
    lines = ["vol %d %d %d" % tuple(vol_size), "brick %d" % brick, "overlap %d" % overlap, "bits %d" % bits,
             "float %d" % int(dtype == orc.F32), "pool %d %d %d" % tuple(pool_size), "max3d %d" % max3d,
             "lods %d" % o.lod_count]
    for lod in range(o.lod_count):
        lines.append("layout %d %d %d %d" % ((lod,) + tuple(o.brick_count(lod))))
    lines.append("sizes %s" % os.path.join(tmp_path, "sizes.bin"))
    lines.append("minmax %s" % os.path.join(tmp_path, "minmax.bin"))
    if with_voxels:
        with open(os.path.join(tmp_path, "bricks.bin"), "wb") as f:
            for k in keys:
                f.write(np.ascontiguousarray(o.brick(*k)).tobytes())
        lines.append("bricks %s" % os.path.join(tmp_path, "bricks.bin"))
    lines.append("create")
    for op in ops:
        if op[0] == "upload":
            ids = np.asarray(op[1], np.int64).reshape(-1, 4)
            lines.append("upload %d %s" % (len(ids), " ".join(str(int(v)) for v in ids.reshape(-1))))
        else:
            lines.append(" ".join([op[0]] + [repr(float(v)) for v in op[1:]]))
    scen = os.path.join(tmp_path, "scenario.txt")
    res = os.path.join(tmp_path, "result.txt")
    atlas_bin = os.path.join(tmp_path, "atlas.bin")
    with open(scen, "w") as f:
        f.write("\n".join(lines) + "\n")
    subprocess.check_call([BIN, scen, res, atlas_bin], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

    out = RefPoolResult()
    with open(res) as f:
        rows = [l.split() for l in f if l.strip()]
    i = 0
    while i < len(rows):
        r = rows[i]
        if r[0] == "create":
            k = r.index("offsets"); m = r.index("metadim")
            out.create = dict(total=int(r[2]), lods=int(r[4]), capacity=tuple(int(v) for v in r[6:9]),
                              offsets=[int(v) for v in r[k + 1:m]], metadim=tuple(int(v) for v in r[m + 1:m + 4]))
        elif r[0] == "counts":
            out.events.append(("counts", tuple(int(v) for v in r[1:5])))
        elif r[0] == "paged":
            out.events.append(("paged", int(r[1])))
        elif r[0] == "first":
            out.events.append(("first",))
        elif r[0] == "meta":
            n = int(r[1])
            d = dict(texture_equals_cpu=bool(int(r[3])), meta=np.array(r[4:4 + n], np.uint32))
            s = rows[i + 1]; ns = int(s[1])
            vals = s[2:2 + 5 * ns]
            d["slot_brick"] = np.array([int(v) for v in vals[0::5]], np.int32)
            d["slot_time"] = np.array([int(v) for v in vals[1::5]], np.uint64)      # UINT64_MAX marks the first brick
            d["slot_pos"] = np.array([[int(vals[j + 2]), int(vals[j + 3]), int(vals[j + 4])] for j in range(0, 5 * ns, 5)],
                                     np.uint32).reshape(ns, 3)
            a = rows[i + 2]
            d["atlas_dim"] = tuple(int(v) for v in a[1:4]); d["atlas_fnv"] = a[5]
            out.events.append(("dump", d))
            i += 2
        i += 1
    dm = out.dumps()
    if dm:
        dim = dm[-1]["atlas_dim"]
        dm[-1]["atlas"] = np.fromfile(atlas_bin, orc.NP_DTYPE[dtype]).reshape(dim[2], dim[1], dim[0])
    return out
