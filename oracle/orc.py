"""ctypes binding of the CPU ORACLE (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by tuvok_b200/.
See oracle/orc.h for the reference citations of every entry point.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

U8, U16, F32, RGBA8 = 0, 1, 2, 3        # RGBA8: 4 x 8 bit colour volume (render side; bricks are [z, y, x, 4] uint8)
RM_1DTRANS, RM_2DTRANS, RM_ISOSURFACE = 0, 1, 2
BI_MISSING, BI_CHILD_EMPTY, BI_EMPTY, BI_FLAG_COUNT = 0, 1, 2, 3
BS_ONLY_NEEDED, BS_REQUEST_ALL, BS_SKIP_ONE, BS_SKIP_TWO = 0, 1, 2, 3
MAX_LOD = 32
NP_DTYPE = {U8: np.uint8, U16: np.uint16, F32: np.float32, RGBA8: np.dtype((np.uint8, 4))}
DTYPE_OF = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.float32): F32}

u32x3 = C.c_uint32 * 3
f32x3 = C.c_float * 3
f32x4 = C.c_float * 4


class RenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("model_view", C.c_float * 16), ("projection", C.c_float * 16),
        ("vol", u32x3), ("scale", f32x3), ("dtype", C.c_int),
        ("pool_size", u32x3), ("capacity", u32x3),
        ("max_total_brick", u32x3), ("max_inner_brick", u32x3),
        ("lod_count", C.c_uint32), ("lod_offset", C.c_uint32 * MAX_LOD),
        ("meta_dim", u32x3),
        ("mode", C.c_int), ("lighting", C.c_int),
        ("sample_rate_modifier", C.c_float), ("trans_scale", C.c_float),
        ("gradient_scale", C.c_float), ("isoval", C.c_float),
        ("ambient", f32x4), ("diffuse", f32x4), ("specular", f32x4),
        ("light_dir", f32x3), ("eye", f32x3), ("iso_color", f32x3),
        ("lod_factor", C.c_float),
        ("tf_w", C.c_uint32), ("tf_h", C.c_uint32),
        ("hash_size", C.c_uint32), ("rehash_count", C.c_uint32), ("strategy", C.c_int),
        ("clip_min", f32x3), ("clip_max", f32x3), ("nearest", C.c_int), ("pipeline", C.c_int),
        ("clip_plane_on", C.c_int), ("clip_plane", f32x4),
    ]


class ClassicBrick(C.Structure):
    _fields_ = [("center", f32x3), ("ext", f32x3), ("tex_min", f32x3), ("tex_max", f32x3), ("n_vox", u32x3),
                ("coord", u32x3), ("index", C.c_uint32), ("distance", C.c_float), ("empty", C.c_int32)]


class RenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64),
                ("brick_visits", C.c_uint64), ("hash_entries", C.c_uint32)]


def build(force=False):
    """Compile liborc.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".cpp", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-j8", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref():
    """Build oracle/_ref tools from the reference sources, if they are present here."""
    if not os.path.isdir("/root/reference/IO"):
        return False
    subprocess.check_call(["make", "-C", _HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)
    return True


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liborc.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    P = C.c_void_p
    sig = {
        "orc_octree_new": (P, [u32x3, u32x3, C.c_uint32, C.c_int]),
        "orc_octree_free": (None, [P]),
        "orc_octree_build": (C.c_int, [P, P, C.c_int, C.c_int]),
        "orc_octree_build_component": (C.c_int, [P, P, C.c_int, C.c_int, P]),
        "orc_octree_lod_count": (C.c_uint32, [P]),
        "orc_octree_largest_single_brick_lod": (C.c_uint32, [P]),
        "orc_octree_lod_size": (None, [P, C.c_uint32, u32x3]),
        "orc_octree_brick_count": (None, [P, C.c_uint32, u32x3]),
        "orc_octree_total_bricks": (C.c_uint64, [P]),
        "orc_octree_brick_index": (C.c_uint64, [P] + [C.c_uint32] * 4),
        "orc_octree_brick_size": (None, [P] + [C.c_uint32] * 4 + [u32x3]),
        "orc_octree_get_brick": (C.c_int, [P] + [C.c_uint32] * 4 + [P]),
        "orc_octree_minmax": (C.POINTER(C.c_double), [P]),
        "orc_octree_lod_volume": (P, [P, C.c_uint32]),
        "orc_tf1d_std": (None, [P, C.c_uint32, C.c_float, C.c_float]),
        "orc_tf1d_bytes": (None, [P, C.c_uint32, P]),
        "orc_tf1d_nonzero": (None, [P, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "orc_tf2d_nonzero": (None, [P, C.c_uint32, C.c_uint32, C.c_uint64 * 4]),
        "orc_pool_size": (None, [C.c_uint64, C.c_uint64, C.c_uint64, u32x3, C.c_uint64, C.c_uint32, u32x3]),
        "orc_fit_1d_to_3d": (C.c_int, [C.c_uint64, C.c_uint32, u32x3]),
        "orc_pool_new": (P, [u32x3, u32x3, u32x3, C.c_uint32, C.c_uint32, C.c_uint32, P]),
        "orc_pool_free": (None, [P]),
        "orc_pool_total_bricks": (C.c_uint32, [P]),
        "orc_pool_meta_count": (C.c_uint32, [P]),
        "orc_pool_meta": (C.POINTER(C.c_uint32), [P]),
        "orc_pool_meta_dim": (None, [P, u32x3]),
        "orc_pool_capacity": (None, [P, u32x3]),
        "orc_pool_lod_offsets": (None, [P, P]),
        "orc_pool_brick_layout": (None, [P, C.c_uint32, u32x3]),
        "orc_pool_float_layout": (None, [P, C.c_uint32, f32x3]),
        "orc_pool_brick_id": (C.c_uint32, [P] + [C.c_uint32] * 4),
        "orc_pool_vector_id": (None, [P, C.c_uint32, C.c_uint32 * 4]),
        "orc_pool_upload_first": (C.c_uint32, [P]),
        "orc_pool_recompute_visibility": (None, [P, C.c_int] + [C.c_double] * 4 + [C.c_uint32 * 4]),
        "orc_pool_upload_bricks": (C.c_uint32, [P, P, C.c_uint32, P]),
        "orc_pool_slot_count": (C.c_uint32, [P]),
        "orc_pool_slots": (None, [P, P, P, P]),
        "orc_ray_setup": (None, [C.POINTER(RenderParams), P, P, P]),
        "orc_raycast": (None, [C.POINTER(RenderParams)] + [P] * 12 + [C.POINTER(RenderStats), C.c_int]),
        "orc_raycast_slots": (None, [C.POINTER(RenderParams)] + [P] * 12 + [C.POINTER(RenderStats), C.POINTER(C.c_uint64), C.c_int]),
        "orc_iso_compose": (None, [C.POINTER(RenderParams), P, P, P]),
        "orc_iso_compose_color": (None, [C.POINTER(RenderParams), P, P, P]),
        "orc_ray_exit_eye": (None, [P, P]),
        "orc_classic_step_scale": (C.c_float, [P, C.c_uint32]),
        "orc_uniforms": (None, [P, P]),
        "orc_hash_decode": (C.c_uint32, [P, C.c_uint32, u32x3, P]),
        "orc_hash_insert": (C.c_uint32, [P, C.c_uint32, C.c_uint32, u32x3, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_rgba8": (None, [P, C.c_uint64, P]),
        "orc_composite_over": (None, [P, P, C.c_uint64, P]),
        "orc_classic_lod": (C.c_uint32, [C.POINTER(RenderParams), C.c_uint32]),
        "orc_classic_brick_list": (C.c_uint32, [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, P, C.c_double * 4, P,
                                                C.c_uint32]),
        "orc_classic_render": (None, [C.POINTER(RenderParams), C.c_uint32, P, C.c_uint32, P, P, P,
                                      C.POINTER(RenderStats), C.c_int]),
        "orc_stereo_view": (None, [P, P, P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, P, P, P, P]),
        "orc_stereo_compose": (None, [C.c_int, P, P, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, P]),
        "orc_classic_iso_render": (None, [C.POINTER(RenderParams), C.c_uint32, P, C.c_uint32, P, P, P,
                                          C.POINTER(RenderStats), C.c_int]),
        "orc_classic_cv_render": (None, [C.POINTER(RenderParams), C.c_uint32, P, C.c_uint32, P, C.c_float, P, P, P, P,
                                         C.POINTER(RenderStats), C.c_int]),
        "orc_cv_compose": (None, [C.POINTER(RenderParams), P, P, P, P, P, P, P, P]),
        "orc_mip_lod": (C.c_uint32, [C.POINTER(RenderParams), C.c_uint32, C.c_int]),
        "orc_mip_brick_list": (C.c_uint32, [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, P, C.c_double * 4, P,
                                            C.c_uint32]),
        "orc_mip_render": (None, [C.POINTER(RenderParams), C.c_uint32, P, C.c_uint32, P, P, C.c_uint32, P, P,
                                  C.POINTER(RenderStats), C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Octree:
    """ExtendedOctree + converter restatement (orc_octree.c)."""

    def __init__(self, flat, max_brick, overlap, clamp=False, median=False, chan0=None):
        """chan0: the Octree of component 0 when `flat` is a further component of a multi-component volume."""
        flat = np.ascontiguousarray(flat)
        assert flat.ndim == 3, "flat volume is indexed [z, y, x]"
        self.dtype = DTYPE_OF[flat.dtype]
        self.vol = (flat.shape[2], flat.shape[1], flat.shape[0])
        if np.isscalar(max_brick):
            max_brick = (max_brick,) * 3
        self.max_brick = tuple(int(b) for b in max_brick)
        self.overlap = int(overlap)
        L = lib()
        self.h = L.orc_octree_new(u32x3(*self.vol), u32x3(*self.max_brick), self.overlap, self.dtype)
        if not self.h:
            raise ValueError("invalid octree geometry")
        if chan0 is None:
            rc = L.orc_octree_build(self.h, _p(flat), int(clamp), int(median))
        else:
            rc = L.orc_octree_build_component(self.h, _p(flat), int(clamp), int(median), chan0.h)
        if rc != 0:
            raise ValueError("octree build failed (%d)" % rc)
        self.lod_count = L.orc_octree_lod_count(self.h)
        self.total_bricks = L.orc_octree_total_bricks(self.h)
        self.largest_single_brick_lod = L.orc_octree_largest_single_brick_lod(self.h)
        mm = L.orc_octree_minmax(self.h)
        self.minmax = np.ctypeslib.as_array(mm, shape=(self.total_bricks, 4)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_octree_free(self.h)
            self.h = None

    def lod_size(self, lod):
        o = u32x3()
        lib().orc_octree_lod_size(self.h, lod, o)
        return tuple(o)

    def brick_count(self, lod):
        o = u32x3()
        lib().orc_octree_brick_count(self.h, lod, o)
        return tuple(o)

    def brick_index(self, x, y, z, lod):
        return lib().orc_octree_brick_index(self.h, x, y, z, lod)

    def brick_size(self, x, y, z, lod):
        o = u32x3()
        lib().orc_octree_brick_size(self.h, x, y, z, lod, o)
        return tuple(o)

    def brick(self, x, y, z, lod):
        """Brick voxels incl. ghost as array [sz, sy, sx]."""
        s = self.brick_size(x, y, z, lod)
        out = np.empty((s[2], s[1], s[0]), NP_DTYPE[self.dtype])
        lib().orc_octree_get_brick(self.h, x, y, z, lod, _p(out))
        return out

    def lod_volume(self, lod):
        s = self.lod_size(lod)
        n = s[0] * s[1] * s[2]
        ptr = lib().orc_octree_lod_volume(self.h, lod)
        buf = (C.c_char * (n * np.dtype(NP_DTYPE[self.dtype]).itemsize)).from_address(ptr)
        return np.frombuffer(buf, NP_DTYPE[self.dtype]).reshape(s[2], s[1], s[0]).copy()

    def iter_bricks(self, max_lod=None):
        n = self.lod_count if max_lod is None else max_lod + 1
        for lod in range(n):
            bc = self.brick_count(lod)
            for z in range(bc[2]):
                for y in range(bc[1]):
                    for x in range(bc[0]):
                        yield (x, y, z, lod)


class ColorOctree:
    """A 4-component (RGBA, 8 bit) volume as four scalar conversions interleaved: the converter filters the components of a
    voxel alike (same filter, same ghost rule) -- except for its odd-corner rule, which writes component 0 into every
    component (orc_octree.c build_impl) -- and the min / max a renderer sees for colour data is that of the ALPHA component
    (UVFDataset::MaxMinForKey: GetValue(i, 3), uvfDataset.cpp:1144 / :1188).  Same interface as Octree; bricks are
    [sz, sy, sx, 4].  Pinned to the unmodified converter run with four components
    (tests/test_octree_ref.py::test_colour_octree_matches_reference_multi_component_converter)."""

    def __init__(self, rgba, max_brick, overlap, clamp=False):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        assert rgba.ndim == 4 and rgba.shape[3] == 4, "colour volume is indexed [z, y, x, channel]"
        self.ch = [Octree(np.ascontiguousarray(rgba[..., 0]), max_brick, overlap, clamp=clamp)]
        for k in range(1, 4):      # the converter's odd-corner rule couples the components (orc_octree.c build_impl)
            self.ch.append(Octree(np.ascontiguousarray(rgba[..., k]), max_brick, overlap, clamp=clamp, chan0=self.ch[0]))
        a = self.ch[3]
        self.dtype, self.vol, self.max_brick, self.overlap = RGBA8, a.vol, a.max_brick, a.overlap
        self.lod_count, self.total_bricks, self.largest_single_brick_lod = a.lod_count, a.total_bricks, a.largest_single_brick_lod
        self.minmax = a.minmax

    def lod_size(self, lod): return self.ch[3].lod_size(lod)
    def brick_count(self, lod): return self.ch[3].brick_count(lod)
    def brick_index(self, x, y, z, lod): return self.ch[3].brick_index(x, y, z, lod)
    def brick_size(self, x, y, z, lod): return self.ch[3].brick_size(x, y, z, lod)
    def iter_bricks(self, max_lod=None): return self.ch[3].iter_bricks(max_lod)

    def brick(self, x, y, z, lod):
        return np.ascontiguousarray(np.stack([c.brick(x, y, z, lod) for c in self.ch], axis=-1))


class Rebricked:
    """DynamicBrickingDS restatement (IO/DynamicBrickingDS.cpp) over a source `Octree`: the dataset a renderer sees when a
    file converted with large bricks is opened through IOManager::LoadRebrickedDataset (IO/IOManager.cpp:1281-1320).
    Pinned to the unmodified reference class by tests/test_rebrick.py (oracle/_ref/ref_dynbrick) and to the KATs of
    IO/test/rebricking.h.  numpy, small cases only -- test infrastructure."""

    def __init__(self, src, target_brick):
        self.src = src
        g = 2 * src.overlap                                             # ghost(), DynamicBrickingDS.cpp:105-111
        if np.isscalar(target_brick):
            target_brick = (target_brick,) * 3
        # IOManager.cpp:1301-1305: the target is clamped to the source's brick size
        self.max_brick = tuple(min(int(t), int(s)) for t, s in zip(target_brick, src.max_brick))
        self.src_inner = tuple(b - g for b in src.max_brick)            # SourceMaxBrickSize, :348-355
        self.inner = tuple(b - g for b in self.max_brick)                   # BrickSansGhost, :208-215
        for a in range(3):                                              # Rebrick(), :1092-1105
            if self.inner[a] <= 0 or self.src_inner[a] % self.inner[a] != 0:
                raise ValueError("%s dimension is not an integer multiple of original brick size." % "xyz"[a])
        self.ratio = tuple(s // t for s, t in zip(self.src_inner, self.inner))   # TargetBricksPerSource, :249-260
        self.lod_count = src.lod_count                                  # no level below the source's, :1136-1139
        self.ghost = g

    def lod_size(self, lod):
        return self.src.lod_size(lod)

    def brick_count(self, lod):
        # layout(): ceil in SINGLE precision, DynamicBrickingDS.cpp:113-121
        v = self.lod_size(lod)
        return tuple(int(np.ceil(np.float32(v[a]) / np.float32(self.inner[a]))) for a in range(3))

    def total_bricks(self):
        return sum(int(np.prod(self.brick_count(l))) for l in range(self.lod_count))

    def brick_size(self, x, y, z, lod):
        # ComputedTargetBrickSize, :414-434 (the "4 +" is the reference's literal: a ghost of 2 on either side)
        v, bl, idx = self.lod_size(lod), self.brick_count(lod), (x, y, z)
        out = []
        for a in range(3):
            extra = v[a] % self.inner[a]
            out.append(4 + extra if (idx[a] == bl[a] - 1 and extra) else self.max_brick[a])
        return tuple(out)

    def brick(self, x, y, z, lod):
        """CopyBrick (:495-538) of the sub-box at OffsetIntoSource (:305-335) of source brick SourceBrickIndex (:273-303)."""
        idx = (x, y, z)
        sb = tuple(idx[a] // self.ratio[a] for a in range(3))
        off = tuple((idx[a] % self.ratio[a]) * self.inner[a] for a in range(3))
        s = self.src.brick(sb[0], sb[1], sb[2], lod)
        n = self.brick_size(x, y, z, lod)
        return s[off[2]:off[2] + n[2], off[1]:off[1] + n[1], off[0]:off[0] + n[0]].copy()

    def minmax(self, x, y, z, lod):
        """MM_PRECOMPUTE / MM_DYNAMIC: minmax_brick (BMinMax.cpp:8-14) = extrema of every voxel GetBrick returns."""
        b = self.brick(x, y, z, lod)
        return float(b.min()), float(b.max())

    def iter_bricks(self):
        for lod in range(self.lod_count):
            bc = self.brick_count(lod)
            for z in range(bc[2]):
                for y in range(bc[1]):
                    for x in range(bc[0]):
                        yield (x, y, z, lod)


class Pool:
    """GLVolumePool bookkeeping restatement (orc_pool.cpp)."""

    def __init__(self, pool_size, vol, max_brick, overlap, pool_lod_count, minmax4, max_3d_dim=16384):
        L = lib()
        self.pool_size = tuple(int(v) for v in pool_size)
        self.vol = tuple(vol)
        self.max_brick = tuple(max_brick)
        self.overlap = overlap
        self.lod_count = pool_lod_count
        mm = np.ascontiguousarray(minmax4, np.float64)
        self.h = L.orc_pool_new(u32x3(*self.pool_size), u32x3(*self.vol), u32x3(*self.max_brick), overlap,
                                pool_lod_count, max_3d_dim, _p(mm))
        if not self.h:
            raise ValueError("pool creation failed")
        self.total_bricks = L.orc_pool_total_bricks(self.h)
        o = u32x3(); L.orc_pool_capacity(self.h, o); self.capacity = tuple(o)
        o = u32x3(); L.orc_pool_meta_dim(self.h, o); self.meta_dim = tuple(o)
        offs = np.zeros(pool_lod_count, np.uint32)
        L.orc_pool_lod_offsets(self.h, _p(offs))
        self.lod_offsets = offs

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pool_free(self.h)
            self.h = None

    @property
    def meta(self):
        L = lib()
        n = L.orc_pool_meta_count(self.h)
        return np.ctypeslib.as_array(L.orc_pool_meta(self.h), shape=(n,)).copy()

    def brick_layout(self, lod):
        o = u32x3(); lib().orc_pool_brick_layout(self.h, lod, o); return tuple(o)

    def float_layout(self, lod):
        o = f32x3(); lib().orc_pool_float_layout(self.h, lod, o); return tuple(o)

    def brick_id(self, x, y, z, lod):
        return lib().orc_pool_brick_id(self.h, x, y, z, lod)

    def vector_id(self, i):
        o = (C.c_uint32 * 4)(); lib().orc_pool_vector_id(self.h, i, o); return tuple(o)

    def upload_first(self):
        return lib().orc_pool_upload_first(self.h)

    def recompute_visibility(self, mode, a, b=0.0, c=0.0, d=0.0):
        counts = (C.c_uint32 * 4)()
        lib().orc_pool_recompute_visibility(self.h, mode, a, b, c, d, counts)
        return tuple(counts)

    def upload_bricks(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32).reshape(-1, 4)
        out = np.zeros(len(ids), np.uint32)
        n = lib().orc_pool_upload_bricks(self.h, _p(ids), len(ids), _p(out))
        return n, out

    def slots(self):
        L = lib()
        n = L.orc_pool_slot_count(self.h)
        ids = np.zeros(n, np.int32); t = np.zeros(n, np.uint64); pos = np.zeros((n, 3), np.uint32)
        L.orc_pool_slots(self.h, _p(ids), _p(t), _p(pos))
        return ids, t, pos


def pool_size(max_gpu_mem, bit_width, comp_count, max_brick, total_bricks, max_3d_dim=16384):
    o = u32x3()
    lib().orc_pool_size(max_gpu_mem, bit_width, comp_count, u32x3(*max_brick), total_bricks, max_3d_dim, o)
    return tuple(o)


def fit_1d_to_3d(n, max_dim):
    o = u32x3()
    if lib().orc_fit_1d_to_3d(n, max_dim, o) != 0:
        raise ValueError("index exceeds array")
    return tuple(o)


def tf1d_std(n, center=0.5, inv_gradient=0.5):
    rgba = np.zeros((n, 4), np.float32)
    lib().orc_tf1d_std(_p(rgba), n, center, inv_gradient)
    return rgba


def tf1d_bytes(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    out = np.zeros(rgba.shape, np.uint8)
    lib().orc_tf1d_bytes(_p(rgba), rgba.shape[0], _p(out))
    return out


def tf1d_nonzero(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().orc_tf1d_nonzero(_p(rgba), rgba.shape[0], C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def tf2d_nonzero(rgba8):
    rgba8 = np.ascontiguousarray(rgba8, np.uint8)
    h, w = rgba8.shape[:2]
    o = (C.c_uint64 * 4)()
    lib().orc_tf2d_nonzero(_p(rgba8), w, h, o)
    return tuple(o)


def ray_setup(params):
    n = params.width * params.height
    entry = np.zeros((n, 4), np.float32); exit_ = np.zeros((n, 4), np.float32)
    cov = np.zeros(n, np.uint8)
    lib().orc_ray_setup(C.byref(params), _p(entry), _p(exit_), _p(cov))
    return entry, exit_, cov


def ray_exit_eye(params):
    out = np.zeros((params.height * params.width, 3), np.float32)
    lib().orc_ray_exit_eye(C.byref(params), _p(out))
    return out


def uniforms(params):
    o = np.zeros(84, np.float32)
    lib().orc_uniforms(C.byref(params), _p(o))
    return dict(emm=o[:16].copy(), domain_scale=o[16:19].copy(), ambient=o[19:22].copy(), diffuse=o[22:25].copy(),
                specular=o[25:28].copy(), light_dir_m=o[28:31].copy(), eye_m=o[31:34].copy(), lzwse=float(o[34]),
                norm=float(o[35]), model_to_eye=o[36:52].copy(), mv_inv=o[52:68].copy(),
                inv_proj=o[68:84].copy())


def raycast(params, pool_atlas, meta, tf, ray_start, start_color, exit_, covered, hash_table=None, threads=1):
    n = params.width * params.height
    outs = [np.zeros((n, 4), np.float32) for _ in range(4)]
    st = RenderStats()
    meta = np.ascontiguousarray(meta, np.uint32)
    tf = np.ascontiguousarray(tf, np.uint8)
    hp = _p(hash_table) if hash_table is not None else None
    lib().orc_raycast(C.byref(params), _p(pool_atlas), _p(meta), _p(tf), _p(ray_start), _p(start_color),
                      _p(exit_), _p(covered), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]),
                      hp, C.byref(st), threads)
    return outs, st


def raycast_slots(params, slots, meta, tf, ray_start, start_color, exit_, covered, threads=1):
    """orc_raycast on a sparse slot-linear pool.  slots: {linear pool coordinate: ndarray of max_total_brick^3 voxels}.
    Returns (outs, stats, absent_reads)."""
    n = params.width * params.height
    cap = params.capacity
    n_slots = cap[0] * cap[1] * cap[2]
    ptrs = (C.c_void_p * n_slots)()
    keep = {}
    for k, a in slots.items():
        a = np.ascontiguousarray(a, NP_DTYPE[params.dtype])
        assert a.size == params.max_total_brick[0] * params.max_total_brick[1] * params.max_total_brick[2]
        keep[k] = a
        ptrs[int(k)] = a.ctypes.data
    outs = [np.zeros((n, 4), np.float32) for _ in range(4)]
    st = RenderStats()
    absent = C.c_uint64(0)
    meta = np.ascontiguousarray(meta, np.uint32)
    tf = np.ascontiguousarray(tf, np.uint8)
    lib().orc_raycast_slots(C.byref(params), C.cast(ptrs, C.c_void_p), _p(meta), _p(tf), _p(ray_start), _p(start_color),
                            _p(exit_), _p(covered), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]),
                            None, C.byref(st), C.byref(absent), threads)
    return outs, st, int(absent.value)


def iso_compose(params, hit_pos, hit_normal):
    out = np.zeros_like(hit_pos)
    if params.dtype == RGBA8:      # GLRenderer::ComposeSurfaceImage picks Compose-Color-FS for colour data
        lib().orc_iso_compose_color(C.byref(params), _p(hit_pos), _p(hit_normal), _p(out))
    else:
        lib().orc_iso_compose(C.byref(params), _p(hit_pos), _p(hit_normal), _p(out))
    return out


def hash_insert(hash_table, rehash_count, finest_layout, x, y, z, lod):
    """One miss report into the u32 table (in place); returns the rehash count used."""
    return lib().orc_hash_insert(_p(hash_table), len(hash_table), rehash_count, u32x3(*finest_layout), x, y, z, lod)


def hash_decode(hash_table, finest_layout):
    out = np.zeros((len(hash_table), 4), np.uint32)
    n = lib().orc_hash_decode(_p(hash_table), len(hash_table), u32x3(*finest_layout), _p(out))
    return out[:n].copy()


def rgba8(img):
    img = np.ascontiguousarray(img, np.float32)
    out = np.zeros(img.shape, np.uint8)
    lib().orc_rgba8(_p(img), img.size // 4, _p(out))
    return out


def composite_over(front, back):
    out = np.zeros_like(front)
    lib().orc_composite_over(_p(front), _p(back), front.size // 4, _p(out))
    return out


def classic_lod(params, lod_count):
    """AbstrRenderer::ComputeMinLODForCurrentView."""
    return int(lib().orc_classic_lod(C.byref(params), lod_count))


def classic_step_scale(params, lod):
    return float(lib().orc_classic_step_scale(C.byref(params), lod))


def classic_brick_list(params, lod, overlap, minmax_lod, vis):
    """AbstrRenderer::BuildSubFrameBrickList for one LoD -> ctypes array of ClassicBrick (depth sorted)."""
    mm = np.ascontiguousarray(minmax_lod, np.float64)
    v = (C.c_double * 4)(*[float(x) for x in vis])
    n = lib().orc_classic_brick_list(C.byref(params), lod, overlap, _p(mm), v, None, 0)
    arr = (ClassicBrick * max(n, 1))()
    lib().orc_classic_brick_list(C.byref(params), lod, overlap, _p(mm), v, C.cast(arr, C.c_void_p), n)
    return arr, n


def classic_render(params, lod, bricks, n, brick_arrays, tf, threads=1):
    """One classic GLRaycaster frame.  brick_arrays[i]: ndarray of list entry i (None for empty bricks)."""
    keep = [np.ascontiguousarray(a) if a is not None else None for a in brick_arrays]
    ptrs = (C.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in keep])
    out = np.zeros((params.height * params.width, 4), np.float32)
    st = RenderStats()
    tfb = np.ascontiguousarray(tf, np.uint8)
    lib().orc_classic_render(C.byref(params), lod, C.cast(bricks, C.c_void_p), n, C.cast(ptrs, C.c_void_p), _p(tfb), _p(out),
                             C.byref(st), threads)
    return out, st


def mip_lod(params, lod_count, use_mip_lod=True):
    """AbstrRenderer::PlanHQMIPFrame's LoD."""
    return int(lib().orc_mip_lod(C.byref(params), lod_count, 1 if use_mip_lod else 0))


def mip_brick_list(params, lod, overlap, minmax_lod, vis):
    """BuildSubFrameBrickList(true) of a HQ MIP frame (no culling, key order) -> ctypes array of ClassicBrick."""
    mm = np.ascontiguousarray(minmax_lod, np.float64)
    v = (C.c_double * 4)(*[float(x) for x in vis])
    n = lib().orc_mip_brick_list(C.byref(params), lod, overlap, _p(mm), v, None, 0)
    arr = (ClassicBrick * max(n, 1))()
    lib().orc_mip_brick_list(C.byref(params), lod, overlap, _p(mm), v, C.cast(arr, C.c_void_p), n)
    return arr, n


def mip_render(params, lod, bricks, n, brick_arrays, tf1d, threads=1):
    """One HQ MIP frame.  Returns (rgba [h*w, 4], max image [h*w, 2] = (maximum, coverage), stats)."""
    keep = [np.ascontiguousarray(a) if a is not None else None for a in brick_arrays]
    ptrs = (C.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in keep])
    out = np.zeros((params.height * params.width, 4), np.float32)
    mx = np.zeros((params.height * params.width, 2), np.float32)
    st = RenderStats()
    tfb = np.ascontiguousarray(tf1d, np.uint8).reshape(-1, 4)
    lib().orc_mip_render(C.byref(params), lod, C.cast(bricks, C.c_void_p), n, C.cast(ptrs, C.c_void_p), _p(tfb), len(tfb),
                         _p(mx), _p(out), C.byref(st), threads)
    return out, mx, st


SM_RB, SM_SCANLINE, SM_SBS, SM_AF = 0, 1, 2, 3          # AbstrRenderer::EStereoMode


def stereo_view(eye, at, up, fov_deg, aspect, z_near, z_far, focal_length, eye_dist):
    """FLOATMATRIX4::BuildStereoLookAtAndProjection -> (view_left, view_right, proj_left, proj_right), 4x4 float32."""
    e, a, u = (np.ascontiguousarray(v, np.float32) for v in (eye, at, up))
    out = [np.zeros((4, 4), np.float32) for _ in range(4)]
    lib().orc_stereo_view(_p(e), _p(a), _p(u), fov_deg, aspect, z_near, z_far, focal_length, eye_dist, *[_p(o) for o in out])
    return tuple(out)


def stereo_compose(mode, left, right, eye_swap=False, alternating_frame_id=0, split_coord=0.5):
    """GLRenderer::EndFrame's eye composition (Compose-{Anaglyphs,Scanline,SBS,AF}-FS.glsl) of two (h, w, 4) images."""
    l, r = np.ascontiguousarray(left, np.float32), np.ascontiguousarray(right, np.float32)
    assert l.shape == r.shape and l.ndim == 3 and l.shape[2] == 4
    out = np.zeros_like(l)
    lib().orc_stereo_compose(mode, _p(l), _p(r), l.shape[1], l.shape[0], int(bool(eye_swap)), alternating_frame_id,
                             split_coord, _p(out))
    return out


def classic_iso_render(params, lod, bricks, n, brick_arrays, threads=1):
    """Classic isosurface frame -> (hit_pos [h*w, 4], hit_normal [h*w, 4], stats); compose with iso_compose()."""
    keep = [np.ascontiguousarray(a) if a is not None else None for a in brick_arrays]
    ptrs = (C.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in keep])
    hp = np.zeros((params.height * params.width, 4), np.float32)
    hn = np.zeros((params.height * params.width, 4), np.float32)
    st = RenderStats()
    lib().orc_classic_iso_render(C.byref(params), lod, C.cast(bricks, C.c_void_p), n, C.cast(ptrs, C.c_void_p), _p(hp), _p(hn),
                                 C.byref(st), threads)
    return hp, hn, st


def classic_cv_render(params, lod, bricks, n, brick_arrays, cv_isoval, threads=1):
    """Classic isosurface frame with ClearView -> (hit_pos, hit_normal, cv_pos, cv_normal, stats), each [h*w, 4]."""
    keep = [np.ascontiguousarray(a) if a is not None else None for a in brick_arrays]
    ptrs = (C.c_void_p * max(n, 1))(*[(a.ctypes.data if a is not None else None) for a in keep])
    bufs = [np.zeros((params.height * params.width, 4), np.float32) for _ in range(4)]
    st = RenderStats()
    lib().orc_classic_cv_render(C.byref(params), lod, C.cast(bricks, C.c_void_p), n, C.cast(ptrs, C.c_void_p), cv_isoval,
                                _p(bufs[0]), _p(bufs[1]), _p(bufs[2]), _p(bufs[3]), C.byref(st), threads)
    return bufs[0], bufs[1], bufs[2], bufs[3], st


def cv_compose(params, hit_pos, hit_normal, cv_pos, cv_normal, cv_color, cv_param, pick):
    """Compose-CV-FS.glsl -> rgba [h*w, 4]."""
    out = np.zeros_like(hit_pos)
    col, prm, pk = (np.ascontiguousarray(v, np.float32) for v in (cv_color, cv_param, pick))
    lib().orc_cv_compose(C.byref(params), _p(hit_pos), _p(hit_normal), _p(cv_pos), _p(cv_normal), _p(col), _p(prm), _p(pk), _p(out))
    return out


# ---- value quantiser (IO/Quantize.h, AbstrConverter::Process8Bits) ------------------------------------------------
ST_I8, ST_U8, ST_I16, ST_U16, ST_I32, ST_U32, ST_F32, ST_F64 = range(8)
ST_OF = {np.dtype(np.int8): ST_I8, np.dtype(np.uint8): ST_U8, np.dtype(np.int16): ST_I16, np.dtype(np.uint16): ST_U16,
         np.dtype(np.int32): ST_I32, np.dtype(np.uint32): ST_U32, np.dtype(np.float32): ST_F32, np.dtype(np.float64): ST_F64}


class QuantizeInfo(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("factor", C.c_double), ("bin_count", C.c_uint64),
                ("changed", C.c_int32), ("hist_set", C.c_int32)]


def quantize(src, out_bits=16):
    """-> (dst or None when the input is used as is, histogram uint64[256 | 4096], QuantizeInfo)"""
    src = np.ascontiguousarray(src).reshape(-1)
    if src.dtype in (np.dtype(np.int8), np.dtype(np.uint8)):
        out_bits = 8
    dst = np.zeros(src.size, np.uint8 if out_bits == 8 else np.uint16)
    hist = np.zeros(256 if out_bits == 8 else 4096, np.uint64)
    info = QuantizeInfo()
    rc = lib().orc_quantize(_p(src), ST_OF[src.dtype], C.c_uint64(src.size), out_bits, _p(dst), _p(hist), C.byref(info))
    if rc:
        raise ValueError("orc_quantize refused the input")
    return (dst if info.changed else None), hist, info
