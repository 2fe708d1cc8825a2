"""CudaGridLeaper -- host-side mirror of the reference renderer interface for the hot path.

Method names and argument meaning follow tuvok::AbstrRenderer / GLGridLeaper
(Renderer/AbstrRenderer.h:112-881, Renderer/GL/GLGridLeaper.h) so the parity tests read like
client code of the reference: LoadDataset/RegisterDataset, Set1DTrans/Changed1DTrans,
SetRendermode, SetUseLighting, SetSampleRateModifier, SetIsoValue, Resize, SetRotation /
SetTranslation, Paint, CheckForRedraw, RecomputeBrickVisibility.  Everything below the method
bodies is a call into the C ABI of libtvkcuda.so (include/tvk.h); no pixel, voxel or page-table
entry is computed in Python.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from .tf import TransferFunction1D, TransferFunction2D

RM_1DTRANS, RM_2DTRANS, RM_ISOSURFACE = L.RM_1DTRANS, L.RM_2DTRANS, L.RM_ISOSURFACE
_NP_OF = {L.U8: np.uint8, L.U16: np.uint16, L.F32: np.float32, L.RGBA8: np.uint8}
_DT_OF = {np.dtype(np.uint8): L.U8, np.dtype(np.uint16): L.U16, np.dtype(np.float32): L.F32}
IDENTITY = np.eye(4, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def rotation_x(deg):
    """FLOATMATRIX4::RotationX (Basics/Vectors.h), row-vector convention."""
    a = np.float32(np.deg2rad(deg))
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    m = np.eye(4, dtype=np.float32)
    m[1, 1] = c; m[1, 2] = s; m[2, 1] = -s; m[2, 2] = c
    return m


def rotation_y(deg):
    a = np.float32(np.deg2rad(deg))
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = c; m[0, 2] = -s; m[2, 0] = s; m[2, 2] = c
    return m


def matmul4(a, b):
    """FLOATMATRIX4::operator* (Basics/Vectors.h:936-946): row-vector product, fp32, summed left to right."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    r = a[:, 0:1] * b[0:1, :]
    for k in (1, 2, 3):
        r = (r + a[:, k:k + 1] * b[k:k + 1, :]).astype(np.float32)
    return r


def mip_rotation(window_mode="coronal", angle_deg=0.0, flip=(False, False), region_rotation=None):
    """m_maMIPRotation of GLRenderer::RenderHQMIPPreLoop (GLRenderer.cpp:1256-1285): matRotDir * region.rotation *
    matFlipX * matFlipY * RotationY(angle); sines / cosines in double, then cast (FLOATMATRIX4::RotationX/Y)."""
    def rx(a):
        m = np.eye(4, dtype=np.float32)
        c, s_ = np.float32(np.cos(a)), np.float32(np.sin(a))
        m[1, 1] = c; m[1, 2] = s_; m[2, 1] = -s_; m[2, 2] = c
        return m

    def ry(a):
        m = np.eye(4, dtype=np.float32)
        c, s_ = np.float32(np.cos(a)), np.float32(np.sin(a))
        m[0, 0] = c; m[0, 2] = -s_; m[2, 0] = s_; m[2, 2] = c
        return m

    pi = 3.141592653589793238462643383
    if window_mode == "sagittal":
        rot_dir = matmul4(rx(-pi / 2.0), ry(-pi / 2.0))
    elif window_mode == "axial":
        rot_dir = rx(-pi / 2.0)
    elif window_mode == "coronal":
        rot_dir = np.eye(4, dtype=np.float32)
    else:
        raise ValueError("Invalid windowmode set")
    flip_y = np.diag([-1.0 if flip[0] else 1.0, 1.0, 1.0, 1.0]).astype(np.float32)
    flip_x = np.diag([1.0, -1.0 if flip[1] else 1.0, 1.0, 1.0]).astype(np.float32)
    reg = np.eye(4, dtype=np.float32) if region_rotation is None else np.asarray(region_rotation, np.float32)
    return matmul4(matmul4(matmul4(matmul4(rot_dir, reg), flip_x), flip_y), ry(pi * float(angle_deg) / 180.0))


def mip_ortho_projection(width, height):
    """The parallel projection GLRenderer::Render2DView sets for an HQ MIP frame when m_bOrthoView is on
    (GLRenderer.cpp:1183-1197): FLOATMATRIX4::Ortho (Vectors.h:1279-1284) over +-0.5 * root2scale / aspect, z in [-100, 100];
    aspect ratios in double, the rest in single precision as written there."""
    f = np.float32
    ax, ay = 1.0 / float(width), 1.0 / float(height)
    m = max(ax, ay)
    ax, ay = ax / m, ay / m
    root2 = max(f(1.0), f(f(1.414213) * f(ax / ay))) if ax < ay else f(1.414213)
    l, r = f(f(f(-0.5) * root2) / f(ax)), f(f(f(0.5) * root2) / f(ax))
    b, t = f(f(f(-0.5) * root2) / f(ay)), f(f(f(0.5) * root2) / f(ay))
    n, fa = f(-100.0), f(100.0)
    p = np.zeros((4, 4), np.float32)                   # array[4 * row + col], row-vector convention
    p[0, 0] = f(2.0) / f(r - l); p[3, 0] = f(-f(r + l)) / f(r - l)
    p[1, 1] = f(2.0) / f(t - b); p[3, 1] = f(-f(t + b)) / f(t - b)
    p[2, 2] = f(-f(2.0)) / f(fa - n); p[3, 2] = f(-f(fa + n)) / f(fa - n)
    p[3, 3] = f(1.0)
    return p


def translation(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (x, y, z)
    return m


class CudaGridLeaper:
    """One renderer == one tvk_ctx (MasterController::RequestNewVolumeRenderer + GLGridLeaper)."""

    def __init__(self, device=0, max_gpu_mem=0, max_pool_dim=0, hash_table_size=0, rehash_count=0,
                 brick_strategy=L.BS_SKIP_TWO_LEVELS):
        self._lib = L.lib()
        cfg = L.DeviceCfg(device, max_gpu_mem, max_pool_dim, hash_table_size, rehash_count, brick_strategy)
        h = C.c_void_p()
        rc = self._lib.tvk_create(C.byref(cfg), C.byref(h))
        if rc != L.OK:
            raise L.TvkError(rc, (self._lib.tvk_last_error(None) or b"").decode())
        self._h = h
        self._cb = None
        self._keep = []
        self.params = L.RenderParams()
        self._lib.tvk_default_params(C.byref(self.params), 512, 512)
        # AbstrRenderer view state (AbstrRenderer.cpp:64-69)
        self._eye, self._at, self._up = (0.0, 0.0, 1.6), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)
        self._fov, self._znear, self._zfar = 50.0, 0.01, 1000.0
        self._rotation, self._translation = IDENTITY.copy(), IDENTITY.copy()
        self._user_matrices = None
        self._clip_plane, self._clip_on, self._clip_model = (0.0, 0.0, 1.0, 0.0), False, None   # ExtendedPlane default: z = 0, disabled
        self._dirty = True
        self._converged = False
        self.tf1d = None
        self.tf2d = None
        self.last_stats = L.FrameStats()

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != L.OK:
            raise L.TvkError(rc, (self._lib.tvk_last_error(self._h) or b"").decode())

    def Cleanup(self):
        if getattr(self, "_h", None):
            for k in getattr(self, "_keep", []):
                if isinstance(k, tuple) and k[0] == "pinned":
                    self._lib.tvk_host_free(self._h, C.c_void_p(k[1]))
            self._keep = []
            self._lib.tvk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Cleanup()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Launch all kernels on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self._lib.tvk_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self._lib.tvk_synchronize(self._h))

    def enable_counters(self, on=True):
        self._ck(self._lib.tvk_enable_counters(self._h, int(on)))

    # ------------------------------------------------------------------ dataset
    def OpenOctreeFile(self, path, offset=0, uvf_file_version=5, scale=None, minmax=None, range_max=0.0,
                       max_gradient_magnitude=0.0):
        """UVFDataset stand-in for a TOC block: the ExtendedOctree file at `path` (header at byte `offset`) becomes the
        brick source of the streaming path (parallel pread into pinned staging, zlib / lz4 decoded there).  minmax:
        MaxMinDataBlock table (n, 4) in TOC order, or None to compute it on the device in one streaming pass.
        Returns the parsed header (OctreeFileInfo)."""
        info = L.OctreeFileInfo()
        sc = L.f32x3(*scale) if scale is not None else None
        mm, n = None, 0
        if minmax is not None:
            mm = np.ascontiguousarray(minmax, np.float64).reshape(-1, 4)
            n = mm.shape[0]
        self._ck(self._lib.tvk_open_octree_file(self._h, os.fsencode(path), int(offset), int(uvf_file_version),
                                                C.cast(sc, C.c_void_p) if sc is not None else None,
                                                mm.ctypes.data_as(C.c_void_p) if mm is not None else None, n,
                                                float(range_max), float(max_gradient_magnitude), C.byref(info)))
        self._dirty = True
        return info

    def OpenRebrickedOctreeFile(self, path, target_brick_size, offset=0, uvf_file_version=5, scale=None, range_max=0.0,
                                max_gradient_magnitude=0.0):
        """IOManager::LoadRebrickedDataset stand-in (IO/IOManager.cpp:1281-1320, DynamicBrickingDS with MM_PRECOMPUTE): the
        file's large bricks are re-cut on the device into bricks of `target_brick_size` (ghost included) while it is
        loaded; the result is a resident brick store with its min / max table."""
        info = L.OctreeFileInfo()
        sc = L.f32x3(*scale) if scale is not None else None
        if np.isscalar(target_brick_size):
            target_brick_size = (target_brick_size,) * 3
        self._ck(self._lib.tvk_open_octree_file_rebricked(self._h, os.fsencode(path), int(offset), int(uvf_file_version),
                                                          C.cast(sc, C.c_void_p) if sc is not None else None,
                                                          L.u32x3(*[int(v) for v in target_brick_size]), float(range_max),
                                                          float(max_gradient_magnitude), C.byref(info)))
        self._dirty = True
        return info

    def OpenUVF(self, path, timestep=0, scale=None, range_max=0.0, max_gradient_magnitude=0.0):
        """UVFDataset stand-in: walk the .uvf container and stream the `timestep`-th TOC block (with its MaxMin block)."""
        info = L.OctreeFileInfo()
        sc = L.f32x3(*scale) if scale is not None else None
        self._ck(self._lib.tvk_open_uvf(self._h, os.fsencode(path), int(timestep),
                                        C.cast(sc, C.c_void_p) if sc is not None else None, float(range_max),
                                        float(max_gradient_magnitude), C.byref(info)))
        self._dirty = True
        return info

    def RegisterDataset(self, domain_size, max_brick_size, overlap, dtype, minmax, get_brick, scale=(1, 1, 1),
                        range_max=0.0, max_gradient_magnitude=0.0):
        """LinearIndexDataset stand-in: `get_brick(x, y, z, lod) -> ndarray [sz, sy, sx]` plays
        Dataset::GetBrick, `minmax` is MaxMinForKey for all bricks in TOC order (n, 4)."""
        mm = np.ascontiguousarray(minmax, np.float64).reshape(-1, 4)
        if np.isscalar(max_brick_size):
            max_brick_size = (max_brick_size,) * 3
        d = L.VolumeDesc()
        d.domain_size = L.u32x3(*domain_size)
        d.scale = L.f32x3(*scale)
        d.max_brick_size = L.u32x3(*max_brick_size)
        d.overlap = overlap
        d.dtype = dtype
        d.range_max = range_max if range_max > 0 else {L.U8: 255.0, L.U16: 65535.0, L.F32: 1.0, L.RGBA8: 255.0}[dtype]
        d.max_gradient_magnitude = max_gradient_magnitude
        d.brick_count = mm.shape[0]
        d.minmax = mm.ctypes.data_as(C.POINTER(C.c_double))
        np_dt = _NP_OF[dtype]

        def _cb(user, x, y, z, lod, dst, cap):
            try:
                b = np.ascontiguousarray(get_brick(x, y, z, lod), np_dt)
                if b.nbytes > cap:
                    return 2
                C.memmove(dst, b.ctypes.data, b.nbytes)
                return 0
            except Exception:   # Dataset::GetBrick returning false
                return 1

        self._cb = L.BRICK_CB(_cb)
        self._ck(self._lib.tvk_set_volume(self._h, C.byref(d), self._cb, None))
        self._dirty = True

    def BuildVolume(self, raw, max_brick_size, overlap, scale=(1, 1, 1), clamp_to_edge=False, range_max=0.0,
                    max_gradient_magnitude=0.0, size=None, dtype=None, median=False):
        """Brick a raw volume on the GPU (replaces the offline ExtendedOctreeConverter).  `raw` is a
        numpy array [z, y, x] (host; [z, y, x, 4] uint8 = a colour volume) or an int device pointer with `size=(nx,ny,nz)`
        and `dtype`."""
        if np.isscalar(max_brick_size):
            max_brick_size = (max_brick_size,) * 3
        if isinstance(raw, np.ndarray):
            raw = np.ascontiguousarray(raw)
            size = (raw.shape[2], raw.shape[1], raw.shape[0])
            if raw.ndim == 4:
                if raw.shape[3] != 4 or raw.dtype != np.uint8:
                    raise ValueError("a colour volume is [z, y, x, 4] uint8")
                dtype = L.RGBA8
            else:
                dtype = _DT_OF[raw.dtype]
            ptr, on_dev = _ptr(raw), 0
        else:
            ptr, on_dev = C.c_void_p(int(raw)), 1
        self._ck(self._lib.tvk_set_pyramid_filter(self._h, int(median)))
        self._ck(self._lib.tvk_build_volume(self._h, ptr, on_dev, L.u32x3(*size), dtype, L.f32x3(*scale),
                                            L.u32x3(*max_brick_size), overlap, int(clamp_to_edge), range_max,
                                            max_gradient_magnitude))
        self._dirty = True

    def SetProceduralVolume(self, kind, size, dtype, max_brick_size, overlap, seed=0x5EED, scale=(1, 1, 1), range_max=0.0,
                            max_gradient_magnitude=0.0, minmax=None, host_cache_bytes=0, threads=0):
        """Procedural multi-resolution dataset (tvk_set_procedural_volume): stands where a UVFDataset would; bricks are
        generated on demand by host threads into pinned staging memory and paged in like any streamed brick."""
        if np.isscalar(max_brick_size):
            max_brick_size = (max_brick_size,) * 3
        mm, n = None, 0
        if minmax is not None:
            mm = np.ascontiguousarray(minmax, np.float64).reshape(-1, 4)
            n = mm.shape[0]
            self._keep.append(mm)
        self._ck(self._lib.tvk_set_procedural_volume(self._h, kind, L.u32x3(*size), dtype, seed, L.f32x3(*scale),
                                                     L.u32x3(*max_brick_size), overlap, range_max, max_gradient_magnitude,
                                                     _ptr(mm) if mm is not None else None, n, int(host_cache_bytes), threads))
        self._dirty = True

    def procedural_minmax(self, kind, size, dtype, max_brick_size, overlap, first, count, seed=0x5EED):
        """min/max table slice [first, first + count) of a procedural dataset, evaluated on the device (no dataset needed)."""
        if np.isscalar(max_brick_size):
            max_brick_size = (max_brick_size,) * 3
        out = np.zeros((int(count), 4), np.float64)
        self._ck(self._lib.tvk_procedural_minmax(self._h, kind, L.u32x3(*size), dtype, seed, L.u32x3(*max_brick_size), overlap,
                                                 int(first), int(count), _ptr(out)))
        return out

    def stream_stats(self):
        st = L.StreamStats()
        self._ck(self._lib.tvk_get_stream_stats(self._h, C.byref(st)))
        return st

    def Quantize(self, src_device_ptr, scalar_type, n, out_bits, dst_device_ptr):
        """tvk_quantize: Quantize<T, U> / Process8Bits of the import path on data in device memory.
        -> (histogram uint64[256 | 4096], QuantizeInfo)"""
        bits = 8 if scalar_type in (L.ST_I8, L.ST_U8) else int(out_bits)
        hist = np.zeros(256 if bits == 8 else 4096, np.uint64)
        info = L.QuantizeInfo()
        self._ck(self._lib.tvk_quantize(self._h, C.c_void_p(int(src_device_ptr)), int(scalar_type), int(n), bits,
                                        C.c_void_p(int(dst_device_ptr)) if dst_device_ptr else None, _ptr(hist), C.byref(info)))
        return hist, info

    def synth_volume(self, dst_device_ptr, kind, size, dtype, seed=0x5EED):
        self._ck(self._lib.tvk_synth_volume(self._h, C.c_void_p(int(dst_device_ptr)), kind, L.u32x3(*size), dtype, seed))

    def info(self):
        o = L.Info()
        self._ck(self._lib.tvk_get_info(self._h, C.byref(o)))
        return o

    def minmax(self, n=None):
        n = int(self.info().total_bricks) if n is None else int(n)
        out = np.zeros((n, 4), np.float64)
        self._ck(self._lib.tvk_get_minmax(self._h, _ptr(out), n))
        return out

    def brick_size(self, x, y, z, lod):
        o = L.u32x3()
        self._ck(self._lib.tvk_get_brick_size(self._h, x, y, z, lod, o))
        return tuple(o)

    def brick(self, x, y, z, lod, dtype):
        s = self.brick_size(x, y, z, lod)
        out = np.zeros((s[2], s[1], s[0]) + ((4,) if dtype == L.RGBA8 else ()), _NP_OF[dtype])
        self._ck(self._lib.tvk_read_brick(self._h, x, y, z, lod, _ptr(out), out.nbytes))
        return out

    # ------------------------------------------------------------------ transfer functions
    def Set1DTrans(self, tf):
        """tf: TransferFunction1D, or float rgba array (n, 4)."""
        if not isinstance(tf, TransferFunction1D):
            t = TransferFunction1D(len(tf))
            t.Set(tf)
            tf = t
        self.tf1d = tf
        self.Changed1DTrans()

    def Changed1DTrans(self):
        b = np.ascontiguousarray(self.tf1d.GetByteArray())
        lo, hi = self.tf1d.GetNonZeroLimits()
        self._ck(self._lib.tvk_set_tf1d(self._h, _ptr(b), b.shape[0], lo, hi))
        self._dirty = True

    def Set2DTrans(self, tf):
        if not isinstance(tf, TransferFunction2D):
            tf = TransferFunction2D(tf)
        self.tf2d = tf
        self.Changed2DTrans()

    def Changed2DTrans(self):
        b = np.ascontiguousarray(self.tf2d.GetByteArray())
        nz = (C.c_uint64 * 4)(*self.tf2d.GetNonZeroLimits())
        self._ck(self._lib.tvk_set_tf2d(self._h, _ptr(b), b.shape[1], b.shape[0], nz))
        self._dirty = True

    # ------------------------------------------------------------------ pool
    def CreateVolumePool(self, pool_size=None):
        p = L.u32x3(*pool_size) if pool_size is not None else None
        self._ck(self._lib.tvk_create_pool(self._h, p))
        self._dirty = True

    def RecomputeBrickVisibility(self, force=True):
        self._push_params()
        counts = (C.c_uint32 * 4)()
        self._ck(self._lib.tvk_recompute_visibility(self._h, int(force), counts))
        return tuple(counts)

    def UploadBricks(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32).reshape(-1, 4)
        out = np.zeros(len(ids), np.uint32)
        n = C.c_uint32()
        self._ck(self._lib.tvk_upload_bricks(self._h, _ptr(ids), len(ids), _ptr(out), C.byref(n)))
        return n.value, out

    def page_table(self):
        n = int(self.info().meta_count)
        out = np.zeros(n, np.uint32)
        self._ck(self._lib.tvk_get_page_table(self._h, _ptr(out), n))
        return out

    def slots(self):
        i = self.info()
        n = int(i.pool_capacity[0]) * int(i.pool_capacity[1]) * int(i.pool_capacity[2])
        ids = np.zeros(n, np.int32); t = np.zeros(n, np.uint64); pos = np.zeros((n, 3), np.uint32)
        self._ck(self._lib.tvk_get_slots(self._h, _ptr(ids), _ptr(t), _ptr(pos), n))
        return ids, t, pos

    def pool_slot(self, slot, dtype, brick):
        out = np.zeros((brick[2], brick[1], brick[0]), _NP_OF[dtype])
        self._ck(self._lib.tvk_read_pool_slot(self._h, slot, _ptr(out), out.nbytes))
        return out

    def touched_bricks(self):
        """Page-table indices of the bricks the last counted subframe sampled (enable_counters)."""
        n = C.c_uint64()
        self._ck(self._lib.tvk_get_touched_bricks(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self._lib.tvk_get_touched_bricks(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def missing_list(self):
        n = C.c_uint32()
        self._ck(self._lib.tvk_get_missing_list(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.uint32)
        if n.value:
            self._ck(self._lib.tvk_get_missing_list(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    # ------------------------------------------------------------------ AbstrRenderer state
    def Resize(self, w, h):
        self.params.width, self.params.height = int(w), int(h)
        self._dirty = True

    def SetRendermode(self, mode):
        self.params.mode = int(mode)
        self._dirty = True

    def SetUseLighting(self, on):
        self.params.lighting = int(bool(on))
        self._dirty = True

    def SetSampleRateModifier(self, v):
        self.params.sample_rate_modifier = float(v)
        self._dirty = True

    def SetIsoValue(self, v):
        self.params.isovalue = float(v)
        self._dirty = True

    def SetInterpolant(self, nearest):
        self.params.nearest = int(bool(nearest))
        self._dirty = True

    def SetLightColors(self, ambient, diffuse, specular, direction):
        self.params.ambient = L.f32x4(*ambient)
        self.params.diffuse = L.f32x4(*diffuse)
        self.params.specular = L.f32x4(*specular)
        self.params.light_dir = L.f32x3(*direction)
        self._dirty = True

    def SetIsosurfaceColor(self, rgb):
        self.params.iso_color = L.f32x3(*rgb)
        self._dirty = True

    def SetRotation(self, m):
        self._rotation = np.ascontiguousarray(m, np.float32).reshape(4, 4)
        self._dirty = True

    def SetTranslation(self, m):
        self._translation = np.ascontiguousarray(m, np.float32).reshape(4, 4)
        self._dirty = True

    def SetViewParameters(self, fov, znear, zfar, eye, ref, vup):
        self._fov, self._znear, self._zfar = float(fov), float(znear), float(zfar)
        self._eye, self._at, self._up = tuple(eye), tuple(ref), tuple(vup)
        self._user_matrices = None
        self._dirty = True

    def SetUserMatrices(self, view, projection, lod_factor=None):
        """AbstrRenderer::SetUserMatrices (AbstrRenderer.cpp:1557-1575): bypass BuildLookAt/Perspective."""
        self._user_matrices = (np.ascontiguousarray(view, np.float32).reshape(4, 4),
                               np.ascontiguousarray(projection, np.float32).reshape(4, 4), lod_factor)
        self._dirty = True

    # ------------------------------------------------------------------ stereo (AbstrRenderer.cpp:1383-1412)
    def SetStereo(self, on):
        self._stereo = bool(on)
        self._dirty = True

    def SetStereoMode(self, mode):
        """AbstrRenderer::EStereoMode: 0 SM_RB, 1 SM_SCANLINE, 2 SM_SBS, 3 SM_AF."""
        self._stereo_mode = int(mode)

    def SetStereoEyeDist(self, d):
        self._stereo_eye_dist = float(d)
        self._dirty = True

    def SetStereoFocalLength(self, f):
        self._stereo_focal = float(f)
        self._dirty = True

    def SetStereoEyeSwap(self, swap):
        self._stereo_swap = bool(swap)

    def ToggleStereoFrame(self):
        """AbstrRenderer::ToggleStereoFrame (AbstrRenderer.cpp:1581-1584)."""
        self._alt_frame = 1 - getattr(self, "_alt_frame", 0)

    def stereo_params(self):
        """(left, right) tvk_render_params of GLRenderer::ComputeViewAndProjection in stereo."""
        p = self.params
        left, right = L.RenderParams(), L.RenderParams()
        C.memmove(C.byref(left), C.byref(p), C.sizeof(p))
        rot = L.f32x16(*self._rotation.reshape(-1))
        tra = L.f32x16(*self._translation.reshape(-1))
        self._ck(self._lib.tvk_compute_stereo_view(C.byref(left), C.byref(right), p.width, p.height, rot, tra,
                                                   L.f32x3(*self._eye), L.f32x3(*self._at), L.f32x3(*self._up), self._fov,
                                                   self._znear, self._zfar, 1.0, getattr(self, "_stereo_focal", 1.0),
                                                   getattr(self, "_stereo_eye_dist", 0.02)))
        return left, right

    def PaintStereoEye(self, eye, max_subframes=0):
        """One eye (0 left / 1 right) rendered to convergence and kept as m_pFBO3DImageNext[eye]."""
        prm = self.stereo_params()[eye]
        self._ck(self._lib.tvk_set_params(self._h, C.byref(prm)))
        st = L.FrameStats()
        self._ck(self._lib.tvk_paint(self._h, max_subframes, C.byref(st)))
        self._ck(self._lib.tvk_stereo_keep_eye(self._h, eye))
        self._dirty = True          # the mono parameters are pushed again by the next mono frame
        self.last_stats = st
        self._converged = bool(st.converged)
        return st

    def ComposeStereo(self):
        """GLRenderer::EndFrame's composition of the two kept eyes; the composed frame is what Read* return next."""
        self._ck(self._lib.tvk_stereo_compose(self._h, getattr(self, "_stereo_mode", 0), int(getattr(self, "_stereo_swap", False)),
                                              getattr(self, "_alt_frame", 0), 0.5))

    def PaintStereoUntilConverged(self, max_subframes=0):
        """A finished stereo frame: both eyes rendered to convergence (GLGridLeaper renders the eyes through the same
        pool), then the composition.  Returns the two eyes' frame stats."""
        stats = [self.PaintStereoEye(0, max_subframes), self.PaintStereoEye(1, max_subframes)]
        self.ComposeStereo()
        self._converged = bool(stats[0].converged and stats[1].converged)
        return stats

    # ------------------------------------------------------------------ clip plane (AbstrRenderer.h:215-219)
    def SetClipPlane(self, plane):
        """AbstrRenderer::SetClipPlane: plane = ExtendedPlane::Plane() = (normal, d) in WORLD space; the kept side is
        dot(normal, p) + d <= 0 (GLGridLeaper::FillBBoxVBO -> Clipper::BoxPlane)."""
        self._clip_plane = tuple(float(v) for v in plane)
        self._clip_model = None
        self._dirty = True

    def SetClipPlaneModel(self, plane):
        """The plane already in the box's model space (what a C++ shim computes with the reference's own PLANE / FLOATMATRIX4
        classes); overrides the world-space plane until SetClipPlane is called again."""
        self._clip_model = tuple(float(v) for v in plane)
        self._dirty = True

    def EnableClipPlane(self):
        self._clip_on = True
        self._dirty = True

    def DisableClipPlane(self):
        self._clip_on = False
        self._dirty = True

    def IsClipPlaneEnabled(self):
        return self._clip_on

    def clip_plane_model(self):
        """The plane FillBBoxVBO hands to Clipper::BoxPlane: Plane() * inverse(rotation * translation), normal normalised."""
        if self._clip_model is not None:
            return self._clip_model
        out = L.f32x4()
        rc = self._lib.tvk_clip_plane_to_model(L.f32x4(*self._clip_plane), L.f32x16(*self._rotation.reshape(-1)),
                                               L.f32x16(*self._translation.reshape(-1)), out)
        if rc != L.OK:
            raise L.TvkError(rc, (self._lib.tvk_last_error(None) or b"").decode())
        return tuple(out)

    def Pick(self, mouse_pos):
        """GLRenderer::Pick (GLRenderer.cpp:2856-2872): isosurface hit position under a window position (y from the top);
        raises as the reference throws (wrong mode / no intersection)."""
        out = L.f32x3()
        self._ck(self._lib.tvk_pick(self._h, int(mouse_pos[0]), int(mouse_pos[1]), out))
        return tuple(out)

    def SetShardBox(self, clip_min, clip_max):
        """Sort-last: restrict rays to this rank's convex brick block (normalised volume coords)."""
        self.params.clip_min = L.f32x3(*clip_min)
        self.params.clip_max = L.f32x3(*clip_max)
        self._dirty = True

    def _push_params(self):
        if not self._dirty:
            return
        p = self.params
        w, h = p.width, p.height
        rot = L.f32x16(*self._rotation.reshape(-1))
        tra = L.f32x16(*self._translation.reshape(-1))
        self._ck(self._lib.tvk_compute_view(C.byref(p), w, h, rot, tra, L.f32x3(*self._eye), L.f32x3(*self._at),
                                            L.f32x3(*self._up), self._fov, self._znear, self._zfar, 1.0))
        if self._user_matrices is not None:
            view, proj, lf = self._user_matrices
            mv = (self._rotation @ self._translation @ view).astype(np.float32)
            p.model_view = L.f32x16(*mv.reshape(-1))
            p.projection = L.f32x16(*proj.reshape(-1))
            if lf is not None:
                p.lod_factor = lf
        self._ck(self._lib.tvk_set_params(self._h, C.byref(p)))
        self._ck(self._lib.tvk_set_clip_plane(self._h, int(self._clip_on), L.f32x4(*(self.clip_plane_model() if self._clip_on else (0.0, 0.0, 1.0, 0.0)))))
        self._dirty = False
        self._converged = False

    def _push_ortho_mip(self, mip_rot):
        """m_bOrthoView: projection = Ortho, model view = the MIP rotation alone (no view matrix)."""
        p = self.params
        p.model_view = L.f32x16(*np.asarray(mip_rot, np.float32).reshape(-1))
        p.projection = L.f32x16(*mip_ortho_projection(p.width, p.height).reshape(-1))
        self._ck(self._lib.tvk_set_params(self._h, C.byref(p)))

    # ------------------------------------------------------------------ frame
    def Paint(self):
        """One subframe (GLGridLeaper::Render3DRegion)."""
        self._push_params()
        st = L.FrameStats()
        self._ck(self._lib.tvk_render(self._h, C.byref(st)))
        self.last_stats = st
        self._converged = bool(st.converged)
        return st

    def PaintClassic(self):
        """One converged frame of the classic per-brick raycaster (GLRaycaster): AbstrRenderer::PlanFrame at
        ComputeMinLODForCurrentView + the brick loop of GLRenderer::Render3DView / GLRaycaster::Render3DInLoop."""
        self._push_params()
        st = L.FrameStats()
        self._ck(self._lib.tvk_render_classic(self._h, C.byref(st)))
        self.last_stats = st
        self._converged = True
        return st

    # ------------------------------------------------------------------ ClearView (AbstrRenderer.cpp:1247-1360)
    def _push_cv(self):
        cv = self._cv = getattr(self, "_cv", dict(on=False, iso=0.8, color=(1.0, 0.0, 0.0), size=5.5, context=1.0, border=60.0,
                                                  pos=(0.0, 0.0, 0.5, 1.0)))
        self._ck(self._lib.tvk_set_clearview(self._h, int(cv["on"]), float(cv["iso"]), L.f32x3(*cv["color"]), cv["size"],
                                             cv["context"], cv["border"], L.f32x4(*cv["pos"])))

    def _set_cv(self, **kw):
        self._push_cv()
        self._cv.update(kw)
        self._push_cv()

    def SetCV(self, on):
        self._set_cv(on=bool(on))

    def SetCVIsoValue(self, v):
        self._set_cv(iso=float(v))

    def SetCVColor(self, rgb):
        self._set_cv(color=tuple(float(c) for c in rgb))

    def SetCVSize(self, v):
        self._set_cv(size=float(v))

    def SetCVContextScale(self, v):
        self._set_cv(context=float(v))

    def SetCVBorderScale(self, v):
        self._set_cv(border=float(v))

    def SetCVFocusPos(self, pos4):
        self._set_cv(pos=tuple(float(c) for c in pos4))

    def ReadCVBuffers(self):
        """(cv_pos, cv_normal) of the last ClearView frame, each (h, w, 4) float32 (m_pFBOCVHit)."""
        h, w = self.params.height, self.params.width
        a, b = np.empty((h, w, 4), np.float32), np.empty((h, w, 4), np.float32)
        self._ck(self._lib.tvk_read_cv_buffers(self._h, _ptr(a), _ptr(b)))
        return a, b

    # ------------------------------------------------------------------ depth pipeline (tvk_render_stage)
    def RenderStage(self, in_resume_pos=0, in_resume_color=0):
        """One stage of a depth-pipelined frame on this renderer's slab (SetShardBox); inputs are DEVICE pointers of the
        two hand-over images of the stage in front (0 for the first stage).  Pages in what the stage missed."""
        self._push_params()
        st = L.FrameStats()
        self._ck(self._lib.tvk_render_stage(self._h, C.c_void_p(in_resume_pos or None), C.c_void_p(in_resume_color or None),
                                            C.byref(st)))
        self.last_stats = st
        return st

    def stage_output_ptrs(self):
        """(image, resume_color, resume_pos) device pointers of the last stage."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self._lib.tvk_get_stage_outputs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def SetMIPRotationAngle(self, angle_deg):
        """AbstrRenderer::SetMIPRotationAngle (AbstrRenderer.h:545-547)."""
        self._mip_angle = float(angle_deg)

    def SetMIPLOD(self, on):
        """AbstrRenderer::SetMIPLOD (AbstrRenderer.h:385)."""
        self._mip_lod = bool(on)

    def SetOrthoView(self, on=True):
        """AbstrRenderer::SetOrthoView (AbstrRenderer.cpp:1264-1269): HQ MIP frames use a parallel projection."""
        self._ortho = bool(on)
        self._dirty = True

    def PaintHQMIP(self, window_mode="coronal", flip=(False, False), region_rotation=None):
        """One HQ MIP frame of a 2D window in MIP mode (GLRenderer.cpp:1183-1253): modelView = m_maMIPRotation * view
        (GLRaycaster::RenderHQMIPPreLoop), PlanHQMIPFrame's LoD, per-brick maximum, BE_MAX blending, Transfer-MIP.
        Under SetOrthoView the projection is mip_ortho_projection and the model view the MIP rotation alone
        (GLRaycaster.cpp:486-487)."""
        keep = (self._rotation, self._translation)
        self._rotation = mip_rotation(window_mode, getattr(self, "_mip_angle", 0.0), flip, region_rotation)
        self._translation = np.eye(4, dtype=np.float32)
        self._dirty = True
        try:
            self._push_params()
            if getattr(self, "_ortho", False):
                self._push_ortho_mip(self._rotation)
            st = L.FrameStats()
            self._ck(self._lib.tvk_render_mip(self._h, 1 if getattr(self, "_mip_lod", True) else 0, C.byref(st)))
        finally:
            self._rotation, self._translation = keep
            self._dirty = True
        self.last_stats = st
        self._converged = True
        return st

    def mip_max_image(self):
        """(h, w, 2) float32: blended maximum and coverage of the last MIP frame (parity tap)."""
        out = np.empty((self.params.height, self.params.width, 2), np.float32)
        self._ck(self._lib.tvk_read_mip_max(self._h, _ptr(out)))
        return out

    def classic_brick_list(self):
        """(lod, ndarray [n, 2] of (BrickKey index, bIsEmpty), distances) of the last classic frame, depth sorted
        (AbstrRenderer::m_vCurrentBrickList)."""
        n, lod = C.c_uint32(0), C.c_uint32(0)
        self._ck(self._lib.tvk_get_classic_brick_list(self._h, C.byref(lod), None, 0, C.byref(n)))
        arr = (L.ClassicBrick * max(1, n.value))()
        self._ck(self._lib.tvk_get_classic_brick_list(self._h, C.byref(lod), C.cast(arr, C.c_void_p), n.value, C.byref(n)))
        order = np.array([[arr[i].index, arr[i].empty] for i in range(n.value)], np.int64).reshape(-1, 2)
        dist = np.array([arr[i].distance for i in range(n.value)], np.float32)
        return lod.value, order, dist

    def CheckForRedraw(self):
        return self._dirty or not self._converged

    def PaintUntilConverged(self, max_subframes=0):
        """`while (ren.checkForRedraw()) paint()` done inside the library."""
        self._push_params()
        st = L.FrameStats()
        self._ck(self._lib.tvk_paint(self._h, max_subframes, C.byref(st)))
        self.last_stats = st
        self._converged = bool(st.converged)
        return st

    def probe_fetch(self, steps=256, direction=(0.3, 0.2, 0.93)):
        """Fetch-path ceiling of the traversal kernel on this pool: (samples / s, ms) of loads + filter trees only."""
        self._push_params()
        ms, n = C.c_float(), C.c_uint64()
        self._ck(self._lib.tvk_probe_fetch(self._h, int(steps), L.f32x3(*direction), C.byref(ms), C.byref(n)))
        return n.value / (ms.value * 1e-3), ms.value

    def RaycastOnly(self):
        self._push_params()
        self._ck(self._lib.tvk_raycast_only(self._h))

    def ReadRGBA8(self, out=None):
        w, h = self.params.width, self.params.height
        if out is None:
            out = np.zeros((h, w, 4), np.uint8)
        self._ck(self._lib.tvk_read_rgba8(self._h, _ptr(out), 0))
        return out

    def host_alloc(self, shape, dtype=np.uint8):
        """ndarray over page-locked host memory owned by the library (freed with the renderer)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._ck(self._lib.tvk_host_alloc(self._h, n, C.byref(p)))
        self._keep.append(("pinned", p.value))
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def ReadRGBA8Async(self, out_pinned):
        """Queue the read-back of the last frame into page-locked memory (PBO-style); overlaps the next Paint()."""
        self._ck(self._lib.tvk_read_rgba8_async(self._h, _ptr(out_pinned), 0))

    def WaitRead(self, pending_allowed=0):
        """Block until at most `pending_allowed` queued read-backs are still in flight."""
        self._ck(self._lib.tvk_read_wait(self._h, int(pending_allowed)))

    def ReadRGBA32F(self):
        w, h = self.params.width, self.params.height
        out = np.zeros((h, w, 4), np.float32)
        self._ck(self._lib.tvk_read_rgba32f(self._h, _ptr(out), 0))
        return out

    def ReadIsoBuffers(self):
        w, h = self.params.width, self.params.height
        a = np.zeros((h, w, 4), np.float32); b = np.zeros((h, w, 4), np.float32)
        self._ck(self._lib.tvk_read_iso_buffers(self._h, _ptr(a), _ptr(b)))
        return a, b

    def device_image_ptr(self):
        p = C.c_void_p()
        self._ck(self._lib.tvk_get_device_image(self._h, C.byref(p)))
        return p.value

    # ------------------------------------------------------------------ sort-last inside the library (tvk_sortlast_*)
    @staticmethod
    def sortlast_unique_id():
        """ncclGetUniqueId through the library (rank 0; hand the 128 bytes to the other ranks)."""
        buf = (C.c_uint8 * L.COMM_ID_BYTES)()
        rc = L.lib().tvk_sortlast_unique_id(buf)
        if rc != L.OK:
            raise L.TvkError(rc, (L.lib().tvk_last_error(None) or b"").decode())
        return bytes(buf)

    def SortLastInit(self, comm_id, rank, n_ranks, policy=L.SL_OCTANT):
        buf = (C.c_uint8 * L.COMM_ID_BYTES)(*comm_id)
        self._ck(self._lib.tvk_sortlast_init(self._h, buf, int(rank), int(n_ranks), int(policy)))
        self._sl_n = int(n_ranks) * (2 if int(policy) == L.SL_PAIRED else 1)     # blocks in the plan = entries of `order`
        self._sl_blocks_per_rank = 2 if int(policy) == L.SL_PAIRED else 1

    def SortLastShutdown(self):
        self._ck(self._lib.tvk_sortlast_shutdown(self._h))

    def SortLastBlock(self):
        """(clip_min, clip_max, order, slice_lo, slice_hi) of this rank for the current view."""
        self._push_params()
        a, b = L.f32x3(), L.f32x3()
        order = (C.c_int * self._sl_n)()
        lo, hi = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.tvk_sortlast_get_block(self._h, a, b, C.cast(order, C.c_void_p), C.byref(lo), C.byref(hi)))
        return tuple(a), tuple(b), list(order), lo.value, hi.value

    def SortLastBlockOf(self, which=0):
        """(clip_min, clip_max) of this rank's `which`-th block (paired policy: 0 and 1) for the current view."""
        self._push_params()
        a, b = L.f32x3(), L.f32x3()
        nb = C.c_int()
        self._ck(self._lib.tvk_sortlast_get_block_of(self._h, int(which), a, b, C.byref(nb)))
        return tuple(a), tuple(b)

    def SortLastFrame(self):
        """One subframe on every rank + direct-send compositing + RGBA8 gather on rank 0 (collective)."""
        self._push_params()
        st = L.SortLastStats()
        self._ck(self._lib.tvk_sortlast_frame(self._h, C.byref(st)))
        self.last_stats = st.frame
        self._converged = bool(st.frame.converged)
        return st

    def SortLastFlush(self):
        """The render stream waits for everything queued on the exchange stream (overlapped exchange, tvk_sortlast_flush)."""
        self._ck(self._lib.tvk_sortlast_flush(self._h))

    def SortLastReadRGBA8(self, out=None):
        w, h = self.params.width, self.params.height
        if out is None:
            out = np.zeros((h, w, 4), np.uint8)
        self._ck(self._lib.tvk_sortlast_read_rgba8(self._h, _ptr(out), 0))
        return out

    def SortLastReadRGBA8Async(self, out_pinned):
        self._ck(self._lib.tvk_sortlast_read_rgba8_async(self._h, _ptr(out_pinned), 0))

    def SortLastReadSlice(self, n_pixels):
        out = np.zeros((n_pixels, 4), np.float32)
        self._ck(self._lib.tvk_sortlast_read_slice(self._h, _ptr(out)))
        return out

    def SetStoreShard(self, clip_min, clip_max):
        """Sort-last at the source: BuildVolume keeps only the bricks that touch this box (call before BuildVolume)."""
        self._ck(self._lib.tvk_set_store_shard(self._h, L.f32x3(*clip_min), L.f32x3(*clip_max)))

    def composite_nway(self, slice_ptrs, out_f_ptr, out8_ptr, n_pixels):
        arr = (C.c_void_p * len(slice_ptrs))(*slice_ptrs)
        self._ck(self._lib.tvk_composite_nway(self._h, C.cast(arr, C.c_void_p), len(slice_ptrs), C.c_void_p(out_f_ptr or None),
                                              C.c_void_p(out8_ptr), n_pixels))

    def composite_over(self, front_ptr, back_ptr, out_ptr, n_pixels):
        self._ck(self._lib.tvk_composite_over(self._h, C.c_void_p(front_ptr), C.c_void_p(back_ptr),
                                              C.c_void_p(out_ptr), n_pixels))

    def quantize_rgba8(self, src_ptr, dst_ptr, n_pixels):
        self._ck(self._lib.tvk_quantize_rgba8(self._h, C.c_void_p(src_ptr), C.c_void_p(dst_ptr), n_pixels))
