"""ctypes binding of libtvkcuda.so (include/tvk.h).

The library is the product; this module only declares its C ABI.  There is no CPU
fallback: if the shared library is missing or no sm_100 device is present, loading /
`tvk_create` fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVK_LIB: developer override to A/B-test kernel build variants (still an in-tree libtvkcuda build)
LIB_PATH = os.environ.get("TVK_LIB") or os.path.join(_HERE, "libtvkcuda.so")

TVK_MAX_LOD = 16
U8, U16, F32, RGBA8 = 0, 1, 2, 3     # RGBA8: colour volume, bricks are [z, y, x, 4] uint8 (tvk.h TVK_RGBA8)
RM_1DTRANS, RM_2DTRANS, RM_ISOSURFACE = 0, 1, 2
BS_ONLY_NEEDED, BS_REQUEST_ALL, BS_SKIP_ONE_LEVEL, BS_SKIP_TWO_LEVELS = 0, 1, 2, 3
BI_MISSING, BI_CHILD_EMPTY, BI_EMPTY, BI_FLAG_COUNT = 0, 1, 2, 3
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM, ERR_SOURCE = 0, 1, 2, 3, 4, 5

u32x3 = C.c_uint32 * 3
f32x3 = C.c_float * 3
f32x4 = C.c_float * 4
f32x16 = C.c_float * 16


class DeviceCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_gpu_mem", C.c_uint64), ("max_pool_dim", C.c_uint32),
                ("hash_table_size", C.c_uint32), ("rehash_count", C.c_uint32), ("brick_strategy", C.c_int32)]


class VolumeDesc(C.Structure):
    _fields_ = [("domain_size", u32x3), ("scale", f32x3), ("max_brick_size", u32x3), ("overlap", C.c_uint32),
                ("dtype", C.c_int32), ("range_max", C.c_double), ("max_gradient_magnitude", C.c_float),
                ("brick_count", C.c_uint64), ("minmax", C.POINTER(C.c_double))]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("model_view", f32x16), ("projection", f32x16), ("lod_factor", C.c_float),
                ("mode", C.c_int32), ("lighting", C.c_int32), ("sample_rate_modifier", C.c_float),
                ("isovalue", C.c_double),
                ("ambient", f32x4), ("diffuse", f32x4), ("specular", f32x4),
                ("light_dir", f32x3), ("eye", f32x3), ("iso_color", f32x3),
                ("nearest", C.c_int32), ("clip_min", f32x3), ("clip_max", f32x3)]


class FrameStats(C.Structure):
    _fields_ = [("converged", C.c_int32), ("missing_reported", C.c_uint32), ("bricks_paged", C.c_uint32),
                ("samples", C.c_uint64), ("rays", C.c_uint64), ("brick_visits", C.c_uint64),
                ("bricks_touched", C.c_uint64), ("alive_lane_iters", C.c_uint64), ("warp_iters", C.c_uint64),
                ("max_lane_iters", C.c_uint64),
                ("ms_raycast", C.c_float), ("ms_read_htable", C.c_float), ("ms_upload_bricks", C.c_float),
                ("ms_total", C.c_float)]


class Info(C.Structure):
    _fields_ = [("lod_count", C.c_uint32), ("pool_lod_count", C.c_uint32), ("total_bricks", C.c_uint64),
                ("lod_size", (C.c_uint32 * 3) * TVK_MAX_LOD), ("brick_layout", (C.c_uint32 * 3) * TVK_MAX_LOD),
                ("lod_offset", C.c_uint32 * TVK_MAX_LOD),
                ("pool_size", u32x3), ("pool_capacity", u32x3), ("meta_dim", u32x3), ("meta_count", C.c_uint64)]


class OctreeFileInfo(C.Structure):
    _fields_ = [("domain_size", u32x3), ("aspect", C.c_double * 3), ("max_brick_size", u32x3), ("overlap", C.c_uint32),
                ("dtype", C.c_int32), ("version", C.c_uint32), ("lod_count", C.c_uint32), ("brick_count", C.c_uint64),
                ("payload_bytes", C.c_uint64), ("bricks_by_codec", C.c_uint64 * 6)]


class SortLastStats(C.Structure):
    _fields_ = [("frame", FrameStats), ("ms_exchange", C.c_float), ("ms_frame", C.c_float), ("bytes_sent", C.c_uint64),
                ("slice_lo", C.c_uint64), ("slice_hi", C.c_uint64), ("peer_memory", C.c_int32), ("ms_wait_peers", C.c_float)]


COMM_ID_BYTES = 128
SL_OCTANT, SL_SCREEN, SL_PAIRED = 0, 1, 2


class ClassicBrick(C.Structure):
    _fields_ = [("index", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32),
                ("distance", C.c_float), ("empty", C.c_int32)]


BRICK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t)
LOG_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p)

P = C.c_void_p
# every symbol include/tvk.h declares: name -> (restype, argtypes)
class StreamStats(C.Structure):
    _fields_ = [("bricks_uploaded", C.c_uint64), ("h2d_bytes", C.c_uint64), ("upload_ms", C.c_double), ("h2d_ms", C.c_double),
                ("bricks_generated", C.c_uint64), ("host_cache_hits", C.c_uint64), ("host_cache_evictions", C.c_uint64),
                ("source_thread_ms", C.c_double), ("source_threads", C.c_uint32), ("host_cache_pinned", C.c_uint32)]


class QuantizeInfo(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("factor", C.c_double), ("bin_count", C.c_uint64),
                ("changed", C.c_int32), ("hist_set", C.c_int32), ("ms_range", C.c_float), ("ms_map", C.c_float)]


ST_I8, ST_U8, ST_I16, ST_U16, ST_I32, ST_U32, ST_F32, ST_F64 = range(8)

SIGNATURES = {
    "tvk_abi_version": (C.c_uint32, []),
    "tvk_create": (C.c_int, [C.POINTER(DeviceCfg), C.POINTER(P)]),
    "tvk_destroy": (None, [P]),
    "tvk_last_error": (C.c_char_p, [P]),
    "tvk_set_log_callback": (C.c_int, [P, LOG_CB, P]),
    "tvk_set_stream": (C.c_int, [P, P]),
    "tvk_synchronize": (C.c_int, [P]),
    "tvk_enable_counters": (C.c_int, [P, C.c_int]),
    "tvk_set_volume": (C.c_int, [P, C.POINTER(VolumeDesc), BRICK_CB, P]),
    "tvk_build_volume": (C.c_int, [P, P, C.c_int, u32x3, C.c_int, f32x3, u32x3, C.c_uint32, C.c_int,
                                   C.c_double, C.c_float]),
    "tvk_open_octree_file": (C.c_int, [P, C.c_char_p, C.c_uint64, C.c_uint64, P, P, C.c_uint64, C.c_double, C.c_float,
                                       C.POINTER(OctreeFileInfo)]),
    "tvk_open_octree_file_rebricked": (C.c_int, [P, C.c_char_p, C.c_uint64, C.c_uint64, P, u32x3, C.c_double, C.c_float, P]),
    "tvk_open_uvf": (C.c_int, [P, C.c_char_p, C.c_uint64, P, C.c_double, C.c_float, C.POINTER(OctreeFileInfo)]),
    "tvk_uvf_probe_stats": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_double * 2), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_float), C.POINTER(C.c_uint64 * 2)]),
    "tvk_uvf_probe": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "tvk_octree_file_probe": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(OctreeFileInfo)]),
    "tvk_octree_file_read_brick": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_uint32, P, C.c_size_t, u32x3]),
    "tvk_synth_volume": (C.c_int, [P, P, C.c_int, u32x3, C.c_int, C.c_uint32]),
    "tvk_set_procedural_volume": (C.c_int, [P, C.c_int, u32x3, C.c_int, C.c_uint32, f32x3, u32x3, C.c_uint32, C.c_double, C.c_float,
                                            P, C.c_uint64, C.c_uint64, C.c_uint32]),
    "tvk_procedural_brick_count": (C.c_int, [u32x3, u32x3, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "tvk_procedural_brick": (C.c_int, [C.c_int, u32x3, C.c_int, C.c_uint32, u32x3, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, P, C.c_size_t, u32x3]),
    "tvk_procedural_minmax": (C.c_int, [P, C.c_int, u32x3, C.c_int, C.c_uint32, u32x3, C.c_uint32, C.c_uint64, C.c_uint64, P]),
    "tvk_get_stream_stats": (C.c_int, [P, C.POINTER(StreamStats)]),
    "tvk_quantize": (C.c_int, [P, P, C.c_int, C.c_uint64, C.c_int, P, P, C.POINTER(QuantizeInfo)]),
    "tvk_get_info": (C.c_int, [P, C.POINTER(Info)]),
    "tvk_get_minmax": (C.c_int, [P, P, C.c_uint64]),
    "tvk_get_brick_size": (C.c_int, [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32x3]),
    "tvk_read_brick": (C.c_int, [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P, C.c_size_t]),
    "tvk_set_tf1d": (C.c_int, [P, P, C.c_uint32, C.c_uint64, C.c_uint64]),
    "tvk_set_tf2d": (C.c_int, [P, P, C.c_uint32, C.c_uint32, C.c_uint64 * 4]),
    "tvk_create_pool": (C.c_int, [P, P]),
    "tvk_recompute_visibility": (C.c_int, [P, C.c_int, C.c_uint32 * 4]),
    "tvk_upload_bricks": (C.c_int, [P, P, C.c_uint32, P, C.POINTER(C.c_uint32)]),
    "tvk_get_page_table": (C.c_int, [P, P, C.c_uint64]),
    "tvk_get_slots": (C.c_int, [P, P, P, P, C.c_uint32]),
    "tvk_read_pool_slot": (C.c_int, [P, C.c_uint32, P, C.c_size_t]),
    "tvk_get_touched_bricks": (C.c_int, [P, P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "tvk_get_missing_list": (C.c_int, [P, P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "tvk_compute_view": (C.c_int, [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, f32x16, f32x16, f32x3, f32x3,
                                   f32x3, C.c_float, C.c_float, C.c_float, C.c_float]),
    "tvk_default_params": (C.c_int, [C.POINTER(RenderParams), C.c_uint32, C.c_uint32]),
    "tvk_set_params": (C.c_int, [P, C.POINTER(RenderParams)]),
    "tvk_set_pyramid_filter": (C.c_int, [P, C.c_int]),
    "tvk_set_clip_plane": (C.c_int, [P, C.c_int, f32x4]),
    "tvk_clip_plane_to_model": (C.c_int, [f32x4, f32x16, f32x16, f32x4]),
    "tvk_pick": (C.c_int, [P, C.c_uint32, C.c_uint32, f32x3]),
    "tvk_render": (C.c_int, [P, C.POINTER(FrameStats)]),
    "tvk_paint": (C.c_int, [P, C.c_uint32, C.POINTER(FrameStats)]),
    "tvk_raycast_only": (C.c_int, [P]),
    "tvk_read_rgba8": (C.c_int, [P, P, C.c_size_t]),
    "tvk_read_rgba32f": (C.c_int, [P, P, C.c_size_t]),
    "tvk_read_rgba8_async": (C.c_int, [P, P, C.c_size_t]),
    "tvk_read_wait": (C.c_int, [P, C.c_int]),
    "tvk_host_alloc": (C.c_int, [P, C.c_size_t, C.POINTER(P)]),
    "tvk_host_free": (C.c_int, [P, P]),
    "tvk_get_device_image": (C.c_int, [P, C.POINTER(P)]),
    "tvk_read_iso_buffers": (C.c_int, [P, P, P]),
    "tvk_render_classic": (C.c_int, [P, C.POINTER(FrameStats)]),
    "tvk_compute_stereo_view": (C.c_int, [C.POINTER(RenderParams), C.POINTER(RenderParams), C.c_uint32, C.c_uint32, f32x16, f32x16,
                                          f32x3, f32x3, f32x3, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "tvk_stereo_keep_eye": (C.c_int, [P, C.c_int]),
    "tvk_stereo_compose": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_float]),
    "tvk_set_clearview": (C.c_int, [P, C.c_int, C.c_double, f32x3, C.c_float, C.c_float, C.c_float, f32x4]),
    "tvk_read_cv_buffers": (C.c_int, [P, P, P]),
    "tvk_render_stage": (C.c_int, [P, P, P, C.POINTER(FrameStats)]),
    "tvk_get_stage_outputs": (C.c_int, [P, C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "tvk_render_mip": (C.c_int, [P, C.c_int, C.POINTER(FrameStats)]),
    "tvk_read_mip_max": (C.c_int, [P, P]),
    "tvk_get_classic_brick_list": (C.c_int, [P, C.POINTER(C.c_uint32), P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "tvk_composite_over": (C.c_int, [P, P, P, P, C.c_uint64]),
    "tvk_probe_fetch": (C.c_int, [P, C.c_uint32, f32x3, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "tvk_sortlast_unique_id": (C.c_int, [C.c_uint8 * COMM_ID_BYTES]),
    "tvk_sortlast_init": (C.c_int, [P, C.c_uint8 * COMM_ID_BYTES, C.c_int, C.c_int, C.c_int]),
    "tvk_sortlast_shutdown": (C.c_int, [P]),
    "tvk_sortlast_get_block": (C.c_int, [P, f32x3, f32x3, P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tvk_sortlast_get_block_of": (C.c_int, [P, C.c_int, f32x3, f32x3, C.POINTER(C.c_int)]),
    "tvk_sortlast_frame": (C.c_int, [P, C.POINTER(SortLastStats)]),
    "tvk_sortlast_flush": (C.c_int, [P]),
    "tvk_sortlast_read_rgba8": (C.c_int, [P, P, C.c_size_t]),
    "tvk_sortlast_read_rgba8_async": (C.c_int, [P, P, C.c_size_t]),
    "tvk_sortlast_read_slice": (C.c_int, [P, P]),
    "tvk_sortlast_plan": (C.c_int, [u32x3, f32x3, C.c_double * 3, f32x16, C.c_int, C.c_int, P, P, P]),
    "tvk_composite_nway": (C.c_int, [P, P, C.c_int, P, P, C.c_uint64]),
    "tvk_set_store_shard": (C.c_int, [P, f32x3, f32x3]),
    "tvk_quantize_rgba8": (C.c_int, [P, P, P, C.c_uint64]),
}

_LIB = None


class TvkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tvk error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libtvkcuda.so; raises if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C tuvok_b200/csrc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L
