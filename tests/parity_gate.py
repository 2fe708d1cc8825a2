"""Parity gate for FULL-SIZE frames: re-trace a strided subset of the pixels of a frame the CUDA renderer has just
produced with the CPU oracle (oracle/orc_render.c, the checker), on the renderer's OWN page table and on the pool slots
the frame touched, read back from the device -- no second copy of the dataset is needed, so this works on the 2048^3
benchmark volume (28 GB pool, ~3000 touched bricks = 280 MB per view).

Used by bench.py (the `parity` object of the JSON line: max |delta| in 1/255 units, PSNR, whether the float images are
bit-identical) and by tests/test_gpu_render.py.  TEST INFRASTRUCTURE: imports oracle/, never imported by tuvok_b200/.

Reference behaviour being checked: Shaders/GLGridLeaper-blend.glsl:65-228 / -iso.glsl:68-200 through the oracle, tolerance
of BASELINE.json's north_star: RGBA8 max |delta| <= 2/255 per channel, PSNR >= 45 dB.
"""
import ctypes as C
import math

import numpy as np

from oracle import orc

MAX_ABS_255 = 2
MIN_PSNR_DB = 45.0


def oracle_params(r, vol, scale, dtype, brick, overlap, range_max, max_grad):
    """orc.RenderParams of the renderer's CURRENT frame state (tvk_render_params + tvk_info)."""
    r._push_params()
    q, info = r.params, r.info()
    p = orc.RenderParams()
    p.width, p.height = q.width, q.height
    p.model_view = (C.c_float * 16)(*q.model_view)
    p.projection = (C.c_float * 16)(*q.projection)
    p.vol = orc.u32x3(*vol)
    p.scale = orc.f32x3(*scale)
    p.dtype = dtype
    p.pool_size = orc.u32x3(*info.pool_size)
    p.capacity = orc.u32x3(*info.pool_capacity)
    p.max_total_brick = orc.u32x3(brick, brick, brick)
    p.max_inner_brick = orc.u32x3(*[brick - 2 * overlap] * 3)
    p.lod_count = info.pool_lod_count
    for i in range(info.pool_lod_count):
        p.lod_offset[i] = info.lod_offset[i]
    p.meta_dim = orc.u32x3(*info.meta_dim)
    p.mode, p.lighting = q.mode, q.lighting
    p.sample_rate_modifier = q.sample_rate_modifier
    full = {orc.U8: 255.0, orc.U16: 65535.0, orc.F32: 1.0}[dtype]
    p.trans_scale = np.float32(full / range_max)
    p.gradient_scale = np.float32(1.0) if max_grad == 0 else np.float32(1.0) / np.float32(max_grad)
    iso = q.isovalue
    p.isoval = {orc.U8: np.float32(iso / 256.0), orc.U16: np.float32(iso / 65536.0), orc.F32: np.float32(iso)}[dtype]
    p.ambient, p.diffuse, p.specular = orc.f32x4(*q.ambient), orc.f32x4(*q.diffuse), orc.f32x4(*q.specular)
    p.light_dir, p.eye, p.iso_color = orc.f32x3(*q.light_dir), orc.f32x3(*q.eye), orc.f32x3(*q.iso_color)
    p.lod_factor = q.lod_factor
    if q.mode == orc.RM_2DTRANS:
        p.tf_w, p.tf_h = r.tf2d.GetSize()
    else:
        p.tf_w, p.tf_h = r.tf1d.GetSize(), 1
    p.hash_size, p.rehash_count, p.strategy = 0, 10, orc.BS_SKIP_TWO
    p.clip_min, p.clip_max = orc.f32x3(*q.clip_min), orc.f32x3(*q.clip_max)
    p.nearest = q.nearest
    return p


def stride_mask(width, height, stride, covered):
    m = np.zeros((height, width), np.uint8)
    m[stride // 2::stride, stride // 2::stride] = 1
    return (m.reshape(-1) & covered).astype(np.uint8)


def image_metrics(a8, b8):
    a = a8.astype(np.int32).reshape(-1)
    b = b8.astype(np.int32).reshape(-1)
    d = np.abs(a - b)
    mse = float(np.mean((a - b).astype(np.float64) ** 2)) if a.size else 0.0
    psnr = float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
    return int(d.max()) if d.size else 0, psnr


def check_frame(r, vol, dtype, brick, overlap, scale=(1.0, 1.0, 1.0), range_max=None, max_grad=0.25, stride=8, threads=8):
    """Render the renderer's current view to convergence with counters on, read back the page table and the touched pool
    slots, re-trace every `stride`-th pixel with the oracle and compare.  Returns a dict (JSON-serialisable)."""
    range_max = range_max or {orc.U8: 255.0, orc.U16: 65535.0, orc.F32: 1.0}[dtype]
    r.enable_counters(True)
    try:
        st = r.PaintUntilConverged()
        if not st.converged:
            return {"ok": False, "error": "frame did not converge"}
        r._dirty = True                     # a new frame (blank region), not a resumed one: every ray is traced again
        st = r.Paint()                      # one resident pass: its touched-brick bitmap is the frame's
        if not st.converged:
            return {"ok": False, "error": "resident pass reported missing bricks"}
        ids = r.touched_bricks()
    finally:
        r.enable_counters(False)
    gpu_f = r.ReadRGBA32F().reshape(-1, 4)
    gpu_8 = r.ReadRGBA8().reshape(-1, 4)
    meta = r.page_table()
    p = oracle_params(r, vol, scale, dtype, brick, overlap, range_max, max_grad)
    slots = {}
    not_resident = 0
    for i in ids:
        v = int(meta[int(i)])
        if v < orc.BI_FLAG_COUNT:
            not_resident += 1
            continue
        slots[v - orc.BI_FLAG_COUNT] = r.pool_slot(v - orc.BI_FLAG_COUNT, dtype, (brick, brick, brick)).reshape(-1)
    entry, exit_, cov = orc.ray_setup(p)
    mask = stride_mask(p.width, p.height, stride, cov)
    tf = np.ascontiguousarray(r.tf2d.GetByteArray() if p.mode == orc.RM_2DTRANS else r.tf1d.GetByteArray())
    outs, ost, absent = orc.raycast_slots(p, slots, meta, tf, entry, np.zeros_like(entry), exit_, mask, threads)
    img = orc.iso_compose(p, outs[0], outs[1]) if p.mode == orc.RM_ISOSURFACE else outs[0]
    sel = np.flatnonzero(mask)
    ref8 = orc.rgba8(img)[sel]
    mx, psnr = image_metrics(gpu_8[sel], ref8)
    bit = bool(np.array_equal(gpu_f[sel].view(np.uint32), img[sel].view(np.uint32)))
    ok = absent == 0 and not_resident == 0 and mx <= MAX_ABS_255 and psnr >= MIN_PSNR_DB
    return {"ok": bool(ok), "max_abs_255": mx, "psnr_db": ("inf" if math.isinf(psnr) else round(psnr, 2)), "pixels": int(sel.size),
            "float_bit_identical": bit, "max_abs_float": float(np.abs(gpu_f[sel] - img[sel]).max()) if sel.size else 0.0,
            "oracle_samples": int(ost.samples), "slots_read": len(slots), "absent_texel_reads": int(absent),
            "touched_not_resident": int(not_resident), "stride": int(stride)}
