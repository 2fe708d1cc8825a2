"""Shared scene description for the parity tests: builds the SAME seeded scene for the CPU oracle
(oracle/, the checker) and for the CUDA renderer (tuvok_b200, the product).

The oracle side restates what a Tuvok client + GLGridLeaper would do: convert the volume
(ExtendedOctreeConverter), size and create the pool (GPUMemMan::GetVolumePool / GLVolumePool),
UploadFirstBrick, RecomputeBrickVisibility, then `while (CheckForRedraw()) Paint()` with the hash
table -> UploadBricks paging loop.  View matrices are computed here in numpy (float32) following
Basics/Vectors.h:1250-1277 and GLRenderer.cpp:627, independently of the library's tvk_compute_view.
"""
import math

import numpy as np

from oracle import orc
from tuvok_b200 import synth
from tuvok_b200.tf import TransferFunction1D, TransferFunction2D

F = np.float32


def look_at(eye, at, up):
    eye, at, up = (np.asarray(v, F) for v in (eye, at, up))
    f = at - eye
    s = np.cross(f, up).astype(F)
    u = np.cross(s, f).astype(F)
    f, u, s = (v / F(math.sqrt(float(np.dot(v, v)))) for v in (f, u, s))
    f, u, s = f.astype(F), u.astype(F), s.astype(F)
    m = np.zeros(16, F)
    m[0], m[4], m[8], m[12] = s[0], s[1], s[2], -F(np.dot(s, eye))
    m[1], m[5], m[9], m[13] = u[0], u[1], u[2], -F(np.dot(u, eye))
    m[2], m[6], m[10], m[14] = -f[0], -f[1], -f[2], F(np.dot(f, eye))
    m[15] = 1
    return m.reshape(4, 4)


def perspective(fovy_deg, aspect, n, f):
    fovy = F(fovy_deg) * F(3.14159265358979323846 / 180.0)
    cotan = F(1.0 / math.tan(float(fovy) / 2.0))
    n, f, aspect = F(n), F(f), F(aspect)
    m = np.zeros(16, F)
    m[0] = cotan / aspect
    m[5] = cotan
    m[10] = -(f + n) / (f - n)
    m[14] = F(-2) * (f * n) / (f - n)
    m[11] = -1
    return m.reshape(4, 4)


def lod_factor(fov_deg, height, screen_space_error=1.0):
    return F(2.0) * F(math.tan(float(F(fov_deg) * F((3.1416 / 180.0) / 2.0)))) * F(screen_space_error) / F(height)


class Scene:
    """A fully specified test scene; `.oracle_*` methods run the checker, `.make_renderer()` the product."""

    def __init__(self, kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=36, overlap=2, mode=orc.RM_1DTRANS,
                 lighting=False, width=96, height=96, rotation=None, translation=None, tf1d_size=None,
                 tf_center=0.4, tf_inv_gradient=0.4, tf2d=None, isovalue=None, sample_rate=1.0,
                 max_gpu_mem=1 << 30, seed=0x5EED, scale=(1.0, 1.0, 1.0), pool_size=None, hash_size=None,
                 strategy=orc.BS_SKIP_TWO, clip=((0, 0, 0), (1, 1, 1)), nearest=False, eye=(0, 0, 1.6), fov=50.0,
                 max_grad=0.25, clamp=False, clip_plane_model=None):
        self.kind, self.size, self.dtype = kind, tuple(size), dtype
        self.brick = (brick,) * 3 if np.isscalar(brick) else tuple(brick)
        self.overlap, self.mode, self.lighting = overlap, mode, bool(lighting)
        self.width, self.height = width, height
        self.rotation = np.eye(4, dtype=F) if rotation is None else np.asarray(rotation, F)
        self.translation = np.eye(4, dtype=F) if translation is None else np.asarray(translation, F)
        self.sample_rate, self.seed, self.scale = sample_rate, seed, tuple(scale)
        self.max_gpu_mem, self.strategy, self.clip, self.nearest = max_gpu_mem, strategy, clip, nearest
        self.eye, self.fov, self.max_grad, self.clamp = tuple(eye), fov, max_grad, clamp
        # user clip plane as GLGridLeaper::FillBBoxVBO hands it to Clipper::BoxPlane: (normal, d) in the box's model space
        self.clip_plane_model = None if clip_plane_model is None else tuple(float(F(v)) for v in clip_plane_model)
        self.bits = {orc.U8: 8, orc.U16: 16, orc.F32: 32}[dtype]
        self.range_max = {orc.U8: 255.0, orc.U16: 65535.0, orc.F32: 1.0}[dtype]
        self.isovalue = isovalue if isovalue is not None else self.range_max / 2
        n_tf = tf1d_size or {orc.U8: 256, orc.U16: 4096, orc.F32: 4096}[dtype]
        self.tf1d = TransferFunction1D(n_tf)
        self.tf1d.SetStdFunction(tf_center, tf_inv_gradient)
        # the reference sizes the 2D TF like the 1D one along the value axis (Get2DHistogram()->GetFilledSize())
        self.tf2d = tf2d if tf2d is not None else TransferFunction2D.rectangle(w=n_tf, h=64, x0=0.02, x1=0.9, alpha_max=64)
        self.volume = synth.synth_volume(kind, self.size, dtype, seed)
        self._pool_size = pool_size
        self._hash_size = hash_size
        self._oct = None

    # ------------------------------------------------------------- view
    def matrices(self):
        view = look_at(self.eye, (0, 0, 0), (0, 1, 0))
        proj = perspective(self.fov, F(self.width) / F(self.height), 0.01, 1000.0)
        if getattr(self, "stereo_eye", None) is not None:   # (eye 0 = left / 1 = right, focal length, eye distance)
            e, focal, dist = self.stereo_eye
            vl, vr, pl, pr = orc.stereo_view(self.eye, (0, 0, 0), (0, 1, 0), self.fov, float(F(self.width) / F(self.height)),
                                             0.01, 1000.0, focal, dist)
            view, proj = (vl, pl) if e == 0 else (vr, pr)
        if getattr(self, "ortho_mip", False):
            # m_bOrthoView: an HQ MIP frame's model view is the MIP rotation alone, its projection FLOATMATRIX4::Ortho
            # (GLRaycaster.cpp:486-487, GLRenderer.cpp:1183-1197)
            import tuvok_b200 as tb
            return self.rotation.astype(F), tb.mip_ortho_projection(self.width, self.height)
        mv = ((self.rotation @ self.translation).astype(F) @ view).astype(F)
        return mv, proj

    # ------------------------------------------------------------- oracle side
    @property
    def octree(self):
        if self._oct is None:
            self._oct = orc.Octree(self.volume, self.brick, self.overlap, clamp=self.clamp)
        return self._oct

    def pool_lod_count(self):
        return self.octree.largest_single_brick_lod + 1

    def pool_size(self):
        if self._pool_size is not None:
            return tuple(self._pool_size)
        return orc.pool_size(self.max_gpu_mem, self.bits, 1, self.brick, self.octree.total_bricks)

    def hash_size(self):
        if self._hash_size is not None:
            return self._hash_size
        # > every serialised brick id, so the open-addressing table never collides (deterministic order)
        o = self.octree
        bc = o.brick_count(0)
        return 1 + bc[0] * bc[1] * bc[2] * self.pool_lod_count() + 7

    def visibility_args(self):
        """GLGridLeaper::RecomputeBrickVisibility (GLGridLeaper.cpp:647-687)."""
        rescale = self.range_max / float(self.tf1d.GetSize() - 1)
        if self.mode == orc.RM_1DTRANS:
            lo, hi = self.tf1d.GetNonZeroLimits()
            return (lo * rescale, hi * rescale, 0.0, 0.0)
        if self.mode == orc.RM_2DTRANS:
            x0, x1, y0, y1 = self.tf2d.GetNonZeroLimits()
            return (x0 * rescale, x1 * rescale, float(y0), float(y1))
        return (float(self.isovalue), 0.0, 0.0, 0.0)

    def oracle_pool(self):
        o = self.octree
        pool = orc.Pool(self.pool_size(), self.size, self.brick, self.overlap, self.pool_lod_count(), o.minmax)
        pool.upload_first()
        counts = pool.recompute_visibility(self.mode, *self.visibility_args())
        return pool, counts

    def oracle_params(self, pool):
        p = orc.RenderParams()
        p.width, p.height = self.width, self.height
        mv, proj = self.matrices()
        p.model_view = (orc.C.c_float * 16)(*mv.reshape(-1))
        p.projection = (orc.C.c_float * 16)(*proj.reshape(-1))
        p.vol = orc.u32x3(*self.size)
        p.scale = orc.f32x3(*self.scale)
        p.dtype = self.dtype
        p.pool_size = orc.u32x3(*pool.pool_size)
        p.capacity = orc.u32x3(*pool.capacity)
        p.max_total_brick = orc.u32x3(*self.brick)
        p.max_inner_brick = orc.u32x3(*[b - 2 * self.overlap for b in self.brick])
        p.lod_count = pool.lod_count
        for i, v in enumerate(pool.lod_offsets):
            p.lod_offset[i] = int(v)
        p.meta_dim = orc.u32x3(*pool.meta_dim)
        p.mode, p.lighting = self.mode, int(self.lighting)
        p.sample_rate_modifier = self.sample_rate
        full = {orc.U8: 255.0, orc.U16: 65535.0, orc.F32: 1.0, orc.RGBA8: 255.0}[self.dtype]
        p.trans_scale = F(full / self.range_max)
        p.gradient_scale = F(1.0) if self.max_grad == 0 else F(1.0) / F(self.max_grad)
        p.isoval = {orc.U8: F(self.isovalue / 256.0), orc.U16: F(self.isovalue / 65536.0),
                    orc.F32: F(self.isovalue), orc.RGBA8: F(self.isovalue / 256.0)}[self.dtype]
        p.ambient = orc.f32x4(1, 1, 1, 0.1)
        p.diffuse = orc.f32x4(1, 1, 1, 1)
        p.specular = orc.f32x4(1, 1, 1, 1)
        p.light_dir = orc.f32x3(0, 0, -1)
        p.eye = orc.f32x3(*self.eye)
        p.iso_color = orc.f32x3(0.5, 0.5, 0.5)
        p.lod_factor = lod_factor(self.fov, self.height)
        if self.mode == orc.RM_2DTRANS:
            p.tf_w, p.tf_h = self.tf2d.GetSize()
        else:
            p.tf_w, p.tf_h = self.tf1d.GetSize(), 1
        p.hash_size, p.rehash_count, p.strategy = self.hash_size(), 10, self.strategy
        p.clip_min = orc.f32x3(*self.clip[0])
        p.clip_max = orc.f32x3(*self.clip[1])
        p.nearest = int(self.nearest)
        if self.clip_plane_model is not None:
            p.clip_plane_on = 1
            p.clip_plane = orc.f32x4(*self.clip_plane_model)
        return p

    def tf_bytes(self):
        if self.mode == orc.RM_2DTRANS:
            return np.ascontiguousarray(self.tf2d.GetByteArray())
        return np.ascontiguousarray(self.tf1d.GetByteArray())

    def oracle_render(self, threads=8, max_subframes=64, single_pass=False, warm=None):
        """The reference's convergence loop on the CPU.  Returns dict(image, rgba8, pool, meta, stats, ...).
        warm: the dict of an earlier frame of the same dataset -- the loop continues on its pool / page table (a new
        view of a renderer that has already paged bricks in) instead of starting from an empty pool."""
        o = self.octree
        if warm is None:
            pool, counts = self.oracle_pool()
        else:
            pool, counts = warm["pool"], None
        p = self.oracle_params(pool)
        ps = pool.pool_size
        atlas = np.zeros((ps[2], ps[1], ps[0]), orc.NP_DTYPE[self.dtype]) if warm is None else warm["atlas"]
        b3 = self.brick

        def put(slot_coord, key):
            cap = pool.capacity
            sx, sy, sz = slot_coord % cap[0], (slot_coord // cap[0]) % cap[1], slot_coord // (cap[0] * cap[1])
            b = o.brick(*key)
            atlas[sz * b3[2]:sz * b3[2] + b.shape[0], sy * b3[1]:sy * b3[1] + b.shape[1],
                  sx * b3[0]:sx * b3[0] + b.shape[2]] = b

        last = pool.lod_count - 1
        if warm is None:
            put(pool.capacity[0] * pool.capacity[1] * pool.capacity[2] - 1, (0, 0, 0, last))
        entry, exit_, cov = orc.ray_setup(p)
        ray_start, start_color = entry.copy(), np.zeros_like(entry)
        tf = self.tf_bytes()
        finest = o.brick_count(0)
        total = orc.RenderStats()
        subframes, paged_total, requests = 0, 0, []
        while True:
            hash_table = np.zeros(p.hash_size, np.uint32)
            outs, st = orc.raycast(p, atlas, pool.meta, tf, ray_start, start_color, exit_, cov, hash_table, threads)
            subframes += 1
            total.samples += st.samples
            total.brick_visits += st.brick_visits
            total.rays = st.rays
            ids = orc.hash_decode(hash_table, finest)
            requests.append(ids)
            if len(ids) == 0 or single_pass:
                break
            n, slots = pool.upload_bricks(ids)
            paged_total += n
            for key, s in zip(ids, slots):
                if s != 0xFFFFFFFF:
                    put(int(s), tuple(int(v) for v in key))
            if self.mode == orc.RM_ISOSURFACE:
                ray_start, start_color = outs[2], outs[3]
            else:
                ray_start, start_color = outs[2], outs[1]
            if subframes >= max_subframes or n == 0:
                break
        if self.mode == orc.RM_ISOSURFACE:
            image = orc.iso_compose(p, outs[0], outs[1])
        else:
            image = outs[0]
        image = image.reshape(self.height, self.width, 4)
        return dict(image=image, rgba8=orc.rgba8(image), pool=pool, meta=pool.meta, stats=total, counts=counts,
                    subframes=subframes, paged=paged_total, requests=requests, outs=outs, params=p, atlas=atlas,
                    covered=cov, entry=entry, exit=exit_, tf=tf, last_stats=st)

    def oracle_pipeline(self, n_stages, align=1, threads=8, fresh=False):
        """The depth pipeline on the oracle (orc_render.c `pipeline`): the stages of one frame run one after the other on
        one pool, stage s on the s-th slab from the eye with the two hand-over images of stage s-1 as inputs.
        fresh = False: the stages run on the pool a whole single-renderer frame left behind; fresh = True: on an empty
        pool that every stage pages its slab into, pass by pass -- the paging history of a renderer that only ever ran
        stages (the converged floats depend on that history, like every resumed GridLeaper frame).
        Returns dict(image, rgba8, stages=[dict(outs, samples, box)], resume_pos of the last stage)."""
        from tuvok_b200 import sortlast
        o = self.octree
        warm = self.oracle_render(threads=threads)          # the single-renderer frame (and, unless fresh, its pool)
        pool, atlas = warm["pool"], warm["atlas"]
        b3 = self.brick
        if fresh:
            pool, _ = self.oracle_pool()
            ps = pool.pool_size
            atlas = np.zeros((ps[2], ps[1], ps[0]), orc.NP_DTYPE[self.dtype])

        def put(slot_coord, key):
            cap = pool.capacity
            sx, sy, sz = slot_coord % cap[0], (slot_coord // cap[0]) % cap[1], slot_coord // (cap[0] * cap[1])
            b = o.brick(*key)
            atlas[sz * b3[2]:sz * b3[2] + b.shape[0], sy * b3[1]:sy * b3[1] + b.shape[1], sx * b3[0]:sx * b3[0] + b.shape[2]] = b

        if fresh:
            put(pool.capacity[0] * pool.capacity[1] * pool.capacity[2] - 1, (0, 0, 0, pool.lod_count - 1))
        inner = [b - 2 * self.overlap for b in self.brick]
        finest = [-(-v // i) for v, i in zip(self.size, inner)]
        fl = [np.float32(v) / np.float32(i) for v, i in zip(self.size, inner)]
        fl = [f - f * np.finfo(np.float32).eps if float(int(f)) == float(f) else f for f in fl]
        mx = max(float(v) * float(s_) for v, s_ in zip(self.size, self.scale))
        ext = np.array([float(v) * float(s_) / mx for v, s_ in zip(self.size, self.scale)], np.float64)
        mv, _ = self.matrices()
        eye = sortlast.eye_in_volume(mv, ext)
        axis, boxes = sortlast.depth_slabs(finest, n_stages, (0.5 - eye) * ext, None, align)
        tf = self.tf_bytes()
        fin = o.brick_count(0)
        pos = col = None
        stages = []
        for s_ in range(n_stages):
            cmin, cmax = sortlast.box_to_clip(boxes[s_], finest, fl)
            keep = self.clip
            self.clip = (cmin, cmax)
            p = self.oracle_params(pool)
            self.clip = keep
            p.pipeline = 1
            entry, exit_, cov = orc.ray_setup(p)
            rs = entry.copy() if pos is None else pos
            sc = np.zeros_like(entry) if col is None else col
            for _ in range(32):
                hash_table = np.zeros(p.hash_size, np.uint32)
                outs, st = orc.raycast(p, atlas, pool.meta, tf, rs, sc, exit_, cov, hash_table, threads)
                ids = orc.hash_decode(hash_table, fin)
                if len(ids) == 0:
                    break
                _, slots = pool.upload_bricks(ids)
                for key, sl in zip(ids, slots):
                    if sl != 0xFFFFFFFF:
                        put(int(sl), tuple(int(v) for v in key))
            else:
                raise RuntimeError("stage %d did not converge" % s_)
            stages.append(dict(outs=outs, samples=int(st.samples), box=boxes[s_], covered=cov))
            pos, col = outs[2].copy(), outs[1].copy()
        image = stages[-1]["outs"][0].reshape(self.height, self.width, 4)
        # one whole frame on the single renderer's pool: the sample count to compare with
        p1 = warm["params"]
        _, st1 = orc.raycast(p1, warm["atlas"], warm["pool"].meta, tf, warm["entry"], np.zeros_like(warm["entry"]), warm["exit"],
                             warm["covered"], None, threads)
        return dict(image=image, rgba8=orc.rgba8(image), stages=stages, resume_pos=pos, single=warm, single_samples=int(st1.samples))

    def oracle_classic(self, threads=8):
        """The classic GLRaycaster frame (AbstrRenderer::PlanFrame at the converged LoD + per-brick passes)."""
        o = self.octree
        pool, _ = self.oracle_pool()
        p = self.oracle_params(pool)
        lod = orc.classic_lod(p, self.pool_lod_count())
        bc = o.brick_count(lod)
        first = o.brick_index(0, 0, 0, lod)
        mm = o.minmax[first:first + bc[0] * bc[1] * bc[2]]
        vis = self.visibility_args()
        bricks, n = orc.classic_brick_list(p, lod, self.overlap, mm, vis)
        data = [None if bricks[i].empty else o.brick(*bricks[i].coord, lod) for i in range(n)]
        extra = {}
        cv = getattr(self, "clearview", None)
        if self.mode == orc.RM_ISOSURFACE and cv:
            # AbstrRenderer ClearView state: m_fCVIsovalue, m_vCVColor, m_fCVSize, m_fCVContextScale, m_fCVBorderScale, m_vCVPos
            iso2 = {orc.U8: F(cv["isovalue"] / 256.0), orc.U16: F(cv["isovalue"] / 65536.0), orc.F32: F(cv["isovalue"])}[self.dtype]
            hp, hn, cp, cn, st = orc.classic_cv_render(p, lod, bricks, n, data, float(iso2), threads)
            mv = np.array(list(p.model_view), F).reshape(4, 4)
            q = np.asarray(cv.get("pos", (0.0, 0.0, 0.5, 1.0)), F)
            pick = [F(F(F(q[0] * mv[0, c] + q[1] * mv[1, c]) + q[2] * mv[2, c]) + q[3] * mv[3, c]) for c in range(3)]
            prm = (cv.get("size", 5.5), cv.get("context_scale", 1.0), cv.get("border_scale", 60.0))
            img = orc.cv_compose(p, hp, hn, cp, cn, cv.get("color", (1.0, 0.0, 0.0)), prm, pick)
            extra = dict(hit_pos=hp, hit_normal=hn, cv_pos=cp, cv_normal=cn, cv_isoval=float(iso2), pick=pick, cv_param=prm)
        elif self.mode == orc.RM_ISOSURFACE:
            hp, hn, st = orc.classic_iso_render(p, lod, bricks, n, data, threads)
            img = orc.iso_compose(p, hp, hn)
            extra = dict(hit_pos=hp, hit_normal=hn)
        else:
            img, st = orc.classic_render(p, lod, bricks, n, data, self.tf_bytes(), threads)
        img = img.reshape(self.height, self.width, 4)
        order = np.array([[bricks[i].index, bricks[i].empty] for i in range(n)], np.int64).reshape(-1, 2)
        dist = np.array([bricks[i].distance for i in range(n)], np.float32)
        return dict(image=img, rgba8=orc.rgba8(img), lod=lod, order=order, distance=dist, samples=st.samples, bricks=bricks,
                    n=n, data=data, params=p, **extra)

    def oracle_mip(self, use_mip_lod=True, threads=8):
        """One HQ MIP frame (AbstrRenderer::PlanHQMIPFrame + GLRaycaster::RenderHQMIPInLoop per brick + Transfer-MIP)."""
        o = self.octree
        pool, _ = self.oracle_pool()
        p = self.oracle_params(pool)
        lod = orc.mip_lod(p, self.pool_lod_count(), use_mip_lod)
        bc = o.brick_count(lod)
        first = o.brick_index(0, 0, 0, lod)
        mm = o.minmax[first:first + bc[0] * bc[1] * bc[2]]
        bricks, n = orc.mip_brick_list(p, lod, self.overlap, mm, self.visibility_args())
        data = [None if bricks[i].empty else o.brick(*bricks[i].coord, lod) for i in range(n)]
        img, mx, st = orc.mip_render(p, lod, bricks, n, data, self.tf1d.GetByteArray(), threads)
        img = img.reshape(self.height, self.width, 4)
        order = np.array([[bricks[i].index, bricks[i].empty] for i in range(n)], np.int64).reshape(-1, 2)
        return dict(image=img, rgba8=orc.rgba8(img), max=mx.reshape(self.height, self.width, 2), lod=lod, order=order,
                    samples=st.samples, bricks=bricks, n=n, data=data, params=p)

    # ------------------------------------------------------------- product side
    def make_renderer(self, source="device", device=0):
        """CUDA renderer for the same scene.  source: 'device' = GPU bricker (tvk_build_volume),
        'callback' = host Dataset::GetBrick stand-in fed from the oracle's octree."""
        import tuvok_b200 as tb
        r = tb.CudaGridLeaper(device=device, max_gpu_mem=self.max_gpu_mem, hash_table_size=self.hash_size(),
                              brick_strategy=self.strategy)
        if source == "device":
            r.BuildVolume(self.volume, self.brick, self.overlap, scale=self.scale, clamp_to_edge=self.clamp,
                          max_gradient_magnitude=self.max_grad)
        else:
            o = self.octree
            r.RegisterDataset(self.size, self.brick, self.overlap, self.dtype, o.minmax,
                              lambda x, y, z, lod: o.brick(x, y, z, lod), scale=self.scale,
                              max_gradient_magnitude=self.max_grad)
        r.Set1DTrans(self.tf1d)
        r.Set2DTrans(self.tf2d)
        r.SetRendermode(self.mode)
        r.SetUseLighting(self.lighting)
        r.SetSampleRateModifier(self.sample_rate)
        r.SetIsoValue(self.isovalue)
        r.SetInterpolant(self.nearest)
        r.Resize(self.width, self.height)
        r.SetRotation(self.rotation)
        r.SetTranslation(self.translation)
        r.SetViewParameters(self.fov, 0.01, 1000.0, self.eye, (0, 0, 0), (0, 1, 0))
        r.SetShardBox(*self.clip)
        if self.clip_plane_model is not None:
            r.SetClipPlaneModel(self.clip_plane_model)
            r.EnableClipPlane()
        r.CreateVolumePool(self._pool_size)
        return r


class ColorScene(Scene):
    """A 4-component (RGBA8) volume on the GridLeaper path (GLGridLeaper-Method-*-color.glsl, Compose-Color-FS.glsl): the
    alpha channel is the scene's synthetic field, the colour channels are three other seeded fields.  Bricks, pool and page
    table work as for scalar data with 4 bytes per voxel; visibility uses the alpha channel's min / max."""

    def __init__(self, **kw):
        kw["dtype"] = orc.U8
        super().__init__(**kw)
        self.dtype = orc.RGBA8
        a = self.volume
        chans = [synth.synth_volume(synth.V_NOISE, self.size, orc.U8, self.seed + 11 * (k + 1)) for k in range(3)]
        self.volume = np.ascontiguousarray(np.stack(chans + [a], axis=-1))
        self._oct = None

    @property
    def octree(self):
        if self._oct is None:
            self._oct = orc.ColorOctree(self.volume, self.brick, self.overlap, clamp=self.clamp)
        return self._oct

    def pool_size(self):
        if self._pool_size is not None:
            return tuple(self._pool_size)
        return orc.pool_size(self.max_gpu_mem, 8, 4, self.brick, self.octree.total_bricks)

    def make_renderer(self, source="callback", device=0):
        """source: 'callback' = registered dataset fed from the oracle's colour octree, 'device' = the GPU bricker."""
        return super().make_renderer(source, device)


def image_diff(a8, b8):
    """max |delta| per channel (in 1/255 units) and PSNR (dB) of two RGBA8 images."""
    a = a8.astype(np.int32).reshape(-1)
    b = b8.astype(np.int32).reshape(-1)
    d = np.abs(a - b)
    mse = float(np.mean((a - b).astype(np.float64) ** 2))
    psnr = float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
    return int(d.max()), psnr
