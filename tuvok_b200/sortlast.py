"""Sort-last multi-GPU rendering: brick blocks sharded across ranks, binary-swap compositing.

New design (the reference is single-GPU): the volume is cut into G = 2^k convex
axis-aligned blocks by recursive bisection along the longest axis (in finest-level brick units, so
block faces coincide with brick faces).  Rank g walks every ray that meets its block from the ray's true
entry point (so its sample positions are those of the single-GPU ray; bricks of other ranks are stepped
through without sampling or paging) and takes only the samples inside its block, into a
full-resolution premultiplied RGBA32F image; log2(G) binary-swap rounds then exchange half of the
remaining image region with partner `rank ^ 2^r` (torch.distributed P2P = ncclSend/ncclRecv in one
group) and blend with the library's over-operator kernel in the front-to-back order given by the
camera side of the split plane.  After the last round every rank owns 1/G of the final image.

The exchange schedule is pure host logic and backend-agnostic (tested with gloo on CPU); the
blend is always the CUDA kernel `tvk_composite_over` on GPU ranks.
"""
import math

import numpy as np


def split_axes(view_dir, n_ranks, policy="screen"):
    """Axis to bisect at every level for a camera looking along `view_dir` (volume space).

    policy "screen": the axes most PERPENDICULAR to the view first (k = 3: 4 slabs along the best axis x 2 along
    the second best), so the blocks lie side by side on screen.  A ray crosses one or two blocks, early ray
    termination works as on one GPU and no rank renders samples a front rank already made invisible -- but every
    rank still traverses full-length rays, and the launch is bound by the latency of its longest rays
    (DESIGN.md section 5), so this does not scale.
    policy "depth": every cut along the axis most PARALLEL to the view: the blocks are slabs behind each other, each
    rank sees all pixels but only 1/G of every ray, which divides the critical path by G.  Costs the samples that a
    single GPU would have skipped behind an early-terminated front slab.
    policy "octant": longest axis of the block at every level (view independent).
    policy "depthw" / "octantw": as "depth" / "octant", but every cut is placed at the weighted median of the block
    (SortLastRenderer.set_weights: non-empty finest-level bricks), so the slabs hold equal amounts of visible data
    instead of equal numbers of bricks."""
    k = int(round(math.log2(n_ranks)))
    if policy in ("octant", "octantw"):
        return None
    order = sorted(range(3), key=lambda i: (abs(float(view_dir[i])), i))
    if policy in ("depth", "depthw"):
        return [order[2]] * k
    if policy == "depth2":      # two cuts in depth, the third across the screen
        return [order[2], order[2], order[0]][:k]
    return [order[0], order[1], order[0]][:k]


def shard_boxes(finest_layout, n_ranks, axes=None, weights=None):
    """Recursive bisection of the finest brick grid into n_ranks (power of two) blocks.  axes[level] = axis cut at
    that level (split_axes); None or an axis with fewer than 2 bricks left: the longest axis of the block.
    weights: optional array [x, y, z] over the finest brick grid; a cut then halves the block's weight (weighted
    median along the cut axis) instead of its brick count.
    Returns (boxes, splits): boxes[g] = (lo[3], hi[3]) in brick units; splits = list, one entry per
    bisection LEVEL l (0 = first cut) of dict{prefix -> (axis, cut_brick)} keyed by the rank's top-l bits."""
    k = int(round(math.log2(n_ranks)))
    if 1 << k != n_ranks:
        raise ValueError("sort-last sharding needs a power-of-two rank count; use replicas otherwise")
    boxes = {0: ([0, 0, 0], [int(v) for v in finest_layout])}
    splits = []
    for level in range(k):
        nxt, lvl = {}, {}
        for prefix, (lo, hi) in boxes.items():
            ext = [hi[i] - lo[i] for i in range(3)]
            axis = int(np.argmax(ext))
            if axes is not None and ext[axes[level]] >= 2:
                axis = int(axes[level])
            if ext[axis] < 2:
                raise ValueError("volume has too few bricks to shard %d ways" % n_ranks)
            cut = lo[axis] + (ext[axis] + 1) // 2
            if weights is not None:
                blk = np.asarray(weights, np.float64)[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
                layers = blk.sum(axis=tuple(a for a in range(3) if a != axis))
                total = float(layers.sum())
                if total > 0.0:
                    cum = np.cumsum(layers)[:-1]                       # weight below a cut after layer i
                    d = np.abs(cum - 0.5 * total)
                    cut = lo[axis] + 1 + int(len(d) - 1 - np.argmin(d[::-1]))   # ties: the upper cut, like (ext + 1) // 2
            lvl[prefix] = (axis, cut)
            hi0 = list(hi); hi0[axis] = cut
            lo1 = list(lo); lo1[axis] = cut
            nxt[prefix * 2] = (list(lo), hi0)
            nxt[prefix * 2 + 1] = (lo1, list(hi))
        boxes = nxt
        splits.append(lvl)
    return [boxes[g] for g in range(n_ranks)], splits


def box_to_clip(box, finest_layout, float_layout):
    """Brick-unit box -> normalised volume coordinates.  Brick b of the finest level spans
    [b/L, (b+1)/L] with L = vLODLayout[0] (volume/innerBrick as float, GLVolumePool.cpp:89-107)."""
    lo, hi = box
    cmin, cmax = [], []
    for i in range(3):
        cmin.append(0.0 if lo[i] == 0 else float(np.float32(lo[i]) / np.float32(float_layout[i])))
        cmax.append(1.0 if hi[i] == finest_layout[i] else float(np.float32(hi[i]) / np.float32(float_layout[i])))
    return tuple(cmin), tuple(cmax)


def eye_in_volume(model_view, scale_extent):
    """Camera centre in normalised volume space: (0,0,0,1) * inverse(modelView) * S(1/extent) * T(.5)."""
    imv = np.linalg.inv(np.asarray(model_view, np.float64).reshape(4, 4))
    e = imv[3, :3] / imv[3, 3]
    return e / np.asarray(scale_extent, np.float64) + 0.5


def swap_plan(rank, n_ranks, splits, boxes, finest_layout, float_layout, eye_norm, n_pixels):
    """Binary-swap schedule of one rank: list of rounds, each
    dict(partner, keep=(lo,hi), send=(lo,hi), i_am_front).  Round r pairs ranks differing in bit r;
    bit r is the cut made at bisection level k-1-r."""
    k = len(splits)
    lo, hi = 0, n_pixels
    rounds = []
    for r in range(k):
        partner = rank ^ (1 << r)
        level = k - 1 - r
        prefix = rank >> (r + 1)
        axis, cut = splits[level][prefix]
        plane = 0.0 if cut == 0 else float(np.float32(cut) / np.float32(float_layout[axis]))
        low_side = ((rank >> r) & 1) == 0           # my block lies below the cut plane
        eye_low = eye_norm[axis] < plane
        i_am_front = low_side == eye_low
        mid = lo + (hi - lo) // 2
        if low_side:
            keep, send = (lo, mid), (mid, hi)
        else:
            keep, send = (mid, hi), (lo, mid)
        rounds.append(dict(partner=partner, keep=keep, send=send, i_am_front=i_am_front))
        lo, hi = keep
    return rounds


def final_ranges(n_ranks, n_pixels):
    """Pixel range owned by every rank after the last round (same arithmetic as swap_plan)."""
    k = int(round(math.log2(n_ranks)))
    out = []
    for rank in range(n_ranks):
        lo, hi = 0, n_pixels
        for r in range(k):
            mid = lo + (hi - lo) // 2
            if ((rank >> r) & 1) == 0:
                hi = mid
            else:
                lo = mid
        out.append((lo, hi))
    return out


def binary_swap(image, plan, dist, over, recv_buf):
    """Run the schedule.  image: flat [n_pixels, 4] float32 tensor (this rank's partial image, modified in
    place); over(front, back, out) blends tensors; recv_buf: scratch tensor [>= n_pixels/2, 4].
    Returns (lo, hi): the range of `image` that now holds final pixels."""
    lo, hi = 0, image.shape[0]
    for rd in plan:
        (klo, khi), (slo, shi) = rd["keep"], rd["send"]
        n_keep = khi - klo
        recv = recv_buf[:n_keep]
        ops = [dist.P2POp(dist.isend, image[slo:shi], rd["partner"]),
               dist.P2POp(dist.irecv, recv, rd["partner"])]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        mine = image[klo:khi]
        if rd["i_am_front"]:
            over(mine, recv, mine)
        else:
            over(recv, mine, mine)
        lo, hi = klo, khi
    return lo, hi


class SortLastRenderer:
    """One rank of the sort-last renderer: a CudaGridLeaper restricted to its block + the compositor."""

    def __init__(self, renderer, rank, n_ranks, finest_layout, float_layout, extent, view_dependent=True,
                 policy="screen"):
        import torch
        import torch.distributed as dist
        self.r, self.rank, self.n = renderer, rank, n_ranks
        self.torch, self.dist = torch, dist
        self.finest, self.flayout, self.extent = tuple(finest_layout), tuple(float_layout), tuple(extent)
        self.view_dependent = view_dependent and policy not in ("octant", "octantw")
        self.policy = policy
        self._axes = None
        self.weights = None
        self._partition(None)
        self._img = None
        self._recv = None
        self._gather = None
        self._rb = None
        # one stream for the traversal, the blend kernels and the NCCL exchange: ordering comes from the stream,
        # not from host synchronisation
        self.stream = torch.cuda.current_stream()
        renderer.set_stream(self.stream.cuda_stream)

    def _partition(self, axes):
        """(Re)cut the brick grid; bricks of the new block are paged in by the renderer's normal miss path."""
        self._axes = axes
        self.boxes, self.splits = shard_boxes(self.finest, self.n, axes, self.weights if self.policy.endswith("w") else None)
        cmin, cmax = box_to_clip(self.boxes[self.rank], self.finest, self.flayout)
        self.r.SetShardBox(cmin, cmax)

    def set_weights(self, weights=None):
        """Per-brick weights of the finest level for the balanced policies ("depthw", "octantw").  Default: 1 for every
        brick the page table does not flag empty (the flags come from brick min/max against the transfer function or
        isovalue -- identical on every rank), 0 otherwise.  Call after the first frame (visibility is computed there)."""
        if weights is None:
            from . import _lib as L
            n = int(self.finest[0]) * int(self.finest[1]) * int(self.finest[2])
            off = int(self.r.info().lod_offset[0])
            meta = self.r.page_table()[off:off + n]                      # finest level, x fastest
            nonempty = (meta != L.BI_EMPTY) & (meta != L.BI_CHILD_EMPTY)
            weights = nonempty.reshape(self.finest[2], self.finest[1], self.finest[0]).transpose(2, 1, 0).astype(np.float64)
        self.weights = np.asarray(weights, np.float64)
        self._partition(self._axes)

    def update_partition(self):
        """Re-cut the brick grid for the renderer's current view if the view-dependent axes changed; returns the
        camera position in normalised volume space.  Every rank derives the same cut from the same view."""
        r = self.r
        r._push_params()
        eye = eye_in_volume(np.array(list(r.params.model_view)), self.extent)
        if self.view_dependent and self.n > 1:
            axes = split_axes((0.5 - eye) * np.asarray(self.extent, np.float64), self.n, self.policy)
            if axes != self._axes:
                self._partition(axes)
        return eye

    def _wrap(self, ptr, n_pixels):
        torch = self.torch

        class _Dev:   # zero-copy view of library-owned device memory
            __cuda_array_interface__ = {"shape": (n_pixels, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Dev(), device="cuda")

    def render(self):
        """Render this rank's block (paging until converged) and composite.  Returns (lo, hi, image):
        `image[lo:hi]` are this rank's final pixels (flat RGBA32F tensor on the device)."""
        r = self.r
        eye = self.update_partition()
        st = r.PaintUntilConverged()
        p = r.params
        n_pixels = p.width * p.height
        ptr = r.device_image_ptr()
        if self._img is None or self._img.data_ptr() != ptr or self._img.shape[0] != n_pixels:
            self._img = self._wrap(ptr, n_pixels)
            self._recv = self.torch.empty((n_pixels // 2 + 1, 4), dtype=self.torch.float32, device="cuda")
        if self.n == 1:
            return 0, n_pixels, self._img, st
        plan = swap_plan(self.rank, self.n, self.splits, self.boxes, self.finest, self.flayout, eye, n_pixels)

        def over(front, back, out):
            r.composite_over(front.data_ptr(), back.data_ptr(), out.data_ptr(), front.shape[0])

        lo, hi = binary_swap(self._img, plan, self.dist, over, self._recv)
        return lo, hi, self._img, st

    def gather(self, lo, hi, image, dst_rank=0):
        """Collect the 1/G slices on dst_rank -> full flat RGBA32F image there (None elsewhere)."""
        torch, dist = self.torch, self.dist
        n_pixels = image.shape[0]
        if self.n == 1:
            return image
        ranges = final_ranges(self.n, n_pixels)
        if self.rank == dst_rank:
            if self._gather is None or self._gather.shape[0] != n_pixels:
                self._gather = torch.empty((n_pixels, 4), dtype=torch.float32, device=image.device)
            full = self._gather
            full[lo:hi] = image[lo:hi]
            ops = [dist.P2POp(dist.irecv, full[a:b], g) for g, (a, b) in enumerate(ranges) if g != dst_rank]
        else:
            full = None
            ops = [dist.P2POp(dist.isend, image[lo:hi], dst_rank)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return full

    def read_rgba8_async(self, full):
        """PBO-style read-back of the gathered frame on the destination rank: float -> unorm8 on the device, then an
        asynchronous copy into one of two page-locked host images on a side stream, so the copy overlaps the next
        frame's traversal.  Returns (host_image_uint8[n_pixels, 4], event); the image is valid once the event has
        completed (event.synchronize()).  At most two reads are in flight: a buffer is reused two calls later."""
        torch = self.torch
        n_pixels = full.shape[0]
        if self._rb is None or self._rb["n"] != n_pixels:
            self._rb = dict(n=n_pixels, i=0, copy=torch.cuda.Stream(),
                            dev=[torch.empty((n_pixels, 4), dtype=torch.uint8, device=full.device) for _ in range(2)],
                            host=[torch.empty((n_pixels, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                            done=[None, None])
        rb = self._rb
        k = rb["i"] % 2
        rb["i"] += 1
        if rb["done"][k] is not None:
            rb["done"][k].synchronize()       # the copy that last used this buffer pair has landed
        self.r.quantize_rgba8(full.data_ptr(), rb["dev"][k].data_ptr(), n_pixels)
        ready = torch.cuda.Event()
        ready.record(self.stream)
        rb["copy"].wait_event(ready)
        with torch.cuda.stream(rb["copy"]):
            rb["host"][k].copy_(rb["dev"][k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(rb["copy"])
        rb["done"][k] = done
        return rb["host"][k].numpy(), done


# ---------------------------------------------------------------------------------------------------------------------
# Depth pipeline: the alternative to binary swap for latency-bound frames.  The brick grid is cut into n slabs behind
# each other along the axis most parallel to the view; rank s is STAGE s: it marches every ray through the s-th slab
# from the eye only, starting from the colour and position the stage in front handed over (tvk_render_stage), and
# passes both images on over NCCL.  Rays that terminated early stay terminated (no rank renders what a front slab hid),
# the critical path of a launch is 1/n of the ray, and with consecutive frames in flight all ranks are busy.
# ---------------------------------------------------------------------------------------------------------------------
def depth_slabs(finest_layout, n, view_dir, weights=None, align=1, layer_weights=None):
    """-> (axis, boxes): boxes[s] = (lo[3], hi[3]) in finest-brick units of the s-th slab FROM THE EYE for a camera looking
    along view_dir (volume space).  weights [x, y, z]: cuts at the k/n quantiles of the weight along the axis (default:
    equal thickness); align: cuts are multiples of `align` bricks (bricks of coarser LoDs must not straddle a cut:
    a stage hands a ray on where it leaves its last brick)."""
    axis = int(np.argmax(np.abs(np.asarray(view_dir, np.float64))))
    L_ = int(finest_layout[axis])
    if n > L_:
        raise ValueError("volume has too few bricks to cut %d slabs" % n)
    if layer_weights is not None and float(np.asarray(layer_weights).sum()) > 0.0:
        layers = np.asarray(layer_weights, np.float64)          # per brick layer along the axis (measured, see rebalance)
    elif weights is not None and float(np.asarray(weights).sum()) > 0.0:
        layers = np.asarray(weights, np.float64).sum(axis=tuple(a for a in range(3) if a != axis))
    else:
        layers = np.ones(L_)
    cum = np.concatenate([[0.0], np.cumsum(layers)])
    cuts = [0]
    for k in range(1, n):
        c = int(np.argmin(np.abs(cum - cum[-1] * k / n)))
        if align > 1:
            c = int(round(c / align)) * align
        c = max(c, cuts[-1] + 1)                  # every slab keeps at least one brick layer
        c = min(c, L_ - (n - k))
        cuts.append(c)
    cuts.append(L_)
    boxes = []
    for k in range(n):
        lo, hi = [0, 0, 0], [int(v) for v in finest_layout]
        lo[axis], hi[axis] = cuts[k], cuts[k + 1]
        boxes.append((lo, hi))
    if float(view_dir[axis]) < 0.0:               # the camera looks down the axis: the far end comes first
        boxes.reverse()
    return axis, boxes


class DepthPipeline:
    """One rank (= stage) of the depth-pipelined renderer."""

    def __init__(self, renderer, rank, n_ranks, finest_layout, float_layout, extent, align=4, device="cuda"):
        """device "cpu" is for the host-logic tests (gloo, a stand-in renderer with host buffers)."""
        import torch
        import torch.distributed as dist
        self.r, self.rank, self.n = renderer, rank, n_ranks
        self.torch, self.dist = torch, dist
        self.device = device
        self.finest, self.flayout, self.extent = tuple(finest_layout), tuple(float_layout), tuple(extent)
        self.align = align
        self.weights = None
        self._in = None
        self._key = None
        self.view_id = None            # key of the per-view measured layer weights (rebalance)
        self.measured = {}
        self.stream = None
        if device == "cuda":
            self.stream = torch.cuda.current_stream()
            renderer.set_stream(self.stream.cuda_stream)

    def rebalance(self, stage_samples):
        """Load feedback for the current view (self.view_id): stage_samples[s] = the cost stage s measured with the current
        cuts -- its kernel time (or sample count) -- gathered from all ranks, identical everywhere.  Early ray
        termination puts most samples into the front of the volume while the longest (never terminated) rays bound the
        back slabs, which no view-independent weight can know; the measured cost of every stage is spread over its
        brick layers and the cuts move to the new quantiles.  A few rounds converge; the result is kept per view."""
        boxes = self.update_slab()
        axis = self._key[0]
        layers = np.zeros(int(self.finest[axis]))
        for s_, (lo, hi) in enumerate(boxes):
            layers[lo[axis]:hi[axis]] = float(stage_samples[s_]) / max(1, hi[axis] - lo[axis])
        layers += 1e-6 * max(1.0, layers.max())                    # keep every layer cuttable
        self.measured[self.view_id] = (axis, layers)
        self._key = None

    def set_weights(self, weights=None):
        """Non-empty finest-level bricks (SortLastRenderer.set_weights); call after one frame."""
        if weights is None:
            from . import _lib as L
            n = int(self.finest[0]) * int(self.finest[1]) * int(self.finest[2])
            off = int(self.r.info().lod_offset[0])
            meta = self.r.page_table()[off:off + n]
            nonempty = (meta != L.BI_EMPTY) & (meta != L.BI_CHILD_EMPTY)
            weights = nonempty.reshape(self.finest[2], self.finest[1], self.finest[0]).transpose(2, 1, 0).astype(np.float64)
        self.weights = np.asarray(weights, np.float64)
        self._key = None

    def update_slab(self):
        """This stage's slab for the renderer's current view (every rank derives the same cuts from the same view)."""
        r = self.r
        r._push_params()
        eye = eye_in_volume(np.array(list(r.params.model_view)), self.extent)
        view_dir = (0.5 - eye) * np.asarray(self.extent, np.float64)
        lw = self.measured.get(self.view_id)
        axis0 = int(np.argmax(np.abs(view_dir)))
        axis, boxes = depth_slabs(self.finest, self.n, view_dir, self.weights, self.align,
                                  lw[1] if lw is not None and lw[0] == axis0 else None)
        key = (axis, tuple(map(tuple, boxes[self.rank])))
        if key != self._key:
            self._key = key
            cmin, cmax = box_to_clip(boxes[self.rank], self.finest, self.flayout)
            r.SetShardBox(cmin, cmax)
        return boxes

    def _wrap(self, ptr, n_pixels):
        torch = self.torch
        if self.device != "cuda":
            import ctypes
            buf = (ctypes.c_float * (n_pixels * 4)).from_address(ptr)
            return torch.from_numpy(np.frombuffer(buf, np.float32).reshape(n_pixels, 4))

        class _Dev:
            __cuda_array_interface__ = {"shape": (n_pixels, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Dev(), device="cuda")

    def render_frame(self):
        """This stage's part of one frame: receive the hand-over images (stage > 0), march the slab, send them on
        (stage < n-1).  Returns (stats, image): image is the finished frame (flat RGBA32F device tensor) on the last
        stage, None elsewhere."""
        r, torch, dist = self.r, self.torch, self.dist
        self.update_slab()
        n_pixels = r.params.width * r.params.height
        pos_in = col_in = 0
        if self.rank > 0:
            if self._in is None or self._in[0].shape[0] != n_pixels:
                self._in = [torch.empty((n_pixels, 4), dtype=torch.float32, device=self.device) for _ in range(2)]
            dist.recv(self._in[0], src=self.rank - 1)
            dist.recv(self._in[1], src=self.rank - 1)
            if self.stream is not None:
                self.stream.synchronize()         # the library launches on this stream, but pages from the host
            pos_in, col_in = self._in[0].data_ptr(), self._in[1].data_ptr()
        st = r.RenderStage(pos_in, col_in)
        img, col, pos = r.stage_output_ptrs()
        if self.rank < self.n - 1:
            dist.send(self._wrap(pos, n_pixels), dst=self.rank + 1)
            dist.send(self._wrap(col, n_pixels), dst=self.rank + 1)
            return st, None
        return st, self._wrap(img, n_pixels)


# ---------------------------------------------------------------------------------------------------------------------
# The in-library path (tvk_sortlast_*, csrc/tvk_sortlast.inc): partition, direct-send exchange over NCCL, n-way blend and
# RGBA8 gather all happen inside libtvkcuda.so on the renderer's stream.  What is left here is the rendezvous (the
# NCCL unique id travels over whatever process group the host already has) and host-side mirrors of the plan for tests.
# ---------------------------------------------------------------------------------------------------------------------
def plan(finest_layout, float_layout, extent, model_view, n_ranks, policy=0):
    """tvk_sortlast_plan: (clip_min [n,3], clip_max [n,3], order [n]) -- host only, no device."""
    import ctypes as C
    from . import _lib as L
    cmin = np.zeros((n_ranks, 3), np.float32)
    cmax = np.zeros((n_ranks, 3), np.float32)
    order = np.zeros(n_ranks, np.int32)
    mv = np.ascontiguousarray(model_view, np.float32).reshape(-1)
    rc = L.lib().tvk_sortlast_plan(L.u32x3(*[int(v) for v in finest_layout]), L.f32x3(*[float(v) for v in float_layout]),
                                   (C.c_double * 3)(*[float(v) for v in extent]), L.f32x16(*mv), int(n_ranks), int(policy),
                                   cmin.ctypes.data_as(C.c_void_p), cmax.ctypes.data_as(C.c_void_p),
                                   order.ctypes.data_as(C.c_void_p))
    if rc != L.OK:
        raise L.TvkError(rc, (L.lib().tvk_last_error(None) or b"").decode())
    return cmin, cmax, order


def bsp_order(n_ranks, splits, float_layout, eye_norm):
    """Front-to-back order of the blocks of shard_boxes() for a camera at eye_norm (host mirror of sl_order)."""
    k = len(splits)
    out = []

    def walk(level, prefix):
        if level == k:
            out.append(prefix)
            return
        axis, cut = splits[level][prefix]
        plane = 0.0 if cut == 0 else float(np.float32(cut) / np.float32(float_layout[axis]))
        low = eye_norm[axis] < plane
        walk(level + 1, prefix * 2 + (0 if low else 1))
        walk(level + 1, prefix * 2 + (1 if low else 0))

    walk(0, 0)
    return out


def slice_range(n_pixels, n_ranks, rank):
    return n_pixels * rank // n_ranks, n_pixels * (rank + 1) // n_ranks


def direct_send(image, rank, n_ranks, order, dist, over):
    """Host mirror of the library's exchange for the gloo tests: every rank sends slice p of its partial image to rank p,
    receives the partials of its own slice and folds them front to back in `order` (over(front, back) -> tensor).
    Returns (lo, hi, composited slice)."""
    import torch
    n_pixels = image.shape[0]
    lo, hi = slice_range(n_pixels, n_ranks, rank)
    recv = {p: torch.empty((hi - lo, image.shape[1]), dtype=image.dtype) for p in range(n_ranks) if p != rank}
    ops = []
    for p in range(n_ranks):
        if p == rank:
            continue
        plo, phi = slice_range(n_pixels, n_ranks, p)
        ops.append(dist.P2POp(dist.isend, image[plo:phi].contiguous(), p))
        ops.append(dist.P2POp(dist.irecv, recv[p], p))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    acc = None
    for r in order:
        part = image[lo:hi] if r == rank else recv[r]
        acc = part.clone() if acc is None else over(acc, part)
    return lo, hi, acc


def init_library_sortlast(renderer, rank, n_ranks, dist=None, policy=0):
    """Rendezvous for tvk_sortlast_init: rank 0 creates the NCCL unique id, the host's process group (torch.distributed,
    any backend) carries its 128 bytes to the other ranks."""
    import torch
    from .renderer import CudaGridLeaper
    if n_ranks > 1:
        ids = [CudaGridLeaper.sortlast_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        cid = ids[0]
    else:
        cid = bytes(128)
    renderer.SortLastInit(cid, rank, n_ranks, policy)
