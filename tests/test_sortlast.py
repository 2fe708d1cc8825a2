"""Sort-last host logic (tuvok_b200/sortlast.py) on the CPU: sharding, the binary-swap schedule, and a
world_size-2 gloo run.  The blend used here is the ORACLE's over operator (test infrastructure); on GPU
ranks the same schedule drives the CUDA kernel tvk_composite_over (tests/test_gpu_sortlast.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_scenes
from oracle import orc
from scene import image_diff
from tuvok_b200 import sortlast


def scene_layout(s):
    inner = [b - 2 * s.overlap for b in s.brick]
    finest = [-(-v // i) for v, i in zip(s.size, inner)]
    fl = []
    for v, i in zip(s.size, inner):
        c = np.float32(v) / np.float32(i)
        if np.float32(int(c)) == c:
            c = c - c * np.finfo(np.float32).eps
        fl.append(np.float32(c))
    ext = np.array(s.size, np.float64) * np.array(s.scale, np.float64)
    return finest, fl, ext / ext.max()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_shard_boxes_tile_the_brick_grid(n):
    boxes, splits = sortlast.shard_boxes((7, 4, 5), n)
    assert len(boxes) == n and len(splits) == int(np.log2(n))
    cover = np.zeros((7, 4, 5), np.int32)
    for lo, hi in boxes:
        assert all(h > l for l, h in zip(lo, hi))
        cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
    assert (cover == 1).all()           # disjoint and complete


def test_view_dependent_axes_put_blocks_side_by_side():
    # looking mostly along z: never cut z; 8 ranks = 4 slabs along the most perpendicular axis x 2 along the next
    assert sortlast.split_axes((0.1, -0.3, 0.9), 2) == [0]
    assert sortlast.split_axes((0.1, -0.3, 0.9), 4) == [0, 1]
    assert sortlast.split_axes((0.1, -0.3, 0.9), 8) == [0, 1, 0]
    boxes, _ = sortlast.shard_boxes((8, 8, 8), 8, [0, 1, 0])
    assert all(hi[2] - lo[2] == 8 for lo, hi in boxes)
    assert sorted({(lo[0], hi[0]) for lo, hi in boxes}) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    # an axis that cannot be cut any more falls back to the longest one
    boxes, splits = sortlast.shard_boxes((1, 6, 4), 4, [0, 0])
    assert [a for lvl in splits for a, _ in lvl.values()] == [1, 2, 2] or len(boxes) == 4


def test_non_power_of_two_is_refused():
    with pytest.raises(ValueError):
        sortlast.shard_boxes((4, 4, 4), 3)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_swap_plan_is_consistent(n):
    finest, fl = (6, 5, 4), (np.float32(5.9), np.float32(4.7), np.float32(3.9))
    boxes, splits = sortlast.shard_boxes(finest, n)
    eye = np.array([1.7, -0.3, 0.4])
    n_pix = 1001
    plans = [sortlast.swap_plan(r, n, splits, boxes, finest, fl, eye, n_pix) for r in range(n)]
    for r in range(n):
        for k, rd in enumerate(plans[r]):
            other = plans[rd["partner"]][k]
            assert other["partner"] == r
            assert rd["keep"] == other["send"] and rd["send"] == other["keep"]
            assert rd["i_am_front"] != other["i_am_front"]
    ranges = sortlast.final_ranges(n, n_pix)
    assert sorted(ranges)[0][0] == 0 and sorted(ranges)[-1][1] == n_pix
    assert sum(b - a for a, b in ranges) == n_pix
    for r in range(n):
        assert plans[r][-1]["keep"] == ranges[r]


class FakeDist:
    """In-process stand-in for torch.distributed P2P: all ranks advance round by round."""

    class P2POp:
        def __init__(self, op, tensor, peer):
            self.op, self.tensor, self.peer = op, tensor, peer

    isend, irecv = "isend", "irecv"


def run_in_process(images, plans):
    """Execute the schedules of all ranks in lock step with numpy; returns the per-rank final ranges."""
    n = len(images)
    imgs = [im.copy() for im in images]
    for k in range(len(plans[0])):
        sends = {r: imgs[r][plans[r][k]["send"][0]:plans[r][k]["send"][1]].copy() for r in range(n)}
        for r in range(n):
            rd = plans[r][k]
            recv = sends[rd["partner"]]
            lo, hi = rd["keep"]
            mine = imgs[r][lo:hi]
            imgs[r][lo:hi] = orc.composite_over(mine, recv) if rd["i_am_front"] else orc.composite_over(recv, mine)
    return imgs


def partial_images(s, n, view_dependent=False):
    finest, fl, ext = scene_layout(s)
    axes = None
    if view_dependent:
        mv, _ = s.matrices()
        axes = sortlast.split_axes((0.5 - sortlast.eye_in_volume(mv, ext)) * ext, n)
    boxes, splits = sortlast.shard_boxes(finest, n, axes)
    outs = []
    for r in range(n):
        cmin, cmax = sortlast.box_to_clip(boxes[r], finest, fl)
        sr = golden_scenes.make(s._name, clip=(cmin, cmax))
        outs.append(sr.oracle_render()["image"].reshape(-1, 4).copy())
    mv, _ = s.matrices()
    eye = sortlast.eye_in_volume(mv, ext)
    n_pix = s.width * s.height
    plans = [sortlast.swap_plan(r, n, splits, boxes, finest, fl, eye, n_pix) for r in range(n)]
    return outs, plans


@pytest.mark.parametrize("name,n", [("c2_bricked36_1d_ert", 2), ("ragged_1d_lit", 4), ("inside_aniso_2d", 8),
                                    ("c3_bricked36_2d_lit", 2), ("c4_f32_iso", 4)])
@pytest.mark.parametrize("view_dependent", [False, True])
def test_sort_last_image_matches_single_renderer(name, n, view_dependent):
    s = golden_scenes.make(name)
    s._name = name
    single = s.oracle_render()
    parts, plans = partial_images(s, n, view_dependent)
    done = run_in_process(parts, plans)
    n_pix = s.width * s.height
    final = np.zeros((n_pix, 4), np.float32)
    for r, (a, b) in enumerate(sortlast.final_ranges(n, n_pix)):
        final[a:b] = done[r][a:b]
    mx, psnr = image_diff(orc.rgba8(final), single["rgba8"].reshape(-1, 4))
    assert mx <= 2 and psnr >= 45.0, (mx, psnr)
    assert orc.rgba8(final)[:, 3].any()


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = golden_scenes.make(name)
        s._name = name
        parts, plans = partial_images(s, world)
        img = torch.from_numpy(parts[rank].copy())
        recv = torch.empty((img.shape[0] // 2 + 1, 4), dtype=torch.float32)

        def over(front, back, out):
            out.copy_(torch.from_numpy(orc.composite_over(front.numpy().copy(), back.numpy().copy())))

        lo, hi = sortlast.binary_swap(img, plans[rank], dist, over, recv)
        q.put((rank, lo, hi, img[lo:hi].numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_binary_swap_over_gloo_world_size_2():
    name, world = "c2_bricked36_1d_ert", 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = golden_scenes.make(name)
    s._name = name
    parts, plans = partial_images(s, world)
    expect = run_in_process(parts, plans)
    n_pix = s.width * s.height
    assert sorted((lo, hi) for _, lo, hi, _ in got) == sorted(sortlast.final_ranges(world, n_pix))
    for rank, lo, hi, px in got:
        np.testing.assert_array_equal(px, expect[rank][lo:hi])


def test_weighted_cuts_balance_the_visible_bricks():
    """Policies "depthw" / "octantw": every cut halves the weight (non-empty bricks) of its block, blocks stay a disjoint
    cover of the brick grid, and uniform weights reproduce the unweighted cuts."""
    from tuvok_b200 import sortlast
    finest = (16, 12, 10)
    rng = np.random.default_rng(5)
    x, y, z = np.meshgrid(np.arange(16), np.arange(12), np.arange(10), indexing="ij")
    w = (((x - 11) ** 2 + (y - 6) ** 2 + (z - 5) ** 2) < 16).astype(np.float64)      # an off-centre ball of visible bricks
    for n in (2, 4, 8):
        for axes in (None, [0] * 3, [2, 2, 0]):
            boxes, splits = sortlast.shard_boxes(finest, n, axes, weights=w)
            cover = np.zeros(finest, int)
            sums = []
            for lo, hi in boxes:
                cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
                sums.append(w[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].sum())
                assert all(hi[i] > lo[i] for i in range(3))
            assert (cover == 1).all()
            assert len(splits) == int(np.log2(n))
            plain, _ = sortlast.shard_boxes(finest, n, axes)
            psums = [w[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].sum() for lo, hi in plain]
            assert max(sums) <= max(psums)                          # never worse balanced than the midpoint cuts
            if axes == [0] * 3:
                assert max(sums) - min(sums) <= w.sum(axis=(1, 2)).max() + 1e-9    # within one brick layer of even
    assert sortlast.shard_boxes(finest, 8, None, weights=np.ones(finest)) == sortlast.shard_boxes(finest, 8, None)
    assert sortlast.shard_boxes(finest, 4, [1, 1], weights=np.zeros(finest)) == sortlast.shard_boxes(finest, 4, [1, 1])
    assert sortlast.split_axes((0.1, 0.2, -0.9), 4, "depthw") == [2, 2] and sortlast.split_axes((1, 0, 0), 8, "octantw") is None


def test_depth_slabs_order_cover_and_load_feedback():
    """Depth pipeline planning (no GPU): slabs cover the grid once, come in front-to-back order for either viewing
    direction, respect the alignment, and the load feedback (DepthPipeline.rebalance's arithmetic) equalises a
    front-heavy cost profile within a few rounds."""
    from tuvok_b200 import sortlast
    finest = (64, 48, 32)
    for n in (1, 2, 3, 4, 8):
        for vd in ((0.9, 0.1, 0.2), (-0.9, 0.1, 0.2), (0.1, 0.2, -0.95)):
            axis, boxes = sortlast.depth_slabs(finest, n, vd)
            assert axis == int(np.argmax(np.abs(vd))) and len(boxes) == n
            cover = np.zeros(finest, int)
            for lo, hi in boxes:
                cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
                assert all(hi[i] > lo[i] for i in range(3))
            assert (cover == 1).all()
            starts = [b[0][axis] for b in boxes]
            assert starts == sorted(starts, reverse=vd[axis] < 0)    # stage 0 is the slab nearest to the eye
    _, boxes = sortlast.depth_slabs(finest, 4, (1, 0, 0), None, align=8)
    assert all(b[0][0] % 8 == 0 for b in boxes)
    with pytest.raises(ValueError):
        sortlast.depth_slabs((3, 3, 3), 4, (1, 0, 0))
    # weights: only the layers 20..49 hold visible bricks -> the cuts fall inside them
    w = np.zeros(finest); w[20:50] = 1.0
    _, boxes = sortlast.depth_slabs(finest, 3, (1, 0, 0), w)
    assert [b[0][0] for b in boxes] == [0, 30, 40]
    # load feedback: a cost profile that decays with depth (early ray termination)
    cost = np.exp(-np.arange(64) / 6.0)
    lw, spread = None, []
    for _ in range(5):
        _, boxes = sortlast.depth_slabs(finest, 4, (1.0, 0, 0), None, 1, lw)
        stage = [cost[b[0][0]:b[1][0]].sum() for b in boxes]
        spread.append(max(stage) / (sum(stage) / 4))
        lw = np.zeros(64)
        for s_, (lo, hi) in enumerate(boxes):
            lw[lo[0]:hi[0]] = stage[s_] / (hi[0] - lo[0])
        lw += 1e-6 * lw.max()
    assert spread[0] > 3.0 and spread[-1] < 1.3                     # from one stage doing 93 % to within 30 % of even


class _StageStandIn:
    """Plays CudaGridLeaper for DepthPipeline on the host: a stage adds (rank + 1) * tag to the accumulated colour it is
    handed, counts the stages a ray went through in the resume position, and records the slab it was given."""

    class _P:
        width, height = 6, 4
        model_view = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, -1.6, 1]

    def __init__(self, rank):
        self.rank, self.params, self.tag = rank, self._P(), 0.0
        n = self.params.width * self.params.height
        self.out = [np.zeros((n, 4), np.float32) for _ in range(3)]      # image, resume colour, resume position
        self.boxes = []

    def _push_params(self):
        pass

    def SetShardBox(self, cmin, cmax):
        self.boxes.append((cmin, cmax))

    def RenderStage(self, pos_ptr, col_ptr):
        import ctypes
        n = self.params.width * self.params.height

        def view(ptr):
            return np.frombuffer((ctypes.c_float * (n * 4)).from_address(ptr), np.float32).reshape(n, 4)
        pos = view(pos_ptr).copy() if pos_ptr else np.zeros((n, 4), np.float32)
        col = view(col_ptr).copy() if col_ptr else np.zeros((n, 4), np.float32)
        self.out[1][:] = col + (self.rank + 1) * self.tag
        self.out[2][:] = pos + 1.0
        self.out[0][:] = self.out[1]

        class _St:
            converged, samples, ms_raycast, bricks_paged = 1, 0, 0.0, 0
        return _St()

    def stage_output_ptrs(self):
        return tuple(o.ctypes.data for o in self.out)


def _pipeline_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ren = _StageStandIn(rank)
        pipe = sortlast.DepthPipeline(ren, rank, world, (8, 8, 8), (8.0, 8.0, 8.0), (1.0, 1.0, 1.0), align=1, device="cpu")
        frames = []
        for f in range(5):                                     # consecutive frames in flight
            ren.tag = float(10 ** (f % 3))
            st, img = pipe.render_frame()
            if img is not None:
                frames.append((img.numpy().copy(), ren.out[2].copy()))
        q.put((rank, frames, ren.boxes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_depth_pipeline_hand_over_over_gloo(world):
    """The N > 1 orchestration of the depth pipeline on CPU: stage s receives the two hand-over images of stage s-1, the last
    stage holds the frame, frames do not mix, and every rank cuts the same slabs (stage 0 nearest to the eye)."""
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, frames, boxes = q.get(timeout=240)
        got[rank] = (frames, boxes)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = got[world - 1][0]
    assert len(frames) == 5 and all(len(got[r][0]) == 0 for r in range(world - 1))     # only the last stage holds frames
    for f, (img, pos) in enumerate(frames):
        tag = float(10 ** (f % 3))
        assert np.all(img == sum(r + 1 for r in range(world)) * tag)     # every stage added its part exactly once
        assert np.all(pos == float(world))                                 # the ray went through all stages in order
    # the eye is at z = +1.6 (model_view above): slabs along z, stage 0 = the slab with the largest z
    zs = [got[r][1][0][0][2] for r in range(world)]
    assert zs == sorted(zs, reverse=True) and all(len(got[r][1]) == 1 for r in range(world))


@pytest.mark.parametrize("name,n", [("c2_bricked36_1d_ert", 2), ("c2_bricked36_1d_ert", 3), ("c3_bricked36_2d_lit", 2),
                                    ("ragged_1d_lit", 3), ("inside_aniso_2d", 2), ("ragged_1d_lit", 2)])
def test_oracle_depth_pipeline_equals_the_single_frame(name, n):
    """The oracle's restatement of a pipeline stage (orc_render.c `pipeline`, the checker of raycast_kernel<PIPE>): the
    stages of a frame, run one after the other with the hand-over images in between, end in the single-renderer frame
    up to the resume arithmetic; every ray is finished after the last stage; rays that terminated early in a front slab
    stay terminated, so the stages together take the single renderer's samples (binary swap: up to 1.6x)."""
    s = golden_scenes.make(name)
    r = s.oracle_pipeline(n)
    single = r["single"]
    d = np.abs(r["image"] - single["image"])
    mx, psnr = image_diff(r["rgba8"], single["rgba8"])
    assert float(d.max()) <= 0.0101 and mx <= 3 and psnr >= 60.0, (float(d.max()), mx, psnr)
    assert (r["resume_pos"].reshape(-1, 4)[:, 3] == 1000.0).all()
    total = sum(st["samples"] for st in r["stages"])
    assert total <= r["single_samples"] * 1.02 + n * s.width * s.height, (total, r["single_samples"])
    assert total >= r["single_samples"] * 0.98
    # hand-over images of the first stage: a ray is either finished or waits at/behind the slab's far side
    if n > 1:
        pos = r["stages"][0]["outs"][2].reshape(-1, 4)
        cov = r["stages"][0]["covered"].reshape(-1).astype(bool)
        live = cov & (pos[:, 3] != 1000.0)
        assert (pos[~cov, 3] == 1000.0).all()                       # pixels outside the volume are marked finished
        if live.any():
            assert (pos[live, 3] < 0).all()                         # a depth in front of the eye (eye-space z < 0)


def test_oracle_pipeline_paging_history():
    """A renderer that only ever ran stages pages each slab in pass by pass; one that rendered a whole frame first has
    every brick resident.  Both converge and both reproduce the single frame, but -- like every resumed GridLeaper frame
    -- not to the same bits (this difference, 2.9e-3 on this scene, is exactly what the CUDA stages showed against the
    warm-pool oracle: profiles/r1z_pipe_stage_check.txt)."""
    s = golden_scenes.make("c2_bricked36_1d_ert")
    warm, fresh = s.oracle_pipeline(2), golden_scenes.make("c2_bricked36_1d_ert").oracle_pipeline(2, fresh=True)
    single = warm["single"]["image"]
    for r in (warm, fresh):
        assert float(np.abs(r["image"] - single).max()) <= 0.0101
        assert (r["resume_pos"].reshape(-1, 4)[:, 3] == 1000.0).all()
    d = float(np.abs(warm["image"] - fresh["image"]).max())
    assert 1e-4 < d < 0.0101
    assert sum(st["samples"] for st in fresh["stages"]) == 507713      # the count the CUDA stages reported


# ---------------------------------------------------------------------------------------------------------------------
# the in-library sort-last path (csrc/tvk_sortlast.inc): its host-only plan against the Python mirror, and the
# direct-send exchange + front-to-back fold (host mirror, oracle over operator) against the single renderer
# ---------------------------------------------------------------------------------------------------------------------
def _mirror_plan(finest, fl, ext, mv, n, policy):
    eye = sortlast.eye_in_volume(mv, ext)
    axes = None
    if policy == 1:      # SCREEN: the axes most perpendicular to the view first (sortlast.split_axes, continued to k = 4)
        vd = (0.5 - eye) * ext
        o = sorted(range(3), key=lambda i: (abs(float(vd[i])), i))
        axes = [o[0], o[1], o[0], o[1]][:int(round(np.log2(n)))]
        assert axes[:3] == sortlast.split_axes(vd, min(n, 8), "screen")
    boxes, splits = sortlast.shard_boxes(finest, n, axes)
    clips = [sortlast.box_to_clip(b, finest, fl) for b in boxes]
    return clips, sortlast.bsp_order(n, [[lvl[p] for p in sorted(lvl)] for lvl in splits], fl, eye), boxes, splits, eye


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("policy", [0, 1])
def test_library_plan_equals_the_host_mirror(n, policy):
    from tuvok_b200.renderer import rotation_x, rotation_y
    import scene as sc
    finest, ext = (64, 23, 17), np.array([1.0, 0.36, 0.27])
    fl = [np.float32(63.99999), np.float32(22.6), np.float32(17.0) - np.float32(17.0) * np.finfo(np.float32).eps]
    for step in range(0, 36, 5):
        rot = (rotation_y(10.0 * step) @ rotation_x(20.0 + step)).astype(np.float32)
        mv = (rot @ sc.look_at((0, 0, 1.6), (0, 0, 0), (0, 1, 0))).astype(np.float32)
        cmin, cmax, order = sortlast.plan(finest, fl, ext, mv, n, policy)
        clips, morder, boxes, _, _ = _mirror_plan(finest, fl, ext, mv, n, policy)
        for r in range(n):
            np.testing.assert_array_equal(cmin[r], np.array(clips[r][0], np.float32))
            np.testing.assert_array_equal(cmax[r], np.array(clips[r][1], np.float32))
        assert list(order) == morder and sorted(order) == list(range(n))
        # the frontmost block holds the camera side of every cut: its box is the nearest to the eye
        eye = sortlast.eye_in_volume(mv, ext)
        centre = lambda r: np.array([(a + b) / 2 for a, b in zip(clips[r][0], clips[r][1])])
        d = [np.linalg.norm((centre(r) - eye) * ext) for r in order]
        assert d[0] == min(d)


def test_library_plan_refuses_bad_rank_counts():
    from tuvok_b200 import _lib as L
    with pytest.raises(L.TvkError):
        sortlast.plan((4, 4, 4), (4.0, 4.0, 4.0), (1.0, 1.0, 1.0), np.eye(4, dtype=np.float32), 3, 0)
    with pytest.raises(L.TvkError):
        sortlast.plan((1, 1, 2), (1.0, 1.0, 2.0), (1.0, 1.0, 1.0), np.eye(4, dtype=np.float32), 4, 0)


def _direct_send_in_process(parts, order):
    """The fold every rank performs on its slice, run for all ranks in this process."""
    n, n_pix = len(parts), parts[0].shape[0]
    final = np.zeros((n_pix, 4), np.float32)
    for r in range(n):
        lo, hi = sortlast.slice_range(n_pix, n, r)
        acc = parts[order[0]][lo:hi].copy()
        for k in order[1:]:
            acc = orc.composite_over(acc, parts[k][lo:hi].copy())
        final[lo:hi] = acc
    return final


@pytest.mark.parametrize("name,n,policy", [("c2_bricked36_1d_ert", 2, 0), ("ragged_1d_lit", 4, 1), ("inside_aniso_2d", 8, 0),
                                            ("c3_bricked36_2d_lit", 4, 0), ("c3_bricked36_2d_lit", 8, 1)])
def test_direct_send_image_matches_single_renderer(name, n, policy):
    s = golden_scenes.make(name)
    s._name = name
    single = s.oracle_render()
    finest, fl, ext = scene_layout(s)
    mv, _ = s.matrices()
    cmin, cmax, order = sortlast.plan(finest, fl, ext, mv, n, policy)
    parts = []
    for r in range(n):
        sr = golden_scenes.make(name, clip=(tuple(float(v) for v in cmin[r]), tuple(float(v) for v in cmax[r])))
        parts.append(sr.oracle_render()["image"].reshape(-1, 4).copy())
    final = _direct_send_in_process(parts, list(order))
    mx, psnr = image_diff(orc.rgba8(final), single["rgba8"].reshape(-1, 4))
    assert mx <= 2 and psnr >= 45.0, (mx, psnr)
    assert orc.rgba8(final)[:, 3].any()


def _ds_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)           # every rank draws all partial images, keeps its own
        parts = []
        for r in range(world):
            a = rng.random((1001, 1), dtype=np.float32) * (rng.random((1001, 1)) < 0.7)
            parts.append(np.concatenate([rng.random((1001, 3), dtype=np.float32) * a, a], axis=1).astype(np.float32))
        order = [1, 0] if world == 2 else list(range(world))[::-1]

        def over(front, back):
            return torch.from_numpy(orc.composite_over(front.numpy().copy(), back.numpy().copy()))

        lo, hi, acc = sortlast.direct_send(torch.from_numpy(parts[rank].copy()), rank, world, order, dist, over)
        q.put((rank, lo, hi, acc.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_direct_send_over_gloo(world):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ds_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(11)
    parts = []
    for r in range(world):
        a = rng.random((1001, 1), dtype=np.float32) * (rng.random((1001, 1)) < 0.7)
        parts.append(np.concatenate([rng.random((1001, 3), dtype=np.float32) * a, a], axis=1).astype(np.float32))
    order = [1, 0] if world == 2 else list(range(world))[::-1]
    expect = _direct_send_in_process(parts, order)
    covered = np.zeros(1001, bool)
    for rank, lo, hi, px in got:
        assert (lo, hi) == sortlast.slice_range(1001, world, rank)
        np.testing.assert_array_equal(px, expect[lo:hi])
        covered[lo:hi] = True
    assert covered.all()
