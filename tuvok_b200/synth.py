"""Seeded synthetic volumes in integer arithmetic (numpy), bit-identical to the device generator
`tvk_synth_volume` (tuvok_b200/csrc/k_bricker.cu).  Used to feed the CPU oracle and the host
brick-callback path with exactly the voxels the GPU path renders (V_sph, V_noise, V_ramp).
Arrays are indexed [z, y, x].
"""
import numpy as np

V_SPH, V_NOISE, V_RAMP = 0, 1, 2
_NP = {0: np.uint8, 1: np.uint16, 2: np.float32}


def _hash32(x, y, z, seed):
    with np.errstate(over="ignore"):
        h = np.uint32(seed) ^ (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ (z * np.uint32(0xC2B2AE3D))
        h = h ^ (h >> np.uint32(15))
        h = h * np.uint32(0x2C1B3C6D)
        h = h ^ (h >> np.uint32(12))
        h = h * np.uint32(0x297A2D39)
        h = h ^ (h >> np.uint32(15))
    return h


def _lattice_noise(x, y, z, shift, seed):
    s = np.uint32(shift)
    ix, iy, iz = x >> s, y >> s, z >> s
    one = np.uint64(1) << np.uint64(shift)
    m = one - np.uint64(1)
    fx, fy, fz = x.astype(np.uint64) & m, y.astype(np.uint64) & m, z.astype(np.uint64) & m
    acc = np.zeros(np.broadcast(x, y, z).shape, np.uint64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                v = (_hash32(ix + np.uint32(dx), iy + np.uint32(dy), iz + np.uint32(dz), seed) >> np.uint32(16)).astype(np.uint64)
                w = (fx if dx else one - fx) * (fy if dy else one - fy) * (fz if dz else one - fz)
                acc = acc + v * w
    return acc >> np.uint64(3 * shift)


def field_u16(kind, x, y, z, size, seed=0x5EED):
    """The analytic field at integer grid positions x, y, z (uint32 arrays, broadcastable, all inside the grid) of a grid of
    size = (nx, ny, nz): uint32 values 0..65535."""
    nx, ny, nz = (int(v) for v in size)
    shape = np.broadcast(x, y, z).shape
    if kind == V_RAMP:
        return np.broadcast_to((x + np.uint32(8) * y + np.uint32(64) * z) & np.uint32(0xFFFF), shape).astype(np.uint32)

    def axis(c, n):
        a = np.abs(2 * c.astype(np.int64) + 1 - n).astype(np.uint64)
        return a * np.uint64(4096) // np.uint64(n)

    ax, ay, az = axis(x, nx), axis(y, ny), axis(z, nz)
    d2 = ax * ax + ay * ay + az * az
    R2 = np.uint64(13589545)
    inside = d2 < R2
    d2c = np.where(inside, d2, np.uint64(0))
    w = ((R2 - d2c) << np.uint64(16)) // R2
    if kind == V_SPH:
        t = (d2c << np.uint64(16)) // R2
        ph = ((t * np.uint64(3)) & np.uint64(0xFFFF)).astype(np.int64) - 32768
        tri = np.abs(ph).astype(np.uint64) * np.uint64(2)
        v = np.minimum((w * tri) >> np.uint64(16), np.uint64(65535))
        return np.broadcast_to(np.where(inside, v, np.uint64(0)), shape).astype(np.uint32)
    mx = max(nx, ny, nz)
    lg = mx.bit_length() - 1
    shift0 = lg - 3 if lg > 3 else 0
    total = np.zeros(shape, np.uint64)
    for o in range(4):
        s = shift0 - o if shift0 > o else 0
        total = total + (_lattice_noise(x, y, z, s, (seed + o) & 0xFFFFFFFF) >> np.uint64(o))
    noise = total * np.uint64(8) // np.uint64(15)
    v = (noise * w) >> np.uint64(16)
    t0 = np.uint64(14000)
    r = np.minimum((np.maximum(v, t0) - t0) * np.uint64(2), np.uint64(65535))
    return np.where(inside & (v > t0), r, np.uint64(0)).astype(np.uint32)


def synth_u16(kind, size, seed=0x5EED):
    """size = (nx, ny, nz) -> uint32 array [z, y, x] with values 0..65535."""
    nx, ny, nz = (int(v) for v in size)
    x = np.arange(nx, dtype=np.uint32)[None, None, :]
    y = np.arange(ny, dtype=np.uint32)[None, :, None]
    z = np.arange(nz, dtype=np.uint32)[:, None, None]
    return field_u16(kind, x, y, z, size, seed)


def _to_dtype(v, dtype):
    if dtype == 0:
        return (v >> 8).astype(np.uint8)
    if dtype == 1:
        return v.astype(np.uint16)
    return (v.astype(np.float32) / np.float32(65535.0)).astype(np.float32)


def procedural_geometry(size, brick, overlap):
    """Level sizes, brick layouts and page-table offsets of the pool LoDs of a procedural dataset
    (tvk_set_procedural_volume): sizes ceil-halve per level; the pool ends at the first single-brick level."""
    brick = (brick,) * 3 if np.isscalar(brick) else tuple(brick)
    s = [int(v) for v in size]
    sizes, layouts, offsets = [], [], [0]
    while True:
        if sizes:
            s = [(v + 1) // 2 if v > 1 else v for v in s]
        lay = [-(-v // (b - 2 * overlap)) for v, b in zip(s, brick)]
        sizes.append(tuple(s)); layouts.append(tuple(lay))
        offsets.append(offsets[-1] + lay[0] * lay[1] * lay[2])
        if lay[0] * lay[1] * lay[2] == 1:
            return sizes, layouts, offsets


def procedural_brick(kind, size, dtype, seed, brick, overlap, x, y, z, lod):
    """Brick (x, y, z, lod) of the procedural multi-resolution dataset: the field sampled on level `lod`'s grid, `overlap`
    ghost voxels per side, 0 outside the grid; array [z, y, x] at the brick's own size."""
    brick = (brick,) * 3 if np.isscalar(brick) else tuple(brick)
    sizes, layouts, _ = procedural_geometry(size, brick, overlap)
    n, lay = sizes[lod], layouts[lod]
    bs, org = [], []
    for c, nn, l, b in zip((x, y, z), n, lay, brick):
        inner = b - 2 * overlap
        rem = nn % inner
        bs.append(2 * overlap + rem if (c == l - 1 and rem) else b)
        org.append(c * inner - overlap)
    gx = org[0] + np.arange(bs[0], dtype=np.int64)[None, None, :]
    gy = org[1] + np.arange(bs[1], dtype=np.int64)[None, :, None]
    gz = org[2] + np.arange(bs[2], dtype=np.int64)[:, None, None]
    ok = (gx >= 0) & (gx < n[0]) & (gy >= 0) & (gy < n[1]) & (gz >= 0) & (gz < n[2])
    cx, cy, cz = (np.clip(g, 0, nn - 1).astype(np.uint32) for g, nn in zip((gx, gy, gz), n))
    v = field_u16(kind, cx, cy, cz, n, seed)
    return _to_dtype(np.where(ok, v, np.uint32(0)).astype(np.uint32), dtype)


def synth_volume(kind, size, dtype, seed=0x5EED):
    """dtype: tvk dtype code (0 u8, 1 u16, 2 f32)."""
    return _to_dtype(synth_u16(kind, size, seed), dtype)
