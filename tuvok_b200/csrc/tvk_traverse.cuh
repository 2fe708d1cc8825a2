// tvk_traverse.cuh -- device code shared by the traversal kernels (k_raycast.cu: scalar volumes, k_color.cu: colour volumes):
// brick reference, filter footprints, transfer-function fetch, page-table walk (GetBrick), LoD selection, opacity
// correction and the analytic ray set-up.  Included INSIDE namespace tvk { namespace { ... } } after tvk_math.cuh.

struct BrickRef {
  f3 pool_entry, pool_exit, norm_exit, scale, trans;
  bool empty;
  int where;             // sort-last: 0 brick inside the shard box, 1 straddles it, 2 outside (never sampled)
  uint32_t bx, by, bz, bl;
  uint32_t ox, oy, oz;   // slot origin in virtual-atlas texels
  uint32_t id;           // page-table index
  uint32_t slot;         // linear pool coordinate (slot s starts at voxel s * slot_voxels)
};

// Filter footprint of one sample position inside a slot.  Integer pools are in the x-pair layout (k_pool.cu): element x
// of a row is the pair (voxel x, voxel x+1).
//   FastFoot (linear filter, ghost >= 2): the 4x4x4 neighbourhood [X-1..X+2]^3 always lies inside the
//         slot, so one clamped centre address + uniform row strides address everything.  A centre row of the
//         footprint (voxels X-1..X+2) is the two pairs at X-1 and X+1, a side row (voxels X, X+1) the pair at X:
//         16 loads fetch the 32 distinct voxels of the 7 overlapping trilinear footprints (4 loads for one footprint).
//         BS != 0 bakes a cubic brick size in (like the #defines of the reference's generated GLSL,
//         GLVolumePool.cpp:364-400), turning the addresses into immediate offsets.
//         The lerp trees run on packed fp32 (two rows per instruction, tvk_math.cuh): every lerp has the operands and
//         the rounding of tri(), so the result is bit-identical to seven independent tri() calls.
//         The footprint is DATA (the loaded words + the three filter fractions): the kernel fetches the footprint of a
//         ray's NEXT sample before it shades the current one, so the loads' L1 / L2 latency is covered by a whole
//         sample of arithmetic instead of stalling the warp (software pipelining; the launch is latency-bound at four
//         resident warps per scheduler).
//   SlowFoot (nearest filter or ghost < 2): texel indices are taken in the reference's VIRTUAL ATLAS
//         (capacity * brick texels, clamp-to-edge at the atlas border like GL_CLAMP_TO_EDGE) and then
//         split into (slot, texel-in-slot), so taps that leave a brick with a 1-voxel ghost read the
//         atlas neighbour exactly as the reference's 3D texture does.  One voxel per load (the pair's first half).
template <typename T, int BS, bool GRAD>
struct FastFoot {
  typedef typename PairOf<T>::W W;
  typedef typename PairOf<T>::V V;
  typedef PairCvt<T> CV;
  V w[GRAD ? 16 : 4];
  float fx, fy, fz;

  __device__ __forceinline__ void fetch(const RayConsts& P, const W* vox, uint32_t ox, uint32_t oy, uint32_t oz, f3 tc) {
    const float ux = fmaf(tc.x, P.pool_size_f[0], -0.5f);
    const float uy = fmaf(tc.y, P.pool_size_f[1], -0.5f);
    const float uz = fmaf(tc.z, P.pool_size_f[2], -0.5f);
    const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    fx = ux - x0; fy = uy - y0; fz = uz - z0;
    int X = (int)x0 - (int)ox, Y = (int)y0 - (int)oy, Z = (int)z0 - (int)oz;
    const int sy = BS ? BS : (int)P.total[0];
    const int sz = BS ? BS * BS : (int)(P.total[0] * P.total[1]);
    X = min(max(X, 1), (BS ? BS : (int)P.total[0]) - 3);
    Y = min(max(Y, 1), (BS ? BS : (int)P.total[1]) - 3);
    Z = min(max(Z, 1), (BS ? BS : (int)P.total[2]) - 3);
    const W* c = vox + (X + Y * sy + Z * sz);
    // the pair at element offset (i, j, k) from the footprint origin (immediate offsets when BS is baked in)
    auto ld = [&](int i, int j, int k) -> V { return load_pair(c + (i + j * sy + k * sz)); };
    if (!GRAD) {
      w[0] = ld(0, 0, 0); w[1] = ld(0, 1, 0); w[2] = ld(0, 0, 1); w[3] = ld(0, 1, 1);   // rows (y, z)
    } else {
#pragma unroll
      for (int j = 0; j < 2; j++) {   // centre rows y = j: pairs at x = -1 and x = +1, z = 0 / 1
        w[4 * j + 0] = ld(-1, j, 0); w[4 * j + 1] = ld(-1, j, 1); w[4 * j + 2] = ld(1, j, 0); w[4 * j + 3] = ld(1, j, 1);
      }
      w[8] = ld(0, -1, 0); w[9] = ld(0, -1, 1); w[10] = ld(0, 2, 0); w[11] = ld(0, 2, 1);   // rows y = -1, y = 2
#pragma unroll
      for (int j = 0; j < 2; j++) { w[12 + 2 * j] = ld(0, j, -1); w[13 + 2 * j] = ld(0, j, 2); }   // rows z = -1, z = 2
    }
  }
  // texture(volumePool, coords).r at the sample position (!GRAD footprints)
  __device__ __forceinline__ float centre(const RayConsts& P) const {
    static_assert(!GRAD, "a gradient footprint has no x = 0 pairs: use sample_with_gradient");
    // rows (y, z) of the footprint: pair = (voxel X, voxel X+1); the two z-slices share an instruction
    const f2 x0 = xlerp2<CV::kBiased>(F2(CV::lo(w[0]), CV::lo(w[2])), F2(CV::hi(w[0]), CV::hi(w[2])), fx);   // y = 0, z = (0, 1)
    const f2 x1 = xlerp2<CV::kBiased>(F2(CV::lo(w[1]), CV::lo(w[3])), F2(CV::hi(w[1]), CV::hi(w[3])), fx);   // y = 1
    const f2 y = lerp2(x0, x1, fy);
    return lerp1(y.x, y.y, fz) * P.norm;
  }
  // centre value + central-difference gradient (GLGridLeaper-GradientTools.glsl:6-16; the "Yp" tap is
  // fetched at -delta) from the 32 distinct voxels of the 7 overlapping footprints
  __device__ __forceinline__ void sample_with_gradient(const RayConsts& P, float& data, f3& grad) const {
    const float n = P.norm;
    constexpr bool B = CV::kBiased;
    // ---- x-lerps.  Centre rows (y, z in {0,1}): voxels m, a, b, p at x = -1, 0, 1, 2 from the pairs at -1 and +1;
    // the rows z = 0 and z = 1 of one y share the packed instructions.  xm / xc / xp = the x-lerps of the taps at
    // x-1, x, x+1.
    f2 xm[2], xc[2], xp[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const V l0 = w[4 * j + 0], l1 = w[4 * j + 1], h0 = w[4 * j + 2], h1 = w[4 * j + 3];
      const f2 m = F2(CV::lo(l0), CV::lo(l1)), a = F2(CV::hi(l0), CV::hi(l1));
      const f2 b = F2(CV::lo(h0), CV::lo(h1)), p = F2(CV::hi(h0), CV::hi(h1));
      xm[j] = xlerp2<B>(m, a, fx);
      xc[j] = xlerp2<B>(a, b, fx);
      xp[j] = xlerp2<B>(b, p, fx);
    }
    // side rows: y = -1 and y = 2 (z = 0, 1 packed), z = -1 and z = 2 (packed with each other, per y)
    const f2 xyl = xlerp2<B>(F2(CV::lo(w[8]), CV::lo(w[9])), F2(CV::hi(w[8]), CV::hi(w[9])), fx);       // row y = -1, z = (0, 1)
    const f2 xyh = xlerp2<B>(F2(CV::lo(w[10]), CV::lo(w[11])), F2(CV::hi(w[10]), CV::hi(w[11])), fx);   // row y = 2
    f2 xz[2];
#pragma unroll
    for (int j = 0; j < 2; j++)   // row y = j, z = (-1, 2)
      xz[j] = xlerp2<B>(F2(CV::lo(w[12 + 2 * j]), CV::lo(w[13 + 2 * j])), F2(CV::hi(w[12 + 2 * j]), CV::hi(w[13 + 2 * j])), fx);
    // ---- y-lerps, lanes = z slices
    const f2 yc = lerp2(xc[0], xc[1], fy);       // centre tap, z = (0, 1)
    const f2 yxm = lerp2(xm[0], xm[1], fy);      // tap at x-1
    const f2 yxp = lerp2(xp[0], xp[1], fy);      // tap at x+1
    const f2 yym = lerp2(xc[1], xyh, fy);        // tap at y+1 (fetched by the shader as "Ym"): rows y = 1, 2
    const f2 yyp = lerp2(xyl, xc[0], fy);        // tap at y-1 ("Yp"): rows y = -1, 0
    const f2 yz = lerp2(xz[0], xz[1], fy);       // z = (-1, 2)
    // ---- z-lerps
    data = lerp1(yc.x, yc.y, fz) * n;
    const float txm = lerp1(yxm.x, yxm.y, fz) * n, txp = lerp1(yxp.x, yxp.y, fz) * n;
    const float tym = lerp1(yym.x, yym.y, fz) * n, typ = lerp1(yyp.x, yyp.y, fz) * n;
    const float tzp = lerp1(yc.y, yz.y, fz) * n;     // tap at z+1: slices z = 1, 2
    const float tzm = lerp1(yz.x, yc.x, fz) * n;     // tap at z-1: slices z = -1, 0
    grad = F3((txm - txp) / 2.0f, (typ - tym) / 2.0f, (tzm - tzp) / 2.0f);
  }
};

template <typename T>
struct SlowFoot {
  typedef typename PairOf<T>::W W;
  const W* c;            // first element of the pool
  uint64_t xo[4], yo[4], zo[4];   // element offsets (slot part + in-slot part) of X-1..X+2 etc.
  float fx, fy, fz;
  bool nearest;

  __device__ __forceinline__ void fetch(const RayConsts& P, const W* pool, f3 tc) {
    int X, Y, Z;
    nearest = P.nearest != 0;
    if (nearest) {
      X = (int)floorf(tc.x * P.pool_size_f[0]);
      Y = (int)floorf(tc.y * P.pool_size_f[1]);
      Z = (int)floorf(tc.z * P.pool_size_f[2]);
      fx = fy = fz = 0.0f;
    } else {
      const float ux = fmaf(tc.x, P.pool_size_f[0], -0.5f);
      const float uy = fmaf(tc.y, P.pool_size_f[1], -0.5f);
      const float uz = fmaf(tc.z, P.pool_size_f[2], -0.5f);
      const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
      fx = ux - x0; fy = uy - y0; fz = uz - z0;
      X = (int)x0; Y = (int)y0; Z = (int)z0;
    }
    const uint32_t sy = P.total[0], sz = P.total[0] * P.total[1];
    c = pool;
    const int ax = (int)(P.capacity[0] * P.total[0]) - 1, ay = (int)(P.capacity[1] * P.total[1]) - 1,
              az = (int)(P.capacity[2] * P.total[2]) - 1;
    const uint64_t slot_y = (uint64_t)P.capacity[0] * P.slot_voxels, slot_z = slot_y * P.capacity[1];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t gx = (uint32_t)min(max(X - 1 + i, 0), ax), gy = (uint32_t)min(max(Y - 1 + i, 0), ay),
                     gz = (uint32_t)min(max(Z - 1 + i, 0), az);
      xo[i] = (uint64_t)(gx / P.total[0]) * P.slot_voxels + gx % P.total[0];
      yo[i] = (uint64_t)(gy / P.total[1]) * slot_y + (uint64_t)(gy % P.total[1]) * sy;
      zo[i] = (uint64_t)(gz / P.total[2]) * slot_z + (uint64_t)(gz % P.total[2]) * sz;
    }
  }
  // voxel at texel offset (i, j, k) in [-1, 2]^3 from the footprint origin
  __device__ __forceinline__ float v(int i, int j, int k) const {
    return first_voxel<T>(__ldg(c + (xo[1 + i] + yo[1 + j] + zo[1 + k])));
  }
  // texture(volumePool, coords).r at texel offset (dx,dy,dz)
  __device__ __forceinline__ float tap(const RayConsts& P, int dx, int dy, int dz) const {
    if (nearest) return v(dx, dy, dz) * P.norm;
    return tri(v(dx, dy, dz), v(dx + 1, dy, dz), v(dx, dy + 1, dz), v(dx + 1, dy + 1, dz), v(dx, dy, dz + 1),
               v(dx + 1, dy, dz + 1), v(dx, dy + 1, dz + 1), v(dx + 1, dy + 1, dz + 1), fx, fy, fz) * P.norm;
  }
  __device__ __forceinline__ float centre(const RayConsts& P) const { return tap(P, 0, 0, 0); }
  __device__ __forceinline__ void sample_with_gradient(const RayConsts& P, float& data, f3& grad) const {
    data = tap(P, 0, 0, 0);
    const float xp = tap(P, 1, 0, 0), xm = tap(P, -1, 0, 0);
    const float yp = tap(P, 0, -1, 0), ym = tap(P, 0, 1, 0);
    const float zp = tap(P, 0, 0, 1), zm = tap(P, 0, 0, -1);
    grad = F3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
  }
};

// one footprint type for the kernel: GRAD selects what a FastFoot holds (the slow path loads at use)
template <typename T, bool FAST, int BS, bool GRAD>
struct Foot {
  typedef typename PairOf<T>::W W;
  FastFoot<T, BS, GRAD> fast;
  SlowFoot<T> slow;
  __device__ __forceinline__ void fetch(const RayConsts& P, const W* pool, const W* vox, uint32_t ox, uint32_t oy, uint32_t oz, f3 tc) {
    if (FAST) fast.fetch(P, vox, ox, oy, oz, tc); else slow.fetch(P, pool, tc);
  }
  __device__ __forceinline__ float centre(const RayConsts& P) const { return FAST ? fast.centre(P) : slow.centre(P); }
  __device__ __forceinline__ void sample_with_gradient(const RayConsts& P, float& data, f3& grad) const {
    if (FAST) fast.sample_with_gradient(P, data, grad); else slow.sample_with_gradient(P, data, grad);
  }
};

// RGBA8 table, GL_NEAREST, clamp-to-edge (GPUMemMan.cpp:398-401, GLTexture1D.h:48-51).  The table stays RGBA8
// (4 bytes per entry: a 4096 x 256 2D table is 4 MB and L2-resident); the texture unit's unorm8 -> float conversion
// byte / 255.0f is done per fetch, exactly (unorm8x4, tvk_math.cuh).
__device__ __forceinline__ f4 tf_lookup(const RayConsts& P, float s, float t) {
  const int w = (int)P.tf_w, h = (int)P.tf_h;
  int ix = (int)floorf(s * (float)w);
  ix = min(max(ix, 0), w - 1);
  int iy = 0;
  if (h > 1) {
    iy = (int)floorf(t * (float)h);
    iy = min(max(iy, 0), h - 1);
  }
  return unorm8x4(__ldg(P.tf + ((uint32_t)iy * (uint32_t)w + (uint32_t)ix)));
}

__device__ __forceinline__ void brick_coords(const RayConsts& P, f3 pos, uint32_t lod, uint32_t& x, uint32_t& y,
                                             uint32_t& z) {
  x = (uint32_t)(pos.x * P.lod_layout[lod][0]);
  y = (uint32_t)(pos.y * P.lod_layout[lod][1]);
  z = (uint32_t)(pos.z * P.lod_layout[lod][2]);
}
__device__ __forceinline__ uint32_t brick_index(const RayConsts& P, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  return P.lod_offset[lod] + x + y * P.lod_layout_sz[lod][0] + z * P.lod_layout_sz[lod][1];
}
__device__ __forceinline__ uint32_t brick_info(const RayConsts& P, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  return __ldg(P.meta + brick_index(P, x, y, z, lod));
}

// GLHashTable.cpp:140-182.  Rays of a warp that miss the same brick elect one reporter first
// (warp vote) so the table sees one CAS chain per distinct brick per warp.
__device__ __forceinline__ void report_missing(const RayConsts& P, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  if (!P.hash || P.hash_size == 0) return;
  const uint32_t ser = 1 + x + y * P.finest[0] + z * P.finest[0] * P.finest[1] +
                       lod * P.finest[0] * P.finest[1] * P.finest[2];
  const unsigned act = __activemask();
  const unsigned same = __match_any_sync(act, ser);
  if ((unsigned)(__ffs(same) - 1) != (threadIdx.x & 31u)) return;
  uint32_t rehash = 0;
  do {
    uint32_t h = (ser + rehash) % P.hash_size;
    uint32_t old = atomicCAS(P.hash + h, 0u, ser);
    if (old == 0 || old == ser) return;
  } while (++rehash < P.rehash_count);
}

enum { IN_SHARD = 0, PARTLY_IN_SHARD = 1, OUTSIDE_SHARD = 2 };

// position of the brick box [c0,c1] relative to the sort-last shard box
__device__ __forceinline__ int classify_brick(const RayConsts& P, f3 c0, f3 c1) {
  if (!P.shard) return IN_SHARD;
  if (c1.x <= P.sh_lo[0] || c0.x >= P.sh_hi[0] || c1.y <= P.sh_lo[1] || c0.y >= P.sh_hi[1] || c1.z <= P.sh_lo[2] ||
      c0.z >= P.sh_hi[2])
    return OUTSIDE_SHARD;
  const bool inside = c0.x >= P.sh_lo[0] && c1.x <= P.sh_hi[0] && c0.y >= P.sh_lo[1] && c1.y <= P.sh_hi[1] &&
                      c0.z >= P.sh_lo[2] && c1.z <= P.sh_hi[2];
  return inside ? IN_SHARD : PARTLY_IN_SHARD;
}

// Returns 1 (brick of the requested LOD present), 0 (it was missing: reported, a coarser one is returned) or
// -1 (SPEC only: the brick is missing and this is a look-ahead call -- nothing was reported or changed; the
// caller retries when the ray really stands at this brick, so miss reports keep the shader's order).
template <bool SPEC>
__device__ __forceinline__ int get_brick(const RayConsts& P, f3 pos, uint32_t& lod, f3 dir, f3 dv, BrickRef& o) {
  const uint32_t max_lod = P.lod_count - 1;
  pos = F3(clampf(pos.x, 0.0f, 1.0f), clampf(pos.y, 0.0f, 1.0f), clampf(pos.z, 0.0f, 1.0f));
  int found = 1;
  uint32_t bx, by, bz;
  brick_coords(P, pos, lod, bx, by, bz);
  uint32_t info = brick_info(P, bx, by, bz, lod);
  // sort-last: a missing brick that does not touch this rank's block lives on another rank.  It is walked
  // through (same step arithmetic, nominal slot 0) but never requested, sampled or replaced by a coarser
  // level, so the ray reaches this rank's block at the single-GPU ray's sample phase.
  bool foreign = false;
  if (P.shard && info == TVK_BI_MISSING) {
    const f3 fl = F3(P.lod_layout[lod]);
    foreign = classify_brick(P, div3(F3((float)bx, (float)by, (float)bz), fl),
                             div3(F3((float)(bx + 1), (float)(by + 1), (float)(bz + 1)), fl)) == OUTSIDE_SHARD;
  }
  if (info == TVK_BI_MISSING && !foreign) {
    if (SPEC) return -1;
    const uint32_t start = lod;
    report_missing(P, bx, by, bz, lod);
    found = 0;
    // the reference loops `do {...} while (brickInfo == BI_MISSING)`: the coarsest brick is always
    // resident (UploadFirstBrick), so the bound only guards a corrupted table
    while (info == TVK_BI_MISSING && lod < max_lod) {
      lod++;
      brick_coords(P, pos, lod, bx, by, bz);
      info = brick_info(P, bx, by, bz, lod);
      if (info == TVK_BI_MISSING) {
        if (P.strategy == TVK_BS_REQUEST_ALL) report_missing(P, bx, by, bz, lod);
        else if (P.strategy == TVK_BS_SKIP_ONE_LEVEL && start + 1 == lod) report_missing(P, bx, by, bz, lod);
        else if (P.strategy == TVK_BS_SKIP_TWO_LEVELS && start + 2 == lod) report_missing(P, bx, by, bz, lod);
      }
    }
  }
  o.empty = !foreign && info <= TVK_BI_EMPTY;
  if (o.empty) {
    for (uint32_t lo = lod + 1; lo < max_lod; ++lo) {   // strict <, GLVolumePool.cpp:593
      uint32_t lx, ly, lz;
      brick_coords(P, pos, lo, lx, ly, lz);
      uint32_t li = brick_info(P, lx, ly, lz, lo);
      if (li == TVK_BI_CHILD_EMPTY) { bx = lx; by = ly; bz = lz; info = li; lod = lo; }
      else break;
    }
  }
  // GetBrickCorners / BrickExit
  const f3 lay = F3(P.lod_layout[lod]);
  const f3 c0 = div3(F3((float)bx, (float)by, (float)bz), lay);
  const f3 c1 = div3(F3((float)(bx + 1), (float)(by + 1), (float)(bz + 1)), lay);
  float tx = ((dv.x < 0.0f ? c0.x : c1.x) - pos.x) * dv.x;
  float ty = ((dv.y < 0.0f ? c0.y : c1.y) - pos.y) * dv.y;
  float tz = ((dv.z < 0.0f ? c0.z : c1.z) - pos.z) * dv.z;
  float tm = fminf(fminf(tx, ty), tz);
  o.norm_exit = add3(pos, scl3(dir, tm));
  o.bx = bx; o.by = by; o.bz = bz; o.bl = lod;
  o.where = IN_SHARD;
  if (o.empty) return found;
  o.where = foreign ? OUTSIDE_SHARD : classify_brick(P, c0, c1);
  o.id = brick_index(P, bx, by, bz, lod);   // only used by the counting kernels
  // InfoToCoords / BrickPoolCoords / NormCoordsToPoolCoords
  // a brick outside the shard box is only stepped through: always in the pool coordinates of slot 0, resident or not,
  // so the ray's positions behind it do not depend on what other views have paged into this pool
  const uint32_t index = o.where == OUTSIDE_SHARD ? 0u : info - TVK_BI_FLAG_COUNT;
  const uint32_t sx = index % P.capacity[0], sy = (index / P.capacity[0]) % P.capacity[1],
                 sz = index / (P.capacity[0] * P.capacity[1]);
  o.ox = sx * P.total[0]; o.oy = sy * P.total[1]; o.oz = sz * P.total[2];
  o.slot = index;
  const f3 ps = F3(P.pool_size_f), ov = F3(P.overlap_tc);
  const f3 vp = F3((float)o.ox, (float)o.oy, (float)o.oz);
  const f3 vq = F3((float)(o.ox + P.total[0]), (float)(o.oy + P.total[1]), (float)(o.oz + P.total[2]));
  const f3 pc0 = add3(div3(vp, ps), ov);
  const f3 pc1 = sub3(div3(vq, ps), ov);
  o.scale = div3(sub3(pc1, pc0), sub3(c1, c0));
  o.trans = sub3(pc0, mul3(c0, o.scale));
  o.pool_entry = add3(mul3(pos, o.scale), o.trans);
  o.pool_exit = add3(mul3(o.norm_exit, o.scale), o.trans);
  return found;
}

// min(iMaxLOD, uint(log2(fLoDFactor*(-dist)/fLevelZeroWorldSpaceError))); uint(log2 x) = exponent of x
__device__ __forceinline__ uint32_t compute_lod(const RayConsts& P, float dist) {
  float x = P.lod_factor * (-dist) / P.lzwse;
  const uint32_t max_lod = P.lod_count - 1;
  if (!(x >= 1.0f)) return 0;
  if (isinf(x)) return max_lod;
  uint32_t l = ((__float_as_uint(x) >> 23) & 0xffu) - 127u;
  return min(l, max_lod);
}

__device__ __forceinline__ float opacity_correct(const RayConsts& P, float a) {
  if (P.oc == 1.0f) return a;
  return 1.0f - powf(1.0f - a, P.oc);
}

// analytic ray/box entry + exit at the pixel centre (what the rasterised bbox front/back faces and
// the near-plane quad deliver per fragment)
__device__ __forceinline__ bool ray_setup(const RayConsts& P, uint32_t px, uint32_t py, f4& entry, f4& exit_, bool shard_test = true) {
  float nx = ((float)px + 0.5f) / (float)P.width * 2.0f - 1.0f;
  float ny = ((float)py + 0.5f) / (float)P.height * 2.0f - 1.0f;
  f4 nr = xform4(P.inv_proj, nx, ny, -1.0f, 1.0f);
  f3 pn = F3(nr.x / nr.w, nr.y / nr.w, nr.z / nr.w);
  f4 o4 = xform4(P.emm, 0.0f, 0.0f, 0.0f, 1.0f);
  f4 n4 = xform4(P.emm, pn.x, pn.y, pn.z, 1.0f);
  const float o[3] = {o4.x, o4.y, o4.z};
  const float d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
  float s_in = -INFINITY, s_out = INFINITY;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (d[i] == 0.0f) {
      if (o[i] < 0.0f || o[i] > 1.0f) return false;
      continue;
    }
    float t0 = (0.0f - o[i]) / d[i], t1 = (1.0f - o[i]) / d[i];
    s_in = fmaxf(s_in, fminf(t0, t1));
    s_out = fminf(s_out, fmaxf(t0, t1));
  }
  if (P.clip_plane_on) {
    // the bbox cut by the clip plane (Clipper::BoxPlane keeps f <= 0): f(s) = a + s * b along the ray
    const float a = fmaf(P.clip_plane[2], o[2], fmaf(P.clip_plane[1], o[1], P.clip_plane[0] * o[0])) + P.clip_plane[3];
    const float b = fmaf(P.clip_plane[2], d[2], fmaf(P.clip_plane[1], d[1], P.clip_plane[0] * d[0]));
    if (b > 0.0f) s_out = fminf(s_out, (0.0f - a) / b);
    else if (b < 0.0f) s_in = fmaxf(s_in, (0.0f - a) / b);
    else if (a > 0.0f) return false;
  }
  const float s0 = fmaxf(s_in, 1.0f);
  if (!(s_out > s0)) return false;
  if (P.shard && shard_test) {
    // sort-last: the ray keeps its whole-volume entry/exit (its sample positions are those of the single-GPU
    // ray); a pixel whose ray never meets this rank's brick block is simply not shaded
    float a_in = -INFINITY, a_out = INFINITY;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float lo = P.clip_min[i], hi = P.clip_max[i];
      if (d[i] == 0.0f) {
        if (o[i] < lo || o[i] > hi) return false;
        continue;
      }
      float t0 = (lo - o[i]) / d[i], t1 = (hi - o[i]) / d[i];
      a_in = fmaxf(a_in, fminf(t0, t1));
      a_out = fminf(a_out, fmaxf(t0, t1));
    }
    if (!(fminf(a_out, s_out) > fmaxf(a_in, s0))) return false;
  }
  const f3 pe = scl3(pn, s0), px_ = scl3(pn, s_out);
  f4 e = xform4(P.emm, pe.x, pe.y, pe.z, 1.0f);
  f4 x = xform4(P.emm, px_.x, px_.y, px_.z, 1.0f);
  entry.x = e.x; entry.y = e.y; entry.z = e.z; entry.w = pe.z;
  exit_.x = x.x; exit_.y = x.y; exit_.z = x.z; exit_.w = px_.z;
  return true;
}

__device__ __forceinline__ float4 to4(f4 v) { return make_float4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ f4 from4(float4 v) { f4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }

