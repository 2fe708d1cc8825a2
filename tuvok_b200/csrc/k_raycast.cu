// k_raycast.cu -- the GridLeaper traversal kernel for sm_100a.
//
// One thread = one ray (one fragment of GLGridLeaper-blend.glsl / -iso.glsl main()); a warp is
// an 8x4 pixel tile so neighbouring rays walk the same bricks and share cache lines.  Replaces
// (reference file:line):
//   ray entry/exit     GLGridLeaper.cpp:560-620, GLGridLeaper-NearPlane-VS.glsl:9-14,
//                      GLGridLeaper-entry-VS.glsl:10-14, GLGridLeaper-frontfaces-FS.glsl:6-8
//   main() DVR / ISO   Shaders/GLGridLeaper-blend.glsl:65-228, GLGridLeaper-iso.glsl:68-200
//   page-table walk    generated GLSL, Renderer/GL/GLVolumePool.cpp:484-656
//   miss reports       generated GLSL, Renderer/GL/GLHashTable.cpp:136-182
//   classification     GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}.glsl, GLGridLeaper-GradientTools.glsl:6-23,
//                      lighting.glsl:33-43, Compositing.glsl:33-38
//
// HBM layout: the pool is SLOT-LINEAR -- slot s (= the reference's linear pool coordinate
// x + y*capX + z*capX*capY) is one contiguous maxTotalBrickSize^3 block, x fastest -- instead of
// the reference's 3D-texture atlas.  The shader's pool texture coordinates are kept (virtual
// atlas of capacity*brick voxels) so the arithmetic stays the reference's; only the final
// texel address is slot-local.
//
// Control flow: the shader's nested loops (bricks along the ray / samples inside a brick) are run
// as one flat per-warp loop with a "fetch next brick" phase and a "take one sample" phase, so
// lanes whose brick ends early do not idle until the slowest lane of the warp finishes its brick;
// per ray the sequence of operations is exactly the shader's.
//
// Arithmetic contract: IEEE fp32, no implicit FMA contraction (this file is compiled
// with -fmad=false); fmaf() exactly where the contract names it (texel-coordinate map, trilinear
// lerps, dot products, under-compositing).  Gradient taps sit exactly +-1 texel from the centre
// sample and share its filter fractions (the GLSL adds sampleDelta = 1/poolSize in texture
// coordinates, i.e. one texel; the precision of that is implementation-defined in GL).
#include "tvk_dev.h"

namespace tvk {
namespace {

#include "tvk_math.cuh"

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

#ifndef TVK_FETCH_LANES
#define TVK_FETCH_LANES 32
#endif
constexpr int kFetchLanes = TVK_FETCH_LANES;
#ifndef TVK_MIN_BLOCKS
#define TVK_MIN_BLOCKS 8
#endif
#ifndef TVK_PREFETCH
#define TVK_PREFETCH 0
#endif
constexpr bool kPrefetch = TVK_PREFETCH != 0;   // fetch the next sample's footprint before shading the current one
// CTA = TVK_WX x TVK_WY warps, each an 8x4 pixel tile (CTA covers 8*WX x 4*WY pixels)
#ifndef TVK_WX
#define TVK_WX 2
#endif
#ifndef TVK_WY
#define TVK_WY 2
#endif
constexpr int kWX = TVK_WX, kWY = TVK_WY, kThreads = 32 * TVK_WX * TVK_WY;
#ifndef TVK_SKIP_CLEAR
#define TVK_SKIP_CLEAR 1
#endif
constexpr bool kSkipClear = TVK_SKIP_CLEAR != 0;   // A/B switch for the zero-alpha shortcut in shade()
#ifndef TVK_CENTRE_OUT
#define TVK_CENTRE_OUT 1
#endif
constexpr bool kCentreOut = TVK_CENTRE_OUT != 0;
// k-th element of the sequence c, c-1, c+1, c-2, c+2, ... (c = n/2): a bijection of [0, n)
__device__ __forceinline__ uint32_t centre_out(uint32_t k, uint32_t n) {
  const uint32_t c = n >> 1, d = (k + 1) >> 1;
  return (k & 1u) ? c - d : c + d;
}

// State of the brick chain (page-table walk).  None of it is needed while a lane samples, so between two chain
// phases it is PARKED in shared memory (TVK_PARK) instead of occupying ~37 registers of every thread for the
// whole kernel; the sample phase then fits a smaller register budget and more warps are resident per SM.
struct ChainSt {
  f3 entry, nexit, dir, dv, nudge, cur;
  float ray_len, step, entry_depth, exit_depth, t;
  f4 resume_pos, resume_col;
  uint32_t j, lbx, lby, lbz, lbl;
  bool optimal;
};
#ifndef TVK_PARK
#define TVK_PARK 0   // measured on B200 (gpurun_out/r1_ab3.log): parking + more resident warps is SLOWER (the extra
#endif               // shared memory shrinks L1 and more warps thrash it): 229 fps vs 238; 10/12 CTAs: 179/167 fps
constexpr bool kPark = TVK_PARK != 0;
constexpr int kChainWords = 37;
__device__ __forceinline__ void park_const(const ChainSt& c, float (*m)[kThreads], int tid) {   // written once per ray
  m[0][tid] = c.entry.x; m[1][tid] = c.entry.y; m[2][tid] = c.entry.z;
  m[3][tid] = c.nexit.x; m[4][tid] = c.nexit.y; m[5][tid] = c.nexit.z;
  m[6][tid] = c.dir.x; m[7][tid] = c.dir.y; m[8][tid] = c.dir.z;
  m[9][tid] = c.dv.x; m[10][tid] = c.dv.y; m[11][tid] = c.dv.z;
  m[12][tid] = c.nudge.x; m[13][tid] = c.nudge.y; m[14][tid] = c.nudge.z;
  m[15][tid] = c.ray_len; m[16][tid] = c.step; m[17][tid] = c.entry_depth; m[18][tid] = c.exit_depth;
}
__device__ __forceinline__ void park_var(const ChainSt& c, float (*m)[kThreads], int tid) {     // after every chain phase
  m[19][tid] = c.cur.x; m[20][tid] = c.cur.y; m[21][tid] = c.cur.z; m[22][tid] = c.t;
  m[23][tid] = c.resume_pos.x; m[24][tid] = c.resume_pos.y; m[25][tid] = c.resume_pos.z; m[26][tid] = c.resume_pos.w;
  m[27][tid] = c.resume_col.x; m[28][tid] = c.resume_col.y; m[29][tid] = c.resume_col.z; m[30][tid] = c.resume_col.w;
  m[31][tid] = __uint_as_float(c.j); m[32][tid] = __uint_as_float(c.lbx); m[33][tid] = __uint_as_float(c.lby);
  m[34][tid] = __uint_as_float(c.lbz); m[35][tid] = __uint_as_float(c.lbl);
  m[36][tid] = __uint_as_float(c.optimal ? 1u : 0u);
}
__device__ __forceinline__ void unpark_result(ChainSt& c, float (*m)[kThreads], int tid) {      // what TerminateRay needs
  c.resume_pos.x = m[23][tid]; c.resume_pos.y = m[24][tid]; c.resume_pos.z = m[25][tid]; c.resume_pos.w = m[26][tid];
  c.resume_col.x = m[27][tid]; c.resume_col.y = m[28][tid]; c.resume_col.z = m[29][tid]; c.resume_col.w = m[30][tid];
  c.optimal = __float_as_uint(m[36][tid]) != 0u;
}
__device__ __forceinline__ void unpark(ChainSt& c, float (*m)[kThreads], int tid) {
  c.entry = F3(m[0][tid], m[1][tid], m[2][tid]);
  c.nexit = F3(m[3][tid], m[4][tid], m[5][tid]);
  c.dir = F3(m[6][tid], m[7][tid], m[8][tid]);
  c.dv = F3(m[9][tid], m[10][tid], m[11][tid]);
  c.nudge = F3(m[12][tid], m[13][tid], m[14][tid]);
  c.ray_len = m[15][tid]; c.step = m[16][tid]; c.entry_depth = m[17][tid]; c.exit_depth = m[18][tid];
  c.cur = F3(m[19][tid], m[20][tid], m[21][tid]); c.t = m[22][tid];
  c.j = __float_as_uint(m[31][tid]); c.lbx = __float_as_uint(m[32][tid]); c.lby = __float_as_uint(m[33][tid]);
  c.lbz = __float_as_uint(m[34][tid]); c.lbl = __float_as_uint(m[35][tid]);
  unpark_result(c, m, tid);
}

// MODE: 0 = 1D TF, 1 = 2D TF, 2 = isosurface
// PIPE (DVR modes): this launch is one STAGE of the depth pipeline: the rank's block is a slab
// of the volume; a ray comes in with the colour the stages in front of it accumulated (start colour / resume position,
// exactly the inputs of a resumed GridLeaper subframe), is marched through the slab, and leaves with its resume
// position at the slab's far side -- or finished (w = 1000) if it terminated early or left the volume.  Early ray
// termination therefore works across ranks as on one GPU.
// TVK_PERSIST (persistent warps that hand the pixels of a tile queue to their idle lanes) was built and MEASURED SLOWER on
// B200 / C3 (profiles/r2x_persist_ab.txt): non-persistent 304 fps; persistent, a tile only when the whole warp is idle
// (TVK_REFILL=32) 293 fps; refill as soon as 8 / 16 / 24 lanes are idle 178 / 185 / 180 fps -- images bit-identical in every
// variant (208 GPU tests).  Lanes that start rays at different times need their brick-chain phases at different turns, and
// the chain phase (page-table walk) runs for the whole warp whenever ONE lane needs it: refilling trades idle lanes for
// many more, thinner chain phases.  Kept as a build switch, off by default.
#ifndef TVK_PERSIST
#define TVK_PERSIST 0
#endif
#ifndef TVK_REFILL
#define TVK_REFILL 8
#endif
constexpr int kRefill = TVK_REFILL;
constexpr int kSMs = 148;
#if TVK_PERSIST
// Persistent warps with lane refill (TVK_PERSIST=1, see the measurement above): the launch is one CTA slot per SM x resident CTAs, and a warp
// does not own one 8x4 tile: it draws tiles from a global counter (the old dispatch order: 2x2 tile groups, centre-out)
// and gives the pixels of the open tile to its IDLE lanes -- a lane whose ray has terminated (early ray termination,
// left the volume) starts the next pixel while its neighbours are still marching, instead of idling until the slowest ray
// of the tile is done (ncu before: 24.4 of 32 lanes active per issued instruction).  Rays, their arithmetic and their
// outputs are unchanged: only WHICH lane traces WHICH pixel WHEN differs.  A warp interrupts its live rays for a refill
// only when kRefill lanes are idle (the set-up runs with the idle lanes only).
template <typename T, int MODE, bool LIT, bool FAST, int BS, bool COUNT, bool PIPE = false>
__global__ void __launch_bounds__(kThreads, TVK_MIN_BLOCKS * 64 / kThreads) raycast_kernel(const __grid_constant__ RayConsts P) {
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr bool ISO = MODE == 2;
  typedef typename PairOf<T>::W W;           // pool element: the x-pair (voxel x, voxel x+1)
  const W* pool = (const W*)P.pool;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  __shared__ float park[kPark ? kChainWords : 1][kThreads];
  __shared__ float seg_f[9][kThreads];       // next segment: pool entry, trans, 1/scale
  __shared__ uint32_t seg_u[COUNT ? 7 : 6][kThreads];    // slot origin (3), slot index, steps, flags (, page-table index)
  // tile queue: 2x2 groups of 8x4 tiles, groups in the centre-out order of the non-persistent launch
  const uint32_t gx = (P.width + 15u) / 16u, gy = (P.height + 7u) / 8u, n_tiles = gx * gy * 4u;
  uint32_t open_tile = 0u, open_mask = 0u;   // warp-uniform: the tile being handed out, its pixels not yet taken
  bool more = true;                          // warp-uniform: the global queue has not run dry
  // ---- state of the lane's current ray
  size_t pix = 0;
  bool have_ray = false;           // a ray has been started and its outputs are not written yet
  bool done = false;
  unsigned long long n_samples = 0, n_bricks = 0;
  unsigned long long n_alive_iters = 0, n_warp_iters = 0;   // lane-utilisation diagnostics (count mode)
  ChainSt c;
  f4 acc = from4(zero4);
  f4& resume_col = c.resume_col;
  f4& resume_pos = c.resume_pos;
  f4 hit_pos = from4(zero4), hit_nrm = from4(zero4), resume_nrm = from4(zero4);
  bool handoff = false;            // PIPE: the ray left this stage's slab alive
  f4 hand_pos = from4(zero4);
  f3 vdir = F3(0.f, 0.f, 0.f);
  const float voxel_size = 0.125f / 2000.0f;
  const f3 dscale = F3(P.domain_scale), eye_m = F3(P.eye_m), la = F3(P.light_a), ld = F3(P.light_d),
           ls = F3(P.light_s), ldir = F3(P.light_dir_m);
  bool ray_live = false;                  // the ray has not terminated (ERT / iso hit)
  bool chain = false;                     // the brick chain has not reached the end of the ray
  bool have_next = false;                 // a prefetched segment waits in shared memory
  int steps_left = 0;      // samples left in the current brick
  bool b_partial = false;  // sort-last: the current brick straddles the shard box (ownership per sample)
  f3 pc = F3(0.f, 0.f, 0.f), b_trans = pc, b_inv = pc;
  uint32_t b_ox = 0, b_oy = 0, b_oz = 0;
  const W* vox = pool;
  constexpr bool GRAD = !ISO && (MODE == 1 || LIT);   // what a sample needs: the 7-tap footprint or the centre tap
  Foot<T, FAST, BS, GRAD> cur;   // footprint of the sample at pc, fetched one turn ahead
  bool cur_ok = false;
  unsigned long long pend = 0;   // COUNT: brick visits of the chain that the sampling has not reached yet

  for (;;) {
    // ---- refill: finished rays are written out, idle lanes take the next pixels
    const unsigned full = 0xffffffffu;
    const unsigned idle_m = __ballot_sync(full, !ray_live);
    if (idle_m == full || ((more || open_mask != 0u) && __popc(idle_m) >= kRefill)) {
      if (!ray_live && have_ray) {
        have_ray = false;
        if (!done) {
          if (kPark) unpark_result(c, park, tid);
          // TerminateRay
          if (!ISO) {
            if (c.optimal) {
              // ray_live is false only after early termination; a ray that ran out of bricks in this slab is handed on
              if (PIPE && handoff && !(acc.w > 0.99f)) { resume_pos = hand_pos; resume_col = acc; }
              else { resume_pos.w = 1000.0f; resume_col = acc; }
            }
          } else {
            if (c.optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
            resume_nrm = hit_nrm;
          }
        }
        if (!ISO) {
          P.out0[pix] = to4(acc); P.out1[pix] = to4(resume_col); P.out2[pix] = to4(resume_pos);
        } else {
          P.out0[pix] = to4(hit_pos); P.out1[pix] = to4(hit_nrm); P.out2[pix] = to4(resume_pos);
          P.out3[pix] = to4(resume_nrm);
        }
        if (COUNT) {
          atomicAdd(P.counters + 0, n_samples); atomicAdd(P.counters + 1, 1ull); atomicAdd(P.counters + 2, n_bricks);
          atomicAdd(P.counters + 3, n_alive_iters); atomicAdd(P.counters + 4, n_warp_iters);
          atomicMax(P.counters + 5, n_alive_iters);
        }
      }
      // the r-th waiting lane takes the r-th open pixel of the tile
      unsigned want = idle_m;
      bool got = false;
      uint32_t px = 0, py = 0;
      while (want != 0u) {
        if (open_mask == 0u) {
          if (!more) break;
          uint32_t t = 0;
          if (lane == 0) t = atomicAdd(P.tile_counter, 1u);
          t = __shfl_sync(full, t, 0);
          if (t >= n_tiles) { more = false; break; }
          open_tile = t; open_mask = full;
        }
        const int n_take = min(__popc(open_mask), __popc(want));
        const int r = __popc(want & ((1u << lane) - 1u));
        const bool take = ((want >> lane) & 1u) != 0u && r < n_take;
        uint32_t bit = 0;
        if (take) {
          bit = __fns(open_mask, 0, r + 1);
          const uint32_t g = open_tile >> 2, sub = open_tile & 3u;
          const uint32_t sx = g % gx, sy = g / gx;
          const uint32_t tx = (kCentreOut ? centre_out(sx, gx) : sx) * 2u + (sub & 1u);
          const uint32_t ty = (kCentreOut ? centre_out(sy, gy) : sy) * 2u + (sub >> 1);
          px = tx * 8u + (bit & 7u); py = ty * 4u + (bit >> 3);
          got = true;
        }
        open_mask &= ~__reduce_or_sync(full, take ? (1u << bit) : 0u);
        want &= ~__ballot_sync(full, take);
      }
      if (got && px < P.width && py < P.height) {
        pix = (size_t)py * P.width + px;
        f4 entry4, exit4;
        const bool covered = ray_setup(P, px, py, entry4, exit4, !PIPE);   // a stage takes every ray that meets the VOLUME
        if (!covered) {   // render targets are cleared where no back face is rasterised (GLGridLeaper.cpp:837)
          P.out0[pix] = zero4; P.out1[pix] = zero4; P.out2[pix] = PIPE ? make_float4(0.f, 0.f, 0.f, 1000.0f) : zero4;
          if (ISO) P.out3[pix] = zero4;
        } else {
          have_ray = true; done = false; handoff = false;
          hand_pos = from4(zero4); hit_pos = from4(zero4); hit_nrm = from4(zero4); resume_nrm = from4(zero4);
          n_samples = 0; n_bricks = 0; n_alive_iters = 0; n_warp_iters = 0; pend = 0;
          if (P.first_pass) {
            resume_pos = entry4;
            acc = from4(zero4);
          } else {
            resume_pos = from4(P.ray_start[pix]);
            acc = from4(P.start_color[pix]);
          }
          if (!ISO) {
            resume_col = acc;
            if (resume_pos.w == 1000.0f) done = true;
          } else {
            if (floorf(resume_pos.w) == 1000.0f) done = true;
            else if (floorf(resume_pos.w) == 500.0f) {
              hit_pos = xform4(P.m2e, resume_pos.x, resume_pos.y, resume_pos.z, 1.0f);
              hit_pos.w = resume_pos.w - floorf(resume_pos.w) + 1.0f;
              hit_nrm = acc;   // rayStartNormal
              resume_nrm = hit_nrm;
              done = true;
            }
          }
          if (!done) {
            c.entry = F3(resume_pos.x, resume_pos.y, resume_pos.z);
            c.entry_depth = resume_pos.w;
            c.nexit = F3(exit4.x, exit4.y, exit4.z);
            c.exit_depth = exit4.w;
            c.dir = sub3(c.nexit, c.entry);
            c.ray_len = len3(c.dir);
            // TransformToPoolSpace
            const f3 ps = F3(P.pool_size_f);
            vdir = norm3(mul3(c.dir, F3(P.vol_f)));
            vdir = div3(vdir, ps);
            const float den = 2.0f * P.sample_rate;
            vdir = F3(vdir.x / den, vdir.y / den, vdir.z / den);
            c.step = len3(vdir);
            c.t = 0.0f;
            c.optimal = true;
            c.cur = c.entry;
            c.j = 0;   // bricks visited by the chain (the shader's j < 100 bound)
            c.lbx = 0; c.lby = 0; c.lbz = 0; c.lbl = 9999;
        
            c.dv = F3(1.0f / c.dir.x, 1.0f / c.dir.y, 1.0f / c.dir.z);   // BrickExit's 1.0/dir
            // the empty-brick advance voxelSize*direction/rayLength
            c.nudge = F3(voxel_size * c.dir.x / c.ray_len, voxel_size * c.dir.y / c.ray_len, voxel_size * c.dir.z / c.ray_len);
            ray_live = c.ray_len > voxel_size;
            chain = ray_live; have_next = false; steps_left = 0; b_partial = false; cur_ok = false;
            if (kPark) { park_const(c, park, tid); park_var(c, park, tid); }
          }
        }
      }
      if (__ballot_sync(full, ray_live || have_ray) == 0u && !more && open_mask == 0u) break;
    }
    if (ray_live) {
      // ---- chain phase: runs for the whole warp when some lane can neither sample nor pick up a segment
      const unsigned act = __activemask();
      const bool need = steps_left == 0 && !have_next && chain;
      if (__ballot_sync(act, need) != 0u && chain && !have_next) {
        if (kPark) unpark(c, park, tid);
#pragma unroll 1
        for (int f = 0; f < 4 && chain && !have_next; f++) {
          if (c.j >= 100) { chain = false; break; }
          if (P.shard) {   // the block is convex: once the ray has left it there is nothing more to do on this rank
            const bool gone = (c.dir.x > 0.0f && c.cur.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && c.cur.x <= P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && c.cur.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && c.cur.y <= P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && c.cur.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && c.cur.z <= P.sh_lo[2]);
            if (gone) {
              if (PIPE) {   // where the next stage picks the ray up
                handoff = true;
                hand_pos.x = c.cur.x; hand_pos.y = c.cur.y; hand_pos.z = c.cur.z;
                hand_pos.w = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
              }
              chain = false; break;
            }
          }
          const float cur_depth = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
          uint32_t lod = compute_lod(P, cur_depth);
          BrickRef b;
          int ok;
          if (steps_left > 0) {   // look-ahead: the ray is still sampling the previous segment
            ok = get_brick<true>(P, c.cur, lod, c.dir, c.dv, b);
            if (ok < 0) break;    // missing brick: handled when the ray stands here
          } else {
            ok = get_brick<false>(P, c.cur, lod, c.dir, c.dv, b);
            if (!ok && c.optimal) {
              c.optimal = false;
              resume_pos.x = c.cur.x; resume_pos.y = c.cur.y; resume_pos.z = c.cur.z; resume_pos.w = cur_depth;
              if (!ISO) resume_col = acc;
            }
          }
          if (COUNT) pend++;
          if (!b.empty && !(c.lbx == b.bx && c.lby == b.by && c.lbz == b.bz && c.lbl == b.bl)) {
            int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / c.step);
            const int s2 = (int)ceilf(len3(mul3(sub3(c.nexit, c.cur), b.scale)) / c.step);
            steps = min(steps, s2);
            const f3 inv = F3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
            f3 pe = b.pool_entry;
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
            if (b.where == OUTSIDE_SHARD) {   // another rank's brick: advance by its steps, take no sample
              const float n = (float)max(steps, 0);
              pe = F3(fmaf(n, vdir.x, pe.x), fmaf(n, vdir.y, pe.y), fmaf(n, vdir.z, pe.z));
              steps = 0;
            }
            if (steps > 0) {
              seg_f[0][tid] = pe.x; seg_f[1][tid] = pe.y; seg_f[2][tid] = pe.z;
              seg_f[3][tid] = b.trans.x; seg_f[4][tid] = b.trans.y; seg_f[5][tid] = b.trans.z;
              seg_f[6][tid] = inv.x; seg_f[7][tid] = inv.y; seg_f[8][tid] = inv.z;
              seg_u[0][tid] = b.ox; seg_u[1][tid] = b.oy; seg_u[2][tid] = b.oz;
              seg_u[3][tid] = b.slot; seg_u[4][tid] = (uint32_t)steps;
              seg_u[5][tid] = b.where == PARTLY_IN_SHARD ? 1u : 0u;
              if (COUNT) seg_u[6][tid] = b.id;
              have_next = true;
              // where the sample loop will leave pc: `steps` SEQUENTIAL adds (fp32 addition is not associative, so
              // there is no closed form); 4x unrolled -- this loop was 6.5 % of the launch's warp instructions
#pragma unroll 4
              for (int i = 0; i < steps; i++) pe = add3(pe, vdir);
            }
            c.cur = mul3(sub3(pe, b.trans), inv);
          } else {
            c.cur = add3(b.norm_exit, c.nudge);
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
          }
          c.t = len3(sub3(c.entry, b.norm_exit)) / c.ray_len;
          c.j++;
          if (c.t > 0.9999f) chain = false;
        }
        if (kPark) park_var(c, park, tid);
      }
      // ---- pick up the waiting segment
      if (steps_left == 0 && have_next) {
        pc = F3(seg_f[0][tid], seg_f[1][tid], seg_f[2][tid]);
        b_trans = F3(seg_f[3][tid], seg_f[4][tid], seg_f[5][tid]);
        b_inv = F3(seg_f[6][tid], seg_f[7][tid], seg_f[8][tid]);
        b_ox = seg_u[0][tid]; b_oy = seg_u[1][tid]; b_oz = seg_u[2][tid];
        vox = pool + (uint64_t)seg_u[3][tid] * P.slot_voxels;
        steps_left = (int)seg_u[4][tid];
        b_partial = seg_u[5][tid] != 0u;
        have_next = false;
        cur_ok = false;
        if (COUNT) {
          n_bricks += pend; pend = 0;
          if (P.visited) { const uint32_t id = seg_u[6][tid]; atomicOr(P.visited + (id >> 5), 1u << (id & 31)); }
        }
      }
      if (steps_left == 0 && !chain) {   // the ray left the volume (or its 100-brick budget) unterminated
        if (COUNT) n_bricks += pend;
        ray_live = false;
      }
      if (COUNT) {
        n_alive_iters += ray_live ? 1 : 0;
        if (__ffs(__activemask()) - 1 == (tid & 31)) n_warp_iters++;
      }
      // ---- sample phase: one sample for every lane that is inside a brick ----
      if (ray_live && steps_left > 0) {
        bool terminated = false;
        // software pipeline: the footprint of THIS sample was fetched during the previous turn (cur_ok), the one of the
        // next sample of the brick is fetched now, before this sample's arithmetic, so its load latency is covered
        if (!kPrefetch || !FAST || !cur_ok) cur.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
        const f3 pc_next = add3(pc, vdir);
        const bool pf = kPrefetch && FAST && steps_left > 1;
        Foot<T, FAST, BS, GRAD> nxt;
        if (pf) nxt.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc_next);
        bool mine = true;
        if (b_partial) {
          const f3 mq = mul3(sub3(pc, b_trans), b_inv);
          mine = mq.x >= P.sh_lo[0] && mq.x < P.sh_hi[0] && mq.y >= P.sh_lo[1] && mq.y < P.sh_hi[1] &&
                 mq.z >= P.sh_lo[2] && mq.z < P.sh_hi[2];
          if (PIPE && !ISO && !mine) {
            // a brick of a coarser LoD straddles the slab's far side: the ray is handed on AT the side, not
            // behind the brick, so the next stage takes the brick's remaining samples
            const bool gone = (c.dir.x > 0.0f && mq.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && mq.x < P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && mq.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && mq.y < P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && mq.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && mq.z < P.sh_lo[2]);
            if (gone) {
              const float tq = len3(sub3(mq, c.entry)) / c.ray_len;
              handoff = true;
              hand_pos.x = mq.x; hand_pos.y = mq.y; hand_pos.z = mq.z;
              hand_pos.w = c.entry_depth * (1.0f - tq) + c.exit_depth * tq;
              terminated = true;   // leaves the loop; TerminateRay sees alpha <= 0.99 and hands the ray on
            }
          }
        }
        if constexpr (!ISO) {
          if (mine) {
            if (COUNT) n_samples++;
            // ComputeColorFromVolume + OpacityCorrectColor at pool position pc
            f4 col;
            bool clear = false;
            if constexpr (MODE == 0 && !LIT) {
              col = tf_lookup(P, cur.centre(P) * P.trans_scale, 0.0f);
            } else if constexpr (MODE == 0) {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              col = tf_lookup(P, data * P.trans_scale, 0.0f);
              // A sample whose transfer-function alpha is exactly 0 leaves the ray unchanged bit for bit
              // (UnderCompositing adds colour * (1-a) * 0 = 0 to every channel; table colours and lit colours are
              // finite, and opacity correction maps 0 to 0), so its normal and lighting are not computed when the
              // whole warp agrees.  The shader cannot branch this cheaply; the result is identical.
              clear = kSkipClear && col.w == 0.0f;
              if (!clear) {
                f3 n = mul3(g, dscale);   // ComputeNormal
                const float l = len3(n);
                if (l > 0.0f) n = scl3(n, 1.0f / l);
                const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                col.x = lit.x; col.y = lit.y; col.z = lit.z;
              }
            } else {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              const float gm = len3(g);
              col = tf_lookup(P, data * P.trans_scale, 1.0f - gm * P.gradient_scale);
              if (LIT) {
                clear = kSkipClear && col.w == 0.0f;
                if (!clear) {
                  const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                  const f3 n = mul3(dscale, gn);
                  const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                  float dl, sp;
                  light_terms(eye_m, mp, n, ldir, dl, sp);
                  const f3 lit = light_apply(la, mul3(F3(col.x, col.y, col.z), ld), ls, dl, sp);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              }
            }
            if (!clear) {
              col.w = opacity_correct(P, col.w);
              // UnderCompositing
              const float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) terminated = true;
            }
          }
        } else if (mine) {   // isosurface march
          if (COUNT) n_samples++;
          if (cur.centre(P) >= P.isoval) {
            // RefineIsosurface
            f3 rd = F3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
            pc = sub3(pc, rd);
            Foot<T, FAST, BS, false> rf;
#pragma unroll 1
            for (int k = 0; k < 5; k++) {
              rd = F3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
              rf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
              if (rf.centre(P) >= P.isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
            }
            const f3 hp = mul3(sub3(pc, b_trans), b_inv);
            hit_pos = xform4(P.m2e, hp.x, hp.y, hp.z, 1.0f);
            hit_pos.w = 1.0f + 1.0f;   // color.r + 1
            Foot<T, FAST, BS, true> gf;
            gf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
            float dummy; f3 g;
            gf.sample_with_gradient(P, dummy, g);
            f3 n = mul3(g, dscale);
            const float l = len3(n);
            if (l > 0.0f) n = scl3(n, 1.0f / l);
            const float* m = P.mv_inv;   // mModelViewIT * vec4(n, 0)
            hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
            hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
            hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
            hit_nrm.w = floorf(1.0f * 512.0f) + 1.0f;   // floor(color.g*512)+color.b
            terminated = true;
          } else {
            hit_pos = from4(zero4);
          }
        }
        steps_left -= 1;
        if (terminated) ray_live = false;
        else {
          pc = pc_next;
          if (pf) cur = nxt;
          cur_ok = pf;
        }
      }
    }
  }
}

#else
template <typename T, int MODE, bool LIT, bool FAST, int BS, bool COUNT, bool PIPE = false>
__global__ void __launch_bounds__(kThreads, TVK_MIN_BLOCKS * 64 / kThreads) raycast_kernel(const __grid_constant__ RayConsts P) {
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  // CTAs are dispatched in blockIdx order (x fastest).  The rays through the middle of the volume are the
  // longest: tile rows / columns are mapped centre-out so those start first (measured: +1 % on C3, the launch
  // is bound by per-warp latency, not by its tail; kept because it costs nothing).
  const uint32_t tx = kCentreOut ? centre_out(blockIdx.x, gridDim.x) : blockIdx.x;
  const uint32_t ty = kCentreOut ? centre_out(blockIdx.y, gridDim.y) : blockIdx.y;
  const uint32_t px = tx * (8 * kWX) + (wid % kWX) * 8 + (lane & 7);
  const uint32_t py = ty * (4 * kWY) + (wid / kWX) * 4 + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const size_t pix = (size_t)py * P.width + px;
  constexpr bool ISO = MODE == 2;
  typedef typename PairOf<T>::W W;           // pool element: the x-pair (voxel x, voxel x+1)
  const W* pool = (const W*)P.pool;
  unsigned long long n_samples = 0, n_bricks = 0;
  unsigned long long n_alive_iters = 0, n_warp_iters = 0;   // lane-utilisation diagnostics (count mode)

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  f4 entry4, exit4;
  const bool covered = ray_setup(P, px, py, entry4, exit4, !PIPE);   // a stage takes every ray that meets the VOLUME
  if (!covered) {   // render targets are cleared where no back face is rasterised (GLGridLeaper.cpp:837)
    P.out0[pix] = zero4; P.out1[pix] = zero4; P.out2[pix] = PIPE ? make_float4(0.f, 0.f, 0.f, 1000.0f) : zero4;
    if (ISO) P.out3[pix] = zero4;
    return;
  }
  __shared__ float park[kPark ? kChainWords : 1][kThreads];
  ChainSt c;
  f4 acc;
  f4& resume_col = c.resume_col;
  f4& resume_pos = c.resume_pos;
  f4 hit_pos = from4(zero4), hit_nrm = from4(zero4), resume_nrm = from4(zero4);
  bool done = false;
  bool handoff = false;            // PIPE: the ray left this stage's slab alive
  f4 hand_pos = from4(zero4);
  if (P.first_pass) {
    resume_pos = entry4;
    acc = from4(zero4);
  } else {
    resume_pos = from4(P.ray_start[pix]);
    acc = from4(P.start_color[pix]);
  }
  if (!ISO) {
    resume_col = acc;
    if (resume_pos.w == 1000.0f) done = true;
  } else {
    if (floorf(resume_pos.w) == 1000.0f) done = true;
    else if (floorf(resume_pos.w) == 500.0f) {
      hit_pos = xform4(P.m2e, resume_pos.x, resume_pos.y, resume_pos.z, 1.0f);
      hit_pos.w = resume_pos.w - floorf(resume_pos.w) + 1.0f;
      hit_nrm = acc;   // rayStartNormal
      resume_nrm = hit_nrm;
      done = true;
    }
  }
  if (!done) {
    c.entry = F3(resume_pos.x, resume_pos.y, resume_pos.z);
    c.entry_depth = resume_pos.w;
    c.nexit = F3(exit4.x, exit4.y, exit4.z);
    c.exit_depth = exit4.w;
    c.dir = sub3(c.nexit, c.entry);
    c.ray_len = len3(c.dir);
    // TransformToPoolSpace
    const f3 ps = F3(P.pool_size_f);
    f3 vdir = norm3(mul3(c.dir, F3(P.vol_f)));
    vdir = div3(vdir, ps);
    const float den = 2.0f * P.sample_rate;
    vdir = F3(vdir.x / den, vdir.y / den, vdir.z / den);
    c.step = len3(vdir);
    c.t = 0.0f;
    c.optimal = true;
    const float voxel_size = 0.125f / 2000.0f;
    c.cur = c.entry;
    c.j = 0;   // bricks visited by the chain (the shader's j < 100 bound)
    c.lbx = 0; c.lby = 0; c.lbz = 0; c.lbl = 9999;
    const f3 dscale = F3(P.domain_scale), eye_m = F3(P.eye_m), la = F3(P.light_a), ld = F3(P.light_d),
             ls = F3(P.light_s), ldir = F3(P.light_dir_m);

    c.dv = F3(1.0f / c.dir.x, 1.0f / c.dir.y, 1.0f / c.dir.z);   // BrickExit's 1.0/dir
    // the empty-brick advance voxelSize*direction/rayLength
    c.nudge = F3(voxel_size * c.dir.x / c.ray_len, voxel_size * c.dir.y / c.ray_len, voxel_size * c.dir.z / c.ray_len);
    // ---- flat loop state ---------------------------------------------------------------------
    // The shader's nested loops (bricks along the ray / samples inside a brick) run as ONE per-warp loop with
    // two decoupled parts: a brick CHAIN that walks the page table one segment ahead of the sampling (the next
    // segment waits in shared memory), and the SAMPLE phase.  A chain step needs no voxel data -- the position
    // after a brick is its entry + steps sequential adds of the step vector, repeated here exactly as the
    // sample loop does -- so all lanes of a warp can look ahead TOGETHER whenever one lane runs dry, and a lane
    // that finishes its brick early just picks up its waiting segment instead of idling until the slowest lane
    // is done.  Look-ahead never has side effects: a missing brick is only processed (reported, resume state)
    // when the ray really stands at it, so reports, resume points and counters are the shader's.
    __shared__ float seg_f[9][kThreads];       // next segment: pool entry, trans, 1/scale
    __shared__ uint32_t seg_u[COUNT ? 7 : 6][kThreads];    // slot origin (3), slot index, steps, flags (, page-table index)
    bool ray_live = c.ray_len > voxel_size;   // the ray has not terminated (ERT / iso hit)
    bool chain = ray_live;                  // the brick chain has not reached the end of the ray
    bool have_next = false;                 // a prefetched segment waits in shared memory
    int steps_left = 0;      // samples left in the current brick
    bool b_partial = false;  // sort-last: the current brick straddles the shard box (ownership per sample)
    f3 pc = c.entry, b_trans = c.entry, b_inv = c.entry;
    uint32_t b_ox = 0, b_oy = 0, b_oz = 0;
    const W* vox = pool;
    constexpr bool GRAD = !ISO && (MODE == 1 || LIT);   // what a sample needs: the 7-tap footprint or the centre tap
    Foot<T, FAST, BS, GRAD> cur;   // footprint of the sample at pc, fetched one turn ahead
    bool cur_ok = false;
    unsigned long long pend = 0;   // COUNT: brick visits of the chain that the sampling has not reached yet

    if (kPark) { park_const(c, park, tid); park_var(c, park, tid); }
    while (ray_live) {
      // ---- chain phase: runs for the whole warp when some lane can neither sample nor pick up a segment
      const unsigned act = __activemask();
      const bool need = steps_left == 0 && !have_next && chain;
      if (__ballot_sync(act, need) != 0u && chain && !have_next) {
        if (kPark) unpark(c, park, tid);
#pragma unroll 1
        for (int f = 0; f < 4 && chain && !have_next; f++) {
          if (c.j >= 100) { chain = false; break; }
          if (P.shard) {   // the block is convex: once the ray has left it there is nothing more to do on this rank
            const bool gone = (c.dir.x > 0.0f && c.cur.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && c.cur.x <= P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && c.cur.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && c.cur.y <= P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && c.cur.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && c.cur.z <= P.sh_lo[2]);
            if (gone) {
              if (PIPE) {   // where the next stage picks the ray up
                handoff = true;
                hand_pos.x = c.cur.x; hand_pos.y = c.cur.y; hand_pos.z = c.cur.z;
                hand_pos.w = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
              }
              chain = false; break;
            }
          }
          const float cur_depth = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
          uint32_t lod = compute_lod(P, cur_depth);
          BrickRef b;
          int ok;
          if (steps_left > 0) {   // look-ahead: the ray is still sampling the previous segment
            ok = get_brick<true>(P, c.cur, lod, c.dir, c.dv, b);
            if (ok < 0) break;    // missing brick: handled when the ray stands here
          } else {
            ok = get_brick<false>(P, c.cur, lod, c.dir, c.dv, b);
            if (!ok && c.optimal) {
              c.optimal = false;
              resume_pos.x = c.cur.x; resume_pos.y = c.cur.y; resume_pos.z = c.cur.z; resume_pos.w = cur_depth;
              if (!ISO) resume_col = acc;
            }
          }
          if (COUNT) pend++;
          if (!b.empty && !(c.lbx == b.bx && c.lby == b.by && c.lbz == b.bz && c.lbl == b.bl)) {
            int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / c.step);
            const int s2 = (int)ceilf(len3(mul3(sub3(c.nexit, c.cur), b.scale)) / c.step);
            steps = min(steps, s2);
            const f3 inv = F3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
            f3 pe = b.pool_entry;
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
            if (b.where == OUTSIDE_SHARD) {   // another rank's brick: advance by its steps, take no sample
              const float n = (float)max(steps, 0);
              pe = F3(fmaf(n, vdir.x, pe.x), fmaf(n, vdir.y, pe.y), fmaf(n, vdir.z, pe.z));
              steps = 0;
            }
            if (steps > 0) {
              seg_f[0][tid] = pe.x; seg_f[1][tid] = pe.y; seg_f[2][tid] = pe.z;
              seg_f[3][tid] = b.trans.x; seg_f[4][tid] = b.trans.y; seg_f[5][tid] = b.trans.z;
              seg_f[6][tid] = inv.x; seg_f[7][tid] = inv.y; seg_f[8][tid] = inv.z;
              seg_u[0][tid] = b.ox; seg_u[1][tid] = b.oy; seg_u[2][tid] = b.oz;
              seg_u[3][tid] = b.slot; seg_u[4][tid] = (uint32_t)steps;
              seg_u[5][tid] = b.where == PARTLY_IN_SHARD ? 1u : 0u;
              if (COUNT) seg_u[6][tid] = b.id;
              have_next = true;
              // where the sample loop will leave pc: `steps` SEQUENTIAL adds (fp32 addition is not associative, so
              // there is no closed form); 4x unrolled -- this loop was 6.5 % of the launch's warp instructions
#pragma unroll 4
              for (int i = 0; i < steps; i++) pe = add3(pe, vdir);
            }
            c.cur = mul3(sub3(pe, b.trans), inv);
          } else {
            c.cur = add3(b.norm_exit, c.nudge);
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
          }
          c.t = len3(sub3(c.entry, b.norm_exit)) / c.ray_len;
          c.j++;
          if (c.t > 0.9999f) chain = false;
        }
        if (kPark) park_var(c, park, tid);
      }
      // ---- pick up the waiting segment
      if (steps_left == 0 && have_next) {
        pc = F3(seg_f[0][tid], seg_f[1][tid], seg_f[2][tid]);
        b_trans = F3(seg_f[3][tid], seg_f[4][tid], seg_f[5][tid]);
        b_inv = F3(seg_f[6][tid], seg_f[7][tid], seg_f[8][tid]);
        b_ox = seg_u[0][tid]; b_oy = seg_u[1][tid]; b_oz = seg_u[2][tid];
        vox = pool + (uint64_t)seg_u[3][tid] * P.slot_voxels;
        steps_left = (int)seg_u[4][tid];
        b_partial = seg_u[5][tid] != 0u;
        have_next = false;
        cur_ok = false;
        if (COUNT) {
          n_bricks += pend; pend = 0;
          if (P.visited) { const uint32_t id = seg_u[6][tid]; atomicOr(P.visited + (id >> 5), 1u << (id & 31)); }
        }
      }
      if (steps_left == 0 && !chain) {   // the ray left the volume (or its 100-brick budget) unterminated
        if (COUNT) n_bricks += pend;
        ray_live = false;
      }
      if (COUNT) {
        n_alive_iters += ray_live ? 1 : 0;
        if (__ffs(__activemask()) - 1 == (tid & 31)) n_warp_iters++;
      }
      // ---- sample phase: one sample for every lane that is inside a brick ----
      if (ray_live && steps_left > 0) {
        bool terminated = false;
        // software pipeline: the footprint of THIS sample was fetched during the previous turn (cur_ok), the one of the
        // next sample of the brick is fetched now, before this sample's arithmetic, so its load latency is covered
        if (!kPrefetch || !FAST || !cur_ok) cur.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
        const f3 pc_next = add3(pc, vdir);
        const bool pf = kPrefetch && FAST && steps_left > 1;
        Foot<T, FAST, BS, GRAD> nxt;
        if (pf) nxt.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc_next);
        bool mine = true;
        if (b_partial) {
          const f3 mq = mul3(sub3(pc, b_trans), b_inv);
          mine = mq.x >= P.sh_lo[0] && mq.x < P.sh_hi[0] && mq.y >= P.sh_lo[1] && mq.y < P.sh_hi[1] &&
                 mq.z >= P.sh_lo[2] && mq.z < P.sh_hi[2];
          if (PIPE && !ISO && !mine) {
            // a brick of a coarser LoD straddles the slab's far side: the ray is handed on AT the side, not
            // behind the brick, so the next stage takes the brick's remaining samples
            const bool gone = (c.dir.x > 0.0f && mq.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && mq.x < P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && mq.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && mq.y < P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && mq.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && mq.z < P.sh_lo[2]);
            if (gone) {
              const float tq = len3(sub3(mq, c.entry)) / c.ray_len;
              handoff = true;
              hand_pos.x = mq.x; hand_pos.y = mq.y; hand_pos.z = mq.z;
              hand_pos.w = c.entry_depth * (1.0f - tq) + c.exit_depth * tq;
              terminated = true;   // leaves the loop; TerminateRay sees alpha <= 0.99 and hands the ray on
            }
          }
        }
        if constexpr (!ISO) {
          if (mine) {
            if (COUNT) n_samples++;
            // ComputeColorFromVolume + OpacityCorrectColor at pool position pc
            f4 col;
            bool clear = false;
            if constexpr (MODE == 0 && !LIT) {
              col = tf_lookup(P, cur.centre(P) * P.trans_scale, 0.0f);
            } else if constexpr (MODE == 0) {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              col = tf_lookup(P, data * P.trans_scale, 0.0f);
              // A sample whose transfer-function alpha is exactly 0 leaves the ray unchanged bit for bit
              // (UnderCompositing adds colour * (1-a) * 0 = 0 to every channel; table colours and lit colours are
              // finite, and opacity correction maps 0 to 0), so its normal and lighting are not computed when the
              // whole warp agrees.  The shader cannot branch this cheaply; the result is identical.
              clear = kSkipClear && col.w == 0.0f;
              if (!clear) {
                f3 n = mul3(g, dscale);   // ComputeNormal
                const float l = len3(n);
                if (l > 0.0f) n = scl3(n, 1.0f / l);
                const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                col.x = lit.x; col.y = lit.y; col.z = lit.z;
              }
            } else {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              const float gm = len3(g);
              col = tf_lookup(P, data * P.trans_scale, 1.0f - gm * P.gradient_scale);
              if (LIT) {
                clear = kSkipClear && col.w == 0.0f;
                if (!clear) {
                  const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                  const f3 n = mul3(dscale, gn);
                  const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                  float dl, sp;
                  light_terms(eye_m, mp, n, ldir, dl, sp);
                  const f3 lit = light_apply(la, mul3(F3(col.x, col.y, col.z), ld), ls, dl, sp);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              }
            }
            if (!clear) {
              col.w = opacity_correct(P, col.w);
              // UnderCompositing
              const float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) terminated = true;
            }
          }
        } else if (mine) {   // isosurface march
          if (COUNT) n_samples++;
          if (cur.centre(P) >= P.isoval) {
            // RefineIsosurface
            f3 rd = F3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
            pc = sub3(pc, rd);
            Foot<T, FAST, BS, false> rf;
#pragma unroll 1
            for (int k = 0; k < 5; k++) {
              rd = F3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
              rf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
              if (rf.centre(P) >= P.isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
            }
            const f3 hp = mul3(sub3(pc, b_trans), b_inv);
            hit_pos = xform4(P.m2e, hp.x, hp.y, hp.z, 1.0f);
            hit_pos.w = 1.0f + 1.0f;   // color.r + 1
            Foot<T, FAST, BS, true> gf;
            gf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
            float dummy; f3 g;
            gf.sample_with_gradient(P, dummy, g);
            f3 n = mul3(g, dscale);
            const float l = len3(n);
            if (l > 0.0f) n = scl3(n, 1.0f / l);
            const float* m = P.mv_inv;   // mModelViewIT * vec4(n, 0)
            hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
            hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
            hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
            hit_nrm.w = floorf(1.0f * 512.0f) + 1.0f;   // floor(color.g*512)+color.b
            terminated = true;
          } else {
            hit_pos = from4(zero4);
          }
        }
        steps_left -= 1;
        if (terminated) ray_live = false;
        else {
          pc = pc_next;
          if (pf) cur = nxt;
          cur_ok = pf;
        }
      }
    }
    if (kPark) unpark_result(c, park, tid);
    // TerminateRay
    if (!ISO) {
      if (c.optimal) {
        // ray_live is false only after early termination; a ray that ran out of bricks in this slab is handed on
        if (PIPE && handoff && !(acc.w > 0.99f)) { resume_pos = hand_pos; resume_col = acc; }
        else { resume_pos.w = 1000.0f; resume_col = acc; }
      }
    } else {
      if (c.optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
      resume_nrm = hit_nrm;
    }
  }
  if (!ISO) {
    P.out0[pix] = to4(acc); P.out1[pix] = to4(resume_col); P.out2[pix] = to4(resume_pos);
  } else {
    P.out0[pix] = to4(hit_pos); P.out1[pix] = to4(hit_nrm); P.out2[pix] = to4(resume_pos);
    P.out3[pix] = to4(resume_nrm);
  }
  if (COUNT) {
    atomicAdd(P.counters + 0, n_samples); atomicAdd(P.counters + 1, 1ull); atomicAdd(P.counters + 2, n_bricks);
    atomicAdd(P.counters + 3, n_alive_iters); atomicAdd(P.counters + 4, n_warp_iters);
    atomicMax(P.counters + 5, n_alive_iters);
  }
}

#endif

// ---- fetch-path ceiling ------------------------------------------------------------------------------------------
// What the traversal kernel's OWN fetch path can deliver when nothing else is in the loop: the same warp tiles (8x4
// rays, one voxel apart), the same 0.5-voxel steps, the same FastFoot loads + packed filter trees on the resident pool,
// but no page-table walk, classification, shading or compositing.  Rays march through a slot along `dir` and hop to
// another slot when they leave it (so the working set is the whole pool, far larger than L2).  bench.py divides the
// kernel's sample rate by this rate: the fetch fraction of the roofline object.
template <typename T, int BS, bool GRAD>
__global__ void __launch_bounds__(kThreads) fetch_probe_kernel(const __grid_constant__ RayConsts P, uint32_t n_slots, uint32_t steps,
                                                               float dx, float dy, float dz, float* out) {
  typedef typename PairOf<T>::W W;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t px = blockIdx.x * (8 * kWX) + (wid % kWX) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * (4 * kWY) + (wid / kWX) * 4 + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const W* pool = (const W*)P.pool;
  const uint32_t tile = (blockIdx.y * gridDim.x + blockIdx.x) * (kWX * kWY) + wid;
  uint32_t slot = (tile * 2654435761u) % n_slots;
  const float bs = (float)(BS ? BS : (int)P.total[0]);
  // lane offsets across the ray bundle: perpendicular to the march direction's dominant axis
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  f3 u, v;
  if (az >= ax && az >= ay) { u = F3(1.f, 0.f, 0.f); v = F3(0.f, 1.f, 0.f); }
  else if (ay >= ax) { u = F3(1.f, 0.f, 0.f); v = F3(0.f, 0.f, 1.f); }
  else { u = F3(0.f, 1.f, 0.f); v = F3(0.f, 0.f, 1.f); }
  const float lu = (float)(lane & 7), lv = (float)(lane >> 3);
  const float inv = 0.5f / sqrtf(dx * dx + dy * dy + dz * dz);
  const f3 step = F3(dx * inv, dy * inv, dz * inv);                  // 0.5 voxel per step
  const f3 start = F3(4.f + u.x * lu + v.x * lv, 4.f + u.y * lu + v.y * lv, 4.f + u.z * lu + v.z * lv);
  f3 q = start;                                                       // voxel coordinates inside the slot
  float acc = 0.0f;
  for (uint32_t i = 0; i < steps; i++) {
    if (q.x < 2.f || q.y < 2.f || q.z < 2.f || q.x > bs - 3.f || q.y > bs - 3.f || q.z > bs - 3.f) {
      q = start;                                                      // left the brick: next slot (warp-coherent hop)
      slot = (slot + 9973u) % n_slots;
    }
    const W* vox = pool + (uint64_t)slot * P.slot_voxels;
    FastFoot<T, BS, GRAD> f;
    f.fetch(P, vox, 0u, 0u, 0u, F3(q.x / bs, q.y / bs, q.z / bs));
    if constexpr (GRAD) {
      float data; f3 g;
      f.sample_with_gradient(P, data, g);
      acc += data + g.x + g.y + g.z;
    } else {
      acc += f.centre(P);
    }
    q = add3(q, step);
  }
  out[(size_t)py * P.width + px] = acc;
}

template <typename T, int MODE, bool LIT>
void launch_t(const RayConsts& rc, cudaStream_t s) {
  dim3 block(kThreads);
#if TVK_PERSIST
  // persistent warps: every CTA slot of the device once (launch bounds: TVK_MIN_BLOCKS * 64 / kThreads CTAs per SM), never
  // more warps than there are tiles
  const uint32_t n_tiles = ((rc.width + 15u) / 16u) * ((rc.height + 7u) / 8u) * 4u;
  dim3 grid(std::max(1u, std::min((uint32_t)kSMs * (uint32_t)(TVK_MIN_BLOCKS * 64 / kThreads), (n_tiles + (kThreads / 32) - 1) / (kThreads / 32))));
  cudaMemsetAsync(rc.tile_counter, 0, sizeof(uint32_t), s);
#else
  dim3 grid((rc.width + 8 * kWX - 1) / (8 * kWX), (rc.height + 4 * kWY - 1) / (4 * kWY));
#endif
  // FAST addressing needs the +-1 gradient taps of every legal sample position inside the slot
  bool fast = !rc.nearest;
  for (int i = 0; i < 3; i++) fast = fast && rc.total[i] >= 4 && rc.ghost[i] >= 2;
  const bool b36 = rc.total[0] == 36 && rc.total[1] == 36 && rc.total[2] == 36;
  if (rc.count && rc.pipeline && MODE != 2) {
    if (fast) raycast_kernel<T, MODE, LIT, true, 0, true, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, true, true><<<grid, block, 0, s>>>(rc);
  } else if (rc.count) {
    if (fast) raycast_kernel<T, MODE, LIT, true, 0, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, true><<<grid, block, 0, s>>>(rc);
  } else if (rc.pipeline && MODE != 2) {   // depth-pipeline stage (DVR modes)
    if (fast && b36) raycast_kernel<T, MODE, LIT, true, 36, false, true><<<grid, block, 0, s>>>(rc);
    else if (fast) raycast_kernel<T, MODE, LIT, true, 0, false, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, false, true><<<grid, block, 0, s>>>(rc);
  } else if (fast && b36) raycast_kernel<T, MODE, LIT, true, 36, false><<<grid, block, 0, s>>>(rc);
  else if (fast) raycast_kernel<T, MODE, LIT, true, 0, false><<<grid, block, 0, s>>>(rc);
  else raycast_kernel<T, MODE, LIT, false, 0, false><<<grid, block, 0, s>>>(rc);
}

template <typename T>
void launch_d(const RayConsts& rc, int mode, int lighting, cudaStream_t s) {
  if (mode == TVK_RM_ISOSURFACE) launch_t<T, 2, false>(rc, s);
  else if (mode == TVK_RM_1DTRANS) { if (lighting) launch_t<T, 0, true>(rc, s); else launch_t<T, 0, false>(rc, s); }
  else { if (lighting) launch_t<T, 1, true>(rc, s); else launch_t<T, 1, false>(rc, s); }
}

}  // namespace

// rc: pool, slot_voxels, total, norm, width, height are used; the pool is addressed as a one-slot atlas per brick
void launch_fetch_probe(const RayConsts& rc_in, int dtype, bool grad, uint32_t n_slots, uint32_t steps, const float dir[3],
                        float* out, cudaStream_t s) {
  RayConsts rc = rc_in;
  for (int i = 0; i < 3; i++) rc.pool_size_f[i] = (float)rc.total[i];
  dim3 block(kThreads);
  dim3 grid((rc.width + 8 * kWX - 1) / (8 * kWX), (rc.height + 4 * kWY - 1) / (4 * kWY));
  const bool b36 = rc.total[0] == 36 && rc.total[1] == 36 && rc.total[2] == 36;
#define TVK_PROBE(T, BSV)                                                                                              \
  do {                                                                                                                 \
    if (grad) fetch_probe_kernel<T, BSV, true><<<grid, block, 0, s>>>(rc, n_slots, steps, dir[0], dir[1], dir[2], out); \
    else fetch_probe_kernel<T, BSV, false><<<grid, block, 0, s>>>(rc, n_slots, steps, dir[0], dir[1], dir[2], out);    \
  } while (0)
  switch (dtype) {
    case TVK_U8: if (b36) TVK_PROBE(uint8_t, 36); else TVK_PROBE(uint8_t, 0); break;
    case TVK_U16: if (b36) TVK_PROBE(uint16_t, 36); else TVK_PROBE(uint16_t, 0); break;
    default: if (b36) TVK_PROBE(float, 36); else TVK_PROBE(float, 0); break;
  }
#undef TVK_PROBE
}

void launch_raycast(const RayConsts& rc, int mode, int lighting, int dtype, cudaStream_t s) {
  switch (dtype) {
    case TVK_U8: launch_d<uint8_t>(rc, mode, lighting, s); break;
    case TVK_U16: launch_d<uint16_t>(rc, mode, lighting, s); break;
    default: launch_d<float>(rc, mode, lighting, s); break;
  }
}

}  // namespace tvk
