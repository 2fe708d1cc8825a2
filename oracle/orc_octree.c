/*
 * orc_octree.c -- ORACLE (test infrastructure): restatement of the reference's
 * bricked LOD hierarchy builder.
 *
 * Follows (reference file:line):
 *   ExtendedOctree::ComputeMetadata        IO/UVF/ExtendedOctree/ExtendedOctree.cpp:188-243
 *   ExtendedOctree::ComputeBrickSize       ExtendedOctree.cpp:276-285
 *   ExtendedOctree::BrickCoordsToIndex     ExtendedOctree.cpp:394-401
 *   ExtendedOctreeConverter::GetInputBrick ExtendedOctreeConverter.cpp:288-373
 *   ExtendedOctreeConverter::ClampToEdge   ExtendedOctreeConverter.cpp:376-462
 *   DownsampleBricktoBrick / DownsampleBrick ExtendedOctreeConverter.inc:1-356
 *   FillOverlap                            ExtendedOctreeConverter.cpp:1203-1380
 *   ComputeBrickStats                      ExtendedOctreeConverter.inc:408-444
 *   VolumeTools::Filter                    VolumeTools.h:168-262
 *   MaxMinDataBlock::SetDataFromFlatVector IO/UVF/MaxMinDataBlock.cpp:175-195
 *
 * Inner voxels: the reference downsamples brick-by-brick; this restatement
 * downsamples whole LOD volumes: LOD l+1 voxel = Filter over the existing
 * voxels of the 2x2x2 block of LOD l (argument order dx-major, dz-minor =
 * the p0..p7 order of the reference).  Identical when the inner brick size is
 * even (asserted).
 *
 * Ghost voxels: LOD 0 bricks are cut from the flat input with ghost voxels
 * zero (or clamp-to-edge) outside the volume (GetInputBrick).  LOD >= 1 bricks
 * get their ghost voxels from FillOverlap, which is emulated copy-by-copy in
 * the reference's order, because that order leaves three ghost corner regions
 * of a brick (left&bottom&back, right&top&back, right&bottom&front) holding
 * the not-yet-filled ghost of a later brick, i.e. ZERO in the default
 * zero-border mode ("Q1", found by diffing against the compiled reference).
 * This is deterministic reference behaviour and feeds the per-brick min/max,
 * so it is reproduced.  orc_octree_q1_closed_form() states the same thing as
 * a per-voxel rule (what the GPU bricker implements); tests check both agree.
 *
 * Outside the contract (the reference reads stale memory there, nothing to
 * match): a last-brick remainder smaller than `overlap` ("Q2"), and the
 * not-yet-filled ghost corners in clamp-to-edge mode (emulated as zero).
 * tests/test_octree_ref.py diffs everything else bit-exactly against the
 * compiled reference converter (oracle/_ref/ref_octree).
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

struct orc_octree {
  uint32_t vol[3], brick[3], overlap;
  int dtype;
  uint32_t lod_count;
  uint32_t lod_size[ORC_MAX_LOD][3];
  uint32_t lod_bricks[ORC_MAX_LOD][3];
  uint64_t lod_offset[ORC_MAX_LOD];
  uint64_t n_bricks;
  void* lod_vol[ORC_MAX_LOD];
  uint8_t** bricks[ORC_MAX_LOD];   /* per LOD >= 1: FillOverlap-emulated bricks */
  double* minmax;
  int clamp;
};

static size_t esize(int dtype) { return dtype == ORC_U8 ? 1 : dtype == ORC_U16 ? 2 : 4; }

static uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static size_t vol3(const uint32_t s[3]) { return (size_t)s[0] * s[1] * s[2]; }

orc_octree* orc_octree_new(const uint32_t vol[3], const uint32_t max_brick[3],
                           uint32_t overlap, int dtype) {
  for (int i = 0; i < 3; i++)
    if (max_brick[i] <= 2 * overlap || vol[i] == 0) return NULL;
  orc_octree* t = (orc_octree*)calloc(1, sizeof(*t));
  memcpy(t->vol, vol, sizeof(t->vol));
  memcpy(t->brick, max_brick, sizeof(t->brick));
  t->overlap = overlap;
  t->dtype = dtype;
  /* ComputeMetadata: LOD sizes = ceil(prev/2) per axis (axes of size 1 stay 1) until 1^3 */
  uint32_t s[3] = {vol[0], vol[1], vol[2]};
  uint32_t l = 0;
  for (;;) {
    if (l > 0)
      for (int i = 0; i < 3; i++)
        if (s[i] > 1) s[i] = (s[i] + 1) / 2;
    for (int i = 0; i < 3; i++) {
      t->lod_size[l][i] = s[i];
      t->lod_bricks[l][i] = cdiv(s[i], max_brick[i] - 2 * overlap);
    }
    l++;
    if (!(s[0] > 1 || s[1] > 1 || s[2] > 1) || l >= ORC_MAX_LOD) break;
  }
  t->lod_count = l;
  t->lod_offset[0] = 0;
  for (uint32_t i = 1; i < l; i++)
    t->lod_offset[i] = t->lod_offset[i - 1] +
      (uint64_t)t->lod_bricks[i - 1][0] * t->lod_bricks[i - 1][1] * t->lod_bricks[i - 1][2];
  t->n_bricks = t->lod_offset[l - 1] +
      (uint64_t)t->lod_bricks[l - 1][0] * t->lod_bricks[l - 1][1] * t->lod_bricks[l - 1][2];
  return t;
}

static void free_bricks(orc_octree* t) {
  for (uint32_t l = 0; l < t->lod_count; l++) {
    if (!t->bricks[l]) continue;
    size_t n = (size_t)t->lod_bricks[l][0] * t->lod_bricks[l][1] * t->lod_bricks[l][2];
    for (size_t i = 0; i < n; i++) free(t->bricks[l][i]);
    free(t->bricks[l]);
    t->bricks[l] = NULL;
  }
}

void orc_octree_free(orc_octree* t) {
  if (!t) return;
  free_bricks(t);
  for (uint32_t i = 0; i < t->lod_count; i++) free(t->lod_vol[i]);
  free(t->minmax);
  free(t);
}

uint32_t orc_octree_lod_count(const orc_octree* t) { return t->lod_count; }

uint32_t orc_octree_largest_single_brick_lod(const orc_octree* t) {
  /* first (finest) LOD that consists of exactly one brick; cf. UVFDataset::GetLargestSingleBrickLOD */
  for (uint32_t l = 0; l < t->lod_count; l++)
    if (t->lod_bricks[l][0] * t->lod_bricks[l][1] * t->lod_bricks[l][2] == 1) return l;
  return t->lod_count - 1;
}

void orc_octree_lod_size(const orc_octree* t, uint32_t lod, uint32_t out[3]) {
  memcpy(out, t->lod_size[lod], 3 * sizeof(uint32_t));
}
void orc_octree_brick_count(const orc_octree* t, uint32_t lod, uint32_t out[3]) {
  memcpy(out, t->lod_bricks[lod], 3 * sizeof(uint32_t));
}
uint64_t orc_octree_total_bricks(const orc_octree* t) { return t->n_bricks; }

uint64_t orc_octree_brick_index(const orc_octree* t, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  const uint32_t* c = t->lod_bricks[lod];
  return t->lod_offset[lod] + x + (uint64_t)y * c[0] + (uint64_t)z * c[0] * c[1];
}

void orc_octree_brick_size(const orc_octree* t, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]) {
  const uint32_t co[3] = {x, y, z};
  for (int i = 0; i < 3; i++) {
    uint32_t core = t->brick[i] - 2 * t->overlap;
    int last = co[i] == t->lod_bricks[lod][i] - 1;
    uint32_t rem = t->lod_size[lod][i] % core;
    out[i] = (last && rem) ? 2 * t->overlap + rem : t->brick[i];
  }
}

/* ---- typed helpers ------------------------------------------------- */
#define DEF_FILTER(T, NAME)                                                          \
  static void order_##NAME(T* a, T* b) { if (*a > *b) { T x = *a; *a = *b; *b = x; } } \
  static void insq_##NAME(T* a, T* b, T* c, T* d, T* p) {                            \
    if (*p > *c) { order_##NAME(d, p); }                                             \
    else if (*p < *b) { *d = *c; *c = *b; *b = *p; order_##NAME(a, b); }             \
    else { *d = *c; *c = *p; }                                                       \
  }                                                                                  \
  static T filter_##NAME(const T* v, int n, int median) {                            \
    if (n == 1) return v[0];                                                         \
    if (median) {                                                                    \
      if (n == 2) return v[0];                                                       \
      if (n == 4) { T a = v[0], b = v[1], c = v[2];                                  \
        order_##NAME(&a, &b); order_##NAME(&b, &c); return a > b ? a : b; }          \
      T a = v[0], b = v[1], c = v[2], d = v[3], e = v[4], f = v[5], g = v[6];        \
      order_##NAME(&a, &b); order_##NAME(&c, &d); order_##NAME(&a, &c);              \
      order_##NAME(&b, &d); order_##NAME(&b, &c);                                    \
      insq_##NAME(&a, &b, &c, &d, &e); insq_##NAME(&a, &b, &c, &d, &f);              \
      T m = d < g ? d : g; return m > c ? m : c;                                     \
    }                                                                                \
    double s = (double)v[0];                                                         \
    for (int i = 1; i < n; i++) s = s + (double)v[i];                                \
    return (T)(s / (double)n);                                                       \
  }                                                                                  \
  static void downsample_##NAME(const T* src, const uint32_t ss[3], T* dst,          \
                                const uint32_t ds[3], int median) {                  \
    for (uint32_t z = 0; z < ds[2]; z++)                                             \
      for (uint32_t y = 0; y < ds[1]; y++)                                           \
        for (uint32_t x = 0; x < ds[0]; x++) {                                       \
          T v[8]; int n = 0;                                                         \
          /* axes that were not halved (size 1 stays 1) map 1:1 */                   \
          uint32_t bx = ss[0] > 1 ? 2 * x : x, by = ss[1] > 1 ? 2 * y : y,           \
                   bz = ss[2] > 1 ? 2 * z : z;                                       \
          uint32_t nx = (ss[0] > 1 && bx + 1 < ss[0]) ? 2 : 1;                       \
          uint32_t ny = (ss[1] > 1 && by + 1 < ss[1]) ? 2 : 1;                       \
          uint32_t nz = (ss[2] > 1 && bz + 1 < ss[2]) ? 2 : 1;                       \
          for (uint32_t dx = 0; dx < nx; dx++)                                       \
            for (uint32_t dy = 0; dy < ny; dy++)                                     \
              for (uint32_t dz = 0; dz < nz; dz++)                                   \
                v[n++] = src[(size_t)(bx + dx) + (size_t)ss[0] * ((by + dy) +        \
                             (size_t)ss[1] * (bz + dz))];                            \
          dst[(size_t)x + (size_t)ds[0] * (y + (size_t)ds[1] * z)] =                 \
              filter_##NAME(v, n, median);                                           \
        }                                                                            \
  }

DEF_FILTER(uint8_t, u8)
DEF_FILTER(uint16_t, u16)
DEF_FILTER(float, f32)

/* cut a brick out of the LOD volume; ghost voxels outside the volume are zero or clamped.
   inner_only: leave everything but the inner region zero (state after DownsampleBrick). */
static void cut_brick(const orc_octree* t, uint32_t bx, uint32_t by, uint32_t bz, uint32_t lod,
                      void* dst, int inner_only);

int orc_octree_get_brick(const orc_octree* t, uint32_t bx, uint32_t by, uint32_t bz, uint32_t lod, void* dst) {
  if (!t->lod_vol[lod]) return -1;
  if (lod > 0 && t->bricks[lod]) {
    uint32_t bs[3];
    orc_octree_brick_size(t, bx, by, bz, lod, bs);
    const uint32_t* c = t->lod_bricks[lod];
    memcpy(dst, t->bricks[lod][bx + (size_t)c[0] * (by + (size_t)c[1] * bz)], vol3(bs) * esize(t->dtype));
    return 0;
  }
  cut_brick(t, bx, by, bz, lod, dst, 0);
  return 0;
}

static void cut_brick(const orc_octree* t, uint32_t bx, uint32_t by, uint32_t bz, uint32_t lod,
                      void* dst, int inner_only) {
  uint32_t bs[3];
  orc_octree_brick_size(t, bx, by, bz, lod, bs);
  const uint32_t* ls = t->lod_size[lod];
  const size_t es = esize(t->dtype);
  const int64_t ov = t->overlap;
  const int64_t ox = (int64_t)bx * (t->brick[0] - 2 * ov) - ov;
  const int64_t oy = (int64_t)by * (t->brick[1] - 2 * ov) - ov;
  const int64_t oz = (int64_t)bz * (t->brick[2] - 2 * ov) - ov;
  const uint8_t* src = (const uint8_t*)t->lod_vol[lod];
  uint8_t* d = (uint8_t*)dst;
  for (uint32_t z = 0; z < bs[2]; z++)
    for (uint32_t y = 0; y < bs[1]; y++)
      for (uint32_t x = 0; x < bs[0]; x++) {
        int64_t gx = ox + x, gy = oy + y, gz = oz + z;
        int inside = gx >= 0 && gy >= 0 && gz >= 0 && gx < ls[0] && gy < ls[1] && gz < ls[2];
        uint8_t* o = d + es * ((size_t)x + (size_t)bs[0] * (y + (size_t)bs[1] * z));
        if (inner_only) {
          int in = x >= ov && y >= ov && z >= ov && x < bs[0] - ov && y < bs[1] - ov && z < bs[2] - ov;
          if (!in) { memset(o, 0, es); continue; }
        }
        if (!inside && !t->clamp) { memset(o, 0, es); continue; }
        if (!inside) {
          gx = gx < 0 ? 0 : gx >= ls[0] ? ls[0] - 1 : gx;
          gy = gy < 0 ? 0 : gy >= ls[1] ? ls[1] - 1 : gy;
          gz = gz < 0 ? 0 : gz >= ls[2] ? ls[2] - 1 : gz;
        }
        memcpy(o, src + es * ((size_t)gx + (size_t)ls[0] * (gy + (size_t)ls[1] * gz)), es);
      }
}

/* CopyBrickToBrick, ExtendedOctreeConverter.cpp:1153-1181 */
static void copy_region(const uint8_t* src, const uint32_t ss[3], uint8_t* dst, const uint32_t ds[3],
                        uint32_t sx, uint32_t sy, uint32_t sz, uint32_t dx, uint32_t dy, uint32_t dz,
                        uint32_t rx, uint32_t ry, uint32_t rz, size_t es) {
  for (uint32_t z = 0; z < rz; z++)
    for (uint32_t y = 0; y < ry; y++)
      memcpy(dst + es * (dx + (size_t)(dy + y) * ds[0] + (size_t)(dz + z) * ds[0] * ds[1]),
             src + es * (sx + (size_t)(sy + y) * ss[0] + (size_t)(sz + z) * ss[0] * ss[1]), es * rx);
}

/* ClampToEdge, ExtendedOctreeConverter.cpp:376-462 (applied after the copies, :1352-1362) */
static void clamp_edges(uint8_t* d, const uint32_t bs[3], uint32_t ov, size_t es,
                        int xs, int ys, int zs, int xe, int ye, int ze) {
#define AT(x, y, z) (d + es * ((size_t)(x) + (size_t)bs[0] * ((y) + (size_t)bs[1] * (z))))
  if (xs) for (uint32_t z = 0; z < bs[2]; z++) for (uint32_t y = 0; y < bs[1]; y++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(o, y, z), AT(ov, y, z), es);
  if (xe) for (uint32_t z = 0; z < bs[2]; z++) for (uint32_t y = 0; y < bs[1]; y++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(bs[0] - 1 - o, y, z), AT(bs[0] - 1 - ov, y, z), es);
  if (ys) for (uint32_t z = 0; z < bs[2]; z++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(0, o, z), AT(0, ov, z), es * bs[0]);
  if (ye) for (uint32_t z = 0; z < bs[2]; z++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(0, bs[1] - 1 - o, z), AT(0, bs[1] - 1 - ov, z), es * bs[0]);
  if (zs) for (uint32_t y = 0; y < bs[1]; y++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(0, y, o), AT(0, y, ov), es * bs[0]);
  if (ze) for (uint32_t y = 0; y < bs[1]; y++)
    for (uint32_t o = 0; o < ov; o++) memcpy(AT(0, y, bs[2] - 1 - o), AT(0, y, bs[2] - 1 - ov), es * bs[0]);
#undef AT
}

/* FillOverlap, ExtendedOctreeConverter.cpp:1203-1380, copy by copy in the reference's order */
static void fill_overlap(orc_octree* t, uint32_t lod) {
  const uint32_t* c = t->lod_bricks[lod];
  const size_t es = esize(t->dtype);
  const uint32_t ov = t->overlap;
  size_t n = (size_t)c[0] * c[1] * c[2];
  uint8_t** B = (uint8_t**)calloc(n, sizeof(uint8_t*));
  t->bricks[lod] = B;
#define BR(x, y, z) B[(x) + (size_t)c[0] * ((y) + (size_t)c[1] * (z))]
  for (uint32_t z = 0; z < c[2]; z++) for (uint32_t y = 0; y < c[1]; y++) for (uint32_t x = 0; x < c[0]; x++) {
    uint32_t bs[3];
    orc_octree_brick_size(t, x, y, z, lod, bs);
    BR(x, y, z) = (uint8_t*)malloc(vol3(bs) * es);
    cut_brick(t, x, y, z, lod, BR(x, y, z), 1);   /* state after DownsampleBrick: inner only */
  }
  for (uint32_t z = 0; z < c[2]; z++) for (uint32_t y = 0; y < c[1]; y++) for (uint32_t x = 0; x < c[0]; x++) {
    int L = x > 0, R = x < c[0] - 1, T = y > 0, Bo = y < c[1] - 1, F = z > 0, K = z < c[2] - 1;
    uint32_t ts[3], ss[3];
    orc_octree_brick_size(t, x, y, z, lod, ts);
    uint8_t* tg = BR(x, y, z);
#define SRC(X, Y, Z) const uint8_t* sr = BR(X, Y, Z); orc_octree_brick_size(t, X, Y, Z, lod, ss)
    if (R)  { SRC(x + 1, y, z); copy_region(sr, ss, tg, ts, ov, 0, 0, ts[0] - ov, 0, 0, ov, ss[1], ss[2], es); }
    if (Bo) { SRC(x, y + 1, z); copy_region(sr, ss, tg, ts, 0, ov, 0, 0, ts[1] - ov, 0, ss[0], ov, ss[2], es); }
    if (K)  { SRC(x, y, z + 1); copy_region(sr, ss, tg, ts, 0, 0, ov, 0, 0, ts[2] - ov, ss[0], ss[1], ov, es); }
    if (L)  { SRC(x - 1, y, z); copy_region(sr, ss, tg, ts, ss[0] - 2 * ov, 0, 0, 0, 0, 0, ov, ss[1], ss[2], es); }
    if (T)  { SRC(x, y - 1, z); copy_region(sr, ss, tg, ts, 0, ss[1] - 2 * ov, 0, 0, 0, 0, ss[0], ov, ss[2], es); }
    if (F)  { SRC(x, y, z - 1); copy_region(sr, ss, tg, ts, 0, 0, ss[2] - 2 * ov, 0, 0, 0, ss[0], ss[1], ov, es); }
    if (Bo && R) { SRC(x + 1, y + 1, z); copy_region(sr, ss, tg, ts, ov, ov, 0, ts[0] - ov, ts[1] - ov, 0, ov, ov, ss[2], es); }
    if (R && K)  { SRC(x + 1, y, z + 1); copy_region(sr, ss, tg, ts, ov, 0, ov, ts[0] - ov, 0, ts[2] - ov, ov, ss[1], ov, es); }
    if (Bo && K) { SRC(x, y + 1, z + 1); copy_region(sr, ss, tg, ts, 0, ov, ov, 0, ts[1] - ov, ts[2] - ov, ss[0], ov, ov, es); }
    if (R && Bo && K) { SRC(x + 1, y + 1, z + 1); copy_region(sr, ss, tg, ts, ov, ov, ov, ts[0] - ov, ts[1] - ov, ts[2] - ov, ov, ov, ov, es); }
#undef SRC
    if (t->clamp) clamp_edges(tg, ts, ov, es, !L, !T, !F, !R, !Bo, !K);
  }
#undef BR
}

/* chan0 != NULL: `t` holds component c > 0 of a multi-component volume whose component 0 is `chan0` (already built).  The
 * converter filters every component alike -- except for the single voxel at the x/y/z corner of a source brick whose three
 * inner sizes are odd, where DownsampleBricktoBrick writes COMPONENT 0 of the source voxel into every component
 * (`*(pTargetData+c) = *p0;`, ExtendedOctreeConverter.inc:232-246).  Inner brick sizes are even here, so only the last brick
 * of a level whose three sizes are odd has that corner: the last voxel of the next level. */
static int build_impl(orc_octree* t, const void* flat, int clamp_to_edge, int median, const orc_octree* chan0) {
  const size_t es = esize(t->dtype);
  t->clamp = clamp_to_edge;
  for (int i = 0; i < 3; i++)
    if ((t->brick[i] - 2 * t->overlap) % 2) return -2;   /* odd inner sizes: not restated */
  free_bricks(t);
  for (uint32_t l = 0; l < t->lod_count; l++) { free(t->lod_vol[l]); t->lod_vol[l] = NULL; }
  t->lod_vol[0] = malloc(vol3(t->lod_size[0]) * es);
  memcpy(t->lod_vol[0], flat, vol3(t->lod_size[0]) * es);
  for (uint32_t l = 1; l < t->lod_count; l++) {
    t->lod_vol[l] = malloc(vol3(t->lod_size[l]) * es);
    switch (t->dtype) {
      case ORC_U8:  downsample_u8((const uint8_t*)t->lod_vol[l - 1], t->lod_size[l - 1], (uint8_t*)t->lod_vol[l], t->lod_size[l], median); break;
      case ORC_U16: downsample_u16((const uint16_t*)t->lod_vol[l - 1], t->lod_size[l - 1], (uint16_t*)t->lod_vol[l], t->lod_size[l], median); break;
      default:      downsample_f32((const float*)t->lod_vol[l - 1], t->lod_size[l - 1], (float*)t->lod_vol[l], t->lod_size[l], median); break;
    }
    if (chan0 && (t->lod_size[l - 1][0] & 1u) && (t->lod_size[l - 1][1] & 1u) && (t->lod_size[l - 1][2] & 1u))
      memcpy((uint8_t*)t->lod_vol[l] + (vol3(t->lod_size[l]) - 1) * es,
             (const uint8_t*)chan0->lod_vol[l - 1] + (vol3(t->lod_size[l - 1]) - 1) * es, es);
  }
  for (uint32_t l = 1; l < t->lod_count; l++) fill_overlap(t, l);
  /* ComputeBrickStats: min/max over EVERY stored voxel of the brick incl. ghost, as double;
     gradient range (-DBL_MAX, DBL_MAX) (MaxMinDataBlock.cpp:175-195) */
  free(t->minmax);
  t->minmax = (double*)malloc(sizeof(double) * 4 * t->n_bricks);
  size_t maxvox = (size_t)t->brick[0] * t->brick[1] * t->brick[2];
  void* tmp = malloc(maxvox * es);
  for (uint32_t l = 0; l < t->lod_count; l++)
    for (uint32_t z = 0; z < t->lod_bricks[l][2]; z++)
      for (uint32_t y = 0; y < t->lod_bricks[l][1]; y++)
        for (uint32_t x = 0; x < t->lod_bricks[l][0]; x++) {
          uint32_t bs[3];
          orc_octree_brick_size(t, x, y, z, l, bs);
          orc_octree_get_brick(t, x, y, z, l, tmp);
          size_t n = vol3(bs);
          double mn = DBL_MAX, mx = -DBL_MAX;
          for (size_t i = 0; i < n; i++) {
            double c = t->dtype == ORC_U8 ? (double)((uint8_t*)tmp)[i]
                     : t->dtype == ORC_U16 ? (double)((uint16_t*)tmp)[i]
                     : (double)((float*)tmp)[i];
            mn = c < mn ? c : mn;
            mx = c > mx ? c : mx;
          }
          double* o = t->minmax + 4 * orc_octree_brick_index(t, x, y, z, l);
          o[0] = mn; o[1] = mx; o[2] = -DBL_MAX; o[3] = DBL_MAX;
        }
  free(tmp);
  return 0;
}

int orc_octree_build(orc_octree* t, const void* flat, int clamp_to_edge, int median) {
  return build_impl(t, flat, clamp_to_edge, median, NULL);
}
/* component c > 0 of a multi-component volume (see build_impl); same geometry and type as chan0 */
int orc_octree_build_component(orc_octree* t, const void* flat, int clamp_to_edge, int median, const orc_octree* chan0) {
  if (!chan0 || chan0->dtype != t->dtype || chan0->lod_count != t->lod_count) return -3;
  return build_impl(t, flat, clamp_to_edge, median, chan0);
}

const double* orc_octree_minmax(const orc_octree* t) { return t->minmax; }
const void* orc_octree_lod_volume(const orc_octree* t, uint32_t lod) { return t->lod_vol[lod]; }
