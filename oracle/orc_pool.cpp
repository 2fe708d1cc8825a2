/*
 * orc_pool.cpp -- ORACLE (test infrastructure): brick pool bookkeeping, page
 * table, min/max driven visibility and paging of the GridLeaper path.
 *
 * Follows (reference file:line):
 *   GPUMemMan::GetVolumePool                 Renderer/GPUMemMan/GPUMemMan.cpp:766-844
 *   BrickIDFlags                             Renderer/GL/GLVolumePool.cpp:25-30
 *   GetLoDSize/GetFloatBrickLayout/GetBrickLayout   GLVolumePool.cpp:82-117
 *   ctor: slot table, LoD offset table, min/max copy  GLVolumePool.cpp:119-259
 *   GetIntegerBrickID / GetVectorBrickID      GLVolumePool.cpp:288-303
 *   UploadBrick (slot bookkeeping)            GLVolumePool.cpp:673-717
 *   UploadFirstBrick / UploadBrick(BrickElemInfo)  GLVolumePool.cpp:763-778
 *   Fit1DIndexTo3DArray / CreateGLResources   GLVolumePool.cpp:816-904
 *   PrepareForPaging (std::sort by creation time)  GLVolumePool.cpp:906-910,955-959
 *   ContainsData<mode>                        GLVolumePool.cpp:962-989
 *   RecomputeVisibilityForBrickPool           GLVolumePool.cpp:991-1018
 *   RecomputeVisibilityForOctree              GLVolumePool.cpp:1020-1359
 *   UploadBricksToBrickPoolT                  GLVolumePool.cpp:1361-1393
 *   RecomputeVisibility                       GLVolumePool.cpp:1580-1718
 *   PoolSlotData                              Renderer/GL/GLVolumePool.h:26-60
 *
 * C++ only because the reference's slot replacement order is defined by
 * std::sort on equal keys; the same call is made here.
 */
#include "orc.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Slot {
  int32_t brick_id;
  uint64_t time;
  uint64_t orig_time;
  uint32_t pos[3];
  bool was_ever_used() const { return brick_id != -1; }
  bool contains_visible() const { return time > 1; }
  void flag_empty() { orig_time = time; time = 1; }
  void restore() { time = orig_time; }
};

struct MinMax { double mn, mx; };

uint32_t pow2(uint32_t e) { return 1u << e; }

void lod_size(const uint32_t v[3], uint32_t lod, uint32_t out[3]) {
  for (int i = 0; i < 3; i++) out[i] = uint32_t(std::ceil(double(v[i]) / pow2(lod)));
}

}  // namespace

struct orc_pool {
  uint32_t pool_size[3], vol[3], inner[3], total[3], capacity[3];
  uint32_t lod_count;
  std::vector<uint32_t> lod_offset;
  uint32_t total_bricks;
  uint32_t meta_dim[3];
  std::vector<uint32_t> meta;
  std::vector<Slot> slots;
  std::vector<MinMax> mm_scalar, mm_grad;
  uint64_t time_of_creation;
  size_t insert_pos;

  void brick_layout(uint32_t lod, uint32_t out[3]) const {
    uint32_t base[3];
    for (int i = 0; i < 3; i++) base[i] = uint32_t(std::ceil(double(vol[i]) / inner[i]));
    lod_size(base, lod, out);
  }
  void float_layout(uint32_t lod, float out[3]) const {
    for (int i = 0; i < 3; i++) {
      float c = float(vol[i]) / inner[i];
      c /= float(pow2(lod));
      if (float(uint32_t(c)) == c) c -= c * std::numeric_limits<float>::epsilon();
      out[i] = c;
    }
  }
  uint32_t brick_id(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const {
    uint32_t b[3];
    brick_layout(lod, b);
    return x + y * b[0] + z * b[0] * b[1] + lod_offset[lod];
  }
  void vector_id(uint32_t id, uint32_t out[4]) const {
    auto up = std::upper_bound(lod_offset.cbegin(), lod_offset.cend(), id);
    uint32_t lod = uint32_t(up - lod_offset.cbegin()) - 1;
    uint32_t b[3];
    brick_layout(lod, b);
    id -= lod_offset[lod];
    out[0] = id % b[0];
    out[1] = (id % (b[0] * b[1])) / b[0];
    out[2] = id / (b[0] * b[1]);
    out[3] = lod;
  }
  uint32_t pool_coord(const Slot& s) const {
    return s.pos[0] + s.pos[1] * capacity[0] + s.pos[2] * capacity[0] * capacity[1];
  }
  void upload_brick(uint32_t id, size_t pos, uint64_t toc) {
    Slot& s = slots[pos];
    if (s.contains_visible()) meta[s.brick_id] = ORC_BI_MISSING;
    s.brick_id = int32_t(id);
    s.time = toc;
    meta[s.brick_id] = pool_coord(s) + ORC_BI_FLAG_COUNT;
  }
  bool contains(int mode, const double v[4], uint32_t id) const {
    switch (mode) {
      case ORC_RM_1DTRANS:
        return v[1] >= mm_scalar[id].mn && v[0] <= mm_scalar[id].mx;
      case ORC_RM_2DTRANS:
        return (v[1] >= mm_scalar[id].mn && v[0] <= mm_scalar[id].mx) &&
               (v[3] >= mm_grad[id].mn && v[2] <= mm_grad[id].mx);
      default:
        return v[0] >= mm_scalar[id].mn && v[0] <= mm_scalar[id].mx;
    }
  }
};

extern "C" {

void orc_pool_size(uint64_t max_gpu_mem, uint64_t bit_width, uint64_t comp_count,
                   const uint32_t bs[3], uint64_t total_brick_count,
                   uint32_t max_dim, uint32_t out[3]) {
  const uint64_t max_voxels = max_gpu_mem / (comp_count * bit_width / 8);
  const uint64_t r3 = uint64_t(std::pow(double(max_voxels), 1.0 / 3.0));
  uint32_t g[3];
  uint64_t m = uint64_t(((float)r3 / bs[0]) + 0.5f) * bs[0];
  if (m > max_dim) m = (max_dim / bs[0]) * bs[0];
  g[0] = uint32_t(m);
  m = ((max_voxels / (uint64_t(g[0]) * g[0])) / bs[1]) * bs[1];
  if (m > max_dim) m = (max_dim / bs[1]) * bs[1];
  g[1] = uint32_t(m);
  m = ((max_voxels / (uint64_t(g[0]) * g[1])) / bs[2]) * bs[2];
  if (m > max_dim) m = (max_dim / bs[2]) * bs[2];
  g[2] = uint32_t(m);

  const uint64_t r3b = uint64_t(std::pow(double(total_brick_count), 1.0 / 3.0));
  uint64_t d[3];
  m = bs[0] * r3b;
  if (m > max_dim) m = (max_dim / bs[0]) * bs[0];
  d[0] = uint32_t(m);
  m = bs[1] * uint64_t(std::ceil(float(total_brick_count) / ((d[0] / bs[0]) * (d[0] / bs[0]))));
  if (m > max_dim) m = (max_dim / bs[1]) * bs[1];
  d[1] = uint32_t(m);
  m = bs[2] * uint64_t(std::ceil(float(total_brick_count) / ((d[0] / bs[0]) * (d[1] / bs[1]))));
  if (m > max_dim) m = (max_dim / bs[2]) * bs[2];
  d[2] = uint32_t(m);

  const bool use_d = d[0] * d[1] * d[2] < uint64_t(g[0]) * g[1] * g[2];
  for (int i = 0; i < 3; i++) out[i] = use_d ? uint32_t(d[i]) : g[i];
}

int orc_fit_1d_to_3d(uint64_t max_idx, uint32_t max_array, uint32_t out[3]) {
  const uint64_t max_elems = uint64_t(max_array) * max_array * max_array;
  if (max_idx > max_elems) return -1;
  if (max_idx < uint64_t(max_array)) {
    out[0] = uint32_t(max_idx); out[1] = 1; out[2] = 1;
  } else if (max_idx < uint64_t(max_array) * max_array) {
    out[0] = uint32_t(std::ceil(std::sqrt(double(max_idx))));
    out[1] = uint32_t(std::ceil(double(max_idx) / double(out[0])));
    out[2] = 1;
  } else {
    out[0] = uint32_t(std::ceil(std::pow(double(max_idx), 1.0 / 3.0)));
    out[1] = uint32_t(std::ceil(double(max_idx) / double(out[0] * out[0])));
    out[2] = uint32_t(std::ceil(double(max_idx) / double(out[0] * out[1])));
  }
  return 0;
}

orc_pool* orc_pool_new(const uint32_t pool_size[3], const uint32_t vol[3],
                       const uint32_t max_brick[3], uint32_t overlap,
                       uint32_t pool_lod_count, uint32_t max_3d_dim,
                       const double* minmax4) {
  orc_pool* p = new orc_pool();
  for (int i = 0; i < 3; i++) {
    p->pool_size[i] = pool_size[i];
    p->vol[i] = vol[i];
    p->total[i] = max_brick[i];
    p->inner[i] = max_brick[i] - 2 * overlap;
    p->capacity[i] = pool_size[i] / max_brick[i];
  }
  p->lod_count = pool_lod_count;
  p->time_of_creation = 2;
  p->insert_pos = 0;
  for (uint32_t z = 0; z < p->capacity[2]; z++)
    for (uint32_t y = 0; y < p->capacity[1]; y++)
      for (uint32_t x = 0; x < p->capacity[0]; x++) {
        Slot s = {-1, 0, 0, {x, y, z}};
        p->slots.push_back(s);
      }
  uint32_t off = 0;
  p->lod_offset.resize(pool_lod_count);
  for (uint32_t i = 0; i < pool_lod_count; i++) {
    p->lod_offset[i] = off;
    uint32_t b[3];
    p->brick_layout(i, b);
    off += b[0] * b[1] * b[2];
  }
  p->total_bricks = p->lod_offset.back() + 1;
  if (orc_fit_1d_to_3d(p->total_bricks, max_3d_dim, p->meta_dim) != 0) { delete p; return NULL; }
  p->meta.assign(size_t(p->meta_dim[0]) * p->meta_dim[1] * p->meta_dim[2], ORC_BI_MISSING);
  p->mm_scalar.resize(p->total_bricks);
  p->mm_grad.resize(p->total_bricks);
  /* the dataset's TOC order is LOD-major z,y,x -- identical to GetIntegerBrickID for the pool LoDs */
  for (uint32_t i = 0; i < p->total_bricks; i++) {
    p->mm_scalar[i].mn = minmax4[4 * i + 0];
    p->mm_scalar[i].mx = minmax4[4 * i + 1];
    p->mm_grad[i].mn = minmax4[4 * i + 2];
    p->mm_grad[i].mx = minmax4[4 * i + 3];
  }
  return p;
}

void orc_pool_free(orc_pool* p) { delete p; }
uint32_t orc_pool_total_bricks(const orc_pool* p) { return p->total_bricks; }
uint32_t orc_pool_meta_count(const orc_pool* p) { return uint32_t(p->meta.size()); }
const uint32_t* orc_pool_meta(const orc_pool* p) { return p->meta.data(); }
void orc_pool_meta_dim(const orc_pool* p, uint32_t out[3]) { memcpy(out, p->meta_dim, 12); }
void orc_pool_capacity(const orc_pool* p, uint32_t out[3]) { memcpy(out, p->capacity, 12); }
void orc_pool_lod_offsets(const orc_pool* p, uint32_t* out) {
  memcpy(out, p->lod_offset.data(), 4 * p->lod_count);
}
void orc_pool_brick_layout(const orc_pool* p, uint32_t lod, uint32_t out[3]) { p->brick_layout(lod, out); }
void orc_pool_float_layout(const orc_pool* p, uint32_t lod, float out[3]) { p->float_layout(lod, out); }
uint32_t orc_pool_brick_id(const orc_pool* p, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  return p->brick_id(x, y, z, lod);
}
void orc_pool_vector_id(const orc_pool* p, uint32_t id, uint32_t out[4]) { p->vector_id(id, out); }

uint32_t orc_pool_upload_first(orc_pool* p) {
  const uint32_t last = p->lod_offset.back();
  p->upload_brick(last, p->slots.size() - 1, std::numeric_limits<uint64_t>::max());
  return p->pool_coord(p->slots.back());
}

void orc_pool_recompute_visibility(orc_pool* p, int mode, double a, double b,
                                   double c, double d, uint32_t counts[4]) {
  const double v[4] = {a, b, c, d};
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  std::fill(p->meta.begin(), p->meta.end(), uint32_t(ORC_BI_MISSING));

  /* RecomputeVisibilityForBrickPool */
  for (auto s = p->slots.begin(); s < p->slots.end(); s++) {
    if (!s->was_ever_used()) continue;
    const bool has = p->contains(mode, v, s->brick_id);
    const bool had = s->contains_visible();
    if (has) {
      if (!had) s->restore();
      p->meta[s->brick_id] = p->pool_coord(*s) + ORC_BI_FLAG_COUNT;
    } else {
      if (had) s->flag_empty();
      p->meta[s->brick_id] = ORC_BI_EMPTY;
    }
  }

  /* RecomputeVisibilityForOctree<false, mode> */
  uint32_t child[3];
  p->brick_layout(0, child);
  for (uint32_t z = 0; z < child[2]; z++)
    for (uint32_t y = 0; y < child[1]; y++)
      for (uint32_t x = 0; x < child[0]; x++) {
        counts[0]++;
        const uint32_t id = p->brick_id(x, y, z, 0);
        if (p->meta[id] < ORC_BI_FLAG_COUNT && !p->contains(mode, v, id)) {
          p->meta[id] = ORC_BI_CHILD_EMPTY;
          counts[3]++;
        }
      }
  for (uint32_t lod = 1; lod < p->lod_count; lod++) {
    uint32_t lay[3];
    p->brick_layout(lod, lay);
    /* every parent (x,y,z) is visited exactly once by the reference's even region +
       odd planes / lines / corner; its existing children are (2x+dx,2y+dy,2z+dz) < child layout */
    for (uint32_t z = 0; z < lay[2]; z++)
      for (uint32_t y = 0; y < lay[1]; y++)
        for (uint32_t x = 0; x < lay[0]; x++) {
          counts[0]++;
          const uint32_t id = p->brick_id(x, y, z, lod);
          if (p->meta[id] >= ORC_BI_FLAG_COUNT) continue;
          if (p->contains(mode, v, id)) continue;
          p->meta[id] = ORC_BI_CHILD_EMPTY;
          bool all_child_empty = true;
          for (uint32_t dz = 0; dz < 2; dz++)
            for (uint32_t dy = 0; dy < 2; dy++)
              for (uint32_t dx = 0; dx < 2; dx++) {
                const uint32_t cx = 2 * x + dx, cy = 2 * y + dy, cz = 2 * z + dz;
                if (cx >= child[0] || cy >= child[1] || cz >= child[2]) continue;
                if (p->meta[p->brick_id(cx, cy, cz, lod - 1)] != ORC_BI_CHILD_EMPTY)
                  all_child_empty = false;
              }
          if (!all_child_empty) { p->meta[id] = ORC_BI_EMPTY; counts[1]++; }
          else counts[2]++;
        }
    memcpy(child, lay, sizeof(child));
  }
}

uint32_t orc_pool_upload_bricks(orc_pool* p, const uint32_t* ids, uint32_t n, uint32_t* out_slots) {
  uint32_t paged = 0;
  for (uint32_t i = 0; i < n; i++) out_slots[i] = 0xFFFFFFFFu;
  if (n == 0) return 0;
  /* PrepareForPaging */
  std::sort(p->slots.begin(), p->slots.end(),
            [](const Slot& i, const Slot& j) { return i.time < j.time; });
  p->insert_pos = 0;
  for (uint32_t i = 0; i < n; i++) {
    /* UploadBrick(BrickElemInfo): stop once all slots but the last were replaced in this frame */
    if (p->insert_pos >= p->slots.size() - 1) break;
    const uint32_t id = p->brick_id(ids[4 * i], ids[4 * i + 1], ids[4 * i + 2], ids[4 * i + 3]);
    p->upload_brick(id, p->insert_pos, p->time_of_creation++);
    out_slots[i] = p->pool_coord(p->slots[p->insert_pos]);
    p->insert_pos++;
    paged++;
  }
  return paged;
}

uint32_t orc_pool_slot_count(const orc_pool* p) { return uint32_t(p->slots.size()); }
void orc_pool_slots(const orc_pool* p, int32_t* brick_ids, uint64_t* times, uint32_t* pos3) {
  for (size_t i = 0; i < p->slots.size(); i++) {
    brick_ids[i] = p->slots[i].brick_id;
    times[i] = p->slots[i].time;
    memcpy(pos3 + 3 * i, p->slots[i].pos, 12);
  }
}

}  // extern "C"
