// tvk_host.h -- host-side state of one renderer (tvk_ctx) and small math helpers.
#ifndef TVK_HOST_H
#define TVK_HOST_H
#include <cmath>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "tvk_dev.h"
#include "octree_file.h"

namespace tvk {

// ---- double precision 4x4, Tuvok storage (row vectors) --------------------------------------
inline bool inv4(const double* a, double* out) {   // Gauss-Jordan, partial pivoting
  double m[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) { m[r][c] = a[r * 4 + c]; m[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++) if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
    if (m[p][c] == 0.0) return false;
    if (p != c) for (int k = 0; k < 8; k++) { double t = m[c][k]; m[c][k] = m[p][k]; m[p][k] = t; }
    const double d = m[c][c];
    for (int k = 0; k < 8; k++) m[c][k] = m[c][k] / d;
    for (int r = 0; r < 4; r++) if (r != c) {
      const double f = m[r][c];
      for (int k = 0; k < 8; k++) m[r][k] = m[r][k] - f * m[c][k];
    }
  }
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = m[r][4 + c];
  return true;
}
inline void mul4(const double* a, const double* b, double* out) {
  double t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s = s + a[r * 4 + k] * b[k * 4 + c];
      t[r * 4 + c] = s;
    }
  std::memcpy(out, t, sizeof(t));
}

// PoolSlotData (Renderer/GL/GLVolumePool.h:26-60)
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

// VisibilityState (Renderer/VisibilityState.h:15-50)
struct VisState {
  int mode = -1;
  double v[4] = {0, 0, 0, 0};
  bool needs_update(int m, double a, double b, double c, double d) {
    const bool changed = m != mode || a != v[0] || b != v[1] || c != v[2] || d != v[3];
    mode = m; v[0] = a; v[1] = b; v[2] = c; v[3] = d;
    return changed;
  }
};

// procedural brick source (tvk_procedural.inc): generator parameters, the host-side brick cache in front of it, counters
struct ProcState {
  bool on = false;
  int kind = 0;
  uint32_t seed = 0, threads = 1;
  unsigned char* arena = nullptr;          // cache_slots bricks of slot_bytes each
  bool pinned = false;                     // ... page-locked: brick copies DMA straight out of the cache
  uint32_t cache_slots = 0, used = 0;
  std::mutex mu;
  std::unordered_map<uint32_t, uint32_t> map;   // page-table index -> cache slot
  std::vector<uint32_t> owner;
  std::vector<uint64_t> stamp;
  std::vector<uint8_t> ready;
  std::vector<uint32_t> pins;
  uint64_t clock = 0, evicted = 0;
  std::atomic<uint64_t> generated{0}, hits{0}, bytes{0}, source_ns{0};
};

}  // namespace tvk

struct tvk_sortlast;

struct tvk_ctx {
  tvk_device_cfg cfg{};
  std::string err;
  tvk_log_cb log_cb = nullptr;
  void* log_user = nullptr;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool counters_on = false;
  // LPT tile schedule of the traversal kernel (TVK_TILE_LPT / tvk_set_tile_schedule): tiles sorted by last frame's cost
  int tile_lpt = 0;
  uint32_t tile_split_cost = 0;     // persistent build: tiles that cost at least this many turns are split (0 = never)
  uint32_t* tile_cost_d = nullptr;
  uint32_t* tile_order_d[2] = {nullptr, nullptr};
  uint32_t tile_n = 0; int tile_cur = 0; bool tile_valid = false;
  uint32_t launch_seq = 0;         // traversal launches so far (selects the tile counter of a launch)

  // ---- dataset ----
  bool have_volume = false;
  uint32_t vol[3] = {0, 0, 0}, brick[3] = {0, 0, 0}, inner[3] = {0, 0, 0}, overlap = 0;
  float scale[3] = {1, 1, 1};
  int dtype = 0;
  uint32_t esize = 1;
  double range_max = 0;
  float max_grad = 0;
  uint32_t lod_count = 0, pool_lod_count = 0;
  uint32_t lod_size[TVK_MAX_LOD][3]{};
  uint32_t layout[TVK_MAX_LOD][3]{};      // octree brick counts per LOD
  uint64_t toc_offset[TVK_MAX_LOD + 1]{}; // first TOC index of each LOD
  uint64_t n_bricks_all = 0;              // all LODs
  std::vector<double> minmax_h;           // 4 per brick, TOC order
  double* minmax_d = nullptr;
  tvk_brick_cb cb = nullptr;
  void* cb_user = nullptr;
  tvk::OctreeFile* file = nullptr;        // ExtendedOctree file source (tvk_open_octree_file); cb then points at it
  uint32_t io_threads = 8;                // parallel pread/decode workers of the file source
  bool pyramid_median = false;            // tvk_set_pyramid_filter: median instead of mean when tvk_build_volume halves a level
  tvk::ProcState proc;                    // procedural source (tvk_set_procedural_volume); cb then points at it
  uint64_t up_bricks = 0, up_bytes = 0; double up_ms = 0, up_h2d_ms = 0;   // streaming totals (tvk_get_stream_stats)
  void* store_d = nullptr;                // device brick store (slot layout, TOC order) or null
  uint64_t slot_voxels = 0, slot_bytes = 0;   // per slot: voxels, bytes of the plain brick (store / staging)
  uint64_t pool_slot_bytes = 0;               // bytes of one POOL slot (x-pair layout: 2 x slot_bytes for 8 / 16-bit data)

  // ---- transfer functions ----
  uint32_t* tf1d_d = nullptr; uint32_t tf1d_n = 0; uint64_t tf1d_nz[2] = {0, 0};
  uint32_t* tf2d_d = nullptr; uint32_t tf2d_w = 0, tf2d_h = 0; uint64_t tf2d_nz[4] = {0, 0, 0, 0};

  // ---- pool ----
  bool have_pool = false;
  uint32_t pool_size[3] = {0, 0, 0}, capacity[3] = {0, 0, 0}, n_slots = 0;
  void* pool_d = nullptr;
  uint32_t lod_offset[TVK_MAX_LOD]{};     // pool brick-id offsets
  uint32_t pool_layout[TVK_MAX_LOD][3]{};
  uint32_t total_bricks = 0;
  uint32_t meta_dim[3] = {0, 0, 0};
  uint64_t meta_count = 0;
  uint32_t* meta_d = nullptr;
  std::vector<uint32_t> meta_h;
  std::vector<tvk::Slot> slots;
  int32_t* slot_brick_d = nullptr;
  uint64_t time_of_creation = 2;
  size_t insert_pos = 0;
  tvk::VisState vis;
  uint32_t* counts_d = nullptr;
  tvk::PageOp* ops_d = nullptr; size_t ops_cap = 0;
  // staging for callback-sourced bricks
  void* stage_h = nullptr; void* stage_d = nullptr; size_t stage_bricks = 0;

  // ---- miss reports ----
  uint32_t hash_size = 0;
  uint32_t* hash_d = nullptr;
  uint32_t* miss_d = nullptr;       // (index, value) pairs + count at [2*hash_size]
  uint32_t* miss_h = nullptr;       // pinned mirror
  std::vector<uint32_t> last_missing;   // decoded (x,y,z,lod)

  // ---- classic per-brick path (GLRaycaster) ----
  std::vector<tvk_classic_brick> classic_list;   // last planned brick list (depth sorted)
  uint32_t classic_lod = 0;
  // plan of the last HQ MIP frame: its brick list depends on (LoD, mode, transfer function, isovalue, dataset) only --
  // not on the view -- so a MIP turntable re-plans nothing (AbstrRenderer re-plans every frame; same list)
  struct MipPlan { bool valid = false; uint32_t lod = 0; int mode = 0; uint64_t tf_gen = 0, data_gen = 0; double iso = 0; float sample_rate = 0; } mip_plan;
  uint64_t tf_gen = 0, data_gen = 0;              // bumped by tvk_set_tf1d/2d and by every dataset registration
  std::vector<uint32_t> classic_table_h;          // host copy of the brick -> slot table on the device
  float* classic_axis_d = nullptr; size_t classic_axis_cap = 0;   // per-axis tables (floats + nvox)
  uint32_t* classic_table_d = nullptr; size_t classic_table_cap = 0;

  // ---- frame ----
  tvk_render_params params{};
  bool have_params = false;
  bool blank = true;
  uint32_t img_w = 0, img_h = 0;
  float4* buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // buf: 0 acc/hitpos, 1,2 resume colour/normal ping-pong, 3,4 resume pos ping-pong, 5 hit normal,
  //      6 composed iso image, 7 unused
  int cur = 0;
  uchar4* rgba8_d = nullptr;
  // ClearView state (AbstrRenderer: m_bDoClearView, m_fCVIsovalue, m_vCVColor, m_fCVSize, m_fCVContextScale, m_fCVBorderScale, m_vCVPos)
  struct ClearView { bool on = false; double iso = 0.8; float color[3] = {1, 0, 0}; float size = 5.5f, context = 1.0f, border = 60.0f;
                     float pos[4] = {0, 0, 0.5f, 1.0f}; } cv;
  bool cv_frame = false;          // the last classic frame filled the ClearView targets
  bool clip_plane_on = false; float clip_plane[4] = {0, 0, 1, 0};   // model-space clip plane of the bbox (tvk_set_clip_plane)
  bool stage_mode = false;        // tvk_render_stage in progress: raycast_pass launches a depth-pipeline stage
  const float4* stage_ray_start = nullptr; const float4* stage_color = nullptr;   // its inputs (nullptr: first stage)
  float4* result_buf = nullptr;   // set by frames whose result does not follow the mode rule of result_image() (MIP, stereo)
  float4* stereo_d[2] = {nullptr, nullptr};   // kept eye images (m_pFBO3DImageNext[0 / 1]) of a stereo frame
  uchar4* rgba8_async_d[2] = {nullptr, nullptr};   // double-buffered unorm8 images of the async read-back
  cudaEvent_t read_ev[2] = {nullptr, nullptr}; cudaEvent_t quant_ev = nullptr;
  int read_slot = 0;
  void* read_h = nullptr; size_t read_cap = 0;   // pinned read-back staging
  unsigned long long* counters_d = nullptr;
  unsigned long long* counters_h = nullptr;
  uint32_t* visited_d = nullptr;
  std::vector<uint32_t> visited_h;
  void* unpair_d = nullptr;        // one slot of plain voxels (tvk_read_pool_slot)
  // ---- sort-last (tvk_sortlast.inc) ----
  tvk_sortlast* sl = nullptr;
  float store_clip_min[3] = {0.0f, 0.0f, 0.0f}, store_clip_max[3] = {1.0f, 1.0f, 1.0f};   // tvk_set_store_shard
  std::vector<int32_t> store_index;   // sharded device brick store: TOC index -> store slot or -1 (empty: identity)
  int32_t* store_index_d = nullptr;
  uint64_t store_count = 0;           // bricks kept in the store
};

#endif
