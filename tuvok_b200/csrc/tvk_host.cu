// tvk_host.cu -- the C ABI of libtvkcuda.so (include/tvk.h) and the host-side renderer logic:
// dataset geometry, pool sizing, slot (LRU) table, paging, visibility scheduling, the per-frame
// subframe loop.  Mirrors (reference file:line):
//   GLGridLeaper::{CreateVolumePool,Initialize,RecomputeBrickVisibility,SetupRaycastShader,Raycast,
//                  Render3DRegion}           Renderer/GL/GLGridLeaper.cpp:83-103,247-264,647-870,914-1154
//   GLVolumePool ctor / UploadBrick / PrepareForPaging / RecomputeVisibility / UploadBricks
//                                             Renderer/GL/GLVolumePool.cpp:119-259,673-778,955-959,1580-1789
//   GPUMemMan::GetVolumePool                  Renderer/GPUMemMan/GPUMemMan.cpp:766-844
//   GLHashTable::{ClearData,GetData,Int2Vector}  Renderer/GL/GLHashTable.cpp:65-110
//   ExtendedOctree::ComputeMetadata           IO/UVF/ExtendedOctree/ExtendedOctree.cpp:188-243
//   GLRenderer::ComputeViewAndProjection, CullingLOD::SetScreenParams
// There is no CPU fallback: every data-path operation is a CUDA kernel in k_*.cu.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <chrono>
#include "tvk_host.h"
#include "tvk_synth.cuh"

using namespace tvk;

namespace {

thread_local std::string g_create_err;

void proc_free(tvk_ctx* ctx);   // tvk_procedural.inc
int sl_second_pass(tvk_ctx* ctx);
bool sl_two_launches(const tvk_ctx* ctx);   // PAIRED policy: two concurrent traversal launches per frame   // tvk_sortlast.inc: the rank's second brick block (paired policy), concurrent with the first
int proc_brick_cb(void* user, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap);
const unsigned char* proc_acquire(tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t* cache_slot);
void proc_release(tvk_ctx* ctx, const std::vector<uint32_t>& cache_slots);
void proc_generate(const tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst);

int fail(tvk_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) {
    c->err = buf;
    if (c->log_cb) c->log_cb(c->log_user, 2, "tvk", buf);
  } else {
    g_create_err = buf;
  }
  return code;
}

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? TVK_ERR_OOM : TVK_ERR_CUDA,     \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)


uint32_t esize_of(int dtype) { return dtype == TVK_U8 ? 1u : dtype == TVK_U16 ? 2u : 4u; }

// ExtendedOctree::ComputeMetadata: LOD sizes = ceil(prev/2) per axis (size-1 axes stay) until 1^3
int compute_geometry(tvk_ctx* ctx) {
  uint32_t s[3] = {ctx->vol[0], ctx->vol[1], ctx->vol[2]};
  uint32_t l = 0;
  ctx->toc_offset[0] = 0;
  for (;;) {
    if (l > 0)
      for (int i = 0; i < 3; i++) if (s[i] > 1) s[i] = (s[i] + 1) / 2;
    uint64_t n = 1;
    for (int i = 0; i < 3; i++) {
      ctx->lod_size[l][i] = s[i];
      ctx->layout[l][i] = (s[i] + ctx->inner[i] - 1) / ctx->inner[i];
      n *= ctx->layout[l][i];
    }
    ctx->toc_offset[l + 1] = ctx->toc_offset[l] + n;
    l++;
    if (!(s[0] > 1 || s[1] > 1 || s[2] > 1)) break;
    if (l >= TVK_MAX_LOD) return fail(ctx, TVK_ERR_INVALID, "volume needs more than %d LODs", TVK_MAX_LOD);
  }
  ctx->lod_count = l;
  ctx->n_bricks_all = ctx->toc_offset[l];
  // GetLargestSingleBrickLOD: the finest LOD that is a single brick
  uint32_t single = l - 1;
  for (uint32_t i = 0; i < l; i++)
    if (ctx->layout[i][0] * ctx->layout[i][1] * ctx->layout[i][2] == 1) { single = i; break; }
  ctx->pool_lod_count = single + 1;
  // pool brick layout (GLVolumePool.cpp:109-117): ceil(ceil(vol/inner) / 2^lod)
  uint32_t off = 0;
  for (uint32_t i = 0; i < ctx->pool_lod_count; i++) {
    ctx->lod_offset[i] = off;
    for (int a = 0; a < 3; a++) {
      const uint32_t base = (uint32_t)std::ceil(double(ctx->vol[a]) / ctx->inner[a]);
      ctx->pool_layout[i][a] = (uint32_t)std::ceil(double(base) / double(1u << i));
      if (ctx->pool_layout[i][a] != ctx->layout[i][a])
        return fail(ctx, TVK_ERR_INVALID, "pool and octree brick layouts disagree at LOD %u axis %d (%u vs %u)", i,
                    a, ctx->pool_layout[i][a], ctx->layout[i][a]);
    }
    off += ctx->pool_layout[i][0] * ctx->pool_layout[i][1] * ctx->pool_layout[i][2];
  }
  ctx->total_bricks = ctx->lod_offset[ctx->pool_lod_count - 1] + 1;
  ctx->slot_voxels = (uint64_t)ctx->brick[0] * ctx->brick[1] * ctx->brick[2];
  ctx->slot_bytes = ctx->slot_voxels * ctx->esize;
  // x-pair layout of the scalar integer pools (k_pool.cu); float and colour (uchar4) pools are plain
  ctx->pool_slot_bytes = ctx->slot_bytes * ((ctx->dtype == TVK_F32 || ctx->dtype == TVK_RGBA8) ? 1 : 2);
  return TVK_OK;
}

void brick_size(const tvk_ctx* ctx, const uint32_t co[3], uint32_t lod, uint32_t out[3]) {
  for (int i = 0; i < 3; i++) {
    const uint32_t core = ctx->inner[i];
    const bool last = co[i] == ctx->layout[lod][i] - 1;
    const uint32_t rem = ctx->lod_size[lod][i] % core;
    out[i] = (last && rem) ? 2 * ctx->overlap + rem : ctx->brick[i];
  }
}

void free_dataset(tvk_ctx* c) {
  if (c->minmax_d) cudaFree(c->minmax_d);
  if (c->store_d) cudaFree(c->store_d);
  if (c->store_index_d) cudaFree(c->store_index_d);
  c->minmax_d = nullptr; c->store_d = nullptr; c->store_index_d = nullptr;
  c->store_index.clear(); c->store_count = 0;
  c->minmax_h.clear();
  if (c->file) { delete c->file; c->file = nullptr; }
  proc_free(c);
  c->cb = nullptr; c->cb_user = nullptr;
  c->have_volume = false;
}

void free_pool(tvk_ctx* c) {
  void* p[] = {c->pool_d, c->meta_d, c->slot_brick_d, c->counts_d, c->ops_d, c->stage_d, c->hash_d, c->miss_d, c->visited_d,
               c->unpair_d};
  c->visited_d = nullptr; c->unpair_d = nullptr;
  for (void* q : p) if (q) cudaFree(q);
  if (c->stage_h) cudaFreeHost(c->stage_h);
  if (c->miss_h) cudaFreeHost(c->miss_h);
  c->pool_d = nullptr; c->meta_d = nullptr; c->slot_brick_d = nullptr; c->counts_d = nullptr; c->ops_d = nullptr;
  c->stage_d = nullptr; c->stage_h = nullptr; c->hash_d = nullptr; c->miss_d = nullptr; c->miss_h = nullptr;
  c->ops_cap = 0; c->stage_bricks = 0;
  c->slots.clear(); c->meta_h.clear();
  c->have_pool = false;
}

void free_frame(tvk_ctx* c) {
  for (auto& b : c->buf) { if (b) cudaFree(b); b = nullptr; }
  if (c->rgba8_d) cudaFree(c->rgba8_d);
  c->rgba8_d = nullptr;
  for (auto& b : c->rgba8_async_d) { if (b) cudaFree(b); b = nullptr; }
  for (auto& b : c->stereo_d) { if (b) cudaFree(b); b = nullptr; }
  c->result_buf = nullptr;
  c->img_w = c->img_h = 0;
  if (c->classic_axis_d) cudaFree(c->classic_axis_d);
  if (c->classic_table_d) cudaFree(c->classic_table_d);
  c->classic_axis_d = nullptr; c->classic_table_d = nullptr; c->classic_axis_cap = c->classic_table_cap = 0;
  c->mip_plan.valid = false;      // its device tables are gone
}

int ensure_frame(tvk_ctx* ctx, uint32_t w, uint32_t h) {
  if (ctx->img_w == w && ctx->img_h == h) return TVK_OK;
  free_frame(ctx);
  const size_t n = (size_t)w * h;
  for (int i = 0; i < 8; i++) CU(cudaMalloc(&ctx->buf[i], n * sizeof(float4)));   // [7]: composed stereo frame
  CU(cudaMalloc(&ctx->rgba8_d, n * 4));
  ctx->img_w = w; ctx->img_h = h;
  ctx->blank = true;
  return TVK_OK;
}

int ensure_read(tvk_ctx* ctx, size_t bytes) {
  if (ctx->read_cap >= bytes) return TVK_OK;
  if (ctx->read_h) cudaFreeHost(ctx->read_h);
  ctx->read_h = nullptr; ctx->read_cap = 0;
  CU(cudaMallocHost(&ctx->read_h, bytes));
  ctx->read_cap = bytes;
  return TVK_OK;
}

uint32_t pool_coord(const tvk_ctx* c, const Slot& s) {
  return s.pos[0] + s.pos[1] * c->capacity[0] + s.pos[2] * c->capacity[0] * c->capacity[1];
}
uint32_t brick_id(const tvk_ctx* c, uint32_t x, uint32_t y, uint32_t z, uint32_t lod) {
  return x + y * c->pool_layout[lod][0] + z * c->pool_layout[lod][0] * c->pool_layout[lod][1] + c->lod_offset[lod];
}

bool contains_h(const tvk_ctx* c, const VisState& v, uint32_t id) {
  const double* m = &c->minmax_h[4 * (size_t)id];
  switch (v.mode) {
    case TVK_RM_1DTRANS: return v.v[1] >= m[0] && v.v[0] <= m[1];
    case TVK_RM_2DTRANS: return (v.v[1] >= m[0] && v.v[0] <= m[1]) && (v.v[3] >= m[2] && v.v[2] <= m[3]);
    default: return v.v[0] >= m[0] && v.v[0] <= m[1];
  }
}

// GPUMemMan::GetVolumePool sizing (GPUMemMan.cpp:766-844)
void size_pool(uint64_t max_gpu_mem, uint64_t bit_width, const uint32_t bs[3], uint64_t brick_count, uint32_t max_dim,
               uint32_t out[3]) {
  const uint64_t max_voxels = max_gpu_mem / (bit_width / 8);
  const uint64_t r3v = uint64_t(std::pow(double(max_voxels), 1.0 / 3.0));
  uint64_t gpu[3], ds[3];
  uint64_t m = uint64_t(((float)r3v / bs[0]) + 0.5f) * bs[0];
  if (m > max_dim) m = (max_dim / bs[0]) * bs[0];
  gpu[0] = uint32_t(m);
  m = ((max_voxels / (gpu[0] * gpu[0])) / bs[1]) * bs[1];
  if (m > max_dim) m = (max_dim / bs[1]) * bs[1];
  gpu[1] = uint32_t(m);
  m = ((max_voxels / (gpu[0] * gpu[1])) / bs[2]) * bs[2];
  if (m > max_dim) m = (max_dim / bs[2]) * bs[2];
  gpu[2] = uint32_t(m);
  const uint64_t r3b = uint64_t(std::pow(double(brick_count), 1.0 / 3.0));
  m = bs[0] * r3b;
  if (m > max_dim) m = (max_dim / bs[0]) * bs[0];
  ds[0] = uint32_t(m);
  m = bs[1] * uint64_t(std::ceil(float(brick_count) / ((ds[0] / bs[0]) * (ds[0] / bs[0]))));
  if (m > max_dim) m = (max_dim / bs[1]) * bs[1];
  ds[1] = uint32_t(m);
  m = bs[2] * uint64_t(std::ceil(float(brick_count) / ((ds[0] / bs[0]) * (ds[1] / bs[1]))));
  if (m > max_dim) m = (max_dim / bs[2]) * bs[2];
  ds[2] = uint32_t(m);
  const bool use_ds = ds[0] * ds[1] * ds[2] < gpu[0] * gpu[1] * gpu[2];
  for (int i = 0; i < 3; i++) out[i] = uint32_t(use_ds ? ds[i] : gpu[i]);
}

// Fit1DIndexTo3DArray (GLVolumePool.cpp:816-848)
bool fit_1d_to_3d(uint64_t max_idx, uint32_t max_array, uint32_t out[3]) {
  const uint64_t max_elems = uint64_t(max_array) * max_array * max_array;
  if (max_idx > max_elems) return false;
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
  return true;
}

int ensure_ops(tvk_ctx* ctx, size_t n) {
  if (ctx->ops_cap >= n) return TVK_OK;
  if (ctx->ops_d) cudaFree(ctx->ops_d);
  ctx->ops_d = nullptr; ctx->ops_cap = 0;
  const size_t cap = std::max<size_t>(n, 4096);
  CU(cudaMalloc(&ctx->ops_d, cap * sizeof(PageOp)));
  ctx->ops_cap = cap;
  return TVK_OK;
}

// scatter (index,value) pairs into a device u32 array through the PageOp list
int scatter_u32(tvk_ctx* ctx, uint32_t* dst, const std::vector<std::pair<uint32_t, uint32_t>>& kv) {
  if (kv.empty()) return TVK_OK;
  std::vector<PageOp> ops(kv.size());
  for (size_t i = 0; i < kv.size(); i++) { ops[i] = PageOp{}; ops[i].new_id = kv[i].first; ops[i].slot = kv[i].second; }
  int rc = ensure_ops(ctx, ops.size());
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->ops_d, ops.data(), ops.size() * sizeof(PageOp), cudaMemcpyHostToDevice, ctx->stream));
  launch_page_meta(dst, ctx->ops_d, (uint32_t)ops.size(), ctx->stream);
  CU(cudaStreamSynchronize(ctx->stream));   // ops (pageable host memory) must outlive the copy
  return TVK_OK;
}

struct CopyReq { uint32_t id; uint32_t slot; uint32_t co[4]; };

// Dataset::GetBrick stand-in of the file source: brick (x,y,z,lod) of the open ExtendedOctree file into dst
int file_brick_cb(void* user, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap) {
  tvk_ctx* ctx = static_cast<tvk_ctx*>(user);
  const OctreeFile* f = ctx->file;
  uint32_t bs[3];
  f->brick_size(x, y, z, lod, bs);
  std::string err;
  if (!f->read_brick(f->brick_index(x, y, z, lod), (size_t)bs[0] * bs[1] * bs[2] * ctx->esize, dst, cap, &err)) {
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    ctx->file->error = err;
    return 1;
  }
  return 0;
}

// bricks reqs[0..n) -> hb + i * slot_bytes.  Returns the index of the first failed brick or -1.
long fill_stage(tvk_ctx* ctx, const CopyReq* reqs, size_t n, unsigned char* hb) {
  auto one = [&](size_t i) {
    const CopyReq& r = reqs[i];
    return ctx->cb(ctx->cb_user, r.co[0], r.co[1], r.co[2], r.co[3], hb + i * ctx->slot_bytes, ctx->slot_bytes) == 0;
  };
  // the file source and the procedural source are thread-safe; a user callback (Dataset::GetBrick) is not assumed to be
  const size_t workers = ctx->file ? std::min<size_t>(ctx->io_threads, n) : ctx->proc.on ? std::min<size_t>(ctx->proc.threads, n) : 1;
  if (workers <= 1) {
    for (size_t i = 0; i < n; i++) if (!one(i)) return (long)i;
    return -1;
  }
  std::atomic<size_t> next{0};
  std::atomic<long> bad{-1};
  std::vector<std::thread> pool;
  for (size_t t = 0; t < workers; t++)
    pool.emplace_back([&] {
      for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1))
        if (!one(i)) { long exp = -1; bad.compare_exchange_strong(exp, (long)i); }
    });
  for (std::thread& th : pool) th.join();
  return bad.load();
}

// move the voxels of the requested bricks into their slots
int copy_bricks(tvk_ctx* ctx, const std::vector<CopyReq>& reqs) {
  if (reqs.empty()) return TVK_OK;
  if (ctx->store_d) {
    std::vector<PageOp> ops(reqs.size());
    for (size_t i = 0; i < reqs.size(); i++) {
      ops[i] = PageOp{};
      ops[i].slot = reqs[i].slot;
      uint64_t at = reqs[i].id;                                   // pool id == TOC index for the pool LoDs
      if (!ctx->store_index.empty()) {
        const int32_t k = ctx->store_index[reqs[i].id];
        if (k < 0)
          return fail(ctx, TVK_ERR_SOURCE, "brick (%u,%u,%u,%u) is not in this rank's brick store (tvk_set_store_shard)",
                      reqs[i].co[0], reqs[i].co[1], reqs[i].co[2], reqs[i].co[3]);
        at = (uint64_t)k;
      }
      ops[i].src_off = at * ctx->slot_bytes;
    }
    int rc = ensure_ops(ctx, ops.size());
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->ops_d, ops.data(), ops.size() * sizeof(PageOp), cudaMemcpyHostToDevice, ctx->stream));
    launch_page_copy(ctx->pool_d, ctx->store_d, ctx->ops_d, (uint32_t)ops.size(), ctx->slot_voxels, ctx->esize,
                     ctx->brick, 1, ctx->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return TVK_OK;
  }
  if (!ctx->cb) return fail(ctx, TVK_ERR_INVALID, "no brick source");
  // Dataset::GetBrick -> pinned staging -> async H2D on the copy stream -> scatter into slots.
  // Two half-buffers: the callback fills one half while the other is in flight.
  const size_t half = ctx->stage_bricks / 2;
  const auto wall0 = std::chrono::steady_clock::now();
  cudaEvent_t done[2], c0[2], c1[2];
  bool timed[2] = {false, false};
  CU(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) { CU(cudaEventCreate(&c0[i])); CU(cudaEventCreate(&c1[i])); }
  auto harvest = [&](int i) {   // H2D time of the half's last brick copy (events on the copy stream)
    float ms = 0.0f;
    if (timed[i] && cudaEventElapsedTime(&ms, c0[i], c1[i]) == cudaSuccess) ctx->up_h2d_ms += ms;
    timed[i] = false;
  };
  std::vector<PageOp> ops;
  std::vector<uint32_t> held[2];   // host-cache slots the in-flight copies of each half read from
  double t_wait = 0.0, t_src = 0.0;   // host time waiting for the device / producing bricks (TVK_UPLOAD_TRACE)
  int rc = TVK_OK;
  size_t pos = 0;
  int h = 0;
  while (pos < reqs.size() && rc == TVK_OK) {
    const size_t n = std::min(half, reqs.size() - pos);
    const auto tw0 = std::chrono::steady_clock::now();
    cudaEventSynchronize(done[h]);   // the kernels that read this half last time have finished
    t_wait += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
    harvest(h);
    const auto ts0 = std::chrono::steady_clock::now();
    unsigned char* hb = (unsigned char*)ctx->stage_h + (size_t)h * half * ctx->slot_bytes;
    unsigned char* db = (unsigned char*)ctx->stage_d + (size_t)h * half * ctx->slot_bytes;
    ops.assign(n, PageOp{});
    for (size_t i = 0; i < n; i++) {
      const CopyReq& r = reqs[pos + i];
      uint32_t bs[3];
      brick_size(ctx, r.co, r.co[3], bs);
      ops[i].slot = r.slot;
      ops[i].src_off = (uint64_t)i * ctx->slot_bytes;
      ops[i].size[0] = bs[0]; ops[i].size[1] = bs[1]; ops[i].size[2] = bs[2];
    }
    // fill the pinned half: the file source reads (pread + decode) with several workers straight into pinned
    // memory; a user callback (Dataset::GetBrick) is called from this thread, one brick at a time
    // procedural source with a page-locked host cache: the DMA reads the brick where the cache keeps it -- no staging
    // copy on the host; bricks the cache cannot take (all slots busy) are generated into the pinned half as usual
    const bool direct = ctx->proc.on && ctx->proc.pinned;
    std::vector<const unsigned char*> src;
    if (direct) {
      proc_release(ctx, held[h]);
      held[h].assign(n, 0xFFFFFFFFu);
      src.assign(n, nullptr);
      const CopyReq* rq = &reqs[pos];
      std::atomic<size_t> next{0};
      auto work = [&] {
        for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
          const CopyReq& r = rq[i];
          const auto g0 = std::chrono::steady_clock::now();
          src[i] = proc_acquire(ctx, r.co[0], r.co[1], r.co[2], r.co[3], &held[h][i]);
          if (!src[i]) {
            proc_generate(ctx, r.co[0], r.co[1], r.co[2], r.co[3], hb + i * ctx->slot_bytes);
            ctx->proc.generated.fetch_add(1);
            src[i] = hb + i * ctx->slot_bytes;
          }
          ctx->proc.source_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - g0).count());
        }
      };
      const size_t workers = std::min<size_t>(ctx->proc.threads, n);
      std::vector<std::thread> tp;
      for (size_t t = 1; t < workers; t++) tp.emplace_back(work);
      work();
      for (std::thread& th : tp) th.join();
    } else {
    const long bad = fill_stage(ctx, &reqs[pos], n, hb);
    if (bad >= 0) {
      const CopyReq& r = reqs[pos + (size_t)bad];
      rc = fail(ctx, TVK_ERR_SOURCE, "brick source failed for (%u,%u,%u,%u)%s%s", r.co[0], r.co[1], r.co[2], r.co[3],
                ctx->file ? ": " : "", ctx->file ? ctx->file->error.c_str() : "");
      break;
    }
    }
    t_src += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count();
    // ops travel in the same pinned half (tail) so the copy is truly asynchronous
    const size_t ops_off = (ctx->stage_bricks * ctx->slot_bytes + 15) & ~size_t(15);
    PageOp* hops = (PageOp*)((unsigned char*)ctx->stage_h + ops_off) + (size_t)h * half;
    PageOp* dops = (PageOp*)((unsigned char*)ctx->stage_d + ops_off) + (size_t)h * half;
    std::memcpy(hops, ops.data(), n * sizeof(PageOp));
    cudaEventRecord(c0[h], ctx->copy_stream);
    if (direct) {
      for (size_t i = 0; i < n; i++)
        cudaMemcpyAsync(db + i * ctx->slot_bytes, src[i], (size_t)ops[i].size[0] * ops[i].size[1] * ops[i].size[2] * ctx->esize,
                        cudaMemcpyHostToDevice, ctx->copy_stream);
    } else {
      cudaMemcpyAsync(db, hb, n * ctx->slot_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
    }
    cudaEventRecord(c1[h], ctx->copy_stream);
    timed[h] = true;
    cudaMemcpyAsync(dops, hops, n * sizeof(PageOp), cudaMemcpyHostToDevice, ctx->copy_stream);
    launch_page_copy(ctx->pool_d, db, dops, (uint32_t)n, ctx->slot_voxels, ctx->esize, ctx->brick, 0, ctx->copy_stream);
    cudaEventRecord(done[h], ctx->copy_stream);
    ctx->up_bricks += n; ctx->up_bytes += (uint64_t)n * ctx->slot_bytes;
    pos += n;
    h ^= 1;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->proc.on) { proc_release(ctx, held[0]); proc_release(ctx, held[1]); }
  harvest(0); harvest(1);
  cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
  for (int i = 0; i < 2; i++) { cudaEventDestroy(c0[i]); cudaEventDestroy(c1[i]); }
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  ctx->up_ms += wall;
  static const bool trace = std::getenv("TVK_UPLOAD_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[tvk upload] %zu bricks: wall %.2f ms, source %.2f ms, waiting for the device %.2f ms\n", reqs.size(), wall, t_src, t_wait);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, TVK_ERR_CUDA, "brick upload failed: %s", cudaGetErrorString(e));
  return TVK_OK;
}

// GLVolumePool::UploadBrick bookkeeping (GLVolumePool.cpp:673-717)
void assign_slot(tvk_ctx* c, uint32_t id, size_t pos, uint64_t toc, std::vector<std::pair<uint32_t, uint32_t>>& meta_kv,
                 std::vector<std::pair<uint32_t, uint32_t>>& slot_kv) {
  Slot& s = c->slots[pos];
  if (s.contains_visible()) {
    c->meta_h[s.brick_id] = TVK_BI_MISSING;
    meta_kv.emplace_back((uint32_t)s.brick_id, (uint32_t)TVK_BI_MISSING);
  }
  s.brick_id = (int32_t)id;
  s.time = toc;
  const uint32_t pc = pool_coord(c, s);
  c->meta_h[id] = pc + TVK_BI_FLAG_COUNT;
  meta_kv.emplace_back(id, pc + TVK_BI_FLAG_COUNT);
  slot_kv.emplace_back(pc, id);
}

// later entries for the same index win (sequential semantics of the reference's texel uploads)
void dedup_last(std::vector<std::pair<uint32_t, uint32_t>>& kv) {
  std::vector<std::pair<uint32_t, uint32_t>> out;
  std::stable_sort(kv.begin(), kv.end(), [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) {
    return a.first < b.first;
  });
  for (size_t i = 0; i < kv.size(); i++)
    if (i + 1 == kv.size() || kv[i + 1].first != kv[i].first) out.push_back(kv[i]);
  kv.swap(out);
}

int upload_bricks(tvk_ctx* ctx, const uint32_t* ids, uint32_t n, uint32_t* out_slots, uint32_t* n_paged) {
  uint32_t paged = 0;
  if (out_slots) for (uint32_t i = 0; i < n; i++) out_slots[i] = 0xFFFFFFFFu;
  if (n_paged) *n_paged = 0;
  if (n == 0) return TVK_OK;
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t* q = ids + 4 * i;
    if (q[3] >= ctx->pool_lod_count || q[0] >= ctx->pool_layout[q[3]][0] || q[1] >= ctx->pool_layout[q[3]][1] ||
        q[2] >= ctx->pool_layout[q[3]][2])
      return fail(ctx, TVK_ERR_INVALID, "brick id (%u,%u,%u,%u) out of range", q[0], q[1], q[2], q[3]);
  }
  // PrepareForPaging: oldest first; never used = 0, flagged empty = 1, first brick = UINT64_MAX
  std::sort(ctx->slots.begin(), ctx->slots.end(), [](const Slot& i, const Slot& j) { return i.time < j.time; });
  ctx->insert_pos = 0;
  std::vector<std::pair<uint32_t, uint32_t>> meta_kv, slot_kv;
  std::vector<CopyReq> reqs;
  for (uint32_t i = 0; i < n; i++) {
    if (ctx->insert_pos >= ctx->slots.size() - 1) break;   // all slots but the last replaced this frame
    const uint32_t* q = ids + 4 * i;
    const uint32_t id = brick_id(ctx, q[0], q[1], q[2], q[3]);
    assign_slot(ctx, id, ctx->insert_pos, ctx->time_of_creation++, meta_kv, slot_kv);
    const uint32_t pc = pool_coord(ctx, ctx->slots[ctx->insert_pos]);
    if (out_slots) out_slots[i] = pc;
    CopyReq r; r.id = id; r.slot = pc; r.co[0] = q[0]; r.co[1] = q[1]; r.co[2] = q[2]; r.co[3] = q[3];
    reqs.push_back(r);
    ctx->insert_pos++;
    paged++;
  }
  dedup_last(meta_kv);
  dedup_last(slot_kv);
  int rc = copy_bricks(ctx, reqs);
  if (rc) {
    // The brick source or a copy failed part-way: some target slots are already overwritten, none of the new bricks
    // can be trusted.  Leave a CONSISTENT state behind: every brick that was evicted and every brick that was to be
    // paged in is MISSING on the host mirror and on the device, the touched slots are empty ("never used"), so the next
    // frame simply requests them again.  (The error text of the failure is kept.)
    const std::string why = ctx->err;
    std::vector<std::pair<uint32_t, uint32_t>> undo_meta, undo_slot;
    for (const auto& kv : meta_kv) {
      ctx->meta_h[kv.first] = TVK_BI_MISSING;
      undo_meta.emplace_back(kv.first, (uint32_t)TVK_BI_MISSING);
    }
    for (size_t i = 0; i < reqs.size(); i++) {
      Slot& sl = ctx->slots[i];          // slots [0, insert_pos) of the sorted table took the requests in order
      sl.brick_id = -1; sl.time = 0;
      undo_slot.emplace_back(pool_coord(ctx, sl), 0xFFFFFFFFu);
    }
    ctx->insert_pos = 0;
    if (out_slots) for (uint32_t i = 0; i < n; i++) out_slots[i] = 0xFFFFFFFFu;
    scatter_u32(ctx, ctx->meta_d, undo_meta);
    scatter_u32(ctx, (uint32_t*)ctx->slot_brick_d, undo_slot);
    ctx->blank = true;
    ctx->err = why;
    return rc;
  }
  rc = scatter_u32(ctx, ctx->meta_d, meta_kv);
  if (rc) return rc;
  rc = scatter_u32(ctx, (uint32_t*)ctx->slot_brick_d, slot_kv);
  if (rc) return rc;
  if (n_paged) *n_paged = paged;
  return TVK_OK;
}

int recompute_visibility(tvk_ctx* ctx, int force, uint32_t counts[4]) {
  if (counts) counts[0] = counts[1] = counts[2] = counts[3] = 0;
  if (!ctx->have_pool) return fail(ctx, TVK_ERR_INVALID, "no pool");
  if (!ctx->have_params) return fail(ctx, TVK_ERR_INVALID, "no render params");
  if (!ctx->tf1d_d) return fail(ctx, TVK_ERR_INVALID, "no 1D transfer function (needed for the rescale factor)");
  // GLGridLeaper::RecomputeBrickVisibility (GLGridLeaper.cpp:647-687); note the 1D TF size is used
  // for the rescale factor in every mode
  const double max_value = ctx->range_max;
  const double rescale = max_value / double(ctx->tf1d_n - 1);
  int mode = ctx->params.mode;
  double a = 0, b = 0, c = 0, d = 0;
  switch (mode) {
    case TVK_RM_1DTRANS: a = double(ctx->tf1d_nz[0]) * rescale; b = double(ctx->tf1d_nz[1]) * rescale; break;
    case TVK_RM_2DTRANS:
      if (!ctx->tf2d_d) return fail(ctx, TVK_ERR_INVALID, "no 2D transfer function");
      a = double(ctx->tf2d_nz[0]) * rescale; b = double(ctx->tf2d_nz[1]) * rescale;
      c = double(ctx->tf2d_nz[2]); d = double(ctx->tf2d_nz[3]);
      break;
    case TVK_RM_ISOSURFACE: a = ctx->params.isovalue; break;
    default: return fail(ctx, TVK_ERR_INVALID, "Unhandled rendering mode.");
  }
  if (!ctx->vis.needs_update(mode, a, b, c, d) && !force) return TVK_OK;

  VisConsts vc{};
  vc.mode = mode;
  vc.v[0] = a; vc.v[1] = b; vc.v[2] = c; vc.v[3] = d;
  vc.lod_count = ctx->pool_lod_count;
  for (uint32_t l = 0; l < ctx->pool_lod_count; l++) {
    vc.lod_offset[l] = ctx->lod_offset[l];
    for (int i = 0; i < 3; i++) vc.layout[l][i] = ctx->pool_layout[l][i];
  }
  // host side of RecomputeVisibilityForBrickPool: only the LRU flags (the table is written on the device)
  for (Slot& s : ctx->slots) {
    if (!s.was_ever_used()) continue;
    const bool has = contains_h(ctx, ctx->vis, (uint32_t)s.brick_id);
    const bool had = s.contains_visible();
    if (has) { if (!had) s.restore(); }
    else { if (had) s.flag_empty(); }
  }
  CU(cudaMemsetAsync(ctx->counts_d, 0, 4 * sizeof(uint32_t), ctx->stream));
  launch_vis_clear(ctx->meta_d, ctx->meta_count, ctx->stream);
  launch_vis_pool(ctx->meta_d, ctx->slot_brick_d, ctx->n_slots, ctx->minmax_d, vc, ctx->stream);
  for (uint32_t l = 0; l < ctx->pool_lod_count; l++)
    launch_vis_level(ctx->meta_d, ctx->minmax_d, vc, l, ctx->counts_d, ctx->stream);
  CU(cudaGetLastError());
  uint32_t cnt[4];
  CU(cudaMemcpyAsync(cnt, ctx->counts_d, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->meta_h.data(), ctx->meta_d, ctx->meta_count * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (counts) std::memcpy(counts, cnt, sizeof(cnt));
  ctx->blank = true;   // a new table invalidates the resume state
  return TVK_OK;
}

// SetupRaycastShader + GLVolumePool::Enable + the #defines of the generated pool GLSL
int derive(tvk_ctx* ctx, RayConsts& u) {
  const tvk_render_params& p = ctx->params;
  std::memset(&u, 0, sizeof(u));
  u.width = p.width; u.height = p.height;
  double mv[16], pr[16], imv[16], ipr[16], emm[16], m2e[16];
  for (int i = 0; i < 16; i++) { mv[i] = p.model_view[i]; pr[i] = p.projection[i]; }
  if (!inv4(mv, imv) || !inv4(pr, ipr)) return fail(ctx, TVK_ERR_INVALID, "singular view or projection matrix");
  float ex[3], sc[3];
  for (int i = 0; i < 3; i++) ex[i] = (float)ctx->vol[i] * ctx->scale[i];
  const float mx = fmaxf(ex[0], fmaxf(ex[1], ex[2]));
  for (int i = 0; i < 3; i++) ex[i] = ex[i] / mx;
  const float mn = fminf(ctx->scale[0], fminf(ctx->scale[1], ctx->scale[2]));
  for (int i = 0; i < 3; i++) sc[i] = ctx->scale[i] / mn;
  for (int i = 0; i < 3; i++) u.domain_scale[i] = 1.0f / sc[i];
  // mEyeToModel = inverse(MV) * T(-centre = 0) * S(1/extend) * T(.5)  (GLGridLeaper.cpp:560-573)
  double s[16] = {0}, t[16] = {0};
  s[0] = (double)(1.0f / ex[0]); s[5] = (double)(1.0f / ex[1]); s[10] = (double)(1.0f / ex[2]); s[15] = 1;
  t[0] = t[5] = t[10] = t[15] = 1; t[12] = t[13] = t[14] = 0.5;
  mul4(imv, s, emm);
  mul4(emm, t, emm);
  if (!inv4(emm, m2e)) return fail(ctx, TVK_ERR_INVALID, "singular eye-to-model matrix");
  for (int i = 0; i < 16; i++) {
    u.emm[i] = (float)emm[i]; u.inv_proj[i] = (float)ipr[i]; u.m2e[i] = (float)m2e[i]; u.mv_inv[i] = (float)imv[i];
  }
  for (int i = 0; i < 3; i++) {
    u.light_a[i] = p.ambient[i] * p.ambient[3];
    u.light_d[i] = p.diffuse[i] * p.diffuse[3];
    u.light_s[i] = p.specular[i] * p.specular[3];
  }
  {  // (vec4(lightDir,0) * emm).xyz normalised; (vec4(eye,1) * emm).xyz  (GLGridLeaper.cpp:741-742)
    const float* m = u.emm;
    float l[3], e[3];
    for (int c = 0; c < 3; c++) {
      l[c] = p.light_dir[0] * m[c] + p.light_dir[1] * m[4 + c] + p.light_dir[2] * m[8 + c] + 0.0f * m[12 + c];
      e[c] = p.eye[0] * m[c] + p.eye[1] * m[4 + c] + p.eye[2] * m[8 + c] + 1.0f * m[12 + c];
    }
    // normalize() of the arithmetic contract: v * (1 / sqrt(fma(z,z, fma(y,y, x*x))))
    const float inv = 1.0f / sqrtf(fmaf(l[2], l[2], fmaf(l[1], l[1], l[0] * l[0])));
    for (int c = 0; c < 3; c++) { u.light_dir_m[c] = l[c] * inv; u.eye_m[c] = e[c]; }
  }
  u.lzwse = fmaxf(ex[0] / (float)ctx->vol[0], fmaxf(ex[1] / (float)ctx->vol[1], ex[2] / (float)ctx->vol[2]));
  u.lod_factor = p.lod_factor;
  for (int i = 0; i < 3; i++) {
    u.pool_size_f[i] = (float)ctx->pool_size[i];
    u.vol_f[i] = (float)ctx->vol[i];
    u.overlap_tc[i] = (ctx->brick[i] - ctx->inner[i]) / (2.0f * ctx->pool_size[i]);
    u.capacity[i] = ctx->capacity[i];
    u.total[i] = ctx->brick[i];
    u.ghost[i] = ctx->overlap;
    u.finest[i] = ctx->pool_layout[0][i];
    u.clip_min[i] = p.clip_min[i]; u.clip_max[i] = p.clip_max[i];
    if (p.clip_min[i] > 0.0f || p.clip_max[i] < 1.0f) u.shard = 1;
    u.sh_lo[i] = p.clip_min[i] > 0.0f ? p.clip_min[i] : -INFINITY;
    u.sh_hi[i] = p.clip_max[i] < 1.0f ? p.clip_max[i] : INFINITY;
  }
  if (ctx->clip_plane_on) {
    // model space (centre 0, extent ex) -> the [0,1]^3 coordinates of ray_setup: p_model = (p - 0.5) * ex
    u.clip_plane_on = 1;
    const float* q = ctx->clip_plane;
    for (int i = 0; i < 3; i++) u.clip_plane[i] = q[i] * ex[i];
    u.clip_plane[3] = q[3] - 0.5f * (u.clip_plane[0] + u.clip_plane[1] + u.clip_plane[2]);
  }
  u.lod_count = ctx->pool_lod_count;
  for (uint32_t l = 0; l < ctx->pool_lod_count; l++) {
    u.lod_offset[l] = ctx->lod_offset[l];
    float c[3];
    for (int i = 0; i < 3; i++) {   // GetFloatBrickLayout (GLVolumePool.cpp:89-107)
      c[i] = (float)ctx->vol[i] / ctx->inner[i];
      c[i] = c[i] / (float)(1u << l);
      if ((float)(uint32_t)c[i] == c[i]) c[i] = c[i] - c[i] * std::numeric_limits<float>::epsilon();
      u.lod_layout[l][i] = c[i];
    }
    u.lod_layout_sz[l][0] = (uint32_t)ceilf(c[0]);
    u.lod_layout_sz[l][1] = (uint32_t)ceilf(c[0]) * (uint32_t)ceilf(c[1]);
  }
  u.norm = (ctx->dtype == TVK_U8 || ctx->dtype == TVK_RGBA8) ? 1.0f / 255.0f : ctx->dtype == TVK_U16 ? 1.0f / 65535.0f : 1.0f;
  u.sample_rate = p.sample_rate_modifier;
  u.oc = 1.0f / p.sample_rate_modifier;
  // GLRenderer::CalculateScaling (GLRenderer.cpp:1813-1819): (2^bits - 1) / maxValue; float data: 1/maxValue (H7)
  const double full = (ctx->dtype == TVK_U8 || ctx->dtype == TVK_RGBA8) ? 255.0 : ctx->dtype == TVK_U16 ? 65535.0 : 1.0;
  u.trans_scale = (float)(full / ctx->range_max);
  u.gradient_scale = ctx->max_grad == 0.0f ? 1.0f : 1.0f / ctx->max_grad;
  // GetNormalizedIsovalue (AbstrRenderer.cpp:412-424): iso / 2^bits; float data: iso itself (H7)
  u.isoval = (ctx->dtype == TVK_U8 || ctx->dtype == TVK_RGBA8) ? (float)(p.isovalue / 256.0) : ctx->dtype == TVK_U16 ? (float)(p.isovalue / 65536.0)
                                                                                       : (float)p.isovalue;
  if (p.mode == TVK_RM_2DTRANS) { u.tf = ctx->tf2d_d; u.tf_w = ctx->tf2d_w; u.tf_h = ctx->tf2d_h; }
  else { u.tf = ctx->tf1d_d; u.tf_w = ctx->tf1d_n; u.tf_h = 1; }
  u.hash_size = ctx->hash_size;
  u.rehash_count = ctx->cfg.rehash_count;
  u.strategy = ctx->cfg.brick_strategy;
  u.nearest = p.nearest;
  u.count = ctx->counters_on ? 1 : 0;
  u.pool = ctx->pool_d;
  u.slot_voxels = ctx->slot_voxels;
  u.meta = ctx->meta_d;
  u.hash = ctx->hash_d;
  u.counters = ctx->counters_d;
  u.visited = ctx->visited_d;
  // one tile counter per launch in flight (concurrent launches of the PAIRED policy must not share one)
  u.tile_counter = reinterpret_cast<uint32_t*>(ctx->counters_d + 8) + (ctx->launch_seq++ & 15u);
  return TVK_OK;
}

int check_renderable(tvk_ctx* ctx) {
  if (!ctx->have_volume) return fail(ctx, TVK_ERR_INVALID, "no dataset registered");
  if (!ctx->have_pool) return fail(ctx, TVK_ERR_INVALID, "no volume pool");
  if (!ctx->have_params) return fail(ctx, TVK_ERR_INVALID, "no render params");
  const int m = ctx->params.mode;
  if (m == TVK_RM_1DTRANS && !ctx->tf1d_d) return fail(ctx, TVK_ERR_INVALID, "no 1D transfer function");
  if (m == TVK_RM_2DTRANS && !ctx->tf2d_d) return fail(ctx, TVK_ERR_INVALID, "no 2D transfer function");
  return TVK_OK;
}

// one raycast pass into the frame buffers (GLGridLeaper::Raycast, GLGridLeaper.cpp:754-870)
int raycast_pass(tvk_ctx* ctx, bool with_hash) {
  int rc = ensure_frame(ctx, ctx->params.width, ctx->params.height);
  if (rc) return rc;
  RayConsts u;
  rc = derive(ctx, u);
  if (rc) return rc;
  if (!with_hash) { u.hash = nullptr; u.hash_size = 0; }
  const bool iso = ctx->params.mode == TVK_RM_ISOSURFACE;
  const int cur = ctx->cur, nxt = cur ^ 1;
  u.first_pass = ctx->blank ? 1 : 0;
  u.ray_start = ctx->buf[3 + cur];
  u.start_color = ctx->buf[1 + cur];
  u.out0 = ctx->buf[0];
  if (!iso) { u.out1 = ctx->buf[1 + nxt]; u.out2 = ctx->buf[3 + nxt]; u.out3 = nullptr; }
  else { u.out1 = ctx->buf[5]; u.out2 = ctx->buf[3 + nxt]; u.out3 = ctx->buf[1 + nxt]; }
  if (ctx->stage_mode) {   // one stage of the depth pipeline (tvk_render_stage): inputs from the stage in front
    if (iso) return fail(ctx, TVK_ERR_INVALID, "depth pipeline: transfer-function modes only");
    u.pipeline = 1;
    u.first_pass = ctx->stage_ray_start ? 0 : 1;
    u.ray_start = ctx->stage_ray_start ? ctx->stage_ray_start : ctx->buf[3];
    u.start_color = ctx->stage_color ? ctx->stage_color : ctx->buf[1];
    u.out1 = ctx->buf[1]; u.out2 = ctx->buf[3];
  }
  // LPT tile schedule: this launch runs the tiles in the order of the cost the previous launch measured, and measures anew
  u.tile_order = nullptr; u.tile_cost = nullptr;
  const bool lpt = ctx->tile_lpt && ctx->dtype != TVK_RGBA8 && !sl_two_launches(ctx);
  if (lpt) {
    const uint32_t n_tiles = raycast_tiles(u.width, u.height);
    if (n_tiles != ctx->tile_n) {
      if (ctx->tile_cost_d) cudaFree(ctx->tile_cost_d);
      for (auto& o : ctx->tile_order_d) { if (o) cudaFree(o); o = nullptr; }
      ctx->tile_cost_d = nullptr; ctx->tile_n = 0; ctx->tile_valid = false;
      CU(cudaMalloc(&ctx->tile_cost_d, (size_t)n_tiles * 4));
      for (auto& o : ctx->tile_order_d) CU(cudaMalloc(&o, ((size_t)n_tiles + 1) * 4));   // + the split count
      ctx->tile_n = n_tiles;
    }
    CU(cudaMemsetAsync(ctx->tile_cost_d, 0, (size_t)n_tiles * 4, ctx->stream));
    u.tile_cost = ctx->tile_cost_d;
    if (ctx->tile_valid) u.tile_order = ctx->tile_order_d[ctx->tile_cur];
  }
  if (ctx->dtype == TVK_RGBA8) {
    if (u.shard || u.pipeline || ctx->counters_on)
      return fail(ctx, TVK_ERR_INVALID, "colour volumes: sort-last shards, pipeline stages and counters are not built");
    launch_raycast_color(u, ctx->params.mode, ctx->params.lighting, ctx->stream);
  } else {
    launch_raycast(u, ctx->params.mode, ctx->params.lighting, ctx->dtype, ctx->stream);
  }
  if (lpt) {
    launch_tile_order(ctx->tile_cost_d, ctx->tile_n, ctx->tile_order_d[ctx->tile_cur ^ 1], 2, ctx->tile_split_cost, ctx->stream);
    ctx->tile_cur ^= 1;
    ctx->tile_valid = true;
  }
  CU(cudaGetLastError());
  if (ctx->stage_mode) { ctx->blank = true; return TVK_OK; }   // no resume state of its own: every stage frame starts anew
  if (iso) {   // GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2830)
    const tvk_render_params& p = ctx->params;
    float a[3], d[3], s[3];
    for (int i = 0; i < 3; i++) {
      a[i] = p.ambient[i] * p.ambient[3];
      // colour data: Compose-Color-FS with the plain diffuse light, the surface colour comes out of the hit buffers
      // (GLRenderer.cpp:2796-2810)
      d[i] = p.diffuse[i] * p.diffuse[3] * (ctx->dtype == TVK_RGBA8 ? 1.0f : p.iso_color[i]);
      s[i] = p.specular[i] * p.specular[3];
    }
    launch_iso_compose(ctx->buf[0], ctx->buf[5], ctx->buf[6], p.width, p.height, a, d, s, p.light_dir, ctx->stream,
                       ctx->dtype == TVK_RGBA8);
    CU(cudaGetLastError());
  }
  ctx->cur = nxt;      // swap current/next resume buffers (GLGridLeaper.cpp:861-862)
  ctx->blank = false;
  return TVK_OK;
}

float4* result_image(tvk_ctx* ctx) {
  if (ctx->result_buf) return ctx->result_buf;
  return ctx->params.mode == TVK_RM_ISOSURFACE ? ctx->buf[6] : ctx->buf[0];
}

}  // namespace

// =================================================================================================
extern "C" {

uint32_t tvk_abi_version(void) { return TVK_ABI_VERSION; }

const char* tvk_last_error(const tvk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tvk_create(const tvk_device_cfg* cfg, tvk_ctx** out) {
  if (!out) return fail(nullptr, TVK_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, TVK_ERR_NO_DEVICE, "no CUDA device (%s); libtvkcuda has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  tvk_ctx* ctx = new tvk_ctx();
  if (cfg) ctx->cfg = *cfg;
  if (ctx->cfg.device < 0 || ctx->cfg.device >= n_dev) {
    fail(nullptr, TVK_ERR_INVALID, "device %d out of range (0..%d)", ctx->cfg.device, n_dev - 1);
    delete ctx;
    return TVK_ERR_INVALID;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, ctx->cfg.device);
  if (prop.major != 10) {
    fail(nullptr, TVK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", ctx->cfg.device,
         prop.major, prop.minor);
    delete ctx;
    return TVK_ERR_NO_DEVICE;
  }
  if (ctx->cfg.max_gpu_mem == 0) ctx->cfg.max_gpu_mem = 8ull << 30;   // SystemInfo default
  if (ctx->cfg.max_pool_dim == 0) ctx->cfg.max_pool_dim = 16384;
  // LPT tile schedule of the traversal kernel: on unless TVK_TILE_LPT=0 (measured: C3 303.7 -> 313.5 fps on one GPU,
  // 735 -> 770 fps at N = 8, images bit-identical; profiles/r3a_lpt_ab.txt)
  { const char* e = std::getenv("TVK_TILE_LPT"); ctx->tile_lpt = (e && e[0] == '0') ? 0 : 1; }
  { const char* e = std::getenv("TVK_SPLIT_COST"); ctx->tile_split_cost = e ? (uint32_t)std::atoi(e) : 0u; }
  if (ctx->cfg.hash_table_size == 0) ctx->cfg.hash_table_size = 509;
  if (ctx->cfg.rehash_count == 0) ctx->cfg.rehash_count = 10;
  if (!cfg) ctx->cfg.brick_strategy = TVK_BS_SKIP_TWO_LEVELS;
  bool ok = cudaSetDevice(ctx->cfg.device) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& ev : ctx->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->counters_d, 8 * sizeof(unsigned long long) + 16 * sizeof(uint32_t)) == cudaSuccess;   // + tile counters
  ok = ok && cudaMallocHost(&ctx->counters_h, 8 * sizeof(unsigned long long)) == cudaSuccess;
  if (!ok) {
    fail(nullptr, TVK_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return TVK_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return TVK_OK;
}

void tvk_destroy(tvk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaDeviceSynchronize();
  tvk_sortlast_shutdown(ctx);
  free_frame(ctx);
  free_pool(ctx);
  free_dataset(ctx);
  if (ctx->tf1d_d) cudaFree(ctx->tf1d_d);
  if (ctx->tf2d_d) cudaFree(ctx->tf2d_d);
  if (ctx->read_h) cudaFreeHost(ctx->read_h);
  if (ctx->counters_d) cudaFree(ctx->counters_d);
  if (ctx->tile_cost_d) cudaFree(ctx->tile_cost_d);
  for (auto& o : ctx->tile_order_d) if (o) cudaFree(o);
  if (ctx->counters_h) cudaFreeHost(ctx->counters_h);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& e : ctx->read_ev) if (e) cudaEventDestroy(e);
  if (ctx->quant_ev) cudaEventDestroy(ctx->quant_ev);
  delete ctx;
}

int tvk_set_log_callback(tvk_ctx* ctx, tvk_log_cb cb, void* user) {
  if (!ctx) return TVK_ERR_INVALID;
  ctx->log_cb = cb; ctx->log_user = user;
  return TVK_OK;
}

int tvk_set_stream(tvk_ctx* ctx, void* cuda_stream) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return TVK_OK;
}

int tvk_synchronize(tvk_ctx* ctx) {
  if (!ctx) return TVK_ERR_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  return TVK_OK;
}

int tvk_enable_counters(tvk_ctx* ctx, int enable) {
  if (!ctx) return TVK_ERR_INVALID;
  ctx->counters_on = enable != 0;
  return TVK_OK;
}

// ---- dataset ------------------------------------------------------------------------------------
static int set_geometry(tvk_ctx* ctx, const uint32_t size[3], const float scale[3], const uint32_t max_brick[3],
                        uint32_t overlap, int dtype, double range_max, float max_grad) {
  if (dtype < TVK_U8 || dtype > TVK_RGBA8) return fail(ctx, TVK_ERR_INVALID, "unsupported dtype %d", dtype);
  for (int i = 0; i < 3; i++) {
    if (size[i] == 0) return fail(ctx, TVK_ERR_INVALID, "empty domain");
    if (max_brick[i] <= 2 * overlap) return fail(ctx, TVK_ERR_INVALID, "brick size must exceed 2*overlap");
    if ((max_brick[i] - 2 * overlap) % 2) return fail(ctx, TVK_ERR_INVALID, "odd inner brick sizes are not supported");
  }
  cudaSetDevice(ctx->cfg.device);
  free_pool(ctx);
  free_dataset(ctx);
  for (int i = 0; i < 3; i++) {
    ctx->vol[i] = size[i]; ctx->scale[i] = scale ? scale[i] : 1.0f; ctx->brick[i] = max_brick[i];
    ctx->inner[i] = max_brick[i] - 2 * overlap;
  }
  ctx->overlap = overlap;
  ctx->dtype = dtype;
  ctx->esize = esize_of(dtype);
  ctx->range_max = range_max > 0 ? range_max : ((dtype == TVK_U8 || dtype == TVK_RGBA8) ? 255.0 : dtype == TVK_U16 ? 65535.0 : 1.0);
  ctx->max_grad = max_grad;
  return compute_geometry(ctx);
}

int tvk_set_volume(tvk_ctx* ctx, const tvk_volume_desc* d, tvk_brick_cb cb, void* user) {
  if (!ctx || !d || !cb || !d->minmax) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  int rc = set_geometry(ctx, d->domain_size, d->scale, d->max_brick_size, d->overlap, d->dtype, d->range_max,
                        d->max_gradient_magnitude);
  if (rc) return rc;
  if (d->brick_count < ctx->total_bricks)
    return fail(ctx, TVK_ERR_INVALID, "min/max table has %llu entries, the pool LoDs need %u",
                (unsigned long long)d->brick_count, ctx->total_bricks);
  ctx->n_bricks_all = d->brick_count;
  ctx->minmax_h.assign(d->minmax, d->minmax + 4 * (size_t)ctx->total_bricks);
  CU(cudaMalloc(&ctx->minmax_d, ctx->minmax_h.size() * sizeof(double)));
  CU(cudaMemcpy(ctx->minmax_d, ctx->minmax_h.data(), ctx->minmax_h.size() * sizeof(double), cudaMemcpyHostToDevice));
  ctx->cb = cb; ctx->cb_user = user;
  ctx->have_volume = true; ctx->data_gen++;
  return TVK_OK;
}

static void fill_file_info(const OctreeFile& g, tvk_octree_file_info* info);

// min/max table of a streamed dataset that comes without a MaxMin block: every brick of the pool LoDs goes once
// through the pinned staging path (double buffered, parallel reads) and is reduced on the device
static int minmax_from_source(tvk_ctx* ctx) {
  const size_t nb_total = ctx->total_bricks;
  size_t nb = std::max<size_t>(2, ((64ull << 20) / ctx->slot_bytes) & ~size_t(1));
  nb = std::min(nb, std::max<size_t>(2, (nb_total + 1) & ~size_t(1)));
  const size_t half = nb / 2;
  const size_t ops_off = (nb * ctx->slot_bytes + 15) & ~size_t(15);
  const size_t bytes = ops_off + nb * sizeof(PageOp);
  unsigned char *sh = nullptr, *sd = nullptr;
  CU(cudaMallocHost(&sh, bytes));
  if (cudaMalloc(&sd, bytes) != cudaSuccess) { cudaFreeHost(sh); return fail(ctx, TVK_ERR_OOM, "staging allocation failed"); }
  cudaEvent_t done[2];
  cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
  int rc = TVK_OK;
  std::vector<CopyReq> reqs;
  size_t pos = 0;
  int h = 0;
  while (pos < nb_total && rc == TVK_OK) {
    const size_t n = std::min(half, nb_total - pos);
    cudaEventSynchronize(done[h]);
    unsigned char* hb = sh + (size_t)h * half * ctx->slot_bytes;
    PageOp* hops = (PageOp*)(sh + ops_off) + (size_t)h * half;
    reqs.assign(n, CopyReq{});
    for (size_t i = 0; i < n; i++) {
      const uint32_t id = (uint32_t)(pos + i);
      uint32_t lod = 0;
      while (lod + 1 < ctx->pool_lod_count && id >= ctx->lod_offset[lod + 1]) lod++;
      const uint32_t local = id - ctx->lod_offset[lod];
      const uint32_t* l = ctx->pool_layout[lod];
      CopyReq& r = reqs[i];
      r.id = id; r.slot = 0;
      r.co[0] = local % l[0]; r.co[1] = (local / l[0]) % l[1]; r.co[2] = local / (l[0] * l[1]); r.co[3] = lod;
      uint32_t bs[3];
      brick_size(ctx, r.co, lod, bs);
      hops[i] = PageOp{};
      hops[i].new_id = id;
      hops[i].src_off = (uint64_t)i * ctx->slot_bytes;
      hops[i].size[0] = bs[0]; hops[i].size[1] = bs[1]; hops[i].size[2] = bs[2];
    }
    const long bad = fill_stage(ctx, reqs.data(), n, hb);
    if (bad >= 0) {
      const CopyReq& r = reqs[(size_t)bad];
      rc = fail(ctx, TVK_ERR_SOURCE, "brick source failed for (%u,%u,%u,%u)%s%s", r.co[0], r.co[1], r.co[2], r.co[3],
                ctx->file ? ": " : "", ctx->file ? ctx->file->error.c_str() : "");
      break;
    }
    unsigned char* db = sd + (size_t)h * half * ctx->slot_bytes;
    PageOp* dops = (PageOp*)(sd + ops_off) + (size_t)h * half;
    cudaMemcpyAsync(db, hb, n * ctx->slot_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
    cudaMemcpyAsync(dops, hops, n * sizeof(PageOp), cudaMemcpyHostToDevice, ctx->copy_stream);
    launch_brick_minmax(db, dops, (uint32_t)n, ctx->minmax_d, ctx->dtype, ctx->copy_stream);
    cudaEventRecord(done[h], ctx->copy_stream);
    pos += n;
    h ^= 1;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
  cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
  cudaFreeHost(sh); cudaFree(sd);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, TVK_ERR_CUDA, "min/max pass failed: %s", cudaGetErrorString(e));
  return TVK_OK;
}

static int tvk_open_octree_file_impl(tvk_ctx* ctx, const char* path, uint64_t offset, uint64_t uvf_file_version, const float scale[3],
                         const double* minmax, uint64_t n_minmax, double range_max, float max_gradient_magnitude,
                         tvk_octree_file_info* info) {
  if (!ctx || !path) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  std::unique_ptr<OctreeFile> f(new OctreeFile());
  if (!f->open(path, offset, uvf_file_version)) return fail(ctx, TVK_ERR_SOURCE, "%s: %s", path, f->error.c_str());
  int dtype = -1;
  switch (f->component_type) {   // ExtendedOctree::COMPONENT_TYPE (ExtendedOctree.h:137-148)
    case 0: dtype = TVK_U8; break;
    case 1: dtype = TVK_U16; break;
    case 8: dtype = TVK_F32; break;
    default: break;
  }
  // colour data (AbstrRenderer::ColorData): four interleaved 8-bit components -> the colour kernels (k_color.cu)
  // (a four-component file with the precomputed-normals flag holds value + normal, not colour: refused)
  if (f->component_type == 0 && f->component_count == 4 && !f->precomputed_normals) dtype = TVK_RGBA8;
  else if (f->component_count != 1) dtype = -1;
  if (dtype < 0)
    return fail(ctx, TVK_ERR_INVALID, "%s: component type %u x %llu is not on the hot path (u8 / u16 / f32 scalar, 4 x u8 colour)", path,
                f->component_type, (unsigned long long)f->component_count);
  uint32_t size[3], brick[3];
  float sc[3];
  for (int i = 0; i < 3; i++) {
    if (f->vol[i] > 0xffffffffull || f->brick[i] > 0xffffffffull) return fail(ctx, TVK_ERR_INVALID, "%s: size overflow", path);
    size[i] = (uint32_t)f->vol[i]; brick[i] = (uint32_t)f->brick[i];
    sc[i] = scale ? scale[i] : (float)f->aspect[i];      // UVFDataset: domain scale = the octree's volume aspect
  }
  int rc = set_geometry(ctx, size, sc, brick, f->overlap, dtype, range_max, max_gradient_magnitude);
  if (rc) return rc;
  if (f->toc.size() != ctx->n_bricks_all || f->lod_count() != ctx->lod_count)
    return fail(ctx, TVK_ERR_INVALID, "%s: table of contents has %zu bricks / %u LoDs, the geometry implies %llu / %u", path,
                f->toc.size(), f->lod_count(), (unsigned long long)ctx->n_bricks_all, ctx->lod_count);
  ctx->file = f.release();
  ctx->cb = file_brick_cb; ctx->cb_user = ctx;
  CU(cudaMalloc(&ctx->minmax_d, 4 * (size_t)ctx->total_bricks * sizeof(double)));
  ctx->minmax_h.resize(4 * (size_t)ctx->total_bricks);
  if (minmax) {
    if (n_minmax < ctx->total_bricks) {
      free_dataset(ctx);
      return fail(ctx, TVK_ERR_INVALID, "min/max table has %llu entries, the pool LoDs need %u",
                  (unsigned long long)n_minmax, ctx->total_bricks);
    }
    std::memcpy(ctx->minmax_h.data(), minmax, ctx->minmax_h.size() * sizeof(double));
    CU(cudaMemcpy(ctx->minmax_d, ctx->minmax_h.data(), ctx->minmax_h.size() * sizeof(double), cudaMemcpyHostToDevice));
  } else {
    rc = minmax_from_source(ctx);
    if (rc) { free_dataset(ctx); return rc; }
    CU(cudaMemcpy(ctx->minmax_h.data(), ctx->minmax_d, ctx->minmax_h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (info) { fill_file_info(*ctx->file, info); info->dtype = dtype; }
  ctx->have_volume = true; ctx->data_gen++;
  return TVK_OK;
}
int tvk_open_octree_file(tvk_ctx* ctx, const char* path, uint64_t offset, uint64_t uvf_file_version, const float scale[3],
                         const double* minmax, uint64_t n_minmax, double range_max, float max_gradient_magnitude,
                         tvk_octree_file_info* info) {
  try { return tvk_open_octree_file_impl(ctx, path, offset, uvf_file_version, scale, minmax, n_minmax, range_max, max_gradient_magnitude, info); }
  catch (const std::bad_alloc&) { return fail(ctx, TVK_ERR_OOM, "tvk_open_octree_file: out of host memory"); }
  catch (const std::exception& e) { return fail(ctx, TVK_ERR_SOURCE, "tvk_open_octree_file: %s", e.what()); }
}

static int tvk_open_uvf_impl(tvk_ctx* ctx, const char* path, uint64_t timestep, const float scale[3], double range_max,
                 float max_gradient_magnitude, tvk_octree_file_info* info) {
  if (!ctx || !path) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  UvfScan sc;
  if (!uvf_scan(path, timestep, &sc)) return fail(ctx, TVK_ERR_SOURCE, "%s: %s", path, sc.error.c_str());
  // what UVFDataset takes from the file when the caller does not override it: the 2D histogram block's maximum
  // gradient magnitude (UVFDataset::GetMaxGradMagnitude -> fGradientScale of the 2D-TF shaders) and the value range
  // (UVFDataset::ComputeRange, IO/uvfDataset.cpp:1120-1155: min / max over the LoD-0 bricks of the MaxMin block)
  if (!(max_gradient_magnitude > 0.0f) && sc.have_hist2d) max_gradient_magnitude = sc.max_grad_magnitude;
  if (!(range_max > 0.0) && sc.have_maxmin) {
    OctreeFile f;
    if (!f.open(path, sc.toc_payload_offset, sc.file_version)) return fail(ctx, TVK_ERR_SOURCE, "%s: %s", path, f.error.c_str());
    double lo, hi;
    if (uvf_range(sc, f.lod0_brick_count(), &lo, &hi)) range_max = hi;
  }
  return tvk_open_octree_file(ctx, path, sc.toc_payload_offset, sc.file_version, scale,
                              sc.have_maxmin ? sc.maxmin.data() : nullptr, sc.maxmin.size() / 4, range_max,
                              max_gradient_magnitude, info);
}
int tvk_open_uvf(tvk_ctx* ctx, const char* path, uint64_t timestep, const float scale[3], double range_max,
                 float max_gradient_magnitude, tvk_octree_file_info* info) {
  try { return tvk_open_uvf_impl(ctx, path, timestep, scale, range_max, max_gradient_magnitude, info); }
  catch (const std::bad_alloc&) { return fail(ctx, TVK_ERR_OOM, "tvk_open_uvf: out of host memory"); }
  catch (const std::exception& e) { return fail(ctx, TVK_ERR_SOURCE, "tvk_open_uvf: %s", e.what()); }
}

static int tvk_uvf_probe_stats_impl(const char* path, uint64_t timestep, double range[2], uint64_t* hist1d_size, uint64_t* hist1d_filled,
                        float* max_gradient_magnitude, uint64_t hist2d_size[2]) {
  if (!path) return TVK_ERR_INVALID;
  UvfScan sc;
  if (!uvf_scan(path, timestep, &sc)) { g_create_err = std::string(path) + ": " + sc.error; return TVK_ERR_SOURCE; }
  if (range) {
    range[0] = 1.0; range[1] = -1.0;                  // "not known": second < first, the reference's convention
    if (sc.have_maxmin) {
      OctreeFile f;
      if (!f.open(path, sc.toc_payload_offset, sc.file_version)) { g_create_err = std::string(path) + ": " + f.error; return TVK_ERR_SOURCE; }
      uvf_range(sc, f.lod0_brick_count(), &range[0], &range[1]);
    }
  }
  if (hist1d_size) *hist1d_size = sc.have_hist1d ? sc.hist1d_size : 0;
  if (hist1d_filled) *hist1d_filled = sc.have_hist1d ? sc.hist1d_filled : 0;
  if (max_gradient_magnitude) *max_gradient_magnitude = sc.have_hist2d ? sc.max_grad_magnitude : 0.0f;
  if (hist2d_size) { hist2d_size[0] = sc.hist2d_size[0]; hist2d_size[1] = sc.hist2d_size[1]; }
  return TVK_OK;
}
int tvk_uvf_probe_stats(const char* path, uint64_t timestep, double range[2], uint64_t* hist1d_size, uint64_t* hist1d_filled,
                        float* max_gradient_magnitude, uint64_t hist2d_size[2]) {
  try { return tvk_uvf_probe_stats_impl(path, timestep, range, hist1d_size, hist1d_filled, max_gradient_magnitude, hist2d_size); }
  catch (const std::bad_alloc&) { g_create_err = "tvk_uvf_probe_stats: out of host memory"; return TVK_ERR_OOM; }
  catch (const std::exception& e) { g_create_err = std::string("tvk_uvf_probe_stats: ") + e.what(); return TVK_ERR_SOURCE; }
}

static int tvk_uvf_probe_impl(const char* path, uint64_t timestep, uint64_t* toc_payload_offset, uint64_t* file_version,
                  uint64_t* n_blocks, uint64_t* n_timesteps, double* maxmin, uint64_t maxmin_cap, uint64_t* n_maxmin) {
  if (!path) return TVK_ERR_INVALID;
  UvfScan sc;
  if (!uvf_scan(path, timestep, &sc)) { g_create_err = std::string(path) + ": " + sc.error; return TVK_ERR_SOURCE; }
  if (toc_payload_offset) *toc_payload_offset = sc.toc_payload_offset;
  if (file_version) *file_version = sc.file_version;
  if (n_blocks) *n_blocks = sc.n_blocks;
  if (n_timesteps) *n_timesteps = sc.n_toc;
  if (n_maxmin) *n_maxmin = sc.maxmin.size() / 4;
  if (maxmin && maxmin_cap) std::memcpy(maxmin, sc.maxmin.data(), std::min<size_t>(maxmin_cap, sc.maxmin.size() / 4) * 32);
  return TVK_OK;
}
int tvk_uvf_probe(const char* path, uint64_t timestep, uint64_t* toc_payload_offset, uint64_t* file_version,
                  uint64_t* n_blocks, uint64_t* n_timesteps, double* maxmin, uint64_t maxmin_cap, uint64_t* n_maxmin) {
  try { return tvk_uvf_probe_impl(path, timestep, toc_payload_offset, file_version, n_blocks, n_timesteps, maxmin, maxmin_cap, n_maxmin); }
  catch (const std::bad_alloc&) { g_create_err = "tvk_uvf_probe: out of host memory"; return TVK_ERR_OOM; }
  catch (const std::exception& e) { g_create_err = std::string("tvk_uvf_probe: ") + e.what(); return TVK_ERR_SOURCE; }
}

static void fill_file_info(const OctreeFile& g, tvk_octree_file_info* info) {
  std::memset(info, 0, sizeof(*info));
  for (int i = 0; i < 3; i++) {
    info->domain_size[i] = (uint32_t)g.vol[i]; info->aspect[i] = g.aspect[i]; info->max_brick_size[i] = (uint32_t)g.brick[i];
  }
  info->overlap = g.overlap; info->version = g.version; info->lod_count = g.lod_count();
  info->dtype = (g.component_count == 4 && g.component_type == 0 && !g.precomputed_normals) ? TVK_RGBA8
              : g.component_count != 1 ? -1 : g.component_type == 0 ? TVK_U8 : g.component_type == 1 ? TVK_U16
              : g.component_type == 8 ? TVK_F32 : -1;
  info->brick_count = g.toc.size();
  for (const OctreeToc& t : g.toc) {
    info->payload_bytes += t.length;
    info->bricks_by_codec[t.codec < 6 ? t.codec : 5]++;
  }
}

// host-only helpers (no device, no ctx): header/TOC probe and single-brick read of an ExtendedOctree file
static int tvk_octree_file_probe_impl(const char* path, uint64_t offset, uint64_t uvf_file_version, tvk_octree_file_info* info) {
  if (!path || !info) return TVK_ERR_INVALID;
  OctreeFile f;
  if (!f.open(path, offset, uvf_file_version)) { g_create_err = std::string(path) + ": " + f.error; return TVK_ERR_SOURCE; }
  fill_file_info(f, info);
  return TVK_OK;
}
int tvk_octree_file_probe(const char* path, uint64_t offset, uint64_t uvf_file_version, tvk_octree_file_info* info) {
  try { return tvk_octree_file_probe_impl(path, offset, uvf_file_version, info); }
  catch (const std::bad_alloc&) { g_create_err = "tvk_octree_file_probe: out of host memory"; return TVK_ERR_OOM; }
  catch (const std::exception& e) { g_create_err = std::string("tvk_octree_file_probe: ") + e.what(); return TVK_ERR_SOURCE; }
}

static int tvk_octree_file_read_brick_impl(const char* path, uint64_t offset, uint64_t uvf_file_version, uint32_t x, uint32_t y,
                               uint32_t z, uint32_t lod, void* dst, size_t cap, uint32_t out_size[3]) {
  if (!path || !dst) return TVK_ERR_INVALID;
  OctreeFile f;
  if (!f.open(path, offset, uvf_file_version)) { g_create_err = std::string(path) + ": " + f.error; return TVK_ERR_SOURCE; }
  if (lod >= f.lod_count() || x >= f.lod_layout[3 * lod] || y >= f.lod_layout[3 * lod + 1] || z >= f.lod_layout[3 * lod + 2]) {
    g_create_err = "brick coordinates out of range";
    return TVK_ERR_INVALID;
  }
  uint32_t bs[3];
  f.brick_size(x, y, z, lod, bs);
  if (out_size) { out_size[0] = bs[0]; out_size[1] = bs[1]; out_size[2] = bs[2]; }
  std::string err;
  if (!f.read_brick(f.brick_index(x, y, z, lod), (size_t)bs[0] * bs[1] * bs[2] * f.element_bytes(), dst, cap, &err)) {
    g_create_err = std::string(path) + ": " + err;
    return TVK_ERR_SOURCE;
  }
  return TVK_OK;
}
int tvk_octree_file_read_brick(const char* path, uint64_t offset, uint64_t uvf_file_version, uint32_t x, uint32_t y,
                               uint32_t z, uint32_t lod, void* dst, size_t cap, uint32_t out_size[3]) {
  try { return tvk_octree_file_read_brick_impl(path, offset, uvf_file_version, x, y, z, lod, dst, cap, out_size); }
  catch (const std::bad_alloc&) { g_create_err = "tvk_octree_file_read_brick: out of host memory"; return TVK_ERR_OOM; }
  catch (const std::exception& e) { g_create_err = std::string("tvk_octree_file_read_brick: ") + e.what(); return TVK_ERR_SOURCE; }
}

int tvk_build_volume(tvk_ctx* ctx, const void* raw, int raw_on_device, const uint32_t size[3], int dtype,
                     const float scale[3], const uint32_t max_brick_size[3], uint32_t overlap, int clamp_to_edge,
                     double range_max, float max_gradient_magnitude) {
  if (!ctx || !raw || !size || !max_brick_size) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  if (dtype == TVK_RGBA8) {   // colour: mean pyramid, unsharded store (sort-last refuses colour data anyway)
    if (ctx->pyramid_median) return fail(ctx, TVK_ERR_INVALID, "tvk_build_volume: the median pyramid is built for scalar volumes");
    for (int i = 0; i < 3; i++)
      if (ctx->store_clip_min[i] > 0.0f || ctx->store_clip_max[i] < 1.0f)
        return fail(ctx, TVK_ERR_INVALID, "tvk_build_volume: a sharded brick store is built for scalar volumes");
  }
  int rc = set_geometry(ctx, size, scale, max_brick_size, overlap, dtype, range_max, max_gradient_magnitude);
  if (rc) return rc;
  ctx->cb = nullptr; ctx->cb_user = nullptr;
  const uint64_t n0 = (uint64_t)size[0] * size[1] * size[2];
  void* lod_prev = nullptr;
  void* lod_own0 = nullptr;
  if (raw_on_device) lod_prev = const_cast<void*>(raw);
  else {
    CU(cudaMalloc(&lod_own0, n0 * ctx->esize));
    CU(cudaMemcpyAsync(lod_own0, raw, n0 * ctx->esize, cudaMemcpyHostToDevice, ctx->stream));
    lod_prev = lod_own0;
  }
  // sort-last at the source (tvk_set_store_shard): keep only the bricks that touch this rank's box
  ctx->store_count = ctx->n_bricks_all;
  bool sharded = false;
  for (int i = 0; i < 3; i++) sharded = sharded || ctx->store_clip_min[i] > 0.0f || ctx->store_clip_max[i] < 1.0f;
  if (sharded) {
    ctx->store_index.assign(ctx->n_bricks_all, -1);
    uint64_t kept = 0;
    for (uint32_t l = 0; l < ctx->pool_lod_count; l++) {
      float lay[3];
      for (int i = 0; i < 3; i++) {   // vLODLayout[l], as derive() hands it to the kernel
        float c = (float)ctx->vol[i] / ctx->inner[i];
        c = c / (float)(1u << l);
        if ((float)(uint32_t)c == c) c = c - c * std::numeric_limits<float>::epsilon();
        lay[i] = c;
      }
      const uint32_t* n = ctx->layout[l];
      for (uint32_t z = 0; z < n[2]; z++)
        for (uint32_t y = 0; y < n[1]; y++)
          for (uint32_t x = 0; x < n[0]; x++) {
            const uint32_t co[3] = {x, y, z};
            bool outside = false;
            for (int i = 0; i < 3; i++) {   // classify_brick of the traversal kernel: OUTSIDE_SHARD
              const float c0 = (float)co[i] / lay[i], c1 = (float)(co[i] + 1) / lay[i];
              const float lo = ctx->store_clip_min[i] > 0.0f ? ctx->store_clip_min[i] : -INFINITY;
              const float hi = ctx->store_clip_max[i] < 1.0f ? ctx->store_clip_max[i] : INFINITY;
              if (c1 <= lo || c0 >= hi) outside = true;
            }
            if (!outside) ctx->store_index[ctx->toc_offset[l] + x + (uint64_t)y * n[0] + (uint64_t)z * n[0] * n[1]] = (int32_t)kept++;
          }
    }
    ctx->store_count = kept;
  }
  cudaError_t e = cudaMalloc(&ctx->store_d, std::max<uint64_t>(ctx->store_count, 1) * ctx->slot_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->minmax_d, ctx->n_bricks_all * 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->store_d, 0, std::max<uint64_t>(ctx->store_count, 1) * ctx->slot_bytes, ctx->stream);
  if (e == cudaSuccess && sharded) {
    e = cudaMalloc(&ctx->store_index_d, ctx->store_index.size() * sizeof(int32_t));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ctx->store_index_d, ctx->store_index.data(), ctx->store_index.size() * sizeof(int32_t),
                          cudaMemcpyHostToDevice, ctx->stream);
  }
  void* lod_cur = nullptr;
  for (uint32_t l = 0; l < ctx->lod_count && e == cudaSuccess; l++) {
    if (l > 0) {
      const uint64_t n = (uint64_t)ctx->lod_size[l][0] * ctx->lod_size[l][1] * ctx->lod_size[l][2];
      e = cudaMalloc(&lod_cur, n * ctx->esize);
      if (e != cudaSuccess) break;
      launch_downsample(lod_prev, ctx->lod_size[l - 1], lod_cur, ctx->lod_size[l], dtype, ctx->pyramid_median ? 1 : 0, ctx->stream);
      if (lod_prev != raw) { cudaStreamSynchronize(ctx->stream); cudaFree(lod_prev); }
      lod_prev = lod_cur;
    }
    CutConsts cc{};
    for (int i = 0; i < 3; i++) { cc.lod_size[i] = ctx->lod_size[l][i]; cc.layout[i] = ctx->layout[l][i]; cc.brick[i] = ctx->brick[i]; }
    cc.overlap = overlap; cc.clamp = clamp_to_edge; cc.lod = (int32_t)l; cc.first_brick = ctx->toc_offset[l];
    static const bool trace = std::getenv("TVK_BUILD_TRACE") != nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (trace) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, ctx->stream); }
    launch_cut_bricks(lod_prev, ctx->store_d, ctx->store_index_d, ctx->minmax_d, cc, dtype, ctx->slot_bytes, ctx->stream);
    e = cudaGetLastError();
    if (trace) {
      cudaEventRecord(t1, ctx->stream); cudaEventSynchronize(t1);
      float ms = 0.0f; cudaEventElapsedTime(&ms, t0, t1);
      const double nb = (double)cc.layout[0] * cc.layout[1] * cc.layout[2];
      const double bytes = (double)cc.lod_size[0] * cc.lod_size[1] * cc.lod_size[2] * ctx->esize + nb * ctx->slot_bytes;
      fprintf(stderr, "[tvk build] level %u: %.0f bricks cut in %.3f ms = %.0f GB/s (volume read once + bricks written)\n", l, nb, ms,
              bytes / 1e9 / (ms * 1e-3));
      cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (lod_prev && lod_prev != raw) cudaFree(lod_prev);
  if (e != cudaSuccess) {
    free_dataset(ctx);
    return fail(ctx, e == cudaErrorMemoryAllocation ? TVK_ERR_OOM : TVK_ERR_CUDA, "tvk_build_volume: %s",
                cudaGetErrorString(e));
  }
  ctx->minmax_h.resize(4 * (size_t)ctx->n_bricks_all);
  CU(cudaMemcpy(ctx->minmax_h.data(), ctx->minmax_d, ctx->minmax_h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  ctx->have_volume = true; ctx->data_gen++;
  return TVK_OK;
}

int tvk_synth_volume(tvk_ctx* ctx, void* dst_device, int kind, const uint32_t size[3], int dtype, uint32_t seed) {
  if (!ctx || !dst_device || !size) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  if (kind < 0 || kind > 2 || dtype < TVK_U8 || dtype > TVK_F32) return fail(ctx, TVK_ERR_INVALID, "bad kind/dtype");
  cudaSetDevice(ctx->cfg.device);
  launch_synth(dst_device, kind, size, dtype, seed, ctx->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return TVK_OK;
}

int tvk_set_store_shard(tvk_ctx* ctx, const float clip_min[3], const float clip_max[3]) {
  if (!ctx || !clip_min || !clip_max) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  for (int i = 0; i < 3; i++)
    if (!(clip_min[i] >= 0.0f && clip_max[i] <= 1.0f && clip_min[i] < clip_max[i])) return fail(ctx, TVK_ERR_INVALID, "bad shard box");
  std::memcpy(ctx->store_clip_min, clip_min, 12);
  std::memcpy(ctx->store_clip_max, clip_max, 12);
  return TVK_OK;
}

int tvk_get_info(const tvk_ctx* ctx, tvk_info* o) {
  if (!ctx || !o) return TVK_ERR_INVALID;
  std::memset(o, 0, sizeof(*o));
  if (!ctx->have_volume) return TVK_ERR_INVALID;
  o->lod_count = ctx->lod_count; o->pool_lod_count = ctx->pool_lod_count; o->total_bricks = ctx->total_bricks;
  for (uint32_t l = 0; l < ctx->lod_count; l++)
    for (int i = 0; i < 3; i++) { o->lod_size[l][i] = ctx->lod_size[l][i]; o->brick_layout[l][i] = ctx->layout[l][i]; }
  for (uint32_t l = 0; l < ctx->pool_lod_count; l++) o->lod_offset[l] = ctx->lod_offset[l];
  for (int i = 0; i < 3; i++) { o->pool_size[i] = ctx->pool_size[i]; o->pool_capacity[i] = ctx->capacity[i]; o->meta_dim[i] = ctx->meta_dim[i]; }
  o->meta_count = ctx->meta_count;
  return TVK_OK;
}

int tvk_get_minmax(tvk_ctx* ctx, double* dst, uint64_t n_bricks) {
  if (!ctx || !dst || !ctx->have_volume) return fail(ctx, TVK_ERR_INVALID, "no dataset");
  if (4 * n_bricks > ctx->minmax_h.size()) return fail(ctx, TVK_ERR_INVALID, "only %zu bricks", ctx->minmax_h.size() / 4);
  std::memcpy(dst, ctx->minmax_h.data(), 4 * n_bricks * sizeof(double));
  return TVK_OK;
}

int tvk_get_brick_size(const tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]) {
  if (!ctx || !ctx->have_volume || lod >= ctx->lod_count) return TVK_ERR_INVALID;
  const uint32_t co[3] = {x, y, z};
  for (int i = 0; i < 3; i++) if (co[i] >= ctx->layout[lod][i]) return TVK_ERR_INVALID;
  brick_size(ctx, co, lod, out);
  return TVK_OK;
}

int tvk_read_brick(tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap) {
  uint32_t bs[3];
  if (!ctx || !dst) return TVK_ERR_INVALID;
  if (tvk_get_brick_size(ctx, x, y, z, lod, bs) != TVK_OK) return fail(ctx, TVK_ERR_INVALID, "bad brick coordinates");
  const size_t bytes = (size_t)bs[0] * bs[1] * bs[2] * ctx->esize;
  if (cap < bytes) return fail(ctx, TVK_ERR_INVALID, "buffer too small");
  if (ctx->proc.on) return proc_brick_cb(ctx, x, y, z, lod, dst, cap) == 0 ? TVK_OK : fail(ctx, TVK_ERR_SOURCE, "procedural brick failed");
  if (!ctx->store_d) return fail(ctx, TVK_ERR_INVALID, "no device brick store (dataset comes from a callback)");
  uint64_t idx = ctx->toc_offset[lod] + x + (uint64_t)y * ctx->layout[lod][0] +
                 (uint64_t)z * ctx->layout[lod][0] * ctx->layout[lod][1];
  if (!ctx->store_index.empty()) {
    if (ctx->store_index[idx] < 0) return fail(ctx, TVK_ERR_SOURCE, "brick is not in this rank's brick store");
    idx = (uint64_t)ctx->store_index[idx];
  }
  cudaMemcpy3DParms p{};
  p.srcPtr = make_cudaPitchedPtr((unsigned char*)ctx->store_d + idx * ctx->slot_bytes, ctx->brick[0] * ctx->esize,
                                 ctx->brick[0] * ctx->esize, ctx->brick[1]);
  p.dstPtr = make_cudaPitchedPtr(dst, bs[0] * ctx->esize, bs[0] * ctx->esize, bs[1]);
  p.extent = make_cudaExtent(bs[0] * ctx->esize, bs[1], bs[2]);
  p.kind = cudaMemcpyDeviceToHost;
  CU(cudaMemcpy3D(&p));
  return TVK_OK;
}

// ---- transfer functions -------------------------------------------------------------------------
int tvk_set_tf1d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t n, uint64_t nz_lo, uint64_t nz_hi) {
  if (!ctx || !rgba || n < 2) return fail(ctx, TVK_ERR_INVALID, "bad 1D transfer function");
  cudaSetDevice(ctx->cfg.device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->tf1d_n != n) {
    if (ctx->tf1d_d) cudaFree(ctx->tf1d_d);
    ctx->tf1d_d = nullptr; ctx->tf1d_n = 0;
    CU(cudaMalloc(&ctx->tf1d_d, (size_t)n * 4));
  }
  // the table stays RGBA8 on the device, like the reference's GL_RGBA8 texture (GPUMemMan.cpp:398-401)
  CU(cudaMemcpy(ctx->tf1d_d, rgba, (size_t)n * 4, cudaMemcpyHostToDevice));
  ctx->tf1d_n = n; ctx->tf1d_nz[0] = nz_lo; ctx->tf1d_nz[1] = nz_hi;
  ctx->tf_gen++;
  ctx->blank = true;
  return TVK_OK;
}

int tvk_set_tf2d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, const uint64_t nz[4]) {
  if (!ctx || !rgba || !nz || w == 0 || h == 0) return fail(ctx, TVK_ERR_INVALID, "bad 2D transfer function");
  cudaSetDevice(ctx->cfg.device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->tf2d_w != w || ctx->tf2d_h != h) {
    if (ctx->tf2d_d) cudaFree(ctx->tf2d_d);
    ctx->tf2d_d = nullptr; ctx->tf2d_w = ctx->tf2d_h = 0;
    CU(cudaMalloc(&ctx->tf2d_d, (size_t)w * h * 4));
  }
  CU(cudaMemcpy(ctx->tf2d_d, rgba, (size_t)w * h * 4, cudaMemcpyHostToDevice));
  ctx->tf2d_w = w; ctx->tf2d_h = h;
  for (int i = 0; i < 4; i++) ctx->tf2d_nz[i] = nz[i];
  ctx->tf_gen++;
  ctx->blank = true;
  return TVK_OK;
}

// ---- pool ---------------------------------------------------------------------------------------
int tvk_create_pool(tvk_ctx* ctx, const uint32_t* pool_size) {
  if (!ctx || !ctx->have_volume) return fail(ctx, TVK_ERR_INVALID, "no dataset registered");
  cudaSetDevice(ctx->cfg.device);
  free_pool(ctx);
  if (pool_size) for (int i = 0; i < 3; i++) ctx->pool_size[i] = pool_size[i];
  else size_pool(ctx->cfg.max_gpu_mem, ctx->esize * 8, ctx->brick, ctx->n_bricks_all, ctx->cfg.max_pool_dim, ctx->pool_size);
  for (int i = 0; i < 3; i++) {
    ctx->capacity[i] = ctx->pool_size[i] / ctx->brick[i];
    if (ctx->capacity[i] == 0) return fail(ctx, TVK_ERR_INVALID, "pool smaller than one brick");
  }
  ctx->n_slots = ctx->capacity[0] * ctx->capacity[1] * ctx->capacity[2];
  if (ctx->n_slots < 2) return fail(ctx, TVK_ERR_INVALID, "pool needs at least two slots");
  ctx->slots.clear();
  ctx->slots.reserve(ctx->n_slots);
  for (uint32_t z = 0; z < ctx->capacity[2]; z++)
    for (uint32_t y = 0; y < ctx->capacity[1]; y++)
      for (uint32_t x = 0; x < ctx->capacity[0]; x++) ctx->slots.push_back(Slot{-1, 0, 0, {x, y, z}});
  ctx->time_of_creation = 2;
  ctx->insert_pos = 0;
  if (!fit_1d_to_3d(ctx->total_bricks, ctx->cfg.max_pool_dim, ctx->meta_dim))
    return fail(ctx, TVK_ERR_INVALID, "Unable to create brick metadata texture, as it needs more than the max texture size");
  ctx->meta_count = (uint64_t)ctx->meta_dim[0] * ctx->meta_dim[1] * ctx->meta_dim[2];
  ctx->meta_h.assign(ctx->meta_count, TVK_BI_MISSING);
  // 8- and 16-bit pools are stored in the x-pair layout (k_pool.cu: element x = (voxel x, voxel x+1)): 2 * slot_bytes per slot.  The
  // pool is still SIZED in voxels exactly as the reference sizes its atlas (same slot count and page table for the same
  // budget); the second copy of each voxel is the price of halving the traversal kernel's load count.
  // one extra slot of padding so a vector load at the very end never leaves the allocation
  CU(cudaMalloc(&ctx->pool_d, ((uint64_t)ctx->n_slots + 1) * ctx->pool_slot_bytes));
  CU(cudaMalloc(&ctx->meta_d, ctx->meta_count * 4));
  CU(cudaMalloc(&ctx->slot_brick_d, (size_t)ctx->n_slots * 4));
  CU(cudaMalloc(&ctx->counts_d, 4 * sizeof(uint32_t)));
  CU(cudaMalloc(&ctx->visited_d, (ctx->meta_count / 32 + 1) * 4));
  ctx->visited_h.assign(ctx->meta_count / 32 + 1, 0);
  // The clears run on the render stream; ctx->stream and copy_stream are non-blocking streams, so nothing orders them
  // against the legacy NULL stream -- the first upload below (either stream) must not overtake a multi-GB memset.
  CU(cudaMemsetAsync(ctx->meta_d, 0, ctx->meta_count * 4, ctx->stream));
  CU(cudaMemsetAsync(ctx->slot_brick_d, 0xFF, (size_t)ctx->n_slots * 4, ctx->stream));
  CU(cudaMemsetAsync(ctx->pool_d, 0, ((uint64_t)ctx->n_slots + 1) * ctx->pool_slot_bytes, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  // miss-report table (GLGridLeaper::InitHashTable, GLGridLeaper.cpp:266-291)
  ctx->hash_size = ctx->cfg.hash_table_size;
  CU(cudaMalloc(&ctx->hash_d, (size_t)ctx->hash_size * 4));
  CU(cudaMalloc(&ctx->miss_d, ((size_t)ctx->hash_size * 2 + 1) * 4));
  CU(cudaMallocHost(&ctx->miss_h, ((size_t)ctx->hash_size * 2 + 1) * 4));
  if (ctx->cb) {   // pinned + device staging for callback-sourced bricks: <= 64 MiB, >= 2 bricks
    size_t nb = (64ull << 20) / ctx->slot_bytes;
    nb = std::max<size_t>(2, nb & ~size_t(1));
    ctx->stage_bricks = nb;
    const size_t bytes = ((nb * ctx->slot_bytes + 15) & ~size_t(15)) + nb * sizeof(PageOp);
    CU(cudaMallocHost(&ctx->stage_h, bytes));
    CU(cudaMalloc(&ctx->stage_d, bytes));
  }
  ctx->have_pool = true;
  ctx->vis = VisState();
  // UploadFirstBrick: the single coarsest brick goes to the last slot and is never evicted
  {
    const uint32_t last_id = ctx->lod_offset[ctx->pool_lod_count - 1];
    std::vector<std::pair<uint32_t, uint32_t>> meta_kv, slot_kv;
    assign_slot(ctx, last_id, ctx->slots.size() - 1, std::numeric_limits<uint64_t>::max(), meta_kv, slot_kv);
    CopyReq r; r.id = last_id; r.slot = slot_kv[0].first; r.co[0] = r.co[1] = r.co[2] = 0; r.co[3] = ctx->pool_lod_count - 1;
    int rc = copy_bricks(ctx, std::vector<CopyReq>(1, r));
    if (rc == TVK_OK) rc = scatter_u32(ctx, ctx->meta_d, meta_kv);
    if (rc == TVK_OK) rc = scatter_u32(ctx, (uint32_t*)ctx->slot_brick_d, slot_kv);
    if (rc) { free_pool(ctx); return rc; }
  }
  ctx->blank = true;
  // RecomputeBrickVisibility() follows in the reference; it needs TF + mode, so it runs here only
  // when they are already known and otherwise with the first tvk_set_params / tvk_render
  if (ctx->have_params && ctx->tf1d_d) {
    uint32_t counts[4];
    if (!(ctx->params.mode == TVK_RM_2DTRANS && !ctx->tf2d_d)) return recompute_visibility(ctx, 1, counts);
  }
  return TVK_OK;
}

int tvk_recompute_visibility(tvk_ctx* ctx, int force, uint32_t counts[4]) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaSetDevice(ctx->cfg.device);
  return recompute_visibility(ctx, force, counts);
}

int tvk_upload_bricks(tvk_ctx* ctx, const uint32_t* ids, uint32_t n, uint32_t* out_slots, uint32_t* n_paged) {
  if (!ctx || !ctx->have_pool || (!ids && n)) return fail(ctx, TVK_ERR_INVALID, "no pool");
  cudaSetDevice(ctx->cfg.device);
  int rc = upload_bricks(ctx, ids, n, out_slots, n_paged);
  return rc;
}

int tvk_get_page_table(tvk_ctx* ctx, uint32_t* dst, uint64_t n) {
  if (!ctx || !dst || !ctx->have_pool) return fail(ctx, TVK_ERR_INVALID, "no pool");
  if (n > ctx->meta_count) return fail(ctx, TVK_ERR_INVALID, "table has %llu entries", (unsigned long long)ctx->meta_count);
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy(dst, ctx->meta_d, n * 4, cudaMemcpyDeviceToHost));   // the DEVICE table, not the host mirror
  return TVK_OK;
}

int tvk_get_slots(tvk_ctx* ctx, int32_t* brick_ids, uint64_t* times, uint32_t* pos3, uint32_t n_slots) {
  if (!ctx || !ctx->have_pool || n_slots > ctx->slots.size()) return fail(ctx, TVK_ERR_INVALID, "no pool / too many slots");
  for (uint32_t i = 0; i < n_slots; i++) {
    if (brick_ids) brick_ids[i] = ctx->slots[i].brick_id;
    if (times) times[i] = ctx->slots[i].time;
    if (pos3) std::memcpy(pos3 + 3 * i, ctx->slots[i].pos, 12);
  }
  return TVK_OK;
}

int tvk_read_pool_slot(tvk_ctx* ctx, uint32_t slot, void* dst, size_t cap) {
  if (!ctx || !dst || !ctx->have_pool || slot >= ctx->n_slots || cap < ctx->slot_bytes)
    return fail(ctx, TVK_ERR_INVALID, "bad slot / buffer");
  // the plain voxels of the slot: the first halves of its pairs, gathered on the device
  if (!ctx->unpair_d) CU(cudaMalloc(&ctx->unpair_d, ctx->slot_bytes));
  launch_slot_unpair((unsigned char*)ctx->pool_d + (uint64_t)slot * ctx->pool_slot_bytes, ctx->unpair_d, ctx->slot_voxels, ctx->esize,
                     ctx->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(dst, ctx->unpair_d, ctx->slot_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TVK_OK;
}

int tvk_get_touched_bricks(tvk_ctx* ctx, uint32_t* ids, uint64_t cap, uint64_t* n) {
  if (!ctx || !n || !ctx->have_pool) return fail(ctx, TVK_ERR_INVALID, "no pool");
  uint64_t k = 0;
  for (size_t w = 0; w < ctx->visited_h.size(); w++) {
    uint32_t bits = ctx->visited_h[w];
    while (bits) {
      const uint32_t b = (uint32_t)__builtin_ctz(bits);
      bits &= bits - 1;
      if (ids && k < cap) ids[k] = (uint32_t)(w * 32 + b);
      k++;
    }
  }
  *n = k;
  return TVK_OK;
}

int tvk_get_missing_list(tvk_ctx* ctx, uint32_t* ids, uint32_t cap, uint32_t* n) {
  if (!ctx || !n) return TVK_ERR_INVALID;
  const uint32_t have = (uint32_t)(ctx->last_missing.size() / 4);
  *n = have;
  if (ids) std::memcpy(ids, ctx->last_missing.data(), (size_t)std::min(have, cap) * 16);
  return TVK_OK;
}

// ---- rendering ----------------------------------------------------------------------------------
int tvk_default_params(tvk_render_params* p, uint32_t width, uint32_t height) {
  if (!p) return TVK_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->mode = TVK_RM_1DTRANS;
  p->lighting = 1;                       // m_bUseLighting(true)
  p->sample_rate_modifier = 1.0f;
  const float amb[4] = {1, 1, 1, 0.1f}, dif[4] = {1, 1, 1, 1.0f}, spe[4] = {1, 1, 1, 1.0f};
  std::memcpy(p->ambient, amb, 16); std::memcpy(p->diffuse, dif, 16); std::memcpy(p->specular, spe, 16);
  p->light_dir[0] = 0; p->light_dir[1] = 0; p->light_dir[2] = -1;
  p->iso_color[0] = p->iso_color[1] = p->iso_color[2] = 0.5f;
  for (int i = 0; i < 3; i++) { p->clip_min[i] = 0.0f; p->clip_max[i] = 1.0f; }
  const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const float eye[3] = {0, 0, 1.6f}, at[3] = {0, 0, 0}, up[3] = {0, 1, 0};
  return tvk_compute_view(p, width, height, ident, ident, eye, at, up, 50.0f, 0.01f, 1000.0f, 1.0f);
}

int tvk_compute_view(tvk_render_params* p, uint32_t width, uint32_t height, const float rotation[16],
                     const float translation[16], const float eye[3], const float at[3], const float up[3],
                     float fov_deg, float z_near, float z_far, float screen_space_error) {
  if (!p || !rotation || !translation || !eye || !at || !up || width == 0 || height == 0) return TVK_ERR_INVALID;
  // FLOATMATRIX4::BuildLookAt (Basics/Vectors.h:1250-1265), float arithmetic like the reference
  float F[3] = {at[0] - eye[0], at[1] - eye[1], at[2] - eye[2]};
  float U[3] = {up[0], up[1], up[2]};
  float S[3] = {F[1] * U[2] - F[2] * U[1], F[2] * U[0] - F[0] * U[2], F[0] * U[1] - F[1] * U[0]};
  U[0] = S[1] * F[2] - S[2] * F[1]; U[1] = S[2] * F[0] - S[0] * F[2]; U[2] = S[0] * F[1] - S[1] * F[0];
  auto normalize = [](float* v) {
    const float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (l != 0.0f) { v[0] /= l; v[1] /= l; v[2] /= l; }
  };
  normalize(F); normalize(U); normalize(S);
  auto dot = [](const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  float view[16];
  view[0] = S[0]; view[4] = S[1]; view[8] = S[2]; view[12] = -dot(S, eye);
  view[1] = U[0]; view[5] = U[1]; view[9] = U[2]; view[13] = -dot(U, eye);
  view[2] = -F[0]; view[6] = -F[1]; view[10] = -F[2]; view[14] = dot(F, eye);
  view[3] = 0; view[7] = 0; view[11] = 0; view[15] = 1;
  // Perspective (Vectors.h:1267-1277)
  const float aspect = (float)width / (float)height;
  const float fovy = fov_deg * float(3.14159265358979323846 / 180.0);
  const float cotan = float(1.0 / tan(double(fovy) / 2.0));
  float* pr = p->projection;
  std::memset(pr, 0, 64);
  pr[0] = cotan / aspect; pr[5] = cotan;
  pr[10] = -(z_far + z_near) / (z_far - z_near); pr[14] = -2.0f * (z_far * z_near) / (z_far - z_near);
  pr[11] = -1.0f;
  // modelView = rotation * translation * view (GLRenderer.cpp:627), float products like FLOATMATRIX4::operator*
  auto mulf = [](const float* a, const float* b, float* o) {
    float t[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        float s = 0.0f;
        for (int k = 0; k < 4; k++) s += a[r * 4 + k] * b[k * 4 + c];
        t[r * 4 + c] = s;
      }
    std::memcpy(o, t, 64);
  };
  float rt[16];
  mulf(rotation, translation, rt);
  mulf(rt, view, p->model_view);
  // CullingLOD::SetScreenParams (CullingLOD.cpp:57-67), incl. its 3.1416 literal
  p->lod_factor = 2.0f * tanf(fov_deg * ((3.1416f / 180.0f) / 2.0f)) * screen_space_error / float(height);
  p->width = width; p->height = height;
  p->eye[0] = eye[0]; p->eye[1] = eye[1]; p->eye[2] = eye[2];
  return TVK_OK;
}

int tvk_set_params(tvk_ctx* ctx, const tvk_render_params* p) {
  if (!ctx || !p) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  if (p->width == 0 || p->height == 0 || p->width > 32768 || p->height > 32768) return fail(ctx, TVK_ERR_INVALID, "bad image size");
  if (p->mode < TVK_RM_1DTRANS || p->mode > TVK_RM_ISOSURFACE) return fail(ctx, TVK_ERR_INVALID, "Unhandled rendering mode.");
  if (!(p->sample_rate_modifier > 0.0f)) return fail(ctx, TVK_ERR_INVALID, "sample rate modifier must be > 0");
  ctx->params = *p;
  ctx->have_params = true;
  ctx->blank = true;
  return TVK_OK;
}

int tvk_set_pyramid_filter(tvk_ctx* ctx, int median) {
  if (!ctx) return TVK_ERR_INVALID;
  ctx->pyramid_median = median != 0;
  return TVK_OK;
}

int tvk_set_clip_plane(tvk_ctx* ctx, int enabled, const float plane_model[4]) {
  if (!ctx) return TVK_ERR_INVALID;
  if (enabled) {
    if (!plane_model) return fail(ctx, TVK_ERR_INVALID, "NULL clip plane");
    const float len = sqrtf(plane_model[0] * plane_model[0] + plane_model[1] * plane_model[1] + plane_model[2] * plane_model[2]);
    if (!(len > 0.0f) || !std::isfinite(len) || !std::isfinite(plane_model[3])) return fail(ctx, TVK_ERR_INVALID, "degenerate clip plane");
    std::memcpy(ctx->clip_plane, plane_model, 16);
  }
  ctx->clip_plane_on = enabled != 0;
  ctx->blank = true;
  return TVK_OK;
}

// PLANE<T>::transform(m): this = this * transpose(inverse(m)), normalised by |xyz| (Basics/Vectors.h:1459-1478) with
// m = inverse(rotation * translation) (GLGridLeaper.cpp:518-520), so transpose(inverse(m)) = transpose(rotation * translation);
// FillBBoxVBO then normalises xyz once more and keeps d
int tvk_clip_plane_to_model(const float plane_world[4], const float rotation[16], const float translation[16], float plane_model[4]) {
  if (!plane_world || !rotation || !translation || !plane_model) return fail(nullptr, TVK_ERR_INVALID, "NULL argument");
  float rt[16], inv[16], back[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float a = 0.0f;
      for (int k = 0; k < 4; k++) a += rotation[r * 4 + k] * translation[k * 4 + c];
      rt[r * 4 + c] = a;
    }
  // the reference inverts twice in float (inverse() of the product, then transform() inverts its argument again)
  double d1[16], d2[16], d3[16];
  for (int i = 0; i < 16; i++) d1[i] = rt[i];
  if (!inv4(d1, d2)) return fail(nullptr, TVK_ERR_INVALID, "singular rotation * translation");
  for (int i = 0; i < 16; i++) { inv[i] = (float)d2[i]; d2[i] = inv[i]; }
  if (!inv4(d2, d3)) return fail(nullptr, TVK_ERR_INVALID, "singular rotation * translation");
  for (int i = 0; i < 16; i++) back[i] = (float)d3[i];
  float v[4];
  for (int c = 0; c < 4; c++) {   // row vector times transpose(back): v[c] = sum_k plane[k] * back[c][k]
    float a = 0.0f;
    for (int k = 0; k < 4; k++) a += plane_world[k] * back[c * 4 + k];
    v[c] = a;
  }
  float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (!(len > 0.0f)) return fail(nullptr, TVK_ERR_INVALID, "degenerate clip plane");
  for (int c = 0; c < 4; c++) v[c] = v[c] / len;
  len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int c = 0; c < 3; c++) plane_model[c] = v[c] / len;
  plane_model[3] = v[3];
  return TVK_OK;
}

int tvk_pick(tvk_ctx* ctx, uint32_t mouse_x, uint32_t mouse_y, float out[3]) {
  if (!ctx || !out) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  if (!ctx->have_params || ctx->params.mode != TVK_RM_ISOSURFACE)
    return fail(ctx, TVK_ERR_INVALID, "Can only determine pick locations in isosurface rendering mode.");
  if (!ctx->img_w || !ctx->buf[0]) return fail(ctx, TVK_ERR_INVALID, "no frame");
  // GL window coordinates: row m_vWinSize.y - mousePos.y counted from the bottom; rows outside the target read nothing
  if (mouse_x >= ctx->img_w || mouse_y == 0 || mouse_y > ctx->img_h) return fail(ctx, TVK_ERR_INVALID, "No intersection.");
  cudaSetDevice(ctx->cfg.device);
  const size_t i = (size_t)(ctx->img_h - mouse_y) * ctx->img_w + mouse_x;
  float v[4];
  CU(cudaMemcpyAsync(v, ctx->buf[0] + i, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (v[3] == 0.0f) return fail(ctx, TVK_ERR_INVALID, "No intersection.");
  out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
  return TVK_OK;
}

int tvk_probe_fetch(tvk_ctx* ctx, uint32_t steps, const float dir[3], float* ms, uint64_t* samples) {
  if (!ctx || !dir || !ms || steps == 0) return fail(ctx, TVK_ERR_INVALID, "bad argument");
  cudaSetDevice(ctx->cfg.device);
  int rc = check_renderable(ctx);
  if (rc) return rc;
  if (dir[0] == 0.0f && dir[1] == 0.0f && dir[2] == 0.0f) return fail(ctx, TVK_ERR_INVALID, "zero direction");
  for (int i = 0; i < 3; i++)
    if (ctx->overlap < 2 || ctx->brick[i] < 12) return fail(ctx, TVK_ERR_INVALID, "the fetch probe needs the fast-path geometry (ghost >= 2)");
  rc = ensure_frame(ctx, ctx->params.width, ctx->params.height);
  if (rc) return rc;
  RayConsts u;
  rc = derive(ctx, u);
  if (rc) return rc;
  const bool grad = ctx->params.mode == TVK_RM_2DTRANS || (ctx->params.mode == TVK_RM_1DTRANS && ctx->params.lighting);
  cudaStream_t s = ctx->stream;
  launch_fetch_probe(u, ctx->dtype, grad, ctx->n_slots, 8, dir, (float*)ctx->buf[7], s);   // warm-up (instruction cache)
  CU(cudaEventRecord(ctx->ev[0], s));
  launch_fetch_probe(u, ctx->dtype, grad, ctx->n_slots, steps, dir, (float*)ctx->buf[7], s);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->ev[1], s));
  CU(cudaEventSynchronize(ctx->ev[1]));
  cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]);
  if (samples) *samples = (uint64_t)ctx->params.width * ctx->params.height * steps;
  return TVK_OK;
}

int tvk_raycast_only(tvk_ctx* ctx) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaSetDevice(ctx->cfg.device);
  int rc = check_renderable(ctx);
  if (rc) return rc;
  ctx->blank = true;
  return raycast_pass(ctx, false);
}

// One subframe in two halves, so that a caller (the sort-last frame) can queue more work behind the traversal before
// the host waits for the miss table: render_enqueue puts the raycast pass, the miss-table compaction and the
// read-backs of the bookkeeping on the stream; render_finish waits, decodes the requests, pages the bricks in.
static int render_enqueue(tvk_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  CU(cudaEventRecord(ctx->ev[0], s));
  CU(cudaMemsetAsync(ctx->hash_d, 0, (size_t)ctx->hash_size * 4, s));           // GLHashTable::ClearData
  CU(cudaMemsetAsync(ctx->miss_d + 2 * (size_t)ctx->hash_size, 0, 4, s));
  if (ctx->counters_on) {
    CU(cudaMemsetAsync(ctx->counters_d, 0, 8 * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(ctx->visited_d, 0, ctx->visited_h.size() * 4, s));
  }
  int rc = raycast_pass(ctx, true);
  if (rc) return rc;
  rc = sl_second_pass(ctx);
  if (rc) return rc;
  CU(cudaEventRecord(ctx->ev[1], s));
  // GLHashTable::GetData: compact on the device, read back count + (index,value) pairs
  launch_hash_compact(ctx->hash_d, ctx->hash_size, ctx->miss_d, ctx->miss_d + 2 * (size_t)ctx->hash_size, s);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->miss_h + 2 * (size_t)ctx->hash_size, ctx->miss_d + 2 * (size_t)ctx->hash_size, 4,
                     cudaMemcpyDeviceToHost, s));
  if (ctx->counters_on) {
    CU(cudaMemcpyAsync(ctx->counters_h, ctx->counters_d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ctx->visited_h.data(), ctx->visited_d, ctx->visited_h.size() * 4, cudaMemcpyDeviceToHost, s));
  }
  return TVK_OK;
}

static int render_finish(tvk_ctx* ctx, tvk_frame_stats* st) {
  cudaStream_t s = ctx->stream;
  CU(cudaStreamSynchronize(s));
  const uint32_t n_miss = ctx->miss_h[2 * (size_t)ctx->hash_size];
  ctx->last_missing.clear();
  if (n_miss) {
    CU(cudaMemcpyAsync(ctx->miss_h, ctx->miss_d, (size_t)n_miss * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    // table order (ascending slot index) is the reference's request order
    std::vector<std::pair<uint32_t, uint32_t>> kv(n_miss);
    for (uint32_t i = 0; i < n_miss; i++) kv[i] = std::make_pair(ctx->miss_h[2 * i], ctx->miss_h[2 * i + 1]);
    std::sort(kv.begin(), kv.end());
    const uint32_t* f = ctx->pool_layout[0];
    ctx->last_missing.resize((size_t)n_miss * 4);
    for (uint32_t i = 0; i < n_miss; i++) {   // GLHashTable::Int2Vector (GLHashTable.cpp:65-77)
      uint32_t idx = kv[i].second - 1;
      const uint32_t v = f[0] * f[1] * f[2];
      const uint32_t w = idx / v; idx -= w * v;
      const uint32_t z = idx / (f[0] * f[1]); idx -= z * (f[0] * f[1]);
      const uint32_t y = idx / f[0]; idx -= y * f[0];
      uint32_t* o = &ctx->last_missing[4 * (size_t)i];
      o[0] = idx; o[1] = y; o[2] = z; o[3] = w;
    }
  }
  CU(cudaEventRecord(ctx->ev[2], s));
  uint32_t paged = 0;
  if (n_miss) {
    int rc = upload_bricks(ctx, ctx->last_missing.data(), n_miss, nullptr, &paged);
    if (rc) return rc;
  }
  CU(cudaEventRecord(ctx->ev[3], s));
  CU(cudaEventSynchronize(ctx->ev[3]));
  if (st) {
    st->converged = n_miss == 0;
    st->missing_reported = n_miss;
    st->bricks_paged = paged;
    if (ctx->counters_on) {
      st->samples = ctx->counters_h[0]; st->rays = ctx->counters_h[1]; st->brick_visits = ctx->counters_h[2];
      st->alive_lane_iters = ctx->counters_h[3]; st->warp_iters = ctx->counters_h[4];
      st->max_lane_iters = ctx->counters_h[5];
      uint64_t t = 0;
      for (uint32_t w : ctx->visited_h) t += (uint64_t)__builtin_popcount(w);
      st->bricks_touched = t;
    }
    cudaEventElapsedTime(&st->ms_raycast, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&st->ms_read_htable, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&st->ms_upload_bricks, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&st->ms_total, ctx->ev[0], ctx->ev[3]);
  }
  return TVK_OK;
}

int tvk_render(tvk_ctx* ctx, tvk_frame_stats* st) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaSetDevice(ctx->cfg.device);
  if (st) std::memset(st, 0, sizeof(*st));
  int rc = check_renderable(ctx);
  if (rc) return rc;
  ctx->result_buf = nullptr;
  // visibility follows TF / mode / isovalue changes (Changed1DTrans etc. -> RecomputeBrickVisibility)
  uint32_t counts[4];
  rc = recompute_visibility(ctx, 0, counts);
  if (rc) return rc;
  rc = render_enqueue(ctx);
  if (rc) return rc;
  return render_finish(ctx, st);
}

int tvk_render_stage(tvk_ctx* ctx, const void* in_resume_pos, const void* in_resume_color, tvk_frame_stats* st) {
  if (!ctx) return TVK_ERR_INVALID;
  if ((in_resume_pos == nullptr) != (in_resume_color == nullptr)) return fail(ctx, TVK_ERR_INVALID, "stage inputs come in pairs");
  ctx->stage_mode = true;
  ctx->stage_ray_start = static_cast<const float4*>(in_resume_pos);
  ctx->stage_color = static_cast<const float4*>(in_resume_color);
  ctx->blank = true;
  const int rc = tvk_render(ctx, st);      // visibility, miss table, the stage launch, paging of what it missed
  ctx->stage_mode = false;
  ctx->stage_ray_start = ctx->stage_color = nullptr;
  return rc;
}

int tvk_get_stage_outputs(tvk_ctx* ctx, void** image, void** resume_color, void** resume_pos) {
  if (!ctx || !ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  if (image) *image = ctx->buf[0];
  if (resume_color) *resume_color = ctx->buf[1];
  if (resume_pos) *resume_pos = ctx->buf[3];
  return TVK_OK;
}

int tvk_paint(tvk_ctx* ctx, uint32_t max_subframes, tvk_frame_stats* st) {
  if (!ctx) return TVK_ERR_INVALID;
  tvk_frame_stats acc{}, one{};
  if (max_subframes == 0) max_subframes = 1u << 20;
  for (uint32_t i = 0; i < max_subframes; i++) {
    int rc = tvk_render(ctx, &one);
    if (rc) return rc;
    acc.converged = one.converged;
    acc.missing_reported += one.missing_reported;
    acc.bricks_paged += one.bricks_paged;
    acc.samples += one.samples; acc.rays = one.rays; acc.brick_visits += one.brick_visits;
    acc.bricks_touched = one.bricks_touched;
    acc.ms_raycast += one.ms_raycast; acc.ms_read_htable += one.ms_read_htable;
    acc.ms_upload_bricks += one.ms_upload_bricks; acc.ms_total += one.ms_total;
    if (one.converged) break;
    if (one.bricks_paged == 0) break;   // nothing could be paged (pool exhausted this frame): avoid spinning
  }
  if (st) *st = acc;
  return TVK_OK;
}

int tvk_read_rgba8(tvk_ctx* ctx, uint8_t* dst, size_t pitch) {
  if (!ctx || !dst || !ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  cudaSetDevice(ctx->cfg.device);
  const size_t n = (size_t)ctx->img_w * ctx->img_h, row = (size_t)ctx->img_w * 4;
  if (pitch == 0) pitch = row;
  if (pitch < row) return fail(ctx, TVK_ERR_INVALID, "pitch too small");
  int rc = ensure_read(ctx, n * 4);
  if (rc) return rc;
  launch_quantize_rgba8(result_image(ctx), ctx->rgba8_d, n, ctx->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->read_h, ctx->rgba8_d, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (pitch == row) std::memcpy(dst, ctx->read_h, n * 4);
  else for (uint32_t y = 0; y < ctx->img_h; y++) std::memcpy(dst + y * pitch, (uint8_t*)ctx->read_h + y * row, row);
  return TVK_OK;
}

int tvk_read_rgba8_async(tvk_ctx* ctx, uint8_t* dst, size_t pitch) {
  if (!ctx || !dst || !ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  cudaSetDevice(ctx->cfg.device);
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, dst) != cudaSuccess || at.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return fail(ctx, TVK_ERR_INVALID, "tvk_read_rgba8_async needs page-locked memory (tvk_host_alloc)");
  }
  const size_t n = (size_t)ctx->img_w * ctx->img_h, row = (size_t)ctx->img_w * 4;
  if (pitch == 0) pitch = row;
  if (pitch < row) return fail(ctx, TVK_ERR_INVALID, "pitch too small");
  const int k = ctx->read_slot;
  if (!ctx->rgba8_async_d[k]) CU(cudaMalloc(&ctx->rgba8_async_d[k], n * 4));
  if (!ctx->read_ev[k]) CU(cudaEventCreateWithFlags(&ctx->read_ev[k], cudaEventDisableTiming));
  if (!ctx->quant_ev) CU(cudaEventCreateWithFlags(&ctx->quant_ev, cudaEventDisableTiming));
  // the copy that last used this staging image must be done before it is overwritten
  CU(cudaStreamWaitEvent(ctx->stream, ctx->read_ev[k], 0));
  launch_quantize_rgba8(result_image(ctx), ctx->rgba8_async_d[k], n, ctx->stream);
  CU(cudaGetLastError());
  CU(cudaEventRecord(ctx->quant_ev, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->quant_ev, 0));
  CU(cudaMemcpy2DAsync(dst, pitch, ctx->rgba8_async_d[k], row, row, ctx->img_h, cudaMemcpyDeviceToHost, ctx->copy_stream));
  CU(cudaEventRecord(ctx->read_ev[k], ctx->copy_stream));
  ctx->read_slot = k ^ 1;
  return TVK_OK;
}

int tvk_read_wait(tvk_ctx* ctx, int pending_allowed) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaSetDevice(ctx->cfg.device);
  if (pending_allowed >= 1) {   // everything but the newest read: the slot that will be reused next
    if (ctx->read_ev[ctx->read_slot]) CU(cudaEventSynchronize(ctx->read_ev[ctx->read_slot]));
  } else {
    CU(cudaStreamSynchronize(ctx->copy_stream));
  }
  return TVK_OK;
}

int tvk_host_alloc(tvk_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out || !bytes) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  cudaSetDevice(ctx->cfg.device);
  CU(cudaMallocHost(out, bytes));
  return TVK_OK;
}

int tvk_host_free(tvk_ctx* ctx, void* p) {
  if (!ctx) return TVK_ERR_INVALID;
  if (p) CU(cudaFreeHost(p));
  return TVK_OK;
}

int tvk_read_rgba32f(tvk_ctx* ctx, float* dst, size_t pitch) {
  if (!ctx || !dst || !ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  cudaSetDevice(ctx->cfg.device);
  const size_t row = (size_t)ctx->img_w * 16;
  if (pitch == 0) pitch = row;
  if (pitch < row) return fail(ctx, TVK_ERR_INVALID, "pitch too small");
  CU(cudaMemcpy2DAsync(dst, pitch, result_image(ctx), row, row, ctx->img_h, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TVK_OK;
}

int tvk_get_device_image(tvk_ctx* ctx, void** dptr) {
  if (!ctx || !dptr || !ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  *dptr = result_image(ctx);
  return TVK_OK;
}

int tvk_read_iso_buffers(tvk_ctx* ctx, float* hit_pos, float* hit_normal) {
  if (!ctx || !ctx->img_w || ctx->params.mode != TVK_RM_ISOSURFACE) return fail(ctx, TVK_ERR_INVALID, "no iso frame");
  const size_t bytes = (size_t)ctx->img_w * ctx->img_h * 16;
  CU(cudaStreamSynchronize(ctx->stream));
  if (hit_pos) CU(cudaMemcpy(hit_pos, ctx->buf[0], bytes, cudaMemcpyDeviceToHost));
  if (hit_normal) CU(cudaMemcpy(hit_normal, ctx->buf[5], bytes, cudaMemcpyDeviceToHost));
  return TVK_OK;
}

int tvk_composite_over(tvk_ctx* ctx, const void* front, const void* back, void* out, uint64_t n_pixels) {
  if (!ctx || !front || !back || !out) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  cudaSetDevice(ctx->cfg.device);
  launch_composite_over((const float4*)front, (const float4*)back, (float4*)out, n_pixels, ctx->stream);
  CU(cudaGetLastError());
  return TVK_OK;
}

int tvk_quantize_rgba8(tvk_ctx* ctx, const void* rgba32f, void* rgba8, uint64_t n_pixels) {
  if (!ctx || !rgba32f || !rgba8) return fail(ctx, TVK_ERR_INVALID, "NULL argument");
  cudaSetDevice(ctx->cfg.device);
  launch_quantize_rgba8((const float4*)rgba32f, (uchar4*)rgba8, n_pixels, ctx->stream);
  CU(cudaGetLastError());
  return TVK_OK;
}

}  // extern "C"

// =================================================================================================
// classic per-brick path (GLRaycaster): frame planning on the host, one traversal kernel
// =================================================================================================
namespace {

struct F3h { float x, y, z; };
inline float maxv(F3h a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }

// ExtendedOctree::ComputeMetadata's per-LoD aspect (anisotropic downsampling, ExtendedOctree.cpp:188-243)
void lod_aspect(const tvk_ctx* ctx, uint32_t lod, double out[3]) {
  double asp[3] = {1.0, 1.0, 1.0};
  for (uint32_t l = 1; l <= lod; l++) {
    for (int i = 0; i < 3; i++) {
      const uint64_t s = ctx->lod_size[l - 1][i], n = ctx->lod_size[l][i];
      if (s > 1) asp[i] *= (s % 2) ? float(s) / float(n) : 2;
    }
    const double mx = std::max(asp[0], std::max(asp[1], asp[2]));
    for (int i = 0; i < 3; i++) asp[i] /= mx;
  }
  for (int i = 0; i < 3; i++) out[i] = asp[i];
}

// vScale of BuildSubFrameBrickList / RegionNeedsBrick for a given domain size (AbstrRenderer.cpp:1003-1012,886-891)
F3h corrected_scale(const tvk_ctx* ctx, const uint32_t dom[3]) {
  const float dmax = (float)std::max(dom[0], std::max(dom[1], dom[2]));
  F3h c = {ctx->scale[0] * (float)dom[0] / dmax, ctx->scale[1] * (float)dom[1] / dmax, ctx->scale[2] * (float)dom[2] / dmax};
  const float m = maxv(c);
  return F3h{ctx->scale[0] / m, ctx->scale[1] / m, ctx->scale[2] / m};
}

struct AxisTab { std::vector<float> center_md, ext_md; std::vector<uint32_t> nvox; };

// UVFDataset::ComputeMetadataTOC (IO/uvfDataset.cpp:242-288) per axis: the x/y/z loops accumulate the brick
// corner independently per axis, so centre and extent of brick (x,y,z) are (cx[x], cy[y], cz[z]) etc.
void axis_tables(const tvk_ctx* ctx, uint32_t lod, AxisTab tab[3], float asp_f[3]) {
  double asp[3];
  lod_aspect(ctx, lod, asp);
  float nds[3];
  for (int i = 0; i < 3; i++) { asp_f[i] = (float)asp[i]; nds[i] = (float)ctx->lod_size[lod][i] * asp_f[i]; }
  const float max_val = std::fmax(nds[0], std::fmax(nds[1], nds[2]));
  for (int i = 0; i < 3; i++) nds[i] = nds[i] / max_val;
  for (int a = 0; a < 3; a++) {
    const uint32_t n = ctx->layout[lod][a];
    tab[a].center_md.resize(n); tab[a].ext_md.resize(n); tab[a].nvox.resize(n);
    float corner = 0.0f;
    for (uint32_t i = 0; i < n; i++) {
      uint32_t co[3] = {0, 0, 0};
      co[a] = i;
      uint32_t bs[3];
      brick_size(ctx, co, lod, bs);
      const float eff = (float)(bs[a] - 2 * ctx->overlap);
      const float ext = eff * asp_f[a] / max_val;
      tab[a].nvox[i] = bs[a];
      tab[a].ext_md[i] = ext;
      tab[a].center_md[i] = (corner + ext / 2.0f) - nds[a] * 0.5f;
      corner += ext;
    }
  }
}

int ensure_classic(tvk_ctx* ctx, size_t axis_words, size_t table_words) {
  if (ctx->classic_axis_cap < axis_words) {
    if (ctx->classic_axis_d) cudaFree(ctx->classic_axis_d);
    ctx->classic_axis_d = nullptr; ctx->classic_axis_cap = 0;
    CU(cudaMalloc(&ctx->classic_axis_d, axis_words * 4));
    ctx->classic_axis_cap = axis_words;
  }
  if (ctx->classic_table_cap < table_words) {
    if (ctx->classic_table_d) cudaFree(ctx->classic_table_d);
    ctx->classic_table_d = nullptr; ctx->classic_table_cap = 0;
    CU(cudaMalloc(&ctx->classic_table_d, table_words * 4));
    ctx->classic_table_cap = table_words;
  }
  return TVK_OK;
}

}  // namespace

extern "C" {

// One frame of the per-brick renderer.  mip = false: the 3D view (GLRenderer::Render3DView); mip = true: the HQ MIP
// frame of a 2D window (GLRenderer.cpp:1183-1253) -- the caller has put m_maMIPRotation * view into model_view
// (GLRaycaster::RenderHQMIPPreLoop, GLRaycaster.cpp:481-492).
static int render_per_brick(tvk_ctx* ctx, tvk_frame_stats* st, bool mip, int use_mip_lod) {
  if (!ctx) return TVK_ERR_INVALID;
  cudaSetDevice(ctx->cfg.device);
  if (st) std::memset(st, 0, sizeof(*st));
  int rc = check_renderable(ctx);
  if (rc) return rc;
  if (ctx->dtype == TVK_RGBA8)
    return fail(ctx, TVK_ERR_INVALID, "colour volumes are rendered on the GridLeaper path (GLRaycaster-Color-FS is not built)");
  const tvk_render_params& p = ctx->params;
  if (mip && !ctx->tf1d_d) return fail(ctx, TVK_ERR_INVALID, "MIP: no 1D transfer function set (Transfer-MIP-FS needs it)");
  rc = ensure_frame(ctx, p.width, p.height);
  if (rc) return rc;
  // ---- AbstrRenderer::ComputeMinLODForCurrentView (AbstrRenderer.cpp:789-803, CullingLOD.cpp:126-138) ----
  float ex[3];
  for (int i = 0; i < 3; i++) ex[i] = (float)ctx->vol[i] * ctx->scale[i];
  const float emax = std::fmax(ex[0], std::fmax(ex[1], ex[2]));
  for (int i = 0; i < 3; i++) ex[i] = ex[i] / emax;
  const float lzwse = std::fmax(ex[0] / (float)ctx->vol[0], std::fmax(ex[1] / (float)ctx->vol[1], ex[2] / (float)ctx->vol[2]));
  const float z_near = (float)((double)p.projection[14] / ((double)p.projection[10] - 1.0));
  const float fz = std::fmax(z_near, -p.model_view[14]);
  int lod_i = (int)floorf(logf(p.lod_factor * fz / lzwse) / logf(2.0f));
  lod_i = std::max(0, std::min(lod_i, (int)ctx->pool_lod_count - 1));   // the brick store holds the pool LoDs
  if (mip) {   // ---- AbstrRenderer::PlanHQMIPFrame (AbstrRenderer.cpp:1214-1245) ----
    uint32_t vc[3] = {ctx->vol[0], ctx->vol[1], ctx->vol[2]};
    uint64_t l = 0;
    if (use_mip_lod) {
      const uint32_t win = std::max(p.width, p.height);
      while (std::min(vc[0], std::min(vc[1], vc[2])) >= win) {
        for (int i = 0; i < 3; i++) vc[i] /= 2;
        l++;
      }
    }
    if (l > 0) l = std::min<uint64_t>(ctx->pool_lod_count - 1, l - 1);
    lod_i = (int)l;
  }
  const uint32_t lod = (uint32_t)lod_i;

  // ---- AbstrRenderer::BuildSubFrameBrickList (AbstrRenderer.cpp:999-1100) ----
  AxisTab tab[3];
  float asp[3];
  axis_tables(ctx, lod, tab, asp);
  const F3h s_list = corrected_scale(ctx, ctx->lod_size[0]), s_cull = corrected_scale(ctx, ctx->lod_size[lod]);
  const float sl[3] = {s_list.x, s_list.y, s_list.z}, sc[3] = {s_cull.x, s_cull.y, s_cull.z};
  float mvp[16], planes[6][4];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float s = 0.0f;
      for (int k = 0; k < 4; k++) s += p.model_view[r * 4 + k] * p.projection[k * 4 + c];
      mvp[r * 4 + c] = s;
    }
  {   // CullingLOD::Update (CullingLOD.cpp:89-124): right, left, top, bottom, far, near
    const int col[6] = {0, 0, 1, 1, 2, 2};
    const float sg[6] = {-1.0f, 1.0f, -1.0f, 1.0f, 1.0f, -1.0f};
    for (int i = 0; i < 6; i++)
      for (int r = 0; r < 4; r++) planes[i][r] = sg[i] * mvp[r * 4 + col[i]] + mvp[r * 4 + 3];
  }
  // AbstrRenderer::ContainsData (AbstrRenderer.cpp:953-997) with the dataset's legacy tests (uvfDataset.cpp:1201-1227)
  // the rescale factor always comes from the 1D table's size (GLGridLeaper.cpp:652-653 / AbstrRenderer.cpp:961), so a
  // frame without one cannot be planned (tf1d_n - 1 would wrap and every brick would fail ContainsData)
  if (!ctx->tf1d_d || ctx->tf1d_n < 2) return fail(ctx, TVK_ERR_INVALID, "no 1D transfer function (needed for the rescale factor)");
  const double rescale = ctx->range_max / double(ctx->tf1d_n - 1);
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
  if (p.mode == TVK_RM_1DTRANS) { v0 = double(ctx->tf1d_nz[0]) * rescale; v1 = double(ctx->tf1d_nz[1]) * rescale; }
  else { v0 = double(ctx->tf2d_nz[0]) * rescale; v1 = double(ctx->tf2d_nz[1]) * rescale; v2 = double(ctx->tf2d_nz[2]); v3 = double(ctx->tf2d_nz[3]); }
  const uint32_t* lay = ctx->layout[lod];
  std::vector<tvk_classic_brick>& list = ctx->classic_list;
  tvk_ctx::MipPlan key;
  key.valid = true; key.lod = lod; key.mode = p.mode; key.tf_gen = ctx->tf_gen; key.data_gen = ctx->data_gen;
  key.iso = p.isovalue; key.sample_rate = p.sample_rate_modifier;
  const tvk_ctx::MipPlan& have = ctx->mip_plan;
  const bool plan_hit = mip && have.valid && have.lod == key.lod && have.mode == key.mode && have.tf_gen == key.tf_gen &&
                        have.data_gen == key.data_gen && have.iso == key.iso && have.sample_rate == key.sample_rate;
  ctx->mip_plan.valid = false;
  if (!plan_hit) {
  list.clear();
  for (uint32_t z = 0; z < lay[2]; z++)
    for (uint32_t y = 0; y < lay[1]; y++)
      for (uint32_t x = 0; x < lay[0]; x++) {   // TOC index order == BrickTable key order
        const uint32_t ci[3] = {x, y, z};
        float cm[3], em[3];
        for (int a = 0; a < 3; a++) { cm[a] = tab[a].center_md[ci[a]]; em[a] = tab[a].ext_md[ci[a]]; }
        // CullingLOD::IsVisible (CullingLOD.cpp:141-160) on the box scaled like RegionNeedsBrick does
        bool visible = true;
        for (int i = 0; i < 6 && visible && !mip; i++) {   // HQ MIP: m_FrustumCullingLOD.SetPassAll(true)
          const float* pl = planes[i];
          const float cx = cm[0] * sc[0], cy = cm[1] * sc[1], cz = cm[2] * sc[2];
          const float hx = 0.5f * (em[0] * sc[0]), hy = 0.5f * (em[1] * sc[1]), hz = 0.5f * (em[2] * sc[2]);
          if (pl[0] * cx + pl[1] * cy + pl[2] * cz + pl[3] <= -(hx * std::fabs(pl[0]) + hy * std::fabs(pl[1]) + hz * std::fabs(pl[2])))
            visible = false;
        }
        if (!visible) continue;
        tvk_classic_brick b;
        b.index = z * lay[0] * lay[1] + y * lay[0] + x;
        b.x = x; b.y = y; b.z = z;
        const double* mm = &ctx->minmax_h[4 * (ctx->toc_offset[lod] + b.index)];
        bool has;
        if (p.mode == TVK_RM_1DTRANS) has = v1 >= mm[0] && v0 <= mm[1];
        else if (p.mode == TVK_RM_2DTRANS) has = (v1 >= mm[0] && v0 <= mm[1]) && (v3 >= mm[2] && v2 <= mm[3]);
        else has = p.isovalue <= mm[1];   // legacy one-sided iso test (uvfDataset.cpp:1201-1210)
        b.empty = has ? 0 : 1;
        b.distance = 0.0f;
        // HQ MIP: BuildSubFrameBrickList(true) orders by residency only (0 / 1) and BE_MAX blending is order-free
        if (has && !mip) {   // brick_distance (AbstrRenderer.cpp:808-841)
          float dmin = std::numeric_limits<float>::max();
          for (int k = 0; k < 8; k++) {
            float q[3];
            for (int a = 0; a < 3; a++) {
              const float sg = (k >> (2 - a)) & 1 ? 1.0f : -1.0f;
              q[a] = cm[a] * sl[a] + (sg * (em[a] * sl[a])) * 0.4999f;
            }
            const float* m = p.model_view;
            const float tx = q[0] * m[0] + q[1] * m[4] + q[2] * m[8] + 1.0f * m[12];
            const float ty = q[0] * m[1] + q[1] * m[5] + q[2] * m[9] + 1.0f * m[13];
            const float tz = q[0] * m[2] + q[1] * m[6] + q[2] * m[10] + 1.0f * m[14];
            dmin = std::fmin(dmin, sqrtf(fmaf(tz, tz, fmaf(ty, ty, tx * tx))));
          }
          b.distance = dmin;
        }
        list.push_back(b);
      }
  // depth sort; bricks at the same distance keep key order (the reference's std::sort leaves ties open)
  std::stable_sort(list.begin(), list.end(), [](const tvk_classic_brick& a, const tvk_classic_brick& b) { return a.distance < b.distance; });
  }   // !plan_hit
  ctx->classic_lod = lod;

  // ---- residency: every listed, non-empty brick must sit in the pool (GPUMemMan::GetVolume's job) ----
  cudaStream_t s = ctx->stream;
  CU(cudaEventRecord(ctx->ev[0], s));
  std::vector<uint32_t> want;
  size_t n_needed = 0;
  for (const tvk_classic_brick& b : list) {
    if (b.empty) continue;
    n_needed++;
    const uint32_t id = brick_id(ctx, b.x, b.y, b.z, lod);
    if (ctx->meta_h[id] >= TVK_BI_FLAG_COUNT) continue;
    want.push_back(b.x); want.push_back(b.y); want.push_back(b.z); want.push_back(lod);
  }
  if (n_needed > ctx->slots.size() - 1)
    return fail(ctx, TVK_ERR_OOM, "classic path: %zu bricks of LoD %u do not fit the pool (%zu slots)", n_needed, lod,
                ctx->slots.size());
  uint32_t paged = 0;
  if (!want.empty()) {
    // LRU touch: bricks of this frame that are already resident must not be the ones replaced
    for (tvk::Slot& sl_ : ctx->slots) {
      if (!sl_.was_ever_used() || sl_.time == UINT64_MAX) continue;
      // pool ids of one LoD are contiguous
      const uint32_t first = ctx->lod_offset[lod], count = lay[0] * lay[1] * lay[2];
      if ((uint32_t)sl_.brick_id >= first && (uint32_t)sl_.brick_id < first + count && sl_.contains_visible())
        sl_.time = ctx->time_of_creation++;
    }
    rc = upload_bricks(ctx, want.data(), (uint32_t)(want.size() / 4), nullptr, &paged);
    if (rc) return rc;
  }
  CU(cudaEventRecord(ctx->ev[1], s));

  // ---- per-brick uniforms as per-axis tables + the brick -> slot table ----
  uint32_t nmax = std::max(lay[0], std::max(lay[1], lay[2]));
  const uint32_t S = nmax + 1;
  std::vector<float> ax((size_t)6 * 3 * S, 0.0f);
  std::vector<uint32_t> nv((size_t)3 * S, 1u);
  float* plane = ax.data(); float* pmin = plane + 3 * S; float* pmax = pmin + 3 * S;
  float* tmax = pmax + 3 * S; float* tsc = tmax + 3 * S; float* rstep = tsc + 3 * S;
  for (int a = 0; a < 3; a++) {
    for (uint32_t i = 0; i < lay[a]; i++) {
      const float c = tab[a].center_md[i] * sl[a], e = tab[a].ext_md[i] * sl[a];   // Brick::vCenter / vExtension
      const float lo = c - e / 2.0f, hi = c + e / 2.0f;                           // RenderBox (GLRaycaster.cpp:309-311)
      pmin[a * S + i] = lo; pmax[a * S + i] = hi;
      plane[a * S + i] = lo;
      if (i + 1 == lay[a]) plane[a * S + i + 1] = hi;
      const float tmin_ = (float)ctx->overlap / (float)tab[a].nvox[i];            // UVFDataset::GetTextCoords
      const float tmax_ = (1.0f - tmin_) * asp[a];
      tmax[a * S + i] = tmax_;
      tsc[a * S + i] = (tmin_ - tmax_) / (lo - hi);                                // ComputeEyeToTextureMatrix
      rstep[a * S + i] = (e * (1.0f / (float)tab[a].nvox[i])) * (0.5f * 1.0f / p.sample_rate_modifier);
      nv[a * S + i] = tab[a].nvox[i];
    }
  }
  const size_t n_cells = (size_t)lay[0] * lay[1] * lay[2];
  std::vector<uint32_t>& table = ctx->classic_table_h;
  bool table_changed = !plan_hit || table.size() != n_cells;
  if (table_changed) table.assign(n_cells, 0u);
  for (const tvk_classic_brick& b : list) {
    if (b.empty) continue;
    const uint32_t m = ctx->meta_h[brick_id(ctx, b.x, b.y, b.z, lod)];
    if (m < TVK_BI_FLAG_COUNT) return fail(ctx, TVK_ERR_OOM, "classic path: brick (%u,%u,%u,%u) could not be paged in", b.x, b.y, b.z, lod);
    const uint32_t v = (m - TVK_BI_FLAG_COUNT) + 1u;
    if (table[b.index] != v) { table[b.index] = v; table_changed = true; }   // a cached plan whose bricks moved slots
  }
  const bool iso = !mip && p.mode == TVK_RM_ISOSURFACE;
  rc = ensure_classic(ctx, ax.size() + nv.size(), 2 * n_cells);
  if (rc) return rc;
  if (iso) {   // iTileID = position in m_vCurrentBrickList (GLRaycaster.cpp:285-288), empty bricks counted
    std::vector<uint32_t> pos(n_cells, 0u);
    for (size_t i = 0; i < list.size(); i++) pos[list[i].index] = (uint32_t)i;
    CU(cudaMemcpyAsync(ctx->classic_table_d + n_cells, pos.data(), n_cells * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));   // `pos` is a local
  }
  if (!plan_hit) {
    CU(cudaMemcpyAsync(ctx->classic_axis_d, ax.data(), ax.size() * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->classic_axis_d + ax.size(), nv.data(), nv.size() * 4, cudaMemcpyHostToDevice, s));
  }
  if (table_changed) CU(cudaMemcpyAsync(ctx->classic_table_d, table.data(), n_cells * 4, cudaMemcpyHostToDevice, s));

  ClassicConsts c;
  std::memset(&c, 0, sizeof(c));
  c.width = p.width; c.height = p.height;
  double mv[16], pr[16], imv[16], ipr[16];
  for (int i = 0; i < 16; i++) { mv[i] = p.model_view[i]; pr[i] = p.projection[i]; }
  if (!inv4(mv, imv) || !inv4(pr, ipr)) return fail(ctx, TVK_ERR_INVALID, "singular view or projection matrix");
  for (int i = 0; i < 16; i++) { c.imv[i] = (float)imv[i]; c.inv_proj[i] = (float)ipr[i]; }
  c.ortho = c.inv_proj[11] == 0.0f ? 1 : 0;   // FLOATMATRIX4::Ortho (Vectors.h:1279-1284): array[11] = 0
  if (c.ortho && !mip)
    return fail(ctx, TVK_ERR_INVALID, "a parallel projection is built for HQ MIP frames only (m_bOrthoView, GLRenderer.cpp:1183-1197)");
  const float mn = std::fmin(ctx->scale[0], std::fmin(ctx->scale[1], ctx->scale[2]));
  for (int i = 0; i < 3; i++) {
    c.domain_scale[i] = 1.0f / (ctx->scale[i] / mn);
    c.light_a[i] = p.ambient[i] * p.ambient[3];
    c.light_d[i] = p.diffuse[i] * p.diffuse[3];
    c.light_s[i] = p.specular[i] * p.specular[3];
    c.light_dir[i] = p.light_dir[i];
    c.layout[i] = lay[i];
    c.total[i] = ctx->brick[i];
  }
  c.norm = ctx->dtype == TVK_U8 ? 1.0f / 255.0f : ctx->dtype == TVK_U16 ? 1.0f / 65535.0f : 1.0f;
  const double full = ctx->dtype == TVK_U8 ? 255.0 : ctx->dtype == TVK_U16 ? 65535.0 : 1.0;
  c.trans_scale = (float)(full / ctx->range_max);
  c.gradient_scale = ctx->max_grad == 0.0f ? 1.0f : 1.0f / ctx->max_grad;
  // fStepScale = 1/sampleRate * max(domain(0)/domain(lod)) (GLRaycaster.cpp:257)
  c.step_scale = 1.0f / p.sample_rate_modifier *
                 std::fmax((float)ctx->vol[0] / (float)ctx->lod_size[lod][0],
                           std::fmax((float)ctx->vol[1] / (float)ctx->lod_size[lod][1], (float)ctx->vol[2] / (float)ctx->lod_size[lod][2]));
  if (p.mode == TVK_RM_2DTRANS && !mip) { c.tf = ctx->tf2d_d; c.tf_w = ctx->tf2d_w; c.tf_h = ctx->tf2d_h; }
  else { c.tf = ctx->tf1d_d; c.tf_w = ctx->tf1d_n; c.tf_h = 1; }
  c.nearest = p.nearest;
  c.count = ctx->counters_on ? 1 : 0;
  c.axis_stride = S;
  const float* base = ctx->classic_axis_d;
  c.plane = base; c.pmin = base + 3 * S; c.pmax = base + 6 * S; c.tmax = base + 9 * S; c.tsc = base + 12 * S;
  c.rstep = base + 15 * S;
  c.nvox = (const uint32_t*)(base + 18 * S);
  c.table = ctx->classic_table_d;
  c.pool = ctx->pool_d;
  c.slot_voxels = ctx->slot_voxels;
  c.out = ctx->buf[0];
  c.out_max = mip ? reinterpret_cast<float2*>(ctx->buf[1]) : nullptr;   // m_pFBO3DImageNext[1] of the MIP frame
  c.out_nrm = ctx->buf[5];                                             // m_pFBOIsoHit, second target
  c.list_pos = ctx->classic_table_d + n_cells;
  // fIsoval = GetNormalizedIsovalue (AbstrRenderer.cpp:412-424); vProjParam (GLRaycaster.cpp:213-217) with near / far
  // recovered from the projection matrix
  const bool clearview = iso && ctx->cv.on;   // ClearView belongs to the isosurface mode (AbstrRenderer::SetCV)
  if (clearview) {
    c.out_cv = ctx->buf[1]; c.out_cv_nrm = ctx->buf[2];   // m_pFBOCVHit (the GridLeaper resume buffers are free: blank)
    c.cv_isoval = ctx->dtype == TVK_U8 ? (float)(ctx->cv.iso / 256.0) : ctx->dtype == TVK_U16 ? (float)(ctx->cv.iso / 65536.0) : (float)ctx->cv.iso;
  }
  c.isoval = ctx->dtype == TVK_U8 ? (float)(p.isovalue / 256.0) : ctx->dtype == TVK_U16 ? (float)(p.isovalue / 65536.0) : (float)p.isovalue;
  {
    const double zn = pr[14] / (pr[10] - 1.0), zf = pr[14] / (pr[10] + 1.0);
    c.proj_param[0] = (float)(zf / (zf - zn));
    c.proj_param[1] = (float)(zf * zn / (zn - zf));
  }
  c.counters = ctx->counters_d;
  if (ctx->counters_on) CU(cudaMemsetAsync(ctx->counters_d, 0, 8 * sizeof(unsigned long long), s));
  launch_classic(c, mip ? TVK_CLASSIC_MIP : p.mode, p.lighting, ctx->dtype, s);
  CU(cudaGetLastError());
  if (iso) {   // GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2830)
    float a[3], d[3], sp[3];
    for (int i = 0; i < 3; i++) {
      a[i] = p.ambient[i] * p.ambient[3];
      d[i] = p.diffuse[i] * p.diffuse[3] * p.iso_color[i];
      sp[i] = p.specular[i] * p.specular[3];
    }
    if (clearview) {   // m_pProgramCVCompose (GLRenderer.cpp:2777-2795)
      float d2[3], prm[3] = {ctx->cv.size, ctx->cv.context, ctx->cv.border}, pick[3];
      for (int i = 0; i < 3; i++) {
        d2[i] = p.diffuse[i] * p.diffuse[3] * ctx->cv.color[i];
        const float* m = p.model_view;   // m_vCVPos * modelView (FLOATVECTOR4 * FLOATMATRIX4, Vectors.h:434-439)
        const float* q = ctx->cv.pos;
        pick[i] = q[0] * m[i] + q[1] * m[4 + i] + q[2] * m[8 + i] + q[3] * m[12 + i];
      }
      launch_cv_compose(ctx->buf[0], ctx->buf[5], ctx->buf[1], ctx->buf[2], ctx->buf[6], p.width, p.height, a, d, d2, sp,
                        p.light_dir, prm, pick, s);
    } else {
      launch_iso_compose(ctx->buf[0], ctx->buf[5], ctx->buf[6], p.width, p.height, a, d, sp, p.light_dir, s);
    }
    CU(cudaGetLastError());
  }
  ctx->cv_frame = clearview;
  CU(cudaEventRecord(ctx->ev[2], s));
  if (ctx->counters_on) CU(cudaMemcpyAsync(ctx->counters_h, ctx->counters_d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));   // the host tables above must outlive their copies
  ctx->blank = true;              // the GridLeaper resume buffers no longer describe this image
  if (mip) ctx->mip_plan = key;   // list + device tables describe this plan (a classic frame leaves it invalid)
  ctx->result_buf = mip ? ctx->buf[0] : nullptr;   // (an isosurface-mode MIP frame does not end in the deferred-shading buffer)
  if (st) {
    st->converged = 1;
    st->bricks_paged = paged;
    if (ctx->counters_on) st->samples = ctx->counters_h[0];
    st->bricks_touched = n_needed;   // listed, non-empty bricks (each is raycast once)
    cudaEventElapsedTime(&st->ms_upload_bricks, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&st->ms_raycast, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&st->ms_total, ctx->ev[0], ctx->ev[2]);
  }
  return TVK_OK;
}

int tvk_render_classic(tvk_ctx* ctx, tvk_frame_stats* st) { return render_per_brick(ctx, st, false, 0); }

int tvk_render_mip(tvk_ctx* ctx, int use_mip_lod, tvk_frame_stats* st) { return render_per_brick(ctx, st, true, use_mip_lod); }

int tvk_read_mip_max(tvk_ctx* ctx, float* dst) {
  if (!ctx || !dst) return TVK_ERR_INVALID;
  if (!ctx->buf[1]) return fail(ctx, TVK_ERR_INVALID, "no MIP frame rendered");
  cudaSetDevice(ctx->cfg.device);
  CU(cudaMemcpyAsync(dst, ctx->buf[1], (size_t)ctx->params.width * ctx->params.height * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return TVK_OK;
}

// ---- stereo (GLRenderer::ComputeViewAndProjection in stereo, GLRenderer::EndFrame) ----------------------------
int tvk_compute_stereo_view(tvk_render_params* left, tvk_render_params* right, uint32_t width, uint32_t height,
                            const float rotation[16], const float translation[16], const float eye[3], const float at[3],
                            const float up[3], float fov_deg, float z_near, float z_far, float screen_space_error,
                            float focal_length, float eye_dist) {
  if (!left || !right) return TVK_ERR_INVALID;
  // the mono call fills lod_factor, width, height, eye and BuildLookAt's view (with identity rotation / translation)
  const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  int rc = tvk_compute_view(left, width, height, ident, ident, eye, at, up, fov_deg, z_near, z_far, screen_space_error);
  if (rc) return rc;
  float view[16];
  std::memcpy(view, left->model_view, 64);
  *right = *left;
  // FLOATMATRIX4::BuildStereoLookAtAndProjection (Basics/Vectors.h:1215-1248), float arithmetic like the reference
  const float aspect = (float)width / (float)height;
  const float radians = float(3.14159265358979323846 / 180.0) * fov_deg / 2;
  const float wd2 = z_near * float(tan(radians));
  const float nfdl = z_near / focal_length;
  const float shift = eye_dist * nfdl;
  auto off_center = [&](float l, float r, float b, float t, float* m) {   // MatrixPerspectiveOffCenter :1286-1291
    std::memset(m, 0, 64);
    m[0] = 2.0f * z_near / (r - l); m[8] = (r + l) / (r - l);
    m[5] = 2.0f * z_near / (t - b); m[9] = (t + b) / (t - b);
    m[10] = -(z_far + z_near) / (z_far - z_near); m[14] = -2.0f * (z_far * z_near) / (z_far - z_near);
    m[11] = -1.0f;
  };
  off_center(-aspect * wd2 + shift, aspect * wd2 + shift, -wd2, wd2, left->projection);
  off_center(-aspect * wd2 - shift, aspect * wd2 - shift, -wd2, wd2, right->projection);
  auto mulf = [](const float* a, const float* b, float* o) {
    float t[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++)
        t[r * 4 + c] = a[r * 4] * b[c] + a[r * 4 + 1] * b[4 + c] + a[r * 4 + 2] * b[8 + c] + a[r * 4 + 3] * b[12 + c];
    std::memcpy(o, t, 64);
  };
  float tr[16], vl[16], vr[16], rt[16];
  std::memcpy(tr, ident, 64);
  tr[12] = eye_dist;  mulf(tr, view, vl);     // eye translation: mTranslate * mView
  tr[12] = -eye_dist; mulf(tr, view, vr);
  mulf(rotation, translation, rt);            // modelView[eye] = rotation * translation * view[eye] (GLRenderer.cpp:627-630)
  mulf(rt, vl, left->model_view);
  mulf(rt, vr, right->model_view);
  return TVK_OK;
}

int tvk_stereo_keep_eye(tvk_ctx* ctx, int eye) {
  if (!ctx || eye < 0 || eye > 1) return TVK_ERR_INVALID;
  if (!ctx->img_w) return fail(ctx, TVK_ERR_INVALID, "nothing rendered");
  cudaSetDevice(ctx->cfg.device);
  const size_t bytes = (size_t)ctx->img_w * ctx->img_h * sizeof(float4);
  if (!ctx->stereo_d[eye]) CU(cudaMalloc(&ctx->stereo_d[eye], bytes));
  CU(cudaMemcpyAsync(ctx->stereo_d[eye], result_image(ctx), bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return TVK_OK;
}

int tvk_stereo_compose(tvk_ctx* ctx, int mode, int eye_swap, int alternating_frame_id, float split_coord) {
  if (!ctx) return TVK_ERR_INVALID;
  if (mode < TVK_SM_RB || mode > TVK_SM_AF) return fail(ctx, TVK_ERR_INVALID, "invalid stereo mode");
  if (!ctx->img_w || !ctx->stereo_d[0] || !ctx->stereo_d[1]) return fail(ctx, TVK_ERR_INVALID, "stereo: both eye images must be kept first");
  cudaSetDevice(ctx->cfg.device);
  const float4* l = ctx->stereo_d[eye_swap ? 1 : 0];   // m_bStereoEyeSwap exchanges the bound units (GLRenderer.cpp:773-779)
  const float4* r = ctx->stereo_d[eye_swap ? 0 : 1];
  launch_stereo_compose(mode, l, r, ctx->buf[7], ctx->img_w, ctx->img_h, alternating_frame_id, split_coord, ctx->stream);
  CU(cudaGetLastError());
  ctx->result_buf = ctx->buf[7];
  return TVK_OK;
}

int tvk_set_clearview(tvk_ctx* ctx, int enable, double cv_isovalue, const float color[3], float size, float context_scale,
                      float border_scale, const float focus_pos[4]) {
  if (!ctx) return TVK_ERR_INVALID;
  ctx->cv.on = enable != 0;
  ctx->cv.iso = cv_isovalue;
  if (color) for (int i = 0; i < 3; i++) ctx->cv.color[i] = color[i];
  ctx->cv.size = size; ctx->cv.context = context_scale; ctx->cv.border = border_scale;
  if (focus_pos) for (int i = 0; i < 4; i++) ctx->cv.pos[i] = focus_pos[i];
  return TVK_OK;
}

int tvk_read_cv_buffers(tvk_ctx* ctx, float* cv_pos, float* cv_normal) {
  if (!ctx || !ctx->img_w || !ctx->cv_frame) return fail(ctx, TVK_ERR_INVALID, "no ClearView frame");
  cudaSetDevice(ctx->cfg.device);
  const size_t bytes = (size_t)ctx->img_w * ctx->img_h * 16;
  CU(cudaStreamSynchronize(ctx->stream));
  if (cv_pos) CU(cudaMemcpy(cv_pos, ctx->buf[1], bytes, cudaMemcpyDeviceToHost));
  if (cv_normal) CU(cudaMemcpy(cv_normal, ctx->buf[2], bytes, cudaMemcpyDeviceToHost));
  return TVK_OK;
}

int tvk_get_classic_brick_list(tvk_ctx* ctx, uint32_t* lod, tvk_classic_brick* dst, uint32_t cap, uint32_t* n) {
  if (!ctx || !n) return TVK_ERR_INVALID;
  if (lod) *lod = ctx->classic_lod;
  *n = (uint32_t)ctx->classic_list.size();
  if (dst) std::memcpy(dst, ctx->classic_list.data(), std::min<size_t>(cap, ctx->classic_list.size()) * sizeof(tvk_classic_brick));
  return TVK_OK;
}

}  // extern "C"

#include "tvk_sortlast.inc"
#include "tvk_procedural.inc"
#include "tvk_quantize.inc"
#include "tvk_rebrick.inc"
