// k_pool.cu -- page-table kernels: min/max-driven brick visibility (bit-exact with the reference's
// CPU pass), page-table patches, brick paging into the slot-linear pool, miss-list compaction.
// Replaces (reference file:line):
//   ContainsData<mode>                 Renderer/GL/GLVolumePool.cpp:962-989
//   RecomputeVisibilityForBrickPool    GLVolumePool.cpp:991-1018
//   RecomputeVisibilityForOctree       GLVolumePool.cpp:1020-1359 (one launch per level: a level
//                                      only reads the finished table of the level below)
//   UploadMetadataTexel / UploadBrick  GLVolumePool.cpp:673-717,944-953 (batched per subframe)
//   GLHashTable::GetData               Renderer/GL/GLHashTable.cpp:90-105 (device-side compaction
//                                      instead of reading back the whole table)
// All integer / fp64-compare work; HBM-bound, coalesced, grid-stride.
#include <algorithm>
#include "tvk_dev.h"

namespace tvk {
namespace {

constexpr int kSMs = 148;
inline int grid_for(uint64_t n, int block, int per_sm = 8) {
  uint64_t g = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)kSMs * per_sm;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

__device__ __forceinline__ bool contains(const VisConsts& V, const double* __restrict__ mm, uint32_t id) {
  const double smin = mm[4 * (size_t)id + 0], smax = mm[4 * (size_t)id + 1];
  switch (V.mode) {
    case TVK_RM_1DTRANS: return V.v[1] >= smin && V.v[0] <= smax;
    case TVK_RM_2DTRANS: {
      const double gmin = mm[4 * (size_t)id + 2], gmax = mm[4 * (size_t)id + 3];
      return (V.v[1] >= smin && V.v[0] <= smax) && (V.v[3] >= gmin && V.v[2] <= gmax);
    }
    default: return V.v[0] >= smin && V.v[0] <= smax;
  }
}

__global__ void vis_clear_kernel(uint32_t* meta, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    meta[i] = TVK_BI_MISSING;
}

// resident bricks: slot+3 when visible, BI_EMPTY otherwise (slot index == linear pool coordinate)
__global__ void vis_pool_kernel(uint32_t* meta, const int32_t* __restrict__ slot_brick, uint32_t n_slots,
                                const double* __restrict__ mm, const VisConsts V) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    const int32_t id = slot_brick[s];
    if (id < 0) continue;
    meta[id] = contains(V, mm, (uint32_t)id) ? s + TVK_BI_FLAG_COUNT : (uint32_t)TVK_BI_EMPTY;
  }
}

// counts: 0 total, 1 empty, 2 childEmpty, 3 emptyLeaf
__global__ void vis_level_kernel(uint32_t* meta, const double* __restrict__ mm, const VisConsts V, uint32_t lod,
                                 uint32_t* counts) {
  const uint32_t lx = V.layout[lod][0], ly = V.layout[lod][1], lz = V.layout[lod][2];
  const uint32_t n = lx * ly * lz;
  uint32_t c_empty = 0, c_child = 0, c_leaf = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t id = V.lod_offset[lod] + i;
    if (meta[id] >= TVK_BI_FLAG_COUNT) continue;
    if (contains(V, mm, id)) continue;
    if (lod == 0) {
      meta[id] = TVK_BI_CHILD_EMPTY;
      c_leaf++;
      continue;
    }
    const uint32_t x = i % lx, y = (i / lx) % ly, z = i / (lx * ly);
    const uint32_t cx = V.layout[lod - 1][0], cy = V.layout[lod - 1][1], cz = V.layout[lod - 1][2];
    bool all_child_empty = true;
    for (uint32_t dz = 0; dz < 2; dz++)
      for (uint32_t dy = 0; dy < 2; dy++)
        for (uint32_t dx = 0; dx < 2; dx++) {
          const uint32_t px = 2 * x + dx, py = 2 * y + dy, pz = 2 * z + dz;
          if (px >= cx || py >= cy || pz >= cz) continue;   // odd layouts: missing children do not count
          if (meta[V.lod_offset[lod - 1] + px + py * cx + pz * cx * cy] != TVK_BI_CHILD_EMPTY) all_child_empty = false;
        }
    if (all_child_empty) { meta[id] = TVK_BI_CHILD_EMPTY; c_child++; }
    else { meta[id] = TVK_BI_EMPTY; c_empty++; }
  }
  // warp-aggregated counters
  for (int o = 16; o > 0; o >>= 1) {
    c_empty += __shfl_down_sync(0xffffffffu, c_empty, o);
    c_child += __shfl_down_sync(0xffffffffu, c_child, o);
    c_leaf += __shfl_down_sync(0xffffffffu, c_leaf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c_empty) atomicAdd(counts + 1, c_empty);
    if (c_child) atomicAdd(counts + 2, c_child);
    if (c_leaf) atomicAdd(counts + 3, c_leaf);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counts + 0, n);
}

// final page-table values of the entries touched by one paging batch: ops[i].new_id <- ops[i].slot
// (evict_id is used as "value" here: see host) -- plain scatter of (index, value) pairs
__global__ void page_meta_kernel(uint32_t* meta, const PageOp* __restrict__ ops, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    meta[ops[i].new_id] = ops[i].slot;
}

// One CTA per paged brick: move the brick (own size, x fastest) into its slot (strides of the max brick size).
// POOL LAYOUT ("x-pair"): pool element x of a slot row holds the PAIR (voxel x, voxel x+1), so the
// traversal kernel fetches both operands of every x-lerp of its trilinear filters with one load (16 loads per sample in
// the 7-tap modes instead of 32).  The second half of a pair is exactly what the plain layout holds at x+1 -- also where
// that texel is stale (a smaller brick written over a larger one leaves the old texels in place, as glTexSubImage3D
// does in the reference's atlas) -- so every filter footprint reads the values it read before.  8- and 16-bit volumes
// only: float pools stay plain (tvk_math.cuh PairOf).
template <typename T, typename W>
__device__ __forceinline__ W make_pair(T lo, T hi);
template <> __device__ __forceinline__ uint16_t make_pair<uint8_t, uint16_t>(uint8_t lo, uint8_t hi) { return (uint16_t)(lo | ((uint16_t)hi << 8)); }
template <> __device__ __forceinline__ uint32_t make_pair<uint16_t, uint32_t>(uint16_t lo, uint16_t hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
template <> __device__ __forceinline__ float make_pair<float, float>(float lo, float) { return lo; }   // float pools stay plain
__device__ __forceinline__ uint8_t first_of(uint16_t w) { return (uint8_t)(w & 0xffu); }
__device__ __forceinline__ uint16_t first_of(uint32_t w) { return (uint16_t)(w & 0xffffu); }
__device__ __forceinline__ float first_of(float w) { return w; }

template <typename T, typename W>
__global__ void page_copy_kernel(W* pool, const T* __restrict__ store, const PageOp* __restrict__ ops, uint64_t slot_voxels,
                                 uint32_t tx, uint32_t ty, uint32_t tz, int src_is_slot_layout) {
  const PageOp op = ops[blockIdx.x];
  W* dst = pool + (uint64_t)op.slot * slot_voxels;
  const T* src = (const T*)((const unsigned char*)store + op.src_off);
  // source geometry: the device brick store keeps bricks padded to the slot shape, staged bricks come at their own size
  const uint32_t sx = src_is_slot_layout ? tx : op.size[0], sy = src_is_slot_layout ? ty : op.size[1],
                 sz = src_is_slot_layout ? tz : op.size[2];
  const uint32_t rows = sy * sz;
  // blockIdx.y = one of gridDim.y row chunks of the brick: a 128^3 brick is 16 384 rows, far too many for one CTA
  for (uint32_t r = blockIdx.y * (blockDim.x / 32) + threadIdx.x / 32; r < rows; r += gridDim.y * (blockDim.x / 32)) {
    const uint32_t y = r % sy, z = r / sy;
    const T* s = src + (uint64_t)r * sx;
    W* d = dst + ((uint64_t)z * ty + y) * tx;
    for (uint32_t x = threadIdx.x % 32; x < sx; x += 32) {
      const T lo = s[x];
      T hi;
      if (x + 1 < sx) hi = s[x + 1];
      else if (x + 1 < tx) hi = first_of(d[x + 1]);   // the texel behind a short row keeps what the slot held before
      else hi = lo;                                   // last texel of the slot row: never the low operand of a lerp
      d[x] = make_pair<T, W>(lo, hi);
    }
    // the pair in front of the row start is not touched: a row always starts at x = 0
  }
}

// parity tap (tvk_read_pool_slot): the plain voxels of one slot
template <typename T, typename W>
__global__ void slot_unpair_kernel(const W* __restrict__ slot, T* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = first_of(slot[i]);
}

// GLHashTable::GetData: gather the non-zero entries as (table index, value) pairs (the host orders
// them by index = the reference's scan order).  Warp-aggregated append.
__global__ void hash_compact_kernel(const uint32_t* __restrict__ hash, uint32_t n, uint32_t* out, uint32_t* count) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (n + stride - 1) / stride;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t r = 0; r < rounds; r++, i += stride) {
    const uint32_t e = i < n ? hash[i] : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, e != 0);
    if (m) {
      const int lane = threadIdx.x & 31;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (e) {
        const uint32_t k = base + __popc(m & ((1u << lane) - 1u));
        out[2 * k] = i;
        out[2 * k + 1] = e;
      }
    }
  }
}

}  // namespace

void launch_vis_clear(uint32_t* meta, uint64_t n, cudaStream_t s) {
  vis_clear_kernel<<<grid_for(n, 256), 256, 0, s>>>(meta, n);
}
void launch_vis_pool(uint32_t* meta, const int32_t* slot_brick, uint32_t n_slots, const double* minmax,
                     const VisConsts& vc, cudaStream_t s) {
  vis_pool_kernel<<<grid_for(n_slots, 256), 256, 0, s>>>(meta, slot_brick, n_slots, minmax, vc);
}
void launch_vis_level(uint32_t* meta, const double* minmax, const VisConsts& vc, uint32_t lod, uint32_t* counts,
                      cudaStream_t s) {
  const uint64_t n = (uint64_t)vc.layout[lod][0] * vc.layout[lod][1] * vc.layout[lod][2];
  vis_level_kernel<<<grid_for(n, 256), 256, 0, s>>>(meta, minmax, vc, lod, counts);
}
void launch_page_meta(uint32_t* meta, const PageOp* ops, uint32_t n, cudaStream_t s) {
  if (n) page_meta_kernel<<<grid_for(n, 256), 256, 0, s>>>(meta, ops, n);
}
void launch_page_copy(void* pool, const void* store, const PageOp* ops, uint32_t n, uint64_t slot_voxels,
                      uint32_t esize, const uint32_t total[3], int src_is_slot_layout, cudaStream_t s) {
  if (!n) return;
  // enough CTAs to fill the machine even for a handful of large bricks: one chunk per 64 rows, at most 32 per brick
  const uint32_t rows = total[1] * total[2];
  const dim3 g(n, std::max(1u, std::min(32u, rows / 64u)));
  if (esize == 1)
    page_copy_kernel<uint8_t, uint16_t><<<g, 256, 0, s>>>((uint16_t*)pool, (const uint8_t*)store, ops, slot_voxels, total[0],
                                                          total[1], total[2], src_is_slot_layout);
  else if (esize == 2)
    page_copy_kernel<uint16_t, uint32_t><<<g, 256, 0, s>>>((uint32_t*)pool, (const uint16_t*)store, ops, slot_voxels, total[0],
                                                           total[1], total[2], src_is_slot_layout);
  else
    page_copy_kernel<float, float><<<g, 256, 0, s>>>((float*)pool, (const float*)store, ops, slot_voxels, total[0], total[1],
                                                     total[2], src_is_slot_layout);
}
void launch_slot_unpair(const void* slot, void* out, uint64_t n_voxels, uint32_t esize, cudaStream_t s) {
  const unsigned g = (unsigned)std::min<uint64_t>((n_voxels + 255) / 256, 1184);
  if (esize == 1) slot_unpair_kernel<uint8_t, uint16_t><<<g, 256, 0, s>>>((const uint16_t*)slot, (uint8_t*)out, n_voxels);
  else if (esize == 2) slot_unpair_kernel<uint16_t, uint32_t><<<g, 256, 0, s>>>((const uint32_t*)slot, (uint16_t*)out, n_voxels);
  else slot_unpair_kernel<float, float><<<g, 256, 0, s>>>((const float*)slot, (float*)out, n_voxels);
}
// LPT tile schedule of the traversal kernel: counting sort of the tiles by the cost measured in the previous frame,
// largest first; one CTA (a frame has at most a few ten thousand tiles), ties in arrival order
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ order,
                                                          uint32_t shift, uint32_t split_cost) {
  __shared__ uint32_t hist[256], base[256];
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[min(255u, cost[i] >> shift)], 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int b = 255; b >= 0; b--) { base[b] = acc; acc += hist[b]; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&base[min(255u, cost[i] >> shift)], 1u)] = i;
  // order[n] = how many tiles (the first ones of the order) cost at least split_cost: the persistent kernel hands those
  // out in quarter tiles with helper lanes; at most a quarter of the frame
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t cnt = 0;
    const uint32_t b0 = min(255u, split_cost >> shift);
    for (int b = 255; b >= (int)b0 && split_cost != 0u; b--) cnt += hist[b];
    order[n] = min(cnt, n / 4u);
  }
}
void launch_tile_order(const uint32_t* cost, uint32_t n, uint32_t* order, uint32_t shift, uint32_t split_cost, cudaStream_t s) {
  if (n) tile_order_kernel<<<1, 1024, 0, s>>>(cost, n, order, shift, split_cost);
}
void launch_hash_compact(const uint32_t* hash, uint32_t n, uint32_t* out_list, uint32_t* out_count, cudaStream_t s) {
  hash_compact_kernel<<<grid_for(n, 256), 256, 0, s>>>(hash, n, out_list, out_count);
}

}  // namespace tvk
