// k_bricker.cu -- device-side data producer: seeded synthetic volumes, the 2x2x2 LOD pyramid,
// brick cutting with ghost cells and per-brick min/max.  Bit-exact with the reference converter
// (checked against the reference's own ExtendedOctreeConverter through oracle/_ref/ref_octree).
// Replaces (reference file:line):
//   ExtendedOctreeConverter::GetInputBrick / ClampToEdge   ExtendedOctreeConverter.cpp:288-462
//   DownsampleBricktoBrick / DownsampleBrick               ExtendedOctreeConverter.inc:1-356
//   VolumeTools::Filter (mean: sum as double / n, truncating cast)  VolumeTools.h:168-262
//   FillOverlap                                            ExtendedOctreeConverter.cpp:1203-1380
//   ComputeBrickStats / BrickStat                          ExtendedOctreeConverter.inc:408-444, .cpp:954-1000
//   MaxMinDataBlock::SetDataFromFlatVector                 IO/UVF/MaxMinDataBlock.cpp:175-195
// HBM-bound integer work: coalesced x-fastest reads/writes, one CTA per brick.
#include <cfloat>
#include <cuda.h>
#include "tvk_dev.h"
#include "tvk_synth.cuh"

namespace tvk {
namespace {

constexpr int kSMs = 148;

template <typename T>
__global__ void synth_kernel(T* dst, int kind, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t shift0, uint32_t seed) {
  const uint64_t n = (uint64_t)nx * ny * nz;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % nx), y = (uint32_t)((i / nx) % ny), z = (uint32_t)(i / ((uint64_t)nx * ny));
    dst[i] = from_u16<T>(synth_u16(kind, x, y, z, nx, ny, nz, shift0, seed));
  }
}

// ---------------------------------------------------------------------------------------------
// LOD pyramid: voxel of LOD l+1 = Filter over the existing voxels of the 2x2x2 block of LOD l
// ---------------------------------------------------------------------------------------------
// VolumeTools::Filter<T, F, bComputeMedian = true> (VolumeTools.h:168-262): 2 values -> the first; 4 -> the median of the
// first three; 8 -> a median of the first SEVEN (the eighth is ignored), by the reference's own compare-exchange network
template <typename T> __device__ __forceinline__ void order2(T& a, T& b) { if (a > b) { const T x = a; a = b; b = x; } }
template <typename T> __device__ __forceinline__ void insert_sorted4(T& a, T& b, T& c, T& d, T p) {
  if (p > c) { order2(d, p); }
  else if (p < b) { d = c; c = b; b = p; order2(a, b); }
  else { d = c; c = p; }
}
template <typename T> __device__ __forceinline__ T median_of(const T* v, int n) {
  if (n <= 2) return v[0];
  if (n == 4) { T a = v[0], b = v[1], c = v[2]; order2(a, b); order2(b, c); return a > b ? a : b; }
  T a = v[0], b = v[1], c = v[2], d = v[3];
  order2(a, b); order2(c, d); order2(a, c); order2(b, d); order2(b, c);
  insert_sorted4(a, b, c, d, v[4]);
  insert_sorted4(a, b, c, d, v[5]);
  const T m = d < v[6] ? d : v[6];
  return m > c ? m : c;
}

template <typename T, bool MEDIAN = false>
__global__ void downsample_kernel(const T* __restrict__ src, uint32_t sx, uint32_t sy, uint32_t sz, T* dst,
                                  uint32_t dx_, uint32_t dy_, uint32_t dz_) {
  const uint64_t n = (uint64_t)dx_ * dy_ * dz_;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % dx_), y = (uint32_t)((i / dx_) % dy_), z = (uint32_t)(i / ((uint64_t)dx_ * dy_));
    // axes of size 1 are not halved
    const uint32_t bx = sx > 1 ? 2 * x : x, by = sy > 1 ? 2 * y : y, bz = sz > 1 ? 2 * z : z;
    const uint32_t nx = (sx > 1 && bx + 1 < sx) ? 2 : 1;
    const uint32_t ny = (sy > 1 && by + 1 < sy) ? 2 : 1;
    const uint32_t nz = (sz > 1 && bz + 1 < sz) ? 2 : 1;
    double s = 0.0;
    T first = 0;
    T all[8];
    int cnt = 0;
    for (uint32_t a = 0; a < nx; a++)        // p0..p7 order of the reference: x-major, z-minor
      for (uint32_t b = 0; b < ny; b++)
        for (uint32_t c = 0; c < nz; c++) {
          const T v = src[(uint64_t)(bx + a) + (uint64_t)sx * ((by + b) + (uint64_t)sy * (bz + c))];
          if (cnt == 0) { first = v; s = (double)v; } else s = s + (double)v;
          if (MEDIAN) all[cnt] = v;
          cnt++;
        }
    if (MEDIAN) dst[i] = median_of(all, cnt);
    else dst[i] = cnt == 1 ? first : (T)(s / (double)cnt);
  }
}

// 16-bit fast path of the pyramid (all three source extents even, x extent a multiple of 4): one thread makes two
// x-adjacent voxels of the coarser level from four 8-byte loads (4 voxels of each of the 4 source rows) and writes
// them as one 32-bit word.  The eight summands are integers < 2^16, their sum is exact in double in any order, so
// T(sum / 8.0) equals the generic kernel's result bit for bit.
__global__ void downsample_u16x2_kernel(const uint2* __restrict__ src, uint32_t sx, uint32_t sy, uint32_t* dst,
                                        uint32_t dx_, uint32_t dy_, uint32_t dz_) {
  const uint32_t hx = dx_ / 2;                 // output words per row
  const uint64_t n = (uint64_t)hx * dy_ * dz_;
  const uint64_t row8 = sx / 4;                // uint2 per source row
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % hx), y = (uint32_t)((i / hx) % dy_), z = (uint32_t)(i / ((uint64_t)hx * dy_));
    const uint64_t r00 = ((uint64_t)(2 * z) * sy + 2 * y) * row8 + x;
    const uint2 a = __ldg(src + r00), b = __ldg(src + r00 + row8);
    const uint2 c = __ldg(src + r00 + (uint64_t)sy * row8), d = __ldg(src + r00 + (uint64_t)sy * row8 + row8);
    // .x holds voxels 0,1 (low, high half), .y voxels 2,3 of the 4-voxel run
    const uint32_t s0 = (a.x & 0xffffu) + (a.x >> 16) + (b.x & 0xffffu) + (b.x >> 16) + (c.x & 0xffffu) + (c.x >> 16) +
                        (d.x & 0xffffu) + (d.x >> 16);
    const uint32_t s1 = (a.y & 0xffffu) + (a.y >> 16) + (b.y & 0xffffu) + (b.y >> 16) + (c.y & 0xffffu) + (c.y >> 16) +
                        (d.y & 0xffffu) + (d.y >> 16);
    const uint32_t v0 = (uint32_t)(uint16_t)((double)s0 / 8.0), v1 = (uint32_t)(uint16_t)((double)s1 / 8.0);
    dst[i] = v0 | (v1 << 16);
  }
}

// ---------------------------------------------------------------------------------------------
// brick cutting + stats
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cut_brick_generic(const T* __restrict__ vol, T* store, const int32_t* __restrict__ store_index,
                                                  double* minmax, const CutConsts& C, uint64_t slot_voxels, const uint32_t b) {
  const uint32_t bx = b % C.layout[0], by = (b / C.layout[0]) % C.layout[1], bz = b / (C.layout[0] * C.layout[1]);
  const uint32_t bc[3] = {bx, by, bz};
  uint32_t bs[3];
  bool lo_border[3], hi_border[3];
  const int ov = (int)C.overlap;
#pragma unroll
  for (int i = 0; i < 3; i++) {   // ExtendedOctree::ComputeBrickSize (ExtendedOctree.cpp:276-285)
    const uint32_t core = C.brick[i] - 2 * C.overlap;
    const bool last = bc[i] == C.layout[i] - 1;
    const uint32_t rem = C.lod_size[i] % core;
    bs[i] = (last && rem) ? 2 * C.overlap + rem : C.brick[i];
    lo_border[i] = bc[i] == 0;
    hi_border[i] = last;
  }
  const int64_t org[3] = {(int64_t)bx * (C.brick[0] - 2 * ov) - ov, (int64_t)by * (C.brick[1] - 2 * ov) - ov,
                          (int64_t)bz * (C.brick[2] - 2 * ov) - ov};
  // sharded store (sort-last at the source): a brick another rank owns is only measured (min/max are global), not kept
  const int64_t at = store_index ? (int64_t)store_index[C.first_brick + b] : (int64_t)(C.first_brick + b);
  const bool keep = at >= 0;
  T* dst = store + (uint64_t)(keep ? at : 0) * slot_voxels;
  const uint32_t n = bs[0] * bs[1] * bs[2];
  T mn = 0, mx = 0;
  bool any = false;
  // Fast path (the bulk of a large 16-bit volume): a full 36^3 brick with a 2-voxel ghost that touches no domain
  // border.  Its slot is one contiguous 93 312-byte block and every source row is a 4-byte-aligned run of 72 bytes,
  // so the brick is moved as 32-bit words (two voxels each): coalesced 128-byte stores, no per-voxel div/mod.
  // Same voxels, same stale-corner rule, same min/max as the generic loop below.
  if (sizeof(T) == 2 && ov == 2 && C.brick[0] == 36 && C.brick[1] == 36 && C.brick[2] == 36 && bs[0] == 36 &&
      bs[1] == 36 && bs[2] == 36 && !lo_border[0] && !lo_border[1] && !lo_border[2] && !hi_border[0] && !hi_border[1] &&
      !hi_border[2] && (C.lod_size[0] & 1u) == 0u) {
    const uint64_t row_words = C.lod_size[0] / 2;
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(vol) +
                            ((uint64_t)org[2] * C.lod_size[1] + (uint64_t)org[1]) * row_words + (uint64_t)org[0] / 2;
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(dst);
    uint32_t wmn = 0xffffu, wmx = 0u;
    for (uint32_t w = threadIdx.x; w < 36u * 36u * 18u; w += blockDim.x) {
      const uint32_t row = w / 18u, i = w - row * 18u, lz = row / 36u, ly = row - lz * 36u;
      const int gx = i == 0u ? -1 : i == 17u ? 1 : 0, gy = ly < 2u ? -1 : ly >= 34u ? 1 : 0, gz = lz < 2u ? -1 : lz >= 34u ? 1 : 0;
      const bool stale = C.lod > 0 && ((gx == 1 && gy == 1 && gz == -1) || (gx == 1 && gy == -1 && gz == 1) ||
                                       (gx == -1 && gy == 1 && gz == 1));
      const uint32_t v = stale ? 0u : __ldg(src32 + ((uint64_t)lz * C.lod_size[1] + ly) * row_words + i);
      if (keep) dst32[w] = v;
      const uint32_t lo = v & 0xffffu, hi = v >> 16;
      wmn = min(wmn, min(lo, hi));
      wmx = max(wmx, max(lo, hi));
    }
    mn = (T)wmn; mx = (T)wmx; any = true;
  } else
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    int l[3] = {(int)(i % bs[0]), (int)((i / bs[0]) % bs[1]), (int)(i / (bs[0] * bs[1]))};
    const uint32_t di = (uint32_t)l[0] + C.brick[0] * ((uint32_t)l[1] + C.brick[1] * (uint32_t)l[2]);
    int g[3];   // ghost class per axis: -1 low ghost, 0 inner, +1 high ghost
    bool outside = false;
    int64_t gc[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      g[a] = l[a] < ov ? -1 : l[a] >= (int)bs[a] - ov ? 1 : 0;
      if (C.clamp) {   // ClampToEdge on the domain-border sides replicates the first/last inner plane
        if (g[a] < 0 && lo_border[a]) { l[a] = ov; g[a] = 0; }
        if (g[a] > 0 && hi_border[a]) { l[a] = (int)bs[a] - 1 - ov; g[a] = 0; }
      }
      gc[a] = org[a] + l[a];
      if (gc[a] < 0 || gc[a] >= (int64_t)C.lod_size[a]) outside = true;
    }
    T v = 0;
    // FillOverlap's copy order leaves three ghost corners of every LOD>=1 brick holding the still
    // unfilled ghost of a later brick: (right,bottom,front), (right,top,back), (left,bottom,back)
    const bool stale = C.lod > 0 && ((g[0] == 1 && g[1] == 1 && g[2] == -1) || (g[0] == 1 && g[1] == -1 && g[2] == 1) ||
                                     (g[0] == -1 && g[1] == 1 && g[2] == 1));
    if (!outside && !stale)
      v = vol[(uint64_t)gc[0] + (uint64_t)C.lod_size[0] * ((uint64_t)gc[1] + (uint64_t)C.lod_size[1] * (uint64_t)gc[2])];
    if (keep) dst[di] = v;
    if (!any) { mn = mx = v; any = true; }
    else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  }
  // block min/max (every stored voxel incl. ghost)
  __shared__ T s_mn[8], s_mx[8];
  __shared__ int s_any[8];
  const unsigned full = 0xffffffffu;
  for (int o = 16; o > 0; o >>= 1) {
    const T omn = __shfl_down_sync(full, mn, o), omx = __shfl_down_sync(full, mx, o);
    const int oany = __shfl_down_sync(full, (int)any, o);
    if (oany) {
      if (!any) { mn = omn; mx = omx; any = true; }
      else { mn = omn < mn ? omn : mn; mx = omx > mx ? omx : mx; }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_any[warp] = any; }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool have = false;
    T a = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      if (!s_any[w]) continue;
      if (!have) { a = s_mn[w]; c = s_mx[w]; have = true; }
      else { a = s_mn[w] < a ? s_mn[w] : a; c = s_mx[w] > c ? s_mx[w] : c; }
    }
    double* o = minmax + 4 * (uint64_t)(C.first_brick + b);
    o[0] = (double)a; o[1] = (double)c; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cut_bricks_kernel(const T* __restrict__ vol, T* store, const int32_t* __restrict__ store_index,
                                                         double* minmax, const CutConsts C, uint64_t slot_voxels) {
  cut_brick_generic<T>(vol, store, store_index, minmax, C, slot_voxels, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// The same brick cut through the TENSOR MEMORY ACCELERATOR (sm_100a; round 2).  A full 36^3 brick is
// a 3-D box of the LoD volume whose origin is (brick * inner - ghost): what one cp.async.bulk.tensor.3d delivers -- with
// the ghost cells that hang over the domain border zero-filled by the TMA unit itself (out-of-bound box elements read as 0
// = the reference's border rule without clamping).  Two alignment rules of the unit shape the kernel, the second one
// MEASURED on the B200 (scripts/probes/tma_probe.cu, profiles/r2o_tma_probe.txt): the box's inner extent must be a multiple
// of 16 bytes, and so must the inner START coordinate (x = 8 u16 voxels loads, x = 4 or x = -2 raises "illegal
// instruction"; y and z are free, negative included).  The box therefore starts at x0 rounded down to 16 bytes and is
// KBX = 64 / 48 / 40 voxels wide (u8 / u16 / f32); the 36 wanted voxels of a row begin `skip` voxels into the box row
// (skip is even for an even ghost width) and leave shared memory as 4-voxel words assembled from two 2-voxel reads,
// stored coalesced into the slot, which is one contiguous block.  The brick moves in 6 z-chunks of 6 slices through a ring
// of kStages shared-memory stages, each with its mbarrier: thread 0 issues the box loads, all 256 threads drain -- the copy
// engine fetches chunk c + kStages while chunk c is stored.  Min/max (every stored voxel incl. ghost) and the stale-corner
// rule of FillOverlap for LoD >= 1 ride along in the drain loop.  Bricks the box cannot describe (ragged last bricks,
// clamped borders) take the generic path inside the same launch.
// ---------------------------------------------------------------------------------------------
constexpr int kTB = 36, kChunkZ = 6, kChunks = kTB / kChunkZ, kStages = 3;
template <typename T> struct Vec4;
template <> struct Vec4<uint8_t> { using type = uchar4; using half = uchar2; };
template <> struct Vec4<uint16_t> { using type = ushort4; using half = ushort2; };
template <> struct Vec4<float> { using type = float4; using half = float2; };
// voxels per 16 bytes; box width = 36 + the largest even skip, rounded up to 16 bytes
template <typename T> struct TmaBoxX {
  static constexpr int align = (int)(16 / sizeof(T));
  static constexpr int value = (kTB + align - 2 + align - 1) / align * align;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256) cut_bricks_tma_kernel(const CUtensorMap* __restrict__ tmp, const T* __restrict__ vol, T* store,
                                                             const int32_t* __restrict__ store_index, double* minmax,
                                                             const CutConsts C, uint64_t slot_voxels) {
  constexpr int KBX = TmaBoxX<T>::value;
  constexpr uint32_t kStageBytes = KBX * kTB * kChunkZ * sizeof(T);
  constexpr int KA = TmaBoxX<T>::align;
  using V = typename Vec4<T>::type;
  using H = typename Vec4<T>::half;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ __align__(8) uint64_t bars[kStages];
  const uint32_t b = blockIdx.x;
  const uint32_t bx = b % C.layout[0], by = (b / C.layout[0]) % C.layout[1], bz = b / (C.layout[0] * C.layout[1]);
  const uint32_t bc[3] = {bx, by, bz};
  bool full = true, border = false;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint32_t core = C.brick[i] - 2 * C.overlap;
    const bool last = bc[i] == C.layout[i] - 1;
    full = full && !(last && (C.lod_size[i] % core));
    border = border || bc[i] == 0 || last;
  }
  if (!full || (C.clamp && border)) {          // block-uniform: ragged or clamped bricks go the generic way
    cut_brick_generic<T>(vol, store, store_index, minmax, C, slot_voxels, b);
    return;
  }
  const int ov = (int)C.overlap;
  const int xw = (int)(bx * (kTB - 2 * ov)) - ov, y0 = (int)(by * (kTB - 2 * ov)) - ov, z0 = (int)(bz * (kTB - 2 * ov)) - ov;
  const int skip = ((xw % KA) + KA) % KA;      // the wanted row starts `skip` voxels into the 16-byte aligned box row
  const int x0 = xw - skip;
  const int64_t at = store_index ? (int64_t)store_index[C.first_brick + b] : (int64_t)(C.first_brick + b);
  const bool keep = at >= 0;
  V* dst = reinterpret_cast<V*>(store + (uint64_t)(keep ? at : 0) * slot_voxels);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int c = 0; c < kStages && c < kChunks; c++) {
      mbar_expect_tx(&bars[c], kStageBytes);
      tma_load_3d(tma_smem + (size_t)c * kStageBytes, tmp, x0, y0, z0 + c * kChunkZ, &bars[c]);
    }
  }
  __syncthreads();
  T mn = 0, mx = 0;
  bool any = false;
  const bool stale_rule = C.lod > 0;
  constexpr uint32_t kWordsPerChunk = kTB * kChunkZ * (kTB / 4);   // 9 words per row
  const uint32_t trow = threadIdx.x / (kTB / 4), g = threadIdx.x - trow * (kTB / 4);
  for (int c = 0; c < kChunks; c++) {
    const int stage = c % kStages;
    mbar_wait(&bars[stage], (uint32_t)((c / kStages) & 1));
    const T* tile = reinterpret_cast<const T*>(tma_smem + (size_t)stage * kStageBytes);
    // 252 of the 256 threads: thread = (row within a group of 28 rows, word of the row); no division in the loop
    for (uint32_t row = trow; row < (uint32_t)(kTB * kChunkZ) && threadIdx.x < 252u; row += 28u) {
      const uint32_t w = row * (kTB / 4) + g;
      const T* src = tile + (size_t)row * KBX + skip + g * 4;
      const H h0 = *reinterpret_cast<const H*>(src), h1 = *reinterpret_cast<const H*>(src + 2);
      V v;
      v.x = h0.x; v.y = h0.y; v.z = h1.x; v.w = h1.y;
      if (stale_rule) {
        // FillOverlap's copy order leaves three ghost corners of every LoD >= 1 brick zero (see the generic path)
        const uint32_t lz = (uint32_t)c * kChunkZ + row / kTB, ly = row % kTB;
        const int gy = (int)ly < ov ? -1 : (int)ly >= kTB - ov ? 1 : 0, gz = (int)lz < ov ? -1 : (int)lz >= kTB - ov ? 1 : 0;
        if (gy != 0 && gz != 0 && (g == 0 || g == kTB / 4 - 1)) {
          T e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int lx = (int)g * 4 + k;
            const int gx = lx < ov ? -1 : lx >= kTB - ov ? 1 : 0;
            if ((gx == 1 && gy == 1 && gz == -1) || (gx == 1 && gy == -1 && gz == 1) || (gx == -1 && gy == 1 && gz == 1)) e[k] = 0;
          }
          v.x = e[0]; v.y = e[1]; v.z = e[2]; v.w = e[3];
        }
      }
      if (keep) dst[(size_t)c * kWordsPerChunk + w] = v;
      const T lo = min(min(v.x, v.y), min(v.z, v.w)), hi = max(max(v.x, v.y), max(v.z, v.w));
      if (!any) { mn = lo; mx = hi; any = true; }
      else { mn = lo < mn ? lo : mn; mx = hi > mx ? hi : mx; }
    }
    __syncthreads();                                         // every thread is done with this stage
    if (threadIdx.x == 0 && c + kStages < kChunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before the async-proxy refill
      mbar_expect_tx(&bars[stage], kStageBytes);
      tma_load_3d(tma_smem + (size_t)stage * kStageBytes, tmp, x0, y0, z0 + (c + kStages) * kChunkZ, &bars[stage]);
    }
  }
  // block min/max (every stored voxel incl. ghost), as in the generic path
  __shared__ T t_mn[8], t_mx[8];
  const unsigned fullm = 0xffffffffu;
  for (int o = 16; o > 0; o >>= 1) {
    const T omn = __shfl_down_sync(fullm, mn, o), omx = __shfl_down_sync(fullm, mx, o);
    const int oany = __shfl_down_sync(fullm, (int)any, o);
    if (oany) {                                            // threads 252..255 hold no word
      if (!any) { mn = omn; mx = omx; any = true; }
      else { mn = omn < mn ? omn : mn; mx = omx > mx ? omx : mx; }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { t_mn[warp] = mn; t_mx[warp] = mx; }   // lane 0 of every warp holds words (thread 224 < 252)
  __syncthreads();
  if (threadIdx.x == 0) {
    T a = t_mn[0], cmax = t_mx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { a = t_mn[w] < a ? t_mn[w] : a; cmax = t_mx[w] > cmax ? t_mx[w] : cmax; }
    double* o = minmax + 4 * (uint64_t)(C.first_brick + b);
    o[0] = (double)a; o[1] = (double)cmax; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

// ---------------------------------------------------------------------------------------------
// min/max of bricks that arrive through the streaming path (file reader / callback) without a MaxMin block:
// one CTA per staged brick (tightly packed at its own size), every stored voxel incl. ghost, as doubles
// (ComputeBrickStats, ExtendedOctreeConverter.inc:408-444; MaxMinDataBlock layout)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) brick_minmax_kernel(const unsigned char* __restrict__ staged, const PageOp* ops,
                                                           double* minmax) {
  const PageOp op = ops[blockIdx.x];
  const T* src = reinterpret_cast<const T*>(staged + op.src_off);
  const uint32_t n = op.size[0] * op.size[1] * op.size[2];
  T mn = src[0], mx = src[0];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const T v = src[i];
    mn = v < mn ? v : mn; mx = v > mx ? v : mx;
  }
  const unsigned full = 0xffffffffu;
  for (int o = 16; o > 0; o >>= 1) {
    const T omn = __shfl_down_sync(full, mn, o), omx = __shfl_down_sync(full, mx, o);
    mn = omn < mn ? omn : mn; mx = omx > mx ? omx : mx;
  }
  __shared__ T s_mn[8], s_mx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { mn = s_mn[w] < mn ? s_mn[w] : mn; mx = s_mx[w] > mx ? s_mx[w] : mx; }
    double* o = minmax + 4 * (uint64_t)op.new_id;
    o[0] = (double)mn; o[1] = (double)mx; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

// ---------------------------------------------------------------------------------------------
// colour (RGBA8) volumes on the device bricker.  A colour voxel is a 4-byte word: the brick CUT (gather with ghost, zero
// fill, clamp, stale corners -- and the TMA box loads for 36^3 bricks) is the float instantiation used as a pure 4-byte
// mover (no arithmetic touches the payload; its float min / max are meaningless and are overwritten by the pass below).
// The pyramid filters every component like a scalar (mean in double, truncating cast) -- except at the single voxel whose
// source block is ONE voxel (all three source sizes odd, last index), where the converter writes component 0 into every
// component (ExtendedOctreeConverter.inc:232-246, orc_octree.c build_impl).
// ---------------------------------------------------------------------------------------------
__global__ void downsample_rgba_kernel(const uchar4* __restrict__ src, uint32_t sx, uint32_t sy, uint32_t sz, uchar4* dst,
                                       uint32_t dx_, uint32_t dy_, uint32_t dz_) {
  const uint64_t n = (uint64_t)dx_ * dy_ * dz_;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % dx_), y = (uint32_t)((i / dx_) % dy_), z = (uint32_t)(i / ((uint64_t)dx_ * dy_));
    const uint32_t bx = sx > 1 ? 2 * x : x, by = sy > 1 ? 2 * y : y, bz = sz > 1 ? 2 * z : z;
    const uint32_t nx = (sx > 1 && bx + 1 < sx) ? 2 : 1;
    const uint32_t ny = (sy > 1 && by + 1 < sy) ? 2 : 1;
    const uint32_t nz = (sz > 1 && bz + 1 < sz) ? 2 : 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    uchar4 first = make_uchar4(0, 0, 0, 0);
    int cnt = 0;
    for (uint32_t a = 0; a < nx; a++)
      for (uint32_t b = 0; b < ny; b++)
        for (uint32_t c = 0; c < nz; c++) {
          const uchar4 v = src[(uint64_t)(bx + a) + (uint64_t)sx * ((by + b) + (uint64_t)sy * (bz + c))];
          if (cnt == 0) { first = v; s0 = (double)v.x; s1 = (double)v.y; s2 = (double)v.z; s3 = (double)v.w; }
          else { s0 = s0 + (double)v.x; s1 = s1 + (double)v.y; s2 = s2 + (double)v.z; s3 = s3 + (double)v.w; }
          cnt++;
        }
    // cnt == 1 <=> every source size is odd (or 1) and this is the last voxel: the converter's corner rule
    if (cnt == 1) dst[i] = make_uchar4(first.x, first.x, first.x, first.x);
    else {
      const double k = (double)cnt;
      dst[i] = make_uchar4((unsigned char)(s0 / k), (unsigned char)(s1 / k), (unsigned char)(s2 / k), (unsigned char)(s3 / k));
    }
  }
}

// alpha statistics of the bricks of one level, read back from the brick store (every stored voxel incl. ghost)
__global__ void __launch_bounds__(256) store_alpha_minmax_kernel(const uchar4* __restrict__ store, const CutConsts C, double* minmax,
                                                                 uint64_t slot_voxels) {
  const uint32_t b = blockIdx.x;
  const uint32_t bc[3] = {b % C.layout[0], (b / C.layout[0]) % C.layout[1], b / (C.layout[0] * C.layout[1])};
  uint32_t bs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {   // ExtendedOctree::ComputeBrickSize (ExtendedOctree.cpp:276-285)
    const uint32_t core = C.brick[i] - 2 * C.overlap;
    const uint32_t rem = C.lod_size[i] % core;
    bs[i] = (bc[i] == C.layout[i] - 1 && rem) ? 2 * C.overlap + rem : C.brick[i];
  }
  const uchar4* src = store + (uint64_t)(C.first_brick + b) * slot_voxels;
  const uint32_t n = bs[0] * bs[1] * bs[2];
  uint32_t mn = 255u, mx = 0u;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t lx = i % bs[0], ly = (i / bs[0]) % bs[1], lz = i / (bs[0] * bs[1]);
    const uint32_t v = src[lx + C.brick[0] * (ly + C.brick[1] * lz)].w;
    mn = min(mn, v); mx = max(mx, v);
  }
  mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
  __shared__ uint32_t s_mn[8], s_mx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { mn = min(mn, s_mn[w]); mx = max(mx, s_mx[w]); }
    double* o = minmax + 4 * (uint64_t)(C.first_brick + b);
    o[0] = (double)mn; o[1] = (double)mx; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

// colour bricks (four interleaved 8-bit components): the statistics a renderer sees are the ALPHA component's
// (UVFDataset::MaxMinForKey -> GetValue(i, 3), uvfDataset.cpp:1188)
__global__ void __launch_bounds__(256) brick_minmax_alpha_kernel(const unsigned char* __restrict__ staged, const PageOp* ops,
                                                                 double* minmax) {
  const PageOp op = ops[blockIdx.x];
  const uchar4* src = reinterpret_cast<const uchar4*>(staged + op.src_off);
  const uint32_t n = op.size[0] * op.size[1] * op.size[2];
  uint32_t mn = src[0].w, mx = src[0].w;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t v = src[i].w;
    mn = min(mn, v); mx = max(mx, v);
  }
  mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
  __shared__ uint32_t s_mn[8], s_mx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { mn = min(mn, s_mn[w]); mx = max(mx, s_mx[w]); }
    double* o = minmax + 4 * (uint64_t)op.new_id;
    o[0] = (double)mn; o[1] = (double)mx; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

// ---------------------------------------------------------------------------------------------
// procedural dataset: min/max of a brick WITHOUT materialising it -- one CTA per brick evaluates the analytic field at
// the brick's voxels (ghost included; 0 outside the level's grid) and reduces.  Compute-bound integer work (~500 integer
// instructions per voxel); the table it fills is what MaxMinDataBlock holds for a converted file.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) proc_minmax_kernel(const ProcConsts C, uint32_t first, double* minmax) {
  const uint32_t id = first + blockIdx.x;
  uint32_t lod = 0;
  while (lod + 1 < C.lod_count && id >= C.lod_offset[lod + 1]) lod++;
  const uint32_t local = id - C.lod_offset[lod];
  const uint32_t* L = C.layout[lod];
  const uint32_t* N = C.lod_size[lod];
  const uint32_t bc[3] = {local % L[0], (local / L[0]) % L[1], local / (L[0] * L[1])};
  uint32_t bs[3];
  int64_t org[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint32_t core = C.brick[i] - 2 * C.overlap;
    const bool last = bc[i] == L[i] - 1;
    const uint32_t rem = N[i] % core;
    bs[i] = (last && rem) ? 2 * C.overlap + rem : C.brick[i];
    org[i] = (int64_t)bc[i] * core - C.overlap;
  }
  const uint32_t shift0 = synth_shift0(max(N[0], max(N[1], N[2])));
  const uint32_t n = bs[0] * bs[1] * bs[2];
  T mn = 0, mx = 0;
  bool any = false;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t x = org[0] + (int64_t)(i % bs[0]), y = org[1] + (int64_t)((i / bs[0]) % bs[1]), z = org[2] + (int64_t)(i / (bs[0] * bs[1]));
    T v = 0;
    if (x >= 0 && y >= 0 && z >= 0 && x < (int64_t)N[0] && y < (int64_t)N[1] && z < (int64_t)N[2])
      v = from_u16<T>(synth_u16(C.kind, (uint32_t)x, (uint32_t)y, (uint32_t)z, N[0], N[1], N[2], shift0, C.seed));
    if (!any) { mn = mx = v; any = true; }
    else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  }
  __shared__ T s_mn[8], s_mx[8];
  __shared__ int s_any[8];
  const unsigned full = 0xffffffffu;
  for (int o = 16; o > 0; o >>= 1) {
    const T omn = __shfl_down_sync(full, mn, o), omx = __shfl_down_sync(full, mx, o);
    const int oany = __shfl_down_sync(full, (int)any, o);
    if (oany) {
      if (!any) { mn = omn; mx = omx; any = true; }
      else { mn = omn < mn ? omn : mn; mx = omx > mx ? omx : mx; }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_any[warp] = any; }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool have = false;
    T a = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      if (!s_any[w]) continue;
      if (!have) { a = s_mn[w]; c = s_mx[w]; have = true; }
      else { a = s_mn[w] < a ? s_mn[w] : a; c = s_mx[w] > c ? s_mx[w] : c; }
    }
    double* o = minmax + 4 * (uint64_t)blockIdx.x;
    o[0] = (double)a; o[1] = (double)c; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

// ---------------------------------------------------------------------------------------------
// DynamicBrickingDS on the device (IO/DynamicBrickingDS.cpp): a dataset stored in LARGE bricks is re-cut
// into the pool's small bricks when it is loaded.  A target brick (ghost included) is a sub-box of exactly ONE source
// brick -- the target's inner size divides the source's and both carry the same ghost width (IOManager.cpp:1296-1313,
// DynamicBrickingDS.cpp:1092-1105) -- so ghost voxels at a source brick's border come from that source brick's own
// ghost, as in DynamicBrickingDS::GetBrick.  One CTA per target brick of the staged source brick: copy into the store
// slot, min / max of every stored voxel (MinMaxMode MM_PRECOMPUTE: minmax_brick, BMinMax.cpp:8-14).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rebrick_kernel(const T* __restrict__ src, const RebrickConsts C, T* store, double* minmax,
                                                      uint64_t slot_voxels) {
  const uint32_t b = blockIdx.x;
  const uint32_t t[3] = {b % C.ratio[0], (b / C.ratio[0]) % C.ratio[1], b / (C.ratio[0] * C.ratio[1])};
  uint32_t gt[3], bs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    gt[i] = C.src_brick[i] * C.ratio[i] + t[i];
    if (gt[i] >= C.layout[i]) return;                     // ragged source brick: fewer target bricks along this axis
    const uint32_t core = C.brick[i] - 2 * C.overlap;
    const uint32_t rem = C.lod_size[i] % core;
    bs[i] = (gt[i] == C.layout[i] - 1 && rem) ? 2 * C.overlap + rem : C.brick[i];
  }
  const uint32_t core[3] = {C.brick[0] - 2 * C.overlap, C.brick[1] - 2 * C.overlap, C.brick[2] - 2 * C.overlap};
  const uint64_t slot = C.first_brick + gt[0] + (uint64_t)C.layout[0] * (gt[1] + (uint64_t)C.layout[1] * gt[2]);
  T* dst = store + slot * slot_voxels;
  const uint32_t n = bs[0] * bs[1] * bs[2];
  T mn = 0, mx = 0;
  bool any = false;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t lx = i % bs[0], ly = (i / bs[0]) % bs[1], lz = i / (bs[0] * bs[1]);
    const uint64_t si = (uint64_t)(t[0] * core[0] + lx) + (uint64_t)C.src_size[0] * ((t[1] * core[1] + ly) + (uint64_t)C.src_size[1] * (t[2] * core[2] + lz));
    const T v = src[si];
    dst[lx + C.brick[0] * (ly + C.brick[1] * lz)] = v;
    if (!any) { mn = mx = v; any = true; }
    else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  }
  __shared__ T s_mn[8], s_mx[8];
  __shared__ int s_any[8];
  for (int o = 16; o > 0; o >>= 1) {
    const T omn = __shfl_down_sync(0xffffffffu, mn, o), omx = __shfl_down_sync(0xffffffffu, mx, o);
    const int oany = __shfl_down_sync(0xffffffffu, (int)any, o);
    if (oany) {
      if (!any) { mn = omn; mx = omx; any = true; }
      else { mn = omn < mn ? omn : mn; mx = omx > mx ? omx : mx; }
    }
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; s_any[threadIdx.x >> 5] = any; }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool have = false;
    T a = 0, c = 0;
    for (int w = 0; w < 8; w++) {
      if (!s_any[w]) continue;
      if (!have) { a = s_mn[w]; c = s_mx[w]; have = true; }
      else { a = s_mn[w] < a ? s_mn[w] : a; c = s_mx[w] > c ? s_mx[w] : c; }
    }
    double* o = minmax + 4 * slot;
    o[0] = (double)a; o[1] = (double)c; o[2] = -DBL_MAX; o[3] = DBL_MAX;
  }
}

inline int grid_for(uint64_t n, int block) {
  uint64_t g = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)kSMs * 16;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace

void launch_synth(void* dst, int kind, const uint32_t size[3], int dtype, uint32_t seed, cudaStream_t s) {
  const uint64_t n = (uint64_t)size[0] * size[1] * size[2];
  uint32_t mx = size[0] > size[1] ? size[0] : size[1];
  mx = mx > size[2] ? mx : size[2];
  uint32_t lg = 0;
  while ((2u << lg) <= mx) lg++;          // floor(log2(max size))
  const uint32_t shift0 = lg > 3 ? lg - 3 : 0;
  const int g = grid_for(n, 256);
  switch (dtype) {
    case TVK_U8: synth_kernel<uint8_t><<<g, 256, 0, s>>>((uint8_t*)dst, kind, size[0], size[1], size[2], shift0, seed); break;
    case TVK_U16: synth_kernel<uint16_t><<<g, 256, 0, s>>>((uint16_t*)dst, kind, size[0], size[1], size[2], shift0, seed); break;
    default: synth_kernel<float><<<g, 256, 0, s>>>((float*)dst, kind, size[0], size[1], size[2], shift0, seed); break;
  }
}

void launch_downsample(const void* src, const uint32_t ss[3], void* dst, const uint32_t ds[3], int dtype, int median,
                       cudaStream_t s) {
  const uint64_t n = (uint64_t)ds[0] * ds[1] * ds[2];
  if (dtype == TVK_RGBA8) {   // mean filter only (the host refuses the median for colour data)
    downsample_rgba_kernel<<<grid_for(n, 256), 256, 0, s>>>((const uchar4*)src, ss[0], ss[1], ss[2], (uchar4*)dst, ds[0], ds[1], ds[2]);
    return;
  }
  if (median) {
    const int g = grid_for(n, 256);
    switch (dtype) {
      case TVK_U8: downsample_kernel<uint8_t, true><<<g, 256, 0, s>>>((const uint8_t*)src, ss[0], ss[1], ss[2], (uint8_t*)dst, ds[0], ds[1], ds[2]); break;
      case TVK_U16: downsample_kernel<uint16_t, true><<<g, 256, 0, s>>>((const uint16_t*)src, ss[0], ss[1], ss[2], (uint16_t*)dst, ds[0], ds[1], ds[2]); break;
      default: downsample_kernel<float, true><<<g, 256, 0, s>>>((const float*)src, ss[0], ss[1], ss[2], (float*)dst, ds[0], ds[1], ds[2]); break;
    }
    return;
  }
  if (dtype == TVK_U16 && ss[0] % 4 == 0 && ss[1] % 2 == 0 && ss[2] % 2 == 0 && ss[0] > 1 && ss[1] > 1 && ss[2] > 1) {
    downsample_u16x2_kernel<<<grid_for(n / 2, 256), 256, 0, s>>>((const uint2*)src, ss[0], ss[1], (uint32_t*)dst, ds[0], ds[1], ds[2]);
    return;
  }
  const int g = grid_for(n, 256);
  switch (dtype) {
    case TVK_U8: downsample_kernel<uint8_t><<<g, 256, 0, s>>>((const uint8_t*)src, ss[0], ss[1], ss[2], (uint8_t*)dst, ds[0], ds[1], ds[2]); break;
    case TVK_U16: downsample_kernel<uint16_t><<<g, 256, 0, s>>>((const uint16_t*)src, ss[0], ss[1], ss[2], (uint16_t*)dst, ds[0], ds[1], ds[2]); break;
    default: downsample_kernel<float><<<g, 256, 0, s>>>((const float*)src, ss[0], ss[1], ss[2], (float*)dst, ds[0], ds[1], ds[2]); break;
  }
}

void launch_rebrick(const void* src, const RebrickConsts& C, void* store, double* minmax, int dtype, uint64_t slot_bytes, cudaStream_t s) {
  const uint32_t n = C.ratio[0] * C.ratio[1] * C.ratio[2];
  switch (dtype) {
    case TVK_U8: rebrick_kernel<uint8_t><<<n, 256, 0, s>>>((const uint8_t*)src, C, (uint8_t*)store, minmax, slot_bytes); break;
    case TVK_U16: rebrick_kernel<uint16_t><<<n, 256, 0, s>>>((const uint16_t*)src, C, (uint16_t*)store, minmax, slot_bytes / 2); break;
    default: rebrick_kernel<float><<<n, 256, 0, s>>>((const float*)src, C, (float*)store, minmax, slot_bytes / 4); break;
  }
}

void launch_proc_minmax(const ProcConsts& pc, uint32_t first, uint32_t count, double* minmax, cudaStream_t s) {
  if (!count) return;
  switch (pc.dtype) {
    case TVK_U8: proc_minmax_kernel<uint8_t><<<count, 256, 0, s>>>(pc, first, minmax); break;
    case TVK_U16: proc_minmax_kernel<uint16_t><<<count, 256, 0, s>>>(pc, first, minmax); break;
    default: proc_minmax_kernel<float><<<count, 256, 0, s>>>(pc, first, minmax); break;
  }
}

void launch_brick_minmax(const void* staged, const PageOp* ops, uint32_t n, double* minmax, int dtype, cudaStream_t s) {
  const unsigned char* p = static_cast<const unsigned char*>(staged);
  switch (dtype) {
    case TVK_U8: brick_minmax_kernel<uint8_t><<<n, 256, 0, s>>>(p, ops, minmax); break;
    case TVK_U16: brick_minmax_kernel<uint16_t><<<n, 256, 0, s>>>(p, ops, minmax); break;
    case TVK_RGBA8: brick_minmax_alpha_kernel<<<n, 256, 0, s>>>(p, ops, minmax); break;
    default: brick_minmax_kernel<float><<<n, 256, 0, s>>>(p, ops, minmax); break;
  }
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {      // the driver entry point, without linking libcuda
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

template <typename T>
bool launch_cut_tma(const void* lod_vol, void* store, const int32_t* store_index, double* minmax, const CutConsts& cc,
                    uint64_t slot_voxels, CUtensorMapDataType dt, cudaStream_t s) {
  constexpr int KBX = TmaBoxX<T>::value;
  const uint64_t pitch = (uint64_t)cc.lod_size[0] * sizeof(T);
  EncodeTiledFn enc = encode_tiled();
  // what a tiled tensor map needs: 16-byte aligned base and strides, a box no larger than the tensor's rank allows
  if (!enc || cc.brick[0] != kTB || cc.brick[1] != kTB || cc.brick[2] != kTB || pitch % 16 != 0 || ((uintptr_t)lod_vol & 15) != 0 ||
      cc.lod_size[0] < (uint32_t)KBX || 2 * cc.overlap >= (uint32_t)kTB || (cc.overlap & 1u))
    return false;
  static const bool off = std::getenv("TVK_BRICKER_TMA") && std::getenv("TVK_BRICKER_TMA")[0] == '0';
  if (off) return false;
  CUtensorMap tm;
  const cuuint64_t dims[3] = {cc.lod_size[0], cc.lod_size[1], cc.lod_size[2]};
  const cuuint64_t strides[2] = {pitch, pitch * cc.lod_size[1]};
  const cuuint32_t box[3] = {(cuuint32_t)KBX, (cuuint32_t)kTB, (cuuint32_t)kChunkZ};
  const cuuint32_t es[3] = {1, 1, 1};
  if (enc(&tm, dt, 3, const_cast<void*>(lod_vol), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  const size_t smem = (size_t)kStages * KBX * kTB * kChunkZ * sizeof(T);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(cut_bricks_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set = true;
  }
  // the descriptor lives in global memory (a small ring: one per level in flight), written by a stream-ordered copy
  static CUtensorMap* ring = nullptr;
  static int next = 0;
  constexpr int kRing = 32;
  if (!ring && cudaMalloc(&ring, kRing * sizeof(CUtensorMap)) != cudaSuccess) { cudaGetLastError(); ring = nullptr; return false; }
  CUtensorMap* slot = ring + (next++ % kRing);
  if (cudaMemcpyAsync(slot, &tm, sizeof(tm), cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return false; }
  const uint32_t n = cc.layout[0] * cc.layout[1] * cc.layout[2];
  cut_bricks_tma_kernel<T><<<n, 256, smem, s>>>(slot, (const T*)lod_vol, (T*)store, store_index, minmax, cc, slot_voxels);
  return true;
}
}  // namespace

void launch_cut_bricks(const void* lod_vol, void* store, const int32_t* store_index, double* minmax, const CutConsts& cc,
                       int dtype, uint64_t slot_bytes, cudaStream_t s) {
  const uint32_t n = cc.layout[0] * cc.layout[1] * cc.layout[2];
  if (dtype == TVK_RGBA8) {   // 4-byte voxels moved by the float instantiation, alpha statistics from the store (no shard)
    if (!launch_cut_tma<float>(lod_vol, store, nullptr, minmax, cc, slot_bytes / 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, s))
      cut_bricks_kernel<float><<<n, 256, 0, s>>>((const float*)lod_vol, (float*)store, nullptr, minmax, cc, slot_bytes / 4);
    store_alpha_minmax_kernel<<<n, 256, 0, s>>>((const uchar4*)store, cc, minmax, slot_bytes / 4);
    return;
  }
  // 36^3 bricks of a level whose rows are 16-byte multiples: 3-D box loads through the TMA unit
  if (dtype == TVK_U8 && launch_cut_tma<uint8_t>(lod_vol, store, store_index, minmax, cc, slot_bytes, CU_TENSOR_MAP_DATA_TYPE_UINT8, s)) return;
  if (dtype == TVK_U16 && launch_cut_tma<uint16_t>(lod_vol, store, store_index, minmax, cc, slot_bytes / 2, CU_TENSOR_MAP_DATA_TYPE_UINT16, s)) return;
  if (dtype == TVK_F32 && launch_cut_tma<float>(lod_vol, store, store_index, minmax, cc, slot_bytes / 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, s)) return;
  switch (dtype) {
    case TVK_U8: cut_bricks_kernel<uint8_t><<<n, 256, 0, s>>>((const uint8_t*)lod_vol, (uint8_t*)store, store_index, minmax, cc, slot_bytes); break;
    case TVK_U16: cut_bricks_kernel<uint16_t><<<n, 256, 0, s>>>((const uint16_t*)lod_vol, (uint16_t*)store, store_index, minmax, cc, slot_bytes / 2); break;
    default: cut_bricks_kernel<float><<<n, 256, 0, s>>>((const float*)lod_vol, (float*)store, store_index, minmax, cc, slot_bytes / 4); break;
  }
}

}  // namespace tvk
