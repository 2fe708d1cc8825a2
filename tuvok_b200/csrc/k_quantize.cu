// k_quantize.cu -- the value quantiser of the import path on the device (round 2).  Replaces, for data
// that is already in HBM (reference file:line):
//   Quantize<T, U>                 IO/Quantize.h:427-577   range pass (io_minmax :294-343), early return for unsigned data
//                                  that already fits (:462-482), QuantizationFactor (:49-72), map + histogram loop (:522-535)
//   AbstrConverter::Process8Bits   IO/AbstrConverter.cpp:73-157   (signed bytes + 128, histogram of the bytes)
// as RAWConverter's quantize() calls them (IO/RAWConverter.cpp:205-300).  Two streaming passes over the input -- the
// reference makes the same two over a file: (1) min / max, (2) map every value to min(max_out, U((v - min) * factor)) and
// count min(bins - 1, U((v - min) * factor_hist)) in a 256 / 4096-bin histogram.  Both are HBM-bound: 16-byte loads, one
// CTA-private histogram in shared memory flushed once.  The arithmetic is the reference's (the difference in the input
// type's own arithmetic, the product in double, truncating conversion), so results are bit-identical to the CPU oracle
// (oracle/orc_quantize.c, itself pinned to Quantize.h compiled in place and to the reference's known-answer tests).
#include <algorithm>
#include <cfloat>
#include "tvk_dev.h"

namespace tvk {
namespace {

constexpr int kSMs = 148;
constexpr int kQBlocks = kSMs * 8;     // persistent grid: 8 CTAs of 256 threads per SM, grid-stride

template <typename T> struct Lim;
template <> struct Lim<int8_t> { static __device__ int8_t lo() { return -128; } static __device__ int8_t hi() { return 127; } };
template <> struct Lim<uint8_t> { static __device__ uint8_t lo() { return 0; } static __device__ uint8_t hi() { return 255; } };
template <> struct Lim<int16_t> { static __device__ int16_t lo() { return -32768; } static __device__ int16_t hi() { return 32767; } };
template <> struct Lim<uint16_t> { static __device__ uint16_t lo() { return 0; } static __device__ uint16_t hi() { return 65535; } };
template <> struct Lim<int32_t> { static __device__ int32_t lo() { return INT32_MIN; } static __device__ int32_t hi() { return INT32_MAX; } };
template <> struct Lim<uint32_t> { static __device__ uint32_t lo() { return 0u; } static __device__ uint32_t hi() { return UINT32_MAX; } };
template <> struct Lim<float> { static __device__ float lo() { return -FLT_MAX; } static __device__ float hi() { return FLT_MAX; } };
template <> struct Lim<double> { static __device__ double lo() { return -DBL_MAX; } static __device__ double hi() { return DBL_MAX; } };

// values per 16-byte load
template <typename T> struct Pack { static constexpr int n = 16 / sizeof(T); T v[16 / sizeof(T)]; };

template <typename T>
__device__ __forceinline__ Pack<T> load16(const T* p) {
  Pack<T> r;
  *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}

// pass 1: per-CTA minimum / maximum -> part[2 * blockIdx.x] (the host folds the kQBlocks pairs)
template <typename T>
__global__ void __launch_bounds__(256) quant_minmax_kernel(const T* __restrict__ src, uint64_t n, T* part) {
  constexpr int W = Pack<T>::n;
  T mn = Lim<T>::hi(), mx = Lim<T>::lo();
  const uint64_t nv = n / W;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x) {
    const Pack<T> q = load16(src + i * W);
#pragma unroll
    for (int k = 0; k < W; k++) { mn = q.v[k] < mn ? q.v[k] : mn; mx = q.v[k] > mx ? q.v[k] : mx; }
  }
  if (blockIdx.x == 0)
    for (uint64_t i = nv * W + threadIdx.x; i < n; i += blockDim.x) { const T v = src[i]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  for (int o = 16; o > 0; o >>= 1) {
    const T a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx;
  }
  __shared__ T s_mn[8], s_mx[8];
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) { mn = s_mn[w] < mn ? s_mn[w] : mn; mx = s_mx[w] > mx ? s_mx[w] : mx; }
    part[2 * blockIdx.x] = mn; part[2 * blockIdx.x + 1] = mx;
  }
}

// (v - min) as C++ evaluates it for the reference's T, then widened to double
template <typename T> __device__ __forceinline__ double diff(T v, T mn) { return (double)((int)v - (int)mn); }              // 8 / 16-bit: int arithmetic
template <> __device__ __forceinline__ double diff<int32_t>(int32_t v, int32_t mn) { return (double)((int64_t)v - (int64_t)mn); }   // (the reference's int - int may overflow: UB)
template <> __device__ __forceinline__ double diff<uint32_t>(uint32_t v, uint32_t mn) { return (double)(uint32_t)(v - mn); }
template <> __device__ __forceinline__ double diff<float>(float v, float mn) { return (double)(v - mn); }
template <> __device__ __forceinline__ double diff<double>(double v, double mn) { return v - mn; }

template <typename U> __device__ __forceinline__ U to_u(double x) { return (U)__double2int_rz(x); }

struct QuantConsts { double f, fh; uint32_t max_out, bins; int32_t mode; };   // mode 0: map; 1: count the values as they are; 2: signed byte + 128

// pass 2: map + histogram.  DIRECT modes only count (and, for signed bytes, bias) -- Process8Bits and the early return.
template <typename T, typename U>
__global__ void __launch_bounds__(256) quant_map_kernel(const T* __restrict__ src, uint64_t n, T mn, const QuantConsts C, U* dst,
                                                        unsigned long long* hist) {
  extern __shared__ uint32_t s_hist[];
  for (uint32_t b = threadIdx.x; b < C.bins; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  constexpr int W = Pack<T>::n;
  const uint64_t nv = n / W;
  auto one = [&](T v, U& out) -> uint32_t {
    if (C.mode == 1) { out = (U)v; return (uint32_t)v; }
    if (C.mode == 2) { const uint8_t b = (uint8_t)((int)v + 128); out = (U)b; return b; }
    const double d = diff<T>(v, mn);
    const U o = to_u<U>(d * C.f);
    out = o < (U)C.max_out ? o : (U)C.max_out;
    const U h = to_u<U>(d * C.fh);
    return h < (U)(C.bins - 1) ? (uint32_t)h : C.bins - 1;
  };
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nv; i += (uint64_t)gridDim.x * blockDim.x) {
    const Pack<T> q = load16(src + i * W);
    U o[W];
#pragma unroll
    for (int k = 0; k < W; k++) atomicAdd(&s_hist[one(q.v[k], o[k])], 1u);
    if (dst) {
#pragma unroll
      for (int k = 0; k < W; k++) dst[i * W + k] = o[k];      // W consecutive values: the compiler merges them into one store
    }
  }
  if (blockIdx.x == 0)
    for (uint64_t i = nv * W + threadIdx.x; i < n; i += blockDim.x) {
      U o;
      atomicAdd(&s_hist[one(src[i], o)], 1u);
      if (dst) dst[i] = o;
    }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < C.bins; b += blockDim.x)
    if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
}

template <typename T>
int run_minmax(const void* src, uint64_t n, void* part_d, cudaStream_t s) {
  quant_minmax_kernel<T><<<kQBlocks, 256, 0, s>>>((const T*)src, n, (T*)part_d);
  return 0;
}

}  // namespace

int quant_blocks() { return kQBlocks; }

void launch_quant_minmax(const void* src, int type, uint64_t n, void* part_d, cudaStream_t s) {
  switch (type) {
    case TVK_ST_I8: run_minmax<int8_t>(src, n, part_d, s); break;
    case TVK_ST_U8: run_minmax<uint8_t>(src, n, part_d, s); break;
    case TVK_ST_I16: run_minmax<int16_t>(src, n, part_d, s); break;
    case TVK_ST_U16: run_minmax<uint16_t>(src, n, part_d, s); break;
    case TVK_ST_I32: run_minmax<int32_t>(src, n, part_d, s); break;
    case TVK_ST_U32: run_minmax<uint32_t>(src, n, part_d, s); break;
    case TVK_ST_F32: run_minmax<float>(src, n, part_d, s); break;
    default: run_minmax<double>(src, n, part_d, s); break;
  }
}

template <typename T>
static void run_map(const void* src, uint64_t n, double mn, const QuantParams& P, void* dst, unsigned long long* hist, cudaStream_t s) {
  QuantConsts C{P.f, P.fh, P.max_out, P.bins, P.mode};
  const size_t smem = (size_t)P.bins * sizeof(uint32_t);
  if (P.out_bits == 8) quant_map_kernel<T, uint8_t><<<kQBlocks, 256, smem, s>>>((const T*)src, n, (T)mn, C, (uint8_t*)dst, hist);
  else quant_map_kernel<T, uint16_t><<<kQBlocks, 256, smem, s>>>((const T*)src, n, (T)mn, C, (uint16_t*)dst, hist);
}

void launch_quant_map(const void* src, int type, uint64_t n, double mn, const QuantParams& P, void* dst, unsigned long long* hist,
                      cudaStream_t s) {
  switch (type) {
    case TVK_ST_I8: run_map<int8_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_U8: run_map<uint8_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_I16: run_map<int16_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_U16: run_map<uint16_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_I32: run_map<int32_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_U32: run_map<uint32_t>(src, n, mn, P, dst, hist, s); break;
    case TVK_ST_F32: run_map<float>(src, n, mn, P, dst, hist, s); break;
    default: run_map<double>(src, n, mn, P, dst, hist, s); break;
  }
}

}  // namespace tvk
