// tex_probe.cu -- measurement behind the fetch-path choice of the traversal kernel:
// hardware-filtered 3D texture (cudaArray, 16-bit unorm, linear filter: 7 TEX per lit sample) against the product's
// manual filter on linear memory (32 LDG.U16 with immediate offsets + 2^23 conversion + fp32 lerps), on the access
// pattern of the ray caster: 8x4-pixel warp tiles, rays marching at 0.5 voxel through a u16 volume far larger than L2,
// centre tap + 6 central-difference taps per sample (lit / 2D-TF modes) or the centre tap alone (1D TF).
// Also reports how far the texture unit's fixed-point filter weights put the filtered value from the fp32 lerp
// (the arithmetic contract of the parity tests).  Developer tool, not part of the product library.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tex_probe tex_probe.cu && ./tex_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NX = 1024, NY = 1024, NZ = 768;      // 1.5 GiB of u16: >> 126 MB L2
constexpr int W = 1920, H = 1080, STEPS = 384;

__device__ __forceinline__ float cvt(uint16_t v) { return __uint_as_float(0x4b000000u | v) - 8388608.0f; }
__device__ __forceinline__ float lerp1(float a, float b, float t) { return fmaf(t, b - a, a); }
__device__ __forceinline__ float tri(float a, float b, float c, float d, float e, float f, float g, float h, float fx, float fy, float fz) {
  return lerp1(lerp1(lerp1(a, b, fx), lerp1(c, d, fx), fy), lerp1(lerp1(e, f, fx), lerp1(g, h, fx), fy), fz);
}

struct Ray { float ox, oy, oz, dx, dy, dz; };
__device__ __forceinline__ Ray make_ray(int px, int py) {
  Ray r;   // slightly diverging bundle entering the z = 2 face, 0.5-voxel steps
  r.ox = 40.0f + (float)px * 0.48f; r.oy = 30.0f + (float)py * 0.85f; r.oz = 2.25f;
  const float ax = ((float)px / W - 0.5f) * 0.12f, ay = ((float)py / H - 0.5f) * 0.08f;
  const float inv = 0.5f * rsqrtf(ax * ax + ay * ay + 1.0f);
  r.dx = ax * inv; r.dy = ay * inv; r.dz = inv;
  return r;
}

template <bool GRAD>
__global__ void __launch_bounds__(64, 8) k_tex(cudaTextureObject_t tex, float* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int px = blockIdx.x * 8 + (lane & 7), py = blockIdx.y * 8 + wid * 4 + (lane >> 3);
  if (px >= W || py >= H) return;
  const Ray r = make_ray(px, py);
  float acc = 0.0f;
#pragma unroll 1
  for (int s = 0; s < STEPS; s++) {
    const float x = r.ox + s * r.dx, y = r.oy + s * r.dy, z = r.oz + s * r.dz;   // texel space (+0.5 = texel centre)
    const float c = tex3D<float>(tex, x, y, z);
    if (GRAD) {
      const float gx = tex3D<float>(tex, x - 1.0f, y, z) - tex3D<float>(tex, x + 1.0f, y, z);
      const float gy = tex3D<float>(tex, x, y - 1.0f, z) - tex3D<float>(tex, x, y + 1.0f, z);
      const float gz = tex3D<float>(tex, x, y, z - 1.0f) - tex3D<float>(tex, x, y, z + 1.0f);
      acc = fmaf(c, gx * gx + gy * gy + gz * gz, acc);
    } else {
      acc += c;
    }
  }
  out[(size_t)py * W + px] = acc;
}

template <bool GRAD>
__global__ void __launch_bounds__(64, 8) k_ldg(const uint16_t* __restrict__ vol, float* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int px = blockIdx.x * 8 + (lane & 7), py = blockIdx.y * 8 + wid * 4 + (lane >> 3);
  if (px >= W || py >= H) return;
  const Ray r = make_ray(px, py);
  const float n = 1.0f / 65535.0f;
  float acc = 0.0f;
#pragma unroll 1
  for (int s = 0; s < STEPS; s++) {
    const float ux = r.ox + s * r.dx - 0.5f, uy = r.oy + s * r.dy - 0.5f, uz = r.oz + s * r.dz - 0.5f;
    const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    const float fx = ux - x0, fy = uy - y0, fz = uz - z0;
    const uint16_t* c = vol + ((size_t)(int)z0 * NY + (int)y0) * NX + (int)x0;
#define V(i, j, k) cvt(__ldg(c + ((i) + (j) * NX + (k) * NX * NY)))
    const float c000 = V(0, 0, 0), c100 = V(1, 0, 0), c010 = V(0, 1, 0), c110 = V(1, 1, 0);
    const float c001 = V(0, 0, 1), c101 = V(1, 0, 1), c011 = V(0, 1, 1), c111 = V(1, 1, 1);
    const float v = tri(c000, c100, c010, c110, c001, c101, c011, c111, fx, fy, fz) * n;
    if (GRAD) {   // 24 more voxels: the 6 shifted footprints share the centre block
      const float xl00 = V(-1, 0, 0), xl10 = V(-1, 1, 0), xl01 = V(-1, 0, 1), xl11 = V(-1, 1, 1);
      const float xh00 = V(2, 0, 0), xh10 = V(2, 1, 0), xh01 = V(2, 0, 1), xh11 = V(2, 1, 1);
      const float yl00 = V(0, -1, 0), yl10 = V(1, -1, 0), yl01 = V(0, -1, 1), yl11 = V(1, -1, 1);
      const float yh00 = V(0, 2, 0), yh10 = V(1, 2, 0), yh01 = V(0, 2, 1), yh11 = V(1, 2, 1);
      const float zl00 = V(0, 0, -1), zl10 = V(1, 0, -1), zl01 = V(0, 1, -1), zl11 = V(1, 1, -1);
      const float zh00 = V(0, 0, 2), zh10 = V(1, 0, 2), zh01 = V(0, 1, 2), zh11 = V(1, 1, 2);
      const float xm = tri(xl00, c000, xl10, c010, xl01, c001, xl11, c011, fx, fy, fz) * n;
      const float xp = tri(c100, xh00, c110, xh10, c101, xh01, c111, xh11, fx, fy, fz) * n;
      const float ym = tri(yl00, yl10, c000, c100, yl01, yl11, c001, c101, fx, fy, fz) * n;
      const float yp = tri(c010, c110, yh00, yh10, c011, c111, yh01, yh11, fx, fy, fz) * n;
      const float zm = tri(zl00, zl10, zl01, zl11, c000, c100, c010, c110, fx, fy, fz) * n;
      const float zp = tri(c001, c101, c011, c111, zh00, zh10, zh01, zh11, fx, fy, fz) * n;
      const float gx = xm - xp, gy = ym - yp, gz = zm - zp;
      acc = fmaf(v, gx * gx + gy * gy + gz * gz, acc);
    } else {
      acc += v;
    }
#undef V
  }
  out[(size_t)py * W + px] = acc;
}

// filter accuracy: |tex3D - fp32 lerp| over random positions
__global__ void k_acc(cudaTextureObject_t tex, const uint16_t* __restrict__ vol, float* maxdiff, float* sumdiff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = 0x9E3779B9u * (i + 1);
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
  const float x = 2.0f + rnd() * (NX - 4), y = 2.0f + rnd() * (NY - 4), z = 2.0f + rnd() * (NZ - 4);
  const float t = tex3D<float>(tex, x, y, z);
  const float ux = x - 0.5f, uy = y - 0.5f, uz = z - 0.5f;
  const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
  const uint16_t* c = vol + ((size_t)(int)z0 * NY + (int)y0) * NX + (int)x0;
#define V(i, j, k) cvt(c[(i) + (j) * NX + (k) * NX * NY])
  const float v = tri(V(0, 0, 0), V(1, 0, 0), V(0, 1, 0), V(1, 1, 0), V(0, 0, 1), V(1, 0, 1), V(0, 1, 1), V(1, 1, 1), ux - x0, uy - y0, uz - z0) * (1.0f / 65535.0f);
#undef V
  const float d = fabsf(t - v);
  atomicMax((int*)maxdiff, __float_as_int(d));
  atomicAdd(sumdiff, d);
}

__global__ void k_fill(uint16_t* vol, size_t n, int smooth) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % NX), y = (int)((i / NX) % NY), z = (int)(i / ((size_t)NX * NY));
    uint32_t h = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u;
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    if (smooth) vol[i] = (uint16_t)(32768.0f + 20000.0f * __sinf(x * 0.05f) * __cosf(y * 0.04f) * __sinf(z * 0.06f) + (h & 255));
    else vol[i] = (uint16_t)(h & 0xffff);
  }
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; i++) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const size_t n = (size_t)NX * NY * NZ;
  uint16_t* vol; CK(cudaMalloc(&vol, n * 2));
  float* out; CK(cudaMalloc(&out, (size_t)W * H * 4));
  float* diff; CK(cudaMalloc(&diff, 8));
  cudaArray_t arr;
  cudaChannelFormatDesc fd = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(NX, NY, NZ)));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  const dim3 block(64), grid((W + 7) / 8, (H + 7) / 8);
  const double samples = (double)W * H * STEPS;
  for (int smooth = 1; smooth >= 0; smooth--) {
    k_fill<<<148 * 8, 256>>>(vol, n, smooth);
    CK(cudaDeviceSynchronize());
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(vol, NX * 2, NX, NY); cp.dstArray = arr; cp.extent = make_cudaExtent(NX, NY, NZ);
    cp.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cp));
    CK(cudaMemset(diff, 0, 8));
    k_acc<<<4096, 256>>>(tex, vol, diff, diff + 1);
    float hd[2]; CK(cudaMemcpy(hd, diff, 8, cudaMemcpyDeviceToHost));
    printf("volume %s: |tex3D - fp32 lerp| max %.3g (%.2f / 65535), mean %.3g  [1/255 = %.3g]\n", smooth ? "smooth" : "noise", hd[0], hd[0] * 65535.0f,
           hd[1] / (4096.0f * 256.0f), 1.0 / 255.0);
    const float t1 = time_ms([&] { k_tex<false><<<grid, block>>>(tex, out); }, 5);
    const float l1 = time_ms([&] { k_ldg<false><<<grid, block>>>(vol, out); }, 5);
    const float t7 = time_ms([&] { k_tex<true><<<grid, block>>>(tex, out); }, 5);
    const float l7 = time_ms([&] { k_ldg<true><<<grid, block>>>(vol, out); }, 5);
    CK(cudaGetLastError());
    printf("  1 tap / sample : texture %7.3f ms = %6.1f Gsamples/s | manual (8 LDG)  %7.3f ms = %6.1f Gsamples/s | tex/manual speed %.2fx\n",
           t1, samples / t1 * 1e-6, l1, samples / l1 * 1e-6, l1 / t1);
    printf("  7 taps / sample: texture %7.3f ms = %6.1f Gsamples/s (%6.1f Gtaps/s) | manual (32 LDG) %7.3f ms = %6.1f Gsamples/s | tex/manual speed %.2fx\n",
           t7, samples / t7 * 1e-6, 7 * samples / t7 * 1e-6, l7, samples / l7 * 1e-6, l7 / t7);
  }
  return 0;
}
