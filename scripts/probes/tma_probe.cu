// Developer probe: which cp.async.bulk.tensor.3d box loads does the TMA unit of this GPU accept?  (scripts/gpu_r2o.sh)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool GRID_CONST>
__global__ void k(const __grid_constant__ CUtensorMap tmv, const CUtensorMap* tmp, int x, int y, int z, uint32_t bytes, uint16_t* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  const CUtensorMap* tm = GRID_CONST ? &tmv : tmp;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(sm)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  } while (!ok);
  for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = ((const uint16_t*)sm)[i];
}
int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int N = 96;
  std::vector<uint16_t> h((size_t)N * N * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint16_t)(i % 65000);
  uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 1 << 20);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap* ring; cudaMalloc(&ring, sizeof(CUtensorMap));
  struct Case { int bx, by, bz, x, y, z; };
  const Case cases[] = {{40, 36, 6, 0, 0, 0}, {40, 36, 6, 30, 30, 30}, {40, 36, 6, 32, 30, 30}, {40, 36, 6, -8, 0, 0}, {40, 36, 6, 0, -2, 0}, {40, 36, 6, 0, 0, -2}, {40, 36, 6, -2, 0, 0},
                        {40, 36, 6, 64, 62, 94}, {40, 36, 6, 62, 62, 94}, {40, 36, 6, 1, 0, 0}, {40, 36, 6, 4, 0, 0}, {40, 36, 6, 8, 0, 0}, {48, 36, 6, -8, -2, -2}};
  int idx = -1;
  for (const Case& c : cases) {
    idx++;
    if (only >= 0 && idx != only) continue;
    for (int gc = 0; gc < 1; gc++) {
      CUtensorMap tm;
      const cuuint64_t dims[3] = {N, N, N}, strides[2] = {N * 2, (cuuint64_t)N * N * 2};
      const cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bz}, es[3] = {1, 1, 1};
      CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("box %d,%d,%d encode failed %d\n", c.bx, c.by, c.bz, (int)r); break; }
      cudaMemcpy(ring, &tm, sizeof(tm), cudaMemcpyHostToDevice);
      const uint32_t bytes = c.bx * c.by * c.bz * 2;
      cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
      cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
      if (gc) k<true><<<1, 128, bytes>>>(tm, ring, c.x, c.y, c.z, bytes, o); else k<false><<<1, 128, bytes>>>(tm, ring, c.x, c.y, c.z, bytes, o);
      cudaError_t e = cudaDeviceSynchronize();
      uint16_t first[4] = {0, 0, 0, 0};
      if (e == cudaSuccess) cudaMemcpy(first, o + (size_t)c.bx * (c.by * 2 + 2) + 2, 8, cudaMemcpyDeviceToHost);
      printf("box %d,%d,%d at %d,%d,%d desc=%s: %s  sample %u %u (expect %u at +2,+2,+2)\n", c.bx, c.by, c.bz, c.x, c.y, c.z, gc ? "grid_constant" : "global",
             cudaGetErrorString(e), first[0], first[1], (unsigned)(((size_t)(c.z + 2) * N * N + (size_t)(c.y + 2) * N + (c.x + 2)) % 65000));
      if (e != cudaSuccess) { printf("context lost, stopping\n"); return 1; }
    }
  }
  return 0;
}
