// tvk_math.cuh -- device helpers shared by the traversal kernels (k_raycast.cu, k_classic.cu): the vector
// arithmetic of the arithmetic contract (IEEE fp32, explicit fmaf only in lerps / dot products),
// voxel conversion, trilinear filter, lighting.  Include inside `namespace tvk { namespace {`.
#ifndef TVK_MATH_CUH
#define TVK_MATH_CUH

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
__device__ __forceinline__ f3 F3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 F3(const float* p) { return F3(p[0], p[1], p[2]); }
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 mul3(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 div3(f3 a, f3 b) { return F3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ f3 scl3(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float len3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ f3 norm3(f3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return scl3(a, inv); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// v' = v * M (row vectors, Basics/Vectors.h:434-439)
__device__ __forceinline__ f4 xform4(const float* m, float x, float y, float z, float w) {
  f4 r;
  r.x = x * m[0] + y * m[4] + z * m[8] + w * m[12];
  r.y = x * m[1] + y * m[5] + z * m[9] + w * m[13];
  r.z = x * m[2] + y * m[6] + z * m[10] + w * m[14];
  r.w = x * m[3] + y * m[7] + z * m[11] + w * m[15];
  return r;
}

// voxel -> float.  Integer voxels are converted with the 2^23 magic number (exact below 2^23) on the
// FMA/ALU pipes instead of the quarter-rate I2F conversion pipe.
__device__ __forceinline__ float cvt(uint8_t v) { return __uint_as_float(0x4B000000u | (uint32_t)v) - 8388608.0f; }
__device__ __forceinline__ float cvt(uint16_t v) { return __uint_as_float(0x4B000000u | (uint32_t)v) - 8388608.0f; }
__device__ __forceinline__ float cvt(float v) { return v; }

__device__ __forceinline__ float tri(float v000, float v100, float v010, float v110, float v001, float v101,
                                     float v011, float v111, float fx, float fy, float fz) {
  float c00 = fmaf(fx, v100 - v000, v000);
  float c10 = fmaf(fx, v110 - v010, v010);
  float c01 = fmaf(fx, v101 - v001, v001);
  float c11 = fmaf(fx, v111 - v011, v011);
  float c0 = fmaf(fy, c10 - c00, c00);
  float c1 = fmaf(fy, c11 - c01, c01);
  return fmaf(fz, c1 - c0, c0);
}

__device__ __forceinline__ float pow8(float x) { float a = x * x; float b = a * a; return b * b; }

// lighting.glsl:33-43
__device__ __forceinline__ f3 lighting(f3 eye, f3 pos, f3 n, f3 amb, f3 dif, f3 spe, f3 ldir) {
  f3 view = norm3(sub3(eye, pos));
  float dn = dot3(n, view);
  f3 refl = norm3(sub3(view, scl3(n, 2.0f * dn)));
  float dl = fmaxf(fabsf(dot3(n, ldir)), 0.0f);
  float sp = pow8(fmaxf(dot3(refl, ldir), 0.0f));
  return F3(clampf(amb.x + dif.x * dl + spe.x * sp, 0.0f, 1.0f),
            clampf(amb.y + dif.y * dl + spe.y * sp, 0.0f, 1.0f),
            clampf(amb.z + dif.z * dl + spe.z * sp, 0.0f, 1.0f));
}

#endif
