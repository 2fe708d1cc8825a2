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

// ---- packed fp32 (sm_100: add / mul / fma .f32x2 = FADD2 / FMUL2 / FFMA2, one issue slot for two IEEE operations).
// Every lane rounds exactly like the scalar instruction (round to nearest even, no flush), so a computation written
// with these is bit-identical to its scalar form; the arithmetic contract is untouched.  ptxas keeps an f2 in an aligned
// register pair; the mov.b64 packs below cost nothing.
struct f2 { float x, y; };
__device__ __forceinline__ f2 F2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ unsigned long long pk2(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 upk2(unsigned long long v) {
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(r);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(r);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(r);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return upk2(r);
}
// two lerps a + f * (b - a) with the same fraction (the lerp of tri() below, lane by lane)
__device__ __forceinline__ f2 lerp2(f2 a, f2 b, float f) { return fma2(F2(f, f), sub2(b, a), a); }

// RGBA8 texel -> the float4 a GL_RGBA8 texture fetch returns: byte / 255.0f as an IEEE division.  q = b * (1/255),
// then one exact residual step r = fma(-q, 255, b), q' = fma(r, 1/255, q): equal to the correctly rounded quotient for
// every byte value (all 256 checked, tests/test_host_logic.py) at 2.5 packed instructions per channel instead of a
// division sequence, so the tables stay 4 bytes per entry (a 4096 x 256 2D table: 4 MB instead of 16.8 MB as float4).
__device__ __forceinline__ f4 unorm8x4(uint32_t q) {
  const uint32_t magic = 0x4B000000u;                     // 2^23: byte b in the low mantissa bits = 2^23 + b
  f2 rg = F2(__uint_as_float(__byte_perm(q, magic, 0x7540)), __uint_as_float(__byte_perm(q, magic, 0x7541)));
  f2 ba = F2(__uint_as_float(__byte_perm(q, magic, 0x7542)), __uint_as_float(__byte_perm(q, magic, 0x7543)));
  const f2 unbias = F2(-8388608.0f, -8388608.0f), inv = F2(1.0f / 255.0f, 1.0f / 255.0f), m255 = F2(-255.0f, -255.0f);
  rg = add2(rg, unbias); ba = add2(ba, unbias);
  const f2 q0 = mul2(rg, inv), q1 = mul2(ba, inv);
  const f2 r0 = fma2(q0, m255, rg), r1 = fma2(q1, m255, ba);
  const f2 a = fma2(r0, inv, q0), b = fma2(r1, inv, q1);
  f4 o; o.x = a.x; o.y = a.y; o.z = b.x; o.w = b.y;
  return o;
}

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

// ---- x-pair pool layout (k_pool.cu page_copy_kernel): pool element x = (voxel x, voxel x+1)
// W = the stored pool element, V = the pair a footprint row is built from.  8 / 16-bit volumes store the pair (one load
// per pair); float volumes stay plain (a stored pair would be 8 bytes per voxel and halve what L1 holds: measured
// slower on the 1024^3 float isosurface workload) and form the pair from two loads.
template <typename T> struct PairOf;
template <> struct PairOf<uint8_t> { typedef uint16_t W; typedef uint16_t V; };
template <> struct PairOf<uint16_t> { typedef uint32_t W; typedef uint32_t V; };
template <> struct PairOf<float> { typedef float W; typedef float2 V; };
__device__ __forceinline__ uint16_t load_pair(const uint16_t* p) { return __ldg(p); }
__device__ __forceinline__ uint32_t load_pair(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ float2 load_pair(const float* p) { return make_float2(__ldg(p), __ldg(p + 1)); }
// The two voxels of a pair as floats that still carry the 2^23 conversion bias (integer types): one PRMT each.
// The bias cancels exactly in differences of two such values (both are integers below 2^24) and is removed from the low
// operand of a lerp with one (packed) add, so a voxel costs 1 - 1.5 instructions to convert instead of 2.
template <typename T> struct PairCvt;
template <> struct PairCvt<uint8_t> {
  static constexpr bool kBiased = true;
  static __device__ __forceinline__ float lo(uint16_t w) { return __uint_as_float(__byte_perm((uint32_t)w, 0x4B000000u, 0x7540)); }
  static __device__ __forceinline__ float hi(uint16_t w) { return __uint_as_float(__byte_perm((uint32_t)w, 0x4B000000u, 0x7541)); }
};
template <> struct PairCvt<uint16_t> {
  static constexpr bool kBiased = true;
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }
};
template <> struct PairCvt<float> {
  static constexpr bool kBiased = false;
  static __device__ __forceinline__ float lo(float2 w) { return w.x; }
  static __device__ __forceinline__ float hi(float2 w) { return w.y; }
};
// first voxel of a pair, fully converted (what cvt(plain voxel) returns)
__device__ __forceinline__ float first_voxel_of(uint16_t w) { return PairCvt<uint8_t>::lo(w) - 8388608.0f; }
__device__ __forceinline__ float first_voxel_of(uint32_t w) { return PairCvt<uint16_t>::lo(w) - 8388608.0f; }
__device__ __forceinline__ float first_voxel_of(float w) { return w; }
template <typename T>
__device__ __forceinline__ float first_voxel(typename PairOf<T>::W w) { return first_voxel_of(w); }
// two x-lerps at once: lanes (a.x -> b.x) and (a.y -> b.y), operands still biased (see PairCvt)
template <bool BIASED>
__device__ __forceinline__ f2 xlerp2(f2 a, f2 b, float f) {
  const f2 d = sub2(b, a);
  const f2 a0 = BIASED ? add2(a, F2(-8388608.0f, -8388608.0f)) : a;
  return fma2(F2(f, f), d, a0);
}
__device__ __forceinline__ float lerp1(float a, float b, float f) { return fmaf(f, b - a, a); }

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

// lighting() in two steps: the geometric terms do not depend on the classified colour, so a kernel computes them while
// the transfer-function fetch is still in flight (same operations, same order per value: bit-identical)
__device__ __forceinline__ void light_terms(f3 eye, f3 pos, f3 n, f3 ldir, float& dl, float& sp) {
  f3 view = norm3(sub3(eye, pos));
  float dn = dot3(n, view);
  f3 refl = norm3(sub3(view, scl3(n, 2.0f * dn)));
  dl = fmaxf(fabsf(dot3(n, ldir)), 0.0f);
  sp = pow8(fmaxf(dot3(refl, ldir), 0.0f));
}
__device__ __forceinline__ f3 light_apply(f3 amb, f3 dif, f3 spe, float dl, float sp) {
  return F3(clampf(amb.x + dif.x * dl + spe.x * sp, 0.0f, 1.0f),
            clampf(amb.y + dif.y * dl + spe.y * sp, 0.0f, 1.0f),
            clampf(amb.z + dif.z * dl + spe.z * sp, 0.0f, 1.0f));
}

#endif
