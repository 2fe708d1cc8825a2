// glsl_emu.h -- TEST INFRASTRUCTURE: a minimal GLSL 4.20 emulation in C++ so that the reference's OWN shader text
// (Shaders/GLGridLeaper-*.glsl, lighting.glsl, Compositing.glsl and the GLSL that GLVolumePool / GLHashTable generate)
// can be compiled with g++ and EXECUTED on the CPU, one fragment at a time.  The shader text is not copied into this
// repository: tests/glsl_ref.py reads it from /root/reference at test time, applies a purely syntactic rewrite
// (qualifiers `in/out/inout/uniform/layout`, array constructors, `main`) and includes it after this header.
//
// This header IS one OpenGL implementation: where GLSL leaves precision implementation-defined it takes the choices of
// the arithmetic contract in DESIGN.md section 4 (IEEE fp32, no contraction: compile with -ffp-contract=off;
// dot = fma chain; normalize(v) = v * (1 / sqrt(dot)); pow(x, 8) by squaring; log2 exact for the integer part;
// GL_LINEAR filtering of integer texels in fp32 with fma lerps, scaled once; GL_NEAREST RGBA8 transfer functions).
// Covers exactly the constructs those shaders use (prefix swizzles .xy/.xyz/.rgb, vec/ivec/uvec 2-4, mat4x4, arrays).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec2; struct uvec3; struct uvec4;

// ---- prefix swizzle proxies (the member overlays the first N components of its parent) -----------------
template <typename V, typename S, int N> struct swz {
  S v[N];
  operator V() const { V r; for (int i = 0; i < N; i++) (&r.x)[i] = v[i]; return r; }
  swz& operator=(const V& o) { for (int i = 0; i < N; i++) v[i] = (&o.x)[i]; return *this; }
  swz& operator+=(const V& o) { for (int i = 0; i < N; i++) v[i] = v[i] + (&o.x)[i]; return *this; }
  swz& operator-=(const V& o) { for (int i = 0; i < N; i++) v[i] = v[i] - (&o.x)[i]; return *this; }
  swz& operator*=(const V& o) { for (int i = 0; i < N; i++) v[i] = v[i] * (&o.x)[i]; return *this; }
  swz& operator-=(S o) { for (int i = 0; i < N; i++) v[i] = v[i] - o; return *this; }   // vector -= scalar
};

struct vec2 {
  float x, y;
  vec2() : x(0), y(0) {}
  vec2(float a, float b) : x(a), y(b) {}
  float& operator[](int i) { return (&x)[i]; }
};
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {} };
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };

struct vec3 {
  union { struct { float x, y, z; }; struct { float r, g, b; }; swz<vec2, float, 2> xy; swz<vec3, float, 3> xyz; swz<vec3, float, 3> rgb; };
  vec3() : x(0), y(0), z(0) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
  explicit vec3(float a) : x(a), y(a), z(a) {}
  explicit vec3(const ivec3& v);
  explicit vec3(const uvec3& v);
  float& operator[](int i) { return (&x)[i]; }
  float operator[](int i) const { return (&x)[i]; }
  vec3& operator+=(const vec3& o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
  vec3& operator-=(const vec3& o) { x = x - o.x; y = y - o.y; z = z - o.z; return *this; }
  vec3& operator*=(const vec3& o) { x = x * o.x; y = y * o.y; z = z * o.z; return *this; }
  vec3& operator/=(const vec3& o) { x = x / o.x; y = y / o.y; z = z / o.z; return *this; }
  vec3& operator/=(float s) { x = x / s; y = y / s; z = z / s; return *this; }
  vec3& operator*=(float s) { x = x * s; y = y * s; z = z * s; return *this; }
};
struct ivec3 {
  int x, y, z;
  ivec3() : x(0), y(0), z(0) {}
  ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
  ivec3(uint a, uint b, uint c) : x((int)a), y((int)b), z((int)c) {}
  explicit ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
  int operator[](int i) const { return (&x)[i]; }
};
struct uvec3 {
  uint x, y, z;
  uvec3() : x(0), y(0), z(0) {}
  uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
  uvec3(int a, int b, int c) : x((uint)a), y((uint)b), z((uint)c) {}
  uint operator[](int i) const { return (&x)[i]; }
};
inline vec3::vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline vec3::vec3(const uvec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}

struct swz_xw { float v[4]; operator vec2() const { return vec2(v[0], v[3]); } };   // read-only .xw (Transfer-MIP-FS.glsl)
struct vec4 {
  union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; swz<vec3, float, 3> xyz; swz<vec3, float, 3> rgb; swz<vec2, float, 2> xy; swz_xw xw; };
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
  vec4(const vec3& v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
  explicit vec4(float a_) : x(a_), y(a_), z(a_), w(a_) {}
  vec4(const vec3& v, int d_) : x(v.x), y(v.y), z(v.z), w((float)d_) {}
  vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
struct uvec4 {
  union { struct { uint x, y, z, w; }; struct { uint r, g, b, a; }; swz<uvec3, uint, 3> xyz; };
  uvec4() : x(0), y(0), z(0), w(0) {}
  uvec4(uint a_, uint b_, uint c_, uint d_) : x(a_), y(b_), z(c_), w(d_) {}
  uvec4(int a_, int b_, int c_, int d_) : x((uint)a_), y((uint)b_), z((uint)c_), w((uint)d_) {}
  uvec4(const vec3& v, uint d_) : x((uint)v.x), y((uint)v.y), z((uint)v.z), w(d_) {}   // float -> uint: truncation
  uvec4(const uvec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  uvec4& operator=(const uvec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
inline bool operator!=(const uvec4& a, const uvec4& b) { return a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w; }
inline bool operator==(const uvec4& a, const uvec4& b) { return !(a != b); }

// ---- arithmetic (component-wise, IEEE, no contraction) ---------------------------------------------------
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline uvec3 operator+(const uvec3& a, int s) { return uvec3(a.x + (uint)s, a.y + (uint)s, a.z + (uint)s); }
inline uvec3 operator*(const uvec3& a, const ivec3& b) { return uvec3(a.x * (uint)b.x, a.y * (uint)b.y, a.z * (uint)b.z); }
inline uvec3 operator+(const uvec3& a, const ivec3& b) { return uvec3(a.x + (uint)b.x, a.y + (uint)b.y, a.z + (uint)b.z); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

// mat4x4 as uploaded by GLSLProgram::Set(name, float[16], 4, false): GL reads the 16 floats column-major, Tuvok stores
// row vectors, so (M * v).x = a[0]*v.x + a[4]*v.y + a[8]*v.z + a[12]*v.w (summed left to right)
struct mat4x4 { float a[16]; };
typedef mat4x4 mat4;
inline vec4 operator*(const mat4x4& m, const vec4& v) {
  const float* a = m.a;
  return vec4(a[0] * v.x + a[4] * v.y + a[8] * v.z + a[12] * v.w, a[1] * v.x + a[5] * v.y + a[9] * v.z + a[13] * v.w,
              a[2] * v.x + a[6] * v.y + a[10] * v.z + a[14] * v.w, a[3] * v.x + a[7] * v.y + a[11] * v.z + a[15] * v.w);
}

// compatibility-profile matrices: gl_NormalMatrix (mat3, rows filled by the driver) and gl_TextureMatrix[] (as mat4x4)
struct mat3 { float a[9]; };
inline vec3 operator*(const mat3& m, const vec3& v) {
  return vec3(m.a[0] * v.x + m.a[1] * v.y + m.a[2] * v.z, m.a[3] * v.x + m.a[4] * v.y + m.a[5] * v.z,
              m.a[6] * v.x + m.a[7] * v.y + m.a[8] * v.z);
}

// ---- built-ins -------------------------------------------------------------------------------------------------
inline float dot(const vec3& a, const vec3& b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { const float inv = 1.0f / sqrtf(dot(a, a)); return vec3(a.x * inv, a.y * inv, a.z * inv); }
inline vec3 reflect(const vec3& i, const vec3& n) { const float d = dot(n, i); return i - n * (2.0f * d); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec4 clamp(const vec4& v, float lo, float hi) { return vec4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
inline vec3 step(float edge, const vec3& v) { return vec3(v.x < edge ? 0.0f : 1.0f, v.y < edge ? 0.0f : 1.0f, v.z < edge ? 0.0f : 1.0f); }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint min(int a, uint b) { return (uint)a < b ? (uint)a : b; }
inline float abs(float a) { return fabsf(a); }
// component-wise forms used by Compose-CV-FS.glsl
inline vec3 abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
inline vec4 max(const vec4& a, float b) { return vec4(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b), fmaxf(a.w, b)); }
inline vec4 min(const vec4& a, float b) { return vec4(fminf(a.x, b), fminf(a.y, b), fminf(a.z, b), fminf(a.w, b)); }
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline float ceil(float a) { return ceilf(a); }
inline float floor(float a) { return floorf(a); }
inline float pow(float x, float e) {
  if (e == 8.0f) { const float a = x * x, b = a * a; return b * b; }
  if (e == 1.0f) return x;
  return powf(x, e);
}
// log2 is only used as uint(log2(x)): the integer part must be exact, so it is taken from the exponent; x < 1 (and
// non-finite garbage) maps to a value whose uint() is 0, +inf to a huge level that min(iMaxLOD, .) cuts
inline float log2(float x) {
  if (!(x >= 1.0f)) return 0.0f;
  if (std::isinf(x)) return 1e9f;
  uint32_t u; std::memcpy(&u, &x, 4);
  return (float)(int)(((u >> 23) & 0xffu) - 127u);
}

// ---- textures ------------------------------------------------------------------------------------------------
struct sampler1D { const uint8_t* rgba8 = nullptr; int w = 0; };
struct sampler2D { const uint8_t* rgba8 = nullptr; int w = 0, h = 0; const float* f32 = nullptr; };   // rgba8 TF or RGBA32F FBO
struct usampler3D { const uint32_t* d = nullptr; int w = 0, h = 0, z = 0; };
struct sampler3D { const void* d = nullptr; int w = 0, h = 0, z = 0, dtype = 0; float norm = 1.0f; bool nearest = false; };

inline vec4 rgba8_texel(const uint8_t* q) { return vec4((float)q[0] / 255.0f, (float)q[1] / 255.0f, (float)q[2] / 255.0f, (float)q[3] / 255.0f); }
// GL_NEAREST, clamp-to-edge
inline vec4 texture(const sampler1D& s, float c) {
  int i = (int)floorf(c * (float)s.w);
  i = i < 0 ? 0 : i >= s.w ? s.w - 1 : i;
  return rgba8_texel(s.rgba8 + 4 * (size_t)i);
}
inline vec4 texture(const sampler2D& s, const vec2& c) {
  int i = (int)floorf(c.x * (float)s.w), j = (int)floorf(c.y * (float)s.h);
  i = i < 0 ? 0 : i >= s.w ? s.w - 1 : i;
  j = j < 0 ? 0 : j >= s.h ? s.h - 1 : j;
  return rgba8_texel(s.rgba8 + 4 * ((size_t)j * s.w + i));
}
// compatibility-profile fetch of an RGBA32F render target (GL_NEAREST): used by Compose-FS.glsl
inline vec4 texture2D(const sampler2D& s, const vec2& c) {
  int i = (int)floorf(c.x * (float)s.w), j = (int)floorf(c.y * (float)s.h);
  i = i < 0 ? 0 : i >= s.w ? s.w - 1 : i;
  j = j < 0 ? 0 : j >= s.h ? s.h - 1 : j;
  if (!s.f32) return rgba8_texel(s.rgba8 + 4 * ((size_t)j * s.w + i));      // RGBA8 2D transfer function
  const float* q = s.f32 + 4 * ((size_t)j * s.w + i);
  return vec4(q[0], q[1], q[2], q[3]);
}
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int) {
  const float* q = s.f32 + 4 * ((size_t)p.y * s.w + p.x);
  return vec4(q[0], q[1], q[2], q[3]);
}
inline uvec4 texelFetch(const usampler3D& s, const ivec3& p, int) {
  const uint32_t v = s.d[(size_t)p.x + (size_t)s.w * ((size_t)p.y + (size_t)s.h * (size_t)p.z)];
  return uvec4(v, 0u, 0u, 1u);
}
inline float pool_texel(const sampler3D& s, int x, int y, int z, int ch = 0) {
  x = x < 0 ? 0 : x >= s.w ? s.w - 1 : x;
  y = y < 0 ? 0 : y >= s.h ? s.h - 1 : y;
  z = z < 0 ? 0 : z >= s.z ? s.z - 1 : z;
  const size_t i = (size_t)x + (size_t)s.w * ((size_t)y + (size_t)s.h * (size_t)z);
  if (s.dtype == 3) return (float)((const uint8_t*)s.d)[4 * i + (size_t)ch];   // GL_RGBA8 pool of a colour volume
  return s.dtype == 0 ? (float)((const uint8_t*)s.d)[i] : s.dtype == 1 ? (float)((const uint16_t*)s.d)[i] : ((const float*)s.d)[i];
}
// one channel of the filtered texel (the arithmetic of texture() below)
inline float pool_filter(const sampler3D& s, const vec3& c, int ch) {
  if (s.nearest)
    return pool_texel(s, (int)floorf(c.x * (float)s.w), (int)floorf(c.y * (float)s.h), (int)floorf(c.z * (float)s.z), ch) * s.norm;
  const float ux = fmaf(c.x, (float)s.w, -0.5f), uy = fmaf(c.y, (float)s.h, -0.5f), uz = fmaf(c.z, (float)s.z, -0.5f);
  const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
  const float fx = ux - x0, fy = uy - y0, fz = uz - z0;
  const int x = (int)x0, y = (int)y0, z = (int)z0;
  const float v000 = pool_texel(s, x, y, z, ch), v100 = pool_texel(s, x + 1, y, z, ch), v010 = pool_texel(s, x, y + 1, z, ch),
              v110 = pool_texel(s, x + 1, y + 1, z, ch), v001 = pool_texel(s, x, y, z + 1, ch), v101 = pool_texel(s, x + 1, y, z + 1, ch),
              v011 = pool_texel(s, x, y + 1, z + 1, ch), v111 = pool_texel(s, x + 1, y + 1, z + 1, ch);
  const float c00 = fmaf(fx, v100 - v000, v000), c10 = fmaf(fx, v110 - v010, v010);
  const float c01 = fmaf(fx, v101 - v001, v001), c11 = fmaf(fx, v111 - v011, v011);
  const float c0 = fmaf(fy, c10 - c00, c00), c1 = fmaf(fy, c11 - c01, c01);
  return fmaf(fz, c1 - c0, c0) * s.norm;
}
// GL_LUMINANCE8/16/32F, GL_LINEAR (or GL_NEAREST), clamp-to-edge: luminance replicates to rgb, alpha = 1
inline vec4 texture(const sampler3D& s, const vec3& c) {
  if (s.dtype == 3) return vec4(pool_filter(s, c, 0), pool_filter(s, c, 1), pool_filter(s, c, 2), pool_filter(s, c, 3));   // GL_RGBA8
  float v;
  if (s.nearest) {
    v = pool_texel(s, (int)floorf(c.x * (float)s.w), (int)floorf(c.y * (float)s.h), (int)floorf(c.z * (float)s.z)) * s.norm;
  } else {
    const float ux = fmaf(c.x, (float)s.w, -0.5f), uy = fmaf(c.y, (float)s.h, -0.5f), uz = fmaf(c.z, (float)s.z, -0.5f);
    const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    const float fx = ux - x0, fy = uy - y0, fz = uz - z0;
    const int x = (int)x0, y = (int)y0, z = (int)z0;
    const float v000 = pool_texel(s, x, y, z), v100 = pool_texel(s, x + 1, y, z), v010 = pool_texel(s, x, y + 1, z),
                v110 = pool_texel(s, x + 1, y + 1, z), v001 = pool_texel(s, x, y, z + 1), v101 = pool_texel(s, x + 1, y, z + 1),
                v011 = pool_texel(s, x, y + 1, z + 1), v111 = pool_texel(s, x + 1, y + 1, z + 1);
    const float c00 = fmaf(fx, v100 - v000, v000), c10 = fmaf(fx, v110 - v010, v010);
    const float c01 = fmaf(fx, v101 - v001, v001), c11 = fmaf(fx, v111 - v011, v011);
    const float c0 = fmaf(fy, c10 - c00, c00), c1 = fmaf(fy, c11 - c01, c01);
    v = fmaf(fz, c1 - c0, c0) * s.norm;
  }
  return vec4(v, v, v, 1.0f);
}
// compatibility-profile spellings
inline vec4 texture1D(const sampler1D& s, float c) { return texture(s, c); }
inline vec4 texture3D(const sampler3D& s, const vec3& c) { return texture(s, c); }
