/*
 * orc_render.c -- ORACLE (test infrastructure): CPU restatement of the
 * GridLeaper raycasting pass.  PARITY UNPINNED for the GLSL arithmetic: the
 * reference ships no golden images and its GL renderer cannot run here.
 *
 * Follows (reference file:line):
 *   ray entry / exit        Renderer/GL/GLGridLeaper.cpp:560-620 (ComputeEyeToModelMatrix,
 *                           FillRayEntryBuffer), Shaders/GLGridLeaper-NearPlane-VS.glsl:9-14,
 *                           GLGridLeaper-entry-VS.glsl:10-14, GLGridLeaper-frontfaces-FS.glsl:6-8
 *   uniforms                GLGridLeaper.cpp:690-752 (SetupRaycastShader),
 *                           GLVolumePool.cpp:790-803 (Enable), AbstrRenderer.cpp:1102-1109
 *   main() DVR              Shaders/GLGridLeaper-blend.glsl:65-228
 *   main() ISO              Shaders/GLGridLeaper-iso.glsl:68-200
 *   pool traversal          generated GLSL in Renderer/GL/GLVolumePool.cpp:484-656
 *   miss reports            generated GLSL in Renderer/GL/GLHashTable.cpp:136-182, decode :65-77
 *   classification/shading  Shaders/GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}.glsl,
 *                           GLGridLeaper-GradientTools.glsl:6-23, lighting.glsl:33-43,
 *                           Compositing.glsl:33-38
 *   iso compose             Shaders/Compose-FS.glsl:49-76, Renderer/GL/GLRenderer.cpp:2763-2830
 *   read-back               Renderer/GL/GLFrameCapture.cpp:72-85
 *
 * Where GLSL leaves precision implementation-defined, this file fixes ONE
 * arithmetic (documented in DESIGN.md "Arithmetic contract") that the CUDA
 * kernels restate independently:
 *   - all float ops IEEE single, evaluated left to right, no implicit FMA;
 *     explicit fmaf() only in the trilinear lerps, the texel-coordinate map, dot products
 *     (fma(z,z', fma(y,y', x*x'))) and the under-compositing update
 *   - normalize(v) = v * (1/sqrt(dot(v,v))), length = sqrt(dot), v/len = v * (1/len)
 *   - uint(log2(x)) = exponent of x (0 for x < 1)
 *   - pow(x, 8) = ((x*x)^2)^2; pow(x, oc) is the identity for oc == 1, powf otherwise
 *   - unorm texels are filtered as raw integers and scaled once by 1/(2^bits-1)
 *   - x / normToPoolScale is evaluated as x * (1/normToPoolScale)
 *   - gradient taps (centre +- sampleDelta, sampleDelta = one pool texel) are taken at texel
 *     index +-1 with the centre sample's filter fractions
 *   - entry/exit positions come from an analytic ray/box slab test at the pixel
 *     centre instead of rasterised bounding-box faces
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 div3(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 scl3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 norm3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return scl3(a, inv); }
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* v' = v * M, Tuvok row-vector convention (Basics/Vectors.h:434-439) */
static inline v4 xform4(const float* m, float x, float y, float z, float w) {
  v4 r;
  r.x = x * m[0] + y * m[4] + z * m[8] + w * m[12];
  r.y = x * m[1] + y * m[5] + z * m[9] + w * m[13];
  r.z = x * m[2] + y * m[6] + z * m[10] + w * m[14];
  r.w = x * m[3] + y * m[7] + z * m[11] + w * m[15];
  return r;
}

/* ---- double precision 4x4 helpers (host-side uniform derivation) ---- */
static int inv4d(const double* a, double* out) {
  double m[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) { m[r][c] = a[r * 4 + c]; m[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    if (m[p][c] == 0.0) return -1;
    if (p != c) for (int k = 0; k < 8; k++) { double t = m[c][k]; m[c][k] = m[p][k]; m[p][k] = t; }
    double d = m[c][c];
    for (int k = 0; k < 8; k++) m[c][k] = m[c][k] / d;
    for (int r = 0; r < 4; r++) if (r != c) {
      double f = m[r][c];
      for (int k = 0; k < 8; k++) m[r][k] = m[r][k] - f * m[c][k];
    }
  }
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = m[r][4 + c];
  return 0;
}
static void mul4d(const double* a, const double* b, double* out) {
  double t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s = s + a[r * 4 + k] * b[k * 4 + c];
      t[r * 4 + c] = s;
    }
  memcpy(out, t, sizeof(t));
}

typedef struct {
  float emm[16];        /* mEyeToModel */
  float inv_proj[16];   /* mInvProjection */
  float model_to_eye[16];
  float mv_inv[16];     /* inverse(modelView); mModelViewIT*v == mv_inv (column-vector product) */
  v3 extend;            /* vExtend */
  v3 domain_scale;      /* vDomainScale */
  v3 light_a, light_d, light_s, light_dir_m, eye_m;
  float lzwse;          /* fLevelZeroWorldSpaceError */
  v3 pool_size_f, vol_f, overlap_tc, total_f;
  v3 lod_layout[ORC_MAX_LOD];
  uint32_t lod_layout_sz[ORC_MAX_LOD][2];
  float norm;           /* unorm -> float factor */
  float oc;             /* ocFactor */
  v3 step;              /* filled per ray */
} uni;

static void derive(const orc_render_params* p, uni* u) {
  double mv[16], pr[16], imv[16], ipr[16], emm[16], m2e[16];
  for (int i = 0; i < 16; i++) { mv[i] = p->model_view[i]; pr[i] = p->projection[i]; }
  inv4d(mv, imv);
  inv4d(pr, ipr);
  /* vExtend = domain*scale / max; vScale /= min (GLGridLeaper.cpp:691-695) */
  float ex[3], sc[3];
  for (int i = 0; i < 3; i++) ex[i] = (float)p->vol[i] * p->scale[i];
  float mx = fmaxf(ex[0], fmaxf(ex[1], ex[2]));
  for (int i = 0; i < 3; i++) ex[i] = ex[i] / mx;
  float mn = fminf(p->scale[0], fminf(p->scale[1], p->scale[2]));
  for (int i = 0; i < 3; i++) sc[i] = p->scale[i] / mn;
  u->extend = V3(ex[0], ex[1], ex[2]);
  u->domain_scale = V3(1.0f / sc[0], 1.0f / sc[1], 1.0f / sc[2]);
  /* mEyeToModel = inverse(MV) * T(-0) * S(1/extend) * T(.5) */
  double s[16] = {0}, t[16] = {0};
  s[0] = (double)(1.0f / ex[0]); s[5] = (double)(1.0f / ex[1]); s[10] = (double)(1.0f / ex[2]); s[15] = 1;
  t[0] = t[5] = t[10] = t[15] = 1; t[12] = t[13] = t[14] = 0.5;
  mul4d(imv, s, emm);
  mul4d(emm, t, emm);
  inv4d(emm, m2e);
  for (int i = 0; i < 16; i++) {
    u->emm[i] = (float)emm[i];
    u->inv_proj[i] = (float)ipr[i];
    u->model_to_eye[i] = (float)m2e[i];
    u->mv_inv[i] = (float)imv[i];
  }
  u->light_a = V3(p->ambient[0] * p->ambient[3], p->ambient[1] * p->ambient[3], p->ambient[2] * p->ambient[3]);
  u->light_d = V3(p->diffuse[0] * p->diffuse[3], p->diffuse[1] * p->diffuse[3], p->diffuse[2] * p->diffuse[3]);
  u->light_s = V3(p->specular[0] * p->specular[3], p->specular[1] * p->specular[3], p->specular[2] * p->specular[3]);
  v4 ld = xform4(u->emm, p->light_dir[0], p->light_dir[1], p->light_dir[2], 0.0f);
  u->light_dir_m = norm3(V3(ld.x, ld.y, ld.z));
  v4 ep = xform4(u->emm, p->eye[0], p->eye[1], p->eye[2], 1.0f);
  u->eye_m = V3(ep.x, ep.y, ep.z);
  u->lzwse = fmaxf(ex[0] / (float)p->vol[0], fmaxf(ex[1] / (float)p->vol[1], ex[2] / (float)p->vol[2]));
  u->pool_size_f = V3((float)p->pool_size[0], (float)p->pool_size[1], (float)p->pool_size[2]);
  u->vol_f = V3((float)p->vol[0], (float)p->vol[1], (float)p->vol[2]);
  u->total_f = V3((float)p->max_total_brick[0], (float)p->max_total_brick[1], (float)p->max_total_brick[2]);
  u->overlap_tc = V3((p->max_total_brick[0] - p->max_inner_brick[0]) / (2.0f * p->pool_size[0]),
                     (p->max_total_brick[1] - p->max_inner_brick[1]) / (2.0f * p->pool_size[1]),
                     (p->max_total_brick[2] - p->max_inner_brick[2]) / (2.0f * p->pool_size[2]));
  for (uint32_t l = 0; l < p->lod_count; l++) {
    float c[3];
    for (int i = 0; i < 3; i++) {
      c[i] = (float)p->vol[i] / p->max_inner_brick[i];
      c[i] = c[i] / (float)(1u << l);
      if ((float)(uint32_t)c[i] == c[i]) c[i] = c[i] - c[i] * 1.1920928955078125e-07f;
    }
    u->lod_layout[l] = V3(c[0], c[1], c[2]);
    u->lod_layout_sz[l][0] = (uint32_t)ceilf(c[0]);
    u->lod_layout_sz[l][1] = (uint32_t)ceilf(c[0]) * (uint32_t)ceilf(c[1]);
  }
  u->norm = (p->dtype == ORC_U8 || p->dtype == ORC_RGBA8) ? 1.0f / 255.0f : p->dtype == ORC_U16 ? 1.0f / 65535.0f : 1.0f;
  u->oc = 1.0f / p->sample_rate_modifier;
}

/* ------------------------------------------------------------------ */
/* ray setup                                                           */
/* ------------------------------------------------------------------ */
/* slab test of the ray o + s*d against [lo,hi]; returns 0 if the ray is parallel to and outside a slab */
static int slab3(const float o[3], const float d[3], const float lo[3], const float hi[3], float* s_in, float* s_out) {
  float a_in = -INFINITY, a_out = INFINITY;
  for (int i = 0; i < 3; i++) {
    if (d[i] == 0.0f) {
      if (o[i] < lo[i] || o[i] > hi[i]) return 0;
      continue;
    }
    float t0 = (lo[i] - o[i]) / d[i], t1 = (hi[i] - o[i]) / d[i];
    float a = fminf(t0, t1), b = fmaxf(t0, t1);
    a_in = fmaxf(a_in, a);
    a_out = fminf(a_out, b);
  }
  *s_in = a_in; *s_out = a_out;
  return 1;
}

static int shard_active(const orc_render_params* p) {
  for (int i = 0; i < 3; i++) if (p->clip_min[i] > 0.0f || p->clip_max[i] < 1.0f) return 1;
  return 0;
}

static int ray_setup_px(const orc_render_params* p, const uni* u, uint32_t px, uint32_t py,
                        v4* entry, v4* exit_) {
  float nx = ((float)px + 0.5f) / (float)p->width * 2.0f - 1.0f;
  float ny = ((float)py + 0.5f) / (float)p->height * 2.0f - 1.0f;
  v4 nr = xform4(u->inv_proj, nx, ny, -1.0f, 1.0f);
  v3 pn = V3(nr.x / nr.w, nr.y / nr.w, nr.z / nr.w);       /* eye-space point on the near plane */
  v4 o4 = xform4(u->emm, 0.0f, 0.0f, 0.0f, 1.0f);
  v4 n4 = xform4(u->emm, pn.x, pn.y, pn.z, 1.0f);
  float o[3] = {o4.x, o4.y, o4.z};
  float d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
  const float zero[3] = {0.0f, 0.0f, 0.0f}, one[3] = {1.0f, 1.0f, 1.0f};
  float s_in, s_out;
  if (!slab3(o, d, zero, one, &s_in, &s_out)) return 0;
  if (p->clip_plane_on) {
    /* FillBBoxVBO: the box is cut by the plane, f <= 0 is kept (Basics/Clipper.cpp:76-101).  Model space (centre 0, extent
     * ex) -> the [0,1]^3 coordinates of the rays: p_model = (p - 0.5) * ex; f(s) = a + s * b along the ray */
    float q[4];
    q[0] = p->clip_plane[0] * u->extend.x; q[1] = p->clip_plane[1] * u->extend.y; q[2] = p->clip_plane[2] * u->extend.z;
    q[3] = p->clip_plane[3] - 0.5f * (q[0] + q[1] + q[2]);
    const float a = fmaf(q[2], o[2], fmaf(q[1], o[1], q[0] * o[0])) + q[3];
    const float b = fmaf(q[2], d[2], fmaf(q[1], d[1], q[0] * d[0]));
    if (b > 0.0f) s_out = fminf(s_out, (0.0f - a) / b);
    else if (b < 0.0f) s_in = fmaxf(s_in, (0.0f - a) / b);
    else if (a > 0.0f) return 0;
  }
  float s0 = fmaxf(s_in, 1.0f);              /* near plane where the camera is inside / in front */
  if (!(s_out > s0)) return 0;               /* no back-face fragment in front of the near plane */
  if (shard_active(p) && !p->pipeline) {
    /* sort-last: the ray keeps its whole-volume entry/exit (so its sample positions are those of the
     * single-GPU ray); a pixel whose ray never meets this rank's brick block is simply not shaded */
    float a_in, a_out;
    if (!slab3(o, d, p->clip_min, p->clip_max, &a_in, &a_out)) return 0;
    if (!(fminf(a_out, s_out) > fmaxf(a_in, s0))) return 0;
  }
  v3 pe = scl3(pn, s0), px_ = scl3(pn, s_out);
  v4 e = xform4(u->emm, pe.x, pe.y, pe.z, 1.0f);
  v4 x = xform4(u->emm, px_.x, px_.y, px_.z, 1.0f);
  entry->x = e.x; entry->y = e.y; entry->z = e.z; entry->w = pe.z;
  exit_->x = x.x; exit_->y = x.y; exit_->z = x.z; exit_->w = px_.z;
  return 1;
}

void orc_ray_setup(const orc_render_params* p, float* entry, float* exit_, uint8_t* covered) {
  uni u;
  derive(p, &u);
  for (uint32_t y = 0; y < p->height; y++)
    for (uint32_t x = 0; x < p->width; x++) {
      size_t i = (size_t)y * p->width + x;
      v4 e = {0, 0, 0, 0}, q = {0, 0, 0, 0};
      covered[i] = (uint8_t)ray_setup_px(p, &u, x, y, &e, &q);
      memcpy(entry + 4 * i, &e, 16);
      memcpy(exit_ + 4 * i, &q, 16);
    }
}

/* For tests/glsl_ref.py (the reference GLSL executed on the CPU needs the SAME inputs): the eye-space position of the
 * back-face fragment of every covered pixel (what the rasteriser interpolates into vPosInViewCoords) ... */
void orc_ray_exit_eye(const orc_render_params* p, float* exit_eye /* w*h*3 */) {
  uni u;
  derive(p, &u);
  for (uint32_t y = 0; y < p->height; y++)
    for (uint32_t x = 0; x < p->width; x++) {
      size_t i = (size_t)y * p->width + x;
      v4 e, q;
      exit_eye[3 * i] = exit_eye[3 * i + 1] = exit_eye[3 * i + 2] = 0.0f;
      if (!ray_setup_px(p, &u, x, y, &e, &q)) continue;
      float nx = ((float)x + 0.5f) / (float)p->width * 2.0f - 1.0f;
      float ny = ((float)y + 0.5f) / (float)p->height * 2.0f - 1.0f;
      v4 nr = xform4(u.inv_proj, nx, ny, -1.0f, 1.0f);
      v3 pn = V3(nr.x / nr.w, nr.y / nr.w, nr.z / nr.w);
      float s = q.w / pn.z;                 /* exit.w = (pn * s_out).z */
      (void)s;
      /* recompute s_out exactly as ray_setup_px did */
      v4 o4 = xform4(u.emm, 0.0f, 0.0f, 0.0f, 1.0f), n4 = xform4(u.emm, pn.x, pn.y, pn.z, 1.0f);
      float o[3] = {o4.x, o4.y, o4.z}, d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
      const float zero[3] = {0.0f, 0.0f, 0.0f}, one[3] = {1.0f, 1.0f, 1.0f};
      float s_in, s_out;
      slab3(o, d, zero, one, &s_in, &s_out);
      v3 px_ = scl3(pn, s_out);
      exit_eye[3 * i] = px_.x; exit_eye[3 * i + 1] = px_.y; exit_eye[3 * i + 2] = px_.z;
    }
}
/* ... and the uniforms GLGridLeaper::SetupRaycastShader would upload (GLGridLeaper.cpp:690-752):
 * out[0..15] mEyeToModel, [16..18] vDomainScale, [19..21] ambient, [22..24] diffuse, [25..27] specular (rgb * w),
 * [28..30] vModelSpaceLightDir, [31..33] vModelSpaceEyePos, [34] fLevelZeroWorldSpaceError, [35] unorm factor,
 * [36..51] mModelToEye, [52..67] inverse(modelView), [68..83] inverse(projection) */
void orc_uniforms(const orc_render_params* p, float* out) {
  uni u;
  derive(p, &u);
  memcpy(out, u.emm, 64);
  const v3* v[6] = {&u.domain_scale, &u.light_a, &u.light_d, &u.light_s, &u.light_dir_m, &u.eye_m};
  for (int i = 0; i < 6; i++) { out[16 + 3 * i] = v[i]->x; out[17 + 3 * i] = v[i]->y; out[18 + 3 * i] = v[i]->z; }
  out[34] = u.lzwse;
  out[35] = u.norm;
  memcpy(out + 36, u.model_to_eye, 64);   /* mModelToEye */
  memcpy(out + 52, u.mv_inv, 64);         /* inverse(modelView): mModelViewIT * v == column-vector product with it */
  memcpy(out + 68, u.inv_proj, 64);       /* inverse(projection) */
}

/* ------------------------------------------------------------------ */
/* pool sampling                                                       */
/* ------------------------------------------------------------------ */
typedef struct {
  const orc_render_params* p;
  const uni* u;
  const void* pool;
  const void* const* slots;  /* orc_raycast_slots: slot-linear pool given as one pointer per slot (NULL = not supplied) */
  uint64_t absent;           /* texel reads that hit a slot that was not supplied */
  const uint32_t* meta;
  const uint8_t* tf;
  uint32_t* hash;
  uint32_t finest[3];
  uint64_t samples, bricks;
  int shard;                 /* sort-last: only samples inside [sh_lo, sh_hi) are taken */
  float sh_lo[3], sh_hi[3];  /* the shard box with faces on the volume border pushed to -/+inf */
} ctx_t;

static inline float texel_ch(const ctx_t* c, int x, int y, int z, int ch) {
  const uint32_t* ps = c->p->pool_size;
  x = x < 0 ? 0 : x >= (int)ps[0] ? (int)ps[0] - 1 : x;
  y = y < 0 ? 0 : y >= (int)ps[1] ? (int)ps[1] - 1 : y;
  z = z < 0 ? 0 : z >= (int)ps[2] ? (int)ps[2] - 1 : z;
  const void* base = c->pool;
  size_t i;
  if (c->slots) {   /* the same atlas texel, addressed slot by slot (slot = the linear pool coordinate) */
    const uint32_t* tb = c->p->max_total_brick;
    const uint32_t* cap = c->p->capacity;
    const uint32_t sx = (uint32_t)x / tb[0], sy = (uint32_t)y / tb[1], sz = (uint32_t)z / tb[2];
    base = c->slots[(size_t)sx + (size_t)cap[0] * ((size_t)sy + (size_t)cap[1] * (size_t)sz)];
    if (!base) { ((ctx_t*)c)->absent++; return 0.0f; }
    i = (size_t)((uint32_t)x % tb[0]) + (size_t)tb[0] * ((size_t)((uint32_t)y % tb[1]) + (size_t)tb[1] * (size_t)((uint32_t)z % tb[2]));
  } else {
    i = (size_t)x + (size_t)ps[0] * ((size_t)y + (size_t)ps[1] * (size_t)z);
  }
  switch (c->p->dtype) {
    case ORC_U8: return (float)((const uint8_t*)base)[i];
    case ORC_U16: return (float)((const uint16_t*)base)[i];
    case ORC_RGBA8: return (float)((const uint8_t*)base)[4 * i + (size_t)ch];   /* GL_RGBA8 pool of a colour volume */
    default: return ((const float*)base)[i];
  }
}
static inline float texel(const ctx_t* c, int x, int y, int z) { return texel_ch(c, x, y, z, 0); }

/* texture(volumePool, coords + (dx,dy,dz)*sampleDelta).r -- GL_LINEAR, clamp-to-edge
 * (GLVolumePool.cpp:637-639).  sampleDelta = 1/poolSize is exactly one texel, so the offset
 * is applied to the texel index and the filter fractions of the centre are reused. */
static float sample_pool_off_ch(const ctx_t* c, v3 tc, int dx, int dy, int dz, int ch) {
#define texel(c_, x_, y_, z_) texel_ch(c_, x_, y_, z_, ch)
  const uni* u = c->u;
  if (c->p->nearest) {
    int x = (int)floorf(tc.x * u->pool_size_f.x), y = (int)floorf(tc.y * u->pool_size_f.y),
        z = (int)floorf(tc.z * u->pool_size_f.z);
    return texel(c, x + dx, y + dy, z + dz) * u->norm;
  }
  float ux = fmaf(tc.x, u->pool_size_f.x, -0.5f);
  float uy = fmaf(tc.y, u->pool_size_f.y, -0.5f);
  float uz = fmaf(tc.z, u->pool_size_f.z, -0.5f);
  float fx0 = floorf(ux), fy0 = floorf(uy), fz0 = floorf(uz);
  float fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;
  int x = (int)fx0 + dx, y = (int)fy0 + dy, z = (int)fz0 + dz;
  float v000 = texel(c, x, y, z), v100 = texel(c, x + 1, y, z);
  float v010 = texel(c, x, y + 1, z), v110 = texel(c, x + 1, y + 1, z);
  float v001 = texel(c, x, y, z + 1), v101 = texel(c, x + 1, y, z + 1);
  float v011 = texel(c, x, y + 1, z + 1), v111 = texel(c, x + 1, y + 1, z + 1);
  float c00 = fmaf(fx, v100 - v000, v000);
  float c10 = fmaf(fx, v110 - v010, v010);
  float c01 = fmaf(fx, v101 - v001, v001);
  float c11 = fmaf(fx, v111 - v011, v011);
  float c0 = fmaf(fy, c10 - c00, c00);
  float c1 = fmaf(fy, c11 - c01, c01);
  return fmaf(fz, c1 - c0, c0) * u->norm;
#undef texel
}
/* samplePool = texture(volumePool, coords).r (GLVolumePool.cpp:637-639) */
static float sample_pool_off(const ctx_t* c, v3 tc, int dx, int dy, int dz) { return sample_pool_off_ch(c, tc, dx, dy, dz, 0); }

static float sample_pool(const ctx_t* c, v3 tc) { return sample_pool_off(c, tc, 0, 0, 0); }
/* samplePool4 = texture(volumePool, coords) of a colour volume (GLVolumePool.cpp:649-651): every channel is filtered
 * with the same fractions */
static v4 sample_pool4(const ctx_t* c, v3 tc) {
  v4 r = {sample_pool_off_ch(c, tc, 0, 0, 0, 0), sample_pool_off_ch(c, tc, 0, 0, 0, 1), sample_pool_off_ch(c, tc, 0, 0, 0, 2),
          sample_pool_off_ch(c, tc, 0, 0, 0, 3)};
  return r;
}

/* GLGridLeaper-GradientTools.glsl:6-16 (note the y taps: "Yp" is fetched at -delta) */
static v3 gradient_ch(const ctx_t* c, v3 ctr, v3 delta, int ch) {
  (void)delta;
  float xp = sample_pool_off_ch(c, ctr, +1, 0, 0, ch);
  float xm = sample_pool_off_ch(c, ctr, -1, 0, 0, ch);
  float yp = sample_pool_off_ch(c, ctr, 0, -1, 0, ch);
  float ym = sample_pool_off_ch(c, ctr, 0, +1, 0, ch);
  float zp = sample_pool_off_ch(c, ctr, 0, 0, +1, ch);
  float zm = sample_pool_off_ch(c, ctr, 0, 0, -1, ch);
  return V3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
}
/* ComputeGradient: the .r channel; ComputeGradientAlpha (GradientTools.glsl:25-37): the same taps on .a */
static v3 gradient(const ctx_t* c, v3 ctr, v3 delta) { return gradient_ch(c, ctr, delta, 0); }

static v3 compute_normal(const ctx_t* c, v3 ctr, v3 delta, v3 domain_scale) {
  v3 g = gradient(c, ctr, delta);
  v3 n = mul3(g, domain_scale);
  float l = len3(n);
  if (l > 0.0f) n = scl3(n, 1.0f / l);
  return n;
}

static inline float pow8(float x) { float a = x * x; float b = a * a; return b * b; }

/* lighting.glsl:33-43 */
static v3 lighting(v3 eye, v3 pos, v3 n, v3 amb, v3 dif, v3 spe, v3 ldir) {
  v3 view = norm3(sub3(eye, pos));
  float dn = dot3(n, view);
  v3 refl = norm3(sub3(view, scl3(n, 2.0f * dn)));
  float dl = fmaxf(fabsf(dot3(n, ldir)), 0.0f);
  float sp = pow8(fmaxf(dot3(refl, ldir), 0.0f));
  v3 r = V3(amb.x + dif.x * dl + spe.x * sp, amb.y + dif.y * dl + spe.y * sp, amb.z + dif.z * dl + spe.z * sp);
  return V3(clampf(r.x, 0.0f, 1.0f), clampf(r.y, 0.0f, 1.0f), clampf(r.z, 0.0f, 1.0f));
}

/* RGBA8, GL_NEAREST, clamp-to-edge (GPUMemMan.cpp:398-401, GLTexture1D.h:48-51) */
static v4 tf_lookup(const ctx_t* c, float s, float t) {
  int w = (int)c->p->tf_w, h = (int)c->p->tf_h;
  int ix = (int)floorf(s * (float)w);
  ix = ix < 0 ? 0 : ix >= w ? w - 1 : ix;
  int iy = 0;
  if (h > 1) {
    iy = (int)floorf(t * (float)h);
    iy = iy < 0 ? 0 : iy >= h ? h - 1 : iy;
  }
  const uint8_t* q = c->tf + 4 * ((size_t)iy * w + ix);
  v4 r = {(float)q[0] / 255.0f, (float)q[1] / 255.0f, (float)q[2] / 255.0f, (float)q[3] / 255.0f};
  return r;
}

/* ComputeColorFromVolume, GLGridLeaper-Method-{1D,1D-L,2D,2D-L}.glsl */
static v4 color_from_volume(ctx_t* c, v3 pc, v3 model_pos, v3 delta) {
  const orc_render_params* p = c->p;
  const uni* u = c->u;
  c->samples++;
  if (p->dtype == ORC_RGBA8) {
    /* GLGridLeaper-Method-{1D,1D-L,2D,2D-L}-color.glsl: the volume's own colour, the transfer function only maps alpha.
     * 1D-L takes its normal from ComputeNormal (the .r channel), the 2D methods their gradient from ComputeGradientAlpha. */
    v4 data = sample_pool4(c, pc);
    if (p->mode == ORC_RM_1DTRANS) {
      data.w = tf_lookup(c, data.w * p->trans_scale, 0.0f).w;
      if (!p->lighting) return data;
      v3 n = compute_normal(c, pc, delta, u->domain_scale);
      v3 lit = lighting(u->eye_m, model_pos, n, u->light_a, mul3(V3(data.x, data.y, data.z), u->light_d), u->light_s,
                        u->light_dir_m);
      data.x = lit.x; data.y = lit.y; data.z = lit.z;
      return data;
    }
    v3 g = gradient_ch(c, pc, delta, 3);
    float gm = len3(g);
    data.w = tf_lookup(c, data.w * p->trans_scale, 1.0f - gm * p->gradient_scale).w;
    if (!p->lighting) return data;
    v3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
    v3 n = mul3(u->domain_scale, gn);
    v3 lit = lighting(u->eye_m, model_pos, n, u->light_a, mul3(V3(data.x, data.y, data.z), u->light_d), u->light_s,
                      u->light_dir_m);
    data.x = lit.x; data.y = lit.y; data.z = lit.z;
    return data;
  }
  float data = sample_pool(c, pc);
  v4 col;
  if (p->mode == ORC_RM_1DTRANS) {
    col = tf_lookup(c, data * p->trans_scale, 0.0f);
    if (!p->lighting) return col;
    v3 n = compute_normal(c, pc, delta, u->domain_scale);
    v3 lit = lighting(u->eye_m, model_pos, n, u->light_a,
                      mul3(V3(col.x, col.y, col.z), u->light_d), u->light_s, u->light_dir_m);
    col.x = lit.x; col.y = lit.y; col.z = lit.z;
    return col;
  }
  v3 g = gradient(c, pc, delta);
  float gm = len3(g);
  col = tf_lookup(c, data * p->trans_scale, 1.0f - gm * p->gradient_scale);
  if (!p->lighting) return col;
  v3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
  v3 n = mul3(u->domain_scale, gn);
  v3 lit = lighting(u->eye_m, model_pos, n, u->light_a,
                    mul3(V3(col.x, col.y, col.z), u->light_d), u->light_s, u->light_dir_m);
  col.x = lit.x; col.y = lit.y; col.z = lit.z;
  return col;
}

/* ------------------------------------------------------------------ */
/* page-table walk                                                     */
/* ------------------------------------------------------------------ */
typedef struct { uint32_t x, y, z, w; } u4;

static inline u4 brick_coords(const uni* u, v3 pos, uint32_t lod) {
  v3 l = u->lod_layout[lod];
  u4 r = {(uint32_t)(pos.x * l.x), (uint32_t)(pos.y * l.y), (uint32_t)(pos.z * l.z), lod};
  return r;
}
static inline uint32_t brick_info(const ctx_t* c, u4 b) {
  uint32_t idx = c->p->lod_offset[b.w] + b.x + b.y * c->u->lod_layout_sz[b.w][0] +
                 b.z * c->u->lod_layout_sz[b.w][1];
  return c->meta[idx];
}

/* GLHashTable.cpp:140-182 */
static void report_missing(ctx_t* c, u4 b) {
  if (!c->hash || c->p->hash_size == 0) return;
  const uint32_t* f = c->finest;
  uint32_t ser = 1 + b.x + b.y * f[0] + b.z * f[0] * f[1] + b.w * f[0] * f[1] * f[2];
  uint32_t rehash = 0;
  do {
    uint32_t h = (ser + rehash) % c->p->hash_size;
    uint32_t old = __sync_val_compare_and_swap(&c->hash[h], 0u, ser);
    if (old == 0 || old == ser) return;
  } while (++rehash < c->p->rehash_count);
}

typedef struct {
  v3 pool_entry, pool_exit, norm_exit, scale, trans;
  int empty;
  int where;   /* sort-last: 0 = brick inside the shard box, 1 = straddles it, 2 = outside (never sampled) */
  u4 bc;
} brick_t;

enum { IN_SHARD = 0, PARTLY_IN_SHARD = 1, OUTSIDE_SHARD = 2 };

/* position of the brick's box [c0,c1] relative to the shard box */
static int classify_brick(const ctx_t* c, v3 c0, v3 c1) {
  if (!c->shard) return IN_SHARD;
  const float a0[3] = {c0.x, c0.y, c0.z}, a1[3] = {c1.x, c1.y, c1.z};
  int inside = 1;
  for (int i = 0; i < 3; i++) {
    if (a1[i] <= c->sh_lo[i] || a0[i] >= c->sh_hi[i]) return OUTSIDE_SHARD;
    if (a0[i] < c->sh_lo[i] || a1[i] > c->sh_hi[i]) inside = 0;
  }
  return inside ? IN_SHARD : PARTLY_IN_SHARD;
}
static void brick_corners(const uni* u, u4 bc, v3* c0, v3* c1) {
  v3 lay = u->lod_layout[bc.w];
  *c0 = div3(V3((float)bc.x, (float)bc.y, (float)bc.z), lay);
  *c1 = div3(V3((float)(bc.x + 1), (float)(bc.y + 1), (float)(bc.z + 1)), lay);
}

static int get_brick(ctx_t* c, v3 pos, uint32_t* lod, v3 dir, brick_t* o) {
  const orc_render_params* p = c->p;
  const uni* u = c->u;
  const uint32_t max_lod = p->lod_count - 1;
  c->bricks++;
  pos = V3(clampf(pos.x, 0.0f, 1.0f), clampf(pos.y, 0.0f, 1.0f), clampf(pos.z, 0.0f, 1.0f));
  int found = 1;
  u4 bc = brick_coords(u, pos, *lod);
  uint32_t info = brick_info(c, bc);
  int foreign = 0;
  if (info == ORC_BI_MISSING && c->shard) {
    /* sort-last: a brick that does not touch this rank's block lives on another rank.  It is walked
     * through (same step arithmetic, nominal slot 0) but never requested, sampled or replaced by a
     * coarser level, so the ray reaches this rank's block at the single-GPU ray's sample phase. */
    v3 f0, f1;
    brick_corners(u, bc, &f0, &f1);
    foreign = classify_brick(c, f0, f1) == OUTSIDE_SHARD;
  }
  if (info == ORC_BI_MISSING && !foreign) {
    uint32_t start = *lod;
    report_missing(c, bc);
    found = 0;
    do {
      (*lod)++;
      bc = brick_coords(u, pos, *lod);
      info = brick_info(c, bc);
      if (info == ORC_BI_MISSING) {
        if (p->strategy == ORC_BS_REQUEST_ALL) report_missing(c, bc);
        else if (p->strategy == ORC_BS_SKIP_ONE && start + 1 == *lod) report_missing(c, bc);
        else if (p->strategy == ORC_BS_SKIP_TWO && start + 2 == *lod) report_missing(c, bc);
      }
    } while (info == ORC_BI_MISSING);
  }
  o->empty = !foreign && info <= ORC_BI_EMPTY;
  if (o->empty) {
    for (uint32_t lo = *lod + 1; lo < max_lod; ++lo) {   /* strict <: coarsest level never leapt to (H5) */
      u4 lb = brick_coords(u, pos, lo);
      uint32_t li = brick_info(c, lb);
      if (li == ORC_BI_CHILD_EMPTY) { bc = lb; info = li; *lod = lo; }
      else break;
    }
  }
  /* GetBrickCorners */
  v3 c0, c1;
  brick_corners(u, bc, &c0, &c1);
  /* BrickExit */
  v3 dv = V3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
  float tx = ((dv.x < 0.0f ? c0.x : c1.x) - pos.x) * dv.x;
  float ty = ((dv.y < 0.0f ? c0.y : c1.y) - pos.y) * dv.y;
  float tz = ((dv.z < 0.0f ? c0.z : c1.z) - pos.z) * dv.z;
  float tm = fminf(fminf(tx, ty), tz);
  o->norm_exit = add3(pos, scl3(dir, tm));
  o->bc = bc;
  o->where = IN_SHARD;
  if (o->empty) return found;
  o->where = foreign ? OUTSIDE_SHARD : classify_brick(c, c0, c1);
  /* NormCoordsToPoolCoords / BrickPoolCoords / InfoToCoords */
  /* a brick outside the shard box is only stepped through: always in the pool coordinates of slot 0, resident or not,
   * so the ray's positions behind it do not depend on what other views have paged into this pool */
  uint32_t index = o->where == OUTSIDE_SHARD ? 0u : info - ORC_BI_FLAG_COUNT;
  uint32_t sx = index % p->capacity[0], sy = (index / p->capacity[0]) % p->capacity[1],
           sz = index / (p->capacity[0] * p->capacity[1]);
  v3 vp = V3((float)(sx * p->max_total_brick[0]), (float)(sy * p->max_total_brick[1]),
             (float)(sz * p->max_total_brick[2]));
  v3 vq = V3((float)(sx * p->max_total_brick[0] + p->max_total_brick[0]),
             (float)(sy * p->max_total_brick[1] + p->max_total_brick[1]),
             (float)(sz * p->max_total_brick[2] + p->max_total_brick[2]));
  v3 pc0 = add3(div3(vp, u->pool_size_f), u->overlap_tc);
  v3 pc1 = sub3(div3(vq, u->pool_size_f), u->overlap_tc);
  o->scale = div3(sub3(pc1, pc0), sub3(c1, c0));
  o->trans = sub3(pc0, mul3(c0, o->scale));
  o->pool_entry = add3(mul3(pos, o->scale), o->trans);
  o->pool_exit = add3(mul3(o->norm_exit, o->scale), o->trans);
  return found;
}

static inline uint32_t compute_lod(const ctx_t* c, float dist) {
  float x = c->p->lod_factor * (-dist) / c->u->lzwse;
  uint32_t max_lod = c->p->lod_count - 1;
  if (!(x >= 1.0f)) return 0;
  if (isinf(x)) return max_lod;
  int e;
  frexpf(x, &e);
  uint32_t l = (uint32_t)(e - 1);
  return l < max_lod ? l : max_lod;
}

static inline float opacity_correct(const uni* u, float a) {
  if (u->oc == 1.0f) return a;
  return 1.0f - powf(1.0f - a, u->oc);
}

/* ------------------------------------------------------------------ */
/* main()                                                              */
/* ------------------------------------------------------------------ */
static void trace_pixel(ctx_t* c, const float* ray_start, const float* start_color,
                        const float* exit_, float* o0, float* o1, float* o2, float* o3) {
  const orc_render_params* p = c->p;
  const uni* u = c->u;
  const int iso = p->mode == ORC_RM_ISOSURFACE;
  v4 acc, resume_col, resume_pos, hit_pos = {0, 0, 0, 0}, hit_nrm = {0, 0, 0, 0}, resume_nrm = {0, 0, 0, 0};
  memcpy(&resume_pos, ray_start, 16);
  if (!iso) {
    memcpy(&acc, start_color, 16);
    resume_col = acc;
    if (resume_pos.w == 1000.0f) goto done;
  } else {
    if (floorf(resume_pos.w) == 1000.0f) goto done;
    if (floorf(resume_pos.w) == 500.0f) {
      hit_pos = xform4(u->model_to_eye, resume_pos.x, resume_pos.y, resume_pos.z, 1.0f);
      hit_pos.w = resume_pos.w - floorf(resume_pos.w) + 1.0f;
      memcpy(&hit_nrm, start_color, 16);   /* rayStartNormal */
      resume_nrm = hit_nrm;
      goto done;
    }
  }
  {
    v3 entry = V3(resume_pos.x, resume_pos.y, resume_pos.z);
    float entry_depth = resume_pos.w;
    v3 nexit = V3(exit_[0], exit_[1], exit_[2]);
    float exit_depth = exit_[3];
    v3 dir = sub3(nexit, entry);
    float ray_len = len3(dir);
    /* TransformToPoolSpace */
    v3 vdir = norm3(mul3(dir, u->vol_f));
    vdir = div3(vdir, u->pool_size_f);
    float den = 2.0f * p->sample_rate_modifier;
    vdir = V3(vdir.x / den, vdir.y / den, vdir.z / den);
    v3 delta = V3(1.0f / u->pool_size_f.x, 1.0f / u->pool_size_f.y, 1.0f / u->pool_size_f.z);
    float step = len3(vdir);
    float t = 0.0f;
    int optimal = 1;
    const float voxel_size = 0.125f / 2000.0f;
    v3 cur = entry;
    u4 last = {0, 0, 0, 9999};
    int terminated = 0;
    int handoff = 0;              /* pipeline stage: the ray left this stage's slab alive, at hand */
    v4 hand = {0, 0, 0, 0};
    if (ray_len > voxel_size) {
      for (uint32_t j = 0; j < 100 && !terminated; ++j) {
        if (c->shard) {   /* the block is convex: once the ray has left it, nothing more to do on this rank */
          const float cp[3] = {cur.x, cur.y, cur.z}, dd[3] = {dir.x, dir.y, dir.z};
          int gone = 0;
          for (int i = 0; i < 3; i++)
            gone |= (dd[i] > 0.0f && cp[i] >= c->sh_hi[i]) || (dd[i] < 0.0f && cp[i] <= c->sh_lo[i]);
          if (gone) {
            if (p->pipeline) {   /* where the next stage picks the ray up */
              handoff = 1;
              hand.x = cur.x; hand.y = cur.y; hand.z = cur.z;
              hand.w = entry_depth * (1.0f - t) + exit_depth * t;
            }
            break;
          }
        }
        float cur_depth = entry_depth * (1.0f - t) + exit_depth * t;
        uint32_t lod = compute_lod(c, cur_depth);
        brick_t b;
        int ok = get_brick(c, cur, &lod, dir, &b);
        /* the shader's GetBrick clamps its own copy of the position; currentPos itself is unchanged */
        if (!ok && optimal) {
          optimal = 0;
          resume_pos.x = cur.x; resume_pos.y = cur.y; resume_pos.z = cur.z; resume_pos.w = cur_depth;
          if (!iso) resume_col = acc;
        }
        if (!b.empty && !(last.x == b.bc.x && last.y == b.bc.y && last.z == b.bc.z && last.w == b.bc.w)) {
          int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / step);
          int s2 = (int)ceilf(len3(mul3(sub3(nexit, cur), b.scale)) / step);
          steps = steps < s2 ? steps : s2;
          v3 inv_scale = V3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
          v3 pc = b.pool_entry;
          if (b.where == OUTSIDE_SHARD) {   /* another rank's brick: advance by its steps, take no sample */
            float n = (float)(steps > 0 ? steps : 0);
            pc = V3(fmaf(n, vdir.x, pc.x), fmaf(n, vdir.y, pc.y), fmaf(n, vdir.z, pc.z));
            steps = 0;
          }
          for (int i = 0; i < steps; ++i) {
            if (b.where == PARTLY_IN_SHARD) {   /* brick straddles the block face: ownership per sample */
              v3 mq = mul3(sub3(pc, b.trans), inv_scale);
              const float q[3] = {mq.x, mq.y, mq.z};
              int mine = 1;
              for (int a = 0; a < 3; a++) mine &= q[a] >= c->sh_lo[a] && q[a] < c->sh_hi[a];
              if (!mine && p->pipeline) {
                /* a brick of a coarser LoD straddles the slab's far side: the ray is handed on AT the side, so
                 * the next stage takes the brick's remaining samples */
                const float dd[3] = {dir.x, dir.y, dir.z};
                int gone = 0;
                for (int a = 0; a < 3; a++)
                  gone |= (dd[a] > 0.0f && q[a] >= c->sh_hi[a]) || (dd[a] < 0.0f && q[a] < c->sh_lo[a]);
                if (gone) {
                  const float tq = len3(sub3(mq, entry)) / ray_len;
                  handoff = 1;
                  hand.x = mq.x; hand.y = mq.y; hand.z = mq.z;
                  hand.w = entry_depth * (1.0f - tq) + exit_depth * tq;
                  terminated = 1;
                  break;
                }
              }
              if (!mine) { pc = add3(pc, vdir); continue; }
            }
            if (!iso) {
              v3 mp = mul3(sub3(pc, b.trans), inv_scale);
              v4 col = color_from_volume(c, pc, mp, delta);
              col.w = opacity_correct(u, col.w);
              /* UnderCompositing */
              float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) { terminated = 1; break; }
            } else {
              c->samples++;
              /* GetVolumeHit: scalar volumes test .r and report white; colour volumes (GLGridLeaper-Method-iso-color.glsl)
               * test .a and report the sampled colour; RefineIsosurface bisects on the same channel */
              const int colour = p->dtype == ORC_RGBA8, ich = colour ? 3 : 0;
              v4 hcol = {1.0f, 1.0f, 1.0f, 1.0f};
              if (colour) hcol = sample_pool4(c, pc);
              if ((colour ? hcol.w : sample_pool(c, pc)) >= p->isoval) {
                /* RefineIsosurface */
                v3 rd = V3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
                pc = sub3(pc, rd);
                for (int k = 0; k < 5; k++) {
                  rd = V3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
                  if (sample_pool_off_ch(c, pc, 0, 0, 0, ich) >= p->isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
                }
                cur = mul3(sub3(pc, b.trans), inv_scale);
                hit_pos = xform4(u->model_to_eye, cur.x, cur.y, cur.z, 1.0f);
                hit_pos.w = hcol.x + 1.0f;               /* color.r + 1 */
                v3 n = compute_normal(c, pc, delta, u->domain_scale);
                /* mModelViewIT * vec4(n,0): column-vector product with inverse(modelView) */
                const float* m = u->mv_inv;
                hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
                hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
                hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
                hit_nrm.w = floorf(hcol.y * 512.0f) + hcol.z;  /* floor(color.g*512)+color.b */
                terminated = 1;
                break;
              } else {
                hit_pos.x = hit_pos.y = hit_pos.z = hit_pos.w = 0.0f;
              }
            }
            pc = add3(pc, vdir);
          }
          if (terminated) break;
          cur = mul3(sub3(pc, b.trans), inv_scale);
        } else {
          float k = voxel_size;
          cur = V3(b.norm_exit.x + k * dir.x / ray_len, b.norm_exit.y + k * dir.y / ray_len,
                   b.norm_exit.z + k * dir.z / ray_len);
        }
        last = b.bc;
        t = len3(sub3(entry, b.norm_exit)) / ray_len;
        if (t > 0.9999f) break;
      }
    }
    /* TerminateRay */
    if (!iso) {
      if (optimal) {
        if (p->pipeline && handoff && !(acc.w > 0.99f)) { resume_pos = hand; resume_col = acc; }
        else { resume_pos.w = 1000.0f; resume_col = acc; }
      }
    } else {
      if (optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
      resume_nrm = hit_nrm;
    }
  }
done:
  if (!iso) {
    memcpy(o0, &acc, 16); memcpy(o1, &resume_col, 16); memcpy(o2, &resume_pos, 16);
  } else {
    memcpy(o0, &hit_pos, 16); memcpy(o1, &hit_nrm, 16); memcpy(o2, &resume_pos, 16);
    if (o3) memcpy(o3, &resume_nrm, 16);
  }
}

static void raycast_impl(const orc_render_params* p, const void* pool, const void* const* slots, const uint32_t* meta,
                         const uint8_t* tf, const float* ray_start, const float* start_color,
                         const float* exit_, const uint8_t* covered,
                         float* out0, float* out1, float* out2, float* out3,
                         uint32_t* hash, orc_render_stats* stats, uint64_t* absent_reads, int n_threads) {
  uni u;
  derive(p, &u);
  size_t n = (size_t)p->width * p->height;
  /* render targets are cleared (GLGridLeaper.cpp:837) */
  memset(out0, 0, n * 16); memset(out1, 0, n * 16); memset(out2, 0, n * 16);
  if (out3) memset(out3, 0, n * 16);
  uint64_t samples = 0, rays = 0, bricks = 0, absent = 0;
  uint32_t finest[3];
  for (int i = 0; i < 3; i++)
    finest[i] = (uint32_t)ceil((double)p->vol[i] / p->max_inner_brick[i]);
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads) reduction(+ : samples, rays, bricks, absent)
  for (int64_t y = 0; y < (int64_t)p->height; y++) {
    ctx_t c;
    c.p = p; c.u = &u; c.pool = pool; c.slots = slots; c.absent = 0; c.meta = meta; c.tf = tf; c.hash = hash;
    memcpy(c.finest, finest, sizeof(finest));
    c.samples = 0; c.bricks = 0;
    c.shard = shard_active(p);
    for (int i = 0; i < 3; i++) {
      c.sh_lo[i] = p->clip_min[i] > 0.0f ? p->clip_min[i] : -INFINITY;
      c.sh_hi[i] = p->clip_max[i] < 1.0f ? p->clip_max[i] : INFINITY;
    }
    for (uint32_t x = 0; x < p->width; x++) {
      size_t i = (size_t)y * p->width + x;
      if (!covered[i]) {
        if (p->pipeline) out2[4 * i + 3] = 1000.0f;   /* a stage marks pixels outside the volume finished */
        continue;
      }
      rays++;
      trace_pixel(&c, ray_start + 4 * i, start_color + 4 * i, exit_ + 4 * i,
                  out0 + 4 * i, out1 + 4 * i, out2 + 4 * i, out3 ? out3 + 4 * i : NULL);
    }
    samples += c.samples;
    bricks += c.bricks;
    absent += c.absent;
  }
  if (absent_reads) *absent_reads = absent;
  if (stats) {
    stats->samples = samples; stats->rays = rays; stats->brick_visits = bricks;
    uint32_t e = 0;
    if (hash) for (uint32_t i = 0; i < p->hash_size; i++) e += hash[i] != 0;
    stats->hash_entries = e;
  }
}

void orc_raycast(const orc_render_params* p, const void* pool, const uint32_t* meta,
                 const uint8_t* tf, const float* ray_start, const float* start_color,
                 const float* exit_, const uint8_t* covered,
                 float* out0, float* out1, float* out2, float* out3,
                 uint32_t* hash, orc_render_stats* stats, int n_threads) {
  raycast_impl(p, pool, NULL, meta, tf, ray_start, start_color, exit_, covered, out0, out1, out2, out3, hash, stats, NULL,
               n_threads);
}

void orc_raycast_slots(const orc_render_params* p, const void* const* slots, const uint32_t* meta,
                       const uint8_t* tf, const float* ray_start, const float* start_color,
                       const float* exit_, const uint8_t* covered,
                       float* out0, float* out1, float* out2, float* out3,
                       uint32_t* hash, orc_render_stats* stats, uint64_t* absent_reads, int n_threads) {
  raycast_impl(p, NULL, slots, meta, tf, ray_start, start_color, exit_, covered, out0, out1, out2, out3, hash, stats,
               absent_reads, n_threads);
}

/* Compose-FS.glsl:49-76; light colours: GLRenderer.cpp:2772-2808 */
void orc_iso_compose(const orc_render_params* p, const float* hit_pos, const float* hit_normal, float* rgba) {
  size_t n = (size_t)p->width * p->height;
  v3 a = V3(p->ambient[0] * p->ambient[3], p->ambient[1] * p->ambient[3], p->ambient[2] * p->ambient[3]);
  v3 d = V3(p->diffuse[0] * p->diffuse[3] * p->iso_color[0], p->diffuse[1] * p->diffuse[3] * p->iso_color[1],
            p->diffuse[2] * p->diffuse[3] * p->iso_color[2]);
  v3 s = V3(p->specular[0] * p->specular[3], p->specular[1] * p->specular[3], p->specular[2] * p->specular[3]);
  v3 l = V3(p->light_dir[0], p->light_dir[1], p->light_dir[2]);
  for (size_t i = 0; i < n; i++) {
    const float* hp = hit_pos + 4 * i;
    float* o = rgba + 4 * i;
    o[0] = o[1] = o[2] = o[3] = 0.0f;
    if (hp[3] == 0.0f) continue;
    v3 nrm = V3(hit_normal[4 * i], hit_normal[4 * i + 1], fabsf(hit_normal[4 * i + 2]));
    v3 view = norm3(V3(0.0f - hp[0], 0.0f - hp[1], 0.0f - hp[2]));
    float dn = dot3(nrm, view);
    v3 refl = norm3(sub3(view, scl3(nrm, 2.0f * dn)));
    float dl = fmaxf(fabsf(dot3(nrm, V3(-l.x, -l.y, -l.z))), 0.0f);
    float sp = pow8(fmaxf(dot3(refl, l), 0.0f));
    o[0] = clampf(a.x + d.x * dl + s.x * sp, 0.0f, 1.0f);
    o[1] = clampf(a.y + d.y * dl + s.y * sp, 0.0f, 1.0f);
    o[2] = clampf(a.z + d.z * dl + s.z * sp, 0.0f, 1.0f);
    o[3] = 1.0f;
  }
}

/* Compose-Color-FS.glsl:60-92 (colour volumes): as Compose-FS, but the diffuse colour is the hit's own colour, recovered from
 * the two alpha channels -- r = pos.a - 1, g = floor(nrm.a / 2) / 256, b = fract(nrm.a) -- times vLightDiffuse (which
 * GLRenderer sets WITHOUT the isosurface colour for colour data, GLRenderer.cpp:2796-2810) */
void orc_iso_compose_color(const orc_render_params* p, const float* hit_pos, const float* hit_normal, float* rgba) {
  size_t n = (size_t)p->width * p->height;
  v3 a = V3(p->ambient[0] * p->ambient[3], p->ambient[1] * p->ambient[3], p->ambient[2] * p->ambient[3]);
  v3 d0 = V3(p->diffuse[0] * p->diffuse[3], p->diffuse[1] * p->diffuse[3], p->diffuse[2] * p->diffuse[3]);
  v3 s = V3(p->specular[0] * p->specular[3], p->specular[1] * p->specular[3], p->specular[2] * p->specular[3]);
  v3 l = V3(p->light_dir[0], p->light_dir[1], p->light_dir[2]);
  for (size_t i = 0; i < n; i++) {
    const float* hp = hit_pos + 4 * i;
    float* o = rgba + 4 * i;
    o[0] = o[1] = o[2] = o[3] = 0.0f;
    if (hp[3] == 0.0f) continue;
    const float na = hit_normal[4 * i + 3];
    v3 col = V3(hp[3] - 1.0f, floorf(na / 2.0f) / 256.0f, na - floorf(na));
    v3 d = mul3(col, d0);
    v3 nrm = V3(hit_normal[4 * i], hit_normal[4 * i + 1], fabsf(hit_normal[4 * i + 2]));
    v3 view = norm3(V3(0.0f - hp[0], 0.0f - hp[1], 0.0f - hp[2]));
    float dn = dot3(nrm, view);
    v3 refl = norm3(sub3(view, scl3(nrm, 2.0f * dn)));
    float dl = fmaxf(fabsf(dot3(nrm, V3(-l.x, -l.y, -l.z))), 0.0f);
    float sp = pow8(fmaxf(dot3(refl, l), 0.0f));
    o[0] = clampf(a.x + d.x * dl + s.x * sp, 0.0f, 1.0f);
    o[1] = clampf(a.y + d.y * dl + s.y * sp, 0.0f, 1.0f);
    o[2] = clampf(a.z + d.z * dl + s.z * sp, 0.0f, 1.0f);
    o[3] = 1.0f;
  }
}

/* one miss report outside a render pass (same arithmetic as report_missing above): returns the number of rehashes, or
 * rehash_count when the probe chain is exhausted -- the return value of the generated GLSL `Hash(uvec4)` */
uint32_t orc_hash_insert(uint32_t* hash, uint32_t hash_size, uint32_t rehash_count, const uint32_t f[3], uint32_t x,
                         uint32_t y, uint32_t z, uint32_t lod) {
  uint32_t ser = 1 + x + y * f[0] + z * f[0] * f[1] + lod * f[0] * f[1] * f[2];
  uint32_t rehash = 0;
  do {
    uint32_t h = (ser + rehash) % hash_size;
    uint32_t old = __sync_val_compare_and_swap(&hash[h], 0u, ser);
    if (old == 0 || old == ser) return rehash;
  } while (++rehash < rehash_count);
  return rehash_count;
}

uint32_t orc_hash_decode(const uint32_t* hash, uint32_t hash_size, const uint32_t f[3], uint32_t* out) {
  uint32_t n = 0;
  for (uint32_t i = 0; i < hash_size; i++) {
    uint32_t e = hash[i];
    if (!e) continue;
    uint32_t idx = e - 1, vol = f[0] * f[1] * f[2];
    uint32_t w = idx / vol; idx -= w * vol;
    uint32_t z = idx / (f[0] * f[1]); idx -= z * (f[0] * f[1]);
    uint32_t y = idx / f[0]; idx -= y * f[0];
    out[4 * n] = idx; out[4 * n + 1] = y; out[4 * n + 2] = z; out[4 * n + 3] = w;
    n++;
  }
  return n;
}

/* GL float -> unorm8: clamp to [0,1], scale by 255, round to nearest */
void orc_rgba8(const float* rgba, uint64_t n_pixels, uint8_t* out) {
  for (uint64_t i = 0; i < n_pixels * 4; i++) {
    float v = rgba[i];
    v = v < 0.0f ? 0.0f : v > 1.0f ? 1.0f : v;
    if (v != v) v = 0.0f;
    out[i] = (uint8_t)(v * 255.0f + 0.5f);
  }
}

void orc_composite_over(const float* front, const float* back, uint64_t n_pixels, float* out) {
  for (uint64_t i = 0; i < n_pixels; i++) {
    const float* f = front + 4 * i;
    const float* b = back + 4 * i;
    /* early-terminated front rays (alpha > 0.99, GLGridLeaper-blend.glsl:180) hide what lies behind */
    float oma = f[3] > 0.99f ? 0.0f : 1.0f - f[3];
    /* a ray that would have crossed 0.99 inside the back block stops there as well: cut the back image at
     * the middle of (0.99, 1.0], the interval in which the single-GPU ray ends (alpha error <= 0.005) */
    float add = oma * b[3];
    if (oma > 0.0f && f[3] + add > 0.995f) oma = oma * ((0.995f - f[3]) / add);
    out[4 * i + 0] = f[0] + oma * b[0];
    out[4 * i + 1] = f[1] + oma * b[1];
    out[4 * i + 2] = f[2] + oma * b[2];
    out[4 * i + 3] = f[3] + oma * b[3];
  }
}
