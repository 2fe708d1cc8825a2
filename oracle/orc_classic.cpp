/*
 * orc_classic.cpp -- ORACLE (test infrastructure): CPU restatement of the CLASSIC per-brick
 * raycaster (GLRaycaster) and of the frame planning in front of it.  PARITY UNPINNED for the
 * GLSL arithmetic (no golden images in the reference, no GL here); the planning side follows the
 * reference's C++ line by line.
 *
 * Follows (reference file:line):
 *   LOD choice            AbstrRenderer::ComputeMinLODForCurrentView  Renderer/AbstrRenderer.cpp:789-803
 *                         CullingLOD::GetLODLevel / SetScreenParams   Renderer/CullingLOD.cpp:57-67,126-138
 *   brick metadata        UVFDataset::ComputeMetadataTOC              IO/uvfDataset.cpp:242-288
 *                         ExtendedOctree::ComputeMetadata (LoD aspect) IO/UVF/ExtendedOctree/ExtendedOctree.cpp:188-243
 *   brick list            AbstrRenderer::BuildSubFrameBrickList       Renderer/AbstrRenderer.cpp:999-1100
 *                         RegionNeedsBrick :874-920, brick_distance :808-841, ContainsData :953-997
 *                         UVFDataset::ContainsData (legacy tests)     IO/uvfDataset.cpp:1201-1227
 *                         CullingLOD::Update / IsVisible              Renderer/CullingLOD.cpp:89-124,141-160
 *                         UVFDataset::GetTextCoords (TOC)             IO/uvfDataset.cpp:1747-1771
 *   per-brick pass        GLRaycaster::Render3DPreLoop / Render3DInLoop / SetBrickDepShaderVars /
 *                         RenderBox / ComputeEyeToTextureMatrix       Renderer/GL/GLRaycaster.cpp:239-300,302-345,348-478,589-612
 *   shaders               GLRaycaster-1D-FS.glsl:51-82, -1D-light-FS.glsl:84-134, -2D-FS.glsl:53-94,
 *                         -2D-light-FS.glsl:61-117, VRender1D.glsl:39-57, VRender1DLit.glsl:49-72,
 *                         Volume3D.glsl:39-60, lighting.glsl:33-50, Compositing.glsl:33-38
 *   blending              GL state `ONE_MINUS_DST_ALPHA, ONE`         Renderer/GL/GLRenderer.cpp:151-153
 *
 * Arithmetic contract: the one of orc_render.c (IEEE fp32, explicit fmaf only in lerps / dot products /
 * compositing, normalize = v * (1/sqrt(dot))), plus:
 *   - ray entry/exit: analytic slab test of the eye ray against the brick's world box at the pixel
 *     centre; the entry position passes through the RGBA16F ray-entry FBO (GLRaycaster.cpp:97), i.e. it
 *     is rounded to half precision; the FBO keeps its previous content where a brick has no visible
 *     front face (near plane first, GLRaycaster.cpp:348-381)
 *   - eye -> texture: world = eye * inverse(modelView) (fp32 matrix), tex = (world - pMax) * s + tMax with
 *     s = (tMin - tMax) / (pMin - pMax)  (ComputeEyeToTextureMatrix called with the max corner first)
 *   - brick texture: GL_LINEAR, clamp-to-edge on the brick's OWN size; gradient taps at +-1 texel
 *     (vVoxelStepsize = 1/voxelCount) with the centre's filter fractions
 *   - gl_NormalMatrix = transpose(inverse(upper 3x3 of modelView)), taken from inverse(modelView)
 *   - bricks with equal distance keep TOC order (std::stable_sort; the reference's std::sort leaves ties
 *     unspecified)
 */
#include "orc.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline v3 div3(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline v3 scl3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
inline v3 norm3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return scl3(a, inv); }
inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
inline float max3(v3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
inline float min3(v3 a) { return fminf(a.x, fminf(a.y, a.z)); }

inline v4 xform4(const float* m, float x, float y, float z, float w) {   /* v' = v * M */
  v4 r;
  r.x = x * m[0] + y * m[4] + z * m[8] + w * m[12];
  r.y = x * m[1] + y * m[5] + z * m[9] + w * m[13];
  r.z = x * m[2] + y * m[6] + z * m[10] + w * m[14];
  r.w = x * m[3] + y * m[7] + z * m[11] + w * m[15];
  return r;
}

bool inv4d(const double* a, double* out) {
  double m[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) { m[r][c] = a[r * 4 + c]; m[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    if (m[p][c] == 0.0) return false;
    if (p != c) for (int k = 0; k < 8; k++) std::swap(m[c][k], m[p][k]);
    double d = m[c][c];
    for (int k = 0; k < 8; k++) m[c][k] = m[c][k] / d;
    for (int r = 0; r < 4; r++) if (r != c) {
      double f = m[r][c];
      for (int k = 0; k < 8; k++) m[r][k] = m[r][k] - f * m[c][k];
    }
  }
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = m[r][4 + c];
  return true;
}

/* IEEE binary16 round trip (round to nearest even), what a GL_RGBA16F render target stores */
float half_round(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = x & 0x80000000u;
  uint32_t a = x & 0x7fffffffu;
  if (a >= 0x7f800000u) return f;                                  /* inf / nan */
  if (a >= 0x477ff000u) { uint32_t r = sign | 0x7f800000u; float o; memcpy(&o, &r, 4); return o; }   /* >= 65520 -> inf */
  if (a < 0x33000001u) { uint32_t r = sign; float o; memcpy(&o, &r, 4); return o; }                   /* < 2^-25 -> 0 */
  float af; memcpy(&af, &a, 4);
  int e; frexpf(af, &e);                                           /* af = m * 2^e, m in [0.5,1) */
  int ulp_exp = (e - 1 < -14 ? -14 : e - 1) - 10;                  /* exponent of one half ulp step */
  float q = ldexpf(nearbyintf(ldexpf(af, -ulp_exp)), ulp_exp);     /* default rounding mode: nearest even */
  uint32_t r; memcpy(&r, &q, 4); r |= sign;
  float o; memcpy(&o, &r, 4);
  return o;
}

struct lod_geo { uint32_t size[3]; uint32_t layout[3]; double aspect[3]; };

/* ExtendedOctree::ComputeMetadata: sizes, brick counts and the LoD aspect (anisotropic downsampling) */
std::vector<lod_geo> lod_table(const uint32_t vol[3], const uint32_t max_brick[3], uint32_t overlap) {
  std::vector<lod_geo> t;
  uint64_t s[3] = {vol[0], vol[1], vol[2]};
  double asp[3] = {1.0, 1.0, 1.0};
  do {
    lod_geo l;
    uint64_t n[3] = {s[0], s[1], s[2]};
    if (!t.empty()) {
      for (int i = 0; i < 3; i++)
        if (s[i] > 1) {
          n[i] = (uint64_t)ceil(s[i] / 2.0);
          asp[i] *= (s[i] % 2) ? float(s[i]) / float(n[i]) : 2;
        }
      double mx = std::max(asp[0], std::max(asp[1], asp[2]));
      for (int i = 0; i < 3; i++) { asp[i] /= mx; s[i] = n[i]; }
    }
    for (int i = 0; i < 3; i++) {
      l.size[i] = (uint32_t)s[i];
      l.aspect[i] = asp[i];
      l.layout[i] = (uint32_t)ceil(s[i] / double(max_brick[i] - 2 * overlap));
    }
    t.push_back(l);
  } while (s[0] > 1 || s[1] > 1 || s[2] > 1);
  return t;
}

/* ExtendedOctree::ComputeBrickSize */
uint32_t brick_extent(uint32_t lod_size, uint32_t max_brick, uint32_t overlap, uint32_t layout, uint32_t c) {
  const uint32_t inner = max_brick - 2 * overlap;
  if (c + 1 < layout) return max_brick;
  const uint32_t rest = lod_size % inner;
  return rest == 0 ? max_brick : 2 * overlap + rest;
}

struct planes6 { float p[6][4]; };
planes6 frustum(const float* mv, const float* pr) {   /* CullingLOD::Update with m = MV * P (row vectors) */
  float m[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float s = 0.0f;
      for (int k = 0; k < 4; k++) s += mv[r * 4 + k] * pr[k * 4 + c];
      m[r * 4 + c] = s;
    }
  planes6 P;
  const int col[6] = {0, 0, 1, 1, 2, 2};
  const float sg[6] = {-1.0f, 1.0f, -1.0f, 1.0f, 1.0f, -1.0f};   /* right, left, top, bottom, far, near */
  for (int i = 0; i < 6; i++)
    for (int r = 0; r < 4; r++) P.p[i][r] = sg[i] * m[r * 4 + col[i]] + m[r * 4 + 3];
  return P;
}
bool is_visible(const planes6& P, v3 c, v3 e) {
  v3 h = scl3(e, 0.5f);
  for (int i = 0; i < 6; i++) {
    const float* p = P.p[i];
    if (p[0] * c.x + p[1] * c.y + p[2] * c.z + p[3] <= -(h.x * fabsf(p[0]) + h.y * fabsf(p[1]) + h.z * fabsf(p[2])))
      return false;
  }
  return true;
}

float near_plane(const float* pr) {   /* Perspective(): m33 = -(f+n)/(f-n), m43 = -2fn/(f-n) */
  return (float)((double)pr[14] / ((double)pr[10] - 1.0));
}

}  // namespace

extern "C" {

/* AbstrRenderer::ComputeMinLODForCurrentView: the LoD a headless / converged frame is rendered at */
uint32_t orc_classic_lod(const orc_render_params* p, uint32_t lod_count) {
  v3 ext = V3((float)p->vol[0] * p->scale[0], (float)p->vol[1] * p->scale[1], (float)p->vol[2] * p->scale[2]);
  float mx = max3(ext);
  ext = V3(ext.x / mx, ext.y / mx, ext.z / mx);
  const float lzwse = max3(div3(ext, V3((float)p->vol[0], (float)p->vol[1], (float)p->vol[2])));
  const float z_center = p->model_view[14];                 /* (0,0,0,1) * MV */
  const float fz = fmaxf(near_plane(p->projection), -z_center);
  int lod = (int)floorf(logf(p->lod_factor * fz / lzwse) / logf(2.0f));
  if (lod < 0) lod = 0;
  if (lod > (int)lod_count - 1) lod = (int)lod_count - 1;
  return (uint32_t)lod;
}

/* BuildSubFrameBrickList for one LoD.  minmax: 4 doubles per brick of THIS LoD in TOC order (x fastest).
 * vis = {tfMin, tfMax, gradMin, gradMax} (1D/2D, already rescaled) or {iso}.  Returns the list length. */
static uint32_t brick_list_impl(const orc_render_params* p, uint32_t lod, uint32_t overlap, const double* minmax,
                                const double vis[4], orc_classic_brick* out, uint32_t cap, bool pass_all) {
  std::vector<lod_geo> lt = lod_table(p->vol, p->max_total_brick, overlap);
  if (lod >= lt.size()) return 0;
  const lod_geo& L = lt[lod];
  /* vScale of BuildSubFrameBrickList (domain size of LoD 0) and of RegionNeedsBrick (current LoD) */
  auto corrected = [&](const uint32_t dom[3]) {
    v3 sc = V3(p->scale[0], p->scale[1], p->scale[2]);
    float dmax = (float)std::max(dom[0], std::max(dom[1], dom[2]));
    v3 c = V3(sc.x * (float)dom[0] / dmax, sc.y * (float)dom[1] / dmax, sc.z * (float)dom[2] / dmax);
    float m = max3(c);
    return V3(sc.x / m, sc.y / m, sc.z / m);
  };
  const v3 scale_list = corrected(lt[0].size), scale_cull = corrected(L.size);
  const planes6 P = frustum(p->model_view, p->projection);

  /* ComputeMetadataTOC */
  const v3 asp = V3((float)L.aspect[0], (float)L.aspect[1], (float)L.aspect[2]);
  v3 nds = mul3(V3((float)L.size[0], (float)L.size[1], (float)L.size[2]), asp);
  const float max_val = max3(nds);
  nds = V3(nds.x / max_val, nds.y / max_val, nds.z / max_val);
  std::vector<orc_classic_brick> list;
  v3 corner = V3(0, 0, 0);
  v3 ext_md = V3(0, 0, 0);
  for (uint32_t x = 0; x < L.layout[0]; x++) {
    corner.y = 0;
    for (uint32_t y = 0; y < L.layout[1]; y++) {
      corner.z = 0;
      for (uint32_t z = 0; z < L.layout[2]; z++) {
        const uint32_t n[3] = {brick_extent(L.size[0], p->max_total_brick[0], overlap, L.layout[0], x),
                               brick_extent(L.size[1], p->max_total_brick[1], overlap, L.layout[1], y),
                               brick_extent(L.size[2], p->max_total_brick[2], overlap, L.layout[2], z)};
        const v3 eff = V3((float)(n[0] - 2 * overlap), (float)(n[1] - 2 * overlap), (float)(n[2] - 2 * overlap));
        v3 e = mul3(eff, asp);
        ext_md = V3(e.x / max_val, e.y / max_val, e.z / max_val);
        const v3 half = V3(ext_md.x / 2.0f, ext_md.y / 2.0f, ext_md.z / 2.0f);
        const v3 ctr = sub3(add3(corner, half), scl3(nds, 0.5f));
        orc_classic_brick b;
        memset(&b, 0, sizeof(b));
        b.coord[0] = x; b.coord[1] = y; b.coord[2] = z;
        b.index = z * L.layout[0] * L.layout[1] + y * L.layout[0] + x;
        for (int i = 0; i < 3; i++) b.n_vox[i] = n[i];
        /* RegionNeedsBrick: frustum test with the current LoD's scale */
        const v3 ce = mul3(ctr, scale_cull), ee = mul3(ext_md, scale_cull);
        bool needed = pass_all || is_visible(P, ce, ee);   /* CullingLOD::SetPassAll(true) for HQ MIP frames */
        if (needed) {
          const double* mm = minmax + 4 * (size_t)b.index;
          bool has;
          if (p->mode == ORC_RM_1DTRANS) has = vis[1] >= mm[0] && vis[0] <= mm[1];
          else if (p->mode == ORC_RM_2DTRANS) has = (vis[1] >= mm[0] && vis[0] <= mm[1]) && (vis[3] >= mm[2] && vis[2] <= mm[3]);
          else has = vis[0] <= mm[1];                       /* legacy iso test: one-sided */
          b.empty = has ? 0 : 1;
          const v3 c = mul3(ctr, scale_list), ex = mul3(ext_md, scale_list);
          b.center[0] = c.x; b.center[1] = c.y; b.center[2] = c.z;
          b.ext[0] = ex.x; b.ext[1] = ex.y; b.ext[2] = ex.z;
          if (!b.empty) {
            for (int i = 0; i < 3; i++) {
              b.tex_min[i] = (float)overlap / (float)n[i];
              b.tex_max[i] = (1.0f - b.tex_min[i]) * (i == 0 ? asp.x : i == 1 ? asp.y : asp.z);
            }
            /* brick_distance: closest of the 8 corners pulled in by 0.4999 */
            float dmin = 3.402823466e+38f;
            for (int k = 0; k < 8; k++) {
              const float sx = (k & 4) ? 1.0f : -1.0f, sy = (k & 2) ? 1.0f : -1.0f, sz = (k & 1) ? 1.0f : -1.0f;
              const v3 q = add3(c, scl3(V3(sx * ex.x, sy * ex.y, sz * ex.z), 0.4999f));
              const v4 t = xform4(p->model_view, q.x, q.y, q.z, 1.0f);
              dmin = fminf(dmin, len3(V3(t.x, t.y, t.z)));
            }
            b.distance = pass_all ? 0.0f : dmin;   /* bUseResidencyAsDistanceCriterion: every brick is resident here */
          }
          list.push_back(b);
        }
        corner.z += ext_md.z;
      }
      corner.y += ext_md.y;
    }
    corner.x += ext_md.x;
  }
  /* the BrickTable iterates in key order = TOC index order; ties keep that order */
  std::stable_sort(list.begin(), list.end(), [](const orc_classic_brick& a, const orc_classic_brick& b) { return a.index < b.index; });
  std::stable_sort(list.begin(), list.end(), [](const orc_classic_brick& a, const orc_classic_brick& b) { return a.distance < b.distance; });
  const uint32_t n = (uint32_t)std::min<size_t>(list.size(), cap);
  if (out) memcpy(out, list.data(), n * sizeof(orc_classic_brick));
  return (uint32_t)list.size();
}

uint32_t orc_classic_brick_list(const orc_render_params* p, uint32_t lod, uint32_t overlap, const double* minmax,
                                const double vis[4], orc_classic_brick* out, uint32_t cap) {
  return brick_list_impl(p, lod, overlap, minmax, vis, out, cap, false);
}

/* AbstrRenderer::PlanHQMIPFrame (AbstrRenderer.cpp:1214-1245): LoD 0, or with m_bMIPLOD the coarsest LoD whose
 * smallest extent is still >= the largest window side, stepped back by one */
uint32_t orc_mip_lod(const orc_render_params* p, uint32_t lod_count, int use_mip_lod) {
  uint32_t vc[3] = {p->vol[0], p->vol[1], p->vol[2]};
  uint64_t lod = 0;
  if (use_mip_lod) {
    const uint32_t win = std::max(p->width, p->height);
    while (std::min(vc[0], std::min(vc[1], vc[2])) >= win) {
      for (int i = 0; i < 3; i++) vc[i] /= 2;
      lod++;
    }
  }
  if (lod > 0) lod = std::min<uint64_t>(lod_count - 1, lod - 1);
  return (uint32_t)lod;
}

/* BuildSubFrameBrickList(true) of a HQ MIP frame: no frustum culling, key order */
uint32_t orc_mip_brick_list(const orc_render_params* p, uint32_t lod, uint32_t overlap, const double* minmax,
                            const double vis[4], orc_classic_brick* out, uint32_t cap) {
  return brick_list_impl(p, lod, overlap, minmax, vis, out, cap, true);
}

}  // extern "C"

namespace {

struct cuni {
  float inv_proj[16], imv[16];
  v3 domain_scale, la, ld, ls, ldir;
  float norm;
};

struct btex {
  const void* data; int dtype; uint32_t n[3]; int nearest;
  float texel(int x, int y, int z) const {
    x = x < 0 ? 0 : x >= (int)n[0] ? (int)n[0] - 1 : x;
    y = y < 0 ? 0 : y >= (int)n[1] ? (int)n[1] - 1 : y;
    z = z < 0 ? 0 : z >= (int)n[2] ? (int)n[2] - 1 : z;
    size_t i = (size_t)x + (size_t)n[0] * ((size_t)y + (size_t)n[1] * (size_t)z);
    switch (dtype) {
      case ORC_U8: return (float)((const uint8_t*)data)[i];
      case ORC_U16: return (float)((const uint16_t*)data)[i];
      default: return ((const float*)data)[i];
    }
  }
  /* texture3D(texVolume, tc + (dx,dy,dz)/n).x */
  float sample(v3 tc, int dx, int dy, int dz, float norm) const {
    if (nearest) {
      int x = (int)floorf(tc.x * (float)n[0]), y = (int)floorf(tc.y * (float)n[1]), z = (int)floorf(tc.z * (float)n[2]);
      return texel(x + dx, y + dy, z + dz) * norm;
    }
    float ux = fmaf(tc.x, (float)n[0], -0.5f), uy = fmaf(tc.y, (float)n[1], -0.5f), uz = fmaf(tc.z, (float)n[2], -0.5f);
    float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
    float fx = ux - x0, fy = uy - y0, fz = uz - z0;
    int x = (int)x0 + dx, y = (int)y0 + dy, z = (int)z0 + dz;
    float v000 = texel(x, y, z), v100 = texel(x + 1, y, z), v010 = texel(x, y + 1, z), v110 = texel(x + 1, y + 1, z);
    float v001 = texel(x, y, z + 1), v101 = texel(x + 1, y, z + 1), v011 = texel(x, y + 1, z + 1), v111 = texel(x + 1, y + 1, z + 1);
    float c00 = fmaf(fx, v100 - v000, v000), c10 = fmaf(fx, v110 - v010, v010);
    float c01 = fmaf(fx, v101 - v001, v001), c11 = fmaf(fx, v111 - v011, v011);
    float c0 = fmaf(fy, c10 - c00, c00), c1 = fmaf(fy, c11 - c01, c01);
    return fmaf(fz, c1 - c0, c0) * norm;
  }
};

inline float pow8(float x) { float a = x * x; float b = a * a; return b * b; }

/* lighting.glsl:33-50 with the eye at the origin */
v3 lighting(v3 pos, v3 n, v3 amb, v3 dif, v3 spe, v3 ldir) {
  v3 view = norm3(sub3(V3(0, 0, 0), pos));
  float dn = dot3(n, view);
  v3 refl = norm3(sub3(view, scl3(n, 2.0f * dn)));
  float dl = fmaxf(fabsf(dot3(n, ldir)), 0.0f);
  float sp = pow8(fmaxf(dot3(refl, ldir), 0.0f));
  return V3(clampf(amb.x + dif.x * dl + spe.x * sp, 0.0f, 1.0f), clampf(amb.y + dif.y * dl + spe.y * sp, 0.0f, 1.0f),
            clampf(amb.z + dif.z * dl + spe.z * sp, 0.0f, 1.0f));
}

v4 tf_lookup(const uint8_t* tf, int w, int h, float s, float t) {
  int ix = (int)floorf(s * (float)w);
  ix = ix < 0 ? 0 : ix >= w ? w - 1 : ix;
  int iy = 0;
  if (h > 1) { iy = (int)floorf(t * (float)h); iy = iy < 0 ? 0 : iy >= h ? h - 1 : iy; }
  const uint8_t* q = tf + 4 * ((size_t)iy * w + ix);
  v4 r = {(float)q[0] / 255.0f, (float)q[1] / 255.0f, (float)q[2] / 255.0f, (float)q[3] / 255.0f};
  return r;
}

}  // namespace

extern "C" {

float orc_classic_step_scale(const orc_render_params* p, uint32_t lod) {
  std::vector<lod_geo> lt = lod_table(p->vol, p->max_total_brick, (p->max_total_brick[0] - p->max_inner_brick[0]) / 2);
  const lod_geo& L = lt[lod];
  return 1.0f / p->sample_rate_modifier *
         fmaxf((float)p->vol[0] / (float)L.size[0], fmaxf((float)p->vol[1] / (float)L.size[1], (float)p->vol[2] / (float)L.size[2]));
}

/* One classic frame: bricks in list order, each raycast per pixel and blended `dst += (1 - dst.a) * src`.
 * brick_data[i]: voxels of list[i] (x fastest, the brick's own size incl. ghost), NULL for empty bricks.
 * out: w*h*4 floats (premultiplied RGBA, cleared to 0).  stats->samples counts VRender* evaluations. */
void orc_classic_render(const orc_render_params* p, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                        const void* const* brick_data, const uint8_t* tf, float* out, orc_render_stats* stats,
                        int n_threads) {
  cuni u;
  double mv[16], pr[16], imv[16], ipr[16];
  for (int i = 0; i < 16; i++) { mv[i] = p->model_view[i]; pr[i] = p->projection[i]; }
  inv4d(mv, imv); inv4d(pr, ipr);
  for (int i = 0; i < 16; i++) { u.imv[i] = (float)imv[i]; u.inv_proj[i] = (float)ipr[i]; }
  const float mn = fminf(p->scale[0], fminf(p->scale[1], p->scale[2]));
  u.domain_scale = V3(1.0f / (p->scale[0] / mn), 1.0f / (p->scale[1] / mn), 1.0f / (p->scale[2] / mn));
  u.la = V3(p->ambient[0] * p->ambient[3], p->ambient[1] * p->ambient[3], p->ambient[2] * p->ambient[3]);
  u.ld = V3(p->diffuse[0] * p->diffuse[3], p->diffuse[1] * p->diffuse[3], p->diffuse[2] * p->diffuse[3]);
  u.ls = V3(p->specular[0] * p->specular[3], p->specular[1] * p->specular[3], p->specular[2] * p->specular[3]);
  u.ldir = V3(p->light_dir[0], p->light_dir[1], p->light_dir[2]);   /* eye space (GLRenderer light uniforms) */
  u.norm = p->dtype == ORC_U8 ? 1.0f / 255.0f : p->dtype == ORC_U16 ? 1.0f / 65535.0f : 1.0f;
  /* fStepScale = 1/sampleRate * max(domain(0) / domain(lod)) (GLRaycaster.cpp:257) */
  std::vector<lod_geo> lt = lod_table(p->vol, p->max_total_brick, (p->max_total_brick[0] - p->max_inner_brick[0]) / 2);
  const lod_geo& L = lt[lod];
  const float step_scale = 1.0f / p->sample_rate_modifier *
                           fmaxf((float)p->vol[0] / (float)L.size[0], fmaxf((float)p->vol[1] / (float)L.size[1], (float)p->vol[2] / (float)L.size[2]));
  const size_t n_pix = (size_t)p->width * p->height;
  memset(out, 0, n_pix * 16);
  std::vector<float> fbo(n_pix * 3);   /* the RGBA16F ray-entry FBO: near plane first (Render3DPreLoop) */
  std::vector<float> near_pt(n_pix * 3);
  for (uint32_t y = 0; y < p->height; y++)
    for (uint32_t x = 0; x < p->width; x++) {
      float nx = ((float)x + 0.5f) / (float)p->width * 2.0f - 1.0f;
      float ny = ((float)y + 0.5f) / (float)p->height * 2.0f - 1.0f;
      v4 nr = xform4(u.inv_proj, nx, ny, -1.0f, 1.0f);
      size_t i = (size_t)y * p->width + x;
      near_pt[3 * i] = nr.x / nr.w; near_pt[3 * i + 1] = nr.y / nr.w; near_pt[3 * i + 2] = nr.z / nr.w;
      for (int k = 0; k < 3; k++) fbo[3 * i + k] = half_round(near_pt[3 * i + k]);
    }
  uint64_t samples = 0;
  const v4 o4 = xform4(u.imv, 0.0f, 0.0f, 0.0f, 1.0f);   /* camera in world space */
  for (uint32_t bi = 0; bi < n_bricks; bi++) {
    const orc_classic_brick& b = list[bi];
    if (b.empty) continue;
    btex T; T.data = brick_data[bi]; T.dtype = p->dtype; T.nearest = p->nearest;
    for (int i = 0; i < 3; i++) T.n[i] = b.n_vox[i];
    const v3 c = V3(b.center[0], b.center[1], b.center[2]), e = V3(b.ext[0], b.ext[1], b.ext[2]);
    const v3 pmin = sub3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f)), pmax = add3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f));
    const v3 tmin = V3(b.tex_min[0], b.tex_min[1], b.tex_min[2]), tmax = V3(b.tex_max[0], b.tex_max[1], b.tex_max[2]);
    const v3 tsc = div3(sub3(tmin, tmax), sub3(pmin, pmax));
    const v3 vstep = V3(1.0f / (float)b.n_vox[0], 1.0f / (float)b.n_vox[1], 1.0f / (float)b.n_vox[2]);
    const float ray_step = min3(scl3(mul3(e, vstep), 0.5f * 1.0f / p->sample_rate_modifier));
    const float lo[3] = {pmin.x, pmin.y, pmin.z}, hi[3] = {pmax.x, pmax.y, pmax.z};
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads < 1 ? 1 : n_threads) reduction(+ : samples)
    for (int64_t py = 0; py < (int64_t)p->height; py++)
      for (uint32_t px = 0; px < p->width; px++) {
        const size_t i = (size_t)py * p->width + px;
        const v3 pn = V3(near_pt[3 * i], near_pt[3 * i + 1], near_pt[3 * i + 2]);
        const v4 n4 = xform4(u.imv, pn.x, pn.y, pn.z, 1.0f);
        const float o[3] = {o4.x, o4.y, o4.z}, d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
        float s_in = -INFINITY, s_out = INFINITY;
        bool miss = false;
        for (int k = 0; k < 3; k++) {
          if (d[k] == 0.0f) { if (o[k] < lo[k] || o[k] > hi[k]) miss = true; continue; }
          float t0 = (lo[k] - o[k]) / d[k], t1 = (hi[k] - o[k]) / d[k];
          s_in = fmaxf(s_in, fminf(t0, t1));
          s_out = fminf(s_out, fmaxf(t0, t1));
        }
        if (miss || !(s_out > fmaxf(s_in, 1.0f))) continue;   /* no back-face fragment in front of the near plane */
        if (s_in > 1.0f) {                                    /* a front face is visible: it overwrites the entry FBO */
          const v3 fe = scl3(pn, s_in);
          fbo[3 * i] = half_round(fe.x); fbo[3 * i + 1] = half_round(fe.y); fbo[3 * i + 2] = half_round(fe.z);
        }
        const v3 entry = V3(fbo[3 * i], fbo[3 * i + 1], fbo[3 * i + 2]);
        const v3 exit_ = scl3(pn, s_out);
        auto to_tex = [&](v3 q) {
          v4 w = xform4(u.imv, q.x, q.y, q.z, 1.0f);
          return add3(mul3(sub3(V3(w.x, w.y, w.z), pmax), tsc), tmax);
        };
        const v3 et = to_tex(entry), xt = to_tex(exit_);
        v3 rd = sub3(exit_, entry);
        const float len = len3(rd);
        const float nsteps = len / ray_step;
        const int count = (int)nsteps + 1;
        const v3 inc_tex = V3((xt.x - et.x) / nsteps, (xt.y - et.y) / nsteps, (xt.z - et.z) / nsteps);
        rd = V3(rd.x / len, rd.y / len, rd.z / len);
        const v3 inc = scl3(rd, ray_step);
        v4 col = {0, 0, 0, 0};
        v3 ct = et, cp = entry;
        for (int s = 0; s < count; s++) {
          samples++;
          const float data = T.sample(ct, 0, 0, 0, u.norm);
          v4 sc;
          if (p->mode == ORC_RM_1DTRANS && !p->lighting) {
            sc = tf_lookup(tf, (int)p->tf_w, (int)p->tf_h, data * p->trans_scale, 0.0f);
          } else {
            const float xp = T.sample(ct, +1, 0, 0, u.norm), xm = T.sample(ct, -1, 0, 0, u.norm);
            const float yp = T.sample(ct, 0, -1, 0, u.norm), ym = T.sample(ct, 0, +1, 0, u.norm);
            const float zp = T.sample(ct, 0, 0, +1, u.norm), zm = T.sample(ct, 0, 0, -1, u.norm);
            const v3 g = V3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
            if (p->mode == ORC_RM_1DTRANS) sc = tf_lookup(tf, (int)p->tf_w, (int)p->tf_h, data * p->trans_scale, 0.0f);
            else sc = tf_lookup(tf, (int)p->tf_w, (int)p->tf_h, data * p->trans_scale, 1.0f - len3(g) * p->gradient_scale);
            if (p->lighting) {
              /* ComputeNormal: gl_NormalMatrix * (gradient * domainScale), safe-normalised (Volume3D.glsl:55-60) */
              const v3 gs = mul3(g, u.domain_scale);
              const float* m = u.imv;   /* normal matrix = transpose(inverse(MV3x3)): n' = inverse(MV)^T * n */
              v3 nr = V3(m[0] * gs.x + m[1] * gs.y + m[2] * gs.z, m[4] * gs.x + m[5] * gs.y + m[6] * gs.z,
                         m[8] * gs.x + m[9] * gs.y + m[10] * gs.z);
              const float l = len3(nr);
              if (l > 0.0f) nr = scl3(nr, 1.0f / l);
              v3 lit = lighting(cp, nr, u.la, mul3(V3(sc.x, sc.y, sc.z), u.ld), u.ls, u.ldir);
              if (p->mode == ORC_RM_2DTRANS) lit = V3(clampf(lit.x, 0, 1), clampf(lit.y, 0, 1), clampf(lit.z, 0, 1));
              sc.x = lit.x; sc.y = lit.y; sc.z = lit.z;
            }
          }
          sc.w = step_scale == 1.0f ? sc.w : 1.0f - powf(1.0f - sc.w, step_scale);
          const float oma = 1.0f - col.w;   /* UnderCompositing */
          col.x = fmaf(sc.x * oma, sc.w, col.x); col.y = fmaf(sc.y * oma, sc.w, col.y);
          col.z = fmaf(sc.z * oma, sc.w, col.z); col.w = fmaf(sc.w, oma, col.w);
          if (col.w >= 0.99f) break;
          cp = add3(cp, inc);
          ct = add3(ct, inc_tex);
        }
        /* GL blending ONE_MINUS_DST_ALPHA, ONE (GLRenderer.cpp:151-153) */
        float* dst = out + 4 * i;
        const float k = 1.0f - dst[3];
        dst[0] = fmaf(k, col.x, dst[0]); dst[1] = fmaf(k, col.y, dst[1]);
        dst[2] = fmaf(k, col.z, dst[2]); dst[3] = fmaf(k, col.w, dst[3]);
      }
  }
  if (stats) { memset(stats, 0, sizeof(*stats)); stats->samples = samples; }
}


/* One HQ MIP frame (GLRenderer.cpp:1183-1253): per non-empty brick the front faces go into the RGBA16F ray-entry
 * FBO, the back-face pass marches GLRaycaster-MIP-Rot-FS.glsl:47-77 (max of texture3D(texVolume).x at
 * iStepCount = int(len / fRayStepsize) + 1 positions, step = min(ext / voxels) * 0.5 / sampleRate,
 * GLRaycaster.cpp:494-530) and GL blends with BE_MAX into (max, max, max, 1); Transfer-MIP-FS.glsl:43-52 then
 * maps the maximum through the 1D transfer function, ignoring its opacity.
 * The view is whatever model_view / projection hold (the caller passes m_maMIPRotation * view, GLRaycaster.cpp:481-492;
 * perspective rays from the eye -- m_bOrthoView defaults to false, AbstrRenderer.cpp:127).  The entry FBO starts out as
 * the near-plane points, as in orc_classic_render (the reference leaves it with the previous 3D frame's content).
 * out_max: w*h*2 floats (maximum, coverage flag = the blended alpha); out: w*h*4 floats RGBA. */
void orc_mip_render(const orc_render_params* p, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                    const void* const* brick_data, const uint8_t* tf1d, uint32_t tf_n, float* out_max, float* out,
                    orc_render_stats* stats, int n_threads) {
  (void)lod;
  double mv[16], pr[16], imv_d[16], ipr_d[16];
  float imv[16], inv_proj[16];
  for (int i = 0; i < 16; i++) { mv[i] = p->model_view[i]; pr[i] = p->projection[i]; }
  inv4d(mv, imv_d); inv4d(pr, ipr_d);
  for (int i = 0; i < 16; i++) { imv[i] = (float)imv_d[i]; inv_proj[i] = (float)ipr_d[i]; }
  const float norm = p->dtype == ORC_U8 ? 1.0f / 255.0f : p->dtype == ORC_U16 ? 1.0f / 65535.0f : 1.0f;
  const size_t n_pix = (size_t)p->width * p->height;
  memset(out_max, 0, n_pix * 8);
  /* m_bOrthoView (GLRenderer.cpp:1183-1197, GLRaycaster.cpp:486-489): the projection is FLOATMATRIX4::Ortho and the model view
   * the MIP rotation alone.  The shaders are projection-agnostic (entry FBO + interpolated back-face position); what changes is
   * the ray a fragment lies on: eye-space points a + s * b with a = 0, b = the near-plane point for a perspective projection
   * (s = 1 on the near plane), and b = far-plane point - near-plane point, a = near-plane point - b for a parallel one.
   * A projection is parallel when w' does not depend on z (Vectors.h:1279-1284: array[11] = 0). */
  const int ortho = inv_proj[11] == 0.0f;
  std::vector<float> fbo(n_pix * 3), near_pt(n_pix * 3), org_pt(n_pix * 3, 0.0f), dir_pt(n_pix * 3);
  for (uint32_t y = 0; y < p->height; y++)
    for (uint32_t x = 0; x < p->width; x++) {
      float nx = ((float)x + 0.5f) / (float)p->width * 2.0f - 1.0f;
      float ny = ((float)y + 0.5f) / (float)p->height * 2.0f - 1.0f;
      v4 nr = xform4(inv_proj, nx, ny, -1.0f, 1.0f);
      size_t i = (size_t)y * p->width + x;
      near_pt[3 * i] = nr.x / nr.w; near_pt[3 * i + 1] = nr.y / nr.w; near_pt[3 * i + 2] = nr.z / nr.w;
      for (int k = 0; k < 3; k++) fbo[3 * i + k] = half_round(near_pt[3 * i + k]);
      for (int k = 0; k < 3; k++) dir_pt[3 * i + k] = near_pt[3 * i + k];
      if (ortho) {
        v4 fr = xform4(inv_proj, nx, ny, 1.0f, 1.0f);
        const float far_pt[3] = {fr.x / fr.w, fr.y / fr.w, fr.z / fr.w};
        for (int k = 0; k < 3; k++) {
          dir_pt[3 * i + k] = far_pt[k] - near_pt[3 * i + k];
          org_pt[3 * i + k] = near_pt[3 * i + k] - dir_pt[3 * i + k];
        }
      }
    }
  uint64_t samples = 0;
  for (uint32_t bi = 0; bi < n_bricks; bi++) {
    const orc_classic_brick& b = list[bi];
    if (b.empty) continue;                 /* "for MIP we do not consider empty bricks" GLRenderer.cpp:1209-1211 */
    btex T; T.data = brick_data[bi]; T.dtype = p->dtype; T.nearest = p->nearest;
    for (int i = 0; i < 3; i++) T.n[i] = b.n_vox[i];
    const v3 c = V3(b.center[0], b.center[1], b.center[2]), e = V3(b.ext[0], b.ext[1], b.ext[2]);
    const v3 pmin = sub3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f)), pmax = add3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f));
    const v3 tmin = V3(b.tex_min[0], b.tex_min[1], b.tex_min[2]), tmax = V3(b.tex_max[0], b.tex_max[1], b.tex_max[2]);
    const v3 tsc = div3(sub3(tmin, tmax), sub3(pmin, pmax));
    const v3 vstep = V3(1.0f / (float)b.n_vox[0], 1.0f / (float)b.n_vox[1], 1.0f / (float)b.n_vox[2]);
    const float ray_step = min3(scl3(mul3(e, vstep), 0.5f * 1.0f / p->sample_rate_modifier));
    const float lo[3] = {pmin.x, pmin.y, pmin.z}, hi[3] = {pmax.x, pmax.y, pmax.z};
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads < 1 ? 1 : n_threads) reduction(+ : samples)
    for (int64_t py = 0; py < (int64_t)p->height; py++)
      for (uint32_t px = 0; px < p->width; px++) {
        const size_t i = (size_t)py * p->width + px;
        const v3 pn = V3(near_pt[3 * i], near_pt[3 * i + 1], near_pt[3 * i + 2]);
        const v3 pa = V3(org_pt[3 * i], org_pt[3 * i + 1], org_pt[3 * i + 2]), pb = V3(dir_pt[3 * i], dir_pt[3 * i + 1], dir_pt[3 * i + 2]);
        const v4 o4 = xform4(imv, pa.x, pa.y, pa.z, 1.0f);
        const v4 n4 = xform4(imv, pn.x, pn.y, pn.z, 1.0f);
        const float o[3] = {o4.x, o4.y, o4.z}, d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
        float s_in = -INFINITY, s_out = INFINITY;
        bool miss = false;
        for (int k = 0; k < 3; k++) {
          if (d[k] == 0.0f) { if (o[k] < lo[k] || o[k] > hi[k]) miss = true; continue; }
          float t0 = (lo[k] - o[k]) / d[k], t1 = (hi[k] - o[k]) / d[k];
          s_in = fmaxf(s_in, fminf(t0, t1));
          s_out = fminf(s_out, fmaxf(t0, t1));
        }
        if (miss || !(s_out > fmaxf(s_in, 1.0f))) continue;
        if (s_in > 1.0f) {
          const v3 fe = ortho ? add3(pa, scl3(pb, s_in)) : scl3(pn, s_in);
          fbo[3 * i] = half_round(fe.x); fbo[3 * i + 1] = half_round(fe.y); fbo[3 * i + 2] = half_round(fe.z);
        }
        const v3 entry = V3(fbo[3 * i], fbo[3 * i + 1], fbo[3 * i + 2]);
        const v3 exit_ = ortho ? add3(pa, scl3(pb, s_out)) : scl3(pn, s_out);
        auto to_tex = [&](v3 q) {
          v4 w = xform4(imv, q.x, q.y, q.z, 1.0f);
          return add3(mul3(sub3(V3(w.x, w.y, w.z), pmax), tsc), tmax);
        };
        const v3 et = to_tex(entry), xt = to_tex(exit_);
        const float len = len3(sub3(exit_, entry));
        const float nsteps = len / ray_step;
        const int count = (int)nsteps + 1;
        const v3 inc_tex = V3((xt.x - et.x) / nsteps, (xt.y - et.y) / nsteps, (xt.z - et.z) / nsteps);
        float mx = 0.0f;
        v3 ct = et;
        for (int s = 0; s < count; s++) {
          samples++;
          mx = fmaxf(mx, T.sample(ct, 0, 0, 0, norm));
          ct = add3(ct, inc_tex);
        }
        float* dst = out_max + 2 * i;      /* BE_MAX of (max, max, max, 1) */
        dst[0] = fmaxf(dst[0], mx);
        dst[1] = 1.0f;
      }
  }
  for (size_t i = 0; i < n_pix; i++) {     /* Transfer-MIP-FS.glsl */
    float* q = out + 4 * i;
    q[0] = q[1] = q[2] = 0.0f; q[3] = 1.0f;
    if (out_max[2 * i + 1] > 0.5f) {
      const v4 t = tf_lookup(tf1d, (int)tf_n, 1, out_max[2 * i] * p->trans_scale, 0.0f);
      q[0] = t.x; q[1] = t.y; q[2] = t.z;
    }
  }
  if (stats) { memset(stats, 0, sizeof(*stats)); stats->samples = samples; }
}


/* Classic isosurface frame (GLRaycaster::Render3DInLoop, RM_ISOSURFACE branch, GLRaycaster.cpp:383-446): per
 * non-empty brick in list order the front faces go into the RGBA16F entry FBO and the back-face pass runs
 * GLRaycaster-ISO-FS.glsl:56-105 -- first sample with value >= fIsoval, RefineIsosurface.glsl:37-52 (5 bisection
 * steps), eye-space hit position by interpolation between ray entry and exit, normal = ComputeNormal
 * (Volume3D.glsl:55-60), gl_FragDepth = vProjParam.x + vProjParam.y / -z -- into the two iso-hit targets under the
 * depth test DF_LESS of the base state (GLRenderer.cpp:139; cleared to 1), so the nearest hit of all bricks stays.
 * hit_pos: w*h*(x, y, z, fInterpolParam); hit_normal: w*h*(nx, ny, nz, brick number in the list).  The image is then
 * orc_iso_compose (Compose-FS.glsl, which treats hit_pos.w == 0 as "no hit" -- a hit exactly at the ray entry is
 * dropped there, in the reference as well).  vProjParam = (far / (far - near), far * near / (near - far)) with near
 * and far recovered from the projection matrix (m33 = -(f+n)/(f-n), m43 = -2fn/(f-n)) in double. */
static void classic_iso_impl(const orc_render_params* p, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                             const void* const* brick_data, float* hit_pos, float* hit_normal, float cv_isoval,
                             float* cv_pos, float* cv_normal, orc_render_stats* stats, int n_threads) {
  (void)lod;
  double mv[16], pr[16], imv_d[16], ipr_d[16];
  float imv[16], inv_proj[16];
  for (int i = 0; i < 16; i++) { mv[i] = p->model_view[i]; pr[i] = p->projection[i]; }
  inv4d(mv, imv_d); inv4d(pr, ipr_d);
  for (int i = 0; i < 16; i++) { imv[i] = (float)imv_d[i]; inv_proj[i] = (float)ipr_d[i]; }
  const float norm = p->dtype == ORC_U8 ? 1.0f / 255.0f : p->dtype == ORC_U16 ? 1.0f / 65535.0f : 1.0f;
  const float mn = fminf(p->scale[0], fminf(p->scale[1], p->scale[2]));
  const v3 dscale = V3(1.0f / (p->scale[0] / mn), 1.0f / (p->scale[1] / mn), 1.0f / (p->scale[2] / mn));
  const double zn = pr[14] / (pr[10] - 1.0), zf = pr[14] / (pr[10] + 1.0);
  const float ppx = (float)(zf / (zf - zn)), ppy = (float)(zf * zn / (zn - zf));
  const size_t n_pix = (size_t)p->width * p->height;
  memset(hit_pos, 0, n_pix * 16);
  memset(hit_normal, 0, n_pix * 16);
  std::vector<float> depth(n_pix, 1.0f), depth2(n_pix, 1.0f), fbo(n_pix * 3), near_pt(n_pix * 3);
  if (cv_pos) { memset(cv_pos, 0, n_pix * 16); memset(cv_normal, 0, n_pix * 16); }
  for (uint32_t y = 0; y < p->height; y++)
    for (uint32_t x = 0; x < p->width; x++) {
      float nx = ((float)x + 0.5f) / (float)p->width * 2.0f - 1.0f;
      float ny = ((float)y + 0.5f) / (float)p->height * 2.0f - 1.0f;
      v4 nr = xform4(inv_proj, nx, ny, -1.0f, 1.0f);
      size_t i = (size_t)y * p->width + x;
      near_pt[3 * i] = nr.x / nr.w; near_pt[3 * i + 1] = nr.y / nr.w; near_pt[3 * i + 2] = nr.z / nr.w;
      for (int k = 0; k < 3; k++) fbo[3 * i + k] = half_round(near_pt[3 * i + k]);
    }
  uint64_t samples = 0;
  const v4 o4 = xform4(imv, 0.0f, 0.0f, 0.0f, 1.0f);
  for (uint32_t bi = 0; bi < n_bricks; bi++) {
    const orc_classic_brick& b = list[bi];
    if (b.empty) continue;
    btex T; T.data = brick_data[bi]; T.dtype = p->dtype; T.nearest = p->nearest;
    for (int i = 0; i < 3; i++) T.n[i] = b.n_vox[i];
    const v3 c = V3(b.center[0], b.center[1], b.center[2]), e = V3(b.ext[0], b.ext[1], b.ext[2]);
    const v3 pmin = sub3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f)), pmax = add3(c, V3(e.x / 2.0f, e.y / 2.0f, e.z / 2.0f));
    const v3 tmin = V3(b.tex_min[0], b.tex_min[1], b.tex_min[2]), tmax = V3(b.tex_max[0], b.tex_max[1], b.tex_max[2]);
    const v3 tsc = div3(sub3(tmin, tmax), sub3(pmin, pmax));
    const v3 vstep = V3(1.0f / (float)b.n_vox[0], 1.0f / (float)b.n_vox[1], 1.0f / (float)b.n_vox[2]);
    const float ray_step = min3(scl3(mul3(e, vstep), 0.5f * 1.0f / p->sample_rate_modifier));
    const float lo[3] = {pmin.x, pmin.y, pmin.z}, hi[3] = {pmax.x, pmax.y, pmax.z};
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads < 1 ? 1 : n_threads) reduction(+ : samples)
    for (int64_t py = 0; py < (int64_t)p->height; py++)
      for (uint32_t px = 0; px < p->width; px++) {
        const size_t i = (size_t)py * p->width + px;
        const v3 pn = V3(near_pt[3 * i], near_pt[3 * i + 1], near_pt[3 * i + 2]);
        const v4 n4 = xform4(imv, pn.x, pn.y, pn.z, 1.0f);
        const float o[3] = {o4.x, o4.y, o4.z}, d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
        float s_in = -INFINITY, s_out = INFINITY;
        bool miss = false;
        for (int k = 0; k < 3; k++) {
          if (d[k] == 0.0f) { if (o[k] < lo[k] || o[k] > hi[k]) miss = true; continue; }
          float t0 = (lo[k] - o[k]) / d[k], t1 = (hi[k] - o[k]) / d[k];
          s_in = fmaxf(s_in, fminf(t0, t1));
          s_out = fminf(s_out, fmaxf(t0, t1));
        }
        if (miss || !(s_out > fmaxf(s_in, 1.0f))) continue;
        if (s_in > 1.0f) {
          const v3 fe = scl3(pn, s_in);
          fbo[3 * i] = half_round(fe.x); fbo[3 * i + 1] = half_round(fe.y); fbo[3 * i + 2] = half_round(fe.z);
        }
        const v3 entry = V3(fbo[3 * i], fbo[3 * i + 1], fbo[3 * i + 2]);
        const v3 exit_ = scl3(pn, s_out);
        auto to_tex = [&](v3 q) {
          v4 w = xform4(imv, q.x, q.y, q.z, 1.0f);
          return add3(mul3(sub3(V3(w.x, w.y, w.z), pmax), tsc), tmax);
        };
        const v3 et = to_tex(entry), xt = to_tex(exit_);
        auto normal_at = [&](v3 q) {                            /* ComputeNormal, Volume3D.glsl:43-60 */
          const float xp = T.sample(q, +1, 0, 0, norm), xm = T.sample(q, -1, 0, 0, norm);
          const float yp = T.sample(q, 0, -1, 0, norm), ym = T.sample(q, 0, +1, 0, norm);
          const float zp = T.sample(q, 0, 0, +1, norm), zm = T.sample(q, 0, 0, -1, norm);
          const v3 g = V3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
          const v3 gs = mul3(g, dscale);
          const float* m = imv;
          v3 nr = V3(m[0] * gs.x + m[1] * gs.y + m[2] * gs.z, m[4] * gs.x + m[5] * gs.y + m[6] * gs.z,
                     m[8] * gs.x + m[9] * gs.y + m[10] * gs.z);
          const float l = len3(nr);
          if (l > 0.0f) nr = scl3(nr, 1.0f / l);
          return nr;
        };
        /* march + RefineIsosurface from (e_eye, e_tex); returns false for `discard` */
        auto first_hit = [&](v3 e_eye, v3 e_tex, float iso, v3& hp, float& f, v3& hit_tex) {
          const float len = len3(sub3(exit_, e_eye));
          const float len_tex = len3(sub3(xt, e_tex));
          const float nsteps = len / ray_step;
          const int count = (int)nsteps + 1;
          const v3 inc_tex = V3((xt.x - e_tex.x) / nsteps, (xt.y - e_tex.y) / nsteps, (xt.z - e_tex.z) / nsteps);
          v3 ct = e_tex;
          bool hit = false;
          for (int s = 0; s < count; s++) {
            samples++;
            if (T.sample(ct, 0, 0, 0, norm) >= iso) { hit = true; break; }
            ct = add3(ct, inc_tex);
          }
          if (!hit) return false;
          v3 rd = V3(inc_tex.x / 2.0f, inc_tex.y / 2.0f, inc_tex.z / 2.0f);   /* RefineIsosurface */
          ct = sub3(ct, rd);
          for (int k = 0; k < 5; k++) {
            rd = V3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
            samples++;
            if (T.sample(ct, 0, 0, 0, norm) >= iso) ct = sub3(ct, rd); else ct = add3(ct, rd);
          }
          f = len3(sub3(ct, e_tex)) / len_tex;
          const float omf = 1.0f - f;
          hp = add3(scl3(e_eye, omf), scl3(exit_, f));
          hit_tex = ct;
          return true;
        };
        float* hp_o = hit_pos + 4 * i; float* hn_o = hit_normal + 4 * i;
        {                                                       /* GLRaycaster-ISO-FS.glsl into m_pFBOIsoHit */
          v3 hp, ht; float f;
          if (first_hit(entry, et, p->isoval, hp, f, ht)) {
            float dz = ppx + (ppy / -hp.z);
            dz = fminf(fmaxf(dz, 0.0f), 1.0f);                  /* GL clamps the written depth to the depth range */
            if (dz < depth[i]) {                                /* DF_LESS */
              depth[i] = dz;
              const v3 nr = normal_at(ht);
              hp_o[0] = hp.x; hp_o[1] = hp.y; hp_o[2] = hp.z; hp_o[3] = f;
              hn_o[0] = nr.x; hn_o[1] = nr.y; hn_o[2] = nr.z; hn_o[3] = (float)bi;
            }
          }
        }
        if (cv_pos) {                                           /* GLRaycaster-ISO-CV-FS.glsl into m_pFBOCVHit */
          v3 e2 = entry, et2 = et;
          if ((float)bi == hn_o[3]) {                           /* the kept first-pass hit lies in this brick: resume there */
            const float fl = hp_o[3], om = 1.0f - fl;
            e2 = add3(scl3(entry, om), scl3(entry, fl));        /* sic: the shader blends the entry with itself (:68) */
            et2 = add3(scl3(et, om), scl3(xt, fl));
          }
          v3 hp, ht; float f;
          if (first_hit(e2, et2, cv_isoval, hp, f, ht)) {
            float dz = ppx + (ppy / -exit_.z);                  /* depth of the ray EXIT (:99) */
            dz = fminf(fmaxf(dz, 0.0f), 1.0f);
            if (dz < depth2[i]) {
              depth2[i] = dz;
              const v3 nr = normal_at(ht);
              float* q = cv_pos + 4 * i; float* qn = cv_normal + 4 * i;
              q[0] = hp.x; q[1] = hp.y; q[2] = hp.z; q[3] = f;
              qn[0] = nr.x; qn[1] = nr.y; qn[2] = nr.z; qn[3] = (float)bi;
            }
          }
        }
      }
  }
  if (stats) { memset(stats, 0, sizeof(*stats)); stats->samples = samples; }
}


void orc_classic_iso_render(const orc_render_params* p, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                            const void* const* brick_data, float* hit_pos, float* hit_normal, orc_render_stats* stats,
                            int n_threads) {
  classic_iso_impl(p, lod, list, n_bricks, brick_data, hit_pos, hit_normal, 0.0f, nullptr, nullptr, stats, n_threads);
}

/* ClearView (m_bDoClearView, GLRaycaster.cpp:429-444): right after a brick's first pass the same back faces run
 * GLRaycaster-ISO-CV-FS.glsl:56-105 with the focus isovalue (GetNormalizedCVIsovalue) into m_pFBOCVHit -- the ray
 * resumes at the kept first-pass hit when that hit lies in this brick (texLastHit / texLastHitPos), and the depth
 * written is that of the ray exit, so the first brick along the ray with a focus hit stays. */
void orc_classic_cv_render(const orc_render_params* p, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                           const void* const* brick_data, float cv_isoval, float* hit_pos, float* hit_normal,
                           float* cv_pos, float* cv_normal, orc_render_stats* stats, int n_threads) {
  classic_iso_impl(p, lod, list, n_bricks, brick_data, hit_pos, hit_normal, cv_isoval, cv_pos, cv_normal, stats, n_threads);
}

/* Compose-CV-FS.glsl:57-98 over the four hit targets; light colours and parameters GLRenderer.cpp:2772-2795:
 * cv_param = (m_fCVSize, m_fCVContextScale, m_fCVBorderScale), pick = m_vCVPos * modelView (eye space).
 * The targets are GL_NEAREST / clamped; neighbours for the curvature estimate at +-1 pixel. */
void orc_cv_compose(const orc_render_params* p, const float* hit_pos, const float* hit_normal, const float* cv_pos,
                    const float* cv_normal, const float cv_color[3], const float cv_param[3], const float pick[3],
                    float* rgba) {
  const int w = (int)p->width, h = (int)p->height;
  const v3 a = V3(p->ambient[0] * p->ambient[3], p->ambient[1] * p->ambient[3], p->ambient[2] * p->ambient[3]);
  const v3 dd = V3(p->diffuse[0] * p->diffuse[3], p->diffuse[1] * p->diffuse[3], p->diffuse[2] * p->diffuse[3]);
  const v3 d1 = V3(dd.x * p->iso_color[0], dd.y * p->iso_color[1], dd.z * p->iso_color[2]);
  const v3 d2 = V3(dd.x * cv_color[0], dd.y * cv_color[1], dd.z * cv_color[2]);
  const v3 sp = V3(p->specular[0] * p->specular[3], p->specular[1] * p->specular[3], p->specular[2] * p->specular[3]);
  const v3 l = V3(p->light_dir[0], p->light_dir[1], p->light_dir[2]);
  auto light = [&](v3 pos, v3 nrm, v3 dif) {
    nrm.z = fabsf(nrm.z);
    v3 view = norm3(V3(0.0f - pos.x, 0.0f - pos.y, 0.0f - pos.z));
    float dn = dot3(nrm, view);
    v3 refl = norm3(sub3(view, scl3(nrm, 2.0f * dn)));
    float dl = fmaxf(fabsf(dot3(nrm, V3(-l.x, -l.y, -l.z))), 0.0f);
    float s8 = pow8(fmaxf(dot3(refl, l), 0.0f));
    return V3(clampf(a.x + dif.x * dl + sp.x * s8, 0.0f, 1.0f), clampf(a.y + dif.y * dl + sp.y * s8, 0.0f, 1.0f),
              clampf(a.z + dif.z * dl + sp.z * s8, 0.0f, 1.0f));
  };
  auto nrm_at = [&](int x, int y) {
    x = x < 0 ? 0 : x >= w ? w - 1 : x; y = y < 0 ? 0 : y >= h ? h - 1 : y;
    const float* q = hit_normal + 4 * ((size_t)y * w + x);
    return V3(q[0], q[1], q[2]);
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      float* o = rgba + 4 * i;
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      const float* hp = hit_pos + 4 * i;
      if (hp[3] == 0.0f) continue;
      const v3 pos = V3(hp[0], hp[1], hp[2]);
      const v3 n = nrm_at(x, y);
      const v3 ctx = light(pos, n, d1);
      auto absd = [&](v3 q) { return V3(fabsf(q.x - n.x), fabsf(q.y - n.y), fabsf(q.z - n.z)); };
      const v3 cs = add3(add3(add3(absd(nrm_at(x + 1, y)), absd(nrm_at(x - 1, y))), absd(nrm_at(x, y + 1))), absd(nrm_at(x, y - 1)));
      const float curv = len3(cs);
      const float dist_w = len3(sub3(pos, V3(pick[0], pick[1], pick[2]))) * cv_param[0];
      const float blend = clampf(fmaxf(curv * cv_param[1], clampf(dist_w, 0.0f, 1.0f)), 0.0f, 1.0f);
      v4 focus = {0, 0, 0, 0};
      const float* hp2 = cv_pos + 4 * i;
      if (hp2[3] != 0.0f) {
        const v3 f3_ = light(V3(hp2[0], hp2[1], hp2[2]), V3(cv_normal[4 * i], cv_normal[4 * i + 1], cv_normal[4 * i + 2]), d2);
        focus.x = f3_.x; focus.y = f3_.y; focus.z = f3_.z; focus.w = 1.0f;
      }
      const float omb = 1.0f - blend;
      float c[4] = {ctx.x * blend + focus.x * omb, ctx.y * blend + focus.y * omb, ctx.z * blend + focus.z * omb,
                    1.0f * blend + focus.w * omb};
      for (int k = 0; k < 4; k++) c[k] = fminf(fmaxf(c[k], 0.0f), 1.0f);
      const float border = 0.5f * (1.0f - fminf(fmaxf(fabsf(dist_w - 1.0f) * cv_param[2], 0.0f), 1.0f));
      o[0] = c[0] - border; o[1] = c[1] - border; o[2] = c[2] - border; o[3] = c[3];
    }
}

}  // extern "C"
