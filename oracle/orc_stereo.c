/*
 * orc_stereo.c -- ORACLE (test infrastructure): CPU restatement of the stereo side of the 3D view: the per-eye
 * view / projection matrices and the composition of the two eye images.  Pinned: the matrices against the
 * reference's own FLOATMATRIX4::BuildStereoLookAtAndProjection (oracle/_ref/ref_host `stereo`, bit for bit), the
 * composition against the reference's shader text executed per fragment (tests/glsl_ref.py build_stereo).
 *
 * Follows (reference file:line):
 *   matrices     FLOATMATRIX4::BuildStereoLookAtAndProjection   Basics/Vectors.h:1215-1248
 *                BuildLookAt :1250-1265, MatrixPerspectiveOffCenter :1286-1291, Translation :970-975,
 *                operator* :936-946; called by GLRenderer::ComputeViewAndProjection GLRenderer.cpp:904-910
 *   composition  GLRenderer::EndFrame                            Renderer/GL/GLRenderer.cpp:758-812
 *                Compose-Anaglyphs-FS.glsl:41-47, Compose-Scanline-FS.glsl:41-46, Compose-SBS-FS.glsl:40-45,
 *                Compose-AF-FS.glsl:40-43; both eye FBOs are GL_NEAREST (GLRenderer.cpp:1775-1785), the quad's
 *                texture coordinate of pixel (x, y) is ((x + 0.5) / w, (y + 0.5) / h) (FullscreenQuadRegion :688-713)
 * Arithmetic: IEEE fp32, sums left to right, dot products of the anaglyph shader as fmaf chains (the contract of
 * orc_render.c for dot()).
 */
#include "orc.h"
#include <math.h>
#include <string.h>

static void look_at(const float eye[3], const float at[3], const float up[3], float m[16]) {
  float F[3] = {at[0] - eye[0], at[1] - eye[1], at[2] - eye[2]};
  float U[3] = {up[0], up[1], up[2]};
  float S[3] = {F[1] * U[2] - F[2] * U[1], F[2] * U[0] - F[0] * U[2], F[0] * U[1] - F[1] * U[0]};
  U[0] = S[1] * F[2] - S[2] * F[1]; U[1] = S[2] * F[0] - S[0] * F[2]; U[2] = S[0] * F[1] - S[1] * F[0];
  float* v[3] = {F, U, S};
  for (int i = 0; i < 3; i++) {
    const float l = sqrtf(v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2]);
    if (l != 0.0f) { v[i][0] /= l; v[i][1] /= l; v[i][2] /= l; }
  }
  m[0] = S[0]; m[4] = S[1]; m[8] = S[2];    m[12] = -(S[0] * eye[0] + S[1] * eye[1] + S[2] * eye[2]);
  m[1] = U[0]; m[5] = U[1]; m[9] = U[2];    m[13] = -(U[0] * eye[0] + U[1] * eye[1] + U[2] * eye[2]);
  m[2] = -F[0]; m[6] = -F[1]; m[10] = -F[2]; m[14] = (F[0] * eye[0] + F[1] * eye[1] + F[2] * eye[2]);
  m[3] = 0.0f; m[7] = 0.0f; m[11] = 0.0f; m[15] = 1.0f;
}

static void off_center(float l, float r, float b, float t, float n, float f, float m[16]) {
  memset(m, 0, 64);
  m[0] = 2.0f * n / (r - l); m[8] = (r + l) / (r - l);
  m[5] = 2.0f * n / (t - b); m[9] = (t + b) / (t - b);
  m[10] = -(f + n) / (f - n); m[14] = -2.0f * (f * n) / (f - n);
  m[11] = -1.0f;
}

static void mul4(const float* a, const float* b, float* o) {
  float t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
      t[r * 4 + c] = a[r * 4 + 0] * b[0 + c] + a[r * 4 + 1] * b[4 + c] + a[r * 4 + 2] * b[8 + c] + a[r * 4 + 3] * b[12 + c];
  memcpy(o, t, 64);
}

void orc_stereo_view(const float eye[3], const float at[3], const float up[3], float fov_deg, float aspect, float z_near,
                     float z_far, float focal_length, float eye_dist, float view_l[16], float view_r[16],
                     float proj_l[16], float proj_r[16]) {
  const float radians = (float)(3.14159265358979323846 / 180.0) * fov_deg / 2;
  const float wd2 = z_near * (float)tan(radians);
  const float nfdl = z_near / focal_length;
  const float shift = eye_dist * nfdl;
  off_center(-aspect * wd2 + shift, aspect * wd2 + shift, -wd2, wd2, z_near, z_far, proj_l);
  off_center(-aspect * wd2 - shift, aspect * wd2 - shift, -wd2, wd2, z_near, z_far, proj_r);
  float v[16], t[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  look_at(eye, at, up, v);
  t[12] = eye_dist;
  mul4(t, v, view_l);
  t[12] = -eye_dist;
  mul4(t, v, view_r);
}

static const float* fetch(const float* img, uint32_t w, uint32_t h, float s, float t) {   /* texture2D, GL_NEAREST, clamp */
  int i = (int)floorf(s * (float)w), j = (int)floorf(t * (float)h);
  i = i < 0 ? 0 : i >= (int)w ? (int)w - 1 : i;
  j = j < 0 ? 0 : j >= (int)h ? (int)h - 1 : j;
  return img + 4 * ((size_t)j * w + i);
}

/* mode: 0 SM_RB, 1 SM_SCANLINE, 2 SM_SBS, 3 SM_AF (AbstrRenderer.h:128-134).  left / right: the two eye images
 * m_pFBO3DImageNext[0 / 1] (w*h RGBA32F); eye_swap binds them the other way round (GLRenderer.cpp:773-779). */
void orc_stereo_compose(int mode, const float* left, const float* right, uint32_t w, uint32_t h, int eye_swap,
                        int alternating_frame_id, float split_coord, float* out) {
  const float* L = eye_swap ? right : left;
  const float* R = eye_swap ? left : right;
  for (uint32_t y = 0; y < h; y++)
    for (uint32_t x = 0; x < w; x++) {
      const float s = ((float)x + 0.5f) / (float)w, t = ((float)y + 0.5f) / (float)h;
      float* o = out + 4 * ((size_t)y * w + x);
      if (mode == 0) {
        const float* a = fetch(L, w, h, s, t);
        const float* b = fetch(R, w, h, s, t);
        const float gl = fmaf(a[2], 0.11f, fmaf(a[1], 0.59f, a[0] * 0.3f));
        const float gr = fmaf(b[2], 0.11f, fmaf(b[1], 0.59f, b[0] * 0.3f));
        o[0] = gl; o[1] = gr * 0.5f; o[2] = gr; o[3] = fmaxf(a[3], b[3]);
      } else if (mode == 1) {
        const float line = floorf(t * (float)h);
        const float* a = (line / 2.0f == floorf(line / 2.0f)) ? fetch(L, w, h, s, t) : fetch(R, w, h, s, t);
        memcpy(o, a, 16);
      } else if (mode == 2) {
        const float* a = (s < split_coord) ? fetch(L, w, h, s * 2.0f, t) : fetch(R, w, h, (s - split_coord) * 2.0f, t);
        memcpy(o, a, 16);
      } else {
        memcpy(o, alternating_frame_id == 0 ? fetch(L, w, h, s, t) : fetch(R, w, h, s, t), 16);
      }
    }
}
