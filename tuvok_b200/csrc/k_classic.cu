// k_classic.cu -- the CLASSIC per-brick raycaster (GLRaycaster) as one sm_100a kernel.
//
// The reference renders one LoD brick by brick: CPU-sorted brick list, per brick a front-face pass into an
// RGBA16F ray-entry FBO and a back-face pass that marches the brick's own 3D texture, GL blending
// `dst += (1 - dst.a) * src` between bricks and a glFinish after every brick
// (Renderer/GL/GLRaycaster.cpp:348-478, Renderer/GL/GLRenderer.cpp:151-153,2663-2748).  Here one thread is
// one pixel ray: it walks the LoD's brick grid front to back (cell to cell across the shared box planes --
// for a regular grid that IS the depth-sorted order restricted to the bricks the ray meets), and for every
// listed, non-empty brick does what the brick's two GL passes do for that pixel:
//   ray entry/exit          analytic slab test against the brick's world box center +- extension/2
//                           (RenderBox, GLRaycaster.cpp:302-345); the entry goes through the FBO semantics:
//                           half-precision rounding (GLRaycaster.cpp:97) and "keeps its previous content where
//                           no front face is visible" (near plane first, Render3DPreLoop :348-381)
//   eye -> texture          ComputeEyeToTextureMatrix (GLRaycaster.cpp:589-612)
//   march                   GLRaycaster-1D-FS.glsl:51-82, -1D-light-FS.glsl:84-134, -2D-FS.glsl:53-94,
//                           -2D-light-FS.glsl:61-117 with VRender1D(.Lit).glsl, Volume3D.glsl:39-60,
//                           lighting.glsl:33-50 (eye at the origin), Compositing.glsl:33-38
//   per-brick step / opacity exponent   SetBrickDepShaderVars (GLRaycaster.cpp:239-300)
// MODE 2 is the HQ MIP frame of the 2D windows (GLRenderer.cpp:1183-1253, GLRaycaster::RenderHQMIPInLoop
// GLRaycaster.cpp:494-530): the same two passes per brick, but the back-face pass runs
// GLRaycaster-MIP-Rot-FS.glsl:47-77 (maximum of texture3D(texVolume).x along the ray), bricks are blended with
// BE_MAX (order-free, so the cell walk needs no sorting argument at all), and Transfer-MIP-FS.glsl:43-52
// maps the blended maximum through the 1D transfer function (opacity ignored) -- fused into the same thread.
// MODE 3 is the RM_ISOSURFACE branch of Render3DInLoop (GLRaycaster.cpp:383-446): the back-face pass runs
// GLRaycaster-ISO-FS.glsl:56-105 (first sample >= fIsoval, RefineIsosurface.glsl:37-52, eye-space hit position
// interpolated between ray entry and exit, ComputeNormal, gl_FragDepth) into the two iso-hit targets under the
// base state's depth test DF_LESS (GLRenderer.cpp:139), so the nearest hit of all bricks stays; the image is then
// composed by iso_compose_kernel (Compose-FS.glsl) like the GridLeaper isosurface frame.
// The bricks live in the same slot-linear pool as the GridLeaper path (the reference keeps one 3D texture per
// brick in GPUMemMan's LRU cache, GPUMemMan.cpp:846-996); the per-brick table maps a brick of the LoD to its
// slot.  Same arithmetic contract as k_raycast.cu (-fmad=false, explicit fmaf in lerps / dots / compositing).
#include <cuda_fp16.h>
#include "tvk_dev.h"

namespace tvk {
namespace {

#include "tvk_math.cuh"

__device__ __forceinline__ float half_round(float v) { return __half2float(__float2half_rn(v)); }

// texture3D(texVolume, tc) of ONE brick: GL_LINEAR / GL_NEAREST, clamp-to-edge on the brick's own size,
// voxels at slot strides.  Gradient taps sit +-1 texel from the centre and share its filter fractions.
// Interior samples (the whole footprint inside the brick: always, apart from the first / last samples of a segment,
// when the ghost layer is >= 2 voxels) need no clamping: one centre address + uniform strides, and with GRAD the 32
// distinct voxels of the 7 overlapping footprints are loaded once (as in k_raycast.cu).  Samples whose footprint
// touches the brick border take the clamped path; both paths read the same voxels, so the results are identical.
template <typename T, bool GRAD>
struct BrickTex {
  typedef typename PairOf<T>::W W;   // pool element: the x-pair (voxel x, voxel x+1), k_pool.cu; this kernel reads the first half
  const W* base;
  const W* c;            // interior: element (X, Y, Z)
  uint32_t xo[4], yo[4], zo[4];
  int sy, sz;
  float fx, fy, fz, norm;
  bool nearest, interior;
  __device__ __forceinline__ void set(const W* b, const uint32_t n[3], uint32_t sy_, uint32_t sz_, f3 tc, bool nn, float nrm) {
    base = b; nearest = nn; norm = nrm; sy = (int)sy_; sz = (int)sz_;
    int X, Y, Z;
    if (nn) {
      X = (int)floorf(tc.x * (float)n[0]); Y = (int)floorf(tc.y * (float)n[1]); Z = (int)floorf(tc.z * (float)n[2]);
      fx = fy = fz = 0.0f;
    } else {
      const float ux = fmaf(tc.x, (float)n[0], -0.5f), uy = fmaf(tc.y, (float)n[1], -0.5f), uz = fmaf(tc.z, (float)n[2], -0.5f);
      const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
      fx = ux - x0; fy = uy - y0; fz = uz - z0;
      X = (int)x0; Y = (int)y0; Z = (int)z0;
    }
    const int lo = GRAD ? 1 : 0, hi = GRAD ? 3 : 2;   // footprint [X-lo, X+hi-1] must lie in [0, n-1]
    interior = !nn && X >= lo && Y >= lo && Z >= lo && X <= (int)n[0] - hi && Y <= (int)n[1] - hi && Z <= (int)n[2] - hi;
    if (interior) {
      c = b + (X + Y * sy + Z * sz);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        xo[i] = (uint32_t)min(max(X - 1 + i, 0), (int)n[0] - 1);
        yo[i] = (uint32_t)min(max(Y - 1 + i, 0), (int)n[1] - 1) * sy_;
        zo[i] = (uint32_t)min(max(Z - 1 + i, 0), (int)n[2] - 1) * sz_;
      }
    }
  }
  __device__ __forceinline__ float v(int i, int j, int k) const { return first_voxel<T>(__ldg(base + (xo[1 + i] + yo[1 + j] + zo[1 + k]))); }
  __device__ __forceinline__ float vi(int i, int j, int k) const { return first_voxel<T>(__ldg(c + (i + j * sy + k * sz))); }
  __device__ __forceinline__ float tap(int dx, int dy, int dz) const {
    if (nearest) return v(dx, dy, dz) * norm;
    return tri(v(dx, dy, dz), v(dx + 1, dy, dz), v(dx, dy + 1, dz), v(dx + 1, dy + 1, dz), v(dx, dy, dz + 1),
               v(dx + 1, dy, dz + 1), v(dx, dy + 1, dz + 1), v(dx + 1, dy + 1, dz + 1), fx, fy, fz) * norm;
  }
  // texture3D(texVolume, tc).x
  __device__ __forceinline__ float centre() const {
    if (!interior) return tap(0, 0, 0);
    return tri(vi(0, 0, 0), vi(1, 0, 0), vi(0, 1, 0), vi(1, 1, 0), vi(0, 0, 1), vi(1, 0, 1), vi(0, 1, 1), vi(1, 1, 1), fx, fy, fz) * norm;
  }
  // ComputeGradient (Volume3D.glsl:43-53; the "Yp" tap is fetched at -delta)
  __device__ __forceinline__ f3 gradient() const {
    const float xp = tap(1, 0, 0), xm = tap(-1, 0, 0);
    const float yp = tap(0, -1, 0), ym = tap(0, 1, 0);
    const float zp = tap(0, 0, 1), zm = tap(0, 0, -1);
    return F3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
  }
  // centre value + gradient; interior samples load each of the 32 voxels of the 7 footprints once
  __device__ __forceinline__ void centre_and_gradient(float& data, f3& grad) const {
    if (!interior) { data = tap(0, 0, 0); grad = gradient(); return; }
    float cc[2][2][2];   // [z][y][x] centre block
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 2; i++) cc[k][j][i] = vi(i, j, k);
    float xl[2][2], xh[2][2], yl[2][2], yh[2][2], zl[2][2], zh[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++) {
        xl[a][b] = vi(-1, b, a);   // [z][y]
        xh[a][b] = vi(2, b, a);
        yl[a][b] = vi(b, -1, a);   // [z][x]
        yh[a][b] = vi(b, 2, a);
        zl[a][b] = vi(b, a, -1);   // [y][x]
        zh[a][b] = vi(b, a, 2);
      }
    const float n = norm;
    data = tri(cc[0][0][0], cc[0][0][1], cc[0][1][0], cc[0][1][1], cc[1][0][0], cc[1][0][1], cc[1][1][0], cc[1][1][1], fx, fy, fz) * n;
    const float xp = tri(cc[0][0][1], xh[0][0], cc[0][1][1], xh[0][1], cc[1][0][1], xh[1][0], cc[1][1][1], xh[1][1], fx, fy, fz) * n;
    const float xm = tri(xl[0][0], cc[0][0][0], xl[0][1], cc[0][1][0], xl[1][0], cc[1][0][0], xl[1][1], cc[1][1][0], fx, fy, fz) * n;
    const float ym = tri(cc[0][1][0], cc[0][1][1], yh[0][0], yh[0][1], cc[1][1][0], cc[1][1][1], yh[1][0], yh[1][1], fx, fy, fz) * n;   // +y
    const float yp = tri(yl[0][0], yl[0][1], cc[0][0][0], cc[0][0][1], yl[1][0], yl[1][1], cc[1][0][0], cc[1][0][1], fx, fy, fz) * n;   // -y
    const float zp = tri(cc[1][0][0], cc[1][0][1], cc[1][1][0], cc[1][1][1], zh[0][0], zh[0][1], zh[1][0], zh[1][1], fx, fy, fz) * n;
    const float zm = tri(zl[0][0], zl[0][1], zl[1][0], zl[1][1], cc[0][0][0], cc[0][0][1], cc[0][1][0], cc[0][1][1], fx, fy, fz) * n;
    grad = F3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
  }
};

__device__ __forceinline__ f4 tf_fetch(const ClassicConsts& P, float s, float t) {
  const int w = (int)P.tf_w, h = (int)P.tf_h;
  int ix = (int)floorf(s * (float)w);
  ix = min(max(ix, 0), w - 1);
  int iy = 0;
  if (h > 1) { iy = (int)floorf(t * (float)h); iy = min(max(iy, 0), h - 1); }
  return unorm8x4(__ldg(P.tf + ((uint32_t)iy * (uint32_t)w + (uint32_t)ix)));
}

// GLRaycaster-ISO-FS.glsl:72-98 / -ISO-CV-FS.glsl:72-98 for one fragment: march from (e_eye, e_tex) to the brick exit,
// first sample >= iso, RefineIsosurface.glsl:37-52, eye-space hit by interpolation.  false = `discard`.
template <typename T>
__device__ __forceinline__ bool iso_first_hit(const ClassicConsts& P, const typename PairOf<T>::W* vox, const uint32_t nv[3], uint32_t sy, uint32_t sz,
                                              f3 e_eye, f3 exit_, f3 e_tex, f3 xt, float ray_step, float iso, f3& hp, float& f,
                                              f3& hit_tex, unsigned long long& n_samples) {
  const float len = len3(sub3(exit_, e_eye));
  const float len_tex = len3(sub3(xt, e_tex));
  const float nsteps = len / ray_step;
  const int count = (int)nsteps + 1;
  const f3 inc_tex = F3((xt.x - e_tex.x) / nsteps, (xt.y - e_tex.y) / nsteps, (xt.z - e_tex.z) / nsteps);
  f3 ct = e_tex;
  bool hit = false;
#pragma unroll 1
  for (int s = 0; s < count; s++) {
    n_samples++;
    BrickTex<T, false> tx;
    tx.set(vox, nv, sy, sz, ct, P.nearest != 0, P.norm);
    if (tx.centre() >= iso) { hit = true; break; }
    ct = add3(ct, inc_tex);
  }
  if (!hit) return false;
  f3 rdir = F3(inc_tex.x / 2.0f, inc_tex.y / 2.0f, inc_tex.z / 2.0f);
  ct = sub3(ct, rdir);
#pragma unroll 1
  for (int k = 0; k < 5; k++) {
    rdir = F3(rdir.x / 2.0f, rdir.y / 2.0f, rdir.z / 2.0f);
    n_samples++;
    BrickTex<T, false> tx;
    tx.set(vox, nv, sy, sz, ct, P.nearest != 0, P.norm);
    if (tx.centre() >= iso) ct = sub3(ct, rdir); else ct = add3(ct, rdir);
  }
  f = len3(sub3(ct, e_tex)) / len_tex;
  const float omf = 1.0f - f;
  hp = add3(scl3(e_eye, omf), scl3(exit_, f));
  hit_tex = ct;
  return true;
}

// ComputeNormal (Volume3D.glsl:43-60): gl_NormalMatrix * (gradient * domainScale), safe-normalised
template <typename T>
__device__ __forceinline__ f3 iso_normal(const ClassicConsts& P, const typename PairOf<T>::W* vox, const uint32_t nv[3], uint32_t sy, uint32_t sz, f3 ct,
                                         f3 dscale) {
  BrickTex<T, true> tg;
  tg.set(vox, nv, sy, sz, ct, P.nearest != 0, P.norm);
  float unused;
  f3 g;
  tg.centre_and_gradient(unused, g);
  const f3 gs = mul3(g, dscale);
  const float* m = P.imv;
  f3 nrm = F3(m[0] * gs.x + m[1] * gs.y + m[2] * gs.z, m[4] * gs.x + m[5] * gs.y + m[6] * gs.z, m[8] * gs.x + m[9] * gs.y + m[10] * gs.z);
  const float l = len3(nrm);
  if (l > 0.0f) nrm = scl3(nrm, 1.0f / l);
  return nrm;
}

// MODE: 0 = 1D TF, 1 = 2D TF, 2 = HQ MIP, 3 = isosurface, 4 = isosurface + ClearView second pass (first hit, nearest of all bricks by the depth test)
// ORTHO (HQ MIP frames only): a parallel projection, m_bOrthoView (GLRenderer.cpp:1183-1197, GLRaycaster.cpp:486-489) -- the
// fragment's ray is a + s * b with b = far-plane point - near-plane point, a = near-plane point - b (s = 1 on the near plane,
// as for the perspective ray s * pn)
template <typename T, int MODE, bool LIT, bool ORTHO = false>
__global__ void __launch_bounds__(64) classic_kernel(const __grid_constant__ ClassicConsts P) {
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t px = blockIdx.x * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * 8 + wid * 4 + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const size_t pix = (size_t)py * P.width + px;
  const uint32_t S = P.axis_stride;
  f4 acc; acc.x = acc.y = acc.z = acc.w = 0.0f;   // MIP: x = blended maximum, w = coverage (the blended alpha)
  unsigned long long n_samples = 0;
  float best_depth = 1.0f, best_negz = 0.0f;          // MODE 3: the depth buffer texel (cleared to 1) and -z of its hit
  f4 hit_p, hit_n;                                    // MODE 3/4: the two iso-hit targets (cleared to 0)
  hit_p.x = hit_p.y = hit_p.z = hit_p.w = 0.0f; hit_n = hit_p;
  f4 cv_p = hit_p, cv_n = hit_p;                      // MODE 4: m_pFBOCVHit's two targets and its depth texel
  float cv_depth = 1.0f;
  bool first_done = false, cv_done = false;

  // the eye ray through the pixel centre: eye-space points s * pn, world-space o + s * d
  const float nx = ((float)px + 0.5f) / (float)P.width * 2.0f - 1.0f;
  const float ny = ((float)py + 0.5f) / (float)P.height * 2.0f - 1.0f;
  const f4 nr = xform4(P.inv_proj, nx, ny, -1.0f, 1.0f);
  const f3 pn = F3(nr.x / nr.w, nr.y / nr.w, nr.z / nr.w);
  f3 pa = F3(0.0f, 0.0f, 0.0f), pb = pn;
  if (ORTHO) {
    const f4 fr = xform4(P.inv_proj, nx, ny, 1.0f, 1.0f);
    const f3 pf = F3(fr.x / fr.w, fr.y / fr.w, fr.z / fr.w);
    pb = sub3(pf, pn);
    pa = sub3(pn, pb);
  }
  const f4 o4 = xform4(P.imv, pa.x, pa.y, pa.z, 1.0f);
  const f4 n4 = xform4(P.imv, pn.x, pn.y, pn.z, 1.0f);
  const float o[3] = {o4.x, o4.y, o4.z};
  const float d[3] = {n4.x - o4.x, n4.y - o4.y, n4.z - o4.z};
  const uint32_t lay[3] = {P.layout[0], P.layout[1], P.layout[2]};

  // enter the brick grid (its outer planes), not before the near plane
  float g_in = -INFINITY, g_out = INFINITY;
  bool hit = true;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float lo = P.plane[a * S], hi = P.plane[a * S + lay[a]];
    if (d[a] == 0.0f) { if (o[a] < lo || o[a] > hi) hit = false; continue; }
    const float t0 = (lo - o[a]) / d[a], t1 = (hi - o[a]) / d[a];
    g_in = fmaxf(g_in, fminf(t0, t1));
    g_out = fminf(g_out, fmaxf(t0, t1));
  }
  const float s_start = fmaxf(g_in, 1.0f);
  if (hit && g_out > s_start) {
    int cell[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {   // cell that holds the entry point (planes are monotone)
      const float w = o[a] + s_start * d[a];
      int i = 0;
      while (i + 1 < (int)lay[a] && w >= P.plane[a * S + i + 1]) i++;
      cell[a] = i;
    }
    f3 fbo = F3(half_round(pn.x), half_round(pn.y), half_round(pn.z));   // near-plane pass (Render3DPreLoop)
    const f3 dscale = F3(P.domain_scale), la = F3(P.light_a), ld = F3(P.light_d), ls = F3(P.light_s), ldir = F3(P.light_dir);
    typedef typename PairOf<T>::W W;
    const W* pool = (const W*)P.pool;
    const uint32_t sy = P.total[0], sz = P.total[0] * P.total[1];
    const int max_cells = (int)(lay[0] + lay[1] + lay[2]);
#pragma unroll 1
    for (int it = 0; it < max_cells; it++) {
      const uint32_t slot1 = __ldg(P.table + ((size_t)cell[2] * lay[1] + cell[1]) * lay[0] + cell[0]);
      if (slot1 != 0u) {
        float lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = P.pmin[a * S + cell[a]]; hi[a] = P.pmax[a * S + cell[a]]; }
        float s_in = -INFINITY, s_out = INFINITY;
        bool miss = false;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (d[a] == 0.0f) { if (o[a] < lo[a] || o[a] > hi[a]) miss = true; continue; }
          const float t0 = (lo[a] - o[a]) / d[a], t1 = (hi[a] - o[a]) / d[a];
          s_in = fmaxf(s_in, fminf(t0, t1));
          s_out = fminf(s_out, fmaxf(t0, t1));
        }
        if (!miss && s_out > fmaxf(s_in, 1.0f)) {
          if (s_in > 1.0f) {   // a visible front face overwrites the ray-entry FBO
            const f3 fe = ORTHO ? add3(pa, scl3(pb, s_in)) : scl3(pn, s_in);
            fbo = F3(half_round(fe.x), half_round(fe.y), half_round(fe.z));
          }
          const f3 entry = fbo, exit_ = ORTHO ? add3(pa, scl3(pb, s_out)) : scl3(pn, s_out);
          const f3 pmax = F3(hi[0], hi[1], hi[2]);
          const f3 tsc = F3(P.tsc[cell[0]], P.tsc[S + cell[1]], P.tsc[2 * S + cell[2]]);
          const f3 tmax = F3(P.tmax[cell[0]], P.tmax[S + cell[1]], P.tmax[2 * S + cell[2]]);
          const uint32_t nv[3] = {P.nvox[cell[0]], P.nvox[S + cell[1]], P.nvox[2 * S + cell[2]]};
          const float ray_step = fminf(P.rstep[cell[0]], fminf(P.rstep[S + cell[1]], P.rstep[2 * S + cell[2]]));
          const f4 we = xform4(P.imv, entry.x, entry.y, entry.z, 1.0f), wx = xform4(P.imv, exit_.x, exit_.y, exit_.z, 1.0f);
          const f3 et = add3(mul3(sub3(F3(we.x, we.y, we.z), pmax), tsc), tmax);
          const f3 xt = add3(mul3(sub3(F3(wx.x, wx.y, wx.z), pmax), tsc), tmax);
          f3 rd = sub3(exit_, entry);
          const float len = len3(rd);
          const float nsteps = len / ray_step;
          const int count = (int)nsteps + 1;
          const f3 inc_tex = F3((xt.x - et.x) / nsteps, (xt.y - et.y) / nsteps, (xt.z - et.z) / nsteps);
          rd = F3(rd.x / len, rd.y / len, rd.z / len);
          const f3 inc = scl3(rd, ray_step);
          const W* vox = pool + (uint64_t)(slot1 - 1u) * P.slot_voxels;
          if (MODE == 3 || MODE == 4) {
            // First pass (GLRaycaster-ISO-FS).  A fragment's hit lies between its ray entry and exit, so a brick whose
            // (half-rounded) entry is clearly behind the kept hit cannot pass the depth test -- and neither can any
            // brick after it on the ray.
            if (best_depth < 1.0f && -entry.z > best_negz * 1.002f) first_done = true;
            if (first_done && (MODE == 3 || cv_done)) break;
            if (!first_done) {
              f3 hp, ht; float f;
              if (iso_first_hit<T>(P, vox, nv, sy, sz, entry, exit_, et, xt, ray_step, P.isoval, hp, f, ht, n_samples)) {
                float dz = P.proj_param[0] + (P.proj_param[1] / -hp.z);
                dz = fminf(fmaxf(dz, 0.0f), 1.0f);       // depth-range clamp, then DF_LESS
                if (dz < best_depth) {
                  best_depth = dz; best_negz = -hp.z;
                  const f3 nrm = iso_normal<T>(P, vox, nv, sy, sz, ht, dscale);
                  hit_p.x = hp.x; hit_p.y = hp.y; hit_p.z = hp.z; hit_p.w = f;
                  hit_n.x = nrm.x; hit_n.y = nrm.y; hit_n.z = nrm.z;
                  hit_n.w = (float)__ldg(P.list_pos + ((size_t)cell[2] * lay[1] + cell[1]) * lay[0] + cell[0]);
                }
              }
            }
            if (MODE == 4 && !cv_done) {
              // Second pass of the same brick (GLRaycaster-ISO-CV-FS.glsl:56-105): resumes at the kept first-pass hit
              // when that lies in this brick; its depth is the ray EXIT's, so the first brick with a focus hit stays.
              const float tile = (float)__ldg(P.list_pos + ((size_t)cell[2] * lay[1] + cell[1]) * lay[0] + cell[0]);
              f3 e2 = entry, et2 = et;
              if (tile == hit_n.w) {
                const float fl = hit_p.w, om = 1.0f - fl;
                e2 = add3(scl3(entry, om), scl3(entry, fl));   // sic: the shader blends the entry with itself (:68)
                et2 = add3(scl3(et, om), scl3(xt, fl));
              }
              f3 hp, ht; float f;
              if (iso_first_hit<T>(P, vox, nv, sy, sz, e2, exit_, et2, xt, ray_step, P.cv_isoval, hp, f, ht, n_samples)) {
                float dz = P.proj_param[0] + (P.proj_param[1] / -exit_.z);
                dz = fminf(fmaxf(dz, 0.0f), 1.0f);
                if (dz < cv_depth) {
                  cv_depth = dz; cv_done = true;           // every later brick's exit is farther: it fails DF_LESS
                  const f3 nrm = iso_normal<T>(P, vox, nv, sy, sz, ht, dscale);
                  cv_p.x = hp.x; cv_p.y = hp.y; cv_p.z = hp.z; cv_p.w = f;
                  cv_n.x = nrm.x; cv_n.y = nrm.y; cv_n.z = nrm.z; cv_n.w = tile;
                }
              }
            }
          } else if (MODE == 2) {   // GLRaycaster-MIP-Rot-FS.glsl:64-76, then glBlendEquation(GL_MAX)
            float mx = 0.0f;
            f3 ct = et;
            // A segment whose two ends have interior footprints (with a margin far above the drift of the position
            // accumulation) has only interior samples: a branch-free loop, unrolled so that the loads of several
            // (independent) samples are in flight.  Same arithmetic as BrickTex::set + centre().
            const f3 nf = F3((float)nv[0], (float)nv[1], (float)nv[2]);
            const f3 ue = F3(fmaf(et.x, nf.x, -0.5f), fmaf(et.y, nf.y, -0.5f), fmaf(et.z, nf.z, -0.5f));
            const f3 ux = F3(fmaf(xt.x, nf.x, -0.5f), fmaf(xt.y, nf.y, -0.5f), fmaf(xt.z, nf.z, -0.5f));
            const bool all_interior = P.nearest == 0 &&
                fminf(ue.x, ux.x) >= 0.01f && fmaxf(ue.x, ux.x) <= nf.x - 1.01f &&
                fminf(ue.y, ux.y) >= 0.01f && fmaxf(ue.y, ux.y) <= nf.y - 1.01f &&
                fminf(ue.z, ux.z) >= 0.01f && fmaxf(ue.z, ux.z) <= nf.z - 1.01f;
            if (all_interior) {
              const int isy = (int)sy, isz = (int)sz;
#pragma unroll 4
              for (int s = 0; s < count; s++) {
                const float vx = fmaf(ct.x, nf.x, -0.5f), vy = fmaf(ct.y, nf.y, -0.5f), vz = fmaf(ct.z, nf.z, -0.5f);
                const float x0 = floorf(vx), y0 = floorf(vy), z0 = floorf(vz);
                const float fx = vx - x0, fy = vy - y0, fz = vz - z0;
                const W* c = vox + ((int)x0 + (int)y0 * isy + (int)z0 * isz);
                const float v = tri(first_voxel<T>(__ldg(c)), first_voxel<T>(__ldg(c + 1)), first_voxel<T>(__ldg(c + isy)),
                                    first_voxel<T>(__ldg(c + isy + 1)), first_voxel<T>(__ldg(c + isz)),
                                    first_voxel<T>(__ldg(c + isz + 1)), first_voxel<T>(__ldg(c + isz + isy)),
                                    first_voxel<T>(__ldg(c + isz + isy + 1)), fx, fy, fz) * P.norm;
                mx = fmaxf(mx, v);
                ct = add3(ct, inc_tex);
              }
            } else {
#pragma unroll 1
              for (int s = 0; s < count; s++) {
                BrickTex<T, false> tx;
                tx.set(vox, nv, sy, sz, ct, P.nearest != 0, P.norm);
                mx = fmaxf(mx, tx.centre());
                ct = add3(ct, inc_tex);
              }
            }
            n_samples += (unsigned long long)count;
            acc.x = fmaxf(acc.x, mx);
            acc.w = 1.0f;
          } else {
          f4 col; col.x = col.y = col.z = col.w = 0.0f;
          f3 ct = et, cp = entry;
#pragma unroll 1
          for (int s = 0; s < count; s++) {
            n_samples++;
            BrickTex<T, (MODE != 0 || LIT)> tx;
            tx.set(vox, nv, sy, sz, ct, P.nearest != 0, P.norm);
            f4 sc;
            if (MODE == 0 && !LIT) {
              const float data = tx.centre();
              sc = tf_fetch(P, data * P.trans_scale, 0.0f);
            } else {
              float data;
              f3 g;
              tx.centre_and_gradient(data, g);
              if (MODE == 0) sc = tf_fetch(P, data * P.trans_scale, 0.0f);
              else sc = tf_fetch(P, data * P.trans_scale, 1.0f - len3(g) * P.gradient_scale);
              if (LIT) {
                // ComputeNormal: gl_NormalMatrix * (gradient * domainScale), safe-normalised
                const f3 gs = mul3(g, dscale);
                const float* m = P.imv;
                f3 nrm = F3(m[0] * gs.x + m[1] * gs.y + m[2] * gs.z, m[4] * gs.x + m[5] * gs.y + m[6] * gs.z,
                            m[8] * gs.x + m[9] * gs.y + m[10] * gs.z);
                const float l = len3(nrm);
                if (l > 0.0f) nrm = scl3(nrm, 1.0f / l);
                f3 lit = lighting(F3(0.0f, 0.0f, 0.0f), cp, nrm, la, mul3(F3(sc.x, sc.y, sc.z), ld), ls, ldir);
                if (MODE == 1) lit = F3(clampf(lit.x, 0.0f, 1.0f), clampf(lit.y, 0.0f, 1.0f), clampf(lit.z, 0.0f, 1.0f));
                sc.x = lit.x; sc.y = lit.y; sc.z = lit.z;
              }
            }
            sc.w = P.step_scale == 1.0f ? sc.w : 1.0f - powf(1.0f - sc.w, P.step_scale);
            const float oma = 1.0f - col.w;   // UnderCompositing
            col.x = fmaf(sc.x * oma, sc.w, col.x); col.y = fmaf(sc.y * oma, sc.w, col.y);
            col.z = fmaf(sc.z * oma, sc.w, col.z); col.w = fmaf(sc.w, oma, col.w);
            if (col.w >= 0.99f) break;
            cp = add3(cp, inc);
            ct = add3(ct, inc_tex);
          }
          // GL blending ONE_MINUS_DST_ALPHA, ONE
          const float k = 1.0f - acc.w;
          acc.x = fmaf(k, col.x, acc.x); acc.y = fmaf(k, col.y, acc.y);
          acc.z = fmaf(k, col.z, acc.z); acc.w = fmaf(k, col.w, acc.w);
          }   // DVR modes
        }
      }
      // leave the cell through the nearest of its far planes
      float best = INFINITY; int ax = -1;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        if (d[a] == 0.0f) continue;
        const float pl = P.plane[a * S + cell[a] + (d[a] > 0.0f ? 1 : 0)];
        const float t = (pl - o[a]) / d[a];
        if (t < best) { best = t; ax = a; }
      }
      if (ax < 0) break;
      cell[ax] += d[ax] > 0.0f ? 1 : -1;
      if (cell[ax] < 0 || cell[ax] >= (int)lay[ax]) break;
    }
  }
  if (MODE == 3 || MODE == 4) {
    P.out[pix] = make_float4(hit_p.x, hit_p.y, hit_p.z, hit_p.w);
    P.out_nrm[pix] = make_float4(hit_n.x, hit_n.y, hit_n.z, hit_n.w);
    if (MODE == 4) {
      P.out_cv[pix] = make_float4(cv_p.x, cv_p.y, cv_p.z, cv_p.w);
      P.out_cv_nrm[pix] = make_float4(cv_n.x, cv_n.y, cv_n.z, cv_n.w);
    }
  } else if (MODE == 2) {   // Transfer-MIP-FS.glsl:43-52 (1D transfer function, opacity ignored; uncovered pixels black)
    if (P.out_max) P.out_max[pix] = make_float2(acc.x, acc.w);
    f4 t; t.x = t.y = t.z = 0.0f;
    if (acc.w > 0.5f) t = tf_fetch(P, acc.x * P.trans_scale, 0.0f);
    P.out[pix] = make_float4(t.x, t.y, t.z, 1.0f);
  } else {
    P.out[pix] = make_float4(acc.x, acc.y, acc.z, acc.w);
  }
  if (P.count) atomicAdd(P.counters, n_samples);
}

template <typename T>
void launch_t(const ClassicConsts& c, int mode, int lighting, cudaStream_t s) {
  const dim3 block(64), grid((c.width + 7) / 8, (c.height + 7) / 8);
  if (mode == TVK_CLASSIC_MIP) {
    if (c.ortho) classic_kernel<T, 2, false, true><<<grid, block, 0, s>>>(c);
    else classic_kernel<T, 2, false><<<grid, block, 0, s>>>(c);
  } else if (mode == TVK_RM_ISOSURFACE) {
    if (c.out_cv) classic_kernel<T, 4, false><<<grid, block, 0, s>>>(c);   // ClearView: second (focus) pass per brick
    else classic_kernel<T, 3, false><<<grid, block, 0, s>>>(c);
  } else if (mode == TVK_RM_1DTRANS) {
    if (lighting) classic_kernel<T, 0, true><<<grid, block, 0, s>>>(c);
    else classic_kernel<T, 0, false><<<grid, block, 0, s>>>(c);
  } else {
    if (lighting) classic_kernel<T, 1, true><<<grid, block, 0, s>>>(c);
    else classic_kernel<T, 1, false><<<grid, block, 0, s>>>(c);
  }
}

}  // namespace

void launch_classic(const ClassicConsts& c, int mode, int lighting, int dtype, cudaStream_t s) {
  switch (dtype) {
    case TVK_U8: launch_t<uint8_t>(c, mode, lighting, s); break;
    case TVK_U16: launch_t<uint16_t>(c, mode, lighting, s); break;
    default: launch_t<float>(c, mode, lighting, s); break;
  }
}

}  // namespace tvk
