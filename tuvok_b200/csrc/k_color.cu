// k_color.cu -- the GridLeaper traversal kernel for COLOUR volumes (4 x 8 bit, GL_RGBA8 pool) on sm_100a.
//
// GLGridLeaper picks the "-color" method files when the dataset has four components (GLGridLeaper.cpp:766-795,
// AbstrRenderer::ColorData): the volume carries its own colour and the transfer function only maps ALPHA.  Replaces
// (reference file:line):
//   classification   Shaders/GLGridLeaper-Method-1D-color.glsl, -1D-L-color.glsl, -2D-color.glsl, -2D-L-color.glsl,
//                    -iso-color.glsl; GLGridLeaper-GradientTools.glsl:6-42 (ComputeGradient on .r, ComputeGradientAlpha on .a)
//   main()           Shaders/GLGridLeaper-blend.glsl:65-228, GLGridLeaper-iso.glsl:68-200 (unchanged for colour data)
//   samplePool4 / samplePoolAlpha   generated GLSL, Renderer/GL/GLVolumePool.cpp:637-651
// The shader text's quirks are kept: 1D-L takes its normal from ComputeNormal (the RED channel), the isosurface normal is
// ComputeAlphaNormal, which calls ComputeGradient (red channel again), while the hit test, the refinement and the 2D
// methods use alpha.
//
// HBM layout: the pool is slot-linear like every pool of this library; a colour voxel is one uchar4 (4 bytes, plain, no
// x-pair layout), so a trilinear fetch is 8 four-byte loads that deliver all four channels.  Per sample: 8 loads for the
// colour + 48 for the six gradient taps (the neighbouring footprints overlap; L1 serves the repeats).  This path follows the
// shader's nested loops one ray per thread -- it is the parity-first kernel for the colour methods, not the tuned flat loop
// of k_raycast.cu; page-table walk, LoD selection, miss reports, ray set-up and arithmetic contract are the shared code
// of tvk_traverse.cuh, so positions, requests and resume state are those of the scalar kernel.
#include "tvk_dev.h"

namespace tvk {
namespace {

#include "tvk_math.cuh"
#include "tvk_traverse.cuh"

// texel addresses of the 4x4x4 neighbourhood of a sample in the reference's VIRTUAL ATLAS (clamp-to-edge at the atlas
// border), split into (slot, texel-in-slot) -- the addressing of SlowFoot, valid for every ghost width and for GL_NEAREST
struct ColorFoot {
  const uchar4* c;
  uint64_t xo[4], yo[4], zo[4];
  float fx, fy, fz;
  bool nearest;

  __device__ __forceinline__ void fetch(const RayConsts& P, const uchar4* pool, f3 tc) {
    int X, Y, Z;
    nearest = P.nearest != 0;
    if (nearest) {
      X = (int)floorf(tc.x * P.pool_size_f[0]);
      Y = (int)floorf(tc.y * P.pool_size_f[1]);
      Z = (int)floorf(tc.z * P.pool_size_f[2]);
      fx = fy = fz = 0.0f;
    } else {
      const float ux = fmaf(tc.x, P.pool_size_f[0], -0.5f);
      const float uy = fmaf(tc.y, P.pool_size_f[1], -0.5f);
      const float uz = fmaf(tc.z, P.pool_size_f[2], -0.5f);
      const float x0 = floorf(ux), y0 = floorf(uy), z0 = floorf(uz);
      fx = ux - x0; fy = uy - y0; fz = uz - z0;
      X = (int)x0; Y = (int)y0; Z = (int)z0;
    }
    const uint32_t sy = P.total[0], sz = P.total[0] * P.total[1];
    c = pool;
    const int ax = (int)(P.capacity[0] * P.total[0]) - 1, ay = (int)(P.capacity[1] * P.total[1]) - 1,
              az = (int)(P.capacity[2] * P.total[2]) - 1;
    const uint64_t slot_y = (uint64_t)P.capacity[0] * P.slot_voxels, slot_z = slot_y * P.capacity[1];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t gx = (uint32_t)min(max(X - 1 + i, 0), ax), gy = (uint32_t)min(max(Y - 1 + i, 0), ay),
                     gz = (uint32_t)min(max(Z - 1 + i, 0), az);
      xo[i] = (uint64_t)(gx / P.total[0]) * P.slot_voxels + gx % P.total[0];
      yo[i] = (uint64_t)(gy / P.total[1]) * slot_y + (uint64_t)(gy % P.total[1]) * sy;
      zo[i] = (uint64_t)(gz / P.total[2]) * slot_z + (uint64_t)(gz % P.total[2]) * sz;
    }
  }
  // voxel at texel offset (i, j, k) in [-1, 2]^3 from the footprint origin
  __device__ __forceinline__ uchar4 v(int i, int j, int k) const { return __ldg(c + (xo[1 + i] + yo[1 + j] + zo[1 + k])); }
  template <int CH>
  static __device__ __forceinline__ float ch(uchar4 q) { return (float)(CH == 0 ? q.x : CH == 1 ? q.y : CH == 2 ? q.z : q.w); }
  // one channel of texture(volumePool, coords + (dx,dy,dz) texels): raw integers filtered, scaled once
  template <int CH>
  __device__ __forceinline__ float tap(const RayConsts& P, int dx, int dy, int dz) const {
    if (nearest) return ch<CH>(v(dx, dy, dz)) * P.norm;
    return tri(ch<CH>(v(dx, dy, dz)), ch<CH>(v(dx + 1, dy, dz)), ch<CH>(v(dx, dy + 1, dz)), ch<CH>(v(dx + 1, dy + 1, dz)),
               ch<CH>(v(dx, dy, dz + 1)), ch<CH>(v(dx + 1, dy, dz + 1)), ch<CH>(v(dx, dy + 1, dz + 1)),
               ch<CH>(v(dx + 1, dy + 1, dz + 1)), fx, fy, fz) * P.norm;
  }
  // samplePool4: the eight voxels are loaded once, every channel goes through the same lerp tree
  __device__ __forceinline__ f4 rgba(const RayConsts& P) const {
    f4 r;
    if (nearest) {
      const uchar4 q = v(0, 0, 0);
      r.x = (float)q.x * P.norm; r.y = (float)q.y * P.norm; r.z = (float)q.z * P.norm; r.w = (float)q.w * P.norm;
      return r;
    }
    const uchar4 a = v(0, 0, 0), b = v(1, 0, 0), cc = v(0, 1, 0), d = v(1, 1, 0), e = v(0, 0, 1), f = v(1, 0, 1), g = v(0, 1, 1),
                 h = v(1, 1, 1);
    r.x = tri((float)a.x, (float)b.x, (float)cc.x, (float)d.x, (float)e.x, (float)f.x, (float)g.x, (float)h.x, fx, fy, fz) * P.norm;
    r.y = tri((float)a.y, (float)b.y, (float)cc.y, (float)d.y, (float)e.y, (float)f.y, (float)g.y, (float)h.y, fx, fy, fz) * P.norm;
    r.z = tri((float)a.z, (float)b.z, (float)cc.z, (float)d.z, (float)e.z, (float)f.z, (float)g.z, (float)h.z, fx, fy, fz) * P.norm;
    r.w = tri((float)a.w, (float)b.w, (float)cc.w, (float)d.w, (float)e.w, (float)f.w, (float)g.w, (float)h.w, fx, fy, fz) * P.norm;
    return r;
  }
  // ComputeGradient (CH = 0) / ComputeGradientAlpha (CH = 3), GLGridLeaper-GradientTools.glsl:6-16 / :25-37 (the "Yp" tap
  // is fetched at -delta in both)
  template <int CH>
  __device__ __forceinline__ f3 gradient(const RayConsts& P) const {
    const float xp = tap<CH>(P, 1, 0, 0), xm = tap<CH>(P, -1, 0, 0);
    const float yp = tap<CH>(P, 0, -1, 0), ym = tap<CH>(P, 0, 1, 0);
    const float zp = tap<CH>(P, 0, 0, 1), zm = tap<CH>(P, 0, 0, -1);
    return F3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
  }
};

// MODE: 0 = 1D TF, 1 = 2D TF, 2 = isosurface.  One thread = one ray; a warp = an 8x4 pixel tile.
template <int MODE, bool LIT>
__global__ void __launch_bounds__(128) color_kernel(const __grid_constant__ RayConsts P) {
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t px = blockIdx.x * 16u + (wid & 1) * 8u + (lane & 7);
  const uint32_t py = blockIdx.y * 8u + (wid >> 1) * 4u + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const size_t pix = (size_t)py * P.width + px;
  constexpr bool ISO = MODE == 2;
  const uchar4* pool = (const uchar4*)P.pool;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  f4 entry4, exit4;
  if (!ray_setup(P, px, py, entry4, exit4, true)) {   // render targets are cleared where no back face is rasterised
    P.out0[pix] = zero4; P.out1[pix] = zero4; P.out2[pix] = zero4;
    if (ISO) P.out3[pix] = zero4;
    return;
  }
  f4 acc, resume_pos, resume_col = from4(zero4);
  f4 hit_pos = from4(zero4), hit_nrm = from4(zero4), resume_nrm = from4(zero4);
  bool done = false;
  if (P.first_pass) { resume_pos = entry4; acc = from4(zero4); }
  else { resume_pos = from4(P.ray_start[pix]); acc = from4(P.start_color[pix]); }
  if (!ISO) {
    resume_col = acc;
    if (resume_pos.w == 1000.0f) done = true;
  } else {
    if (floorf(resume_pos.w) == 1000.0f) done = true;
    else if (floorf(resume_pos.w) == 500.0f) {
      hit_pos = xform4(P.m2e, resume_pos.x, resume_pos.y, resume_pos.z, 1.0f);
      hit_pos.w = resume_pos.w - floorf(resume_pos.w) + 1.0f;
      hit_nrm = acc;   // rayStartNormal
      resume_nrm = hit_nrm;
      done = true;
    }
  }
  if (!done) {
    const f3 entry = F3(resume_pos.x, resume_pos.y, resume_pos.z);
    const float entry_depth = resume_pos.w;
    const f3 nexit = F3(exit4.x, exit4.y, exit4.z);
    const float exit_depth = exit4.w;
    const f3 dir = sub3(nexit, entry);
    const float ray_len = len3(dir);
    // TransformToPoolSpace
    f3 vdir = norm3(mul3(dir, F3(P.vol_f)));
    vdir = div3(vdir, F3(P.pool_size_f));
    const float den = 2.0f * P.sample_rate;
    vdir = F3(vdir.x / den, vdir.y / den, vdir.z / den);
    const float step = len3(vdir);
    float t = 0.0f;
    bool optimal = true;
    const float voxel_size = 0.125f / 2000.0f;
    f3 cur = entry;
    uint32_t lbx = 0, lby = 0, lbz = 0, lbl = 9999;
    const f3 dscale = F3(P.domain_scale), eye_m = F3(P.eye_m), la = F3(P.light_a), ld = F3(P.light_d), ls = F3(P.light_s),
             ldir = F3(P.light_dir_m);
    const f3 dv = F3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);   // BrickExit's 1.0/dir
    const f3 nudge = F3(voxel_size * dir.x / ray_len, voxel_size * dir.y / ray_len, voxel_size * dir.z / ray_len);
    bool terminated = false;
    if (ray_len > voxel_size) {
#pragma unroll 1
      for (uint32_t j = 0; j < 100 && !terminated; ++j) {
        const float cur_depth = entry_depth * (1.0f - t) + exit_depth * t;
        uint32_t lod = compute_lod(P, cur_depth);
        BrickRef b;
        const int ok = get_brick<false>(P, cur, lod, dir, dv, b);
        if (!ok && optimal) {
          optimal = false;
          resume_pos.x = cur.x; resume_pos.y = cur.y; resume_pos.z = cur.z; resume_pos.w = cur_depth;
          if (!ISO) resume_col = acc;
        }
        if (!b.empty && !(lbx == b.bx && lby == b.by && lbz == b.bz && lbl == b.bl)) {
          int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / step);
          const int s2 = (int)ceilf(len3(mul3(sub3(nexit, cur), b.scale)) / step);
          steps = min(steps, s2);
          const f3 inv = F3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
          f3 pc = b.pool_entry;
#pragma unroll 1
          for (int i = 0; i < steps; ++i) {
            ColorFoot foot;
            foot.fetch(P, pool, pc);
            if constexpr (!ISO) {
              // ComputeColorFromVolume (colour methods) + OpacityCorrectColor + UnderCompositing
              f4 col = foot.rgba(P);
              if constexpr (MODE == 0) {
                col.w = tf_lookup(P, col.w * P.trans_scale, 0.0f).w;
                if (LIT) {
                  f3 n = mul3(foot.gradient<0>(P), dscale);   // ComputeNormal: the red channel
                  const float l = len3(n);
                  if (l > 0.0f) n = scl3(n, 1.0f / l);
                  const f3 mp = mul3(sub3(pc, b.trans), inv);
                  const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              } else {
                const f3 g = foot.gradient<3>(P);              // ComputeGradientAlpha
                const float gm = len3(g);
                col.w = tf_lookup(P, col.w * P.trans_scale, 1.0f - gm * P.gradient_scale).w;
                if (LIT) {
                  const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                  const f3 n = mul3(dscale, gn);
                  const f3 mp = mul3(sub3(pc, b.trans), inv);
                  const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              }
              col.w = opacity_correct(P, col.w);
              const float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) { terminated = true; break; }
            } else {
              const f4 hcol = foot.rgba(P);                    // GetVolumeHit: the colour, hit when alpha >= isovalue
              if (hcol.w >= P.isoval) {
                // RefineIsosurface (on alpha)
                f3 rd = F3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
                pc = sub3(pc, rd);
#pragma unroll 1
                for (int k = 0; k < 5; k++) {
                  rd = F3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
                  foot.fetch(P, pool, pc);
                  if (foot.tap<3>(P, 0, 0, 0) >= P.isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
                }
                const f3 hp = mul3(sub3(pc, b.trans), inv);
                hit_pos = xform4(P.m2e, hp.x, hp.y, hp.z, 1.0f);
                hit_pos.w = hcol.x + 1.0f;                      // color.r + 1
                foot.fetch(P, pool, pc);
                f3 n = mul3(foot.gradient<0>(P), dscale);      // ComputeAlphaNormal -> ComputeGradient: the red channel
                const float l = len3(n);
                if (l > 0.0f) n = scl3(n, 1.0f / l);
                const float* m = P.mv_inv;                     // mModelViewIT * vec4(n, 0)
                hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
                hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
                hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
                hit_nrm.w = floorf(hcol.y * 512.0f) + hcol.z;   // floor(color.g*512)+color.b
                terminated = true;
                break;
              } else {
                hit_pos = from4(zero4);
              }
            }
            pc = add3(pc, vdir);
          }
          if (terminated) break;
          cur = mul3(sub3(pc, b.trans), inv);
        } else {
          cur = add3(b.norm_exit, nudge);
        }
        lbx = b.bx; lby = b.by; lbz = b.bz; lbl = b.bl;
        t = len3(sub3(entry, b.norm_exit)) / ray_len;
        if (t > 0.9999f) break;
      }
    }
    // TerminateRay
    if (!ISO) {
      if (optimal) { resume_pos.w = 1000.0f; resume_col = acc; }
    } else {
      if (optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
      resume_nrm = hit_nrm;
    }
  }
  if (!ISO) {
    P.out0[pix] = to4(acc); P.out1[pix] = to4(resume_col); P.out2[pix] = to4(resume_pos);
  } else {
    P.out0[pix] = to4(hit_pos); P.out1[pix] = to4(hit_nrm); P.out2[pix] = to4(resume_pos);
    P.out3[pix] = to4(resume_nrm);
  }
}

}  // namespace

// colour volumes (dtype TVK_RGBA8): sort-last shards, pipeline stages and counters are not built for this path
void launch_raycast_color(const RayConsts& rc, int mode, int lighting, cudaStream_t s) {
  const dim3 block(128), grid((rc.width + 15u) / 16u, (rc.height + 7u) / 8u);
  if (mode == TVK_RM_ISOSURFACE) color_kernel<2, false><<<grid, block, 0, s>>>(rc);
  else if (mode == TVK_RM_1DTRANS) {
    if (lighting) color_kernel<0, true><<<grid, block, 0, s>>>(rc); else color_kernel<0, false><<<grid, block, 0, s>>>(rc);
  } else {
    if (lighting) color_kernel<1, true><<<grid, block, 0, s>>>(rc); else color_kernel<1, false><<<grid, block, 0, s>>>(rc);
  }
}

}  // namespace tvk
