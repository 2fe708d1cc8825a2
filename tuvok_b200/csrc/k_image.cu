// k_image.cu -- image-space kernels: deferred isosurface shading, the GL float->unorm8 read-back
// conversion, the stereo eye composition (Compose-{Anaglyphs,Scanline,SBS,AF}-FS.glsl, GLRenderer.cpp:758-812)
// and the over operator of the sort-last compositor.  HBM-bound streaming kernels:
// 128-bit accesses, grid sized to a multiple of the 148 SMs with a grid-stride loop.
// Replaces (reference file:line): Shaders/Compose-FS.glsl:49-76 + GLRenderer::ComposeSurfaceImage
// (GLRenderer.cpp:2763-2830); GLFrameCapture.cpp:72-85 (glReadPixels GL_UNSIGNED_BYTE);
// Compositing.glsl:33-38 / blend state GLRenderer.cpp:151-153.
// Compiled with -fmad=false (same arithmetic contract as k_raycast.cu).
#include <algorithm>
#include "tvk_dev.h"

namespace tvk {
namespace {

constexpr int kSMs = 148;

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float pow8(float x) { float a = x * x; float b = a * a; return b * b; }

// dot products follow the arithmetic contract: fma(z,z', fma(y,y', x*x'))
__device__ __forceinline__ float dot(float ax, float ay, float az, float bx, float by, float bz) {
  return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

struct ComposeConsts { float amb[3], dif[3], spe[3], ldir[3]; };

// COLOR: Compose-Color-FS.glsl:60-92 -- the diffuse colour is the hit's own colour, unpacked from the two alpha channels
// (r = pos.a - 1, g = floor(nrm.a / 2) / 256, b = fract(nrm.a)), times vLightDiffuse
template <bool COLOR>
__global__ void iso_compose_kernel(const float4* __restrict__ hit_pos, const float4* __restrict__ hit_nrm,
                                   float4* __restrict__ rgba, uint64_t n, const ComposeConsts C) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float4 hp = hit_pos[i];
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hp.w != 0.0f) {   // `discard` leaves the cleared texel
      const float4 hn = hit_nrm[i];
      const float nx = hn.x, ny = hn.y, nz = fabsf(hn.z);
      float vx = 0.0f - hp.x, vy = 0.0f - hp.y, vz = 0.0f - hp.z;
      float inv = 1.0f / sqrtf(dot(vx, vy, vz, vx, vy, vz));
      vx = vx * inv; vy = vy * inv; vz = vz * inv;
      const float dn = dot(nx, ny, nz, vx, vy, vz);
      const float k = 2.0f * dn;
      float rx = vx - nx * k, ry = vy - ny * k, rz = vz - nz * k;
      inv = 1.0f / sqrtf(dot(rx, ry, rz, rx, ry, rz));
      rx = rx * inv; ry = ry * inv; rz = rz * inv;
      const float dl = fmaxf(fabsf(dot(nx, ny, nz, -C.ldir[0], -C.ldir[1], -C.ldir[2])), 0.0f);
      const float sp = pow8(fmaxf(dot(rx, ry, rz, C.ldir[0], C.ldir[1], C.ldir[2]), 0.0f));
      float d0 = C.dif[0], d1 = C.dif[1], d2 = C.dif[2];
      if (COLOR) {
        d0 = (hp.w - 1.0f) * d0;
        d1 = (floorf(hn.w / 2.0f) / 256.0f) * d1;
        d2 = (hn.w - floorf(hn.w)) * d2;
      }
      o.x = clamp01(C.amb[0] + d0 * dl + C.spe[0] * sp);
      o.y = clamp01(C.amb[1] + d1 * dl + C.spe[1] * sp);
      o.z = clamp01(C.amb[2] + d2 * dl + C.spe[2] * sp);
      o.w = 1.0f;
    }
    rgba[i] = o;
  }
}

__device__ __forceinline__ unsigned char unorm8(float v) {
  v = v < 0.0f ? 0.0f : v > 1.0f ? 1.0f : v;
  if (v != v) v = 0.0f;
  return (unsigned char)(v * 255.0f + 0.5f);
}

__global__ void quantize_kernel(const float4* __restrict__ src, uchar4* __restrict__ dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uchar4(unorm8(v.x), unorm8(v.y), unorm8(v.z), unorm8(v.w));
  }
}

// front OVER back for one pixel, aware of early ray termination
__device__ __forceinline__ float4 over1(float4 f, float4 b) {
  // a front ray that terminated early (alpha > 0.99, GLGridLeaper-blend.glsl:180) hides everything behind it,
  // exactly as the single-GPU ray would have stopped there
  float oma = f.w > 0.99f ? 0.0f : 1.0f - f.w;
  // ... and a ray that would have crossed 0.99 INSIDE the back block stops there too: the back image (which was
  // accumulated without knowing the front alpha) is cut at the middle of the interval (0.99, 1.0] in which the
  // single-GPU ray ends, which bounds the alpha error by 0.005 (< 1.3/255)
  const float add = oma * b.w;
  if (oma > 0.0f && f.w + add > 0.995f) oma = oma * ((0.995f - f.w) / add);
  return make_float4(f.x + oma * b.x, f.y + oma * b.y, f.z + oma * b.z, f.w + oma * b.w);
}

__global__ void over_kernel(const float4* front, const float4* back, float4* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = over1(front[i], back[i]);
}

// Direct-send compositing of the sort-last frame: this rank owns one contiguous pixel slice and has the partial images
// of all n ranks for it (its own in place, the others received over NVLink).  The slices are folded FRONT TO BACK in
// the visibility order of the brick blocks -- ((s0 over s1) over s2) ... -- which is the order a single ray meets the
// blocks in, and the result is written as RGBA32F (parity tap) and as the RGBA8 the frame is read back as
// (GLFrameCapture conversion fused in: 4x less gather traffic).  Streaming: n x 16 B read, 20 B written per pixel.
__global__ void nway_over_kernel(const NWaySrc a, float4* __restrict__ out_f, uchar4* __restrict__ out8, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    float4 acc = a.src[0][i];
    for (int k = 1; k < a.n; k++) acc = over1(acc, a.src[k][i]);
    if (out_f) out_f[i] = acc;
    out8[i] = make_uchar4(unorm8(acc.x), unorm8(acc.y), unorm8(acc.z), unorm8(acc.w));
  }
}

// ---- peer-memory sort-last (tvk_dev.h SlPeer) -------------------------------------------------------------------
__device__ __forceinline__ uint32_t* sl_flag(const SlPeer& P, int owner, int kind, int from) {
  return P.flags[owner] + kind * TVK_MAX_RANKS + from;
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr long long kSlSpinCycles = 60000000000ll;   // ~30 s at 1.9 GHz (a peer may be paging bricks in for seconds); a peer that
                                                     // never arrives must not hang the GPU for good

// thread q waits until rank q's flag of `kind` in MY block has reached `frame`
__device__ __forceinline__ void sl_wait_all(const SlPeer& P, int kind, uint32_t frame, uint32_t* local) {
  if ((int)threadIdx.x < P.n) {
    const uint32_t* f = sl_flag(P, P.self, kind, (int)threadIdx.x);
    const long long t0 = clock64();
    while ((int32_t)(ld_flag(f) - frame) < 0) {
      if (clock64() - t0 > kSlSpinCycles) { atomicExch(&local[1], frame ? frame : 1u); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__global__ void sl_wait_kernel(const SlPeer P, int kind, uint32_t frame, uint32_t* local) { sl_wait_all(P, kind, frame, local); }

// thread q tells rank q: my flag of `kind` is now `frame` (everything this stream did before is visible first)
__global__ void sl_signal_kernel(const SlPeer P, int kind, uint32_t frame) {
  __threadfence_system();
  if ((int)threadIdx.x < P.n) st_flag(sl_flag(P, (int)threadIdx.x, kind, P.self), frame);
}

__global__ void nway_over_peer_kernel(const NWaySrc a, float4* __restrict__ out_f, uchar4* out8, uint64_t n, const SlPeer P,
                                      uint32_t frame, uint32_t* local) {
  sl_wait_all(P, TVK_SLF_READY, frame, local);          // every partial image of this frame is complete
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    float4 acc = a.src[0][i];
    for (int k = 1; k < a.n; k++) acc = over1(acc, a.src[k][i]);
    if (out_f) out_f[i] = acc;
    out8[i] = make_uchar4(unorm8(acc.x), unorm8(acc.y), unorm8(acc.z), unorm8(acc.w));   // rank 0's frame (peer store)
  }
  // the last block to finish tells the peers: I have read your images (they may render the next frame) and rank 0: my
  // RGBA8 slice has landed
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(&local[0], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) local[0] = 0;
    __threadfence_system();
    if ((int)threadIdx.x < P.n) st_flag(sl_flag(P, (int)threadIdx.x, TVK_SLF_CONSUMED, P.self), frame);
    if (threadIdx.x == 0) st_flag(sl_flag(P, 0, TVK_SLF_GATHERED, P.self), frame);
  }
}

// Compose-CV-FS.glsl:57-98 (ClearView): context surface shaded with the isosurface colour, focus surface (second hit
// targets) with the ClearView colour, blended by the distance to the pick position and by a curvature estimate from the
// four neighbouring normals (targets are GL_NEAREST / clamped), darkened along the lens border.
struct CvConsts { float amb[3], dif[3], dif2[3], spe[3], ldir[3], param[3], pick[3]; };

__device__ __forceinline__ void cv_light(const CvConsts& C, const float* dif, float px, float py, float pz, float nx, float ny, float nz,
                                         float& r, float& g, float& b) {
  nz = fabsf(nz);
  float vx = 0.0f - px, vy = 0.0f - py, vz = 0.0f - pz;
  float inv = 1.0f / sqrtf(dot(vx, vy, vz, vx, vy, vz));
  vx = vx * inv; vy = vy * inv; vz = vz * inv;
  const float dn = dot(nx, ny, nz, vx, vy, vz);
  const float k = 2.0f * dn;
  float rx = vx - nx * k, ry = vy - ny * k, rz = vz - nz * k;
  inv = 1.0f / sqrtf(dot(rx, ry, rz, rx, ry, rz));
  rx = rx * inv; ry = ry * inv; rz = rz * inv;
  const float dl = fmaxf(fabsf(dot(nx, ny, nz, -C.ldir[0], -C.ldir[1], -C.ldir[2])), 0.0f);
  const float sp = pow8(fmaxf(dot(rx, ry, rz, C.ldir[0], C.ldir[1], C.ldir[2]), 0.0f));
  r = clamp01(C.amb[0] + dif[0] * dl + C.spe[0] * sp);
  g = clamp01(C.amb[1] + dif[1] * dl + C.spe[1] * sp);
  b = clamp01(C.amb[2] + dif[2] * dl + C.spe[2] * sp);
}

__global__ void cv_compose_kernel(const float4* __restrict__ hit_pos, const float4* __restrict__ hit_nrm,
                                  const float4* __restrict__ cv_pos, const float4* __restrict__ cv_nrm, float4* __restrict__ rgba,
                                  int w, int h, const CvConsts C) {
  const uint64_t n = (uint64_t)w * h;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (uint64_t)y * w);
    const float4 hp = hit_pos[i];
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hp.w != 0.0f) {
      const float4 nn = hit_nrm[i];
      float cr, cg, cb;
      cv_light(C, C.dif, hp.x, hp.y, hp.z, nn.x, nn.y, nn.z, cr, cg, cb);
      float sx = 0.0f, sy = 0.0f, sz = 0.0f;
      const int nbx[4] = {x + 1, x - 1, x, x}, nby[4] = {y, y, y + 1, y - 1};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int qx = min(max(nbx[k], 0), w - 1), qy = min(max(nby[k], 0), h - 1);
        const float4 q = hit_nrm[(size_t)qy * w + qx];
        const float ax = fabsf(q.x - nn.x), ay = fabsf(q.y - nn.y), az = fabsf(q.z - nn.z);
        if (k == 0) { sx = ax; sy = ay; sz = az; } else { sx = sx + ax; sy = sy + ay; sz = sz + az; }
      }
      const float curv = sqrtf(dot(sx, sy, sz, sx, sy, sz));
      const float dx = hp.x - C.pick[0], dy = hp.y - C.pick[1], dz = hp.z - C.pick[2];
      const float dist_w = sqrtf(dot(dx, dy, dz, dx, dy, dz)) * C.param[0];
      const float blend = clamp01(fmaxf(curv * C.param[1], clamp01(dist_w)));
      float fr = 0.0f, fg = 0.0f, fb = 0.0f, fa = 0.0f;
      const float4 hp2 = cv_pos[i];
      if (hp2.w != 0.0f) {
        const float4 n2 = cv_nrm[i];
        cv_light(C, C.dif2, hp2.x, hp2.y, hp2.z, n2.x, n2.y, n2.z, fr, fg, fb);
        fa = 1.0f;
      }
      const float omb = 1.0f - blend;
      const float r = clamp01(cr * blend + fr * omb), g = clamp01(cg * blend + fg * omb), b = clamp01(cb * blend + fb * omb);
      const float a = clamp01(1.0f * blend + fa * omb);
      const float border = 0.5f * (1.0f - clamp01(fabsf(dist_w - 1.0f) * C.param[2]));
      o = make_float4(r - border, g - border, b - border, a);
    }
    rgba[i] = o;
  }
}

// Stereo eye composition over a full-screen quad: pixel (x, y) carries the texture coordinate ((x+.5)/w, (y+.5)/h) and
// the eye FBOs are GL_NEAREST / clamp (GLRenderer.cpp:688-713,1775-1785).  MODE = AbstrRenderer::EStereoMode.
__device__ __forceinline__ float4 eye_fetch(const float4* img, uint32_t w, uint32_t h, float s, float t) {
  int i = (int)floorf(s * (float)w), j = (int)floorf(t * (float)h);
  i = min(max(i, 0), (int)w - 1);
  j = min(max(j, 0), (int)h - 1);
  return img[(size_t)j * w + i];
}

template <int MODE>
__global__ void stereo_compose_kernel(const float4* __restrict__ left, const float4* __restrict__ right,
                                      float4* __restrict__ out, uint32_t w, uint32_t h, int alt_id, float split) {
  const uint64_t n = (uint64_t)w * h;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (uint64_t)y * w);
    const float s = ((float)x + 0.5f) / (float)w, t = ((float)y + 0.5f) / (float)h;
    float4 o;
    if (MODE == 0) {          // Compose-Anaglyphs-FS.glsl:41-47
      const float4 a = eye_fetch(left, w, h, s, t), b = eye_fetch(right, w, h, s, t);
      const float gl = dot(a.x, a.y, a.z, 0.3f, 0.59f, 0.11f), gr = dot(b.x, b.y, b.z, 0.3f, 0.59f, 0.11f);
      o = make_float4(gl, gr * 0.5f, gr, fmaxf(a.w, b.w));
    } else if (MODE == 1) {   // Compose-Scanline-FS.glsl:41-46
      const float line = floorf(t * (float)h);
      o = (line / 2.0f == floorf(line / 2.0f)) ? eye_fetch(left, w, h, s, t) : eye_fetch(right, w, h, s, t);
    } else if (MODE == 2) {   // Compose-SBS-FS.glsl:40-45
      o = (s < split) ? eye_fetch(left, w, h, s * 2.0f, t) : eye_fetch(right, w, h, (s - split) * 2.0f, t);
    } else {                  // Compose-AF-FS.glsl:40-43
      o = alt_id == 0 ? eye_fetch(left, w, h, s, t) : eye_fetch(right, w, h, s, t);
    }
    out[i] = o;
  }
}

inline int grid_for(uint64_t n, int block) {
  uint64_t g = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)kSMs * 8;   // 8 resident 256-thread CTAs per SM
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace

void launch_iso_compose(const float4* hit_pos, const float4* hit_nrm, float4* rgba, uint32_t w, uint32_t h,
                        const float amb[3], const float dif[3], const float spe[3], const float ldir[3],
                        cudaStream_t s, bool color) {
  ComposeConsts C;
  for (int i = 0; i < 3; i++) { C.amb[i] = amb[i]; C.dif[i] = dif[i]; C.spe[i] = spe[i]; C.ldir[i] = ldir[i]; }
  const uint64_t n = (uint64_t)w * h;
  if (color) iso_compose_kernel<true><<<grid_for(n, 256), 256, 0, s>>>(hit_pos, hit_nrm, rgba, n, C);
  else iso_compose_kernel<false><<<grid_for(n, 256), 256, 0, s>>>(hit_pos, hit_nrm, rgba, n, C);
}

void launch_quantize_rgba8(const float4* src, uchar4* dst, uint64_t n, cudaStream_t s) {
  quantize_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, dst, n);
}

void launch_cv_compose(const float4* hit_pos, const float4* hit_nrm, const float4* cv_pos, const float4* cv_nrm, float4* rgba,
                       uint32_t w, uint32_t h, const float amb[3], const float dif[3], const float dif2[3], const float spe[3],
                       const float ldir[3], const float cv_param[3], const float pick[3], cudaStream_t s) {
  CvConsts C;
  for (int i = 0; i < 3; i++) {
    C.amb[i] = amb[i]; C.dif[i] = dif[i]; C.dif2[i] = dif2[i]; C.spe[i] = spe[i]; C.ldir[i] = ldir[i];
    C.param[i] = cv_param[i]; C.pick[i] = pick[i];
  }
  cv_compose_kernel<<<grid_for((uint64_t)w * h, 256), 256, 0, s>>>(hit_pos, hit_nrm, cv_pos, cv_nrm, rgba, (int)w, (int)h, C);
}

void launch_stereo_compose(int mode, const float4* left, const float4* right, float4* out, uint32_t w, uint32_t h,
                           int alternating_frame_id, float split_coord, cudaStream_t s) {
  const int g = grid_for((uint64_t)w * h, 256);
  switch (mode) {
    case 0: stereo_compose_kernel<0><<<g, 256, 0, s>>>(left, right, out, w, h, alternating_frame_id, split_coord); break;
    case 1: stereo_compose_kernel<1><<<g, 256, 0, s>>>(left, right, out, w, h, alternating_frame_id, split_coord); break;
    case 2: stereo_compose_kernel<2><<<g, 256, 0, s>>>(left, right, out, w, h, alternating_frame_id, split_coord); break;
    default: stereo_compose_kernel<3><<<g, 256, 0, s>>>(left, right, out, w, h, alternating_frame_id, split_coord); break;
  }
}

void launch_composite_over(const float4* front, const float4* back, float4* out, uint64_t n, cudaStream_t s) {
  over_kernel<<<grid_for(n, 256), 256, 0, s>>>(front, back, out, n);
}

void launch_sl_wait(const SlPeer& P, int kind, uint32_t frame, uint32_t* local, cudaStream_t s) {
  sl_wait_kernel<<<1, 32, 0, s>>>(P, kind, frame, local);
}
void launch_sl_signal(const SlPeer& P, int kind, uint32_t frame, cudaStream_t s) { sl_signal_kernel<<<1, 32, 0, s>>>(P, kind, frame); }
void launch_nway_over_peer(const NWaySrc& a, float4* out_f, uchar4* out8, uint64_t n, const SlPeer& P, uint32_t frame,
                           uint32_t* local, cudaStream_t s) {
  // one CTA per SM x 4: every block must be resident while it spins on the flags (no block may wait for a slot)
  const int g = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)kSMs * 4));
  nway_over_peer_kernel<<<g, 256, 0, s>>>(a, out_f, out8, n, P, frame, local);
}

void launch_nway_over(const NWaySrc& a, float4* out_f, uchar4* out8, uint64_t n, cudaStream_t s) {
  if (n) nway_over_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, out_f, out8, n);
}

}  // namespace tvk
