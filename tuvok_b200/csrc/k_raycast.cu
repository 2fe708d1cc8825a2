// k_raycast.cu -- the GridLeaper traversal kernel for sm_100a.
//
// One thread = one ray (one fragment of GLGridLeaper-blend.glsl / -iso.glsl main()); a warp is
// an 8x4 pixel tile so neighbouring rays walk the same bricks and share cache lines.  Replaces
// (reference file:line):
//   ray entry/exit     GLGridLeaper.cpp:560-620, GLGridLeaper-NearPlane-VS.glsl:9-14,
//                      GLGridLeaper-entry-VS.glsl:10-14, GLGridLeaper-frontfaces-FS.glsl:6-8
//   main() DVR / ISO   Shaders/GLGridLeaper-blend.glsl:65-228, GLGridLeaper-iso.glsl:68-200
//   page-table walk    generated GLSL, Renderer/GL/GLVolumePool.cpp:484-656
//   miss reports       generated GLSL, Renderer/GL/GLHashTable.cpp:136-182
//   classification     GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}.glsl, GLGridLeaper-GradientTools.glsl:6-23,
//                      lighting.glsl:33-43, Compositing.glsl:33-38
//
// HBM layout: the pool is SLOT-LINEAR -- slot s (= the reference's linear pool coordinate
// x + y*capX + z*capX*capY) is one contiguous maxTotalBrickSize^3 block, x fastest -- instead of
// the reference's 3D-texture atlas.  The shader's pool texture coordinates are kept (virtual
// atlas of capacity*brick voxels) so the arithmetic stays the reference's; only the final
// texel address is slot-local.
//
// Control flow: the shader's nested loops (bricks along the ray / samples inside a brick) are run
// as one flat per-warp loop with a "fetch next brick" phase and a "take one sample" phase, so
// lanes whose brick ends early do not idle until the slowest lane of the warp finishes its brick;
// per ray the sequence of operations is exactly the shader's.
//
// Arithmetic contract: IEEE fp32, no implicit FMA contraction (this file is compiled
// with -fmad=false); fmaf() exactly where the contract names it (texel-coordinate map, trilinear
// lerps, dot products, under-compositing).  Gradient taps sit exactly +-1 texel from the centre
// sample and share its filter fractions (the GLSL adds sampleDelta = 1/poolSize in texture
// coordinates, i.e. one texel; the precision of that is implementation-defined in GL).
#include "tvk_dev.h"

namespace tvk {
namespace {

#include "tvk_math.cuh"

#include "tvk_traverse.cuh"

#ifndef TVK_FETCH_LANES
#define TVK_FETCH_LANES 32
#endif
constexpr int kFetchLanes = TVK_FETCH_LANES;
#ifndef TVK_MIN_BLOCKS
#define TVK_MIN_BLOCKS 8
#endif
#ifndef TVK_PREFETCH
#define TVK_PREFETCH 0
#endif
constexpr bool kPrefetch = TVK_PREFETCH != 0;   // fetch the next sample's footprint before shading the current one
// CTA = TVK_WX x TVK_WY warps, each an 8x4 pixel tile (CTA covers 8*WX x 4*WY pixels)
#ifndef TVK_WX
#define TVK_WX 2
#endif
#ifndef TVK_WY
#define TVK_WY 2
#endif
constexpr int kWX = TVK_WX, kWY = TVK_WY, kThreads = 32 * TVK_WX * TVK_WY;
#ifndef TVK_SKIP_CLEAR
#define TVK_SKIP_CLEAR 1
#endif
constexpr bool kSkipClear = TVK_SKIP_CLEAR != 0;   // A/B switch for the zero-alpha shortcut in shade()
#ifndef TVK_CENTRE_OUT
#define TVK_CENTRE_OUT 1
#endif
constexpr bool kCentreOut = TVK_CENTRE_OUT != 0;
// k-th element of the sequence c, c-1, c+1, c-2, c+2, ... (c = n/2): a bijection of [0, n)
__device__ __forceinline__ uint32_t centre_out(uint32_t k, uint32_t n) {
  const uint32_t c = n >> 1, d = (k + 1) >> 1;
  return (k & 1u) ? c - d : c + d;
}

// State of the brick chain (page-table walk).  None of it is needed while a lane samples, so between two chain
// phases it is PARKED in shared memory (TVK_PARK) instead of occupying ~37 registers of every thread for the
// whole kernel; the sample phase then fits a smaller register budget and more warps are resident per SM.
struct ChainSt {
  f3 entry, nexit, dir, dv, nudge, cur;
  float ray_len, step, entry_depth, exit_depth, t;
  f4 resume_pos, resume_col;
  uint32_t j, lbx, lby, lbz, lbl;
  bool optimal;
};
#ifndef TVK_PARK
#define TVK_PARK 0   // measured on B200 (gpurun_out/r1_ab3.log): parking + more resident warps is SLOWER (the extra
#endif               // shared memory shrinks L1 and more warps thrash it): 229 fps vs 238; 10/12 CTAs: 179/167 fps
constexpr bool kPark = TVK_PARK != 0;
constexpr int kChainWords = 37;
__device__ __forceinline__ void park_const(const ChainSt& c, float (*m)[kThreads], int tid) {   // written once per ray
  m[0][tid] = c.entry.x; m[1][tid] = c.entry.y; m[2][tid] = c.entry.z;
  m[3][tid] = c.nexit.x; m[4][tid] = c.nexit.y; m[5][tid] = c.nexit.z;
  m[6][tid] = c.dir.x; m[7][tid] = c.dir.y; m[8][tid] = c.dir.z;
  m[9][tid] = c.dv.x; m[10][tid] = c.dv.y; m[11][tid] = c.dv.z;
  m[12][tid] = c.nudge.x; m[13][tid] = c.nudge.y; m[14][tid] = c.nudge.z;
  m[15][tid] = c.ray_len; m[16][tid] = c.step; m[17][tid] = c.entry_depth; m[18][tid] = c.exit_depth;
}
__device__ __forceinline__ void park_var(const ChainSt& c, float (*m)[kThreads], int tid) {     // after every chain phase
  m[19][tid] = c.cur.x; m[20][tid] = c.cur.y; m[21][tid] = c.cur.z; m[22][tid] = c.t;
  m[23][tid] = c.resume_pos.x; m[24][tid] = c.resume_pos.y; m[25][tid] = c.resume_pos.z; m[26][tid] = c.resume_pos.w;
  m[27][tid] = c.resume_col.x; m[28][tid] = c.resume_col.y; m[29][tid] = c.resume_col.z; m[30][tid] = c.resume_col.w;
  m[31][tid] = __uint_as_float(c.j); m[32][tid] = __uint_as_float(c.lbx); m[33][tid] = __uint_as_float(c.lby);
  m[34][tid] = __uint_as_float(c.lbz); m[35][tid] = __uint_as_float(c.lbl);
  m[36][tid] = __uint_as_float(c.optimal ? 1u : 0u);
}
__device__ __forceinline__ void unpark_result(ChainSt& c, float (*m)[kThreads], int tid) {      // what TerminateRay needs
  c.resume_pos.x = m[23][tid]; c.resume_pos.y = m[24][tid]; c.resume_pos.z = m[25][tid]; c.resume_pos.w = m[26][tid];
  c.resume_col.x = m[27][tid]; c.resume_col.y = m[28][tid]; c.resume_col.z = m[29][tid]; c.resume_col.w = m[30][tid];
  c.optimal = __float_as_uint(m[36][tid]) != 0u;
}
__device__ __forceinline__ void unpark(ChainSt& c, float (*m)[kThreads], int tid) {
  c.entry = F3(m[0][tid], m[1][tid], m[2][tid]);
  c.nexit = F3(m[3][tid], m[4][tid], m[5][tid]);
  c.dir = F3(m[6][tid], m[7][tid], m[8][tid]);
  c.dv = F3(m[9][tid], m[10][tid], m[11][tid]);
  c.nudge = F3(m[12][tid], m[13][tid], m[14][tid]);
  c.ray_len = m[15][tid]; c.step = m[16][tid]; c.entry_depth = m[17][tid]; c.exit_depth = m[18][tid];
  c.cur = F3(m[19][tid], m[20][tid], m[21][tid]); c.t = m[22][tid];
  c.j = __float_as_uint(m[31][tid]); c.lbx = __float_as_uint(m[32][tid]); c.lby = __float_as_uint(m[33][tid]);
  c.lbz = __float_as_uint(m[34][tid]); c.lbl = __float_as_uint(m[35][tid]);
  unpark_result(c, m, tid);
}

// MODE: 0 = 1D TF, 1 = 2D TF, 2 = isosurface
// PIPE (DVR modes): this launch is one STAGE of the depth pipeline: the rank's block is a slab
// of the volume; a ray comes in with the colour the stages in front of it accumulated (start colour / resume position,
// exactly the inputs of a resumed GridLeaper subframe), is marched through the slab, and leaves with its resume
// position at the slab's far side -- or finished (w = 1000) if it terminated early or left the volume.  Early ray
// termination therefore works across ranks as on one GPU.
// TVK_PERSIST (persistent warps that hand the pixels of a tile queue to their idle lanes) was built and MEASURED SLOWER on
// B200 / C3 (profiles/r2x_persist_ab.txt): non-persistent 304 fps; persistent, a tile only when the whole warp is idle
// (TVK_REFILL=32) 293 fps; refill as soon as 8 / 16 / 24 lanes are idle 178 / 185 / 180 fps -- images bit-identical in every
// variant (208 GPU tests).  Lanes that start rays at different times need their brick-chain phases at different turns, and
// the chain phase (page-table walk) runs for the whole warp whenever ONE lane needs it: refilling trades idle lanes for
// many more, thinner chain phases.  Kept as a build switch, off by default.
#ifndef TVK_PERSIST
#define TVK_PERSIST 0
#endif
#ifndef TVK_REFILL
#define TVK_REFILL 32   // 32: a warp takes its next tile only when all its lanes are idle (lanes without a ray help, TVK_HELP)
#endif
// TVK_HELP (persistent variant only): lanes without a ray shade later samples of a neighbour's brick segment.  Built,
// bit-identical, MEASURED SLOWER (profiles/r3e_help_ab.txt: 313.7 -> 241.9 fps): the long rays of a frame are neighbours, so
// the warps on the critical path have no idle lanes to help them.
#ifndef TVK_HELP
#define TVK_HELP 1
#endif
#ifndef TVK_HELP_MAX
#define TVK_HELP_MAX 3
#endif
constexpr bool kHelp = TVK_HELP != 0;
constexpr int kHelpMax = TVK_HELP_MAX;
constexpr int kRefill = TVK_REFILL;
constexpr int kSMs = 148;
#if TVK_PERSIST
// Persistent warps with lane refill (TVK_PERSIST=1, see the measurement above): the launch is one CTA slot per SM x resident CTAs, and a warp
// does not own one 8x4 tile: it draws tiles from a global counter (the old dispatch order: 2x2 tile groups, centre-out)
// and gives the pixels of the open tile to its IDLE lanes -- a lane whose ray has terminated (early ray termination,
// left the volume) starts the next pixel while its neighbours are still marching, instead of idling until the slowest ray
// of the tile is done (ncu before: 24.4 of 32 lanes active per issued instruction).  Rays, their arithmetic and their
// outputs are unchanged: only WHICH lane traces WHICH pixel WHEN differs.  A warp interrupts its live rays for a refill
// only when kRefill lanes are idle (the set-up runs with the idle lanes only).
template <typename T, int MODE, bool LIT, bool FAST, int BS, bool COUNT, bool PIPE = false>
__global__ void __launch_bounds__(kThreads, TVK_MIN_BLOCKS * 64 / kThreads) raycast_kernel(const __grid_constant__ RayConsts P) {
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr bool ISO = MODE == 2;
  typedef typename PairOf<T>::W W;           // pool element: the x-pair (voxel x, voxel x+1)
  const W* pool = (const W*)P.pool;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  __shared__ float park[kPark ? kChainWords : 1][kThreads];
  __shared__ float seg_f[9][kThreads];       // next segment: pool entry, trans, 1/scale
  __shared__ uint32_t seg_u[COUNT ? 7 : 6][kThreads];    // slot origin (3), slot index, steps, flags (, page-table index)
  // tile queue: 2x2 groups of 8x4 tiles, groups in the centre-out order of the non-persistent launch
  const uint32_t gx = (P.width + 15u) / 16u, gy = (P.height + 7u) / 8u;
  uint32_t open_tile = 0u, open_mask = 0u;   // warp-uniform: the tile being handed out, its pixels not yet taken
  bool more = true;                          // warp-uniform: the global queue has not run dry
  // SPLIT work units (warp-uniform): the 2x2 tile groups whose rays were LONG in the previous frame (the first n_split entries
  // of the cost-sorted order, tile_order_kernel) are handed out in quarter tiles -- 8 rays per warp -- so that the warp's
  // other 24 lanes help (3 helpers per ray, 4 samples of a ray per turn) and four warps share what one warp would have
  // walked alone: the critical path of the frame's longest rays is cut fourfold, every other tile runs the plain loop.
  bool split = false;
  const uint32_t n_groups = gx * gy;
  const uint32_t n_split = (kHelp && !ISO && !PIPE && P.tile_order) ? min(__ldg(P.tile_order + n_groups), n_groups) : 0u;
  const uint32_t n_units = 16u * n_split + 4u * (n_groups - n_split);
  uint32_t ray_group = 0u, n_work = 0u;      // the ray's tile group and the work it took (samples + sample-less turns): next frame's cost
  // ---- state of the lane's current ray
  size_t pix = 0;
  bool have_ray = false;           // a ray has been started and its outputs are not written yet
  bool done = false;
  unsigned long long n_samples = 0, n_bricks = 0;
  unsigned long long n_alive_iters = 0, n_warp_iters = 0;   // lane-utilisation diagnostics (count mode)
  ChainSt c;
  f4 acc = from4(zero4);
  f4& resume_col = c.resume_col;
  f4& resume_pos = c.resume_pos;
  f4 hit_pos = from4(zero4), hit_nrm = from4(zero4), resume_nrm = from4(zero4);
  bool handoff = false;            // PIPE: the ray left this stage's slab alive
  f4 hand_pos = from4(zero4);
  f3 vdir = F3(0.f, 0.f, 0.f);
  const float voxel_size = 0.125f / 2000.0f;
  const f3 dscale = F3(P.domain_scale), eye_m = F3(P.eye_m), la = F3(P.light_a), ld = F3(P.light_d),
           ls = F3(P.light_s), ldir = F3(P.light_dir_m);
  bool ray_live = false;                  // the ray has not terminated (ERT / iso hit)
  bool chain = false;                     // the brick chain has not reached the end of the ray
  bool have_next = false;                 // a prefetched segment waits in shared memory
  int steps_left = 0;      // samples left in the current brick
  bool b_partial = false;  // sort-last: the current brick straddles the shard box (ownership per sample)
  f3 pc = F3(0.f, 0.f, 0.f), b_trans = pc, b_inv = pc;
  uint32_t b_ox = 0, b_oy = 0, b_oz = 0;
  const W* vox = pool;
  constexpr bool GRAD = !ISO && (MODE == 1 || LIT);   // what a sample needs: the 7-tap footprint or the centre tap
  Foot<T, FAST, BS, GRAD> cur;   // footprint of the sample at pc, fetched one turn ahead
  bool cur_ok = false;
  unsigned long long pend = 0;   // COUNT: brick visits of the chain that the sampling has not reached yet

  for (;;) {
    // ---- refill: finished rays are written out, idle lanes take the next pixels
    const unsigned full = 0xffffffffu;
    const unsigned idle_m = __ballot_sync(full, !ray_live);
    if (idle_m == full || ((more || open_mask != 0u) && __popc(idle_m) >= kRefill)) {
      if (!ray_live && have_ray) {
        have_ray = false;
        if (!done) {
          if (kPark) unpark_result(c, park, tid);
          // TerminateRay
          if (!ISO) {
            if (c.optimal) {
              // ray_live is false only after early termination; a ray that ran out of bricks in this slab is handed on
              if (PIPE && handoff && !(acc.w > 0.99f)) { resume_pos = hand_pos; resume_col = acc; }
              else { resume_pos.w = 1000.0f; resume_col = acc; }
            }
          } else {
            if (c.optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
            resume_nrm = hit_nrm;
          }
        }
        if (!ISO) {
          P.out0[pix] = to4(acc); P.out1[pix] = to4(resume_col); P.out2[pix] = to4(resume_pos);
        } else {
          P.out0[pix] = to4(hit_pos); P.out1[pix] = to4(hit_nrm); P.out2[pix] = to4(resume_pos);
          P.out3[pix] = to4(resume_nrm);
        }
        if (COUNT) {
          atomicAdd(P.counters + 0, n_samples); atomicAdd(P.counters + 1, 1ull); atomicAdd(P.counters + 2, n_bricks);
          atomicAdd(P.counters + 3, n_alive_iters); atomicAdd(P.counters + 4, n_warp_iters);
          atomicMax(P.counters + 5, n_alive_iters);
        }
        if (P.tile_cost) atomicMax(P.tile_cost + ray_group, n_work);
      }
      // the r-th waiting lane takes the r-th open pixel of the tile
      unsigned want = idle_m;
      bool got = false;
      uint32_t px = 0, py = 0;
      while (want != 0u) {
        if (open_mask == 0u) {
          if (!more) break;
          uint32_t t = 0;
          if (lane == 0) t = atomicAdd(P.tile_counter, 1u);
          t = __shfl_sync(full, t, 0);
          if (t >= n_units) { more = false; break; }
          uint32_t j, sub;
          if (t < 16u * n_split) {           // a quarter of a tile of a costly group
            j = t >> 4; sub = (t >> 2) & 3u; split = true;
            open_mask = 0xffu << (8u * (t & 3u));
          } else {
            const uint32_t u = t - 16u * n_split;
            j = n_split + (u >> 2); sub = u & 3u; split = false;
            open_mask = full;
          }
          const uint32_t g = P.tile_order ? __ldg(P.tile_order + j) : j;
          open_tile = (g << 2) | sub;
        }
        const int n_take = min(__popc(open_mask), __popc(want));
        const int r = __popc(want & ((1u << lane) - 1u));
        const bool take = ((want >> lane) & 1u) != 0u && r < n_take;
        uint32_t bit = 0;
        if (take) {
          bit = __fns(open_mask, 0, r + 1);
          const uint32_t g = open_tile >> 2, sub = open_tile & 3u;
          const uint32_t sx = g % gx, sy = g / gx;
          const uint32_t tx = (kCentreOut ? centre_out(sx, gx) : sx) * 2u + (sub & 1u);
          const uint32_t ty = (kCentreOut ? centre_out(sy, gy) : sy) * 2u + (sub >> 1);
          px = tx * 8u + (bit & 7u); py = ty * 4u + (bit >> 3);
          got = true;
        }
        open_mask &= ~__reduce_or_sync(full, take ? (1u << bit) : 0u);
        want &= ~__ballot_sync(full, take);
        if (split) break;   // a split unit is all this warp takes: its other lanes are the helpers
      }
      if (got && px < P.width && py < P.height) {
        pix = (size_t)py * P.width + px;
        f4 entry4, exit4;
        const bool covered = ray_setup(P, px, py, entry4, exit4, !PIPE);   // a stage takes every ray that meets the VOLUME
        if (!covered) {   // render targets are cleared where no back face is rasterised (GLGridLeaper.cpp:837)
          P.out0[pix] = zero4; P.out1[pix] = zero4; P.out2[pix] = PIPE ? make_float4(0.f, 0.f, 0.f, 1000.0f) : zero4;
          if (ISO) P.out3[pix] = zero4;
        } else {
          have_ray = true; done = false; handoff = false;
          ray_group = open_tile >> 2; n_work = 0u;
          hand_pos = from4(zero4); hit_pos = from4(zero4); hit_nrm = from4(zero4); resume_nrm = from4(zero4);
          n_samples = 0; n_bricks = 0; n_alive_iters = 0; n_warp_iters = 0; pend = 0;
          if (P.first_pass) {
            resume_pos = entry4;
            acc = from4(zero4);
          } else {
            resume_pos = from4(P.ray_start[pix]);
            acc = from4(P.start_color[pix]);
          }
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
            c.entry = F3(resume_pos.x, resume_pos.y, resume_pos.z);
            c.entry_depth = resume_pos.w;
            c.nexit = F3(exit4.x, exit4.y, exit4.z);
            c.exit_depth = exit4.w;
            c.dir = sub3(c.nexit, c.entry);
            c.ray_len = len3(c.dir);
            // TransformToPoolSpace
            const f3 ps = F3(P.pool_size_f);
            vdir = norm3(mul3(c.dir, F3(P.vol_f)));
            vdir = div3(vdir, ps);
            const float den = 2.0f * P.sample_rate;
            vdir = F3(vdir.x / den, vdir.y / den, vdir.z / den);
            c.step = len3(vdir);
            c.t = 0.0f;
            c.optimal = true;
            c.cur = c.entry;
            c.j = 0;   // bricks visited by the chain (the shader's j < 100 bound)
            c.lbx = 0; c.lby = 0; c.lbz = 0; c.lbl = 9999;
        
            c.dv = F3(1.0f / c.dir.x, 1.0f / c.dir.y, 1.0f / c.dir.z);   // BrickExit's 1.0/dir
            // the empty-brick advance voxelSize*direction/rayLength
            c.nudge = F3(voxel_size * c.dir.x / c.ray_len, voxel_size * c.dir.y / c.ray_len, voxel_size * c.dir.z / c.ray_len);
            ray_live = c.ray_len > voxel_size;
            chain = ray_live; have_next = false; steps_left = 0; b_partial = false; cur_ok = false;
            if (kPark) { park_const(c, park, tid); park_var(c, park, tid); }
          }
        }
      }
      if (__ballot_sync(full, ray_live || have_ray) == 0u && !more && open_mask == 0u) break;
    }
    if (ray_live) {
      // ---- chain phase: runs for the whole warp when some lane can neither sample nor pick up a segment
      const unsigned act = __activemask();
      const bool need = steps_left == 0 && !have_next && chain;
      if (__ballot_sync(act, need) != 0u && chain && !have_next) {
        if (kPark) unpark(c, park, tid);
#pragma unroll 1
        for (int f = 0; f < 4 && chain && !have_next; f++) {
          if (c.j >= 100) { chain = false; break; }
          if (P.shard) {   // the block is convex: once the ray has left it there is nothing more to do on this rank
            const bool gone = (c.dir.x > 0.0f && c.cur.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && c.cur.x <= P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && c.cur.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && c.cur.y <= P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && c.cur.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && c.cur.z <= P.sh_lo[2]);
            if (gone) {
              if (PIPE) {   // where the next stage picks the ray up
                handoff = true;
                hand_pos.x = c.cur.x; hand_pos.y = c.cur.y; hand_pos.z = c.cur.z;
                hand_pos.w = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
              }
              chain = false; break;
            }
          }
          const float cur_depth = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
          uint32_t lod = compute_lod(P, cur_depth);
          BrickRef b;
          int ok;
          if (steps_left > 0) {   // look-ahead: the ray is still sampling the previous segment
            ok = get_brick<true>(P, c.cur, lod, c.dir, c.dv, b);
            if (ok < 0) break;    // missing brick: handled when the ray stands here
          } else {
            ok = get_brick<false>(P, c.cur, lod, c.dir, c.dv, b);
            if (!ok && c.optimal) {
              c.optimal = false;
              resume_pos.x = c.cur.x; resume_pos.y = c.cur.y; resume_pos.z = c.cur.z; resume_pos.w = cur_depth;
              if (!ISO) resume_col = acc;
            }
          }
          if (COUNT) pend++;
          if (!b.empty && !(c.lbx == b.bx && c.lby == b.by && c.lbz == b.bz && c.lbl == b.bl)) {
            int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / c.step);
            const int s2 = (int)ceilf(len3(mul3(sub3(c.nexit, c.cur), b.scale)) / c.step);
            steps = min(steps, s2);
            const f3 inv = F3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
            f3 pe = b.pool_entry;
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
            if (b.where == OUTSIDE_SHARD) {   // another rank's brick: advance by its steps, take no sample
              const float n = (float)max(steps, 0);
              pe = F3(fmaf(n, vdir.x, pe.x), fmaf(n, vdir.y, pe.y), fmaf(n, vdir.z, pe.z));
              steps = 0;
            }
            if (steps > 0) {
              seg_f[0][tid] = pe.x; seg_f[1][tid] = pe.y; seg_f[2][tid] = pe.z;
              seg_f[3][tid] = b.trans.x; seg_f[4][tid] = b.trans.y; seg_f[5][tid] = b.trans.z;
              seg_f[6][tid] = inv.x; seg_f[7][tid] = inv.y; seg_f[8][tid] = inv.z;
              seg_u[0][tid] = b.ox; seg_u[1][tid] = b.oy; seg_u[2][tid] = b.oz;
              seg_u[3][tid] = b.slot; seg_u[4][tid] = (uint32_t)steps;
              seg_u[5][tid] = b.where == PARTLY_IN_SHARD ? 1u : 0u;
              if (COUNT) seg_u[6][tid] = b.id;
              have_next = true;
              // where the sample loop will leave pc: `steps` SEQUENTIAL adds (fp32 addition is not associative, so
              // there is no closed form); 4x unrolled -- this loop was 6.5 % of the launch's warp instructions
#pragma unroll 4
              for (int i = 0; i < steps; i++) pe = add3(pe, vdir);
            }
            c.cur = mul3(sub3(pe, b.trans), inv);
          } else {
            c.cur = add3(b.norm_exit, c.nudge);
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
          }
          c.t = len3(sub3(c.entry, b.norm_exit)) / c.ray_len;
          c.j++;
          if (c.t > 0.9999f) chain = false;
        }
        if (kPark) park_var(c, park, tid);
      }
      // ---- pick up the waiting segment
      if (steps_left == 0 && have_next) {
        pc = F3(seg_f[0][tid], seg_f[1][tid], seg_f[2][tid]);
        b_trans = F3(seg_f[3][tid], seg_f[4][tid], seg_f[5][tid]);
        b_inv = F3(seg_f[6][tid], seg_f[7][tid], seg_f[8][tid]);
        b_ox = seg_u[0][tid]; b_oy = seg_u[1][tid]; b_oz = seg_u[2][tid];
        vox = pool + (uint64_t)seg_u[3][tid] * P.slot_voxels;
        steps_left = (int)seg_u[4][tid];
        b_partial = seg_u[5][tid] != 0u;
        have_next = false;
        cur_ok = false;
        if (COUNT) {
          n_bricks += pend; pend = 0;
          if (P.visited) { const uint32_t id = seg_u[6][tid]; atomicOr(P.visited + (id >> 5), 1u << (id & 31)); }
        }
      }
      if (steps_left == 0 && !chain) {   // the ray left the volume (or its 100-brick budget) unterminated
        if (COUNT) n_bricks += pend;
        ray_live = false;
      }
      if (COUNT) {
        n_alive_iters += ray_live ? 1 : 0;
        if (__ffs(__activemask()) - 1 == (tid & 31)) n_warp_iters++;
      }
    }   // if (ray_live): chain phase, segment pick-up, end-of-ray test
    if constexpr (ISO || PIPE || !kHelp) {
      // ---- sample phase: one sample for every lane that is inside a brick ----
      if (ray_live && steps_left > 0) {
        bool terminated = false;
        // software pipeline: the footprint of THIS sample was fetched during the previous turn (cur_ok), the one of the
        // next sample of the brick is fetched now, before this sample's arithmetic, so its load latency is covered
        if (!kPrefetch || !FAST || !cur_ok) cur.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
        const f3 pc_next = add3(pc, vdir);
        const bool pf = kPrefetch && FAST && steps_left > 1;
        Foot<T, FAST, BS, GRAD> nxt;
        if (pf) nxt.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc_next);
        bool mine = true;
        if (b_partial) {
          const f3 mq = mul3(sub3(pc, b_trans), b_inv);
          mine = mq.x >= P.sh_lo[0] && mq.x < P.sh_hi[0] && mq.y >= P.sh_lo[1] && mq.y < P.sh_hi[1] &&
                 mq.z >= P.sh_lo[2] && mq.z < P.sh_hi[2];
          if (PIPE && !ISO && !mine) {
            // a brick of a coarser LoD straddles the slab's far side: the ray is handed on AT the side, not
            // behind the brick, so the next stage takes the brick's remaining samples
            const bool gone = (c.dir.x > 0.0f && mq.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && mq.x < P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && mq.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && mq.y < P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && mq.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && mq.z < P.sh_lo[2]);
            if (gone) {
              const float tq = len3(sub3(mq, c.entry)) / c.ray_len;
              handoff = true;
              hand_pos.x = mq.x; hand_pos.y = mq.y; hand_pos.z = mq.z;
              hand_pos.w = c.entry_depth * (1.0f - tq) + c.exit_depth * tq;
              terminated = true;   // leaves the loop; TerminateRay sees alpha <= 0.99 and hands the ray on
            }
          }
        }
        if constexpr (!ISO) {
          if (mine) {
            if (COUNT) n_samples++;
            // ComputeColorFromVolume + OpacityCorrectColor at pool position pc
            f4 col;
            bool clear = false;
            if constexpr (MODE == 0 && !LIT) {
              col = tf_lookup(P, cur.centre(P) * P.trans_scale, 0.0f);
            } else if constexpr (MODE == 0) {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              col = tf_lookup(P, data * P.trans_scale, 0.0f);
              // A sample whose transfer-function alpha is exactly 0 leaves the ray unchanged bit for bit
              // (UnderCompositing adds colour * (1-a) * 0 = 0 to every channel; table colours and lit colours are
              // finite, and opacity correction maps 0 to 0), so its normal and lighting are not computed when the
              // whole warp agrees.  The shader cannot branch this cheaply; the result is identical.
              clear = kSkipClear && col.w == 0.0f;
              if (!clear) {
                f3 n = mul3(g, dscale);   // ComputeNormal
                const float l = len3(n);
                if (l > 0.0f) n = scl3(n, 1.0f / l);
                const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                col.x = lit.x; col.y = lit.y; col.z = lit.z;
              }
            } else {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              const float gm = len3(g);
              col = tf_lookup(P, data * P.trans_scale, 1.0f - gm * P.gradient_scale);
              if (LIT) {
                clear = kSkipClear && col.w == 0.0f;
                if (!clear) {
                  const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                  const f3 n = mul3(dscale, gn);
                  const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                  float dl, sp;
                  light_terms(eye_m, mp, n, ldir, dl, sp);
                  const f3 lit = light_apply(la, mul3(F3(col.x, col.y, col.z), ld), ls, dl, sp);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              }
            }
            if (!clear) {
              col.w = opacity_correct(P, col.w);
              // UnderCompositing
              const float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) terminated = true;
            }
          }
        } else if (mine) {   // isosurface march
          if (COUNT) n_samples++;
          if (cur.centre(P) >= P.isoval) {
            // RefineIsosurface
            f3 rd = F3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
            pc = sub3(pc, rd);
            Foot<T, FAST, BS, false> rf;
  #pragma unroll 1
            for (int k = 0; k < 5; k++) {
              rd = F3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
              rf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
              if (rf.centre(P) >= P.isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
            }
            const f3 hp = mul3(sub3(pc, b_trans), b_inv);
            hit_pos = xform4(P.m2e, hp.x, hp.y, hp.z, 1.0f);
            hit_pos.w = 1.0f + 1.0f;   // color.r + 1
            Foot<T, FAST, BS, true> gf;
            gf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
            float dummy; f3 g;
            gf.sample_with_gradient(P, dummy, g);
            f3 n = mul3(g, dscale);
            const float l = len3(n);
            if (l > 0.0f) n = scl3(n, 1.0f / l);
            const float* m = P.mv_inv;   // mModelViewIT * vec4(n, 0)
            hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
            hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
            hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
            hit_nrm.w = floorf(1.0f * 512.0f) + 1.0f;   // floor(color.g*512)+color.b
            terminated = true;
          } else {
            hit_pos = from4(zero4);
          }
        }
        steps_left -= 1;
        if (terminated) ray_live = false;
        else {
          pc = pc_next;
          if (pf) cur = nxt;
          cur_ok = pf;
        }
      }
    } else {
      // ---- sample phase with HELPER LANES.  The samples of one brick segment are independent until they are blended, and
      // a launch cannot finish before its longest ray has walked the serial chain of one sample (load -> filter -> gradient ->
      // table fetch -> Phong -> blend) hundreds of times.  So a lane WITHOUT a ray (finished, uncovered,
      // waiting for the next tile) adopts a neighbour's segment for one turn and shades that ray's sample k + off at
      // pc + off * vdir -- the very position the owner would reach by its own sequential adds -- with the owner's brick state;
      // the owner then blends its own sample and its helpers' samples IN RAY ORDER, stopping at early termination exactly
      // where it would have stopped alone.  Same samples, same arithmetic, same order: the ray's result is bit-identical,
      // its critical path is up to kHelpMax + 1 times shorter.
      const bool own = ray_live && steps_left > 0;
      bool helper = false;
      int hc = 0, first_rank = 0;   // owner: number of my helpers, rank of my first helper among the idle lanes
      int h = 0;                    // warp-uniform: helpers per owner this turn
      const unsigned idle_h = split ? __ballot_sync(full, !ray_live) : 0u;
      const unsigned owner_m = split ? __ballot_sync(full, own && steps_left >= 2 && !b_partial) : 0u;
      if (idle_h != 0u && owner_m != 0u) {
        const int n_idle = __popc(idle_h), n_own = __popc(owner_m);
        h = max(1, min(kHelpMax, n_idle / n_own));
        const unsigned lt = (1u << lane) - 1u;
        int src = lane, off = 0;
        if (!ray_live) {
          const int r = __popc(idle_h & lt);
          if (r / h < n_own) { src = (int)__fns(owner_m, 0, r / h + 1); off = 1 + r % h; }
        } else if (own && steps_left >= 2 && !b_partial) {
          first_rank = __popc(owner_m & lt) * h;
          hc = max(0, min(h, n_idle - first_rank));
        }
        // the owner's segment state travels to its helpers (every lane executes the shuffles; a lane that helps nobody reads itself)
        const f3 o_pc = F3(__shfl_sync(full, pc.x, src), __shfl_sync(full, pc.y, src), __shfl_sync(full, pc.z, src));
        const f3 o_vd = F3(__shfl_sync(full, vdir.x, src), __shfl_sync(full, vdir.y, src), __shfl_sync(full, vdir.z, src));
        const f3 o_tr = F3(__shfl_sync(full, b_trans.x, src), __shfl_sync(full, b_trans.y, src), __shfl_sync(full, b_trans.z, src));
        const f3 o_in = F3(__shfl_sync(full, b_inv.x, src), __shfl_sync(full, b_inv.y, src), __shfl_sync(full, b_inv.z, src));
        const uint32_t o_ox = __shfl_sync(full, b_ox, src), o_oy = __shfl_sync(full, b_oy, src), o_oz = __shfl_sync(full, b_oz, src);
        const int o_steps = __shfl_sync(full, steps_left, src);
        const unsigned long long o_vox = __shfl_sync(full, (unsigned long long)vox, src);
        if (off > 0 && off < o_steps) {
          helper = true;
          pc = o_pc;
          for (int i = 0; i < off; i++) pc = add3(pc, o_vd);   // the owner's own sequence of adds
          b_trans = o_tr; b_inv = o_in; b_ox = o_ox; b_oy = o_oy; b_oz = o_oz;
          vox = (const W*)o_vox;
        }
      }
      // ---- shade: a lane's own sample and an adopted sample run the same code
      f4 col = from4(zero4);
      bool clear = false, mine = true;
      if (own || helper) {
        cur.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
        if (own && b_partial) {
          const f3 mq = mul3(sub3(pc, b_trans), b_inv);
          mine = mq.x >= P.sh_lo[0] && mq.x < P.sh_hi[0] && mq.y >= P.sh_lo[1] && mq.y < P.sh_hi[1] &&
                 mq.z >= P.sh_lo[2] && mq.z < P.sh_hi[2];
        }
        if (mine) {
          // ComputeColorFromVolume + OpacityCorrectColor at pool position pc
          if constexpr (MODE == 0 && !LIT) {
            col = tf_lookup(P, cur.centre(P) * P.trans_scale, 0.0f);
          } else if constexpr (MODE == 0) {
            float data; f3 g;
            cur.sample_with_gradient(P, data, g);
            col = tf_lookup(P, data * P.trans_scale, 0.0f);
            clear = kSkipClear && col.w == 0.0f;   // see the note in the single-lane path: the result is identical
            if (!clear) {
              f3 n = mul3(g, dscale);   // ComputeNormal
              const float l = len3(n);
              if (l > 0.0f) n = scl3(n, 1.0f / l);
              const f3 mp = mul3(sub3(pc, b_trans), b_inv);
              const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
              col.x = lit.x; col.y = lit.y; col.z = lit.z;
            }
          } else {
            float data; f3 g;
            cur.sample_with_gradient(P, data, g);
            const float gm = len3(g);
            col = tf_lookup(P, data * P.trans_scale, 1.0f - gm * P.gradient_scale);
            if (LIT) {
              clear = kSkipClear && col.w == 0.0f;
              if (!clear) {
                const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                const f3 n = mul3(dscale, gn);
                const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                float dl, sp;
                light_terms(eye_m, mp, n, ldir, dl, sp);
                const f3 lit = light_apply(la, mul3(F3(col.x, col.y, col.z), ld), ls, dl, sp);
                col.x = lit.x; col.y = lit.y; col.z = lit.z;
              }
            }
          }
          if (!clear) col.w = opacity_correct(P, col.w);
        }
      }
      // ---- blend in ray order: the lane's own sample, then its helpers' samples (UnderCompositing)
      if (ray_live) n_work++;
      if (own) {
        if (mine) {
          if (COUNT) n_samples++;
          if (!clear) {
            const float oma = 1.0f - acc.w;
            acc.x = fmaf(col.x * oma, col.w, acc.x);
            acc.y = fmaf(col.y * oma, col.w, acc.y);
            acc.z = fmaf(col.z * oma, col.w, acc.z);
            acc.w = fmaf(col.w, oma, acc.w);
            if (acc.w > 0.99f) ray_live = false;
          }
        }
        steps_left -= 1;
        if (ray_live) pc = add3(pc, vdir);
      }
      for (int j = 0; j < h; j++) {   // warp-uniform trip count
        const int hl = hc > j ? (int)__fns(idle_h, 0, first_rank + j + 1) : lane;
        const float cx = __shfl_sync(full, col.x, hl), cy = __shfl_sync(full, col.y, hl), cz = __shfl_sync(full, col.z, hl),
                    cw = __shfl_sync(full, col.w, hl);
        const int fl = __shfl_sync(full, (helper ? 2 : 0) | (clear ? 1 : 0), hl);
        if (hc > j && (fl & 2) != 0 && ray_live && steps_left > 0) {
          n_work++;
          if (COUNT) n_samples++;
          if ((fl & 1) == 0) {
            const float oma = 1.0f - acc.w;
            acc.x = fmaf(cx * oma, cw, acc.x);
            acc.y = fmaf(cy * oma, cw, acc.y);
            acc.z = fmaf(cz * oma, cw, acc.z);
            acc.w = fmaf(cw, oma, acc.w);
            if (acc.w > 0.99f) ray_live = false;
          }
          steps_left -= 1;
          if (ray_live) pc = add3(pc, vdir);
        }
      }
    }
  }
}
#else
template <typename T, int MODE, bool LIT, bool FAST, int BS, bool COUNT, bool PIPE = false>
__global__ void __launch_bounds__(kThreads, TVK_MIN_BLOCKS * 64 / kThreads) raycast_kernel(const __grid_constant__ RayConsts P) {
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  // CTAs are dispatched in blockIdx order (x fastest).  The rays through the middle of the volume are the
  // longest: tile rows / columns are mapped centre-out so those start first (measured: +1 % on C3, the launch
  // is bound by per-warp latency, not by its tail; kept because it costs nothing).
  // Tile schedule (P.tile_order, optional): CTA k of the launch takes tile order[k] -- the tiles sorted by the cost the
  // PREVIOUS frame measured for them, longest first (LPT), so that the CTAs with the longest rays do not start last and
  // leave the device idle behind them.  Which CTA traces which tile does not change any ray.
  uint32_t lin = blockIdx.y * gridDim.x + blockIdx.x;
  if (P.tile_order) lin = __ldg(P.tile_order + lin);
  const uint32_t bxi = lin % gridDim.x, byi = lin / gridDim.x;
  const uint32_t tx = kCentreOut ? centre_out(bxi, gridDim.x) : bxi;
  const uint32_t ty = kCentreOut ? centre_out(byi, gridDim.y) : byi;
  uint32_t n_turns = 0;   // turns of the flat loop this ray took: the tile's cost is the maximum over its rays
  const uint32_t px = tx * (8 * kWX) + (wid % kWX) * 8 + (lane & 7);
  const uint32_t py = ty * (4 * kWY) + (wid / kWX) * 4 + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const size_t pix = (size_t)py * P.width + px;
  constexpr bool ISO = MODE == 2;
  typedef typename PairOf<T>::W W;           // pool element: the x-pair (voxel x, voxel x+1)
  const W* pool = (const W*)P.pool;
  unsigned long long n_samples = 0, n_bricks = 0;
  unsigned long long n_alive_iters = 0, n_warp_iters = 0;   // lane-utilisation diagnostics (count mode)

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  f4 entry4, exit4;
  const bool covered = ray_setup(P, px, py, entry4, exit4, !PIPE);   // a stage takes every ray that meets the VOLUME
  if (!covered) {   // render targets are cleared where no back face is rasterised (GLGridLeaper.cpp:837)
    P.out0[pix] = zero4; P.out1[pix] = zero4; P.out2[pix] = PIPE ? make_float4(0.f, 0.f, 0.f, 1000.0f) : zero4;
    if (ISO) P.out3[pix] = zero4;
    return;
  }
  __shared__ float park[kPark ? kChainWords : 1][kThreads];
  ChainSt c;
  f4 acc;
  f4& resume_col = c.resume_col;
  f4& resume_pos = c.resume_pos;
  f4 hit_pos = from4(zero4), hit_nrm = from4(zero4), resume_nrm = from4(zero4);
  bool done = false;
  bool handoff = false;            // PIPE: the ray left this stage's slab alive
  f4 hand_pos = from4(zero4);
  if (P.first_pass) {
    resume_pos = entry4;
    acc = from4(zero4);
  } else {
    resume_pos = from4(P.ray_start[pix]);
    acc = from4(P.start_color[pix]);
  }
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
    c.entry = F3(resume_pos.x, resume_pos.y, resume_pos.z);
    c.entry_depth = resume_pos.w;
    c.nexit = F3(exit4.x, exit4.y, exit4.z);
    c.exit_depth = exit4.w;
    c.dir = sub3(c.nexit, c.entry);
    c.ray_len = len3(c.dir);
    // TransformToPoolSpace
    const f3 ps = F3(P.pool_size_f);
    f3 vdir = norm3(mul3(c.dir, F3(P.vol_f)));
    vdir = div3(vdir, ps);
    const float den = 2.0f * P.sample_rate;
    vdir = F3(vdir.x / den, vdir.y / den, vdir.z / den);
    c.step = len3(vdir);
    c.t = 0.0f;
    c.optimal = true;
    const float voxel_size = 0.125f / 2000.0f;
    c.cur = c.entry;
    c.j = 0;   // bricks visited by the chain (the shader's j < 100 bound)
    c.lbx = 0; c.lby = 0; c.lbz = 0; c.lbl = 9999;
    const f3 dscale = F3(P.domain_scale), eye_m = F3(P.eye_m), la = F3(P.light_a), ld = F3(P.light_d),
             ls = F3(P.light_s), ldir = F3(P.light_dir_m);

    c.dv = F3(1.0f / c.dir.x, 1.0f / c.dir.y, 1.0f / c.dir.z);   // BrickExit's 1.0/dir
    // the empty-brick advance voxelSize*direction/rayLength
    c.nudge = F3(voxel_size * c.dir.x / c.ray_len, voxel_size * c.dir.y / c.ray_len, voxel_size * c.dir.z / c.ray_len);
    // ---- flat loop state ---------------------------------------------------------------------
    // The shader's nested loops (bricks along the ray / samples inside a brick) run as ONE per-warp loop with
    // two decoupled parts: a brick CHAIN that walks the page table one segment ahead of the sampling (the next
    // segment waits in shared memory), and the SAMPLE phase.  A chain step needs no voxel data -- the position
    // after a brick is its entry + steps sequential adds of the step vector, repeated here exactly as the
    // sample loop does -- so all lanes of a warp can look ahead TOGETHER whenever one lane runs dry, and a lane
    // that finishes its brick early just picks up its waiting segment instead of idling until the slowest lane
    // is done.  Look-ahead never has side effects: a missing brick is only processed (reported, resume state)
    // when the ray really stands at it, so reports, resume points and counters are the shader's.
    __shared__ float seg_f[9][kThreads];       // next segment: pool entry, trans, 1/scale
    __shared__ uint32_t seg_u[COUNT ? 7 : 6][kThreads];    // slot origin (3), slot index, steps, flags (, page-table index)
    bool ray_live = c.ray_len > voxel_size;   // the ray has not terminated (ERT / iso hit)
    bool chain = ray_live;                  // the brick chain has not reached the end of the ray
    bool have_next = false;                 // a prefetched segment waits in shared memory
    int steps_left = 0;      // samples left in the current brick
    bool b_partial = false;  // sort-last: the current brick straddles the shard box (ownership per sample)
    f3 pc = c.entry, b_trans = c.entry, b_inv = c.entry;
    uint32_t b_ox = 0, b_oy = 0, b_oz = 0;
    const W* vox = pool;
    constexpr bool GRAD = !ISO && (MODE == 1 || LIT);   // what a sample needs: the 7-tap footprint or the centre tap
    Foot<T, FAST, BS, GRAD> cur;   // footprint of the sample at pc, fetched one turn ahead
    bool cur_ok = false;
    unsigned long long pend = 0;   // COUNT: brick visits of the chain that the sampling has not reached yet

    if (kPark) { park_const(c, park, tid); park_var(c, park, tid); }
    while (ray_live) {
      n_turns++;
      // ---- chain phase: runs for the whole warp when some lane can neither sample nor pick up a segment
      const unsigned act = __activemask();
      const bool need = steps_left == 0 && !have_next && chain;
      if (__ballot_sync(act, need) != 0u && chain && !have_next) {
        if (kPark) unpark(c, park, tid);
#pragma unroll 1
        for (int f = 0; f < 4 && chain && !have_next; f++) {
          if (c.j >= 100) { chain = false; break; }
          if (P.shard) {   // the block is convex: once the ray has left it there is nothing more to do on this rank
            const bool gone = (c.dir.x > 0.0f && c.cur.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && c.cur.x <= P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && c.cur.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && c.cur.y <= P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && c.cur.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && c.cur.z <= P.sh_lo[2]);
            if (gone) {
              if (PIPE) {   // where the next stage picks the ray up
                handoff = true;
                hand_pos.x = c.cur.x; hand_pos.y = c.cur.y; hand_pos.z = c.cur.z;
                hand_pos.w = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
              }
              chain = false; break;
            }
          }
          const float cur_depth = c.entry_depth * (1.0f - c.t) + c.exit_depth * c.t;
          uint32_t lod = compute_lod(P, cur_depth);
          BrickRef b;
          int ok;
          if (steps_left > 0) {   // look-ahead: the ray is still sampling the previous segment
            ok = get_brick<true>(P, c.cur, lod, c.dir, c.dv, b);
            if (ok < 0) break;    // missing brick: handled when the ray stands here
          } else {
            ok = get_brick<false>(P, c.cur, lod, c.dir, c.dv, b);
            if (!ok && c.optimal) {
              c.optimal = false;
              resume_pos.x = c.cur.x; resume_pos.y = c.cur.y; resume_pos.z = c.cur.z; resume_pos.w = cur_depth;
              if (!ISO) resume_col = acc;
            }
          }
          if (COUNT) pend++;
          if (!b.empty && !(c.lbx == b.bx && c.lby == b.by && c.lbz == b.bz && c.lbl == b.bl)) {
            int steps = (int)ceilf(len3(sub3(b.pool_exit, b.pool_entry)) / c.step);
            const int s2 = (int)ceilf(len3(mul3(sub3(c.nexit, c.cur), b.scale)) / c.step);
            steps = min(steps, s2);
            const f3 inv = F3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
            f3 pe = b.pool_entry;
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
            if (b.where == OUTSIDE_SHARD) {   // another rank's brick: advance by its steps, take no sample
              const float n = (float)max(steps, 0);
              pe = F3(fmaf(n, vdir.x, pe.x), fmaf(n, vdir.y, pe.y), fmaf(n, vdir.z, pe.z));
              steps = 0;
            }
            if (steps > 0) {
              seg_f[0][tid] = pe.x; seg_f[1][tid] = pe.y; seg_f[2][tid] = pe.z;
              seg_f[3][tid] = b.trans.x; seg_f[4][tid] = b.trans.y; seg_f[5][tid] = b.trans.z;
              seg_f[6][tid] = inv.x; seg_f[7][tid] = inv.y; seg_f[8][tid] = inv.z;
              seg_u[0][tid] = b.ox; seg_u[1][tid] = b.oy; seg_u[2][tid] = b.oz;
              seg_u[3][tid] = b.slot; seg_u[4][tid] = (uint32_t)steps;
              seg_u[5][tid] = b.where == PARTLY_IN_SHARD ? 1u : 0u;
              if (COUNT) seg_u[6][tid] = b.id;
              have_next = true;
              // where the sample loop will leave pc: `steps` SEQUENTIAL adds (fp32 addition is not associative, so
              // there is no closed form); 4x unrolled -- this loop was 6.5 % of the launch's warp instructions
#pragma unroll 4
              for (int i = 0; i < steps; i++) pe = add3(pe, vdir);
            }
            c.cur = mul3(sub3(pe, b.trans), inv);
          } else {
            c.cur = add3(b.norm_exit, c.nudge);
            c.lbx = b.bx; c.lby = b.by; c.lbz = b.bz; c.lbl = b.bl;
          }
          c.t = len3(sub3(c.entry, b.norm_exit)) / c.ray_len;
          c.j++;
          if (c.t > 0.9999f) chain = false;
        }
        if (kPark) park_var(c, park, tid);
      }
      // ---- pick up the waiting segment
      if (steps_left == 0 && have_next) {
        pc = F3(seg_f[0][tid], seg_f[1][tid], seg_f[2][tid]);
        b_trans = F3(seg_f[3][tid], seg_f[4][tid], seg_f[5][tid]);
        b_inv = F3(seg_f[6][tid], seg_f[7][tid], seg_f[8][tid]);
        b_ox = seg_u[0][tid]; b_oy = seg_u[1][tid]; b_oz = seg_u[2][tid];
        vox = pool + (uint64_t)seg_u[3][tid] * P.slot_voxels;
        steps_left = (int)seg_u[4][tid];
        b_partial = seg_u[5][tid] != 0u;
        have_next = false;
        cur_ok = false;
        if (COUNT) {
          n_bricks += pend; pend = 0;
          if (P.visited) { const uint32_t id = seg_u[6][tid]; atomicOr(P.visited + (id >> 5), 1u << (id & 31)); }
        }
      }
      if (steps_left == 0 && !chain) {   // the ray left the volume (or its 100-brick budget) unterminated
        if (COUNT) n_bricks += pend;
        ray_live = false;
      }
      if (COUNT) {
        n_alive_iters += ray_live ? 1 : 0;
        if (__ffs(__activemask()) - 1 == (tid & 31)) n_warp_iters++;
      }
      // ---- sample phase: one sample for every lane that is inside a brick ----
      if (ray_live && steps_left > 0) {
        bool terminated = false;
        // software pipeline: the footprint of THIS sample was fetched during the previous turn (cur_ok), the one of the
        // next sample of the brick is fetched now, before this sample's arithmetic, so its load latency is covered
        if (!kPrefetch || !FAST || !cur_ok) cur.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
        const f3 pc_next = add3(pc, vdir);
        const bool pf = kPrefetch && FAST && steps_left > 1;
        Foot<T, FAST, BS, GRAD> nxt;
        if (pf) nxt.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc_next);
        bool mine = true;
        if (b_partial) {
          const f3 mq = mul3(sub3(pc, b_trans), b_inv);
          mine = mq.x >= P.sh_lo[0] && mq.x < P.sh_hi[0] && mq.y >= P.sh_lo[1] && mq.y < P.sh_hi[1] &&
                 mq.z >= P.sh_lo[2] && mq.z < P.sh_hi[2];
          if (PIPE && !ISO && !mine) {
            // a brick of a coarser LoD straddles the slab's far side: the ray is handed on AT the side, not
            // behind the brick, so the next stage takes the brick's remaining samples
            const bool gone = (c.dir.x > 0.0f && mq.x >= P.sh_hi[0]) || (c.dir.x < 0.0f && mq.x < P.sh_lo[0]) ||
                              (c.dir.y > 0.0f && mq.y >= P.sh_hi[1]) || (c.dir.y < 0.0f && mq.y < P.sh_lo[1]) ||
                              (c.dir.z > 0.0f && mq.z >= P.sh_hi[2]) || (c.dir.z < 0.0f && mq.z < P.sh_lo[2]);
            if (gone) {
              const float tq = len3(sub3(mq, c.entry)) / c.ray_len;
              handoff = true;
              hand_pos.x = mq.x; hand_pos.y = mq.y; hand_pos.z = mq.z;
              hand_pos.w = c.entry_depth * (1.0f - tq) + c.exit_depth * tq;
              terminated = true;   // leaves the loop; TerminateRay sees alpha <= 0.99 and hands the ray on
            }
          }
        }
        if constexpr (!ISO) {
          if (mine) {
            if (COUNT) n_samples++;
            // ComputeColorFromVolume + OpacityCorrectColor at pool position pc
            f4 col;
            bool clear = false;
            if constexpr (MODE == 0 && !LIT) {
              col = tf_lookup(P, cur.centre(P) * P.trans_scale, 0.0f);
            } else if constexpr (MODE == 0) {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              col = tf_lookup(P, data * P.trans_scale, 0.0f);
              // A sample whose transfer-function alpha is exactly 0 leaves the ray unchanged bit for bit
              // (UnderCompositing adds colour * (1-a) * 0 = 0 to every channel; table colours and lit colours are
              // finite, and opacity correction maps 0 to 0), so its normal and lighting are not computed when the
              // whole warp agrees.  The shader cannot branch this cheaply; the result is identical.
              clear = kSkipClear && col.w == 0.0f;
              if (!clear) {
                f3 n = mul3(g, dscale);   // ComputeNormal
                const float l = len3(n);
                if (l > 0.0f) n = scl3(n, 1.0f / l);
                const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                const f3 lit = lighting(eye_m, mp, n, la, mul3(F3(col.x, col.y, col.z), ld), ls, ldir);
                col.x = lit.x; col.y = lit.y; col.z = lit.z;
              }
            } else {
              float data; f3 g;
              cur.sample_with_gradient(P, data, g);
              const float gm = len3(g);
              col = tf_lookup(P, data * P.trans_scale, 1.0f - gm * P.gradient_scale);
              if (LIT) {
                clear = kSkipClear && col.w == 0.0f;
                if (!clear) {
                  const f3 gn = gm > 0.0f ? scl3(g, 1.0f / gm) : g;
                  const f3 n = mul3(dscale, gn);
                  const f3 mp = mul3(sub3(pc, b_trans), b_inv);
                  float dl, sp;
                  light_terms(eye_m, mp, n, ldir, dl, sp);
                  const f3 lit = light_apply(la, mul3(F3(col.x, col.y, col.z), ld), ls, dl, sp);
                  col.x = lit.x; col.y = lit.y; col.z = lit.z;
                }
              }
            }
            if (!clear) {
              col.w = opacity_correct(P, col.w);
              // UnderCompositing
              const float oma = 1.0f - acc.w;
              acc.x = fmaf(col.x * oma, col.w, acc.x);
              acc.y = fmaf(col.y * oma, col.w, acc.y);
              acc.z = fmaf(col.z * oma, col.w, acc.z);
              acc.w = fmaf(col.w, oma, acc.w);
              if (acc.w > 0.99f) terminated = true;
            }
          }
        } else if (mine) {   // isosurface march
          if (COUNT) n_samples++;
          if (cur.centre(P) >= P.isoval) {
            // RefineIsosurface
            f3 rd = F3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);
            pc = sub3(pc, rd);
            Foot<T, FAST, BS, false> rf;
#pragma unroll 1
            for (int k = 0; k < 5; k++) {
              rd = F3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
              rf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
              if (rf.centre(P) >= P.isoval) pc = sub3(pc, rd); else pc = add3(pc, rd);
            }
            const f3 hp = mul3(sub3(pc, b_trans), b_inv);
            hit_pos = xform4(P.m2e, hp.x, hp.y, hp.z, 1.0f);
            hit_pos.w = 1.0f + 1.0f;   // color.r + 1
            Foot<T, FAST, BS, true> gf;
            gf.fetch(P, pool, vox, b_ox, b_oy, b_oz, pc);
            float dummy; f3 g;
            gf.sample_with_gradient(P, dummy, g);
            f3 n = mul3(g, dscale);
            const float l = len3(n);
            if (l > 0.0f) n = scl3(n, 1.0f / l);
            const float* m = P.mv_inv;   // mModelViewIT * vec4(n, 0)
            hit_nrm.x = m[0] * n.x + m[1] * n.y + m[2] * n.z;
            hit_nrm.y = m[4] * n.x + m[5] * n.y + m[6] * n.z;
            hit_nrm.z = m[8] * n.x + m[9] * n.y + m[10] * n.z;
            hit_nrm.w = floorf(1.0f * 512.0f) + 1.0f;   // floor(color.g*512)+color.b
            terminated = true;
          } else {
            hit_pos = from4(zero4);
          }
        }
        steps_left -= 1;
        if (terminated) ray_live = false;
        else {
          pc = pc_next;
          if (pf) cur = nxt;
          cur_ok = pf;
        }
      }
    }
    if (kPark) unpark_result(c, park, tid);
    // TerminateRay
    if (!ISO) {
      if (c.optimal) {
        // ray_live is false only after early termination; a ray that ran out of bricks in this slab is handed on
        if (PIPE && handoff && !(acc.w > 0.99f)) { resume_pos = hand_pos; resume_col = acc; }
        else { resume_pos.w = 1000.0f; resume_col = acc; }
      }
    } else {
      if (c.optimal) resume_pos.w = hit_pos.w == 0.0f ? 1000.0f : 499.0f + hit_pos.w;
      resume_nrm = hit_nrm;
    }
  }
  if (P.tile_cost) {
    const unsigned m = __activemask();
    const uint32_t w = __reduce_max_sync(m, n_turns);
    if ((uint32_t)(__ffs(m) - 1) == (uint32_t)lane) atomicMax(P.tile_cost + lin, w);
  }
  if (!ISO) {
    P.out0[pix] = to4(acc); P.out1[pix] = to4(resume_col); P.out2[pix] = to4(resume_pos);
  } else {
    P.out0[pix] = to4(hit_pos); P.out1[pix] = to4(hit_nrm); P.out2[pix] = to4(resume_pos);
    P.out3[pix] = to4(resume_nrm);
  }
  if (COUNT) {
    atomicAdd(P.counters + 0, n_samples); atomicAdd(P.counters + 1, 1ull); atomicAdd(P.counters + 2, n_bricks);
    atomicAdd(P.counters + 3, n_alive_iters); atomicAdd(P.counters + 4, n_warp_iters);
    atomicMax(P.counters + 5, n_alive_iters);
  }
}

#endif

// ---- fetch-path ceiling ------------------------------------------------------------------------------------------
// What the traversal kernel's OWN fetch path can deliver when nothing else is in the loop: the same warp tiles (8x4
// rays, one voxel apart), the same 0.5-voxel steps, the same FastFoot loads + packed filter trees on the resident pool,
// but no page-table walk, classification, shading or compositing.  Rays march through a slot along `dir` and hop to
// another slot when they leave it (so the working set is the whole pool, far larger than L2).  bench.py divides the
// kernel's sample rate by this rate: the fetch fraction of the roofline object.
template <typename T, int BS, bool GRAD>
__global__ void __launch_bounds__(kThreads) fetch_probe_kernel(const __grid_constant__ RayConsts P, uint32_t n_slots, uint32_t steps,
                                                               float dx, float dy, float dz, float* out) {
  typedef typename PairOf<T>::W W;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t px = blockIdx.x * (8 * kWX) + (wid % kWX) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * (4 * kWY) + (wid / kWX) * 4 + (lane >> 3);
  if (px >= P.width || py >= P.height) return;
  const W* pool = (const W*)P.pool;
  const uint32_t tile = (blockIdx.y * gridDim.x + blockIdx.x) * (kWX * kWY) + wid;
  uint32_t slot = (tile * 2654435761u) % n_slots;
  const float bs = (float)(BS ? BS : (int)P.total[0]);
  // lane offsets across the ray bundle: perpendicular to the march direction's dominant axis
  const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
  f3 u, v;
  if (az >= ax && az >= ay) { u = F3(1.f, 0.f, 0.f); v = F3(0.f, 1.f, 0.f); }
  else if (ay >= ax) { u = F3(1.f, 0.f, 0.f); v = F3(0.f, 0.f, 1.f); }
  else { u = F3(0.f, 1.f, 0.f); v = F3(0.f, 0.f, 1.f); }
  const float lu = (float)(lane & 7), lv = (float)(lane >> 3);
  const float inv = 0.5f / sqrtf(dx * dx + dy * dy + dz * dz);
  const f3 step = F3(dx * inv, dy * inv, dz * inv);                  // 0.5 voxel per step
  const f3 start = F3(4.f + u.x * lu + v.x * lv, 4.f + u.y * lu + v.y * lv, 4.f + u.z * lu + v.z * lv);
  f3 q = start;                                                       // voxel coordinates inside the slot
  float acc = 0.0f;
  for (uint32_t i = 0; i < steps; i++) {
    if (q.x < 2.f || q.y < 2.f || q.z < 2.f || q.x > bs - 3.f || q.y > bs - 3.f || q.z > bs - 3.f) {
      q = start;                                                      // left the brick: next slot (warp-coherent hop)
      slot = (slot + 9973u) % n_slots;
    }
    const W* vox = pool + (uint64_t)slot * P.slot_voxels;
    FastFoot<T, BS, GRAD> f;
    f.fetch(P, vox, 0u, 0u, 0u, F3(q.x / bs, q.y / bs, q.z / bs));
    if constexpr (GRAD) {
      float data; f3 g;
      f.sample_with_gradient(P, data, g);
      acc += data + g.x + g.y + g.z;
    } else {
      acc += f.centre(P);
    }
    q = add3(q, step);
  }
  out[(size_t)py * P.width + px] = acc;
}

template <typename T, int MODE, bool LIT>
void launch_t(const RayConsts& rc, cudaStream_t s) {
  dim3 block(kThreads);
#if TVK_PERSIST
  // persistent warps: every CTA slot of the device once (launch bounds: TVK_MIN_BLOCKS * 64 / kThreads CTAs per SM), never
  // more warps than there are tiles
  const uint32_t n_tiles = ((rc.width + 15u) / 16u) * ((rc.height + 7u) / 8u) * 4u;
  dim3 grid(std::max(1u, std::min((uint32_t)kSMs * (uint32_t)(TVK_MIN_BLOCKS * 64 / kThreads), (n_tiles + (kThreads / 32) - 1) / (kThreads / 32))));
  cudaMemsetAsync(rc.tile_counter, 0, sizeof(uint32_t), s);
#else
  dim3 grid((rc.width + 8 * kWX - 1) / (8 * kWX), (rc.height + 4 * kWY - 1) / (4 * kWY));
#endif
  // FAST addressing needs the +-1 gradient taps of every legal sample position inside the slot
  bool fast = !rc.nearest;
  for (int i = 0; i < 3; i++) fast = fast && rc.total[i] >= 4 && rc.ghost[i] >= 2;
  const bool b36 = rc.total[0] == 36 && rc.total[1] == 36 && rc.total[2] == 36;
  if (rc.count && rc.pipeline && MODE != 2) {
    if (fast) raycast_kernel<T, MODE, LIT, true, 0, true, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, true, true><<<grid, block, 0, s>>>(rc);
  } else if (rc.count) {
    if (fast) raycast_kernel<T, MODE, LIT, true, 0, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, true><<<grid, block, 0, s>>>(rc);
  } else if (rc.pipeline && MODE != 2) {   // depth-pipeline stage (DVR modes)
    if (fast && b36) raycast_kernel<T, MODE, LIT, true, 36, false, true><<<grid, block, 0, s>>>(rc);
    else if (fast) raycast_kernel<T, MODE, LIT, true, 0, false, true><<<grid, block, 0, s>>>(rc);
    else raycast_kernel<T, MODE, LIT, false, 0, false, true><<<grid, block, 0, s>>>(rc);
  } else if (fast && b36) raycast_kernel<T, MODE, LIT, true, 36, false><<<grid, block, 0, s>>>(rc);
  else if (fast) raycast_kernel<T, MODE, LIT, true, 0, false><<<grid, block, 0, s>>>(rc);
  else raycast_kernel<T, MODE, LIT, false, 0, false><<<grid, block, 0, s>>>(rc);
}

template <typename T>
void launch_d(const RayConsts& rc, int mode, int lighting, cudaStream_t s) {
  if (mode == TVK_RM_ISOSURFACE) launch_t<T, 2, false>(rc, s);
  else if (mode == TVK_RM_1DTRANS) { if (lighting) launch_t<T, 0, true>(rc, s); else launch_t<T, 0, false>(rc, s); }
  else { if (lighting) launch_t<T, 1, true>(rc, s); else launch_t<T, 1, false>(rc, s); }
}

}  // namespace

// rc: pool, slot_voxels, total, norm, width, height are used; the pool is addressed as a one-slot atlas per brick
void launch_fetch_probe(const RayConsts& rc_in, int dtype, bool grad, uint32_t n_slots, uint32_t steps, const float dir[3],
                        float* out, cudaStream_t s) {
  RayConsts rc = rc_in;
  for (int i = 0; i < 3; i++) rc.pool_size_f[i] = (float)rc.total[i];
  dim3 block(kThreads);
  dim3 grid((rc.width + 8 * kWX - 1) / (8 * kWX), (rc.height + 4 * kWY - 1) / (4 * kWY));
  const bool b36 = rc.total[0] == 36 && rc.total[1] == 36 && rc.total[2] == 36;
#define TVK_PROBE(T, BSV)                                                                                              \
  do {                                                                                                                 \
    if (grad) fetch_probe_kernel<T, BSV, true><<<grid, block, 0, s>>>(rc, n_slots, steps, dir[0], dir[1], dir[2], out); \
    else fetch_probe_kernel<T, BSV, false><<<grid, block, 0, s>>>(rc, n_slots, steps, dir[0], dir[1], dir[2], out);    \
  } while (0)
  switch (dtype) {
    case TVK_U8: if (b36) TVK_PROBE(uint8_t, 36); else TVK_PROBE(uint8_t, 0); break;
    case TVK_U16: if (b36) TVK_PROBE(uint16_t, 36); else TVK_PROBE(uint16_t, 0); break;
    default: if (b36) TVK_PROBE(float, 36); else TVK_PROBE(float, 0); break;
  }
#undef TVK_PROBE
}

uint32_t raycast_tiles(uint32_t width, uint32_t height) {
  return ((width + 8 * kWX - 1) / (8 * kWX)) * ((height + 4 * kWY - 1) / (4 * kWY));
}

void launch_raycast(const RayConsts& rc, int mode, int lighting, int dtype, cudaStream_t s) {
  switch (dtype) {
    case TVK_U8: launch_d<uint8_t>(rc, mode, lighting, s); break;
    case TVK_U16: launch_d<uint16_t>(rc, mode, lighting, s); break;
    default: launch_d<float>(rc, mode, lighting, s); break;
  }
}

}  // namespace tvk
