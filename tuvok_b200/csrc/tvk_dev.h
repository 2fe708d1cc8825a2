// tvk_dev.h -- structs shared by the host layer and the sm_100a kernels of libtvkcuda.so.
#ifndef TVK_DEV_H
#define TVK_DEV_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tvk.h"

namespace tvk {

// Everything the traversal kernel reads that GLGridLeaper::SetupRaycastShader
// (GLGridLeaper.cpp:690-752), GLVolumePool::Enable (GLVolumePool.cpp:790-803) and the
// generated pool GLSL (#defines at GLVolumePool.cpp:364-463) hand to the shader.
// Passed by value as a __grid_constant__ kernel parameter (constant bank, broadcast reads).
struct RayConsts {
  uint32_t width, height;
  float emm[16];        // mEyeToModel
  float inv_proj[16];   // inverse projection (near-plane entry, GLGridLeaper-NearPlane-VS.glsl)
  float m2e[16];        // mModelToEye
  float mv_inv[16];     // inverse(modelView): mModelViewIT is its transpose
  float domain_scale[3];
  float light_a[3], light_d[3], light_s[3], light_dir_m[3], eye_m[3];
  float lzwse;          // fLevelZeroWorldSpaceError
  float lod_factor;     // fLoDFactor
  float pool_size_f[3]; // iPoolSize (virtual atlas = capacity * max brick)
  float vol_f[3];       // volumeSize
  float overlap_tc[3];  // overlap in pool texcoords
  uint32_t capacity[3];
  uint32_t total[3];    // maxTotalBrickSize
  uint32_t ghost[3];    // brick overlap per side
  uint32_t lod_count;   // pool LoD count
  uint32_t lod_offset[TVK_MAX_LOD];
  float lod_layout[TVK_MAX_LOD][3];     // vLODLayout
  uint32_t lod_layout_sz[TVK_MAX_LOD][2]; // iLODLayoutSize
  float norm;           // unorm -> float
  float oc;             // ocFactor = 1/sampleRateModifier
  float sample_rate;
  float trans_scale, gradient_scale, isoval;
  uint32_t tf_w, tf_h;
  uint32_t hash_size, rehash_count;
  int32_t strategy;
  uint32_t finest[3];   // finest brick layout (hash serialisation)
  float clip_min[3], clip_max[3];   // sort-last shard box (ray hit test)
  int32_t shard;                    // shard box != whole volume
  float sh_lo[3], sh_hi[3];         // shard box with faces on the volume border pushed to -/+inf
  int32_t clip_plane_on;            // bbox cut by the user clip plane (GLGridLeaper::FillBBoxVBO)
  float clip_plane[4];              // ... in the [0,1]^3 coordinates of the ray set-up: kept where dot(xyz, p) + w <= 0
  int32_t nearest;
  int32_t first_pass;   // region is blank: ray entry computed, start colour = 0
  int32_t pipeline;     // this launch is a stage of the depth pipeline (k_raycast.cu PIPE)
  int32_t count;        // accumulate counters
  // device pointers
  const void* pool;       // slot-linear brick pool, x-pair layout: element x = (voxel x, voxel x+1) (k_pool.cu)
  uint64_t slot_voxels;   // voxels per slot
  const uint32_t* meta;   // page table
  const uint32_t* tf;     // RGBA8 table (byte / 255.0f is formed per fetch, tvk_math.cuh unorm8x4)
  uint32_t* hash;         // miss-report table
  const float4* ray_start;   // resume position (in), ignored when first_pass
  const float4* start_color; // resume colour / normal (in)
  float4* out0;  // DVR accRayColor   | ISO rayHitPos
  float4* out1;  // DVR rayResumeColor| ISO rayHitNormal
  float4* out2;  // rayResumePos
  float4* out3;  // ISO rayResumeNormal
  unsigned long long* counters; // samples, rays, brick visits
  uint32_t* visited;            // bitmap over page-table indices of sampled bricks (counting only)
  uint32_t* tile_counter;       // persistent traversal kernel: the next tile of the launch (zeroed by the launcher)
  const uint32_t* tile_order;   // tile schedule of this launch (CTA k -> tile order[k]) or NULL = dispatch order
  uint32_t* tile_cost;          // per tile: the largest number of loop turns one of its rays took (input of the next schedule)
};

// launchers (defined in the .cu files)
void launch_raycast(const RayConsts& rc, int mode, int lighting, int dtype, cudaStream_t s);
// fetch-path ceiling probe (k_raycast.cu): `steps` fetch + filter evaluations per ray on the resident pool, nothing else
void launch_fetch_probe(const RayConsts& rc, int dtype, bool grad, uint32_t n_slots, uint32_t steps, const float dir[3],
                        float* out, cudaStream_t s);
void launch_iso_compose(const float4* hit_pos, const float4* hit_nrm, float4* rgba, uint32_t w, uint32_t h,
                        const float amb[3], const float dif[3], const float spe[3], const float ldir[3],
                        cudaStream_t s, bool color = false);
// colour volumes (k_color.cu): GLGridLeaper-Method-*-color.glsl on a plain uchar4 pool
void launch_raycast_color(const RayConsts& rc, int mode, int lighting, cudaStream_t s);
void launch_quantize_rgba8(const float4* src, uchar4* dst, uint64_t n, cudaStream_t s);
// Compose-CV-FS.glsl over the four hit targets (GLRenderer::ComposeSurfaceImage, ClearView branch)
void launch_cv_compose(const float4* hit_pos, const float4* hit_nrm, const float4* cv_pos, const float4* cv_nrm, float4* rgba,
                       uint32_t w, uint32_t h, const float amb[3], const float dif[3], const float dif2[3], const float spe[3],
                       const float ldir[3], const float cv_param[3], const float pick[3], cudaStream_t s);
// GLRenderer::EndFrame's eye composition (Compose-{Anaglyphs,Scanline,SBS,AF}-FS.glsl); mode = tvk_stereo_mode
void launch_stereo_compose(int mode, const float4* left, const float4* right, float4* out, uint32_t w, uint32_t h,
                           int alternating_frame_id, float split_coord, cudaStream_t s);
void launch_composite_over(const float4* front, const float4* back, float4* out, uint64_t n, cudaStream_t s);
// sort-last direct send: fold the n partial images of one pixel slice front to back (src[0] = frontmost) -> RGBA32F + RGBA8
#define TVK_MAX_RANKS 16
struct NWaySrc { const float4* src[TVK_MAX_RANKS]; int n; };
void launch_nway_over(const NWaySrc& a, float4* out_f, uchar4* out8, uint64_t n, cudaStream_t s);
// Peer-memory variant (all GPUs of one box, NVLink / NVSwitch): the partial images are read where the peers rendered
// them and the RGBA8 slice is stored straight into rank 0's frame -- no copy, no NCCL call on the frame's path.  Ranks
// synchronise through flag words in each other's memory: flags[p] = rank p's flag block (mapped into this process),
// uint32 [3][TVK_MAX_RANKS]: READY[q] = last frame whose partial image rank q has finished, CONSUMED[q] = last frame
// whose slices rank q has read out of rank p's image, GATHERED[q] = last frame whose RGBA8 slice rank q has stored (rank 0).
enum { TVK_SLF_READY = 0, TVK_SLF_CONSUMED = 1, TVK_SLF_GATHERED = 2 };
struct SlPeer { int n, self; uint32_t* flags[TVK_MAX_RANKS]; };
// local[0] = block counter, local[1] = time-out marker (a peer never signalled: the frame is garbage, nothing hangs)
void launch_sl_wait(const SlPeer& P, int kind, uint32_t frame, uint32_t* local, cudaStream_t s);
void launch_sl_signal(const SlPeer& P, int kind, uint32_t frame, cudaStream_t s);
void launch_nway_over_peer(const NWaySrc& a, float4* out_f, uchar4* out8, uint64_t n, const SlPeer& P, uint32_t frame,
                           uint32_t* local, cudaStream_t s);

// classic per-brick raycaster (k_classic.cu): uniforms of GLRaycaster::SetBrickDepShaderVars / RenderBox plus the
// per-axis brick tables of the LoD (the brick boxes of one LoD are a tensor-product grid, so everything the
// per-brick passes need is stored per axis: [3][axis_stride])
struct ClassicConsts {
  uint32_t width, height;
  float inv_proj[16], imv[16];      // inverse projection, inverse(modelView)
  int32_t ortho;                    // the projection is parallel (w' independent of z): HQ MIP frames under m_bOrthoView
  float domain_scale[3], light_a[3], light_d[3], light_s[3], light_dir[3];
  float norm, trans_scale, gradient_scale, step_scale;
  uint32_t tf_w, tf_h;
  const uint32_t* tf;
  int32_t nearest, count;
  uint32_t layout[3];               // bricks per axis of the LoD
  uint32_t total[3];                // slot strides (max brick size)
  uint32_t axis_stride;
  const float* plane;               // [3][n+1] cell planes shared by neighbouring bricks
  const float* pmin; const float* pmax;   // [3][n] brick box center -+ extension/2
  const float* tmax; const float* tsc;    // [3][n] texcoord of the max corner, (tMin - tMax) / (pMin - pMax)
  const float* rstep;               // [3][n] extension * (1/voxelCount) * (0.5/sampleRate); the ray step is their min
  const uint32_t* nvox;             // [3][n] brick size incl. ghost
  const uint32_t* table;            // per brick of the LoD (x fastest): 0 = not rendered, else pool slot + 1
  const void* pool;
  uint64_t slot_voxels;
  float4* out;
  float isoval, proj_param[2];      // isosurface mode: fIsoval (normalised), vProjParam = (f/(f-n), f*n/(n-f))
  float4* out_nrm;                  // isosurface mode: second iso-hit target (normal, brick number in the list)
  float cv_isoval;                  // ClearView: GetNormalizedCVIsovalue
  float4* out_cv; float4* out_cv_nrm;   // ClearView: m_pFBOCVHit's two targets (nullptr = ClearView off)
  const uint32_t* list_pos;         // isosurface mode: per brick of the LoD, its position in the frame's brick list (iTileID)
  float2* out_max;                  // HQ MIP only: (blended maximum, coverage) per pixel, the FBO Transfer-MIP reads
  unsigned long long* counters;
};
// mode: TVK_RM_1DTRANS / TVK_RM_2DTRANS, or TVK_CLASSIC_MIP for the HQ MIP frame
#define TVK_CLASSIC_MIP 3
void launch_classic(const ClassicConsts& c, int mode, int lighting, int dtype, cudaStream_t s);

// page table / visibility (k_pool.cu)
struct VisConsts {
  int32_t mode;
  double v[4];                   // 1D: min,max | 2D: min,max,gmin,gmax | ISO: iso
  uint32_t lod_count;
  uint32_t lod_offset[TVK_MAX_LOD];
  uint32_t layout[TVK_MAX_LOD][3];
};
void launch_vis_clear(uint32_t* meta, uint64_t n, cudaStream_t s);
void launch_vis_pool(uint32_t* meta, const int32_t* slot_brick, uint32_t n_slots, const double* minmax,
                     const VisConsts& vc, cudaStream_t s);
void launch_vis_level(uint32_t* meta, const double* minmax, const VisConsts& vc, uint32_t lod,
                      uint32_t* counts, cudaStream_t s);
struct PageOp { uint32_t evict_id; uint32_t new_id; uint32_t slot; uint32_t pad; uint64_t src_off; uint32_t size[3]; uint32_t pad2; };
void launch_page_meta(uint32_t* meta, const PageOp* ops, uint32_t n, cudaStream_t s);
void launch_page_copy(void* pool, const void* store, const PageOp* ops, uint32_t n, uint64_t slot_voxels,
                      uint32_t esize, const uint32_t total[3], int src_is_slot_layout, cudaStream_t s);
// plain voxels of one pool slot (the first halves of its pairs) -> out (device)
void launch_slot_unpair(const void* slot, void* out, uint64_t n_voxels, uint32_t esize, cudaStream_t s);
// tile schedule: order = the n tiles sorted by cost, largest first (counting sort over cost >> shift, one CTA)
void launch_tile_order(const uint32_t* cost, uint32_t n, uint32_t* order, uint32_t shift, uint32_t split_cost, cudaStream_t s);
// CTAs of a traversal launch (k_raycast.cu)
uint32_t raycast_tiles(uint32_t width, uint32_t height);
void launch_hash_compact(const uint32_t* hash, uint32_t n, uint32_t* out_list, uint32_t* out_count, cudaStream_t s);

// bricker (k_bricker.cu)
void launch_synth(void* dst, int kind, const uint32_t size[3], int dtype, uint32_t seed, cudaStream_t s);
void launch_downsample(const void* src, const uint32_t ss[3], void* dst, const uint32_t ds[3], int dtype, int median,
                       cudaStream_t s);
struct CutConsts {
  uint32_t lod_size[3], layout[3], brick[3], overlap;
  int32_t clamp, lod;
  uint64_t first_brick;   // TOC index of brick (0,0,0) of this LOD
};
// procedural multi-resolution dataset (tvk_procedural.inc): the geometry of every pool LoD + the generator's parameters
struct ProcConsts {
  int32_t kind, dtype;
  uint32_t seed, lod_count, overlap;
  uint32_t brick[3];
  uint32_t lod_size[TVK_MAX_LOD][3], layout[TVK_MAX_LOD][3];
  uint32_t lod_offset[TVK_MAX_LOD];
};
// min/max (incl. ghost, 0 outside the level's grid) of the procedural bricks [first, first + count) -> minmax[4 * (id - first)]
void launch_proc_minmax(const ProcConsts& pc, uint32_t first, uint32_t count, double* minmax, cudaStream_t s);
// min/max of n staged bricks (ops[i]: src_off, size, new_id = TOC index) -> minmax[4 * new_id]
void launch_brick_minmax(const void* staged, const PageOp* ops, uint32_t n, double* minmax, int dtype, cudaStream_t s);
// store_index (device, TOC index -> store slot or -1) or nullptr: the store holds every brick at its TOC index
void launch_cut_bricks(const void* lod_vol, void* store, const int32_t* store_index, double* minmax, const CutConsts& cc,
                       int dtype, uint64_t slot_bytes, cudaStream_t s);

// re-bricking of a staged source brick into the target bricks it holds (k_bricker.cu)
struct RebrickConsts {
  uint32_t lod_size[3], layout[3] /* target */, brick[3] /* target, incl. ghost */, overlap;
  uint32_t src_size[3] /* this source brick's own size incl. ghost */, src_brick[3], src_inner[3], ratio[3];
  uint64_t first_brick;   // first TOC index of the target level
};
void launch_rebrick(const void* src, const RebrickConsts& C, void* store, double* minmax, int dtype, uint64_t slot_bytes, cudaStream_t s);
// value quantiser (k_quantize.cu)
struct QuantParams { double f, fh; uint32_t max_out, bins; int32_t mode, out_bits; };
int quant_blocks();   // CTAs of the range pass = (min, max) pairs it writes
void launch_quant_minmax(const void* src, int type, uint64_t n, void* part_d, cudaStream_t s);
void launch_quant_map(const void* src, int type, uint64_t n, double mn, const QuantParams& P, void* dst, unsigned long long* hist,
                      cudaStream_t s);

}  // namespace tvk
#endif
