/*
 * orc.h -- CPU ORACLE for the Tuvok brick-pool raycaster hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  It is a plain restatement of the reference's
 * algorithm (SCIInstitute/Tuvok, citations are file:line relative to the
 * reference root) used ONLY as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * tuvok_b200/ may include, link or call it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - data side (bricking / LOD pyramid / min-max): PINNED against the
 *     reference's own ExtendedOctreeConverter compiled from /root/reference
 *     (oracle/_ref/ref_octree) and against the IO/test/rebricking.h KATs.
 *   - page table / visibility / paging: PINNED against the reference's own
 *     GLVolumePool.cpp compiled against a null-GL shim (oracle/_ref/ref_pool)
 *     when that builds; otherwise restatement only.
 *   - raycast arithmetic (GLSL): PARITY UNPINNED -- the reference has no golden
 *     images and its GL renderer cannot run in this container (no GL).
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_U8 = 0, ORC_U16 = 1, ORC_F32 = 2 };
enum { ORC_RM_1DTRANS = 0, ORC_RM_2DTRANS = 1, ORC_RM_ISOSURFACE = 2 };
enum { ORC_BI_MISSING = 0, ORC_BI_CHILD_EMPTY = 1, ORC_BI_EMPTY = 2, ORC_BI_FLAG_COUNT = 3 };
/* MissingBrickStrategy, Controller/MasterController.h:61-67 */
enum { ORC_BS_ONLY_NEEDED = 0, ORC_BS_REQUEST_ALL = 1, ORC_BS_SKIP_ONE = 2, ORC_BS_SKIP_TWO = 3 };

#define ORC_MAX_LOD 32

/* ------------------------------------------------------------------ */
/* data side: ExtendedOctree geometry + converter                      */
/* ------------------------------------------------------------------ */
typedef struct orc_octree orc_octree;

orc_octree* orc_octree_new(const uint32_t vol[3], const uint32_t max_brick[3],
                           uint32_t overlap, int dtype);
void        orc_octree_free(orc_octree*);
/* flat: x-fastest raw volume of vol[0]*vol[1]*vol[2] voxels */
int         orc_octree_build(orc_octree*, const void* flat, int clamp_to_edge, int median);
uint32_t    orc_octree_lod_count(const orc_octree*);            /* until 1^3 */
uint32_t    orc_octree_largest_single_brick_lod(const orc_octree*);
void        orc_octree_lod_size(const orc_octree*, uint32_t lod, uint32_t out[3]);
void        orc_octree_brick_count(const orc_octree*, uint32_t lod, uint32_t out[3]);
uint64_t    orc_octree_total_bricks(const orc_octree*);
uint64_t    orc_octree_brick_index(const orc_octree*, uint32_t x, uint32_t y, uint32_t z, uint32_t lod);
void        orc_octree_brick_size(const orc_octree*, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]);
/* copies brick voxels (incl. ghost), x-fastest, tightly packed at the brick's own size */
int         orc_octree_get_brick(const orc_octree*, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst);
/* 4 doubles per brick: minScalar,maxScalar,minGradient,maxGradient; TOC order */
const double* orc_octree_minmax(const orc_octree*);
const void* orc_octree_lod_volume(const orc_octree*, uint32_t lod);

/* ------------------------------------------------------------------ */
/* transfer functions                                                  */
/* ------------------------------------------------------------------ */
/* TransferFunction1D::SetStdFunction(center, invGradient) on n entries -> float rgba[n*4] */
void orc_tf1d_std(float* rgba, uint32_t n, float center, float inv_gradient);
/* TransferFunction1D::GetByteArray (truncating) */
void orc_tf1d_bytes(const float* rgba, uint32_t n, uint8_t* out);
/* TransferFunction1D::ComputeNonZeroLimits: lo = n, hi = 0 when all alpha are zero */
void orc_tf1d_nonzero(const float* rgba, uint32_t n, uint64_t* lo, uint64_t* hi);
/* TransferFunction2D::ComputeNonZeroLimits over an RGBA8 raster (w x h) */
void orc_tf2d_nonzero(const uint8_t* rgba, uint32_t w, uint32_t h, uint64_t out[4]);

/* ------------------------------------------------------------------ */
/* pool sizing, page table, visibility, paging, hash table             */
/* ------------------------------------------------------------------ */
typedef struct orc_pool orc_pool;

/* GPUMemMan::GetVolumePool sizing */
void orc_pool_size(uint64_t max_gpu_mem, uint64_t bit_width, uint64_t comp_count,
                   const uint32_t max_brick[3], uint64_t total_brick_count,
                   uint32_t max_3d_dim, uint32_t out_pool_size[3]);
/* Fit1DIndexTo3DArray */
int  orc_fit_1d_to_3d(uint64_t max_idx, uint32_t max_array, uint32_t out[3]);

orc_pool* orc_pool_new(const uint32_t pool_size[3], const uint32_t vol[3],
                       const uint32_t max_brick[3], uint32_t overlap,
                       uint32_t pool_lod_count, uint32_t max_3d_dim,
                       const double* minmax4 /* TOC order, 4 per brick */);
void      orc_pool_free(orc_pool*);
uint32_t  orc_pool_total_bricks(const orc_pool*);
uint32_t  orc_pool_meta_count(const orc_pool*);
const uint32_t* orc_pool_meta(const orc_pool*);
void      orc_pool_meta_dim(const orc_pool*, uint32_t out[3]);
void      orc_pool_capacity(const orc_pool*, uint32_t out[3]);
void      orc_pool_lod_offsets(const orc_pool*, uint32_t* out /* lod_count */);
void      orc_pool_brick_layout(const orc_pool*, uint32_t lod, uint32_t out[3]);
void      orc_pool_float_layout(const orc_pool*, uint32_t lod, float out[3]);
uint32_t  orc_pool_brick_id(const orc_pool*, uint32_t x, uint32_t y, uint32_t z, uint32_t lod);
void      orc_pool_vector_id(const orc_pool*, uint32_t id, uint32_t out[4]);
/* UploadFirstBrick: returns slot index used (last slot) */
uint32_t  orc_pool_upload_first(orc_pool*);
/* RecomputeVisibility (synchronous); vis = {mode, min, max, gmin, gmax | iso} */
void      orc_pool_recompute_visibility(orc_pool*, int mode, double a, double b,
                                        double c, double d, uint32_t counts[4]);
/* UploadBricks: ids = n x (x,y,z,lod); out_slots[i] = linear pool coordinate used
 * for request i or 0xFFFFFFFF if not paged; returns #paged */
uint32_t  orc_pool_upload_bricks(orc_pool*, const uint32_t* ids, uint32_t n, uint32_t* out_slots);
/* slot table dump: per slot {brickID(int32 as u32), posx,posy,posz} in CURRENT (sorted) order,
 * followed by times -- for parity of the LRU state */
uint32_t  orc_pool_slot_count(const orc_pool*);
void      orc_pool_slots(const orc_pool*, int32_t* brick_ids, uint64_t* times, uint32_t* pos3);

/* ------------------------------------------------------------------ */
/* renderer                                                            */
/* ------------------------------------------------------------------ */
typedef struct {
  /* image */
  uint32_t width, height;
  /* matrices, Tuvok storage (row vectors, v' = v*M, array[r*4+c]) */
  float model_view[16];
  float projection[16];
  /* volume */
  uint32_t vol[3];
  float    scale[3];          /* Dataset::GetScale */
  int      dtype;
  /* pool geometry */
  uint32_t pool_size[3];
  uint32_t capacity[3];
  uint32_t max_total_brick[3];
  uint32_t max_inner_brick[3];
  uint32_t lod_count;          /* pool LoD count */
  uint32_t lod_offset[ORC_MAX_LOD];
  uint32_t meta_dim[3];
  /* modes */
  int   mode;                  /* ORC_RM_* */
  int   lighting;
  float sample_rate_modifier;
  float trans_scale;           /* fTransScale */
  float gradient_scale;        /* fGradientScale */
  float isoval;                /* fIsoval, normalised */
  float ambient[4], diffuse[4], specular[4]; /* rgba as in AbstrRenderer (w = intensity) */
  float light_dir[3];
  float eye[3];                /* m_vEye */
  float iso_color[3];
  float lod_factor;            /* CullingLOD::GetLoDFactor */
  /* TF */
  uint32_t tf_w, tf_h;         /* 1D: h = 1 */
  /* miss reporting */
  uint32_t hash_size, rehash_count;
  int      strategy;           /* ORC_BS_* */
  /* sort-last shard box in normalised volume coords ([0,1]^3 = whole volume).  Rays are NOT clipped to it:
   * they walk the brick chain of the whole volume from their true entry (so sample positions equal the
   * single-GPU ray's) and only samples inside [clip_min, clip_max) are taken; bricks that do not touch the
   * box are stepped through without sampling or paging */
  float clip_min[3], clip_max[3];
  int   nearest;               /* SetInterpolant(NEAREST) */
} orc_render_params;

typedef struct {
  uint64_t samples;       /* ComputeColorFromVolume / GetVolumeHit evaluations */
  uint64_t rays;          /* pixels covered by back faces */
  uint64_t brick_visits;  /* GetBrick calls */
  uint32_t hash_entries;  /* non-zero entries after the pass */
} orc_render_stats;

/* FillRayEntryBuffer: entry[w*h*4] (xyz norm pos, w eye z); coverage via exit[w*h*4]
 * (xyz norm exit, w eye z); covered[w*h] = 1 where back faces produce a fragment */
void orc_ray_setup(const orc_render_params*, float* entry, float* exit_, uint8_t* covered);

/* One GridLeaper raycast pass (GLGridLeaper-blend.glsl / -iso.glsl main()).
 *  pool      : atlas 3D array (pool_size, x fastest) of dtype voxels
 *  meta      : page table
 *  tf        : RGBA8 table (tf_w*tf_h*4)
 *  ray_start / start_color : in  (w*h*4 floats) -- resume state
 *  out0..out3: MRT outputs (w*h*4 floats each):
 *     DVR: out0 accRayColor, out1 rayResumeColor, out2 rayResumePos, out3 unused
 *     ISO: out0 rayHitPos,   out1 rayHitNormal,   out2 rayResumePos, out3 rayResumeNormal
 *  hash      : u32[hash_size] (not cleared here)
 */
void orc_raycast(const orc_render_params*, const void* pool, const uint32_t* meta,
                 const uint8_t* tf, const float* ray_start, const float* start_color,
                 const float* exit_, const uint8_t* covered,
                 float* out0, float* out1, float* out2, float* out3,
                 uint32_t* hash, orc_render_stats* stats, int n_threads);

/* Compose-FS.glsl over the iso hit buffers -> rgba float image (cleared to 0) */
void orc_iso_compose(const orc_render_params*, const float* hit_pos, const float* hit_normal, float* rgba);

/* GLHashTable::GetData decode: returns n, fills out[n*4] */
uint32_t orc_hash_decode(const uint32_t* hash, uint32_t hash_size, const uint32_t finest_layout[3], uint32_t* out);

/* GL float -> unorm8 read-back (GLFrameCapture.cpp:72-85) */
void orc_rgba8(const float* rgba, uint64_t n_pixels, uint8_t* out);

/* ------------------------------------------------------------------ */
/* classic per-brick raycaster (GLRaycaster) -- orc_classic.cpp          */
/* ------------------------------------------------------------------ */
typedef struct {
  float center[3], ext[3];        /* world box = center +- ext/2 (AbstrRenderer.cpp:1003-1028) */
  float tex_min[3], tex_max[3];   /* UVFDataset::GetTextCoords */
  uint32_t n_vox[3];              /* brick size incl. ghost */
  uint32_t coord[3];
  uint32_t index;                 /* z*bx*by + y*bx + x inside the LoD (BrickKey index) */
  float distance;                 /* brick_distance */
  int32_t empty;                  /* bIsEmpty: in the frustum but ContainsData() == false */
} orc_classic_brick;
/* AbstrRenderer::ComputeMinLODForCurrentView, clamped to [0, lod_count-1] */
uint32_t orc_classic_lod(const orc_render_params*, uint32_t lod_count);
/* AbstrRenderer::BuildSubFrameBrickList for one LoD (frustum culled, min/max tested, depth sorted) */
uint32_t orc_classic_brick_list(const orc_render_params*, uint32_t lod, uint32_t overlap, const double* minmax_lod,
                                const double vis[4], orc_classic_brick* out, uint32_t cap);
/* GLRenderer::Render3DView brick loop + GLRaycaster::Render3DInLoop + GL blending */
void orc_classic_render(const orc_render_params*, uint32_t lod, const orc_classic_brick* list, uint32_t n_bricks,
                        const void* const* brick_data, const uint8_t* tf, float* out, orc_render_stats* stats,
                        int n_threads);

/* over-operator of the sort-last compositor: front + (1-front.a)*back, aware of early ray termination
 * (a terminated front hides the back; a back image that would push alpha past 0.995 is cut there) */
void orc_composite_over(const float* front, const float* back, uint64_t n_pixels, float* out);

#ifdef __cplusplus
}
#endif
#endif
