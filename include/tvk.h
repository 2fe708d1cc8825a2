/*
 * tvk.h -- C ABI of libtvkcuda.so, the B200 (sm_100a) brick-pool volume raycaster
 * that drops in for Tuvok's GLGridLeaper / GLRaycaster hot path.
 *
 * The reference (SCIInstitute/Tuvok) has no plugin/FFI mechanism: renderers are C++
 * subclasses of tuvok::AbstrRenderer (Renderer/AbstrRenderer.h:112-881) created by
 * MasterController::RequestNewVolumeRenderer (Controller/MasterController.cpp:144-212).
 * A maintainer adds one enum value + one `case` there that creates the thin
 * `CUDAGridLeaper : AbstrRenderer` shim shown in INTEGRATION.md; that shim (and the
 * Python/ctypes test harness in tuvok_b200/) talks to this ABI only.
 *
 * Conventions: plain C, no exceptions cross the boundary, every call returns a
 * tvk_status (0 = ok) and records a message retrievable with tvk_last_error().
 * Input buffers are borrowed for the duration of the call, output buffers are
 * caller-allocated.  One tvk_ctx per renderer; a ctx is NOT thread-safe (the reference
 * renderer is bound to the one thread owning its GL context).  Matrices use
 * Tuvok's storage: row-major float[16], ROW vectors, v' = v*M (Basics/Vectors.h:434-439,
 * 855-864).  Every entry point cites the reference interface it replaces.
 */
#ifndef TVK_H
#define TVK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVK_ABI_VERSION 1
#define TVK_MAX_LOD 16

typedef enum {
  TVK_OK = 0,
  TVK_ERR_INVALID = 1,      /* bad argument / call order (reference: T_ERROR + return false) */
  TVK_ERR_CUDA = 2,         /* a CUDA runtime call failed */
  TVK_ERR_NO_DEVICE = 3,    /* no sm_100 device: there is NO CPU fallback */
  TVK_ERR_OOM = 4,
  TVK_ERR_SOURCE = 5        /* the brick source callback failed (Dataset::GetBrick returned false) */
} tvk_status;

/* ExtendedOctree::COMPONENT_TYPE subset on the hot path (ExtendedOctree.h:137-148) */
typedef enum { TVK_U8 = 0, TVK_U16 = 1, TVK_F32 = 2,
               /* colour volume: 4 x 8 bit per voxel, bricks are RGBA byte quadruples (UVF component count 4).  The renderer
                * then runs GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}-color.glsl / Compose-Color-FS.glsl (the transfer function
                * maps alpha only; AbstrRenderer::ColorData).  min / max = the ALPHA component's (uvfDataset.cpp:1144, :1188).
                * Registered datasets (tvk_set_volume) and four-component 8-bit ExtendedOctree / UVF files (without a table the alpha
                * statistics are computed on the device like the scalar ones) and tvk_build_volume (mean pyramid, unsharded store) on the GridLeaper
                * path: the re-bricking loader, the classic
                * per-brick path, MIP, sort-last and the depth pipeline refuse it. */
               TVK_RGBA8 = 3 } tvk_dtype;
/* AbstrRenderer::ERenderMode (Renderer/AbstrRenderer.h:142-147) */
typedef enum { TVK_RM_1DTRANS = 0, TVK_RM_2DTRANS = 1, TVK_RM_ISOSURFACE = 2 } tvk_render_mode;
/* RendererState::BrickStrategy (Controller/MasterController.h:61-67) */
typedef enum { TVK_BS_ONLY_NEEDED = 0, TVK_BS_REQUEST_ALL = 1, TVK_BS_SKIP_ONE_LEVEL = 2,
               TVK_BS_SKIP_TWO_LEVELS = 3 } tvk_brick_strategy;
/* page-table flags, BrickIDFlags (Renderer/GL/GLVolumePool.cpp:25-30); >= 3: pool slot + 3 */
enum { TVK_BI_MISSING = 0, TVK_BI_CHILD_EMPTY = 1, TVK_BI_EMPTY = 2, TVK_BI_FLAG_COUNT = 3 };

typedef struct tvk_ctx tvk_ctx;

/* RendererState + SystemInfo knobs (MasterController.h:61-74, SystemInfo.h:48) */
typedef struct {
  int32_t  device;            /* CUDA device ordinal */
  uint64_t max_gpu_mem;       /* pool budget in bytes, 0 = 8 GiB default of SystemInfo; GPUMemMan.cpp:766-844 */
  uint32_t max_pool_dim;      /* stands in for GL_MAX_3D_TEXTURE_SIZE; 0 = 16384 */
  uint32_t hash_table_size;   /* RState.HashTableSize, 0 = 509 */
  uint32_t rehash_count;      /* RState.RehashCount, 0 = 10 */
  int32_t  brick_strategy;    /* tvk_brick_strategy; MasterController.cpp:89 default SKIP_TWO_LEVELS */
} tvk_device_cfg;

/* What GLGridLeaper reads from LinearIndexDataset (SURVEY 8b "Data interfaces consumed") */
typedef struct {
  uint32_t domain_size[3];     /* GetDomainSize(0) */
  float    scale[3];           /* GetScale() */
  uint32_t max_brick_size[3];  /* GetMaxUsedBrickSizes() -- incl. ghost */
  uint32_t overlap;            /* GetBrickOverlapSize() (same on all axes) */
  int32_t  dtype;              /* tvk_dtype: GetBitWidth/GetIsFloat */
  double   range_max;          /* GetRange().second (MaxValue, AbstrRenderer.cpp:860-866) */
  float    max_gradient_magnitude; /* Dataset::MaxGradientMagnitude */
  uint64_t brick_count;        /* GetTotalBrickCount(): entries in minmax */
  const double* minmax;        /* MaxMinForKey for every brick in TOC order (LOD-major, z, y, x):
                                  {minScalar, maxScalar, minGradient, maxGradient} */
} tvk_volume_desc;

/* Dataset::GetBrick(key, vector<T>&) (IO/Dataset.h:93-100).  dst is LIBRARY-OWNED PINNED
 * memory of `cap` bytes; write the brick x-fastest at its own size (incl. ghost).
 * Return 0 on success. */
typedef int (*tvk_brick_cb)(void* user, uint32_t x, uint32_t y, uint32_t z, uint32_t lod,
                            void* dst, size_t cap);
/* AbstrDebugOut channel: 0 message, 1 warning, 2 error, 3 other (Controller/Controller.h:68-84) */
typedef void (*tvk_log_cb)(void* user, int channel, const char* source, const char* msg);

/* The per-frame state the GL renderer reads from AbstrRenderer */
typedef struct {
  uint32_t width, height;        /* AbstrRenderer::Resize */
  float model_view[16];          /* rotation*translation*view, GLRenderer.cpp:627 */
  float projection[16];          /* GLRenderer::ComputeViewAndProjection, GLRenderer.cpp:892-919 */
  float lod_factor;              /* CullingLOD::GetLoDFactor, CullingLOD.cpp:57-67 */
  int32_t mode;                  /* tvk_render_mode */
  int32_t lighting;              /* m_bUseLighting */
  float sample_rate_modifier;    /* m_fSampleRateModifier */
  double isovalue;               /* m_fIsovalue in data units (GetIsoValue) */
  float ambient[4], diffuse[4], specular[4]; /* m_cAmbient/m_cDiffuse/m_cSpecular: rgb, w = intensity */
  float light_dir[3];            /* m_vLightDir */
  float eye[3];                  /* m_vEye */
  float iso_color[3];            /* m_vIsoColor */
  int32_t nearest;               /* SetInterpolant(NearestNeighbor) */
  float clip_min[3], clip_max[3];/* sort-last shard box in normalised volume space; {0,0,0},{1,1,1} = all */
} tvk_render_params;

/* mirrors PERF_* of Basics/PerfCounter.h:7-44 for the phases of GLGridLeaper::Render3DRegion */
typedef struct {
  int32_t  converged;            /* m_bConverged = hash.empty(), GLGridLeaper.cpp:1097-1099 */
  uint32_t missing_reported;     /* decoded hash entries */
  uint32_t bricks_paged;         /* UploadBricks count */
  uint64_t samples;              /* ComputeColorFromVolume/GetVolumeHit evaluations (if counting enabled) */
  uint64_t rays;                 /* pixels covered by the volume */
  uint64_t brick_visits;         /* GetBrick calls */
  uint64_t bricks_touched;       /* distinct non-empty bricks sampled this subframe (if counting enabled) */
  uint64_t alive_lane_iters;     /* diagnostics: sum over lanes of loop turns with a live ray */
  uint64_t warp_iters;           /* diagnostics: loop turns summed over warps (x32 = lane slots) */
  uint64_t max_lane_iters;       /* diagnostics: loop turns of the longest ray = the kernel's critical path */
  float ms_raycast;              /* PERF_RAYCAST (CUDA events) */
  float ms_read_htable;          /* PERF_READ_HTABLE + PERF_CONDENSE_HTABLE */
  float ms_upload_bricks;        /* PERF_UPLOAD_BRICKS */
  float ms_total;                /* PERF_RENDER */
} tvk_frame_stats;

typedef struct {
  uint32_t lod_count;            /* all LODs down to 1^3 */
  uint32_t pool_lod_count;       /* GetLargestSingleBrickLOD()+1, GLVolumePool.cpp:132 */
  uint64_t total_bricks;         /* bricks of the pool LoDs = page-table entries */
  uint32_t lod_size[TVK_MAX_LOD][3];
  uint32_t brick_layout[TVK_MAX_LOD][3];
  uint32_t lod_offset[TVK_MAX_LOD];
  uint32_t pool_size[3];         /* atlas voxels = capacity * max brick */
  uint32_t pool_capacity[3];     /* slots per axis */
  uint32_t meta_dim[3];          /* Fit1DIndexTo3DArray shape of the page table, GLVolumePool.cpp:816-848 */
  uint64_t meta_count;           /* meta_dim volume (>= total_bricks) */
} tvk_info;

/* ---- lifecycle ------------------------------------------------------------------- */
uint32_t    tvk_abi_version(void);
/* MasterController::RequestNewVolumeRenderer + GLGridLeaper ctor (MasterController.cpp:144-212) */
int         tvk_create(const tvk_device_cfg* cfg, tvk_ctx** out);
/* GLGridLeaper::CleanupShaders/Cleanup (GLGridLeaper.cpp:134-160) */
void        tvk_destroy(tvk_ctx* ctx);
const char* tvk_last_error(const tvk_ctx* ctx);   /* ctx may be NULL: error of the last failed tvk_create */
int         tvk_set_log_callback(tvk_ctx* ctx, tvk_log_cb cb, void* user);
/* all kernels are launched on this cudaStream_t (default: the ctx's own stream) */
int         tvk_set_stream(tvk_ctx* ctx, void* cuda_stream);
int         tvk_synchronize(tvk_ctx* ctx);
/* count samples/rays/brick visits in tvk_frame_stats (costs atomics; off by default) */
int         tvk_enable_counters(tvk_ctx* ctx, int enable);

/* ---- dataset --------------------------------------------------------------------- */
/* GLGridLeaper::RegisterDataset (GLGridLeaper.cpp:105-132): host-described dataset whose bricks
 * come from `cb` (Dataset::GetBrick).  Copies the min/max table (GLVolumePool.cpp:225-235). */
int tvk_set_volume(tvk_ctx* ctx, const tvk_volume_desc* desc, tvk_brick_cb cb, void* user);
/* Device-side data producer replacing ExtendedOctreeConverter::Convert
 * (ExtendedOctreeConverter.cpp:128-280): bricks a raw x-fastest volume (host or device pointer),
 * builds the 2x2x2-mean LOD pyramid, ghost cells and per-brick min/max on the GPU and keeps the
 * brick store resident; it then acts as the brick source. range_max <= 0: 2^bits-1 (1 for f32). */
int tvk_build_volume(tvk_ctx* ctx, const void* raw, int raw_on_device, const uint32_t size[3],
                     int dtype, const float scale[3], const uint32_t max_brick_size[3],
                     uint32_t overlap, int clamp_to_edge, double range_max,
                     float max_gradient_magnitude);
/* The filter tvk_build_volume halves a level with: 0 = mean (default), 1 = median -- ExtendedOctreeConverter::Convert's
 * bComputeMedian (ExtendedOctreeConverter.cpp:128, .inc:1-248; VolumeTools::Filter<T, F, true>, VolumeTools.h:168-262: the
 * first of two, the median of the first three of four, a median of the first seven of eight).  Applies to the next
 * tvk_build_volume. */
int tvk_set_pyramid_filter(tvk_ctx* ctx, int median);
/* seeded integer-arithmetic synthetic volumes (bit-identical to tuvok_b200.synth on the CPU);
 * kind 0 = V_sph (shells), 1 = V_noise (value noise x falloff), 2 = V_ramp (x + 8y + 64z).
 * Writes size[0]*size[1]*size[2] voxels to the DEVICE pointer dst. */
int tvk_synth_volume(tvk_ctx* ctx, void* dst_device, int kind, const uint32_t size[3], int dtype,
                     uint32_t seed);
/* ExtendedOctree file source: the payload of a UVF TOC block -- or the file
 * ExtendedOctreeConverter::Convert writes -- becomes the brick source of the streaming path.  Replaces
 * ExtendedOctree::Open / GetBrickData (IO/UVF/ExtendedOctree/ExtendedOctree.cpp:87-165,313-360) and
 * UVFDataset::GetBrick (IO/uvfDataset.cpp:1690-1712): header + table of contents are parsed once, bricks are
 * read with parallel pread() straight into the library's pinned staging memory (zlib / lz4 / LZMA / bzip2 bricks
 * are decoded there; bzip2 needs the system's libbz2 runtime) and copied to the pool on the side stream.  `offset` = byte offset of the
 * octree header in the file, `uvf_file_version` as UVF::ms_ulReaderVersion (>= 5: versioned octree header).
 * scale NULL: the octree's volume aspect.  minmax = MaxMinDataBlock contents (TOC order) or NULL: the table is then
 * computed on the device in one streaming pass over all bricks.  info may be NULL. */
typedef struct {
  uint32_t domain_size[3];
  double   aspect[3];
  uint32_t max_brick_size[3];
  uint32_t overlap;
  int32_t  dtype;
  uint32_t version;              /* octree version (0: UVF <= 4) */
  uint32_t lod_count;
  uint64_t brick_count;          /* all LoDs */
  uint64_t payload_bytes;        /* stored (compressed) bytes of all bricks */
  uint64_t bricks_by_codec[6];   /* ExtendedOctree COMPRESSION_TYPE histogram: none, zlib, lzma, lz4, bzlib, other */
} tvk_octree_file_info;
int tvk_open_octree_file(tvk_ctx* ctx, const char* path, uint64_t offset, uint64_t uvf_file_version,
                         const float scale[3], const double* minmax, uint64_t n_minmax, double range_max,
                         float max_gradient_magnitude, tvk_octree_file_info* info);
/* DynamicBrickingDS on the device (IO/DynamicBrickingDS.cpp; IOManager::LoadRebrickedDataset, IO/IOManager.cpp:1281-1320):
 * a file converted with LARGE bricks is re-cut into bricks of target_brick_size (incl. ghost; clamped to the source's) while
 * it is loaded -- every source brick is streamed once (pinned, double buffered) and cut on the device into the brick store,
 * which then serves the pool like a tvk_build_volume store; min / max per target brick are computed in the same pass
 * (MinMaxMode MM_PRECOMPUTE).  The reference's rules hold: the target's inner size must divide the source's on every axis,
 * the ghost width is the source's, a target brick is a sub-box of one source brick (ghost at a source brick's border =
 * that brick's own ghost, DynamicBrickingDS::GetBrick), the levels are the file's own, target_brick_size is clamped to the
 * source's (IOManager.cpp:1301-1305).  Scalar min / max equal DynamicBrickingDS::MaxMinForKey bit for bit.  One stated
 * deviation: minmax_brick (BMinMax.cpp:13) stores the EMPTY gradient interval (DBL_MAX, -FLT_MAX), which hides every brick
 * of a rebricked dataset from a 2D transfer function; the gradient interval stored here is the unbounded one, as for every
 * other source of this library without gradient statistics. */
int tvk_open_octree_file_rebricked(tvk_ctx* ctx, const char* path, uint64_t offset, uint64_t uvf_file_version,
                                   const float scale[3], const uint32_t target_brick_size[3], double range_max,
                                   float max_gradient_magnitude, tvk_octree_file_info* info);
/* The same for a whole .uvf file: walks the container (magic, global header, data-block list; UVF.cpp:140-290,
 * GlobalHeader.cpp:39-47, DataBlock.cpp:60-72), takes the `timestep`-th TOC block as the brick source and the
 * `timestep`-th MaxMin block (if any) as the min/max table -- what UVFDataset::Open does for TOC-based files
 * (IO/uvfDataset.cpp:640-700).  Legacy raster-data-block UVFs are refused (use tvk_set_volume with GetBrick). */
int tvk_open_uvf(tvk_ctx* ctx, const char* path, uint64_t timestep, const float scale[3], double range_max,
                 float max_gradient_magnitude, tvk_octree_file_info* info);
/* host-only: container walk result; maxmin (4 doubles per brick, TOC order) is filled up to maxmin_cap bricks */
int tvk_uvf_probe(const char* path, uint64_t timestep, uint64_t* toc_payload_offset, uint64_t* file_version,
                  uint64_t* n_blocks, uint64_t* n_timesteps, double* maxmin, uint64_t maxmin_cap, uint64_t* n_maxmin);
/* host-only: what UVFDataset derives from the other blocks of the container -- the value range (ComputeRange,
 * IO/uvfDataset.cpp:1120-1155: min / max over the LoD-0 bricks of the MaxMin block; range[1] < range[0] = unknown), the
 * 1D histogram's length and filled size (index of the last non-zero bin + 1: the 1D transfer function's size,
 * GLRenderer.cpp:188-196) and the 2D histogram's maximum gradient magnitude (fGradientScale = 1 / it) and size.
 * tvk_open_uvf uses the range maximum and the gradient magnitude when it is called with range_max <= 0 /
 * max_gradient_magnitude <= 0. */
int tvk_uvf_probe_stats(const char* path, uint64_t timestep, double range[2], uint64_t* hist1d_size,
                        uint64_t* hist1d_filled, float* max_gradient_magnitude, uint64_t hist2d_size[2]);
/* host-only helpers (no device, no ctx; errors via tvk_last_error(NULL)): parse header + table of contents,
 * read one brick (x fastest, own size incl. ghost, decoded) into host memory */
int tvk_octree_file_probe(const char* path, uint64_t offset, uint64_t uvf_file_version, tvk_octree_file_info* info);
int tvk_octree_file_read_brick(const char* path, uint64_t offset, uint64_t uvf_file_version, uint32_t x, uint32_t y,
                               uint32_t z, uint32_t lod, void* dst, size_t cap, uint32_t out_size[3]);
/* ---- procedural multi-resolution dataset (BASELINE configs[4]: 8192^3 uint8 = 512 GiB, a volume that exists nowhere at
 * once).  Level l of the hierarchy is the seeded analytic field of tvk_synth_volume sampled on that level's grid (domain
 * size ceil-halved l times -- the level sizes of ExtendedOctree, ExtendedOctree.cpp:167-199), bricked like a converted
 * file (max_brick_size incl. `overlap` ghost voxels per side, 0 outside the level's grid).  It stands where a
 * UVFDataset would: Dataset::GetBrick (IO/uvfDataset.cpp:1690-1712) is answered by host threads that generate the
 * requested brick into the library's pinned staging memory, from where it is paged in like any other brick
 * (GLVolumePool::UploadBricks, GLVolumePool.cpp:1720-1789: pinned cudaMemcpyAsync on the copy stream, double buffered);
 * only bricks the LOD-driven traversal asks for are ever produced.  host_cache_bytes > 0 keeps generated bricks in a
 * host-side LRU cache in front of the generator (the role the OS page cache plays for a .uvf file; page-locked when the
 * system allows it, so that the copy engine reads a cached brick where it lies), so a brick the device pool evicted
 * streams back at PCIe speed.  minmax = the MaxMinDataBlock-equivalent table for all pool-LoD bricks in
 * TOC order (4 doubles per brick; from tvk_procedural_minmax, possibly computed in slices by several ranks) or NULL: it
 * is then computed on this device.  threads = host generator threads (0: all cores). */
int tvk_set_procedural_volume(tvk_ctx* ctx, int kind, const uint32_t size[3], int dtype, uint32_t seed, const float scale[3],
                              const uint32_t max_brick_size[3], uint32_t overlap, double range_max,
                              float max_gradient_magnitude, const double* minmax, uint64_t n_minmax,
                              uint64_t host_cache_bytes, uint32_t threads);
/* host-only: number of bricks / LoDs of the pool LoDs (down to the first single-brick level) of such a dataset */
int tvk_procedural_brick_count(const uint32_t size[3], const uint32_t max_brick_size[3], uint32_t overlap, uint64_t* n_bricks,
                               uint32_t* n_lods);
/* host-only parity tap: one brick of such a dataset exactly as the generator threads produce it (x fastest, tightly
 * packed at its own size incl. ghost; out_size may be NULL) */
int tvk_procedural_brick(int kind, const uint32_t size[3], int dtype, uint32_t seed, const uint32_t max_brick_size[3],
                         uint32_t overlap, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap,
                         uint32_t out_size[3]);
/* per-brick min/max (incl. ghost; {min, max, -DBL_MAX, DBL_MAX} like tvk_build_volume) of the bricks [first, first+count)
 * in TOC order, evaluated on the device without materialising the bricks; needs no dataset to be set */
int tvk_procedural_minmax(tvk_ctx* ctx, int kind, const uint32_t size[3], int dtype, uint32_t seed,
                          const uint32_t max_brick_size[3], uint32_t overlap, uint64_t first, uint64_t count, double* dst);
/* totals of the streaming path since tvk_create (callback / file / procedural sources) */
typedef struct {
  uint64_t bricks_uploaded;        /* bricks that went host -> pool */
  uint64_t h2d_bytes;              /* bytes of those copies (slot-sized) */
  double   upload_ms;              /* host wall time inside the upload path (source + copy, overlapped) */
  double   h2d_ms;                 /* device time of the pinned H2D brick copies alone (events on the copy stream) */
  uint64_t bricks_generated;       /* procedural source: bricks produced by the generator */
  uint64_t host_cache_hits;        /* ... served from the host cache instead */
  uint64_t host_cache_evictions;
  double   source_thread_ms;       /* ... thread-milliseconds spent producing / copying bricks on the host */
  uint32_t source_threads;
  uint32_t host_cache_pinned;      /* the cache is page-locked: cached bricks are copied by DMA straight out of it */
} tvk_stream_stats;
int tvk_get_stream_stats(tvk_ctx* ctx, tvk_stream_stats* out);
/* ---- value quantiser of the import path ---------------------------------------------------------------
 * Quantize<T, U> (IO/Quantize.h:427-577) and AbstrConverter::Process8Bits (IO/AbstrConverter.cpp:73-157) as RAWConverter's
 * quantize() calls them (IO/RAWConverter.cpp:205-300), for n values that are already in device memory: a range pass, then
 * every value -> min(max_out, U((v - min) * factor)) (factor = max_out / (max - min), capped at 1 for integer input so
 * that integers are only ever compressed, never stretched) plus the 1D histogram (256 bins for 8-bit output, 4096 for
 * 16-bit) that becomes the Histogram1DDataBlock.  Signed bytes are biased by 128, unsigned bytes are only counted.
 * Unsigned 16-bit data whose maximum is below 4096 needs no processing: info->changed = 0, dst is not written, and
 * info->hist_set = 0 because the reference returns before it sets the histogram (hist then holds the direct value counts
 * the bin count is taken from).  A constant input maps to 0 (the reference divides by zero).  src_device must be 16-byte
 * aligned; dst_device: n values of uint8 / uint16; hist: HOST array of 256 / 4096 counters. */
typedef enum { TVK_ST_I8 = 0, TVK_ST_U8 = 1, TVK_ST_I16 = 2, TVK_ST_U16 = 3, TVK_ST_I32 = 4, TVK_ST_U32 = 5, TVK_ST_F32 = 6,
               TVK_ST_F64 = 7 } tvk_scalar_type;
typedef struct {
  double   min, max;      /* value range of the input */
  double   factor;        /* fQuantFact */
  uint64_t bin_count;     /* *iBinCount: non-zero bins (data used as is) or bins_needed */
  int32_t  changed;       /* Quantize's return value: 1 = dst holds the quantised data, 0 = the input is used as is */
  int32_t  hist_set;      /* 0: the reference leaves the Histogram1DDataBlock untouched (early return) */
  float    ms_range, ms_map;   /* device time of the two passes (CUDA events) */
} tvk_quantize_info;
int tvk_quantize(tvk_ctx* ctx, const void* src_device, int scalar_type, uint64_t n, int out_bits, void* dst_device,
                 uint64_t* hist, tvk_quantize_info* info);
int tvk_get_info(const tvk_ctx* ctx, tvk_info* out);
/* parity taps: MaxMinForKey table and one brick (x-fastest, own size incl. ghost) */
int tvk_get_minmax(tvk_ctx* ctx, double* dst, uint64_t n_bricks);
int tvk_get_brick_size(const tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, uint32_t out[3]);
int tvk_read_brick(tvk_ctx* ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t lod, void* dst, size_t cap);

/* ---- transfer functions ---------------------------------------------------------- */
/* AbstrRenderer::Set1DTrans/Changed1DTrans; rgba = TransferFunction1D::GetByteArray
 * (TransferFunction1D.cpp:311-330), nz = GetNonZeroLimits (:362-371) */
int tvk_set_tf1d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t n, uint64_t nz_lo, uint64_t nz_hi);
/* Changed2DTrans; rgba = TransferFunction2D::GetByteArray (w x h), nz = GetNonZeroLimits
 * (xmin, xmax, ymin, ymax) (TransferFunction2D.cpp:378-397) */
int tvk_set_tf2d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h, const uint64_t nz[4]);

/* ---- pool / page table ----------------------------------------------------------- */
/* GLGridLeaper::CreateVolumePool (GLGridLeaper.cpp:83-103): size the pool (GPUMemMan::GetVolumePool,
 * GPUMemMan.cpp:766-844, unless pool_size != NULL), build slot + page tables, UploadFirstBrick,
 * RecomputeBrickVisibility. */
int tvk_create_pool(tvk_ctx* ctx, const uint32_t* pool_size /* [3] or NULL */);
/* GLGridLeaper::RecomputeBrickVisibility (GLGridLeaper.cpp:647-687) ->
 * GLVolumePool::RecomputeVisibility (GLVolumePool.cpp:1580-1718), always synchronous.
 * counts = (total, empty, childEmpty, emptyLeaf).  force=0 honours VisibilityState::NeedsUpdate. */
int tvk_recompute_visibility(tvk_ctx* ctx, int force, uint32_t counts[4]);
/* GLVolumePool::UploadBricks (GLVolumePool.cpp:1720-1789) for an explicit request list
 * ids[n][4] = (x,y,z,lod); out_slots[i] = linear pool coordinate or 0xFFFFFFFF. */
int tvk_upload_bricks(tvk_ctx* ctx, const uint32_t* ids, uint32_t n, uint32_t* out_slots, uint32_t* n_paged);
/* parity taps */
int tvk_get_page_table(tvk_ctx* ctx, uint32_t* dst, uint64_t n);
int tvk_get_slots(tvk_ctx* ctx, int32_t* brick_ids, uint64_t* times, uint32_t* pos3, uint32_t n_slots);
int tvk_read_pool_slot(tvk_ctx* ctx, uint32_t slot, void* dst, size_t cap);
/* parity tap: page-table indices of the bricks the last counted subframe (tvk_enable_counters) took samples from --
 * what tvk_frame_stats.bricks_touched counts.  ids may be NULL (size query); *n = number of touched bricks.  Together
 * with tvk_get_page_table and tvk_read_pool_slot this is what a full-size frame is re-traced from on the CPU. */
int tvk_get_touched_bricks(tvk_ctx* ctx, uint32_t* ids, uint64_t cap, uint64_t* n);
/* GLHashTable::GetData of the last subframe (GLHashTable.cpp:90-105): ids[n][4] */
int tvk_get_missing_list(tvk_ctx* ctx, uint32_t* ids, uint32_t cap, uint32_t* n);

/* ---- rendering ------------------------------------------------------------------- */
/* helper: GLRenderer::ComputeViewAndProjection + modelView = rotation*translation*view
 * (GLRenderer.cpp:627,892-919) and CullingLOD::SetScreenParams (CullingLOD.cpp:57-67);
 * fills model_view, projection, lod_factor, width, height, eye of `p`. */
int tvk_compute_view(tvk_render_params* p, uint32_t width, uint32_t height,
                     const float rotation[16], const float translation[16],
                     const float eye[3], const float at[3], const float up[3],
                     float fov_deg, float z_near, float z_far, float screen_space_error);
/* AbstrRenderer defaults (AbstrRenderer.cpp:64-171) */
int tvk_default_params(tvk_render_params* p, uint32_t width, uint32_t height);
/* any view/mode change: marks the region blank (new ray-entry buffer, GLGridLeaper.cpp:925-945) */
int tvk_set_params(tvk_ctx* ctx, const tvk_render_params* p);
/* AbstrRenderer::SetClipPlane / EnableClipPlane / DisableClipPlane (Renderer/AbstrRenderer.h:215-219) as GLGridLeaper
 * implements them (GLGridLeaper.cpp:506-532,1337-1356): the bounding box whose front / back faces give every ray its
 * entry and exit is cut by the plane (Clipper::BoxPlane, Basics/Clipper.cpp:137-167); the part with
 * dot(normal, p) + d <= 0 is kept.  plane_model = (normal.xyz normalised, d) in the box's MODEL space (centre 0, extent =
 * GetVolumeAABB), i.e. what FillBBoxVBO passes to BoxPlane: m_ClipPlane.Plane() * inverse(rotation * translation).
 * Here the clipped polytope is not tessellated: the ray interval is cut analytically per pixel.  Marks the region
 * blank.  GridLeaper path (tvk_render / tvk_paint / tvk_sortlast_frame); the classic path ignores it. */
int tvk_set_clip_plane(tvk_ctx* ctx, int enabled, const float plane_model[4]);
/* host-only helper: PLANE<float>::operator*(inverse(rotation * translation)) + the renormalisation of FillBBoxVBO
 * (Basics/Vectors.h:1459-1487, GLGridLeaper.cpp:518-524) for a world-space plane (ExtendedPlane::Plane()). */
int tvk_clip_plane_to_model(const float plane_world[4], const float rotation[16], const float translation[16],
                            float plane_model[4]);
/* GLRenderer::Pick (GLRenderer.cpp:2856-2872; AbstrRenderer.h:245): the isosurface hit position under a window
 * position (mouse coordinates: y counted from the TOP row, as the reference reads m_vWinSize.y - mousePos.y).
 * TVK_ERR_INVALID outside isosurface mode ("Can only determine pick locations in isosurface rendering mode.") or when
 * the ray hit nothing ("No intersection."); out = rayHitPos.xyz of the last frame (eye space). */
int tvk_pick(tvk_ctx* ctx, uint32_t mouse_x, uint32_t mouse_y, float out[3]);
/* one subframe of GLGridLeaper::Render3DRegion (GLGridLeaper.cpp:914-1154): clear hash table,
 * raycast, read + decode hash table, page missing bricks in.  stats may be NULL. */
int tvk_render(tvk_ctx* ctx, tvk_frame_stats* stats);
/* `while (CheckForRedraw()) Paint()` (GLGridLeaper.cpp:872-890); returns after convergence or
 * max_subframes.  stats (may be NULL) accumulates over the subframes. */
int tvk_paint(tvk_ctx* ctx, uint32_t max_subframes, tvk_frame_stats* stats);
/* raycast pass only, no miss read-back / paging (for device-timed benchmarking of the kernel) */
int tvk_raycast_only(tvk_ctx* ctx);
/* GLFrameCapture read-back of GetLastFBO() (GLFrameCapture.cpp:72-85): bottom row first.
 * pitch in bytes, 0 = tight.  dst is host memory. */
int tvk_read_rgba8(tvk_ctx* ctx, uint8_t* dst, size_t pitch);
int tvk_read_rgba32f(tvk_ctx* ctx, float* dst, size_t pitch);
/* Asynchronous variant (the analogue of a GL pixel-buffer-object read-back): the float -> unorm8 conversion is
 * queued behind the frame on the render stream and the copy runs on the library's copy stream, so it overlaps
 * the next frame's traversal.  dst must be page-locked (tvk_host_alloc); at most two reads may be in flight.
 * tvk_read_wait returns when at most `pending_allowed` (0 or 1) queued reads have not yet landed in host memory
 * (1 = wait for the previous frame while the newest one is still being copied). */
int tvk_read_rgba8_async(tvk_ctx* ctx, uint8_t* dst_pinned, size_t pitch);
int tvk_read_wait(tvk_ctx* ctx, int pending_allowed);
/* page-locked host memory for brick sources and read-backs (cudaMallocHost / cudaFreeHost) */
int tvk_host_alloc(tvk_ctx* ctx, size_t bytes, void** out);
int tvk_host_free(tvk_ctx* ctx, void* p);
/* device pointer of the RGBA32F result (width*height float4, premultiplied) for zero-copy
 * consumers such as the sort-last compositor */
int tvk_get_device_image(tvk_ctx* ctx, void** dptr);
/* iso mode parity taps: rayHitPos / rayHitNormal MRTs */
int tvk_read_iso_buffers(tvk_ctx* ctx, float* hit_pos, float* hit_normal);

/* ---- classic per-brick raycaster (GLRaycaster) ---------------------------------------- */
/* one entry of AbstrRenderer::m_vCurrentBrickList (Renderer/AbstrRenderer.h:69-108) */
typedef struct {
  uint32_t index;        /* BrickKey index inside the LoD: z*bx*by + y*bx + x */
  uint32_t x, y, z;
  float    distance;     /* brick_distance (AbstrRenderer.cpp:808-841) */
  int32_t  empty;        /* bIsEmpty: inside the frustum but ContainsData() is false */
} tvk_classic_brick;
/* One converged frame of the classic path: AbstrRenderer::PlanFrame at ComputeMinLODForCurrentView
 * (AbstrRenderer.cpp:789-803,1125-1212), BuildSubFrameBrickList (:999-1100: frustum culling, legacy
 * ContainsData, depth sort), then GLRenderer::Render3DView's brick loop with GLRaycaster::Render3DPreLoop /
 * Render3DInLoop per brick (GLRenderer.cpp:2663-2748, GLRaycaster.cpp:348-478) and GL under-blending.
 * 1D / 2D transfer function modes with and without lighting, and the isosurface mode (RM_ISOSURFACE branch of
 * Render3DInLoop, GLRaycaster.cpp:383-446: GLRaycaster-ISO-FS.glsl + RefineIsosurface.glsl per brick, nearest hit
 * kept by the depth test, then GLRenderer::ComposeSurfaceImage; tvk_read_iso_buffers returns the two hit targets).
 * The result is read with tvk_read_rgba8/32f. */
int tvk_render_classic(tvk_ctx* ctx, tvk_frame_stats* stats);
/* One HQ MIP frame of a 2D window (RenderRegion2D with GetUseMIP()): AbstrRenderer::PlanHQMIPFrame
 * (AbstrRenderer.cpp:1214-1245: no frustum culling; LoD 0, or with use_mip_lod = m_bMIPLOD the coarsest LoD whose
 * smallest extent still covers the larger window side, stepped back by one), the brick loop of
 * GLRenderer.cpp:1183-1230 with GLRaycaster::RenderHQMIPInLoop per non-empty brick (GLRaycaster.cpp:494-530:
 * GLRaycaster-MIP-Rot-FS.glsl, BE_MAX blending) and the Transfer-MIP-FS.glsl pass over the blended maximum
 * (GLRenderer.cpp:1232-1250).  The caller sets model_view = m_maMIPRotation * view as RenderHQMIPPreLoop does
 * (GLRenderer.cpp:1256-1285, GLRaycaster.cpp:481-492; perspective rays -- the m_bOrthoView switch is not built).
 * Brick emptiness follows the current render mode (ContainsData); the colour always comes from the 1D transfer
 * function.  The RGBA result (alpha 1) is read with tvk_read_rgba8/32f. */
int tvk_render_mip(tvk_ctx* ctx, int use_mip_lod, tvk_frame_stats* stats);
/* parity tap: the blended maximum image of the last MIP frame, width*height*(maximum, coverage) floats --
 * what Transfer-MIP-FS reads from m_pFBO3DImageNext[1] */
int tvk_read_mip_max(tvk_ctx* ctx, float* dst);
/* ClearView state of the isosurface mode (AbstrRenderer::SetCV / SetCVIsoValue / SetCVColor / SetCVSize /
 * SetCVContextScale / SetCVBorderScale / SetCVFocusPos, AbstrRenderer.cpp:1247-1360; defaults :131-136,171: colour
 * (1,0,0), size 5.5, context scale 1, border scale 60, focus position (0,0,0.5,1) in object space).  With ClearView
 * on, tvk_render_classic in isosurface mode runs GLRaycaster-ISO-CV-FS.glsl as a second pass per brick with the
 * focus isovalue (GLRaycaster.cpp:429-444) and composes with Compose-CV-FS.glsl (GLRenderer.cpp:2777-2795).
 * GLGridLeaper does not support ClearView (GLGridLeaper.h:47), so tvk_render ignores this state. */
int tvk_set_clearview(tvk_ctx* ctx, int enable, double cv_isovalue, const float color[3], float size,
                      float context_scale, float border_scale, const float focus_pos[4]);
/* parity tap: m_pFBOCVHit's two targets of the last ClearView frame */
int tvk_read_cv_buffers(tvk_ctx* ctx, float* cv_pos, float* cv_normal);
/* parity tap: the brick list of the last classic / MIP frame (depth sorted; MIP: key order) and its LoD */
int tvk_get_classic_brick_list(tvk_ctx* ctx, uint32_t* lod, tvk_classic_brick* dst, uint32_t cap, uint32_t* n);

/* ---- stereo (GLRenderer::ComputeViewAndProjection in stereo, GLRenderer::EndFrame) ------------------- */
/* AbstrRenderer::EStereoMode (Renderer/AbstrRenderer.h:128-134) */
typedef enum { TVK_SM_RB = 0, TVK_SM_SCANLINE = 1, TVK_SM_SBS = 2, TVK_SM_AF = 3 } tvk_stereo_mode;
/* tvk_compute_view for the two eyes: FLOATMATRIX4::BuildStereoLookAtAndProjection (Basics/Vectors.h:1215-1248:
 * off-centre frusta shifted by eye_dist * near / focal_length, views translated by +-eye_dist) as
 * GLRenderer::ComputeViewAndProjection calls it (GLRenderer.cpp:904-910), then modelView[eye] = rotation *
 * translation * view[eye].  Defaults of the reference: focal_length 1.0, eye_dist 0.02 (AbstrRenderer.cpp:142-143). */
int tvk_compute_stereo_view(tvk_render_params* left, tvk_render_params* right, uint32_t width, uint32_t height,
                            const float rotation[16], const float translation[16],
                            const float eye[3], const float at[3], const float up[3],
                            float fov_deg, float z_near, float z_far, float screen_space_error,
                            float focal_length, float eye_dist);
/* keeps the finished image of the current frame as eye 0 (left, m_pFBO3DImageNext[0]) or 1 (right, [1]) */
int tvk_stereo_keep_eye(tvk_ctx* ctx, int eye);
/* GLRenderer::EndFrame (GLRenderer.cpp:758-812): composes the two kept eye images with
 * Compose-Anaglyphs-FS / Compose-Scanline-FS / Compose-SBS-FS (split_coord = fSplitCoord, 0.5 at full resolution) /
 * Compose-AF-FS (alternating_frame_id = m_iAlternatingFrameID); eye_swap = m_bStereoEyeSwap.  The composed frame
 * becomes the image tvk_read_rgba8 / tvk_read_rgba32f / tvk_get_device_image return until the next frame. */
int tvk_stereo_compose(tvk_ctx* ctx, int mode, int eye_swap, int alternating_frame_id, float split_coord);

/* ---- depth pipeline (new; SURVEY 8e, alternative to binary swap) ------------------------- */
/* One STAGE of a depth-pipelined frame.  The volume is cut into slabs behind each other along the view axis
 * (clip_min / clip_max of the render params = this rank's slab); stage s marches every ray through slab s only, starting
 * from the state the stage in front handed over -- resume position + accumulated colour, the same two images a
 * resumed GLGridLeaper subframe starts from (GLGridLeaper-blend.glsl:130-137) -- and leaves the state for the next
 * stage: resume_pos.w = 1000 marks a finished ray (early termination or volume exit), anything else the depth at which
 * the next slab takes over.  in_resume_pos / in_resume_color: width*height float4 DEVICE images, NULL for the first
 * stage.  The last stage's image is the frame.  Early ray termination works across ranks as on one GPU, and with
 * consecutive frames in flight every rank traces 1/N of every ray.  1D / 2D transfer-function modes. */
int tvk_render_stage(tvk_ctx* ctx, const void* in_resume_pos, const void* in_resume_color, tvk_frame_stats* stats);
/* device images of the last stage: accumulated colour so far (= the frame after the last stage), and the two
 * hand-over images for the next stage */
int tvk_get_stage_outputs(tvk_ctx* ctx, void** image, void** resume_color, void** resume_pos);

/* ---- sort-last compositing (new) --------------------------------------- */
/* out = front + (1-front.a)*back on n_pixels premultiplied RGBA32F device pixels
 * (Compositing.glsl:33-38 / blend state GLRenderer.cpp:151-153); a front pixel with alpha > 0.99
 * (early ray termination, GLGridLeaper-blend.glsl:180) is kept as is, and a back image that would push alpha past 0.995 is
 * scaled so that the result ends AT 0.995 -- the middle of the interval (0.99, 1] in which the single-GPU ray would have
 * terminated inside the back block (alpha error <= 0.005).  The cut makes the operator non-associative: partial images
 * must be folded front to back, ((s0 over s1) over s2) ..., which is what tvk_composite_nway and the sort-last frame do.
 * out may alias front or back. */
int tvk_composite_over(tvk_ctx* ctx, const void* front, const void* back, void* out, uint64_t n_pixels);
/* float -> unorm8 (GL read-back conversion) on device buffers */
int tvk_quantize_rgba8(tvk_ctx* ctx, const void* rgba32f, void* rgba8, uint64_t n_pixels);

/* ---- environment switches (read once, at tvk_create / tvk_sortlast_init; all default to the measured-best setting) ----
 *   TVK_TILE_LPT=0      traversal launches run their tiles in dispatch order instead of longest-first by the previous launch's
 *                       per-tile cost (results never depend on the order)
 *   TVK_SL_PEER=0       sort-last: NCCL slice exchange instead of reading the peers' images through CUDA IPC peer memory
 *   TVK_SL_OVERLAP=0    sort-last: blend + gather of frame f in order on the render stream instead of on the internal stream
 *                       while frame f + 1 is traversed
 *   TVK_BRICKER_TMA=0   tvk_build_volume cuts bricks with the generic kernel instead of the TMA box loads
 *   TVK_BUILD_TRACE=1 / TVK_UPLOAD_TRACE=1   per-level build times / per-batch upload times on stderr
 *   TVK_SPLIT_COST=n    only in a -DTVK_PERSIST=1 build (measured slower): tiles above n turns are split */

/* ---- measurement ---------------------------------------------------------------------------------------------------
 * The ceiling of the traversal kernel's own fetch path: width*height rays in the kernel's warp tiles
 * march `steps` 0.5-voxel steps along `dir` through the resident pool doing ONLY the footprint loads and the packed
 * filter trees of the current mode (7-tap footprint for 2D-TF / lit modes, 1 tap otherwise) -- no page table,
 * classification, shading, compositing.  Needs a pool with FAST-path geometry (linear filter, ghost >= 2).
 * ms = device time of the probe launch; the rate is width*height*steps / ms. */
int tvk_probe_fetch(tvk_ctx* ctx, uint32_t steps, const float dir[3], float* ms, uint64_t* samples);

/* ---- sort-last across the GPUs of one box, inside the library (new) -----------------------------------
 * One process (or thread) per GPU, one tvk_ctx each, all with the same dataset description.  The finest brick grid is
 * cut into n_ranks convex blocks by recursive bisection; rank g traverses only its block (rays keep the single-GPU
 * ray's sample positions, see tvk_render_params.clip_min), then the frame is composited by DIRECT SEND: the image is
 * cut into n_ranks pixel slices, every rank sends slice p of its partial RGBA32F image to rank p and receives the
 * n_ranks-1 partials of its own slice in ONE grouped ncclSend/ncclRecv exchange over NVLink, folds them front to back in
 * the visibility order of the blocks with a hand-written kernel (over operator of Compositing.glsl:33-38 / the blend
 * state of GLRenderer.cpp:151-153, GL read-back conversion fused in) and the RGBA8 slices are gathered on rank 0 --
 * traversal, exchange, blend and gather are queued on one stream without host synchronisation in between.  NCCL
 * (libnccl.so.2, loaded on first use) carries the image exchange only.  Transfer-function modes (1D / 2D). */
#define TVK_COMM_ID_BYTES 128
/* how the brick grid is cut: OCTANT = longest axis of the block at every level (view independent: a rank's bricks never
 * change, the brick store can be sharded at the source); SCREEN = the axes most perpendicular to the view first (blocks
 * lie side by side on screen; re-cut when the view's dominant axes change) */
typedef enum { TVK_SL_OCTANT = 0, TVK_SL_SCREEN = 1, TVK_SL_PAIRED = 2 } tvk_sortlast_policy;
/* PAIRED: the grid is cut into 2 x n_ranks blocks (longest axis at every level, view independent) and rank r renders block
 * r AND its mirror block (the side flipped at the first cut of every axis) in two concurrent traversal launches.  For a
 * camera outside the volume one of the two lies on the near side and the other on the far side, so a view's load is
 * balanced by construction.  2 x n_ranks partial images are folded per pixel slice (n_ranks <= 8).  MEASURED SLOWER than
 * OCTANT on C3 (a far block still takes all of its own samples -- early termination does not reach
 * across blocks -- so pairing converts idle time into samples); kept as an option, not the default. */
typedef struct {
  tvk_frame_stats frame;       /* this rank's subframe; frame.ms_raycast = its traversal kernel */
  float ms_exchange;           /* device time from the end of the traversal to the gathered RGBA8 frame: slice exchange +
                                  n-way blend + gather */
  float ms_frame;              /* device time of the whole frame on this rank */
  uint64_t bytes_sent;         /* image bytes this rank sent (exchange + gather) */
  uint64_t slice_lo, slice_hi; /* the pixel range this rank composited */
  int32_t  peer_memory;        /* 1: the frame went through peer memory (no NCCL call on its path), 2: the same with the
                                  exchange overlapped with the next frame (see tvk_sortlast_flush), 0: NCCL exchange */
  float    ms_wait_peers;      /* peer-memory path: the part of ms_exchange spent waiting for the slowest rank's partial image */
} tvk_sortlast_stats;
/* rank 0 creates the id (ncclGetUniqueId) and hands it to the other ranks by whatever means the host has */
int tvk_sortlast_unique_id(uint8_t id[TVK_COMM_ID_BYTES]);
/* collective over all ranks: ncclCommInitRank on ctx's device.  n_ranks must be a power of two <= 16. */
int tvk_sortlast_init(tvk_ctx* ctx, const uint8_t id[TVK_COMM_ID_BYTES], int rank, int n_ranks, int policy);
int tvk_sortlast_shutdown(tvk_ctx* ctx);
/* this rank's block for the current render params (normalised volume space), the front-to-back order of the ranks
 * (order[0] = frontmost, n_ranks entries) and its pixel slice; any pointer may be NULL */
int tvk_sortlast_get_block(tvk_ctx* ctx, float clip_min[3], float clip_max[3], int* order, uint64_t* slice_lo,
                           uint64_t* slice_hi);
/* the `which`-th block of this rank (0, or 1 = its second block under TVK_SL_PAIRED) and the number of blocks in the
 * plan (= entries of `order` above: n_ranks, or 2 x n_ranks when paired); any pointer may be NULL */
int tvk_sortlast_get_block_of(tvk_ctx* ctx, int which, float clip_min[3], float clip_max[3], int* n_blocks);
/* collective: one subframe on every rank + compositing.  Bricks a rank missed are paged in afterwards as in tvk_render;
 * st->frame.converged is this rank's flag (the host ANDs it over the ranks when it needs a global one). */
int tvk_sortlast_frame(tvk_ctx* ctx, tvk_sortlast_stats* st);
/* Overlapped exchange (default on the peer-memory path with OCTANT / SCREEN; TVK_SL_OVERLAP=0 turns it off): the blend and
 * gather of frame f run on an internal high-priority stream while the render stream already traverses frame f + 1 -- a
 * rank runs up to one frame ahead, so waiting for the slowest rank of a view is hidden behind the next view's traversal.
 * Results are unchanged (same kernels, same order of the fold).  Then st->peer_memory = 2, st->ms_exchange /
 * st->ms_wait_peers describe the PREVIOUS frame's exchange (0 if it is still running) and st->ms_frame is the render
 * stream's share of this frame (traversal + publication).  The tvk_sortlast_read_* calls order themselves behind the
 * exchange; tvk_sortlast_flush makes the render stream (tvk_set_stream) wait for everything queued on the exchange stream
 * -- call it before recording an event that is meant to mark "all frames so far are gathered". */
int tvk_sortlast_flush(tvk_ctx* ctx);
/* rank 0: the composited frame, bottom row first (GLFrameCapture.cpp:72-85); async: dst is page-locked, the copy runs
 * on the library's copy stream, tvk_read_wait as for tvk_read_rgba8_async */
int tvk_sortlast_read_rgba8(tvk_ctx* ctx, uint8_t* dst, size_t pitch);
int tvk_sortlast_read_rgba8_async(tvk_ctx* ctx, uint8_t* dst_pinned, size_t pitch);
/* parity tap: this rank's composited slice as RGBA32F, (slice_hi - slice_lo) * 4 floats */
int tvk_sortlast_read_slice(tvk_ctx* ctx, float* dst);
/* host-only (no device, ctx may be NULL): the partition and visibility order tvk_sortlast_frame uses.  finest = bricks
 * per axis of the finest level, float_layout = vLODLayout[0] (GLVolumePool.cpp:89-107), extent = normalised volume
 * extent (AbstrRenderer.cpp:1102-1109).  clip_min / clip_max: n_ranks x 3 floats, order: n_ranks ints. */
int tvk_sortlast_plan(const uint32_t finest[3], const float float_layout[3], const double extent[3],
                      const float model_view[16], int n_ranks, int policy, float* clip_min, float* clip_max, int* order);
/* the n-way front-to-back fold on device images (slices[0] = frontmost): RGBA32F (may be NULL) and RGBA8 results */
int tvk_composite_nway(tvk_ctx* ctx, const void* const* slices, int n, void* out_rgba32f, void* out_rgba8,
                       uint64_t n_pixels);
/* Sort-last at the SOURCE: keep only the bricks of tvk_build_volume's store that touch the box (normalised volume
 * space; all LoDs -- coarser bricks that overlap several blocks are kept by each of them), so the memory of a rank's
 * brick store falls with the rank count.  Call before tvk_build_volume; {0,0,0},{1,1,1} (the default) keeps everything.
 * Min/max are still computed for every brick (visibility is global).  A request for a brick that was not kept fails
 * with TVK_ERR_SOURCE -- the traversal never asks for one (bricks outside the shard box are stepped through). */
int tvk_set_store_shard(tvk_ctx* ctx, const float clip_min[3], const float clip_max[3]);

#ifdef __cplusplus
}
#endif
#endif /* TVK_H */
