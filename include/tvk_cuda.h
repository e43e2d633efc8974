/*
 * tvk_cuda.h -- C ABI of the B200-native bricked ray-caster that replaces the
 * GLSL/GL hot path of Tuvok (ImageVis3D).
 *
 * Everything here is plain C: opaque handle, POD structs, caller-owned buffers,
 * int status returns (0 = TVK_OK), no exceptions, no torch / C++ types.  A
 * `CudaGridLeaper : AbstrRenderer` binds these entry points
 * instead of the GL objects the reference uses.  Each entry point names the
 * reference interface it replaces (paths relative to /root/reference/).
 *
 * Conventions shared with the reference:
 *   - matrices are FLOATMATRIX4 memory order (m11,m12,...,m44), row-vector
 *     convention v' = v*M (Tuvok/Basics/Vectors.h:434,936)
 *   - brick ids are (x,y,z,lod); the 1-D index is
 *     lodOffset[lod] + x + y*Lx + z*Lx*Ly (GLVolumePool.cpp:288-303)
 *   - metadata values: 0 MISSING, 1 CHILD_EMPTY, 2 EMPTY, >=3 pool slot + 3
 *     (GLVolumePool.h:62-72)
 *   - images are RGBA32F, row 0 = bottom row (GL window convention)
 */
#ifndef TVK_CUDA_H
#define TVK_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVK_OK               0
#define TVK_ERR_INVALID      1   /* bad argument / state                      */
#define TVK_ERR_CUDA         2   /* a CUDA runtime call failed                */
#define TVK_ERR_NOMEM        3   /* allocation failed                         */
#define TVK_ERR_UNSUPPORTED  4   /* valid in the reference, not built yet     */

#define TVK_BI_MISSING      0u
#define TVK_BI_CHILD_EMPTY  1u
#define TVK_BI_EMPTY        2u
#define TVK_BI_FLAG_COUNT   3u

#define TVK_MAX_LOD 16

/* AbstrRenderer::ERenderMode (Tuvok/Renderer/AbstrRenderer.h:142-147) */
#define TVK_RM_1DTRANS     0
#define TVK_RM_2DTRANS     1
#define TVK_RM_ISOSURFACE  2

/* GLVolumePool::MissingBrickStrategy (GLVolumePool.h:81-87) */
#define TVK_MBS_ONLY_NEEDED     0
#define TVK_MBS_REQUEST_ALL     1
#define TVK_MBS_SKIP_ONE_LEVEL  2
#define TVK_MBS_SKIP_TWO_LEVELS 3

/* pool sampler back-ends (both give GL_LINEAR / CLAMP_TO_EDGE semantics) */
#define TVK_SAMPLER_EXACT  0   /* fp32 trilinear from linear HBM (oracle arithmetic) */
#define TVK_SAMPLER_TEX3D  1   /* cudaArray + hardware filtering (1.8 fixed-point weights) */
#define TVK_SAMPLER_EXACT_Q4 2 /* EXACT arithmetic; pool stored as (y,z) 2x2 quads per voxel: 4x the
                                  memory, 2 loads per sample instead of 8 (same results as EXACT)   */
#define TVK_SAMPLER_GATHER 3   /* EXACT arithmetic; pool = LAYERED 2-D array (one layer per pool z), a sample's
                                  2x2x2 footprint = two four-texel gathers (tld4) through the texture units:
                                  1x the memory, 2 fetches per sample, same results as EXACT.  Limits: pool
                                  slots are laid out physically as X x Y x (<= 2048 / max_brick.z) layer groups */

typedef struct tvk_ctx tvk_ctx;

/* ------------------------------------------------------------------ context */
/* replaces tuvok::Context / BatchContext (Renderer/Context.h:46-59,
 * BatchRenderer/BatchContext.h:47-66): one context = one device + streams. */
int  tvk_create(int device, tvk_ctx** out);
void tvk_destroy(tvk_ctx* ctx);
const char* tvk_last_error(const tvk_ctx* ctx);
int  tvk_device_synchronize(tvk_ctx* ctx);
/* Run all render / builder / compositor work of this context on an existing CUDA stream
 * (e.g. the caller's, so that NCCL collectives and these kernels are ordered without host
 * synchronisation).  cuda_stream is a cudaStream_t; NULL restores the context's own stream. */
int  tvk_set_stream(tvk_ctx* ctx, void* cuda_stream);
/* Upper bound on the resident CTAs per SM of the persistent ray-march grid (0 = what the occupancy calculator
 * allows, the default).  At full occupancy K1 holds every register of an SM until the frame is done; one CTA less per
 * SM leaves room for the pager's gather / placement kernels, so that the paging of the NEXT frame's bricks (side
 * stream) proceeds while this frame marches. */
int  tvk_set_march_occupancy(tvk_ctx* ctx, uint32_t ctas_per_sm);

/* --------------------------------------------------------------- brick pool */
/* replaces GLVolumePool ctor + CreateGLResources (GLVolumePool.cpp:119-259,
 * 864-910) and the shader constants baked by GetShaderFragment (:317-483). */
typedef struct tvk_pool_desc {
  uint32_t volume_size[3];     /* LOD-0 domain size in voxels (m_volumeSize)            */
  uint32_t max_brick_size[3];  /* total brick size incl. overlap (m_maxTotalBrickSize)  */
  uint32_t overlap;            /* ghost voxels per side                                 */
  uint32_t bytes_per_voxel;    /* 1 (UNORM8) or 2 (UNORM16), single component           */
  uint32_t pool_capacity[3];   /* slots per axis; pool voxels = capacity*max_brick_size */
  uint32_t lod_count;          /* LargestSingleBrickLOD+1 (GLVolumePool.cpp:132)        */
  uint32_t sampler;            /* TVK_SAMPLER_*                                          */
} tvk_pool_desc;

int tvk_pool_create(tvk_ctx* ctx, const tvk_pool_desc* desc);
int tvk_pool_destroy(tvk_ctx* ctx);
/* total bricks over the lod_count pool levels (m_iTotalBrickCount) */
int tvk_pool_total_bricks(const tvk_ctx* ctx, uint32_t* out);

/* replaces GLVolumePool::UploadBrick -> glTexSubImage3D (GLVolumePool.cpp:673-717).
 * Copies n bricks (x-fastest voxels, sizes[3*i..] voxels each, possibly smaller
 * than max_brick_size) into pool slots slot[i] (linear slot index
 * x + y*Cx + z*Cx*Cy) on the pager's side stream.  src pointers may be pinned
 * host memory (on_device == 0; when it is mapped into the device address space a 256-brick batch
 * travels as copy-engine transfers -- one if its bricks follow one another in host memory, one per
 * run of adjacent bricks if the runs average three bricks or more -- and is otherwise gathered over
 * PCIe by a kernel; unmapped memory: one DMA per run of bricks) or device memory (on_device == 1).
 * on_device == 2: the caller states that ALL sources lie in ONE pinned, device-mapped host
 * allocation (e.g. the renderer's brick store): one pointer query instead of one per brick.
 * Metadata is NOT touched; patch it with tvk_pool_patch_metadata once
 * tvk_pool_upload_wait() has returned. */
int tvk_pool_upload_bricks(tvk_ctx* ctx, uint32_t n, const uint32_t* slot,
                           const void* const* src, const uint32_t* sizes,
                           int on_device);
int tvk_pool_upload_wait(tvk_ctx* ctx);
/* LZ4-compressed bricks (raw LZ4 block format, as lz4Compress writes them into a UVF file, ExtendedOctree/Lz4Compression.cpp)
 * inflated ON THE DEVICE: replaces the host-side decompression of ExtendedOctree::GetBrickData (ExtendedOctree.cpp:313-360)
 * for paging.  src[i] = the comp_len[i] compressed bytes of brick i in (pinned) host memory; only those bytes cross PCIe, a
 * kernel (one warp per brick) inflates them in HBM and the placement kernels put them into slot[i].  Blocks until the
 * bricks are in the pool; corrupt data -> TVK_ERR_INVALID, pool untouched.  Metadata as for tvk_pool_upload_bricks. */
int tvk_pool_upload_bricks_lz4(tvk_ctx* ctx, uint32_t n, const uint32_t* slot, const void* const* src, const uint32_t* comp_len,
                               const uint32_t* sizes);
/* the same kernel without a pool: n LZ4 blocks in host memory -> dst[i] (dst_len[i] bytes each, host memory) */
int tvk_lz4_inflate(tvk_ctx* ctx, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len);
/* zlib-compressed bricks (CT_ZLIB: one RFC 1950 stream per brick, zCompress, ExtendedOctree/ZlibCompression.cpp) inflated on
 * the device the same way: one warp per brick runs a DEFLATE decoder (stored, fixed and dynamic blocks; every lane follows
 * the bit stream, lane 0 writes literals, the warp shares the match copies).  codec = TVK_CODEC_LZ4 / TVK_CODEC_ZLIB. */
#define TVK_CODEC_LZ4  3u   /* = TVK_UVF_COMP_LZ4  */
#define TVK_CODEC_ZLIB 1u   /* = TVK_UVF_COMP_ZLIB */
int tvk_pool_upload_bricks_compressed(tvk_ctx* ctx, uint32_t codec, uint32_t n, const uint32_t* slot, const void* const* src,
                                      const uint32_t* comp_len, const uint32_t* sizes);
int tvk_zlib_inflate(tvk_ctx* ctx, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len);
/* The same ordering without blocking the host: everything issued on the render stream after this call (metadata
 * patches, the next ray-march) runs after the upload in flight has been placed.  Only for sources that stay valid
 * until then (the renderer's pinned brick store). */
int tvk_pool_upload_fence(tvk_ctx* ctx);

/* replaces UploadMetadataTexture / UploadMetadataTexel (GLVolumePool.cpp:912-953) */
int tvk_pool_set_metadata(tvk_ctx* ctx, const uint32_t* meta, uint32_t n);
/* every entry TVK_BI_MISSING, ordered on the render stream, no host wait (GLVolumePool::PH_Reset, GLVolumePool.cpp:261-284,
 * clears m_vBrickMetadata and re-uploads it) */
int tvk_pool_clear_metadata(tvk_ctx* ctx);
int tvk_pool_patch_metadata(tvk_ctx* ctx, uint32_t n, const uint32_t* brick_index,
                            const uint32_t* value);
int tvk_pool_get_metadata(tvk_ctx* ctx, uint32_t* meta, uint32_t n);

/* ---------------------------------------------------------------- visibility */
/* replaces GLVolumePool::RecomputeVisibility (+ForBrickPool, +ForOctree)
 * (GLVolumePool.cpp:1580-1718, 992-1018, 1021-1359).  minmax holds
 * {min,max} doubles per brick in pool index order (m_vMinMaxScalar);
 * grad_minmax may be NULL (then -DBL_MAX/+DBL_MAX as written by the TOC path,
 * ExtendedOctreeConverter.cpp:954-1000). */
int tvk_visibility_set_minmax(tvk_ctx* ctx, const double* minmax,
                              const double* grad_minmax, uint32_t n);
typedef struct tvk_visibility {
  uint32_t render_mode;        /* TVK_RM_*                                         */
  double   smin, smax;         /* 1D/2D TF non-zero value range; iso: smin = iso   */
  double   gmin, gmax;         /* 2D TF gradient range                             */
} tvk_visibility;
/* Recomputes every metadata entry on the device: entries >= FLAG_COUNT
 * (resident) keep their slot if the brick contains data, else become EMPTY;
 * all others are derived bottom-up from the min/max table.
 * counts[4] = (processed, inner EMPTY, inner CHILD_EMPTY, leaf empty) as
 * vEmptyBrickCount; may be NULL. */
int tvk_visibility_recompute(tvk_ctx* ctx, const tvk_visibility* vis, uint32_t* counts);

/* --------------------------------------------------------- transfer functions */
/* replaces GLRenderer::Changed1DTrans/Changed2DTrans texture upload
 * (GLRenderer.cpp:465-472, GPUMemMan.cpp:421-427): RGBA8, NEAREST, clamp. */
int tvk_set_tf1d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t n);
int tvk_set_tf2d(tvk_ctx* ctx, const uint8_t* rgba, uint32_t w, uint32_t h);
/* The context keeps the bytes of the last 2-D table it was given in a page-locked block (they travel to the device from
 * there).  tvk_tf2d_staging returns that block; passing it back to tvk_set_tf2d uploads the table again without a host
 * copy -- what a host that switches between the 1-D and the 2-D function does (the device holds ONE table), instead of
 * keeping a second copy of a 4 MB table and copying it twice per edit. */
int tvk_tf2d_staging(tvk_ctx* ctx, const uint8_t** rgba, uint32_t* w, uint32_t* h);

/* ------------------------------------------------------------------ ray-march */
/* replaces GLGridLeaper::Raycast + SetupRaycastShader + FillRayEntryBuffer
 * (GLGridLeaper.cpp:754-870, 690-752, 579-620) and the shaders
 * GLGridLeaper-blend.glsl / -iso.glsl / Method-*.glsl / Compose-FS.glsl. */
typedef struct tvk_raycast_params {
  uint32_t width, height;          /* render target size                              */
  float eye_to_model[16];          /* mEyeToModel (GLGridLeaper.cpp:560-573)          */
  float model_to_eye[16];          /* its inverse (iso: mModelToEye)                  */
  float model_view_it[16];         /* MV^-1 uploaded transposed (iso: mModelViewIT)   */
  float inv_projection[16];        /* m_mProjection^-1 (near-plane / pixel rays)      */
  float lod_factor;                /* CullingLOD::GetLoDFactor (CullingLOD.cpp:65-66) */
  float level_zero_wse;            /* fLevelZeroWorldSpaceError (GLVolumePool.cpp:801)*/
  float sample_rate_modifier;
  float trans_scale;               /* fTransScale (GLRenderer.cpp:1813-1819)          */
  float gradient_scale;            /* fGradientScale (GLGridLeaper.cpp:728-732)       */
  float isoval;                    /* fIsoval, normalised                             */
  uint32_t render_mode;            /* TVK_RM_*                                        */
  uint32_t lighting;               /* m_bUseLighting                                  */
  float light_ambient[3];
  float light_diffuse[3];
  float light_specular[3];
  float model_light_dir[3];        /* vModelSpaceLightDir                             */
  float model_eye_pos[3];          /* vModelSpaceEyePos                               */
  float domain_scale[3];           /* vDomainScale = 1/(scale/min(scale))             */
  float iso_color[3];              /* Compose-FS: vLightDiffuse*isoColor              */
  float compose_light_dir[3];      /* Compose-FS vLightDir (eye space)                */
  uint32_t new_view;               /* 1: rr.isBlank -> recompute ray entry, clear     */
  uint32_t missing_strategy;       /* TVK_MBS_* (default SKIP_TWO_LEVELS)             */
  uint32_t rehash_count;           /* GLHashTable rehash count (default 10)           */
  uint32_t hash_table_size;        /* GLHashTable size (default 509)                  */
  uint32_t clip_enabled;           /* clip plane (model space, normalised coords)     */
  float clip_plane[4];
  float box_min[3];                /* rays are clipped to this sub-box of the unit cube (normalised   */
  float box_max[3];                /* volume coords): {0,0,0}-{1,1,1}, or a sort-last partition       */
  uint32_t count_samples;          /* accumulate executed samples into the stats      */
  /* sort-last: the volume of this context is one box of a larger one.  Lighting is evaluated in the GLOBAL volume's model
   * space -- model_light_dir / model_eye_pos refer to it -- so a sample's (box-local) model position p is first taken there:
   * p * (1 + model_pos_scale_m1) + model_pos_offset.  All zero = one volume, positions used as they are (bit-identical). */
  float model_pos_scale_m1[3];
  float model_pos_offset[3];
} tvk_raycast_params;

/* One sub-frame = one GLGridLeaper::Render3DRegion pass: clears the request
 * table, marches all unfinished rays, leaves image + resume state on device. */
int tvk_raycast_subframe(tvk_ctx* ctx, const tvk_raycast_params* p);

/* replaces GLHashTable::GetData (GLHashTable.cpp:90-110): compacted requests in
 * table order, 4 uint32 (x,y,z,lod) each.  *count <= cap. */
int tvk_read_requests(tvk_ctx* ctx, uint32_t* xyzl, uint32_t cap, uint32_t* count);

/* replaces glReadPixels on m_pFBO3DImage (float RGBA, premultiplied; for iso
 * mode this is the Compose-FS shaded image).  rgba = width*height*4 floats. */
int tvk_read_image(tvk_ctx* ctx, float* rgba);
/* replaces glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) of the same target (GLFrameCapture.cpp:47-86): clamp, *255, round
 * to nearest on the device, width*height*4 bytes copied back.  flip_rows != 0 returns the top row first
 * (FrameCapture::SaveImage, FrameCapture.h:90-98). */
int tvk_read_image_rgba8(tvk_ctx* ctx, uint8_t* rgba8, int flip_rows);
/* resume buffers, for tests: rayResumePos / rayResumeColor (iso: hit normal) */
int tvk_read_resume(tvk_ctx* ctx, float* pos, float* color);
/* device pointer of the current image (for the sort-last compositor / interop) */
int tvk_image_device_ptr(tvk_ctx* ctx, void** dptr);

typedef struct tvk_stats {
  uint64_t samples;        /* executed ComputeColorFromVolume / GetVolumeHit calls  */
  uint64_t rays;           /* rays that entered the volume                          */
  uint64_t brick_steps;    /* GetBrick calls                                        */
  float    last_raycast_ms;/* CUDA-event time of the last ray-march launch          */
  uint32_t launches;       /* kernels launched by this context since creation       */
} tvk_stats;
int tvk_get_stats(tvk_ctx* ctx, tvk_stats* out, int reset);

/* bit-exactness probes used by the parity tests (R3, R2, R5 of SURVEY 8a):
 * evaluate ComputeLOD / ComputeBrickCoords+GetBrickIndex / hash Serialize on
 * the device for n inputs. */
int tvk_probe_lod(tvk_ctx* ctx, const float* depth, uint32_t n, float lod_factor,
                  float level_zero_wse, uint32_t* lod_out);
int tvk_probe_brick_index(tvk_ctx* ctx, const float* norm_xyz, const uint32_t* lod,
                          uint32_t n, uint32_t* index_out, uint32_t* serial_out);

/* ------------------------------------------------------------------- builder */
/* replaces ExtendedOctreeConverter::Convert (ExtendedOctreeConverter.cpp:128-280:
 * PermuteInputData, ComputeHierarchy/DownsampleBrick, FillOverlap,
 * ComputeBrickStats) + Histogram1DDataBlock::Compute +
 * Histogram2DDataBlock::Compute, as streaming kernels over a volume resident
 * in HBM.  Output bricks use the ExtendedOctree ToC order and sizes. */
typedef struct tvk_build_desc {
  uint32_t volume_size[3];
  uint32_t brick_size;         /* cubic max brick size incl. overlap (default 64) */
  uint32_t overlap;            /* default 2                                        */
  uint32_t bytes_per_voxel;    /* 1 or 2 (unsigned)                                */
  uint32_t use_median;         /* 0 mean (default) / 1 median                      */
  uint32_t clamp_to_edge;      /* 0 (default) / 1                                  */
} tvk_build_desc;

typedef struct tvk_build_info {
  uint32_t lod_count;                      /* all LODs down to 1x1x1            */
  uint32_t lod_size[TVK_MAX_LOD][3];       /* voxels                            */
  uint32_t lod_bricks[TVK_MAX_LOD][3];     /* brick counts                      */
  uint64_t lod_offset[TVK_MAX_LOD];        /* first ToC index of each LOD       */
  uint64_t total_bricks;
  uint64_t total_bytes;                    /* size of the brick store           */
} tvk_build_info;

/* metadata only (ExtendedOctree::ComputeMetadata, ExtendedOctree.cpp:188-243) */
int tvk_build_plan(const tvk_build_desc* desc, tvk_build_info* info);

/* Runs the whole builder.  raw: the x-fastest input volume (host pointer, or
 * device pointer if raw_on_device).  On return the context owns a device
 * resident "brick store" (all bricks, ToC order, tightly packed), a min/max
 * table and the histograms; fetch them with the getters below. */
int tvk_build_octree(tvk_ctx* ctx, const tvk_build_desc* desc, const void* raw,
                     int raw_on_device, tvk_build_info* info);
/* per brick: offset into the store (bytes) and size in voxels */
/* Streamed build for raw volumes that do not fit in device memory next to their brick store (replaces the
 * out-of-core loop of ExtendedOctreeConverter::Convert, ExtendedOctreeConverter.cpp:118-300, which pages raw data
 * through a brick cache): LOD 0 is fed in z-slabs of WHOLE BRICK ROWS, in ascending order; every slab yields its LOD-0
 * bricks (ghost cells, min/max), its share of the 1-D histogram and of the flat LOD-1 volume; tvk_build_stream_end builds
 * the coarser levels from LOD 1, which stays resident (1/7 of the raw size for the whole pyramid).  The result is the
 * state tvk_build_octree leaves (same ToC, bricks, min/max, histogram), except that no flat LOD 0 is kept, so
 * tvk_build_hist2d is not available.
 *   store: NULL = the brick store is allocated in device memory; otherwise caller-owned memory of
 *          info.total_bytes bytes -- device memory, or pinned host memory mapped into the device address space
 *          (cudaHostAlloc), in which case bricks are written over PCIe as they are cut.
 *   tvk_build_stream_rows: the raw rows [z0, z0 + rows) a slab must hold to build brick rows [bz0, bz1)
 *          (inner rows plus `overlap` ghost rows either side, clipped to the volume).
 *   tvk_build_stream_slab: raw_slab = device pointer to rows [z0, z0 + rows), x fastest; builds brick rows
 *          [bz0, bz1); bz0 must be where the previous slab stopped. */
int tvk_build_stream_begin(tvk_ctx* ctx, const tvk_build_desc* desc, void* store, tvk_build_info* info);
int tvk_build_stream_rows(const tvk_build_desc* desc, uint32_t bz0, uint32_t bz1, uint32_t* z0, uint32_t* rows);
int tvk_build_stream_slab(tvk_ctx* ctx, const void* raw_slab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1);
int tvk_build_stream_end(tvk_ctx* ctx, tvk_build_info* info);

/* ------------------------------------------------------------------ re-bricking */
/* replaces DynamicBrickingDS (IO/DynamicBrickingDS.cpp; IOManager::LoadRebrickedDataset, IOManager.cpp:1281-1320):
 * a dataset with large bricks is presented with smaller ones.  Every target brick is a sub-box of ONE source brick
 * (ghost voxels included: the target's ghost cells are the source brick's own voxels), target index / source index /
 * offset as SourceBrickIndex (:273-303), OffsetIntoSource (:305-336) and ComputedTargetBrickSize (:414-435) define
 * them; the LOD pyramid is the source's (Rebrick(), :1090-1180, adds no level).  The reference serves target bricks one
 * GetBrick() at a time from a host cache of source bricks; here the whole target store is cut on the device in one
 * launch.  Brick sizes INCLUDE the 2*overlap ghost voxels; overlap must be 2 (the reference hard-codes 4 ghost voxels in
 * BrickSansGhost, :208-218).  Target bricks are ordered LOD by LOD, x fastest (BrickKey index order, :1148). */
#define TVK_MM_SOURCE      0   /* a target brick reports its source brick's min/max (DynamicBrickingDS.h MinMaxMode) */
#define TVK_MM_PRECOMPUTE  1   /* min/max of the target brick's voxels, computed when the dataset is opened          */
#define TVK_MM_DYNAMIC     2   /* the same values, computed when asked for (identical here: one kernel does both)    */
typedef struct tvk_rebrick_desc {
  uint32_t lod_count;
  uint32_t lod_size[TVK_MAX_LOD][3];   /* voxels per LOD of the SOURCE dataset (GetDomainSize)     */
  uint32_t src_brick[3];               /* GetMaxBrickSize() of the source, ghost voxels included   */
  uint32_t tgt_brick[3];               /* wanted brick size, ghost voxels included                 */
  uint32_t overlap;                    /* 2                                                         */
  uint32_t bytes_per_voxel;            /* 1 or 2                                                    */
} tvk_rebrick_desc;
typedef struct tvk_rebrick_info {
  uint32_t tgt_brick[3];                   /* min(tgt, src) per axis (IOManager.cpp:1300-1304)   */
  uint32_t lod_bricks[TVK_MAX_LOD][3];     /* target layout                                      */
  uint64_t lod_offset[TVK_MAX_LOD];        /* first target index of each LOD                     */
  uint32_t src_lod_bricks[TVK_MAX_LOD][3]; /* source layout                                      */
  uint64_t src_lod_offset[TVK_MAX_LOD];
  uint64_t total_bricks, src_total_bricks;
  uint64_t total_bytes;                    /* target store, every brick 256-byte aligned         */
} tvk_rebrick_info;
/* layout only; TVK_ERR_INVALID if a target core size does not divide the source core size (the reference throws) */
int tvk_rebrick_plan(const tvk_rebrick_desc* desc, tvk_rebrick_info* info);
/* per target brick: byte offset in the target store, size, index of its source brick, voxel offset inside it */
int tvk_rebrick_toc(const tvk_rebrick_desc* desc, uint64_t* offsets, uint32_t* sizes3, uint64_t* src_index, uint32_t* src_offset3, uint64_t n);
/* Cuts the target store.  src_store: all source bricks (device memory, or pinned host memory mapped into the device),
 * brick i at src_offsets[i] with src_sizes3[3*i..]; dst_store: device memory of info.total_bytes bytes.
 * minmax2 (host, 2 doubles per target brick, may be NULL): min/max of each target brick's voxels. */
int tvk_rebrick(tvk_ctx* ctx, const tvk_rebrick_desc* desc, const void* src_store, const uint64_t* src_offsets,
                const uint32_t* src_sizes3, uint64_t n_src, void* dst_store, double* minmax2);

/* the caller is about to free the device volume it passed with raw_on_device = 1: the build state drops its pointer
 * (tvk_build_hist2d is refused afterwards) */
int tvk_build_forget_raw(tvk_ctx* ctx);
int tvk_build_get_toc(tvk_ctx* ctx, uint64_t* offsets, uint32_t* sizes3, uint64_t n);
int tvk_build_get_minmax(tvk_ctx* ctx, double* minmax2, uint64_t n);
int tvk_build_get_bricks(tvk_ctx* ctx, void* host_dst, uint64_t offset, uint64_t bytes);
int tvk_build_store_device_ptr(tvk_ctx* ctx, void** dptr);
/* CUDA-event time (ms) of the builder kernels of the last tvk_build_octree (allocations and the
 * host->device copy of the raw volume excluded) */
int tvk_build_last_ms(tvk_ctx* ctx, float* kernels_ms);
/* the same for the last tvk_build_hist2d (max-gradient pass + binning pass) */
int tvk_build_hist2d_last_ms(tvk_ctx* ctx, float* kernels_ms);
/* 1-D histogram of LOD-0 inner voxels, 2^bits bins, *n = trimmed size
 * (Histogram1DDataBlock.cpp:60-160) */
int tvk_build_hist1d(tvk_ctx* ctx, uint64_t* bins, uint32_t cap, uint32_t* n);
/* 2-D histogram (value x gradient magnitude), bins_x * 256 entries
 * (Histogram2DDataBlock.cpp:348-455) */
int tvk_build_hist2d(tvk_ctx* ctx, uint32_t bins_x, double max_nonzero_value,
                     uint64_t* bins, float* max_grad_magnitude);
int tvk_build_release(tvk_ctx* ctx);

/* --------------------------------------------------------------- quantisation */
/* replaces the reference's pre-bricking passes over the raw input: the dispatcher quantize()
 * (IO/RAWConverter.cpp:196-352) with Quantize<T,U> / BinningQuantize<T,U> (IO/Quantize.h:429-686),
 * AbstrConverter::Process8Bits and ::QuantizeTo8Bit (IO/AbstrConverter.cpp:75-246), for one component
 * and one time step.  Input element types: */
#define TVK_DT_I8  0
#define TVK_DT_U8  1
#define TVK_DT_I16 2
#define TVK_DT_U16 3
#define TVK_DT_I32 4
#define TVK_DT_U32 5
#define TVK_DT_F32 6
#define TVK_DT_I64 7
#define TVK_DT_U64 8
#define TVK_DT_F64 9
typedef struct tvk_quantize_result {
  uint32_t out_bits;    /* component size after the call: 8 or 16                                     */
  uint32_t generated;   /* 1: dst holds the quantised stream (uint8 / uint16); 0: the input is used   */
                        /*    as it is (the reference returns its source file in that case)           */
  uint32_t hist_size;   /* bins written to hist; 0 = none (Histogram1DDataBlock::Compute runs later)  */
  uint32_t bin_count;   /* iBinCount of Quantize<T,U> (unsigned integer inputs; diagnostic)           */
  float    device_ms;   /* CUDA-event time of all passes                                              */
} tvk_quantize_result;
/* src: n elements (host pointer, or 16-byte aligned device pointer if src_on_device).
 * dst_device: 16-byte aligned device buffer of at least 2*n bytes.  hist: host, 4096 entries. */
int tvk_quantize(tvk_ctx* ctx, uint32_t type, uint64_t n, const void* src, int src_on_device,
                 int to8bit, void* dst_device, uint64_t* hist, tvk_quantize_result* result);

/* -------------------------------------------------------- sort-last compositor */
/* new in this build (no reference counterpart): ordered "under"
 * blend of n premultiplied RGBA32F partial images, front-to-back in the given
 * order, using Compositing.glsl:33-38 arithmetic for premultiplied input.
 * Asynchronous on the context's stream (n <= 64); synchronise before reading out_dptr. */
int tvk_composite_ordered(tvk_ctx* ctx, const void* const* partial_dptrs, uint32_t n,
                          uint64_t pixels, void* out_dptr);

/* ------------------------------------------------- classic per-brick ray-caster */
/* replaces GLRaycaster::Render3DPreLoop / Render3DInLoop / SetBrickDepShaderVars / RenderBox /
 * ComputeEyeToTextureMatrix (Renderer/GL/GLRaycaster.cpp:239-300,327-479,589-612) with the shaders
 * GLRaycaster-1D-FS.glsl / GLRaycaster-1D-light-FS.glsl (+ VRender1D[Lit].glsl, Volume3D.glsl, lighting.glsl,
 * Compositing.glsl) and the frame buffer's ONE_MINUS_DST_ALPHA / ONE blend (GLRenderer.cpp:152-153): the bricks
 * of ONE level of detail, sorted front to back by the host (AbstrRenderer::BuildSubFrameBrickList), are marched
 * one after the other for every pixel.  Like the reference, ray entry points pass through half precision (the
 * RGBA16F entry buffer, GLRaycaster.cpp:97) and early ray termination works per brick.  Render modes: 1D and 2D
 * transfer function with or without lighting (GLRaycaster-2D[-light]-FS.glsl), and the isosurface pass
 * (GLRaycaster-ISO-FS.glsl + RefineIsosurface.glsl: per brick first hit, refined, kept if it passes the GL_LESS depth
 * test against the hits of the bricks before it) followed by Compose-FS.glsl.  Bricks are read from a
 * device-resident brick store (tvk_build_store_device_ptr). */
typedef struct tvk_classic_brick {
  float center[3], extension[3];     /* Brick::vCenter / vExtension                                  */
  float eye_to_tex[16];              /* ComputeEyeToTextureMatrix for this brick (row-vector order)   */
  float ray_step;                    /* fRayStepsize (SetBrickDepShaderVars)                          */
  uint32_t voxels[3];                /* brick size incl. overlap                                      */
  uint64_t store_offset;             /* byte offset of the brick in the store                         */
} tvk_classic_brick;
typedef struct tvk_classic_params {
  uint32_t width, height;
  float inv_model_view[16];          /* renderRegion.modelView^-1                                     */
  float normal_matrix[9];            /* gl_NormalMatrix = (modelView^-1)^T, upper 3x3, row-vector use */
  float inv_projection[16];
  float step_scale;                  /* fStepScale (opacity correction exponent)                      */
  float trans_scale;                 /* fTransScale                                                   */
  uint32_t lighting;
  float light_ambient[3], light_diffuse[3], light_specular[3], light_dir[3];
  float domain_scale[3];             /* vDomainScale = 1 / scale                                      */
  uint32_t bytes_per_voxel;
  uint32_t render_mode;              /* TVK_RM_1DTRANS / TVK_RM_2DTRANS / TVK_RM_ISOSURFACE             */
  float gradient_scale;              /* fGradientScale, 2D TF (GLRenderer.cpp:1840-1841)               */
  float isoval;                      /* fIsoval = GetNormalizedIsovalue()                              */
  float proj_param[2];               /* vProjParam (CullingLOD::GetDepthScaleParams, CullingLOD.cpp:69-72) */
  float iso_color[3];                /* isosurface colour (ComposeSurfaceImage multiplies the diffuse by it) */
  uint32_t clip_enabled;             /* m_bClipPlaneOn: RenderBox clips the brick's box with the plane (GLRaycaster.cpp:327-337) */
  float clip_plane[4];               /* model-space plane, unit normal; n.p + d <= 0 is kept (Clipper::BoxPlane)         */
} tvk_classic_params;
/* bricks: host array, front to back.  Result: the context's image (tvk_read_image). */
int tvk_raycast_classic(tvk_ctx* ctx, const tvk_classic_params* p, const tvk_classic_brick* bricks, uint32_t n,
                        const void* store_dptr);

/* ClearView (isosurface mode of the classic renderer): replaces the second pass of GLRaycaster::Render3DInLoop
 * (GLRaycaster.cpp:438-452, GLRaycaster-ISO-CV-FS.glsl: per brick a second surface at the ClearView isovalue, searched from
 * the first surface's hit when that lies in the same brick; its own depth test on the ray EXIT depth; hit targets in half
 * precision like m_pFBOCVHit) and the ClearView branch of GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2776-2795,
 * Compose-CV-FS.glsl: context surface, focus surface, blended by the distance to the picked point and by a screen-space
 * curvature estimate of the context surface's normals, dark ring at the lens border).  The march writes four hit buffers,
 * a second launch composes. */
typedef struct tvk_clearview_params {
  float cv_isoval;               /* GetNormalizedCVIsovalue (AbstrRenderer.cpp:426-438)                        */
  float cv_color[3];             /* m_vCVColor (vLightDiffuse2 = diffuse * colour)                             */
  float size;                    /* vCVParam.x = m_fCVSize (default 5.5)                                       */
  float context_scale;           /* vCVParam.y = m_fCVContextScale (1)                                         */
  float border_scale;            /* vCVParam.z = m_fCVBorderScale (60)                                         */
  float pick_pos[3];             /* vCVPickPos = m_vCVPos * modelView (eye space)                              */
} tvk_clearview_params;
int tvk_raycast_classic_cv(tvk_ctx* ctx, const tvk_classic_params* p, const tvk_clearview_params* cv, const tvk_classic_brick* bricks,
                           uint32_t n, const void* store_dptr);
/* the first surface's hit (eye-space position, w != 0: hit) of the last isosurface frame -- classic renderer with or without
 * ClearView, or the ray-guided renderer (rayHitPos of GLGridLeaper-iso.glsl) -- at pixel (x, y), y from the bottom: what GLRenderer::Pick (GLRenderer.cpp:2856-2872) and
 * GLRenderer::CVFocusHasChanged (:2833-2851) read back from m_pFBOIsoHit under the mouse */
int tvk_read_iso_hit(tvk_ctx* ctx, uint32_t x, uint32_t y, float out[4]);

/* ------------------------------------------------------------ 2-D slice / MIP views */
/* replaces GLRenderer::Render2DView's slice and MIP-preview path (GLRenderer.cpp:1057-1176: the finest LOD that is ONE
 * brick, GLRenderer::RenderSlice :922-1014 for the textured quad, and the shaders 1D-slice-FS.glsl / 2D-slice-FS.glsl /
 * MIP-slice-FS.glsl + Transfer-MIP-FS.glsl).  One launch: every pixel of the region finds its texture coordinate on the
 * quad, samples the brick (GL_LINEAR, clamp to edge) and classifies -- slice: transfer function, colour multiplied by
 * its opacity, alpha 1; MIP: maximum over all slices of the brick along the view axis, then the 1-D transfer function's
 * colour with alpha 1 (Transfer-MIP ignores opacity).  Pixels outside the quad keep the clear colour (0,0,0,1).
 * The full-resolution rotating MIP of capture mode is tvk_raycast_mip below. */
#define TVK_WM_SAGITTAL 0   /* slices along x (RenderRegion::EWindowMode, RenderRegion.h:53-59) */
#define TVK_WM_AXIAL    1   /* along y */
#define TVK_WM_CORONAL  2   /* along z */
typedef struct tvk_slice_params {
  uint32_t width, height;        /* the region = the whole target                                              */
  uint32_t window_mode;          /* TVK_WM_*                                                                   */
  uint32_t use_mip;              /* 0 slice, 1 MIP preview                                                     */
  float    slice;                /* fSliceIndex: texture coordinate along the view axis (slice index / domain) */
  uint32_t voxels[3];            /* vVoxelCount: size of the brick, ghost voxels included                      */
  uint32_t bytes_per_voxel;
  float    min_coords[3];        /* vMinCoords / vMaxCoords (:1141-1142), already swapped for flipView          */
  float    max_coords[3];
  float    quad[2];              /* v2AspectRatio: half extent of the quad in normalised device coordinates    */
  uint32_t render_mode;          /* slice: TVK_RM_1DTRANS or TVK_RM_2DTRANS; MIP always uses the 1-D function  */
  float    trans_scale;          /* fTransScale                                                                */
  float    gradient_scale;       /* fGradientScale (2-D function)                                              */
} tvk_slice_params;
/* brick: the single brick (x fastest), device memory or host memory (copied up).  Result: the context's image. */
int tvk_render_slice(tvk_ctx* ctx, const tvk_slice_params* p, const void* brick);

/* ------------------------------------------------------------ full-resolution MIP (capture mode) */
/* replaces the other branch of GLRenderer::Render2DView (GLRenderer.cpp:1182-1250): RenderHQMIPPreLoop (:1256-1283),
 * GLRaycaster::RenderHQMIPPreLoop / RenderHQMIPInLoop (GLRaycaster.cpp:481-529: per brick the front faces into the RGBA16F
 * entry buffer, the back faces with GLRaycaster-MIP-Rot-FS.glsl under GL_MAX blending) and Transfer-MIP-FS.glsl.  One launch
 * walks the brick list per pixel (the maximum does not depend on the order) and classifies: colour of the 1-D transfer
 * function at the maximum, opacity ignored, alpha 1; pixels no brick covers are black.  bricks: every brick of the LOD
 * AbstrRenderer::PlanHQMIPFrame picks (AbstrRenderer.cpp:1214-1245), described like the classic path's
 * (tvk_classic_brick: centre, extension, eye-to-texture matrix, fRayStepsize, voxel counts, offset in the store).
 * inv_projection: the inverse of the orthographic projection of m_bOrthoView (:1183-1197) or of the perspective one;
 * inv_model_view: (m_maMIPRotation [* m_mView[0]])^-1. */
typedef struct tvk_mip_params {
  uint32_t width, height;
  float inv_model_view[16];
  float inv_projection[16];
  uint32_t bytes_per_voxel;
  float trans_scale;             /* fTransScale */
} tvk_mip_params;
int tvk_raycast_mip(tvk_ctx* ctx, const tvk_mip_params* p, const tvk_classic_brick* bricks, uint32_t n, const void* store_dptr);

/* ------------------------------------------------------------------------ stereo */
/* replaces the per-eye frame buffers of GLGridLeaper / GLRenderer (m_pFBO3DImageNext[2], m_pFBORayStart[2],
 * m_pFBOStartColor[2]; GLGridLeaper.cpp:922-958) and GLRenderer::EndFrame's stereo compose (GLRenderer.cpp:757-812 with
 * Compose-Anaglyphs-FS.glsl, Compose-Scanline-FS.glsl, Compose-SBS-FS.glsl, Compose-AF-FS.glsl).
 * tvk_select_eye: the following sub-frames, read-backs and resume state refer to that eye's frame state (eye 0 = left is
 * what a context starts with; eye 1's buffers appear with its first sub-frame, which must be a new view).  Both eyes share
 * ONE request table: a sub-frame of eye 0 clears it, a sub-frame of eye 1 adds to what eye 0's left, and
 * tvk_read_requests after the pair returns the union -- as Render3DRegion clears the table once, marches both eyes and
 * reads it once.
 * tvk_compose_stereo: both eyes' pictures -> one, which tvk_read_image / _rgba8 / tvk_image_device_ptr return until the
 * next launch that draws.  swap_eyes = m_bStereoEyeSwap; split_coord = fSplitCoord of the side-by-side mode (0.5 at full
 * resolution); alternating_frame_id = m_iAlternatingFrameID of the alternating-frame mode.  The eye pictures are
 * GL_NEAREST textures drawn one to one; the side-by-side mode reads at twice the x coordinate, i.e. texel 2x+1 of the
 * left picture in the left half, texel 2x+1-2*split*width of the right one in the right half. */
#define TVK_SM_RB       0   /* red/blue anaglyph: (grey(left), grey(right)/2, grey(right), max alpha), grey = (0.3,0.59,0.11).rgb */
#define TVK_SM_SCANLINE 1   /* even lines left, odd lines right */
#define TVK_SM_SBS      2   /* side by side */
#define TVK_SM_AF       3   /* alternating frames */
int tvk_select_eye(tvk_ctx* ctx, uint32_t eye);
/* one eye's own picture (m_pFBO3DImageNext[eye]), whichever eye is selected and whether or not a composition is shown */
int tvk_read_eye_image(tvk_ctx* ctx, uint32_t eye, float* rgba);
int tvk_compose_stereo(tvk_ctx* ctx, uint32_t mode, uint32_t swap_eyes, float split_coord, uint32_t alternating_frame_id);

/* ---------------------------------------------- sort-last exchange over peer memory */
/* Direct-send compositing fused into the kernels (no reference counterpart): rank r owns scan-line
 * stripe r.  After tvk_exchange_begin_frame the next tvk_raycast_subframe stores every finished pixel into the
 * stripe owner's receive buffer with peer stores (NVLink), tvk_exchange_composite blends the `world` partial
 * stripes of this rank front to back (order[k] = rank of the k-th nearest slab) once they have all arrived and
 * sends the result to rank 0, tvk_exchange_final orders rank 0's stream behind the arrival of all stripes.
 * One process per GPU: exchange the 64-byte handle of tvk_exchange_create between the ranks and pass all of
 * them (rank order) to tvk_exchange_open_peers.  Several ranks inside one process (tests): pass the pointers of
 * tvk_exchange_local_base to tvk_exchange_set_peers.  frame_id >= 1, the same on all ranks, +1 per frame.
 * Waits are bounded (4 s): tvk_exchange_status reports a time-out instead of a hang.  world <= 16. */
int tvk_exchange_create(tvk_ctx* ctx, uint32_t world, uint32_t rank, uint32_t width, uint32_t height,
                        void* ipc_handle64);
int tvk_exchange_open_peers(tvk_ctx* ctx, const void* handles);
int tvk_exchange_set_peers(tvk_ctx* ctx, void* const* bases);
/* Several exchange members in one process (one GPU rendering more than one box): bases[r] != NULL = region r is already
 * addressable here (a sibling context's region from tvk_exchange_local_base, or a mapping a sibling opened, from
 * tvk_exchange_peer_base); every other region is opened from handles[r] (64 bytes each).  Either argument may be NULL. */
int tvk_exchange_connect(tvk_ctx* ctx, const void* handles, void* const* bases);
int tvk_exchange_peer_base(tvk_ctx* ctx, uint32_t r, void** base);
int tvk_exchange_local_base(tvk_ctx* ctx, void** base);
int tvk_exchange_begin_frame(tvk_ctx* ctx, uint32_t frame_id);
int tvk_exchange_composite(tvk_ctx* ctx, const uint32_t* order, uint32_t frame_id);
int tvk_exchange_final(tvk_ctx* ctx, uint32_t frame_id, void** final_dptr);
/* glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) of the sort-last result (rank 0): tvk_exchange_final + the conversion of
 * tvk_read_image_rgba8 + one device-to-host copy of width*height*4 bytes; blocks until the frame is in rgba8. */
int tvk_exchange_read_final_rgba8(tvk_ctx* ctx, uint32_t frame_id, uint8_t* rgba8, int flip_rows);
int tvk_exchange_status(tvk_ctx* ctx, uint32_t* timed_out);
int tvk_exchange_destroy(tvk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TVK_CUDA_H */
