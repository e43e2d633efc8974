/*
 * tvk_renderer.h -- C handle API of the C++ host class tvk::CudaGridLeaper
 * (imagevis3d_b200/csrc/host/tvk_host.h), for callers that cannot include C++
 * (the Python mirror, ctypes tests, a Lua binding).  One function per
 * AbstrRenderer / Lua method the reference's scripts use; the Lua name each
 * replaces is given (registrations: Tuvok/Renderer/AbstrRenderer.cpp:1793-2267,
 * Tuvok/Renderer/RenderRegion.cpp:51-127).  Return 0 = success, non-zero =
 * failure (tvkr_last_error gives the T_ERROR-style message); predicates return
 * 0/1.
 */
#ifndef TVK_RENDERER_H
#define TVK_RENDERER_H

#include "tvk_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvkr_renderer tvkr_renderer;

typedef struct tvkr_dataset_info {
  uint32_t domain[3];
  uint32_t pool_capacity[3];
  uint32_t lod_count;          /* pool LODs = LargestSingleBrickLOD + 1 */
  uint32_t total_bricks;
  uint32_t tf_size;            /* 1-D transfer function size = 1-D histogram size (<= 4096) */
  double   range[2];           /* UVFDataset::ComputeRange */
  float    max_grad_magnitude;
  double   iso_value;
  uint32_t out_of_core;
} tvkr_dataset_info;

/* tuvok.renderer.new(...) + renderer.initialize(ctx) : MasterController::RequestNewVolumeRenderer
 * (MasterController.cpp:144-215), GLGridLeaper::Initialize (GLGridLeaper.cpp:247-316) */
tvkr_renderer* tvkr_create(int device);
/* renderer.cleanup() + deleteClass */
void tvkr_destroy(tvkr_renderer* r);
const char* tvkr_last_error(tvkr_renderer* r);

/* renderer.loadDataset(...) for an in-memory raw volume: runs the K2 builder
 * (the reference converts to UVF first: IOManager::ConvertDataset, IOManager.cpp:1061).
 * out_of_core: brick store in pinned host memory, pool sized by gpu_mem_bytes
 * (tuvok.state.gpuMem; 0 = the whole dataset fits). */
int tvkr_load_raw_volume(tvkr_renderer* r, const void* raw, const uint32_t size[3], uint32_t bytes_per_voxel,
                         const float scale[3], uint32_t brick_size, uint32_t overlap, int use_median,
                         int out_of_core, uint64_t gpu_mem_bytes, int compute_hist2d);
/* renderer.loadDataset("file.uvf") (AbstrRenderer::LoadFile, AbstrRenderer.cpp:215-263 -> UVFDataset::Open,
 * uvfDataset.cpp:83-216): a UVF v5 file with a TOC block; bricks are read (zlib / lz4 inflated) into pinned
 * host memory and paged from there.  gpu_mem_bytes as above. */
int tvkr_load_dataset(tvkr_renderer* r, const char* uvf_file, uint64_t gpu_mem_bytes, int verify_checksum);
/* the same, but a file whose bricks are all LZ4-compressed keeps them COMPRESSED in the pinned host store: paging moves the
 * compressed bytes over PCIe and the device inflates them (tvk_pool_upload_bricks_lz4); replaces the host-side inflate of
 * ExtendedOctree::GetBrickData (ExtendedOctree.cpp:313-360).  Other files load exactly as tvkr_load_dataset. */
int tvkr_load_dataset_compressed(tvkr_renderer* r, const char* uvf_file, uint64_t gpu_mem_bytes, int verify_checksum);
int tvkr_store_is_compressed(tvkr_renderer* r);
/* renderer.loadRebricked("file.uvf", {bs,bs,bs}, minmaxMode) (AbstrRenderer::LoadRebricked, AbstrRenderer.cpp:266-273 ->
 * IOManager::LoadRebrickedDataset, IOManager.cpp:1281-1320 -> DynamicBrickingDS): the file's bricks (e.g. 256^3) are
 * presented as bricks of brick_size (ghost voxels included; clamped to the file's brick size; its core must divide the
 * file's core).  minmax_mode: TVK_MM_SOURCE / TVK_MM_PRECOMPUTE / TVK_MM_DYNAMIC (tvk_cuda.h).  The target store is cut on
 * the device (tvk_rebrick) and paged like any loaded dataset. */
int tvkr_load_rebricked(tvkr_renderer* r, const char* uvf_file, uint32_t brick_size, int minmax_mode, uint64_t gpu_mem_bytes,
                        int verify_checksum);
/* what IOManager::ConvertDataset leaves on disk (RAWConverter::ConvertRAWDataset, RAWConverter.cpp:412-681):
 * the loaded volume as a UVF v5 file; compression = TVK_UVF_COMP_NONE / _ZLIB (tvk_uvf.h) */
int tvkr_save_uvf(tvkr_renderer* r, const char* uvf_file, uint32_t compression, uint32_t level);
int tvkr_get_dataset_info(tvkr_renderer* r, tvkr_dataset_info* out);

int tvkr_resize(tvkr_renderer* r, uint32_t w, uint32_t h);                      /* renderer.resize({w,h})       */
int tvkr_set_render_mode(tvkr_renderer* r, int mode);                           /* renderer.setRenderMode       */
int tvkr_set_lighting_enabled(tvkr_renderer* r, int on);                        /* renderer.setLightingEnabled  */
int tvkr_set_sample_rate_modifier(tvkr_renderer* r, float f);                   /* renderer.setSampleRateModifier */
int tvkr_set_fov(tvkr_renderer* r, float fov_deg);                              /* renderer.setFoV              */
int tvkr_set_view_pos(tvkr_renderer* r, const float p[3]);                      /* renderer.setViewPos          */
int tvkr_set_view_dir(tvkr_renderer* r, const float d[3]);                      /* renderer.setViewDir          */
int tvkr_set_up_dir(tvkr_renderer* r, const float u[3]);                        /* renderer.setUpDir            */
/* AbstrRenderer::SetViewParameters (AbstrRenderer.h:250-261, C++ only in the reference) */
int tvkr_set_view_parameters(tvkr_renderer* r, float angle, float znear, float zfar, const float eye[3],
                             const float ref[3], const float up[3]);
int tvkr_set_rotation4x4(tvkr_renderer* r, const float m[16]);                  /* region.setRotation4x4        */
int tvkr_set_translation4x4(tvkr_renderer* r, const float m[16]);               /* region.setTranslation4x4     */
int tvkr_set_iso_value(tvkr_renderer* r, float v);                              /* renderer.setIsoValue         */
int tvkr_set_iso_value_relative(tvkr_renderer* r, float v);                     /* renderer.setIsoValueRelative */
int tvkr_set_isosurface_color(tvkr_renderer* r, const float rgb[3]);            /* renderer.setIsosurfaceColor  */
/* AbstrRenderer::Set1DTrans (AbstrRenderer.h:189): RGBA8 table */
int tvkr_set_1d_trans(tvkr_renderer* r, const uint8_t* rgba, uint32_t n);
/* renderer.get1DTrans() proxy edit + tuvok.gpu.changed1DTrans */
int tvkr_set_1d_trans_float(tvkr_renderer* r, const float* rgba, uint32_t n);
int tvkr_load_1d_trans(tvkr_renderer* r, const char* filename_1dt);
/* 2-D TF as a rasterised RGBA8 table (w = value bins, h = 256 gradient bins) + its non-zero
 * limits (TransferFunction2D::GetNonZeroLimits); + tuvok.gpu.changed2DTrans */
int tvkr_set_2d_trans(tvkr_renderer* r, const uint8_t* rgba, uint32_t w, uint32_t h, const double nz_limits[4]);
/* A 2-D render region over the whole target (RenderRegion2D: windowMode TVK_WM_SAGITTAL / _AXIAL / _CORONAL, slice index in
 * voxels of the finest level, MIP flag, flipView): GLRenderer::Render2DView's slice view and MIP preview
 * (GLRenderer.cpp:1057-1176) of the finest single-brick LOD, classified with the 1-D transfer function (the 2-D one in 2-D
 * mode; MIP: 1-D colour only).  The result is the frame (tvkr_get_frame_*, tvkr_capture_single_frame); the next
 * tvkr_paint() starts the 3-D view again.  tvkr_get_slice_params: what the view would be drawn with (tests). */
int tvkr_render_2d_view(tvkr_renderer* r, int window_mode, uint64_t slice_index, int use_mip, int flip_x, int flip_y);
int tvkr_get_slice_params(tvkr_renderer* r, int window_mode, uint64_t slice_index, int use_mip, int flip_x, int flip_y,
                          tvk_slice_params* out);
/* The full-resolution MIP a 2-D region shows in capture mode (the rotating MIP of ImageVis3D's capture dialog;
 * GLRenderer::Render2DView :1182-1250, AbstrRenderer::PlanHQMIPFrame :1214-1245, GLRenderer::RenderHQMIPPreLoop :1256-1283,
 * GLRaycaster.cpp:481-529): every brick of the LOD PlanHQMIPFrame picks (mip_lod = m_bMIPLOD, default on in the reference)
 * is marched for its maximum under the rotation windowMode / flipView / angle_deg (= m_fMIPRotationAngle) define, with
 * the orthographic projection of m_bOrthoView (ortho != 0) or the camera's perspective one; the maximum is classified with
 * the 1-D transfer function's colour (opacity ignored).  The result is the frame.  tvkr_plan_hq_mip: what it is drawn with. */
int tvkr_render_hq_mip(tvkr_renderer* r, int window_mode, int flip_x, int flip_y, float angle_deg, int ortho, int mip_lod);
int tvkr_plan_hq_mip(tvkr_renderer* r, int window_mode, int flip_x, int flip_y, float angle_deg, int ortho, int mip_lod,
                     tvk_mip_params* p, tvk_classic_brick* bricks, uint32_t cap, uint32_t* n, uint32_t* lod);
/* renderer.load2DTrans("file.2dt") (TransferFunction2D::Load(file, size), TransferFunction2D.cpp:102-152): the swatch list is
 * rasterised on the host (the reference paints it with Qt's QPainter, :295-361; here: the same rules without Qt -- 1-D
 * function stretched over every row, polygons filled without anti-aliasing, linear / radial gradient, source-over) into a
 * table of (1-D table size) x 256 entries, then as tvkr_set_2d_trans */
int tvkr_load_2d_trans(tvkr_renderer* r, const char* filename_2dt);
/* the same without a renderer: rasterises file.2dt into a w x h RGBA8 table + its non-zero limits; save_as != NULL/"" also
 * writes the function back out with the file's own size (TransferFunction2D::Load(file) + Save, :156-220: the text round trip) */
int tvkr_tf2d_render(const char* filename_2dt, uint32_t w, uint32_t h, uint8_t* rgba, double nz_limits[4], const char* save_as);
/* tuvok.state.brickStrategy / hashTableSize / rehashCount (MasterController.cpp:458-496) + sampler */
int tvkr_set_state(tvkr_renderer* r, uint32_t brick_strategy, uint32_t hash_table_size, uint32_t rehash_count,
                   uint32_t sampler);

int tvkr_paint(tvkr_renderer* r);                                               /* renderer.paint               */
int tvkr_check_for_redraw(tvkr_renderer* r);                                    /* renderer.checkForRedraw      */
int tvkr_converged(tvkr_renderer* r);                                           /* renderer.converged           */
int tvkr_paint_until_converged(tvkr_renderer* r, uint32_t max_subframes, uint32_t* subframes);
int tvkr_clear_working_set(tvkr_renderer* r);                                   /* renderer.clearWorkingSet     */
int tvkr_recalc_visibility(tvkr_renderer* r);                                   /* renderer.recalcVisibility    */
/* renderer.captureSingleFrame(file, preserveTransparency): PNG */
int tvkr_capture_single_frame(tvkr_renderer* r, const char* filename, int preserve_transparency);
/* glReadPixels equivalents: RGBA8 top row first / premultiplied RGBA32F bottom row first */
int tvkr_get_frame_rgba8(tvkr_renderer* r, uint8_t* out);
int tvkr_get_frame_float(tvkr_renderer* r, float* out);

/* introspection for tests */
int tvkr_get_raycast_params(tvkr_renderer* r, tvk_raycast_params* out);
int tvkr_get_pool_slots(tvkr_renderer* r, int32_t* brick, uint64_t* time, uint32_t cap, uint32_t* n);
/* tuvok.renderer.new(type, ...): 0 = OpenGL_GridLeaper-like ray-guided path (default), 1 = OpenGL_Raycaster-like classic
 * per-brick path (MasterController::RequestNewVolumeRenderer, MasterController.cpp:144-215).  The classic path
 * renders the final sub-frame of a view (AbstrRenderer::PlanFrame's last LOD) per paint(), 1D-TF modes, in-core data. */
int tvkr_set_renderer_type(tvkr_renderer* r, int type);
/* the classic path's frame plan, for tests: parameters, front-to-back brick list (cap entries), chosen LOD */
int tvkr_plan_classic_frame(tvkr_renderer* r, tvk_classic_params* p, tvk_classic_brick* bricks, uint32_t cap,
                            uint32_t* n, uint32_t* lod);
/* getFirst3DRenderRegion().setClipPlane / enableClipPlane: world-space plane (a,b,c,d), n.p + d <= 0 is kept
 * (GLGridLeaper::FillBBoxVBO, GLGridLeaper.cpp:515-526 + Clipper::BoxPlane) */
int tvkr_set_clip_plane(tvkr_renderer* r, const float plane[4]);
int tvkr_enable_clip_plane(tvkr_renderer* r, int on);
/* ClearView (renderer.setClearViewEnabled / setCVIsoValue / setCVColor / setCVSize / setCVContextScale / setCVBorderScale /
 * setCVFocusPos; AbstrRenderer.cpp:1247-1256,1304-1361): the classic renderer's isosurface mode draws a second surface at the
 * ClearView isovalue (data units, default 1.5 x the isovalue) and shows it through a lens around the focus point (defaults:
 * colour (1,0,0), size 5.5, context scale 1, border scale 60, focus (0,0,0.5,1)).  tvkr_set_cv_focus_pos: window
 * coordinates of the mouse, y from the top -- the first surface's hit there becomes the focus (needs a ClearView frame;
 * GLRenderer::CVFocusHasChanged, GLRenderer.cpp:2833-2851); tvkr_set_cv_focus_pos_fvec: the world-space point itself.
 * The ray-guided renderer has no ClearView (GLGridLeaper::SupportsClearView is false); it ignores the switch. */
/* renderer.pick (GLRenderer::Pick, GLRenderer.cpp:2856-2872): isosurface mode (either renderer type), after a frame: the
 * eye-space position of the surface under window position (x, y from the top); 1 = no intersection / not available
 * (tvkr_last_error says which) */
int tvkr_pick(tvkr_renderer* r, int x, int y, float out[3]);
int tvkr_set_clearview(tvkr_renderer* r, int on);
int tvkr_get_clearview(tvkr_renderer* r);
int tvkr_set_cv_iso_value(tvkr_renderer* r, float iso);
int tvkr_get_cv_iso_value(tvkr_renderer* r, float* iso);
int tvkr_set_cv_color(tvkr_renderer* r, const float rgb[3]);
int tvkr_set_cv_size(tvkr_renderer* r, float size);
int tvkr_set_cv_context_scale(tvkr_renderer* r, float scale);
int tvkr_set_cv_border_scale(tvkr_renderer* r, float scale);
int tvkr_set_cv_focus_pos(tvkr_renderer* r, int x, int y);
int tvkr_set_cv_focus_pos_fvec(tvkr_renderer* r, const float pos[4]);
int tvkr_get_clearview_params(tvkr_renderer* r, int eye, tvk_clearview_params* out);
/* Stereo (renderer.setStereoEnabled / setStereoEyeDist / setStereoFocalLength / setStereoMode / setStereoEyeSwap /
 * initStereoFrame / toggleStereoFrame; AbstrRenderer.cpp:1387-1410,1577-1585): every paint() marches the left and the right
 * eye (views and off-axis projections of FLOATMATRIX4::BuildStereoLookAtAndProjection, Vectors.h:1215-1248), both report
 * into one request table, and the converged frame is the composition GLRenderer::EndFrame draws (GLRenderer.cpp:757-812):
 * mode TVK_SM_RB anaglyph (default), _SCANLINE, _SBS, _AF.  Defaults: eye distance 0.02, focal length 1.  Both renderer
 * types (the classic one draws eye 0's brick set in each eye's own front-to-back order); not for sort-last members.  tvkr_get_raycast_params_eye: the parameters of eye 0 (left) / 1 (right).
 * tvkr_stereo_matrices: the four matrices (left view, right view, left projection, right projection; 16 floats each, as
 * the reference stores them), no renderer needed. */
int tvkr_set_stereo(tvkr_renderer* r, int on);
int tvkr_get_stereo(tvkr_renderer* r);
int tvkr_set_stereo_eye_dist(tvkr_renderer* r, float dist);
int tvkr_set_stereo_focal_length(tvkr_renderer* r, float focal_length);
int tvkr_set_stereo_mode(tvkr_renderer* r, int mode);
int tvkr_set_stereo_eye_swap(tvkr_renderer* r, int swap);
int tvkr_init_stereo_frame(tvkr_renderer* r);
int tvkr_toggle_stereo_frame(tvkr_renderer* r);
int tvkr_get_raycast_params_eye(tvkr_renderer* r, int eye, tvk_raycast_params* out);
/* the classic path's frame plan of one eye of a stereo pair (tvkr_plan_classic_frame = eye 0): the same bricks as eye 0,
 * sorted front to back for this eye (AbstrRenderer::BuildLeftEyeSubFrameBrickList, AbstrRenderer.cpp:843-858) */
int tvkr_plan_classic_frame_eye(tvkr_renderer* r, int eye, tvk_classic_params* p, tvk_classic_brick* bricks, uint32_t cap,
                                uint32_t* n, uint32_t* lod);
int tvkr_get_eye_frame_float(tvkr_renderer* r, int eye, float* out);   /* one eye's own picture, premultiplied RGBA32F */
int tvkr_stereo_matrices(const float eye[3], const float at[3], const float up[3], float fovy, float aspect, float znear, float zfar,
                         float focal_length, float eye_dist, float out[64]);
int tvkr_set_light_colors(tvkr_renderer* r, const float ambient[4], const float diffuse[4], const float specular[4],
                          const float dir[3]);                                    /* renderer.setLightColors   */
int tvkr_get_view(tvkr_renderer* r, float pos[3], float dir[3], float up[3]);    /* getViewPos/Dir/UpDir      */
int tvkr_reset_view(tvkr_renderer* r);                                           /* resetViewPos/Dir/UpDir    */
int tvkr_set_timestep(tvkr_renderer* r, uint64_t t);                             /* renderer.setTimestep      */
/* dataset.getNumberOfTimesteps (UVFDataset::GetNumberOfTimesteps, uvfDataset.cpp:621-623) and the step being shown; a
 * dataset loaded from memory has one */
int tvkr_get_timesteps(tvkr_renderer* r, uint64_t* count, uint64_t* current);
int tvkr_set_lod_limits(tvkr_renderer* r, uint32_t lo, uint32_t hi);             /* renderer.setLODLimits     */
int tvkr_set_bg_colors(tvkr_renderer* r, const float top[3], const float bottom[3]);   /* renderer.setBGColors */
int tvkr_set_blend_precision(tvkr_renderer* r, int precision);   /* renderer.setBlendPrecision: accumulation is always fp32 here */
int tvkr_get_perf(tvkr_renderer* r, double out[6]);     /* tuvok.perf: subframes, raycast ms, paged bricks, uploaded bytes, samples, frames */
/* tuvok.perf(id) (MasterController::PerfQuery, MasterController.cpp:305-310, registered :502): id = enum PerfCounter of
 * Basics/PerfCounter.h:7-44 (0 PERF_SUBFRAMES, 1 PERF_RENDER, 2 PERF_RAYCAST, 3 PERF_READ_HTABLE, 6 PERF_UPLOAD_BRICKS,
 * 7 PERF_POOL_SORT, 8 PERF_POOL_UPLOADED_MEM, 9 PERF_POOL_GET_BRICK, 19 PERF_POOL_UPLOAD_BRICK,
 * 21 PERF_POOL_UPLOAD_METADATA, 22 PERF_EO_BRICKS ...): the amount since the last query of that counter, which is reset */
double tvkr_perf_query(tvkr_renderer* r, int counter);
/* ---- raw-volume front ends: QVIS .dat/.raw (IO/QVISConverter.cpp) and NRRD .nrrd/.nhdr, raw or gzip encoded
 * (IO/NRRDConverter.cpp).  tvk_raw_probe fills what AbstrConverter::ConvertToRAW returns; tvk_raw_read delivers the voxels
 * in host byte order; tvkr_load_volume_file = probe + read + quantisation on the device (RAWConverter::ConvertRAWDataset's
 * dispatcher, tvk_quantize) + octree build, i.e. what "open a .dat / .nrrd file" does in ImageVis3D.  No device needed for
 * the first two. */
#define TVK_RAW_ENC_RAW  0
#define TVK_RAW_ENC_GZIP 1
typedef struct tvk_raw_desc {
  uint64_t header_skip;          /* iHeaderSkip: bytes before the voxels in data_file                              */
  uint64_t inflated_skip;        /* gzip encoding: bytes to drop at the start of the INFLATED stream ("byte skip")   */
  uint32_t component_bits;       /* iComponentSize: 8, 16, 32, 64                                                  */
  uint32_t component_count;      /* iComponentCount                                                                */
  int32_t  convert_endianness;   /* bConvertEndianess: the file's byte order is not the host's                     */
  int32_t  is_signed, is_float;
  uint32_t encoding;             /* TVK_RAW_ENC_*                                                                  */
  uint64_t size[3];              /* vVolumeSize                                                                    */
  float    aspect[3];            /* vVolumeAspect                                                                  */
  char     title[64];
  char     data_file[1024];      /* strIntermediateFile: where the voxels are                                      */
} tvk_raw_desc;
/* ---- sort-last over several GPUs: what `tuvok.state.gpuCount` + a compositor setting stand for.  Every GPU (or
 * every box, when a GPU renders several) runs one renderer = one MEMBER; the dataset it loaded is one box of a global volume:
 * voxels [start, start + domain) of it, of which the member owns [lo, hi) (the rest is halo for the filter, the gradients and
 * the coarser LODs).  After tvkr_sortlast_configure the camera, LOD selection and TF scaling refer to the global volume and
 * rays are clipped to the owned box.  Exchange = fused direct send (tvk_exchange_*): _exchange_create returns this member's
 * 64-byte CUDA IPC handle, _exchange_connect takes all members' handles and / or base pointers already valid in this process
 * (bases[i] != NULL wins; a sibling member's base comes from _exchange_base), exactly as tvk_exchange_connect.
 * Per displayed frame, on every member: _begin_frame; paint() until converged() -- each sub-frame's pixels travel to the stripe
 * owners while it marches; when the APPLICATION knows every member has converged: _composite(order) with the front-to-back
 * order _depth_order gives for the current view (boxes: members x {lo xyz, hi xyz}); member 0 reads the frame. */
int tvkr_sortlast_configure(tvkr_renderer* r, uint32_t members, uint32_t member, const uint32_t global_size[3], const uint32_t lo[3],
                            const uint32_t hi[3], const uint32_t start[3], double global_max_value);
int tvkr_sortlast_exchange_create(tvkr_renderer* r, void* handle64);
int tvkr_sortlast_exchange_connect(tvkr_renderer* r, const void* handles, void* const* bases);
int tvkr_sortlast_exchange_base(tvkr_renderer* r, uint32_t member, void** base);
int tvkr_sortlast_begin_frame(tvkr_renderer* r);
int tvkr_sortlast_composite(tvkr_renderer* r, const uint32_t* order);
int tvkr_sortlast_depth_order(tvkr_renderer* r, const uint32_t* boxes_lo_hi, uint32_t* order);
int tvkr_sortlast_read_frame_rgba8(tvkr_renderer* r, uint8_t* rgba8, int flip_rows);

/* Volumes larger than device memory (configs[4]): the flat volume is fed in z-slabs of whole LOD-0 brick rows (DEVICE
 * pointers; tvkr_load_streamed_rows tells which raw rows a slab must hold), the streamed builder cuts the bricks straight into
 * the pinned host store the pager reads from.  _end finishes the coarser levels, creates the pool (gpu_mem_bytes budget) and
 * reports the builder's kernel time.  No 2-D histogram (it needs the flat volume). */
int tvkr_load_streamed_begin(tvkr_renderer* r, const uint32_t size[3], uint32_t bytes_per_voxel, const float scale[3], uint32_t brick_size,
                             uint32_t overlap, int median);
int tvkr_load_streamed_rows(tvkr_renderer* r, uint32_t bz0, uint32_t bz1, uint32_t* z0, uint32_t* rows);
int tvkr_load_streamed_slab(tvkr_renderer* r, const void* device_slab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1);
int tvkr_load_streamed_end(tvkr_renderer* r, uint64_t gpu_mem_bytes, float* builder_kernel_ms);
int tvk_raw_probe(const char* path, tvk_raw_desc* out, char* err, size_t errlen);
int tvk_raw_read(const tvk_raw_desc* d, void* dst, uint64_t dst_bytes, char* err, size_t errlen);
int tvkr_load_volume_file(tvkr_renderer* r, const char* path, uint32_t brick_size, uint32_t overlap, int median, int out_of_core,
                          uint64_t gpu_mem_bytes, int compute_hist2d);

/* The coarse-to-fine sub-frame ladder of AbstrRenderer (classic renderer): which LOD a view starts at
 * (ComputeMaxLODForCurrentView, AbstrRenderer.cpp:681-787, fed by the times the first two sub-frames of the previous view
 * took) and how PlanFrame steps down to the LOD the view needs (:1125-1212).  The state is a POD so that the logic can be
 * driven -- and checked against the reference's statements -- without a device. */
typedef struct tvkr_lod_state {
  int32_t  target;               /* ERendererTarget: 0 interactive, 1 capture, 2 headless                              */
  float    msec_passed[2];       /* time of the first / second sub-frame of the last view, -1 = none (RestartTimer)     */
  float    max_ms_per_frame;     /* m_fMaxMSPerFrame (default 10000)                                                   */
  uint32_t lod_not_ok_counter;   /* m_iLODNotOKCounter                                                                 */
  uint64_t perf_lod_skip;        /* m_iPerformanceBasedLODSkip                                                         */
  int32_t  low_res_intermediate; /* m_bRenderLowResIntermediateResults ("use all means")                               */
  int32_t  decrease_sampling_rate, decrease_screen_res;
  uint64_t brick_list_size, bricks_rendered;   /* m_vCurrentBrickList.size(), m_iBricksRenderedInThisSubFrame            */
  uint64_t max_lod_index;        /* m_iMaxLODIndex = LOD level count - 1                                               */
  uint64_t min_lod_for_view;     /* m_iMinLODForCurrentView                                                            */
  uint64_t lod_limit_x;          /* m_iLODLimits.x                                                                     */
  uint64_t start_lod_offset, current_lod_offset;   /* out: m_iStartLODOffset, m_iCurrentLODOffset                        */
} tvkr_lod_state;
int tvkr_compute_max_lod_for_view(tvkr_lod_state* s);
/* classic renderer: frame-time budget of the ladder, a time to use instead of the measured kernel time of the next
 * sub-frames (ms < 0: measure again; for deterministic tests), and the ladder's state */
int tvkr_set_renderer_target(tvkr_renderer* r, int target);   /* AbstrRenderer::SetRendererTarget: 0 interactive, 1 capture, 2 headless */
int tvkr_set_max_ms_per_frame(tvkr_renderer* r, float ms);
int tvkr_override_subframe_ms(tvkr_renderer* r, float ms);
int tvkr_get_lod_state(tvkr_renderer* r, tvkr_lod_state* out, uint32_t* current_lod);

/* TransferFunction1D::Resample (IO/TransferFunction1D.cpp:198-231): n RGBA float entries -> target entries (adaptive box
 * filter down, linear interpolation up); no device needed. */
int tvkr_tf1d_resample(const float* rgba, uint32_t n, uint32_t target, float* out);
/* The 1-D histogram as the transfer-function editor gets it (UVFDataset::GetHistograms, IO/uvfDataset.cpp:747-800): the
 * file's bins weighted into `target` bins; the last bin, which the reference's loop never writes, is 0. */
int tvkr_resample_hist1d(const uint64_t* bins, uint64_t n, uint32_t target, uint32_t* out);
/* Dataset::Get1DHistogram of the loaded dataset: *n = min(file bins, 4096, 2^bits); out == NULL only queries n. */
int tvkr_get_histogram1d(tvkr_renderer* r, uint32_t* out, uint32_t cap, uint32_t* n);
/* GPUMemMan::GetVolumePool's pool sizing (GPUMemMan.cpp:766-830) on its own, no device needed: slots per axis of the
 * brick pool for bricks of max_brick voxels, total_bricks bricks in the dataset, gpu_mem_bytes of pool memory and a
 * 3-D texture limit of max_3d_texture texels per axis. */
int tvkr_compute_pool_capacity(const uint32_t max_brick[3], uint32_t bytes_per_voxel, uint64_t total_bricks, uint64_t gpu_mem_bytes,
                               uint32_t max_3d_texture, uint32_t capacity[3]);
tvk_ctx* tvkr_context(tvkr_renderer* r);

#ifdef __cplusplus
}
#endif
#endif /* TVK_RENDERER_H */
