// C++ host of the B200 ray-caster: the AbstrRenderer-shaped class a Tuvok build
// registers as a new EVolumeRendererType.  It owns no CUDA
// code: everything device-side goes through the C ABI of include/tvk_cuda.h.
//
// Reference counterparts (paths relative to /root/reference/Tuvok/):
//   Matrix4 / Vec3          Basics/Vectors.h (row-vector convention, BuildLookAt :1250,
//                           Perspective :1267, inverse :1042)
//   TransferFunction1D      IO/TransferFunction1D.{h,cpp}
//   BrickedVolume           what GLGridLeaper needs from LinearIndexDataset / UVFDataset
//                           (IO/LinearIndexDataset.h:14-27, IO/uvfDataset.cpp:590-596,1120-1199)
//   BrickPool               Renderer/GL/GLVolumePool.{h,cpp} bookkeeping (slots, LRU,
//                           UploadBricks, RecomputeVisibility) + GPUMemMan::GetVolumePool sizing
//   CudaGridLeaper          Renderer/AbstrRenderer.{h,cpp} state + Renderer/GL/GLGridLeaper.cpp
//                           frame loop (Render3DRegion :914-1154) + GLRenderer camera set-up
#pragma once
#include <cstdint>
#include <array>
#include <string>
#include <vector>
#include "tvk_cuda.h"
#include "tvk_renderer.h"
#include <cstring>

namespace tvk {

struct Vec3 {
  float x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator*(float s) const { return Vec3(x * s, y * s, z * s); }
  float dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
  Vec3 cross(const Vec3& o) const { return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
  float length() const;
  Vec3 normalized() const;
};

// FLOATMATRIX4: memory order m11..m44, v' = v*M
struct Matrix4 {
  float m[16];
  Matrix4();                                     // identity
  Matrix4 operator*(const Matrix4& o) const;
  Matrix4 inverse() const;
  Matrix4 transposed() const;
  static Matrix4 Translation(float x, float y, float z);
  static Matrix4 Scaling(float x, float y, float z);
  static Matrix4 RotationX(double a);
  static Matrix4 RotationY(double a);
  static Matrix4 RotationZ(double a);
  static Matrix4 LookAt(const Vec3& eye, const Vec3& at, const Vec3& up);
  static Matrix4 Perspective(float fovy, float aspect, float n, float f);
  static Matrix4 Ortho(float left, float right, float bottom, float top, float n, float f);
  static Matrix4 PerspectiveOffCenter(float left, float right, float bottom, float top, float n, float f);
  // FLOATMATRIX4::BuildStereoLookAtAndProjection (Vectors.h:1215-1248): [0] = left eye, [1] = right eye
  static void StereoLookAtAndProjection(const Vec3& eye, const Vec3& at, const Vec3& up, float fovy, float aspect, float n, float f,
                                        float focalLength, float eyeDist, Matrix4 view[2], Matrix4 proj[2]);
  void transform(const float v[4], float out[4]) const;   // v*M
};

class TransferFunction1D {
 public:
  explicit TransferFunction1D(size_t n = 0) : m_color(4 * n, 0.0f) { ComputeNonZeroLimits(); }
  size_t GetSize() const { return m_color.size() / 4; }
  void Resize(size_t n) { m_color.resize(4 * n, 0.0f); ComputeNonZeroLimits(); }
  void Resample(size_t n);                                 // TransferFunction1D::Resample (box filter down, linear up)
  bool Load(const std::string& filename);                  // ".1dt" text format
  bool Save(const std::string& filename) const;
  bool Load(std::istream& in);                             // TransferFunction1D.cpp:279-295 (the 1-D part of a ".2dt" file)
  bool Save(std::ostream& out) const;                      // :298-309, byte for byte
  void Set(const float* rgba, size_t n);
  void SetStdFunction(float center = 0.5f, float invGradient = 0.5f);
  void GetByteArray(std::vector<uint8_t>& out, uint8_t usedRange = 255) const;
  void ComputeNonZeroLimits();
  uint64_t NonZeroMin() const { return m_lo; }
  uint64_t NonZeroMax() const { return m_hi; }
  std::vector<float>& ColorData() { return m_color; }

 private:
  void SetStdFunction(float center, float invGradient, int component);
  std::vector<float> m_color;     // r g b a per entry
  uint64_t m_lo = 0, m_hi = 0;
};

// One swatch of a 2-D transfer function (TFPolygon, TransferFunction2D.h:52-66): a polygon in normalised (value,
// gradient) coordinates filled with a linear or radial colour gradient.
struct TFPolygon {
  bool radial = false;
  std::vector<std::array<float, 2>> points;
  float gradientCoords[2][2] = {{0, 0}, {0, 0}};
  std::vector<std::pair<float, std::array<float, 4>>> stops;
};

// TransferFunction2D (IO/TransferFunction2D.cpp): the ".2dt" file (size, a 1-D function, the swatch list) with the
// reference's Load / Save, and a rasteriser for the table the GPU samples.  The reference rasterises with Qt's QPainter
// (RenderTransferFunction8Bit, :295-361: the 1-D function stretched over every row, then each swatch as a polygon without
// pen or anti-aliasing, QLinearGradient / QRadialGradient brush, source-over); Qt is not available here, so Render()
// restates those rules (pixel-centre coverage, odd-even fill, pad spread, stops interpolated linearly) -- equal in
// structure, not guaranteed equal in the last bit to what a particular Qt version paints.
class TransferFunction2D {
 public:
  bool Load(const std::string& filename);                                  // :156-198 (size from the file)
  bool Load(const std::string& filename, size_t targetW, size_t targetH);  // :102-152 (1-D part resampled to the target)
  bool Save(const std::string& filename) const;                            // :200-220, byte for byte
  // GetByteArray (:222-236, RGBA8 rows) + ComputeNonZeroLimits (:376-397: min x, max x, min y, max y of alpha != 0)
  void Render(std::vector<uint8_t>& rgba, double limits[4]) const;
  size_t Width() const { return m_w; }
  size_t Height() const { return m_h; }
  TransferFunction1D& Trans1D() { return m_trans1d; }
  std::vector<TFPolygon>& Swatches() { return m_swatches; }

 private:
  bool LoadRest(std::istream& in, bool allowNoSwatch);
  size_t m_w = 0, m_h = 0;
  TransferFunction1D m_trans1d;
  std::vector<TFPolygon> m_swatches;
};

// The 1-D histogram the transfer-function editor draws: the file's histogram, "resampled" to at most
// MAX_TRANSFERFUNCTION_SIZE (4096) bins when it is larger (UVFDataset::GetHistograms, uvfDataset.cpp:747-800).
// The reference's loop stops before it writes the last bin (which stays an uninitialised new[] value there): it is 0 here.
void ResampleHistogram1D(const std::vector<uint64_t>& in, size_t target, std::vector<uint32_t>& out);

// A bricked multi-resolution volume in ExtendedOctree layout (ToC order).
// Brick payloads live EITHER in the device brick store left by tvk_build_octree
// (in-core) OR in one pinned host allocation (out-of-core paging source).
struct BrickedVolume {
  uint32_t domain[3] = {0, 0, 0};
  float    scale[3] = {1, 1, 1};
  uint32_t bpv = 1, brick_size = 64, overlap = 2;
  tvk_build_info info{};
  std::vector<uint64_t> offsets;        // bytes, per brick
  std::vector<uint32_t> sizes;          // 3 per brick
  std::vector<double>   minmax;         // 2 per brick
  std::vector<uint64_t> hist1d;
  float    max_grad_magnitude = 0.0f;   // 0 = unknown (2-D histogram not computed)
  std::vector<uint64_t> hist2d;         // [hist2d_x][hist2d_y] when the file (or the builder) provided one
  uint64_t hist2d_x = 0, hist2d_y = 0;
  void*    classic_store = nullptr;     // device copy of host_store, made on demand for the classic ray-caster
  // host_store holds the bricks STILL LZ4-COMPRESSED (as the file stores them): offsets[] index the compressed store,
  // comp_len[] are their lengths, and the pager inflates on the device (tvk_pool_upload_bricks_lz4)
  bool     lz4_store = false;
  uint32_t store_codec = 0;                // TVK_CODEC_LZ4 / TVK_CODEC_ZLIB when the store holds the compressed bricks of the file
  std::vector<uint32_t> comp_len;
  uint8_t* host_store = nullptr;        // pinned (cudaHostAlloc) when out of core
  void*    device_store = nullptr;      // owned by the tvk_ctx build state when in core
  uint64_t store_bytes = 0;

  uint32_t LargestSingleBrickLOD() const;                 // uvfDataset.cpp:590-596
  void     MaxUsedBrickSizes(uint32_t out[3]) const;      // GetMaxUsedBrickSizes
  uint64_t TocIndex(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const;
  void     Range(double& lo, double& hi) const;           // ComputeRange, uvfDataset.cpp:1120-1181
};

struct PoolSlot {                        // PoolSlotData, GLVolumePool.h:25-60
  int32_t  brick = -1;
  uint64_t time = 0, orig_time = 0;
  uint32_t pos = 0;                      // linear slot index = position in pool
  bool WasEverUsed() const { return brick != -1; }
  bool ContainsVisibleBrick() const { return time > 1; }
  void FlagEmpty() { orig_time = time; time = 1; }
  void Restore() { time = orig_time; }
};

class BrickPool {
 public:
  BrickPool(tvk_ctx* ctx, const BrickedVolume* vol) : m_ctx(ctx), m_vol(vol) {}
  // GPUMemMan::GetVolumePool (GPUMemMan.cpp:766-844); max3d = GL_MAX_3D_TEXTURE_SIZE stand-in
  static void ComputePoolCapacity(const BrickedVolume& v, uint64_t gpuMemBytes, uint32_t max3d, uint32_t cap[3]);
  static void ComputePoolCapacity(const uint32_t maxBrick[3], uint32_t bpv, uint64_t totalBricks, uint64_t gpuMemBytes, uint32_t max3d,
                                  uint32_t cap[3]);
  int  Create(const uint32_t capacity[3], uint32_t sampler);
  int  UploadFirstBrick(bool wait = true);                                     // :763-766
  // wait = false: the upload and the page-table patch are only ORDERED before later work on the render stream (the
  // next sub-frame), the host does not block on the PCIe transfer (sources live in the pinned brick store)
  int  UploadBricks(const std::vector<uint32_t>& xyzl, uint32_t& paged, bool wait = true);   // :1720-1789
  int  Reset();                                                  // PH_Reset (:261-284): empty pool, coarsest brick pinned again
  int  RecomputeVisibility(const tvk_visibility& vis, uint32_t counts[4]); // :1580-1718
  bool IsBrickResident(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const;
  uint32_t BrickIndex(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const;   // GetIntegerBrickID
  void     BrickCoords(uint32_t index, uint32_t xyzl[4]) const;                  // GetVectorBrickID
  uint32_t LodCount() const { return m_lodCount; }
  uint32_t TotalBricks() const { return m_total; }
  const std::vector<PoolSlot>& Slots() const { return m_slots; }
  const uint32_t* Capacity() const { return m_cap; }
  const uint32_t* MaxBrick() const { return m_maxBrick; }
  uint64_t UploadedBytes() const { return m_uploadedBytes; }
  // host milliseconds of the last UploadBricks: slot sort, source look-up, transfer issue, page-table patch
  double LastSortMs() const { return m_lastMs[0]; }
  double LastGetBrickMs() const { return m_lastMs[1]; }
  double LastUploadMs() const { return m_lastMs[2]; }
  double LastMetadataMs() const { return m_lastMs[3]; }

 private:
  bool ContainsData(const tvk_visibility& vis, uint32_t index) const;
  void PrepareForPaging();                                        // :955-959
  tvk_ctx* m_ctx;
  const BrickedVolume* m_vol;
  uint32_t m_cap[3] = {0, 0, 0}, m_maxBrick[3] = {0, 0, 0};
  uint32_t m_lodCount = 0, m_total = 0;
  std::vector<uint32_t> m_lodOffset;
  std::vector<uint32_t> m_lodBricks;     // 3 per lod
  std::vector<double>   m_minmax;        // pool index order
  std::vector<PoolSlot> m_slots;
  uint64_t m_time = 2;                   // m_iTimeOfCreation
  size_t   m_insertPos = 0;
  // LRU order = the ring m_slots[m_head .. m_movable) ++ m_slots[0 .. m_head), then the pinned tail (PrepareForPaging)
  enum { ORDER_SORTED = 0, ORDER_INSERTED = 1, ORDER_TOUCHED = 2 };
  int      m_order = ORDER_TOUCHED;      // what happened to the slots' times since the order was last established
  size_t   m_head = 0, m_movable = 0;
  PoolSlot& SlotAt(size_t i) { return m_slots[i < m_movable ? (m_head + i) % m_movable : i]; }
  uint64_t m_uploadedBytes = 0;
  double   m_lastMs[4] = {0, 0, 0, 0};
};

// host-level choice: the layered-array pool (TVK_SAMPLER_GATHER)
static const uint32_t kSamplerAuto = 0xFFFFFFFFu;

enum ERenderMode { RM_1DTRANS = 0, RM_2DTRANS = 1, RM_ISOSURFACE = 2 };
enum ERendererTarget { RT_INTERACTIVE = 0, RT_CAPTURE = 1, RT_HEADLESS = 2 };

struct PerfCounters {                     // running totals (tvkr_get_perf)
  double subframes = 0, raycast_ms = 0, paged_bricks = 0, uploaded_bytes = 0, samples = 0, frames = 0;
};

// enum PerfCounter of the reference, same order and values (Basics/PerfCounter.h:7-44): what `tuvok.perf(id)` takes.
// PerfQuery returns the amount accumulated since the last query of that counter and resets it
// (MasterController::PerfQuery, MasterController.cpp:305-310).  Filled here: SUBFRAMES, RENDER, RAYCAST, READ_HTABLE,
// CONDENSE_HTABLE, UPLOAD_BRICKS, POOL_SORT, POOL_UPLOADED_MEM, POOL_GET_BRICK (host time of resolving the brick sources),
// POOL_UPLOAD_BRICK (issue of the transfers), POOL_UPLOAD_METADATA, EO_BRICKS; the DY_* / disk counters stay 0.
enum EPerfCounter {
  PERF_SUBFRAMES = 0, PERF_RENDER, PERF_RAYCAST, PERF_READ_HTABLE, PERF_CONDENSE_HTABLE, PERF_SORT_HTABLE, PERF_UPLOAD_BRICKS,
  PERF_POOL_SORT, PERF_POOL_UPLOADED_MEM, PERF_POOL_GET_BRICK, PERF_DY_GET_BRICK, PERF_DY_CACHE_LOOKUPS, PERF_DY_CACHE_LOOKUP,
  PERF_DY_RESERVE_BRICK, PERF_DY_LOAD_BRICK, PERF_DY_CACHE_ADDS, PERF_DY_CACHE_ADD, PERF_DY_BRICK_COPIED, PERF_DY_BRICK_COPY,
  PERF_POOL_UPLOAD_BRICK, PERF_POOL_UPLOAD_TEXEL, PERF_POOL_UPLOAD_METADATA, PERF_EO_BRICKS, PERF_EO_DISK_READ,
  PERF_EO_DECOMPRESSION, PERF_MM_PRECOMPUTE, PERF_SOMETHING, PERF_END
};

class CudaGridLeaper {
 public:
  CudaGridLeaper();
  ~CudaGridLeaper();

  // --- life cycle (AbstrRenderer::LoadFile / Initialize / Resize / Cleanup)
  bool Initialize(int device);
  // raw x-fastest volume in host memory -> K2 builder -> bricked volume.  out_of_core moves the
  // brick store to pinned host memory and sizes the pool by gpuMemBytes (0 = whole dataset).
  bool LoadRawVolume(const void* raw, const uint32_t size[3], uint32_t bytesPerVoxel, const float scale[3],
                     uint32_t brickSize, uint32_t overlap, bool median, bool outOfCore, uint64_t gpuMemBytes,
                     bool computeHist2D, bool rawOnDevice = false);
  // ---- sort-last over several GPUs (`tuvok.state.gpuCount` / compositor in MasterController terms): the loaded
  // volume is ONE BOX of a larger global volume -- voxels [start, start + domain) of it, of which this member owns [lo, hi)
  // (the rest is halo) -- and the member is one of `members` exchange members (one per box; several per process are fine).
  // The camera, the LOD selection and the transfer-function scaling then refer to the GLOBAL volume, rays are clipped to the
  // owned box, every Paint() sends its pixels to the scan-line stripe owners (fused direct send, tvk_exchange_*), and
  // SortLastComposite blends this member's stripe once every member has converged.
  bool SortLastConfigure(uint32_t members, uint32_t member, const uint32_t globalSize[3], const uint32_t lo[3], const uint32_t hi[3],
                         const uint32_t start[3], double globalMaxValue);
  bool SortLastExchangeCreate(void* handle64);                         // after Resize: tvk_exchange_create
  bool SortLastExchangeConnect(const void* handles, void* const* bases);
  bool SortLastBeginFrame();                                           // once per displayed frame, on every member
  bool SortLastComposite(const uint32_t* order);                       // after all members have converged
  void SortLastDepthOrder(const uint32_t* boxesLoHi, uint32_t* order); // members x 6 (lo xyz, hi xyz) -> front-to-back order
  bool SortLastReadFrameRGBA8(uint8_t* out, bool flipRows);            // member 0: the finished frame
  bool SortLastActive() const { return m_sl.on; }
  // Volumes larger than device memory: the flat volume arrives in z-slabs of whole brick rows (device pointers), the streamed
  // builder (tvk_build_stream_*) cuts the bricks STRAIGHT INTO THE PINNED HOST STORE the pager reads from; nothing but one
  // slab, the coarser flat levels (1/7 of the raw size) and the pool is ever resident.  Begin -> Rows/Slab... -> End.
  bool LoadStreamedBegin(const uint32_t size[3], uint32_t bytesPerVoxel, const float scale[3], uint32_t brickSize, uint32_t overlap,
                         bool median);
  bool LoadStreamedRows(uint32_t bz0, uint32_t bz1, uint32_t& z0, uint32_t& rows) const;
  bool LoadStreamedSlab(const void* deviceSlab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1);
  bool LoadStreamedEnd(uint64_t gpuMemBytes);
  float LastBuildMs() const { return m_lastBuildMs; }
  // "open a .dat / .nrrd file": raw front end (tvk_raw_probe / tvk_raw_read) -> quantisation on the device (tvk_quantize)
  // -> LoadRawVolume; single-component volumes of any of the ten element types
  bool LoadVolumeFile(const std::string& path, uint32_t brickSize, uint32_t overlap, bool median, bool outOfCore, uint64_t gpuMemBytes,
                      bool computeHist2D);
  // renderer.loadDataset(file.uvf): a UVF with a TOC block (8/16-bit unsigned); bricks go to pinned host memory
  // keepCompressed: a file whose bricks are all LZ4 keeps them compressed in the pinned store; they cross PCIe compressed
  // and are inflated on the device (ray-guided renderer only; any other file loads as usual)
  bool LoadDataset(const std::string& uvfFile, uint64_t gpuMemBytes, bool verifyChecksum, bool keepCompressed = false);
  // A bricked volume some OTHER reader produced (the AbstrRenderer adapter walks a tuvok LinearIndexDataset once):
  // the caller fills domain / scale / bpv / brick_size / overlap and, per brick in ToC order, sizes (3 each) and
  // minmax (2 each), plus hist1d and max_grad_magnitude; PrepareHostVolume plans the octree, lays the bricks out in ONE
  // pinned host allocation (host_store + offsets); the caller copies the brick payloads there; AdoptVolume takes
  // ownership and sets up pool, transfer function and visibility like LoadDataset does.
  // renderer.loadRebricked(file, {bs,bs,bs}, minmaxMode) (AbstrRenderer.cpp:266-273): brickSize includes the ghost voxels;
  // minmaxMode = TVK_MM_SOURCE / _PRECOMPUTE / _DYNAMIC
  // a 2-D render region over the whole target: slice view or MIP preview (GLRenderer::Render2DView)
  bool Render2DView(int windowMode, uint64_t sliceIndex, bool useMIP, bool flipX, bool flipY);
  bool ComputeSliceParams(int windowMode, uint64_t sliceIndex, bool useMIP, bool flipX, bool flipY, tvk_slice_params& p, uint64_t& toc) const;
  // renderer.load2DTrans(file.2dt): Load with the renderer's table size + rasterisation + Set2DTrans
  bool Load2DTrans(const std::string& filename);
  bool LoadRebricked(const std::string& uvfFile, uint32_t brickSize, int minmaxMode, uint64_t gpuMemBytes, bool verifyChecksum);
  bool PrepareHostVolume(BrickedVolume* v);
  bool AdoptVolume(BrickedVolume* v, uint64_t gpuMemBytes);
  // the loaded volume as a UVF v5 file (TOC, histograms, max/min, key/values, MD5); compression: TVK_UVF_COMP_*
  bool SaveUVF(const std::string& uvfFile, uint32_t compression, uint32_t level);
  void Resize(uint32_t w, uint32_t h);
  void Cleanup();
  const char* LastError() const { return m_err.c_str(); }

  // --- state (names of AbstrRenderer.h; Lua: setRenderMode, setViewPos, ...)
  bool SetRendermode(int mode);
  int  GetRendermode() const { return m_mode; }
  void SetUseLighting(bool b);
  void SetSampleRateModifier(float f);
  void SetFoV(float f);
  void SetViewPos(const Vec3& p);
  void SetViewDir(const Vec3& d);
  void SetUpDir(const Vec3& u);
  void ResetViewPos(); void ResetViewDir(); void ResetUpDir();
  void SetViewParameters(float angle, float znear, float zfar, const Vec3& eye, const Vec3& ref, const Vec3& up);
  void SetRotation4x4(const Matrix4& m);        // first 3-D render region (RenderRegion.cpp:51-127)
  void SetTranslation4x4(const Matrix4& m);
  void SetIsoValue(float v);
  void SetIsoValueRelative(float v);
  double GetIsoValue() const { return m_isoValue; }
  void SetIsosurfaceColor(const Vec3& c);
  void SetLightColors(const float ambient[4], const float diffuse[4], const float specular[4], const Vec3& lightDir);
  bool Set1DTrans(const std::vector<uint8_t>& rgba);           // AbstrRenderer.h:189 (pure virtual there)
  TransferFunction1D& Get1DTrans() { return m_tf1d; }
  void Changed1DTrans();                                        // tuvok.gpu.changed1DTrans
  bool Set2DTrans(const uint8_t* rgba, uint32_t w, uint32_t h, double nzLimits[4]);
  // RenderRegion::SetClipPlane / EnableClipPlane (world-space plane a,b,c,d; n.p + d <= 0 is kept)
  void SetClipPlane(const float plane[4]);
  void EnableClipPlane(bool on);
  // stereo (AbstrRenderer.cpp:1387-1410, :1577-1585; the Lua names setStereoEnabled, setStereoEyeDist, setStereoFocalLength,
  // setStereoMode, setStereoEyeSwap, initStereoFrame, toggleStereoFrame): both eyes are marched in every sub-frame, share the
  // request table and the pool, and are composed into one picture when the frame has converged (GLRenderer::EndFrame)
  // ClearView (classic renderer, isosurface mode; AbstrRenderer.cpp:1247-1256,1304-1361, GLRenderer.cpp:2833-2851; the Lua names
  // setClearViewEnabled, setCVIsoValue, setCVColor, setCVSize, setCVContextScale, setCVBorderScale, setCVFocusPos)
  void SetCV(bool on) { if (on != m_cv) { m_cv = on; if (m_mode == RM_ISOSURFACE) ScheduleCompleteRedraw(); } }
  bool GetCV() const { return m_cv && m_classic; }                           // SupportsClearView: the classic renderer only
  void SetCVIsoValue(float v) { if (v != m_cvIsoValue) { m_cvIsoValue = v; if (m_cv && m_mode == RM_ISOSURFACE) ScheduleCompleteRedraw(); } }
  float GetCVIsoValue() const { return m_cvIsoValue; }
  void SetCVColor(const Vec3& c) { m_cvColor = c; CVChanged(); }
  void SetCVSize(float f) { m_cvSize = f; CVChanged(); }
  void SetCVContextScale(float f) { m_cvContextScale = f; CVChanged(); }
  void SetCVBorderScale(float f) { m_cvBorderScale = f; CVChanged(); }
  void SetCVFocusPosFVec(const float p[4]) { memcpy(m_cvPos, p, 16); CVChanged(); }
  bool Pick(int x, int y, float out[3]);            // GLRenderer::Pick: eye-space isosurface position under a window position
  bool SetCVFocusPos(int x, int y);                 // window coordinates, y from the top (the mouse position)
  bool ComputeClearViewParams(tvk_clearview_params& cv, int eye = 0);
  Matrix4 ModelView(int eye = 0) const;
  void SetStereo(bool on) { if (on != m_stereo) { m_stereo = on; ScheduleCompleteRedraw(); } }
  bool GetStereo() const { return m_stereo; }
  void SetStereoEyeDist(float d) { m_stereoEyeDist = d; if (m_stereo) ScheduleCompleteRedraw(); }
  void SetStereoFocalLength(float f) { m_stereoFocalLength = f; if (m_stereo) ScheduleCompleteRedraw(); }
  bool SetStereoMode(int mode);
  void SetStereoEyeSwap(bool swap);
  void InitStereoFrame() { m_afId = 0; ScheduleCompleteRedraw(); }
  bool ToggleStereoFrame();
  bool ComposeStereo();
  Vec3 GetViewPos() const { return m_eye; }
  Vec3 GetViewDir() const { return (m_at - m_eye).normalized(); }  // AbstrRenderer.cpp:1608-1610
  Vec3 GetUpDir() const { return m_up; }
  bool SetTimestep(size_t t);
  size_t Timesteps() const { return m_timesteps; }
  size_t Timestep() const { return m_timestep; }
  void SetLODLimits(uint32_t lo, uint32_t hi);                     // kept for the classic path; the ray-guided path ignores it
  void SetBGColors(const float top[3], const float bottom[3]);     // display path only (captures read the FBO)
  void SetRendererTarget(int t) { m_target = t; }
  void SetMissingBrickStrategy(uint32_t s) { m_strategy = s; ScheduleCompleteRedraw(); }
  void SetHashTable(uint32_t size, uint32_t rehash) { m_hashSize = size; m_rehash = rehash; }
  void SetSampler(uint32_t s) { m_sampler = s; }
  void ScheduleCompleteRedraw() { m_blank = true; m_converged = false; }

  // --- renderer type: false = ray-guided GLGridLeaper path (default), true = classic per-brick GLRaycaster path
  void SetClassic(bool on) { m_classic = on; ScheduleCompleteRedraw(); }
  // forceLod < 0: the LOD the view needs (m_iMinLODForCurrentView; 0 for RT_CAPTURE); else the brick list of that LOD.
  // minLodOut: m_iMinLODForCurrentView of the view.
  // ov: model-view / projection of the caller instead of the camera's, and (allBricks) every brick of the LOD without the
  // frustum, clip-plane and contains-data tests -- the brick list of the full-resolution MIP
  struct ClassicPlanOverride { Matrix4 mv, proj; bool allBricks = false; };
  bool PlanClassicFrame(tvk_classic_params& p, std::vector<tvk_classic_brick>& bricks, uint32_t& lod, int forceLod = -1,
                        uint32_t* minLodOut = nullptr, int eye = 0, const ClassicPlanOverride* ov = nullptr);
  // the capture mode's full-resolution MIP of a 2-D region (window mode, flipView, m_fMIPRotationAngle, m_bOrthoView, m_bMIPLOD)
  bool PlanHQMIP(int windowMode, bool flipX, bool flipY, float angleDeg, bool ortho, bool mipLOD, tvk_mip_params& p,
                 std::vector<tvk_classic_brick>& bricks, uint32_t& lod);
  bool RenderHQMIP(int windowMode, bool flipX, bool flipY, float angleDeg, bool ortho, bool mipLOD);
  static void ComputeMaxLODForCurrentView(tvkr_lod_state& s);        // AbstrRenderer.cpp:681-787
  void SetMaxMSPerFrame(float ms) { m_ladder.max_ms_per_frame = ms; }
  void OverrideSubframeMs(float ms) { m_msOverride = ms; }
  const tvkr_lod_state& LodState() const { return m_ladder; }
  uint32_t CurrentLOD() const { return m_currentLOD; }
  bool PaintClassic();
  const void* ClassicStore();        // the brick store in device memory (copied up once for datasets kept in pinned host memory)
  bool PaintClassicWith(const void* store, tvk_classic_params& p, std::vector<tvk_classic_brick>& bricks, uint32_t& lod, uint32_t& minLod);

  // --- frame loop
  bool Paint();                      // one sub-frame (GLGridLeaper::Render3DRegion, or the classic path's final sub-frame)
  bool CheckForRedraw() const;       // GLGridLeaper.cpp:872-890
  bool Converged() const { return m_converged; }        // PH_Converged
  bool PaintUntilConverged(uint32_t maxSubframes, uint32_t& subframes);
  void ClearWorkingSet();            // PH_ClearWorkingSet: pool back to the pinned coarsest brick
  bool RecalcVisibility();           // PH_RecalculateVisibility
  // GLFrameCapture.cpp:47-86 + FrameCapture.h:83-159: RGBA8, top row first
  bool GetFrameRGBA8(std::vector<uint8_t>& out);
  bool GetFrameRGBA8(uint8_t* out);
  bool GetFrameFloat(std::vector<float>& out);             // premultiplied RGBA32F, bottom row first
  bool CaptureSingleFrame(const std::string& filename, bool preserveTransparency);

  // --- introspection used by tests / the C wrapper
  bool ComputeRaycastParams(tvk_raycast_params& p, int eye = 0);    // eye: which of the two stereo views (ignored in mono)
  const BrickedVolume* Volume() const { return m_vol; }
  BrickPool* Pool() { return m_pool; }
  tvk_ctx* Context() { return m_ctx; }
  PerfCounters& Perf() { return m_perf; }
  double PerfQuery(int counter);                        // tuvok.perf(id): value since the last query, then reset
  void IncrementPerfCounter(int counter, double amount) { if (counter >= 0 && counter < PERF_END) m_perfTable[counter] += amount; }
  uint32_t Width() const { return m_w; }
  uint32_t Height() const { return m_h; }

 private:
  bool Fail(const std::string& what, int rc);
  bool FinishLoad(BrickedVolume* v, uint64_t gpuMemBytes, uint32_t tfSize);
  BrickedVolume* Rebrick(const BrickedVolume* src, uint32_t brickSize, int minmaxMode);
  std::string m_file;                  // the UVF the dataset came from (SetTimestep reads another step of it)
  uint64_t m_fileGpuMem = 0;
  bool m_fileVerify = false, m_fileKeepCompressed = false;
  uint32_t m_fileRebrick = 0;
  int m_fileRebrickMinMax = 0;
  size_t m_timestep = 0, m_timesteps = 1, m_loadTimestep = 0;
  uint32_t m_rebrickSize = 0;          // != 0 while LoadRebricked runs LoadDataset
  int m_rebrickMinMax = 0;
  bool RecomputeBrickVisibility(bool force);            // GLGridLeaper.cpp:632-688
  tvk_ctx* m_ctx = nullptr;
  BrickedVolume* m_vol = nullptr;
  BrickPool* m_pool = nullptr;
  std::string m_err;
  uint32_t m_w = 0, m_h = 0;
  int m_mode = RM_1DTRANS, m_target = RT_INTERACTIVE;
  bool m_lighting = true;                                // AbstrRenderer.cpp:81
  float m_sampleRate = 1.0f, m_fov = 50.0f, m_znear = 0.01f, m_zfar = 1000.0f;
  Vec3 m_eye{0, 0, 1.6f}, m_at{0, 0, 0}, m_up{0, 1, 0};
  Matrix4 m_rotation, m_translation;
  Vec3 m_lightDir{0, 0, -1}, m_isoColor{0.5f, 0.5f, 0.5f};
  std::vector<uint8_t> m_tf2dBytes;                                   // host copy of the rasterised 2D table while the context does not hold it
  bool m_tf2dInCtx = false;                                           // the context's staging block holds the table (tvk_tf2d_staging)
  int Upload2DTable();
  uint32_t m_tf2dW = 0, m_tf2dH = 0;
  float m_ambient[4] = {1, 1, 1, 0.1f}, m_diffuse[4] = {1, 1, 1, 1}, m_specular[4] = {1, 1, 1, 1};
  double m_isoValue = 0.5;
  TransferFunction1D m_tf1d;
  bool m_has2d = false;
  double m_tf2dLimits[4] = {0, 0, 0, 0};
  uint32_t m_strategy = TVK_MBS_SKIP_TWO_LEVELS, m_hashSize = 509, m_rehash = 10, m_sampler = kSamplerAuto;
  uint32_t m_activeSampler = TVK_SAMPLER_EXACT;
  bool m_blank = true, m_converged = false;
  tvk_visibility m_lastVis{};
  bool m_visValid = false;
  PerfCounters m_perf;
  double m_perfTable[PERF_END] = {};
  bool m_classic = false;
  bool m_clipEnabled = false;
  float m_clipPlane[4] = {0, 0, 1, 0};
  bool m_cv = false;                                              // AbstrRenderer.cpp:130-136,171
  bool m_cvFrame = false;                                         // the last classic frame was a ClearView one (hit buffers readable)
  float m_cvIsoValue = 0.8f, m_cvSize = 5.5f, m_cvContextScale = 1.0f, m_cvBorderScale = 60.0f;
  Vec3 m_cvColor{1, 0, 0};
  float m_cvPos[4] = {0, 0, 0.5f, 1.0f};
  void CVChanged() { if (m_cv && m_mode == RM_ISOSURFACE) ScheduleCompleteRedraw(); }   // (the reference only recomposes)
  bool m_stereo = false, m_stereoEyeSwap = false;                 // AbstrRenderer.cpp:141-145
  float m_stereoEyeDist = 0.02f, m_stereoFocalLength = 1.0f;
  int m_stereoMode = TVK_SM_RB;
  uint32_t m_afId = 0;
  uint32_t m_lodLimits[2] = {0, 0};
  struct SortLast { bool on = false; uint32_t members = 1, member = 0, global[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0},
                    start[3] = {0, 0, 0}, frame = 0; double maxValue = 0; bool exchange = false; } m_sl;
  BrickedVolume* m_streamVol = nullptr;      // between LoadStreamedBegin and LoadStreamedEnd
  tvk_build_desc m_streamDesc{};
  float m_lastBuildMs = 0.0f;
  tvkr_lod_state m_ladder{0, {-1.0f, -1.0f}, 10000.0f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t m_currentLOD = 0;
  float m_msOverride = -1.0f, m_msCurrentFrame = 0.0f;
  // a view whose start LOD is already the coarsest and still too slow first renders a pass at half the sampling rate
  // (decreaseSamplingRateNow); the reduced-resolution pass (decreaseScreenResNow) is counted but rendered at full size
  bool m_decRateNow = false, m_decResNow = false;
  float m_bg[2][3] = {{0, 0, 0}, {0, 0, 0}};
};

}  // namespace tvk
