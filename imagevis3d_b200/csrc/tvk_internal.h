// Internal definitions shared by the CUDA translation units behind include/tvk_cuda.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "tvk_cuda.h"

// Constants the reference bakes into the generated GLSL (GLVolumePool.cpp:346-483,
// GLHashTable.cpp:112-135); passed to the kernels as one __grid_constant__ struct.
struct PoolConst {
  uint32_t lod_count;
  uint32_t total_bricks;
  float    lod_layout[TVK_MAX_LOD][3];   // vLODLayout
  uint32_t lod_layout_x[TVK_MAX_LOD];    // iLODLayoutSize.x
  uint32_t lod_layout_xy[TVK_MAX_LOD];   // iLODLayoutSize.y
  uint32_t lod_offset[TVK_MAX_LOD];      // vLODOffset
  uint32_t lod_bricks[TVK_MAX_LOD][3];   // integer GetBrickLayout (host tables)
  uint32_t l0_bricks[3];                 // hash Serialize strides
  uint32_t capacity[3];                  // poolCapacity
  uint32_t max_brick[3];                 // maxTotalBrickSize
  uint32_t pool_dim[3];                  // iPoolSize (voxels)
  float    pool_size[3];                 // vec3(iPoolSize)
  float    overlap_tc[3];                // "overlap" in pool tex coords
  float    volume_size[3];               // volumeSize
  uint32_t bpv;
  uint32_t phys_cap[3];                  // TVK_SAMPLER_GATHER: slots per axis of the layered array (z = layer groups)
};

struct BuildState {
  tvk_build_desc desc{};
  tvk_build_info info{};
  std::vector<uint64_t> offsets;     // byte offset per brick
  std::vector<uint32_t> sizes;       // 3 per brick
  void*     d_store = nullptr;       // brick store
  uint64_t* d_offsets = nullptr;
  uint32_t* d_sizes = nullptr;
  double*   d_minmax = nullptr;      // 2 per brick
  unsigned long long* d_hist1d = nullptr;
  void*     d_lod[TVK_MAX_LOD] = {}; // flat LOD volumes (lod 0 = raw when raw_on_device)
  bool      own_lod0 = false;
  bool      own_store = true;        // false: caller-owned store (tvk_build_stream_begin)
  uint32_t* d_mn = nullptr;          // packed min | max | global max scratch while a build is in progress
  bool      hist_fused = false, hist_full = false, lod0_fused = false, hist_pending = false, hist_upper_pending = false;
  bool      streaming = false;       // between tvk_build_stream_begin and _end
  uint32_t  next_bz = 0;             // first LOD-0 brick row not fed yet
  bool      valid = false;
  float     kernels_ms = 0.f;        // CUDA-event time of the builder kernels of the last tvk_build_octree
  float     hist2d_ms = 0.f;         // ... and of the last tvk_build_hist2d
};

// sort-last exchange over peer memory (tvk_exchange.cu)
#define TVK_XCHG_MAX_WORLD 16
struct ExchangeState {
  uint32_t world = 0, rank = 0, width = 0, height = 0, rows = 0;
  void* local = nullptr;                          // this rank's region (cudaMalloc)
  void* base[TVK_XCHG_MAX_WORLD] = {};            // every rank's region as seen from this device
  uint32_t* flags[TVK_XCHG_MAX_WORLD] = {};
  bool ipc_opened[TVK_XCHG_MAX_WORLD] = {};
  uint32_t send_frame = 0;                        // != 0: the next ray-march launch also sends its stripes
};
struct ExchangeSend {                             // what K1 needs; world == 0: local image only
  float4* recv[TVK_XCHG_MAX_WORLD];               // owner r's buffer for this rank's stripe (rows*W pixels)
  uint32_t* flag[TVK_XCHG_MAX_WORLD];             // owner r's arrival flag for this rank
  uint32_t world, rank, rows, frame;
  uint32_t flag_index;                            // parity*16 + rank: this rank's entry in an owner's flag / rectangle tables
  uint32_t rect[4];                               // x0, y0, x1, y1: the pixels this rank sends (everything else reads as zero)
};

// Pinned host ring for the small parameter blocks that accompany asynchronous work (placement / gather job lists,
// page-table patches).  A cudaMemcpyAsync FROM PAGEABLE memory first synchronises the host with the stream; from this
// ring it does not, so issuing an upload never waits for the ray-march in flight.
#define TVK_RING_SEGMENTS 16
#define TVK_RING_SEGMENT_BYTES (256u << 10)
struct PinnedRing {
  uint8_t* base = nullptr;
  cudaEvent_t done[TVK_RING_SEGMENTS] = {};
  bool in_flight[TVK_RING_SEGMENTS] = {};
  int next = 0;
};

struct tvk_ctx {
  int device = 0;
  PinnedRing ring;
  cudaStream_t stream = nullptr;       // render / builder stream (own_stream unless tvk_set_stream)
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // pager side stream
  cudaStream_t place_stream = nullptr; // pager: placement of batch b runs beside the gather of batch b+1
  cudaEvent_t ev_gather[2] = {nullptr, nullptr}, ev_place[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, copy_done = nullptr, render_done = nullptr, bev0 = nullptr, bev1 = nullptr;
  std::string err;
  uint32_t march_ctas_per_sm = 0;      // tvk_set_march_occupancy (0 = occupancy calculator)

  // pool
  bool has_pool = false;
  tvk_pool_desc pdesc{};
  PoolConst pc{};
  void* d_pool = nullptr;              // linear pool (TVK_SAMPLER_EXACT)
  cudaArray_t pool_array = nullptr;    // 3-D array (TVK_SAMPLER_TEX3D)
  cudaTextureObject_t pool_tex = 0;
  cudaSurfaceObject_t pool_surf = 0;   // TVK_SAMPLER_GATHER: placement target (kept for the life of the pool)
  uint32_t* d_meta = nullptr;
  double* d_minmax = nullptr;
  double* d_gminmax = nullptr;
  uint32_t n_minmax = 0;
  uint32_t* d_counts = nullptr;

  // transfer functions
  uchar4* d_tf = nullptr;
  float4* d_tf_f = nullptr;            // float copy of the table (byte * (1/255))
  float4* d_tf_f1 = nullptr;           // the same with alpha already opacity-corrected for sampleRateModifier == 1: 1 - (1 - a)
  uint32_t* d_tf_colmask = nullptr;    // bit x: column x of the table holds a non-zero alpha (any row); 4096 bits max... sized tf_w
  uint32_t tf_w = 0, tf_h = 0;
  // pinned staging for table edits (a 4096x256 2-D table is 4 MB): the caller's bytes are copied here and travel
  // asynchronously, so that an edit costs no host synchronisation; tf_stage_done guards the buffer's reuse
  uint8_t* h_tf_stage = nullptr;
  size_t tf_stage_bytes = 0;
  uint32_t tf_stage_w = 0, tf_stage_h = 0;   // the 2-D table the block holds (tvk_tf2d_staging)
  cudaEvent_t tf_stage_done = nullptr;
  bool tf_stage_busy = false;
  uint8_t* h_tf1_stage = nullptr;      // the same for 1-D tables (their own block: the 2-D one doubles as the host's copy of the 2-D table)
  size_t tf1_stage_bytes = 0;
  cudaEvent_t tf1_stage_done = nullptr;
  bool tf1_stage_busy = false;

  // frame state
  uint32_t width = 0, height = 0;
  float4* d_image = nullptr;           // accRayColor / composed iso image
  uchar4* d_rgba8 = nullptr;           // tvk_read_image_rgba8 staging (allocated on first use)
  size_t rgba8_px = 0;                 // pixels d_rgba8 holds
  uint8_t* h_rgba8 = nullptr;          // pinned landing buffer of the RGBA8 read-back (the caller's memory may be pageable)
  size_t h_rgba8_px = 0;
  uint64_t lz4_bytes_in = 0, lz4_bytes_out = 0;   // compressed bytes moved over PCIe / bytes inflated on the device
  float4* d_pos[2] = {nullptr, nullptr};   // rayStartPoint ping-pong
  float4* d_aux[2] = {nullptr, nullptr};   // rayStartColor (iso: normal) ping-pong
  int cur = 0;
  uint32_t* d_hash = nullptr;
  uint32_t hash_size = 0;
  uint32_t* d_tile_counter = nullptr;
  unsigned long long* d_stats = nullptr;   // samples, rays, brick steps
  float last_raycast_ms = 0.f;
  uint32_t launches = 0;
  // request read-back: one pinned staging block for the table AND the counters, so that a sub-frame costs ONE host
  // synchronisation (tvk_read_requests); tvk_get_stats right after it is answered from the cached copy
  uint32_t* h_readback = nullptr;
  uint32_t h_readback_words = 0;
  uint64_t k1_serial = 0, stats_serial = ~0ull;
  unsigned long long cached_stats[3] = {0, 0, 0};

  // stereo (tvk_select_eye): the frame state of the eye that is NOT selected, and the composed picture
  struct EyeState {
    float4* image = nullptr;
    float4* pos[2] = {nullptr, nullptr};
    float4* aux[2] = {nullptr, nullptr};
    uint32_t width = 0, height = 0;
  } other_eye;
  uint32_t eye = 0;
  float4* d_stereo = nullptr;          // tvk_compose_stereo target
  size_t stereo_px = 0;
  bool show_stereo = false;            // the read-back calls return d_stereo (until the next launch that draws)
  const float4* iso_hits = nullptr;    // the first surface's hit positions of the last isosurface frame (tvk_read_iso_hit), or nullptr

  BuildState build;
  ExchangeState xchg;
};

// the picture the read-back calls return
static inline float4* tvk_display_image(const tvk_ctx* c) { return c->show_stereo ? c->d_stereo : c->d_image; }

#define TVK_CHECK_CUDA(ctx, call)                                                       \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                  \
      return (_e == cudaErrorMemoryAllocation) ? TVK_ERR_NOMEM : TVK_ERR_CUDA;          \
    }                                                                                   \
  } while (0)

#define TVK_FAIL(ctx, code, msg) do { (ctx)->err = (msg); return (code); } while (0)

void tvk_compute_pool_const(const tvk_pool_desc& d, PoolConst& pc);
int  tvk_ensure_frame(tvk_ctx* ctx, uint32_t w, uint32_t h, uint32_t hash_size);
void tvk_exchange_fill(const tvk_ctx* ctx, ExchangeSend& s);
// H2D copy of a small host block on stream st through the pinned ring (falls back to a plain copy for large blocks)
int  tvk_copy_params_async(tvk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st);
