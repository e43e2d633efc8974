// Context, brick pool (storage + pager upload path), metadata table and transfer
// functions behind include/tvk_cuda.h.
//
// HBM layout of the pool (replaces the GL 3-D texture of GLVolumePool.cpp:864-870):
//   TVK_SAMPLER_EXACT : one linear array [pool_dim.z][pool_dim.y][pool_dim.x] of
//                       UNORM8/16 voxels; slot (sx,sy,sz) starts at voxel
//                       (sx,sy,sz)*max_brick_size -- identical addressing to the
//                       reference's texture so the shader constants carry over.
//   TVK_SAMPLER_TEX3D : the same voxels in a cudaArray (surface-writable) bound
//                       to a normalised-float, linear-filter texture object.
//   TVK_SAMPLER_GATHER: the same voxels in a LAYERED 2-D cudaArray, one layer per pool z
//                       plane, read with the four-texel gather (PTX tld4.a2d) through a point-
//                       sampled, element-type texture object: the texture units do the 2-D
//                       addressing and caching, the kernel does the exact fp32 trilinear filter.
//                       1x the memory of the volume (the quad-packed pool needs 4x).  The slot
//                       grid of the array is PHYSICAL (pc.phys_cap, <= 2048 layers); the logical
//                       capacity the reference's shader arithmetic sees is unchanged.
// Brick upload = contiguous DMA (pinned host -> staging ring in HBM, side stream)
// followed by a placement kernel that scatters each brick into its slot with
// 16-byte accesses; the reference does one synchronous glTexSubImage3D per brick
// from pageable memory (GLVolumePool.cpp:673-717).
#include "tvk_internal.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <algorithm>

// ---------------------------------------------------------------------------
// host tables: GLVolumePool.cpp:82-117 (GetLoDSize, GetFloatBrickLayout,
// GetBrickLayout) and the LoD offset table of the ctor (:222-230)
// ---------------------------------------------------------------------------
void tvk_compute_pool_const(const tvk_pool_desc& d, PoolConst& pc) {
  memset(&pc, 0, sizeof(pc));
  pc.lod_count = d.lod_count;
  pc.bpv = d.bytes_per_voxel;
  uint32_t inner[3], base[3];
  for (int a = 0; a < 3; ++a) {
    inner[a] = d.max_brick_size[a] - 2 * d.overlap;
    base[a] = (uint32_t)std::ceil((double)d.volume_size[a] / inner[a]);
  }
  uint32_t off = 0;
  for (uint32_t l = 0; l < d.lod_count; ++l) {
    const uint32_t pw = 1u << l;
    for (int a = 0; a < 3; ++a) {
      float f = (float)d.volume_size[a] / (float)inner[a];
      f /= (float)pw;
      if ((float)(uint32_t)f == f) f -= f * FLT_EPSILON;   // the epsilon nudge, :99-104
      pc.lod_layout[l][a] = f;
      pc.lod_bricks[l][a] = (uint32_t)std::ceil((double)base[a] / pw);
    }
    pc.lod_layout_x[l] = (uint32_t)std::ceil(pc.lod_layout[l][0]);
    pc.lod_layout_xy[l] = (uint32_t)std::ceil(pc.lod_layout[l][0]) * (uint32_t)std::ceil(pc.lod_layout[l][1]);
    pc.lod_offset[l] = off;
    off += pc.lod_bricks[l][0] * pc.lod_bricks[l][1] * pc.lod_bricks[l][2];
  }
  pc.total_bricks = pc.lod_offset[d.lod_count - 1] + 1;    // m_iTotalBrickCount, :877
  for (int a = 0; a < 3; ++a) {
    pc.l0_bricks[a] = pc.lod_bricks[0][a];
    pc.capacity[a] = d.pool_capacity[a];
    pc.max_brick[a] = d.max_brick_size[a];
    pc.pool_dim[a] = d.pool_capacity[a] * d.max_brick_size[a];
    pc.pool_size[a] = (float)pc.pool_dim[a];
    pc.overlap_tc[a] = (float)(d.max_brick_size[a] - inner[a]) / (2.0f * pc.pool_size[a]);
    pc.volume_size[a] = (float)d.volume_size[a];
  }
}

extern "C" int tvk_create(int device, tvk_ctx** out) {
  if (!out) return TVK_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return TVK_ERR_CUDA;
  tvk_ctx* c = new tvk_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->place_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_gather[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_gather[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_place[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_place[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaEventCreate(&c->bev0) != cudaSuccess || cudaEventCreate(&c->bev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->render_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaMalloc(&c->d_tile_counter, 64) != cudaSuccess ||
      cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&c->d_counts, 4 * sizeof(uint32_t)) != cudaSuccess) {
    delete c;
    return TVK_ERR_CUDA;
  }
  c->own_stream = c->stream;
  if (cudaHostAlloc((void**)&c->ring.base, (size_t)TVK_RING_SEGMENTS * TVK_RING_SEGMENT_BYTES, cudaHostAllocDefault) != cudaSuccess) {
    c->ring.base = nullptr;                     // not fatal: parameter blocks then travel from pageable memory
    (void)cudaGetLastError();
  }
  for (int i = 0; i < TVK_RING_SEGMENTS; ++i) cudaEventCreateWithFlags(&c->ring.done[i], cudaEventDisableTiming);
  cudaMemset(c->d_tile_counter, 0, 64);
  cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long));
  // keep stream-ordered allocations (pager staging ring, job lists) cached between frames
  cudaMemPool_t mp;
  if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = c;
  return TVK_OK;
}

int tvk_copy_params_async(tvk_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return TVK_OK;
  PinnedRing& r = c->ring;
  if (bytes > TVK_RING_SEGMENT_BYTES || !r.base) {
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return TVK_OK;
  }
  const int k = r.next;
  r.next = (r.next + 1) % TVK_RING_SEGMENTS;
  if (r.in_flight[k]) TVK_CHECK_CUDA(c, cudaEventSynchronize(r.done[k]));     // 16 blocks ago: long finished in practice
  uint8_t* seg = r.base + (size_t)k * TVK_RING_SEGMENT_BYTES;
  memcpy(seg, src, bytes);
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(dst, seg, bytes, cudaMemcpyHostToDevice, st));
  TVK_CHECK_CUDA(c, cudaEventRecord(r.done[k], st));
  r.in_flight[k] = true;
  return TVK_OK;
}

static void free_frame(tvk_ctx* c) {
  c->iso_hits = nullptr;
  cudaFree(c->d_image); c->d_image = nullptr;
  cudaFree(c->d_rgba8); c->d_rgba8 = nullptr; c->rgba8_px = 0;
  for (int i = 0; i < 2; ++i) { cudaFree(c->d_pos[i]); cudaFree(c->d_aux[i]); c->d_pos[i] = c->d_aux[i] = nullptr; }
  c->width = c->height = 0;
}

extern "C" void tvk_destroy(tvk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  tvk_exchange_destroy(c);
  tvk_build_release(c);
  tvk_pool_destroy(c);
  free_frame(c);
  cudaFree(c->other_eye.image);
  for (int i = 0; i < 2; ++i) { cudaFree(c->other_eye.pos[i]); cudaFree(c->other_eye.aux[i]); }
  cudaFree(c->d_stereo);
  cudaFree(c->d_tf); cudaFree(c->d_tf_f); cudaFree(c->d_tf_f1); cudaFree(c->d_tf_colmask); cudaFree(c->d_hash); cudaFree(c->d_tile_counter); cudaFree(c->d_stats); cudaFree(c->d_counts);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->bev0); cudaEventDestroy(c->bev1); cudaEventDestroy(c->copy_done); cudaEventDestroy(c->render_done);
  for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->ev_gather[i]); cudaEventDestroy(c->ev_place[i]); }
  cudaStreamDestroy(c->own_stream); cudaStreamDestroy(c->copy_stream); cudaStreamDestroy(c->place_stream);
  for (int i = 0; i < TVK_RING_SEGMENTS; ++i) if (c->ring.done[i]) cudaEventDestroy(c->ring.done[i]);
  if (c->ring.base) cudaFreeHost(c->ring.base);
  if (c->h_readback) cudaFreeHost(c->h_readback);
  if (c->h_tf_stage) cudaFreeHost(c->h_tf_stage);
  if (c->tf_stage_done) cudaEventDestroy(c->tf_stage_done);
  if (c->h_tf1_stage) cudaFreeHost(c->h_tf1_stage);
  if (c->tf1_stage_done) cudaEventDestroy(c->tf1_stage_done);
  if (c->h_rgba8) cudaFreeHost(c->h_rgba8);
  delete c;
}

extern "C" const char* tvk_last_error(const tvk_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int tvk_set_stream(tvk_ctx* c, void* cuda_stream) {
  if (!c) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return TVK_OK;
}

extern "C" int tvk_set_march_occupancy(tvk_ctx* c, uint32_t ctas_per_sm) {
  if (!c) return TVK_ERR_INVALID;
  c->march_ctas_per_sm = ctas_per_sm;
  return TVK_OK;
}

extern "C" int tvk_device_synchronize(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

int tvk_ensure_frame(tvk_ctx* c, uint32_t w, uint32_t h, uint32_t hash_size) {
  if (w == 0 || h == 0) TVK_FAIL(c, TVK_ERR_INVALID, "zero-sized render target");
  if (c->width != w || c->height != h) {
    free_frame(c);
    const size_t bytes = (size_t)w * h * sizeof(float4);
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_image, bytes));
    for (int i = 0; i < 2; ++i) {
      TVK_CHECK_CUDA(c, cudaMalloc(&c->d_pos[i], bytes));
      TVK_CHECK_CUDA(c, cudaMalloc(&c->d_aux[i], bytes));
    }
    c->width = w; c->height = h; c->cur = 0;
  }
  if (c->hash_size != hash_size) {
    cudaFree(c->d_hash); c->d_hash = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_hash, sizeof(uint32_t) * hash_size));
    c->hash_size = hash_size;
  }
  return TVK_OK;
}

// ---------------------------------------------------------------------------
// pool
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clear_layered_kernel(cudaSurfaceObject_t surf, uint32_t row_bytes, uint32_t rows, uint32_t layers) {
  const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  if (x >= row_bytes) return;
  const bool whole = x + 4u <= row_bytes;                               // rows of odd-sized slots end in a partial word
  for (uint32_t l = blockIdx.z; l < layers; l += gridDim.z)
    for (uint32_t y = blockIdx.y; y < rows; y += gridDim.y) {
      if (whole) surf2DLayeredwrite(0u, surf, (int)x, (int)y, (int)l);
      else for (uint32_t k = x; k < row_bytes; ++k) surf2DLayeredwrite((uint8_t)0, surf, (int)k, (int)y, (int)l);
    }
}

extern "C" int tvk_pool_create(tvk_ctx* c, const tvk_pool_desc* d) {
  if (!c || !d) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  if (d->bytes_per_voxel != 1 && d->bytes_per_voxel != 2) TVK_FAIL(c, TVK_ERR_UNSUPPORTED, "only 8/16 bit single component volumes");
  if (d->lod_count == 0 || d->lod_count > TVK_MAX_LOD) TVK_FAIL(c, TVK_ERR_INVALID, "bad lod_count");
  for (int a = 0; a < 3; ++a)
    if (d->max_brick_size[a] <= 2 * d->overlap || d->pool_capacity[a] == 0 || d->volume_size[a] == 0)
      TVK_FAIL(c, TVK_ERR_INVALID, "bad pool description");
  if (d->sampler > TVK_SAMPLER_GATHER) TVK_FAIL(c, TVK_ERR_INVALID, "bad sampler");
  tvk_pool_destroy(c);
  c->pdesc = *d;
  tvk_compute_pool_const(*d, c->pc);
  const size_t voxels = (size_t)c->pc.pool_dim[0] * c->pc.pool_dim[1] * c->pc.pool_dim[2];
  if (d->sampler == TVK_SAMPLER_EXACT || d->sampler == TVK_SAMPLER_EXACT_Q4) {
    const size_t elem = (d->sampler == TVK_SAMPLER_EXACT_Q4 ? 4 : 1) * (size_t)d->bytes_per_voxel;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_pool, voxels * elem));
    TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_pool, 0, voxels * elem, c->stream));
  } else if (d->sampler == TVK_SAMPLER_GATHER) {
    // physical slot grid: as many layer groups as the 2048-layer limit allows, the rest spread over a near-square
    // X x Y grid (<= 32768 texels per side)
    int max_w = 0, max_h = 0, max_l = 0;
    TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&max_w, cudaDevAttrMaxTexture2DLayeredWidth, c->device));
    TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&max_h, cudaDevAttrMaxTexture2DLayeredHeight, c->device));
    TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&max_l, cudaDevAttrMaxTexture2DLayeredLayers, c->device));
    const uint64_t nslots = (uint64_t)d->pool_capacity[0] * d->pool_capacity[1] * d->pool_capacity[2];
    uint64_t pz = std::max<uint64_t>(1, std::min<uint64_t>(nslots, (uint64_t)max_l / d->max_brick_size[2]));
    // fewer layer groups than the limit when that keeps the grid closer to a cube (small pools)
    while (pz > 1 && pz * pz * pz > nslots * 8) --pz;
    const uint64_t per_layer = (nslots + pz - 1) / pz;
    uint64_t px = (uint64_t)std::ceil(std::sqrt((double)per_layer));
    px = std::max<uint64_t>(1, std::min<uint64_t>(px, (uint64_t)max_w / d->max_brick_size[0]));
    const uint64_t py = (per_layer + px - 1) / px;
    if (py * d->max_brick_size[1] > (uint64_t)max_h || d->max_brick_size[2] > (uint32_t)max_l)
      TVK_FAIL(c, TVK_ERR_UNSUPPORTED, "pool too large for a layered array (use TVK_SAMPLER_EXACT)");
    c->pc.phys_cap[0] = (uint32_t)px; c->pc.phys_cap[1] = (uint32_t)py; c->pc.phys_cap[2] = (uint32_t)pz;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc(8 * d->bytes_per_voxel, 0, 0, 0, cudaChannelFormatKindUnsigned);
    // (cudaArrayTextureGather is rejected together with cudaArrayLayered; the gather instruction does not need it:
    //  benchmarks/gather_probe.cu, 4 M gathers checked on a B200)
    TVK_CHECK_CUDA(c, cudaMalloc3DArray(&c->pool_array, &fd,
                                        make_cudaExtent(px * d->max_brick_size[0], py * d->max_brick_size[1], pz * d->max_brick_size[2]),
                                        cudaArrayLayered | cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->pool_array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    TVK_CHECK_CUDA(c, cudaCreateTextureObject(&c->pool_tex, &rd, &td, nullptr));
    TVK_CHECK_CUDA(c, cudaCreateSurfaceObject(&c->pool_surf, &rd));
    // arrays are not zero-initialised; the linear pools are, and a fresh slot must read as zeros in all of them
    const uint32_t row_bytes = (uint32_t)(px * d->max_brick_size[0] * d->bytes_per_voxel);
    const uint32_t rows = (uint32_t)(py * d->max_brick_size[1]), layers = (uint32_t)(pz * d->max_brick_size[2]);
    clear_layered_kernel<<<dim3(((row_bytes + 3) / 4 + 255) / 256, std::min<uint32_t>(rows, 65535u), std::min<uint32_t>(layers, 65535u)), 256, 0, c->stream>>>(
        c->pool_surf, row_bytes, rows, layers);
    TVK_CHECK_CUDA(c, cudaGetLastError());
  } else {
    cudaChannelFormatDesc fd = cudaCreateChannelDesc(8 * d->bytes_per_voxel, 0, 0, 0, cudaChannelFormatKindUnsigned);
    TVK_CHECK_CUDA(c, cudaMalloc3DArray(&c->pool_array, &fd,
                                        make_cudaExtent(c->pc.pool_dim[0], c->pc.pool_dim[1], c->pc.pool_dim[2]),
                                        cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->pool_array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    TVK_CHECK_CUDA(c, cudaCreateTextureObject(&c->pool_tex, &rd, &td, nullptr));
  }
  TVK_CHECK_CUDA(c, cudaMalloc(&c->d_meta, sizeof(uint32_t) * c->pc.total_bricks));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_meta, 0, sizeof(uint32_t) * c->pc.total_bricks, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->has_pool = true;
  return TVK_OK;
}

extern "C" int tvk_pool_destroy(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  cudaSetDevice(c->device);
  if (c->pool_tex) { cudaDestroyTextureObject(c->pool_tex); c->pool_tex = 0; }
  if (c->pool_surf) { cudaDestroySurfaceObject(c->pool_surf); c->pool_surf = 0; }
  if (c->pool_array) { cudaFreeArray(c->pool_array); c->pool_array = nullptr; }
  cudaFree(c->d_pool); c->d_pool = nullptr;
  cudaFree(c->d_meta); c->d_meta = nullptr;
  cudaFree(c->d_minmax); c->d_minmax = nullptr;
  cudaFree(c->d_gminmax); c->d_gminmax = nullptr;
  c->n_minmax = 0;
  c->has_pool = false;
  return TVK_OK;
}

extern "C" int tvk_pool_total_bricks(const tvk_ctx* c, uint32_t* out) {
  if (!c || !out || !c->has_pool) return TVK_ERR_INVALID;
  *out = c->pc.total_bricks;
  return TVK_OK;
}

// Placement kernel: one CTA row-group per brick; copies brick b (tightly packed,
// x fastest) into its pool slot.  VEC = bytes moved per thread access.
struct PlaceJob {
  const uint8_t* src;
  uint32_t slot;
  uint32_t sx, sy, sz;   // voxels
};

template <int VEC>
__global__ void __launch_bounds__(256) place_bricks_linear(const PlaceJob* __restrict__ jobs, uint8_t* __restrict__ pool,
                                                            const __grid_constant__ PoolConst pc) {
  const PlaceJob j = jobs[blockIdx.x];
  const uint32_t bpv = pc.bpv;
  const uint32_t row_bytes = j.sx * bpv;
  const uint32_t units = row_bytes / VEC;                 // per row
  const uint32_t rows = j.sy * j.sz;
  const uint32_t s0 = j.slot % pc.capacity[0], s1 = (j.slot / pc.capacity[0]) % pc.capacity[1],
                 s2 = j.slot / (pc.capacity[0] * pc.capacity[1]);
  const size_t px = (size_t)pc.pool_dim[0] * bpv, pxy = px * pc.pool_dim[1];
  uint8_t* dst0 = pool + (size_t)s0 * pc.max_brick[0] * bpv + (size_t)s1 * pc.max_brick[1] * px + (size_t)s2 * pc.max_brick[2] * pxy;
  const uint32_t total = units * rows;
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const uint32_t r = i / units, u = i - r * units;
    const uint32_t y = r % j.sy, z = r / j.sy;
    const uint8_t* s = j.src + (size_t)r * row_bytes + (size_t)u * VEC;
    uint8_t* d = dst0 + (size_t)y * px + (size_t)z * pxy + (size_t)u * VEC;
    if (VEC == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
    else *d = *s;
  }
}

// TVK_SAMPLER_EXACT_Q4: pool element (x,y,z) = { v(x,y,z), v(x,y+1,z), v(x,y,z+1), v(x,y+1,z+1) }
// (y+1 / z+1 clamped inside the brick; the clamped copies are never sampled with weight > 0).
template <typename T, typename T4>
__global__ void __launch_bounds__(256) place_bricks_q4(const PlaceJob* __restrict__ jobs, T4* __restrict__ pool,
                                                        const __grid_constant__ PoolConst pc) {
  const PlaceJob j = jobs[blockIdx.x];
  const uint32_t s0 = j.slot % pc.capacity[0], s1 = (j.slot / pc.capacity[0]) % pc.capacity[1],
                 s2 = j.slot / (pc.capacity[0] * pc.capacity[1]);
  const size_t px = pc.pool_dim[0], pxy = px * pc.pool_dim[1];
  T4* dst0 = pool + (size_t)s0 * pc.max_brick[0] + (size_t)s1 * pc.max_brick[1] * px + (size_t)s2 * pc.max_brick[2] * pxy;
  const T* src = reinterpret_cast<const T*>(j.src);
  const uint32_t total = j.sx * j.sy * j.sz;
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const uint32_t x = i % j.sx, y = (i / j.sx) % j.sy, z = i / (j.sx * j.sy);
    const uint32_t y1 = min(y + 1, j.sy - 1), z1 = min(z + 1, j.sz - 1);
    T4 q;
    q.x = src[x + j.sx * (y + j.sy * z)];
    q.y = src[x + j.sx * (y1 + j.sy * z)];
    q.z = src[x + j.sx * (y + j.sy * z1)];
    q.w = src[x + j.sx * (y1 + j.sy * z1)];
    dst0[(size_t)x + (size_t)y * px + (size_t)z * pxy] = q;
  }
}

// Vector form of place_bricks_q4 for bricks whose rows are a multiple of four voxels: a thread builds four
// consecutive quads from four voxels of each of the rows (y,z), (y+1,z), (y,z+1), (y+1,z+1) -- four 4/8-byte
// loads, one (u8) or two (u16) 16-byte stores, no per-voxel division.
template <typename T>
__global__ void __launch_bounds__(256) place_bricks_q4_vec(const PlaceJob* __restrict__ jobs, void* __restrict__ pool_v,
                                                            const __grid_constant__ PoolConst pc) {
  const PlaceJob j = jobs[blockIdx.x];
  const uint32_t s0 = j.slot % pc.capacity[0], s1 = (j.slot / pc.capacity[0]) % pc.capacity[1],
                 s2 = j.slot / (pc.capacity[0] * pc.capacity[1]);
  const size_t px = pc.pool_dim[0], pxy = px * pc.pool_dim[1];
  const size_t base = (size_t)s0 * pc.max_brick[0] + (size_t)s1 * pc.max_brick[1] * px + (size_t)s2 * pc.max_brick[2] * pxy;   // in quads
  const uint32_t gx = j.sx / 4, rows = j.sy * j.sz, total = gx * rows;
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const uint32_t r = i / gx, x = (i - r * gx) * 4;
    const uint32_t y = r % j.sy, z = r / j.sy;
    const uint32_t y1 = min(y + 1, j.sy - 1), z1 = min(z + 1, j.sz - 1);
    const size_t quad = base + x + (size_t)y * px + (size_t)z * pxy;
    if (sizeof(T) == 1) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(j.src);
      const uint32_t a = src[(x + j.sx * (y + j.sy * z)) >> 2], b = src[(x + j.sx * (y1 + j.sy * z)) >> 2];
      const uint32_t c = src[(x + j.sx * (y + j.sy * z1)) >> 2], d = src[(x + j.sx * (y1 + j.sy * z1)) >> 2];
      // quad k = { a.byte[k], b.byte[k], c.byte[k], d.byte[k] }
      const uint32_t ab0 = __byte_perm(a, b, 0x5140), ab1 = __byte_perm(a, b, 0x7362);     // a0 b0 a1 b1 | a2 b2 a3 b3
      const uint32_t cd0 = __byte_perm(c, d, 0x5140), cd1 = __byte_perm(c, d, 0x7362);
      uint4 o;
      o.x = __byte_perm(ab0, cd0, 0x5410); o.y = __byte_perm(ab0, cd0, 0x7632);
      o.z = __byte_perm(ab1, cd1, 0x5410); o.w = __byte_perm(ab1, cd1, 0x7632);
      reinterpret_cast<uint4*>(static_cast<uchar4*>(pool_v) + quad)[0] = o;
    } else {
      const uint2* src = reinterpret_cast<const uint2*>(j.src);
      const uint2 a = src[(x + j.sx * (y + j.sy * z)) >> 2], b = src[(x + j.sx * (y1 + j.sy * z)) >> 2];
      const uint2 c = src[(x + j.sx * (y + j.sy * z1)) >> 2], d = src[(x + j.sx * (y1 + j.sy * z1)) >> 2];
      // quad k = { a.half[k], b.half[k], c.half[k], d.half[k] } (8 bytes each)
      uint4 o0, o1;
      o0.x = __byte_perm(a.x, b.x, 0x5410); o0.y = __byte_perm(c.x, d.x, 0x5410);
      o0.z = __byte_perm(a.x, b.x, 0x7632); o0.w = __byte_perm(c.x, d.x, 0x7632);
      o1.x = __byte_perm(a.y, b.y, 0x5410); o1.y = __byte_perm(c.y, d.y, 0x5410);
      o1.z = __byte_perm(a.y, b.y, 0x7632); o1.w = __byte_perm(c.y, d.y, 0x7632);
      uint4* dst = reinterpret_cast<uint4*>(static_cast<ushort4*>(pool_v) + quad);
      dst[0] = o0; dst[1] = o1;
    }
  }
}

// Pager gather: pulls the bricks of one batch out of PINNED host memory straight over PCIe (zero-copy loads,
// 16 bytes per thread, every SM issuing) into the staging area -- one launch instead of one cudaMemcpyAsync per
// brick, whose set-up cost is comparable to the 5 us a 256 KB brick takes on the wire.
struct GatherJob {
  const uint8_t* src;      // device alias of the pinned host address
  uint8_t* dst;            // staging
  uint64_t bytes;
};
__global__ void __launch_bounds__(256) gather_from_host_kernel(const GatherJob* __restrict__ jobs) {
  const GatherJob j = jobs[blockIdx.x];
  const uint64_t tid = (uint64_t)blockIdx.y * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.y * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(j.dst)) & 15) == 0) {
    const uint64_t nv = j.bytes / 16;
    const uint4* s = reinterpret_cast<const uint4*>(j.src);
    uint4* d = reinterpret_cast<uint4*>(j.dst);
    uint64_t i = tid;
    for (; i + 3 * nth < nv; i += 4 * nth) {            // four loads in flight per thread
      const uint4 a = __ldcs(s + i), b = __ldcs(s + i + nth), c = __ldcs(s + i + 2 * nth), e = __ldcs(s + i + 3 * nth);
      d[i] = a; d[i + nth] = b; d[i + 2 * nth] = c; d[i + 3 * nth] = e;
    }
    for (; i < nv; i += nth) d[i] = __ldcs(s + i);
    for (uint64_t k = nv * 16 + tid; k < j.bytes; k += nth) j.dst[k] = j.src[k];
  } else {
    for (uint64_t k = tid; k < j.bytes; k += nth) j.dst[k] = j.src[k];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) place_bricks_surface(const PlaceJob* __restrict__ jobs, cudaSurfaceObject_t surf,
                                                             const __grid_constant__ PoolConst pc) {
  const PlaceJob j = jobs[blockIdx.x];
  const uint32_t s0 = j.slot % pc.capacity[0], s1 = (j.slot / pc.capacity[0]) % pc.capacity[1],
                 s2 = j.slot / (pc.capacity[0] * pc.capacity[1]);
  const uint32_t total = j.sx * j.sy * j.sz;
  const T* src = reinterpret_cast<const T*>(j.src);
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const uint32_t x = i % j.sx, y = (i / j.sx) % j.sy, z = i / (j.sx * j.sy);
    surf3Dwrite(src[i], surf, (int)((s0 * pc.max_brick[0] + x) * sizeof(T)), (int)(s1 * pc.max_brick[1] + y),
                (int)(s2 * pc.max_brick[2] + z));
  }
}

// TVK_SAMPLER_GATHER: brick rows go into the layered array through surface stores; VEC bytes per store (16 when
// the row length, the source and the slot pitch allow it).  Slot (physical) = (p0, p1, p2): texel origin
// (p0, p1) * max_brick in every layer of the group p2 * max_brick.z ...
template <int VEC>
__global__ void __launch_bounds__(256) place_bricks_layered(const PlaceJob* __restrict__ jobs, cudaSurfaceObject_t surf,
                                                             const __grid_constant__ PoolConst pc) {
  const PlaceJob j = jobs[blockIdx.x];
  const uint32_t bpv = pc.bpv;
  const uint32_t p0 = j.slot % pc.phys_cap[0], p1 = (j.slot / pc.phys_cap[0]) % pc.phys_cap[1],
                 p2 = j.slot / (pc.phys_cap[0] * pc.phys_cap[1]);
  const uint32_t row_bytes = j.sx * bpv, units = row_bytes / VEC, rows = j.sy * j.sz, total = units * rows;
  const int x0 = (int)(p0 * pc.max_brick[0] * bpv), y0 = (int)(p1 * pc.max_brick[1]), l0 = (int)(p2 * pc.max_brick[2]);
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const uint32_t r = i / units, u = i - r * units;
    const uint32_t y = r % j.sy, z = r / j.sy;
    const uint8_t* s = j.src + (size_t)r * row_bytes + (size_t)u * VEC;
    if (VEC == 16) surf2DLayeredwrite(*reinterpret_cast<const uint4*>(s), surf, x0 + (int)(u * 16), y0 + (int)y, l0 + (int)z);
    else if (VEC == 2) surf2DLayeredwrite(*reinterpret_cast<const uint16_t*>(s), surf, x0 + (int)(u * 2), y0 + (int)y, l0 + (int)z);
    else surf2DLayeredwrite(*s, surf, x0 + (int)u, y0 + (int)y, l0 + (int)z);
  }
}

static int place_batch(tvk_ctx* c, std::vector<PlaceJob>& jobs, cudaStream_t st) {
  if (jobs.empty()) return TVK_OK;
  const uint32_t bpv = c->pc.bpv;
  // 16-byte accesses need 16 B aligned sources, rows AND pool slots (slot pitch = max_brick.x*bpv)
  const bool slot_aligned = ((c->pc.max_brick[0] * bpv) & 15) == 0;
  auto can_vec = [bpv, slot_aligned](const PlaceJob& j) {
    return slot_aligned && !(((uintptr_t)j.src & 15) || ((j.sx * bpv) & 15));
  };
  std::stable_partition(jobs.begin(), jobs.end(), can_vec);
  PlaceJob* d_jobs = nullptr;
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_jobs, sizeof(PlaceJob) * jobs.size(), st));
  if (c->pdesc.sampler != TVK_SAMPLER_EXACT_Q4) { int rc = tvk_copy_params_async(c, d_jobs, jobs.data(), sizeof(PlaceJob) * jobs.size(), st); if (rc) return rc; }
  const dim3 grid((unsigned)jobs.size(), 8);
  if (c->pdesc.sampler == TVK_SAMPLER_EXACT_Q4) {
    // jobs with rows of a multiple of four voxels (and a source aligned to four voxels) take the vector kernel
    auto q4_vec = [bpv, &c](const PlaceJob& j) {
      return (j.sx & 3) == 0 && (((uintptr_t)j.src) & (4 * bpv - 1)) == 0 && (c->pc.max_brick[0] & 3) == 0;
    };
    std::vector<PlaceJob> sorted(jobs);
    const size_t nv = std::stable_partition(sorted.begin(), sorted.end(), q4_vec) - sorted.begin();
    { int rc = tvk_copy_params_async(c, d_jobs, sorted.data(), sizeof(PlaceJob) * sorted.size(), st); if (rc) return rc; }
    if (nv) {
      if (bpv == 1) place_bricks_q4_vec<uint8_t><<<dim3((unsigned)nv, 8), 256, 0, st>>>(d_jobs, c->d_pool, c->pc);
      else place_bricks_q4_vec<uint16_t><<<dim3((unsigned)nv, 8), 256, 0, st>>>(d_jobs, c->d_pool, c->pc);
    }
    if (nv < sorted.size()) {
      const dim3 g2((unsigned)(sorted.size() - nv), 8);
      if (bpv == 1) place_bricks_q4<uint8_t, uchar4><<<g2, 256, 0, st>>>(d_jobs + nv, (uchar4*)c->d_pool, c->pc);
      else place_bricks_q4<uint16_t, ushort4><<<g2, 256, 0, st>>>(d_jobs + nv, (ushort4*)c->d_pool, c->pc);
      if (nv) c->launches++;
    }
  } else if (c->pdesc.sampler == TVK_SAMPLER_EXACT) {
    // jobs arrive sorted: [0, nvec) can move 16 B per access, the rest go byte-wise
    size_t nvec = 0;
    while (nvec < jobs.size() && can_vec(jobs[nvec])) ++nvec;
    if (nvec) place_bricks_linear<16><<<dim3((unsigned)nvec, 8), 256, 0, st>>>(d_jobs, (uint8_t*)c->d_pool, c->pc);
    if (nvec < jobs.size()) {
      place_bricks_linear<1><<<dim3((unsigned)(jobs.size() - nvec), 8), 256, 0, st>>>(d_jobs + nvec, (uint8_t*)c->d_pool, c->pc);
      c->launches++;
    }
  } else if (c->pdesc.sampler == TVK_SAMPLER_GATHER) {
    size_t nvec = 0;
    while (nvec < jobs.size() && can_vec(jobs[nvec])) ++nvec;
    if (nvec) place_bricks_layered<16><<<dim3((unsigned)nvec, 8), 256, 0, st>>>(d_jobs, c->pool_surf, c->pc);
    if (nvec < jobs.size()) {
      const dim3 g2((unsigned)(jobs.size() - nvec), 8);
      // element-sized stores: bricks are arrays of whole voxels, so a source is always bpv-aligned
      if (bpv == 2) place_bricks_layered<2><<<g2, 256, 0, st>>>(d_jobs + nvec, c->pool_surf, c->pc);
      else place_bricks_layered<1><<<g2, 256, 0, st>>>(d_jobs + nvec, c->pool_surf, c->pc);
      if (nvec) c->launches++;
    }
  } else {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->pool_array;
    cudaSurfaceObject_t surf = 0;
    TVK_CHECK_CUDA(c, cudaCreateSurfaceObject(&surf, &rd));
    if (c->pc.bpv == 1) place_bricks_surface<uint8_t><<<grid, 256, 0, st>>>(d_jobs, surf, c->pc);
    else place_bricks_surface<uint16_t><<<grid, 256, 0, st>>>(d_jobs, surf, c->pc);
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(st));
    cudaDestroySurfaceObject(surf);
  }
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaFreeAsync(d_jobs, st));
  return TVK_OK;
}

extern "C" int tvk_pool_upload_bricks(tvk_ctx* c, uint32_t n, const uint32_t* slot, const void* const* src,
                                      const uint32_t* sizes, int on_device) {
  if (!c || !c->has_pool || (n && (!slot || !src || !sizes))) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const uint32_t nslots = c->pc.capacity[0] * c->pc.capacity[1] * c->pc.capacity[2];
  for (uint32_t i = 0; i < n; ++i) {
    if (slot[i] >= nslots) TVK_FAIL(c, TVK_ERR_INVALID, "slot out of range");
    for (int a = 0; a < 3; ++a)
      if (sizes[3 * i + a] == 0 || sizes[3 * i + a] > c->pc.max_brick[a]) TVK_FAIL(c, TVK_ERR_INVALID, "brick larger than a pool slot");
  }
  cudaStream_t st = c->copy_stream;
  // the side stream must not overwrite slots a ray-march in flight still samples
  TVK_CHECK_CUDA(c, cudaEventRecord(c->render_done, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamWaitEvent(st, c->render_done, 0));
  if (on_device == 1) {
    std::vector<PlaceJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) jobs[i] = PlaceJob{(const uint8_t*)src[i], slot[i], sizes[3 * i], sizes[3 * i + 1], sizes[3 * i + 2]};
    int rc = place_batch(c, jobs, st);
    if (rc) return rc;
  } else {
    // staging ring in HBM: contiguous DMA per brick, then one placement launch per batch
    const size_t slot_bytes = (size_t)c->pc.max_brick[0] * c->pc.max_brick[1] * c->pc.max_brick[2] * c->pc.bpv;
    const uint32_t batch = 256;
    uint8_t* staging = nullptr;
    const uint32_t cap = std::min(n, batch);
    // two staging areas when there is more than one batch: the placement of batch b (place_stream) then runs beside
    // the PCIe gather of batch b+1
    const uint32_t nbuf = n > batch ? 2u : 1u;
    if (cap) TVK_CHECK_CUDA(c, cudaMallocAsync(&staging, slot_bytes * cap * nbuf, st));
    // zero-copy gather when every source is pinned, device-mapped host memory; else one DMA per (run of) brick(s)
    std::vector<const uint8_t*> alias(n, nullptr);
    bool mapped = true;
    if (on_device == 2 && n) {
      // one allocation (caller's promise): the device alias of every source is the host pointer plus one offset
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, src[0]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        const ptrdiff_t delta = static_cast<const uint8_t*>(at.devicePointer) - static_cast<const uint8_t*>(src[0]);
        for (uint32_t i = 0; i < n; ++i) alias[i] = static_cast<const uint8_t*>(src[i]) + delta;
      } else { mapped = false; (void)cudaGetLastError(); }
    } else
    for (uint32_t i = 0; i < n && mapped; ++i) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, src[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) { mapped = false; (void)cudaGetLastError(); }
      else alias[i] = static_cast<const uint8_t*>(at.devicePointer);
    }
    for (uint32_t b0 = 0; b0 < n; b0 += batch) {
      const uint32_t bn = std::min(batch, n - b0);
      std::vector<PlaceJob> jobs(bn);
      if (mapped) {
        const uint32_t bi = b0 / batch, buf = nbuf == 2 ? (bi & 1u) : 0u;
        uint8_t* stg = staging + (size_t)buf * slot_bytes * cap;
        if (nbuf == 2 && bi >= 2) TVK_CHECK_CUDA(c, cudaStreamWaitEvent(st, c->ev_place[buf], 0));   // batch b-2 has left this area
        // A batch whose bricks follow one another in host memory (a bulk load in ToC order; the store aligns bricks to
        // 256 bytes) goes out as ONE copy-engine transfer, which gets closer to the link rate than SM-issued reads
        // (55 vs 49 GB/s); scattered bricks -- what paging on demand asks for -- are gathered by the kernel.
        // A batch that is only PARTLY contiguous -- the non-empty bricks of a bulk load, which come in rows along x --
        // goes out as one transfer per run when the runs are long enough (three bricks on average), so that the
        // upload stays on the copy engine and off the SMs the ray-march kernel is using.
        auto brick_bytes = [&](uint32_t k) { return (size_t)sizes[3 * k] * sizes[3 * k + 1] * sizes[3 * k + 2] * c->pc.bpv; };
        struct Run { uint32_t first, count; size_t span; };
        std::vector<Run> runs;
        for (uint32_t i = 0; i < bn;) {
          uint32_t j = i + 1;
          while (j < bn) {
            const uint32_t k = b0 + j;
            const ptrdiff_t gap = static_cast<const uint8_t*>(src[k]) - (static_cast<const uint8_t*>(src[k - 1]) + brick_bytes(k - 1));
            if (gap < 0 || gap >= 256) break;
            ++j;
          }
          const uint32_t kl = b0 + j - 1;
          runs.push_back(Run{i, j - i, (size_t)(static_cast<const uint8_t*>(src[kl]) - static_cast<const uint8_t*>(src[b0 + i])) + brick_bytes(kl)});
          i = j;
        }
        size_t staged = 0;
        for (const Run& r : runs) staged += (r.span + 255u) & ~(size_t)255u;
        const bool by_runs = bn >= 16 && runs.size() * 3u <= bn && staged <= slot_bytes * cap;
        if (by_runs) {
          size_t at = 0;
          for (const Run& r : runs) {
            const uint8_t* first = static_cast<const uint8_t*>(src[b0 + r.first]);
            TVK_CHECK_CUDA(c, cudaMemcpyAsync(stg + at, first, r.span, cudaMemcpyHostToDevice, st));
            for (uint32_t i = r.first; i < r.first + r.count; ++i) {
              const uint32_t k = b0 + i;
              jobs[i] = PlaceJob{stg + at + (static_cast<const uint8_t*>(src[k]) - first), slot[k], sizes[3 * k], sizes[3 * k + 1], sizes[3 * k + 2]};
            }
            at += (r.span + 255u) & ~(size_t)255u;
          }
        } else {
          std::vector<GatherJob> gj(bn);
          for (uint32_t i = 0; i < bn; ++i) {
            const uint32_t k = b0 + i;
            gj[i] = GatherJob{alias[k], stg + slot_bytes * i, (uint64_t)sizes[3 * k] * sizes[3 * k + 1] * sizes[3 * k + 2] * c->pc.bpv};
            jobs[i] = PlaceJob{stg + slot_bytes * i, slot[k], sizes[3 * k], sizes[3 * k + 1], sizes[3 * k + 2]};
          }
          GatherJob* d_gj = nullptr;
          TVK_CHECK_CUDA(c, cudaMallocAsync(&d_gj, sizeof(GatherJob) * bn, st));
          { int rc = tvk_copy_params_async(c, d_gj, gj.data(), sizeof(GatherJob) * bn, st); if (rc) return rc; }
          gather_from_host_kernel<<<dim3(bn, 8), 256, 0, st>>>(d_gj);
          c->launches++;
          TVK_CHECK_CUDA(c, cudaGetLastError());
          TVK_CHECK_CUDA(c, cudaFreeAsync(d_gj, st));
        }
        if (nbuf == 2) {
          TVK_CHECK_CUDA(c, cudaEventRecord(c->ev_gather[buf], st));
          TVK_CHECK_CUDA(c, cudaStreamWaitEvent(c->place_stream, c->ev_gather[buf], 0));
          int rc = place_batch(c, jobs, c->place_stream);
          if (rc) return rc;
          TVK_CHECK_CUDA(c, cudaEventRecord(c->ev_place[buf], c->place_stream));
          if (b0 + batch >= n) {                             // last batch: the side stream ends after both placements in flight
            TVK_CHECK_CUDA(c, cudaStreamWaitEvent(st, c->ev_place[buf], 0));
            TVK_CHECK_CUDA(c, cudaStreamWaitEvent(st, c->ev_place[buf ^ 1u], 0));
          }
        } else {
          int rc = place_batch(c, jobs, st);
          if (rc) return rc;
        }
        continue;
      }
      for (uint32_t i = 0; i < bn;) {
        // full-size bricks that are contiguous in host memory go out as ONE DMA
        uint32_t run = 1;
        size_t bytes = (size_t)sizes[3 * (b0 + i)] * sizes[3 * (b0 + i) + 1] * sizes[3 * (b0 + i) + 2] * c->pc.bpv;
        if (bytes == slot_bytes)
          while (i + run < bn &&
                 (const uint8_t*)src[b0 + i + run] == (const uint8_t*)src[b0 + i] + slot_bytes * run &&
                 (size_t)sizes[3 * (b0 + i + run)] * sizes[3 * (b0 + i + run) + 1] * sizes[3 * (b0 + i + run) + 2] * c->pc.bpv == slot_bytes)
            ++run;
        if (run > 1) bytes = slot_bytes * run;
        TVK_CHECK_CUDA(c, cudaMemcpyAsync(staging + slot_bytes * i, src[b0 + i], bytes, cudaMemcpyHostToDevice, st));
        for (uint32_t r = 0; r < run; ++r) {
          const uint32_t k = b0 + i + r;
          jobs[i + r] = PlaceJob{staging + slot_bytes * (i + r), slot[k], sizes[3 * k], sizes[3 * k + 1], sizes[3 * k + 2]};
        }
        i += run;
      }
      int rc = place_batch(c, jobs, st);
      if (rc) return rc;
    }
    if (staging) TVK_CHECK_CUDA(c, cudaFreeAsync(staging, st));
  }
  TVK_CHECK_CUDA(c, cudaEventRecord(c->copy_done, st));
  return TVK_OK;
}

// ------------------------------------------------------------------ LZ4 bricks inflated on the device
// ExtendedOctree::GetBrickData inflates a compressed brick on the host (ExtendedOctree.cpp:313-360, lz4Decompress ->
// LZ4_decompress_fast: raw LZ4 BLOCK format, decompressed size known from the ToC).  Here the COMPRESSED bytes cross
// PCIe -- the link is what bounds paging -- and one warp per brick inflates them in HBM: every lane parses the sequence
// header (token, length bytes, offset: uniform, broadcast loads), the 32 lanes share the literal copy and the match
// copy.  A match whose offset is shorter than 32 bytes repeats a pattern that is already complete (out[i] = m[i % off]), a
// longer one is copied 32 bytes per trip with a warp barrier between trips, so that no lane reads a byte another lane of
// the same trip writes.  Malformed input (a length or offset that leaves the buffers) raises the error flag.
struct Lz4Job { const uint8_t* src; uint8_t* dst; uint32_t src_len, dst_len; };

__global__ void __launch_bounds__(128) lz4_inflate_kernel(const Lz4Job* __restrict__ jobs, uint32_t njobs, uint32_t* __restrict__ err) {
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (warp >= njobs) return;
  const Lz4Job j = jobs[warp];
  const uint8_t* ip = j.src;
  const uint8_t* const iend = j.src + j.src_len;
  uint8_t* op = j.dst;
  uint8_t* const oend = j.dst + j.dst_len;
  bool bad = false;
  for (;;) {
    if (ip >= iend) { bad = true; break; }
    const uint32_t token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) { uint32_t b; do { if (ip >= iend) { bad = true; break; } b = *ip++; lit += b; } while (b == 255); if (bad) break; }
    if (lit > (size_t)(iend - ip) || lit > (size_t)(oend - op)) { bad = true; break; }
    for (size_t i = lane; i < lit; i += 32) op[i] = ip[i];
    op += lit; ip += lit;
    if (op == oend) break;                               // the last sequence has literals only
    if (iend - ip < 2) { bad = true; break; }
    const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
    ip += 2;
    if (off == 0 || off > (size_t)(op - j.dst)) { bad = true; break; }
    size_t ml = token & 15u;
    if (ml == 15) { uint32_t b; do { if (ip >= iend) { bad = true; break; } b = *ip++; ml += b; } while (b == 255); if (bad) break; }
    ml += 4;
    if (ml > (size_t)(oend - op)) { bad = true; break; }
    const uint8_t* m = op - off;
    __syncwarp();                                        // the literals just written may be the match source
    if (off < 32) {
      for (size_t i = lane; i < ml; i += 32) op[i] = m[i % off];
    } else {
      for (size_t i0 = 0; i0 < ml; i0 += 32) {
        if (i0 + lane < ml) op[i0 + lane] = m[i0 + lane];
        __syncwarp();
      }
    }
    op += ml;
    __syncwarp();
  }
  if (bad && lane == 0) atomicExch(err, 1u);
}

// ------------------------------------------------------------------ zlib bricks inflated on the device
// A brick stored CT_ZLIB is one zlib stream (RFC 1950: 2 header bytes, DEFLATE blocks, Adler-32; zCompress / zDecompress,
// ExtendedOctree/ZlibCompression.cpp:14-70 -- the reference inflates it on the host).  One warp per brick, same job
// layout as the LZ4 kernel.  DEFLATE (RFC 1951) is a serial bit stream: EVERY lane reads it (private, identical bit
// position; the loads are broadcasts), so control flow stays uniform without any shuffle; lane 0 alone writes literals
// and the shared-memory Huffman tables (canonical codes decoded by counting, one bit at a time); the 32 lanes share the
// copy of every match and of stored blocks, as in the LZ4 kernel.  The Adler-32 trailer is not checked (the UVF carries an
// MD5 of the whole file).  Throughput is that of a serial decoder per brick (a batch of bricks inflates in parallel); what it
// buys is that the compressed bytes, not the inflated ones, sit in pinned memory and cross PCIe.
__constant__ unsigned short lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ unsigned short lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ unsigned short dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ unsigned short dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ unsigned char zorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
#define TVK_Z_LITBITS 10
#define TVK_Z_DISTBITS 8
struct ZTables {
  unsigned short lencnt[16], lensym[288], distcnt[16], distsym[32];
  unsigned char lengths[320];
  // first-level look-up: the next TVK_Z_*BITS bits of the stream -> (code length << 9) | symbol, 0 = a longer code (decoded by counting)
  unsigned short litfast[1 << TVK_Z_LITBITS], distfast[1 << TVK_Z_DISTBITS];
};
struct ZBits { const uint8_t* ip; const uint8_t* iend; unsigned long long buf; int cnt; bool bad; };

// more than 32 bits in the buffer (zeros behind the end of the input: reading them is an error only once they are CONSUMED)
__device__ __forceinline__ void z_fill(ZBits& z) {
  if (z.cnt <= 32) {                                       // four bytes at a time: one load latency per ~3 symbols, not per symbol
    uint32_t w;
    if (z.iend - z.ip >= 4) w = (uint32_t)z.ip[0] | ((uint32_t)z.ip[1] << 8) | ((uint32_t)z.ip[2] << 16) | ((uint32_t)z.ip[3] << 24);
    else {
      w = 0u;
      for (int k = 0; k < 4; ++k) if (z.ip + k < z.iend) w |= (uint32_t)z.ip[k] << (8 * k);
    }
    z.ip += 4;
    z.buf |= (unsigned long long)w << z.cnt;
    z.cnt += 32;
  }
}
__device__ __forceinline__ void z_drop(ZBits& z, int n) {
  z.buf >>= n; z.cnt -= n;
  // bits taken from behind the end of the stream: ip ran past iend by more bytes than the buffer still holds
  if (z.ip > z.iend && (long long)(z.ip - z.iend) * 8 > (long long)z.cnt) z.bad = true;
}
__device__ __forceinline__ uint32_t z_bits(ZBits& z, int need) {             // need <= 16
  z_fill(z);
  const uint32_t v = (uint32_t)z.buf & ((1u << need) - 1u);
  z_drop(z, need);
  return v;
}
// Canonical Huffman decode by counting (one bit per length), with the 15 per-length counts in REGISTERS (packed two per word,
// the loop is unrolled so that every access has a constant index): the only shared-memory access is the symbol fetch.
struct ZCounts { uint32_t w[8]; };
__device__ __forceinline__ void z_load_counts(ZCounts& c, const unsigned short* cnt) {
#pragma unroll
  for (int k = 0; k < 8; ++k) c.w[k] = (uint32_t)cnt[2 * k] | ((uint32_t)cnt[2 * k + 1] << 16);
}
__device__ __forceinline__ int z_decode(ZBits& z, const ZCounts& c, const unsigned short* sym) {
  z_fill(z);
  uint32_t bits = (uint32_t)z.buf;
  int code = 0, first = 0, index = 0;
#pragma unroll
  for (int len = 1; len <= 15; ++len) {
    code |= (int)(bits & 1u);
    bits >>= 1;
    const int count = (int)((c.w[len >> 1] >> (16 * (len & 1))) & 0xFFFFu);
    if (code - count < first) { z_drop(z, len); return z.bad ? -1 : (int)sym[index + (code - first)]; }
    index += count; first += count; first <<= 1; code <<= 1;
  }
  z.bad = true;
  return -1;
}
// the look-up table of one code: every lane decodes its share of the 2^nbits bit patterns by counting
__device__ __forceinline__ void z_build_fast(unsigned short* fast, int nbits, const ZCounts& c, const unsigned short* sym, uint32_t lane) {
  for (uint32_t e = lane; e < (1u << nbits); e += 32u) {
    uint32_t bits = e;
    int code = 0, first = 0, index = 0;
    unsigned short out = 0;
#pragma unroll
    for (int len = 1; len <= 15; ++len) {
      if (len <= nbits && !out) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int count = (int)((c.w[len >> 1] >> (16 * (len & 1))) & 0xFFFFu);
        if (code - count < first) out = (unsigned short)((len << 9) | sym[index + (code - first)]);
        index += count; first += count; first <<= 1; code <<= 1;
      }
    }
    fast[e] = out;
  }
}
__device__ __forceinline__ int z_decode_fast(ZBits& z, const unsigned short* fast, int nbits, const ZCounts& c, const unsigned short* sym) {
  z_fill(z);
  const uint32_t e = fast[(uint32_t)z.buf & ((1u << nbits) - 1u)];
  if (e) { z_drop(z, (int)(e >> 9)); return z.bad ? -1 : (int)(e & 511u); }
  return z_decode(z, c, sym);
}
// canonical Huffman table from code lengths (lane 0 only); returns < 0 for an over-subscribed set
__device__ int z_construct(unsigned short* cnt, unsigned short* sym, const unsigned char* length, int n) {
  unsigned short offs[16];
  for (int l = 0; l <= 15; ++l) cnt[l] = 0;
  for (int s = 0; s < n; ++s) cnt[length[s]]++;
  if (cnt[0] == n) return 0;
  int left = 1;
  for (int l = 1; l <= 15; ++l) { left <<= 1; left -= cnt[l]; if (left < 0) return left; }
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + cnt[l];
  for (int s = 0; s < n; ++s) if (length[s]) sym[offs[length[s]]++] = (unsigned short)s;
  return left;
}

__global__ void __launch_bounds__(128) zlib_inflate_kernel(const Lz4Job* __restrict__ jobs, uint32_t njobs, uint32_t* __restrict__ err) {
  __shared__ ZTables tables[4];
  const uint32_t wib = threadIdx.x >> 5, warp = blockIdx.x * (blockDim.x >> 5) + wib, lane = threadIdx.x & 31u;
  if (warp >= njobs) return;
  ZTables& T = tables[wib];
  const Lz4Job j = jobs[warp];
  uint8_t* op = j.dst;
  uint8_t* const oend = j.dst + j.dst_len;
  ZBits z;
  z.ip = j.src; z.iend = j.src + j.src_len; z.buf = 0ull; z.cnt = 0; z.bad = false;
  bool bad = false;
  // zlib header: CM = 8 (deflate), no preset dictionary, (CMF * 256 + FLG) % 31 == 0
  if (j.src_len < 2 || (j.src[0] & 0x0F) != 8 || (j.src[1] & 0x20) || ((((uint32_t)j.src[0]) << 8) | j.src[1]) % 31u) bad = true;
  z.ip += 2;
  uint32_t last = 0;
  while (!bad && !last) {
    last = z_bits(z, 1);
    const uint32_t type = z_bits(z, 2);
    if (z.bad || type == 3) { bad = true; break; }
    if (type == 0) {                                       // stored
      // back to the byte boundary: drop the bits left of the current byte, then hand the whole bytes still buffered back
      z_drop(z, z.cnt & 7);
      z.ip -= z.cnt >> 3; z.buf = 0ull; z.cnt = 0;
      if (z.bad || z.ip > z.iend || z.iend - z.ip < 4) { bad = true; break; }
      const uint32_t len = (uint32_t)z.ip[0] | ((uint32_t)z.ip[1] << 8), nlen = (uint32_t)z.ip[2] | ((uint32_t)z.ip[3] << 8);
      z.ip += 4;
      if (len != (~nlen & 0xFFFFu) || len > (size_t)(z.iend - z.ip) || len > (size_t)(oend - op)) { bad = true; break; }
      for (uint32_t i = lane; i < len; i += 32) op[i] = z.ip[i];
      op += len; z.ip += len;
      __syncwarp();
      continue;
    }
    __syncwarp();                                          // nobody still reads the previous block's tables
    if (type == 1) {                                       // fixed codes
      if (lane == 0) {
        int s = 0;
        for (; s < 144; ++s) T.lengths[s] = 8;
        for (; s < 256; ++s) T.lengths[s] = 9;
        for (; s < 280; ++s) T.lengths[s] = 7;
        for (; s < 288; ++s) T.lengths[s] = 8;
        z_construct(T.lencnt, T.lensym, T.lengths, 288);
        for (s = 0; s < 30; ++s) T.lengths[s] = 5;
        z_construct(T.distcnt, T.distsym, T.lengths, 30);
      }
    } else {                                               // dynamic codes
      const int nlen = (int)z_bits(z, 5) + 257, ndist = (int)z_bits(z, 5) + 1, ncode = (int)z_bits(z, 4) + 4;
      if (z.bad || nlen > 286 || ndist > 30) { bad = true; break; }
      unsigned char cl[19];
      for (int i = 0; i < 19; ++i) cl[i] = 0;
      for (int i = 0; i < ncode; ++i) cl[zorder[i]] = (unsigned char)z_bits(z, 3);
      if (z.bad) { bad = true; break; }
      int ok = 0;
      if (lane == 0) ok = z_construct(T.lencnt, T.lensym, cl, 19);       // the code-length code, in the literal/length table's place
      ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
      if (ok != 0) { bad = true; break; }                  // must be complete
      __syncwarp();
      int index = 0;
      int prev = 0;
      bool hdr_bad = false;
      ZCounts cc;
      z_load_counts(cc, T.lencnt);
      while (index < nlen + ndist) {
        int sym = z_decode(z, cc, T.lensym);
        if (sym < 0) { hdr_bad = true; break; }
        if (sym < 16) { if (lane == 0) T.lengths[index] = (unsigned char)sym; prev = sym; ++index; }
        else {
          int len = 0, rep;
          if (sym == 16) { if (index == 0) { hdr_bad = true; break; } len = prev; rep = 3 + (int)z_bits(z, 2); }
          else if (sym == 17) rep = 3 + (int)z_bits(z, 3);
          else rep = 11 + (int)z_bits(z, 7);
          if (z.bad || index + rep > nlen + ndist) { hdr_bad = true; break; }
          if (lane == 0) for (int k = 0; k < rep; ++k) T.lengths[index + k] = (unsigned char)len;
          index += rep;
          prev = len;
        }
      }
      if (hdr_bad) { bad = true; break; }
      __syncwarp();                                        // lengths complete; the code-length table is no longer read
      int e1 = 0, e2 = 0, eob = 0;
      if (lane == 0) {
        eob = T.lengths[256];
        e1 = z_construct(T.lencnt, T.lensym, T.lengths, nlen);
        e2 = z_construct(T.distcnt, T.distsym, T.lengths + nlen, ndist);
        // incomplete codes are allowed only for a single code of length one (RFC 1951 / zlib)
        if (e1 < 0 || (e1 > 0 && nlen - T.lencnt[0] != 1)) e1 = -1; else e1 = 0;
        if (e2 < 0 || (e2 > 0 && ndist - T.distcnt[0] != 1)) e2 = -1; else e2 = 0;
        if (eob == 0) e1 = -1;
      }
      e1 = __shfl_sync(0xFFFFFFFFu, e1 | e2, 0);
      if (e1 != 0) { bad = true; break; }
    }
    __syncwarp();
    ZCounts lc, dc;
    z_load_counts(lc, T.lencnt);
    z_load_counts(dc, T.distcnt);
    z_build_fast(T.litfast, TVK_Z_LITBITS, lc, T.lensym, lane);
    z_build_fast(T.distfast, TVK_Z_DISTBITS, dc, T.distsym, lane);
    __syncwarp();
    for (;;) {                                             // the block's symbols
      const int sym = z_decode_fast(z, T.litfast, TVK_Z_LITBITS, lc, T.lensym);
      if (sym < 0) { bad = true; break; }
      if (sym < 256) {
        if (op >= oend) { bad = true; break; }
        if (lane == 0) *op = (uint8_t)sym;
        ++op;
      } else if (sym == 256) {
        break;
      } else {
        const int ls = sym - 257;
        if (ls >= 29) { bad = true; break; }
        const uint32_t ml = lbase[ls] + z_bits(z, lext[ls]);
        const int ds = z_decode_fast(z, T.distfast, TVK_Z_DISTBITS, dc, T.distsym);
        if (ds < 0 || ds >= 30) { bad = true; break; }
        const size_t off = dbase[ds] + z_bits(z, dext[ds]);
        if (z.bad || off > (size_t)(op - j.dst) || ml > (size_t)(oend - op)) { bad = true; break; }
        const uint8_t* m = op - off;
        __syncwarp();                                      // the literals lane 0 wrote may be the match source
        if (off < 32) {
          for (uint32_t i = lane; i < ml; i += 32) op[i] = m[i % off];
        } else {
          for (uint32_t i0 = 0; i0 < ml; i0 += 32) {
            if (i0 + lane < ml) op[i0 + lane] = m[i0 + lane];
            __syncwarp();
          }
        }
        op += ml;
        __syncwarp();
      }
    }
  }
  if (!bad && op != oend) bad = true;                      // the ToC's brick size is the inflated size
  if (bad && lane == 0) atomicExch(err, 1u);
}

static int upload_bricks_compressed(tvk_ctx* c, uint32_t codec, uint32_t n, const uint32_t* slot, const void* const* src, const uint32_t* comp_len,
                                    const uint32_t* sizes);
extern "C" int tvk_pool_upload_bricks_lz4(tvk_ctx* c, uint32_t n, const uint32_t* slot, const void* const* src, const uint32_t* comp_len,
                                          const uint32_t* sizes) {
  return upload_bricks_compressed(c, TVK_CODEC_LZ4, n, slot, src, comp_len, sizes);
}
extern "C" int tvk_pool_upload_bricks_compressed(tvk_ctx* c, uint32_t codec, uint32_t n, const uint32_t* slot, const void* const* src,
                                                 const uint32_t* comp_len, const uint32_t* sizes) {
  if (codec != TVK_CODEC_LZ4 && codec != TVK_CODEC_ZLIB) return TVK_ERR_INVALID;
  return upload_bricks_compressed(c, codec, n, slot, src, comp_len, sizes);
}
static int upload_bricks_compressed(tvk_ctx* c, uint32_t codec, uint32_t n, const uint32_t* slot, const void* const* src, const uint32_t* comp_len,
                                    const uint32_t* sizes) {
  if (!c || !c->has_pool || (n && (!slot || !src || !comp_len || !sizes))) return TVK_ERR_INVALID;
  if (!n) return TVK_OK;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t st = c->copy_stream;
  // staging: [compressed bricks, 16-byte aligned][inflated bricks, 256-byte aligned][jobs][error flag]
  std::vector<size_t> coff(n), doff(n);
  size_t cbytes = 0, dbytes = 0;
  for (uint32_t i = 0; i < n; ++i) {
    coff[i] = cbytes; cbytes += ((size_t)comp_len[i] + 15) & ~(size_t)15;
    doff[i] = dbytes; dbytes += (((size_t)sizes[3 * i] * sizes[3 * i + 1] * sizes[3 * i + 2] * c->pc.bpv) + 255) & ~(size_t)255;
  }
  uint8_t* d_comp = nullptr; uint8_t* d_out = nullptr; Lz4Job* d_jobs = nullptr; uint32_t* d_err = nullptr;
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_comp, cbytes + 16, st));
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_out, dbytes, st));
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_jobs, sizeof(Lz4Job) * n, st));
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_err, sizeof(uint32_t), st));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_err, 0, sizeof(uint32_t), st));
  std::vector<Lz4Job> jobs(n);
  std::vector<const void*> dsrc(n);
  for (uint32_t i = 0; i < n; ) {
    // bricks that follow one another in host memory (a ToC-order bulk load) travel as one transfer
    uint32_t k = i + 1;
    while (k < n && static_cast<const uint8_t*>(src[k]) == static_cast<const uint8_t*>(src[k - 1]) + comp_len[k - 1] &&
           coff[k] == coff[k - 1] + comp_len[k - 1]) ++k;
    size_t span = 0;
    for (uint32_t q = i; q < k; ++q) span += comp_len[q];
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_comp + coff[i], src[i], span, cudaMemcpyHostToDevice, st));
    i = k;
  }
  for (uint32_t i = 0; i < n; ++i) {
    jobs[i].src = d_comp + coff[i]; jobs[i].dst = d_out + doff[i]; jobs[i].src_len = comp_len[i];
    jobs[i].dst_len = sizes[3 * i] * sizes[3 * i + 1] * sizes[3 * i + 2] * c->pc.bpv;
    dsrc[i] = jobs[i].dst;
  }
  { int rc = tvk_copy_params_async(c, d_jobs, jobs.data(), sizeof(Lz4Job) * n, st); if (rc) return rc; }
  if (codec == TVK_CODEC_ZLIB) zlib_inflate_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_jobs, n, d_err);
  else lz4_inflate_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_jobs, n, d_err);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  uint32_t h_err = 0;
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(&h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(st));          // the inflated bricks are placement sources from here on
  c->lz4_bytes_in += cbytes; c->lz4_bytes_out += dbytes;
  int rc = TVK_OK;
  if (h_err) { c->err = codec == TVK_CODEC_ZLIB ? "corrupt zlib brick" : "corrupt LZ4 brick"; rc = TVK_ERR_INVALID; }
  else rc = tvk_pool_upload_bricks(c, n, slot, dsrc.data(), sizes, 1);
  if (rc == TVK_OK) rc = tvk_pool_upload_wait(c);
  cudaFreeAsync(d_comp, st); cudaFreeAsync(d_out, st); cudaFreeAsync(d_jobs, st); cudaFreeAsync(d_err, st);
  return rc;
}

// the same kernel without a pool: n LZ4 blocks (host pointers) -> their inflated bytes in host memory (tests, tools)
static int inflate_blocks(tvk_ctx* c, uint32_t codec, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len);
extern "C" int tvk_lz4_inflate(tvk_ctx* c, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len) {
  return inflate_blocks(c, TVK_CODEC_LZ4, n, src, comp_len, dst, dst_len);
}
extern "C" int tvk_zlib_inflate(tvk_ctx* c, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len) {
  return inflate_blocks(c, TVK_CODEC_ZLIB, n, src, comp_len, dst, dst_len);
}
static int inflate_blocks(tvk_ctx* c, uint32_t codec, uint32_t n, const void* const* src, const uint32_t* comp_len, void* const* dst, const uint32_t* dst_len) {
  if (!c || (n && (!src || !comp_len || !dst || !dst_len))) return TVK_ERR_INVALID;
  if (!n) return TVK_OK;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t st = c->copy_stream;
  std::vector<size_t> coff(n), doff(n);
  size_t cbytes = 0, dbytes = 0;
  for (uint32_t i = 0; i < n; ++i) { coff[i] = cbytes; cbytes += ((size_t)comp_len[i] + 15) & ~(size_t)15; doff[i] = dbytes; dbytes += ((size_t)dst_len[i] + 255) & ~(size_t)255; }
  uint8_t* d_comp = nullptr; uint8_t* d_out = nullptr; Lz4Job* d_jobs = nullptr; uint32_t* d_err = nullptr;
  TVK_CHECK_CUDA(c, cudaMalloc(&d_comp, cbytes + 16));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_out, dbytes));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_jobs, sizeof(Lz4Job) * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_err, sizeof(uint32_t)));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_err, 0, sizeof(uint32_t), st));
  std::vector<Lz4Job> jobs(n);
  for (uint32_t i = 0; i < n; ++i) {
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_comp + coff[i], src[i], comp_len[i], cudaMemcpyHostToDevice, st));
    jobs[i].src = d_comp + coff[i]; jobs[i].dst = d_out + doff[i]; jobs[i].src_len = comp_len[i]; jobs[i].dst_len = dst_len[i];
  }
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(Lz4Job) * n, cudaMemcpyHostToDevice, st));
  if (codec == TVK_CODEC_ZLIB) zlib_inflate_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_jobs, n, d_err);
  else lz4_inflate_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_jobs, n, d_err);
  c->launches++;
  uint32_t h_err = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
  for (uint32_t i = 0; i < n && e == cudaSuccess; ++i) e = cudaMemcpyAsync(dst[i], d_out + doff[i], dst_len[i], cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_comp); cudaFree(d_out); cudaFree(d_jobs); cudaFree(d_err);
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return TVK_ERR_CUDA; }
  if (h_err) { c->err = codec == TVK_CODEC_ZLIB ? "corrupt zlib stream" : "corrupt LZ4 block"; return TVK_ERR_INVALID; }
  return TVK_OK;
}

extern "C" int tvk_pool_upload_fence(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  // device-side ordering only: work issued on the render stream from now on sees the placed bricks; the host does not wait
  TVK_CHECK_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
  return TVK_OK;
}

extern "C" int tvk_pool_upload_wait(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  // later work on the render stream sees the placed bricks; the host may reuse its buffers
  TVK_CHECK_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  return TVK_OK;
}

// ---------------------------------------------------------------------------
// metadata
// ---------------------------------------------------------------------------
__global__ void patch_meta_kernel(uint32_t* __restrict__ meta, const uint32_t* __restrict__ idx,
                                  const uint32_t* __restrict__ val, uint32_t n, uint32_t total) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] < total) meta[idx[i]] = val[i];
}

extern "C" int tvk_pool_set_metadata(tvk_ctx* c, const uint32_t* meta, uint32_t n) {
  if (!c || !c->has_pool || !meta || n != c->pc.total_bricks) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->d_meta, meta, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

extern "C" int tvk_pool_clear_metadata(tvk_ctx* c) {
  if (!c || !c->has_pool) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  static_assert(TVK_BI_MISSING == 0u, "the clear is a memset");
  TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_meta, 0, sizeof(uint32_t) * c->pc.total_bricks, c->stream));
  return TVK_OK;
}

extern "C" int tvk_pool_patch_metadata(tvk_ctx* c, uint32_t n, const uint32_t* brick_index, const uint32_t* value) {
  if (!c || !c->has_pool || (n && (!brick_index || !value))) return TVK_ERR_INVALID;
  if (!n) return TVK_OK;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  uint32_t* d = nullptr;
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d, sizeof(uint32_t) * 2 * n, c->stream));
  { int rc = tvk_copy_params_async(c, d, brick_index, sizeof(uint32_t) * n, c->stream); if (rc) return rc; }
  { int rc = tvk_copy_params_async(c, d + n, value, sizeof(uint32_t) * n, c->stream); if (rc) return rc; }
  patch_meta_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_meta, d, d + n, n, c->pc.total_bricks);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaFreeAsync(d, c->stream));
  // no host synchronisation: both arrays have been copied into the pinned parameter ring (or staged by the runtime), so
  // the caller may reuse them; the patch itself is ordered on the render stream
  return TVK_OK;
}

extern "C" int tvk_pool_get_metadata(tvk_ctx* c, uint32_t* meta, uint32_t n) {
  if (!c || !c->has_pool || !meta || n != c->pc.total_bricks) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(meta, c->d_meta, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

// ---------------------------------------------------------------------------
// transfer functions (RGBA8, fetched NEAREST in the kernels)
// ---------------------------------------------------------------------------
// per value column of a 2-D table: does ANY gradient row give the sample a non-zero alpha?  (the 2D modes skip the
// gradient where not)  One warp per 32 columns, lanes read adjacent entries of a row.
__global__ void tf_colmask_kernel(const uchar4* __restrict__ tf, uint32_t w, uint32_t h, uint32_t* __restrict__ mask) {
  const uint32_t x = blockIdx.x * 32 + threadIdx.x;
  bool any = false;
  if (x < w)
    for (uint32_t y = 0; y < h; ++y) any |= (tf[(size_t)y * w + x].w != 0);
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, any);
  if (threadIdx.x == 0) mask[blockIdx.x] = m;
}

// 1-D table: the float copies the march fetches (byte * (1/255); and the same with the opacity already corrected for
// sampleRateModifier == 1: 1 - (1 - a), which is not always a in fp32 -- OpacityCorrectColor, blend.glsl:37-39, evaluated
// once per table entry instead of once per sample), made on the device from the bytes
__global__ void __launch_bounds__(256) tf1d_expand_kernel(const uchar4* __restrict__ tf, uint32_t n, float4* __restrict__ tf_f,
                                                         float4* __restrict__ tf_f1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uchar4 b = tf[i];
  const float k = 1.0f / 255.0f;
  const float4 f = make_float4((float)b.x * k, (float)b.y * k, (float)b.z * k, (float)b.w * k);
  tf_f[i] = f;
  const float om = __fsub_rn(1.0f, f.w);
  tf_f1[i] = make_float4(f.x, f.y, f.z, __fsub_rn(1.0f, om));
}

static int set_tf(tvk_ctx* c, const uint8_t* rgba, uint32_t w, uint32_t h) {
  if (!c || !rgba || w == 0 || h == 0) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)w * h;
  if (n != (size_t)c->tf_w * c->tf_h || (h == 1) != (c->d_tf_f != nullptr)) {
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));            // a march in flight may still read the old tables
    cudaFree(c->d_tf); c->d_tf = nullptr;
    cudaFree(c->d_tf_f); c->d_tf_f = nullptr;
    cudaFree(c->d_tf_f1); c->d_tf_f1 = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_tf, sizeof(uchar4) * n));
    // float copy with the kernel's own conversion (byte * (1/255)): one 16-byte fetch per sample -- 1D tables only
    // (the 2D modes fetch the byte table; a 4096x256 table would be 16 MB more to convert and upload per edit)
    if (h == 1) { TVK_CHECK_CUDA(c, cudaMalloc(&c->d_tf_f, sizeof(float4) * n)); TVK_CHECK_CUDA(c, cudaMalloc(&c->d_tf_f1, sizeof(float4) * n)); }
  }
  if (((w + 31) / 32) != ((c->tf_w + 31) / 32) || !c->d_tf_colmask) {
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_tf_colmask); c->d_tf_colmask = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_tf_colmask, sizeof(uint32_t) * ((w + 31) / 32)));
  }
  c->tf_w = w; c->tf_h = h;
  if (h > 1) {
    // 2-D table: the bytes go through the pinned staging block and the column mask is built on the device, so the edit
    // returns without waiting for the stream (the march in flight keeps the old table until the copy is reached)
    const size_t bytes = sizeof(uchar4) * n;
    if (c->tf_stage_busy) { TVK_CHECK_CUDA(c, cudaEventSynchronize(c->tf_stage_done)); c->tf_stage_busy = false; }
    if (c->tf_stage_bytes < bytes) {
      if (c->h_tf_stage) cudaFreeHost(c->h_tf_stage);
      c->h_tf_stage = nullptr; c->tf_stage_bytes = 0;
      TVK_CHECK_CUDA(c, cudaMallocHost(&c->h_tf_stage, bytes));
      c->tf_stage_bytes = bytes;
    }
    if (!c->tf_stage_done) TVK_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->tf_stage_done, cudaEventDisableTiming));
    if (rgba != c->h_tf_stage) memcpy(c->h_tf_stage, rgba, bytes);   // (tvk_tf2d_staging: the block itself may be handed back)
    c->tf_stage_w = w; c->tf_stage_h = h;
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->d_tf, c->h_tf_stage, bytes, cudaMemcpyHostToDevice, c->stream));
    TVK_CHECK_CUDA(c, cudaEventRecord(c->tf_stage_done, c->stream));
    c->tf_stage_busy = true;
    tf_colmask_kernel<<<(w + 31) / 32, 32, 0, c->stream>>>(c->d_tf, w, h, c->d_tf_colmask);
    c->launches++;
    TVK_CHECK_CUDA(c, cudaGetLastError());
    return TVK_OK;
  }
  // 1-D table: the bytes travel through their own page-locked block (the 2-D block doubles as the host's copy of the 2-D
  // table) and the float copies + the column mask are made on the device: no host loop over the table, no stream wait
  {
    const size_t bytes = sizeof(uchar4) * n;
    if (c->tf1_stage_busy) { TVK_CHECK_CUDA(c, cudaEventSynchronize(c->tf1_stage_done)); c->tf1_stage_busy = false; }
    if (c->tf1_stage_bytes < bytes) {
      if (c->h_tf1_stage) cudaFreeHost(c->h_tf1_stage);
      c->h_tf1_stage = nullptr; c->tf1_stage_bytes = 0;
      TVK_CHECK_CUDA(c, cudaMallocHost(&c->h_tf1_stage, bytes));
      c->tf1_stage_bytes = bytes;
    }
    if (!c->tf1_stage_done) TVK_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->tf1_stage_done, cudaEventDisableTiming));
    memcpy(c->h_tf1_stage, rgba, bytes);
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->d_tf, c->h_tf1_stage, bytes, cudaMemcpyHostToDevice, c->stream));
    TVK_CHECK_CUDA(c, cudaEventRecord(c->tf1_stage_done, c->stream));
    c->tf1_stage_busy = true;
    tf1d_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_tf, (uint32_t)n, c->d_tf_f, c->d_tf_f1);
    tf_colmask_kernel<<<(w + 31) / 32, 32, 0, c->stream>>>(c->d_tf, w, 1, c->d_tf_colmask);
    c->launches += 2;
    TVK_CHECK_CUDA(c, cudaGetLastError());
  }
  return TVK_OK;
}
extern "C" int tvk_set_tf1d(tvk_ctx* c, const uint8_t* rgba, uint32_t n) { return set_tf(c, rgba, n, 1); }
extern "C" int tvk_set_tf2d(tvk_ctx* c, const uint8_t* rgba, uint32_t w, uint32_t h) { return set_tf(c, rgba, w, h); }
extern "C" int tvk_tf2d_staging(tvk_ctx* c, const uint8_t** rgba, uint32_t* w, uint32_t* h) {
  if (!c || !rgba || !w || !h) return TVK_ERR_INVALID;
  if (!c->h_tf_stage || !c->tf_stage_w || !c->tf_stage_h) TVK_FAIL(c, TVK_ERR_INVALID, "no 2-D transfer function has been set");
  *rgba = c->h_tf_stage; *w = c->tf_stage_w; *h = c->tf_stage_h;
  return TVK_OK;
}
