// Re-bricking on the device: a dataset with large bricks presented with smaller ones.
//
// Reference (read for behaviour only; nothing is copied):
//   Tuvok/IO/DynamicBrickingDS.cpp   GenericBrickLayout :229-247, TargetBricksPerSource :249-262, SourceBrickIndex
//                                    :273-303, OffsetIntoSource :305-336, ComputedTargetBrickSize :414-435, CopyBrick
//                                    :495-538, Rebrick :1090-1180, MaxMinForKey :862-876
//   Tuvok/IO/IOManager.cpp:1281-1320 LoadRebrickedDataset (size clamp + divisibility check)
//   Tuvok/IO/BMinMax.cpp:8-17        minmax_brick (min/max over all voxels of the brick, ghost cells included)
//
// The reference copies one target brick per GetBrick() call out of a host-side cache of source bricks.  Here the
// mapping target -> (source brick, offset) is a table, and ONE launch cuts every target brick of every LOD: a CTA per
// (brick, z-chunk), a warp per row, min/max reduced per warp.  Pure data movement: bound by HBM (or by PCIe when the
// source store is mapped host memory).
#include "tvk_internal.h"
#include <algorithm>
#include <vector>

namespace {

struct Layout {
  uint32_t tgt[3], tcore[3], score[3], per_src[3];
};

bool make_layout(const tvk_rebrick_desc* d, Layout& L) {
  if (!d || d->overlap != 2 || d->lod_count == 0 || d->lod_count > TVK_MAX_LOD || (d->bytes_per_voxel != 1 && d->bytes_per_voxel != 2)) return false;
  const uint32_t g = 2 * d->overlap;
  for (int a = 0; a < 3; ++a) {
    if (d->src_brick[a] <= g || d->tgt_brick[a] <= g) return false;
    L.tgt[a] = std::min(d->tgt_brick[a], d->src_brick[a]);          // IOManager.cpp:1300-1304
    L.tcore[a] = L.tgt[a] - g;
    L.score[a] = d->src_brick[a] - g;
    if (L.score[a] % L.tcore[a]) return false;                       // IOManager.cpp:1305-1312, Rebrick() :1094-1105
    L.per_src[a] = L.score[a] / L.tcore[a];
  }
  return true;
}

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

struct RebrickJob {
  uint64_t src;        // byte offset of the first voxel to copy in the source store
  uint64_t dst;        // byte offset of the target brick
  uint32_t sx, sxy;    // source brick row / slice pitch in voxels
  uint32_t tx, ty, tz; // target brick size
};

template <typename T>
__global__ void __launch_bounds__(256) rebrick_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       const RebrickJob* __restrict__ jobs, uint32_t* __restrict__ mn,
                                                       uint32_t* __restrict__ mx) {
  const RebrickJob j = jobs[blockIdx.x];
  const T* s = reinterpret_cast<const T*>(src + j.src);
  T* d = reinterpret_cast<T*>(dst + j.dst);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.y * blockDim.x) >> 5;
  const uint32_t rows = j.ty * j.tz;
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  for (uint32_t r = warp; r < rows; r += nwarps) {
    const uint32_t y = r % j.ty, z = r / j.ty;
    const T* srow = s + (size_t)z * j.sxy + (size_t)y * j.sx;
    T* drow = d + (size_t)r * j.tx;
    for (uint32_t x = lane; x < j.tx; x += 32u) {
      const T v = srow[x];
      drow[x] = v;
      lo = min(lo, (uint32_t)v);
      hi = max(hi, (uint32_t)v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0 && lo <= hi) { atomicMin(mn + blockIdx.x, lo); atomicMax(mx + blockIdx.x, hi); }
}

}  // namespace

extern "C" int tvk_rebrick_plan(const tvk_rebrick_desc* d, tvk_rebrick_info* info) {
  Layout L;
  if (!info || !make_layout(d, L)) return TVK_ERR_INVALID;
  memset(info, 0, sizeof(*info));
  uint64_t n = 0, ns = 0, bytes = 0;
  for (uint32_t l = 0; l < d->lod_count; ++l) {
    info->lod_offset[l] = n; info->src_lod_offset[l] = ns;
    uint32_t last[3];
    for (int a = 0; a < 3; ++a) {
      if (d->lod_size[l][a] == 0) return TVK_ERR_INVALID;
      info->lod_bricks[l][a] = ceil_div(d->lod_size[l][a], L.tcore[a]);        // GenericBrickLayout
      info->src_lod_bricks[l][a] = ceil_div(d->lod_size[l][a], L.score[a]);
      const uint32_t extra = d->lod_size[l][a] % L.tcore[a];                    // ComputedTargetBrickSize
      last[a] = extra ? 4u + extra : L.tgt[a];
    }
    const uint32_t* nb = info->lod_bricks[l];
    // bytes of the level: bricks differ in size only when they are last along an axis
    for (uint32_t cz = 0; cz < 2; ++cz) for (uint32_t cy = 0; cy < 2; ++cy) for (uint32_t cx = 0; cx < 2; ++cx) {
      const uint64_t cnt = (uint64_t)(cx ? 1 : nb[0] - 1) * (cy ? 1 : nb[1] - 1) * (cz ? 1 : nb[2] - 1);
      const uint64_t vox = (uint64_t)(cx ? last[0] : L.tgt[0]) * (cy ? last[1] : L.tgt[1]) * (cz ? last[2] : L.tgt[2]);
      bytes += cnt * ((vox * d->bytes_per_voxel + 255) / 256 * 256);
    }
    n += (uint64_t)nb[0] * nb[1] * nb[2];
    ns += (uint64_t)info->src_lod_bricks[l][0] * info->src_lod_bricks[l][1] * info->src_lod_bricks[l][2];
  }
  for (int a = 0; a < 3; ++a) info->tgt_brick[a] = L.tgt[a];
  info->total_bricks = n; info->src_total_bricks = ns; info->total_bytes = bytes;
  return TVK_OK;
}

extern "C" int tvk_rebrick_toc(const tvk_rebrick_desc* d, uint64_t* offsets, uint32_t* sizes3, uint64_t* src_index, uint32_t* src_offset3, uint64_t n) {
  Layout L;
  tvk_rebrick_info info;
  if (!make_layout(d, L) || tvk_rebrick_plan(d, &info) != TVK_OK || n != info.total_bricks) return TVK_ERR_INVALID;
  uint64_t i = 0, off = 0;
  for (uint32_t l = 0; l < d->lod_count; ++l) {
    const uint32_t* nb = info.lod_bricks[l];
    const uint32_t* sb = info.src_lod_bricks[l];
    for (uint32_t z = 0; z < nb[2]; ++z) for (uint32_t y = 0; y < nb[1]; ++y) for (uint32_t x = 0; x < nb[0]; ++x, ++i) {
      const uint32_t c[3] = {x, y, z};
      uint32_t sz[3], si[3];
      for (int a = 0; a < 3; ++a) {
        const uint32_t extra = d->lod_size[l][a] % L.tcore[a];
        sz[a] = (c[a] == nb[a] - 1 && extra) ? 4u + extra : L.tgt[a];
        si[a] = c[a] / L.per_src[a];                                            // SourceBrickIndex
        if (src_offset3) src_offset3[3 * i + a] = (c[a] % L.per_src[a]) * L.tcore[a];   // OffsetIntoSource
        if (sizes3) sizes3[3 * i + a] = sz[a];
      }
      if (src_index) src_index[i] = info.src_lod_offset[l] + si[0] + (uint64_t)si[1] * sb[0] + (uint64_t)si[2] * sb[0] * sb[1];
      if (offsets) offsets[i] = off;
      off += ((uint64_t)sz[0] * sz[1] * sz[2] * d->bytes_per_voxel + 255) / 256 * 256;
    }
  }
  return TVK_OK;
}

extern "C" int tvk_rebrick(tvk_ctx* c, const tvk_rebrick_desc* d, const void* src_store, const uint64_t* src_offsets,
                           const uint32_t* src_sizes3, uint64_t n_src, void* dst_store, double* minmax2) {
  if (!c || !src_store || !src_offsets || !src_sizes3 || !dst_store) return TVK_ERR_INVALID;
  tvk_rebrick_info info;
  if (tvk_rebrick_plan(d, &info) != TVK_OK) TVK_FAIL(c, TVK_ERR_INVALID, "re-bricking: the target brick size (without ghost voxels) must divide the source brick size");
  if (n_src != info.src_total_bricks) TVK_FAIL(c, TVK_ERR_INVALID, "re-bricking: source brick count does not match the source layout");
  if (info.total_bricks > 0x7FFFFFFFull) TVK_FAIL(c, TVK_ERR_UNSUPPORTED, "re-bricking: too many target bricks for one launch");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const uint64_t n = info.total_bricks;
  std::vector<uint64_t> off(n), sidx(n);
  std::vector<uint32_t> sz(3 * n), so(3 * n);
  tvk_rebrick_toc(d, off.data(), sz.data(), sidx.data(), so.data(), n);
  const uint32_t bpv = d->bytes_per_voxel;
  std::vector<RebrickJob> jobs(n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t* ss = src_sizes3 + 3 * sidx[i];
    for (int a = 0; a < 3; ++a)
      if (so[3 * i + a] + sz[3 * i + a] > ss[a]) TVK_FAIL(c, TVK_ERR_INVALID, "re-bricking: a target brick does not fit its source brick (source ToC does not match the layout)");
    jobs[i] = RebrickJob{src_offsets[sidx[i]] + ((uint64_t)so[3 * i] + (uint64_t)so[3 * i + 1] * ss[0] + (uint64_t)so[3 * i + 2] * ss[0] * ss[1]) * bpv,
                         off[i], ss[0], ss[0] * ss[1], sz[3 * i], sz[3 * i + 1], sz[3 * i + 2]};
  }
  // the source may be pinned host memory: the kernel needs its device alias
  const uint8_t* src = static_cast<const uint8_t*>(src_store);
  cudaPointerAttributes at;
  const bool known = cudaPointerGetAttributes(&at, src_store) == cudaSuccess;
  if (!known) (void)cudaGetLastError();
  if (known && at.type == cudaMemoryTypeHost && at.devicePointer) src = static_cast<const uint8_t*>(at.devicePointer);
  else if (!known || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged))
    TVK_FAIL(c, TVK_ERR_INVALID, "re-bricking: the source store must be device memory or pinned (mapped) host memory");
  RebrickJob* d_jobs = nullptr;
  uint32_t* d_mm = nullptr;
  TVK_CHECK_CUDA(c, cudaMalloc(&d_jobs, sizeof(RebrickJob) * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_mm, sizeof(uint32_t) * 2 * n));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(RebrickJob) * n, cudaMemcpyHostToDevice, c->stream));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_mm, 0xFF, sizeof(uint32_t) * n, c->stream));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_mm + n, 0, sizeof(uint32_t) * n, c->stream));
  // enough CTAs per brick to fill the device when there are few bricks
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const unsigned chunks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(8, ((uint64_t)sms * 8 + n - 1) / n));
  if (bpv == 1) rebrick_kernel<uint8_t><<<dim3((unsigned)n, chunks), 256, 0, c->stream>>>(src, static_cast<uint8_t*>(dst_store), d_jobs, d_mm, d_mm + n);
  else rebrick_kernel<uint16_t><<<dim3((unsigned)n, chunks), 256, 0, c->stream>>>(src, static_cast<uint8_t*>(dst_store), d_jobs, d_mm, d_mm + n);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  std::vector<uint32_t> mm(2 * n);
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(mm.data(), d_mm, sizeof(uint32_t) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_jobs); cudaFree(d_mm);
  if (minmax2)
    for (uint64_t i = 0; i < n; ++i) { minmax2[2 * i] = (double)mm[i]; minmax2[2 * i + 1] = (double)mm[n + i]; }
  return TVK_OK;
}
