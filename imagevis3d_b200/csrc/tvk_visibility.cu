// Brick visibility (empty-space) recomputation on the device.
//
// Replaces GLVolumePool::RecomputeVisibility and its helpers
// (Tuvok/Renderer/GL/GLVolumePool.cpp:1580-1718, RecomputeVisibilityForBrickPool
// :992-1018, RecomputeVisibilityForOctree :1021-1359, ContainsData :963-989),
// which the reference runs single-threaded on the CPU at ~7500 bricks/ms and then
// re-uploads as a whole texture.  Here the min/max table (16 B/brick) and the
// metadata (4 B/brick) stay in HBM; one tiny launch per octree level walks the
// hierarchy bottom-up (a parent needs its children's final flags).
#include "tvk_internal.h"
#include <cfloat>

struct VisParams {
  uint32_t mode;
  double smin, smax, gmin, gmax;
};

__device__ __forceinline__ bool contains_data(const VisParams& v, const double* __restrict__ mm,
                                              const double* __restrict__ gmm, uint32_t i) {
  const double lo = mm[2 * i], hi = mm[2 * i + 1];
  if (v.mode == TVK_RM_1DTRANS) return v.smax >= lo && v.smin <= hi;
  if (v.mode == TVK_RM_2DTRANS) {
    const double glo = gmm ? gmm[2 * i] : -DBL_MAX, ghi = gmm ? gmm[2 * i + 1] : DBL_MAX;
    return (v.smax >= lo && v.smin <= hi) && (v.gmax >= glo && v.gmin <= ghi);
  }
  return v.smin >= lo && v.smin <= hi;
}

// one launch per level; level 0 first
__global__ void __launch_bounds__(256) visibility_level_kernel(uint32_t* __restrict__ meta, const double* __restrict__ mm,
                                                                const double* __restrict__ gmm, const __grid_constant__ PoolConst pc,
                                                                const __grid_constant__ VisParams vis, uint32_t lod,
                                                                uint32_t* __restrict__ counts) {
  const uint32_t bx = pc.lod_bricks[lod][0], by = pc.lod_bricks[lod][1], bz = pc.lod_bricks[lod][2];
  const uint32_t n = bx * by * bz;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t i = pc.lod_offset[lod] + t;
  atomicAdd(counts + 0, 1u);
  uint32_t m = meta[i];
  const bool cd = contains_data(vis, mm, gmm, i);
  if (m >= TVK_BI_FLAG_COUNT) {
    if (cd) return;                      // resident and visible: keeps its slot
    m = TVK_BI_EMPTY;                    // RecomputeVisibilityForBrickPool
  } else {
    m = TVK_BI_MISSING;
  }
  if (!cd) {
    if (lod == 0) {
      m = TVK_BI_CHILD_EMPTY;
      atomicAdd(counts + 3, 1u);
    } else {
      const uint32_t x = t % bx, y = (t / bx) % by, z = t / (bx * by);
      const uint32_t cx = pc.lod_bricks[lod - 1][0], cy = pc.lod_bricks[lod - 1][1], cz = pc.lod_bricks[lod - 1][2];
      bool all_child_empty = true;
#pragma unroll
      for (uint32_t d = 0; d < 8; ++d) {
        const uint32_t xx = 2 * x + (d & 1), yy = 2 * y + ((d >> 1) & 1), zz = 2 * z + (d >> 2);
        if (xx < cx && yy < cy && zz < cz)
          all_child_empty = all_child_empty && (meta[pc.lod_offset[lod - 1] + xx + yy * cx + zz * cx * cy] == TVK_BI_CHILD_EMPTY);
      }
      m = all_child_empty ? TVK_BI_CHILD_EMPTY : TVK_BI_EMPTY;
      atomicAdd(counts + (all_child_empty ? 2 : 1), 1u);
    }
  }
  meta[i] = m;
}

extern "C" int tvk_visibility_set_minmax(tvk_ctx* c, const double* minmax, const double* grad_minmax, uint32_t n) {
  if (!c || !c->has_pool || !minmax || n != c->pc.total_bricks) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  if (!c->d_minmax) TVK_CHECK_CUDA(c, cudaMalloc(&c->d_minmax, sizeof(double) * 2 * n));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->d_minmax, minmax, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, c->stream));
  if (grad_minmax) {
    if (!c->d_gminmax) TVK_CHECK_CUDA(c, cudaMalloc(&c->d_gminmax, sizeof(double) * 2 * n));
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->d_gminmax, grad_minmax, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, c->stream));
  } else {
    cudaFree(c->d_gminmax); c->d_gminmax = nullptr;
  }
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_minmax = n;
  return TVK_OK;
}

extern "C" int tvk_visibility_recompute(tvk_ctx* c, const tvk_visibility* vis, uint32_t* counts) {
  if (!c || !c->has_pool || !vis) return TVK_ERR_INVALID;
  if (!c->d_minmax || c->n_minmax != c->pc.total_bricks) TVK_FAIL(c, TVK_ERR_INVALID, "min/max table not set");
  if (vis->render_mode > TVK_RM_ISOSURFACE) TVK_FAIL(c, TVK_ERR_INVALID, "bad render mode");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  VisParams vp{vis->render_mode, vis->smin, vis->smax, vis->gmin, vis->gmax};
  TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_counts, 0, 4 * sizeof(uint32_t), c->stream));
  for (uint32_t l = 0; l < c->pc.lod_count; ++l) {
    const uint32_t n = c->pc.lod_bricks[l][0] * c->pc.lod_bricks[l][1] * c->pc.lod_bricks[l][2];
    visibility_level_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_meta, c->d_minmax, c->d_gminmax, c->pc, vp, l, c->d_counts);
    c->launches++;
  }
  TVK_CHECK_CUDA(c, cudaGetLastError());
  if (counts) {
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(counts, c->d_counts, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return TVK_OK;
}

// ---------------------------------------------------------------------------
// sort-last compositor: ordered "under" of premultiplied partial images
// (dst += (1-dst.a)*src, the premultiplied form of Compositing.glsl:33-38).
// Pure streaming: 16 B/pixel read per partial, 16 B/pixel written.
// ---------------------------------------------------------------------------
struct PartialList { const float4* p[64]; };

__global__ void __launch_bounds__(256) composite_kernel(const __grid_constant__ PartialList partials, uint32_t n, uint64_t pixels,
                                                         float4* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (uint64_t)gridDim.x * blockDim.x) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < n; ++k) {
      const float4 s = partials.p[k][i];
      const float w = 1.0f - d.w;
      d.x += w * s.x; d.y += w * s.y; d.z += w * s.z; d.w += w * s.w;
    }
    out[i] = d;
  }
}

// Asynchronous: enqueued on the context's stream (see tvk_set_stream); the caller synchronises.
extern "C" int tvk_composite_ordered(tvk_ctx* c, const void* const* partial_dptrs, uint32_t n, uint64_t pixels, void* out_dptr) {
  if (!c || !partial_dptrs || !out_dptr || n == 0 || n > 64) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  PartialList pl{};
  for (uint32_t k = 0; k < n; ++k) pl.p[k] = static_cast<const float4*>(partial_dptrs[k]);
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  uint64_t blocks = (pixels + 255) / 256;
  if (blocks > (uint64_t)sms * 8) blocks = (uint64_t)sms * 8;
  composite_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(pl, n, pixels, (float4*)out_dptr);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  return TVK_OK;
}
