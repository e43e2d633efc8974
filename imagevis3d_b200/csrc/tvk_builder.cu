// K2 -- ExtendedOctree / UVF brick-and-LOD builder as streaming kernels (sm_100a).
//
// Replaces ExtendedOctreeConverter::Convert (Tuvok/IO/UVF/ExtendedOctree/
// ExtendedOctreeConverter.cpp:128-280) and the histogram blocks:
//   plan              ExtendedOctree::ComputeMetadata / ComputeBrickSize (ExtendedOctree.cpp:188-285)
//   downsample_kernel ComputeHierarchy -> DownsampleBrick -> DownsampleBricktoBrick
//                     (ExtendedOctreeConverter.inc:1-398), VolumeTools::Filter (VolumeTools.h:168-262)
//   cut_bricks_kernel PermuteInputData/GetInputBrick (.cpp:1381-1414,288-373) for LOD 0,
//                     FillOverlap/CopyBrickToBrick (.cpp:1203-1372,1153-1181) for LOD >= 1,
//                     ComputeBrickStats (.inc:408-444) fused in
//   hist1d_kernel     Histogram1DDataBlock::Compute (IO/UVF/Histogram1DDataBlock.cpp:60-160)
//   grad_*_kernel     Histogram2DDataBlock::ComputeTemplate (IO/UVF/Histogram2DDataBlock.cpp:348-455)
//
// The reference works brick by brick through an LRU file cache; the result is
// nevertheless a pure function of GLOBAL voxel coordinates, which is what lets
// the work stream:
//  * the inner region of every LOD-(l+1) brick is the 2x2x2 (edge: 4/2/1 tap)
//    reduction of LOD-l voxels 2g..2g+1 -- valid because the inner brick size is
//    even, so child bricks never straddle a parent voxel.  We therefore keep a
//    flat volume per LOD in HBM (total 1/7 of the input) and reduce flat->flat;
//  * every brick voxel, ghost cells included, is then a gather from the flat
//    volume of its LOD.  LOD 0 follows GetInputBrick's linear-file-offset reads
//    (a scan line of a next-to-last brick may run into the following row when the
//    last brick is thinner than the overlap).  LOD >= 1 follows FillOverlap's
//    fixed order of 10 neighbour copies: resolve_overlap() replays "which copy
//    wrote this ghost voxel last, and had its source brick been completed yet",
//    which reproduces the reference's zeroed corner ghosts bit for bit.
// Algorithmic bytes per input voxel: read bpv + write 1.2136*(8/7)*bpv (bricks)
// + 1/7*bpv (pyramid) -- DESIGN.md, "K2".
#include "tvk_internal.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

static const uint64_t kBrickAlign = 256;   // device store: every brick starts 256 B aligned

extern "C" int tvk_build_plan(const tvk_build_desc* d, tvk_build_info* info) {
  if (!d || !info) return TVK_ERR_INVALID;
  if (d->brick_size <= 2 * d->overlap || (d->bytes_per_voxel != 1 && d->bytes_per_voxel != 2)) return TVK_ERR_INVALID;
  for (int a = 0; a < 3; ++a) if (d->volume_size[a] == 0) return TVK_ERR_INVALID;
  memset(info, 0, sizeof(*info));
  const uint32_t usable = d->brick_size - 2 * d->overlap;
  uint32_t v[3] = {d->volume_size[0], d->volume_size[1], d->volume_size[2]};
  uint32_t n = 0;
  do {
    if (n > 0)
      for (int a = 0; a < 3; ++a) if (v[a] > 1) v[a] = (uint32_t)std::ceil(v[a] / 2.0);
    if (n >= TVK_MAX_LOD) return TVK_ERR_UNSUPPORTED;
    for (int a = 0; a < 3; ++a) {
      info->lod_size[n][a] = v[a];
      info->lod_bricks[n][a] = (uint32_t)std::ceil((double)v[a] / (double)usable);
    }
    ++n;
  } while (v[0] > 1 || v[1] > 1 || v[2] > 1);
  info->lod_count = n;
  uint64_t off = 0;
  for (uint32_t l = 0; l < n; ++l) {
    info->lod_offset[l] = off;
    off += (uint64_t)info->lod_bricks[l][0] * info->lod_bricks[l][1] * info->lod_bricks[l][2];
  }
  info->total_bricks = off;
  return TVK_OK;
}

struct BuildConst {
  uint32_t lod_size[TVK_MAX_LOD][3];
  uint32_t lod_bricks[TVK_MAX_LOD][3];
  uint32_t brick, overlap, core;
};

// What a brick-cut kernel sees of its source: rows [z0, z0 + rows) of the flat LOD volume (all of it: z0 = 0, rows =
// the LOD's depth) and the first brick row of the launch -- so that LOD 0 can be fed through in z-slabs.
struct SlabView { uint32_t z0, rows, bz_first; };

__host__ __device__ inline uint32_t brick_dim(const BuildConst& bc, uint32_t lod, int axis, uint32_t coord) {
  const bool last = coord >= bc.lod_bricks[lod][axis] - 1;
  const uint32_t rem = bc.lod_size[lod][axis] % bc.core;
  return (last && rem) ? (2 * bc.overlap + rem) : bc.brick;
}

// ------------------------------------------------------------------ filters
template <typename T> __device__ __forceinline__ void order2(T& a, T& b) { if (a > b) { T t = a; a = b; b = t; } }
template <typename T> __device__ __forceinline__ void insert4(T& a, T& b, T& c, T& d, T& p) {
  if (p > c) { order2(d, p); }
  else if (p < b) { d = c; c = b; b = p; order2(a, b); }
  else { d = c; c = p; }
}
template <typename T, bool MEDIAN> __device__ __forceinline__ T filter8(T a, T b, T c, T d, T e, T f, T g, T h) {
  if (MEDIAN) {
    order2(a, b); order2(c, d); order2(a, c); order2(b, d); order2(b, c);
    insert4(a, b, c, d, e);
    insert4(a, b, c, d, f);
    const T m = d < g ? d : g;
    return m > c ? m : c;
  }
  // T((double sum)/8.0): the sum is exact, the quotient truncates toward zero
  return (T)(((uint32_t)a + b + c + d + e + f + g + h) >> 3);
}
template <typename T, bool MEDIAN> __device__ __forceinline__ T filter4(T a, T b, T c, T d) {
  if (MEDIAN) { order2(a, b); order2(b, c); return a > b ? a : b; }
  return (T)(((uint32_t)a + b + c + d) >> 2);
}
template <typename T, bool MEDIAN> __device__ __forceinline__ T filter2(T a, T b) {
  if (MEDIAN) return a;
  return (T)(((uint32_t)a + b) >> 1);
}

// flat LOD l -> flat LOD l+1; one thread per output voxel, x fastest (coalesced:
// a warp reads 64 consecutive input voxels from each of 4 rows).
template <typename T, bool MEDIAN>
__global__ void __launch_bounds__(256) downsample_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          uint32_t sx, uint32_t sy, uint32_t sz,
                                                          uint32_t dx, uint32_t dy, uint32_t dz) {
  const uint64_t n = (uint64_t)dx * dy * dz;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(i % dx), y = (uint32_t)((i / dx) % dy), z = (uint32_t)(i / ((uint64_t)dx * dy));
    // an axis of extent 1 is not reduced any further (ExtendedOctree.cpp:204-217)
    const uint32_t x0 = (sx > 1) ? 2 * x : x, y0 = (sy > 1) ? 2 * y : y, z0 = (sz > 1) ? 2 * z : z;
    const bool hx = (sx > 1) && (x0 + 1 < sx), hy = (sy > 1) && (y0 + 1 < sy), hz = (sz > 1) && (z0 + 1 < sz);
    const uint64_t rs = sx, ss = (uint64_t)sx * sy;
    const T* p = src + x0 + y0 * rs + z0 * ss;
    T out;
    // argument order = p0..p7 of DownsampleBricktoBrick: p1 = +z, p2 = +y, p3 = +y+z, p4.. = +x
    if (hx && hy && hz) out = filter8<T, MEDIAN>(p[0], p[ss], p[rs], p[rs + ss], p[1], p[1 + ss], p[1 + rs], p[1 + rs + ss]);
    else if (!hx && hy && hz) out = filter4<T, MEDIAN>(p[0], p[ss], p[rs], p[rs + ss]);
    else if (hx && !hy && hz) out = filter4<T, MEDIAN>(p[0], p[ss], p[1], p[1 + ss]);
    else if (hx && hy && !hz) out = filter4<T, MEDIAN>(p[0], p[rs], p[1], p[1 + rs]);
    else if (!hx && !hy && hz) out = filter2<T, MEDIAN>(p[0], p[ss]);
    else if (hx && !hy && !hz) out = filter2<T, MEDIAN>(p[0], p[1]);
    else if (!hx && hy && !hz) out = filter2<T, MEDIAN>(p[0], p[rs]);
    else out = p[0];
    dst[i] = out;
  }
}


// ------------------------------------------------------------------ 16-byte vector helpers
// The streaming kernels move 16 bytes per thread.  Brick rows start at x = bx*core - overlap, so the
// source of a destination chunk is almost never 16-byte aligned: load16() reads the (up to) five
// aligned 32-bit words that cover [byte_off, byte_off+16) and funnel-shifts them into place.  Bytes
// before the array or at/after nbytes read as zero (GetInputBrick's "offset past the end of the file").
__device__ __forceinline__ uint32_t load_word_guarded(const uint8_t* __restrict__ base, int64_t w, uint64_t nbytes) {
  if (w < 0) return 0u;
  const uint64_t b = (uint64_t)w * 4u;
  if (b + 4u <= nbytes) return __ldg(reinterpret_cast<const uint32_t*>(base + b));
  uint32_t v = 0u;
  for (uint32_t k = 0; k < 4u; ++k) if (b + k < nbytes) v |= (uint32_t)__ldg(base + b + k) << (8u * k);
  return v;
}
__device__ __forceinline__ uint4 load16(const uint8_t* __restrict__ base, int64_t byte_off, uint64_t nbytes) {
  const int64_t w0 = byte_off >> 2;
  const uint32_t sh = ((uint32_t)byte_off & 3u) * 8u;
  uint32_t w[5];
  if (byte_off >= 0 && (uint64_t)(w0 * 4 + 20) <= nbytes) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base) + w0;
    if (((uint32_t)byte_off & 15u) == 0u) return __ldg(reinterpret_cast<const uint4*>(p));
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = __ldg(p + k);
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = load_word_guarded(base, w0 + k, nbytes);
  }
  return make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh),
                    __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
}
// CB = 16, 8 or 4 bytes at an arbitrary byte offset, zero-extended to a uint4
template <int CB>
__device__ __forceinline__ uint4 load_chunk(const uint8_t* __restrict__ base, int64_t byte_off, uint64_t nbytes) {
  if (CB == 16) return load16(base, byte_off, nbytes);
  constexpr int NW = CB / 4;
  const int64_t w0 = byte_off >> 2;
  const uint32_t sh = ((uint32_t)byte_off & 3u) * 8u;
  uint32_t w[NW + 1];
  if (byte_off >= 0 && (uint64_t)(w0 * 4 + 4 * (NW + 1)) <= nbytes) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base) + w0;
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = __ldg(p + k);
  } else {
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = load_word_guarded(base, w0 + k, nbytes);
  }
  uint4 v = make_uint4(__funnelshift_r(w[0], w[1], sh), 0u, 0u, 0u);
  if (NW == 2) v.y = __funnelshift_r(w[1], w[NW], sh);
  return v;
}
// keep bytes [lo, hi) of one 32-bit word whose first byte has index `first`
__device__ __forceinline__ uint32_t keep_bytes(uint32_t v, int first, int lo, int hi) {
  const int a = lo - first, b = hi - first;
  if (a >= 4 || b <= 0) return 0u;
  uint32_t m = 0xFFFFFFFFu;
  if (a > 0) m &= 0xFFFFFFFFu << (8 * a);
  if (b < 4) m &= 0xFFFFFFFFu >> (8 * (4 - b));
  return v & m;
}
__device__ __forceinline__ uint4 keep16(uint4 v, int lo, int hi) {
  return make_uint4(keep_bytes(v.x, 0, lo, hi), keep_bytes(v.y, 4, lo, hi), keep_bytes(v.z, 8, lo, hi), keep_bytes(v.w, 12, lo, hi));
}
template <typename T> __device__ __forceinline__ uint32_t vec_get(const uint4& v, int e) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  if (sizeof(T) == 1) return (w[e >> 2] >> (8 * (e & 3))) & 0xFFu;
  return (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
}
template <typename T> __device__ __forceinline__ void vec_set(uint4& v, int e, uint32_t val) {
  const int wi = (sizeof(T) == 1) ? (e >> 2) : (e >> 1);
  const uint32_t s = (sizeof(T) == 1) ? val << (8 * (e & 3)) : val << (16 * (e & 1));
  if (wi == 0) v.x |= s; else if (wi == 1) v.y |= s; else if (wi == 2) v.z |= s; else v.w |= s;
}
template <typename T> __device__ __forceinline__ void vec_minmax(const uint4& v, uint32_t& pmn, uint32_t& pmx) {
  if (sizeof(T) == 1) {
    pmn = __vminu4(__vminu4(pmn, v.x), __vminu4(__vminu4(v.y, v.z), v.w));
    pmx = __vmaxu4(__vmaxu4(pmx, v.x), __vmaxu4(__vmaxu4(v.y, v.z), v.w));
  } else {
    pmn = __vminu2(__vminu2(pmn, v.x), __vminu2(__vminu2(v.y, v.z), v.w));
    pmx = __vmaxu2(__vmaxu2(pmx, v.x), __vmaxu2(__vmaxu2(v.y, v.z), v.w));
  }
}

// flat LOD l -> flat LOD l+1, 16 bytes of each of the (up to) four source rows per thread, 8 bytes
// out.  Requires 16-byte aligned source rows (sx*sizeof(T) % 16 == 0), which also makes the
// destination rows 8-byte aligned and every x pair complete.
template <typename T, bool MEDIAN>
__global__ void __launch_bounds__(256) downsample_vec_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                              uint32_t sx, uint32_t sy, uint32_t sz,
                                                              uint32_t dx, uint32_t dy, uint32_t dz) {
  constexpr int E = 16 / sizeof(T), V = E / 2;
  const uint32_t ncx = sx / E;
  const uint64_t n = (uint64_t)ncx * dy * dz;
  const uint64_t rs = sx, ss = (uint64_t)sx * sy;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t cx = (uint32_t)(idx % ncx);
    const uint64_t row = idx / ncx;
    const uint32_t y = (uint32_t)(row % dy), z = (uint32_t)(row / dy);
    const uint32_t y0 = (sy > 1) ? 2 * y : y, z0 = (sz > 1) ? 2 * z : z;
    const bool hy = (sy > 1) && (y0 + 1 < sy), hz = (sz > 1) && (z0 + 1 < sz);
    const T* p = src + (uint64_t)cx * E + y0 * rs + z0 * ss;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    uint4 b = a, c = a, d = a;
    if (hz) b = __ldg(reinterpret_cast<const uint4*>(p + ss));
    if (hy) c = __ldg(reinterpret_cast<const uint4*>(p + rs));
    if (hy && hz) d = __ldg(reinterpret_cast<const uint4*>(p + rs + ss));
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int m = 0; m < V; ++m) {
      const T a0 = (T)vec_get<T>(a, 2 * m), a1 = (T)vec_get<T>(a, 2 * m + 1);
      const T b0 = (T)vec_get<T>(b, 2 * m), b1 = (T)vec_get<T>(b, 2 * m + 1);
      const T c0 = (T)vec_get<T>(c, 2 * m), c1 = (T)vec_get<T>(c, 2 * m + 1);
      const T d0 = (T)vec_get<T>(d, 2 * m), d1 = (T)vec_get<T>(d, 2 * m + 1);
      T r;
      // p0..p7 of DownsampleBricktoBrick: p1 = +z, p2 = +y, p3 = +y+z, p4.. = +x
      if (hy && hz) r = filter8<T, MEDIAN>(a0, b0, c0, d0, a1, b1, c1, d1);
      else if (hz) r = filter4<T, MEDIAN>(a0, b0, a1, b1);
      else if (hy) r = filter4<T, MEDIAN>(a0, c0, a1, c1);
      else r = filter2<T, MEDIAN>(a0, a1);
      vec_set<T>(o, m, (uint32_t)r);
    }
    *reinterpret_cast<uint2*>(dst + (uint64_t)cx * V + (uint64_t)y * dx + (uint64_t)z * dx * dy) = make_uint2(o.x, o.y);
  }
}

// The LOD 0 -> 1 reduction reads every raw voxel exactly once, 16 bytes at a time -- so it also counts the 1-D
// histogram on the way (Histogram1DDataBlock::Compute is a histogram of exactly these voxels), which saves one full
// read of the volume.  One CTA of 1024 threads per SM; the table lives in shared memory: 32 per-warp 256-bin tables
// for 8-bit data, one 32768-bin window [0, 32768) for 16-bit data (values above it are counted afterwards by
// hist16_window_kernel, which returns at once when the volume has none).
// downsample_hist_kernel itself packs TWO 16-bit counters into each word of the same 128 KB, so its table spans all
// 65536 bins and full-range 16-bit volumes are not read a second time (see hist_add16_packed).
#define HIST_WINDOW 32768u
// One packed 16-bit counter per bin: bin v lives in half v>>15 of word v&0x7FFF (neighbouring values, which smooth
// data sends to the table together, stay in different words and banks).  A half may never reach 0x10000 (the low
// one would carry into the other bin), which the caller guarantees like this: the CTA adds at most 0x8000 voxels
// between two barriers, and the one add that takes a half from below 0x8000 to 0x8000 or above (bit 15 of the half
// goes 0 -> 1: `~old & new & 0x80008000`) moves 0x8000 counts to the global bin before its thread reaches the
// barrier.  While that move is pending the half stays >= 0x8000, so no second add can cross (at most one move is
// pending per half, and it never borrows); at a barrier every half is therefore < 0x8000, and inside a round it is
// at most 0x7FFF + 0x8000 = 0xFFFF.
__device__ __forceinline__ void packed_hist_add(unsigned int* __restrict__ table, unsigned long long* __restrict__ bins,
                                                uint32_t v, uint32_t cnt) {
  const uint32_t inc = cnt << ((v >> 11) & 16u);
  unsigned int* p = table + (v & 0x7FFFu);
  const uint32_t old = atomicAdd(p, inc);
  const uint32_t crossed = ~old & (old + inc) & 0x80008000u;
  if (crossed) {
    atomicSub(p, crossed);
    atomicAdd(&bins[v], 0x8000ull);
  }
}
__device__ __forceinline__ void hist_add16_packed(const uint4& q, unsigned int* __restrict__ table,
                                                  unsigned long long* __restrict__ bins) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  uint32_t cur = w[0] & 0xFFFFu, cnt = 0;          // equal neighbours are merged into one atomic
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t v = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
    if (v == cur) cnt++;
    else { packed_hist_add(table, bins, cur, cnt); cur = v; cnt = 1; }
  }
  packed_hist_add(table, bins, cur, cnt);
}
__device__ __forceinline__ bool same16(const uint4& p, const uint4& q) { return p.x == q.x && p.y == q.y && p.z == q.z && p.w == q.w; }
template <typename T>
__device__ __forceinline__ void hist_add16(const uint4& q, unsigned int* __restrict__ table) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  if (sizeof(T) == 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&table[w[k] & 0xFFu], 1u); atomicAdd(&table[(w[k] >> 8) & 0xFFu], 1u);
      atomicAdd(&table[(w[k] >> 16) & 0xFFu], 1u); atomicAdd(&table[w[k] >> 24], 1u);
    }
  } else {
    uint32_t cur = w[0] & 0xFFFFu, cnt = 0;        // equal neighbours are merged into one atomic
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t v = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
      if (v == cur) cnt++;
      else {
        if (cur < HIST_WINDOW) atomicAdd(&table[cur], cnt);
        cur = v; cnt = 1;
      }
    }
    if (cur < HIST_WINDOW) atomicAdd(&table[cur], cnt);
  }
}

template <typename T, bool MEDIAN>
__global__ void __launch_bounds__(1024, 1) downsample_hist_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                                  uint32_t sx, uint32_t sy, uint32_t sz,
                                                                  uint32_t dx, uint32_t dy, uint32_t dz,
                                                                  unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned int hsm[];
  constexpr uint32_t TABLE = sizeof(T) == 1 ? 32u * 256u : HIST_WINDOW;
  for (uint32_t i = threadIdx.x; i < TABLE; i += blockDim.x) hsm[i] = 0u;
  __syncthreads();
  unsigned int* mine = sizeof(T) == 1 ? hsm + (threadIdx.x >> 5) * 256u : hsm;
  constexpr int E = 16 / sizeof(T), V = E / 2;
  const uint32_t ncx = sx / E;
  const uint64_t n = (uint64_t)ncx * dy * dz;
  const uint64_t rs = sx, ss = (uint64_t)sx * sy;
  // uniform trip count: the 16-bit table needs a barrier per round (32 voxels x 1024 threads = 0x8000 adds at most)
  const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (n + gstride - 1) / gstride;
  // (cx, y, z) of this thread's 16-byte column, advanced by the grid stride with carries: the two 64-bit divisions
  // it replaces were a quarter of the instructions of a round
  const uint64_t idx0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cx = (uint32_t)(idx0 % ncx), y = (uint32_t)((idx0 / ncx) % dy);
  uint64_t z = (idx0 / ncx) / dy;                 // may start (and run) past dz: those rounds are idle
  const uint32_t step_cx = (uint32_t)(gstride % ncx), step_y = (uint32_t)((gstride / ncx) % dy);
  const uint64_t step_z = (gstride / ncx) / dy;
  for (uint64_t round = 0; round < rounds; ++round) {
    if (round) {
      cx += step_cx;
      const uint32_t carry_x = cx >= ncx ? 1u : 0u;
      cx -= carry_x ? ncx : 0u;
      y += step_y + carry_x;
      const uint32_t carry_y = y >= dy ? 1u : 0u;
      y -= carry_y ? dy : 0u;
      z += step_z + carry_y;
    }
    const bool act = z < dz;
    bool hy = false, hz = false;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a, c = a, d = a;
    if (act) {
      const uint32_t y0 = (sy > 1) ? 2 * y : y, z0 = (sz > 1) ? 2 * z : z;
      hy = (sy > 1) && (y0 + 1 < sy); hz = (sz > 1) && (z0 + 1 < sz);
      const T* p = src + (uint64_t)cx * E + y0 * rs + z0 * ss;
      a = __ldg(reinterpret_cast<const uint4*>(p));
      b = a; c = a; d = a;
      if (hz) b = __ldg(reinterpret_cast<const uint4*>(p + ss));
      if (hy) c = __ldg(reinterpret_cast<const uint4*>(p + rs));
      if (hy && hz) d = __ldg(reinterpret_cast<const uint4*>(p + rs + ss));
    }
    if (sizeof(T) == 2) {
      __syncthreads();                           // every move to the global bins of the previous round is done
      // a warp whose 32 x 32 voxels are one single value (background, saturated regions) issues one atomic
      const uint32_t v0 = a.x & 0xFFFFu;
      const bool uni = act && a.x == (v0 | (v0 << 16)) && a.y == a.x && a.z == a.x && a.w == a.x &&
                       same16(b, a) && same16(c, a) && same16(d, a);      // rows that do not exist are copies of a
      const uint32_t first = __shfl_sync(0xffffffffu, v0, 0);
      if (__all_sync(0xffffffffu, uni && v0 == first)) {
        const uint32_t mine_cnt = 8u * (1u + (hz ? 1u : 0u) + (hy ? 1u : 0u) + ((hy && hz) ? 1u : 0u));
        const uint32_t tot = __reduce_add_sync(0xffffffffu, mine_cnt);
        if ((threadIdx.x & 31u) == 0u) packed_hist_add(hsm, bins, v0, tot);
      } else if (act) {
        hist_add16_packed(a, hsm, bins);
        if (hz) hist_add16_packed(b, hsm, bins);
        if (hy) hist_add16_packed(c, hsm, bins);
        if (hy && hz) hist_add16_packed(d, hsm, bins);
      }
    }
    if (!act) continue;
    if (sizeof(T) == 1) {
      hist_add16<T>(a, mine);                    // only rows that exist: the stand-ins b = c = d = a are not counted
      if (hz) hist_add16<T>(b, mine);
      if (hy) hist_add16<T>(c, mine);
      if (hy && hz) hist_add16<T>(d, mine);
    }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int m = 0; m < V; ++m) {
      const T a0 = (T)vec_get<T>(a, 2 * m), a1 = (T)vec_get<T>(a, 2 * m + 1);
      const T b0 = (T)vec_get<T>(b, 2 * m), b1 = (T)vec_get<T>(b, 2 * m + 1);
      const T c0 = (T)vec_get<T>(c, 2 * m), c1 = (T)vec_get<T>(c, 2 * m + 1);
      const T d0 = (T)vec_get<T>(d, 2 * m), d1 = (T)vec_get<T>(d, 2 * m + 1);
      T r;
      if (hy && hz) r = filter8<T, MEDIAN>(a0, b0, c0, d0, a1, b1, c1, d1);
      else if (hz) r = filter4<T, MEDIAN>(a0, b0, a1, b1);
      else if (hy) r = filter4<T, MEDIAN>(a0, c0, a1, c1);
      else r = filter2<T, MEDIAN>(a0, a1);
      vec_set<T>(o, m, (uint32_t)r);
    }
    *reinterpret_cast<uint2*>(dst + (uint64_t)cx * V + (uint64_t)y * dx + (uint64_t)z * dx * dy) = make_uint2(o.x, o.y);
  }
  __syncthreads();
  if (sizeof(T) == 1) {
    for (uint32_t b = threadIdx.x; b < 256u; b += blockDim.x) {
      unsigned int s = 0;
      for (uint32_t w = 0; w < 32u; ++w) s += hsm[w * 256u + b];
      if (s) atomicAdd(&bins[b], (unsigned long long)s);
    }
  } else {
    for (uint32_t b = threadIdx.x; b < HIST_WINDOW; b += blockDim.x) {
      const unsigned int s = hsm[b];
      if (s & 0xFFFFu) atomicAdd(&bins[b], (unsigned long long)(s & 0xFFFFu));
      if (s >> 16) atomicAdd(&bins[b + HIST_WINDOW], (unsigned long long)(s >> 16));
    }
  }
}

// FillOverlap replay for one ghost voxel of brick (bx,by,bz) at LOD >= 1.
// Returns true and the global voxel coordinate to gather, or false for "zero".
__device__ __forceinline__ bool resolve_overlap(const BuildConst& bc, uint32_t lod, int bx, int by, int bz,
                                                int i, int j, int k, uint32_t& gx, uint32_t& gy, uint32_t& gz) {
  const int o = (int)bc.overlap;
  const int nbx = (int)bc.lod_bricks[lod][0], nby = (int)bc.lod_bricks[lod][1], nbz = (int)bc.lod_bricks[lod][2];
  bool processed = true;
#pragma unroll 1
  for (int iter = 0; iter < 8; ++iter) {
    const int tx = (int)brick_dim(bc, lod, 0, bx), ty = (int)brick_dim(bc, lod, 1, by), tz = (int)brick_dim(bc, lod, 2, bz);
    const int cx = (i < o) ? -1 : (i >= tx - o ? 1 : 0);
    const int cy = (j < o) ? -1 : (j >= ty - o ? 1 : 0);
    const int cz = (k < o) ? -1 : (k >= tz - o ? 1 : 0);
    if (cx == 0 && cy == 0 && cz == 0) {
      gx = (uint32_t)(bx * (int)bc.core + i - o);
      gy = (uint32_t)(by * (int)bc.core + j - o);
      gz = (uint32_t)(bz * (int)bc.core + k - o);
      return true;
    }
    if (!processed) return false;       // ghost cell of a brick FillOverlap had not reached yet: still zero
    const bool R = bx < nbx - 1, B = by < nby - 1, K = bz < nbz - 1, L = bx > 0, T = by > 0, F = bz > 0;
    // last writer first: ops 10..1 of FillOverlap (.cpp:1234-1340)
    if (R && B && K && cx == 1 && cy == 1 && cz == 1) { i = o + (i - (tx - o)); j = o + (j - (ty - o)); k = o + (k - (tz - o)); bx++; by++; bz++; processed = false; }
    else if (B && K && cy == 1 && cz == 1) { j = o + (j - (ty - o)); k = o + (k - (tz - o)); by++; bz++; processed = false; }
    else if (R && K && cx == 1 && cz == 1) { i = o + (i - (tx - o)); k = o + (k - (tz - o)); bx++; bz++; processed = false; }
    else if (R && B && cx == 1 && cy == 1) { i = o + (i - (tx - o)); j = o + (j - (ty - o)); bx++; by++; processed = false; }
    else if (F && cz == -1) { bz--; k = (int)brick_dim(bc, lod, 2, bz) - 2 * o + k; }
    else if (T && cy == -1) { by--; j = (int)brick_dim(bc, lod, 1, by) - 2 * o + j; }
    else if (L && cx == -1) { bx--; i = (int)brick_dim(bc, lod, 0, bx) - 2 * o + i; }
    else if (K && cz == 1) { k = o + (k - (tz - o)); bz++; processed = false; }
    else if (B && cy == 1) { j = o + (j - (ty - o)); by++; processed = false; }
    else if (R && cx == 1) { i = o + (i - (tx - o)); bx++; processed = false; }
    else return false;                  // no copy ever touches it: DownsampleBrick's memset
  }
  return false;
}

// One CTA per (brick, z-chunk): gathers the brick (ghost cells included) from the
// flat LOD volume into the store and reduces min/max over the whole brick.
template <typename T>
__global__ void __launch_bounds__(256) cut_bricks_kernel(const T* __restrict__ flat, uint8_t* __restrict__ store,
                                                          const uint64_t* __restrict__ offsets, uint64_t first_brick,
                                                          uint32_t lod, const __grid_constant__ BuildConst bc,
                                                          uint32_t* __restrict__ mn, uint32_t* __restrict__ mx,
                                                          uint32_t x_first, uint32_t x_count, SlabView sv) {
  const uint32_t nbx = bc.lod_bricks[lod][0], nby = bc.lod_bricks[lod][1];
  const int bx = (int)(x_first + blockIdx.x % x_count), by = (int)((blockIdx.x / x_count) % nby), bz = (int)(sv.bz_first + blockIdx.x / (x_count * nby));
  const uint32_t b = (uint32_t)bx + (uint32_t)by * nbx + (uint32_t)bz * nbx * nby;
  const uint32_t tx = brick_dim(bc, lod, 0, bx), ty = brick_dim(bc, lod, 1, by), tz = brick_dim(bc, lod, 2, bz);
  const uint32_t sx = bc.lod_size[lod][0], sy = bc.lod_size[lod][1];
  const uint64_t nvox = (uint64_t)sx * sy * sv.rows;
  T* dst = reinterpret_cast<T*>(store + offsets[first_brick + b]);
  const int o = (int)bc.overlap;
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  // A warp owns one brick row (j,k) at a time: everything that depends on (j,k) -- row validity,
  // the source row pointer, ghost classification -- is warp-uniform and computed once per row; the
  // lanes then stream x.  No per-voxel division, loads and stores are both unit-stride.
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.y * blockDim.x) >> 5;
  const uint32_t rows = ty * tz;
  const int nbz = (int)bc.lod_bricks[lod][2];
  const bool fx = bx == 0, fy = by == 0, fz = bz == 0;
  const bool lx = bx == (int)nbx - 1, ly = by == (int)nby - 1, lz = bz == nbz - 1;
  for (uint32_t r = warp; r < rows; r += nwarps) {
    const int j = (int)(r % ty), k = (int)(r / ty);
    T* drow = dst + (size_t)r * tx;
    if (lod == 0) {
      // GetInputBrick: rows/lines that exist in the copy loops, addressed by linear file offset
      const bool in_y = j >= (fy ? o : 0) && j < (int)ty - (ly ? o : 0);
      const bool in_z = k >= (fz ? o : 0) && k < (int)tz - (lz ? o : 0);
      const int xa = fx ? o : 0, xb = (int)tx - (lx ? o : 0);
      const uint64_t lin0 = (uint64_t)(int64_t)(bx * (int)bc.core - o) + (uint64_t)(by * (int)bc.core - o + j) * sx +
                            (uint64_t)(int64_t)(bz * (int)bc.core - o + k - (int)sv.z0) * sx * sy;      // wraps for i < xa only; never read there
      for (int i = (int)lane; i < (int)tx; i += 32) {
        T val = 0;
        if (in_y && in_z && i >= xa && i < xb) {
          const uint64_t lin = lin0 + (uint64_t)i;
          if (lin < nvox) val = flat[lin];
        }
        drow[i] = val;
        lo = min(lo, (uint32_t)val);
        hi = max(hi, (uint32_t)val);
      }
    } else {
      const bool inner_row = j >= o && j < (int)ty - o && k >= o && k < (int)tz - o;
      const T* srow = flat + (uint64_t)(by * (int)bc.core + j - o) * sx + (uint64_t)(bz * (int)bc.core + k - o) * sx * sy +
                      (uint64_t)(int64_t)(bx * (int)bc.core - o);
      for (int i = (int)lane; i < (int)tx; i += 32) {
        T val = 0;
        if (inner_row && i >= o && i < (int)tx - o) {
          val = srow[i];
        } else {
          uint32_t gx, gy, gz;
          if (resolve_overlap(bc, lod, bx, by, bz, i, j, k, gx, gy, gz)) val = flat[gx + (uint64_t)gy * sx + (uint64_t)gz * sx * sy];
        }
        drow[i] = val;
        lo = min(lo, (uint32_t)val);
        hi = max(hi, (uint32_t)val);
      }
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(mn + first_brick + b, lo);
    atomicMax(mx + first_brick + b, hi);
  }
}

// Clamp-to-edge borders (bClampToEdge, ExtendedOctreeConverter.cpp:288-373 GetInputBrick + :376-462 ClampToEdge, and
// FillOverlap + ClampToEdge for the coarser levels, :1203-1372).  What the reference's sequence of neighbour copies
// followed by the x, y, z replications leaves in a brick is, voxel for voxel, the LOD's flat volume at the brick voxel's
// global coordinate CLAMPED to the domain (checked against the compiled reference, tests/test_gpu_builder.py) -- with one
// exception: the ghost corners FillOverlap never writes (high ghost on two axes, low ghost on the third, all neighbours
// present; 8 voxels each) keep whatever the converter's reused, un-cleared temp buffer held from an earlier brick in
// that mode (ExtendedOctreeConverter.inc:261).  That content is an accident of processing history; this kernel writes
// the voxel at the global coordinate there as everywhere else.
template <typename T>
__global__ void __launch_bounds__(256) cut_bricks_clamp_kernel(const T* __restrict__ flat, uint8_t* __restrict__ store,
                                                                const uint64_t* __restrict__ offsets, uint64_t first_brick,
                                                                uint32_t lod, const __grid_constant__ BuildConst bc,
                                                                uint32_t* __restrict__ mn, uint32_t* __restrict__ mx,
                                                                uint32_t x_first, uint32_t x_count, SlabView sv) {
  const uint32_t nbx = bc.lod_bricks[lod][0], nby = bc.lod_bricks[lod][1];
  const int bx = (int)(x_first + blockIdx.x % x_count), by = (int)((blockIdx.x / x_count) % nby), bz = (int)(sv.bz_first + blockIdx.x / (x_count * nby));
  const uint32_t b = (uint32_t)bx + (uint32_t)by * nbx + (uint32_t)bz * nbx * nby;
  const uint32_t tx = brick_dim(bc, lod, 0, bx), ty = brick_dim(bc, lod, 1, by), tz = brick_dim(bc, lod, 2, bz);
  const int sx = (int)bc.lod_size[lod][0], sy = (int)bc.lod_size[lod][1], sz = (int)bc.lod_size[lod][2];
  T* dst = reinterpret_cast<T*>(store + offsets[first_brick + b]);
  const int o = (int)bc.overlap;
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.y * blockDim.x) >> 5;
  const uint32_t rows = ty * tz;
  const int nbz = (int)bc.lod_bricks[lod][2];
  const bool fx = bx == 0, fy = by == 0, fz = bz == 0;
  const bool lx = bx == (int)nbx - 1, ly = by == (int)nby - 1, lz = bz == nbz - 1;
  const uint64_t nvox = (uint64_t)sx * sy * sv.rows;
  for (uint32_t r = warp; r < rows; r += nwarps) {
    const int j = (int)(r % ty), k = (int)(r / ty);
    const int gy = min(max(by * (int)bc.core - o + j, 0), sy - 1);
    const int gz = min(max(bz * (int)bc.core - o + k, 0), sz - 1) - (int)sv.z0;      // (the slab holds every row the bricks' clamped range needs)
    const T* srow = flat + (uint64_t)gy * sx + (uint64_t)gz * sx * sy;
    T* drow = dst + (size_t)r * tx;
    const int x0 = bx * (int)bc.core - o;
    if (lod == 0) {
      // GetInputBrick reads scan lines by LINEAR file offset (a line towards a neighbour whose interior is thinner than
      // the overlap runs on into the next row / slice), then ClampToEdge replicates on the domain faces: the brick
      // voxel clamped into the box the copy loops fill, read at its linear offset; past the end of the file (where the
      // reference's buffer keeps stale bytes) the voxel at the clamped coordinate.
      const int jc = min(max(j, fy ? o : 0), (int)ty - (ly ? o : 0) - 1), kc = min(max(k, fz ? o : 0), (int)tz - (lz ? o : 0) - 1);
      const int xa = fx ? o : 0, xb = (int)tx - (lx ? o : 0);
      const uint64_t lin0 = (uint64_t)(int64_t)x0 + (uint64_t)(by * (int)bc.core - o + jc) * sx +
                            (uint64_t)(int64_t)(bz * (int)bc.core - o + kc - (int)sv.z0) * sx * sy;
      for (int i = (int)lane; i < (int)tx; i += 32) {
        const uint64_t lin = lin0 + (uint64_t)min(max(i, xa), xb - 1);
        const T val = lin < nvox ? flat[lin] : srow[min(max(x0 + i, 0), sx - 1)];
        drow[i] = val;
        lo = min(lo, (uint32_t)val);
        hi = max(hi, (uint32_t)val);
      }
      continue;
    }
    for (int i = (int)lane; i < (int)tx; i += 32) {
      const T val = srow[min(max(x0 + i, 0), sx - 1)];
      drow[i] = val;
      lo = min(lo, (uint32_t)val);
      hi = max(hi, (uint32_t)val);
    }
  }
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(mn + first_brick + b, lo);
    atomicMax(mx + first_brick + b, hi);
  }
}

// Vector form of cut_bricks_kernel: a thread produces one aligned CB-byte chunk (16, 8 or 4 bytes,
// whatever divides the brick's row length) per trip, so loads and stores are word-granular.
//  * LOD 0: the chunk is CB bytes at a linear file offset, masked to the x range GetInputBrick copies;
//  * LOD >= 1: the FillOverlap replay of resolve_overlap() has a closed form (checked exhaustively
//    against the replay; needs core >= overlap): with P / M the number of axes on which the voxel is a
//    high / low ghost, the voxel is ZERO iff its global coordinate lies outside the LOD's domain or
//    (P == 2 and M == 1) -- a diagonal two-axis copy runs before the low-side copies and finds the
//    third axis still unfilled -- and otherwise it is the voxel at its global coordinate.  Along a row
//    that is a domain mask plus "drop the low-x ghosts if both y and z are high ghosts" / "drop the
//    high-x ghosts if one of y,z is a high and the other a low ghost".
// Min/max accumulate packed (vminu4 / vminu2) and unpack once per thread.
// The grid covers the bricks bx in [x_first, x_first + x_count) of every brick row, so that the
// full-width bricks and the (narrower) last column get a kernel each, without idle CTAs.
template <int CB> struct Chunk;
template <> struct Chunk<16> { typedef uint4 type; };
template <> struct Chunk<8> { typedef uint2 type; };
template <> struct Chunk<4> { typedef uint32_t type; };

template <typename T, int CB>
__global__ void __launch_bounds__(256) cut_bricks_vec_kernel(const T* __restrict__ flat, uint8_t* __restrict__ store,
                                                              const uint64_t* __restrict__ offsets, uint64_t first_brick,
                                                              uint32_t lod, const __grid_constant__ BuildConst bc,
                                                              uint32_t* __restrict__ mn, uint32_t* __restrict__ mx,
                                                              uint32_t x_first, uint32_t x_count, SlabView sv) {
  constexpr int E = CB / (int)sizeof(T), NW = CB / 4;
  const uint32_t nbx = bc.lod_bricks[lod][0], nby = bc.lod_bricks[lod][1];
  const int nbz = (int)bc.lod_bricks[lod][2];
  const int bx = (int)(x_first + blockIdx.x % x_count), by = (int)((blockIdx.x / x_count) % nby), bz = (int)(sv.bz_first + blockIdx.x / (x_count * nby));
  const uint32_t b = (uint32_t)bx + (uint32_t)by * nbx + (uint32_t)bz * nbx * nby;
  const uint32_t tx = brick_dim(bc, lod, 0, bx), ty = brick_dim(bc, lod, 1, by), tz = brick_dim(bc, lod, 2, bz);
  const uint32_t sx = bc.lod_size[lod][0], sy = bc.lod_size[lod][1], sz = bc.lod_size[lod][2];
  const uint64_t nbytes = (uint64_t)sx * sy * sv.rows * sizeof(T);      // of the buffer `flat` points at (a z-slab of the LOD or all of it)
  const uint8_t* fb = reinterpret_cast<const uint8_t*>(flat);
  typename Chunk<CB>::type* dst = reinterpret_cast<typename Chunk<CB>::type*>(store + offsets[first_brick + b]);
  const int o = (int)bc.overlap, core = (int)bc.core;
  const bool fx = bx == 0, fy = by == 0, fz = bz == 0;
  const bool lx = bx == (int)nbx - 1, ly = by == (int)nby - 1, lz = bz == nbz - 1;
  const uint32_t cpr = tx / E, nchunks = cpr * ty * tz;
  uint32_t pmn = 0xFFFFFFFFu, pmx = 0u;
  for (uint32_t c = blockIdx.y * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.y * blockDim.x) {
    const uint32_t r = c / cpr;
    const int i0 = (int)(c - r * cpr) * E, j = (int)(r % ty), k = (int)(r / ty);
    int lo, hi;            // element range of the row that is copied; everything else stays zero
    bool row_ok;
    if (lod == 0) {
      row_ok = j >= (fy ? o : 0) && j < (int)ty - (ly ? o : 0) && k >= (fz ? o : 0) && k < (int)tz - (lz ? o : 0);
      lo = fx ? o : 0; hi = (int)tx - (lx ? o : 0);
    } else {
      const int gy = by * core + j - o, gz = bz * core + k - o;
      row_ok = gy >= 0 && gy < (int)sy && gz >= 0 && gz < (int)sz;
      const int py = j >= (int)ty - o, pz = k >= (int)tz - o, my = j < o, mz = k < o;
      lo = max(0, o - bx * core); hi = min((int)tx, (int)sx - bx * core + o);      // the domain along x
      if (py + pz == 2) lo = max(lo, o);                                            // P == 2, M == 1 at the low-x ghosts
      if (py + pz == 1 && my + mz == 1) hi = min(hi, (int)tx - o);                  // ... and at the high-x ghosts
    }
    lo = max(lo, i0); hi = min(hi, i0 + E);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok && lo < hi) {
      const int64_t lin = (int64_t)(bx * core - o + i0) + (int64_t)(by * core - o + j) * sx + (int64_t)(bz * core - o + k - (int)sv.z0) * sx * sy;
      v = load_chunk<CB>(fb, lin * (int64_t)sizeof(T), nbytes);
      if (lo > i0 || hi < i0 + E) v = keep16(v, (lo - i0) * (int)sizeof(T), (hi - i0) * (int)sizeof(T));
    }
    if (NW == 4) reinterpret_cast<uint4*>(dst)[c] = v;
    else if (NW == 2) reinterpret_cast<uint2*>(dst)[c] = make_uint2(v.x, v.y);
    else reinterpret_cast<uint32_t*>(dst)[c] = v.x;
    if (NW == 4) vec_minmax<T>(v, pmn, pmx);
    else if (NW == 2) vec_minmax<T>(make_uint4(v.x, v.y, v.x, v.y), pmn, pmx);
    else vec_minmax<T>(make_uint4(v.x, v.x, v.x, v.x), pmn, pmx);
  }
  uint32_t lo, hi;
  if (sizeof(T) == 1) {
    lo = min(min(pmn & 0xFFu, (pmn >> 8) & 0xFFu), min((pmn >> 16) & 0xFFu, pmn >> 24));
    hi = max(max(pmx & 0xFFu, (pmx >> 8) & 0xFFu), max((pmx >> 16) & 0xFFu, pmx >> 24));
  } else {
    lo = min(pmn & 0xFFFFu, pmn >> 16);
    hi = max(pmx & 0xFFFFu, pmx >> 16);
  }
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(mn + first_brick + b, lo);
    atomicMax(mx + first_brick + b, hi);
  }
}

// ------------------------------------------------------------------ fused LOD-0 pass
// One read of the raw volume produces everything LOD 0 contributes: the LOD-0 bricks (ghost layers and min/max
// included), the 1-D histogram and the flat LOD-1 volume.  A thread owns, for one CB-byte x chunk of a brick,
// the 2x2 rows (j, j+1) x (k, k+1): the brick's inner region starts at the (even) overlap and the inner size of
// a brick is even, so those rows and the chunk's voxel pairs are exactly the 2x2x2 blocks of the reduction
// (ghost rows / voxels pair up among themselves and are only copied).  Every raw voxel is an inner voxel of
// exactly one brick, which makes the histogram and the LOD-1 volume complete.  Persistent: one CTA of 1024
// threads per SM walks the bricks bx in [x_first, x_first + x_count) of the lattice; histogram table in shared
// memory as in downsample_hist_kernel.  Needs an even overlap and LOD-0 extents > 1.
// STATUS: bit-exact, ~35 % less DRAM traffic, but one 1024-thread CTA per SM at 64 registers leaves too few loads
// in flight: 2048^3 u16 takes 25.0 ms fused vs 22.4 ms with the separate kernels.  Opt-in (TVK_BUILD_FUSED=1).
template <typename T, int CB>
__device__ __forceinline__ uint4 lod0_chunk(const uint8_t* __restrict__ fb, uint64_t nbytes, uint32_t sx, uint32_t sy,
                                            int bx, int by, int bz, int core, int o, int i0, int j, int k,
                                            bool row_ok, int xlo, int xhi) {
  constexpr int E = CB / (int)sizeof(T);
  const int lo = max(xlo, i0), hi = min(xhi, i0 + E);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (row_ok && lo < hi) {
    const int64_t lin = (int64_t)(bx * core - o + i0) + (int64_t)(by * core - o + j) * sx + (int64_t)(bz * core - o + k) * sx * sy;
    v = load_chunk<CB>(fb, lin * (int64_t)sizeof(T), nbytes);
    if (lo > i0 || hi < i0 + E) v = keep16(v, (lo - i0) * (int)sizeof(T), (hi - i0) * (int)sizeof(T));
  }
  return v;
}
template <int CB> __device__ __forceinline__ void store_chunk(void* dst, uint32_t c, const uint4& v) {
  if (CB == 16) reinterpret_cast<uint4*>(dst)[c] = v;
  else if (CB == 8) reinterpret_cast<uint2*>(dst)[c] = make_uint2(v.x, v.y);
  else reinterpret_cast<uint32_t*>(dst)[c] = v.x;
}
template <typename T, int CB> __device__ __forceinline__ void chunk_minmax(const uint4& v, uint32_t& pmn, uint32_t& pmx) {
  if (CB == 16) vec_minmax<T>(v, pmn, pmx);
  else if (CB == 8) vec_minmax<T>(make_uint4(v.x, v.y, v.x, v.y), pmn, pmx);
  else vec_minmax<T>(make_uint4(v.x, v.x, v.x, v.x), pmn, pmx);
}
// histogram of the chunk's elements [e_lo, e_hi)
template <typename T, int CB>
__device__ __forceinline__ void hist_add_range(const uint4& q, int e_lo, int e_hi, unsigned int* __restrict__ table) {
  constexpr int E = CB / (int)sizeof(T);
  if (sizeof(T) == 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) if (e >= e_lo && e < e_hi) atomicAdd(&table[vec_get<T>(q, e)], 1u);
  } else {
    uint32_t cur = 0xFFFFFFFFu, cnt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (e < e_lo || e >= e_hi) continue;
      const uint32_t v = vec_get<T>(q, e);
      if (v == cur) cnt++;
      else {
        if (cnt && cur < HIST_WINDOW) atomicAdd(&table[cur], cnt);
        cur = v; cnt = 1;
      }
    }
    if (cnt && cur < HIST_WINDOW) atomicAdd(&table[cur], cnt);
  }
}

template <typename T, int CB, bool MEDIAN>
__global__ void __launch_bounds__(1024, 1) lod0_fused_kernel(const T* __restrict__ raw, uint8_t* __restrict__ store,
                                                             const uint64_t* __restrict__ offsets, const __grid_constant__ BuildConst bc,
                                                             uint32_t* __restrict__ mn, uint32_t* __restrict__ mx,
                                                             uint32_t x_first, uint32_t x_count, T* __restrict__ lod1,
                                                             unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned int hsm[];
  constexpr uint32_t TABLE = sizeof(T) == 1 ? 32u * 256u : HIST_WINDOW;
  constexpr int E = CB / (int)sizeof(T);
  for (uint32_t i = threadIdx.x; i < TABLE; i += blockDim.x) hsm[i] = 0u;
  __syncthreads();
  unsigned int* mine = sizeof(T) == 1 ? hsm + (threadIdx.x >> 5) * 256u : hsm;
  const uint32_t nbx = bc.lod_bricks[0][0], nby = bc.lod_bricks[0][1], nbz = bc.lod_bricks[0][2];
  const uint32_t sx = bc.lod_size[0][0], sy = bc.lod_size[0][1], sz = bc.lod_size[0][2];
  const uint32_t dx = bc.lod_size[1][0], dy = bc.lod_size[1][1];
  const uint64_t nbytes = (uint64_t)sx * sy * sz * sizeof(T);
  const uint8_t* fb = reinterpret_cast<const uint8_t*>(raw);
  const int o = (int)bc.overlap, core = (int)bc.core;
  const uint32_t nbricks = x_count * nby * nbz;
  for (uint32_t bi = blockIdx.x; bi < nbricks; bi += gridDim.x) {
    const int bx = (int)(x_first + bi % x_count), by = (int)((bi / x_count) % nby), bz = (int)(bi / (x_count * nby));
    const uint32_t b = (uint32_t)bx + (uint32_t)by * nbx + (uint32_t)bz * nbx * nby;
    const uint32_t tx = brick_dim(bc, 0, 0, bx), ty = brick_dim(bc, 0, 1, by), tz = brick_dim(bc, 0, 2, bz);
    void* dst = store + offsets[b];
    const bool fx = bx == 0, fy = by == 0, fz = bz == 0;
    const bool lx = bx == (int)nbx - 1, ly = by == (int)nby - 1, lz = bz == (int)nbz - 1;
    const int xlo = fx ? o : 0, xhi = (int)tx - (lx ? o : 0);
    const uint32_t cpr = tx / E, jp_n = (ty + 1) / 2, kp_n = (tz + 1) / 2, units = cpr * jp_n * kp_n;
    uint32_t pmn = 0xFFFFFFFFu, pmx = 0u;
    for (uint32_t u = threadIdx.x; u < units; u += blockDim.x) {
      const uint32_t cx = u % cpr, rp = u / cpr;
      const int i0 = (int)cx * E, j = 2 * (int)(rp % jp_n), k = 2 * (int)(rp / jp_n);
      const bool j1 = j + 1 < (int)ty, k1 = k + 1 < (int)tz;
      auto ok = [&](int jj, int kk) {
        return jj >= (fy ? o : 0) && jj < (int)ty - (ly ? o : 0) && kk >= (fz ? o : 0) && kk < (int)tz - (lz ? o : 0);
      };
      const uint4 a = lod0_chunk<T, CB>(fb, nbytes, sx, sy, bx, by, bz, core, o, i0, j, k, ok(j, k), xlo, xhi);
      uint4 bq = a, c = a, d = a;
      store_chunk<CB>(dst, cx + cpr * ((uint32_t)j + ty * (uint32_t)k), a);
      chunk_minmax<T, CB>(a, pmn, pmx);
      if (k1) { bq = lod0_chunk<T, CB>(fb, nbytes, sx, sy, bx, by, bz, core, o, i0, j, k + 1, ok(j, k + 1), xlo, xhi);
                store_chunk<CB>(dst, cx + cpr * ((uint32_t)j + ty * (uint32_t)(k + 1)), bq); chunk_minmax<T, CB>(bq, pmn, pmx); }
      if (j1) { c = lod0_chunk<T, CB>(fb, nbytes, sx, sy, bx, by, bz, core, o, i0, j + 1, k, ok(j + 1, k), xlo, xhi);
                store_chunk<CB>(dst, cx + cpr * ((uint32_t)(j + 1) + ty * (uint32_t)k), c); chunk_minmax<T, CB>(c, pmn, pmx); }
      if (j1 && k1) { d = lod0_chunk<T, CB>(fb, nbytes, sx, sy, bx, by, bz, core, o, i0, j + 1, k + 1, ok(j + 1, k + 1), xlo, xhi);
                      store_chunk<CB>(dst, cx + cpr * ((uint32_t)(j + 1) + ty * (uint32_t)(k + 1)), d); chunk_minmax<T, CB>(d, pmn, pmx); }
      // inner part: rows j / k inner?  (pairs never straddle the inner boundary: the overlap is even)
      const bool inner_j = j >= o && j < (int)ty - o, inner_k = k >= o && k < (int)tz - o;
      if (!inner_j || !inner_k) continue;
      const bool hy = j + 1 < (int)ty - o, hz = k + 1 < (int)tz - o;
      const int e_lo = max(o - i0, 0), e_hi = min((int)tx - o - i0, E);
      if (e_lo >= e_hi) continue;
      hist_add_range<T, CB>(a, e_lo, e_hi, mine);
      if (hz) hist_add_range<T, CB>(bq, e_lo, e_hi, mine);
      if (hy) hist_add_range<T, CB>(c, e_lo, e_hi, mine);
      if (hy && hz) hist_add_range<T, CB>(d, e_lo, e_hi, mine);
      const uint32_t Y = (uint32_t)(by * core + j - o) >> 1, Z = (uint32_t)(bz * core + k - o) >> 1;
      T* orow = lod1 + (uint64_t)Y * dx + (uint64_t)Z * dx * dy;
#pragma unroll
      for (int m = 0; m < E / 2; ++m) {
        const int e0 = 2 * m;
        if (e0 < e_lo || e0 >= e_hi) continue;
        const bool hx = e0 + 1 < e_hi;
        const T a0 = (T)vec_get<T>(a, e0), a1 = (T)vec_get<T>(a, e0 + 1), b0 = (T)vec_get<T>(bq, e0), b1 = (T)vec_get<T>(bq, e0 + 1);
        const T c0 = (T)vec_get<T>(c, e0), c1 = (T)vec_get<T>(c, e0 + 1), d0 = (T)vec_get<T>(d, e0), d1 = (T)vec_get<T>(d, e0 + 1);
        T r;
        // argument order = p0..p7 of DownsampleBricktoBrick (see downsample_kernel)
        if (hx && hy && hz) r = filter8<T, MEDIAN>(a0, b0, c0, d0, a1, b1, c1, d1);
        else if (!hx && hy && hz) r = filter4<T, MEDIAN>(a0, b0, c0, d0);
        else if (hx && !hy && hz) r = filter4<T, MEDIAN>(a0, b0, a1, b1);
        else if (hx && hy && !hz) r = filter4<T, MEDIAN>(a0, c0, a1, c1);
        else if (!hx && !hy && hz) r = filter2<T, MEDIAN>(a0, b0);
        else if (hx && !hy && !hz) r = filter2<T, MEDIAN>(a0, a1);
        else if (!hx && hy && !hz) r = filter2<T, MEDIAN>(a0, c0);
        else r = a0;
        orow[(uint32_t)(bx * core + i0 + e0 - o) >> 1] = r;
      }
    }
    uint32_t lo, hi;
    if (sizeof(T) == 1) {
      lo = min(min(pmn & 0xFFu, (pmn >> 8) & 0xFFu), min((pmn >> 16) & 0xFFu, pmn >> 24));
      hi = max(max(pmx & 0xFFu, (pmx >> 8) & 0xFFu), max((pmx >> 16) & 0xFFu, pmx >> 24));
    } else {
      lo = min(pmn & 0xFFFFu, pmn >> 16);
      hi = max(pmx & 0xFFFFu, pmx >> 16);
    }
    for (int s = 16; s > 0; s >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
      atomicMin(mn + b, lo);
      atomicMax(mx + b, hi);
    }
  }
  __syncthreads();
  if (sizeof(T) == 1) {
    for (uint32_t t = threadIdx.x; t < 256u; t += blockDim.x) {
      unsigned int sum = 0;
      for (uint32_t w = 0; w < 32u; ++w) sum += hsm[w * 256u + t];
      if (sum) atomicAdd(&bins[t], (unsigned long long)sum);
    }
  } else {
    for (uint32_t t = threadIdx.x; t < HIST_WINDOW; t += blockDim.x) {
      const unsigned int sum = hsm[t];
      if (sum) atomicAdd(&bins[t], (unsigned long long)sum);
    }
  }
}

__global__ void minmax_to_double_kernel(const uint32_t* mn, const uint32_t* mx, double* out, uint64_t n, uint32_t* gmax) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0;
  if (i < n) { out[2 * i] = (double)mn[i]; out[2 * i + 1] = (double)mx[i]; m = mx[i]; }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(gmax, m);
}

// 16-bit histogram, one pass per window of HIST_WINDOW bins: every CTA (one per SM, 1024 threads)
// counts its share of the volume into a private 128 KB shared-memory table and adds the non-zero
// entries to the global bins at the end, so the atomics that smooth data aims at a handful of bins
// never leave the SM.  Eight voxels per 16-byte load; equal neighbours are run-length merged and a
// warp that sees one single value issues one atomic.  A window above the volume's maximum (known from
// the brick min/max table, read on the device) returns at once.
__global__ void __launch_bounds__(1024, 1) hist16_window_kernel(const uint16_t* __restrict__ raw, uint64_t n, uint32_t lo,
                                                                 const uint32_t* __restrict__ gmax,
                                                                 unsigned long long* __restrict__ bins) {
  extern __shared__ uint32_t sh[];
  if (lo > *gmax) return;
  for (uint32_t i = threadIdx.x; i < HIST_WINDOW; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const uint64_t head = min((unsigned long long)n, (unsigned long long)(((16u - (reinterpret_cast<uintptr_t>(raw) & 15u)) & 15u) / 2u));   // voxels before 16-byte alignment
  const uint64_t nvec = (n - head) / 8u;
  const uint4* pv = reinterpret_cast<const uint4*>(raw + head);
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gstride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (nvec + gstride - 1) / gstride;
  uint64_t i = gtid;
  for (uint64_t r = 0; r < rounds; ++r, i += gstride) {
    const bool act = i < nvec;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (act) q = __ldg(pv + i);
    const uint32_t v0 = q.x & 0xFFFFu;
    const bool uni = act && q.x == (v0 | (v0 << 16)) && q.y == q.x && q.z == q.x && q.w == q.x;
    const uint32_t first = __shfl_sync(0xffffffffu, v0, 0);      // (outside the &&: every lane must take part)
    if (__all_sync(0xffffffffu, uni && v0 == first)) {
      if ((threadIdx.x & 31) == 0 && v0 - lo < HIST_WINDOW) atomicAdd(&sh[v0 - lo], 256u);
      continue;
    }
    if (!act) continue;
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t cur = v0, cnt = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t v = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
      if (v == cur) { cnt++; }
      else {
        if (cur - lo < HIST_WINDOW) atomicAdd(&sh[cur - lo], cnt);
        cur = v; cnt = 1;
      }
    }
    if (cur - lo < HIST_WINDOW) atomicAdd(&sh[cur - lo], cnt);
  }
  // unaligned head and the tail
  for (uint64_t t = gtid; t < head; t += gstride) { const uint32_t v = raw[t]; if (v - lo < HIST_WINDOW) atomicAdd(&sh[v - lo], 1u); }
  for (uint64_t t = head + nvec * 8u + gtid; t < n; t += gstride) { const uint32_t v = raw[t]; if (v - lo < HIST_WINDOW) atomicAdd(&sh[v - lo], 1u); }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < HIST_WINDOW; t += blockDim.x) {
    const uint32_t cnt = sh[t];
    if (cnt) atomicAdd(&bins[lo + t], (unsigned long long)cnt);
  }
}

// exact histogram of the raw volume (== LOD-0 inner voxels); per-CTA shared
// histogram for 8 bit, warp-aggregated global atomics for 16 bit
template <typename T>
__global__ void __launch_bounds__(256) hist1d_kernel(const T* __restrict__ raw, uint64_t n, unsigned long long* __restrict__ bins) {
  if (sizeof(T) == 1) {
    // one private 256-bin histogram per warp (8 KB per CTA): smooth data sends a whole warp to a
    // couple of bins, which a single shared table would serialise CTA-wide; 16-byte loads
    __shared__ unsigned int sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    unsigned int* mine = sh[threadIdx.x >> 5];
    const uint8_t* p = reinterpret_cast<const uint8_t*>(raw);
    const uint64_t nvec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? n / 16 : 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
      const uint4 q = reinterpret_cast<const uint4*>(p)[i];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&mine[w[k] & 0xFFu], 1u); atomicAdd(&mine[(w[k] >> 8) & 0xFFu], 1u);
        atomicAdd(&mine[(w[k] >> 16) & 0xFFu], 1u); atomicAdd(&mine[w[k] >> 24], 1u);
      }
    }
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
      atomicAdd(&mine[p[i]], 1u);
    __syncthreads();
    unsigned int s = 0;
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    if (s) atomicAdd(&bins[threadIdx.x], (unsigned long long)s);
  } else {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, i += stride) {
      const bool act = i < n;
      const unsigned am = __ballot_sync(0xffffffffu, act);
      if (act) {
        const uint32_t v = raw[i];
        const unsigned peers = __match_any_sync(am, v);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&bins[v], (unsigned long long)__popc(peers));
      }
    }
  }
}

// Histogram2DDataBlock::ComputeGradient on the raw volume: the +-1 neighbours of an
// inner LOD-0 brick voxel are the raw neighbours, zero outside the domain.
// All arithmetic in fp64 with explicit round-to-nearest ops (no FMA contraction)
// so that it matches the reference's x86-64 doubles bit for bit.
// `clamp`: the LOD-0 bricks were built with clamp-to-edge borders, so a neighbour outside the domain is the voxel itself
// (Histogram2DDataBlock reads the bricks' ghost cells, whatever the converter put there) instead of zero.
template <typename T>
__device__ __forceinline__ double grad_mag(const T* __restrict__ raw, uint32_t x, uint32_t y, uint32_t z,
                                           uint32_t sx, uint32_t sy, uint32_t sz, double norm2, bool clamp = false) {
  const uint64_t rs = sx, ss = (uint64_t)sx * sy;
  const uint64_t c = x + y * rs + z * ss;
  const double e = clamp ? (double)raw[c] : 0.0;
  const double l = x > 0 ? (double)raw[c - 1] : e, r = x + 1 < sx ? (double)raw[c + 1] : e;
  const double t = y > 0 ? (double)raw[c - rs] : e, b = y + 1 < sy ? (double)raw[c + rs] : e;
  const double f = z > 0 ? (double)raw[c - ss] : e, k = z + 1 < sz ? (double)raw[c + ss] : e;
  const double gx = __ddiv_rn(__dsub_rn(l, r), norm2), gy = __ddiv_rn(__dsub_rn(t, b), norm2), gz = __ddiv_rn(__dsub_rn(f, k), norm2);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz)));
}

// The fp64 chain of grad_mag() (3 divisions + a square root per voxel, twice over the volume) is what the
// reference evaluates, and it must be matched bit for bit -- but it is needed for very few voxels:
//  * the gradient length is, up to ~4 ulp, sqrt(S)/norm2 with the INTEGER S = dx^2+dy^2+dz^2 of the three
//    central differences; two different S differ by a relative 1/S >= 7.7e-11, far beyond the rounding noise,
//    so the maximum length is attained at a voxel with the maximum S.  Pass 1 therefore runs in integers and
//    evaluates the fp64 chain only for voxels that tie or beat the thread's current best S;
//  * the gradient bin floor(len/max*255) follows from S/Smax in fp32 whenever that estimate is not within
//    1e-3 of a bin boundary (fp32 error < 1e-4); only the boundary cases run the exact fp64 chain.
template <typename T>
__device__ __forceinline__ unsigned long long grad_s(const T* __restrict__ raw, uint32_t x, uint32_t y, uint32_t z,
                                                     uint32_t sx, uint32_t sy, uint32_t sz) {
  const uint64_t rs = sx, ss = (uint64_t)sx * sy;
  const uint64_t c = x + y * rs + z * ss;
  const long long l = x > 0 ? (long long)raw[c - 1] : 0, r = x + 1 < sx ? (long long)raw[c + 1] : 0;
  const long long t = y > 0 ? (long long)raw[c - rs] : 0, b = y + 1 < sy ? (long long)raw[c + rs] : 0;
  const long long f = z > 0 ? (long long)raw[c - ss] : 0, k = z + 1 < sz ? (long long)raw[c + ss] : 0;
  return (unsigned long long)((l - r) * (l - r) + (t - b) * (t - b) + (f - k) * (f - k));
}

// the same S with everything that depends on (y,z) hoisted out of the row loop: pointers to the four
// neighbouring rows (nullptr outside the domain) and 32-bit arithmetic (a difference is < 2^16)
template <typename T> struct RowCtx { const T *line, *up, *down, *front, *back; bool clamp; };
template <typename T>
__device__ __forceinline__ RowCtx<T> row_ctx(const T* __restrict__ raw, uint32_t y, uint32_t z, uint32_t sx, uint32_t sy, uint32_t sz,
                                             bool clamp = false) {
  const uint64_t rs = sx, ss = (uint64_t)sx * sy;
  RowCtx<T> r;
  r.line = raw + y * rs + z * ss;
  const T* outside = clamp ? r.line : nullptr;          // clamp-to-edge: the missing neighbour row is the row itself
  r.up = y > 0 ? r.line - rs : outside; r.down = y + 1 < sy ? r.line + rs : outside;
  r.front = z > 0 ? r.line - ss : outside; r.back = z + 1 < sz ? r.line + ss : outside;
  r.clamp = clamp;
  return r;
}
template <typename T>
__device__ __forceinline__ unsigned long long grad_s_row(const RowCtx<T>& r, uint32_t x, uint32_t sx, unsigned long long* key = nullptr) {
  const int e = r.clamp ? (int)r.line[x] : 0;
  const int l = x > 0 ? (int)r.line[x - 1] : e, rr = x + 1 < sx ? (int)r.line[x + 1] : e;
  const int t = r.up ? (int)r.up[x] : 0, b = r.down ? (int)r.down[x] : 0;
  const int f = r.front ? (int)r.front[x] : 0, k = r.back ? (int)r.back[x] : 0;
  const int dx = l - rr, dy = t - b, dz = f - k;
  // the fp64 length depends on (|dx|,|dy|,|dz|) only: this key identifies it
  if (key) *key = (unsigned long long)abs(dx) | ((unsigned long long)abs(dy) << 17) | ((unsigned long long)abs(dz) << 34);
  if (sizeof(T) == 1) return (unsigned long long)(unsigned)(dx * dx + dy * dy + dz * dz);
  // 16-bit: a square reaches 2^32 - 2^17, beyond int -- multiply as unsigned (signed overflow is undefined and the
  // compiler did sign-extend the wrapped product into the 64-bit sum)
  const unsigned ux = (unsigned)abs(dx), uy = (unsigned)abs(dy), uz = (unsigned)abs(dz);
  return (unsigned long long)ux * ux + (unsigned long long)uy * uy + (unsigned long long)uz * uz;
}

struct GradBest { unsigned long long s; unsigned long long len_bits; };

template <typename T>
__global__ void __launch_bounds__(256) grad_max_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                        double norm2, GradBest* __restrict__ partial, bool clamp) {
  // a warp walks one row (y,z) at a time, 32 voxels per trip: no per-voxel index division
  const uint32_t rows = sy * sz, lane = threadIdx.x & 31u;
  unsigned long long best_s = 0, best_bits = 0, last_key = ~0ull;
  for (uint32_t row = blockIdx.x * 8u + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8u) {
    const uint32_t y = row % sy, z = row / sy;
    const RowCtx<T> rc = row_ctx(raw, y, z, sx, sy, sz, clamp);
    for (uint32_t x = lane; x < sx; x += 32u) {
      unsigned long long key;
      const unsigned long long s = grad_s_row(rc, x, sx, &key);
      if (s == 0 || s < best_s || key == last_key) continue;      // (a tie with the triple just evaluated adds nothing)
      last_key = key;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(grad_mag(raw, x, y, z, sx, sy, sz, norm2, clamp));
      if (s > best_s) { best_s = s; best_bits = bits; }
      else if (bits > best_bits) best_bits = bits;          // non-negative doubles order like integers
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long os = __shfl_xor_sync(0xffffffffu, best_s, o), ob = __shfl_xor_sync(0xffffffffu, best_bits, o);
    if (os > best_s || (os == best_s && ob > best_bits)) { best_s = os; best_bits = ob; }
  }
  __shared__ GradBest sh[8];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5].s = best_s; sh[threadIdx.x >> 5].len_bits = best_bits; }
  __syncthreads();
  if (threadIdx.x == 0) {
    GradBest g = sh[0];
    for (int w = 1; w < 8; ++w) if (sh[w].s > g.s || (sh[w].s == g.s && sh[w].len_bits > g.len_bits)) g = sh[w];
    partial[blockIdx.x] = g;
  }
}

// value -> histogram column, exactly as Histogram2DDataBlock.cpp:432-435 computes it (one fp64 mul + div per VALUE)
__global__ void hist2d_value_lut_kernel(uint32_t range, uint32_t bins_x, double max_nonzero, uint16_t* __restrict__ lut) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= range) return;
  uint32_t vi = (max_nonzero <= (double)(bins_x - 1)) ? v : (uint32_t)__ddiv_rn(__dmul_rn((double)v, (double)(bins_x - 1)), max_nonzero);
  if (vi > bins_x - 1) vi = bins_x - 1;
  lut[v] = (uint16_t)vi;
}

template <typename T>
__global__ void __launch_bounds__(256) hist2d_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                      double norm2, double max_grad, unsigned long long smax, float inv_smax,
                                                      const uint16_t* __restrict__ vlut,
                                                      unsigned long long* __restrict__ bins, bool clamp) {
  // a warp's votes for one bin go out as one atomic; a leader whose bin repeats from trip to trip (the empty
  // background sends every warp to bin 0) keeps counting in a register and flushes when the bin changes
  uint32_t pend_bin = 0xFFFFFFFFu, pend_cnt = 0;
  const uint32_t rows = sy * sz, lane = threadIdx.x & 31u;
  for (uint32_t row = blockIdx.x * 8u + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8u) {
    const uint32_t y = row % sy, z = row / sy;
    const RowCtx<T> rc = row_ctx(raw, y, z, sx, sy, sz, clamp);
    const T* line = rc.line;
    for (uint32_t x0 = 0; x0 < sx; x0 += 32u) {
      const uint32_t x = x0 + lane;
      const bool act = x < sx;
      const unsigned am = __ballot_sync(0xffffffffu, act);
      if (!act) continue;
      const unsigned long long s = grad_s_row(rc, x, sx);
      // a volume without any gradient (Smax == 0): 0/0 = NaN, and size_t(NaN) is what x86-64 makes of it -- 2^63, clamped to
      // bin 255 by the std::min (Histogram2DDataBlock.cpp:428); every other zero gradient lands in bin 0
      uint32_t gi = smax == 0 ? 255u : 0u;
      if (s != 0) {
        const float q = 255.0f * sqrtf((float)s * inv_smax);
        const float fl = floorf(q);
        // S < Smax: the length is below the maximum by a relative 0.5/Smax >= 3.8e-11, so len/max*255 < 255 in fp64
        // too and everything from 254.5 up lands in bin 254; only ties with the maximum need the exact chain there
        if (s != smax && q >= 254.5f) gi = 254u;
        else if (s != smax && q - fl > 1e-3f && fl + 1.0f - q > 1e-3f) gi = (uint32_t)fl;
        else {
          // size_t(len/max*255.0f)
          const double e = __dmul_rn(__ddiv_rn(grad_mag(raw, x, y, z, sx, sy, sz, norm2, clamp), max_grad), 255.0);
          gi = (e >= 255.0) ? 255u : ((e > 0.0) ? (uint32_t)e : 0u);
        }
      }
      const uint32_t bin = (uint32_t)vlut[line[x]] * 256u + gi;
      const unsigned peers = __match_any_sync(am, bin);
      if (lane == (unsigned)(__ffs(peers) - 1)) {
        if (bin == pend_bin) pend_cnt += __popc(peers);
        else {
          if (pend_cnt) atomicAdd(&bins[pend_bin], (unsigned long long)pend_cnt);
          pend_bin = bin; pend_cnt = __popc(peers);
        }
      }
    }
  }
  if (pend_cnt) atomicAdd(&bins[pend_bin], (unsigned long long)pend_cnt);
}

// ---- 4 voxels per thread (rows that are a multiple of 4 voxels long, base 4/8-byte aligned) --------------
// One 4-/8-byte load per neighbouring row instead of one element load per neighbour and voxel, the three
// central differences of all four voxels with SIMD-in-a-word |a-b|, S by dp4a (8-bit) / 32x32->64 multiplies.
// Same decisions as the scalar kernels above; the voxels that need the exact fp64 chain are queued per warp
// and evaluated 32 at a time, so a warp never runs the chain for a single lane.
template <typename T> struct QuadOf;
template <> struct QuadOf<uint8_t> { typedef uint32_t type; };
template <> struct QuadOf<uint16_t> { typedef uint2 type; };
__device__ __forceinline__ uint32_t quad_zero(uint32_t) { return 0u; }
__device__ __forceinline__ uint2 quad_zero(uint2) { return make_uint2(0u, 0u); }
template <typename T>
__device__ __forceinline__ typename QuadOf<T>::type quad_load(const T* row, uint32_t q) {
  typedef typename QuadOf<T>::type Q;
  return row ? __ldg(reinterpret_cast<const Q*>(row) + q) : quad_zero(Q());
}
__device__ __forceinline__ uint32_t quad_shfl_up(uint32_t v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ uint32_t quad_shfl_down(uint32_t v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ uint2 quad_shfl_up(uint2 v) { return make_uint2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1)); }
__device__ __forceinline__ uint2 quad_shfl_down(uint2 v) { return make_uint2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1)); }

// S and the (|dx|,|dy|,|dz|) key of the four voxels of quad q; val[] = the voxels themselves.  All 32 lanes call.
template <typename T>
__device__ __forceinline__ void quad_grad(const RowCtx<T>& rc, uint32_t q, uint32_t nq, uint32_t lane, bool act,
                                          unsigned long long s[4], unsigned long long key[4], uint32_t val[4]) {
  typedef typename QuadOf<T>::type Q;
  const Q z = quad_zero(Q());
  const Q c = act ? quad_load<T>(rc.line, q) : z;
  Q prev = quad_shfl_up(c), next = quad_shfl_down(c);
  if (lane == 0) prev = (act && q > 0) ? quad_load<T>(rc.line, q - 1) : z;
  const bool next_act = __shfl_down_sync(0xffffffffu, (int)act, 1) != 0;       // (a tile may end before the row does)
  if (lane == 31 || !next_act) next = (act && q + 1 < nq) ? quad_load<T>(rc.line, q + 1) : z;
  const Q up = act ? quad_load<T>(rc.up, q) : z, dn = act ? quad_load<T>(rc.down, q) : z;
  const Q fr = act ? quad_load<T>(rc.front, q) : z, bk = act ? quad_load<T>(rc.back, q) : z;
  if constexpr (sizeof(T) == 1) {
    const uint32_t l = __byte_perm(prev, c, 0x6543), r = __byte_perm(c, next, 0x4321);
    const uint32_t adx = __vabsdiffu4(l, r), ady = __vabsdiffu4(up, dn), adz = __vabsdiffu4(fr, bk);
    const uint32_t t0 = __byte_perm(adx, ady, 0x5140), t1 = __byte_perm(adx, ady, 0x7362);
    const uint32_t u0 = __byte_perm(adz, 0u, 0x4140), u1 = __byte_perm(adz, 0u, 0x4342);
    const uint32_t k0 = __byte_perm(t0, u0, 0x5410), k1 = __byte_perm(t0, u0, 0x7632);
    const uint32_t k2 = __byte_perm(t1, u1, 0x5410), k3 = __byte_perm(t1, u1, 0x7632);
    key[0] = k0; key[1] = k1; key[2] = k2; key[3] = k3;
    s[0] = __dp4a(k0, k0, 0u); s[1] = __dp4a(k1, k1, 0u); s[2] = __dp4a(k2, k2, 0u); s[3] = __dp4a(k3, k3, 0u);
    val[0] = c & 0xffu; val[1] = (c >> 8) & 0xffu; val[2] = (c >> 16) & 0xffu; val[3] = c >> 24;
  } else {
    const uint32_t lx = __byte_perm(prev.y, c.x, 0x5432), ly = __byte_perm(c.x, c.y, 0x5432), ry = __byte_perm(c.y, next.x, 0x5432);
    const uint32_t ax[2] = {__vabsdiffu2(lx, ly), __vabsdiffu2(ly, ry)};
    const uint32_t ay[2] = {__vabsdiffu2(up.x, dn.x), __vabsdiffu2(up.y, dn.y)};
    const uint32_t az[2] = {__vabsdiffu2(fr.x, bk.x), __vabsdiffu2(fr.y, bk.y)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t dx = (i & 1) ? ax[i >> 1] >> 16 : ax[i >> 1] & 0xffffu;
      const uint32_t dy = (i & 1) ? ay[i >> 1] >> 16 : ay[i >> 1] & 0xffffu;
      const uint32_t dz = (i & 1) ? az[i >> 1] >> 16 : az[i >> 1] & 0xffffu;
      s[i] = (unsigned long long)dx * dx + (unsigned long long)dy * dy + (unsigned long long)dz * dz;
      key[i] = (unsigned long long)(dx | (dy << 16)) | ((unsigned long long)dz << 32);
    }
    val[0] = c.x & 0xffffu; val[1] = c.x >> 16; val[2] = c.y & 0xffffu; val[3] = c.y >> 16;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) grad_max_vec_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                            double norm2, GradBest* __restrict__ partial,
                                                            unsigned long long* __restrict__ shared_best) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nq = sx / 4u;
  unsigned long long best_s = 0, best_bits = 0, last_key = ~0ull;
  // a CTA works through blocks of 8 rows (y) x 16 slices (z): a row's z neighbours are then rows the same CTA touches one
  // trip earlier / later, and its y neighbours belong to the warps next to it -- they come from L2 / L1 instead of DRAM
  // (a plain row-major walk read 2.8x the volume on 2048^3)
  const uint32_t nyb = (sy + 7u) / 8u, nzb = (sz + 15u) / 16u, ntiles = nyb * nzb;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  for (uint32_t z = (tile / nyb) * 16u; z < min(sz, (tile / nyb) * 16u + 16u); ++z) {
    const uint32_t y = (tile % nyb) * 8u + warp;
    const bool row_valid = y < sy;
    const RowCtx<T> rc = row_ctx(raw, min(y, sy - 1u), z, sx, sy, sz);
    // only voxels that tie or beat the LARGEST S seen anywhere can hold the maximum: adopt the grid-wide bound once
    // per row, so that the exact chain stops running for every thread's local records and ties
    const unsigned long long gb = *reinterpret_cast<volatile unsigned long long*>(shared_best);
    if (gb > best_s) { best_s = gb; best_bits = 0; }
    for (uint32_t q0 = 0; q0 < nq; q0 += 32u) {
      const uint32_t q = q0 + lane;
      const bool act = row_valid && q < nq;
      unsigned long long s[4], key[4];
      uint32_t val[4];
      quad_grad<T>(rc, q, nq, lane, act, s, key, val);
      const unsigned long long m = max(max(s[0], s[1]), max(s[2], s[3]));
      if (m == 0 || m < best_s) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (s[i] == 0 || s[i] < best_s || key[i] == last_key) continue;
        last_key = key[i];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(grad_mag(raw, 4u * q + i, y, z, sx, sy, sz, norm2));
        if (s[i] > best_s) { best_s = s[i]; best_bits = bits; atomicMax(shared_best, best_s); }
        else if (bits > best_bits) best_bits = bits;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long os = __shfl_xor_sync(0xffffffffu, best_s, o), ob = __shfl_xor_sync(0xffffffffu, best_bits, o);
    if (os > best_s || (os == best_s && ob > best_bits)) { best_s = os; best_bits = ob; }
  }
  __shared__ GradBest sh[8];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5].s = best_s; sh[threadIdx.x >> 5].len_bits = best_bits; }
  __syncthreads();
  if (threadIdx.x == 0) {
    GradBest g = sh[0];
    for (int w = 1; w < 8; ++w) if (sh[w].s > g.s || (sh[w].s == g.s && sh[w].len_bits > g.len_bits)) g = sh[w];
    partial[blockIdx.x] = g;
  }
}

// exact bin of one voxel (the chain of hist2d_kernel's fall-back) and its vote
template <typename T>
__device__ __forceinline__ void hist2d_exact_vote(const T* __restrict__ raw, unsigned long long idx, uint32_t sx, uint32_t sy, uint32_t sz,
                                                  double norm2, double max_grad, const uint16_t* __restrict__ vlut,
                                                  unsigned long long* __restrict__ bins) {
  const uint32_t x = (uint32_t)(idx % sx), y = (uint32_t)((idx / sx) % sy), z = (uint32_t)(idx / ((unsigned long long)sx * sy));
  const double e = __dmul_rn(__ddiv_rn(grad_mag(raw, x, y, z, sx, sy, sz, norm2), max_grad), 255.0);
  const uint32_t gi = (e >= 255.0) ? 255u : ((e > 0.0) ? (uint32_t)e : 0u);
  atomicAdd(&bins[(uint32_t)vlut[raw[idx]] * 256u + gi], 1ull);
}

// Votes go through a per-CTA cache in shared memory: 4096 slots of (bin + 1) << 32 | count, direct mapped; a vote for
// the bin that owns the slot is one shared-memory CAS, a vote for another bin evicts the owner (its count goes to the
// global histogram with one atomic) and takes the slot.  Together with the tile traversal below -- a CTA works through
// compact 128 x 8 x 16-voxel blocks, whose voxels fall into few bins when the data is spatially coherent -- this takes
// most of the 64-bit L2 atomics, which bound this pass, off the table.
#define TVK_H2D_SLOTS 4096u
#ifndef TVK_H2D_ZT
#define TVK_H2D_ZT 16u
#endif
#ifndef TVK_H2D_XQ
#define TVK_H2D_XQ 0u
#endif
template <bool CACHED>
__device__ __forceinline__ void hist2d_cache_vote(unsigned long long* slots, unsigned long long* __restrict__ bins, uint32_t bin, uint32_t w) {
  if (!CACHED) { atomicAdd(&bins[bin], (unsigned long long)w); return; }
  unsigned long long* slot = slots + ((bin * 2654435761u) >> 20);
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(slot);
  for (;;) {
    const bool mine = (uint32_t)(old >> 32) == bin + 1u;
    const unsigned long long want = mine ? old + w : (((unsigned long long)(bin + 1u) << 32) | w);
    const unsigned long long seen = atomicCAS(slot, old, want);
    if (seen == old) {
      if (!mine && old) atomicAdd(&bins[(uint32_t)(old >> 32) - 1u], old & 0xffffffffull);
      return;
    }
    old = seen;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) hist2d_vec_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                          double norm2, double max_grad, unsigned long long smax, float inv_smax,
                                                          const uint16_t* __restrict__ vlut,
                                                          unsigned long long* __restrict__ bins) {
  // 8-bit volumes: at most 64 Ki bins in an L2-resident table, where the cache's CAS loop costs more than it saves
  // (measured 8.9 vs 8.0 ms at 1024^3); 16-bit: 1 Mi bins, 95 -> 82 ms at 2048^3 with the cache
  constexpr bool CACHED = sizeof(T) == 2;
  __shared__ unsigned long long slots[CACHED ? TVK_H2D_SLOTS : 1u];
  __shared__ unsigned long long queue[8][64];
  if (CACHED) {
    for (uint32_t i = threadIdx.x; i < TVK_H2D_SLOTS; i += blockDim.x) slots[i] = 0ull;
    __syncthreads();
  }
  unsigned long long* wq = queue[threadIdx.x >> 5];
  uint32_t qn = 0;                                      // queued voxels of this warp (same value in every lane)
  uint32_t pend_bin = 0xFFFFFFFFu, pend_cnt = 0;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nq = sx / 4u;
  const uint32_t lt = (1u << lane) - 1u;
  // tile width in quads: whole rows for short rows (better load efficiency), 32-quad chunks otherwise (TVK_H2D_XQ)
  const uint32_t xq = TVK_H2D_XQ ? TVK_H2D_XQ : nq;
  const uint32_t nxc = (nq + xq - 1u) / xq, nyb = (sy + 7u) / 8u, nzb = (sz + TVK_H2D_ZT - 1u) / TVK_H2D_ZT;
  const uint32_t ntiles = nxc * nyb * nzb;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t xc = tile % nxc, yb = (tile / nxc) % nyb, zb = tile / (nxc * nyb);
    const uint32_t y = yb * 8u + warp;
    const uint32_t z_end = min(sz, (zb + 1u) * TVK_H2D_ZT), q_end = min(nq, (xc + 1u) * xq);
    for (uint32_t z = zb * TVK_H2D_ZT; z < z_end; ++z)
    for (uint32_t q0 = xc * xq; q0 < q_end; q0 += 32u) {
      const uint32_t q = q0 + lane;
      const bool act = y < sy && q < q_end;
      const RowCtx<T> rc = row_ctx(raw, min(y, sy - 1u), z, sx, sy, sz);
      const unsigned long long row_base = ((unsigned long long)z * sy + y) * sx;
      unsigned long long s[4], key[4];
      uint32_t val[4], bin[4];
      quad_grad<T>(rc, q, nq, lane, act, s, key, val);
      uint32_t need = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t gi = smax == 0 ? 255u : 0u;             // (no gradient anywhere: NaN -> bin 255, see hist2d_kernel)
        if (s[i] != 0) {
          // 255*sqrt(S/Smax) in fp32 (relative error < 5e-7, i.e. < 1.3e-4 absolute); floor by a round-down magic add
          const float t = (float)s[i] * inv_smax;
          const float qf = 255.0f * (t * rsqrtf(t));
          const float mg = __fadd_rd(qf, 8388608.0f);
          const float fl = mg - 8388608.0f;
          // S < Smax: everything from 254.5 up lands in bin 254 (see hist2d_kernel); ties with the maximum and
          // estimates within 5e-4 of a bin boundary take the exact chain
          if (s[i] == smax) need |= 1u << i;
          else if (qf >= 254.5f) gi = 254u;
          else if (qf - fl > 5e-4f && fl + 1.0f - qf > 5e-4f) gi = __float_as_uint(mg) & 0x7FFFFFu;
          else need |= 1u << i;
        }
        bin[i] = (uint32_t)vlut[val[i]] * 256u + gi;
      }
      if (!act) need = 0;
      // ---- exact-chain queue
      if (__any_sync(0xffffffffu, need != 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned m = __ballot_sync(0xffffffffu, (need >> i) & 1u);
          if (!m) continue;
          if ((need >> i) & 1u) wq[qn + __popc(m & lt)] = row_base + 4u * q + i;
          qn += __popc(m);
          __syncwarp();
          if (qn >= 32u) {
            hist2d_exact_vote<T>(raw, wq[lane], sx, sy, sz, norm2, max_grad, vlut, bins);
            const unsigned long long mv = (lane + 32u < qn) ? wq[lane + 32u] : 0ull;
            __syncwarp();
            wq[lane] = mv;
            qn -= 32u;
            __syncwarp();
          }
        }
      }
      // ---- votes of the decided voxels: one vote per distinct bin of the warp (MATCH.ANY), repeated bins kept in a
      // register (the empty background sends every warp to bin 0), the rest through the shared-memory cache
      const bool uni = act && need == 0 && bin[0] == bin[1] && bin[1] == bin[2] && bin[2] == bin[3];
      const unsigned um = __ballot_sync(0xffffffffu, uni), am = __ballot_sync(0xffffffffu, act);
      if (um == am) {                                    // every active thread votes four times for one bin
        if (act) {
          const unsigned peers = __match_any_sync(am, bin[0]);
          if (lane == (unsigned)(__ffs(peers) - 1)) {
            const uint32_t n = 4u * __popc(peers);
            if (bin[0] == pend_bin) pend_cnt += n;
            else { if (pend_cnt) hist2d_cache_vote<CACHED>(slots, bins, pend_bin, pend_cnt); pend_bin = bin[0]; pend_cnt = n; }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool v = act && !((need >> i) & 1u);
          const unsigned vm = __ballot_sync(0xffffffffu, v);
          if (!v) continue;
          const unsigned peers = __match_any_sync(vm, bin[i]);
          if (lane == (unsigned)(__ffs(peers) - 1)) {
            const uint32_t n = __popc(peers);
            if (bin[i] == pend_bin) pend_cnt += n;
            else { if (pend_cnt) hist2d_cache_vote<CACHED>(slots, bins, pend_bin, pend_cnt); pend_bin = bin[i]; pend_cnt = n; }
          }
        }
      }
    }
  }
  if (pend_cnt) hist2d_cache_vote<CACHED>(slots, bins, pend_bin, pend_cnt);
  if (lane < qn) hist2d_exact_vote<T>(raw, wq[lane], sx, sy, sz, norm2, max_grad, vlut, bins);
  if (CACHED) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < TVK_H2D_SLOTS; i += blockDim.x) {
      const unsigned long long e = slots[i];
      if (e) atomicAdd(&bins[(uint32_t)(e >> 32) - 1u], e & 0xffffffffull);
    }
  }
}


// ---- 8 voxels per thread, squares in fp32 (rows a multiple of 8 voxels long, base 16-byte aligned) ---------
// What the 4-voxel kernels above spend their instructions on is the exact 64-bit S of EVERY voxel (three 32x32->64
// multiplies, 64-bit adds, a 64-bit int->float conversion) and a MATCH.ANY + 64-bit shared-memory CAS loop per vote.
// Neither is needed:
//  * a voxel value v becomes the float 2^23 + v with ONE byte permute (0x4B000000 | v); the difference of two such
//    floats is the exact central difference, and S' = dx*dx + dy*dy + dz*dz in fp32 is S to within 1.8e-7 relative
//    (exactly S below 2^24).  Pass 1 keeps the fp32 maximum as a FILTER: only voxels with S' >= best*(1 - 4e-7) --
//    the ones that can tie or beat the exact maximum -- evaluate the exact integer S and the fp64 chain.  Pass 2 bins
//    on S' and sends the voxels within 5e-4 of a bin boundary or within 1e-6 of the maximum down the exact chain, as
//    before (fp32 error of the estimate < 1.5e-4);
//  * votes go to a per-CTA table of (tag, count) pairs whose tags cannot change within a PHASE (= one tile): a slot is
//    claimed once (32-bit CAS on an empty tag) and then counted with plain 32-bit shared-memory adds; a vote whose slot
//    belongs to another bin goes to the global table directly.  Slots are flushed and cleared between tiles.
//  * a warp whose 5 x 32 x 16 bytes are all one value (empty background, saturated regions) votes once.
template <typename T> struct V8;
template <> struct V8<uint16_t> {
  typedef uint4 type;
  static __device__ __forceinline__ float m(const uint4& v, int i) {           // 2^23 + voxel i as a float
    const uint32_t w = (i >> 1) == 0 ? v.x : (i >> 1) == 1 ? v.y : (i >> 1) == 2 ? v.z : v.w;
    return __uint_as_float(__byte_perm(w, 0x4B00u, (i & 1) ? 0x5432 : 0x5410));
  }
  static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ uint4 load(const uint16_t* row, uint32_t o) { return __ldg(reinterpret_cast<const uint4*>(row) + o); }
  static __device__ __forceinline__ uint32_t first(const uint4& v) { return v.x & 0xffffu; }
  static __device__ __forceinline__ uint32_t last(const uint4& v) { return v.w >> 16; }
  static __device__ __forceinline__ bool all_equal(const uint4& v, uint32_t val) {
    const uint32_t rep = val | (val << 16);
    return ((v.x ^ rep) | (v.y ^ rep) | (v.z ^ rep) | (v.w ^ rep)) == 0u;
  }
  static __device__ __forceinline__ uint32_t shfl_last_up(const uint4& v) { return __shfl_up_sync(0xffffffffu, v.w, 1) >> 16; }
  static __device__ __forceinline__ uint32_t shfl_first_down(const uint4& v) { return __shfl_down_sync(0xffffffffu, v.x, 1) & 0xffffu; }
};
template <> struct V8<uint8_t> {
  typedef uint2 type;
  static __device__ __forceinline__ float m(const uint2& v, int i) {
    return __uint_as_float(__byte_perm((i >> 2) ? v.y : v.x, 0x4B000000u, 0x7440 | (i & 3)));
  }
  static __device__ __forceinline__ uint2 zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ uint2 load(const uint8_t* row, uint32_t o) { return __ldg(reinterpret_cast<const uint2*>(row) + o); }
  static __device__ __forceinline__ uint32_t first(const uint2& v) { return v.x & 0xffu; }
  static __device__ __forceinline__ uint32_t last(const uint2& v) { return v.y >> 24; }
  static __device__ __forceinline__ bool all_equal(const uint2& v, uint32_t val) {
    const uint32_t rep = val * 0x01010101u;
    return ((v.x ^ rep) | (v.y ^ rep)) == 0u;
  }
  static __device__ __forceinline__ uint32_t shfl_last_up(const uint2& v) { return __shfl_up_sync(0xffffffffu, v.y, 1) >> 24; }
  static __device__ __forceinline__ uint32_t shfl_first_down(const uint2& v) { return __shfl_down_sync(0xffffffffu, v.x, 1) & 0xffu; }
};

// The 8 voxels of octet o of row rc and everything around them, as loaded (all 32 lanes call; inactive lanes get zeros).
// `uniform` = every byte this thread looked at (own octet, x halo, four neighbouring octets) is the value `v0`.
template <typename T> struct Octet {
  typename V8<T>::type c, up, dn, fr, bk;
  uint32_t prev, next, v0;
  bool uniform;
};
template <typename T>
__device__ __forceinline__ Octet<T> octet_load(const RowCtx<T>& rc, uint32_t o, uint32_t o_first, uint32_t o_end, uint32_t no, uint32_t lane, bool act) {
  typedef V8<T> V;
  typedef typename V::type Q;
  const Q z = V::zero();
  Octet<T> r;
  r.c = act ? V::load(rc.line, o) : z;
  r.up = (act && rc.up) ? V::load(rc.up, o) : z; r.dn = (act && rc.down) ? V::load(rc.down, o) : z;
  r.fr = (act && rc.front) ? V::load(rc.front, o) : z; r.bk = (act && rc.back) ? V::load(rc.back, o) : z;
  r.prev = V::shfl_last_up(r.c); r.next = V::shfl_first_down(r.c);
  // the x halo of the first / last lane of the chunk comes from memory (zero outside the row)
  if (lane == 0 || o == o_first) r.prev = (act && o > 0) ? (uint32_t)__ldg(rc.line + 8u * o - 1u) : 0u;
  if (lane == 31 || o + 1 == o_end) r.next = (act && o + 1 < no) ? (uint32_t)__ldg(rc.line + 8u * o + 8u) : 0u;
  if (!act) r.prev = r.next = 0u;                       // (an idle lane next to the row's last octet: S' stays 0)
  r.v0 = V::first(r.c);
  r.uniform = V::all_equal(r.c, r.v0) && V::all_equal(r.up, r.v0) && V::all_equal(r.dn, r.v0) && V::all_equal(r.fr, r.v0) &&
              V::all_equal(r.bk, r.v0) && r.prev == r.v0 && r.next == r.v0;
  return r;
}
// S' per voxel and 2^23 + value per voxel of a loaded octet
template <typename T>
__device__ __forceinline__ void octet_s(const Octet<T>& q, float sf[8], float cm[8]) {
  typedef V8<T> V;
  float xm[10];
  xm[0] = __uint_as_float(0x4B000000u | q.prev);
  xm[9] = __uint_as_float(0x4B000000u | q.next);
#pragma unroll
  for (int i = 0; i < 8; ++i) { cm[i] = V::m(q.c, i); xm[i + 1] = cm[i]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float dx = __fsub_rn(xm[i], xm[i + 2]);                 // exact: both are 2^23 + integer
    const float dy = __fsub_rn(V::m(q.up, i), V::m(q.dn, i));
    const float dz = __fsub_rn(V::m(q.fr, i), V::m(q.bk, i));
    sf[i] = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) grad_max_v8_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                           double norm2, GradBest* __restrict__ partial,
                                                           unsigned long long* __restrict__ shared_best) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, no = sx / 8u;
  unsigned long long best_s = 0, best_bits = 0, last_key = ~0ull;
  float thr = 0.0f;                                     // a voxel below it cannot tie or beat best_s
  const uint32_t nyb = (sy + 7u) / 8u, nzb = (sz + 15u) / 16u, ntiles = nyb * nzb;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  for (uint32_t z = (tile / nyb) * 16u; z < min(sz, (tile / nyb) * 16u + 16u); ++z) {
    const uint32_t y = (tile % nyb) * 8u + warp;
    const bool row_valid = y < sy;
    const RowCtx<T> rc = row_ctx(raw, min(y, sy - 1u), z, sx, sy, sz);
    const unsigned long long gb = *reinterpret_cast<volatile unsigned long long*>(shared_best);
    if (gb > best_s) { best_s = gb; best_bits = 0; thr = __fmul_rd(__ull2float_rd(best_s), 1.0f - 4e-7f); }
    for (uint32_t o0 = 0; o0 < no; o0 += 32u) {
      const uint32_t o = o0 + lane;
      const bool act = row_valid && o < no;
      const Octet<T> oc = octet_load<T>(rc, o, o0, no, no, lane, act);
      if (__all_sync(0xffffffffu, oc.uniform)) continue;            // one value everywhere the warp looked: no gradient
      float sf[8], cm[8];
      octet_s<T>(oc, sf, cm);
      const float m = fmaxf(fmaxf(fmaxf(sf[0], sf[1]), fmaxf(sf[2], sf[3])), fmaxf(fmaxf(sf[4], sf[5]), fmaxf(sf[6], sf[7])));
      if (m == 0.0f || m < thr) continue;
      uint32_t cand = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) cand |= (sf[i] != 0.0f && sf[i] >= thr) ? (1u << i) : 0u;
      while (cand) {
        const uint32_t i = __ffs(cand) - 1u;
        cand &= cand - 1u;
        const uint32_t x = 8u * o + i;
        unsigned long long key;
        const unsigned long long s = grad_s_row(rc, x, sx, &key);
        if (s == 0 || s < best_s || key == last_key) continue;
        last_key = key;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(grad_mag(raw, x, y, z, sx, sy, sz, norm2));
        if (s > best_s) {
          best_s = s; best_bits = bits; atomicMax(shared_best, best_s);
          thr = __fmul_rd(__ull2float_rd(best_s), 1.0f - 4e-7f);
        } else if (bits > best_bits) best_bits = bits;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long os = __shfl_xor_sync(0xffffffffu, best_s, o), ob = __shfl_xor_sync(0xffffffffu, best_bits, o);
    if (os > best_s || (os == best_s && ob > best_bits)) { best_s = os; best_bits = ob; }
  }
  __shared__ GradBest sh[8];
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5].s = best_s; sh[threadIdx.x >> 5].len_bits = best_bits; }
  __syncthreads();
  if (threadIdx.x == 0) {
    GradBest g = sh[0];
    for (int w = 1; w < 8; ++w) if (sh[w].s > g.s || (sh[w].s == g.s && sh[w].len_bits > g.len_bits)) g = sh[w];
    partial[blockIdx.x] = g;
  }
}

#ifndef TVK_H2D8_LOG_SLOTS
#define TVK_H2D8_LOG_SLOTS 13u   // 8192 slots of 4 bytes
#endif
#ifndef TVK_H2D8_XO
#define TVK_H2D8_XO 32u          // tile width in octets (one warp trip = 256 voxels)
#endif
#ifndef TVK_H2D8_ZT
#define TVK_H2D8_ZT 16u
#endif
#define TVK_H2D8_SLOTS (1u << TVK_H2D8_LOG_SLOTS)
#define TVK_H2D8_BIN_BITS 20u                           // bins_x * 256 <= 2^20 (4096 value columns), else the 4-voxel kernels
static_assert(TVK_H2D8_XO * 8u * 8u * TVK_H2D8_ZT < 65536u, "a tile's votes must fit the 16-bit count of a slot");
// A slot is ONE 32-bit word: tag << 16 | count.  bin -> m = bin * K mod 2^20 is a bijection (K odd); the slot index is the
// top bits of m, the tag the remaining low bits + 1 (never 0), so (slot, tag) names the bin and the flush can invert it.
constexpr uint32_t kH2d8K = 0x779B1u;
constexpr uint32_t h2d8_inverse(uint32_t k) { uint32_t x = k; for (int i = 0; i < 6; ++i) x *= 2u - k * x; return x & ((1u << TVK_H2D8_BIN_BITS) - 1u); }
constexpr uint32_t kH2d8Kinv = h2d8_inverse(kH2d8K);
static_assert(((kH2d8K * kH2d8Kinv) & ((1u << TVK_H2D8_BIN_BITS) - 1u)) == 1u, "modular inverse");
constexpr uint32_t kH2d8TagBits = TVK_H2D8_BIN_BITS - TVK_H2D8_LOG_SLOTS;

__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// One vote.  The tag of a slot cannot change before the end of the phase (tile), so "tag matches" followed by a plain
// shared-memory add is safe; an empty slot is claimed with one CAS; a slot owned by another bin sends the vote to the
// global table.  Tags carry a valid bit (0x80), so a word with count 0 still names its bin: at the end of a phase the
// counts are flushed and the tags of the slots that were voted for STAY (the next tile of the CTA is its neighbour and
// mostly votes for the same bins -- re-claiming them cost a divergent CAS path in two warps out of three), slots that
// saw no vote are freed.
// Shared-state-space instructions on 32-bit addresses (no generic-address arithmetic).
__device__ __forceinline__ uint32_t lds_volatile(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void reds_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(a) : "memory"); }
__device__ __forceinline__ uint32_t atoms_cas(uint32_t a, uint32_t cmp, uint32_t val) {
  uint32_t o; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(a), "r"(cmp), "r"(val) : "memory"); return o;
}
__device__ __noinline__ void h2d8_vote_slow(uint32_t addr, uint32_t tagw, uint32_t w, unsigned long long* __restrict__ bins, uint32_t bin) {
  if (w == 0u) {
    w = atoms_cas(addr, 0u, tagw | 1u);
    if (w == 0u) return;
    if ((w ^ tagw) < 0x10000u) { reds_inc(addr); return; }
  }
  atomicAdd(&bins[bin], 1ull);
}
__device__ __forceinline__ void h2d8_vote(uint32_t slots_saddr, unsigned long long* __restrict__ bins, uint32_t bin) {
  const uint32_t m = bin * kH2d8K;                                       // low 20 bits: the bijective mix
  const uint32_t addr = slots_saddr + ((m >> (kH2d8TagBits - 2u)) & ((TVK_H2D8_SLOTS - 1u) << 2));
  const uint32_t tagw = ((m << 16) & (((1u << kH2d8TagBits) - 1u) << 16)) | (1u << (16 + kH2d8TagBits));
  const uint32_t w = lds_volatile(addr);
  if ((w ^ tagw) < 0x10000u) reds_inc(addr);                             // same tag
  else h2d8_vote_slow(addr, tagw, w, bins, bin);
}

template <typename T>
__global__ void __launch_bounds__(256) hist2d_v8_kernel(const T* __restrict__ raw, uint32_t sx, uint32_t sy, uint32_t sz,
                                                         double norm2, double max_grad, unsigned long long smax, float k_smax, float tie_thr,
                                                         const uint16_t* __restrict__ vlut,
                                                         unsigned long long* __restrict__ bins) {
  extern __shared__ unsigned long long h2d8_smem[];       // 8 x 64 queue entries, then the slots
  uint32_t* slots = reinterpret_cast<uint32_t*>(h2d8_smem + 8 * 64);
  for (uint32_t i = threadIdx.x; i < TVK_H2D8_SLOTS; i += blockDim.x) slots[i] = 0u;
  __syncthreads();
  const uint32_t slots_saddr = (uint32_t)__cvta_generic_to_shared(slots);
  unsigned long long* wq = h2d8_smem + 64 * (threadIdx.x >> 5);
  uint32_t qn = 0;                                      // queued voxels of this warp (same value in every lane)
  uint32_t pend_bin = 0xFFFFFFFFu;
  unsigned long long pend_cnt = 0;                      // lane 0: votes of uniform warps, kept across tiles
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, no = sx / 8u;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t nxc = (no + TVK_H2D8_XO - 1u) / TVK_H2D8_XO, nyb = (sy + 7u) / 8u, nzb = (sz + TVK_H2D8_ZT - 1u) / TVK_H2D8_ZT;
  const uint32_t ntiles = nxc * nyb * nzb;
  const uint32_t gi_zero = smax == 0 ? 255u : 0u;       // (no gradient anywhere: NaN -> bin 255, see hist2d_kernel)
  // a CTA takes a run of consecutive tiles (neighbours along x, then y): the bins of one tile are mostly those of the next
  const uint32_t per_cta = (ntiles + gridDim.x - 1u) / gridDim.x;
  const uint32_t tile_end = min(ntiles, (blockIdx.x + 1u) * per_cta);
  for (uint32_t tile = blockIdx.x * per_cta; tile < tile_end; ++tile) {
    const uint32_t xc = tile % nxc, yb = (tile / nxc) % nyb, zb = tile / (nxc * nyb);
    const uint32_t y = yb * 8u + warp;
    const uint32_t z_end = min(sz, (zb + 1u) * TVK_H2D8_ZT), o_end = min(no, (xc + 1u) * TVK_H2D8_XO);
    for (uint32_t z = zb * TVK_H2D8_ZT; z < z_end; ++z)
    for (uint32_t o0 = xc * TVK_H2D8_XO; o0 < o_end; o0 += 32u) {
      const uint32_t o = o0 + lane;
      const bool act = y < sy && o < o_end;
      const RowCtx<T> rc = row_ctx(raw, min(y, sy - 1u), z, sx, sy, sz);
      const Octet<T> oc = octet_load<T>(rc, o, o0, o_end, no, lane, act);
      const unsigned am = __ballot_sync(0xffffffffu, act);
      if (am == 0u) continue;
      // ---- one value everywhere the warp looked: one vote for the lot (before any of the per-voxel arithmetic)
      const uint32_t lead = __ffs(am) - 1;
      const uint32_t vl = __shfl_sync(0xffffffffu, oc.v0, lead);
      if (__all_sync(0xffffffffu, !act || (oc.uniform && oc.v0 == vl))) {
        if (lane == 0) {
          const uint32_t bin = (uint32_t)vlut[vl] * 256u + gi_zero;
          const uint32_t n = 8u * __popc(am);
          if (bin == pend_bin) pend_cnt += n;
          else { if (pend_cnt) atomicAdd(&bins[pend_bin], pend_cnt); pend_bin = bin; pend_cnt = n; }
        }
        continue;
      }
      float sf[8], cm[8];
      octet_s<T>(oc, sf, cm);
      // ---- value columns: eight table reads in flight while the gradient bins are worked out
      uint32_t vb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) vb[i] = (uint32_t)__ldg(vlut + (__float_as_uint(cm[i]) & 0xFFFFu)) << 8;
      uint32_t need = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // 255*sqrt(S/Smax) = sqrt(t), t = S' * (65025/Smax); floor by a round-down magic add.  Branch-free: S' == 0
        // makes t * rsqrt(t) a NaN, which every comparison below rejects before the zero case overrides the result.
        const float t = sf[i] * k_smax;
        const float qf = t * rsqrt_ftz(t);
        const float mg = __fadd_rd(qf, 8388608.0f);
        const float fr = qf - (mg - 8388608.0f);
        const bool top = qf >= 254.5f;                                       // S < Smax: bin 254 (see hist2d_kernel)
        const bool ok = sf[i] < tie_thr && (top || fabsf(fr - 0.5f) < 0.4995f);   // else: boundary or may tie with the maximum
        uint32_t gi = top ? 254u : (__float_as_uint(mg) & 0x7FFFFFu);
        const bool zero = sf[i] == 0.0f;
        gi = zero ? gi_zero : gi;
        vb[i] |= gi;
        need |= (zero || ok) ? 0u : (1u << i);
      }
      if (!act) need = 0xFFu;                            // idle lanes: no votes (the queue below takes active lanes only)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (!((need >> i) & 1u)) h2d8_vote(slots_saddr, bins, vb[i]);
      if (!act) need = 0;
      // ---- exact-chain queue
      if (__any_sync(0xffffffffu, need != 0)) {
        const unsigned long long row_base = ((unsigned long long)z * sy + y) * sx + 8ull * o;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const unsigned m = __ballot_sync(0xffffffffu, (need >> i) & 1u);
          if (!m) continue;
          if ((need >> i) & 1u) wq[qn + __popc(m & lt)] = row_base + i;
          qn += __popc(m);
          __syncwarp();
          if (qn >= 32u) {
            hist2d_exact_vote<T>(raw, wq[lane], sx, sy, sz, norm2, max_grad, vlut, bins);
            const unsigned long long mv = (lane + 32u < qn) ? wq[lane + 32u] : 0ull;
            __syncwarp();
            wq[lane] = mv;
            qn -= 32u;
            __syncwarp();
          }
        }
      }
    }
    // ---- end of the phase: counts go to the global table, slots are free again
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < TVK_H2D8_SLOTS; i += blockDim.x) {
      const uint32_t w = slots[i];
      if (w & 0xFFFFu) {
        const uint32_t m = (i << kH2d8TagBits) | ((w >> 16) & ((1u << kH2d8TagBits) - 1u));
        atomicAdd(&bins[(m * kH2d8Kinv) & ((1u << TVK_H2D8_BIN_BITS) - 1u)], (unsigned long long)(w & 0xFFFFu));
        slots[i] = w & 0xFFFF0000u;                     // the bin keeps its slot
      } else if (w) slots[i] = 0u;                      // no vote in this phase: the slot is free again
    }
    __syncthreads();
  }
  if (lane == 0 && pend_cnt) atomicAdd(&bins[pend_bin], pend_cnt);
  if (lane < qn) hist2d_exact_vote<T>(raw, wq[lane], sx, sy, sz, norm2, max_grad, vlut, bins);
}

static void release(tvk_ctx* c) {
  BuildState& b = c->build;
  if (b.own_store) cudaFree(b.d_store);
  cudaFree(b.d_mn);
  cudaFree(b.d_offsets); cudaFree(b.d_sizes); cudaFree(b.d_minmax); cudaFree(b.d_hist1d);
  for (uint32_t l = 0; l < TVK_MAX_LOD; ++l) {
    if (l > 0 || b.own_lod0) cudaFree(b.d_lod[l]);
    b.d_lod[l] = nullptr;
  }
  b = BuildState();
}

extern "C" int tvk_build_release(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  cudaSetDevice(c->device);
  release(c);
  return TVK_OK;
}

// ------------------------------------------------------------------ host side of a build
// A build is begin -> LOD-0 rows -> finish.  tvk_build_octree feeds the whole raw volume as one "slab";
// tvk_build_stream_* feeds it in z-slabs of whole brick rows (raw volumes that do not fit next to their brick store).
static BuildConst make_const(const BuildState& b) {
  BuildConst bc{};
  memcpy(bc.lod_size, b.info.lod_size, sizeof(bc.lod_size));
  memcpy(bc.lod_bricks, b.info.lod_bricks, sizeof(bc.lod_bricks));
  bc.brick = b.desc.brick_size; bc.overlap = b.desc.overlap; bc.core = b.desc.brick_size - 2 * b.desc.overlap;
  return bc;
}

template <typename T>
static int build_begin(tvk_ctx* c, BuildState& b) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  // all allocations first, so that the event-timed regions below hold kernels only
  for (uint32_t l = 1; l < info.lod_count; ++l) {
    const uint64_t n = (uint64_t)info.lod_size[l][0] * info.lod_size[l][1] * info.lod_size[l][2];
    TVK_CHECK_CUDA(c, cudaMalloc(&b.d_lod[l], n * sizeof(T)));
  }
  TVK_CHECK_CUDA(c, cudaMalloc(&b.d_mn, sizeof(uint32_t) * (2 * info.total_bricks + 1)));
  const uint32_t range = 1u << (8 * sizeof(T));
  TVK_CHECK_CUDA(c, cudaMalloc(&b.d_hist1d, sizeof(unsigned long long) * range));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(b.d_mn, 0xFF, sizeof(uint32_t) * info.total_bricks, st));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(b.d_mn + info.total_bricks, 0, sizeof(uint32_t) * (info.total_bricks + 1), st));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(b.d_hist1d, 0, sizeof(unsigned long long) * range, st));
  b.hist_fused = false; b.hist_full = false; b.lod0_fused = false; b.hist_pending = false; b.hist_upper_pending = false;
  b.kernels_ms = 0.f;
  return TVK_OK;
}

// launches one brick-cut pass: bricks bz in [bz0, bz1) of LOD `l`, read from `src` = rows [z0, z0 + rows) of the LOD
template <typename T>
static int cut_rows(tvk_ctx* c, BuildState& b, const BuildConst& bc, int sms, uint32_t l, const T* src, uint32_t z0, uint32_t rows,
                    uint32_t bz0, uint32_t bz1) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  uint32_t* d_mn = b.d_mn;
  uint32_t* d_mx = b.d_mn + info.total_bricks;
  // one launch for the full-width bricks, one for the last column if it is narrower
  const uint32_t nbx = info.lod_bricks[l][0];
  const uint32_t last_tx = brick_dim(bc, l, 0, nbx - 1);
  const uint32_t groups[2][3] = {{0u, last_tx == bc.brick ? nbx : nbx - 1, bc.brick}, {nbx - 1, last_tx == bc.brick ? 0u : 1u, last_tx}};
  const SlabView sv{z0, rows, bz0};
  for (int g = 0; g < 2; ++g) {
    const uint32_t x_first = groups[g][0], x_count = groups[g][1], tx = groups[g][2];
    if (!x_count) continue;
    const uint64_t nb = (uint64_t)x_count * info.lod_bricks[l][1] * (bz1 - bz0);
    // a few z-chunks per brick when there are few bricks, to fill the machine
    const unsigned chunks = nb >= (uint64_t)sms * 8 ? 2 : (nb >= (uint64_t)sms ? 8 : 32);
    const dim3 grid((unsigned)nb, chunks);
    const uint32_t row_bytes = tx * (uint32_t)sizeof(T);
    // word-granular path: needs a 4-byte aligned volume, and core >= overlap for the closed-form ghosts
    const bool vec = (reinterpret_cast<uintptr_t>(src) & 3) == 0 && bc.core >= bc.overlap && row_bytes % 4 == 0;
    uint8_t* store = (uint8_t*)b.d_store;
    if (b.desc.clamp_to_edge)
      cut_bricks_clamp_kernel<T><<<grid, 256, 0, st>>>(src, store, b.d_offsets, info.lod_offset[l], l, bc, d_mn, d_mx, x_first, x_count, sv);
    else if (vec && row_bytes % 16 == 0)
      cut_bricks_vec_kernel<T, 16><<<grid, 256, 0, st>>>(src, store, b.d_offsets, info.lod_offset[l], l, bc, d_mn, d_mx, x_first, x_count, sv);
    else if (vec && row_bytes % 8 == 0)
      cut_bricks_vec_kernel<T, 8><<<grid, 256, 0, st>>>(src, store, b.d_offsets, info.lod_offset[l], l, bc, d_mn, d_mx, x_first, x_count, sv);
    else if (vec)
      cut_bricks_vec_kernel<T, 4><<<grid, 256, 0, st>>>(src, store, b.d_offsets, info.lod_offset[l], l, bc, d_mn, d_mx, x_first, x_count, sv);
    else
      cut_bricks_kernel<T><<<grid, 256, 0, st>>>(src, store, b.d_offsets, info.lod_offset[l], l, bc, d_mn, d_mx, x_first, x_count, sv);
    c->launches++;
  }
  TVK_CHECK_CUDA(c, cudaGetLastError());
  return TVK_OK;
}

// launches the reduction of rows [s_z0, s_z0 + s_rows) of LOD l-1 (at `src`) into LOD l (rows from s_z0 / 2 on); for
// l == 1 it also counts the 1-D histogram of those rows when the fused kernel applies
template <typename T>
static int downsample_rows(tvk_ctx* c, BuildState& b, int sms, uint32_t l, const T* src, uint32_t s_z0, uint32_t s_rows, bool* counted) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  constexpr uint32_t E = 16 / sizeof(T);
  const uint32_t s[3] = {info.lod_size[l - 1][0], info.lod_size[l - 1][1], s_rows};
  const uint32_t d[3] = {info.lod_size[l][0], info.lod_size[l][1], info.lod_size[l - 1][2] > 1 ? (s_rows + 1) / 2 : s_rows};
  const uint32_t d_z0 = info.lod_size[l - 1][2] > 1 ? s_z0 / 2 : s_z0;
  T* dst = (T*)b.d_lod[l] + (uint64_t)d_z0 * d[0] * d[1];
  const bool vec = (s[0] % E) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const uint64_t n = vec ? (uint64_t)(s[0] / E) * d[1] * d[2] : (uint64_t)d[0] * d[1] * d[2];
  const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 32);
  // LOD 0 -> 1 also counts the 1-D histogram (all 65536 bins for 16-bit data: packed shared-memory counters)
  const bool fuse_hist = vec && l == 1 && ((uint64_t)s[0] * s[1] * s[2]) >= (1u << 18) && info.lod_size[0][2] > 1;
  if (counted) *counted = fuse_hist;
  if (fuse_hist) {
    const size_t smem = sizeof(unsigned int) * (sizeof(T) == 1 ? 32u * 256u : HIST_WINDOW);
    static bool attr_set = false;
    if (!attr_set) {
      TVK_CHECK_CUDA(c, cudaFuncSetAttribute(downsample_hist_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      TVK_CHECK_CUDA(c, cudaFuncSetAttribute(downsample_hist_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    if (b.desc.use_median) downsample_hist_kernel<T, true><<<sms, 1024, smem, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2], b.d_hist1d);
    else downsample_hist_kernel<T, false><<<sms, 1024, smem, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2], b.d_hist1d);
  } else if (vec) {
    if (b.desc.use_median) downsample_vec_kernel<T, true><<<blocks, 256, 0, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2]);
    else downsample_vec_kernel<T, false><<<blocks, 256, 0, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2]);
  } else {
    if (b.desc.use_median) downsample_kernel<T, true><<<blocks, 256, 0, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2]);
    else downsample_kernel<T, false><<<blocks, 256, 0, st>>>(src, dst, s[0], s[1], s[2], d[0], d[1], d[2]);
  }
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  return TVK_OK;
}

// LOD-0 work for brick rows [bz0, bz1): `slab` holds raw rows [z0, z0 + rows), which must include the rows those bricks
// touch.  whole == true: the slab is the entire volume (one-shot build) and the histogram may be finished later from it.
template <typename T>
static int build_lod0_rows(tvk_ctx* c, BuildState& b, const BuildConst& bc, int sms, const T* slab, uint32_t z0, uint32_t rows,
                           uint32_t bz0, uint32_t bz1, bool whole) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  const uint32_t* s0 = info.lod_size[0];
  const uint32_t cz0 = bz0 * bc.core, cz1 = std::min<uint32_t>(bz1 * bc.core, s0[2]);   // inner rows of these bricks
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev0, st));
  bool counted = false;
  if (info.lod_count > 1 && cz1 > cz0) {
    int rc = downsample_rows<T>(c, b, sms, 1, slab + (uint64_t)(cz0 - z0) * s0[0] * s0[1], cz0, cz1 - cz0, &counted);
    if (rc) return rc;
  }
  int rc = cut_rows<T>(c, b, bc, sms, 0, slab, z0, rows, bz0, bz1);
  if (rc) return rc;
  if (counted) { b.hist_fused = true; }
  else if (!whole && cz1 > cz0) {
    // streamed slab the fused kernel does not take (unaligned rows / tiny slab): plain global-atomic histogram of its inner rows
    const uint64_t n = (uint64_t)(cz1 - cz0) * s0[0] * s0[1];
    hist1d_kernel<T><<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>(slab + (uint64_t)(cz0 - z0) * s0[0] * s0[1], n, b.d_hist1d);
    c->launches++;
    TVK_CHECK_CUDA(c, cudaGetLastError());
  } else if (whole) {
    b.hist_pending = true;                       // finished in build_finish from the whole volume
  }
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev1, st));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->bev0, c->bev1) == cudaSuccess) b.kernels_ms += ms;
  return TVK_OK;
}

template <typename T>
static int build_finish(tvk_ctx* c, BuildState& b, const BuildConst& bc, int sms) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  uint32_t* d_mn = b.d_mn;
  uint32_t* d_mx = d_mn + info.total_bricks;
  uint32_t* d_gmax = d_mx + info.total_bricks;
  const uint32_t range = 1u << (8 * sizeof(T));
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev0, st));
  // the rest of the flat pyramid, then its bricks
  for (uint32_t l = 2; l < info.lod_count; ++l) {
    int rc = downsample_rows<T>(c, b, sms, l, (const T*)b.d_lod[l - 1], 0, info.lod_size[l - 1][2], nullptr);
    if (rc) return rc;
  }
  for (uint32_t l = 1; l < info.lod_count; ++l) {
    int rc = cut_rows<T>(c, b, bc, sms, l, (const T*)b.d_lod[l], 0, info.lod_size[l][2], 0, info.lod_bricks[l][2]);
    if (rc) return rc;
  }
  minmax_to_double_kernel<<<(unsigned)((info.total_bricks + 255) / 256), 256, 0, st>>>(d_mn, d_mx, b.d_minmax, info.total_bricks, d_gmax);
  c->launches++;
  if (b.hist_upper_pending) {                    // fused LOD-0 pass, 16-bit: bins [HIST_WINDOW, 65536) from a second read
    const uint64_t n0 = (uint64_t)info.lod_size[0][0] * info.lod_size[0][1] * info.lod_size[0][2];
    for (uint32_t lo = HIST_WINDOW; lo < range; lo += HIST_WINDOW) {
      hist16_window_kernel<<<sms, 1024, HIST_WINDOW * sizeof(uint32_t), st>>>((const uint16_t*)b.d_lod[0], n0, lo, d_gmax, b.d_hist1d);
      c->launches++;
    }
    b.hist_upper_pending = false;
  }
  // 1-D histogram of a one-shot build that the fused LOD 0 -> 1 pass did not count
  if (b.hist_pending && !b.hist_fused) {
    const uint64_t n0 = (uint64_t)info.lod_size[0][0] * info.lod_size[0][1] * info.lod_size[0][2];
    if (sizeof(T) == 2 && n0 >= (1u << 18)) {
      static bool attr_set = false;
      if (!attr_set) {
        TVK_CHECK_CUDA(c, cudaFuncSetAttribute(hist16_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HIST_WINDOW * sizeof(uint32_t))));
        attr_set = true;
      }
      for (uint32_t lo = 0u; lo < range; lo += HIST_WINDOW) {
        hist16_window_kernel<<<sms, 1024, HIST_WINDOW * sizeof(uint32_t), st>>>((const uint16_t*)b.d_lod[0], n0, lo, d_gmax, b.d_hist1d);
        c->launches++;
      }
    } else {
      hist1d_kernel<T><<<(unsigned)std::min<uint64_t>((n0 + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>((const T*)b.d_lod[0], n0, b.d_hist1d);
      c->launches++;
    }
  }
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev1, st));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->bev0, c->bev1) == cudaSuccess) b.kernels_ms += ms;
  cudaFree(b.d_mn); b.d_mn = nullptr;
  return TVK_OK;
}

// plan + ToC + the allocations every build needs; `store` != nullptr: caller-owned brick store (device memory, or pinned
// host memory mapped into the device address space -- bricks are then written over PCIe)
static int build_setup(tvk_ctx* c, const tvk_build_desc* d, void* store) {
  tvk_build_info info;
  if (tvk_build_plan(d, &info) != TVK_OK) TVK_FAIL(c, TVK_ERR_INVALID, "bad build description");
  if ((d->brick_size - 2 * d->overlap) % 2) TVK_FAIL(c, TVK_ERR_UNSUPPORTED, "odd inner brick size");
  release(c);
  BuildState& b = c->build;
  b.desc = *d;
  b.info = info;
  const BuildConst bc = make_const(b);
  // ToC
  b.offsets.resize(info.total_bricks);
  b.sizes.resize(3 * info.total_bricks);
  uint64_t off = 0, i = 0;
  for (uint32_t l = 0; l < info.lod_count; ++l)
    for (uint32_t z = 0; z < info.lod_bricks[l][2]; ++z)
      for (uint32_t y = 0; y < info.lod_bricks[l][1]; ++y)
        for (uint32_t x = 0; x < info.lod_bricks[l][0]; ++x, ++i) {
          const uint32_t s[3] = {brick_dim(bc, l, 0, x), brick_dim(bc, l, 1, y), brick_dim(bc, l, 2, z)};
          b.offsets[i] = off;
          b.sizes[3 * i] = s[0]; b.sizes[3 * i + 1] = s[1]; b.sizes[3 * i + 2] = s[2];
          off += ((uint64_t)s[0] * s[1] * s[2] * d->bytes_per_voxel + kBrickAlign - 1) / kBrickAlign * kBrickAlign;
        }
  b.info.total_bytes = off;
  if (store) { b.d_store = store; b.own_store = false; }
  else { TVK_CHECK_CUDA(c, cudaMalloc(&b.d_store, off)); b.own_store = true; }
  TVK_CHECK_CUDA(c, cudaMalloc(&b.d_offsets, sizeof(uint64_t) * info.total_bricks));
  TVK_CHECK_CUDA(c, cudaMalloc(&b.d_sizes, sizeof(uint32_t) * 3 * info.total_bricks));
  TVK_CHECK_CUDA(c, cudaMalloc(&b.d_minmax, sizeof(double) * 2 * info.total_bricks));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(b.d_offsets, b.offsets.data(), sizeof(uint64_t) * info.total_bricks, cudaMemcpyHostToDevice, c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(b.d_sizes, b.sizes.data(), sizeof(uint32_t) * 3 * info.total_bricks, cudaMemcpyHostToDevice, c->stream));
  return TVK_OK;
}

// Opt-in (TVK_BUILD_FUSED=1) fused LOD-0 pass of a one-shot build: bricks + min/max + 1-D histogram (bins below
// HIST_WINDOW for 16-bit data) + flat LOD 1 from ONE read of the raw volume.  Measured slower than the separate kernels
// (DESIGN.md); kept because it is the lower-traffic variant.  *done = false: conditions not met, nothing launched.
template <typename T>
static int lod0_fused_pass(tvk_ctx* c, BuildState& b, const BuildConst& bc, int sms, bool* done) {
  const tvk_build_info& info = b.info;
  cudaStream_t st = c->stream;
  *done = false;
  const uint32_t* s0 = info.lod_size[0];
  const uint32_t nbx = info.lod_bricks[0][0];
  const uint32_t last_tx = brick_dim(bc, 0, 0, nbx - 1);
  const uint32_t groups[2][3] = {{0u, last_tx == bc.brick ? nbx : nbx - 1, bc.brick}, {nbx - 1, last_tx == bc.brick ? 0u : 1u, last_tx}};
  bool can = info.lod_count >= 2 && s0[0] > 1 && s0[1] > 1 && s0[2] > 1 && (bc.overlap % 2) == 0 && bc.core >= bc.overlap &&
             (reinterpret_cast<uintptr_t>(b.d_lod[0]) & 3) == 0 && (uint64_t)s0[0] * s0[1] * s0[2] >= (1u << 18) &&
             getenv("TVK_BUILD_FUSED") != nullptr && !b.desc.clamp_to_edge;
  for (int g = 0; g < 2; ++g) if (groups[g][1] && (groups[g][2] * sizeof(T)) % 4 != 0) can = false;
  if (!can) return TVK_OK;
  uint32_t* d_mn = b.d_mn;
  uint32_t* d_mx = b.d_mn + info.total_bricks;
  const size_t smem = sizeof(unsigned int) * (sizeof(T) == 1 ? 32u * 256u : HIST_WINDOW);
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev0, st));
  for (int g = 0; g < 2; ++g) {
    const uint32_t x_first = groups[g][0], x_count = groups[g][1], row_bytes = groups[g][2] * (uint32_t)sizeof(T);
    if (!x_count) continue;
    const uint64_t nb = (uint64_t)x_count * info.lod_bricks[0][1] * info.lod_bricks[0][2];
    const unsigned grid = (unsigned)std::min<uint64_t>(nb, (uint64_t)sms);
    const T* src = (const T*)b.d_lod[0];
    uint8_t* store = (uint8_t*)b.d_store;
    T* l1 = (T*)b.d_lod[1];
#define TVK_LAUNCH_FUSED(CBV)                                                                                                  \
    do {                                                                                                                   \
      if (b.desc.use_median) {                                                                                             \
        TVK_CHECK_CUDA(c, cudaFuncSetAttribute(lod0_fused_kernel<T, CBV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lod0_fused_kernel<T, CBV, true><<<grid, 1024, smem, st>>>(src, store, b.d_offsets, bc, d_mn, d_mx, x_first, x_count, l1, b.d_hist1d); \
      } else {                                                                                                             \
        TVK_CHECK_CUDA(c, cudaFuncSetAttribute(lod0_fused_kernel<T, CBV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lod0_fused_kernel<T, CBV, false><<<grid, 1024, smem, st>>>(src, store, b.d_offsets, bc, d_mn, d_mx, x_first, x_count, l1, b.d_hist1d); \
      }                                                                                                                    \
    } while (0)
    if (row_bytes % 16 == 0) TVK_LAUNCH_FUSED(16);
    else if (row_bytes % 8 == 0) TVK_LAUNCH_FUSED(8);
    else TVK_LAUNCH_FUSED(4);
#undef TVK_LAUNCH_FUSED
    c->launches++;
  }
  TVK_CHECK_CUDA(c, cudaGetLastError());
  if (sizeof(T) == 2) {                          // the fused table holds the bins below HIST_WINDOW: count the rest in a second pass
    static bool attr_set = false;
    if (!attr_set) {
      TVK_CHECK_CUDA(c, cudaFuncSetAttribute(hist16_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HIST_WINDOW * sizeof(uint32_t))));
      attr_set = true;
    }
    b.hist_upper_pending = true;                 // launched in build_finish, once the global maximum is known
  }
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev1, st));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(st));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->bev0, c->bev1) == cudaSuccess) b.kernels_ms += ms;
  b.hist_fused = true; b.lod0_fused = true;
  *done = true;
  return TVK_OK;
}

template <typename T>
static int run_build(tvk_ctx* c, BuildState& b, const BuildConst& bc, int sms) {
  int rc = build_begin<T>(c, b);
  if (rc) return rc;
  const tvk_build_info& info = b.info;
  bool fused = false;
  rc = lod0_fused_pass<T>(c, b, bc, sms, &fused);
  if (rc) return rc;
  if (!fused) {
    rc = build_lod0_rows<T>(c, b, bc, sms, (const T*)b.d_lod[0], 0, info.lod_size[0][2], 0, info.lod_bricks[0][2], true);
    if (rc) return rc;
  }
  return build_finish<T>(c, b, bc, sms);
}

extern "C" int tvk_build_octree(tvk_ctx* c, const tvk_build_desc* d, const void* raw, int raw_on_device, tvk_build_info* info_out) {
  if (!c || !d || !raw) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int rc = build_setup(c, d, nullptr);
  if (rc) { release(c); return rc; }
  BuildState& b = c->build;
  const BuildConst bc = make_const(b);
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const uint64_t n0 = (uint64_t)d->volume_size[0] * d->volume_size[1] * d->volume_size[2];
  if (raw_on_device) { b.d_lod[0] = const_cast<void*>(raw); b.own_lod0 = false; }
  else {
    TVK_CHECK_CUDA(c, cudaMalloc(&b.d_lod[0], n0 * d->bytes_per_voxel));
    b.own_lod0 = true;
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(b.d_lod[0], raw, n0 * d->bytes_per_voxel, cudaMemcpyHostToDevice, c->stream));
  }
  rc = (d->bytes_per_voxel == 1) ? run_build<uint8_t>(c, b, bc, sms) : run_build<uint16_t>(c, b, bc, sms);
  if (rc) { release(c); return rc; }
  b.valid = true;
  if (info_out) *info_out = b.info;
  return TVK_OK;
}

// ---- streamed build: LOD 0 arrives in z-slabs of whole brick rows
extern "C" int tvk_build_stream_begin(tvk_ctx* c, const tvk_build_desc* d, void* store, tvk_build_info* info_out) {
  if (!c || !d) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int rc = build_setup(c, d, store);
  if (rc) { release(c); return rc; }
  BuildState& b = c->build;
  rc = (d->bytes_per_voxel == 1) ? build_begin<uint8_t>(c, b) : build_begin<uint16_t>(c, b);
  if (rc) { release(c); return rc; }
  b.streaming = true; b.next_bz = 0;
  if (info_out) *info_out = b.info;
  return TVK_OK;
}

extern "C" int tvk_build_stream_rows(const tvk_build_desc* d, uint32_t bz0, uint32_t bz1, uint32_t* z0, uint32_t* rows) {
  tvk_build_info info;
  if (!d || !z0 || !rows || tvk_build_plan(d, &info) != TVK_OK || bz0 >= bz1 || bz1 > info.lod_bricks[0][2]) return TVK_ERR_INVALID;
  const uint32_t core = d->brick_size - 2 * d->overlap, sz = d->volume_size[2];
  const uint32_t lo = bz0 * core >= d->overlap ? bz0 * core - d->overlap : 0u;
  const uint32_t hi = std::min<uint64_t>((uint64_t)bz1 * core + d->overlap, sz);
  *z0 = lo; *rows = hi - lo;
  return TVK_OK;
}

extern "C" int tvk_build_stream_slab(tvk_ctx* c, const void* raw_slab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1) {
  if (!c || !raw_slab) return TVK_ERR_INVALID;
  BuildState& b = c->build;
  if (!b.streaming) TVK_FAIL(c, TVK_ERR_INVALID, "tvk_build_stream_slab without tvk_build_stream_begin");
  if (bz0 != b.next_bz || bz1 <= bz0 || bz1 > b.info.lod_bricks[0][2]) TVK_FAIL(c, TVK_ERR_INVALID, "slabs must cover the brick rows in ascending order, without gaps");
  uint32_t need_z0 = 0, need_rows = 0;
  tvk_build_stream_rows(&b.desc, bz0, bz1, &need_z0, &need_rows);
  if (z0 > need_z0 || (uint64_t)z0 + rows < (uint64_t)need_z0 + need_rows || (uint64_t)z0 + rows > b.info.lod_size[0][2])
    TVK_FAIL(c, TVK_ERR_INVALID, "the slab does not hold the raw rows its brick rows read (tvk_build_stream_rows)");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const BuildConst bc = make_const(b);
  const int rc = (b.desc.bytes_per_voxel == 1)
      ? build_lod0_rows<uint8_t>(c, b, bc, sms, (const uint8_t*)raw_slab, z0, rows, bz0, bz1, false)
      : build_lod0_rows<uint16_t>(c, b, bc, sms, (const uint16_t*)raw_slab, z0, rows, bz0, bz1, false);
  if (rc) return rc;
  b.next_bz = bz1;
  return TVK_OK;
}

extern "C" int tvk_build_stream_end(tvk_ctx* c, tvk_build_info* info_out) {
  if (!c) return TVK_ERR_INVALID;
  BuildState& b = c->build;
  if (!b.streaming) TVK_FAIL(c, TVK_ERR_INVALID, "tvk_build_stream_end without tvk_build_stream_begin");
  if (b.next_bz != b.info.lod_bricks[0][2]) TVK_FAIL(c, TVK_ERR_INVALID, "not every brick row of LOD 0 has been fed");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const BuildConst bc = make_const(b);
  const int rc = (b.desc.bytes_per_voxel == 1) ? build_finish<uint8_t>(c, b, bc, sms) : build_finish<uint16_t>(c, b, bc, sms);
  if (rc) { release(c); return rc; }
  b.streaming = false;
  b.valid = true;                                // no flat LOD 0: tvk_build_hist2d needs the raw volume and refuses
  if (info_out) *info_out = b.info;
  return TVK_OK;
}

extern "C" int tvk_build_forget_raw(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  BuildState& b = c->build;
  if (!b.own_lod0) b.d_lod[0] = nullptr;           // caller-owned flat volume: it may be freed from now on
  return TVK_OK;
}

extern "C" int tvk_build_get_toc(tvk_ctx* c, uint64_t* offsets, uint32_t* sizes3, uint64_t n) {
  if (!c || !c->build.valid || n != c->build.info.total_bricks) return TVK_ERR_INVALID;
  if (offsets) memcpy(offsets, c->build.offsets.data(), sizeof(uint64_t) * n);
  if (sizes3) memcpy(sizes3, c->build.sizes.data(), sizeof(uint32_t) * 3 * n);
  return TVK_OK;
}

extern "C" int tvk_build_get_minmax(tvk_ctx* c, double* minmax2, uint64_t n) {
  if (!c || !c->build.valid || !minmax2 || n != c->build.info.total_bricks) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpy(minmax2, c->build.d_minmax, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  return TVK_OK;
}

extern "C" int tvk_build_get_bricks(tvk_ctx* c, void* host_dst, uint64_t offset, uint64_t bytes) {
  if (!c || !c->build.valid || !host_dst || offset + bytes > c->build.info.total_bytes) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpy(host_dst, (const uint8_t*)c->build.d_store + offset, bytes, cudaMemcpyDeviceToHost));
  return TVK_OK;
}

extern "C" int tvk_build_last_ms(tvk_ctx* c, float* kernels_ms) {
  if (!c || !c->build.valid || !kernels_ms) return TVK_ERR_INVALID;
  *kernels_ms = c->build.kernels_ms;
  return TVK_OK;
}

extern "C" int tvk_build_hist2d_last_ms(tvk_ctx* c, float* kernels_ms) {
  if (!c || !c->build.valid || !kernels_ms) return TVK_ERR_INVALID;
  *kernels_ms = c->build.hist2d_ms;
  return TVK_OK;
}

extern "C" int tvk_build_store_device_ptr(tvk_ctx* c, void** dptr) {
  if (!c || !c->build.valid || !dptr) return TVK_ERR_INVALID;
  *dptr = c->build.d_store;
  return TVK_OK;
}

extern "C" int tvk_build_hist1d(tvk_ctx* c, uint64_t* bins, uint32_t cap, uint32_t* n) {
  if (!c || !c->build.valid || !bins || !n) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const uint32_t range = 1u << (8 * c->build.desc.bytes_per_voxel);
  std::vector<unsigned long long> h(range);
  TVK_CHECK_CUDA(c, cudaMemcpy(h.data(), c->build.d_hist1d, sizeof(unsigned long long) * range, cudaMemcpyDeviceToHost));
  uint32_t size = 0;
  for (uint32_t i = 0; i < range; ++i) if (h[i]) size = i + 1;    // trim after the last non-zero bin
  if (size > cap) TVK_FAIL(c, TVK_ERR_INVALID, "histogram buffer too small");
  for (uint32_t i = 0; i < size; ++i) bins[i] = h[i];
  *n = size;
  return TVK_OK;
}

template <typename T>
static int run_hist2d(tvk_ctx* c, uint32_t bins_x, double max_nonzero, uint64_t* bins, float* max_grad, int sms) {
  BuildState& b = c->build;
  const uint32_t* s = b.info.lod_size[0];
  const double norm2 = (sizeof(T) == 1 ? 255.0 : 65535.0) * 2;
  const uint32_t range = 1u << (8 * sizeof(T));
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)s[1] * s[2] + 7) / 8, (uint64_t)sms * 16);
  GradBest* d_part = nullptr;
  unsigned long long* d_bins = nullptr;
  uint16_t* d_lut = nullptr;
  TVK_CHECK_CUDA(c, cudaMalloc(&d_part, sizeof(GradBest) * blocks));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_part, 0, sizeof(GradBest) * blocks, c->stream));    // (the tiled kernel launches fewer CTAs)
  TVK_CHECK_CUDA(c, cudaMalloc(&d_bins, sizeof(unsigned long long) * 256 * bins_x));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_lut, sizeof(uint16_t) * range));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(d_bins, 0, sizeof(unsigned long long) * 256 * bins_x, c->stream));
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev0, c->stream));
  // 4 voxels per thread when rows are whole quads (cudaMalloc'ed / caller-aligned base); TVK_HIST2D_SCALAR=1 forces the
  // one-voxel-per-thread kernels
  // (clamp-to-edge volumes: the one-voxel-per-thread kernels, which know that a neighbour outside the domain is the voxel itself)
  const bool clamp = b.desc.clamp_to_edge != 0;
  const bool vec = (s[0] % 4u) == 0 && (reinterpret_cast<uintptr_t>(b.d_lod[0]) % (4 * sizeof(T))) == 0 && !getenv("TVK_HIST2D_SCALAR") && !clamp;
  // 8 voxels per thread + fp32 squares when rows are whole octets and 16-byte aligned (TVK_HIST2D_V4=1: the 4-voxel kernels)
  const bool v8 = vec && (s[0] % 8u) == 0 && (reinterpret_cast<uintptr_t>(b.d_lod[0]) % (8 * sizeof(T))) == 0 &&
                  (uint64_t)bins_x * 256u <= (1ull << TVK_H2D8_BIN_BITS) && !getenv("TVK_HIST2D_V4");
  if (v8) grad_max_v8_kernel<T><<<(unsigned)std::min<uint64_t>((uint64_t)((s[1] + 7) / 8) * ((s[2] + 15) / 16), (uint64_t)sms * 8), 256, 0, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, d_part, d_bins);
  else if (vec) grad_max_vec_kernel<T><<<(unsigned)std::min<uint64_t>((uint64_t)((s[1] + 7) / 8) * ((s[2] + 15) / 16), (uint64_t)sms * 8), 256, 0, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, d_part, d_bins);   // d_bins[0]: zeroed scratch, re-zeroed below
  else grad_max_kernel<T><<<blocks, 256, 0, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, d_part, clamp);
  hist2d_value_lut_kernel<<<(range + 255) / 256, 256, 0, c->stream>>>(range, bins_x, max_nonzero, d_lut);
  c->launches += 2;
  std::vector<GradBest> part(blocks);
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(part.data(), d_part, sizeof(GradBest) * blocks, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  GradBest g = part[0];
  for (unsigned i = 1; i < blocks; ++i) if (part[i].s > g.s || (part[i].s == g.s && part[i].len_bits > g.len_bits)) g = part[i];
  double mg;
  memcpy(&mg, &g.len_bits, sizeof(mg));
  const float inv_smax = g.s ? 1.0f / (float)g.s : 0.0f;
  if (vec) TVK_CHECK_CUDA(c, cudaMemsetAsync(d_bins, 0, sizeof(unsigned long long), c->stream));
  if (v8) {
    const uint64_t tiles = (uint64_t)((s[0] / 8 + TVK_H2D8_XO - 1) / TVK_H2D8_XO) * ((s[1] + 7) / 8) * ((s[2] + TVK_H2D8_ZT - 1) / TVK_H2D8_ZT);
    const size_t smem = sizeof(unsigned long long) * 8 * 64 + sizeof(uint32_t) * TVK_H2D8_SLOTS;
    TVK_CHECK_CUDA(c, cudaFuncSetAttribute(hist2d_v8_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TVK_CHECK_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hist2d_v8_kernel<T>, 256, smem));
    // runs of consecutive tiles per CTA; 4 waves of CTAs, so that runs through empty space do not leave SMs idle
    const unsigned tb = (unsigned)std::min<uint64_t>(tiles, (uint64_t)sms * std::max(per_sm, 1) * 4);
    const float k_smax = g.s ? (float)(65025.0 / (double)g.s) : 0.0f;
    const float tie_thr = (float)((double)g.s * (1.0 - 1e-6));
    hist2d_v8_kernel<T><<<tb, 256, smem, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, mg, g.s, k_smax, tie_thr, d_lut, d_bins);
  }
  else if (vec) {
    const uint32_t xq = TVK_H2D_XQ ? TVK_H2D_XQ : s[0] / 4;
    const uint64_t tiles = (uint64_t)((s[0] / 4 + xq - 1) / xq) * ((s[1] + 7) / 8) * ((s[2] + TVK_H2D_ZT - 1) / TVK_H2D_ZT);
    const unsigned tb = (unsigned)std::min<uint64_t>(tiles, (uint64_t)sms * (sizeof(T) == 1 ? 16 : 6));
    hist2d_vec_kernel<T><<<tb, 256, 0, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, mg, g.s, inv_smax, d_lut, d_bins);
  }
  else hist2d_kernel<T><<<blocks, 256, 0, c->stream>>>((const T*)b.d_lod[0], s[0], s[1], s[2], norm2, mg, g.s, inv_smax, d_lut, d_bins, clamp);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaEventRecord(c->bev1, c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(bins, d_bins, sizeof(unsigned long long) * 256 * bins_x, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->bev0, c->bev1) == cudaSuccess) b.hist2d_ms = ms;
  cudaFree(d_part); cudaFree(d_bins); cudaFree(d_lut);
  *max_grad = (float)mg;
  return TVK_OK;
}

extern "C" int tvk_build_hist2d(tvk_ctx* c, uint32_t bins_x, double max_nonzero_value, uint64_t* bins, float* max_grad_magnitude) {
  if (!c || !c->build.valid || !bins || !max_grad_magnitude || bins_x == 0) return TVK_ERR_INVALID;
  if (!c->build.d_lod[0]) TVK_FAIL(c, TVK_ERR_UNSUPPORTED, "the 2-D histogram reads the flat raw volume, which a streamed build does not keep");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  return c->build.desc.bytes_per_voxel == 1 ? run_hist2d<uint8_t>(c, bins_x, max_nonzero_value, bins, max_grad_magnitude, sms)
                                            : run_hist2d<uint16_t>(c, bins_x, max_nonzero_value, bins, max_grad_magnitude, sms);
}
