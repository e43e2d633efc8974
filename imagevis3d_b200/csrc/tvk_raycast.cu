// K1 -- persistent-warp ray-march kernel (sm_100a).
//
// Replaces, in one launch per sub-frame, what the reference does with two raster
// passes and a fragment shader (paths relative to /root/reference/Tuvok/):
//   ray entry / exit      GLGridLeaper::FillRayEntryBuffer + bbox back faces
//                         (Renderer/GL/GLGridLeaper.cpp:579-620,835-843)          -> ray_setup()
//   traversal             Shaders/GLGridLeaper-blend.glsl:65-228, -iso.glsl:68-200 -> march_*()
//   page-table walk       generated GLSL, Renderer/GL/GLVolumePool.cpp:484-656    -> get_brick()
//   request table         generated GLSL, Renderer/GL/GLHashTable.cpp:136-182     -> hash_report()
//   sample -> colour      Shaders/GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}.glsl,
//                         GLGridLeaper-GradientTools.glsl, lighting.glsl, Compositing.glsl
//   iso shading           Shaders/Compose-FS.glsl:47-75 (fused into the epilogue)
//
// Execution model: a persistent grid (CTAs x 4 warps, sized from the SM count);
// every warp repeatedly claims an 8x4 pixel tile from a global counter, so
// early-terminated / empty-space tiles cost nothing and the tail is balanced.
// Neighbouring rays of a tile walk the same bricks, so their 2x2x2 voxel
// footprints share L1 lines.  The page table (<= 6 MB) and the transfer function
// stay L2/L1 resident.  Ray state (resume pos / colour) lives in HBM as float4
// planes and is updated in place: each pixel only ever touches its own entry.
#include "tvk_internal.h"
#include <cmath>
#include <cstring>

// Tuning knobs of the 1D-TF sample loop (swept on a B200, see profiles/README.md): samples in flight
// per ray and resident CTAs per SM, for the hardware-filtered and the exact (LDG) samplers.
#ifndef TVK_NB_TEX
#define TVK_NB_TEX 8
#endif
#ifndef TVK_NB_LDG
#define TVK_NB_LDG 4
#endif
#ifndef TVK_OCC_TEX
#define TVK_OCC_TEX 6
#endif
#ifndef TVK_OCC_LDG
#define TVK_OCC_LDG 4
#endif
#ifndef TVK_OCC_LIT
#define TVK_OCC_LIT 4
#endif
#ifndef TVK_NB_GATHER
#define TVK_NB_GATHER 4
#endif
#ifndef TVK_OCC_GATHER
#define TVK_OCC_GATHER 4
#endif
// lit / 2D-TF modes: centre samples whose footprints (and table look-ups) are in flight together
#ifndef TVK_NB_LIT
#define TVK_NB_LIT 2
#endif

struct RayParams {
  tvk_raycast_params p;
  PoolConst pc;
  const void* pool;
  uint32_t elem_bytes;   // bytes per pool element (bpv, or 4*bpv for the quad-packed pool)
  cudaTextureObject_t tex;
  const uint32_t* meta;
  const uchar4* tf;
  const float4* tf_f;   // the same table pre-converted to float (byte * (1/255)), 1D mode
  const uint32_t* tf_colmask;   // bit x: some row of column x has alpha != 0
  uint32_t tf_w, tf_h;
  uint32_t* hash;
  float4* pos;      // rayStartPoint / rayResumePos
  float4* aux;      // rayStartColor / rayResumeColor (iso: normal)
  float4* image;    // accRayColor (iso: composed image)
  float4* hit_buf;  // iso: rayHitPos (m_pFBOIsoHit buffer 0) kept for Pick; nullptr otherwise
  uint32_t* tile_counter;
  unsigned long long* stats;
  uint32_t tiles_x, tiles_y;     // 8x4-pixel tiles the launch covers ...
  uint32_t tile_x0, tile_y0;     // ... starting at this tile: the screen rectangle of the (clipped) volume box
  ExchangeSend xs;       // sort-last: finished pixels also go to the stripe owner's receive buffer (peer store)
};

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
// Floating-point contract (shared with oracle/tvk_oracle_render.c): this file is compiled with
// -fmad=false, so the ONLY fused multiply-adds are the explicit fmaf() calls, and they sit exactly
// where the oracle has them.  Every other operation is a single IEEE fp32 op in the oracle's order -- with one
// exception: the SHADING block (normalisations in lighting(), the normal, the model-space position) uses
// rsqrtf / a per-brick reciprocal instead of IEEE sqrt + divisions (12 divisions and 3 square roots per lit
// sample otherwise).  Nothing discrete depends on it -- transfer-function indices, LOD, brick walk and step
// counts stay exact -- and the colour changes by ~2 ulp (the reference's GLSL normalize() is no more precise).
__device__ __forceinline__ float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float len3(f3 v) { return sqrtf(dot3(v, v)); }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// v' = v*M, FLOATMATRIX4 memory order (Basics/Vectors.h:434)
__device__ __forceinline__ float4 xform(const float* m, f3 v, float w) {
  float4 o;
  o.x = fmaf(w, m[12], fmaf(v.z, m[8], fmaf(v.y, m[4], v.x * m[0])));
  o.y = fmaf(w, m[13], fmaf(v.z, m[9], fmaf(v.y, m[5], v.x * m[1])));
  o.z = fmaf(w, m[14], fmaf(v.z, m[10], fmaf(v.y, m[6], v.x * m[2])));
  o.w = fmaf(w, m[15], fmaf(v.z, m[11], fmaf(v.y, m[7], v.x * m[3])));
  return o;
}

// ComputeLOD, GLVolumePool.cpp:653-656.  log2 is evaluated in fp64 and rounded
// to fp32 so that the truncation agrees with a correctly rounded log2f on the
// CPU side (bit-exact LOD selection); negative / NaN saturate to 0 like NVIDIA's
// GLSL uint().
__device__ __forceinline__ uint32_t compute_lod(float dist, float lod_factor, float wse, uint32_t max_lod) {
  const float x = lod_factor * (-dist) / wse;
  const float v = (float)log2((double)x);
  const uint32_t l = (v > 0.0f) ? ((v >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t)v) : 0u;
  return min(l, max_lod);
}

__device__ __forceinline__ uint4 brick_coords(const PoolConst& pc, f3 p, uint32_t lod) {
  return make_uint4((uint32_t)(p.x * pc.lod_layout[lod][0]), (uint32_t)(p.y * pc.lod_layout[lod][1]),
                    (uint32_t)(p.z * pc.lod_layout[lod][2]), lod);
}
__device__ __forceinline__ uint32_t brick_index(const PoolConst& pc, uint4 bc) {
  return pc.lod_offset[bc.w] + bc.x + bc.y * pc.lod_layout_x[bc.w] + bc.z * pc.lod_layout_xy[bc.w];
}
__device__ __forceinline__ uint32_t serialize(const PoolConst& pc, uint4 bc) {
  return 1u + bc.x + bc.y * pc.l0_bricks[0] + bc.z * pc.l0_bricks[0] * pc.l0_bricks[1] +
         bc.w * pc.l0_bricks[0] * pc.l0_bricks[1] * pc.l0_bricks[2];
}

// Hash(), GLHashTable.cpp:170-182.  A plain read first: once a brick is in the
// table (the common case: thousands of rays report the same brick) no atomic is
// issued at all; semantics are those of the CAS loop.
__device__ __forceinline__ void hash_report(const RayParams& rp, uint4 bc) {
  const uint32_t serial = serialize(rp.pc, bc);
  for (uint32_t rehash = 0; rehash < rp.p.rehash_count; ++rehash) {
    const uint32_t h = (serial + rehash) % rp.p.hash_table_size;
    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(rp.hash + h);
    if (cur == serial) return;
    if (cur == 0u) {
      cur = atomicCAS(rp.hash + h, 0u, serial);
      if (cur == 0u || cur == serial) return;
    }
  }
}

// Where the current brick lives in the pool: pointer to the slot's first element and the slot's
// origin in pool voxels (pre-biased by the floor-trick constant), so that tap addresses are one
// 64-bit base + a 32-bit in-slot offset.
struct SlotCtx {
  const uint8_t* base;
  int ox, oy, oz;
  // TVK_SAMPLER_GATHER: what turns a LOGICAL pool voxel coordinate (the reference's texture addressing) into the
  // PHYSICAL position of that voxel in the layered array: gather coordinate = floor(u) + gdx (exact in fp32: small
  // integers; includes the +1 that makes a gather at (i+1, j+1) return texels i..i+1 x j..j+1), layer = floor(w) + ldz.
  float gdx, gdy;
  int ldz;               // added to the float BITS of floor(w) + 1.5*2^23 (the round-down trick's raw result)
};

struct BrickHit {
  f3 pool_entry, pool_exit, norm_exit, scale, trans, inv_scale;   // inv_scale: for the shading block only
  uint4 bc;
  SlotCtx slot;
  bool empty, found;
};

// GetBrick, GLVolumePool.cpp:552-620
template <int SAMP>
__device__ __forceinline__ void get_brick(const RayParams& rp, f3 entry, uint32_t& lod, f3 dir, BrickHit& b) {
  const PoolConst& pc = rp.pc;
  const uint32_t max_lod = pc.lod_count - 1;
  const f3 e = mk3(clamp01(entry.x), clamp01(entry.y), clamp01(entry.z));
  b.found = true;
  b.bc = brick_coords(pc, e, lod);
  uint32_t info = __ldg(rp.meta + brick_index(pc, b.bc));
  if (info == TVK_BI_MISSING) {
    const uint32_t start = lod;
    hash_report(rp, b.bc);
    b.found = false;
    do {
      lod++;
      b.bc = brick_coords(pc, e, lod);
      info = __ldg(rp.meta + brick_index(pc, b.bc));
      if (info == TVK_BI_MISSING) {
        const uint32_t s = rp.p.missing_strategy;
        if (s == TVK_MBS_REQUEST_ALL || (s == TVK_MBS_SKIP_ONE_LEVEL && start + 1 == lod) ||
            (s == TVK_MBS_SKIP_TWO_LEVELS && start + 2 == lod))
          hash_report(rp, b.bc);
      }
    } while (info == TVK_BI_MISSING && lod < max_lod);   // the coarsest brick is pinned resident
  }
  b.empty = (info <= TVK_BI_EMPTY);
  if (b.empty) {
    for (uint32_t l = lod + 1; l < max_lod; ++l) {
      const uint4 lbc = brick_coords(pc, e, l);
      const uint32_t linfo = __ldg(rp.meta + brick_index(pc, lbc));
      if (linfo == TVK_BI_CHILD_EMPTY) { b.bc = lbc; info = linfo; lod = l; }
      else break;
    }
  }
  const float lx = pc.lod_layout[b.bc.w][0], ly = pc.lod_layout[b.bc.w][1], lz = pc.lod_layout[b.bc.w][2];
  const f3 c0 = mk3((float)b.bc.x / lx, (float)b.bc.y / ly, (float)b.bc.z / lz);
  const f3 c1 = mk3((float)(b.bc.x + 1) / lx, (float)(b.bc.y + 1) / ly, (float)(b.bc.z + 1) / lz);
  // BrickExit: side = step(0, 1/dir)
  const float dvx = 1.0f / dir.x, dvy = 1.0f / dir.y, dvz = 1.0f / dir.z;
  const float tx = ((dvx >= 0.0f ? c1.x : c0.x) - e.x) * dvx;
  const float ty = ((dvy >= 0.0f ? c1.y : c0.y) - e.y) * dvy;
  const float tz = ((dvz >= 0.0f ? c1.z : c0.z) - e.z) * dvz;
  const float tmin = fminf(fminf(fminf(INFINITY, tx), ty), tz);   // fminf drops NaN (0*inf)
  b.norm_exit = mk3(fmaf(tmin, dir.x, e.x), fmaf(tmin, dir.y, e.y), fmaf(tmin, dir.z, e.z));
  if (b.empty) return;
  const uint32_t idx = info - TVK_BI_FLAG_COUNT;
  const uint32_t s0 = idx % pc.capacity[0], s1 = (idx / pc.capacity[0]) % pc.capacity[1],
                 s2 = idx / (pc.capacity[0] * pc.capacity[1]);
  {
    const uint32_t vx = s0 * pc.max_brick[0], vy = s1 * pc.max_brick[1], vz = s2 * pc.max_brick[2];
    const size_t sx = pc.pool_dim[0], sxy = sx * pc.pool_dim[1];
    b.slot.base = static_cast<const uint8_t*>(rp.pool) + ((size_t)vx + (size_t)vy * sx + (size_t)vz * sxy) * rp.elem_bytes;
    b.slot.ox = 0x4B400000 + (int)vx; b.slot.oy = 0x4B400000 + (int)vy; b.slot.oz = 0x4B400000 + (int)vz;
    if (SAMP == TVK_SAMPLER_GATHER) {
      const uint32_t p0 = idx % pc.phys_cap[0], p1 = (idx / pc.phys_cap[0]) % pc.phys_cap[1], p2 = idx / (pc.phys_cap[0] * pc.phys_cap[1]);
      b.slot.gdx = ((float)(p0 * pc.max_brick[0]) + 1.0f) - (float)vx;
      b.slot.gdy = ((float)(p1 * pc.max_brick[1]) + 1.0f) - (float)vy;
      b.slot.ldz = (int)(p2 * pc.max_brick[2]) - b.slot.oz;
    }
  }
  {
    const uint32_t v = s0 * pc.max_brick[0];
    const float p0 = (float)v / pc.pool_size[0] + pc.overlap_tc[0];
    const float p1 = (float)(v + pc.max_brick[0]) / pc.pool_size[0] - pc.overlap_tc[0];
    b.scale.x = (p1 - p0) / (c1.x - c0.x);
    b.trans.x = fmaf(-c0.x, b.scale.x, p0);
  }
  {
    const uint32_t v = s1 * pc.max_brick[1];
    const float p0 = (float)v / pc.pool_size[1] + pc.overlap_tc[1];
    const float p1 = (float)(v + pc.max_brick[1]) / pc.pool_size[1] - pc.overlap_tc[1];
    b.scale.y = (p1 - p0) / (c1.y - c0.y);
    b.trans.y = fmaf(-c0.y, b.scale.y, p0);
  }
  {
    const uint32_t v = s2 * pc.max_brick[2];
    const float p0 = (float)v / pc.pool_size[2] + pc.overlap_tc[2];
    const float p1 = (float)(v + pc.max_brick[2]) / pc.pool_size[2] - pc.overlap_tc[2];
    b.scale.z = (p1 - p0) / (c1.z - c0.z);
    b.trans.z = fmaf(-c0.z, b.scale.z, p0);
  }
  b.inv_scale = mk3(1.0f / b.scale.x, 1.0f / b.scale.y, 1.0f / b.scale.z);
  b.pool_entry = mk3(fmaf(e.x, b.scale.x, b.trans.x), fmaf(e.y, b.scale.y, b.trans.y), fmaf(e.z, b.scale.z, b.trans.z));
  b.pool_exit = mk3(fmaf(b.norm_exit.x, b.scale.x, b.trans.x), fmaf(b.norm_exit.y, b.scale.y, b.trans.y),
                    fmaf(b.norm_exit.z, b.scale.z, b.trans.z));
}

// Exact small-integer -> float without the quarter-rate I2F pipe: OR the value into the
// mantissa of 2^23 and subtract 2^23 (PRMT/LOP3 + FADD).
__device__ __forceinline__ float u2f(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

// Four-texel gather from one layer of the layered pool array (PTX tld4, 2-D array geometry): returns the texels
// around the shared corner (x, y) in the order  .x = (i, j+1)  .y = (i+1, j+1)  .z = (i+1, j)  .w = (i, j)  with
// i = x - 1, j = y - 1 (unnormalised coordinates; the corner is half a texel away from every texel boundary, so the
// hardware's fixed-point footprint selection cannot disagree with ours).  Raw integer texels, zero-extended.
__device__ __forceinline__ uint4 gather4(cudaTextureObject_t t, uint32_t layer, float x, float y) {
  uint4 r;
  asm("tld4.r.a2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(t), "r"(layer), "f"(x), "f"(y));
  return r;
}

// The 2x2x2 voxel footprint of one sample: raw values + lerp weights.
struct Taps {
  float v[4];       // base taps 000 010 001 011 (the x = 0 corner of each x-edge)
  float d[4];       // x-differences v(1,y,z) - v(0,y,z) of the same four edges
  float a, b, g;
};

// Address + fetch half of texture(volumePool, p): GL_LINEAR footprint of p.
// floor() is taken with a round-down magic add (valid for |u| < 2^22, i.e. any pool).
// The base voxel is clamped to [0, dim-2]: inside a brick the footprint (and the
// gradient taps one texel further) never leaves the slot -- 2 ghost voxels per side --
// so this equals GL_CLAMP_TO_EDGE for every position the march can produce, and it lets
// the eight loads share one address.  (The clamp is applied to the in-slot coordinate.)
template <int BPV, int SAMP>
__device__ __forceinline__ void fetch_taps(const RayParams& rp, const SlotCtx& sl, f3 p, Taps& t) {
  const PoolConst& pc = rp.pc;
  const float M = 12582912.0f;   // 1.5 * 2^23
  const float u = fmaf(p.x, pc.pool_size[0], -0.5f), v = fmaf(p.y, pc.pool_size[1], -0.5f), w = fmaf(p.z, pc.pool_size[2], -0.5f);
  const float tu = __fadd_rd(u, M), tv = __fadd_rd(v, M), tw = __fadd_rd(w, M);
  t.a = u - (tu - M); t.b = v - (tv - M); t.g = w - (tw - M);
  if (SAMP == TVK_SAMPLER_GATHER) {
    // Two gathers at the footprint's base corner.  No clamp to the slot: a position the march can produce keeps its
    // footprint inside the brick's ghost layers (the integer path below clamps only to be able to share one
    // address), and the texture unit itself clamps coordinates and layer index at the borders of the array.
    const float gx = (tu - M) + sl.gdx, gy = (tv - M) + sl.gdy;
    const uint32_t l = (uint32_t)(__float_as_int(tw) + sl.ldz);
    const uint4 g0 = gather4(rp.tex, l, gx, gy), g1 = gather4(rp.tex, l + 1u, gx, gy);
    // biased values 2^23 + texel: differences of biased values are exact, only the base taps drop the bias
    const float b000 = __uint_as_float(0x4B000000u | g0.w), b100 = __uint_as_float(0x4B000000u | g0.z);
    const float b010 = __uint_as_float(0x4B000000u | g0.x), b110 = __uint_as_float(0x4B000000u | g0.y);
    const float b001 = __uint_as_float(0x4B000000u | g1.w), b101 = __uint_as_float(0x4B000000u | g1.z);
    const float b011 = __uint_as_float(0x4B000000u | g1.x), b111 = __uint_as_float(0x4B000000u | g1.y);
    t.d[0] = b100 - b000; t.d[1] = b110 - b010; t.d[2] = b101 - b001; t.d[3] = b111 - b011;
    t.v[0] = b000 - 8388608.0f; t.v[1] = b010 - 8388608.0f; t.v[2] = b001 - 8388608.0f; t.v[3] = b011 - 8388608.0f;
    return;
  }
  // in-slot voxel coordinates of the footprint's base corner
  const int x = min(max(__float_as_int(tu) - sl.ox, 0), (int)pc.max_brick[0] - 2);
  const int y = min(max(__float_as_int(tv) - sl.oy, 0), (int)pc.max_brick[1] - 2);
  const int z = min(max(__float_as_int(tw) - sl.oz, 0), (int)pc.max_brick[2] - 2);
  const uint32_t sx = pc.pool_dim[0], sxy = sx * pc.pool_dim[1];
  const uint32_t o = (uint32_t)x + (uint32_t)y * sx + (uint32_t)z * sxy;
  if (SAMP == TVK_SAMPLER_EXACT_Q4) {
    // two loads: the (y,z) quads of columns x and x+1
    if (BPV == 1) {
      const uint32_t* d = reinterpret_cast<const uint32_t*>(sl.base) + o;
      const uint32_t q0 = __ldg(d), q1 = __ldg(d + 1);
      // the x-differences are taken on the biased values (2^23 + byte: the difference is exact either way), so only
      // the four base taps need the bias removed
      const float b0 = __uint_as_float(__byte_perm(q0, 0x4B000000u, 0x7650)), b1 = __uint_as_float(__byte_perm(q1, 0x4B000000u, 0x7650));
      const float b2 = __uint_as_float(__byte_perm(q0, 0x4B000000u, 0x7651)), b3 = __uint_as_float(__byte_perm(q1, 0x4B000000u, 0x7651));
      const float b4 = __uint_as_float(__byte_perm(q0, 0x4B000000u, 0x7652)), b5 = __uint_as_float(__byte_perm(q1, 0x4B000000u, 0x7652));
      const float b6 = __uint_as_float(__byte_perm(q0, 0x4B000000u, 0x7653)), b7 = __uint_as_float(__byte_perm(q1, 0x4B000000u, 0x7653));
      t.d[0] = b1 - b0; t.d[1] = b3 - b2; t.d[2] = b5 - b4; t.d[3] = b7 - b6;
      t.v[0] = b0 - 8388608.0f; t.v[1] = b2 - 8388608.0f; t.v[2] = b4 - 8388608.0f; t.v[3] = b6 - 8388608.0f;
    } else {
      const uint2* d = reinterpret_cast<const uint2*>(sl.base) + o;
      const uint2 q0 = __ldg(d), q1 = __ldg(d + 1);
      t.v[0] = u2f(q0.x & 0xFFFFu); t.v[1] = u2f(q0.x >> 16); t.v[2] = u2f(q0.y & 0xFFFFu); t.v[3] = u2f(q0.y >> 16);
      t.d[0] = u2f(q1.x & 0xFFFFu) - t.v[0]; t.d[1] = u2f(q1.x >> 16) - t.v[1];
      t.d[2] = u2f(q1.y & 0xFFFFu) - t.v[2]; t.d[3] = u2f(q1.y >> 16) - t.v[3];
    }
  } else if (BPV == 1) {
    const uint8_t* d = sl.base + o;
    const uint8_t *d1 = d + sx, *d2 = d + sxy, *d3 = d2 + sx;
    const uint32_t r0 = __ldg(d), r1 = __ldg(d + 1), r2 = __ldg(d1), r3 = __ldg(d1 + 1);
    const uint32_t r4 = __ldg(d2), r5 = __ldg(d2 + 1), r6 = __ldg(d3), r7 = __ldg(d3 + 1);
    t.v[0] = u2f(r0); t.v[1] = u2f(r2); t.v[2] = u2f(r4); t.v[3] = u2f(r6);
    t.d[0] = u2f(r1) - t.v[0]; t.d[1] = u2f(r3) - t.v[1]; t.d[2] = u2f(r5) - t.v[2]; t.d[3] = u2f(r7) - t.v[3];
  } else {
    const uint16_t* d = reinterpret_cast<const uint16_t*>(sl.base) + o;
    const uint16_t *d1 = d + sx, *d2 = d + sxy, *d3 = d2 + sx;
    const uint32_t r0 = __ldg(d), r1 = __ldg(d + 1), r2 = __ldg(d1), r3 = __ldg(d1 + 1);
    const uint32_t r4 = __ldg(d2), r5 = __ldg(d2 + 1), r6 = __ldg(d3), r7 = __ldg(d3 + 1);
    t.v[0] = u2f(r0); t.v[1] = u2f(r2); t.v[2] = u2f(r4); t.v[3] = u2f(r6);
    t.d[0] = u2f(r1) - t.v[0]; t.d[1] = u2f(r3) - t.v[1]; t.d[2] = u2f(r5) - t.v[2]; t.d[3] = u2f(r7) - t.v[3];
  }
}

// Filter half: fp32 trilinear on the raw values, then the UNORM scale.
template <int BPV>
__device__ __forceinline__ float filter_taps(const Taps& t) {
  const float c00 = fmaf(t.a, t.d[0], t.v[0]), c10 = fmaf(t.a, t.d[1], t.v[1]);
  const float c01 = fmaf(t.a, t.d[2], t.v[2]), c11 = fmaf(t.a, t.d[3], t.v[3]);
  const float c0 = fmaf(t.b, c10 - c00, c00), c1 = fmaf(t.b, c11 - c01, c01);
  return fmaf(t.g, c1 - c0, c0) * (BPV == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}

// texture(volumePool, p).r -- GL_LINEAR, CLAMP_TO_EDGE, UNORM8/16.
template <int BPV, int SAMP>
__device__ __forceinline__ float sample_pool(const RayParams& rp, const SlotCtx& sl, f3 p) {
  if (SAMP == TVK_SAMPLER_TEX3D) {
    return tex3D<float>(rp.tex, p.x, p.y, p.z);
  } else {
    Taps t;
    fetch_taps<BPV, SAMP>(rp, sl, p, t);
    return filter_taps<BPV>(t);
  }
}

// GL_NEAREST + clamp-to-edge index: floor(s*n) clamped to [0, n-1].  F2I truncation equals
// floor for s*n >= 0, everything below clamps to 0, NaN converts to 0, +inf saturates.
__device__ __forceinline__ int nearest_index(float s, uint32_t n) {
  return min(max(__float2int_rz(s * (float)n), 0), (int)n - 1);
}
__device__ __forceinline__ float4 tf_fetch(const RayParams& rp, int i) {
  const uchar4 t = __ldg(rp.tf + i);
  const float k = 1.0f / 255.0f;
  return make_float4(u2f(t.x) * k, u2f(t.y) * k, u2f(t.z) * k, u2f(t.w) * k);
}

// ComputeGradient, GLGridLeaper-GradientTools.glsl:6-16 (y sign flipped there)
template <int BPV, int SAMP>
__device__ __forceinline__ f3 gradient(const RayParams& rp, const SlotCtx& sl, f3 p) {
  const float dx = 1.0f / rp.pc.pool_size[0], dy = 1.0f / rp.pc.pool_size[1], dz = 1.0f / rp.pc.pool_size[2];
  const float xp = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x + dx, p.y, p.z)), xm = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x - dx, p.y, p.z));
  const float yp = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x, p.y - dy, p.z)), ym = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x, p.y + dy, p.z));
  const float zp = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x, p.y, p.z + dz)), zm = sample_pool<BPV, SAMP>(rp, sl, mk3(p.x, p.y, p.z - dz));
  return mk3((xm - xp) / 2.0f, (yp - ym) / 2.0f, (zm - zp) / 2.0f);
}


// Centre sample AND central-difference gradient from ONE shared footprint (quad-packed pool only).
// gradient() above evaluates six more trilinear samples at +-1 texel, each with its own address arithmetic and
// eight-tap footprint: 7 footprints = 14 quad loads and ~500 instructions per lit / 2D-TF sample.  All seven
// share the fractional weights, and their footprints overlap: the union is a 4x2x2 + 2x4x2 + 2x2x4 cross of 32
// voxels, which the (y,z)-quad layout delivers in 12 loads around one base address.  The CENTRE value is filtered
// exactly as filter_taps does (x, then y, then z: it selects the transfer-function entry); the six offset samples
// are filtered with the same weights in whatever order needs the fewest lerps, so the gradient differs from the
// seven-sample evaluation by rounding only (~1e-7 relative; the image tolerance is 2/255).
template <int BPV>
__device__ __forceinline__ void unpack_quad(const void* __restrict__ base, uint32_t o, float q[4]) {
  if (BPV == 1) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(base) + o);
    q[0] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388608.0f;
    q[1] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388608.0f;
    q[2] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388608.0f;
    q[3] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388608.0f;
  } else {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(base) + o);
    q[0] = u2f(w.x & 0xFFFFu); q[1] = u2f(w.x >> 16); q[2] = u2f(w.y & 0xFFFFu); q[3] = u2f(w.y >> 16);
  }
}
__device__ __forceinline__ float lerpf(float t, float a, float b) { return fmaf(t, b - a, a); }

// The 32 voxels the centre sample and its six central-difference neighbours touch, named by what the arithmetic
// below needs: four x-columns of (y,z) quads {v(y,z), v(y+1,z), v(y,z+1), v(y+1,z+1)} at x-1, x, x+1, x+2, and the
// rows y-1 / y+2 and planes z-1 / z+2 over the 2x2 of the other two axes.
struct Cross {
  float X0[4], X1[4], X2[4], X3[4];
  float ym[4], yp[4];     // (x,y-1,z) (x+1,y-1,z) (x,y-1,z+1) (x+1,y-1,z+1);  the same at y+2
  float zm[4], zp[4];     // (x,y,z-1) (x+1,y,z-1) (x,y+1,z-1) (x+1,y+1,z-1);  the same at z+2
};

__device__ __forceinline__ void cross_filter(const Cross& k, float a, float b, float g, float scale, float& data, f3& grad) {
  const float* X0 = k.X0; const float* X1 = k.X1; const float* X2 = k.X2; const float* X3 = k.X3;
  // centre: filter_taps order
  const float c00 = fmaf(a, X2[0] - X1[0], X1[0]), c10 = fmaf(a, X2[1] - X1[1], X1[1]);
  const float c01 = fmaf(a, X2[2] - X1[2], X1[2]), c11 = fmaf(a, X2[3] - X1[3], X1[3]);
  const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
  data = fmaf(g, c1 - c0, c0) * scale;
  // x: bilinear (b, g) value of each of the four columns, then the two offset samples
  const float B0 = lerpf(g, lerpf(b, X0[0], X0[1]), lerpf(b, X0[2], X0[3])), B1 = lerpf(g, lerpf(b, X1[0], X1[1]), lerpf(b, X1[2], X1[3]));
  const float B2 = lerpf(g, lerpf(b, X2[0], X2[1]), lerpf(b, X2[2], X2[3])), B3 = lerpf(g, lerpf(b, X3[0], X3[1]), lerpf(b, X3[2], X3[3]));
  const float txm = lerpf(a, B0, B1), txp = lerpf(a, B2, B3);
  // y: rows y-1, y, y+1, y+2, each bilinear in (a, g)
  const float R0 = lerpf(g, lerpf(a, k.ym[0], k.ym[1]), lerpf(a, k.ym[2], k.ym[3]));
  const float R1 = lerpf(g, c00, c01), R2 = lerpf(g, c10, c11);
  const float R3 = lerpf(g, lerpf(a, k.yp[0], k.yp[1]), lerpf(a, k.yp[2], k.yp[3]));
  const float tym = lerpf(b, R0, R1), typ = lerpf(b, R2, R3);
  // z: planes z-1, z, z+1, z+2, each bilinear in (a, b)
  const float P0 = lerpf(b, lerpf(a, k.zm[0], k.zm[1]), lerpf(a, k.zm[2], k.zm[3]));
  const float P1 = c0, P2 = c1;
  const float P3 = lerpf(b, lerpf(a, k.zp[0], k.zp[1]), lerpf(a, k.zp[2], k.zp[3]));
  const float tzm = lerpf(g, P0, P1), tzp = lerpf(g, P2, P3);
  // ComputeGradient: (xm - xp, yp - ym, zm - zp) / 2 with yp = S(y - dy), ym = S(y + dy)  (GradientTools.glsl:6-16)
  grad = mk3((txm - txp) * scale / 2.0f, (tym - typ) * scale / 2.0f, (tzm - tzp) * scale / 2.0f);
}

template <int BPV, int SAMP>
__device__ __forceinline__ void center_gradient(const RayParams& rp, const SlotCtx& sl, f3 p, float& data, f3& grad) {
  const PoolConst& pc = rp.pc;
  const float M = 12582912.0f;   // 1.5 * 2^23
  const float u = fmaf(p.x, pc.pool_size[0], -0.5f), v = fmaf(p.y, pc.pool_size[1], -0.5f), w = fmaf(p.z, pc.pool_size[2], -0.5f);
  const float tu = __fadd_rd(u, M), tv = __fadd_rd(v, M), tw = __fadd_rd(w, M);
  const float a = u - (tu - M), b = v - (tv - M), g = w - (tw - M);
  const int mx = (int)pc.max_brick[0], my = (int)pc.max_brick[1], mz = (int)pc.max_brick[2];
  const float scale = (BPV == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
  Cross k;
  if (SAMP == TVK_SAMPLER_GATHER) {
    // the 4x4 block (x-1..x+2, y-1..y+2) of planes z and z+1 = four diagonal gathers each; planes z-1 and z+2
    // need the centre 2x2 only: 10 gathers.  x / y neighbours are not clamped to the slot: a position the march
    // can produce keeps its taps inside the brick's ghost layers (see fetch_taps).
    const float gx = (tu - M) + sl.gdx, gy = (tv - M) + sl.gdy;
    const int zs = min(max(__float_as_int(tw) - sl.oz, 0), mz - 2);          // in-slot plane (the z-1 / z+2 planes are clamped to the slot)
    const int lz0 = sl.ldz + sl.oz;                                           // first layer of the slot
    const uint32_t l = (uint32_t)(lz0 + zs);
    const uint4 mm0 = gather4(rp.tex, l, gx - 1.0f, gy - 1.0f), pm0 = gather4(rp.tex, l, gx + 1.0f, gy - 1.0f);
    const uint4 mp0 = gather4(rp.tex, l, gx - 1.0f, gy + 1.0f), pp0 = gather4(rp.tex, l, gx + 1.0f, gy + 1.0f);
    const uint4 mm1 = gather4(rp.tex, l + 1u, gx - 1.0f, gy - 1.0f), pm1 = gather4(rp.tex, l + 1u, gx + 1.0f, gy - 1.0f);
    const uint4 mp1 = gather4(rp.tex, l + 1u, gx - 1.0f, gy + 1.0f), pp1 = gather4(rp.tex, l + 1u, gx + 1.0f, gy + 1.0f);
    const uint4 zm = gather4(rp.tex, (uint32_t)(lz0 + max(zs - 1, 0)), gx, gy), zp = gather4(rp.tex, (uint32_t)(lz0 + min(zs + 2, mz - 1)), gx, gy);
    // gather(gx-1, gy-1): .w (x-1,y-1) .z (x,y-1) .x (x-1,y) .y (x,y);   gather(gx+1, gy-1): .w (x+1,y-1) .z (x+2,y-1) .x (x+1,y) .y (x+2,y)
    // gather(gx-1, gy+1): .w (x-1,y+1) .z (x,y+1) .x (x-1,y+2) .y (x,y+2); gather(gx+1, gy+1): .w (x+1,y+1) .z (x+2,y+1) .x (x+1,y+2) .y (x+2,y+2)
    k.X0[0] = u2f(mm0.x); k.X0[1] = u2f(mp0.w); k.X0[2] = u2f(mm1.x); k.X0[3] = u2f(mp1.w);
    k.X1[0] = u2f(mm0.y); k.X1[1] = u2f(mp0.z); k.X1[2] = u2f(mm1.y); k.X1[3] = u2f(mp1.z);
    k.X2[0] = u2f(pm0.x); k.X2[1] = u2f(pp0.w); k.X2[2] = u2f(pm1.x); k.X2[3] = u2f(pp1.w);
    k.X3[0] = u2f(pm0.y); k.X3[1] = u2f(pp0.z); k.X3[2] = u2f(pm1.y); k.X3[3] = u2f(pp1.z);
    k.ym[0] = u2f(mm0.z); k.ym[1] = u2f(pm0.w); k.ym[2] = u2f(mm1.z); k.ym[3] = u2f(pm1.w);
    k.yp[0] = u2f(mp0.y); k.yp[1] = u2f(pp0.x); k.yp[2] = u2f(mp1.y); k.yp[3] = u2f(pp1.x);
    k.zm[0] = u2f(zm.w); k.zm[1] = u2f(zm.z); k.zm[2] = u2f(zm.x); k.zm[3] = u2f(zm.y);
    k.zp[0] = u2f(zp.w); k.zp[1] = u2f(zp.z); k.zp[2] = u2f(zp.x); k.zp[3] = u2f(zp.y);
  } else {
    // (y,z)-quad pool: the cross arrives in 12 quad loads around one base address
    const int x = min(max(__float_as_int(tu) - sl.ox, 0), mx - 2);
    const int y = min(max(__float_as_int(tv) - sl.oy, 0), my - 2);
    const int z = min(max(__float_as_int(tw) - sl.oz, 0), mz - 2);
    const uint32_t sx = pc.pool_dim[0], sxy = sx * pc.pool_dim[1];
    const uint32_t o = (uint32_t)x + (uint32_t)y * sx + (uint32_t)z * sxy;
    const uint32_t oxm = o - (x > 0 ? 1u : 0u), oxp = o + (x + 2 < mx ? 2u : 1u);         // columns x-1 and x+2 (clamped)
    const uint32_t dym = y > 0 ? sx : 0u, dyp = sx, dzm = z > 0 ? sxy : 0u, dzp = sxy;     // (y+1, z+1 <= dim-1 always)
    float Ym0[4], Ym1[4], Yp0[4], Yp1[4], Zm0[4], Zm1[4], Zp0[4], Zp1[4];
    unpack_quad<BPV>(sl.base, oxm, k.X0); unpack_quad<BPV>(sl.base, o, k.X1); unpack_quad<BPV>(sl.base, o + 1, k.X2); unpack_quad<BPV>(sl.base, oxp, k.X3);
    unpack_quad<BPV>(sl.base, o - dym, Ym0); unpack_quad<BPV>(sl.base, o + 1 - dym, Ym1);
    unpack_quad<BPV>(sl.base, o + dyp, Yp0); unpack_quad<BPV>(sl.base, o + 1 + dyp, Yp1);
    unpack_quad<BPV>(sl.base, o - dzm, Zm0); unpack_quad<BPV>(sl.base, o + 1 - dzm, Zm1);
    unpack_quad<BPV>(sl.base, o + dzp, Zp0); unpack_quad<BPV>(sl.base, o + 1 + dzp, Zp1);
    k.ym[0] = y > 0 ? Ym0[0] : k.X1[0]; k.ym[1] = y > 0 ? Ym1[0] : k.X2[0]; k.ym[2] = y > 0 ? Ym0[2] : k.X1[2]; k.ym[3] = y > 0 ? Ym1[2] : k.X2[2];
    k.yp[0] = Yp0[1]; k.yp[1] = Yp1[1]; k.yp[2] = Yp0[3]; k.yp[3] = Yp1[3];
    k.zm[0] = z > 0 ? Zm0[0] : k.X1[0]; k.zm[1] = z > 0 ? Zm1[0] : k.X2[0]; k.zm[2] = z > 0 ? Zm0[1] : k.X1[1]; k.zm[3] = z > 0 ? Zm1[1] : k.X2[1];
    k.zp[0] = Zp0[2]; k.zp[1] = Zp1[2]; k.zp[2] = Zp0[3]; k.zp[3] = Zp1[3];
  }
  cross_filter(k, a, b, g, scale, data, grad);
}

__device__ __forceinline__ float pow8(float x) { x *= x; x *= x; return x * x; }

// Lighting, lighting.glsl:33-43
__device__ __forceinline__ f3 lighting(const tvk_raycast_params& p, f3 pos, f3 n, f3 diffuse) {
  // (sort-last: box-local model position -> the global volume's; identity when the fields are zero: fmaf(x, 1, 0) == x)
  pos = mk3(fmaf(pos.x, 1.0f + p.model_pos_scale_m1[0], p.model_pos_offset[0]), fmaf(pos.y, 1.0f + p.model_pos_scale_m1[1], p.model_pos_offset[1]),
            fmaf(pos.z, 1.0f + p.model_pos_scale_m1[2], p.model_pos_offset[2]));
  f3 v = mk3(p.model_eye_pos[0] - pos.x, p.model_eye_pos[1] - pos.y, p.model_eye_pos[2] - pos.z);
  float il = rsqrtf(dot3(v, v));
  v = mk3(v.x * il, v.y * il, v.z * il);
  const float d2 = 2.0f * dot3(n, v);
  f3 r = mk3(fmaf(-d2, n.x, v.x), fmaf(-d2, n.y, v.y), fmaf(-d2, n.z, v.z));
  il = rsqrtf(dot3(r, r));
  r = mk3(r.x * il, r.y * il, r.z * il);
  const f3 L = mk3(p.model_light_dir[0], p.model_light_dir[1], p.model_light_dir[2]);
  const float ndl = fmaxf(fabsf(dot3(n, L)), 0.0f);
  const float spec = pow8(fmaxf(dot3(r, L), 0.0f));
  return mk3(clamp01(fmaf(p.light_specular[0], spec, fmaf(diffuse.x, ndl, p.light_ambient[0]))),
             clamp01(fmaf(p.light_specular[1], spec, fmaf(diffuse.y, ndl, p.light_ambient[1]))),
             clamp01(fmaf(p.light_specular[2], spec, fmaf(diffuse.z, ndl, p.light_ambient[2]))));
}

// ComputeColorFromVolume, GLGridLeaper-Method-{1D,1D-L,2D,2D-L}.glsl; `data` is the
// already filtered centre sample (fetched one iteration ahead by the march loop).
template <int MODE, bool LIT, int BPV, int SAMP>
__device__ __forceinline__ float4 color_from_volume(const RayParams& rp, const SlotCtx& sl, float data, f3 pc, f3 scale /* 1 / brick scale */, f3 trans,
                                                     const f3* grad = nullptr) {
  const tvk_raycast_params& p = rp.p;
  float4 col;
  f3 n;
  if (MODE == TVK_RM_1DTRANS) {
    col = __ldg(rp.tf_f + nearest_index(data * p.trans_scale, rp.tf_w));
    if (LIT && col.w != 0.0f) {                       // (see below: transparent samples are not shaded)
      const f3 g = grad ? *grad : gradient<BPV, SAMP>(rp, sl, pc);
      n = mk3(g.x * p.domain_scale[0], g.y * p.domain_scale[1], g.z * p.domain_scale[2]);
      const float d = dot3(n, n);
      if (d > 0.0f) { const float il = rsqrtf(d); n = mk3(n.x * il, n.y * il, n.z * il); }
    }
  } else {
    const int cx = nearest_index(data * p.trans_scale, rp.tf_w);
    // a table column that is transparent in every gradient row: no gradient needed to know the sample adds nothing
    if (!grad && !((__ldg(rp.tf_colmask + (cx >> 5)) >> (cx & 31)) & 1u)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const f3 g = grad ? *grad : gradient<BPV, SAMP>(rp, sl, pc);
    const float gm = len3(g);
    col = tf_fetch(rp, nearest_index(fmaf(-gm, p.gradient_scale, 1.0f), rp.tf_h) * (int)rp.tf_w + cx);
    if (LIT) {
      const float ig = 1.0f / gm;
      const f3 gn = (gm > 0.0f) ? mk3(g.x * ig, g.y * ig, g.z * ig) : g;
      n = mk3(p.domain_scale[0] * gn.x, p.domain_scale[1] * gn.y, p.domain_scale[2] * gn.z);
    }
  }
  // a sample the transfer function makes fully transparent adds (1-a_acc)*0*rgb = 0 to the ray whatever its shading
  // (rgb is clamped to [0,1]): the lighting is skipped for it
  if (LIT && col.w != 0.0f) {
    const f3 mp = mk3((pc.x - trans.x) * scale.x, (pc.y - trans.y) * scale.y, (pc.z - trans.z) * scale.z);   // scale = 1 / brick scale here
    const f3 lit = lighting(p, mp, n, mk3(col.x * p.light_diffuse[0], col.y * p.light_diffuse[1], col.z * p.light_diffuse[2]));
    col.x = lit.x; col.y = lit.y; col.z = lit.z;
  }
  return col;
}

// Analytic ray set-up (pixel-centre ray against the unit box in normalised volume
// coordinates); same arithmetic as oracle ray_setup().
__device__ __forceinline__ bool ray_setup(const tvk_raycast_params& p, uint32_t px, uint32_t py, float4& entry, float4& exitp) {
  const f3 ndc = mk3(fmaf(((float)px + 0.5f) / (float)p.width, 2.0f, -1.0f), fmaf(((float)py + 0.5f) / (float)p.height, 2.0f, -1.0f), -1.0f);
  const float4 n4 = xform(p.inv_projection, ndc, 1.0f);
  const f3 n = mk3(n4.x / n4.w, n4.y / n4.w, n4.z / n4.w);
  const float4 O = xform(p.eye_to_model, mk3(0.f, 0.f, 0.f), 1.0f);
  const float4 D = xform(p.eye_to_model, n, 0.0f);
  float s0 = -INFINITY, s1 = INFINITY;
  const float o[3] = {O.x, O.y, O.z}, d[3] = {D.x, D.y, D.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (d[a] == 0.0f) {
      if (o[a] < p.box_min[a] || o[a] > p.box_max[a]) return false;
      continue;
    }
    const float inv = 1.0f / d[a];
    float ta = (p.box_min[a] - o[a]) * inv, tb = (p.box_max[a] - o[a]) * inv;
    if (ta > tb) { const float t = ta; ta = tb; tb = t; }
    s0 = fmaxf(s0, ta);
    s1 = fminf(s1, tb);
  }
  if (p.clip_enabled) {
    const f3 cp = mk3(p.clip_plane[0], p.clip_plane[1], p.clip_plane[2]);
    const float dn = dot3(cp, mk3(D.x, D.y, D.z));
    const float on = dot3(cp, mk3(O.x, O.y, O.z)) + p.clip_plane[3];
    if (dn == 0.0f) { if (on > 0.0f) return false; }
    else {
      const float t = -on / dn;
      if (dn > 0.0f) s1 = fminf(s1, t); else s0 = fmaxf(s0, t);
    }
  }
  const float se = fmaxf(s0, 1.0f);
  if (!(s1 > se)) return false;
  const f3 pe = mk3(n.x * se, n.y * se, n.z * se), px1 = mk3(n.x * s1, n.y * s1, n.z * s1);
  float4 t4 = xform(p.eye_to_model, pe, 1.0f);
  entry = make_float4(t4.x, t4.y, t4.z, pe.z);
  t4 = xform(p.eye_to_model, px1, 1.0f);
  exitp = make_float4(t4.x, t4.y, t4.z, px1.z);
  return true;
}

// per-ray counters (a ray takes <= 100 bricks); the kernel folds them into 64-bit totals after each ray
struct Counters { uint32_t samples, bricks; };

// GLGridLeaper-blend.glsl main()
template <int MODE, bool LIT, int BPV, int SAMP, bool SRM1>
__device__ __forceinline__ void march_ray(const RayParams& rp, float4 exitp, float4& pos, float4& col, float4& acc, Counters& cn) {
  const tvk_raycast_params& p = rp.p;
  const PoolConst& pc = rp.pc;
  acc = col;
  const f3 entry = mk3(pos.x, pos.y, pos.z);
  const float entry_depth = pos.w, exit_depth = exitp.w;
  const f3 dir = mk3(exitp.x - entry.x, exitp.y - entry.y, exitp.z - entry.z);
  const float ray_len = len3(dir);
  f3 vdir = mk3(dir.x * pc.volume_size[0], dir.y * pc.volume_size[1], dir.z * pc.volume_size[2]);
  const float vl = len3(vdir);
  const float den = 2.0f * p.sample_rate_modifier;
  vdir = mk3(vdir.x / vl / pc.pool_size[0] / den, vdir.y / vl / pc.pool_size[1] / den, vdir.z / vl / pc.pool_size[2] / den);
  const float step_size = len3(vdir);
  float t = 0.0f;
  bool optimal = true;
  const float voxel_size = 0.125f / 2000.0f;
  f3 cur = entry;
  uint4 last_bc = make_uint4(0, 0, 0, 9999);
  const float oc = 1.0f / p.sample_rate_modifier;
  if (ray_len > voxel_size) {
    for (uint32_t j = 0; j < 100; ++j) {
      const float depth = fmaf(exit_depth, t, entry_depth * (1.0f - t));
      uint32_t lod = compute_lod(depth, p.lod_factor, p.level_zero_wse, pc.lod_count - 1);
      BrickHit b;
      get_brick<SAMP>(rp, cur, lod, dir, b);
      cn.bricks++;
      if (!b.found && optimal) {
        optimal = false;
        pos = make_float4(cur.x, cur.y, cur.z, depth);
        col = acc;
      }
      if (!b.empty && (last_bc.x != b.bc.x || last_bc.y != b.bc.y || last_bc.z != b.bc.z || last_bc.w != b.bc.w)) {
        int steps = (int)ceilf(len3(mk3(b.pool_exit.x - b.pool_entry.x, b.pool_exit.y - b.pool_entry.y, b.pool_exit.z - b.pool_entry.z)) / step_size);
        const int s2 = (int)ceilf(len3(mk3((exitp.x - cur.x) * b.scale.x, (exitp.y - cur.y) * b.scale.y, (exitp.z - cur.z) * b.scale.z)) / step_size);
        steps = min(steps, s2);
        f3 q = b.pool_entry;
        // NB samples per trip: their footprints (and then their TF fetches) are in flight together, which
        // multiplies the memory-level parallelism of a ray without any register rotation.  Samples are
        // still classified, counted and composited strictly in order, with the ERT test after each.
        // The hardware-filtered path needs one register per sample in flight, so it batches deeper.
        const int NB = ((MODE == TVK_RM_1DTRANS) && !LIT) ? (SAMP == TVK_SAMPLER_TEX3D ? TVK_NB_TEX : (SAMP == TVK_SAMPLER_GATHER ? TVK_NB_GATHER : TVK_NB_LDG)) : 1;
        int i = 0;
#define TVK_COMPOSITE(S)                                                                     \
        {                                                                                    \
          cn.samples++;                                                                      \
          if (!SRM1) (S).w = (oc == 1.0f) ? 1.0f - (1.0f - (S).w) : 1.0f - powf(1.0f - (S).w, oc);   /* SRM1: the table holds 1-(1-a) */ \
          const float w_ = (1.0f - acc.w) * (S).w;    /* UnderCompositing, Compositing.glsl:33-38 */ \
          acc.x = fmaf((S).x, w_, acc.x); acc.y = fmaf((S).y, w_, acc.y); acc.z = fmaf((S).z, w_, acc.z); \
          acc.w = fmaf((S).w, 1.0f - acc.w, acc.w);                                          \
          if (acc.w > 0.99f) {                                                               \
            if (optimal) { pos.w = 1000.0f; col = acc; }                                     \
            return;                                                                          \
          }                                                                                  \
        }
        if (NB > 1) {
          for (; i + NB <= steps; i += NB) {
            f3 qs[NB];
            float d[NB];
            qs[0] = q;
#pragma unroll
            for (int k = 1; k < NB; ++k) qs[k] = mk3(qs[k - 1].x + vdir.x, qs[k - 1].y + vdir.y, qs[k - 1].z + vdir.z);
            if (SAMP == TVK_SAMPLER_TEX3D) {
#pragma unroll
              for (int k = 0; k < NB; ++k) d[k] = tex3D<float>(rp.tex, qs[k].x, qs[k].y, qs[k].z);
            } else {
              Taps tp[NB];
#pragma unroll
              for (int k = 0; k < NB; ++k) fetch_taps<BPV, SAMP>(rp, b.slot, qs[k], tp[k]);
#pragma unroll
              for (int k = 0; k < NB; ++k) d[k] = filter_taps<BPV>(tp[k]);
            }
            float4 sc[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) sc[k] = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, d[k], qs[k], b.inv_scale, b.trans);
#pragma unroll
            for (int k = 0; k < NB; ++k) TVK_COMPOSITE(sc[k])
            q = mk3(qs[NB - 1].x + vdir.x, qs[NB - 1].y + vdir.y, qs[NB - 1].z + vdir.z);
          }
        }
        if ((SAMP == TVK_SAMPLER_EXACT_Q4 || SAMP == TVK_SAMPLER_GATHER) && TVK_NB_LIT > 1 && ((LIT && MODE == TVK_RM_1DTRANS) || MODE == TVK_RM_2DTRANS)) {
          // Lit and 2D-TF modes, pipelined: the CENTRE samples of NL consecutive steps (two fetches each) are issued
          // together, then their table look-ups (1D: the entry itself, 2D: the column's "some row is visible" bit), and
          // only then -- per sample, in order -- the gradient footprint and the shading of the samples that are not
          // transparent.  Same arithmetic per sample as the scalar loop below, same order of compositing, early ray
          // termination after each sample (later samples of the group were fetched speculatively and are dropped).
          const int NL = TVK_NB_LIT;
          for (; i + NL <= steps; i += NL) {
            f3 qs[NL];
            float d[NL];
            qs[0] = q;
#pragma unroll
            for (int k = 1; k < NL; ++k) qs[k] = mk3(qs[k - 1].x + vdir.x, qs[k - 1].y + vdir.y, qs[k - 1].z + vdir.z);
            {
              Taps tp[NL];
#pragma unroll
              for (int k = 0; k < NL; ++k) fetch_taps<BPV, SAMP>(rp, b.slot, qs[k], tp[k]);
#pragma unroll
              for (int k = 0; k < NL; ++k) d[k] = filter_taps<BPV>(tp[k]);
            }
            float4 sc[NL];
            bool vis[NL];
#pragma unroll
            for (int k = 0; k < NL; ++k) {
              const int cx = nearest_index(d[k] * p.trans_scale, rp.tf_w);
              if (MODE == TVK_RM_1DTRANS) { sc[k] = __ldg(rp.tf_f + cx); vis[k] = sc[k].w != 0.0f; }
              else { sc[k] = make_float4(0.f, 0.f, 0.f, 0.f); vis[k] = ((__ldg(rp.tf_colmask + (cx >> 5)) >> (cx & 31)) & 1u) != 0u; }
            }
#pragma unroll
            for (int k = 0; k < NL; ++k) {
              if (vis[k]) {
                float dc;
                f3 g;
                center_gradient<BPV, SAMP>(rp, b.slot, qs[k], dc, g);
                sc[k] = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, d[k], qs[k], b.inv_scale, b.trans, &g);
              }
              TVK_COMPOSITE(sc[k])
            }
            q = mk3(qs[NL - 1].x + vdir.x, qs[NL - 1].y + vdir.y, qs[NL - 1].z + vdir.z);
          }
        }
        for (; i < steps; ++i) {
          float4 s0;
          if ((SAMP == TVK_SAMPLER_EXACT_Q4 || SAMP == TVK_SAMPLER_GATHER) && LIT && MODE == TVK_RM_1DTRANS) {
            // 1D-TF lit: the table needs the centre sample only (2 quad loads); where it returns alpha 0 -- empty space
            // inside visible bricks -- the gradient (10 more quad loads) and the shading are not needed either
            const float data = sample_pool<BPV, SAMP>(rp, b.slot, q);
            s0 = __ldg(rp.tf_f + nearest_index(data * p.trans_scale, rp.tf_w));
            if (s0.w != 0.0f) {
              float dc;
              f3 g;
              center_gradient<BPV, SAMP>(rp, b.slot, q, dc, g);
              s0 = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, data, q, b.inv_scale, b.trans, &g);
            }
          } else if ((SAMP == TVK_SAMPLER_EXACT_Q4 || SAMP == TVK_SAMPLER_GATHER) && MODE == TVK_RM_2DTRANS) {
            // 2D-TF: the table row depends on the gradient, the column on the centre sample alone; a column that is
            // transparent in every row makes the sample transparent whatever its gradient
            const float data = sample_pool<BPV, SAMP>(rp, b.slot, q);
            const int cx = nearest_index(data * p.trans_scale, rp.tf_w);
            s0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((__ldg(rp.tf_colmask + (cx >> 5)) >> (cx & 31)) & 1u) {
              float dc;
              f3 g;
              center_gradient<BPV, SAMP>(rp, b.slot, q, dc, g);
              s0 = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, data, q, b.inv_scale, b.trans, &g);
            }
          } else {
            const float data = sample_pool<BPV, SAMP>(rp, b.slot, q);
            s0 = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, data, q, b.inv_scale, b.trans);
          }
          TVK_COMPOSITE(s0)
          q = mk3(q.x + vdir.x, q.y + vdir.y, q.z + vdir.z);
        }
#undef TVK_COMPOSITE
        cur = mk3((q.x - b.trans.x) / b.scale.x, (q.y - b.trans.y) / b.scale.y, (q.z - b.trans.z) / b.scale.z);
      } else {
        cur = mk3(b.norm_exit.x + voxel_size * dir.x / ray_len, b.norm_exit.y + voxel_size * dir.y / ray_len,
                  b.norm_exit.z + voxel_size * dir.z / ray_len);
      }
      last_bc = b.bc;
      t = len3(mk3(entry.x - b.norm_exit.x, entry.y - b.norm_exit.y, entry.z - b.norm_exit.z)) / ray_len;
      if (t > 0.9999f) break;
    }
  }
  if (optimal) { pos.w = 1000.0f; col = acc; }
}

// GLGridLeaper-iso.glsl main() + GLGridLeaper-Method-iso.glsl
template <int BPV, int SAMP>
__device__ __forceinline__ void march_ray_iso(const RayParams& rp, float4 exitp, float4& pos, float4& nrm, float4& hit, float4& hitn, Counters& cn) {
  const tvk_raycast_params& p = rp.p;
  const PoolConst& pc = rp.pc;
  hit = make_float4(0.f, 0.f, 0.f, 0.f);
  hitn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (floorf(pos.w) == 1000.0f) return;
  if (floorf(pos.w) == 500.0f) {
    const float4 h4 = xform(p.model_to_eye, mk3(pos.x, pos.y, pos.z), 1.0f);
    hit = make_float4(h4.x, h4.y, h4.z, pos.w - floorf(pos.w) + 1.0f);
    hitn = nrm;
    return;
  }
  const f3 entry = mk3(pos.x, pos.y, pos.z);
  const float entry_depth = pos.w, exit_depth = exitp.w;
  const f3 dir = mk3(exitp.x - entry.x, exitp.y - entry.y, exitp.z - entry.z);
  const float ray_len = len3(dir);
  f3 vdir = mk3(dir.x * pc.volume_size[0], dir.y * pc.volume_size[1], dir.z * pc.volume_size[2]);
  const float vl = len3(vdir);
  const float den = 2.0f * p.sample_rate_modifier;
  vdir = mk3(vdir.x / vl / pc.pool_size[0] / den, vdir.y / vl / pc.pool_size[1] / den, vdir.z / vl / pc.pool_size[2] / den);
  const float step_size = len3(vdir);
  float t = 0.0f;
  bool optimal = true;
  const float voxel_size = 0.125f / 2000.0f;
  f3 cur = entry;
  uint4 last_bc = make_uint4(0, 0, 0, 9999);
  if (ray_len > voxel_size) {
    for (uint32_t j = 0; j < 100; ++j) {
      const float depth = fmaf(exit_depth, t, entry_depth * (1.0f - t));
      uint32_t lod = compute_lod(depth, p.lod_factor, p.level_zero_wse, pc.lod_count - 1);
      BrickHit b;
      get_brick<SAMP>(rp, cur, lod, dir, b);
      cn.bricks++;
      if (!b.found && optimal) {
        optimal = false;
        pos = make_float4(cur.x, cur.y, cur.z, depth);
      }
      if (!b.empty && (last_bc.x != b.bc.x || last_bc.y != b.bc.y || last_bc.z != b.bc.z || last_bc.w != b.bc.w)) {
        int steps = (int)ceilf(len3(mk3(b.pool_exit.x - b.pool_entry.x, b.pool_exit.y - b.pool_entry.y, b.pool_exit.z - b.pool_entry.z)) / step_size);
        const int s2 = (int)ceilf(len3(mk3((exitp.x - cur.x) * b.scale.x, (exitp.y - cur.y) * b.scale.y, (exitp.z - cur.z) * b.scale.z)) / step_size);
        steps = min(steps, s2);
        f3 q = b.pool_entry;
        for (int i = 0; i < steps; ++i) {
          cn.samples++;
          if (sample_pool<BPV, SAMP>(rp, b.slot, q) >= p.isoval) {
            f3 rd = mk3(vdir.x / 2.0f, vdir.y / 2.0f, vdir.z / 2.0f);       // RefineIsosurface
            q = mk3(q.x - rd.x, q.y - rd.y, q.z - rd.z);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              rd = mk3(rd.x / 2.0f, rd.y / 2.0f, rd.z / 2.0f);
              if (sample_pool<BPV, SAMP>(rp, b.slot, q) >= p.isoval) q = mk3(q.x - rd.x, q.y - rd.y, q.z - rd.z);
              else q = mk3(q.x + rd.x, q.y + rd.y, q.z + rd.z);
            }
            cur = mk3((q.x - b.trans.x) / b.scale.x, (q.y - b.trans.y) / b.scale.y, (q.z - b.trans.z) / b.scale.z);
            const float4 h4 = xform(p.model_to_eye, cur, 1.0f);
            hit = make_float4(h4.x, h4.y, h4.z, 2.0f);                       // color.r + 1
            const f3 g = gradient<BPV, SAMP>(rp, b.slot, q);
            f3 n = mk3(g.x * p.domain_scale[0], g.y * p.domain_scale[1], g.z * p.domain_scale[2]);
            const float l = len3(n);
            if (l > 0.0f) n = mk3(n.x / l, n.y / l, n.z / l);
            const float4 n4 = xform(p.model_view_it, n, 0.0f);
            hitn = make_float4(n4.x, n4.y, n4.z, 513.0f);                    // floor(color.g*512)+color.b
            if (optimal) pos.w = 499.0f + hit.w;                             // TerminateRay; xyz untouched
            nrm = hitn;
            return;
          }
          q = mk3(q.x + vdir.x, q.y + vdir.y, q.z + vdir.z);
        }
        cur = mk3((q.x - b.trans.x) / b.scale.x, (q.y - b.trans.y) / b.scale.y, (q.z - b.trans.z) / b.scale.z);
      } else {
        cur = mk3(b.norm_exit.x + voxel_size * dir.x / ray_len, b.norm_exit.y + voxel_size * dir.y / ray_len,
                  b.norm_exit.z + voxel_size * dir.z / ray_len);
      }
      last_bc = b.bc;
      t = len3(mk3(entry.x - b.norm_exit.x, entry.y - b.norm_exit.y, entry.z - b.norm_exit.z)) / ray_len;
      if (t > 0.9999f) break;
    }
  }
  if (optimal) pos.w = 1000.0f;
  nrm = hitn;
}

// Compose-FS.glsl:47-75 with the uniforms of GLRenderer::ComposeSurfaceImage (:2763-2811)
__device__ __forceinline__ float4 compose_iso(const tvk_raycast_params& p, float4 hit, float4 hitn) {
  if (hit.w == 0.0f) return make_float4(0.f, 0.f, 0.f, 0.f);
  const f3 n = mk3(hitn.x, hitn.y, fabsf(hitn.z));
  f3 v = mk3(0.0f - hit.x, 0.0f - hit.y, 0.0f - hit.z);
  float l = len3(v);
  v = mk3(v.x / l, v.y / l, v.z / l);
  const float d2 = 2.0f * dot3(n, v);
  f3 r = mk3(fmaf(-d2, n.x, v.x), fmaf(-d2, n.y, v.y), fmaf(-d2, n.z, v.z));
  l = len3(r);
  r = mk3(r.x / l, r.y / l, r.z / l);
  const f3 L = mk3(p.compose_light_dir[0], p.compose_light_dir[1], p.compose_light_dir[2]);
  const float ndl = fmaxf(fabsf(dot3(n, mk3(-L.x, -L.y, -L.z))), 0.0f);
  const float spec = pow8(fmaxf(dot3(r, L), 0.0f));
  return make_float4(clamp01(fmaf(p.light_specular[0], spec, fmaf(p.light_diffuse[0] * p.iso_color[0], ndl, p.light_ambient[0]))),
                     clamp01(fmaf(p.light_specular[1], spec, fmaf(p.light_diffuse[1] * p.iso_color[1], ndl, p.light_ambient[1]))),
                     clamp01(fmaf(p.light_specular[2], spec, fmaf(p.light_diffuse[2] * p.iso_color[2], ndl, p.light_ambient[2]))), 1.0f);
}

// Result store.  With the sort-last exchange active the pixel also goes, by a peer store over NVLink, into the
// receive buffer of the rank that owns its scan-line stripe -- while the warp moves on to its next tile.
__device__ __forceinline__ void store_pixel(const RayParams& rp, uint32_t px, uint32_t py, size_t i, float4 v) {
  rp.image[i] = v;
  if (rp.xs.world) {
    const uint32_t owner = py / rp.xs.rows;
    rp.xs.recv[owner][(size_t)(py - owner * rp.xs.rows) * rp.p.width + px] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Transfer-function modes: LANES REFILL THEIR OWN RAYS.
//
// In the tile kernel below a warp owns an 8x4 pixel tile until its LAST ray has finished; rays of one tile end at
// very different depths (early ray termination, silhouettes, bricks clipped at a corner), and ncu showed only 18-20
// of 32 lanes active in the sample loop.  Here every lane carries its own ray through a small state machine --
//   idle -> (fetch a pixel, set the ray up) -> need brick -> (GetBrick) -> sampling -> need brick -> ... -> finished -> idle
// -- and a lane whose ray has finished fetches the next pixel while the others go on.  Pixels are handed out from one
// global counter in tile-major order (32 consecutive numbers = one 8x4 tile), so the lanes of a warp stay on
// neighbouring pixels.  The two divergent phases (ray set-up, page-table walk) only run when at least
// TVK_REFILL_MIN / TVK_BRICK_MIN lanes wait for them (or nobody can sample), which keeps them from running with one
// or two lanes.  The arithmetic of a ray -- and so every pixel, request and counter -- is that of march_ray.
#ifndef TVK_REFILL_MIN
#define TVK_REFILL_MIN 8
#endif
#ifndef TVK_BRICK_MIN
#define TVK_BRICK_MIN 8
#endif

template <int MODE, bool LIT, int BPV, int SAMP, bool SRM1>
__global__ void __launch_bounds__(128, (MODE == TVK_RM_1DTRANS && !LIT) ? (SAMP == TVK_SAMPLER_TEX3D ? TVK_OCC_TEX : (SAMP == TVK_SAMPLER_GATHER ? TVK_OCC_GATHER : TVK_OCC_LDG)) : TVK_OCC_LIT) raycast_refill_kernel(const __grid_constant__ RayParams rp) {
  const tvk_raycast_params& p = rp.p;
  const PoolConst& pc = rp.pc;
  const uint32_t lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;
  const uint32_t total_slots = rp.tiles_x * rp.tiles_y * 32u;
  const float voxel_size = 0.125f / 2000.0f;
  const float oc = 1.0f / p.sample_rate_modifier;
  const float den = 2.0f * p.sample_rate_modifier;
  unsigned long long nrays = 0, nsamples = 0, nbricks = 0;

  // ---- per-lane ray state
  bool active = false, optimal = true;
  size_t pi = 0;
  uint32_t px = 0, py = 0, j = 0, cn_samples = 0, cn_bricks = 0;
  int steps_left = 0;
  float4 exitp = make_float4(0.f, 0.f, 0.f, 0.f), pos = exitp, col = exitp, acc = exitp;
  f3 entry = mk3(0.f, 0.f, 0.f), dir = entry, vdir = entry, cur = entry, q = entry;
  float ray_len = 1.0f, step_size = 1.0f, t = 0.0f, entry_depth = 0.0f;
  uint4 last_bc = make_uint4(0, 0, 0, 9999);
  BrickHit b;
  b.scale = b.trans = b.inv_scale = b.norm_exit = entry;
  b.bc = last_bc;
  b.slot.base = nullptr; b.slot.ox = b.slot.oy = b.slot.oz = 0; b.slot.gdx = b.slot.gdy = 0.f; b.slot.ldz = 0;
  bool exhausted = false;                         // warp-uniform: the pixel counter has run past the image

  // a finished ray writes its pixel and resume state (rayResumePos.w = 1000 when it was optimal all the way)
#define TVK_FINISH_RAY()                                                   \
  {                                                                        \
    if (optimal) { pos.w = 1000.0f; col = acc; }                           \
    store_pixel(rp, px, py, pi, acc);                                      \
    rp.pos[pi] = pos;                                                      \
    rp.aux[pi] = col;                                                      \
    nsamples += cn_samples; nbricks += cn_bricks;                          \
    active = false; steps_left = 0;                                        \
  }
  // what GLGridLeaper-blend.glsl does after a brick: new position, brick bookkeeping, progress along the ray
#define TVK_AFTER_BRICK()                                                                                  \
  {                                                                                                        \
    last_bc = b.bc;                                                                                        \
    t = len3(mk3(entry.x - b.norm_exit.x, entry.y - b.norm_exit.y, entry.z - b.norm_exit.z)) / ray_len;   \
    ++j;                                                                                                   \
    if (t > 0.9999f || j >= 100u) TVK_FINISH_RAY()                                                        \
  }

  for (;;) {
    // ------------------------------------------------------------------ refill: idle lanes fetch pixels
    const unsigned idle = __ballot_sync(FULL, !active);
    if (!exhausted && (idle == FULL || __popc(idle) >= TVK_REFILL_MIN)) {
      const int n = __popc(idle), leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(rp.tile_counter, (uint32_t)n);
      base = __shfl_sync(FULL, base, leader);
      if (base + (uint32_t)n >= total_slots) exhausted = true;
      if (!active) {
        const uint32_t slot = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if (slot < total_slots) {
          const uint32_t tile = slot >> 5, l = slot & 31u;
          px = (rp.tile_x0 + tile % rp.tiles_x) * 8u + (l & 7u); py = (rp.tile_y0 + tile / rp.tiles_x) * 4u + (l >> 3);
          if (px < p.width && py < p.height) {
            pi = (size_t)py * p.width + px;
            float4 e4;
            if (!ray_setup(p, px, py, e4, exitp)) {
              if (p.new_view) {
                store_pixel(rp, px, py, pi, make_float4(0.f, 0.f, 0.f, 0.f));
                rp.aux[pi] = make_float4(0.f, 0.f, 0.f, 0.f);
                rp.pos[pi] = make_float4(0.f, 0.f, 0.f, 1000.0f);
              } else if (rp.xs.world) {
                store_pixel(rp, px, py, pi, rp.image[pi]);       // the stripe owner needs every pixel of every frame
              }
            } else {
              bool go = true;
              if (p.new_view) { pos = e4; col = make_float4(0.f, 0.f, 0.f, 0.f); }
              else {
                pos = rp.pos[pi];
                if (pos.w == 1000.0f) {                           // finished: image already holds rayResumeColor
                  go = false;
                  if (rp.xs.world) store_pixel(rp, px, py, pi, rp.image[pi]);
                } else col = rp.aux[pi];
              }
              nrays++;
              if (go) {
                // GLGridLeaper-blend.glsl main(), the part before the brick loop
                acc = col;
                entry = mk3(pos.x, pos.y, pos.z);
                entry_depth = pos.w;
                dir = mk3(exitp.x - entry.x, exitp.y - entry.y, exitp.z - entry.z);
                ray_len = len3(dir);
                vdir = mk3(dir.x * pc.volume_size[0], dir.y * pc.volume_size[1], dir.z * pc.volume_size[2]);
                const float vl = len3(vdir);
                vdir = mk3(vdir.x / vl / pc.pool_size[0] / den, vdir.y / vl / pc.pool_size[1] / den, vdir.z / vl / pc.pool_size[2] / den);
                step_size = len3(vdir);
                t = 0.0f; optimal = true; cur = entry; last_bc = make_uint4(0, 0, 0, 9999); j = 0; steps_left = 0;
                cn_samples = 0; cn_bricks = 0;
                active = true;
                if (!(ray_len > voxel_size)) TVK_FINISH_RAY()
              }
            }
          }
        }
      }
    }
    const unsigned act = __ballot_sync(FULL, active);
    if (act == 0u) { if (exhausted) break; continue; }

    // ------------------------------------------------------------------ brick phase: GetBrick for the lanes that need one
    const bool need_brick = active && steps_left == 0;
    const unsigned nb = __ballot_sync(FULL, need_brick);
    const unsigned can_sample = act & ~nb;
    if (nb != 0u && (can_sample == 0u || __popc(nb) >= TVK_BRICK_MIN)) {
      if (need_brick) {
        const float depth = fmaf(exitp.w, t, entry_depth * (1.0f - t));
        uint32_t lod = compute_lod(depth, p.lod_factor, p.level_zero_wse, pc.lod_count - 1);
        get_brick<SAMP>(rp, cur, lod, dir, b);
        cn_bricks++;
        if (!b.found && optimal) {
          optimal = false;
          pos = make_float4(cur.x, cur.y, cur.z, depth);
          col = acc;
        }
        if (!b.empty && (last_bc.x != b.bc.x || last_bc.y != b.bc.y || last_bc.z != b.bc.z || last_bc.w != b.bc.w)) {
          int steps = (int)ceilf(len3(mk3(b.pool_exit.x - b.pool_entry.x, b.pool_exit.y - b.pool_entry.y, b.pool_exit.z - b.pool_entry.z)) / step_size);
          const int s2 = (int)ceilf(len3(mk3((exitp.x - cur.x) * b.scale.x, (exitp.y - cur.y) * b.scale.y, (exitp.z - cur.z) * b.scale.z)) / step_size);
          steps = min(steps, s2);
          q = b.pool_entry;
          steps_left = max(steps, 0);
          if (steps_left == 0) {
            cur = mk3((q.x - b.trans.x) / b.scale.x, (q.y - b.trans.y) / b.scale.y, (q.z - b.trans.z) / b.scale.z);
            TVK_AFTER_BRICK()
          }
        } else {
          cur = mk3(b.norm_exit.x + voxel_size * dir.x / ray_len, b.norm_exit.y + voxel_size * dir.y / ray_len,
                    b.norm_exit.z + voxel_size * dir.z / ray_len);
          TVK_AFTER_BRICK()
        }
      }
      continue;                                   // lanes may have finished: give the refill a look before sampling
    }

    // ------------------------------------------------------------------ sample phase: one group of samples per lane
    if (active && steps_left > 0) {
      bool ended = false;                         // early ray termination inside this group
#define TVK_COMPOSITE_L(S)                                                                   \
      {                                                                                      \
        cn_samples++;                                                                        \
        if (!SRM1) (S).w = (oc == 1.0f) ? 1.0f - (1.0f - (S).w) : 1.0f - powf(1.0f - (S).w, oc);   /* SRM1: the table holds 1-(1-a) */ \
        const float w_ = (1.0f - acc.w) * (S).w;    /* UnderCompositing, Compositing.glsl:33-38 */ \
        acc.x = fmaf((S).x, w_, acc.x); acc.y = fmaf((S).y, w_, acc.y); acc.z = fmaf((S).z, w_, acc.z); \
        acc.w = fmaf((S).w, 1.0f - acc.w, acc.w);                                            \
        if (acc.w > 0.99f) ended = true;                                                     \
      }
      const bool shaded = (SAMP == TVK_SAMPLER_EXACT_Q4 || SAMP == TVK_SAMPLER_GATHER) && ((LIT && MODE == TVK_RM_1DTRANS) || MODE == TVK_RM_2DTRANS);
      const int NG = shaded ? TVK_NB_LIT : (((MODE == TVK_RM_1DTRANS) && !LIT) ? (SAMP == TVK_SAMPLER_TEX3D ? TVK_NB_TEX : (SAMP == TVK_SAMPLER_GATHER ? TVK_NB_GATHER : TVK_NB_LDG)) : 1);
      f3 qs[NG + 1];
      float d[NG];
      qs[0] = q;
#pragma unroll
      for (int k = 1; k <= NG; ++k) qs[k] = mk3(qs[k - 1].x + vdir.x, qs[k - 1].y + vdir.y, qs[k - 1].z + vdir.z);
      // the footprints of the whole group are in flight together (positions past the brick's last step are fetched
      // speculatively -- addresses stay inside the slot -- and never composited)
      if (SAMP == TVK_SAMPLER_TEX3D) {
#pragma unroll
        for (int k = 0; k < NG; ++k) d[k] = tex3D<float>(rp.tex, qs[k].x, qs[k].y, qs[k].z);
      } else {
        Taps tp[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) fetch_taps<BPV, SAMP>(rp, b.slot, qs[k], tp[k]);
#pragma unroll
        for (int k = 0; k < NG; ++k) d[k] = filter_taps<BPV>(tp[k]);
      }
      float4 sc[NG];
      const int n_do = min(NG, steps_left);
      int done = 0;
      if (shaded) {
        bool vis[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          const int cx = nearest_index(d[k] * p.trans_scale, rp.tf_w);
          if (MODE == TVK_RM_1DTRANS) { sc[k] = __ldg(rp.tf_f + cx); vis[k] = sc[k].w != 0.0f; }
          else { sc[k] = make_float4(0.f, 0.f, 0.f, 0.f); vis[k] = ((__ldg(rp.tf_colmask + (cx >> 5)) >> (cx & 31)) & 1u) != 0u; }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          if (k < n_do && !ended) {
            if (vis[k]) {
              float dc;
              f3 g;
              center_gradient<BPV, SAMP>(rp, b.slot, qs[k], dc, g);
              sc[k] = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, d[k], qs[k], b.inv_scale, b.trans, &g);
            }
            TVK_COMPOSITE_L(sc[k])
            done = k + 1;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < NG; ++k) sc[k] = color_from_volume<MODE, LIT, BPV, SAMP>(rp, b.slot, d[k], qs[k], b.inv_scale, b.trans);
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          if (k < n_do && !ended) {
            TVK_COMPOSITE_L(sc[k])
            done = k + 1;
          }
        }
      }
#undef TVK_COMPOSITE_L
      if (ended) {
        TVK_FINISH_RAY()
      } else {
        // q after `done` samples: qs[done] (built by repeated addition, exactly like the scalar loop's q += step)
        f3 qn = qs[0];
#pragma unroll
        for (int k = 1; k <= NG; ++k) if (k == done) qn = qs[k];
        q = qn;
        steps_left -= done;
        if (steps_left == 0) {
          cur = mk3((q.x - b.trans.x) / b.scale.x, (q.y - b.trans.y) / b.scale.y, (q.z - b.trans.z) / b.scale.z);
          TVK_AFTER_BRICK()
        }
      }
    }
  }
#undef TVK_FINISH_RAY
#undef TVK_AFTER_BRICK
  if (p.count_samples) {
    for (int o = 16; o > 0; o >>= 1) {
      nsamples += __shfl_down_sync(FULL, nsamples, o);
      nbricks += __shfl_down_sync(FULL, nbricks, o);
      nrays += __shfl_down_sync(FULL, nrays, o);
    }
    if (lane == 0) {
      atomicAdd(rp.stats + 0, nsamples);
      atomicAdd(rp.stats + 1, nrays);
      atomicAdd(rp.stats + 2, nbricks);
    }
  }
  if (rp.xs.world) {
    // one system-scope fence per CTA (cumulative over the peer stores the barrier orders before it), see tvk_exchange.cu
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned prev = atomicAdd(rp.tile_counter + 1, 1u);
      if (prev == gridDim.x - 1) {
        rp.tile_counter[1] = 0u;
        // the rectangle this rank sent goes next to the flag (rects[parity*16 + rank] follow the 64 flag words)
        for (uint32_t o = 0; o < rp.xs.world; ++o) {
          volatile uint32_t* r = reinterpret_cast<volatile uint32_t*>(rp.xs.flag[o] - rp.xs.flag_index + 64u + 4u * rp.xs.flag_index);
          r[0] = rp.xs.rect[0]; r[1] = rp.xs.rect[1]; r[2] = rp.xs.rect[2]; r[3] = rp.xs.rect[3];
        }
        __threadfence_system();
        for (uint32_t o = 0; o < rp.xs.world; ++o) *reinterpret_cast<volatile uint32_t*>(rp.xs.flag[o]) = rp.xs.frame;
      }
    }
  }
}

// SRM1: sampleRateModifier == 1 known at compile time (the opacity correction is then the identity
// 1-(1-a) and the pow path is not even compiled in); instantiated for the 1D-TF kernels only.
template <int MODE, bool LIT, int BPV, int SAMP, bool SRM1>
__global__ void __launch_bounds__(128, (MODE == TVK_RM_1DTRANS && !LIT) ? (SAMP == TVK_SAMPLER_TEX3D ? TVK_OCC_TEX : (SAMP == TVK_SAMPLER_GATHER ? TVK_OCC_GATHER : TVK_OCC_LDG)) : TVK_OCC_LIT) raycast_kernel(const __grid_constant__ RayParams rp) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ntiles = rp.tiles_x * rp.tiles_y;
  Counters cn;
  cn.samples = 0; cn.bricks = 0;
  unsigned long long nrays = 0, nsamples = 0, nbricks = 0;
  for (;;) {
    nsamples += cn.samples; nbricks += cn.bricks;
    cn.samples = 0; cn.bricks = 0;
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(rp.tile_counter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= ntiles) break;
    const uint32_t px = (rp.tile_x0 + tile % rp.tiles_x) * 8u + (lane & 7u), py = (rp.tile_y0 + tile / rp.tiles_x) * 4u + (lane >> 3);
    if (px >= rp.p.width || py >= rp.p.height) continue;
    const size_t i = (size_t)py * rp.p.width + px;
    float4 entry, exitp;
    if (!ray_setup(rp.p, px, py, entry, exitp)) {
      if (rp.p.new_view) {
        store_pixel(rp, px, py, i, make_float4(0.f, 0.f, 0.f, 0.f));
        rp.aux[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        rp.pos[i] = make_float4(0.f, 0.f, 0.f, 1000.0f);
      } else if (rp.xs.world) {
        store_pixel(rp, px, py, i, rp.image[i]);      // sort-last: the stripe owner needs EVERY pixel of every frame
      }
      continue;
    }
    float4 pos, aux;
    if (rp.p.new_view) { pos = entry; aux = make_float4(0.f, 0.f, 0.f, 0.f); }
    else { pos = rp.pos[i]; }
    nrays++;
    if (MODE == TVK_RM_ISOSURFACE) {
      if (!rp.p.new_view) {
        if (floorf(pos.w) == 1000.0f) {              // finished without a hit: image stays 0
          if (rp.xs.world) store_pixel(rp, px, py, i, rp.image[i]);
          continue;
        }
        aux = rp.aux[i];
      }
      float4 hit, hitn;
      march_ray_iso<BPV, SAMP>(rp, exitp, pos, aux, hit, hitn, cn);
      if (rp.hit_buf) rp.hit_buf[i] = hit;           // (rays that found the surface march again in every sub-frame: the buffer is current)
      store_pixel(rp, px, py, i, compose_iso(rp.p, hit, hitn));
    } else {
      if (!rp.p.new_view) {
        if (pos.w == 1000.0f) {                      // finished: image already holds rayResumeColor
          if (rp.xs.world) store_pixel(rp, px, py, i, rp.image[i]);
          continue;
        }
        aux = rp.aux[i];
      }
      float4 acc;
      march_ray<MODE, LIT, BPV, SAMP, SRM1>(rp, exitp, pos, aux, acc, cn);
      store_pixel(rp, px, py, i, acc);
    }
    rp.pos[i] = pos;
    rp.aux[i] = aux;
  }
  if (rp.p.count_samples) {
    for (int o = 16; o > 0; o >>= 1) {
      nsamples += __shfl_down_sync(0xffffffffu, nsamples, o);
      nbricks += __shfl_down_sync(0xffffffffu, nbricks, o);
      nrays += __shfl_down_sync(0xffffffffu, nrays, o);
    }
    if (lane == 0) {
      atomicAdd(rp.stats + 0, nsamples);
      atomicAdd(rp.stats + 1, nrays);
      atomicAdd(rp.stats + 2, nbricks);
    }
  }
  if (rp.xs.world) {
    // all peer stores of this CTA are performed system-wide (one cumulative fence after the CTA barrier); the last CTA
    // to retire raises this rank's arrival flag on every stripe owner
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned prev = atomicAdd(rp.tile_counter + 1, 1u);
      if (prev == gridDim.x - 1) {
        rp.tile_counter[1] = 0u;
        // the rectangle this rank sent goes next to the flag (rects[parity*16 + rank] follow the 64 flag words)
        for (uint32_t o = 0; o < rp.xs.world; ++o) {
          volatile uint32_t* r = reinterpret_cast<volatile uint32_t*>(rp.xs.flag[o] - rp.xs.flag_index + 64u + 4u * rp.xs.flag_index);
          r[0] = rp.xs.rect[0]; r[1] = rp.xs.rect[1]; r[2] = rp.xs.rect[2]; r[3] = rp.xs.rect[3];
        }
        __threadfence_system();
        for (uint32_t o = 0; o < rp.xs.world; ++o) *reinterpret_cast<volatile uint32_t*>(rp.xs.flag[o]) = rp.xs.frame;
      }
    }
  }
}

typedef void (*raycast_fn)(const RayParams);

// Transfer-function modes run the tile kernel.  -DTVK_K1_REFILL builds the ray-refill kernel instead: same images
// (the GPU tests pass with it), but measured SLOWER on the B200 -- 2048^3 2D-TF lit 5.7-7.5 ms against 3.9 ms, 1D-TF
// 1.9-2.5 ms against 1.3 ms over the threshold sweep (profiles/r02_k1_sweep2.jsonl): the page-table walk and the ray
// set-up then run many times with a quarter of the lanes, which costs more than the idle lanes it removes.
#ifdef TVK_K1_REFILL
#define TVK_TF_KERNEL raycast_refill_kernel
#else
#define TVK_TF_KERNEL raycast_kernel
#endif
template <int BPV, int SAMP>
static raycast_fn pick_kernel(uint32_t mode, bool lit, bool srm1) {
  if (mode == TVK_RM_ISOSURFACE) return raycast_kernel<TVK_RM_ISOSURFACE, false, BPV, SAMP, false>;
  if (mode == TVK_RM_1DTRANS) {
    if (srm1) return lit ? TVK_TF_KERNEL<TVK_RM_1DTRANS, true, BPV, SAMP, true> : TVK_TF_KERNEL<TVK_RM_1DTRANS, false, BPV, SAMP, true>;
    return lit ? TVK_TF_KERNEL<TVK_RM_1DTRANS, true, BPV, SAMP, false> : TVK_TF_KERNEL<TVK_RM_1DTRANS, false, BPV, SAMP, false>;
  }
  return lit ? TVK_TF_KERNEL<TVK_RM_2DTRANS, true, BPV, SAMP, false> : TVK_TF_KERNEL<TVK_RM_2DTRANS, false, BPV, SAMP, false>;
}
template <int BPV>
static raycast_fn pick_sampler(uint32_t sampler, uint32_t mode, bool lit, bool srm1) {
  if (sampler == TVK_SAMPLER_TEX3D) return pick_kernel<BPV, TVK_SAMPLER_TEX3D>(mode, lit, srm1);
  if (sampler == TVK_SAMPLER_EXACT_Q4) return pick_kernel<BPV, TVK_SAMPLER_EXACT_Q4>(mode, lit, srm1);
  if (sampler == TVK_SAMPLER_GATHER) return pick_kernel<BPV, TVK_SAMPLER_GATHER>(mode, lit, srm1);
  return pick_kernel<BPV, TVK_SAMPLER_EXACT>(mode, lit, srm1);
}

// ---- screen rectangle of the ray-marched box -----------------------------------------------------------------
static bool invert4(const double m[16], double inv[16]) {
  double a[4][8];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = m[4 * i + j]; a[i][4 + j] = i == j ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
    if (fabs(a[piv][col]) < 1e-300) return false;
    if (piv != col) for (int j = 0; j < 8; ++j) std::swap(a[piv][j], a[col][j]);
    const double d = a[col][col];
    for (int j = 0; j < 8; ++j) a[col][j] /= d;
    for (int r = 0; r < 4; ++r) if (r != col) { const double f = a[r][col]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[col][j]; }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[4 * i + j] = a[i][4 + j];
  return true;
}

// Conservative pixel rectangle [x0, x1) x [y0, y1) that contains the projection of the box [box_min, box_max] (normalised
// volume coordinates): the box is convex, so with all eight corners in front of the eye its image lies inside the bounding
// rectangle of the projected corners (+2 pixels of slack for the fp32 arithmetic of ray_setup).  Anything else -- a corner
// beside or behind the eye, a singular matrix -- falls back to the whole frame.  TVK_FULL_FRAME=1 forces the fall-back.
static void box_screen_rect(const tvk_raycast_params* p, uint32_t rect[4]) {
  rect[0] = 0; rect[1] = 0; rect[2] = p->width; rect[3] = p->height;
  static const bool full = getenv("TVK_FULL_FRAME") != nullptr;
  if (full) return;
  double ip[16], proj[16];
  for (int i = 0; i < 16; ++i) ip[i] = p->inv_projection[i];
  if (!invert4(ip, proj)) return;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int k = 0; k < 8; ++k) {
    const double v[4] = {(k & 1) ? p->box_max[0] : p->box_min[0], (k & 2) ? p->box_max[1] : p->box_min[1],
                         (k & 4) ? p->box_max[2] : p->box_min[2], 1.0};
    double e[4], c[4];                                     // row vector times matrix, as xform_point
    for (int j = 0; j < 4; ++j) e[j] = v[0] * p->model_to_eye[j] + v[1] * p->model_to_eye[4 + j] + v[2] * p->model_to_eye[8 + j] + v[3] * p->model_to_eye[12 + j];
    for (int j = 0; j < 4; ++j) c[j] = e[0] * proj[j] + e[1] * proj[4 + j] + e[2] * proj[8 + j] + e[3] * proj[12 + j];
    if (!(c[3] > 1e-6) || !(fabs(e[3]) > 1e-12)) return;   // at or behind the eye plane
    const double sx = (c[0] / c[3] * 0.5 + 0.5) * p->width, sy = (c[1] / c[3] * 0.5 + 0.5) * p->height;
    if (!(sx == sx) || !(sy == sy)) return;
    x0 = std::min(x0, sx); x1 = std::max(x1, sx); y0 = std::min(y0, sy); y1 = std::max(y1, sy);
  }
  const double W = p->width, H = p->height;
  x0 = std::max(0.0, floor(x0) - 2.0); y0 = std::max(0.0, floor(y0) - 2.0);
  x1 = std::min(W, ceil(x1) + 2.0); y1 = std::min(H, ceil(y1) + 2.0);
  if (x1 <= x0 || y1 <= y0) { rect[0] = rect[1] = rect[2] = rect[3] = 0; return; }      // the box is off screen
  rect[0] = (uint32_t)x0; rect[1] = (uint32_t)y0; rect[2] = (uint32_t)x1; rect[3] = (uint32_t)y1;
}

// new view: the values the march leaves at a pixel no ray of which meets the box (cleared targets, ray finished)
__global__ void __launch_bounds__(256) fill_outside_kernel(float4* __restrict__ image, float4* __restrict__ pos, float4* __restrict__ aux,
                                                          uint32_t w, uint32_t h, uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t ry1) {
  const size_t n = (size_t)w * h;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), done = make_float4(0.f, 0.f, 0.f, 1000.0f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (size_t)y * w);
    if (x >= rx0 && x < rx1 && y >= ry0 && y < ry1) continue;
    image[i] = zero; aux[i] = zero; pos[i] = done;
  }
}

// sort-last with an empty rectangle: no march is launched, but the stripe owners still wait for this rank's flag
__global__ void publish_empty_kernel(const ExchangeSend xs) {
  if (threadIdx.x != 0) return;
  for (uint32_t o = 0; o < xs.world; ++o) {
    volatile uint32_t* r = reinterpret_cast<volatile uint32_t*>(xs.flag[o] - xs.flag_index + 64u + 4u * xs.flag_index);
    r[0] = 0u; r[1] = 0u; r[2] = 0u; r[3] = 0u;
  }
  __threadfence_system();
  for (uint32_t o = 0; o < xs.world; ++o) *reinterpret_cast<volatile uint32_t*>(xs.flag[o]) = xs.frame;
}

extern "C" int tvk_raycast_subframe(tvk_ctx* c, const tvk_raycast_params* p) {
  if (!c || !p) return TVK_ERR_INVALID;
  if (!c->has_pool) TVK_FAIL(c, TVK_ERR_INVALID, "no brick pool");
  if (p->render_mode > TVK_RM_ISOSURFACE) TVK_FAIL(c, TVK_ERR_INVALID, "bad render mode");
  if (p->render_mode != TVK_RM_ISOSURFACE && !c->d_tf) TVK_FAIL(c, TVK_ERR_INVALID, "no transfer function");
  if (p->render_mode == TVK_RM_1DTRANS && c->tf_h != 1) TVK_FAIL(c, TVK_ERR_INVALID, "1D mode needs a 1D transfer function");
  if (p->hash_table_size == 0 || p->rehash_count == 0) TVK_FAIL(c, TVK_ERR_INVALID, "bad hash table parameters");
  if (!(p->sample_rate_modifier > 0.0f)) TVK_FAIL(c, TVK_ERR_INVALID, "bad sample rate modifier");
  for (int a = 0; a < 3; ++a)
    if (!(p->box_min[a] >= 0.0f && p->box_max[a] <= 1.0f && p->box_min[a] < p->box_max[a]))
      TVK_FAIL(c, TVK_ERR_INVALID, "bad clip box (box_min/box_max must be a sub-box of the unit cube)");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const bool fresh = (c->width != p->width || c->height != p->height);
  int rc = tvk_ensure_frame(c, p->width, p->height, p->hash_table_size);
  if (rc) return rc;
  if (fresh && !p->new_view) TVK_FAIL(c, TVK_ERR_INVALID, "first sub-frame after a resize must be a new view");

  RayParams rp;
  rp.p = *p;
  rp.pc = c->pc;
  rp.pool = c->d_pool;
  rp.elem_bytes = c->pc.bpv * (c->pdesc.sampler == TVK_SAMPLER_EXACT_Q4 ? 4u : 1u);
  rp.tex = c->pool_tex;
  rp.meta = c->d_meta;
  rp.tf = c->d_tf;
  rp.tf_f = c->d_tf_f;       // (replaced below by the opacity-corrected copy when the SRM1 kernels run)
  rp.tf_colmask = c->d_tf_colmask;
  rp.tf_w = c->tf_w; rp.tf_h = c->tf_h;
  rp.hash = c->d_hash;
  rp.pos = c->d_pos[0]; rp.aux = c->d_aux[0]; rp.image = c->d_image;
  // isosurface mode: the hit positions (eye space; w == 0: no hit) are kept in the second position buffer, which this path
  // does not use otherwise -- GLRenderer::Pick reads them back (m_pFBOIsoHit[0]: the first eye's)
  rp.hit_buf = (p->render_mode == TVK_RM_ISOSURFACE && c->eye == 0) ? c->d_pos[1] : nullptr;
  rp.tile_counter = c->d_tile_counter;
  rp.stats = c->d_stats;
  // Screen rectangle of the (sort-last clipped) volume box: the reference rasterises the box's faces, so its fragment
  // shader only runs where the box projects; pixels outside get their cleared values from a fill kernel on a new view and
  // are not visited again
  uint32_t rect[4];
  box_screen_rect(p, rect);
  rp.tile_x0 = rect[0] / 8u; rp.tile_y0 = rect[1] / 4u;
  rp.tiles_x = (rect[2] + 7u) / 8u - rp.tile_x0; rp.tiles_y = (rect[3] + 3u) / 4u - rp.tile_y0;
  const uint32_t rx0 = rp.tile_x0 * 8u, ry0 = rp.tile_y0 * 4u;
  const uint32_t rx1 = std::min(p->width, (rp.tile_x0 + rp.tiles_x) * 8u), ry1 = std::min(p->height, (rp.tile_y0 + rp.tiles_y) * 4u);
  tvk_exchange_fill(c, rp.xs);
  rp.xs.rect[0] = rx0; rp.xs.rect[1] = ry0; rp.xs.rect[2] = rx1; rp.xs.rect[3] = ry1;
  if (rp.xs.world && (c->xchg.width != p->width || c->xchg.height != p->height))
    TVK_FAIL(c, TVK_ERR_INVALID, "exchange was created for another image size");

  const bool srm1 = (1.0f / p->sample_rate_modifier == 1.0f);
  if (srm1 && p->render_mode == TVK_RM_1DTRANS) rp.tf_f = c->d_tf_f1;
  const raycast_fn fn = (c->pc.bpv == 1) ? pick_sampler<1>(c->pdesc.sampler, p->render_mode, p->lighting != 0, srm1)
                                         : pick_sampler<2>(c->pdesc.sampler, p->render_mode, p->lighting != 0, srm1);

  int sms = 0, per_sm = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  TVK_CHECK_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 128, 0));
  if (c->march_ctas_per_sm && (int)c->march_ctas_per_sm < per_sm) per_sm = (int)c->march_ctas_per_sm;
  if (per_sm < 1) per_sm = 1;
  const uint32_t ntiles = rp.tiles_x * rp.tiles_y;
  uint32_t grid = (uint32_t)(sms * per_sm);
  if (grid > (ntiles + 3) / 4) grid = (ntiles + 3) / 4;
  if (grid < 1) grid = 1;

  // GLHashTable::ClearData (:106-110) + tile counter reset, then the march; timed with events
  // (stereo: the second eye's march adds its requests to the table the first eye's left, GLGridLeaper.cpp:947-958)
  if (c->eye == 0) TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_hash, 0, sizeof(uint32_t) * p->hash_table_size, c->stream));
  TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_tile_counter, 0, sizeof(uint32_t), c->stream));
  c->show_stereo = false;
  if (c->eye == 0) c->iso_hits = rp.hit_buf;
  if (rp.hit_buf && p->new_view) TVK_CHECK_CUDA(c, cudaMemsetAsync(rp.hit_buf, 0, sizeof(float4) * (size_t)p->width * p->height, c->stream));
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  if (p->new_view && (rx0 > 0 || ry0 > 0 || rx1 < p->width || ry1 < p->height)) {
    const size_t npx = (size_t)p->width * p->height;
    fill_outside_kernel<<<(unsigned)std::min<size_t>((npx + 255) / 256, (size_t)sms * 8), 256, 0, c->stream>>>(
        rp.image, rp.pos, rp.aux, p->width, p->height, rx0, ry0, rx1, ry1);
    c->launches++;
  }
  if (ntiles) fn<<<grid, 128, 0, c->stream>>>(rp);
  else if (rp.xs.world) publish_empty_kernel<<<1, 32, 0, c->stream>>>(rp.xs);     // nothing to march: the owners still wait for the flag
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  c->launches++;
  c->k1_serial++;
  return TVK_OK;
}

extern "C" int tvk_read_requests(tvk_ctx* c, uint32_t* xyzl, uint32_t cap, uint32_t* count) {
  if (!c || !count || (cap && !xyzl) || !c->d_hash || !c->has_pool) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const uint32_t words = c->hash_size + 8;
  if (c->h_readback_words < words) {
    if (c->h_readback) cudaFreeHost(c->h_readback);
    c->h_readback = nullptr; c->h_readback_words = 0;
    TVK_CHECK_CUDA(c, cudaHostAlloc((void**)&c->h_readback, sizeof(uint32_t) * words, cudaHostAllocDefault));
    c->h_readback_words = words;
  }
  uint32_t* h = c->h_readback;
  unsigned long long* hs = reinterpret_cast<unsigned long long*>(h + ((c->hash_size + 1) & ~1u));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(h, c->d_hash, sizeof(uint32_t) * c->hash_size, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(hs, c->d_stats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->cached_stats[0] = hs[0]; c->cached_stats[1] = hs[1]; c->cached_stats[2] = hs[2];
  c->stats_serial = c->k1_serial;
  // GLHashTable::GetData + Int2Vector (GLHashTable.cpp:65-77,90-104)
  const uint32_t* B = c->pc.l0_bricks;
  const uint32_t vol = B[0] * B[1] * B[2];
  uint32_t n = 0;
  for (uint32_t i = 0; i < c->hash_size && n < cap; ++i) {
    if (!h[i]) continue;
    uint32_t idx = h[i] - 1;
    uint32_t* o = xyzl + 4 * n++;
    o[3] = idx / vol; idx -= o[3] * vol;
    o[2] = idx / (B[0] * B[1]); idx -= o[2] * (B[0] * B[1]);
    o[1] = idx / B[0]; idx -= o[1] * B[0];
    o[0] = idx;
  }
  *count = n;
  return TVK_OK;
}

extern "C" int tvk_read_image(tvk_ctx* c, float* rgba) {
  if (!c || !rgba || !c->d_image) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(rgba, tvk_display_image(c), sizeof(float4) * (size_t)c->width * c->height, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

// glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) of the float target: clamp to [0,1], scale by 255, round to nearest
// (GLFrameCapture.cpp:47-86); flip = the row reversal FrameCapture::SaveImage applies (FrameCapture.h:90-98).
__global__ void __launch_bounds__(256) image_to_rgba8_kernel(const float4* __restrict__ img, uchar4* __restrict__ out,
                                                            uint32_t w, uint32_t h, int flip) {
  const uint32_t n = w * h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t y = i / w, x = i - y * w;
    const float4 v = __ldcs(img + (size_t)(flip ? h - 1 - y : y) * w + x);
    uchar4 o;
    o.x = (unsigned char)(fminf(1.0f, fmaxf(0.0f, v.x)) * 255.0f + 0.5f);
    o.y = (unsigned char)(fminf(1.0f, fmaxf(0.0f, v.y)) * 255.0f + 0.5f);
    o.z = (unsigned char)(fminf(1.0f, fmaxf(0.0f, v.z)) * 255.0f + 0.5f);
    o.w = (unsigned char)(fminf(1.0f, fmaxf(0.0f, v.w)) * 255.0f + 0.5f);
    out[i] = o;
  }
}

// d_rgba8 -> the caller's buffer.  The caller's memory is usually pageable (a std::vector, a numpy array): a direct
// cudaMemcpy from the device to it is staged by the runtime in small pieces; one DMA into a pinned block the context
// keeps + one host memcpy is 3-4x quicker for a 1080p frame.
static int rgba8_to_host(tvk_ctx* c, uint8_t* rgba8, size_t n) {
  // a caller whose frame buffer is page-locked (cudaHostAlloc / cudaHostRegister) gets the frame by DMA, without the
  // landing block and the host copy out of it
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, rgba8) == cudaSuccess && at.type == cudaMemoryTypeHost) {
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(rgba8, c->d_rgba8, n * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream));
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return TVK_OK;
  }
  (void)cudaGetLastError();
  if (!c->h_rgba8 || c->h_rgba8_px < n) {
    if (c->h_rgba8) cudaFreeHost(c->h_rgba8);
    c->h_rgba8 = nullptr; c->h_rgba8_px = 0;
    TVK_CHECK_CUDA(c, cudaMallocHost(&c->h_rgba8, n * sizeof(uchar4)));
    c->h_rgba8_px = n;
  }
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(c->h_rgba8, c->d_rgba8, n * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(rgba8, c->h_rgba8, n * sizeof(uchar4));
  return TVK_OK;
}

extern "C" int tvk_read_image_rgba8(tvk_ctx* c, uint8_t* rgba8, int flip_rows) {
  if (!c || !rgba8 || !c->d_image) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)c->width * c->height;
  if (!c->d_rgba8 || c->rgba8_px < n) {
    cudaFree(c->d_rgba8); c->d_rgba8 = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_rgba8, n * sizeof(uchar4)));
    c->rgba8_px = n;
  }
  int sm_count = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->device));
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count * 8);
  image_to_rgba8_kernel<<<blocks, 256, 0, c->stream>>>(tvk_display_image(c), c->d_rgba8, c->width, c->height, flip_rows);
  TVK_CHECK_CUDA(c, cudaGetLastError());
  return rgba8_to_host(c, rgba8, n);
}

// The same read-back for the sort-last result: the finished image of the fused exchange (rank 0, tvk_exchange_final) goes
// through the same conversion kernel and ONE device-to-host copy of width*height*4 bytes.
extern "C" int tvk_exchange_read_final_rgba8(tvk_ctx* c, uint32_t frame_id, uint8_t* rgba8, int flip_rows) {
  if (!c || !rgba8 || !c->xchg.world) return TVK_ERR_INVALID;
  void* fin = nullptr;
  int rc = tvk_exchange_final(c, frame_id, &fin);            // rank 0: orders the stream behind the landing of every stripe
  if (rc) return rc;
  const uint32_t w = c->xchg.width, h = c->xchg.height;
  const size_t n = (size_t)w * h;
  if (!c->d_rgba8 || c->rgba8_px < n) {
    cudaFree(c->d_rgba8); c->d_rgba8 = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_rgba8, n * sizeof(uchar4)));
    c->rgba8_px = n;
  }
  int sm_count = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->device));
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count * 8);
  image_to_rgba8_kernel<<<blocks, 256, 0, c->stream>>>(static_cast<const float4*>(fin), c->d_rgba8, w, h, flip_rows);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  return rgba8_to_host(c, rgba8, n);
}

extern "C" int tvk_read_resume(tvk_ctx* c, float* pos, float* color) {
  if (!c || !c->d_image) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = sizeof(float4) * (size_t)c->width * c->height;
  if (pos) TVK_CHECK_CUDA(c, cudaMemcpyAsync(pos, c->d_pos[0], bytes, cudaMemcpyDeviceToHost, c->stream));
  if (color) TVK_CHECK_CUDA(c, cudaMemcpyAsync(color, c->d_aux[0], bytes, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

extern "C" int tvk_image_device_ptr(tvk_ctx* c, void** dptr) {
  if (!c || !dptr || !c->d_image) return TVK_ERR_INVALID;
  *dptr = tvk_display_image(c);
  return TVK_OK;
}

extern "C" int tvk_get_stats(tvk_ctx* c, tvk_stats* out, int reset) {
  if (!c || !out) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  unsigned long long h[3];
  if (!reset && c->stats_serial == c->k1_serial) {
    // counters were fetched together with the request table after the last ray-march: no device access needed
    h[0] = c->cached_stats[0]; h[1] = c->cached_stats[1]; h[2] = c->cached_stats[2];
  } else {
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (reset) { TVK_CHECK_CUDA(c, cudaMemsetAsync(c->d_stats, 0, sizeof(h), c->stream)); c->stats_serial = ~0ull; }
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  float ms = 0.f;
  if (c->launches && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_raycast_ms = ms;
  (void)cudaGetLastError();
  out->samples = h[0]; out->rays = h[1]; out->brick_steps = h[2];
  out->last_raycast_ms = c->last_raycast_ms;
  out->launches = c->launches;
  return TVK_OK;
}

// ---------------------------------------------------------------------------
// bit-exactness probes
// ---------------------------------------------------------------------------
__global__ void probe_lod_kernel(const float* depth, uint32_t n, float lod_factor, float wse, uint32_t max_lod, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = compute_lod(depth[i], lod_factor, wse, max_lod);
}
__global__ void probe_index_kernel(const float* xyz, const uint32_t* lod, uint32_t n, const __grid_constant__ PoolConst pc,
                                   uint32_t* index_out, uint32_t* serial_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const f3 e = mk3(clamp01(xyz[3 * i]), clamp01(xyz[3 * i + 1]), clamp01(xyz[3 * i + 2]));
  const uint4 bc = brick_coords(pc, e, lod[i]);
  index_out[i] = brick_index(pc, bc);
  serial_out[i] = serialize(pc, bc);
}

extern "C" int tvk_probe_lod(tvk_ctx* c, const float* depth, uint32_t n, float lod_factor, float wse, uint32_t* lod_out) {
  if (!c || !c->has_pool || !depth || !lod_out) return TVK_ERR_INVALID;
  if (!n) return TVK_OK;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  float* d_in = nullptr; uint32_t* d_out = nullptr;
  TVK_CHECK_CUDA(c, cudaMalloc(&d_in, sizeof(float) * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_out, sizeof(uint32_t) * n));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_in, depth, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
  probe_lod_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_in, n, lod_factor, wse, c->pc.lod_count - 1, d_out);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(lod_out, d_out, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_in); cudaFree(d_out);
  return TVK_OK;
}

extern "C" int tvk_probe_brick_index(tvk_ctx* c, const float* norm_xyz, const uint32_t* lod, uint32_t n,
                                     uint32_t* index_out, uint32_t* serial_out) {
  if (!c || !c->has_pool || !norm_xyz || !lod || !index_out || !serial_out) return TVK_ERR_INVALID;
  if (!n) return TVK_OK;
  for (uint32_t i = 0; i < n; ++i) if (lod[i] >= c->pc.lod_count) TVK_FAIL(c, TVK_ERR_INVALID, "lod out of range");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  float* d_xyz = nullptr; uint32_t *d_lod = nullptr, *d_i = nullptr, *d_s = nullptr;
  TVK_CHECK_CUDA(c, cudaMalloc(&d_xyz, sizeof(float) * 3 * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_lod, sizeof(uint32_t) * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_i, sizeof(uint32_t) * n));
  TVK_CHECK_CUDA(c, cudaMalloc(&d_s, sizeof(uint32_t) * n));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_xyz, norm_xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_lod, lod, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
  probe_index_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_xyz, d_lod, n, c->pc, d_i, d_s);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(index_out, d_i, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(serial_out, d_s, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_xyz); cudaFree(d_lod); cudaFree(d_i); cudaFree(d_s);
  return TVK_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Stereo (GLGridLeaper::Render3DRegion's loop over EStereoID, GLGridLeaper.cpp:922-958; GLRenderer::EndFrame :757-812;
// Compose-Anaglyphs-FS.glsl, Compose-Scanline-FS.glsl, Compose-SBS-FS.glsl, Compose-AF-FS.glsl).  The reference keeps one
// set of frame buffers per eye (m_pFBO3DImageNext[2], m_pFBORayStart[2], m_pFBOStartColor[2]) and ONE request table.
// ---------------------------------------------------------------------------------------------------------
extern "C" int tvk_select_eye(tvk_ctx* c, uint32_t eye) {
  if (!c || eye > 1) return TVK_ERR_INVALID;
  if (eye == c->eye) return TVK_OK;
  tvk_ctx::EyeState& o = c->other_eye;
  std::swap(c->d_image, o.image);
  for (int i = 0; i < 2; ++i) { std::swap(c->d_pos[i], o.pos[i]); std::swap(c->d_aux[i], o.aux[i]); }
  std::swap(c->width, o.width);
  std::swap(c->height, o.height);
  c->eye = eye;
  return TVK_OK;
}

extern "C" int tvk_read_eye_image(tvk_ctx* c, uint32_t eye, float* rgba) {
  if (!c || eye > 1 || !rgba) return TVK_ERR_INVALID;
  const float4* src = eye == c->eye ? c->d_image : c->other_eye.image;
  const uint32_t w = eye == c->eye ? c->width : c->other_eye.width, h = eye == c->eye ? c->height : c->other_eye.height;
  if (!src || !w || !h) TVK_FAIL(c, TVK_ERR_INVALID, "that eye has no picture yet");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(rgba, src, sizeof(float4) * (size_t)w * h, cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

struct StereoArgs {
  const float4* left;
  const float4* right;
  float4* out;
  uint32_t w, h, mode, af_id;
  double right_shift;                  // SBS: 2 * fSplitCoord * width
  float split;
};

__global__ void __launch_bounds__(256) compose_stereo_kernel(const __grid_constant__ StereoArgs a) {
  const uint32_t n = a.w * a.h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t y = i / a.w, x = i - y * a.w;
    float4 o;
    if (a.mode == TVK_SM_RB) {
      const float4 l = a.left[i], r = a.right[i];
      const float gl = fmaf(l.z, 0.11f, fmaf(l.y, 0.59f, l.x * 0.3f)), gr = fmaf(r.z, 0.11f, fmaf(r.y, 0.59f, r.x * 0.3f));
      o = make_float4(gl, gr * 0.5f, gr, fmaxf(l.w, r.w));
    } else if (a.mode == TVK_SM_SCANLINE) {
      o = (y & 1u) ? a.right[i] : a.left[i];                           // fLine/2 == floor(fLine/2): even lines show the left eye
    } else if (a.mode == TVK_SM_SBS) {
      // both halves show a whole eye image squeezed to half the width: GL_NEAREST at x*2 is texel 2x+1
      const float tx = ((float)x + 0.5f) / (float)a.w;
      const bool lhs = tx < a.split;
      const double t = lhs ? (double)(2u * x + 1u) : (double)(2u * x + 1u) - a.right_shift;
      const int sx = min(max((int)floor(t), 0), (int)a.w - 1);
      o = (lhs ? a.left : a.right)[(size_t)y * a.w + sx];
    } else {
      o = a.af_id == 0 ? a.left[i] : a.right[i];
    }
    a.out[i] = o;
  }
}

extern "C" int tvk_compose_stereo(tvk_ctx* c, uint32_t mode, uint32_t swap_eyes, float split_coord, uint32_t alternating_frame_id) {
  if (!c || mode > TVK_SM_AF) return TVK_ERR_INVALID;
  if (!c->d_image || !c->other_eye.image) TVK_FAIL(c, TVK_ERR_INVALID, "stereo compose needs a picture for each eye (tvk_select_eye)");
  if (c->width != c->other_eye.width || c->height != c->other_eye.height) TVK_FAIL(c, TVK_ERR_INVALID, "the two eyes' pictures differ in size");
  if (mode == TVK_SM_SBS && !(split_coord > 0.0f && split_coord <= 0.5f)) TVK_FAIL(c, TVK_ERR_INVALID, "bad split coordinate");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)c->width * c->height;
  if (!c->d_stereo || c->stereo_px < n) {
    cudaFree(c->d_stereo); c->d_stereo = nullptr; c->stereo_px = 0;
    TVK_CHECK_CUDA(c, cudaMalloc(&c->d_stereo, n * sizeof(float4)));
    c->stereo_px = n;
  }
  const float4* eye0 = c->eye == 0 ? c->d_image : c->other_eye.image;
  const float4* eye1 = c->eye == 0 ? c->other_eye.image : c->d_image;
  StereoArgs a;
  a.left = swap_eyes ? eye1 : eye0;                                   // m_bStereoEyeSwap binds the textures the other way round
  a.right = swap_eyes ? eye0 : eye1;
  a.out = c->d_stereo; a.w = c->width; a.h = c->height; a.mode = mode; a.af_id = alternating_frame_id;
  a.split = split_coord; a.right_shift = 2.0 * (double)split_coord * (double)c->width;
  int sm_count = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->device));
  compose_stereo_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count * 8), 256, 0, c->stream>>>(a);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  c->show_stereo = true;
  return TVK_OK;
}
