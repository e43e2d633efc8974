/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's
 * ray-guided single-pass ray-caster (GLGridLeaper) and of the brick-pool
 * bookkeeping around it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Parity status: the INTEGER parts (brick indexing, LOD selection, request
 * serialisation, visibility flags) follow the reference's C++/GLSL literally and
 * are what the CUDA path must match bit-exactly.  The IMAGE part is
 * "parity unpinned": the reference's pixels come out of an OpenGL driver
 * (rasteriser, texture filtering, float blending) that is not in /root/reference
 * and cannot run in this container (no libGL/Mesa/Qt), and the reference's own
 * tests hold no golden image (Tuvok/test/render/render.cpp writes test.png and
 * compares nothing).  This file therefore restates the GLSL in scalar fp32 with
 * GL_LINEAR/CLAMP_TO_EDGE texture filtering in full fp32 and GL_NEAREST
 * transfer-function fetches, and is the arbiter the CUDA images are held to
 * (PSNR >= 45 dB, max channel error <= 2/255).
 *
 * Floating-point contract: every operation below is a single IEEE-754 fp32 operation (add, sub,
 * mul, div, sqrt, fma) in a fixed order -- fmaf() where a multiply-add is fused, nothing else is
 * contracted (-ffp-contract=off) -- and the CUDA kernel issues the same operations in the same
 * order (nvcc -fmad=false + explicit fmaf).  The two therefore agree bit for bit except where
 * powf/log2 enter (sample-rate modifier != 1, LOD boundaries), and in the shading block of the
 * lit / 2D-TF modes, where the kernel normalises with rsqrtf and multiplies by a per-brick
 * reciprocal while this file keeps sqrtf and the divisions of the GLSL (tolerance 2e-3 in the tests,
 * measured 3e-5).
 *
 * Restated items (paths relative to /root/reference/Tuvok/):
 *   orc_pool_layout     GetLoDSize/GetFloatBrickLayout/GetBrickLayout, LoD offset
 *                       table                Renderer/GL/GLVolumePool.cpp:82-117,222-230
 *   get_brick           generated GLSL GetBrick & helpers   GLVolumePool.cpp:484-620
 *   compute_lod         ComputeLOD                           GLVolumePool.cpp:653-656
 *   to_pool_space       TransformToPoolSpace/GetSampleDelta  GLVolumePool.cpp:622-635
 *   sample_pool         texture(volumePool,..) GL_LINEAR     GLVolumePool.cpp:637-651
 *   hash_report         Hash/Serialize/AccessHashTable       Renderer/GL/GLHashTable.cpp:136-182
 *   march_ray           main()                               Shaders/GLGridLeaper-blend.glsl:65-228
 *   march_ray_iso       main()                               Shaders/GLGridLeaper-iso.glsl:68-200
 *   color_*             ComputeColorFromVolume               Shaders/GLGridLeaper-Method-{1D,1D-L,2D,2D-L}.glsl
 *   gradient/normal     Shaders/GLGridLeaper-GradientTools.glsl:6-23
 *   lighting            Shaders/lighting.glsl:33-43
 *   iso hit/refine      Shaders/GLGridLeaper-Method-iso.glsl:9-31
 *   compose_iso         Shaders/Compose-FS.glsl:38-75
 *   ray entry/exit      GLGridLeaper::FillRayEntryBuffer (GLGridLeaper.cpp:579-620),
 *                       GLGridLeaper-NearPlane-VS/-entry-VS/-frontfaces-FS.glsl;
 *                       the rasterised bbox faces are replaced by the analytic
 *                       ray/box intersection of the same pixel-centre ray
 *   orc_visibility      RecomputeVisibilityForOctree/ContainsData GLVolumePool.cpp:1021-1359,963-989
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include "tvk_cuda.h"

typedef struct orc_pool_t {
  const void* data;            /* pool voxels, x fastest, pool_size[0..2]        */
  uint32_t bpv;
  uint32_t volume_size[3];
  uint32_t max_brick[3];
  uint32_t overlap;
  uint32_t capacity[3];
  uint32_t lod_count;
  const uint32_t* meta;        /* metadata, pool index order                      */
} orc_pool_t;

typedef struct layout_t {
  uint32_t lod_count;
  float    lod_layout[TVK_MAX_LOD][3];   /* vLODLayout (float, epsilon nudged)       */
  uint32_t lod_layout_x[TVK_MAX_LOD];    /* iLODLayoutSize.x = ceil(layout.x)        */
  uint32_t lod_layout_xy[TVK_MAX_LOD];   /* iLODLayoutSize.y = ceil(x)*ceil(y)       */
  uint32_t lod_offset[TVK_MAX_LOD];      /* vLODOffset                               */
  uint32_t lod_bricks[TVK_MAX_LOD][3];   /* integer GetBrickLayout                   */
  uint32_t total_bricks;
  uint32_t l0_bricks[3];                 /* hash Serialize strides (m_maxBrickCount) */
} layout_t;

/* GLVolumePool.cpp:82-117 and ctor :222-230 */
static void pool_layout(const uint32_t vol[3], const uint32_t inner[3], uint32_t lod_count,
                        layout_t* L) {
  memset(L, 0, sizeof(*L));
  L->lod_count = lod_count;
  uint32_t base[3];
  for (int a = 0; a < 3; ++a) base[a] = (uint32_t)ceil((double)vol[a] / inner[a]);
  uint32_t off = 0;
  for (uint32_t l = 0; l < lod_count; ++l) {
    const uint32_t pw = 1u << l;
    for (int a = 0; a < 3; ++a) {
      float f = (float)vol[a] / (float)inner[a];
      f /= (float)pw;
      if ((float)(uint32_t)f == f) f -= f * FLT_EPSILON;
      L->lod_layout[l][a] = f;
      L->lod_bricks[l][a] = (uint32_t)ceil((double)base[a] / pw);
    }
    L->lod_layout_x[l] = (uint32_t)ceilf(L->lod_layout[l][0]);
    L->lod_layout_xy[l] = (uint32_t)ceilf(L->lod_layout[l][0]) * (uint32_t)ceilf(L->lod_layout[l][1]);
    L->lod_offset[l] = off;
    off += L->lod_bricks[l][0] * L->lod_bricks[l][1] * L->lod_bricks[l][2];
  }
  L->total_bricks = L->lod_offset[lod_count - 1] + 1; /* m_iTotalBrickCount, :877 */
  for (int a = 0; a < 3; ++a) L->l0_bricks[a] = L->lod_bricks[0][a];
}

/* exported for tests: layout tables of the pool */
void orc_pool_layout(const uint32_t vol[3], const uint32_t max_brick[3], uint32_t overlap,
                     uint32_t lod_count, float* lod_layout, uint32_t* lod_offset,
                     uint32_t* lod_bricks, uint32_t* total) {
  uint32_t inner[3] = {max_brick[0] - 2 * overlap, max_brick[1] - 2 * overlap, max_brick[2] - 2 * overlap};
  layout_t L;
  pool_layout(vol, inner, lod_count, &L);
  for (uint32_t l = 0; l < lod_count; ++l) {
    for (int a = 0; a < 3; ++a) { lod_layout[3 * l + a] = L.lod_layout[l][a]; lod_bricks[3 * l + a] = L.lod_bricks[l][a]; }
    lod_offset[l] = L.lod_offset[l];
  }
  *total = L.total_bricks;
}

/* ComputeLOD, GLVolumePool.cpp:653-656.  uint() of a negative / NaN float is
 * undefined in GLSL; NVIDIA saturates to 0 and so do we. */
static uint32_t compute_lod(float dist, float lod_factor, float wse, uint32_t max_lod) {
  const float v = log2f(lod_factor * (-dist) / wse);
  uint32_t l = (v > 0.0f) ? ((v >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t)v) : 0u;
  return l < max_lod ? l : max_lod;
}

void orc_probe_lod(const float* depth, uint32_t n, float lod_factor, float wse, uint32_t max_lod,
                   uint32_t* out) {
  for (uint32_t i = 0; i < n; ++i) out[i] = compute_lod(depth[i], lod_factor, wse, max_lod);
}

typedef struct ctx_t {
  const tvk_raycast_params* p;
  const orc_pool_t* pool;
  layout_t L;
  float pool_size[3];
  float overlap_tc[3];
  const uint8_t* tf;
  uint32_t tf_w, tf_h;
  uint32_t* hash;
  uint64_t samples, bricks, rays, drops;   /* drops: reports that found no free entry within rehash_count probes */
} ctx_t;

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* Hash(), GLHashTable.cpp:136-182 (sequential compare-and-swap) */
static void hash_report(ctx_t* c, const uint32_t bc[4]) {
  const uint32_t* B = c->L.l0_bricks;
  const uint32_t serial = 1 + bc[0] + bc[1] * B[0] + bc[2] * B[0] * B[1] + bc[3] * B[0] * B[1] * B[2];
  uint32_t rehash = 0;
  do {
    const uint32_t h = (serial + rehash) % c->p->hash_table_size;
    const uint32_t cur = __sync_val_compare_and_swap(&c->hash[h], 0u, serial); /* imageAtomicCompSwap */
    if (cur == 0 || cur == serial) return;
  } while (++rehash < c->p->rehash_count);
  c->drops++;
}

static inline void brick_coords(const ctx_t* c, const float p[3], uint32_t lod, uint32_t bc[4]) {
  for (int a = 0; a < 3; ++a) bc[a] = (uint32_t)(p[a] * c->L.lod_layout[lod][a]);
  bc[3] = lod;
}
static inline uint32_t brick_index(const ctx_t* c, const uint32_t bc[4]) {
  return c->L.lod_offset[bc[3]] + bc[0] + bc[1] * c->L.lod_layout_x[bc[3]] + bc[2] * c->L.lod_layout_xy[bc[3]];
}
static inline uint32_t brick_info(const ctx_t* c, const uint32_t bc[4]) {
  return c->pool->meta[brick_index(c, bc)];
}

void orc_probe_brick_index(const uint32_t vol[3], const uint32_t max_brick[3], uint32_t overlap,
                           uint32_t lod_count, const float* xyz, const uint32_t* lod, uint32_t n,
                           uint32_t* index_out, uint32_t* serial_out) {
  ctx_t c;
  memset(&c, 0, sizeof(c));
  uint32_t inner[3] = {max_brick[0] - 2 * overlap, max_brick[1] - 2 * overlap, max_brick[2] - 2 * overlap};
  pool_layout(vol, inner, lod_count, &c.L);
  for (uint32_t i = 0; i < n; ++i) {
    float p[3] = {clampf(xyz[3 * i], 0, 1), clampf(xyz[3 * i + 1], 0, 1), clampf(xyz[3 * i + 2], 0, 1)};
    uint32_t bc[4];
    brick_coords(&c, p, lod[i], bc);
    index_out[i] = brick_index(&c, bc);
    const uint32_t* B = c.L.l0_bricks;
    serial_out[i] = 1 + bc[0] + bc[1] * B[0] + bc[2] * B[0] * B[1] + bc[3] * B[0] * B[1] * B[2];
  }
}

/* GetBrick, GLVolumePool.cpp:552-620 */
static int get_brick(ctx_t* c, const float entry_in[3], uint32_t* lod, const float dir[3],
                     float pool_entry[3], float pool_exit[3], float norm_exit[3], int* empty,
                     float scale[3], float trans[3], uint32_t bc[4]) {
  const uint32_t max_lod = c->L.lod_count - 1;
  float e[3] = {clampf(entry_in[0], 0, 1), clampf(entry_in[1], 0, 1), clampf(entry_in[2], 0, 1)};
  int found = 1;
  brick_coords(c, e, *lod, bc);
  uint32_t info = brick_info(c, bc);
  c->bricks++;
  if (info == TVK_BI_MISSING) {
    const uint32_t start = *lod;
    hash_report(c, bc);
    found = 0;
    do {
      (*lod)++;
      brick_coords(c, e, *lod, bc);
      info = brick_info(c, bc);
      switch (c->p->missing_strategy) {
        case TVK_MBS_REQUEST_ALL: if (info == TVK_BI_MISSING) hash_report(c, bc); break;
        case TVK_MBS_SKIP_ONE_LEVEL: if (info == TVK_BI_MISSING && start + 1 == *lod) hash_report(c, bc); break;
        case TVK_MBS_SKIP_TWO_LEVELS: if (info == TVK_BI_MISSING && start + 2 == *lod) hash_report(c, bc); break;
        default: break;
      }
    } while (info == TVK_BI_MISSING);
  }
  *empty = (info <= TVK_BI_EMPTY);
  if (*empty) {
    for (uint32_t l = *lod + 1; l < max_lod; ++l) {
      uint32_t lbc[4];
      brick_coords(c, e, l, lbc);
      const uint32_t linfo = brick_info(c, lbc);
      if (linfo == TVK_BI_CHILD_EMPTY) { memcpy(bc, lbc, sizeof(lbc)); info = linfo; *lod = l; }
      else break;
    }
  }
  /* GetBrickCorners + BrickExit */
  float c0[3], c1[3];
  for (int a = 0; a < 3; ++a) {
    c0[a] = (float)bc[a] / c->L.lod_layout[bc[3]][a];
    c1[a] = (float)(bc[a] + 1) / c->L.lod_layout[bc[3]][a];
  }
  float tmin = INFINITY;
  for (int a = 0; a < 3; ++a) {
    const float div = 1.0f / dir[a];
    const float corner = (div >= 0.0f) ? c1[a] : c0[a];   /* step(0.0, div) */
    const float t = (corner - e[a]) * div;
    tmin = fminf(tmin, t);                                /* NaN (0*inf) is ignored */
  }
  for (int a = 0; a < 3; ++a) norm_exit[a] = fmaf(tmin, dir[a], e[a]);
  if (*empty) return found;
  /* InfoToCoords + BrickPoolCoords + NormCoordsToPoolCoords */
  const uint32_t idx = info - TVK_BI_FLAG_COUNT;
  const uint32_t* cap = c->pool->capacity;
  const uint32_t s[3] = {idx % cap[0], (idx / cap[0]) % cap[1], idx / (cap[0] * cap[1])};
  for (int a = 0; a < 3; ++a) {
    const uint32_t vp = s[a] * c->pool->max_brick[a];
    const float pc0 = (float)vp / c->pool_size[a] + c->overlap_tc[a];
    const float pc1 = (float)(vp + c->pool->max_brick[a]) / c->pool_size[a] - c->overlap_tc[a];
    scale[a] = (pc1 - pc0) / (c1[a] - c0[a]);
    trans[a] = fmaf(-c0[a], scale[a], pc0);
    pool_entry[a] = fmaf(e[a], scale[a], trans[a]);
    pool_exit[a] = fmaf(norm_exit[a], scale[a], trans[a]);
  }
  return found;
}

/* texture(volumePool, p).r with GL_LINEAR / GL_CLAMP_TO_EDGE on UNORM8/16 */
static inline float voxel(const ctx_t* c, int x, int y, int z) {
  const int sx = (int)c->pool_size[0], sy = (int)c->pool_size[1], sz = (int)c->pool_size[2];
  x = x < 0 ? 0 : (x >= sx ? sx - 1 : x);
  y = y < 0 ? 0 : (y >= sy ? sy - 1 : y);
  z = z < 0 ? 0 : (z >= sz ? sz - 1 : z);
  const size_t i = (size_t)x + (size_t)sx * ((size_t)y + (size_t)sy * (size_t)z);
  return c->pool->bpv == 1 ? (float)((const uint8_t*)c->pool->data)[i] : (float)((const uint16_t*)c->pool->data)[i];
}
/* trilinear filter on the raw texel values, then the UNORM scale (1/255 or 1/65535) */
static float sample_pool(const ctx_t* c, const float p[3]) {
  const float u = fmaf(p[0], c->pool_size[0], -0.5f), v = fmaf(p[1], c->pool_size[1], -0.5f), w = fmaf(p[2], c->pool_size[2], -0.5f);
  const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
  const int x = (int)fu, y = (int)fv, z = (int)fw;
  const float a = u - fu, b = v - fv, g = w - fw;
  const float c000 = voxel(c, x, y, z), c100 = voxel(c, x + 1, y, z);
  const float c010 = voxel(c, x, y + 1, z), c110 = voxel(c, x + 1, y + 1, z);
  const float c001 = voxel(c, x, y, z + 1), c101 = voxel(c, x + 1, y, z + 1);
  const float c011 = voxel(c, x, y + 1, z + 1), c111 = voxel(c, x + 1, y + 1, z + 1);
  const float c00 = fmaf(a, c100 - c000, c000), c10 = fmaf(a, c110 - c010, c010);
  const float c01 = fmaf(a, c101 - c001, c001), c11 = fmaf(a, c111 - c011, c011);
  const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
  return fmaf(g, c1 - c0, c0) * (c->pool->bpv == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}

/* texture(transferFunction, s): RGBA8, GL_NEAREST, clamp to edge */
static inline int nearest_index(float s, uint32_t n) {
  const float f = floorf(s * (float)n);
  if (!(f > 0.0f)) return 0;
  return f >= (float)n ? (int)n - 1 : (int)f;
}
static void tf_fetch1d(const ctx_t* c, float s, float out[4]) {
  const uint8_t* t = c->tf + 4 * nearest_index(s, c->tf_w);
  for (int k = 0; k < 4; ++k) out[k] = (float)t[k] * (1.0f / 255.0f);
}
static void tf_fetch2d(const ctx_t* c, float s, float g, float out[4]) {
  const uint8_t* t = c->tf + 4 * ((size_t)nearest_index(g, c->tf_h) * c->tf_w + nearest_index(s, c->tf_w));
  for (int k = 0; k < 4; ++k) out[k] = (float)t[k] * (1.0f / 255.0f);
}

static inline float dot3(const float a[3], const float b[3]) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }
static inline float len3(const float v[3]) { return sqrtf(dot3(v, v)); }
static inline float pow8(float x) { x *= x; x *= x; return x * x; }

/* ComputeGradient, GLGridLeaper-GradientTools.glsl:6-16 (note the y sign) */
static void gradient(const ctx_t* c, const float p[3], float g[3]) {
  const float dx = 1.0f / c->pool_size[0], dy = 1.0f / c->pool_size[1], dz = 1.0f / c->pool_size[2];
  float q[3];
#define S(ox, oy, oz) (q[0] = p[0] + (ox), q[1] = p[1] + (oy), q[2] = p[2] + (oz), sample_pool(c, q))
  const float xp = S(+dx, 0, 0), xm = S(-dx, 0, 0);
  const float yp = S(0, -dy, 0), ym = S(0, +dy, 0);
  const float zp = S(0, 0, +dz), zm = S(0, 0, -dz);
#undef S
  g[0] = (xm - xp) / 2.0f; g[1] = (yp - ym) / 2.0f; g[2] = (zm - zp) / 2.0f;
}

/* Lighting, lighting.glsl:33-43 */
static void lighting(const ctx_t* c, const float pos[3], const float n[3], const float diffuse[3],
                     float out[3]) {
  const tvk_raycast_params* p = c->p;
  /* sort-last: box-local model position -> the global volume's (identity when the fields are zero) */
  const float gp[3] = {fmaf(pos[0], 1.0f + p->model_pos_scale_m1[0], p->model_pos_offset[0]),
                       fmaf(pos[1], 1.0f + p->model_pos_scale_m1[1], p->model_pos_offset[1]),
                       fmaf(pos[2], 1.0f + p->model_pos_scale_m1[2], p->model_pos_offset[2])};
  float v[3] = {p->model_eye_pos[0] - gp[0], p->model_eye_pos[1] - gp[1], p->model_eye_pos[2] - gp[2]};
  float l = len3(v);
  for (int a = 0; a < 3; ++a) v[a] /= l;
  /* reflect(I,N) = I - 2*dot(N,I)*N */
  const float d2 = 2.0f * dot3(n, v);
  float r[3] = {fmaf(-d2, n[0], v[0]), fmaf(-d2, n[1], v[1]), fmaf(-d2, n[2], v[2])};
  l = len3(r);
  for (int a = 0; a < 3; ++a) r[a] /= l;
  const float ndl = fmaxf(fabsf(dot3(n, p->model_light_dir)), 0.0f);
  const float spec = pow8(fmaxf(dot3(r, p->model_light_dir), 0.0f));   /* pow(x, 8.0) */
  for (int a = 0; a < 3; ++a)
    out[a] = clampf(fmaf(p->light_specular[a], spec, fmaf(diffuse[a], ndl, p->light_ambient[a])), 0.0f, 1.0f);
}

/* ComputeColorFromVolume for the four scalar variants */
static void color_from_volume(const ctx_t* c, const float pc[3], const float model_pos[3], float col[4]) {
  const tvk_raycast_params* p = c->p;
  const float data = sample_pool(c, pc);
  if (p->render_mode == TVK_RM_1DTRANS) {
    tf_fetch1d(c, data * p->trans_scale, col);
    if (p->lighting) {
      float g[3], n[3];
      gradient(c, pc, g);
      for (int a = 0; a < 3; ++a) n[a] = g[a] * p->domain_scale[a];
      const float l = len3(n);
      if (l > 0.0f) for (int a = 0; a < 3; ++a) n[a] /= l;
      const float diff[3] = {col[0] * p->light_diffuse[0], col[1] * p->light_diffuse[1], col[2] * p->light_diffuse[2]};
      lighting(c, model_pos, n, diff, col);
    }
  } else {
    float g[3];
    gradient(c, pc, g);
    const float gm = len3(g);
    tf_fetch2d(c, data * p->trans_scale, fmaf(-gm, p->gradient_scale, 1.0f), col);
    if (p->lighting) {
      float n[3];
      for (int a = 0; a < 3; ++a) n[a] = p->domain_scale[a] * ((gm > 0) ? g[a] / gm : g[a]);
      const float diff[3] = {col[0] * p->light_diffuse[0], col[1] * p->light_diffuse[1], col[2] * p->light_diffuse[2]};
      lighting(c, model_pos, n, diff, col);
    }
  }
}

/* v' = v*M, FLOATMATRIX4 memory order (Vectors.h:434) */
static void xform_point(const float m[16], const float v[3], float w, float out[4]) {
  for (int j = 0; j < 4; ++j) out[j] = fmaf(w, m[12 + j], fmaf(v[2], m[8 + j], fmaf(v[1], m[4 + j], v[0] * m[j])));
}

/* Analytic ray set-up for pixel (px,py): replaces the rasterised near-plane quad
 * and bbox faces (GLGridLeaper.cpp:579-620, 835-843).  Returns 0 if the ray
 * misses the box.  entry/exit = (normalised volume coords, eye-space z). */
static int ray_setup_eye(const tvk_raycast_params* p, uint32_t px, uint32_t py, float entry[4], float exitp[4], float* exit_eye) {
  const float ndc[3] = {fmaf(((float)px + 0.5f) / (float)p->width, 2.0f, -1.0f),
                        fmaf(((float)py + 0.5f) / (float)p->height, 2.0f, -1.0f), -1.0f};
  float n4[4];
  xform_point(p->inv_projection, ndc, 1.0f, n4);
  const float n[3] = {n4[0] / n4[3], n4[1] / n4[3], n4[2] / n4[3]};   /* near-plane point, eye space */
  float O[4], D[4];
  const float zero[3] = {0, 0, 0};
  xform_point(p->eye_to_model, zero, 1.0f, O);
  xform_point(p->eye_to_model, n, 0.0f, D);
  float s0 = -INFINITY, s1 = INFINITY;
  for (int a = 0; a < 3; ++a) {
    const float inv = 1.0f / D[a];
    float ta = (p->box_min[a] - O[a]) * inv, tb = (p->box_max[a] - O[a]) * inv;
    if (D[a] == 0.0f) {
      if (O[a] < p->box_min[a] || O[a] > p->box_max[a]) return 0;
      continue;
    }
    if (ta > tb) { const float t = ta; ta = tb; tb = t; }
    s0 = fmaxf(s0, ta);
    s1 = fminf(s1, tb);
  }
  if (p->clip_enabled) {
    /* keep the half space dot(plane.xyz, x) + plane.w <= 0 ... (model space) */
    const float dn = dot3(p->clip_plane, D);
    const float on = dot3(p->clip_plane, O) + p->clip_plane[3];
    if (dn == 0.0f) { if (on > 0.0f) return 0; }
    else {
      const float t = -on / dn;
      if (dn > 0.0f) s1 = fminf(s1, t); else s0 = fmaxf(s0, t);
    }
  }
  const float se = fmaxf(s0, 1.0f);               /* eye inside the box: near plane */
  if (!(s1 > se)) return 0;
  const float pe[3] = {n[0] * se, n[1] * se, n[2] * se};
  const float px1[3] = {n[0] * s1, n[1] * s1, n[2] * s1};
  float t4[4];
  xform_point(p->eye_to_model, pe, 1.0f, t4);
  entry[0] = t4[0]; entry[1] = t4[1]; entry[2] = t4[2]; entry[3] = pe[2];
  xform_point(p->eye_to_model, px1, 1.0f, t4);
  exitp[0] = t4[0]; exitp[1] = t4[1]; exitp[2] = t4[2]; exitp[3] = px1[2];
  if (exit_eye) { exit_eye[0] = px1[0]; exit_eye[1] = px1[1]; exit_eye[2] = px1[2]; }
  return 1;
}
static int ray_setup(const tvk_raycast_params* p, uint32_t px, uint32_t py, float entry[4], float exitp[4]) {
  return ray_setup_eye(p, px, py, entry, exitp, NULL);
}

/* The rasteriser's part of a sub-frame for the GLSL harness (oracle/glslref.py): per pixel, whether the bbox back faces
 * cover it, the ray entry (what FillRayEntryBuffer leaves in rayStartPoint) and the back-face position in EYE space
 * (the vertex shader's vPosInViewCoords; the fragment shader transforms it itself). */
void orc_ray_setup_image(const tvk_raycast_params* p, float* entry, float* exit_eye, uint8_t* covered) {
  for (uint32_t py = 0; py < p->height; ++py)
    for (uint32_t px = 0; px < p->width; ++px) {
      const size_t q = (size_t)py * p->width + px;
      float e[4], x[4];
      float xe[3];
      covered[q] = (uint8_t)ray_setup_eye(p, px, py, e, x, xe);
      memset(entry + 4 * q, 0, 16);
      memset(exit_eye + 3 * q, 0, 12);
      if (!covered[q]) continue;
      memcpy(entry + 4 * q, e, 16);
      memcpy(exit_eye + 3 * q, xe, 12);
    }
}

static void under(const float src[4], float dst[4]) {   /* Compositing.glsl:33-38 */
  const float w = (1.0f - dst[3]) * src[3];
  dst[0] = fmaf(src[0], w, dst[0]); dst[1] = fmaf(src[1], w, dst[1]); dst[2] = fmaf(src[2], w, dst[2]);
  dst[3] = fmaf(src[3], 1.0f - dst[3], dst[3]);
}

/* GLGridLeaper-blend.glsl main().  pos/col: in = rayStartPoint/rayStartColor,
 * out = rayResumePos/rayResumeColor; acc = accRayColor. */
static void march_ray(ctx_t* c, const float exitp[4], float pos[4], float col[4], float acc[4]) {
  const tvk_raycast_params* p = c->p;
  memcpy(acc, col, 16);
  if (pos[3] == 1000.0f) return;
  const float entry[3] = {pos[0], pos[1], pos[2]};
  const float entry_depth = pos[3], exit_depth = exitp[3];
  const float dir[3] = {exitp[0] - entry[0], exitp[1] - entry[1], exitp[2] - entry[2]};
  const float ray_len = len3(dir);
  /* TransformToPoolSpace */
  float vdir[3] = {dir[0] * (float)c->pool->volume_size[0], dir[1] * (float)c->pool->volume_size[1], dir[2] * (float)c->pool->volume_size[2]};
  const float vl = len3(vdir);
  for (int a = 0; a < 3; ++a) vdir[a] = vdir[a] / vl / c->pool_size[a] / (2.0f * p->sample_rate_modifier);
  const float step_size = len3(vdir);
  float t = 0.0f;
  int optimal = 1;
  const float voxel_size = 0.125f / 2000.0f;
  float cur[3] = {entry[0], entry[1], entry[2]};
  uint32_t last_bc[4] = {0, 0, 0, 9999};
  const float oc = 1.0f / p->sample_rate_modifier;
  if (ray_len > voxel_size) {
    for (uint32_t j = 0; j < 100; ++j) {
      const float depth = fmaf(exit_depth, t, entry_depth * (1.0f - t));   /* mix */
      uint32_t lod = compute_lod(depth, p->lod_factor, p->level_zero_wse, c->L.lod_count - 1);
      float pe[3], px[3], nexit[3], scale[3], trans[3];
      uint32_t bc[4];
      int empty;
      const int ok = get_brick(c, cur, &lod, dir, pe, px, nexit, &empty, scale, trans, bc);
      if (!ok && optimal) {
        optimal = 0;
        pos[0] = cur[0]; pos[1] = cur[1]; pos[2] = cur[2]; pos[3] = depth;
        memcpy(col, acc, 16);
      }
      if (!empty && memcmp(last_bc, bc, sizeof(bc)) != 0) {
        const float d0[3] = {px[0] - pe[0], px[1] - pe[1], px[2] - pe[2]};
        int steps = (int)ceilf(len3(d0) / step_size);
        const float d1[3] = {(exitp[0] - cur[0]) * scale[0], (exitp[1] - cur[1]) * scale[1], (exitp[2] - cur[2]) * scale[2]};
        const int s2 = (int)ceilf(len3(d1) / step_size);
        if (s2 < steps) steps = s2;
        float pc[3] = {pe[0], pe[1], pe[2]};
        for (int i = 0; i < steps; ++i) {
          const float mp[3] = {(pc[0] - trans[0]) / scale[0], (pc[1] - trans[1]) / scale[1], (pc[2] - trans[2]) / scale[2]};
          float s[4];
          color_from_volume(c, pc, mp, s);
          c->samples++;
          s[3] = 1.0f - powf(1.0f - s[3], oc);
          under(s, acc);
          if (acc[3] > 0.99f) {
            if (optimal) { pos[3] = 1000.0f; memcpy(col, acc, 16); }
            return;
          }
          for (int a = 0; a < 3; ++a) pc[a] += vdir[a];
        }
        for (int a = 0; a < 3; ++a) cur[a] = (pc[a] - trans[a]) / scale[a];
      } else {
        for (int a = 0; a < 3; ++a) cur[a] = nexit[a] + voxel_size * dir[a] / ray_len;
      }
      memcpy(last_bc, bc, sizeof(bc));
      const float dd[3] = {entry[0] - nexit[0], entry[1] - nexit[1], entry[2] - nexit[2]};
      t = len3(dd) / ray_len;
      if (t > 0.9999f) break;
    }
  }
  if (optimal) { pos[3] = 1000.0f; memcpy(col, acc, 16); }
}

/* GLGridLeaper-iso.glsl main() + Method-iso.glsl.  pos/nrm: in = rayStartPoint /
 * rayStartNormal, out = rayResumePos / rayResumeNormal; hit/hitn = MRT 0/1. */
static void march_ray_iso(ctx_t* c, const float exitp[4], float pos[4], float nrm[4], float hit[4], float hitn[4]) {
  const tvk_raycast_params* p = c->p;
  memset(hit, 0, 16);
  memset(hitn, 0, 16);
  if (floorf(pos[3]) == 1000.0f) return;
  if (floorf(pos[3]) == 500.0f) {
    float h4[4];
    xform_point(p->model_to_eye, pos, 1.0f, h4);
    hit[0] = h4[0]; hit[1] = h4[1]; hit[2] = h4[2];
    hit[3] = pos[3] - floorf(pos[3]) + 1.0f;
    memcpy(hitn, nrm, 16);
    return;
  }
  const float entry[3] = {pos[0], pos[1], pos[2]};
  const float entry_depth = pos[3], exit_depth = exitp[3];
  const float dir[3] = {exitp[0] - entry[0], exitp[1] - entry[1], exitp[2] - entry[2]};
  const float ray_len = len3(dir);
  float vdir[3] = {dir[0] * (float)c->pool->volume_size[0], dir[1] * (float)c->pool->volume_size[1], dir[2] * (float)c->pool->volume_size[2]};
  const float vl = len3(vdir);
  for (int a = 0; a < 3; ++a) vdir[a] = vdir[a] / vl / c->pool_size[a] / (2.0f * p->sample_rate_modifier);
  const float step_size = len3(vdir);
  float t = 0.0f;
  int optimal = 1;
  const float voxel_size = 0.125f / 2000.0f;
  float cur[3] = {entry[0], entry[1], entry[2]};
  uint32_t last_bc[4] = {0, 0, 0, 9999};
  if (ray_len > voxel_size) {
    for (uint32_t j = 0; j < 100; ++j) {
      const float depth = fmaf(exit_depth, t, entry_depth * (1.0f - t));
      uint32_t lod = compute_lod(depth, p->lod_factor, p->level_zero_wse, c->L.lod_count - 1);
      float pe[3], px[3], nexit[3], scale[3], trans[3];
      uint32_t bc[4];
      int empty;
      const int ok = get_brick(c, cur, &lod, dir, pe, px, nexit, &empty, scale, trans, bc);
      if (!ok && optimal) {
        optimal = 0;
        pos[0] = cur[0]; pos[1] = cur[1]; pos[2] = cur[2]; pos[3] = depth;
      }
      if (!empty && memcmp(last_bc, bc, sizeof(bc)) != 0) {
        const float d0[3] = {px[0] - pe[0], px[1] - pe[1], px[2] - pe[2]};
        int steps = (int)ceilf(len3(d0) / step_size);
        const float d1[3] = {(exitp[0] - cur[0]) * scale[0], (exitp[1] - cur[1]) * scale[1], (exitp[2] - cur[2]) * scale[2]};
        const int s2 = (int)ceilf(len3(d1) / step_size);
        if (s2 < steps) steps = s2;
        float pc[3] = {pe[0], pe[1], pe[2]};
        for (int i = 0; i < steps; ++i) {
          c->samples++;
          if (sample_pool(c, pc) >= p->isoval) {
            /* RefineIsosurface */
            float rd[3] = {vdir[0] / 2.0f, vdir[1] / 2.0f, vdir[2] / 2.0f};
            for (int a = 0; a < 3; ++a) pc[a] -= rd[a];
            for (int k = 0; k < 5; ++k) {
              for (int a = 0; a < 3; ++a) rd[a] /= 2.0f;
              if (sample_pool(c, pc) >= p->isoval) { for (int a = 0; a < 3; ++a) pc[a] -= rd[a]; }
              else { for (int a = 0; a < 3; ++a) pc[a] += rd[a]; }
            }
            for (int a = 0; a < 3; ++a) cur[a] = (pc[a] - trans[a]) / scale[a];
            float h4[4];
            xform_point(p->model_to_eye, cur, 1.0f, h4);
            hit[0] = h4[0]; hit[1] = h4[1]; hit[2] = h4[2];
            hit[3] = 1.0f + 1.0f;                      /* color.r + 1.0, color = (1,1,1,1) */
            float g[3], n[3];
            gradient(c, pc, g);
            for (int a = 0; a < 3; ++a) n[a] = g[a] * p->domain_scale[a];
            const float l = len3(n);
            if (l > 0.0f) for (int a = 0; a < 3; ++a) n[a] /= l;
            float n4[4];
            xform_point(p->model_view_it, n, 0.0f, n4);
            hitn[0] = n4[0]; hitn[1] = n4[1]; hitn[2] = n4[2];
            hitn[3] = floorf(1.0f * 512.0f) + 1.0f;    /* floor(color.g*512)+color.b */
            /* TerminateRay (iso.glsl:52-58): rayResumePos.xyz is NOT updated, only w.
             * With color = (1,1,1,1) this stores w = 501, so a later pass re-marches
             * the ray from the same start point (reference quirk, kept). */
            if (optimal) pos[3] = (hit[3] == 0.0f) ? 1000.0f : 499.0f + hit[3];
            memcpy(nrm, hitn, 16);
            return;
          }
          for (int a = 0; a < 3; ++a) pc[a] += vdir[a];
        }
        for (int a = 0; a < 3; ++a) cur[a] = (pc[a] - trans[a]) / scale[a];
      } else {
        for (int a = 0; a < 3; ++a) cur[a] = nexit[a] + voxel_size * dir[a] / ray_len;
      }
      memcpy(last_bc, bc, sizeof(bc));
      const float dd[3] = {entry[0] - nexit[0], entry[1] - nexit[1], entry[2] - nexit[2]};
      t = len3(dd) / ray_len;
      if (t > 0.9999f) break;
    }
  }
  if (optimal) pos[3] = 1000.0f;     /* rayHitPos.a == 0 */
  memcpy(nrm, hitn, 16);
}

/* Compose-FS.glsl:47-75 deferred shading of the iso hit buffers, uniforms as set by
 * GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2811): vLightDiffuse =
 * diffuse*isoColor, vLightDir = m_vLightDir (eye space).  Only vNormal.z is made
 * positive and the normal is NOT re-normalised, as in the shader. */
static void compose_iso(const tvk_raycast_params* p, const float hit[4], const float hitn[4], float out[4]) {
  if (hit[3] == 0.0f) { memset(out, 0, 16); return; }   /* discard -> cleared target */
  const float n[3] = {hitn[0], hitn[1], fabsf(hitn[2])};
  float v[3] = {0.0f - hit[0], 0.0f - hit[1], 0.0f - hit[2]};
  float l = len3(v);
  for (int a = 0; a < 3; ++a) v[a] /= l;
  const float d2 = 2.0f * dot3(n, v);
  float r[3] = {fmaf(-d2, n[0], v[0]), fmaf(-d2, n[1], v[1]), fmaf(-d2, n[2], v[2])};
  l = len3(r);
  for (int a = 0; a < 3; ++a) r[a] /= l;
  const float* L = p->compose_light_dir;
  const float negL[3] = {-L[0], -L[1], -L[2]};
  const float ndl = fmaxf(fabsf(dot3(n, negL)), 0.0f);
  const float spec = pow8(fmaxf(dot3(r, L), 0.0f));
  for (int a = 0; a < 3; ++a)
    out[a] = clampf(fmaf(p->light_specular[a], spec, fmaf(p->light_diffuse[a] * p->iso_color[a], ndl, p->light_ambient[a])), 0.0f, 1.0f);
  out[3] = 1.0f;
}

/* One sub-frame (GLGridLeaper::Render3DRegion -> Raycast).  ray_pos / ray_aux are
 * the W*H*4 resume buffers (in/out; ignored on input when new_view), image the
 * W*H*4 result, hash the request table (cleared here), stats[4] = samples,
 * GetBrick calls, rays, dropped brick reports (table full within the probe limit). */
int orc_raycast_rect(const tvk_raycast_params* p, const orc_pool_t* pool, const uint8_t* tf,
                     uint32_t tf_w, uint32_t tf_h, float* ray_pos, float* ray_aux, float* image,
                     uint32_t* hash, uint64_t* stats, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                     int clear_hash);

int orc_raycast(const tvk_raycast_params* p, const orc_pool_t* pool, const uint8_t* tf,
                uint32_t tf_w, uint32_t tf_h, float* ray_pos, float* ray_aux, float* image,
                uint32_t* hash, uint64_t* stats) {
  return orc_raycast_rect(p, pool, tf, tf_w, tf_h, ray_pos, ray_aux, image, hash, stats, 0, 0, p->width, p->height, 1);
}

/* The same sub-frame restricted to the pixel rectangle [x0,x1) x [y0,y1) of the W x H target (the buffers stay
 * full-frame; pixels outside the rectangle are not touched).  Used to check crops of frames whose full CPU render
 * would take too long.  clear_hash == 0 accumulates the requests of several rectangles in one table; stats are
 * those of this call. */
int orc_raycast_rect(const tvk_raycast_params* p, const orc_pool_t* pool, const uint8_t* tf,
                     uint32_t tf_w, uint32_t tf_h, float* ray_pos, float* ray_aux, float* image,
                     uint32_t* hash, uint64_t* stats, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                     int clear_hash) {
  uint32_t inner[3] = {pool->max_brick[0] - 2 * pool->overlap, pool->max_brick[1] - 2 * pool->overlap,
                       pool->max_brick[2] - 2 * pool->overlap};
  if (x1 > p->width) x1 = p->width;
  if (y1 > p->height) y1 = p->height;
  if (clear_hash) memset(hash, 0, sizeof(uint32_t) * p->hash_table_size);
  uint64_t tot_s = 0, tot_b = 0, tot_r = 0, tot_d = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot_s, tot_b, tot_r, tot_d)
  for (uint32_t py = y0; py < y1; ++py) {
    ctx_t c;
    memset(&c, 0, sizeof(c));
    c.p = p; c.pool = pool; c.tf = tf; c.tf_w = tf_w; c.tf_h = tf_h; c.hash = hash;
    pool_layout(pool->volume_size, inner, pool->lod_count, &c.L);
    for (int a = 0; a < 3; ++a) {
      c.pool_size[a] = (float)(pool->capacity[a] * pool->max_brick[a]);
      c.overlap_tc[a] = (float)(pool->max_brick[a] - inner[a]) / (2.0f * c.pool_size[a]);
    }
    for (uint32_t px = x0; px < x1; ++px) {
      const size_t i = 4 * ((size_t)py * p->width + px);
      float entry[4], exitp[4];
      const int hit_box = ray_setup(p, px, py, entry, exitp);
      if (!hit_box) {
        /* the bbox back faces do not cover this pixel: the shader never runs */
        memset(image + i, 0, 16);
        if (p->new_view) { memset(ray_aux + i, 0, 16); memset(ray_pos + i, 0, 16); ray_pos[i + 3] = 1000.0f; }
        continue;
      }
      if (p->new_view) { memcpy(ray_pos + i, entry, 16); memset(ray_aux + i, 0, 16); }
      c.rays++;
      if (p->render_mode == TVK_RM_ISOSURFACE) {
        float h[4], hn[4];
        march_ray_iso(&c, exitp, ray_pos + i, ray_aux + i, h, hn);
        compose_iso(p, h, hn, image + i);
      } else {
        march_ray(&c, exitp, ray_pos + i, ray_aux + i, image + i);
      }
    }
    tot_s += c.samples; tot_b += c.bricks; tot_r += c.rays; tot_d += c.drops;
  }
  if (stats) { stats[0] = tot_s; stats[1] = tot_b; stats[2] = tot_r; stats[3] = tot_d; }
  return 0;
}

/* GLHashTable::GetData + Int2Vector, GLHashTable.cpp:65-77,90-104 */
uint32_t orc_read_requests(const uint32_t* hash, uint32_t table_size, const uint32_t l0_bricks[3],
                           uint32_t* xyzl, uint32_t cap) {
  uint32_t n = 0;
  const uint32_t vol = l0_bricks[0] * l0_bricks[1] * l0_bricks[2];
  for (uint32_t i = 0; i < table_size && n < cap; ++i) {
    if (!hash[i]) continue;
    uint32_t idx = hash[i] - 1;
    uint32_t* o = xyzl + 4 * n++;
    o[3] = idx / vol; idx -= o[3] * vol;
    o[2] = idx / (l0_bricks[0] * l0_bricks[1]); idx -= o[2] * (l0_bricks[0] * l0_bricks[1]);
    o[1] = idx / l0_bricks[0]; idx -= o[1] * l0_bricks[0];
    o[0] = idx;
  }
  return n;
}

/* ContainsData<mode>, GLVolumePool.cpp:963-989 */
static int contains_data(const tvk_visibility* v, const double* mm, const double* gmm, uint32_t i) {
  switch (v->render_mode) {
    case TVK_RM_1DTRANS: return v->smax >= mm[2 * i] && v->smin <= mm[2 * i + 1];
    case TVK_RM_2DTRANS:
      return (v->smax >= mm[2 * i] && v->smin <= mm[2 * i + 1]) &&
             (gmm ? (v->gmax >= gmm[2 * i] && v->gmin <= gmm[2 * i + 1]) : (v->gmax >= -DBL_MAX && v->gmin <= DBL_MAX));
    default: return v->smin >= mm[2 * i] && v->smin <= mm[2 * i + 1];
  }
}

/* GLVolumePool::RecomputeVisibility (:1580-1718) for a pool whose resident
 * bricks are given by the entries of `meta` that are >= FLAG_COUNT on input:
 * reset to MISSING, RecomputeVisibilityForBrickPool (:992-1018) for the
 * resident ones, then RecomputeVisibilityForOctree (:1021-1359).
 * counts[4] = vEmptyBrickCount (processed, EMPTY, CHILD_EMPTY inner, leaf empty). */
void orc_visibility(const uint32_t vol[3], const uint32_t max_brick[3], uint32_t overlap,
                    uint32_t lod_count, const tvk_visibility* vis, const double* minmax,
                    const double* grad_minmax, uint32_t* meta, uint32_t* counts) {
  uint32_t inner[3] = {max_brick[0] - 2 * overlap, max_brick[1] - 2 * overlap, max_brick[2] - 2 * overlap};
  layout_t L;
  pool_layout(vol, inner, lod_count, &L);
  uint32_t cnt[4] = {0, 0, 0, 0};
  for (uint32_t i = 0; i < L.total_bricks; ++i) {
    if (meta[i] >= TVK_BI_FLAG_COUNT) {
      if (!contains_data(vis, minmax, grad_minmax, i)) meta[i] = TVK_BI_EMPTY;
    } else meta[i] = TVK_BI_MISSING;
  }
#define IDX(x, y, z, l) (L.lod_offset[l] + (x) + (y) * L.lod_bricks[l][0] + (z) * L.lod_bricks[l][0] * L.lod_bricks[l][1])
  for (uint32_t z = 0; z < L.lod_bricks[0][2]; ++z)
    for (uint32_t y = 0; y < L.lod_bricks[0][1]; ++y)
      for (uint32_t x = 0; x < L.lod_bricks[0][0]; ++x) {
        cnt[0]++;
        const uint32_t i = IDX(x, y, z, 0);
        if (meta[i] < TVK_BI_FLAG_COUNT && !contains_data(vis, minmax, grad_minmax, i)) { meta[i] = TVK_BI_CHILD_EMPTY; cnt[3]++; }
      }
  for (uint32_t l = 1; l < lod_count; ++l) {
    /* the reference enumerates the even core, then the odd planes/lines/corner
     * (:1063-1352); every parent tests exactly its EXISTING children, which is
     * what the loops below do in one sweep. */
    for (uint32_t z = 0; z < L.lod_bricks[l][2]; ++z)
      for (uint32_t y = 0; y < L.lod_bricks[l][1]; ++y)
        for (uint32_t x = 0; x < L.lod_bricks[l][0]; ++x) {
          cnt[0]++;
          const uint32_t i = IDX(x, y, z, l);
          if (meta[i] >= TVK_BI_FLAG_COUNT) continue;
          if (contains_data(vis, minmax, grad_minmax, i)) continue;
          int all_child_empty = 1;
          for (uint32_t dz = 0; dz < 2; ++dz)
            for (uint32_t dy = 0; dy < 2; ++dy)
              for (uint32_t dx = 0; dx < 2; ++dx) {
                const uint32_t cx = 2 * x + dx, cy = 2 * y + dy, cz = 2 * z + dz;
                if (cx >= L.lod_bricks[l - 1][0] || cy >= L.lod_bricks[l - 1][1] || cz >= L.lod_bricks[l - 1][2]) continue;
                if (meta[IDX(cx, cy, cz, l - 1)] != TVK_BI_CHILD_EMPTY) all_child_empty = 0;
              }
          if (all_child_empty) { meta[i] = TVK_BI_CHILD_EMPTY; cnt[2]++; }
          else { meta[i] = TVK_BI_EMPTY; cnt[1]++; }
        }
  }
#undef IDX
  if (counts) memcpy(counts, cnt, sizeof(cnt));
}

/* Ordered "under" blend of premultiplied partial images, front to back
 * (the sort-last compositor's arithmetic; Compositing.glsl:33-38 specialised
 * to premultiplied sources: dst += (1-dst.a)*src). */
void orc_composite_ordered(const float* const* partials, uint32_t n, uint64_t pixels, float* out) {
  for (uint64_t i = 0; i < pixels; ++i) {
    float d[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < n; ++k) {
      const float* s = partials[k] + 4 * i;
      const float w = 1.0f - d[3];
      d[0] += w * s[0]; d[1] += w * s[1]; d[2] += w * s[2]; d[3] += w * s[3];
    }
    memcpy(out + 4 * i, d, 16);
  }
}
