// Classic per-brick ray-caster (see tvk_raycast_classic in include/tvk_cuda.h for the
// reference functions and shaders it replaces).  One thread per pixel walks the host-sorted brick list of one
// level of detail front to back: analytic ray / brick-box intersection instead of the two rasterised passes
// per brick, the fragment shader's march inside the brick, and the frame buffer's "under" blend in registers --
// one launch per frame instead of 2 draws + glFinish per brick.  Floating-point contract as in tvk_raycast.cu
// (-fmad=false, explicit fmaf where the oracle has it).
#include "tvk_internal.h"
#include <cuda_fp16.h>

struct ClassicArgs {
  tvk_classic_params p;
  const tvk_classic_brick* bricks;
  uint32_t nbricks;
  const uint8_t* store;
  const float4* tf_f;                // 1D table as floats
  const uchar4* tf;                  // RGBA8 table (2D: [tf_h][tf_w])
  const uint32_t* tf_colmask;        // bit x: column x has a non-zero alpha in some row
  uint32_t tf_w, tf_h;
  float4* image;
  unsigned long long* stats;
  // ClearView (isosurface mode only): second surface per brick, hit buffers instead of the fused composition
  uint32_t cv_on;
  tvk_clearview_params cv;
  float4* cvb[4];                    // m_pFBOIsoHit buffers 0 / 1 (fp32), m_pFBOCVHit buffers 0 / 1 (half precision)
};

__device__ __forceinline__ float c_dot3(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, ax * bx)); }
__device__ __forceinline__ float c_clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
// v' = (v, w) * M, FLOATMATRIX4 memory order
__device__ __forceinline__ float4 c_xform(const float* m, float x, float y, float z, float w) {
  float4 o;
  o.x = fmaf(w, m[12], fmaf(z, m[8], fmaf(y, m[4], x * m[0])));
  o.y = fmaf(w, m[13], fmaf(z, m[9], fmaf(y, m[5], x * m[1])));
  o.z = fmaf(w, m[14], fmaf(z, m[10], fmaf(y, m[6], x * m[2])));
  o.w = fmaf(w, m[15], fmaf(z, m[11], fmaf(y, m[7], x * m[3])));
  return o;
}

// texture3D(texVolume, t).x on the brick's own texture: GL_LINEAR, CLAMP_TO_EDGE, UNORM8/16
template <int BPV>
__device__ __forceinline__ float brick_sample(const uint8_t* __restrict__ data, const uint32_t vox[3], float tx, float ty, float tz) {
  const float u = fmaf(tx, (float)vox[0], -0.5f), v = fmaf(ty, (float)vox[1], -0.5f), w = fmaf(tz, (float)vox[2], -0.5f);
  const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
  const float a = u - fu, b = v - fv, g = w - fw;
  const int mx = (int)vox[0] - 1, my = (int)vox[1] - 1, mz = (int)vox[2] - 1;
  // (int) of a float beyond the int range is clamped by the conversion; NaN -> 0
  const int x0 = min(max((int)fu, 0), mx), x1 = min(max((int)fu + 1, 0), mx);
  const int y0 = min(max((int)fv, 0), my), y1 = min(max((int)fv + 1, 0), my);
  const int z0 = min(max((int)fw, 0), mz), z1 = min(max((int)fw + 1, 0), mz);
  const size_t sx = vox[0], sxy = (size_t)vox[0] * vox[1];
#define TVK_VOX(X, Y, Z) (BPV == 1 ? (float)data[(X) + (Y) * sx + (Z) * sxy] : (float)reinterpret_cast<const uint16_t*>(data)[(X) + (Y) * sx + (Z) * sxy])
  const float v000 = TVK_VOX(x0, y0, z0), v100 = TVK_VOX(x1, y0, z0), v010 = TVK_VOX(x0, y1, z0), v110 = TVK_VOX(x1, y1, z0);
  const float v001 = TVK_VOX(x0, y0, z1), v101 = TVK_VOX(x1, y0, z1), v011 = TVK_VOX(x0, y1, z1), v111 = TVK_VOX(x1, y1, z1);
#undef TVK_VOX
  const float c00 = fmaf(a, v100 - v000, v000), c10 = fmaf(a, v110 - v010, v010);
  const float c01 = fmaf(a, v101 - v001, v001), c11 = fmaf(a, v111 - v011, v011);
  const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
  return fmaf(g, c1 - c0, c0) * (BPV == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}

// ComputeGradient, Volume3D.glsl:43-53 (y sign flipped there)
template <int BPV>
__device__ __forceinline__ void brick_gradient(const uint8_t* __restrict__ data, const uint32_t vox[3], float tx, float ty, float tz, float g[3]) {
  const float hx = 1.0f / (float)vox[0], hy = 1.0f / (float)vox[1], hz = 1.0f / (float)vox[2];
  const float gxp = brick_sample<BPV>(data, vox, tx + hx, ty, tz), gxm = brick_sample<BPV>(data, vox, tx - hx, ty, tz);
  const float gyp = brick_sample<BPV>(data, vox, tx, ty - hy, tz), gym = brick_sample<BPV>(data, vox, tx, ty + hy, tz);
  const float gzp = brick_sample<BPV>(data, vox, tx, ty, tz + hz), gzm = brick_sample<BPV>(data, vox, tx, ty, tz - hz);
  g[0] = (gxm - gxp) / 2.0f; g[1] = (gyp - gym) / 2.0f; g[2] = (gzm - gzp) / 2.0f;
}

// gl_NormalMatrix * (g * vDomainScale), "secure" normalisation (Volume3D.glsl:55-60, GLRaycaster-*-light-FS.glsl)
__device__ __forceinline__ void classic_normal(const tvk_classic_params& p, const float g[3], float& Nx, float& Ny, float& Nz) {
  const float gx = g[0] * p.domain_scale[0], gy = g[1] * p.domain_scale[1], gz = g[2] * p.domain_scale[2];
  Nx = fmaf(gz, p.normal_matrix[6], fmaf(gy, p.normal_matrix[3], gx * p.normal_matrix[0]));
  Ny = fmaf(gz, p.normal_matrix[7], fmaf(gy, p.normal_matrix[4], gx * p.normal_matrix[1]));
  Nz = fmaf(gz, p.normal_matrix[8], fmaf(gy, p.normal_matrix[5], gx * p.normal_matrix[2]));
  const float nl = sqrtf(c_dot3(Nx, Ny, Nz, Nx, Ny, Nz));
  if (nl > 0.0f) { Nx /= nl; Ny /= nl; Nz /= nl; }
}

__device__ __forceinline__ int c_nearest(float s, uint32_t n) { return min(max(__float2int_rz(s * (float)n), 0), (int)n - 1); }

// Compose-FS.glsl:49-75 with the uniforms of GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2811)
__device__ __forceinline__ float4 classic_compose(const tvk_classic_params& p, float4 hit, float nx, float ny, float nz) {
  if (hit.w == 0.0f) return make_float4(0.f, 0.f, 0.f, 0.f);
  nz = fabsf(nz);
  float vx = 0.0f - hit.x, vy = 0.0f - hit.y, vz = 0.0f - hit.z;
  float l = sqrtf(c_dot3(vx, vy, vz, vx, vy, vz));
  vx /= l; vy /= l; vz /= l;
  const float d2 = 2.0f * c_dot3(nx, ny, nz, vx, vy, vz);
  float rx = fmaf(-d2, nx, vx), ry = fmaf(-d2, ny, vy), rz = fmaf(-d2, nz, vz);
  l = sqrtf(c_dot3(rx, ry, rz, rx, ry, rz));
  rx /= l; ry /= l; rz /= l;
  const float ndl = fmaxf(fabsf(c_dot3(nx, ny, nz, -p.light_dir[0], -p.light_dir[1], -p.light_dir[2])), 0.0f);
  float sp = fmaxf(c_dot3(rx, ry, rz, p.light_dir[0], p.light_dir[1], p.light_dir[2]), 0.0f);
  sp *= sp; sp *= sp; sp *= sp;
  return make_float4(c_clamp01(fmaf(p.light_specular[0], sp, fmaf(p.light_diffuse[0] * p.iso_color[0], ndl, p.light_ambient[0]))),
                     c_clamp01(fmaf(p.light_specular[1], sp, fmaf(p.light_diffuse[1] * p.iso_color[1], ndl, p.light_ambient[1]))),
                     c_clamp01(fmaf(p.light_specular[2], sp, fmaf(p.light_diffuse[2] * p.iso_color[2], ndl, p.light_ambient[2]))), 1.0f);
}

// Per-tile brick culling.  A warp owns an 8x4 pixel tile; the rays of its pixels lie inside the bundle spanned by the four
// corner rays (see tile_corner_pixel).  A brick entirely outside one of the bundle's four side planes is missed by every
// ray of the tile: the warp tests the brick list against those planes 32 bricks at a time and compacts the survivors -- IN
// LIST ORDER, the blend between bricks depends on it -- into shared memory, and every pixel then walks the tile's short list
// instead of the whole frame's (1024^3 in 64^3 bricks: 5832 bricks per frame, a few dozen per tile).  Conservative: the
// planes are pushed out by a relative epsilon, and a pixel still does its own exact slab test.
#define TVK_TILE_LIST 256
struct TilePlanes { float n[4][3]; float d[4]; float r_scale[4]; };

// co / cd: origin and direction of the bundle's four corner rays, going round the tile
__device__ __forceinline__ void tile_planes(const float co[4][3], const float cd[4][3], bool common_origin, TilePlanes& tp) {
  float cp[3];                                             // a point well inside the bundle
#pragma unroll
  for (int a = 0; a < 3; ++a) cp[a] = 0.25f * (co[0][a] + co[1][a] + co[2][a] + co[3][a]) + 0.125f * (cd[0][a] + cd[1][a] + cd[2][a] + cd[3][a]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int k1 = (k + 1) & 3;
    // the side face holds corner k's ray and corner k+1's: its second direction is the other corner's ray direction when the
    // rays share their origin (perspective from the eye), else the edge between the two origins
    float e[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a] = common_origin ? cd[k1][a] : co[k1][a] - co[k][a];
    float n[3] = {cd[k][1] * e[2] - cd[k][2] * e[1], cd[k][2] * e[0] - cd[k][0] * e[2], cd[k][0] * e[1] - cd[k][1] * e[0]};
    const float side = n[0] * (cp[0] - co[k][0]) + n[1] * (cp[1] - co[k][1]) + n[2] * (cp[2] - co[k][2]);
    const float sg = side < 0.0f ? -1.0f : 1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) tp.n[k][a] = sg * n[a];
    tp.d[k] = -(tp.n[k][0] * co[k][0] + tp.n[k][1] * co[k][1] + tp.n[k][2] * co[k][2]);
    tp.r_scale[k] = 1e-5f * sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  }
}
// The corner rays are those of the pixels ONE PIXEL OUTSIDE the tile (x0-1 / x0+8, y0-1 / y0+4): the cross product of two
// nearly parallel ray directions is good to ~1e-4 of a radian at 4K, a pixel is ~8e-4 -- the planes keep a pixel of margin.
__device__ __forceinline__ void tile_corner_pixel(uint32_t tile_x0, uint32_t tile_y0, int k, float& fx, float& fy) {
  fx = (float)tile_x0 + ((k == 1 || k == 2) ? 8.5f : -0.5f);
  fy = (float)tile_y0 + ((k >= 2) ? 4.5f : -0.5f);
}
__device__ __forceinline__ bool tile_may_hit(const TilePlanes& tp, const tvk_classic_brick& b) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float sdist = tp.n[k][0] * b.center[0] + tp.n[k][1] * b.center[1] + tp.n[k][2] * b.center[2] + tp.d[k];
    const float r = 0.5f * (fabsf(tp.n[k][0]) * b.extension[0] + fabsf(tp.n[k][1]) * b.extension[1] + fabsf(tp.n[k][2]) * b.extension[2]);
    if (sdist < -(r * 1.001f + tp.r_scale[k])) return false;
  }
  return true;
}
// the next chunk of the tile's brick list: bricks [base, ...) that may be hit, in list order; returns their number
__device__ __forceinline__ uint32_t tile_list_chunk(const TilePlanes& tp, const tvk_classic_brick* bricks, uint32_t nbricks, uint32_t& base,
                                                    uint32_t* list, uint32_t lane) {
  uint32_t count = 0;
  while (base < nbricks && count + 32u <= TVK_TILE_LIST) {
    const uint32_t bi = base + lane;
    const bool in = bi < nbricks && tile_may_hit(tp, bricks[bi]);
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, in);
    if (in) list[count + __popc(m & ((1u << lane) - 1u))] = bi;
    count += __popc(m);
    base += 32u;
  }
  __syncwarp();
  return count;
}

template <int BPV, int MODE, bool LIT>
__global__ void __launch_bounds__(128) classic_kernel(const __grid_constant__ ClassicArgs a) {
  // 8x4 pixel tiles per warp, as in K1, so that neighbouring rays share cache lines
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t tiles_x = (a.p.width + 7) / 8;
  const uint32_t px = (warp % tiles_x) * 8u + (lane & 7u), py = (warp / tiles_x) * 4u + (lane >> 3);
  if ((warp / tiles_x) * 4u >= a.p.height) return;         // (whole warps only: the lanes of a tile cull its bricks together)
  const bool alive = px < a.p.width && py < a.p.height;    // pixels beyond the right / top edge still span the tile's ray bundle
  __shared__ uint32_t tile_list[4][TVK_TILE_LIST];
  uint32_t* const list = tile_list[threadIdx.x >> 5];
  const tvk_classic_params& p = a.p;
  const float nx = fmaf(((float)px + 0.5f) / (float)p.width, 2.0f, -1.0f), ny = fmaf(((float)py + 0.5f) / (float)p.height, 2.0f, -1.0f);
  const float4 n4 = c_xform(p.inv_projection, nx, ny, -1.0f, 1.0f);
  const float ex = n4.x / n4.w, ey = n4.y / n4.w, ez = n4.z / n4.w;          // near-plane point in eye space
  const float4 O = c_xform(p.inv_model_view, 0.f, 0.f, 0.f, 1.0f);
  const float4 D = c_xform(p.inv_model_view, ex, ey, ez, 0.0f);
  const float o[3] = {O.x, O.y, O.z}, d[3] = {D.x, D.y, D.z};
  TilePlanes tp;
  {
    float co[4][3], cd[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float fx, fy;
      tile_corner_pixel((warp % tiles_x) * 8u, (warp / tiles_x) * 4u, k, fx, fy);
      const float4 c4 = c_xform(p.inv_projection, fmaf(fx / (float)p.width, 2.0f, -1.0f), fmaf(fy / (float)p.height, 2.0f, -1.0f), -1.0f, 1.0f);
      const float4 cD = c_xform(p.inv_model_view, c4.x / c4.w, c4.y / c4.w, c4.z / c4.w, 0.0f);
      co[k][0] = o[0]; co[k][1] = o[1]; co[k][2] = o[2];
      cd[k][0] = cD.x; cd[k][1] = cD.y; cd[k][2] = cD.z;
    }
    tile_planes(co, cd, true, tp);
  }
  float4 dst = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long nsamples = 0;
  // isosurface: the two hit targets and the depth buffer (cleared to 0 / 1.0, depth test GL_LESS)
  float4 hit = make_float4(0.f, 0.f, 0.f, 0.f);
  float hnx = 0.f, hny = 0.f, hnz = 0.f, best_depth = 1.0f;
  float4 hit2 = make_float4(0.f, 0.f, 0.f, 0.f);            // ClearView: the same for the second surface
  float h2nx = 0.f, h2ny = 0.f, h2nz = 0.f, best_depth2 = 1.0f, best_tile = 0.0f;
  for (uint32_t base = 0; base < a.nbricks;) {
   const uint32_t in_tile = tile_list_chunk(tp, a.bricks, a.nbricks, base, list, lane);
   for (uint32_t li = 0; alive && li < in_tile; ++li) {
    const uint32_t bi = list[li];
    const tvk_classic_brick& b = a.bricks[bi];
    float s0 = -INFINITY, s1 = INFINITY;
    bool miss = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float lo = b.center[k] - b.extension[k] / 2.0f, hi = b.center[k] + b.extension[k] / 2.0f;
      if (d[k] == 0.0f) { if (o[k] < lo || o[k] > hi) miss = true; continue; }
      const float inv = 1.0f / d[k];
      float ta = (lo - o[k]) * inv, tb = (hi - o[k]) * inv;
      if (ta > tb) { const float t = ta; ta = tb; tb = t; }
      s0 = fmaxf(s0, ta); s1 = fminf(s1, tb);
    }
    if (p.clip_enabled && !miss) {                          // the box clipped by the plane (Clipper::BoxPlane), as in ray_setup of K1
      const float dn = c_dot3(p.clip_plane[0], p.clip_plane[1], p.clip_plane[2], d[0], d[1], d[2]);
      const float on = c_dot3(p.clip_plane[0], p.clip_plane[1], p.clip_plane[2], o[0], o[1], o[2]) + p.clip_plane[3];
      if (dn == 0.0f) { if (on > 0.0f) miss = true; }
      else {
        const float t = -on / dn;
        if (dn > 0.0f) s1 = fminf(s1, t); else s0 = fmaxf(s0, t);
      }
    }
    const float se = fmaxf(s0, 1.0f);                       // front face, or the near plane if that lies behind it
    if (miss || !(s1 > se)) continue;
    // ray entry through the RGBA16F buffer, ray exit from the interpolated varying
    const float enx = __half2float(__float2half_rn(ex * se)), eny = __half2float(__float2half_rn(ey * se)), enz = __half2float(__float2half_rn(ez * se));
    const float xx = ex * s1, xy = ey * s1, xz = ez * s1;
    const float4 te = c_xform(b.eye_to_tex, enx, eny, enz, 1.0f), tx = c_xform(b.eye_to_tex, xx, xy, xz, 1.0f);
    float dx = xx - enx, dy = xy - eny, dz = xz - enz;
    const float len = sqrtf(c_dot3(dx, dy, dz, dx, dy, dz));
    const int steps = (int)(len / b.ray_step) + 1;
    const float q = len / b.ray_step;
    const float itx = (tx.x - te.x) / q, ity = (tx.y - te.y) / q, itz = (tx.z - te.z) / q;
    dx /= len; dy /= len; dz /= len;
    const float ix = dx * b.ray_step, iy = dy * b.ray_step, iz = dz * b.ray_step;
    const uint8_t* data = a.store + b.store_offset;
    float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
    float ptx = te.x, pty = te.y, ptz = te.z, pex = enx, pey = eny, pez = enz;
    if (MODE == TVK_RM_ISOSURFACE) {                        // GLRaycaster-ISO-FS.glsl:56-104
      bool found = false;
      for (int i = 0; i < steps; ++i) {
        const float v = brick_sample<BPV>(data, b.voxels, ptx, pty, ptz);
        nsamples++;
        if (v >= p.isoval) { found = true; break; }
        ptx += itx; pty += ity; ptz += itz;
      }
      if (found) {                                          // (else: discard)
        // RefineIsosurface.glsl:37-53
        float rx = itx / 2.0f, ry = ity / 2.0f, rz = itz / 2.0f;
        ptx -= rx; pty -= ry; ptz -= rz;
        for (int i = 0; i < 5; ++i) {
          rx /= 2.0f; ry /= 2.0f; rz /= 2.0f;
          const float v = brick_sample<BPV>(data, b.voxels, ptx, pty, ptz);
          if (v >= p.isoval) { ptx -= rx; pty -= ry; ptz -= rz; } else { ptx += rx; pty += ry; ptz += rz; }
        }
        const float ltx = tx.x - te.x, lty = tx.y - te.y, ltz = tx.z - te.z;
        const float len_tex = sqrtf(c_dot3(ltx, lty, ltz, ltx, lty, ltz));
        const float hx = ptx - te.x, hy = pty - te.y, hz = ptz - te.z;
        const float f = sqrtf(c_dot3(hx, hy, hz, hx, hy, hz)) / len_tex;
        const float hpx = fmaf(xx, f, enx * (1.0f - f)), hpy = fmaf(xy, f, eny * (1.0f - f)), hpz = fmaf(xz, f, enz * (1.0f - f));
        const float depth = c_clamp01(p.proj_param[0] + p.proj_param[1] / (0.0f - hpz));
        if (depth < best_depth) {
          best_depth = depth;
          hit = make_float4(hpx, hpy, hpz, f);
          float g[3];
          brick_gradient<BPV>(data, b.voxels, ptx, pty, ptz, g);
          classic_normal(p, g, hnx, hny, hnz);
          best_tile = (float)bi;                            // gl_FragData[1].w = float(iTileID)
        }
      }
      if (a.cv_on) {
        // ClearView's second surface in the same brick (GLRaycaster::Render3DInLoop, GLRaycaster.cpp:438-452;
        // GLRaycaster-ISO-CV-FS.glsl:57-115).  If the first surface recorded for this pixel lies in THIS brick the search
        // starts there.  As in the shader, only the texture-space start moves: the step count and the step length still come
        // from the whole eye-space segment, and the hit position is interpolated between the brick's entry and exit.
        float sx = te.x, sy = te.y, sz = te.z;
        if ((float)bi == best_tile) {                       // (cleared buffer: tile 0 with a hit parameter of 0)
          const float pp = hit.w;
          sx = fmaf(tx.x, pp, te.x * (1.0f - pp)); sy = fmaf(tx.y, pp, te.y * (1.0f - pp)); sz = fmaf(tx.z, pp, te.z * (1.0f - pp));
        }
        const float l2x = tx.x - sx, l2y = tx.y - sy, l2z = tx.z - sz;
        const float len_tex2 = sqrtf(c_dot3(l2x, l2y, l2z, l2x, l2y, l2z));
        const float i2x = l2x / q, i2y = l2y / q, i2z = l2z / q;
        float cx = sx, cy = sy, cz = sz;
        bool found2 = false;
        for (int i = 0; i < steps; ++i) {
          const float v = brick_sample<BPV>(data, b.voxels, cx, cy, cz);
          nsamples++;
          if (v >= a.cv.cv_isoval) { found2 = true; break; }
          cx += i2x; cy += i2y; cz += i2z;
        }
        if (found2) {
          float rx = i2x / 2.0f, ry = i2y / 2.0f, rz = i2z / 2.0f;
          cx -= rx; cy -= ry; cz -= rz;
          for (int i = 0; i < 5; ++i) {
            rx /= 2.0f; ry /= 2.0f; rz /= 2.0f;
            const float v = brick_sample<BPV>(data, b.voxels, cx, cy, cz);
            if (v >= a.cv.cv_isoval) { cx -= rx; cy -= ry; cz -= rz; } else { cx += rx; cy += ry; cz += rz; }
          }
          const float hx = cx - sx, hy = cy - sy, hz = cz - sz;
          const float f2 = sqrtf(c_dot3(hx, hy, hz, hx, hy, hz)) / len_tex2;
          // "as the surface maybe transparent set depth to the ray exit pos"
          const float depth2 = c_clamp01(p.proj_param[0] + p.proj_param[1] / (0.0f - xz));
          if (depth2 < best_depth2) {
            best_depth2 = depth2;
            hit2 = make_float4(fmaf(xx, f2, enx * (1.0f - f2)), fmaf(xy, f2, eny * (1.0f - f2)), fmaf(xz, f2, enz * (1.0f - f2)), f2);
            float g[3];
            brick_gradient<BPV>(data, b.voxels, cx, cy, cz, g);
            classic_normal(p, g, h2nx, h2ny, h2nz);
          }
        }
      }
      continue;
    }
    for (int i = 0; i < steps; ++i) {
      const float v = brick_sample<BPV>(data, b.voxels, ptx, pty, ptz);
      // samples the transfer function makes fully transparent add nothing to the ray ((1-a)*0*rgb): their gradient
      // and shading are skipped (1D: alpha of the entry; 2D: a column that is transparent in every gradient row)
      float4 c;
      float g[3];
      bool shade = LIT;
      if (MODE == TVK_RM_2DTRANS) {                          // GLRaycaster-2D[-light]-FS.glsl
        const int cx = c_nearest(v * p.trans_scale, a.tf_w);
        if ((__ldg(a.tf_colmask + (cx >> 5)) >> (cx & 31)) & 1u) {
          brick_gradient<BPV>(data, b.voxels, ptx, pty, ptz, g);
          const float gm = sqrtf(c_dot3(g[0], g[1], g[2], g[0], g[1], g[2]));
          const uchar4 t = __ldg(a.tf + (size_t)c_nearest(fmaf(-gm, p.gradient_scale, 1.0f), a.tf_h) * a.tf_w + cx);
          const float k = 1.0f / 255.0f;
          c = make_float4((float)t.x * k, (float)t.y * k, (float)t.z * k, (float)t.w * k);
          shade = LIT && c.w != 0.0f;
        } else { c = make_float4(0.f, 0.f, 0.f, 0.f); shade = false; }
      } else {
        const int ti = min(max(__float2int_rz(v * p.trans_scale * (float)a.tf_w), 0), (int)a.tf_w - 1);
        c = __ldg(a.tf_f + ti);
        shade = LIT && c.w != 0.0f;
        if (shade) brick_gradient<BPV>(data, b.voxels, ptx, pty, ptz, g);
      }
      nsamples++;
      if (LIT && shade) {
        float Nx, Ny, Nz;
        classic_normal(p, g, Nx, Ny, Nz);
        // Lighting(vec3(0), position, ...), lighting.glsl
        float vx = 0.0f - pex, vy = 0.0f - pey, vz = 0.0f - pez;
        float l = sqrtf(c_dot3(vx, vy, vz, vx, vy, vz));
        vx /= l; vy /= l; vz /= l;
        const float d2 = 2.0f * c_dot3(Nx, Ny, Nz, vx, vy, vz);
        float rx = fmaf(-d2, Nx, vx), ry = fmaf(-d2, Ny, vy), rz = fmaf(-d2, Nz, vz);
        l = sqrtf(c_dot3(rx, ry, rz, rx, ry, rz));
        rx /= l; ry /= l; rz /= l;
        const float ndl = fmaxf(fabsf(c_dot3(Nx, Ny, Nz, p.light_dir[0], p.light_dir[1], p.light_dir[2])), 0.0f);
        float sp = fmaxf(c_dot3(rx, ry, rz, p.light_dir[0], p.light_dir[1], p.light_dir[2]), 0.0f);
        sp *= sp; sp *= sp; sp *= sp;
        c.x = c_clamp01(fmaf(p.light_specular[0], sp, fmaf(p.light_diffuse[0] * c.x, ndl, p.light_ambient[0])));
        c.y = c_clamp01(fmaf(p.light_specular[1], sp, fmaf(p.light_diffuse[1] * c.y, ndl, p.light_ambient[1])));
        c.z = c_clamp01(fmaf(p.light_specular[2], sp, fmaf(p.light_diffuse[2] * c.z, ndl, p.light_ambient[2])));
      }
      c.w = (p.step_scale == 1.0f) ? 1.0f - (1.0f - c.w) : 1.0f - powf(1.0f - c.w, p.step_scale);
      const float wgt = (1.0f - col.w) * c.w;             // UnderCompositing, Compositing.glsl:33-38
      col.x = fmaf(c.x, wgt, col.x); col.y = fmaf(c.y, wgt, col.y); col.z = fmaf(c.z, wgt, col.z);
      col.w = fmaf(c.w, 1.0f - col.w, col.w);
      pex += ix; pey += iy; pez += iz;
      ptx += itx; pty += ity; ptz += itz;
      if (col.w >= 0.99f) break;
    }
    const float k = 1.0f - dst.w;                          // frame buffer blend: ONE_MINUS_DST_ALPHA, ONE
    dst.x = fmaf(col.x, k, dst.x); dst.y = fmaf(col.y, k, dst.y); dst.z = fmaf(col.z, k, dst.z); dst.w = fmaf(col.w, k, dst.w);
   }
   __syncwarp();                                           // the list is rewritten by the next chunk
  }
  if (!alive) nsamples = 0;
  else if (MODE == TVK_RM_ISOSURFACE && a.cv_on) {
    // the four hit targets, composed by cv_compose_kernel (it reads the neighbours' normals); the ClearView targets are RGBA16F
    const size_t i = (size_t)py * p.width + px;
    a.cvb[0][i] = hit;
    a.cvb[1][i] = make_float4(hnx, hny, hnz, best_tile);
    a.cvb[2][i] = make_float4(__half2float(__float2half_rn(hit2.x)), __half2float(__float2half_rn(hit2.y)), __half2float(__float2half_rn(hit2.z)),
                              __half2float(__float2half_rn(hit2.w)));
    a.cvb[3][i] = make_float4(__half2float(__float2half_rn(h2nx)), __half2float(__float2half_rn(h2ny)), __half2float(__float2half_rn(h2nz)), 0.0f);
  } else {
    if (MODE == TVK_RM_ISOSURFACE) {
      if (a.cvb[0]) a.cvb[0][(size_t)py * p.width + px] = hit;     // m_pFBOIsoHit buffer 0: what Pick / the ClearView focus read back
      dst = classic_compose(p, hit, hnx, hny, hnz);
    }
    a.image[(size_t)py * p.width + px] = dst;
  }
  for (int s = 16; s > 0; s >>= 1) nsamples += __shfl_down_sync(0xFFFFFFFFu, nsamples, s);
  if (lane == 0 && a.stats) atomicAdd(a.stats, nsamples);
}

// Compose-CV-FS.glsl:52-108 with the uniforms of GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2776-2795)
__device__ __forceinline__ void cv_lighting(const tvk_classic_params& p, float4 pos, float nx, float ny, float nz, const float diffuse[3], float out[3]) {
  nz = fabsf(nz);
  float vx = 0.0f - pos.x, vy = 0.0f - pos.y, vz = 0.0f - pos.z;
  float l = sqrtf(c_dot3(vx, vy, vz, vx, vy, vz));
  vx /= l; vy /= l; vz /= l;
  const float d2 = 2.0f * c_dot3(nx, ny, nz, vx, vy, vz);
  float rx = fmaf(-d2, nx, vx), ry = fmaf(-d2, ny, vy), rz = fmaf(-d2, nz, vz);
  l = sqrtf(c_dot3(rx, ry, rz, rx, ry, rz));
  rx /= l; ry /= l; rz /= l;
  const float ndl = fmaxf(fabsf(c_dot3(nx, ny, nz, -p.light_dir[0], -p.light_dir[1], -p.light_dir[2])), 0.0f);
  float sp = fmaxf(c_dot3(rx, ry, rz, p.light_dir[0], p.light_dir[1], p.light_dir[2]), 0.0f);
  sp *= sp; sp *= sp; sp *= sp;
  for (int k = 0; k < 3; ++k) out[k] = c_clamp01(fmaf(p.light_specular[k], sp, fmaf(diffuse[k], ndl, p.light_ambient[k])));
}

__global__ void __launch_bounds__(256) cv_compose_kernel(const __grid_constant__ ClassicArgs a) {
  const tvk_classic_params& p = a.p;
  const uint32_t n = p.width * p.height;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = (int)(i / p.width), x = (int)(i - (uint32_t)y * p.width);
    const float4 pos = a.cvb[0][i];
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos.w != 0.0f) {                                    // (else: discard)
      const float4 nrm = a.cvb[1][i];
      const float d1[3] = {p.light_diffuse[0] * p.iso_color[0], p.light_diffuse[1] * p.iso_color[1], p.light_diffuse[2] * p.iso_color[2]};
      const float d2c[3] = {p.light_diffuse[0] * a.cv.cv_color[0], p.light_diffuse[1] * a.cv.cv_color[1], p.light_diffuse[2] * a.cv.cv_color[2]};
      float ctx[3];
      cv_lighting(p, pos, nrm.x, nrm.y, nrm.z, d1, ctx);
      // curvature estimate of the context surface: the normals one pixel away (GL_NEAREST, clamped at the border)
      const int xs[4] = {min(x + 1, (int)p.width - 1), max(x - 1, 0), x, x}, ys[4] = {y, y, min(y + 1, (int)p.height - 1), max(y - 1, 0)};
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int k = 0; k < 4; ++k) {
        const float4 nb = a.cvb[1][(size_t)ys[k] * p.width + xs[k]];
        sx += fabsf(nb.x - nrm.x); sy += fabsf(nb.y - nrm.y); sz += fabsf(nb.z - nrm.z);
      }
      const float curv = sqrtf(c_dot3(sx, sy, sz, sx, sy, sz));
      const float qx = pos.x - a.cv.pick_pos[0], qy = pos.y - a.cv.pick_pos[1], qz = pos.z - a.cv.pick_pos[2];
      const float dist_w = sqrtf(c_dot3(qx, qy, qz, qx, qy, qz)) * a.cv.size;
      const float blend = c_clamp01(fmaxf(curv * a.cv.context_scale, c_clamp01(dist_w)));
      const float4 pos2 = a.cvb[2][i];
      float foc[4] = {0.f, 0.f, 0.f, 0.f};
      if (pos2.w != 0.0f) {
        const float4 n2 = a.cvb[3][i];
        cv_lighting(p, pos2, n2.x, n2.y, n2.z, d2c, foc);
        foc[3] = 1.0f;
      }
      const float ib = 1.0f - blend;
      float col[4];
      for (int k = 0; k < 3; ++k) col[k] = c_clamp01(fmaf(foc[k], ib, ctx[k] * blend));
      col[3] = c_clamp01(fmaf(foc[3], ib, 1.0f * blend));
      const float border = 0.5f * (1.0f - c_clamp01(fabsf(dist_w - 1.0f) * a.cv.border_scale));
      out = make_float4(col[0] - border, col[1] - border, col[2] - border, col[3]);
    }
    a.image[i] = out;
  }
}

static int classic_launch(tvk_ctx* c, const tvk_classic_params* p, const tvk_clearview_params* cv, const tvk_classic_brick* bricks, uint32_t n,
                          const void* store);

extern "C" int tvk_raycast_classic(tvk_ctx* c, const tvk_classic_params* p, const tvk_classic_brick* bricks, uint32_t n, const void* store) {
  return classic_launch(c, p, nullptr, bricks, n, store);
}

extern "C" int tvk_raycast_classic_cv(tvk_ctx* c, const tvk_classic_params* p, const tvk_clearview_params* cv, const tvk_classic_brick* bricks,
                                      uint32_t n, const void* store) {
  if (!cv) return TVK_ERR_INVALID;
  if (p && p->render_mode != TVK_RM_ISOSURFACE) TVK_FAIL(c, TVK_ERR_INVALID, "ClearView belongs to the isosurface mode");
  return classic_launch(c, p, cv, bricks, n, store);
}

// the first surface's hit buffer of the last ClearView frame at one pixel (GLRenderer::CVFocusHasChanged reads m_pFBOIsoHit there)
extern "C" int tvk_read_iso_hit(tvk_ctx* c, uint32_t x, uint32_t y, float out[4]) {
  if (!c || !out || x >= c->width || y >= c->height) return TVK_ERR_INVALID;
  if (!c->iso_hits) TVK_FAIL(c, TVK_ERR_INVALID, "no isosurface frame has been drawn");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaMemcpyAsync(out, c->iso_hits + (size_t)y * c->width + x, sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TVK_OK;
}

static int classic_launch(tvk_ctx* c, const tvk_classic_params* p, const tvk_clearview_params* cv, const tvk_classic_brick* bricks, uint32_t n,
                          const void* store) {
  if (!c || !p || (n && !bricks) || !store) return TVK_ERR_INVALID;
  if (p->render_mode != TVK_RM_1DTRANS && p->render_mode != TVK_RM_2DTRANS && p->render_mode != TVK_RM_ISOSURFACE)
    TVK_FAIL(c, TVK_ERR_INVALID, "bad render mode");
  if (p->render_mode == TVK_RM_1DTRANS && (!c->d_tf || c->tf_h != 1)) TVK_FAIL(c, TVK_ERR_INVALID, "1D mode needs a 1D transfer function");
  if (p->render_mode == TVK_RM_2DTRANS && (!c->d_tf || c->tf_h < 2)) TVK_FAIL(c, TVK_ERR_INVALID, "2D mode needs a 2D transfer function");
  if (p->bytes_per_voxel != 1 && p->bytes_per_voxel != 2) TVK_FAIL(c, TVK_ERR_INVALID, "bad voxel size");
  if (!p->width || !p->height) TVK_FAIL(c, TVK_ERR_INVALID, "bad image size");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int rc = tvk_ensure_frame(c, p->width, p->height, c->hash_size ? c->hash_size : 509);
  if (rc) return rc;
  c->show_stereo = false;
  if (c->eye == 0) c->iso_hits = nullptr;
  tvk_classic_brick* d_bricks = nullptr;
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_bricks, sizeof(tvk_classic_brick) * (n ? n : 1), c->stream));
  if (n) TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_bricks, bricks, sizeof(tvk_classic_brick) * n, cudaMemcpyHostToDevice, c->stream));
  ClassicArgs a;
  a.p = *p; a.bricks = d_bricks; a.nbricks = n; a.store = static_cast<const uint8_t*>(store);
  a.tf_f = c->d_tf_f; a.tf = c->d_tf; a.tf_colmask = c->d_tf_colmask; a.tf_w = c->tf_w; a.tf_h = c->tf_h; a.image = c->d_image; a.stats = c->d_stats;
  a.cv_on = cv ? 1u : 0u;
  if (cv) { a.cv = *cv; a.cvb[0] = c->d_pos[0]; a.cvb[1] = c->d_aux[0]; a.cvb[2] = c->d_pos[1]; a.cvb[3] = c->d_aux[1]; }
  else {
    memset(&a.cv, 0, sizeof(a.cv));
    a.cvb[0] = p->render_mode == TVK_RM_ISOSURFACE ? c->d_pos[0] : nullptr;      // the first surface's hits are kept in any isosurface frame
    a.cvb[1] = a.cvb[2] = a.cvb[3] = nullptr;
  }
  if (c->eye == 0) c->iso_hits = p->render_mode == TVK_RM_ISOSURFACE ? c->d_pos[0] : nullptr;   // (m_pFBOIsoHit[0]: the first eye's)
  const uint32_t tiles = ((p->width + 7) / 8) * ((p->height + 3) / 4);
  const uint32_t blocks = (tiles + 3) / 4;
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  typedef void (*classic_fn)(const ClassicArgs);
  const bool lit = p->lighting != 0 && p->render_mode != TVK_RM_ISOSURFACE;
  classic_fn fn;
#define TVK_PICK(B) (p->render_mode == TVK_RM_ISOSURFACE ? classic_kernel<B, TVK_RM_ISOSURFACE, false> \
                     : p->render_mode == TVK_RM_2DTRANS ? (lit ? classic_kernel<B, TVK_RM_2DTRANS, true> : classic_kernel<B, TVK_RM_2DTRANS, false>) \
                                                        : (lit ? classic_kernel<B, TVK_RM_1DTRANS, true> : classic_kernel<B, TVK_RM_1DTRANS, false>))
  fn = (p->bytes_per_voxel == 1) ? TVK_PICK(1) : TVK_PICK(2);
#undef TVK_PICK
  fn<<<blocks, 128, 0, c->stream>>>(a);
  TVK_CHECK_CUDA(c, cudaGetLastError());
  if (cv) {
    int sm_count = 0;
    TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c->device));
    const size_t npx = (size_t)p->width * p->height;
    cv_compose_kernel<<<(unsigned)std::min<size_t>((npx + 255) / 256, (size_t)sm_count * 8), 256, 0, c->stream>>>(a);
    c->launches++;
    TVK_CHECK_CUDA(c, cudaGetLastError());
  }
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  c->launches++;
  c->k1_serial++;
  TVK_CHECK_CUDA(c, cudaFreeAsync(d_bricks, c->stream));
  return TVK_OK;
}


// ---------------------------------------------------------------------------------------------------------
// 2-D slice / MIP views (GLRenderer::Render2DView, GLRenderer.cpp:1057-1176; RenderSlice :922-1014; 1D-slice-FS.glsl,
// 2D-slice-FS.glsl, MIP-slice-FS.glsl, Transfer-MIP-FS.glsl).  The reference draws ONE quad textured with the finest
// single-brick LOD; here every pixel computes the texture coordinate the rasteriser would interpolate for its centre.
// ---------------------------------------------------------------------------------------------------------
struct SliceArgs {
  tvk_slice_params p;
  const uint8_t* brick;
  const uchar4* tf;
  uint32_t tf_w, tf_h;
  float4* image;
};

// texture3D(texVolume, c): GL_LINEAR, GL_CLAMP_TO_EDGE, UNORM
template <int BPV>
__device__ __forceinline__ float slice_sample(const SliceArgs& a, float cx, float cy, float cz) {
  const uint32_t* n = a.p.voxels;
  const float u = fmaf(cx, (float)n[0], -0.5f), v = fmaf(cy, (float)n[1], -0.5f), w = fmaf(cz, (float)n[2], -0.5f);
  const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
  const float ta = u - fu, tb = v - fv, tg = w - fw;
  const int x0 = min(max((int)fu, 0), (int)n[0] - 1), x1 = min(max((int)fu + 1, 0), (int)n[0] - 1);
  const int y0 = min(max((int)fv, 0), (int)n[1] - 1), y1 = min(max((int)fv + 1, 0), (int)n[1] - 1);
  const int z0 = min(max((int)fw, 0), (int)n[2] - 1), z1 = min(max((int)fw + 1, 0), (int)n[2] - 1);
  auto at = [&](int x, int y, int z) -> float {
    const size_t i = (size_t)x + (size_t)y * n[0] + (size_t)z * n[0] * n[1];
    return BPV == 1 ? (float)a.brick[i] : (float)reinterpret_cast<const uint16_t*>(a.brick)[i];
  };
  const float c00 = fmaf(ta, at(x1, y0, z0) - at(x0, y0, z0), at(x0, y0, z0)), c10 = fmaf(ta, at(x1, y1, z0) - at(x0, y1, z0), at(x0, y1, z0));
  const float c01 = fmaf(ta, at(x1, y0, z1) - at(x0, y0, z1), at(x0, y0, z1)), c11 = fmaf(ta, at(x1, y1, z1) - at(x0, y1, z1), at(x0, y1, z1));
  const float c0 = fmaf(tb, c10 - c00, c00), c1 = fmaf(tb, c11 - c01, c01);
  return fmaf(tg, c1 - c0, c0) * (BPV == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}
__device__ __forceinline__ int slice_nearest(float s, uint32_t n) { return min(max(__float2int_rz(s * (float)n), 0), (int)n - 1); }
__device__ __forceinline__ float4 slice_tf(const SliceArgs& a, int x, int y) {
  const uchar4 t = a.tf[(size_t)y * a.tf_w + x];
  return make_float4(t.x * (1.0f / 255.0f), t.y * (1.0f / 255.0f), t.z * (1.0f / 255.0f), t.w * (1.0f / 255.0f));
}

template <int BPV>
__global__ void __launch_bounds__(256) slice_kernel(const __grid_constant__ SliceArgs a) {
  const tvk_slice_params& p = a.p;
  const uint32_t px = blockIdx.x * 16u + (threadIdx.x & 15u), py = blockIdx.y * 16u + (threadIdx.x >> 4);
  if (px >= p.width || py >= p.height) return;
  // pixel centre in normalised device coordinates; the quad spans [-quad.x, quad.x] x [-quad.y, quad.y]
  const float nx = (2.0f * ((float)px + 0.5f)) / (float)p.width - 1.0f, ny = (2.0f * ((float)py + 0.5f)) / (float)p.height - 1.0f;
  float4 out = make_float4(0.f, 0.f, 0.f, 1.f);                             // glClearColor(0,0,0,1)
  const bool inside = fabsf(nx) <= p.quad[0] && fabsf(ny) <= p.quad[1];
  // horizontal / vertical position on the quad in [0,1] (left -> right, bottom -> top)
  const float sx = inside ? (nx + p.quad[0]) / (2.0f * p.quad[0]) : 0.f, sy = inside ? (ny + p.quad[1]) / (2.0f * p.quad[1]) : 0.f;
  // which texture axes the screen axes run along (RenderSlice's glTexCoord3d calls)
  const int ah = p.window_mode == TVK_WM_SAGITTAL ? 1 : 0;                  // sagittal: screen x = texture y; axial, coronal: texture x
  const int av = p.window_mode == TVK_WM_CORONAL ? 1 : 2;                   // coronal: screen y = texture y; sagittal, axial: texture z
  const int ad = (int)p.window_mode == TVK_WM_SAGITTAL ? 0 : (p.window_mode == TVK_WM_AXIAL ? 1 : 2);
  float c[3];
  c[ah] = fmaf(sx, p.max_coords[ah] - p.min_coords[ah], p.min_coords[ah]);
  c[av] = fmaf(sy, p.max_coords[av] - p.min_coords[av], p.min_coords[av]);
  c[ad] = p.slice;
  if (!p.use_mip) {
    if (inside) {
      const float val = slice_sample<BPV>(a, c[0], c[1], c[2]);
      float4 t;
      if (p.render_mode == TVK_RM_2DTRANS) {
        const float hx = 1.0f / (float)p.voxels[0], hy = 1.0f / (float)p.voxels[1], hz = 1.0f / (float)p.voxels[2];   // vVoxelStepsize
        const float xp = slice_sample<BPV>(a, c[0] + hx, c[1], c[2]), xm = slice_sample<BPV>(a, c[0] - hx, c[1], c[2]);
        const float yp = slice_sample<BPV>(a, c[0], c[1] - hy, c[2]), ym = slice_sample<BPV>(a, c[0], c[1] + hy, c[2]);
        const float zp = slice_sample<BPV>(a, c[0], c[1], c[2] + hz), zm = slice_sample<BPV>(a, c[0], c[1], c[2] - hz);
        const float gx = (xm - xp) / 2.0f, gy = (yp - ym) / 2.0f, gz = (zm - zp) / 2.0f;
        const float gm = sqrtf(fmaf(gz, gz, fmaf(gy, gy, gx * gx)));
        t = slice_tf(a, slice_nearest(val * p.trans_scale, a.tf_w), slice_nearest(1.0f - gm * p.gradient_scale, a.tf_h));
      } else {
        t = slice_tf(a, slice_nearest(val * p.trans_scale, a.tf_w), 0);
      }
      out = make_float4(t.x * t.w, t.y * t.w, t.z * t.w, 1.0f);             // rgb *= a; a = 1
    }
  } else {
    // MIP preview: GL_MAX blend of sampleVolume().xxxx over the brick's slices; outside the quad the cleared 0 remains
    float m = 0.0f;
    if (inside) {
      const uint32_t ns = p.voxels[ad];
      for (uint32_t i = 0; i < ns; ++i) {
        c[ad] = (float)((double)i / (double)ns);
        m = fmaxf(m, slice_sample<BPV>(a, c[0], c[1], c[2]));
      }
    }
    // Transfer-MIP-FS.glsl: the cleared alpha of 1 marks every pixel of the region -> vec4(tf(max).rgb, 1)
    const float4 t = slice_tf(a, slice_nearest(m * p.trans_scale, a.tf_w), 0);
    out = make_float4(t.x, t.y, t.z, 1.0f);
  }
  a.image[(size_t)py * p.width + px] = out;
}

extern "C" int tvk_render_slice(tvk_ctx* c, const tvk_slice_params* p, const void* brick) {
  if (!c || !p || !brick) return TVK_ERR_INVALID;
  if (p->window_mode > TVK_WM_CORONAL || !p->width || !p->height || !p->voxels[0] || !p->voxels[1] || !p->voxels[2] ||
      (p->bytes_per_voxel != 1 && p->bytes_per_voxel != 2) || !(p->quad[0] > 0.f) || !(p->quad[1] > 0.f))
    TVK_FAIL(c, TVK_ERR_INVALID, "bad slice parameters");
  const bool two_d = !p->use_mip && p->render_mode == TVK_RM_2DTRANS;
  if (!p->use_mip && p->render_mode != TVK_RM_1DTRANS && p->render_mode != TVK_RM_2DTRANS) TVK_FAIL(c, TVK_ERR_INVALID, "slice views classify with the 1-D or the 2-D transfer function");
  if (!c->d_tf || (two_d ? c->tf_h < 2 : c->tf_h != 1)) TVK_FAIL(c, TVK_ERR_INVALID, "the transfer function of the slice view's mode is not set");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int rc = tvk_ensure_frame(c, p->width, p->height, c->hash_size ? c->hash_size : 509);
  if (rc) return rc;
  c->show_stereo = false;
  if (c->eye == 0) c->iso_hits = nullptr;
  const size_t bytes = (size_t)p->voxels[0] * p->voxels[1] * p->voxels[2] * p->bytes_per_voxel;
  const void* d_brick = brick;
  void* tmp = nullptr;
  cudaPointerAttributes at;
  const bool known = cudaPointerGetAttributes(&at, brick) == cudaSuccess;
  if (!known) (void)cudaGetLastError();
  if (!known || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {   // host memory: the brick is small, copy it up
    TVK_CHECK_CUDA(c, cudaMallocAsync(&tmp, bytes, c->stream));
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(tmp, brick, bytes, cudaMemcpyHostToDevice, c->stream));
    d_brick = tmp;
  }
  SliceArgs a;
  a.p = *p; a.brick = static_cast<const uint8_t*>(d_brick); a.tf = c->d_tf; a.tf_w = c->tf_w; a.tf_h = c->tf_h; a.image = c->d_image;
  const dim3 grid((p->width + 15) / 16, (p->height + 15) / 16);
  if (p->bytes_per_voxel == 1) slice_kernel<1><<<grid, 256, 0, c->stream>>>(a);
  else slice_kernel<2><<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  if (tmp) TVK_CHECK_CUDA(c, cudaFreeAsync(tmp, c->stream));
  return TVK_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Full-resolution MIP of the capture mode ("MIP rotation"): GLRenderer::Render2DView's other branch (GLRenderer.cpp:1182-1250),
// GLRaycaster::RenderHQMIPPreLoop / RenderHQMIPInLoop (GLRaycaster.cpp:481-529), GLRaycaster-MIP-Rot-FS.glsl and
// Transfer-MIP-FS.glsl.  The reference draws every brick of the chosen LOD (front faces into the RGBA16F entry buffer, back
// faces with the MIP shader, GL_MAX blending) and classifies the result; here one thread per pixel walks the brick list --
// the maximum does not depend on the order.  Rays are general (near point -> far point through the inverse projection), so
// that the orthographic view of m_bOrthoView and the perspective one share the code.
// ---------------------------------------------------------------------------------------------------------
struct MipArgs {
  tvk_mip_params p;
  const tvk_classic_brick* bricks;
  uint32_t nbricks;
  const uint8_t* store;
  const uchar4* tf;
  uint32_t tf_w;
  float4* image;
  unsigned long long* stats;
};

template <int BPV>
__global__ void __launch_bounds__(128) mip_kernel(const __grid_constant__ MipArgs a) {
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t tiles_x = (a.p.width + 7) / 8;
  const uint32_t px = (warp % tiles_x) * 8u + (lane & 7u), py = (warp / tiles_x) * 4u + (lane >> 3);
  if ((warp / tiles_x) * 4u >= a.p.height) return;
  const bool alive = px < a.p.width && py < a.p.height;
  __shared__ uint32_t tile_list[4][TVK_TILE_LIST];
  uint32_t* const list = tile_list[threadIdx.x >> 5];
  const tvk_mip_params& p = a.p;
  const float nx = fmaf(((float)px + 0.5f) / (float)p.width, 2.0f, -1.0f), ny = fmaf(((float)py + 0.5f) / (float)p.height, 2.0f, -1.0f);
  const float4 n4 = c_xform(p.inv_projection, nx, ny, -1.0f, 1.0f), f4 = c_xform(p.inv_projection, nx, ny, 1.0f, 1.0f);
  const float e0[3] = {n4.x / n4.w, n4.y / n4.w, n4.z / n4.w};                // near point, eye space
  const float ed[3] = {f4.x / f4.w - e0[0], f4.y / f4.w - e0[1], f4.z / f4.w - e0[2]};   // near -> far
  const float4 O = c_xform(p.inv_model_view, e0[0], e0[1], e0[2], 1.0f);
  const float4 D = c_xform(p.inv_model_view, ed[0], ed[1], ed[2], 0.0f);
  const float o[3] = {O.x, O.y, O.z}, d[3] = {D.x, D.y, D.z};
  float vmax = 0.0f;
  bool covered = false;
  unsigned long long nsamples = 0;
  TilePlanes tp;
  {                                                        // (near -> far rays: every pixel has its own origin)
    float co[4][3], cd[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float fx, fy;
      tile_corner_pixel((warp % tiles_x) * 8u, (warp / tiles_x) * 4u, k, fx, fy);
      const float cx = fmaf(fx / (float)p.width, 2.0f, -1.0f), cy = fmaf(fy / (float)p.height, 2.0f, -1.0f);
      const float4 cn = c_xform(p.inv_projection, cx, cy, -1.0f, 1.0f), cf = c_xform(p.inv_projection, cx, cy, 1.0f, 1.0f);
      const float q0[3] = {cn.x / cn.w, cn.y / cn.w, cn.z / cn.w};
      const float4 cO = c_xform(p.inv_model_view, q0[0], q0[1], q0[2], 1.0f);
      const float4 cD = c_xform(p.inv_model_view, cf.x / cf.w - q0[0], cf.y / cf.w - q0[1], cf.z / cf.w - q0[2], 0.0f);
      co[k][0] = cO.x; co[k][1] = cO.y; co[k][2] = cO.z;
      cd[k][0] = cD.x; cd[k][1] = cD.y; cd[k][2] = cD.z;
    }
    tile_planes(co, cd, false, tp);
  }
  for (uint32_t base = 0; base < a.nbricks;) {
   const uint32_t in_tile = tile_list_chunk(tp, a.bricks, a.nbricks, base, list, lane);
   for (uint32_t li = 0; alive && li < in_tile; ++li) {
    const tvk_classic_brick& b = a.bricks[list[li]];
    float s0 = -INFINITY, s1 = INFINITY;
    bool miss = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float lo = b.center[k] - b.extension[k] / 2.0f, hi = b.center[k] + b.extension[k] / 2.0f;
      if (d[k] == 0.0f) { if (o[k] < lo || o[k] > hi) miss = true; continue; }
      const float inv = 1.0f / d[k];
      float ta = (lo - o[k]) * inv, tb = (hi - o[k]) * inv;
      if (ta > tb) { const float t = ta; ta = tb; tb = t; }
      s0 = fmaxf(s0, ta); s1 = fminf(s1, tb);
    }
    const float se = fmaxf(s0, 0.0f);                       // front face (the near plane when it lies behind it)
    s1 = fminf(s1, 1.0f);
    if (miss || !(s1 > se)) continue;
    covered = true;
    // ray entry through the RGBA16F buffer (m_pFBORayEntry), ray exit from the interpolated varying
    const float enx = __half2float(__float2half_rn(fmaf(ed[0], se, e0[0]))), eny = __half2float(__float2half_rn(fmaf(ed[1], se, e0[1]))),
                enz = __half2float(__float2half_rn(fmaf(ed[2], se, e0[2])));
    const float xx = fmaf(ed[0], s1, e0[0]), xy = fmaf(ed[1], s1, e0[1]), xz = fmaf(ed[2], s1, e0[2]);
    const float4 te = c_xform(b.eye_to_tex, enx, eny, enz, 1.0f), tx = c_xform(b.eye_to_tex, xx, xy, xz, 1.0f);
    const float dx = xx - enx, dy = xy - eny, dz = xz - enz;
    const float len = sqrtf(c_dot3(dx, dy, dz, dx, dy, dz));
    const int steps = (int)(len / b.ray_step) + 1;
    const float q = len / b.ray_step;
    const float itx = (tx.x - te.x) / q, ity = (tx.y - te.y) / q, itz = (tx.z - te.z) / q;
    const uint8_t* data = a.store + b.store_offset;
    float ptx = te.x, pty = te.y, ptz = te.z;
    for (int i = 0; i < steps; ++i) {                       // GLRaycaster-MIP-Rot-FS.glsl:60-66
      vmax = fmaxf(vmax, brick_sample<BPV>(data, b.voxels, ptx, pty, ptz));
      ptx += itx; pty += ity; ptz += itz;
    }
    nsamples += (unsigned long long)steps;
   }
   __syncwarp();
  }
  // Transfer-MIP-FS.glsl: colour of the 1-D function, opacity ignored; pixels no brick covers are black; alpha 1 everywhere
  float4 out = make_float4(0.f, 0.f, 0.f, 1.f);
  if (covered) {
    const uchar4 t = __ldg(a.tf + c_nearest(vmax * p.trans_scale, a.tf_w));
    out = make_float4((float)t.x * (1.0f / 255.0f), (float)t.y * (1.0f / 255.0f), (float)t.z * (1.0f / 255.0f), 1.0f);
  }
  if (alive) a.image[(size_t)py * p.width + px] = out;
  else nsamples = 0;
  for (int s = 16; s > 0; s >>= 1) nsamples += __shfl_down_sync(0xFFFFFFFFu, nsamples, s);
  if (lane == 0 && a.stats) atomicAdd(a.stats, nsamples);
}

extern "C" int tvk_raycast_mip(tvk_ctx* c, const tvk_mip_params* p, const tvk_classic_brick* bricks, uint32_t n, const void* store) {
  if (!c || !p || (n && !bricks) || !store) return TVK_ERR_INVALID;
  if (!c->d_tf || c->tf_h != 1) TVK_FAIL(c, TVK_ERR_INVALID, "the MIP is classified with the 1-D transfer function");
  if (p->bytes_per_voxel != 1 && p->bytes_per_voxel != 2) TVK_FAIL(c, TVK_ERR_INVALID, "bad voxel size");
  if (!p->width || !p->height) TVK_FAIL(c, TVK_ERR_INVALID, "bad image size");
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  int rc = tvk_ensure_frame(c, p->width, p->height, c->hash_size ? c->hash_size : 509);
  if (rc) return rc;
  c->show_stereo = false;
  if (c->eye == 0) c->iso_hits = nullptr;
  tvk_classic_brick* d_bricks = nullptr;
  TVK_CHECK_CUDA(c, cudaMallocAsync(&d_bricks, sizeof(tvk_classic_brick) * (n ? n : 1), c->stream));
  if (n) TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_bricks, bricks, sizeof(tvk_classic_brick) * n, cudaMemcpyHostToDevice, c->stream));
  MipArgs a;
  a.p = *p; a.bricks = d_bricks; a.nbricks = n; a.store = static_cast<const uint8_t*>(store);
  a.tf = c->d_tf; a.tf_w = c->tf_w; a.image = c->d_image; a.stats = c->d_stats;
  const uint32_t tiles = ((p->width + 7) / 8) * ((p->height + 3) / 4);
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  if (p->bytes_per_voxel == 1) mip_kernel<1><<<(tiles + 3) / 4, 128, 0, c->stream>>>(a);
  else mip_kernel<2><<<(tiles + 3) / 4, 128, 0, c->stream>>>(a);
  c->launches++;
  c->k1_serial++;                                          // (the sample counter moved: cached statistics are stale)
  TVK_CHECK_CUDA(c, cudaGetLastError());
  TVK_CHECK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  TVK_CHECK_CUDA(c, cudaFreeAsync(d_bricks, c->stream));
  return TVK_OK;
}
