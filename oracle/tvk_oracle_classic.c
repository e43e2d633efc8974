/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CLASSIC per-brick ray-caster (SURVEY 8a rows
 * R12): what GLRaycaster draws for one sub-frame, pixel by pixel.
 *
 * Follows, under /root/reference/Tuvok/:
 *   Renderer/GL/GLRaycaster.cpp:383-479   Render3DPreLoop / Render3DInLoop (near-plane pass, front faces into the
 *                                         RGBA16F entry buffer :97, back faces with the ray-cast shader)
 *   Renderer/GL/GLRaycaster.cpp:239-300   SetBrickDepShaderVars (fRayStepsize, fStepScale, vVoxelStepsize)
 *   Renderer/GL/GLRaycaster.cpp:589-612   ComputeEyeToTextureMatrix (computed by the caller, see oracle/host.py)
 *   Shaders/GLRaycaster-1D-FS.glsl:51-82, GLRaycaster-1D-light-FS.glsl:83-134   main()
 *   Shaders/GLRaycaster-2D-FS.glsl:52-94, GLRaycaster-2D-light-FS.glsl:62-117   main() (2D transfer function)
 *   Shaders/GLRaycaster-ISO-FS.glsl:56-104, RefineIsosurface.glsl:37-53         first hit per brick, refined; the
 *                                         hit survives if it passes the GL_LESS depth test (GLRenderer.cpp:139) against
 *                                         the buffer cleared to 1.0 (GLRaycaster.cpp:391-394); depth compared in fp32
 *   Shaders/Compose-FS.glsl:49-75, GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2811)   isosurface shading
 *   Shaders/VRender1D.glsl, VRender1DLit.glsl, Volume3D.glsl (sampleVolume, ComputeGradient, ComputeNormal),
 *   Shaders/lighting.glsl, Shaders/Compositing.glsl:33-38
 *   Renderer/GL/GLRenderer.cpp:152-153    frame-buffer blend ONE_MINUS_DST_ALPHA / ONE between bricks
 * Parity status: PINNED to the reference's own shader text -- oracle/classicref.py compiles GLRaycaster-*-FS.glsl, Volume3D,
 * VRender1D*, lighting, Compositing, RefineIsosurface, Compose-FS and Compose-CV-FS with g++ against the GLSL shim and
 * tests/test_glsl_classic_reference.py requires this restatement to agree with it (120-158 dB; 1D / 2D, lit / unlit,
 * isosurface, ClearView, clip plane, u8 / u16).  The reference's pixels themselves need an OpenGL driver.
 * Restatement choices: brick faces are intersected analytically with the pixel-centre ray instead of being
 * rasterised; a brick whose front face lies behind the near plane starts at the near plane (the reference reads
 * whatever the previous brick left in the entry buffer there); the entry point is rounded to half precision as the
 * RGBA16F buffer does, the exit point is not (it is an interpolated varying).
 * Floating-point contract as in tvk_oracle_render.c.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "tvk_cuda.h"

static float half_round(float f) {          /* float -> IEEE binary16 (round to nearest even) -> float */
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = x & 0x80000000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return f;                              /* inf / nan */
  if (x >= 0x477FF000u) { x = 0x7F800000u | sign; memcpy(&f, &x, 4); return f; }   /* rounds to >= 65520 -> inf */
  float a;
  memcpy(&a, &x, 4);
  if (x < 0x38800000u) {                                       /* below 2^-14: half subnormal, quantum 2^-24 */
    const float q = 5.9604644775390625e-08f;
    a = rintf(a / q) * q;
  } else {
    const uint32_t rem = x & 0x1FFFu, base = x & ~0x1FFFu;     /* keep 10 mantissa bits */
    uint32_t r = base;
    if (rem > 0x1000u || (rem == 0x1000u && (base & 0x2000u))) r += 0x2000u;
    memcpy(&a, &r, 4);
  }
  uint32_t y;
  memcpy(&y, &a, 4);
  y |= sign;
  memcpy(&a, &y, 4);
  return a;
}

static inline float dot3c(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, ax * bx)); }
static inline float clamp01c(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
static void xformc(const float* m, float x, float y, float z, float w, float o[4]) {
  o[0] = fmaf(w, m[12], fmaf(z, m[8], fmaf(y, m[4], x * m[0])));
  o[1] = fmaf(w, m[13], fmaf(z, m[9], fmaf(y, m[5], x * m[1])));
  o[2] = fmaf(w, m[14], fmaf(z, m[10], fmaf(y, m[6], x * m[2])));
  o[3] = fmaf(w, m[15], fmaf(z, m[11], fmaf(y, m[7], x * m[3])));
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int f2i_sat(float f) {          /* (int)f with the saturation CUDA's conversion has; NaN -> 0 */
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return (int)f;
}

static float brick_sample(const uint8_t* data, uint32_t bpv, const uint32_t vox[3], float tx, float ty, float tz) {
  const float u = fmaf(tx, (float)vox[0], -0.5f), v = fmaf(ty, (float)vox[1], -0.5f), w = fmaf(tz, (float)vox[2], -0.5f);
  const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
  const float a = u - fu, b = v - fv, g = w - fw;
  const int mx = (int)vox[0] - 1, my = (int)vox[1] - 1, mz = (int)vox[2] - 1;
  const int x0 = clampi(f2i_sat(fu), 0, mx), x1 = clampi(f2i_sat(fu) == 2147483647 ? 2147483647 : f2i_sat(fu) + 1, 0, mx);
  const int y0 = clampi(f2i_sat(fv), 0, my), y1 = clampi(f2i_sat(fv) == 2147483647 ? 2147483647 : f2i_sat(fv) + 1, 0, my);
  const int z0 = clampi(f2i_sat(fw), 0, mz), z1 = clampi(f2i_sat(fw) == 2147483647 ? 2147483647 : f2i_sat(fw) + 1, 0, mz);
  const size_t sx = vox[0], sxy = (size_t)vox[0] * vox[1];
#define VOX(X, Y, Z) (bpv == 1 ? (float)data[(X) + (Y) * sx + (Z) * sxy] : (float)((const uint16_t*)data)[(X) + (Y) * sx + (Z) * sxy])
  const float v000 = VOX(x0, y0, z0), v100 = VOX(x1, y0, z0), v010 = VOX(x0, y1, z0), v110 = VOX(x1, y1, z0);
  const float v001 = VOX(x0, y0, z1), v101 = VOX(x1, y0, z1), v011 = VOX(x0, y1, z1), v111 = VOX(x1, y1, z1);
#undef VOX
  const float c00 = fmaf(a, v100 - v000, v000), c10 = fmaf(a, v110 - v010, v010);
  const float c01 = fmaf(a, v101 - v001, v001), c11 = fmaf(a, v111 - v011, v011);
  const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
  return fmaf(g, c1 - c0, c0) * (bpv == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}

static void brick_gradient(const uint8_t* data, uint32_t bpv, const uint32_t vox[3], float tx, float ty, float tz, float g[3]) {
  const float hx = 1.0f / (float)vox[0], hy = 1.0f / (float)vox[1], hz = 1.0f / (float)vox[2];
  const float gxp = brick_sample(data, bpv, vox, tx + hx, ty, tz), gxm = brick_sample(data, bpv, vox, tx - hx, ty, tz);
  const float gyp = brick_sample(data, bpv, vox, tx, ty - hy, tz), gym = brick_sample(data, bpv, vox, tx, ty + hy, tz);
  const float gzp = brick_sample(data, bpv, vox, tx, ty, tz + hz), gzm = brick_sample(data, bpv, vox, tx, ty, tz - hz);
  g[0] = (gxm - gxp) / 2.0f; g[1] = (gyp - gym) / 2.0f; g[2] = (gzm - gzp) / 2.0f;
}

static void classic_normal(const tvk_classic_params* p, const float g[3], float N[3]) {
  const float gx = g[0] * p->domain_scale[0], gy = g[1] * p->domain_scale[1], gz = g[2] * p->domain_scale[2];
  N[0] = fmaf(gz, p->normal_matrix[6], fmaf(gy, p->normal_matrix[3], gx * p->normal_matrix[0]));
  N[1] = fmaf(gz, p->normal_matrix[7], fmaf(gy, p->normal_matrix[4], gx * p->normal_matrix[1]));
  N[2] = fmaf(gz, p->normal_matrix[8], fmaf(gy, p->normal_matrix[5], gx * p->normal_matrix[2]));
  const float nl = sqrtf(dot3c(N[0], N[1], N[2], N[0], N[1], N[2]));
  if (nl > 0.0f) { N[0] /= nl; N[1] /= nl; N[2] /= nl; }
}

static inline int nearest_c(float s, uint32_t n) { return clampi(f2i_sat(s * (float)n), 0, (int)n - 1); }

static void classic_compose(const tvk_classic_params* p, const float hit[4], const float hn[3], float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  if (hit[3] == 0.0f) return;
  const float nx = hn[0], ny = hn[1], nz = fabsf(hn[2]);
  float vx = 0.0f - hit[0], vy = 0.0f - hit[1], vz = 0.0f - hit[2];
  float l = sqrtf(dot3c(vx, vy, vz, vx, vy, vz));
  vx /= l; vy /= l; vz /= l;
  const float d2 = 2.0f * dot3c(nx, ny, nz, vx, vy, vz);
  float rx = fmaf(-d2, nx, vx), ry = fmaf(-d2, ny, vy), rz = fmaf(-d2, nz, vz);
  l = sqrtf(dot3c(rx, ry, rz, rx, ry, rz));
  rx /= l; ry /= l; rz /= l;
  const float ndl = fmaxf(fabsf(dot3c(nx, ny, nz, -p->light_dir[0], -p->light_dir[1], -p->light_dir[2])), 0.0f);
  float sp = fmaxf(dot3c(rx, ry, rz, p->light_dir[0], p->light_dir[1], p->light_dir[2]), 0.0f);
  sp *= sp; sp *= sp; sp *= sp;
  for (int k = 0; k < 3; ++k)
    out[k] = clamp01c(fmaf(p->light_specular[k], sp, fmaf(p->light_diffuse[k] * p->iso_color[k], ndl, p->light_ambient[k])));
  out[3] = 1.0f;
}

/* tf_rgba8: [tf_h][tf_w][4]; tf_h == 1 for the 1D modes */
/* Compose-CV-FS.glsl:52-62 */
static void cv_lighting(const tvk_classic_params* p, const float pos[4], const float nrm[3], const float diffuse[3], float out[3]) {
  const float nx = nrm[0], ny = nrm[1], nz = fabsf(nrm[2]);
  float vx = 0.0f - pos[0], vy = 0.0f - pos[1], vz = 0.0f - pos[2];
  float l = sqrtf(dot3c(vx, vy, vz, vx, vy, vz));
  vx /= l; vy /= l; vz /= l;
  const float d2 = 2.0f * dot3c(nx, ny, nz, vx, vy, vz);
  float rx = fmaf(-d2, nx, vx), ry = fmaf(-d2, ny, vy), rz = fmaf(-d2, nz, vz);
  l = sqrtf(dot3c(rx, ry, rz, rx, ry, rz));
  rx /= l; ry /= l; rz /= l;
  const float ndl = fmaxf(fabsf(dot3c(nx, ny, nz, -p->light_dir[0], -p->light_dir[1], -p->light_dir[2])), 0.0f);
  float sp = fmaxf(dot3c(rx, ry, rz, p->light_dir[0], p->light_dir[1], p->light_dir[2]), 0.0f);
  sp *= sp; sp *= sp; sp *= sp;
  for (int k = 0; k < 3; ++k) out[k] = clamp01c(fmaf(p->light_specular[k], sp, fmaf(diffuse[k], ndl, p->light_ambient[k])));
}

/* cv != NULL (isosurface mode): ClearView -- the second pass of GLRaycaster::Render3DInLoop (GLRaycaster.cpp:438-452) with
 * GLRaycaster-ISO-CV-FS.glsl:57-115 per brick, then Compose-CV-FS.glsl:65-108 with the uniforms of
 * GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2776-2795).  hits: 4 * width * height * 4 floats of scratch (m_pFBOIsoHit
 * buffers 0 / 1 in fp32, m_pFBOCVHit buffers 0 / 1 rounded to half precision, GLRenderer.cpp:1788-1796). */
static int classic_impl(const tvk_classic_params* p, const tvk_clearview_params* cv, float* hits, const tvk_classic_brick* bricks, uint32_t n,
                        const uint8_t* store, const uint8_t* tf_rgba8, uint32_t tf_n, uint32_t tf_h, float* image, uint64_t* samples_out);

int orc_raycast_classic(const tvk_classic_params* p, const tvk_classic_brick* bricks, uint32_t n, const uint8_t* store,
                        const uint8_t* tf_rgba8, uint32_t tf_n, uint32_t tf_h, float* image, uint64_t* samples_out) {
  return classic_impl(p, NULL, NULL, bricks, n, store, tf_rgba8, tf_n, tf_h, image, samples_out);
}

int orc_raycast_classic_cv(const tvk_classic_params* p, const tvk_clearview_params* cv, const tvk_classic_brick* bricks, uint32_t n,
                           const uint8_t* store, float* hits, float* image, uint64_t* samples_out) {
  static const uint8_t no_tf[4] = {0, 0, 0, 0};
  if (!p || !cv || !hits || p->render_mode != TVK_RM_ISOSURFACE) return 1;
  int rc = classic_impl(p, cv, hits, bricks, n, store, no_tf, 1, 1, image, samples_out);
  if (rc) return rc;
  const size_t npx = (size_t)p->width * p->height;
  const float *H0 = hits, *H1 = hits + 4 * npx, *H2 = hits + 8 * npx, *H3 = hits + 12 * npx;
  const float d1[3] = {p->light_diffuse[0] * p->iso_color[0], p->light_diffuse[1] * p->iso_color[1], p->light_diffuse[2] * p->iso_color[2]};
  const float d2[3] = {p->light_diffuse[0] * cv->cv_color[0], p->light_diffuse[1] * cv->cv_color[1], p->light_diffuse[2] * cv->cv_color[2]};
  for (int64_t y = 0; y < (int64_t)p->height; ++y)
    for (int64_t x = 0; x < (int64_t)p->width; ++x) {
      const size_t i = (size_t)y * p->width + (size_t)x;
      float* o = image + 4 * i;
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      const float* pos = H0 + 4 * i;
      if (pos[3] == 0.0f) continue;
      const float* nrm = H1 + 4 * i;
      float ctx[3];
      cv_lighting(p, pos, nrm, d1, ctx);
      const int64_t xs[4] = {x + 1 < (int64_t)p->width ? x + 1 : x, x > 0 ? x - 1 : 0, x, x};
      const int64_t ys[4] = {y, y, y + 1 < (int64_t)p->height ? y + 1 : y, y > 0 ? y - 1 : 0};
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int k = 0; k < 4; ++k) {
        const float* nb = H1 + 4 * ((size_t)ys[k] * p->width + (size_t)xs[k]);
        sx += fabsf(nb[0] - nrm[0]); sy += fabsf(nb[1] - nrm[1]); sz += fabsf(nb[2] - nrm[2]);
      }
      const float curv = sqrtf(dot3c(sx, sy, sz, sx, sy, sz));
      const float qx = pos[0] - cv->pick_pos[0], qy = pos[1] - cv->pick_pos[1], qz = pos[2] - cv->pick_pos[2];
      const float dist_w = sqrtf(dot3c(qx, qy, qz, qx, qy, qz)) * cv->size;
      const float blend = clamp01c(fmaxf(curv * cv->context_scale, clamp01c(dist_w)));
      const float* pos2 = H2 + 4 * i;
      float foc[4] = {0.f, 0.f, 0.f, 0.f};
      if (pos2[3] != 0.0f) { cv_lighting(p, pos2, H3 + 4 * i, d2, foc); foc[3] = 1.0f; }
      const float ib = 1.0f - blend;
      const float border = 0.5f * (1.0f - clamp01c(fabsf(dist_w - 1.0f) * cv->border_scale));
      for (int k = 0; k < 3; ++k) o[k] = clamp01c(fmaf(foc[k], ib, ctx[k] * blend)) - border;
      o[3] = clamp01c(fmaf(foc[3], ib, 1.0f * blend));
    }
  return 0;
}

static int classic_impl(const tvk_classic_params* p, const tvk_clearview_params* cv, float* hits, const tvk_classic_brick* bricks, uint32_t n,
                        const uint8_t* store, const uint8_t* tf_rgba8, uint32_t tf_n, uint32_t tf_h, float* image, uint64_t* samples_out) {
  if (!p || (n && !bricks) || !store || !tf_rgba8 || !image || !tf_h) return 1;
  const int lit = p->lighting && p->render_mode != TVK_RM_ISOSURFACE;
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
  for (int64_t py = 0; py < (int64_t)p->height; ++py) {
    for (uint32_t px = 0; px < p->width; ++px) {
      const float nx = fmaf(((float)px + 0.5f) / (float)p->width, 2.0f, -1.0f), ny = fmaf(((float)py + 0.5f) / (float)p->height, 2.0f, -1.0f);
      float n4[4], O[4], D[4];
      xformc(p->inv_projection, nx, ny, -1.0f, 1.0f, n4);
      const float ex = n4[0] / n4[3], ey = n4[1] / n4[3], ez = n4[2] / n4[3];
      xformc(p->inv_model_view, 0.f, 0.f, 0.f, 1.0f, O);
      xformc(p->inv_model_view, ex, ey, ez, 0.0f, D);
      float dst[4] = {0.f, 0.f, 0.f, 0.f};
      float hit[4] = {0.f, 0.f, 0.f, 0.f}, hn[3] = {0.f, 0.f, 0.f}, best_depth = 1.0f;
      float hit2[4] = {0.f, 0.f, 0.f, 0.f}, hn2[3] = {0.f, 0.f, 0.f}, best_depth2 = 1.0f, best_tile = 0.0f;
      for (uint32_t bi = 0; bi < n; ++bi) {
        const tvk_classic_brick* b = &bricks[bi];
        float s0 = -INFINITY, s1 = INFINITY;
        int miss = 0;
        for (int k = 0; k < 3; ++k) {
          const float lo = b->center[k] - b->extension[k] / 2.0f, hi = b->center[k] + b->extension[k] / 2.0f;
          if (D[k] == 0.0f) { if (O[k] < lo || O[k] > hi) miss = 1; continue; }
          const float inv = 1.0f / D[k];
          float ta = (lo - O[k]) * inv, tb = (hi - O[k]) * inv;
          if (ta > tb) { const float t = ta; ta = tb; tb = t; }
          s0 = fmaxf(s0, ta); s1 = fminf(s1, tb);
        }
        if (p->clip_enabled && !miss) {       /* RenderBox + Clipper::BoxPlane (GLRaycaster.cpp:327-337): n.p + d <= 0 is kept */
          const float dn = dot3c(p->clip_plane[0], p->clip_plane[1], p->clip_plane[2], D[0], D[1], D[2]);
          const float on = dot3c(p->clip_plane[0], p->clip_plane[1], p->clip_plane[2], O[0], O[1], O[2]) + p->clip_plane[3];
          if (dn == 0.0f) { if (on > 0.0f) miss = 1; }
          else {
            const float t = -on / dn;
            if (dn > 0.0f) s1 = fminf(s1, t); else s0 = fmaxf(s0, t);
          }
        }
        const float se = fmaxf(s0, 1.0f);
        if (miss || !(s1 > se)) continue;
        const float enx = half_round(ex * se), eny = half_round(ey * se), enz = half_round(ez * se);
        const float xx = ex * s1, xy = ey * s1, xz = ez * s1;
        float te[4], tx[4];
        xformc(b->eye_to_tex, enx, eny, enz, 1.0f, te);
        xformc(b->eye_to_tex, xx, xy, xz, 1.0f, tx);
        float dx = xx - enx, dy = xy - eny, dz = xz - enz;
        const float len = sqrtf(dot3c(dx, dy, dz, dx, dy, dz));
        const int steps = f2i_sat(len / b->ray_step) + 1;
        const float q = len / b->ray_step;
        const float itx = (tx[0] - te[0]) / q, ity = (tx[1] - te[1]) / q, itz = (tx[2] - te[2]) / q;
        dx /= len; dy /= len; dz /= len;
        const float ix = dx * b->ray_step, iy = dy * b->ray_step, iz = dz * b->ray_step;
        const uint8_t* data = store + b->store_offset;
        float col[4] = {0.f, 0.f, 0.f, 0.f};
        float ptx = te[0], pty = te[1], ptz = te[2], pex = enx, pey = eny, pez = enz;
        const uint32_t bpv = p->bytes_per_voxel;
        if (p->render_mode == TVK_RM_ISOSURFACE) {
          int found = 0;
          for (int i = 0; i < steps; ++i) {
            const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
            total++;
            if (v >= p->isoval) { found = 1; break; }
            ptx += itx; pty += ity; ptz += itz;
          }
          if (found) {
            float rx = itx / 2.0f, ry = ity / 2.0f, rz = itz / 2.0f;
            ptx -= rx; pty -= ry; ptz -= rz;
            for (int i = 0; i < 5; ++i) {
              rx /= 2.0f; ry /= 2.0f; rz /= 2.0f;
              const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
              if (v >= p->isoval) { ptx -= rx; pty -= ry; ptz -= rz; } else { ptx += rx; pty += ry; ptz += rz; }
            }
            const float ltx = tx[0] - te[0], lty = tx[1] - te[1], ltz = tx[2] - te[2];
            const float len_tex = sqrtf(dot3c(ltx, lty, ltz, ltx, lty, ltz));
            const float hx = ptx - te[0], hy = pty - te[1], hz = ptz - te[2];
            const float f = sqrtf(dot3c(hx, hy, hz, hx, hy, hz)) / len_tex;
            const float hpx = fmaf(xx, f, enx * (1.0f - f)), hpy = fmaf(xy, f, eny * (1.0f - f)), hpz = fmaf(xz, f, enz * (1.0f - f));
            const float depth = clamp01c(p->proj_param[0] + p->proj_param[1] / (0.0f - hpz));
            if (depth < best_depth) {
              best_depth = depth;
              hit[0] = hpx; hit[1] = hpy; hit[2] = hpz; hit[3] = f;
              float g[3];
              brick_gradient(data, bpv, b->voxels, ptx, pty, ptz, g);
              classic_normal(p, g, hn);
              best_tile = (float)bi;                     /* gl_FragData[1].w = float(iTileID), GLRaycaster-ISO-FS.glsl */
            }
          }
          if (cv) {
            /* GLRaycaster-ISO-CV-FS.glsl:66-115.  fLastTileID comes from the (possibly just updated) first-pass target;
             * a cleared target reads tile 0 / parameter 0.  Only vRayEntryTex is advanced (:71-72 advance vRayEntry by
             * itself): fRayLength, iStepCount and the interpolation end points stay those of the whole brick segment. */
            float sx = te[0], sy = te[1], sz = te[2];
            if ((float)bi == best_tile) {
              const float pp = hit[3];
              sx = fmaf(tx[0], pp, te[0] * (1.0f - pp)); sy = fmaf(tx[1], pp, te[1] * (1.0f - pp)); sz = fmaf(tx[2], pp, te[2] * (1.0f - pp));
            }
            const float l2x = tx[0] - sx, l2y = tx[1] - sy, l2z = tx[2] - sz;
            const float len_tex2 = sqrtf(dot3c(l2x, l2y, l2z, l2x, l2y, l2z));
            const float i2x = l2x / q, i2y = l2y / q, i2z = l2z / q;
            float cx = sx, cy = sy, cz = sz;
            int found2 = 0;
            for (int i = 0; i < steps; ++i) {
              const float v = brick_sample(data, bpv, b->voxels, cx, cy, cz);
              total++;
              if (v >= cv->cv_isoval) { found2 = 1; break; }
              cx += i2x; cy += i2y; cz += i2z;
            }
            if (found2) {
              float rx = i2x / 2.0f, ry = i2y / 2.0f, rz = i2z / 2.0f;
              cx -= rx; cy -= ry; cz -= rz;
              for (int i = 0; i < 5; ++i) {
                rx /= 2.0f; ry /= 2.0f; rz /= 2.0f;
                const float v = brick_sample(data, bpv, b->voxels, cx, cy, cz);
                if (v >= cv->cv_isoval) { cx -= rx; cy -= ry; cz -= rz; } else { cx += rx; cy += ry; cz += rz; }
              }
              const float hx = cx - sx, hy = cy - sy, hz = cz - sz;
              const float f2 = sqrtf(dot3c(hx, hy, hz, hx, hy, hz)) / len_tex2;
              const float depth2 = clamp01c(p->proj_param[0] + p->proj_param[1] / (0.0f - xz));   /* depth of the ray EXIT (:106) */
              if (depth2 < best_depth2) {
                best_depth2 = depth2;
                hit2[0] = fmaf(xx, f2, enx * (1.0f - f2)); hit2[1] = fmaf(xy, f2, eny * (1.0f - f2)); hit2[2] = fmaf(xz, f2, enz * (1.0f - f2));
                hit2[3] = f2;
                float g[3];
                brick_gradient(data, bpv, b->voxels, cx, cy, cz, g);
                classic_normal(p, g, hn2);
              }
            }
          }
          continue;
        }
        for (int i = 0; i < steps; ++i) {
          const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
          float c[4], g[3];
          const int two_d = p->render_mode == TVK_RM_2DTRANS;
          if (two_d || lit) brick_gradient(data, bpv, b->voxels, ptx, pty, ptz, g);
          size_t ti;
          if (two_d) {
            const float gm = sqrtf(dot3c(g[0], g[1], g[2], g[0], g[1], g[2]));
            ti = (size_t)nearest_c(fmaf(-gm, p->gradient_scale, 1.0f), tf_h) * tf_n + nearest_c(v * p->trans_scale, tf_n);
          } else ti = (size_t)clampi(f2i_sat(v * p->trans_scale * (float)tf_n), 0, (int)tf_n - 1);
          for (int k = 0; k < 4; ++k) c[k] = (float)tf_rgba8[4 * ti + k] * (1.0f / 255.0f);
          total++;
          if (lit) {
            float N[3];
            classic_normal(p, g, N);
            const float Nx = N[0], Ny = N[1], Nz = N[2];
            float vx = 0.0f - pex, vy = 0.0f - pey, vz = 0.0f - pez;
            float l = sqrtf(dot3c(vx, vy, vz, vx, vy, vz));
            vx /= l; vy /= l; vz /= l;
            const float d2 = 2.0f * dot3c(Nx, Ny, Nz, vx, vy, vz);
            float rx = fmaf(-d2, Nx, vx), ry = fmaf(-d2, Ny, vy), rz = fmaf(-d2, Nz, vz);
            l = sqrtf(dot3c(rx, ry, rz, rx, ry, rz));
            rx /= l; ry /= l; rz /= l;
            const float ndl = fmaxf(fabsf(dot3c(Nx, Ny, Nz, p->light_dir[0], p->light_dir[1], p->light_dir[2])), 0.0f);
            float sp = fmaxf(dot3c(rx, ry, rz, p->light_dir[0], p->light_dir[1], p->light_dir[2]), 0.0f);
            sp *= sp; sp *= sp; sp *= sp;
            c[0] = clamp01c(fmaf(p->light_specular[0], sp, fmaf(p->light_diffuse[0] * c[0], ndl, p->light_ambient[0])));
            c[1] = clamp01c(fmaf(p->light_specular[1], sp, fmaf(p->light_diffuse[1] * c[1], ndl, p->light_ambient[1])));
            c[2] = clamp01c(fmaf(p->light_specular[2], sp, fmaf(p->light_diffuse[2] * c[2], ndl, p->light_ambient[2])));
          }
          c[3] = (p->step_scale == 1.0f) ? 1.0f - (1.0f - c[3]) : 1.0f - powf(1.0f - c[3], p->step_scale);
          const float wgt = (1.0f - col[3]) * c[3];
          col[0] = fmaf(c[0], wgt, col[0]); col[1] = fmaf(c[1], wgt, col[1]); col[2] = fmaf(c[2], wgt, col[2]);
          col[3] = fmaf(c[3], 1.0f - col[3], col[3]);
          pex += ix; pey += iy; pez += iz;
          ptx += itx; pty += ity; ptz += itz;
          if (col[3] >= 0.99f) break;
        }
        const float k = 1.0f - dst[3];
        dst[0] = fmaf(col[0], k, dst[0]); dst[1] = fmaf(col[1], k, dst[1]); dst[2] = fmaf(col[2], k, dst[2]); dst[3] = fmaf(col[3], k, dst[3]);
      }
      if (cv) {
        const size_t npx = (size_t)p->width * p->height, i = (size_t)py * p->width + px;
        float* h0 = hits + 4 * i; float* h1 = hits + 4 * (npx + i); float* h2 = hits + 4 * (2 * npx + i); float* h3 = hits + 4 * (3 * npx + i);
        for (int k = 0; k < 4; ++k) { h0[k] = hit[k]; h2[k] = half_round(hit2[k]); }
        for (int k = 0; k < 3; ++k) { h1[k] = hn[k]; h3[k] = half_round(hn2[k]); }
        h1[3] = best_tile; h3[3] = 0.0f;
        continue;
      }
      if (p->render_mode == TVK_RM_ISOSURFACE) classic_compose(p, hit, hn, dst);
      float* o = image + 4 * ((size_t)py * p->width + px);
      o[0] = dst[0]; o[1] = dst[1]; o[2] = dst[2]; o[3] = dst[3];
    }
  }
  if (samples_out) *samples_out = total;
  return 0;
}
